"""The marching-cubes case table of infinitam_b200/csrc/k_mesh.cu (one 64-bit word per case, a nibble per edge index).

Self-contained checks from the cube's topology, plus - when the reference tree is present - equality with the reference's
triangleTable / edgeTable (ITMLib/Engine/DeviceAgnostic/ITMMeshingEngine.h:9-151)."""
import os
import re

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "..", "infinitam_b200", "csrc", "k_mesh.cu")
REF = "/root/reference/InfiniTAM/ITMLib/Engine/DeviceAgnostic/ITMMeshingEngine.h"

EDGE_A = [0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3]
EDGE_B = [1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7]


def _cases():
    src = open(SRC).read()
    body = re.search(r"MC_CASE\[256\]\s*=\s*\{(.*?)\};", src, re.S).group(1)
    words = [int(w, 16) for w in re.findall(r"0x([0-9A-Fa-f]{16})ULL", body)]
    assert len(words) == 256
    out = []
    for w in words:
        edges = []
        for i in range(16):
            n = (w >> (4 * i)) & 0xF
            if n == 0xF:
                break
            edges.append(n)
        assert all(((w >> (4 * i)) & 0xF) == 0xF for i in range(len(edges), 16)), "nibbles after the terminator must be 0xF"
        out.append(edges)
    return out


def test_case_table_is_consistent_with_the_cube():
    cases = _cases()
    for c, edges in enumerate(cases):
        assert len(edges) % 3 == 0 and len(edges) <= 15 and all(e < 12 for e in edges)
        # the edges a case's triangles touch are exactly the edges whose two corners lie on different sides
        crossing = {e for e in range(12) if ((c >> EDGE_A[e]) & 1) != ((c >> EDGE_B[e]) & 1)}
        assert set(edges) == crossing, "case %d" % c
    assert cases[0] == [] and cases[255] == []
    # complementary cases cut the same edges
    for c in range(256):
        assert set(cases[c]) == set(cases[255 - c])


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")
def test_case_table_equals_the_reference():
    src = open(REF).read()
    rows = re.findall(r"\{([^{}]*)\}", re.search(r"triangleTable\[256\]\[16\]\s*=\s*\{(.*?)\};", src, re.S).group(1))
    ref_cases = [[int(x) for x in r.split(",") if int(x) >= 0] for r in rows]
    assert ref_cases == _cases()
    ref_edges = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", re.search(r"edgeTable\[256\]\s*=\s*\{(.*?)\};", src, re.S).group(1))]
    assert ref_edges == [sum(1 << e for e in set(c)) for c in _cases()]
