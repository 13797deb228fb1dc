"""-m gpu: BASELINE configs[2] geometry on one GPU - 1280x720 depth, 2 mm voxels (the truncation band spans 2.5 voxel
blocks, 5-6 allocation steps per pixel), enlarged local block pool - every stage teacher-forced against the oracle."""
import numpy as np
import pytest

import parity
from infinitam_b200 import synth
from oracle import port

pytestmark = pytest.mark.gpu


def test_large_volume_config_matches_oracle():
    W, H = 1280, 720
    o = port.PortEngine(W, H, voxel_size=0.002, n_local=0x20000)
    eng = parity.make_cuda_engine(o)
    seq = synth.sequence(2, W, H)
    for k in range(2):
        r = parity.compare_frame(o, eng, seq[k], k, strict=True)
        assert r["hash_equal"] and r["visible_equal"] and r["voxel_max_dsdf"] <= 1
    assert r["counters_ref"][0] > 30000  # ~40 k visible blocks: 10x the 640x480 / 5 mm working set
    eng.close(); o.close()
