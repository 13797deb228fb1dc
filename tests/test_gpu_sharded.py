"""-m gpu (needs >= 2 GPUs, skipped otherwise): one scene spread over two ranks - replicated index, voxel payload partitioned
into slabs with a one-block halo, per-rank partial ray casts composed by nearest hit over NVLink.  tools/sharded_run.py
--check compares every rank against a private single-GPU engine after every frame (hash positions / links and visible list
bit-identical, ptr >= 0 exactly on resident blocks, resident voxel blocks bit-identical, composed raycast within 1e-4 m with an
equal hit mask up to 0.1 % of the pixels, free-running pose within 1e-4) and exits non-zero otherwise."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


def _json_objects(text):
    """every JSON object in `text` (the ranks' lines may interleave)"""
    dec, out, i = json.JSONDecoder(), [], 0
    while True:
        i = text.find("{", i)
        if i < 0:
            return out
        try:
            obj, end = dec.raw_decode(text, i)
            out.append(obj)
            i = end
        except json.JSONDecodeError:
            i += 1


def _run(extra, nproc=2):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "sharded_run.py")] + extra
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900)


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs with peer access")
def test_two_rank_sharded_scene_matches_single_gpu():
    r = _run(["--check", "--frames", "8", "--size", "320x240", "--voxel", "0.005", "--pool", "0x10000"])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-3000:]
    recs = [x for x in _json_objects(r.stdout) if "frame" in x]
    assert len(recs) == 2 * 2 * 8 and all(x["ok"] for x in recs)
    a = [x for x in recs if x["pass"].startswith("A")]
    # the payload really is partitioned: no rank holds every block, together they hold all of them
    last = [x for x in a if x["frame"] == 7]
    assert all(x["resident_blocks"] < x["allocated_blocks"] for x in last)
    assert sum(x["owned_blocks"] for x in last) == last[0]["allocated_blocks"]


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs with peer access")
def test_slabs_across_the_viewing_direction_still_give_the_single_gpu_raycast():
    """slabs cut along z (the viewing direction), 40 blocks thick: a few percent of the rays pass through allocated blocks of both
    slabs, so that neither rank can march them on its own voxels - they are marched with peer reads over NVLink.  The composed
    raycast image must be the single GPU's in every pixel, and the free-running poses identical."""
    r = _run(["--check", "--frames", "6", "--size", "320x240", "--voxel", "0.005", "--pool", "0x10000", "--layout", "2,0,40"])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-3000:]
    recs = [x for x in _json_objects(r.stdout) if "frame" in x]
    assert len(recs) == 2 * 2 * 6 and all(x["ok"] for x in recs)
    a = [x for x in recs if x["pass"].startswith("A")]
    assert max(x["raycast_unresolved_px"] for x in a) > 500, "the layout was meant to produce rays that cross slabs"
    assert all(x["raycast_px_differing"] == 0 and x["raycast_hit_mismatch"] == 0 for x in a)
    b = [x for x in recs if x["pass"].startswith("B")]
    assert all(x["pose_rot_rad"] == 0.0 and x["pose_trans_m"] == 0.0 for x in b)


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs with peer access")
def test_two_ranks_hold_a_scene_that_overflows_one_pool():
    """per-rank pool of 0x1400 blocks: one GPU runs out (allocation failures), two GPUs hold the scene without any"""
    r = _run(["--frames", "4", "--warmup", "1", "--size", "320x240", "--voxel", "0.005", "--pool", "0x1400"])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    two = [x for x in _json_objects(r.stdout) if "alloc_failures_rank0" in x][-1]
    one = _run(["--frames", "4", "--warmup", "1", "--size", "320x240", "--voxel", "0.005", "--pool", "0x1400"], nproc=1)
    assert one.returncode == 0, one.stdout[-3000:] + one.stderr[-3000:]
    one = [x for x in _json_objects(one.stdout) if "alloc_failures_rank0" in x][-1]
    assert one["alloc_failures_rank0"] > 0 and two["alloc_failures_rank0"] == 0
