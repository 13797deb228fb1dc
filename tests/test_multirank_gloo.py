"""CPU (-m "not gpu"): the N>1 host logic (scene partition + max-over-ranks timing) under gloo, world_size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from infinitam_b200 import multi


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scenes = multi.partition_scenes(5, world, rank)
    local_ms = 10.0 * (rank + 1)  # rank 1 is the slow one
    ms, frames = multi.combine_timing(local_ms, 100 * len(scenes))
    fps = multi.aggregate_frames_per_second(local_ms, 100 * len(scenes))
    q.put((rank, scenes, ms, frames, fps))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_and_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1, 2] and res[1][1] == [3, 4]  # balanced, contiguous, disjoint, complete
    for _, _, ms, frames, fps in res:
        assert ms == 20.0 and frames == 500 and abs(fps - 500 / 0.020) < 1e-6


def test_partition_properties():
    for n in (1, 7, 64):
        for w in (1, 2, 4, 8):
            parts = [multi.partition_scenes(n, w, r) for r in range(w)]
            flat = [s for p in parts for s in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert multi.combine_timing(3.0, 4) == (3.0, 4)  # no process group: identity
