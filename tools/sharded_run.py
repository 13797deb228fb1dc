#!/usr/bin/env python
"""Spatially sharded fusion of ONE scene on the GPUs of a box (BASELINE configs[2]); launch with torchrun:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/sharded_run.py [--check] [--frames K] [--size 1280x720] [--voxel 0.002] [--pool 0x80000]

--check: every rank also runs a private single-GPU engine on the same frames and compares pose, hash table, voxel
blocks, visible list and ICP maps BITWISE after every frame (the sharded run must equal the single-GPU run).
Without --check: timing (device time per frame, max over ranks), L2 flushed between frames."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from infinitam_b200 import capi, synth  # noqa: E402
from infinitam_b200.engines import ITMMainEngine  # noqa: E402
from infinitam_b200.multi import ShardedEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", default="1280x720")
    ap.add_argument("--voxel", type=float, default=0.002)
    ap.add_argument("--pool", default="0x80000", help="SDF_LOCAL_BLOCK_NUM")
    args = ap.parse_args()
    W, H = (int(x) for x in args.size.split("x"))
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    p = capi.default_params(W, H)
    p.voxel_size, p.sdf_local_block_num, p.device = args.voxel, int(args.pool, 0), local_rank
    n = args.frames
    seq = torch.from_numpy(synth.sequence(n, W, H)).cuda() if rank == 0 or args.check else None
    torch.cuda.synchronize()
    # the engine must share a stream with torch so that the NCCL broadcast and the frame are stream-ordered; the legacy
    # default stream has handle 0 (= "make a private one" for the C ABI), so use an explicit stream
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    eng = ShardedEngine(p, stream=tstream.cuda_stream)
    ok = True
    if args.check:
        single = ITMMainEngine(p)
        for k in range(n):
            eng.EnqueueFrame(seq[k] if rank == 0 else None)
            pose_s, cnt_s = eng.Sync()
            single.EnqueueFrameDevice(seq[k].data_ptr())
            pose_1, cnt_1 = single.Sync()
            same = {"pose": np.array_equal(pose_s, pose_1), "counters": np.array_equal(cnt_s[:3], cnt_1[:3])}
            for name, buf in (("hash", capi.BUF_HASH), ("voxels", capi.BUF_VOXELS), ("visible", capi.BUF_VISIBLE_IDS),
                              ("raycast", capi.BUF_RAYCAST_RESULT), ("points", capi.BUF_POINTS), ("normals", capi.BUF_NORMALS)):
                a, b = eng.engine.read(buf), single.read(buf)
                if name == "visible":
                    a, b = a[: cnt_s[0]], b[: cnt_1[0]]
                same[name] = a.tobytes() == b.tobytes()
            ok = ok and all(same.values())
            print("rank %d frame %d nvis %d %s" % (rank, k, cnt_s[0], "BITWISE EQUAL to single GPU" if all(same.values()) else "DIFFERS: %s" % same), flush=True)
        single.close()
    else:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        eng.engine.set_profiling(True)
        tot, stages, cnt = 0.0, np.zeros(8), None
        for k in range(n):
            flush.fill_(k & 0xFF)
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()  # the engine runs on this (torch's current) stream: broadcast + frame are inside the events
            eng.EnqueueFrame(seq[k] if rank == 0 else None)
            e1.record()
            _, cnt = eng.Sync()
            if k >= args.warmup:
                stages += eng.engine.stage_times()
                tot += e0.elapsed_time(e1)
        t = torch.tensor([tot], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            m = n - args.warmup
            names = ["view", "track", "allocate", "integrate+barrier", "expected_depths", "raycast+barrier", "icp_maps", "total"]
            print(json.dumps({"mode": "sharded", "n_gpus": world, "size": args.size, "voxel": args.voxel, "frames": m,
                              "frames_per_s": m / (float(t[0]) * 1e-3), "ms_per_frame": float(t[0]) / m, "visible_blocks": int(cnt[0]),
                              "stage_us_rank0": {a: round(1e3 * v / m, 1) for a, v in zip(names, stages)},
                              "note": "total = CUDA events around NCCL depth broadcast + frame, max over ranks; stage times are rank 0's"}), flush=True)
    eng.close()
    ok_t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(ok_t[0]) == 1 else 1)


if __name__ == "__main__":
    main()
