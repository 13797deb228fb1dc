#!/usr/bin/env python
"""Spatially sharded fusion of ONE scene on the GPUs of a box (BASELINE configs[2]); launch with torchrun:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/sharded_run.py [--check] [--frames K] [--size 1280x720] [--voxel 0.002] [--pool 0x80000]

--check: every rank also runs a private single-GPU engine on the same frames and compares after every frame
  pass A (poses supplied, TRACKER_EXTERNAL on both sides, so that nothing but the sharding differs):
    * hash-table positions and chain links, excess free list, visible list: bit-identical;
    * ptr >= 0 exactly where the block is resident on this rank (owner + one-block halo), -1 elsewhere;
    * every resident block's voxels: bit-identical to the single GPU's block;
    * composed raycast image vs the single GPU's: hit-mask mismatches and the largest point difference in metres;
  pass B (free-running ICP on both sides): per-frame pose difference.
Without --check: timing (device time per frame incl. the NCCL depth broadcast, max over ranks), L2 flushed between frames."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from infinitam_b200 import capi, multi, synth  # noqa: E402
from infinitam_b200.engines import ITMMainEngine  # noqa: E402
from infinitam_b200.multi import ShardedEngine, compare_scene  # noqa: E402


def gt_pose(k):
    return np.ascontiguousarray(synth.ground_truth_pose(k).astype(np.float32).T).reshape(16)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", default="1280x720")
    ap.add_argument("--voxel", type=float, default=0.002)
    ap.add_argument("--pool", default="0x80000", help="SDF_LOCAL_BLOCK_NUM per rank")
    ap.add_argument("--single-pool", default=None, help="pool of the single-GPU engine of --check (default: --pool)")
    ap.add_argument("--out", default=None, help="write the JSON summary (rank 0) to this file")
    ap.add_argument("--layout", default=None, help="axis,origin_block,thickness_blocks of the slabs (default: the room's x extent cut into world slabs)")
    args = ap.parse_args()
    W, H = (int(x) for x in args.size.split("x"))
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    p = capi.default_params(W, H)
    p.voxel_size, p.sdf_local_block_num, p.device = args.voxel, int(args.pool, 0), local_rank
    n = args.frames
    layout = tuple(int(x) for x in args.layout.split(",")) if args.layout else None
    seq = torch.from_numpy(synth.sequence(n, W, H)).cuda() if rank == 0 or args.check else None
    torch.cuda.synchronize()
    # the engine must share a stream with torch so that the NCCL broadcast and the frame are stream-ordered; the legacy
    # default stream has handle 0 (= "make a private one" for the C ABI), so use an explicit stream
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    ok = True
    summary = {"mode": "sharded", "n_gpus": world, "size": args.size, "voxel": args.voxel, "pool_per_rank": int(args.pool, 0)}
    if args.check:
        import copy
        import parity
        p1 = copy.copy(p)
        p1.sdf_local_block_num = int(args.single_pool or args.pool, 0)
        frames_out = []
        for tracker, label in ((capi.TRACKER_EXTERNAL, "A: poses supplied"), (capi.TRACKER_ICP, "B: free-running ICP")):
            ps, pq = copy.copy(p), copy.copy(p1)
            ps.tracker_type = pq.tracker_type = tracker
            eng = ShardedEngine(ps, stream=tstream.cuda_stream, layout=layout)
            single = ITMMainEngine(pq)
            for k in range(n):
                if tracker == capi.TRACKER_EXTERNAL:
                    eng.engine.set_state(pose_d=gt_pose(k))
                    single.set_state(pose_d=gt_pose(k))
                eng.EnqueueFrame(seq[k] if rank == 0 else None)
                pose_s, cnt_s = eng.Sync()
                single.EnqueueFrameDevice(seq[k].data_ptr())
                pose_1, cnt_1 = single.Sync()
                rot, trans = parity.pose_diff(pose_s, pose_1)
                rec = {"pass": label, "frame": k, "rank": rank, "pose_rot_rad": rot, "pose_trans_m": trans,
                       "alloc_failures": [int(cnt_s[3]), int(cnt_1[3])]}
                if tracker == capi.TRACKER_EXTERNAL:
                    rec.update(compare_scene(eng.engine, single, rank, world, eng.layout, args.voxel, eng.halo))
                    good = (rec["hash_pos_offset_equal"] and rec["visible_list_equal"] and rec["excess_counter_equal"] and rec["residency_matches_ptr"]
                            and rec["resident_voxel_blocks_equal"]
                            # every pixel some rank could march completely on its own voxels is bit-identical to the single GPU's;
                            # the others (counted by the engine) are marched with peer reads - then every pixel is - or, without
                            # attached peers, reported as misses
                            and rec["raycast_max_diff_m"] == 0.0
                            and (rec["raycast_px_differing"] == 0 if eng.peers else rec["raycast_px_differing"] <= rec["raycast_unresolved_px"]))
                else:
                    good = rot <= 1e-4 and trans <= 1e-4
                rec["ok"] = bool(good)
                ok = ok and good
                frames_out.append(rec)
                print(json.dumps(rec), flush=True)
            single.close()
            eng.close()
        summary["check"] = frames_out if rank == 0 else None
    else:
        eng = ShardedEngine(p, stream=tstream.cuda_stream, layout=layout)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        eng.engine.set_profiling(True)
        tot, stages, cnt, sh3 = 0.0, np.zeros(8), None, np.zeros(3)
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
                if world > 1:
                    sh3 += eng.engine.shard_times()
                tot += e0.elapsed_time(e1)
        t = torch.tensor([tot], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        m = n - args.warmup
        per_rank = torch.tensor(list(stages) + list(sh3) + [float(p.sdf_local_block_num - 1 - cnt[1])], dtype=torch.float64, device="cuda")
        gathered = [torch.empty_like(per_rank) for _ in range(world)]
        dist.all_gather(gathered, per_rank)
        names = ["view", "track", "allocate", "integrate", "expected_depths", "raycast+barrier+compose", "icp_maps", "total"]
        summary.update({"frames": m, "frames_per_s": m / (float(t[0]) * 1e-3), "ms_per_frame": float(t[0]) / m, "visible_blocks": int(cnt[0]),
                        "free_blocks_used_rank0": int(p.sdf_local_block_num - 1 - cnt[1]), "alloc_failures_rank0": int(cnt[3]),
                        "stage_us_per_rank": [{a: round(1e3 * float(v) / m, 1) for a, v in zip(names + ["partial_raycast", "barrier_wait", "compose"], g[:11])}
                                              for g in gathered],
                        "voxel_blocks_in_use_per_rank": [int(g[11]) for g in gathered],
                        "note": "total = CUDA events around NCCL depth broadcast + frame, max over ranks"})
        eng.close()
    ok_t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
    summary["ok"] = bool(int(ok_t[0]) == 1)
    if rank == 0:
        print(json.dumps({k: v for k, v in summary.items() if k != "check"}), flush=True)
        if args.out:
            with open(args.out, "w") as f:
                json.dump(summary, f)
    dist.destroy_process_group()
    sys.exit(0 if summary["ok"] else 1)


if __name__ == "__main__":
    main()
