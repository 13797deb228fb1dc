#!/usr/bin/env python
"""Free-running trajectory report: the CUDA engine, the reference CPU engines and the ground truth over the
whole synthetic sequence, each engine on its own (no teacher forcing).

  python tools/drift_report.py [W H N] [--json out.json]

Per frame: pose difference CUDA vs reference (the parity signal: both run the same algorithm on the same frames, closed loop),
and both against the analytic ground truth (tracking accuracy; identical for the two if they stay in step)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import parity  # noqa: E402
from infinitam_b200 import synth  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
W, H, N = (int(x) for x in (args[:3] if len(args) >= 3 else (640, 480, 100)))
out_path = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
from oracle import ref  # noqa: E402
if ref.available("parity"):
    o, kind = ref.RefEngine(W, H), "reference CPU engines (oracle/_ref/libitm_ref.so: -O2, serial, no FMA contraction)"
else:
    from oracle import port
    o, kind = port.PortEngine(W, H, fast=True), "C restatement (oracle/itm_oracle.c)"
seq = synth.sequence(N, W, H)
eng = parity.make_cuda_engine(o)
rows = []
for k in range(N):
    o.process_frame(seq[k])
    pose = eng.ProcessFrame(None, seq[k])
    gt = synth.ground_truth_pose(k).T.reshape(16)
    r1 = parity.pose_diff(pose, o.pose_M)
    r2 = parity.pose_diff(pose, gt)
    r3 = parity.pose_diff(o.pose_M, gt)
    _, cnt = eng.Sync()
    rows.append({"frame": k, "cuda_vs_ref_rot_rad": r1[0], "cuda_vs_ref_trans_m": r1[1], "cuda_vs_gt_rot_rad": r2[0], "cuda_vs_gt_trans_m": r2[1],
                 "ref_vs_gt_rot_rad": r3[0], "ref_vs_gt_trans_m": r3[1], "icp_evaluations": int(cnt[5]), "visible_blocks_cuda": int(cnt[0]),
                 "visible_blocks_ref": int(o.counters[0]), "free_list_heads_equal": [int(cnt[1]), int(cnt[2])] == [int(o.counters[1]), int(o.counters[2])]})
    print("%3d cuda-vs-ref %.2e %.2e | cuda-vs-gt %.2e %.2e | ref-vs-gt %.2e %.2e | evals %d nvis %d / %d" % (
        k, *r1, *r2, *r3, cnt[5], cnt[0], o.counters[0]), flush=True)
summary = {"size": "%dx%d" % (W, H), "frames": N, "oracle": kind,
           "max_cuda_vs_ref_rot_rad": max(r["cuda_vs_ref_rot_rad"] for r in rows), "max_cuda_vs_ref_trans_m": max(r["cuda_vs_ref_trans_m"] for r in rows),
           "final_cuda_vs_gt": [rows[-1]["cuda_vs_gt_rot_rad"], rows[-1]["cuda_vs_gt_trans_m"]],
           "final_ref_vs_gt": [rows[-1]["ref_vs_gt_rot_rad"], rows[-1]["ref_vs_gt_trans_m"]],
           "frames_with_identical_visible_count": sum(r["visible_blocks_cuda"] == r["visible_blocks_ref"] for r in rows),
           "frames_with_identical_free_list_heads": sum(r["free_list_heads_equal"] for r in rows), "per_frame": rows}
print(json.dumps({k: v for k, v in summary.items() if k != "per_frame"}))
if out_path:
    with open(out_path, "w") as f:
        json.dump(summary, f, indent=1)
