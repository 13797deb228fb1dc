#!/usr/bin/env python
"""Free-running drift: CUDA engine vs oracle port vs ground truth."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity
from infinitam_b200 import synth
from oracle import port

W, H, N = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (320, 240, 30)))
seq = synth.sequence(N, W, H)
o = port.PortEngine(W, H, fast=True)
eng = parity.make_cuda_engine(o)
for k in range(N):
    o.process_frame(seq[k])
    pose = eng.ProcessFrame(None, seq[k])
    gt = synth.ground_truth_pose(k).T.reshape(16)
    r1 = parity.pose_diff(pose, o.pose_M)
    r2 = parity.pose_diff(pose, gt)
    r3 = parity.pose_diff(o.pose_M, gt)
    _, cnt = eng.Sync()
    print("%2d cuda-vs-port %.2e %.2e | cuda-vs-gt %.2e %.2e | port-vs-gt %.2e %.2e | evals %d nvis %d" % (k, *r1, *r2, *r3, cnt[5], cnt[0]))
