#!/usr/bin/env python
"""BASELINE configs[3] on one GPU: S independent scenes (one engine + stream each) fused concurrently.

  python tools/batched_scenes.py [S=8] [frames=40] [icp_max_ctas=148//S] [threads=0|1]

threads=0: one host thread enqueues all scenes round-robin (one graph launch per frame); threads=1: one host thread per
scene (ctypes releases the GIL inside the C ABI).  Prints aggregate frames/s (host clock, first enqueue to last sync)."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from infinitam_b200 import capi, synth
from infinitam_b200.engines import ITMMainEngine

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 148 // S
threads = int(sys.argv[4]) if len(sys.argv) > 4 else 0
W, H = 640, 480
seq = torch.from_numpy(synth.sequence(n, W, H)).cuda()
p = capi.default_params(W, H)
p.icp_max_ctas = cap
engs = [ITMMainEngine(p) for _ in range(S)]
warm = 4


def run(e, lo, hi):
    for k in range(lo, hi):
        e.EnqueueFrameDevice(seq[k].data_ptr())


for e in engs:
    run(e, 0, warm)
for e in engs:
    e.Sync()
t0 = time.perf_counter()
if threads:
    ts = [threading.Thread(target=run, args=(e, warm, n)) for e in engs]
    [t.start() for t in ts]
    [t.join() for t in ts]
else:
    for k in range(warm, n):
        for e in engs:
            e.EnqueueFrameDevice(seq[k].data_ptr())
t_enq = time.perf_counter() - t0
poses = [e.Sync()[0] for e in engs]
dt = time.perf_counter() - t0
same = all(np.array_equal(poses[0], q) for q in poses)
print("scenes=%d icp_max_ctas=%d threads=%d: %.0f aggregate frames/s (enqueue %.1f ms of %.1f ms; all scenes reach the same pose: %s)" % (
    S, cap, threads, S * (n - warm) / dt, 1e3 * t_enq, 1e3 * dt, same), flush=True)
