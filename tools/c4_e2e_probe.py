#!/usr/bin/env python
"""8 scenes on one GPU through submit_frame / wait_frame (the e2e leg of bench.py's c4_64 record), frames/s."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from infinitam_b200 import capi, synth
from infinitam_b200.engines import ITMMainEngine
S, W, H, n, warm = 8, 640, 480, 24, 4
block = synth.sequence(n + 8 * S, W, H)
pinned = torch.from_numpy(np.ascontiguousarray(block)).pin_memory()
p = capi.default_params(W, H)
p.icp_max_ctas = 148 // S
engs = [ITMMainEngine(p) for _ in range(S)]
addr = [[pinned[i * 7 + k].data_ptr() for k in range(n)] for i in range(S)]
for k in range(warm):
    for i, e in enumerate(engs):
        e.WaitFrame(e.SubmitFrame(None, addr[i][k]))
torch.cuda.synchronize()
t0 = time.perf_counter()
last = [0] * S
tw = 0.0
for k in range(warm, n):
    for i, e in enumerate(engs):
        t = e.SubmitFrame(None, addr[i][k])
        if last[i]:
            a = time.perf_counter(); e.WaitFrame(last[i]); tw += time.perf_counter() - a
        last[i] = t
for i, e in enumerate(engs):
    e.WaitFrame(last[i])
dt = time.perf_counter() - t0
print("c4 e2e %.0f frames/s (%.2f ms per round of %d), time in WaitFrame %.1f ms of %.1f ms" % (S * (n - warm) / dt, 1e3 * dt / (n - warm), S, 1e3 * tw, 1e3 * dt))
