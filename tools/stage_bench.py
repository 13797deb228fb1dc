#!/usr/bin/env python
"""Quick per-stage device times (us) over N frames of the VGA sequence, L2 flushed between frames.  For A/B runs of
kernel variants: ITM_B200_DEFINES="-DX=1" python -m infinitam_b200.build --force && python tools/stage_bench.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from infinitam_b200 import synth
from infinitam_b200.engines import ITMMainEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (640, 480)
tag = sys.argv[4] if len(sys.argv) > 4 else ""
voxel = float(sys.argv[5]) if len(sys.argv) > 5 else 0.005
n_local = int(sys.argv[6], 0) if len(sys.argv) > 6 else 0x10000
seq = torch.from_numpy(synth.sequence(n, W, H)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
from infinitam_b200 import capi
p = capi.default_params(W, H)
p.voxel_size, p.sdf_local_block_num = voxel, n_local
eng = ITMMainEngine(p)
eng.set_profiling(True)
acc = np.zeros(8)
cnt = 0
nvis = 0
for k in range(n):
    flush.fill_(k & 0xFF)
    torch.cuda.synchronize()
    eng.EnqueueFrameDevice(seq[k].data_ptr())
    _, counters = eng.Sync()
    if k >= 5:
        acc += eng.stage_times()
        cnt += 1
        nvis += int(counters[0])
names = ["view", "track", "alloc", "integ", "expd", "ray", "maps", "total"]
print(tag, " ".join("%s=%.1f" % (a, 1e3 * v / cnt) for a, v in zip(names, acc)), "nvis=%d" % (nvis // cnt), "counters", list(counters), flush=True)
integ_bytes = (nvis / cnt) * (2 * 2048 + 20) + 4 * W * H
print(tag, "integrate: %.1f MB algorithmic -> %.0f GB/s" % (integ_bytes / 1e6, integ_bytes / (acc[3] / cnt * 1e-3) / 1e9), flush=True)
