#!/usr/bin/env python
"""Short workload for ncu: N frames of the VGA sequence through ITMMainEngine.ProcessFrame."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from infinitam_b200 import synth
from infinitam_b200.engines import ITMMainEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (640, 480)
eng = ITMMainEngine(width=W, height=H)
seq = synth.sequence(n, W, H)
for k in range(n):
    eng.ProcessFrame(None, seq[k])
print("done", eng.Sync()[1])
