#!/usr/bin/env python
"""Short workload for ncu: N frames of the VGA sequence through ITMMainEngine.ProcessFrame; with "extras" also the
SURVEY 8f rows (approximate raycast frames, a free-view GetImage, UpdateMesh)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from infinitam_b200 import capi, synth
from infinitam_b200.engines import ITMMainEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (640, 480)
extras = "extras" in sys.argv
p = capi.default_params(W, H)
for a in sys.argv:  # voxel=0.002 pool=0x80000 (BASELINE configs[2] shape)
    if a.startswith("voxel="):
        p.voxel_size = float(a[6:])
    if a.startswith("pool="):
        p.sdf_local_block_num = int(a[5:], 0)
if extras:
    p.use_approximate_raycast = 1
eng = ITMMainEngine(p)
seq = synth.sequence(n, W, H)
for k in range(n):
    eng.ProcessFrame(None, seq[k])
print("done", eng.Sync()[1])
if extras:
    M = np.eye(4, dtype=np.float32)
    M[:3, 3] = [0.1, -0.04, 0.06]
    img = eng.GetImage(capi.IMAGE_FREECAMERA_SHADED, M.T.reshape(16), synth.intrinsics_for(W, H))
    tri = eng.UpdateMesh()
    print("extras", img.mean(), len(tri))
