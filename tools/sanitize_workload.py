#!/usr/bin/env python
"""Small workload for compute-sanitizer: every kernel of the frame (incl. the cooperative tracker), the streaming API, the
external-pose path, colour voxels with swapping, and 4 concurrent scenes, at 320x240."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from infinitam_b200 import capi, synth
from infinitam_b200.engines import ITMMainEngine

W, H = 320, 240
seq = synth.sequence(3, W, H)
eng = ITMMainEngine(width=W, height=H)
for k in range(3):
    eng.ProcessFrame(None, seq[k])
t = [eng.SubmitFrame(None, seq[k]) for k in range(3)]
for x in t:
    eng.WaitFrame(x)
eng.ProcessFrameWithPose(None, seq[2], eng.get_state()[0])
eng.close()
p = capi.default_params(W, H)
p.voxel_type, p.use_swapping = capi.VOXEL_S_RGB, 1
rgb = np.full((H, W, 4), 128, np.uint8)
eng = ITMMainEngine(p)
for k in range(3):
    eng.ProcessFrame(rgb, seq[k])
eng.close()
p = capi.default_params(W, H)
p.icp_max_ctas = 16
import torch
dev_seq = torch.from_numpy(seq).cuda()
engs = [ITMMainEngine(p) for _ in range(4)]
for k in range(3):
    for e in engs:
        e.EnqueueFrameDevice(dev_seq[k].data_ptr())
for e in engs:
    e.Sync()
    e.close()
print("sanitize workload done")
