#!/usr/bin/env python
"""Small workload for compute-sanitizer: every kernel of the frame (incl. the cooperative tracker), the streaming API, the
external-pose path, colour voxels with swapping, free-view rendering, the point cloud, meshing, the low-level image helpers,
and 4 concurrent scenes, at 320x240."""
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
M = np.eye(4, dtype=np.float32)
M[:3, 3] = [0.05, -0.02, 0.03]
for t in (capi.IMAGE_FREECAMERA_SHADED, capi.IMAGE_FREECAMERA_COLOUR_FROM_VOLUME, capi.IMAGE_FREECAMERA_COLOUR_FROM_NORMAL):
    eng.GetImage(t, M.T.reshape(16), (290.0, 290.0, 160.0, 120.0), 200, 150)
for skip in (False, True):
    loc, clr = eng.CreatePointCloud(M.T.reshape(16), None, skip)
    assert len(loc) > 1000
assert len(eng.UpdateMesh()) > 1000
eng.close()
p = capi.default_params(W, H)
p.icp_max_ctas = 16
import ctypes as C
import torch
dev_seq = torch.from_numpy(seq).cuda()
# Layer A: the low-level image helpers on caller-owned buffers (odd output sizes)
lib = capi.load()
ctx = C.c_void_p()
capi.check(lib.itm_b200_ctx_create(C.byref(p), None, C.byref(ctx)))
w2, h2 = 322, 242
img = torch.randint(0, 256, (h2 * w2 * 4,), dtype=torch.uint8, device="cuda")
f4 = torch.randn(h2 * w2 * 4, dtype=torch.float32, device="cuda")
half_u8 = torch.zeros((h2 // 2) * (w2 // 2) * 4, dtype=torch.uint8, device="cuda")
half_f4 = torch.zeros((h2 // 2) * (w2 // 2) * 4, dtype=torch.float32, device="cuda")
grad = torch.zeros(h2 * w2 * 4, dtype=torch.int16, device="cuda")
copy = torch.zeros_like(img)
torch.cuda.synchronize()  # the context runs on a stream of its own: torch's fills must have landed
capi.check(lib.itm_b200_copy_image(ctx, copy.data_ptr(), img.data_ptr(), img.numel()))
capi.check(lib.itm_b200_filter_subsample_rgba(ctx, half_u8.data_ptr(), img.data_ptr(), w2, h2))
capi.check(lib.itm_b200_filter_subsample_with_holes_float4(ctx, half_f4.data_ptr(), f4.data_ptr(), w2, h2))
capi.check(lib.itm_b200_gradient_x(ctx, grad.data_ptr(), img.data_ptr(), w2, h2))
capi.check(lib.itm_b200_gradient_y(ctx, grad.data_ptr(), img.data_ptr(), w2, h2))
assert bool((copy == img).all())
lib.itm_b200_ctx_destroy(ctx)
engs = [ITMMainEngine(p) for _ in range(4)]
for k in range(3):
    for e in engs:
        e.EnqueueFrameDevice(dev_seq[k].data_ptr())
for e in engs:
    e.Sync()
    e.close()
print("sanitize workload done")
