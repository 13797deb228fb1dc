#!/usr/bin/env python
"""A/B of the two integration kernels on identical input (debug aid):

  ITM_B200_INTEGRATE=rows python tools/integrate_ab.py save && ITM_B200_INTEGRATE=cols python tools/integrate_ab.py check

Both runs fuse frame 0 normally, then run view + allocate + integrate of frame 3 at its ground-truth pose as single
stages; `save` stores the voxel array under /tmp, `check` compares its own against it and prints where they differ."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from infinitam_b200 import capi, synth
from infinitam_b200.engines import ITMMainEngine

mode = sys.argv[1]
W, H = 320, 240
p = capi.default_params(W, H)
eng = ITMMainEngine(p)
seq = synth.sequence(4, W, H)
eng.ProcessFrame(None, seq[0])
v0 = eng.read(capi.BUF_VOXELS).copy()
M = np.ascontiguousarray(synth.ground_truth_pose(3).astype(np.float32).T).reshape(16)
eng.set_state(pose_d=M)
eng.UploadDepth(seq[3])
eng.RunStage(capi.STAGE_VIEW)
eng.RunStage(capi.STAGE_ALLOCATE)
eng.RunStage(capi.STAGE_INTEGRATE)
v1 = eng.read(capi.BUF_VOXELS)
h = eng.read(capi.BUF_HASH)
if mode == "save":
    np.save("/tmp/vox_a0.npy", v0)
    np.save("/tmp/vox_a1.npy", v1)
    print("saved", v1.shape, int((v1 != v0).sum()), "voxels changed by the second integration")
else:
    a0, a1 = np.load("/tmp/vox_a0.npy"), np.load("/tmp/vox_a1.npy")
    print("frame-0 voxels equal:", np.array_equal(a0, v0))
    d = np.nonzero(a1 != v1)[0]
    print("differing voxels after the second integration:", len(d), "of", int((a1 != a0).sum()), "changed")
    ptr_to_entry = {int(pp): i for i, pp in enumerate(h["ptr"]) if pp >= 0}
    for i in d[:24]:
        blk, lin = divmod(int(i), 512)
        x, y, z = lin & 7, (lin >> 3) & 7, lin >> 6
        e = ptr_to_entry.get(blk, -1)
        pos = h["pos"][e] if e >= 0 else None
        def dec(v):
            return "sdf=%6d w=%3d" % (np.int16(np.uint16(v & 0xFFFF)), (v >> 16) & 0xFF)
        print("block %d pos %s voxel (%d,%d,%d): before %s | rows %s | cols %s" % (blk, pos, x, y, z, dec(int(v0[i])), dec(int(a1[i])), dec(int(v1[i]))))
    if len(d):
        lin = d % 512
        print("x histogram", np.bincount(lin & 7, minlength=8))
        print("y histogram", np.bincount((lin >> 3) & 7, minlength=8))
        print("z histogram", np.bincount(lin >> 6, minlength=8))
