#!/usr/bin/env python
"""Per-iteration timeline of the persistent ICP kernel (needs a build with ITM_B200_DEFINES=-DITM_ICP_TRACE)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from infinitam_b200 import capi, synth
from infinitam_b200.engines import ITMMainEngine

lib = capi.load()
eng = ITMMainEngine(width=640, height=480)
eng.set_profiling(True)
seq = synth.sequence(12, 640, 480)
for k in range(12):
    eng.ProcessFrame(None, seq[k])
    if k >= 10:
        buf = (C.c_ulonglong * 2048)()
        lib.itm_b200_debug_icp_trace(buf)
        t = np.array(buf[:], dtype=np.uint64).reshape(64, 32).astype(np.int64)
        _, cnt = eng.Sync()
        n = int(cnt[5])
        print("frame", k, "evals", n)
        base = t[0, 0]
        print("  kernel begin %+.1f us before eval 0, kernel end %+.1f us after it; stage_ms(track) = %.1f us" % (
            (base - t[63, 0]) / 1e3, (t[63, 1] - base) / 1e3, 1e3 * eng.stage_times()[1]))
        cb = (C.c_ulonglong * (64 * 160 * 2))()
        lib.itm_b200_debug_icp_cta_trace(cb)
        ct = np.array(cb[:], dtype=np.uint64).reshape(64, 160, 2).astype(np.int64)
        for i in range(n):
            r = t[i]
            act = ct[i, :, 1] > r[0] - 1000  # CTAs that stored a row in this evaluation
            st, rw = (ct[i, act, 0] - r[0]) / 1e3, (ct[i, act, 1] - r[0]) / 1e3
            if act.sum():
                print("      %3d CTAs: pose received min %+5.2f med %+5.2f max %+5.2f | row stored min %+5.2f med %+5.2f max %+5.2f | own time med %5.2f max %5.2f" % (
                    act.sum(), st.min(), np.median(st), st.max(), rw.min(), np.median(rw), rw.max(), np.median(rw - st), (rw - st).max()))
            print("  eval %d level %d: start %+6.1f us | cta0 row written +%5.1f | all rows gathered +%5.1f | lm %4.1f | next pose at cta0 +%5.1f" % (
                i, r[6], (r[0] - base) / 1e3, (r[1] - r[0]) / 1e3, (r[3] - r[0]) / 1e3, (r[4] - r[3]) / 1e3, (r[5] - r[0]) / 1e3))
            print("      traced pixel: loop entered +%5.2f | depth here +%5.2f | projected +%5.2f | point taps here +%5.2f | accumulated +%5.2f" % tuple((r[j] - r[0]) / 1e3 for j in (20, 16, 17, 18, 19)))
            print("      thread 0: pixels done +%5.2f | warp reduce +%5.2f | barrier +%5.2f | row stored +%5.2f | lm stages %s" % (
                (r[8] - r[0]) / 1e3, (r[9] - r[0]) / 1e3, (r[10] - r[0]) / 1e3, (r[11] - r[0]) / 1e3,
                " ".join("%.2f" % ((r[j] - r[3]) / 1e3) for j in (12, 21, 13, 14, 15) if r[j] > r[0])))
