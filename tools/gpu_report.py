#!/usr/bin/env python
"""Non-asserting parity report: runs the teacher-forced comparison and prints every measured difference.
Usage (on a GPU box): python tools/gpu_report.py [--frames N] [--size WxH] [--free N]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import parity  # noqa: E402
from infinitam_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--size", default="640x480")
    ap.add_argument("--free", type=int, default=0, help="also run N frames free-running")
    ap.add_argument("--noise", action="store_true")
    ap.add_argument("--voxel", type=float, default=0.005)
    args = ap.parse_args()
    W, H = (int(x) for x in args.size.split("x"))
    seq = synth.sequence(max(args.frames, args.free), W, H, noise=args.noise)
    oracle = ref.RefEngine(W, H, voxel_size=args.voxel)
    eng = parity.make_cuda_engine(oracle)
    for k in range(args.frames):
        t = time.time()
        try:
            r = parity.compare_frame(oracle, eng, seq[k], k, strict=False)
        except Exception as e:  # noqa: BLE001
            print("frame %d: EXCEPTION %r" % (k, e))
            raise
        r["seconds"] = round(time.time() - t, 2)
        print(json.dumps(r, default=str))
    eng.close()
    oracle.close()
    if args.free:
        oracle = ref.RefEngine(W, H, voxel_size=args.voxel)
        eng = parity.make_cuda_engine(oracle)
        for row in parity.compare_free_running(oracle, eng, seq[: args.free]):
            print(json.dumps(row))


if __name__ == "__main__":
    main()
