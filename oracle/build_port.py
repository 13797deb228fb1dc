#!/usr/bin/env python
"""Compiles oracle/itm_oracle.c (the plain-C restatement; test infrastructure) into oracle/libitm_oracle.so."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "itm_oracle.c")
LIB = os.path.join(HERE, "libitm_oracle.so")
LIB_FAST = os.path.join(HERE, "libitm_oracle_fast.so")


def build(force=False):
    for lib, flags in ((LIB, ["-O2", "-ffp-contract=off"]), (LIB_FAST, ["-O3", "-mavx2", "-mfma"])):
        if force or not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(SRC), os.path.getmtime(__file__)):
            subprocess.check_call(["gcc", "-std=c99", "-fPIC", "-shared", "-Wall", "-Wno-unused-function"] + flags + ["-o", lib, SRC, "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
