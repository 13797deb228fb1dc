"""Builds infinitam_b200/libitm_b200.so (CUDA kernels + C ABI) for sm_100a with nvcc.

The library is built IN-TREE so that it travels to the GPU box with the repo snapshot.
-fmad=false: every float expression must round like the reference's scalar CPU build
(SURVEY.md appendix A); IEEE division and square root are nvcc's defaults.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# ITM_B200_VARIANT=name builds an experimental variant next to the product library (libitm_b200_<name>.so, objects in
# build_<name>/) with the extra defines of ITM_B200_DEFINES; tools select it with ITM_B200_LIB (capi.load)
VARIANT = os.environ.get("ITM_B200_VARIANT", "")
LIB = os.path.join(HERE, "libitm_b200%s.so" % ("_" + VARIANT if VARIANT else ""))
SOURCES = ["engine.cu", "k_view.cu", "k_alloc.cu", "k_integrate.cu", "k_render.cu", "k_vis.cu", "k_mesh.cu", "k_icp.cu", "k_swap.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
EXTRA = os.environ.get("ITM_B200_DEFINES", "").split()  # e.g. -DITM_ICP_TRACE for tools/icp_trace.py
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-Xptxas", "-v",
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "itm_b200.h"))
    headers.append(os.path.abspath(__file__))
    objdir = os.path.join(HERE, "build" + ("_" + VARIANT if VARIANT else ""))
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        if force or _stale(obj, [path] + headers):
            jobs.append(([NVCC] + FLAGS + EXTRA + ["-c", path, "-o", obj], src))

    def run(job):
        cmd, src = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        with open(os.path.join(objdir, src + ".ptxas.log"), "w") as f:
            f.write(r.stderr)

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        subprocess.check_call([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
