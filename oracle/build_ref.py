#!/usr/bin/env python
"""Build the REAL reference CPU engines into oracle/_ref/ (test infrastructure).

The reference's own CMake build cannot configure here (it hard-requires
catkin_simple, GLUT, glog ... -- see SURVEY.md section 8c), but the hot path
compiles from a handful of its .cpp files directly.  This recipe compiles those
files WHERE THEY LIE under /root/reference/InfiniTAM (nothing is copied into
the repo) together with oracle/ref_harness.cpp and writes only into
oracle/_ref/:

  libitm_ref.so       parity flavour: serial, -O2 -ffp-contract=off, SSE2 only
                      -> every float op is IEEE fp32 and reproducible on the GPU
  libitm_ref_fast.so  timing flavour: -O3 -mavx2 -mfma -fopenmp -DWITH_OPENMP
                      (the reference's CMake uses -O3 -march=native + optional
                      OpenMP; -march=native is replaced by AVX2+FMA because the
                      binary is built here and executed on the GPU box's host)
  libitm_ref_fast1.so timing flavour, serial (-O3 -mavx2 -mfma, no OpenMP)

  libitm_adapter.so   drop-in proof: the reference's own host objects (scene, render state,
                      tracking state, view, tracking controller, ITMDepthTracker's host loop,
                      ITMPose) compiled WITH CUDA memory placement, driven through
                      include/itm_b200_adapter.hpp -> libitm_b200.so (oracle/adapter_harness.cpp)

oracle/_ref/ is git-ignored but travels to the GPU box with gpurun.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ITM_REFERENCE_ROOT", "/root/reference/InfiniTAM")
OUT = os.path.join(HERE, "_ref")

SOURCES = [
    "ITMLib/Objects/ITMPose.cpp",
    "ITMLib/Utils/ITMLibSettings.cpp",
    "ITMLib/Engine/ITMDepthTracker.cpp",
    "ITMLib/Engine/ITMWeightedICPTracker.cpp",
    "ITMLib/Engine/ITMTrackingController.cpp",
    "ITMLib/Engine/ITMVisualisationEngine.cpp",
    "ITMLib/Engine/DeviceSpecific/CPU/ITMDepthTracker_CPU.cpp",
    "ITMLib/Engine/DeviceSpecific/CPU/ITMWeightedICPTracker_CPU.cpp",
    "ITMLib/Engine/DeviceSpecific/CPU/ITMLowLevelEngine_CPU.cpp",
    "ITMLib/Engine/DeviceSpecific/CPU/ITMViewBuilder_CPU.cpp",
    "ITMLib/Engine/DeviceSpecific/CPU/ITMSceneReconstructionEngine_CPU.cpp",
    "ITMLib/Engine/DeviceSpecific/CPU/ITMVisualisationEngine_CPU.cpp",
    "ITMLib/Engine/DeviceSpecific/CPU/ITMSwappingEngine_CPU.cpp",
    "ITMLib/Engine/DeviceSpecific/CPU/ITMMeshingEngine_CPU.cpp",
]

# sources that ref_harness.cpp compiles itself (by #include) when it instantiates them for another voxel type
VOXEL_DEPENDENT = ["ITMLib/Engine/DeviceSpecific/CPU/ITMSceneReconstructionEngine_CPU.cpp",
                   "ITMLib/Engine/DeviceSpecific/CPU/ITMVisualisationEngine_CPU.cpp",
                   "ITMLib/Engine/DeviceSpecific/CPU/ITMSwappingEngine_CPU.cpp",
                   "ITMLib/Engine/DeviceSpecific/CPU/ITMMeshingEngine_CPU.cpp"]

FLAVOURS = {
    "libitm_ref.so": ["-O2", "-ffp-contract=off"],
    "libitm_ref_rgb.so": ["-O2", "-ffp-contract=off", "-DREF_VOXEL_RGB"],  # ITMVoxel_s_rgb, parity flags
    "libitm_ref_fast.so": ["-O3", "-mavx2", "-mfma", "-fopenmp", "-DWITH_OPENMP"],
    "libitm_ref_fast1.so": ["-O3", "-mavx2", "-mfma"],
}

COMMON = ["-std=c++11", "-fPIC", "-w", "-DCOMPILE_WITHOUT_CUDA", "-include", "iostream",
          "-I" + os.path.join(HERE, "shim"), "-I" + REF]


def available():
    return os.path.isdir(os.path.join(REF, "ITMLib"))


def build(flavours=None, force=False):
    if not available():
        return False
    os.makedirs(OUT, exist_ok=True)
    harness = os.path.join(HERE, "ref_harness.cpp")
    for lib, flags in FLAVOURS.items():
        if flavours and lib not in flavours:
            continue
        target = os.path.join(OUT, lib)
        if (not force and os.path.exists(target)
                and os.path.getmtime(target) > os.path.getmtime(harness)
                and os.path.getmtime(target) > os.path.getmtime(__file__)):
            continue
        objdir = os.path.join(OUT, "obj_" + lib.replace(".so", ""))
        os.makedirs(objdir, exist_ok=True)
        jobs = []
        for src in SOURCES + [harness]:
            if "-DREF_VOXEL_RGB" in flags and src in VOXEL_DEPENDENT:
                continue
            path = src if os.path.isabs(src) else os.path.join(REF, src)
            obj = os.path.join(objdir, os.path.basename(src).replace(".cpp", ".o"))
            jobs.append((["g++"] + COMMON + flags + ["-c", path, "-o", obj], obj))
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            list(ex.map(lambda j: subprocess.check_call(j[0]), jobs))
        link = ["g++", "-shared", "-o", target] + [j[1] for j in jobs]
        if "-fopenmp" in flags:
            link.append("-fopenmp")
        subprocess.check_call(link)
    return True


ADAPTER_SOURCES = [
    "ITMLib/Objects/ITMPose.cpp",
    "ITMLib/Utils/ITMLibSettings.cpp",
    "ITMLib/Engine/ITMDepthTracker.cpp",
    "ITMLib/Engine/ITMWeightedICPTracker.cpp",
    "ITMLib/Engine/ITMTrackingController.cpp",
]


def build_adapter(force=False):
    """reference host objects + include/itm_b200_adapter.hpp, linked against infinitam_b200/libitm_b200.so"""
    repo = os.path.dirname(HERE)
    product = os.path.join(repo, "infinitam_b200", "libitm_b200.so")
    if not available() or not os.path.exists(product):
        return False
    os.makedirs(OUT, exist_ok=True)
    harness = os.path.join(HERE, "adapter_harness.cpp")
    header = os.path.join(repo, "include", "itm_b200_adapter.hpp")
    target = os.path.join(OUT, "libitm_adapter.so")
    deps = [harness, header, os.path.join(repo, "include", "itm_b200.h"), __file__]
    if not force and os.path.exists(target) and all(os.path.getmtime(target) > os.path.getmtime(d) for d in deps):
        return True
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    objdir = os.path.join(OUT, "obj_adapter")
    os.makedirs(objdir, exist_ok=True)
    common = ["-std=c++11", "-fPIC", "-w", "-O2", "-ffp-contract=off", "-include", "iostream", "-I" + os.path.join(HERE, "shim"),
              "-I" + REF, "-I" + os.path.join(repo, "include"), "-I" + os.path.join(cuda, "include")]
    jobs = []
    for src in ADAPTER_SOURCES + [harness]:
        path = src if os.path.isabs(src) else os.path.join(REF, src)
        obj = os.path.join(objdir, os.path.basename(src).replace(".cpp", ".o"))
        jobs.append((["g++"] + common + ["-c", path, "-o", obj], obj))
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(lambda j: subprocess.check_call(j[0]), jobs))
    subprocess.check_call(["g++", "-shared", "-o", target] + [j[1] for j in jobs] + [
        "-L" + os.path.dirname(product), "-litm_b200", "-L" + os.path.join(cuda, "lib64"), "-lcudart",
        "-Wl,-rpath,$ORIGIN/../../infinitam_b200", "-Wl,-rpath," + os.path.join(cuda, "lib64")])
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    build_adapter(force="--force" in sys.argv)
    print("built" if ok else "reference not present; nothing built")
