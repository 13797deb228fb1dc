#!/usr/bin/env python
"""Benchmark of the per-frame dense-fusion hot path (BASELINE.json metric: fused frames/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one ITMMainEngine::ProcessFrame (allocate + integrate + raycast + ICP) on one frame of the synthetic 640x480
sequence (BASELINE.json configs[1]).  With N > 1 (torchrun, one rank per GPU) every rank fuses its own scene - the SAME
sequence on every rank, so per-rank work is identical and the driver's efficiency figure compares like with like ("batches of
independent sequences, one scene per GPU": no data-path collective) - scaling is weak and `value` is the aggregate frames/s.

Headline keys
  value   frames / sum of per-frame device time (CUDA events recorded by the library on the engine's own stream as nodes of
          the frame graph; the raw frame already in HBM; L2 flushed before every frame, outside the events); max over ranks.
  e2e     the same frames through the host-buffer C ABI: itm_b200_engine_submit_frame / _wait_frame (pinned host rgb + depth,
          H2D inside the timed region, pose + counters read back per frame), three frames in flight; host clock from the first
          submit to the last wait minus the device time of the L2 flushes enqueued between the frames (CUDA events).
          `e2e_blocking` is the round-1 figure: the blocking ProcessFrame call, one frame at a time.
  stage_ms / roofline / roofline_other   per-stage device times from a second pass with a stamp at every stage boundary and
          the algorithmic bytes of DESIGN.md section 5.
Sub-records (outside the headline's timed region): `c3` (1280x720, 2 mm voxels, one GPU: integrate / raycast roofline),
`c4_64` (BASELINE configs[3]: 8 scenes per GPU), `c5` (ITMVoxel_s_rgb + host swapping), `next_rows`, and at
N > 1 `sharded_c3` (one 1280x720 / 2 mm scene spread over all ranks, checked against a single-GPU engine).

`--impl reference` times the reference's own CPU engines (oracle/_ref, built from the unmodified sources with -O3 + OpenMP) on
the box's host cores for the same frames; both arms print the same `config` and the same `workload_stats` keys (visible
blocks mean / last, final pose), so the two lines double as a closed-loop comparison.
"""
from __future__ import annotations

import argparse
import copy
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from infinitam_b200 import synth  # noqa: E402

W, H = 640, 480
L2_FLUSH_BYTES = 256 << 20
METRIC = "fused frames/s (allocate+integrate+raycast+ICP) 640x480"


def headline_config(args):
    """identical in both arms: what is computed, not how it is timed"""
    return {"workload": "configs[1]: synthetic 640x480 analytic-room sequence, 5 mm voxels, mu=0.02, ITMVoxel_s, depth ICP tracker",
            "frames": args.steps, "warmup_frames": args.warmup, "first_frame": 0, "image": "%dx%d" % (W, H),
            "voxel_size_m": 0.005, "mu_m": 0.02, "sdf_local_block_num": 0x10000,
            "l2": "our arm: flushed before every frame (256 MiB write on the engine's stream, outside the timed events); "
                  "the reference arm runs on the host CPU"}


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self._stop.is_set():
                    break
        except Exception:  # noqa: BLE001
            pass

    def wait_first_sample(self, timeout_s=5.0):
        t0 = time.perf_counter()
        while not self.rows and time.perf_counter() - t0 < timeout_s:
            time.sleep(0.01)

    def mark(self):
        """rows sampled from here on belong to the timed region"""
        self.first_timed_row = len(self.rows)

    def stop(self):
        self._stop.set()
        if self.proc:
            self.proc.terminate()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[getattr(self, "first_timed_row", 0):]:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:  # noqa: BLE001
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch and stage, from the newest committed `ncu --set full` summary
    (profiles/rNN_ncu_summary.json, written by tools/ncu_summary.py); {} if there is none."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_summary.json")))
    if not files:
        return {}
    try:
        with open(files[-1]) as f:
            return json.load(f).get("stage_traffic_bytes", {})
    except Exception:  # noqa: BLE001
        return {}


def _dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def _omp_threads(n):
    """OpenMP thread count of the already loaded runtime (the environment variable only counts when libgomp loads)"""
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass


def _pose_list(m):
    return [round(float(x), 6) for x in np.asarray(m, np.float64).reshape(16)]


# ======================================================================================================================
# reference arm

class _c_stdout_to_stderr:
    """The reference's ITMLibSettings constructor prints its tracker type on the C++ stdout; this process's stdout carries
    one JSON line and nothing else, so file descriptor 1 points at stderr while a reference object is constructed."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        os.dup2(self.saved, 1)
        os.close(self.saved)


def _reference_engine():
    from oracle import ref

    flavour = "fast" if ref.available("fast") else ("parity" if ref.available("parity") else None)
    if flavour is None:
        from oracle import port
        return port.PortEngine(W, H), "port", 1, "C restatement, serial"
    cores = (os.cpu_count() or 1) if flavour == "fast" else 1
    os.environ["OMP_NUM_THREADS"] = str(cores)  # torchrun exports OMP_NUM_THREADS=1; libgomp reads it when the library loads
    with _c_stdout_to_stderr():
        eng = ref.RefEngine(W, H, flavour=flavour)
    _omp_threads(cores)
    return eng, "reference", cores, "%s build (-O3 -mavx2 -mfma%s)" % (flavour, " -fopenmp" if flavour == "fast" else "")


def run_reference(args):
    """CPU arm: the reference's own engines on the host cores (rank 0 only)."""
    rank, _, world = _dist_env()
    if rank != 0:
        return
    eng, kind, cores, build = _reference_engine()
    n = args.warmup + args.steps
    frames = synth.sequence(n, W, H)
    for k in range(args.warmup):
        eng.process_frame(frames[k])
    t0 = time.perf_counter()
    stage = np.zeros(6)
    nvis = []
    timed = hasattr(eng, "process_frame_timed")
    for k in range(args.warmup, n):
        if timed:
            stage += np.array(eng.process_frame_timed(frames[k]))
        else:
            eng.process_frame(frames[k])
        nvis.append(int(eng.counters[0]))
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    out = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": headline_config(args),
        "workload_stats": {"visible_blocks_mean": float(np.mean(nvis)), "visible_blocks_last": nvis[-1], "final_pose": _pose_list(eng.pose_M)},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": "%d frames after %d warm-up frames, %s" % (args.steps, args.warmup, build),
                         "stage_ms": {k: float(v / args.steps) for k, v in zip(
                             ["view", "track", "allocate", "integrate", "expected_depths", "raycast_icp_maps"], stage)}},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def cpu_baseline_sample(seconds_budget=10.0):
    """reference CPU engines on a bounded sample of the same workload (rank 0, N=1 only)"""
    try:
        eng, kind, cores, build = _reference_engine()
    except Exception:  # noqa: BLE001
        return None
    warm = 3
    frames = synth.sequence(64, W, H)
    for k in range(warm):
        eng.process_frame(frames[k])
    # about 10 s of CPU work: the 64 frames are walked forwards and backwards (a continuous trajectory either way)
    order = list(range(warm, 64)) + list(range(62, -1, -1)) + list(range(1, 64)) + list(range(62, -1, -1))
    t0 = time.perf_counter()
    done = 0
    for k in order:
        if time.perf_counter() - t0 >= seconds_budget:
            break
        eng.process_frame(frames[k])
        done += 1
    dt = time.perf_counter() - t0
    eng.close()
    return {"value": done / dt, "unit": "frames/s", "cores": cores, "kind": kind,
            "sample": "%d frames of the same sequence (frames 3..63, then back and forth; %.1f s of CPU work), %s" % (done, dt, build)}


# ======================================================================================================================
# our arm: helpers

def _alg_bytes(P, E, nv, ev, voxel_bytes=4):
    """ALGORITHMIC bytes per frame of every stage"""
    return {
        "view": P * (2 + 4) + 4 * P * (1 / 4 + 1 / 16 + 1 / 64 + 1 / 256),
        "track": float(sum(ev[l] * (4 * P / 4 ** l + min(128 * P / 4 ** l, 2 * 16 * P)) for l in range(min(5, len(ev))))),
        "allocate": 4 * P + 3 * E + nv * (16 + 4 + 1),
        "integrate": nv * (2 * 512 * voxel_bytes + 16 + 4) + 4 * P,
        "expected_depths": 8 * P / 64 * 2 + nv * (4 + 16),
        "raycast": 16 * P + nv * (512 * voxel_bytes + 16) + 8 * P / 64,
        "icp_maps": P * (16 + 16 + 16 + 4),
    }


KERNELS = {"view": "k_convert_pyramid", "track": "k_icp_track", "allocate": "k_alloc_pixels+k_alloc_assign+k_visible_merge (compact lists; 1280x720 and swapping / sharded engines: k_alloc_pixels+k_alloc_scan+k_visible_scan)",
           "integrate": "k_integrate_cols", "expected_depths": "k_expected_depths", "raycast": "k_raycast", "icp_maps": "k_icp_maps"}
STAGES = ["view", "track", "allocate", "integrate", "expected_depths", "raycast", "icp_maps", "total"]


def _roofline(alg, stage_avg, peak, peak_kind, traffic, kernels=KERNELS):
    roof = {}
    for name, b in alg.items():
        ms = stage_avg[name]
        ach = b / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        roof[name] = {"kernel": kernels[name], "bound": "hbm", "achieved": ach, "peak": peak, "peak_kind": peak_kind + " (burst copy)", "unit": "GB/s",
                      "frac": ach / peak, "traffic": traffic.get(name), "traffic_note": "dram__bytes_read + dram__bytes_write of one cold-cache ncu "
                      "launch; B200's L2 is write-back, so writes that stay in L2 are not in it", "algorithmic_bytes": int(b), "avg_launch_ms": ms,
                      "share_of_step": ms / stage_avg["total"] if stage_avg["total"] else None}
    return roof


def _device_pass(eng, frames_dev, flush, eng_stream, lo, hi, torch, level):
    """frames [lo, hi) with the frame already in the engine's buffer and the L2 flushed, profiling `level`; returns
    (per-frame total ms list, summed stage ms, visible blocks per frame, ICP evaluations per level summed, last pose)"""
    from infinitam_b200 import capi
    eng.set_profiling(level)
    step_ms, stage_ms, nvis, evals = [], np.zeros(8), [], np.zeros(capi.MAX_LEVELS, np.int64)
    pose = None
    for k in range(lo, hi):
        with torch.cuda.stream(eng_stream):
            flush.fill_(k & 0xFF)
        # the frame is placed in the engine's raw-depth buffer before the timed region ("inputs already resident in HBM")
        eng.EnqueueFrameDevice(eng.PlaceDepthDevice(frames_dev[k].data_ptr()))
        pose, counters = eng.Sync()
        ms = eng.stage_times()
        step_ms.append(float(ms[7]))
        stage_ms += ms
        nvis.append(int(counters[0]))
        evals += eng.icp_stats()
    return step_ms, stage_ms, nvis, evals, pose


def _streaming_pass(eng, rgb_addr, frame_addrs, lo, hi, flush, eng_stream, torch, in_flight=3):
    """host-buffer path, submit / wait with `in_flight` frames queued; an L2 flush is enqueued on the engine's stream between
    frames and its device time (CUDA events on that stream) is subtracted from the host clock.  Returns (seconds, flush seconds)"""
    evs = []
    tickets = []
    t0 = time.perf_counter()
    for k in range(lo, hi):
        tickets.append(eng.SubmitFrame(rgb_addr, frame_addrs[k]))
        if flush is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(eng_stream):
                a.record()
                flush.fill_(k & 0xFF)
                b.record()
            evs.append((a, b))
        if len(tickets) >= in_flight:
            eng.WaitFrame(tickets[len(tickets) - in_flight])
    pose = None
    for t in tickets[-(in_flight - 1):] if in_flight > 1 else []:
        pose, _ = eng.WaitFrame(t)
    eng.Sync()
    dt = time.perf_counter() - t0
    flush_s = sum(a.elapsed_time(b) for a, b in evs) * 1e-3
    return dt, flush_s, pose


# ----------------------------------------------------------------------------------------------------------------------
# sub-records

def c3_record(torch, dev, local_rank, flush, peak, peak_kind, n_frames=22, warm=4):
    """BASELINE configs[2] shape on ONE GPU: 1280x720, 2 mm voxels, enlarged pool.  The working set (~175 MB of voxel blocks
    per frame) exceeds the L2, so this is where integrate / raycast are measured against the HBM roofline."""
    from infinitam_b200 import capi
    from infinitam_b200.engines import ITMMainEngine
    w, h = 1280, 720
    p = capi.default_params(w, h)
    p.voxel_size, p.sdf_local_block_num, p.device = 0.002, 0x80000, local_rank
    seq = torch.from_numpy(synth.sequence(n_frames, w, h)).to(dev)
    out = {"workload": "configs[2] shape on one GPU: synthetic 1280x720, 2 mm voxels, SDF_LOCAL_BLOCK_NUM 0x80000", "frames": n_frames - warm}
    res = {}
    for level in (2, 1):
        eng = ITMMainEngine(p)
        es = torch.cuda.ExternalStream(eng.stream(), device=dev)
        _device_pass(eng, seq, flush, es, 0, warm, torch, level)
        res[level] = _device_pass(eng, seq, flush, es, warm, n_frames, torch, level)
        eng.close()
    step_ms, _, nvis, evals, _ = res[2]
    _, stage_ms, _, _, _ = res[1]
    m = n_frames - warm
    stage_avg = {k: float(v / m) for k, v in zip(STAGES, stage_ms)}
    nv = float(np.mean(nvis))
    alg = _alg_bytes(w * h, p.sdf_bucket_num + p.sdf_excess_list_size, nv, evals / m)
    roof = _roofline(alg, stage_avg, peak, peak_kind, {})
    out.update({"frames_per_s": m / (sum(step_ms) * 1e-3), "ms_per_frame": sum(step_ms) / m, "visible_blocks_mean": nv,
                "gvoxel_updates_per_s": nv * 512 / (stage_avg["integrate"] * 1e-3) / 1e9, "stage_ms": stage_avg,
                "roofline_integrate": roof["integrate"], "roofline_raycast": roof["raycast"]})
    return out


def c4_record(torch, dist, dev, local_rank, world, rank, S, n_frames, warm, frames_np_cache):
    """BASELINE configs[3]: S scenes per GPU (64 across 8 GPUs at S = 8), one engine + stream each, phase-shifted trajectories
    (scene i of rank r starts 7 * (r * S + i) mod 100 frames in).  `frames_per_s`: frames resident in HBM, one host thread
    enqueues all scenes round-robin (one graph launch per frame).  `e2e_frames_per_s`: pinned host buffers through
    submit_frame / wait_frame, round-robin over the scenes with two frames per scene in flight.  Max over ranks."""
    from infinitam_b200 import capi
    from infinitam_b200.engines import ITMMainEngine
    from infinitam_b200.multi import sequence_start_for_scene
    n = warm + n_frames
    starts = [sequence_start_for_scene(rank * S + i) for i in range(S)]
    lo, hi = min(starts), max(starts) + n
    if frames_np_cache is not None and lo >= 0 and hi <= len(frames_np_cache):
        block = frames_np_cache[lo:hi]
    else:
        block = synth.sequence(hi - lo, W, H, start=lo)
    pinned = torch.from_numpy(np.ascontiguousarray(block)).pin_memory()
    on_dev = pinned.to(dev)
    params = capi.default_params(W, H)
    params.device = local_rank
    params.icp_max_ctas = max(1, 148 // S)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    engs = [ITMMainEngine(params) for _ in range(S)]
    for k in range(warm):
        for i, e in enumerate(engs):
            e.EnqueueFrameDevice(on_dev[starts[i] - lo + k].data_ptr())
    for e in engs:
        e.Sync()
    barrier()
    t0 = time.perf_counter()
    for k in range(warm, n):
        for i, e in enumerate(engs):
            e.EnqueueFrameDevice(on_dev[starts[i] - lo + k].data_ptr())
    for e in engs:
        e.Sync()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    for e in engs:
        e.close()
    engs = [ITMMainEngine(params) for _ in range(S)]
    addr = [[pinned[starts[i] - lo + k].data_ptr() for k in range(n)] for i in range(S)]
    for k in range(warm):
        for i, e in enumerate(engs):
            e.WaitFrame(e.SubmitFrame(None, addr[i][k]))
    barrier()
    t0 = time.perf_counter()
    last = [0] * S
    for k in range(warm, n):
        for i, e in enumerate(engs):
            t = e.SubmitFrame(None, addr[i][k])
            if last[i]:
                e.WaitFrame(last[i])
            last[i] = t
    for i, e in enumerate(engs):
        e.WaitFrame(last[i])
    dt_e2e = time.perf_counter() - t0
    barrier()
    for e in engs:
        e.close()
    t = torch.tensor([dt, dt_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    frames = world * S * n_frames
    return {"workload": "configs[3]: %d simultaneous synthetic 640x480 sequences (%d per GPU, phase-shifted), 5 mm voxels, ITMVoxel_s, "
                        "depth ICP tracker" % (world * S, S), "scenes_per_gpu": S, "n_gpus": world, "icp_max_ctas": int(params.icp_max_ctas),
            "frames_per_scene": n_frames, "frames_per_s": frames / float(t[0]), "e2e_frames_per_s": frames / float(t[1]),
            "e2e_h2d_bytes_per_frame": W * H * 2, "l2": "not flushed: %d scenes x ~60 MB of per-frame working set per GPU exceed the 126 MB L2" % S,
            "timing": "host clock from the first enqueue / submit to the last sync / wait, ranks barriered on both sides, max over ranks"}


def c5_record(torch, dev, local_rank, peak, peak_kind, n_frames=44, warm=4):
    """BASELINE configs[4]: ITMVoxel_s_rgb with colour integration and the host swapping engine; the camera leaves the first view
    and comes back (frames 0, 3, 6 .. 96, 93 ..), so blocks are swapped out to the host cache and back in."""
    from infinitam_b200 import capi
    from infinitam_b200.engines import ITMMainEngine
    order = (list(range(0, 99, 3)) + list(range(96, -1, -3)))[:n_frames]
    uniq = sorted(set(order))
    depth = {k: synth.render_depth(k, W, H) for k in uniq}
    yy, xx = np.mgrid[0:H, 0:W]
    rgb_np = np.stack([(xx * 7) & 255, (yy * 5) & 255, ((xx ^ yy) * 3) & 255, np.full_like(xx, 255)], -1).astype(np.uint8)
    rgb = torch.from_numpy(rgb_np).pin_memory()
    dpin = {k: torch.from_numpy(v).pin_memory() for k, v in depth.items()}
    out = {"workload": "configs[4]: 640x480, 5 mm, ITMVoxel_s_rgb (8-byte voxels, colour integration), useSwapping with a host global "
                       "cache; trajectory leaves the first view and returns", "frames": n_frames - warm}
    for swapping in (1, 0):
        p = capi.default_params(W, H)
        p.device, p.voxel_type, p.use_swapping = local_rank, capi.VOXEL_S_RGB, swapping
        eng = ITMMainEngine(p)
        eng.set_profiling(1)
        stage = np.zeros(8)
        nvis, n_in, n_out = [], 0, 0
        t_sum = 0.0
        for i, k in enumerate(order):
            t0 = time.perf_counter()
            eng.ProcessFrame(rgb, dpin[k])
            dt = time.perf_counter() - t0
            if i >= warm:
                t_sum += dt
                stage += eng.stage_times()
                nvis.append(int(eng.Sync()[1][0]))
                if swapping:
                    a, b = eng.swap_counts()
                    n_in += a
                    n_out += b
        m = n_frames - warm
        key = "swapping" if swapping else "no_swapping"
        stage_avg = {s: float(v / m) for s, v in zip(STAGES, stage)}
        nv = float(np.mean(nvis))
        alg_int = nv * (2 * 512 * 8 + 16 + 4) + 4 * W * H + 4 * W * H
        ms = stage_avg["integrate"]
        out[key] = {"frames_per_s": m / t_sum, "timing": "host clock around the blocking ProcessFrame (pinned rgb + depth H2D inside)",
                    "visible_blocks_mean": nv, "stage_ms": stage_avg}
        if swapping:
            out[key].update({"blocks_swapped_in": n_in, "blocks_swapped_out": n_out,
                             "note": "stage_ms.integrate includes the swap-in / swap-out stage: the kernels move blocks to / from the host-mapped cache pool themselves, the frame stays one CUDA graph"})
        else:
            out[key]["roofline_integrate_rgb"] = {"kernel": "k_integrate_rgb", "bound": "hbm", "algorithmic_bytes": int(alg_int), "avg_launch_ms": ms,
                                                  "achieved": alg_int / (ms * 1e-3) / 1e9 if ms else None, "peak": peak, "unit": "GB/s",
                                                  "frac": alg_int / (ms * 1e-3) / 1e9 / peak if ms else None, "peak_kind": peak_kind}
        eng.close()
    return out


def sharded_c3_record(torch, dist, dev, local_rank, world, rank, n_frames=14, warm=3, check_frames=3):
    """BASELINE configs[2]: ONE 1280x720 / 2 mm scene spread over all ranks (infinitam_b200.multi.ShardedEngine): NCCL depth
    broadcast from rank 0, replicated index, voxel payload partitioned into slabs (one-block halo), per-rank partial ray casts
    composed by nearest hit over NVLink peer reads.  First `check_frames` frames with supplied poses are compared with a private
    single-GPU engine on every rank (index bit-identical, resident voxel blocks bit-identical, composed raycast within
    1e-4 m); then free-running timing (CUDA events around broadcast + frame, L2 flushed, max over ranks) next to the single-GPU
    time of the same frames."""
    import copy

    from infinitam_b200 import capi
    from infinitam_b200.engines import ITMMainEngine
    from infinitam_b200.multi import ShardedEngine, compare_scene
    w, h = 1280, 720
    p = capi.default_params(w, h)
    p.voxel_size, p.device = 0.002, local_rank
    p.sdf_local_block_num = 0x80000  # single GPU: the whole scene
    ps = copy.copy(p)
    ps.sdf_local_block_num = max(0x10000, 2 * 0x80000 // world)  # per rank: its slab + halo, with slack for uneven slabs
    n = max(n_frames, check_frames)
    seq = torch.from_numpy(synth.sequence(n, w, h)).to(dev)
    tstream = torch.cuda.Stream(device=dev)
    out = {"workload": "configs[2]: one synthetic 1280x720 scene, 2 mm voxels, spread over %d GPUs (pool per rank 0x%x blocks, single GPU 0x80000)"
                       % (world, ps.sdf_local_block_num), "n_gpus": world}
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    with torch.cuda.stream(tstream):
        # ---- parity: poses supplied on both sides, so that nothing but the sharding differs
        pe, p1 = copy.copy(ps), copy.copy(p)
        pe.tracker_type = p1.tracker_type = capi.TRACKER_EXTERNAL
        eng = ShardedEngine(pe, stream=tstream.cuda_stream)
        single = ITMMainEngine(p1)
        worst = {"raycast_hit_mismatch": 0, "raycast_max_diff_m": 0.0, "raycast_over_1e-4_m": 0, "raycast_unresolved_px": 0, "raycast_px_differing": 0}
        ok = True
        rec = {}
        for k in range(check_frames):
            Mk = np.ascontiguousarray(synth.ground_truth_pose(k).astype(np.float32).T).reshape(16)
            eng.engine.set_state(pose_d=Mk)
            single.set_state(pose_d=Mk)
            eng.EnqueueFrame(seq[k] if rank == 0 else None)
            eng.Sync()
            single.EnqueueFrameDevice(seq[k].data_ptr())
            single.Sync()
            rec = compare_scene(eng.engine, single, rank, world, eng.layout, 0.002, eng.halo)
            ok = ok and rec["hash_pos_offset_equal"] and rec["visible_list_equal"] and rec["residency_matches_ptr"] and rec["resident_voxel_blocks_equal"]
            for key in worst:
                worst[key] = max(worst[key], rec[key])
        single.close()
        eng.close()
        layout = eng.layout
        # ---- timing: free-running ICP; the single-GPU engine first
        single = ITMMainEngine(p)
        single.set_profiling(2)
        es = torch.cuda.ExternalStream(single.stream(), device=dev)
        t1 = 0.0
        for k in range(n):
            with torch.cuda.stream(es):
                flush.fill_(k & 0xFF)
            single.EnqueueFrameDevice(single.PlaceDepthDevice(seq[k].data_ptr()))
            pose_1, _ = single.Sync()
            if k >= warm:
                t1 += float(single.stage_times()[7])
        single.close()
        eng = ShardedEngine(ps, stream=tstream.cuda_stream)
        tot, stages, sh3, nvis, m = 0.0, np.zeros(8), np.zeros(3), 0, 0
        pose_s, cnt = None, None
        for k in range(n):
            flush.fill_(k & 0xFF)
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.EnqueueFrame(seq[k] if rank == 0 else None)
            e1.record()
            pose_s, cnt = eng.Sync()
            if k >= warm:
                tot += e0.elapsed_time(e1)
                nvis += int(cnt[0])
                m += 1
        # stage breakdown: the last frames once more with a stamp at every stage boundary (each stamp costs idle device time,
        # so this pass is not the one that is timed above)
        eng.engine.set_profiling(1)
        ms = 0
        for k in range(max(0, n - 6), n):
            eng.EnqueueFrame(seq[k] if rank == 0 else None)
            eng.Sync()
            stages += eng.engine.stage_times()
            sh3 += eng.engine.shard_times()
            ms += 1
        blocks_used = int(ps.sdf_local_block_num - 1 - cnt[1])
        eng.close()
        # the same frames without the peer-read pass: rays no rank can march on its own voxels stay misses (not bit-identical)
        eng = ShardedEngine(ps, stream=tstream.cuda_stream, peers=False)
        tot_np, m_np = 0.0, 0
        for k in range(n):
            flush.fill_(k & 0xFF)
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.EnqueueFrame(seq[k] if rank == 0 else None)
            e1.record()
            eng.Sync()
            if k >= warm:
                tot_np += e0.elapsed_time(e1)
                m_np += 1
        eng.close()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity
    rot, trans = parity.pose_diff(pose_s, pose_1)
    t = torch.tensor([tot, 0.0 if ok else 1.0, float(worst["raycast_hit_mismatch"]), worst["raycast_max_diff_m"], float(worst["raycast_over_1e-4_m"]),
                      rot, trans, float(blocks_used), float(rec.get("owned_blocks", 0)), float(worst["raycast_unresolved_px"]),
                      float(worst["raycast_px_differing"]), tot_np], dtype=torch.float64, device=dev)
    allr = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
    allr = torch.stack(allr).cpu().numpy()
    names = ["view", "track", "allocate", "integrate", "expected_depths", "raycast+barrier+compose", "icp_maps", "total"]
    tot_max = float(allr[:, 0].max())
    out.update({"slab_layout": {"axis": layout[0], "origin_block": layout[1], "thickness_blocks": layout[2]},
                "index_and_resident_voxels_bit_identical_to_single_gpu": bool(allr[:, 1].max() == 0.0), "checked_frames": check_frames,
                "composed_raycast_vs_single_gpu": {"hit_mask_mismatch_px_max": int(allr[:, 2].max()), "max_point_diff_m": float(allr[:, 3].max()),
                                                   "px_over_1e-4_m_max": int(allr[:, 4].max()), "pixels": w * h,
                                                   "unresolved_px_max": int(allr[:, 9].max()), "px_differing_bitwise_max": int(allr[:, 10].max()),
                                                   "note": "a pixel is unresolved when no rank could march its ray completely on its own voxels; those rays are "
                                                           "marched once more with peer reads of the blocks held elsewhere (frames_per_s), or - without the "
                                                           "peer-read pass - reported as misses (frames_per_s_unresolved_as_misses)"},
                "free_running_pose_diff_after_%d_frames" % n: {"rot_rad": float(allr[:, 5].max()), "trans_m": float(allr[:, 6].max())},
                "voxel_blocks_in_use_per_rank": [int(x) for x in allr[:, 7]], "owned_blocks_per_rank_frame_%d" % (check_frames - 1): [int(x) for x in allr[:, 8]],
                "frames": m, "frames_per_s": m / (tot_max * 1e-3), "ms_per_frame": tot_max / m,
                "frames_per_s_unresolved_as_misses": m_np / (float(allr[:, 11].max()) * 1e-3) if m_np else None,
                "single_gpu_frames_per_s": (n - warm) / (t1 * 1e-3) if t1 else None, "visible_blocks_mean": nvis / m,
                "gvoxel_updates_per_s_all_ranks": (nvis / m) * 512 / (stages[3] / ms * 1e-3) / 1e9 if stages[3] else None,
                "stage_us_rank0": {a: round(1e3 * v / ms, 1) for a, v in zip(names + ["partial_raycast", "barrier_wait", "compose"], list(stages) + list(sh3))},
                "collectives": "NCCL broadcast of the raw depth frame (1.8 MB) from rank 0; one flag barrier + peer reads of the partial raycast tiles "
                               "a rank could not complete itself (NVLink); ICP maps and tracker replicated (no pose broadcast / G-H all-reduce needed)",
                "timing": "CUDA events around NCCL depth broadcast + frame on the shared stream, L2 flushed, max over ranks"})
    return out


def next_rows_sample(params, frames_dev, frames_np, n_frames=30):
    """SURVEY.md 8f rows measured beside the hot path (rank 0, N=1; outside the timed region of the headline numbers):
    useApproximateRaycast frames/s, TRACKER_EXTERNAL frames/s, free-view GetImage and MeshScene on the fused scene, each next to
    the reference CPU engines doing the same on a bounded sample."""
    from infinitam_b200 import capi
    from infinitam_b200.engines import ITMMainEngine

    out = {}
    n = min(n_frames, len(frames_np))
    # --- ForwardRender / useApproximateRaycast
    p2 = copy.copy(params)
    p2.use_approximate_raycast = 1
    eng = ITMMainEngine(p2)
    eng.set_profiling(2)
    ms, n_fwd = [], 0
    for k in range(n):
        eng.EnqueueFrameDevice(frames_dev[k].data_ptr())
        eng.Sync()
        if k >= 5:
            ms.append(float(eng.stage_times()[7]))
            n_fwd += 0 if eng.get_state()[2][4] else 1
    out["approximate_raycast"] = {"frames_per_s": 1e3 * len(ms) / sum(ms), "forward_rendered_frames": n_fwd, "frames": len(ms),
                                  "kernels": "k_track_decide+k_fwd_project+k_fwd_gather+k_fwd_cast+k_fwd_shade"}
    # --- free-view rendering and meshing of the scene fused so far
    K = synth.intrinsics_for(W, H)
    M = np.eye(4, dtype=np.float32)
    M[:3, 3] = [0.1, -0.04, 0.06]
    M = M.T.reshape(16)
    eng.GetImage(capi.IMAGE_FREECAMERA_SHADED, M, K)
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        eng.GetImage(capi.IMAGE_FREECAMERA_SHADED, M, K)
    out["get_image_freecamera"] = {"ms": 1e3 * (time.perf_counter() - t0) / reps, "what": "FindVisibleBlocks + CreateExpectedDepths + RenderImage + D2H of the 640x480 image, host clock",
                                   "kernels": "k_find_visible+k_minmax_init+k_expected_depths+k_render_image"}
    tri = eng.UpdateMesh()
    t0 = time.perf_counter()
    reps = 5
    nt = C.c_uint()
    for _ in range(reps):
        capi.check(eng.lib.itm_b200_engine_mesh_scene(eng.h, None, 0, C.byref(nt)))
    mesh_ms = 1e3 * (time.perf_counter() - t0) / reps
    _, _, st = eng.get_state()
    n_blocks = params.sdf_local_block_num - 1 - int(st[1])
    mesh_bytes = 2 * n_blocks * 2048 + len(tri) * 36 + params.sdf_local_block_num * 32 * 36 + 2 * 16 * (params.sdf_bucket_num + params.sdf_excess_list_size)
    out["mesh_scene"] = {"ms": mesh_ms, "triangles": int(len(tri)), "allocated_blocks": n_blocks, "mtriangles_per_s": len(tri) / mesh_ms / 1e3,
                         "algorithmic_bytes": int(mesh_bytes), "achieved_gbs": mesh_bytes / (mesh_ms * 1e-3) / 1e9,
                         "kernels": "memset+k_find_visible+k_mesh_blocks(count)+k_mesh_scan+k_mesh_blocks(emit)"}
    eng.close()
    # --- the fork's deployment mode: TRACKER_EXTERNAL, pose supplied with the frame, no ICP
    p3 = copy.copy(params)
    p3.tracker_type = capi.TRACKER_EXTERNAL
    eng = ITMMainEngine(p3)
    eng.set_profiling(2)
    ms = []
    for k in range(n):
        Mk = np.ascontiguousarray(synth.ground_truth_pose(k).astype(np.float32).T).reshape(16)
        eng.set_state(pose_d=Mk)
        eng.EnqueueFrameDevice(frames_dev[k].data_ptr())
        eng.Sync()
        if k >= 5:
            ms.append(float(eng.stage_times()[7]))
    out["tracker_external"] = {"frames_per_s": 1e3 * len(ms) / sum(ms), "frames": len(ms),
                               "what": "ground-truth pose set before every frame (pose_d->SetM), fusion without the ICP tracker, device time"}
    eng.close()
    # --- the reference CPU engines on the same scene
    try:
        from oracle import ref
        flavour = "fast" if ref.available("fast") else ("parity" if ref.available("parity") else None)
        if flavour:
            with _c_stdout_to_stderr():
                o = ref.RefEngine(W, H, flavour=flavour)
            o.set_use_approximate_raycast(True)
            m = min(n, 12)
            for k in range(3):
                o.process_frame(frames_np[k])
            t0 = time.perf_counter()
            for k in range(3, m):
                o.process_frame(frames_np[k])
            cpu = {"approximate_raycast_frames_per_s": (m - 3) / (time.perf_counter() - t0)}
            t0 = time.perf_counter()
            o.get_image(capi.IMAGE_FREECAMERA_SHADED, M, K)
            cpu["get_image_freecamera_ms"] = 1e3 * (time.perf_counter() - t0)
            t0 = time.perf_counter()
            tr = o.mesh_scene()
            cpu["mesh_scene_ms"] = 1e3 * (time.perf_counter() - t0)
            cpu["mesh_triangles"] = int(len(tr))
            cpu["cores"] = (os.cpu_count() or 1) if flavour == "fast" else 1
            cpu["sample"] = "%d frames, %s build; MeshScene is serial in the reference" % (m, flavour)
            out["cpu_reference"] = cpu
            o.close()
    except Exception as ex:  # noqa: BLE001
        out["cpu_reference"] = {"error": str(ex)}
    return out


# ======================================================================================================================

def run_ours(args):
    import torch
    import torch.distributed as dist

    from infinitam_b200 import capi
    from infinitam_b200.engines import ITMMainEngine

    rank, local_rank, world = _dist_env()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = capi.load()

    n = args.warmup + args.steps
    # every rank fuses its own scene from the SAME sequence: identical work per rank (the efficiency figure the driver derives
    # from the per-N values then compares like with like)
    frames_np = synth.sequence(n, W, H, start=0)
    frames_pinned = torch.from_numpy(frames_np).pin_memory()
    frames_dev = frames_pinned.to(dev)
    rgb_pinned = torch.full((H, W, 4), 128, dtype=torch.uint8).pin_memory()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    params = capi.default_params(W, H)
    params.device = local_rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first_sample()
    # ---------------------------------------------------------------- pass 1: device-resident input (`value`)
    # The engine runs on its own stream; its own CUDA events (event-record nodes of the frame graph) time each step.  `value`
    # uses the frame's start / end stamps only (profiling level 2); the stage breakdown comes from a second, untimed pass with
    # a stamp at every stage boundary (level 1) - each such stamp is a node between two kernels and costs idle device time.
    # The L2 flush is enqueued on the ENGINE's stream right before the frame, with no host synchronisation in between, so the
    # events measure device time only.
    eng = ITMMainEngine(params)
    eng_stream = torch.cuda.ExternalStream(eng.stream(), device=dev)
    _device_pass(eng, frames_dev, flush, eng_stream, 0, args.warmup, torch, 2)
    barrier()
    sampler.mark()
    launches0 = lib.itm_b200_launch_count()
    step_ms, _, nvis, level_evals, pose_value = _device_pass(eng, frames_dev, flush, eng_stream, args.warmup, n, torch, 2)
    launches = lib.itm_b200_launch_count() - launches0
    barrier()
    total_ms = float(np.sum(step_ms))
    eng.close()
    # stage breakdown (same frames, same flush; not part of `value`)
    eng = ITMMainEngine(params)
    eng_stream = torch.cuda.ExternalStream(eng.stream(), device=dev)
    _device_pass(eng, frames_dev, flush, eng_stream, 0, args.warmup, torch, 1)
    _, stage_ms, _, _, _ = _device_pass(eng, frames_dev, flush, eng_stream, args.warmup, n, torch, 1)
    eng.close()

    # ---------------------------------------------------------------- pass 2: end to end through the host API
    rgb_addr = rgb_pinned.data_ptr()
    frame_addr = [frames_pinned[k].data_ptr() for k in range(n)]  # host addresses of the pinned frames
    eng = ITMMainEngine(params)
    eng_stream = torch.cuda.ExternalStream(eng.stream(), device=dev)
    _streaming_pass(eng, rgb_addr, frame_addr, 0, args.warmup, flush, eng_stream, torch)
    barrier()
    e2e_s, e2e_flush_s, pose_e2e = _streaming_pass(eng, rgb_addr, frame_addr, args.warmup, n, flush, eng_stream, torch)
    barrier()
    eng.close()
    # the blocking call, one frame at a time (round-1 definition), flush outside the per-call clock
    eng = ITMMainEngine(params)
    for k in range(args.warmup):
        eng.ProcessFrame(rgb_pinned, frames_pinned[k])
    blocking_s = 0.0
    for k in range(args.warmup, n):
        flush.fill_(k & 0xFF)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.ProcessFrame(rgb_addr, frame_addr[k])
        blocking_s += time.perf_counter() - t0
    barrier()
    sampler.stop()
    eng.close()

    mine = torch.tensor([total_ms, (e2e_s - e2e_flush_s) * 1e3, blocking_s * 1e3, float(np.mean(nvis)), float(nvis[-1]), float(level_evals.sum()) / args.steps],
                        dtype=torch.float64, device=dev)
    if world > 1:
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        allr = torch.stack(allr).cpu().numpy()
    else:
        allr = mine.cpu().numpy()[None, :]
    total_ms_max, e2e_ms_max, blocking_ms_max = float(allr[:, 0].max()), float(allr[:, 1].max()), float(allr[:, 2].max())

    out = None
    if rank == 0:
        peak, peak_kind = _peaks()
        stage_avg = {k: float(v / args.steps) for k, v in zip(STAGES, stage_ms)}
        P = W * H
        E = params.sdf_bucket_num + params.sdf_excess_list_size
        nv = float(np.mean(nvis))  # mean visible blocks per timed frame
        ev = level_evals / args.steps  # mean ComputeGandH evaluations per frame and pyramid level
        roof = _roofline(_alg_bytes(P, E, nv, ev), stage_avg, peak, peak_kind, _ncu_traffic())
        dominant = max(roof, key=lambda k: roof[k]["avg_launch_ms"])
        cfg = headline_config(args)
        out = {
            "metric": METRIC, "value": world * args.steps / (total_ms_max * 1e-3),
            "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "workload_stats": {"visible_blocks_mean": nv, "visible_blocks_last": int(nvis[-1]), "final_pose": _pose_list(pose_value),
                               "final_pose_e2e_path": _pose_list(pose_e2e), "icp_evaluations_per_frame": float(ev.sum())},
            "method": {"parallelism": "replicas only: one independent scene per GPU (the same sequence on every rank), no collective on the data path",
                       "timing": "value: per-frame CUDA events (frame start / end) on the engine stream, summed, max over ranks; stage_ms from a "
                                 "second pass with a stamp at every stage boundary (its total is stage_ms.total); e2e: host clock around "
                                 "submit_frame / wait_frame with 3 frames in flight minus the device time of the L2 flushes between frames"},
            "gvoxel_updates_per_s": world * nv * 512 / (stage_avg["integrate"] * 1e-3) / 1e9 if stage_avg["integrate"] else None,
            "stage_ms": stage_avg,
            "roofline": roof[dominant],
            "roofline_other": {k: v for k, v in roof.items() if k != dominant},
            "e2e": {"value": world * args.steps / (e2e_ms_max * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": P * 2 + P * 4,
                    "d2h_bytes_per_step": int(C.sizeof(C.c_float) * 16 + 4 * 6 + 4 * 8 + 8),
                    "api": "itm_b200_engine_submit_frame / itm_b200_engine_wait_frame (include/itm_b200.h), pinned host rgb + depth"},
            "e2e_blocking": {"value": world * args.steps / (blocking_ms_max * 1e-3), "unit": "frames/s",
                             "api": "itm_b200_engine_process_frame, one blocking call per frame, host clock per call, summed"},
            "per_rank": [{"rank": r, "value_ms_per_frame": float(allr[r, 0]) / args.steps, "e2e_ms_per_frame": float(allr[r, 1]) / args.steps,
                          "visible_blocks_mean": float(allr[r, 3]), "visible_blocks_last": int(allr[r, 4]),
                          "icp_evaluations_per_frame": float(allr[r, 5])} for r in range(world)],
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
    # ---------------------------------------------------------------- sub-records (outside the headline's timed region)
    sub = {}
    if not args.no_sub:
        try:
            c4 = c4_record(torch, dist, dev, local_rank, world, rank, 8, min(args.steps, 30), 3, frames_np)
            sub["c4_64"] = c4
        except Exception as ex:  # noqa: BLE001
            sub["c4_64"] = {"error": repr(ex)}
        if world > 1:
            try:
                sub["sharded_c3"] = sharded_c3_record(torch, dist, dev, local_rank, world, rank)
            except Exception as ex:  # noqa: BLE001
                sub["sharded_c3"] = {"error": repr(ex)}
        elif rank == 0:
            peak, peak_kind = _peaks()
            for name, fn in (("c3", lambda: c3_record(torch, dev, local_rank, flush, peak, peak_kind)),
                             ("c5", lambda: c5_record(torch, dev, local_rank, peak, peak_kind))):
                try:
                    sub[name] = fn()
                except Exception as ex:  # noqa: BLE001
                    sub[name] = {"error": repr(ex)}
    if rank == 0:
        out.update(sub)
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline_sample()
        if world == 1 and not args.no_next_rows:
            try:
                out["next_rows"] = next_rows_sample(params, frames_dev, frames_np)
            except Exception as ex:  # noqa: BLE001
                out["next_rows"] = {"error": repr(ex)}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_batched(args):
    """--scenes-per-gpu S as the headline (BASELINE configs[3])"""
    import torch
    import torch.distributed as dist

    rank, local_rank, world = _dist_env()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first_sample()
    sampler.mark()
    rec = c4_record(torch, dist, dev, local_rank, world, rank, args.scenes_per_gpu, args.steps, args.warmup, None)
    sampler.stop()
    if rank == 0:
        out = {"metric": METRIC, "value": rec["frames_per_s"], "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": 1e3 * world * args.scenes_per_gpu / rec["frames_per_s"], "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {k: rec[k] for k in ("workload", "scenes_per_gpu", "icp_max_ctas", "frames_per_scene", "l2")},
               "method": {"timing": rec["timing"]},
               "e2e": {"value": rec["e2e_frames_per_s"], "unit": "frames/s", "h2d_bytes_per_step": args.scenes_per_gpu * W * H * 2,
                       "d2h_bytes_per_step": args.scenes_per_gpu * 128}, "clocks": sampler.summary()}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=95)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--scenes-per-gpu", type=int, default=1,
                    help="BASELINE configs[3] as the headline: that many independent scenes per GPU (default 1 = configs[1])")
    ap.add_argument("--no-next-rows", action="store_true", help="skip the SURVEY 8f rows (approximate raycast, GetImage, MeshScene)")
    ap.add_argument("--no-sub", action="store_true", help="skip the c3 / c4_64 / c5 / sharded_c3 sub-records")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.scenes_per_gpu > 1:
        run_batched(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
