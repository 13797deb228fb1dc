#!/usr/bin/env python
"""Condenses an `ncu --set full` report into the tracked summaries under profiles/.

  ncu -i gpurun_out/rNN_full.ncu-rep --page raw --csv > /tmp/raw.csv
  python tools/ncu_summary.py /tmp/raw.csv profiles/rNN_ncu_summary

writes <out>.md (one table row per captured launch) and <out>.json (per-kernel means and, per bench stage, the DRAM
traffic per launch that bench.py reports as roofline.traffic)."""
import csv
import json
import sys
from collections import OrderedDict

METRICS = OrderedDict([
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("lts__t_sectors_op_read.sum", "l2_rd_sectors"),
    ("lts__t_sectors_op_write.sum", "l2_wr_sectors"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_tput_pct"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "mem_tput_pct"),
])
STAGE_OF = {"k_convert_pyramid": "view", "k_icp_track": "track", "k_mark_prev_visible": "allocate", "k_alloc_pixels": "allocate",
            "k_alloc_scan": "allocate", "k_visible_scan": "allocate", "k_alloc_assign": "allocate", "k_visible_merge": "allocate", "k_integrate": "integrate", "k_minmax_init": "expected_depths",
            "k_expected_depths": "expected_depths", "k_raycast": "raycast", "k_icp_maps": "icp_maps",
            # SURVEY 8f rows (not stages of the bench step; listed for the per-kernel table only)
            "k_fwd_project": None, "k_fwd_gather": None, "k_fwd_cast": None, "k_fwd_shade": None, "k_track_decide": None,
            "k_find_visible": None, "k_render_image": None, "k_mesh_blocks": None, "k_mesh_scan": None}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    raws, out = sys.argv[1:-1], sys.argv[-1]
    raw = ", ".join(raws)
    launches = []
    for one in raws:
        rows = list(csv.reader(open(one)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        idx = {h: i for i, h in enumerate(hdr)}
        launches += parse(hdr, units, data, idx)
    finish(raw, out, launches)


def parse(hdr, units, data, idx):
    launches = []
    for d in data:
        name = d[idx["Kernel Name"]]
        short = next((k for k in STAGE_OF if k in name), name[:40])
        rec = OrderedDict(kernel=short)
        for m, label in METRICS.items():
            if m not in idx:
                continue
            v, u = d[idx[m]], units[idx[m]]
            if label.endswith("_MB"):
                rec[label] = round(to_bytes(v, u) / 1e6, 3)
            elif label == "time_us":
                f = float(v.replace(",", ""))
                rec[label] = round(f * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1), 2)
            else:
                f = float(v.replace(",", ""))
                rec[label] = round(f, 2) if f < 1e6 else int(f)
        launches.append(rec)
    return launches


def finish(raw, out, launches):
    per_kernel = OrderedDict()
    for r in launches:
        per_kernel.setdefault(r["kernel"], []).append(r)
    means = OrderedDict()
    for k, rs in per_kernel.items():
        m = OrderedDict(launches=len(rs))
        for label in rs[0]:
            if label == "kernel":
                continue
            m[label] = round(sum(float(r[label]) for r in rs) / len(rs), 3)
        means[k] = m
    stage_traffic = {}
    for k, m in means.items():
        st = STAGE_OF.get(k)
        if st:
            stage_traffic[st] = stage_traffic.get(st, 0) + int((m.get("dram_rd_MB", 0) + m.get("dram_wr_MB", 0)) * 1e6)
    json.dump({"source": raw, "note": "ncu --set full --clock-control none; ncu flushes caches before every replay pass, so DRAM bytes are "
               "cold-cache figures", "kernels": means, "stage_traffic_bytes": stage_traffic}, open(out + ".json", "w"), indent=1)
    cols = ["kernel"] + [l for l in METRICS.values() if l in launches[0]]
    with open(out + ".md", "w") as f:
        f.write("ncu --set full --clock-control none (cold caches per replay pass); one row per captured launch\n\n")
        f.write("| " + " | ".join(cols) + " |\n|" + "---|" * len(cols) + "\n")
        for r in launches:
            f.write("| " + " | ".join(str(r.get(c, "")) for c in cols) + " |\n")
    print(json.dumps(stage_traffic))


if __name__ == "__main__":
    main()
