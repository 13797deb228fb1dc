#!/bin/bash
# Round-end profiling pass (one GPU): launch list of a short bench run, `ncu --set full` of one steady-state VGA frame and of
# the 1280x720 / 2 mm integration + raycast kernels.  Reports land in gpurun_out/; condense with tools/ncu_summary.py.
#   tools/ncu_capture.sh <tag>      e.g. r02b
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 8 --warmup 3 --no-cpu --no-next-rows --no-sub > $OUT/${TAG}_bench_under_ncu.log 2>&1
# frame 9 of a 12-frame run: 8 kernels in the first frame (no tracker), 9 per frame afterwards
ncu --set full --clock-control none --import-source on -k regex:"k_convert|k_icp_track|k_alloc|k_visible|k_integrate|k_expected|k_raycast|k_icp_maps" \
  --launch-skip 72 --launch-count 9 -f -o $OUT/${TAG}_full python tools/profile_run.py 12 > $OUT/${TAG}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_integrate|k_raycast|k_alloc|k_visible" \
  --launch-skip 40 --launch-count 5 -f -o $OUT/${TAG}_c3 python tools/profile_run.py 12 1280 720 voxel=0.002 pool=0x80000 > $OUT/${TAG}_ncu_c3.log 2>&1
for r in full c3; do ncu -i $OUT/${TAG}_$r.ncu-rep --page raw --csv > $OUT/${TAG}_$r.csv 2>/dev/null; done
ls -la $OUT/${TAG}_*
