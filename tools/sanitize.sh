#!/bin/bash
# compute-sanitizer passes over the hot path (memcheck, racecheck, synccheck, initcheck); logs under gpurun_out/.
#   tools/sanitize.sh            single GPU: 3 frames 320x240 through ProcessFrame, the streaming API and 4 concurrent scenes
#   tools/sanitize.sh sharded    2 ranks: 2 frames of a sharded 320x240 scene (memcheck + racecheck, all processes)
set -u
OUT=gpurun_out
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
if [ "${1:-}" = "sharded" ]; then
  for tool in memcheck racecheck; do
    $CS --tool $tool --target-processes all --print-limit 20 --log-file $OUT/r2_sanitizer_sharded_${tool}_%p.log \
      python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
      tools/sharded_run.py --frames 2 --warmup 1 --size 320x240 --voxel 0.005 --pool 0x4000 > $OUT/r2_sanitizer_sharded_${tool}.out 2>&1
    echo "sharded $tool rc=$?"
    grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" $OUT/r2_sanitizer_sharded_${tool}_*.log | sort | uniq -c
  done
  exit 0
fi
for tool in memcheck racecheck synccheck initcheck; do
  $CS --tool $tool --print-limit 20 --log-file $OUT/r2_sanitizer_${tool}.log python tools/sanitize_workload.py > $OUT/r2_sanitizer_${tool}.out 2>&1
  echo "$tool rc=$? $(grep -h 'ERROR SUMMARY\|RACECHECK SUMMARY' $OUT/r2_sanitizer_${tool}.log | tail -1)"
done
