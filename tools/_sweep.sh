for d in 0 2 4 6 10 16; do
  ITM_B200_DEFINES="-DRAY_PF_DIST=$d" python -m infinitam_b200.build --force > /dev/null 2>&1
  python tools/stage_bench.py 40 640 480 "pf=$d"
done
