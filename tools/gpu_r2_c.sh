#!/bin/bash
# round 2, GPU call C: same-box A/B of the epilogue / leaf preload / blocks per SM, ray-length distribution
mkdir -p gpurun_out
for rep in 1 2; do
for v in default blockbar nopreload blk14 blk14bar; do
  if [ $v = default ]; then unset MTB_LIB_PATH; else export MTB_LIB_PATH=mythtracer_b200/build/var_$v/lib.so; fi
  timeout 200 python tools/quick_time.py C3 bvh > gpurun_out/r2c_qt_${v}_$rep.log 2>&1
  echo "== $v #$rep"; cut -c1-60 gpurun_out/r2c_qt_${v}_$rep.log; grep -o '"all_ms.*' gpurun_out/r2c_qt_${v}_$rep.log | cut -c1-400
done; done
unset MTB_LIB_PATH
timeout 200 python tools/quick_time.py C3 noorder > gpurun_out/r2c_qt_noorder.log 2>&1; echo "== noorder"; cut -c1-60 gpurun_out/r2c_qt_noorder.log
timeout 300 python tools/quick_time.py C2,C5 bvh > gpurun_out/r2c_qt_c2c5.log 2>&1; echo "== C2 C5"; cut -c1-60 gpurun_out/r2c_qt_c2c5.log; grep -o '"long128.*' gpurun_out/r2c_qt_c2c5.log
