#!/bin/bash
mkdir -p gpurun_out
for r in 16 32 64 100 128; do
  echo "== PLOC radius $r"
  MTB_PLOC_RADIUS=$r timeout 300 python tools/quick_time.py C3,C2,C4 bvh > gpurun_out/r2h_qt_r$r.log 2>&1; cut -c1-60 gpurun_out/r2h_qt_r$r.log; grep -o '"bvh": [0-9.]*' gpurun_out/r2h_qt_r$r.log | tr '\n' ' '; grep -o '"scene_bvh_device_ms": [0-9.]*' gpurun_out/r2h_qt_r$r.log | tr '\n' ' '; echo
done
