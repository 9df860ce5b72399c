#!/bin/bash
# round 2, GPU call F: spatial splits (C4), parity suite, wavefront occupancy variants
mkdir -p gpurun_out
timeout 400 python tools/quick_time.py C4,C1,C2,C3,C5 bvh > gpurun_out/r2f_qt_all.log 2>&1; echo "== all configs, default split"; cut -c1-60 gpurun_out/r2f_qt_all.log; grep -o '"per_ray.*"literal' gpurun_out/r2f_qt_all.log
for div in 0 16 64 128; do
  MTB_SPLIT_DIV=$div timeout 300 python tools/quick_time.py C4,C3 bvh > gpurun_out/r2f_qt_div$div.log 2>&1; echo "== split div $div"; cut -c1-60 gpurun_out/r2f_qt_div$div.log; grep -o '"per_ray.*"literal' gpurun_out/r2f_qt_div$div.log
done
for v in wfblk6 wfblk7 wfblk10 wfblk12; do
  MTB_LIB_PATH=mythtracer_b200/build/var_$v/lib.so timeout 200 python tools/quick_time.py C3 wf > gpurun_out/r2f_qt_$v.log 2>&1; echo "== $v"; cut -c1-60 gpurun_out/r2f_qt_$v.log
  MTB_LIB_PATH=mythtracer_b200/build/var_$v/lib.so timeout 200 python tools/half_frame.py 8 wf > gpurun_out/r2f_half8_$v.log 2>&1; tail -n 2 gpurun_out/r2f_half8_$v.log
done
timeout 200 python tools/quick_time.py C3 wf > gpurun_out/r2f_qt_wfdefault.log 2>&1; echo "== wf default"; cut -c1-60 gpurun_out/r2f_qt_wfdefault.log
timeout 200 python tools/half_frame.py 8 wf | tail -n 2
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/r2f_pytest.log
