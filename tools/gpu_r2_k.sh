#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for v in default ld128; do
  if [ $v = default ]; then unset MTB_LIB_PATH; else export MTB_LIB_PATH=mythtracer_b200/build/var_$v/lib.so; fi
  timeout 200 python tools/quick_time.py C3 bvh,wf > gpurun_out/r2k_qt_${v}_$rep.log 2>&1
  echo "== $v #$rep"; cut -c1-60 gpurun_out/r2k_qt_${v}_$rep.log
done; done
unset MTB_LIB_PATH
timeout 300 python tools/quick_time.py C2,C4,C5 bvh 2>&1 | cut -c1-60
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/r2k_pytest.log
