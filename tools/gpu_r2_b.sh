#!/bin/bash
# round 2, GPU call B: variant timings of the restructured traversal, share-of-frame timings, parity suite
mkdir -p gpurun_out
timeout 300 python tools/quick_time.py C3 bvh,wf,hybrid > gpurun_out/r2b_qt_default.log 2>&1; echo "== default"; cut -c1-200 gpurun_out/r2b_qt_default.log
for v in ray0 ray1stack8 pf pfray0 blk20 blk14; do
  MTB_LIB_PATH=mythtracer_b200/build/var_$v/lib.so timeout 200 python tools/quick_time.py C3 bvh > gpurun_out/r2b_qt_$v.log 2>&1
  echo "== $v"; cut -c1-200 gpurun_out/r2b_qt_$v.log
done
for mb in 24 64; do
  MTB_L2_PERSIST_MB=$mb timeout 200 python tools/quick_time.py C3 bvh > gpurun_out/r2b_qt_l2_$mb.log 2>&1
  echo "== l2 persist $mb"; cut -c1-200 gpurun_out/r2b_qt_l2_$mb.log
done
for w in 2 4 8; do for m in mega wf hybrid; do
  timeout 200 python tools/half_frame.py $w $m > gpurun_out/r2b_half_${w}_$m.log 2>&1; echo "== share 1/$w $m"; cat gpurun_out/r2b_half_${w}_$m.log | tail -n 2
done; done
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/r2b_pytest.log
