#!/bin/bash
mkdir -p gpurun_out
nproc
export MTB_TIMING=1
for rep in 1 2; do
timeout 400 python tools/quick_time.py C3 bvh > gpurun_out/r2i_qt_host_$rep.log 2> gpurun_out/r2i_qt_host_$rep.err; echo "== host build #$rep"; cut -c1-60 gpurun_out/r2i_qt_host_$rep.log; grep -o '"load": {.*' gpurun_out/r2i_qt_host_$rep.log; grep "mtb\]" gpurun_out/r2i_qt_host_$rep.err | head -12
timeout 400 python tools/quick_time.py C3 devbvh > gpurun_out/r2i_qt_dev_$rep.log 2> gpurun_out/r2i_qt_dev_$rep.err; echo "== device build #$rep"; cut -c1-60 gpurun_out/r2i_qt_dev_$rep.log; grep -o '"load": {.*' gpurun_out/r2i_qt_dev_$rep.log
done
timeout 400 python tools/quick_time.py C4,C2 bvh > gpurun_out/r2i_qt_host_c4.log 2>/dev/null; echo "== host build C4 C2"; cut -c1-60 gpurun_out/r2i_qt_host_c4.log; grep -o '"load": {.*' gpurun_out/r2i_qt_host_c4.log
timeout 400 python tools/quick_time.py C4,C2 devbvh > gpurun_out/r2i_qt_dev_c4.log 2>/dev/null; echo "== device build C4 C2"; cut -c1-60 gpurun_out/r2i_qt_dev_c4.log; grep -o '"load": {.*' gpurun_out/r2i_qt_dev_c4.log
unset MTB_TIMING
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r2i_pytest.log
