#!/bin/bash
# round 2, GPU call G: device-built scene BVH (PLOC) vs host SAH, load stages, wavefront packing knob, parity suite
mkdir -p gpurun_out
export MTB_TIMING=1
timeout 400 python tools/quick_time.py C3,C2,C4,C5,C1 bvh > gpurun_out/r2g_qt_devbvh.log 2> gpurun_out/r2g_qt_devbvh.err; echo "== device BVH"; cut -c1-60 gpurun_out/r2g_qt_devbvh.log; grep -o '"per_ray.*"literal' gpurun_out/r2g_qt_devbvh.log; grep -o '"load": {.*' gpurun_out/r2g_qt_devbvh.log; grep "mtb\]" gpurun_out/r2g_qt_devbvh.err | head -40
timeout 400 python tools/quick_time.py C3,C2,C4,C5,C1 hostbvh > gpurun_out/r2g_qt_hostbvh.log 2> gpurun_out/r2g_qt_hostbvh.err; echo "== host BVH"; cut -c1-60 gpurun_out/r2g_qt_hostbvh.log; grep -o '"per_ray.*"literal' gpurun_out/r2g_qt_hostbvh.log; grep -o '"load": {.*' gpurun_out/r2g_qt_hostbvh.log
unset MTB_TIMING
for l in 4 8 16 32; do echo "== wf min lanes $l"; MTB_WF_MIN_LANES=$l timeout 200 python tools/half_frame.py 8 wf | tail -n 2; MTB_WF_MIN_LANES=$l timeout 200 python tools/half_frame.py 4 wf | tail -n 1; done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/r2g_pytest.log
