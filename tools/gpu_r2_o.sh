#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "hybrid or automatic or launch_forms or wavefront or c1_full" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r2o_pytest.log
for w in 1 2 4 8; do for m in hybrid auto; do echo "== share 1/$w $m"; timeout 200 python tools/half_frame.py $w $m 2>&1 | tail -n 2; done; done
