#!/bin/bash
# round 2, GPU call J: steered hybrid split at 1/1 .. 1/8 of the frame, parity of the hybrid tests
mkdir -p gpurun_out
for w in 1 2 4 8; do for m in mega wf hybrid auto; do
  echo "== share 1/$w $m"; timeout 200 python tools/half_frame.py $w $m 2>&1 | tail -n 2
done; done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "hybrid or automatic or launch_forms or wavefront" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r2j_pytest.log
