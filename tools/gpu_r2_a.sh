#!/bin/bash
# round 2, GPU call A: parity suite, variant timings, bench line, one full ncu capture of RenderMega
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 25 gpurun_out/r2a_pytest.log
timeout 300 python tools/quick_time.py C3 bvh,wf,hybrid > gpurun_out/r2a_qt_default.log 2>&1; cat gpurun_out/r2a_qt_default.log | cut -c1-400
for v in stack0 stack4 blk12 blk20 stack0blk20; do
  MTB_LIB_PATH=mythtracer_b200/build/var_$v/lib.so timeout 200 python tools/quick_time.py C3 bvh > gpurun_out/r2a_qt_$v.log 2>&1
  echo "== $v"; cut -c1-300 gpurun_out/r2a_qt_$v.log
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2a_bench.json; tail -n 5 gpurun_out/r2a_bench.err
timeout 200 python tools/quick_time.py C3 bvh > gpurun_out/r2a_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:RenderMega -s 2 -c 1 -o gpurun_out/r2a_prof_mega python tools/quick_time.py C3 bvh > gpurun_out/r2a_ncu.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out | tail -n 15
