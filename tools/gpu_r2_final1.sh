#!/bin/bash
# round 2, final single-GPU capture: parity suite, bench line (both arms), launch list of the bench command, --set full of RenderMega
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r2_final_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2_final_bench_n1.json; tail -n 3 gpurun_out/r2_final_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_reference_arm.json 2> gpurun_out/r2_final_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2_final_bench_reference_arm.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-workloads --pipeline mega"
timeout 300 $CMD > gpurun_out/r2_final_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_final_launches_bench.csv $CMD > gpurun_out/r2_final_ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 200 python tools/quick_time.py C3 bvh > gpurun_out/r2_final_plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:RenderMega -s 2 -c 1 -o gpurun_out/r2_final_prof_mega python tools/quick_time.py C3 bvh > gpurun_out/r2_final_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | grep r2_final
