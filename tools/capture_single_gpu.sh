#!/bin/bash
# single-GPU capture at HEAD: parity suite, smoke, bench line (both arms), ncu launch lists of the bench command
# (megakernel / queue pipeline) and of a 1/8 share, --set full of RenderMega and of WfQueue.   usage: capture_single_gpu.sh [tag]
TAG=${1:-cap}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -n 1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_reference_arm.json
for p in mega queue; do
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-workloads --pipeline $p"
  timeout 300 $CMD > gpurun_out/${TAG}_plain_$p.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches_bench_$p.csv $CMD > gpurun_out/${TAG}_ncu_launch_$p.log 2>&1
  echo "launch list ($p) rc=$?"
done
timeout 200 python tools/half_frame.py 8 queue > gpurun_out/${TAG}_plain_eighth.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches_queue_eighth.csv python tools/half_frame.py 8 queue > gpurun_out/${TAG}_ncu_eighth.log 2>&1
echo "launch list (1/8) rc=$?"
timeout 200 python tools/quick_time.py C3 bvh,queue > gpurun_out/${TAG}_plain_qt.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k RenderMega -s 2 -c 1 -o gpurun_out/${TAG}_prof_mega python tools/quick_time.py C3 bvh > gpurun_out/${TAG}_ncu_full_mega.log 2>&1
echo "ncu full (mega) rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k WfQueue -s 2 -c 1 -o gpurun_out/${TAG}_prof_queue python tools/quick_time.py C3 queue > gpurun_out/${TAG}_ncu_full_queue.log 2>&1
echo "ncu full (queue) rc=$?"
cut -c1-200 gpurun_out/${TAG}_plain_qt.log
ls -la gpurun_out | grep ${TAG}
