#!/bin/bash
# round 2: ncu evidence for the queue pipeline (launch lists of the bench command and of a 1/8 share, --set full of WfQueue)
mkdir -p gpurun_out
CMDQ="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-workloads --pipeline queue"
timeout 300 $CMDQ > gpurun_out/r2v_plain_queue.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2v_launches_bench_queue.csv $CMDQ > gpurun_out/r2v_ncu_launch_queue.log 2>&1
echo "launch list (queue) rc=$?"
CMDM="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-workloads --pipeline mega"
timeout 300 $CMDM > gpurun_out/r2v_plain_mega.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2v_launches_bench_mega.csv $CMDM > gpurun_out/r2v_ncu_launch_mega.log 2>&1
echo "launch list (mega) rc=$?"
timeout 200 python tools/half_frame.py 8 queue > gpurun_out/r2v_plain_eighth.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2v_launches_queue_eighth.csv python tools/half_frame.py 8 queue > gpurun_out/r2v_ncu_eighth.log 2>&1
echo "launch list (1/8) rc=$?"
timeout 200 python tools/quick_time.py C3 queue > gpurun_out/r2v_plain_qt.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:WfQueue -s 2 -c 1 -o gpurun_out/r2v_prof_queue python tools/quick_time.py C3 queue > gpurun_out/r2v_ncu_full_queue.log 2>&1
echo "ncu full (queue) rc=$?"
tail -n 2 gpurun_out/r2v_plain_qt.log | cut -c1-300
ls -la gpurun_out | grep r2v
