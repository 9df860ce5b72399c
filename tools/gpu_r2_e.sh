#!/bin/bash
# round 2, GPU call E: launch lists of the wavefront pipeline (1/8 of the frame, full frame) and --set full of its two traversal kernels
mkdir -p gpurun_out
timeout 200 python tools/share_launches.py 8 wf > gpurun_out/r2e_plain8.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.per_cycle_active --clock-control none -s 54 -c 27 --csv --log-file gpurun_out/r2e_launches_wf_eighth.csv python tools/share_launches.py 8 wf > gpurun_out/r2e_ncu8.log 2>&1
echo "ncu 1/8 rc=$?"
timeout 200 python tools/share_launches.py 1 wf > gpurun_out/r2e_plain1.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.per_cycle_active --clock-control none -s 54 -c 27 --csv --log-file gpurun_out/r2e_launches_wf_full.csv python tools/share_launches.py 1 wf > gpurun_out/r2e_ncu1.log 2>&1
echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"WfTrace|WfShadow" -s 24 -c 4 -o gpurun_out/r2e_prof_wf_eighth python tools/share_launches.py 8 wf > gpurun_out/r2e_ncu_full.log 2>&1
echo "ncu set full rc=$?"
ls -la gpurun_out | grep r2e
