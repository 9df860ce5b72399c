#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/quick_time.py C3 bvh,hybrid,wf 2>&1 | cut -c1-60
for w in 2 8; do for m in mega hybrid auto; do echo "== share 1/$w $m"; timeout 200 python tools/half_frame.py $w $m 2>&1 | tail -n 2; done; done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-workloads --pipeline mega"
timeout 300 $CMD > gpurun_out/r2l_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/r2l_launches_bench.csv $CMD > gpurun_out/r2l_ncu_launch.log 2>&1
grep -E "BuildTileOrder|RenderMega" gpurun_out/r2l_launches_bench.csv | awk -F'","' '{print $5, $NF}' | head -12
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads > gpurun_out/r2l_bench.json 2>/dev/null; python -c "
import json; j=json.load(open('gpurun_out/r2l_bench.json')); print('bench', j['value'], j['ms_per_step'], 'kernel', j['roofline']['kernel_ms'], 'e2e', j['e2e']['ms_per_step'], j['config']['pipeline'], j['frame_equals_reference'])"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2l_pytest.log
