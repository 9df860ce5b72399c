#!/bin/bash
# queue pipeline integrated: full GPU suite, automatic choice by share, occupancy variants of WfQueue
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r2t_pytest.log
run() { timeout -s KILL 200 python tools/half_frame.py $1 $2 2>&1 | tail -n 2 | python -c "
import sys, json
o = []
for l in sys.stdin:
    try: j = json.loads(l); o.append('%.2f' % j['ms'])
    except Exception: o.append(l.strip()[:80])
print(' '.join(o))"; }
for cfg in "1 auto" "2 auto" "4 auto" "8 auto"; do echo "== share 1/$cfg: $(run $cfg)"; done
for v in q10 q12 q16; do
  export MTB_LIB_PATH=$PWD/mythtracer_b200/build/var_$v/lib.so
  for cfg in "1 queue" "8 queue"; do echo "== $v share 1/$cfg: $(run $cfg)"; done
done
unset MTB_LIB_PATH
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads > gpurun_out/r2t_bench.json 2>gpurun_out/r2t_bench.err; python -c "
import json; j=json.load(open('gpurun_out/r2t_bench.json')); print('bench', j['value'], j['ms_per_step'], 'kernel', j['roofline']['kernel_ms'], 'e2e', j['e2e']['ms_per_step'], j['config']['pipeline'], j['frame_equals_reference'])"
