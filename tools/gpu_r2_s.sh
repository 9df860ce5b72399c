#!/bin/bash
# queue pipeline (one persistent kernel over a single ray queue): parity first, bounded; then timings by share
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "queue" > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/r2s_pytest.log
run() { timeout -s KILL 200 python tools/half_frame.py $1 $2 2>&1 | tail -n 2 | python -c "
import sys, json
o = []
for l in sys.stdin:
    try: j = json.loads(l); o.append('%.2f' % j['ms'])
    except Exception: o.append(l.strip()[:80])
print(' '.join(o))"; }
for cfg in "1 mega" "1 wf" "1 queue" "2 queue" "4 queue" "8 queue" "8 wf" "8 hybrid"; do echo "== share 1/$cfg: $(run $cfg)"; done
