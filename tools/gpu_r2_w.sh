#!/bin/bash
# queue pipeline: ticket size while the queue drains (MTB_QUEUE_DRAIN = 0 / 1 / 2), shares 1/1, 1/4, 1/8
run() { timeout -s KILL 200 python tools/half_frame.py $1 $2 2>&1 | tail -n 2 | python -c "
import sys, json
o = []
for l in sys.stdin:
    try: j = json.loads(l); o.append('%.2f' % j['ms'])
    except Exception: o.append(l.strip()[:80])
print(' '.join(o))"; }
for v in d0 default d2; do
  if [ $v = default ]; then unset MTB_LIB_PATH; else export MTB_LIB_PATH=$PWD/mythtracer_b200/build/var_$v/lib.so; fi
  for cfg in "1 queue" "4 queue" "8 queue"; do echo "== $v share 1/$cfg: $(run $cfg)"; done
done
