#!/bin/bash
# A/B of development variants (mythtracer_b200/build/var_<name>/lib.so, see build.py) on one GPU:
#   tools/gpu_ab.sh <tag> <variant> [<variant> ...]   ("base" = the in-tree library)
# per variant: quick_time on C3 / C5 / C4 (megakernel and queue pipeline), 1/8 share of C3 through the queue pipeline;
# then the GPU parity suites with the LAST variant.
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then unset MTB_LIB_PATH; else export MTB_LIB_PATH=$PWD/mythtracer_b200/build/var_$v/lib.so; fi
  timeout 300 python tools/quick_time.py C3,C5,C4 bvh,queue > gpurun_out/${tag}_qt_$v.log 2> gpurun_out/${tag}_qt_$v.err
  echo "== $v rc=$?"; python - gpurun_out/${tag}_qt_$v.log <<'P'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["config"], d["mode"], d["kernel_ms"], "bvh/ray", d["per_ray"]["bvh"], "triaabb", d["per_ray"]["triaabb"], "mt", d["per_ray"]["mt"], "fallback", d["fallback"])
P
  timeout 200 python tools/half_frame.py 8 queue 2>&1 | tail -n 2
done
last="${@: -1}"
if [ "$last" != base ]; then
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/${tag}_pytest_$last.log 2>&1; echo "pytest($last) rc=$?"; tail -n 3 gpurun_out/${tag}_pytest_$last.log
fi
