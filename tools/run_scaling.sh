#!/bin/bash
# bench.py at N = 1, 2, 4, 8 on one box (needs `gpurun --gpus 8`); results -> gpurun_out/scale_*.json
set -u
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/scale_n$n.json"))
    print("N=%d pipeline=%s value=%.1f Mrays/s ms/step=%.2f e2e_ms=%.2f launches=%d kernel_ms=%.2f" % (d["n_gpus"], d["config"]["pipeline"], d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["roofline"]["kernel_ms"]))
except Exception as e:
    print("N=$n failed:", e)
PY
done
