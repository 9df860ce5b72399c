#!/bin/bash
# 8-GPU box: torchrun bench at N = 8 / 4 / 2 (tiles stored straight into rank 0's frame), and the one-process form.
#   usage: capture_scaling_8gpu.sh [tag]
TAG=${1:-scal}
mkdir -p gpurun_out
run() { # name nproc steps
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $2 --steps $3 --warmup 5 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err; echo "$1 rc=$?"
}
run n8 8 20
run n4 4 20
run n2 2 20
timeout 900 python bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_inproc8.json 2> gpurun_out/${TAG}_inproc8.err; echo "inproc8 rc=$?"
timeout 600 python tools/inproc_scaling.py C5 1,8 > gpurun_out/${TAG}_inproc_c5.jsonl 2> gpurun_out/${TAG}_inproc_c5.err; echo "inproc C5 rc=$?"; cat gpurun_out/${TAG}_inproc_c5.jsonl
for f in gpurun_out/${TAG}_n8.json gpurun_out/${TAG}_n4.json gpurun_out/${TAG}_n2.json gpurun_out/${TAG}_inproc8.json; do python - "$f" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], "value %.1f ms %.3f e2e %.1f (%.3f ms) kernel_ms %.3f max %.3f ranks %s ref_eq %s pipeline %s launches %d" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["roofline"]["kernel_ms"], j["roofline"]["kernel_ms_max_over_ranks"], j["roofline"].get("kernel_ms_per_rank"), j["frame_equals_reference"], j["config"]["pipeline"], j["gpu_launches"]))
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
