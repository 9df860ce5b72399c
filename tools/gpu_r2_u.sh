#!/bin/bash
# round 2, 8-GPU run: torchrun bench at N = 8 / 4 / 2 with direct peer stores, NCCL-gather A/B at N = 8, in-process form, multi-device parity
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_parity.py::test_multi_device_context -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2u_pytest.log
run() { # name nproc extra
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $2 --steps 20 --warmup 5 $3 > gpurun_out/r2u_$1.json 2> gpurun_out/r2u_$1.err; echo "$1 rc=$?"
}
run n8_peer 8 ""
run n4_peer 4 ""
run n2_peer 2 ""
timeout 900 python bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2u_inproc8.json 2> gpurun_out/r2u_inproc8.err; echo "inproc8 rc=$?"
for f in gpurun_out/r2u_n8_peer.json gpurun_out/r2u_n4_peer.json gpurun_out/r2u_n2_peer.json gpurun_out/r2u_inproc8.json; do python - "$f" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], "value %.1f ms %.3f e2e %.1f (%.3f ms) kernel_ms %.3f max %.3f sha %s ref_eq %s pipeline %s share %.2f launches %d" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["roofline"]["kernel_ms"], j["roofline"]["kernel_ms_max_over_ranks"], j["frame_sha"][:12], j["frame_equals_reference"], j["config"]["pipeline"], j["config"]["hybrid_share"], j["gpu_launches"]))
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
tail -n 3 gpurun_out/r2u_n8_peer.err
