#!/bin/bash
# round 2, GPU call D (2 GPUs): multi-device parity, torchrun bench with direct peer stores vs NCCL gather, in-process form
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_parity.py::test_multi_device_context tests/test_gpu_fullsize.py::test_tiles_stored_into_another_process_frame -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/r2d_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2d_bench_n2_peer.json 2> gpurun_out/r2d_bench_n2_peer.err; echo "peer rc=$?"; cut -c1-900 gpurun_out/r2d_bench_n2_peer.json; tail -n 4 gpurun_out/r2d_bench_n2_peer.err
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --gather nccl > gpurun_out/r2d_bench_n2_nccl.json 2> gpurun_out/r2d_bench_n2_nccl.err; echo "nccl rc=$?"; cut -c1-400 gpurun_out/r2d_bench_n2_nccl.json; tail -n 4 gpurun_out/r2d_bench_n2_nccl.err
timeout 600 python bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_inproc2.json 2> gpurun_out/r2d_bench_inproc2.err; echo "inproc rc=$?"; cut -c1-400 gpurun_out/r2d_bench_inproc2.json; tail -n 4 gpurun_out/r2d_bench_inproc2.err
MTB_NO_PEER_STORE=1 timeout 600 python bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_inproc2_copy.json 2> gpurun_out/r2d_bench_inproc2_copy.err; echo "inproc copy rc=$?"; cut -c1-400 gpurun_out/r2d_bench_inproc2_copy.json
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2d_ref_n2.json 2> gpurun_out/r2d_ref_n2.err; echo "ref rc=$?"; cut -c1-700 gpurun_out/r2d_ref_n2.json
for f in gpurun_out/r2d_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], "value %.1f ms %.3f e2e %.1f (%.3f ms) kernel_ms %.3f max %.3f sha %s ref_eq %s pipeline %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["roofline"]["kernel_ms"], j["roofline"]["kernel_ms_max_over_ranks"], j["frame_sha"][:12], j["frame_equals_reference"], j["config"]["pipeline"]))
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
