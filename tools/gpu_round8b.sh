#!/bin/bash
# Round-2 multi-GPU pass (run from the repo root ON an 8-GPU box): bench at N = 8, 4, 2, per-phase exchange timing at N = 8
# (push and NCCL), the C-ABI multi-GPU test over all GPUs, the C++ drop-in with --gpus.   tools/gpu_round8b.sh <tag>
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4 2; do
  [ $N -le $NG ] || continue
  timeout 300 $TR --nproc-per-node $N --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/${TAG}_bench_n$N.json") if l.startswith("{")][0]); print("N=$N", round(d["value"]), "QPS", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"]), " kernel_ms", round(d["roofline"]["kernel_ms"],3), "prepass", round(d["roofline"]["prepass_ms"],3), " 3x:", round(d["fp32_3xtf32_path"]["value"]), d["config"]["parallelism"][:90])
except Exception as e: print("N=$N parse failed", e)
PY
done
for m in push nccl; do
  VSB_EXCHANGE=$m timeout 200 $TR --nproc-per-node $NG --master-port 29650 tools/exchange_timing.py 2>$OUT/${TAG}_exchange_${m}_n$NG.err | grep "^{" > $OUT/${TAG}_exchange_${m}_n$NG.json; cut -c1-330 $OUT/${TAG}_exchange_${m}_n$NG.json
done
timeout 300 python -m pytest tests/test_sharded_gpu.py -q -x -k "all_gpus" > $OUT/${TAG}_pytest_mgpu$NG.log 2>&1; echo "pytest mgpu rc=$?"; tail -2 $OUT/${TAG}_pytest_mgpu$NG.log
python - <<PY
import sys; sys.path.insert(0, ".")
import vsb200_loader, numpy as np
vsb = vsb200_loader.load()
vsb.synth.write_fvecs("/tmp/base.fvecs", vsb.synth.make("cont", 2025, 1_000_000))
vsb.synth.write_fvecs("/tmp/query.fvecs", vsb.synth.make("cont", 2026, 10_000))
PY
for N in 1 $NG; do
  for rep in 1 2; do
  timeout 300 hai-25-rag-on-edge_b200/bin/cpu_baseline /tmp/base.fvecs /tmp/query.fvecs 10 /tmp/res_$N.txt --gpus $N > $OUT/${TAG}_cpp_mgpu_n$N.log 2>&1; echo "cpu_baseline --gpus $N rc=$?"
  done
  grep -E "Throughput|GPUs|Total execution" $OUT/${TAG}_cpp_mgpu_n$N.log
done
cmp /tmp/res_1.txt /tmp/res_$NG.txt && echo "C++ multi-GPU results file identical to 1 GPU"
