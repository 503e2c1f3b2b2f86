#!/bin/bash
# One multi-GPU gpurun call (run from the repo root ON an N-GPU box): sharded tests, bench at N = 2, 4, 8 (as available),
# per-phase exchange timing, the C++ single-process multi-GPU drop-in, config 5.   tools/gpu_round8.sh <tag>
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -q -x > $OUT/${TAG}_pytest_sharded${NG}.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest_sharded${NG}.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4 2; do
  [ $N -le $NG ] || continue
  timeout 300 $TR --nproc-per-node $N --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench_n$N.json")); print("N=$N", round(d["value"]), "QPS", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"]), " kernel_ms", round(d["roofline"]["kernel_ms"],3), " 3x:", round(d["fp32_3xtf32_path"]["value"]))
except Exception as e: print("N=$N parse failed", e)
PY
  timeout 200 $TR --nproc-per-node $N --master-port $((29600+N)) tools/exchange_timing.py > $OUT/${TAG}_exchange_n$N.json 2> $OUT/${TAG}_exchange_n$N.err; echo "exchange N=$N rc=$?"; cat $OUT/${TAG}_exchange_n$N.json
done
# C++ host program: single process, all GPUs (vs_exact_mgpu_*)
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
# config 5
timeout 600 $TR --nproc-per-node $NG --master-port 29700 tools/bench_config5.py > $OUT/${TAG}_config5_n$NG.json 2> $OUT/${TAG}_config5_n$NG.err; echo "config5 rc=$?"; cat $OUT/${TAG}_config5_n$NG.json | cut -c1-600
