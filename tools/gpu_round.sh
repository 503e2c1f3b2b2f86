#!/bin/bash
# One gpurun call's worth of validation + evidence (run from the repo root ON the GPU box):
#   tools/gpu_round.sh <tag> [tests] [bench] [paths] [launches] [full]
# Writes everything under gpurun_out/<tag>_*.  ncu passes only start after the same command exited 0 without ncu.
set -u
TAG=${1:-r1}
shift
WHAT=${*:-tests bench paths launches full}
OUT=gpurun_out
mkdir -p $OUT
has() { [[ " $WHAT " == *" $1 "* ]]; }

if has tests; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
  echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
  tail -3 $OUT/${TAG}_pytest_gpu.log
fi
if has bench; then
  timeout 600 python bench.py > $OUT/${TAG}_bench_auto.json 2> $OUT/${TAG}_bench_auto.err; echo "bench auto rc=$?"
  timeout 600 python bench.py --precision 3xtf32 --no-cpu > $OUT/${TAG}_bench_3xtf32.json 2> $OUT/${TAG}_bench_3xtf32.err; echo "bench 3x rc=$?"
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "bench ref rc=$?"
  cat $OUT/${TAG}_bench_auto.json
fi
if has paths; then
  timeout 900 python tools/bench_paths.py batch1 ivf int8 > $OUT/${TAG}_paths.jsonl 2> $OUT/${TAG}_paths.err; echo "paths rc=$?"
  VSB_IVF_LM=0 timeout 600 python tools/bench_paths.py ivf > $OUT/${TAG}_paths_ivf_querymajor.jsonl 2>> $OUT/${TAG}_paths.err; echo "paths (K6) rc=$?"
  cat $OUT/${TAG}_paths.jsonl
fi
NCU_L="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
if has launches; then
  timeout 600 $NCU_L -c 400 --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_bench.log 2>&1
  echo "ncu launches bench rc=$?"
  timeout 900 $NCU_L -c 3000 --log-file $OUT/${TAG}_launches_paths.csv python tools/bench_paths.py ivf int8 > $OUT/${TAG}_ncu_paths.log 2>&1
  echo "ncu launches paths rc=$?"
fi
NCU_F="ncu --set full --clock-control none --import-source on"
if has full; then
  # one launch of each dominant kernel (skip the warm-up launches with -s)
  timeout 600 $NCU_F -k regex:exact_tc_kernel -s 3 -c 1 -f -o $OUT/${TAG}_full_exact_tc_f16 python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_full_f16.log 2>&1
  echo "ncu full f16 rc=$?"
  timeout 600 $NCU_F -k regex:exact_tc_kernel -s 3 -c 1 -f -o $OUT/${TAG}_full_exact_tc_3xtf32 python bench.py --steps 2 --warmup 3 --no-cpu --precision 3xtf32 > $OUT/${TAG}_ncu_full_3x.log 2>&1
  echo "ncu full 3x rc=$?"
  timeout 600 $NCU_F -k regex:exact_stream_kernel -s 4 -c 1 -f -o $OUT/${TAG}_full_exact_stream python tools/bench_paths.py batch1 > $OUT/${TAG}_ncu_full_stream.log 2>&1
  echo "ncu full stream rc=$?"
  VSB_IVF_LM=0 timeout 900 $NCU_F -k regex:ivf_scan_kernel -s 3 -c 1 -f -o $OUT/${TAG}_full_ivf_scan_np8 python tools/bench_paths.py ivf > $OUT/${TAG}_ncu_full_ivf8.log 2>&1
  echo "ncu full ivf K6 np8 rc=$?"
  VSB_IVF_LM=0 timeout 900 $NCU_F -k regex:ivf_scan_kernel -s 13 -c 1 -f -o $OUT/${TAG}_full_ivf_scan_np32 python tools/bench_paths.py ivf > $OUT/${TAG}_ncu_full_ivf32.log 2>&1
  echo "ncu full ivf K6 np32 rc=$?"
  timeout 900 $NCU_F -k regex:ivf_lm_kernel -s 13 -c 1 -f -o $OUT/${TAG}_full_ivf_lm_np32 python tools/bench_paths.py ivf > $OUT/${TAG}_ncu_full_ivflm32.log 2>&1
  echo "ncu full ivf K8 np32 rc=$?"
  timeout 900 $NCU_F -k regex:int8_tc_kernel -s 3 -c 1 -f -o $OUT/${TAG}_full_int8_b32 python tools/bench_paths.py int8 > $OUT/${TAG}_ncu_full_int8_32.log 2>&1
  echo "ncu full int8 b32 rc=$?"
  timeout 900 $NCU_F -k regex:int8_tc_kernel -s 23 -c 1 -f -o $OUT/${TAG}_full_int8_b1024 python tools/bench_paths.py int8 > $OUT/${TAG}_ncu_full_int8_1024.log 2>&1
  echo "ncu full int8 b1024 rc=$?"
fi
ls -la $OUT | grep ${TAG}_ | head -40
