#!/bin/bash
# Round-2 evidence pass (run from the repo root ON the GPU box):  tools/gpu_round2.sh <tag> [tests] [bench] [launches] [full]
# ncu passes only start after the same command exited 0 without ncu.  Everything lands under gpurun_out/<tag>_*.
set -u
TAG=${1:-r2}
shift
WHAT=${*:-tests bench launches full}
OUT=gpurun_out
mkdir -p $OUT
has() { [[ " $WHAT " == *" $1 "* ]]; }
if has tests; then
  timeout 1200 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1
  echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
  tail -4 $OUT/${TAG}_pytest_gpu.log
fi
if has bench; then
  timeout 600 python bench.py > $OUT/${TAG}_bench_auto.json 2> $OUT/${TAG}_bench_auto.err; echo "bench auto rc=$?"
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "bench ref rc=$?"
  cat $OUT/${TAG}_bench_auto.json
fi
NCU_L="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
if has launches; then
  timeout 600 $NCU_L -c 400 --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_bench.log 2>&1
  echo "ncu launches bench rc=$?"
fi
NCU_F="ncu --set full --clock-control none --import-source on"
if has full; then
  # launch order of exact_tc_kernel in bench.py: (sample, filter) per search; 3 warm-up iterations x 2 searches = 12 launches
  timeout 600 $NCU_F -k regex:exact_tc_kernel -s 13 -c 1 -f -o $OUT/${TAG}_full_exact_tc_f16_filter python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_full_f16.log 2>&1
  echo "ncu full f16 filter rc=$?"
  timeout 600 $NCU_F -k regex:exact_tc_kernel -s 12 -c 1 -f -o $OUT/${TAG}_full_exact_tc_f16_sample python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_full_f16s.log 2>&1
  echo "ncu full f16 sample rc=$?"
fi
ls -la $OUT | grep ${TAG}_ | head -40
