#!/bin/bash
# Round-2 GPU-box cycle: parity suite (product library, then the build with in-kernel traps), variant A/B, bench, ncu.
# usage: bash scripts/r02_cycle.sh <tag> [--no-variants] [--no-ncu]
TAG=${1:-r02}; shift
OUT=gpurun_out; mkdir -p $OUT
VAR=1; NCU=1
for a in "$@"; do [ "$a" == "--no-variants" ] && VAR=0; [ "$a" == "--no-ncu" ] && NCU=0; done
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > $OUT/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest_gpu.log
if [ -f variants/dbg/libmlbsim_debug.so ]; then
  MLBSIM_LIB=$PWD/variants/dbg/libmlbsim_debug.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/${TAG}_pytest_debug_bounds.log 2>&1; echo "debug-bounds pytest rc=$?"; tail -5 $OUT/${TAG}_pytest_debug_bounds.log
fi
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -c 2500 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
if [ $VAR == 1 ]; then timeout 1200 bash scripts/variant_bench.sh ${TAG}_variants; fi
if [ $NCU == 1 ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads --sims 4000000"
  $CMD > $OUT/${TAG}_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_list.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlbsim_native_kernel -s 3 -c 1 -o $OUT/${TAG}_native $CMD > $OUT/${TAG}_ncu_full.log 2>&1
  ls -la $OUT | tail -5
fi
# replay kernel: throughput (bit-exactness checked in the same run) and one ncu --set full capture
timeout 300 python scripts/replay_bench.py 800000 > $OUT/${TAG}_replay.json 2> $OUT/${TAG}_replay.err; tail -c 600 $OUT/${TAG}_replay.json
if [ $NCU == 1 ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlbsim_replay_flat -s 3 -c 1 -o $OUT/${TAG}_replay python scripts/replay_bench.py 400000 > $OUT/${TAG}_replay_ncu.log 2>&1
fi
