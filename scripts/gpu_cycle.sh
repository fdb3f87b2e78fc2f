#!/bin/bash
# One GPU-box cycle: parity tests, bench, launch-geometry sweep, ncu launch list + full capture.
# Usage (from the repo root, on the GPU box):  bash scripts/gpu_cycle.sh <tag> [--no-tests] [--no-ncu]
TAG=${1:-run}; shift
OUT=gpurun_out
mkdir -p $OUT
TESTS=1; NCU=1
for a in "$@"; do [ "$a" == "--no-tests" ] && TESTS=0; [ "$a" == "--no-ncu" ] && NCU=0; done
if [ $TESTS == 1 ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/${TAG}_pytest_gpu.log
fi
python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -c 1500 $OUT/${TAG}_bench.json
for cfg in "256 0" "384 0" "512 0" "640 0" "640 1" "320 4"; do
  set -- $cfg
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --threads-per-block $1 --blocks-per-sm $2 2>/dev/null | tail -1 > $OUT/_sweep.json
  python - "$1" "$2" <<'PY' >> $OUT/${TAG}_sweep.log
import json, sys
try:
    d = json.load(open("gpurun_out/_sweep.json"))
    print("tpb", sys.argv[1], "bps", sys.argv[2], "games/s %.4g" % d["value"], "kernel_ms %.3f" % d["roofline"]["kernel_ms"], d["clocks"])
except Exception as e:
    print("tpb", sys.argv[1], "bps", sys.argv[2], "failed", e)
PY
done
cat $OUT/${TAG}_sweep.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload slate 2>/dev/null | tail -1 > $OUT/${TAG}_bench_slate.json; tail -c 600 $OUT/${TAG}_bench_slate.json
if [ $NCU == 1 ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sims 4000000"
  $CMD > $OUT/${TAG}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_list.log 2>&1
  $CMD > $OUT/${TAG}_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mlbsim_native_kernel -s 3 -c 1 -o $OUT/${TAG}_native $CMD > $OUT/${TAG}_ncu_full.log 2>&1
  ls -la $OUT | tail -5
fi
