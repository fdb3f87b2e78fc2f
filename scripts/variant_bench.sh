#!/bin/bash
# On the GPU box: bench every variants/libmlbsim_*.so (built beforehand by scripts/build_variants.py) plus the product
# library, two short runs each, best kept.  usage: bash scripts/variant_bench.sh <log tag> [bench args...]
TAG=${1:-variants}; shift
mkdir -p gpurun_out
LOG=gpurun_out/${TAG}.log
: > $LOG
for lib in mlb_odds_engine_b200/libmlbsim_b200.so variants/libmlbsim_*.so; do
  [ -f "$lib" ] || continue
  for rep in 1 2; do
    MLBSIM_LIB=$PWD/$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-workloads "$@" 2>gpurun_out/_v.err | tail -1 > gpurun_out/_v.json
    python - "$lib" >> $LOG <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/_v.json"))
    print("VARIANT %-50s games/s %.4g  kernel_ms %.4f  clocks %s" % (sys.argv[1], d["value"], d["roofline"]["kernel_ms"], d["clocks"].get("sm_mhz")))
except Exception as e:
    print("VARIANT %-50s FAILED %s" % (sys.argv[1], e), open("gpurun_out/_v.err").read()[-400:])
PY
  done
done
cat $LOG
