#!/bin/bash
# Rebuild the library with each set of -D tuning switches ON THE GPU BOX and bench it (short runs).
# usage: bash scripts/variant_sweep.sh "<defs A>" "<defs B>" ...
mkdir -p gpurun_out
for spec in "$@"; do
  # "<nvcc defs>|<bench args>"
  defs="${spec%%|*}"; bargs=""; [[ "$spec" == *"|"* ]] && bargs="${spec#*|}"
  MLBSIM_NVCC_DEFS="$defs" python -m mlb_odds_engine_b200.build --force 2>&1 | grep -A2 "mlbsim_native_kernel" | grep -E "registers|spill" | tr -d "\n"; echo
  best=0
  for rep in 1 2; do
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline $bargs 2>/dev/null | tail -1 > gpurun_out/_v.json
    python - "$spec" <<'PY'
import json, sys
d = json.load(open("gpurun_out/_v.json"))
print("VARIANT [%s] games/s %.4g kernel_ms %.3f" % (sys.argv[1], d["value"], d["roofline"]["kernel_ms"]))
PY
  done
done | tee gpurun_out/variant_sweep.log
python -m mlb_odds_engine_b200.build --force > /dev/null 2>&1
