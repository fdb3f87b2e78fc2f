#!/bin/bash
# On the GPU box: replay-mode throughput (scripts/replay_bench.py, bit-exactness checked in the same run) for the
# product library and every variants/libmlbsim_rp*.so.  usage: bash scripts/replay_variant_bench.sh <tag> [n_games] [--no-bullpens]
TAG=${1:-replay}; N=${2:-800000}; EXTRA=${3:-}
mkdir -p gpurun_out; LOG=gpurun_out/${TAG}.log; : > $LOG
for lib in mlb_odds_engine_b200/libmlbsim_b200.so variants/libmlbsim_rp*.so; do
  [ -f "$lib" ] || continue
  MLBSIM_LIB=$PWD/$lib python scripts/replay_bench.py $N $EXTRA 2>gpurun_out/_rp.err | tail -1 > gpurun_out/_rp.json
  python - "$lib" >> $LOG <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/_rp.json"))
    print("REPLAY %-44s games/s %.4g  GB/s %.1f  frac_hbm %.3f  bit_exact %s  kernel_ms %.3f" % (sys.argv[1], d["games_per_s"], d["achieved_gbs"], d["frac_of_hbm"] or 0, d["bit_exact_vs_oracle"], d["kernel_ms"]))
except Exception as e:
    print("REPLAY %-44s FAILED %s" % (sys.argv[1], e), open("gpurun_out/_rp.err").read()[-300:])
PY
done
cat $LOG
cp gpurun_out/_rp.json gpurun_out/${TAG}_last.json
