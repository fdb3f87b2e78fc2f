#!/bin/bash
# usage: bash scripts/replay_variants.sh <n_games> "<defs>" ...   (rebuilds on the GPU box per variant)
N=$1; shift
mkdir -p gpurun_out
for defs in "$@"; do
  MLBSIM_NVCC_DEFS="$defs" python -m mlb_odds_engine_b200.build --force 2>&1 | grep -A2 "mlbsim_replay_flat_kernel" | grep -E "registers" | tr -d "\n"; echo
  python scripts/replay_bench.py $N 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('REPLAY [$defs] n=$N ok=%s games/s %.4g GB/s %.1f ms %.3f' % (d['bit_exact_vs_oracle'], d['games_per_s'], d['achieved_gbs'], d['kernel_ms']))"
done | tee -a gpurun_out/replay_variants.log
python -m mlb_odds_engine_b200.build --force > /dev/null 2>&1
