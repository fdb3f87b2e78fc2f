#!/bin/bash
# On an N-GPU box: the driver's scaling run (bench.py at 1, 2, 4, ... GPUs, both arms left to the driver) plus the multi-GPU tests.
# usage: bash scripts/scale_cycle.sh <tag> [max gpus]
TAG=${1:-scale}; MAX=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > $OUT/${TAG}_pytest_multi.log 2>&1; echo "multi pytest rc=$?"; tail -3 $OUT/${TAG}_pytest_multi.log
for N in 1 2 4 8; do
  [ $N -le $MAX ] || continue
  if [ $N == 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err
  fi
  echo "N=$N rc=$?"
  python - $OUT/${TAG}_bench_${N}gpu.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(" value %.4g  e2e %.4g  ms/step %.3f  " % (d["value"], d["e2e"]["value"], d["ms_per_step"]),
      {k: (round(v["games_per_s"] / 1e9, 2), v["digest"]["sha256"][:8], v["digest"]["matches_oracle_twin"]) for k, v in d["workloads"].items()})
PY
done
