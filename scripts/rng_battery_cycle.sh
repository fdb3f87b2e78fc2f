#!/bin/bash
# DESIGN.md §10 generator battery, both halves, on one GPU box.  Stream level: scripts/rng_battery.py (rounds 3..10, two
# seeds).  Application level: scripts/rng_equivalence.py — SIMS games of the bench matchup and of the fatigue matchup per
# library build (variants/libmlbsim_rR.so = -DNATIVE_PHILOX_ROUNDS=R; the product is R = 10), compared with the product.
SIMS=${1:-20000000000}; SCALE=${2:-6}
OUT=gpurun_out; mkdir -p $OUT
for R in 4 5 6 7; do
  [ -f variants/libmlbsim_r$R.so ] || python scripts/build_variants.py r$R="-DNATIVE_PHILOX_ROUNDS=$R" > /dev/null
done
for s in 20250404 777; do
  timeout 600 python scripts/rng_battery.py --scale $SCALE --seed $s --out $OUT/r02_rng_battery_$s.json > $OUT/r02_rng_battery_$s.log 2>&1
  cat $OUT/r02_rng_battery_$s.log
done
python scripts/rng_equivalence.py run --sims $SIMS --seed 1 --out $OUT/eq_r10_s1.npz
python scripts/rng_equivalence.py run --sims $SIMS --seed 2 --out $OUT/eq_r10_s2.npz
for R in 4 5 6 7; do
  python scripts/rng_equivalence.py run --lib variants/libmlbsim_r$R.so --sims $SIMS --seed 1 --out $OUT/eq_r${R}_s1.npz
done
python scripts/rng_equivalence.py run --lib variants/libmlbsim_r7.so --sims $SIMS --seed 2 --out $OUT/eq_r7_s2.npz
python scripts/rng_equivalence.py compare $OUT/eq_r10_s1.npz $OUT/eq_r10_s2.npz $OUT/eq_r7_s1.npz $OUT/eq_r7_s2.npz $OUT/eq_r6_s1.npz $OUT/eq_r5_s1.npz $OUT/eq_r4_s1.npz > $OUT/r02_rng_equivalence.log 2>&1
cat $OUT/r02_rng_equivalence.log
