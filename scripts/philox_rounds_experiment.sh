#!/bin/bash
# DESIGN.md §10 experiment: what the generator costs.  For R in 7 8: the CUDA library and the oracle twin built with
# NATIVE_PHILOX_ROUNDS=R (variants/libmlbsim_rR.so, oracle/_build/libmlbsim_oracle_rR.so), the bit-exact and the
# statistical GPU parity tests run on that pair, then the bench.  The product (R = 10) is not touched.
mkdir -p gpurun_out; LOG=gpurun_out/r02_philox_rounds.log; : > $LOG
for R in 7 8; do
  [ -f variants/libmlbsim_r$R.so ] || python scripts/build_variants.py r$R="-DNATIVE_PHILOX_ROUNDS=$R" > /dev/null
  [ -f oracle/_build/libmlbsim_oracle_r$R.so ] || make -s -C oracle -B OUT=_build/libmlbsim_oracle_r$R.so EXTRA=-DNATIVE_PHILOX_ROUNDS=$R
  export MLBSIM_LIB=$PWD/variants/libmlbsim_r$R.so MLBSIM_ORACLE_LIB=$PWD/oracle/_build/libmlbsim_oracle_r$R.so
  echo "== NATIVE_PHILOX_ROUNDS=$R" >> $LOG
  timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "native or statistics or calibration_outcome or fresh_reference or slate_multi" 2>&1 | tail -3 >> $LOG
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-workloads 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('games/s %.4g  kernel_ms %.4f' % (d['value'], d['roofline']['kernel_ms'])); print('parity', json.dumps(d['parity']['joint_runs_pmf_chi2']), json.dumps({k:v['z'] for k,v in d['parity']['probabilities'].items()}))" >> $LOG
done
unset MLBSIM_LIB MLBSIM_ORACLE_LIB
cat $LOG
