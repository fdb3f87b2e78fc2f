#!/bin/bash
# compute-sanitizer stand-in (the tool is closed on the GPU pool): rebuild the library with in-kernel address and
# state-invariant traps (-DMLBSIM_DEBUG_BOUNDS) ON THE GPU BOX, run the GPU parity suite, restore the normal build.
mkdir -p gpurun_out
MLBSIM_NVCC_DEFS="-DMLBSIM_DEBUG_BOUNDS" python -m mlb_odds_engine_b200.build --force > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/debug_bounds_pytest.log 2>&1; echo "debug-bounds pytest rc=$?"; tail -3 gpurun_out/debug_bounds_pytest.log
python -m mlb_odds_engine_b200.build --force > /dev/null 2>&1
