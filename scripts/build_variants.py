#!/usr/bin/env python
"""Cross-compile tuning variants of libmlbsim_b200.so HERE (no GPU needed) so that one GPU-box call can A/B them:

    python scripts/build_variants.py tag1="-DOPT_X=1 -DOPT_Y=2" tag2="..."

writes variants/libmlbsim_<tag>.so (git-ignored like every .so; travels to the GPU box with the snapshot).
scripts/variant_bench.sh runs bench.py once per library through the MLBSIM_LIB environment variable.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mlb_odds_engine_b200 import build  # noqa: E402

os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
for spec in sys.argv[1:]:
    tag, _, defs = spec.partition("=")
    out = os.path.join(ROOT, "variants", f"libmlbsim_{tag}.so")
    build.build_library(force=True, defs=defs.split(), out=out)
    print("built", out, defs)
