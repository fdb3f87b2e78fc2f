import sys, time, cProfile, pstats, os
sys.path.insert(0, os.getcwd())
import numpy as np
import bench_workloads as bw
from mlb_odds_engine_b200 import tables
from mlb_odds_engine_b200._lib import Context
import scipy.special, scipy.stats
ctx = Context(0); ctx.alias_rows(np.full((1,8),0.125))
for name in ("sweep","slate"):
    for label, fn in (("host", None), ("device", ctx.alias_rows)):
        for cold in (True, False):
            if cold: tables._CLIP_CACHE.clear()
            t=time.perf_counter(); tb=bw.build_tables(name, alias_fn=fn); dt=time.perf_counter()-t
            print(name, label, "cold" if cold else "warm", "%.1f ms" % (dt*1e3))
tables._CLIP_CACHE.clear()
pr=cProfile.Profile(); pr.enable(); tb=bw.build_tables("sweep", alias_fn=ctx.alias_rows); pr.disable(); pstats.Stats(pr).sort_stats('tottime').print_stats(12)
