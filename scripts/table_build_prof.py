"""Table-build timings on the GPU box: host builder (numpy + scipy, cold / warm quadrature memo, host / device alias rows)
and the device builder (params up, tables built in HBM: csrc/table_kernel.cu), plus a byte comparison of the two on the bench
workloads.  `python scripts/table_build_prof.py > gpurun_out/r02_table_build.log`"""
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
from mlb_odds_engine_b200 import sweep
from mlb_odds_engine_b200.simulator import GpuSimulator
sim = GpuSimulator(0)
base, axes = bw.sweep_grid()
for i in range(3):
    t=time.perf_counter(); params,_ = sweep.grid_params(base, "home", **axes); t1=time.perf_counter(); sim.load_params(params); t2=time.perf_counter()
    print("sweep: params %.1f ms, device build+install %.1f ms" % ((t1-t)*1e3, (t2-t1)*1e3))
for name in ("slate","single","fatigue"):
    ms = bw.matchups(name)
    for i in range(2):
        t=time.perf_counter(); p = tables.params_batch(ms); t1=time.perf_counter(); sim.load_params(p); t2=time.perf_counter()
        print(name, "params %.1f ms, device build+install %.1f ms" % ((t1-t)*1e3, (t2-t1)*1e3))
# byte comparison with host tables on the bench workloads
for name in ("single","slate","fatigue","sweep"):
    host = bw.build_tables(name)
    if name == "sweep":
        params,_ = sweep.grid_params(base, "home", **axes)
    else:
        params = tables.params_batch(bw.matchups(name))
    sim.ctx.build_tables(params); dev = sim.ctx.download_tables()
    nd = sum(int((dev[f] != host[f]).sum()) for f in ("alias","fthr","fslope"))
    print(name, "identical bytes:", dev.tobytes()==host.tobytes(), "differing words:", nd, "of", host["alias"].size)
