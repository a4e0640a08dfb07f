"""Diagnostic: where the HOST time of one end-to-end step goes (cProfile over bench.e2e_fused on the tiny C1 workload, whose
device work is negligible: its ms/step is the per-step host overhead of the public API)."""
import cProfile, os, pstats, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench
cx = bench.Ctx()
wl = sys.argv[1] if len(sys.argv) > 1 else "c1"
fz = bench.Fused(cx, wl, bench.WORKLOADS[wl][0], use_graph=False)
r = bench.e2e_fused(cx, fz, 20)
print("e2e ms/step", r["ms_per_step"])
pr = cProfile.Profile()
pr.enable()
r = bench.e2e_fused(cx, fz, 20)
pr.disable()
print("e2e ms/step under cProfile", r["ms_per_step"])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
