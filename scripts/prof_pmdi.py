import cProfile, pstats, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pmdi_b200 import pmdi as host
name = sys.argv[1]; iters = int(sys.argv[2])
cfg = bench.make_workload(name)
with tempfile.TemporaryDirectory() as td:
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    st = host.pmdi(cfg["data"], cfg["types"], cfg["N"], cfg["P"], cfg["rho"], iters, os.path.join(td, "o.csv"), seed=1)
    pr.disable()
    print("total s", time.perf_counter() - t0, st)
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
