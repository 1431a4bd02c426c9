"""Host-side cost of one end-to-end batch solve: ingest (vdo_graph_add_* + finalize: ordering, tiling, H2D), read-back, free.
  VDO_PROFILE=1 python profiles/time_ingest.py [--workload config5]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS
from vdo_slam_b200 import capi
from vdo_slam_b200.synth import make_batch_graph

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config5")
ap.add_argument("--lib", default=None)
a = ap.parse_args()
g = make_batch_graph(**WORKLOADS[a.workload])
ctx = capi.Context(0, lib_path=a.lib) if a.lib else capi.Context(0)
for rep in range(4):
    t = time.perf_counter(); G = capi.BatchGraph(ctx, g); t1 = time.perf_counter() - t
    t = time.perf_counter(); r = G.optimize(max_iterations=int(os.environ.get("ITERS", "2")), gain_threshold=0.0) if int(os.environ.get("ITERS", "2")) > 0 else None; t2 = time.perf_counter() - t
    t = time.perf_counter(); G.vertices(); t3 = time.perf_counter() - t
    t = time.perf_counter(); G.close(); t4 = time.perf_counter() - t
    print(f"rep {rep}: ingest {t1*1e3:.1f} ms | 2 LM iterations {t2*1e3:.1f} ms | read-back {t3*1e3:.1f} ms | free {t4*1e3:.1f} ms", flush=True)
