"""Short driver for ncu captures: config-5 (or --workload) graph, a few LM iterations only, so that a profiled run ends quickly.
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python profiles/ncu_lm.py --iters 4
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS
from vdo_slam_b200 import capi
from vdo_slam_b200.synth import make_batch_graph

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config5")
ap.add_argument("--iters", type=int, default=4)
ap.add_argument("--skip", type=int, default=0, help="LM iterations to run before the measured ones (later iterations have longer PCG solves)")
a = ap.parse_args()
g = make_batch_graph(**WORKLOADS[a.workload])
ctx = capi.Context(0)
G = capi.BatchGraph(ctx, g)
t = time.time()
r = G.optimize(max_iterations=a.skip + a.iters, gain_threshold=0.0)
print({k: v for k, v in r.items() if k != "chi2"}, "wall", time.time() - t)
