"""Per-kernel CUDA-event timings of the batch path through vdo_graph_time_kernel (no profiler attached)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS
from vdo_slam_b200 import capi
from vdo_slam_b200.synth import make_batch_graph, graph_sizes

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config5")
ap.add_argument("--iters", type=int, default=12)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
g = make_batch_graph(**WORKLOADS[a.workload])
ctx = capi.Context(0)
G = capi.BatchGraph(ctx, g)
r = G.optimize(max_iterations=a.iters, gain_threshold=0.0)
out = {"workload": a.workload, "sizes": graph_sizes(g), "lm": {k: v for k, v in r.items() if k != "chi2"}, "ms": {}}
for n in ["lin_tracklets", "chi2_tracklets", "lin_vertex_obs", "lin_vertex_ter", "lin_se3_edges", "linearize", "factor_landmarks", "precond",
          "schur_landmarks", "schur_static", "schur_chains", "lin_static", "lin_chains", "schur_vertex_obs", "schur_vertex_ter", "hpp_mul", "pcg_dot", "pcg_step", "pcg_iterate8"]:
    out["ms"][n] = G.time_kernel(n, a.reps)
print(json.dumps(out))
