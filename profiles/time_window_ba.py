"""Latency of one sliding-window optimisation (PartialBatchOptimization: 20 cameras, static points only) through the C ABI:
ingest, LM solve (phases with VDO_PROFILE=1), read-back.   python profiles/time_window_ba.py [n_static]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vdo_slam_b200 import capi
from vdo_slam_b200.synth import make_batch_graph, PARTIAL_BATCH

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
g = make_batch_graph(n_frames=20, n_objects=0, n_static=ns, n_dynamic=0, seed=9, consts=PARTIAL_BATCH)
ctx = capi.Context(0)
for rep in range(4):
    t0 = time.perf_counter(); G = capi.BatchGraph(ctx, g); t1 = time.perf_counter()
    r = G.optimize(max_iterations=100, gain_threshold=1e-3); t2 = time.perf_counter()
    G.vertices(); t3 = time.perf_counter(); G.close(); t4 = time.perf_counter()
    print(f"rep {rep}: edges {len(g['obs_cp'])} ingest {(t1-t0)*1e3:.2f} ms | solve {(t2-t1)*1e3:.2f} ms ({r['iterations']} LM it, {r['trials']} trials, {r['pcg_iterations']} pcg, "
          f"{r['kernel_launches']} launches, {(t2-t1)*1e3/max(r['iterations'],1):.2f} ms/it) | read-back {(t3-t2)*1e3:.2f} | free {(t4-t3)*1e3:.2f}", flush=True)
print(G.solver_info() if False else "")
