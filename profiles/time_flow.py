"""Config 2 timing: one PoseOptimizationFlow2 problem of 2 000 points (device-resident re-runs and end-to-end calls), plus a
frame-like batch of objects; CPU oracle timed beside it."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vdo_slam_b200 import capi
from vdo_slam_b200.synth import make_flow_problem
from oracle import pyoracle as po

ctx = capi.Context(0)
out = {}
for name, probs, modes in [("config2_n2000", [make_flow_problem(2000, 1234)], [1]),
                           ("frame_like_cam1200_plus_4obj", [make_flow_problem(1200, 1, rot_deg=0.5)] + [make_flow_problem(n, s) for n, s in [(800, 2), (800, 3), (600, 4), (300, 5)]], [0, 1, 1, 1, 1])]:
    r = capi.pose_opt_flow2(ctx, probs, quirk=1, modes=modes)
    dev_ms = capi.pose_opt_flow2_time(ctx, len(probs), quirk=1, reps=50)
    t0 = time.perf_counter()
    for _ in range(50):
        capi.pose_opt_flow2(ctx, probs, quirk=1, modes=modes)
    e2e_ms = (time.perf_counter() - t0) / 50 * 1e3
    t0 = time.perf_counter()
    for _ in range(5):
        o = [po.flow2(p, mode=m, quirk=1) for p, m in zip(probs, modes)]
    cpu_ms = (time.perf_counter() - t0) / 5 * 1e3
    out[name] = {"lm_iters": [x["iters"] for x in r], "device_ms": dev_ms, "e2e_ms": e2e_ms, "cpu_oracle_ms": cpu_ms,
                 "us_per_lm_iter_device": dev_ms * 1e3 / max(1, max(x["iters"] for x in r))}
print(json.dumps(out))
