"""Generates the golden vectors under tests/golden/ from the CPU oracle on seeded synthetic inputs.

The reference cannot be built or imported in this image (no Eigen3 / OpenCV C++ / CSparse), it ships
no fixtures of its own, and its demo data is an external download -- so the vectors below are the ORACLE's outputs, frozen:
they pin the oracle against drift (tests/test_golden.py, CPU) and give the GPU parity tests a committed target that does not
depend on rebuilding the oracle on the GPU box.  Re-run only when the oracle is deliberately changed:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po                                   # noqa: E402
from vdo_slam_b200.synth import make_batch_graph, make_flow_problem  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # batch LM (Optimizer::FullBatchOptimization constants): 10 frames, 1 object, 120 static + 40 dynamic tracks
    g = make_batch_graph(n_frames=10, n_objects=1, n_static=120, n_dynamic=40, seed=42)
    r = po.ba_optimize(g, max_iters=12, gain_threshold=1e-4)
    np.savez_compressed(os.path.join(HERE, "ba_small.npz"), iters=r["iters"], chi2=r["chi2"], se3=r["se3"], pt=r["pt"],
                        **{"g_" + k: v for k, v in g.items() if isinstance(v, np.ndarray)})
    # per-frame flow / pose LM (Optimizer::PoseOptimizationFlow2), both arithmetic modes, object and camera priors
    out = {}
    p = make_flow_problem(n=300, seed=7)
    for mode in (0, 1):
        for quirk in (0, 1):
            f = po.flow2(p, mode=mode, quirk=quirk)
            out[f"T_m{mode}_q{quirk}"] = f["T"]; out[f"iters_m{mode}_q{quirk}"] = f["iters"]; out[f"inlier_m{mode}_q{quirk}"] = f["inlier"]
    np.savez_compressed(os.path.join(HERE, "flow2_small.npz"), **out, **{"p_" + k: np.asarray(v) for k, v in p.items()})
    print("written:", [f for f in sorted(os.listdir(HERE)) if f.endswith(".npz")])


if __name__ == "__main__":
    main()
