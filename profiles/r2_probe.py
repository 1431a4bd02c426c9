"""Round-2 probe (run on the GPU box): per-kernel CUDA-event timings on config 5, LM solve timing, and the effect of the PCG
tolerance on parity with the oracle's frozen config-5 solve (tests/golden/ba_config5.npz)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vdo_slam_b200 import capi
from vdo_slam_b200.synth import make_batch_graph, iso_inv, iso_mul, iso_t, iso_R

cfg = dict(n_frames=1000, n_objects=50, n_static=800000, n_dynamic=200000, seed=5, obj_span=(100, 400))
if len(sys.argv) > 1 and sys.argv[1] == "config4":
    cfg = dict(n_frames=200, n_objects=5, n_static=40000, n_dynamic=10000, seed=4); name = "ba_config4.npz"
else:
    name = "ba_config5.npz"
g = make_batch_graph(**cfg)
d = np.load(os.path.join(ROOT, "tests", "golden", name))
ctx = capi.Context(0)
G = capi.BatchGraph(ctx, g)
out = {}
# variants "tol" or "tol:loose:switch" (forcing schedule: loose tolerance while the previous LM iteration gained more than `switch`)
for var in os.environ.get("TOLS", "1e-8,1e-6,1e-5,1e-4").split(","):
    parts = [float(x) for x in var.split(":")]
    tol = parts[0]
    kw = dict(pcg_rel_tol=tol)
    if len(parts) == 3:
        kw.update(pcg_loose_tol=parts[1], pcg_switch_gain=parts[2])
    G.reset()
    G.optimize(**kw)
    G.reset()
    t0 = time.perf_counter()
    r = G.optimize(**kw)
    dt = time.perf_counter() - t0
    tol = var
    se3, pt = G.vertices()
    dd = iso_mul(iso_inv(se3), d["se3"])
    n = min(len(r["chi2"]), len(d["chi2"]))
    out[f"tol{tol}"] = dict(iters=r["iterations"], golden_iters=int(d["iters"]), pcg=r["pcg_iterations"], ms=dt * 1e3, ms_total=r["ms_total"],
                             max_pose=float(max(np.abs(iso_t(dd)).max(), np.abs(iso_R(dd) - np.eye(3)).max())),
                             max_point=float(np.abs(pt[d["pt_idx"]] - d["pt"]).max()), max_rel_chi2=float(np.abs(r["chi2"][:n] / d["chi2"][:n] - 1).max()),
                             lm_it_per_s=r["iterations"] / dt)
    print(f"tol {tol}:", json.dumps(out[f"tol{tol}"]), flush=True)
G.reset(); G.optimize()
kt = {}
for k in ["schur_static", "schur_chains", "schur_vertex_obs", "hpp_mul", "pcg_step", "pcg_iterate8", "lin_static", "lin_chains", "lin_vertex_obs", "precond", "band_form", "precond_tiles", "pcr_factor", "pcg_step_a", "chi2_tracklets", "factor_landmarks"]:
    try:
        kt[k] = G.time_kernel(k, 20)
    except Exception as e:
        kt[k] = repr(e)
print("kernel ms:", json.dumps(kt))
out["kernels_ms"] = kt
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_probe_%s.json" % name.split(".")[0]), "w"), indent=1)
