"""world_size-2 run of the landmark-sharded batch solve on CPU (gloo): the real host driver + the serial kernel emulation,
with the backend's all-reduce routed through torch.distributed.  Checks that sharding does not change the result."""
import os
import socket
import subprocess
import sys

import numpy as np

from oracle import pyoracle as po
from vdo_slam_b200.synth import make_batch_graph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_two_rank_sharded_solve_matches_single_rank_oracle(tmp_path):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emul"), "libvdo_emul.so"], stdout=subprocess.DEVNULL)
    port, out = _free_port(), str(tmp_path / "r0.npz")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_dist_worker.py"), str(r), "2", str(port), out]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    d = np.load(out)
    g = make_batch_graph(n_frames=14, n_objects=2, n_static=300, n_dynamic=120, seed=1)
    ro = po.ba_optimize(g)
    assert int(d["iters"]) == ro["iters"]
    assert 0 < int(d["n_pt_local"]) < int(d["n_pt"])          # rank 0 really held only a shard
    assert np.abs(d["se3"] - ro["se3"]).max() < 1e-5 and np.abs(d["pt"] - ro["pt"]).max() < 1e-5
