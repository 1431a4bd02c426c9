"""Worker for the world_size-2 gloo test of the sharded batch solve (host logic on the emulated backend)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(rank, world, port, out_path):
    import torch.distributed as dist
    from vdo_slam_b200 import capi
    from vdo_slam_b200.synth import make_batch_graph
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        g = make_batch_graph(n_frames=14, n_objects=2, n_static=300, n_dynamic=120, seed=1)
        ctx = capi.Context(0, lib_path=os.path.join(ROOT, "tests", "emul", "libvdo_emul.so"))
        ctx.set_collective_emul(rank, world, dist)
        G = capi.BatchGraph(ctx, g)
        info = G.info()
        r = G.optimize()
        se3, pt = G.vertices_gathered(dist)
        if rank == 0:
            np.savez(out_path, se3=se3, pt=pt, iters=r["iterations"], chi2=r["chi2"], n_pt_local=info["n_pt"], n_pt=len(g["pt"]))
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
