"""Worker for the 2-GPU NCCL test of the sharded batch solve (real CUDA backend)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    from vdo_slam_b200 import capi
    from vdo_slam_b200.synth import make_batch_graph
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        g = make_batch_graph(n_frames=30, n_objects=2, n_static=1500, n_dynamic=300, seed=1)
        ctx = capi.Context(rank)
        ctx.init_comm(rank, world, dist)
        G = capi.BatchGraph(ctx, g)
        r = G.optimize()
        se3, pt = G.vertices_gathered(dist)
        # BASELINE config 4 (the golden full solve of tests/golden/ba_config4.npz), sharded over the ranks
        g4 = make_batch_graph(n_frames=200, n_objects=5, n_static=40000, n_dynamic=10000, seed=4)
        G4 = capi.BatchGraph(ctx, g4)
        r4 = G4.optimize()
        se3_4, pt_4 = G4.vertices_gathered(dist)
        if rank == 0:
            np.savez(out_path, se3=se3, pt=pt, iters=r["iterations"], chi2=r["chi2"], c4_se3=se3_4, c4_pt=pt_4, c4_iters=r4["iterations"], c4_chi2=r4["chi2"])
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
