"""One rank of a particle-sharded sweep (spawned by tests/test_gpu_sharded.py and usable by hand):
builds the context for its GPU, connects the ranks, runs the sweep, saves what it got."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_rank(rank, world, port, case_kw, seed, outdir, use_tapes, sweeps):
    import torch
    import torch.distributed as dist
    import pmdi_b200  # noqa: F401
    from pmdi_b200 import capi
    from helpers import problem, tapes_for
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        pr = problem(**case_kw, seed=seed)
        tapes = tapes_for(pr) if use_tapes else None
        ctx = capi.Context(pr["data"], pr["types"], pr["N"], pr["P"], device=rank, rank=rank, n_ranks=world)
        ctx.connect()
        s = pr["s"]
        rng = np.random.default_rng(77)
        out = {}
        for it in range(sweeps):
            order = pr["order"] if it == 0 else rng.permutation(pr["n"]) + 1
            r = ctx.sweep_sharded(s, order, pr["n1"], pr["Pi"], pr["phi"], seed=11, it=3 + it,
                                  logweight_init=0.0 if it == 0 else 1.0, tapes=tapes if it == 0 else None,
                                  debug=True)
            s = r["s"]
            for key in ("s", "alloc", "anc", "lw", "lp", "logweight", "cluster_n"):
                out[f"{key}_{it}"] = r[key]
            out[f"p_star_{it}"] = r["p_star"]
            out[f"n_resamples_{it}"] = r["n_resamples"]
            out[f"n_copies_{it}"] = r["n_copies"]
            out[f"n_remote_rows_{it}"] = r["n_remote_rows"]
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), **out)
        dist.barrier()
        ctx.close()
    finally:
        dist.destroy_process_group()


def run_pmdi_rank(rank, world, port, outdir, iters):
    """pmdi() with the particles sharded over `world` GPUs: every rank runs the host loop, rank 0 writes the CSV."""
    import torch
    import torch.distributed as dist
    import pmdi_b200  # noqa: F401
    from pmdi_b200 import pmdi as host
    from test_gpu_pmdi import _separable
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        data, _ = _separable(70, seed=5)
        host.pmdi(data, [0, 2, 1], 6, 24, 0.25, iters, os.path.join(outdir, f"sharded_rank{rank}.csv"), seed=3,
                  device=rank, distributed=True)
        dist.barrier()
    finally:
        dist.destroy_process_group()
