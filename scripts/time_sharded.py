#!/usr/bin/env python
"""Per-sweep kernel time of a configuration with the particles sharded over the ranks of a torchrun launch:
  python -m torch.distributed.run --nproc-per-node 2 scripts/time_sharded.py cfg2_multiomics 12 [P=...]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import pmdi_b200  # noqa
from pmdi_b200 import capi, synth
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
name, sweeps = sys.argv[1], int(sys.argv[2])
over = {}
for a in sys.argv[3:]:
    k, v = a.split("="); over[k] = float(v) if k == "rho" else int(v)
cfg = synth.make_config(name, **over)
K = len(cfg["sets"]); hy = synth.make_hypers(K, cfg["N"], cfg["n"], cfg["seed"])
n1 = int(np.floor(cfg["rho"] * cfg["n"]))
rng = np.random.default_rng(1)
ctx = capi.Context(cfg["data"], cfg["types"], cfg["N"], cfg["P"], device=local, rank=rank, n_ranks=world)
ctx.connect()
s = hy["s"]
for it in range(sweeps):
    r = ctx.sweep_sharded(s, rng.permutation(cfg["n"]) + 1, n1, hy["Pi"], hy["phi"], seed=9, it=it,
                          logweight_init=float(it > 0), time_phases=True)
    s = r["s"]
    if True:
        print(json.dumps(dict(rank=rank, sweep=it, kernel_ms=round(r["sweep_kernel_ms"], 3), resamples=r["n_resamples"],
                              remote_rows=r["n_remote_rows"], rows=sum(r["rows_evaluated"]), engine=r["engine"],
                              rs_wait_plan_ms=round(r["phase_ms_max"][4], 3), rs_maps_pulls_ms=round(r["phase_ms_max"][5], 3),
                              rs_rebuild_ms=round(r["phase_ms_max"][7], 3), rs_total_ms=round(r["phase_ms_max"][6], 3),
                              mean=[round(v, 2) for v in r["phase_ms"][:4]], mx=[round(v, 2) for v in r["phase_ms_max"][:4]])), flush=True)
if rank == 0:
    print(json.dumps(dict(phase_ms=[round(v, 3) for v in r["phase_ms"]], phase_ms_max=[round(v, 3) for v in r["phase_ms_max"]])))
dist.barrier(); ctx.close(); dist.destroy_process_group()
