"""Host-side plumbing of the particle sharding, on CPU with gloo and world_size 2: every rank ends
up with every rank's arena handle in rank order, and a layout mismatch is refused.  (The exchange
itself - NVLink peer stores inside the sweep kernel - needs GPUs: tests/test_gpu_sharded.py.)"""
import os
import socket

import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakeCtx:
    """Stands in for the device side of capi.Context: records what connect() hands to import."""

    def __init__(self, rank, world, arena_bytes):
        self.rank, self.n_ranks, self._bytes = rank, world, arena_bytes
        self.imported = None

    def export_handle(self):
        return bytes([self.rank]) * 64, self._bytes

    def import_handles(self, handles, sizes=None):
        if len(set(sizes)) != 1:
            raise RuntimeError("different arena layout")
        self.imported = (list(handles), list(sizes))


def _worker(rank, world, port, outdir, mismatch):
    import torch.distributed as dist
    import pmdi_b200  # noqa: F401
    from pmdi_b200 import capi
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        ctx = _FakeCtx(rank, world, 1000 + (rank if mismatch else 0))
        try:
            capi.Context.connect(ctx)          # the real exchange code, on the fake device side
            ok = ctx.imported == ([bytes([r]) * 64 for r in range(world)], [1000] * world)
            res = "ok" if ok else f"bad {ctx.imported}"
        except RuntimeError as e:
            res = f"refused: {e}"
        open(os.path.join(outdir, f"r{rank}.txt"), "w").write(res)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mismatch", [False, True])
def test_connect_exchanges_handles_in_rank_order(tmp_path, mismatch):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), mismatch), nprocs=world, join=True)
    got = [open(os.path.join(tmp_path, f"r{r}.txt")).read() for r in range(world)]
    if mismatch:
        assert all(g.startswith("refused") for g in got), got
    else:
        assert got == ["ok"] * world, got


def test_particles_must_divide_over_ranks():
    """pmdi_ctx_set_ranks refuses a particle count that does not split evenly (no device needed)."""
    import ctypes as C
    import pmdi_b200.capi as capi
    assert capi.lib().pmdi_ctx_set_ranks(None, 0, 2) != 0
    assert b"NULL" in capi.lib().pmdi_last_error()
