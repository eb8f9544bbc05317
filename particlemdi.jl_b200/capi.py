"""ctypes binding of ``libpmdi_cuda.so`` (include/pmdi_cuda.h) — the same entry points a Julia
``ccall`` binding uses.  There is no CPU fallback: if the library is missing it
raises, and every compute call fails without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpmdi_cuda.so")
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

GAUSSIAN, CATEGORICAL, NEGBINOM = 0, 1, 2
F64, I64 = 0, 1
SWEEP_DEBUG, SWEEP_SSTAR_COMPAT, SWEEP_TIME_PHASES = 1, 2, 4

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

EXPORTS = [
    "pmdi_last_error", "pmdi_version", "pmdi_device_count", "pmdi_ctx_create", "pmdi_ctx_destroy",
    "pmdi_ctx_set_stream", "pmdi_ctx_get_stream", "pmdi_set_dataset", "pmdi_set_feature_flags",
    "pmdi_sweep", "pmdi_sweep_upload", "pmdi_sweep_run", "pmdi_sweep_download",
    "pmdi_feature_null", "pmdi_feature_select", "pmdi_cluster_eval", "pmdi_uniform",
    "pmdi_ctx_set_ranks", "pmdi_ipc_export", "pmdi_ipc_import",
    "pmdi_psm_begin", "pmdi_psm_add", "pmdi_psm_get",
]
IPC_HANDLE_BYTES = 64


def build(force: bool = False) -> str:
    """nvcc-compile the CUDA extension in-tree for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "pmdi_cuda.h")]
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        subprocess.check_call([nvcc, *NVCC_FLAGS, "-o", LIB_PATH, os.path.join(CSRC, "pmdi_cuda.cu")])
    return LIB_PATH


class SweepArgs(C.Structure):
    _fields_ = [
        ("flags", C.c_uint32), ("iter", C.c_uint32), ("seed", C.c_uint64),
        ("s", C.c_void_p), ("order_obs", C.c_void_p), ("n1", C.c_int64),
        ("Pi", C.c_void_p), ("phi", C.c_void_p), ("logweight_init", C.c_double),
        ("tape_alloc", C.c_void_p), ("tape_resamp", C.c_void_p),
        ("tape_shuffle", C.c_void_p), ("tape_select", C.c_void_p),
    ]


class SweepOut(C.Structure):
    _fields_ = [
        ("s", C.c_void_p), ("p_star", C.c_void_p), ("logweight", C.c_void_p),
        ("n_resamples", C.c_int64), ("n_copies", C.c_int64), ("n_remote_rows", C.c_int64), ("n_evals", C.c_int64),
        ("n_evals_dense", C.c_int64), ("rows_evaluated", C.c_int64 * 8),
        ("device_ms", C.c_double), ("sweep_kernel_ms", C.c_double), ("phase_ms", C.c_double * 8),
        ("phase_ms_max", C.c_double * 8),
        ("dbg_lp", C.c_void_p), ("dbg_lw", C.c_void_p), ("dbg_alloc", C.c_void_p),
        ("dbg_anc", C.c_void_p), ("cluster_n", C.c_void_p),
        ("label_counts", C.c_void_p), ("pair_agree", C.c_void_p),
        ("rows_referenced", C.c_int64 * 8), ("engine", C.c_int32), ("rows_evaluated_ahead", C.c_int64),
        ("rows_computed", C.c_int64 * 8), ("rows_added", C.c_int64 * 8), ("contingency", C.c_void_p),
    ]


_lib = None


def lib():
    """Load the CUDA extension; raise loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (there is no CPU fallback for the sweep)")
        L = C.CDLL(LIB_PATH)
        L.pmdi_last_error.restype = C.c_char_p
        L.pmdi_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int64, C.c_int32,
                                      C.c_int32, C.c_int32]
        L.pmdi_ctx_destroy.argtypes = [C.c_void_p]
        L.pmdi_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.pmdi_ctx_get_stream.argtypes = [C.c_void_p]
        L.pmdi_ctx_get_stream.restype = C.c_void_p
        L.pmdi_set_dataset.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                       C.c_int64, C.c_int64, C.c_int64]
        L.pmdi_set_feature_flags.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.pmdi_sweep.argtypes = [C.c_void_p, C.POINTER(SweepArgs), C.POINTER(SweepOut)]
        L.pmdi_sweep_upload.argtypes = [C.c_void_p, C.POINTER(SweepArgs)]
        L.pmdi_sweep_run.argtypes = [C.c_void_p]
        L.pmdi_sweep_download.argtypes = [C.c_void_p, C.POINTER(SweepOut)]
        L.pmdi_feature_null.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.pmdi_feature_select.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pmdi_cluster_eval.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int64,
                                        C.c_void_p, C.c_void_p]
        L.pmdi_ctx_set_ranks.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.pmdi_ipc_export.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        L.pmdi_ipc_import.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pmdi_psm_begin.argtypes = [C.c_void_p]
        L.pmdi_psm_add.argtypes = [C.c_void_p, C.c_void_p]
        L.pmdi_psm_get.argtypes = [C.c_void_p, C.c_void_p]
        L.pmdi_uniform.restype = C.c_double
        L.pmdi_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                   C.c_uint32]
        _lib = L
    return _lib


class PmdiError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise PmdiError(f"[{rc}] {lib().pmdi_last_error().decode()}")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def device_count() -> int:
    return lib().pmdi_device_count()


def uniform(seed, it, kind, step, k, index) -> float:
    return lib().pmdi_uniform(seed, it, kind, step, k, index)


class Context:
    """One sweep context: K datasets bound to one GPU (``pmdi_ctx``)."""

    def __init__(self, data, types, N: int, particles: int, device: int = 0, rank: int = 0,
                 n_ranks: int = 1):
        """``particles`` is the GLOBAL count; with ``n_ranks > 1`` this rank holds particles / n_ranks
        of them (call :meth:`connect` after construction)."""
        self.rank, self.n_ranks = int(rank), int(n_ranks)
        self.K = len(data)
        self.n = int(data[0].shape[0])
        self.N, self.P = int(N), int(particles)
        self.types = list(types)
        self.D = [int(d.shape[1]) for d in data]
        h = C.c_void_p()
        _check(lib().pmdi_ctx_create(C.byref(h), self.K, self.n, self.N, self.P, device))
        self.h = h
        if self.n_ranks > 1:
            _check(lib().pmdi_ctx_set_ranks(self.h, self.rank, self.n_ranks))
        for k, (d, t) in enumerate(zip(data, types)):
            if d.shape[0] != self.n:
                raise AssertionError("datasets must have the same number of rows")  # src/pmdi.jl:52
            if t == GAUSSIAN:
                a = np.asfortranarray(d, dtype=np.float64)
                kind = F64
            else:
                a = np.asfortranarray(d, dtype=np.int64)
                kind = I64
            _check(lib().pmdi_set_dataset(self.h, k, t, kind, _ptr(a), self.n, self.D[k], self.n))

    # ---- particle sharding over the GPUs of one node ------------------------------------------
    def export_handle(self):
        """(IPC handle bytes, arena size) of this rank's shared arena."""
        buf = (C.c_ubyte * IPC_HANDLE_BYTES)()
        nbytes = C.c_int64(0)
        _check(lib().pmdi_ipc_export(self.h, buf, C.byref(nbytes)))
        return bytes(buf), int(nbytes.value)

    def import_handles(self, handles, sizes=None):
        blob = b"".join(handles)
        assert len(blob) == self.n_ranks * IPC_HANDLE_BYTES
        hb = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        sz = None if sizes is None else (C.c_int64 * self.n_ranks)(*sizes)
        _check(lib().pmdi_ipc_import(self.h, hb, sz))

    def connect(self, group=None):
        """Exchange the arena handles over ``torch.distributed`` (plumbing only: the per-observation
        exchange runs inside the sweep kernel over NVLink peer memory)."""
        import torch.distributed as dist
        self._group = group
        mine = self.export_handle()
        allh = [None] * self.n_ranks
        dist.all_gather_object(allh, mine, group=group)
        self.import_handles([h for h, _ in allh], [n for _, n in allh])
        dist.barrier(group=group)

    def sweep_sharded(self, *args, **kw):
        """``sweep`` on every rank with identical arguments: upload, barrier over the ranks (every
        rank's grid counter is reset before any kernel starts), run, download."""
        import torch.distributed as dist
        if kw.pop("sstar_compat", False):
            raise NotImplementedError("sstar_compat is not offered through the split upload/run path")
        self.upload(*args, **kw)
        if os.environ.get("PMDI_ENGINE") == "dense":  # the dense engine restarts its grid counters every sweep
            dist.barrier(group=getattr(self, "_group", None))
        self.run()
        return self.download()

    def close(self):
        if getattr(self, "h", None):
            lib().pmdi_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_flags(self, k, flags):
        f = np.ascontiguousarray(flags, dtype=np.uint8)
        assert f.shape == (self.D[k],)
        _check(lib().pmdi_set_feature_flags(self.h, k, _ptr(f)))

    # ---- sweep -------------------------------------------------------------------------------
    def _args(self, s, order_obs, n1, Pi, phi, logweight_init, seed, it, tapes, flags):
        n, K, N, P = self.n, self.K, self.N, self.P
        keep = []
        s_in = np.asfortranarray(s, dtype=np.int64)
        oo = np.ascontiguousarray(order_obs, dtype=np.int64)
        Pi_ = np.asfortranarray(Pi, dtype=np.float64)
        assert Pi_.shape == (N, K) and s_in.shape == (n, K) and oo.shape == (n,)
        phi_ = None if phi is None else np.ascontiguousarray(phi, dtype=np.float64)
        steps = n - int(n1) + 1
        tapes = tapes or {}

        def tape(name, shape):
            t = tapes.get(name)
            if t is None:
                return None
            t = np.ascontiguousarray(t, dtype=np.float64)
            assert t.shape == shape, (name, t.shape, shape)
            keep.append(t)
            return t

        a = SweepArgs()
        a.flags, a.iter, a.seed = int(flags), int(it), int(seed)
        a.s, a.order_obs, a.n1 = _ptr(s_in), _ptr(oo), int(n1)
        a.Pi, a.phi = _ptr(Pi_), _ptr(phi_)
        a.logweight_init = float(logweight_init)
        a.tape_alloc = _ptr(tape("alloc", (steps, K, P)))
        a.tape_resamp = _ptr(tape("resamp", (steps,)))
        a.tape_shuffle = _ptr(tape("shuffle", (steps, P)))
        a.tape_select = _ptr(tape("select", (1,)))
        keep += [s_in, oo, Pi_, phi_]
        return a, keep, steps

    def _out(self, steps, debug):
        n, K, N, P = self.n, self.K, self.N, self.P
        res = {
            "s": np.zeros((n, K), dtype=np.int64, order="F"),
            "p_star": np.zeros(1, dtype=np.int64),
            "logweight": np.zeros(P),
            "cluster_n": np.zeros((K, P, N), dtype=np.int64),
            "label_counts": np.zeros((N, K), dtype=np.int64, order="F"),
            "pair_agree": np.zeros(max(1, K * (K - 1) // 2), dtype=np.int64),
            # [pair][lb][la]: entry (la, lb) of an N x N column-major table per dataset pair
            "contingency": np.zeros((max(1, K * (K - 1) // 2), N, N), dtype=np.int64),
        }
        o = SweepOut()
        o.label_counts, o.pair_agree = _ptr(res["label_counts"]), _ptr(res["pair_agree"])
        o.contingency = _ptr(res["contingency"])
        o.s, o.p_star, o.logweight = _ptr(res["s"]), _ptr(res["p_star"]), _ptr(res["logweight"])
        o.cluster_n = _ptr(res["cluster_n"])
        if debug:
            res["lp"] = np.zeros((steps, K, P, N))
            res["lw"] = np.zeros((steps, P))
            res["alloc"] = np.zeros((steps, K, P), dtype=np.int32)
            res["anc"] = np.zeros((steps, P), dtype=np.int32)
            o.dbg_lp, o.dbg_lw = _ptr(res["lp"]), _ptr(res["lw"])
            o.dbg_alloc, o.dbg_anc = _ptr(res["alloc"]), _ptr(res["anc"])
        return o, res

    @staticmethod
    def _finish(o, res):
        res["p_star"] = int(res["p_star"][0])
        res["n_resamples"] = int(o.n_resamples)
        res["n_copies"] = int(o.n_copies)
        res["n_remote_rows"] = int(o.n_remote_rows)
        res["n_evals"] = int(o.n_evals)
        res["n_evals_dense"] = int(o.n_evals_dense)
        res["rows_evaluated"] = [int(v) for v in o.rows_evaluated]
        res["rows_referenced"] = [int(v) for v in o.rows_referenced]
        res["engine"] = {0: "dense", 1: "pool", 2: "spec"}[int(o.engine)]
        res["rows_computed"] = [int(v) for v in o.rows_computed]
        res["rows_added"] = [int(v) for v in o.rows_added]
        res["rows_evaluated_ahead"] = int(o.rows_evaluated_ahead)
        res["device_ms"] = float(o.device_ms)
        res["sweep_kernel_ms"] = float(o.sweep_kernel_ms)
        res["phase_ms"] = [float(v) for v in o.phase_ms]
        res["phase_ms_max"] = [float(v) for v in o.phase_ms_max]
        return res

    def sweep(self, s, order_obs, n1, Pi, phi, *, logweight_init=0.0, seed=0, it=0, tapes=None,
              debug=False, sstar_compat=False, time_phases=False):
        """One conditional-SMC sweep through ``pmdi_sweep`` (host buffers in, host buffers out)."""
        flags = (SWEEP_DEBUG if debug else 0) | (SWEEP_SSTAR_COMPAT if sstar_compat else 0) | \
                (SWEEP_TIME_PHASES if time_phases else 0)
        a, keep, steps = self._args(s, order_obs, n1, Pi, phi, logweight_init, seed, it, tapes, flags)
        o, res = self._out(steps, debug)
        _check(lib().pmdi_sweep(self.h, C.byref(a), C.byref(o)))
        return self._finish(o, res)

    # split form, for timing the device-resident part on its own
    def upload(self, s, order_obs, n1, Pi, phi, *, logweight_init=0.0, seed=0, it=0, tapes=None,
               debug=False, time_phases=False):
        flags = (SWEEP_DEBUG if debug else 0) | (SWEEP_TIME_PHASES if time_phases else 0)
        a, keep, steps = self._args(s, order_obs, n1, Pi, phi, logweight_init, seed, it, tapes, flags)
        _check(lib().pmdi_sweep_upload(self.h, C.byref(a)))
        self._steps, self._debug = steps, debug

    def run(self):
        _check(lib().pmdi_sweep_run(self.h))

    def download(self):
        o, res = self._out(self._steps, self._debug)
        _check(lib().pmdi_sweep_download(self.h, C.byref(o)))
        return self._finish(o, res)

    # ---- posterior similarity matrices (consensus_map.jl:31-65) --------------------------------
    def psm(self, alloc):
        """alloc: (rows, n_obs, K) retained allocations -> (K, n_obs, n_obs) on the GPU."""
        alloc = np.asarray(alloc)
        _check(lib().pmdi_psm_begin(self.h))
        for r in range(alloc.shape[0]):
            s = np.asfortranarray(alloc[r], dtype=np.int64)
            _check(lib().pmdi_psm_add(self.h, _ptr(s)))
        out = np.zeros((self.K, self.n, self.n))
        _check(lib().pmdi_psm_get(self.h, _ptr(out)))
        return out

    # ---- feature selection / plugin contract -------------------------------------------------
    def feature_null(self, k):
        out = np.zeros(self.D[k])
        _check(lib().pmdi_feature_null(self.h, k, _ptr(out)))
        return out

    def feature_select(self, k, labels, feature_null, seed=0, it=0, tape_f=None):
        lab = np.ascontiguousarray(labels, dtype=np.int64)
        fn = np.ascontiguousarray(feature_null, dtype=np.float64)
        tf = None if tape_f is None else np.ascontiguousarray(tape_f, dtype=np.float64)
        prob = np.zeros(self.D[k])
        flags = np.zeros(self.D[k], dtype=np.uint8)
        _check(lib().pmdi_feature_select(self.h, k, _ptr(lab), _ptr(fn), seed, it, _ptr(tf),
                                         _ptr(prob), _ptr(flags)))
        return prob, flags

    def cluster_eval(self, k, rows_1based, obs_1based=None, logmarginal=False):
        """calc_logprob(obs, cluster built from rows) and/or calc_logmarginal(cluster)."""
        rows = np.ascontiguousarray(rows_1based, dtype=np.int64)
        lp = np.zeros(1) if obs_1based is not None else None
        lm = np.zeros(self.D[k]) if logmarginal else None
        _check(lib().pmdi_cluster_eval(self.h, k, _ptr(rows), len(rows),
                                       int(obs_1based) if obs_1based is not None else 1,
                                       _ptr(lp), _ptr(lm)))
        return (None if lp is None else float(lp[0])), lm
