"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of ``libpmdi_oracle.so`` (the CPU restatement of the reference sweep, see
``pmdi_oracle.h``).  Imported only by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpmdi_oracle.so")

GAUSSIAN, CATEGORICAL, NEGBINOM = 0, 1, 2
MODE_DENSE = 0
MODE_DEDUP = 1
MODE_LITERAL_NEWID = 2
MODE_SSTAR_COMPAT = 4

DRAW_ALLOC, DRAW_RESAMP, DRAW_SHUFFLE, DRAW_SELECT, DRAW_FEATURE = 0, 1, 2, 3, 4


def build(force: bool = False) -> str:
    """Compile the oracle (gcc only; no GPU, no reference sources involved)."""
    srcs = [os.path.join(_HERE, f) for f in ("pmdi_oracle.cpp", "pmdi_oracle.h", "philox.h")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpmdi_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _SweepArgs(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("_pad", C.c_int32),
        ("s_in", C.c_void_p), ("order_obs", C.c_void_p), ("n1", C.c_int64),
        ("Pi", C.c_void_p), ("phi", C.c_void_p), ("logweight_init", C.c_double),
        ("seed", C.c_uint64), ("iter", C.c_uint32), ("_pad2", C.c_uint32),
        ("tape_alloc", C.c_void_p), ("tape_resamp", C.c_void_p),
        ("tape_shuffle", C.c_void_p), ("tape_select", C.c_void_p),
        ("s_out", C.c_void_p), ("p_star", C.c_void_p), ("logweight", C.c_void_p),
        ("n_ops", C.c_void_p), ("n_resamples", C.c_void_p),
        ("dbg_lp", C.c_void_p), ("dbg_lw", C.c_void_p), ("dbg_alloc", C.c_void_p),
        ("dbg_anc", C.c_void_p), ("cluster_n", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.or_create.restype = C.c_void_p
        L.or_create.argtypes = [C.c_int] * 4
        L.or_destroy.argtypes = [C.c_void_p]
        L.or_set_dataset.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.or_set_flags.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.or_sweep.argtypes = [C.c_void_p, C.POINTER(_SweepArgs)]
        L.or_feature_null.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.or_feature_select.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64,
                                        C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.or_cl_new.restype = C.c_void_p
        L.or_cl_new.argtypes = [C.c_void_p, C.c_int]
        L.or_cl_free.argtypes = [C.c_void_p]
        L.or_cl_add.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
        L.or_cl_logprob.restype = C.c_double
        L.or_cl_logprob.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
        L.or_cl_logmarginal.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.or_cl_n.restype = C.c_int64
        L.or_cl_n.argtypes = [C.c_void_p]
        L.or_cl_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.or_calc_ess.restype = C.c_double
        L.or_calc_ess.argtypes = [C.c_void_p, C.c_int]
        L.or_draw_partstar.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.or_uniform_c.restype = C.c_double
        L.or_uniform_c.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                   C.c_uint32]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def uniform(seed, it, kind, step, k, index) -> float:
    return lib().or_uniform_c(seed, it, kind, step, k, index)


def calc_ess(logweight) -> float:
    lw = np.ascontiguousarray(logweight, dtype=np.float64)
    return lib().or_calc_ess(_ptr(lw), len(lw))


def draw_partstar(logweight, r, shuffle_u):
    lw = np.ascontiguousarray(logweight, dtype=np.float64)
    su = np.ascontiguousarray(shuffle_u, dtype=np.float64)
    out = np.zeros(len(lw), dtype=np.int64)
    lib().or_draw_partstar(_ptr(lw), len(lw), float(r), _ptr(su), _ptr(out))
    return out


class Cluster:
    """One cluster of dataset k (the plugin contract: add / logprob / logmarginal)."""

    def __init__(self, orc: "Oracle", k: int):
        self.o, self.k = orc, k
        self.h = lib().or_cl_new(orc.h, k)

    def __del__(self):
        if getattr(self, "h", None):
            lib().or_cl_free(self.h)
            self.h = None

    def add(self, obs_1based: int):
        lib().or_cl_add(self.o.h, self.k, self.h, obs_1based)

    def logprob(self, obs_1based: int) -> float:
        return lib().or_cl_logprob(self.o.h, self.k, self.h, obs_1based)

    def logmarginal(self):
        out = np.zeros(self.o.D[self.k])
        lib().or_cl_logmarginal(self.o.h, self.k, self.h, _ptr(out))
        return out

    @property
    def n(self) -> int:
        return lib().or_cl_n(self.h)

    def get(self, field: str):
        D = self.o.D[self.k]
        idx = {"mu": 0, "sum": 1, "lam": 2, "beta": 3, "counts": 4, "isum": 5}[field]
        size = D * self.o.Lmax[self.k] if field == "counts" else D
        out = np.zeros(size)
        rc = lib().or_cl_get(self.h, idx, _ptr(out))
        assert rc == 0
        if field == "counts":
            return out.reshape(D, self.o.Lmax[self.k]).T  # [level, feature]
        return out


class Oracle:
    """CPU oracle context: K datasets, n_obs rows, N clusters, P particles."""

    def __init__(self, data, types, N: int, P: int):
        self.K = len(data)
        self.n = int(data[0].shape[0])
        self.N, self.P = int(N), int(P)
        self.types = list(types)
        self.D = [int(d.shape[1]) for d in data]
        self.Lmax = [0] * self.K
        self.h = lib().or_create(self.K, self.n, self.N, self.P)
        for k, (d, t) in enumerate(zip(data, types)):
            if t == GAUSSIAN:
                a = np.asfortranarray(d, dtype=np.float64)
            else:
                a = np.asfortranarray(d, dtype=np.int64)
                self.Lmax[k] = int(a.max())
            rc = lib().or_set_dataset(self.h, k, t, _ptr(a), self.D[k])
            assert rc == 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().or_destroy(self.h)
            self.h = None

    def set_flags(self, k, flags):
        f = np.ascontiguousarray(flags, dtype=np.uint8)
        assert f.shape == (self.D[k],)
        lib().or_set_flags(self.h, k, _ptr(f))

    def cluster(self, k) -> Cluster:
        return Cluster(self, k)

    def feature_null(self, k):
        out = np.zeros(self.D[k])
        lib().or_feature_null(self.h, k, _ptr(out))
        return out

    def feature_select(self, k, labels, feature_null, seed=0, it=0, tape_f=None):
        lab = np.ascontiguousarray(labels, dtype=np.int64)
        fn = np.ascontiguousarray(feature_null, dtype=np.float64)
        tf = None if tape_f is None else np.ascontiguousarray(tape_f, dtype=np.float64)
        prob = np.zeros(self.D[k])
        flags = np.zeros(self.D[k], dtype=np.uint8)
        lib().or_feature_select(self.h, k, _ptr(lab), _ptr(fn), seed, it, _ptr(tf), _ptr(prob),
                                _ptr(flags))
        return prob, flags

    def sweep(self, s, order_obs, n1, Pi, phi, *, mode=MODE_DENSE, logweight_init=0.0, seed=0,
              it=0, tapes=None, debug=False):
        """One conditional-SMC sweep.  ``s`` is n x K (labels 1..N), ``order_obs`` 1-based."""
        n, K, N, P = self.n, self.K, self.N, self.P
        s_in = np.asfortranarray(s, dtype=np.int64)
        oo = np.ascontiguousarray(order_obs, dtype=np.int64)
        Pi_ = np.asfortranarray(Pi, dtype=np.float64)
        assert Pi_.shape == (N, K) and s_in.shape == (n, K)
        npairs = max(1, K * (K - 1) // 2)
        phi_ = np.ascontiguousarray(phi, dtype=np.float64)
        assert phi_.size >= (npairs if K > 1 else 0)
        steps = n - int(n1) + 1
        tapes = tapes or {}
        keep = []

        def tape(name, shape):
            t = tapes.get(name)
            if t is None:
                return None
            t = np.ascontiguousarray(t, dtype=np.float64)
            assert t.shape == shape, (name, t.shape, shape)
            keep.append(t)
            return t

        a = _SweepArgs()
        a.mode = mode
        a.s_in, a.order_obs, a.n1 = _ptr(s_in), _ptr(oo), int(n1)
        a.Pi, a.phi = _ptr(Pi_), _ptr(phi_)
        a.logweight_init = float(logweight_init)
        a.seed, a.iter = int(seed), int(it)
        a.tape_alloc = _ptr(tape("alloc", (steps, K, P)))
        a.tape_resamp = _ptr(tape("resamp", (steps,)))
        a.tape_shuffle = _ptr(tape("shuffle", (steps, P)))
        a.tape_select = _ptr(tape("select", (1,)))
        out = {
            "s": np.zeros((n, K), dtype=np.int64, order="F"),
            "p_star": np.zeros(1, dtype=np.int64),
            "logweight": np.zeros(P),
            "n_ops": np.zeros(1, dtype=np.int64),
            "n_resamples": np.zeros(1, dtype=np.int64),
            "cluster_n": np.zeros((K, P, N), dtype=np.int64),
        }
        a.s_out, a.p_star, a.logweight = _ptr(out["s"]), _ptr(out["p_star"]), _ptr(out["logweight"])
        a.n_ops, a.n_resamples = _ptr(out["n_ops"]), _ptr(out["n_resamples"])
        a.cluster_n = _ptr(out["cluster_n"])
        if debug:
            out["lp"] = np.zeros((steps, K, P, N))
            out["lw"] = np.zeros((steps, P))
            out["alloc"] = np.zeros((steps, K, P), dtype=np.int32)
            out["anc"] = np.zeros((steps, P), dtype=np.int32)
            a.dbg_lp, a.dbg_lw = _ptr(out["lp"]), _ptr(out["lw"])
            a.dbg_alloc, a.dbg_anc = _ptr(out["alloc"]), _ptr(out["anc"])
        rc = lib().or_sweep(self.h, C.byref(a))
        if rc != 0:
            raise RuntimeError(f"or_sweep failed with code {rc}")
        out["p_star"] = int(out["p_star"][0])
        out["n_ops"] = int(out["n_ops"][0])
        out["n_resamples"] = int(out["n_resamples"][0])
        return out
