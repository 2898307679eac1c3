"""Host-side mirror of the reference's solver interface over the C ABI.

The reference entry point is (SolveQuadraticProgram.jl:14-17)

    SolveQuadraticProgram!(vX, mP, vQ, mA, vL, vU, LinSysSolInit, LinSysSol!;
        numIterations = 5000, ϵAbs = 1e-6, ϵRel = 1e-6, ρ = 1, σ = 1e-6, α = 1.6, δ = 1e-6,
        adptΡ = false, fctrΡ = 5, numItrConv = 25, numItrPolish = 10, ϵMinres = 1e-6, numItrMinres = 500)

Here it is :func:`SolveQuadraticProgram_` (``!`` -> trailing underscore) with the same positional
arguments, the same keyword names (Python accepts the Greek identifiers; ASCII aliases are accepted
too) and the same return value (the ``ConvergenceFlag``), mutating ``vX`` in place.  The plugin pair
is the singleton pair ``(B200Init, B200Sol)``: passing it selects the GPU path, exactly as a Julia
call site would swap ``FacLdlInit, FacLdl!`` for ``B200Init, B200Sol!`` (see julia/QPB200.jl).
``SolveQuadraticProgram(P, q, A, l, u; ...)`` is the convenience form BASELINE.json names.

Julia is not available in this environment, so this module -- not julia/QPB200.jl -- is what the
test-suite drives; both are thin and call the same C symbols in the same order.
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import Info, QPB200Error, Settings  # noqa: F401


class ConvergenceFlag(enum.IntEnum):
    """``@enum ConvergenceFlag convNumItr = 1 convAdmm convPrimDual`` (SolveQuadraticProgram.jl:12)."""
    convNumItr = 1
    convAdmm = 2
    convPrimDual = 3


class _Plugin:
    def __init__(self, name):
        self._name = name

    def __repr__(self):
        return self._name


#: the (Init, Sol!) pair that selects the B200 path (LinearSystemSolvers.jl:16,28 are the CPU pairs)
B200Init = _Plugin("B200Init")
B200Sol = _Plugin("B200Sol!")

_ALIASES = {
    "ϵAbs": "epsAbs", "ϵRel": "epsRel", "ρ": "rho", "σ": "sigma", "α": "alpha", "δ": "delta",
    "adptΡ": "adptRho", "fctrΡ": "fctrRho", "ϵMinres": "epsMinres", "ϵPcg": "epsPcg",
    # Python NFKC-normalises identifiers: the reference's lunate epsilon (U+03F5) arrives as U+03B5
    "εAbs": "epsAbs", "εRel": "epsRel", "εMinres": "epsMinres", "εPcg": "epsPcg",
}
_DEFAULTS = dict(numIterations=5000, epsAbs=1e-6, epsRel=1e-6, rho=1.0, sigma=1e-6, alpha=1.6, delta=1e-6,
                 adptRho=False, fctrRho=5.0, numItrConv=25, numItrPolish=10, epsMinres=1e-6, numItrMinres=500,
                 # plugin kwargs (LinearSystemSolvers.jl:125) and the new ones
                 epsPcg=1e-6, numItrPcg=1000, relPcg=-1.0, linSolver="pcg", precond="jacobi", device=-1,
                 spmvLoader="auto",
                 # Ruiz equilibration iterations (0 = off = the reference's behaviour)
                 numItrScaling=0,
                 # arrangement of the (P)CG recurrence: "standard" (IterativeSolvers' CGIterable / PCGIterable statement
                 # order), "one_reduction" (three grid barriers per CG iteration), "auto" (by problem size)
                 cgRecurrence="auto",
                 # polish after the ADMM loop (SolveQuadraticProgram.m:289-325; the Julia driver ignores its polish
                 # keywords, so False is the reference's behaviour): uses numItrPolish, delta, epsMinres, numItrMinres
                 polish=False)


def make_settings(**kw) -> Settings:
    """Keyword arguments of ``SolveQuadraticProgram!`` -> ``qpb200_settings``."""
    opts = dict(_DEFAULTS)
    for k, v in kw.items():
        k = _ALIASES.get(k, k)
        if k not in opts:
            raise TypeError(f"SolveQuadraticProgram: unknown keyword argument {k!r}")
        opts[k] = v
    s = _lib.default_settings()
    s.max_iter = int(opts["numIterations"])
    s.eps_abs = float(opts["epsAbs"]); s.eps_rel = float(opts["epsRel"])
    s.rho = float(opts["rho"]); s.sigma = float(opts["sigma"]); s.alpha = float(opts["alpha"])
    s.delta = float(opts["delta"])
    s.adaptive_rho = int(bool(opts["adptRho"]))
    s.rho_factor = float(opts["fctrRho"])
    s.check_every = int(opts["numItrConv"])
    s.polish_iter = int(opts["numItrPolish"]); s.minres_eps = float(opts["epsMinres"])
    s.minres_iter = int(opts["numItrMinres"])
    s.pcg_eps = float(opts["epsPcg"]); s.pcg_max_iter = int(opts["numItrPcg"]); s.pcg_rel_eps = float(opts["relPcg"])
    s.lin_solver = {"pcg": _lib.LINSOLVE_PCG, "cholesky": _lib.LINSOLVE_CHOLESKY}[str(opts["linSolver"]).lstrip(":")]
    s.precond = {"none": _lib.PRECOND_NONE, "jacobi": _lib.PRECOND_JACOBI}[str(opts["precond"]).lstrip(":")]
    s.device = int(opts["device"])
    s.spmv_loader = {"auto": 0, "ldg": 1, "tma": 2, "tma_pipe": 3}[str(opts["spmvLoader"])]
    s.reserved_i[2] = int(opts["numItrScaling"])          # QPB200_RSV_SCALING_ITERS
    s.reserved_i[4] = {"auto": 0, "standard": 1, "one_reduction": 2}[str(opts["cgRecurrence"])]   # QPB200_RSV_CG_RECURRENCE
    s.reserved_i[5] = int(bool(opts["polish"]))           # QPB200_RSV_POLISH
    return s


def equality_rho_scale(vL, vU, factor=1e3):
    """OSQP's rho vector as a scale on the scalar rho: ``factor`` on equality rows (``l == u``), 1 elsewhere
    (Stellato et al. 2020, section 5.2: RHO_EQ_OVER_RHO_INEQ = 1e3).  Input for ``rhoScale``."""
    vL = np.asarray(vL, dtype=np.float64)
    vU = np.asarray(vU, dtype=np.float64)
    return np.where(vL == vU, float(factor), 1.0)


def _pop_rho_scale(kw, vL, vU):
    """``rhoScale`` (m positive factors) or ``rhoEqScale`` (factor for the equality rows) -> array or None."""
    rs = kw.pop("rhoScale", None)
    eq = kw.pop("rhoEqScale", None)
    if rs is not None and eq is not None:
        raise TypeError("pass rhoScale or rhoEqScale, not both")
    if eq is not None:
        rs = equality_rho_scale(vL, vU, eq)
    if rs is None:
        return None
    rs = np.ascontiguousarray(rs, dtype=np.float64)
    if rs.shape != np.shape(vL):
        raise ValueError("rhoScale must have one entry per constraint")
    return rs


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _p64(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _csc_arrays(mat):
    """scipy matrix -> (colptr, rowval, nzval) int64/int64/float64, 0-based -- SparseMatrixCSC layout."""
    mat = sp.csc_matrix(mat)
    if not mat.has_sorted_indices:
        mat = mat.sorted_indices()
    return (np.ascontiguousarray(mat.indptr, dtype=np.int64), np.ascontiguousarray(mat.indices, dtype=np.int64),
            np.ascontiguousarray(mat.data, dtype=np.float64))


class QPB200Solver:
    """A ``qpb200_handle``: the problem uploaded once, device-resident state, repeated solves."""

    def __init__(self, mP, vQ, mA, vL, vU, **kw):
        lib = _lib.load()
        self.n = int(mP.shape[0])
        self.m = int(mA.shape[0])
        if mP.shape != (self.n, self.n) or mA.shape[1] != self.n:
            raise ValueError("dimension mismatch between P and A")
        vQ = np.ascontiguousarray(vQ, dtype=np.float64); vL = np.ascontiguousarray(vL, dtype=np.float64)
        vU = np.ascontiguousarray(vU, dtype=np.float64)
        if vQ.shape != (self.n,) or vL.shape != (self.m,) or vU.shape != (self.m,):
            raise ValueError("dimension mismatch in q, l or u")
        Pp, Pi, Pv = _csc_arrays(mP)
        Ap, Ai, Av = _csc_arrays(mA)
        rho_scale = _pop_rho_scale(kw, vL, vU)
        self._kw = {_ALIASES.get(k, k): v for k, v in kw.items()}
        self.settings = make_settings(**kw)
        self._h = C.c_void_p()
        _lib.check(lib.qpb200_create(C.byref(self._h), self.n, self.m, _p64(Pp), _p64(Pi), _pd(Pv), _p64(Ap), _p64(Ai),
                                     _pd(Av), _pd(vQ), _pd(vL), _pd(vU), C.byref(self.settings), 0))
        self.info = None
        if rho_scale is not None:
            self.set_rho_scale(rho_scale)

    def set_rho_scale(self, rho_scale):
        """Per-constraint step size ``rho_i = rho * rho_scale[i]`` (``None`` = the reference's scalar rho)."""
        if rho_scale is None:
            _lib.check(_lib.load().qpb200_set_rho_scale(self._h, None))
            return
        rs = np.ascontiguousarray(rho_scale, dtype=np.float64)
        if rs.shape != (self.m,):
            raise ValueError("rho_scale must have one entry per constraint")
        _lib.check(_lib.load().qpb200_set_rho_scale(self._h, _pd(rs)))

    def solve(self, vX, want_zy: bool = False):
        """Solve from start point ``vX`` (mutated in place).  Returns the ConvergenceFlag."""
        if vX.dtype != np.float64 or not vX.flags.c_contiguous or vX.shape != (self.n,):
            raise ValueError("vX must be a contiguous float64 vector of length n")
        info = Info()
        z = np.empty(self.m) if want_zy else None
        y = np.empty(self.m) if want_zy else None
        _lib.check(_lib.load().qpb200_solve(self._h, _pd(vX), _pd(z) if want_zy else None, _pd(y) if want_zy else None,
                                            C.byref(info)))
        self.info = info.as_dict()
        if want_zy:
            self.info["z"] = z
            self.info["y"] = y
        return ConvergenceFlag(info.conv_flag)

    def update_vectors(self, vQ=None, vL=None, vU=None):
        arr = [None if v is None else np.ascontiguousarray(v, dtype=np.float64) for v in (vQ, vL, vU)]
        _lib.check(_lib.load().qpb200_update_vectors(self._h, *[None if a is None else _pd(a) for a in arr]))

    #: keyword arguments that are fixed when the handle is created
    _IMMUTABLE = ("numItrScaling", "device", "linSolver")

    def update_settings(self, **kw):
        """Change keyword arguments of the handle; the ones not named keep the values given at creation."""
        kw = {_ALIASES.get(k, k): v for k, v in kw.items()}
        for k in self._IMMUTABLE:
            if k in kw and kw[k] != self._kw.get(k, _DEFAULTS[k]):
                raise ValueError(f"{k} cannot change after the handle is created")
        merged = dict(self._kw)
        merged.update(kw)
        settings = make_settings(**merged)
        _lib.check(_lib.load().qpb200_update_settings(self._h, C.byref(settings)))
        self._kw, self.settings = merged, settings

    def apply(self, which: int, x):
        """Operators of the path: 0: P x, 1: A x, 2: A' x, 3: (P + sigma I + rho A'A) x."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.m if which == 1 else self.n)
        _lib.check(_lib.load().qpb200_apply(self._h, which, _pd(x), _pd(y)))
        return y

    def time_apply(self, which: int, reps: int = 20, flush_l2: bool = True) -> float:
        ms = C.c_double(0.0)
        _lib.check(_lib.load().qpb200_time_apply(self._h, which, reps, int(flush_l2), C.byref(ms)))
        return ms.value

    def apply_bytes(self, which: int) -> int:
        return int(_lib.load().qpb200_apply_bytes(self._h, which))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.load().qpb200_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def SolveQuadraticProgram_(vX, mP, vQ, mA, vL, vU, LinSysSolInit=B200Init, LinSysSol=B200Sol, **kw):
    """``SolveQuadraticProgram!`` on the B200 (SolveQuadraticProgram.jl:14-76).  Mutates ``vX``
    (start point in, solution out) and returns the ``ConvergenceFlag`` -- nothing else, like the
    reference.  Only the ``(B200Init, B200Sol)`` plugin pair exists here: there is no CPU path."""
    if LinSysSolInit is not B200Init or LinSysSol is not B200Sol:
        raise TypeError("quadraticprogramsolver_b200 only provides the (B200Init, B200Sol) plugin pair; "
                        "the CPU plugins live in the reference")
    with QPB200Solver(mP, vQ, mA, vL, vU, **kw) as s:
        return s.solve(vX)


def SolveQuadraticProgram(P, q, A, l, u, x0=None, **kw):
    """Convenience form named by BASELINE.json: returns ``(x, flag, info)``."""
    x = np.zeros(P.shape[0]) if x0 is None else np.array(x0, dtype=np.float64)
    with QPB200Solver(P, q, A, l, u, **kw) as s:
        flag = s.solve(x, want_zy=True)
        return x, flag, s.info


class QPB200Batch:
    """A ``qpb200_batch``: ``batch`` independent small dense QPs on one GPU (configs[2]).

    ``P[batch, n, n]`` and ``A[batch, n, m]`` hold each problem's block column-major (``A[b, j, i] =
    A_b[i, j]``; see ``problems.config_cfg3_batch``); ``q[batch, n]``, ``l, u[batch, m]``."""

    def __init__(self, P, q, A_cm, l, u, **kw):
        """``P[n, n]`` and ``A_cm[n, m]`` (ONE pair for the whole batch, MPC-style: only q, l, u differ) select the
        shared-matrix engine (``qpb200_batch_create_shared``: one factor, tensor-pipe GEMMs over 16 problems)."""
        lib = _lib.load()
        P = np.ascontiguousarray(P, dtype=np.float64); A_cm = np.ascontiguousarray(A_cm, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64); l = np.ascontiguousarray(l, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        self.shared = P.ndim == 2
        self.batch, self.n = int(q.shape[0]), int(q.shape[1])
        self.m = int(A_cm.shape[-1])
        lead = () if self.shared else (self.batch,)
        if P.shape != lead + (self.n, self.n) or A_cm.shape != lead + (self.n, self.m):
            raise ValueError("P must be [batch, n, n] and A [batch, n, m] (column-major m x n blocks), or one [n, n] / [n, m] pair")
        if q.shape != (self.batch, self.n) or l.shape != (self.batch, self.m) or u.shape != (self.batch, self.m):
            raise ValueError("dimension mismatch in q, l or u")
        self.settings = _batch_settings(kw)
        self._h = C.c_void_p()
        create = lib.qpb200_batch_create_shared if self.shared else lib.qpb200_batch_create
        _lib.check(create(C.byref(self._h), self.batch, self.n, self.m, _pd(P), _pd(A_cm), _pd(q), _pd(l), _pd(u),
                          C.byref(self.settings)))
        self.info = None

    def solve(self, X=None):
        """Returns ``(X, flags, iters)``; ``X`` (start points, default zeros) is mutated in place."""
        if X is None:
            X = np.zeros((self.batch, self.n))
        if X.dtype != np.float64 or not X.flags.c_contiguous or X.shape != (self.batch, self.n):
            raise ValueError("X must be a contiguous float64 array [batch, n]")
        flags = np.zeros(self.batch, dtype=np.int32)
        iters = np.zeros(self.batch, dtype=np.int64)
        info = Info()
        _lib.check(_lib.load().qpb200_batch_solve(self._h, _pd(X), flags.ctypes.data_as(C.POINTER(C.c_int32)), _p64(iters),
                                                  C.byref(info)))
        self.info = info.as_dict()
        return X, flags, iters

    def update_vectors(self, q=None, l=None, u=None):
        """New q [batch, n] / l, u [batch, m] for the matrices already on the device (MPC-style re-solve)."""
        arr = []
        for v, shape in ((q, (self.batch, self.n)), (l, (self.batch, self.m)), (u, (self.batch, self.m))):
            if v is None:
                arr.append(None)
                continue
            v = np.ascontiguousarray(v, dtype=np.float64)
            if v.shape != shape:
                raise ValueError(f"expected shape {shape}, got {v.shape}")
            arr.append(v)
        _lib.check(_lib.load().qpb200_batch_update_vectors(self._h, *[None if a is None else _pd(a) for a in arr]))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.load().qpb200_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _batch_settings(kw):
    kw = dict(kw)
    kw.setdefault("linSolver", "cholesky")
    unblocked = bool(kw.pop("unblockedCholesky", False))
    # A/B switch of the n = 64, m <= 96 kernel: "smem" = products out of shared memory, "regs" = A in registers
    variant = {"auto": 0, "smem": 1, "regs": 2, "regs_ak": 3}[kw.pop("denseVariant", "auto")]
    chunk = int(kw.pop("batchChunk", 0))
    s = make_settings(**kw)
    s.reserved_i[0] = 1 if unblocked else 0
    s.reserved_i[3] = variant                # QPB200_RSV_DENSE_VARIANT
    s.reserved_i[6] = chunk                  # QPB200_RSV_BATCH_CHUNK
    return s


def SolveQuadraticProgramBatch(P, q, A_cm, l, u, X0=None, pipelined=True, **kw):
    """Batched form: every problem is solved as ``SolveQuadraticProgram!`` with a direct (exact-solve)
    plugin would.  Returns ``(X, flags, iters, info)``.  ``pipelined`` (default): one ``qpb200_batch_solve_once``
    call -- chunked upload overlapped with the solve; ``False``: create + solve + destroy (everything resident)."""
    if not pipelined or np.ndim(P) == 2:           # (a shared (P, A) pair uploads next to nothing: nothing to pipeline)
        with QPB200Batch(P, q, A_cm, l, u, **kw) as b:
            X = None if X0 is None else np.array(X0, dtype=np.float64)
            X, flags, iters = b.solve(X)
            return X, flags, iters, b.info
    P = np.ascontiguousarray(P, dtype=np.float64); A_cm = np.ascontiguousarray(A_cm, dtype=np.float64)
    q = np.ascontiguousarray(q, dtype=np.float64); l = np.ascontiguousarray(l, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    batch, n = int(P.shape[0]), int(P.shape[1])
    m = int(A_cm.shape[2])
    if P.shape != (batch, n, n) or A_cm.shape != (batch, n, m):
        raise ValueError("P must be [batch, n, n] and A [batch, n, m] (column-major m x n blocks)")
    if q.shape != (batch, n) or l.shape != (batch, m) or u.shape != (batch, m):
        raise ValueError("dimension mismatch in q, l or u")
    X = np.zeros((batch, n)) if X0 is None else np.array(X0, dtype=np.float64)
    flags = np.zeros(batch, dtype=np.int32)
    iters = np.zeros(batch, dtype=np.int64)
    info = Info()
    settings = _batch_settings(kw)
    _lib.check(_lib.load().qpb200_batch_solve_once(batch, n, m, _pd(P), _pd(A_cm), _pd(q), _pd(l), _pd(u), C.byref(settings),
                                                   _pd(X), flags.ctypes.data_as(C.POINTER(C.c_int32)), _p64(iters), C.byref(info)))
    return X, flags, iters, info.as_dict()


class QPB200DistSolver:
    """One large sparse QP row-partitioned over the ranks of a ``torch.distributed`` group (one rank per
    GPU).  Every rank constructs it with the full problem (or pre-sliced parts via ``presliced``) and calls
    :meth:`solve` collectively.  The n-vector all-reduces run inside the persistent kernel over NVLink peer
    memory (``distMode="auto"``/``"peer"``) or through NCCL (``"nccl"``)."""
    _comm_key = None

    def __init__(self, mP, vQ, mA, vL, vU, group=None, presliced=None, arrays=None, **kw):
        """``presliced``: a slice made by ``partition.slice_problem`` (-> ``qpb200_dist_create``); default: the whole QP
        goes to ``qpb200_dist_create_full`` and the library partitions it.  ``arrays = ((Pp, Pi, Pv), (Ap, Ai, Av))``:
        the CSC arrays already in Julia's Int64 layout (what a ``ccall`` passes), instead of scipy matrices."""
        import torch
        import torch.distributed as dist
        lib = _lib.load()
        self.rank = dist.get_rank(group)
        self.nranks = dist.get_world_size(group)
        self.n = int(np.shape(vQ)[0])
        # rank 0 creates the NCCL id; the group (any backend) carries its 128 bytes
        idbuf = np.zeros(128, dtype=np.uint8)
        key = (id(group), self.rank, self.nranks)
        if QPB200DistSolver._comm_key != key:
            # first handle of this process for this group: make an NCCL communicator (~1 s); later handles pass
            # an all-zero id, which libqpb200 reads as "reuse the cached communicator"
            if self.rank == 0:
                _lib.check(lib.qpb200_dist_unique_id(idbuf.ctypes.data_as(C.c_void_p)))
            box = [idbuf.tobytes()]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            idbuf = np.frombuffer(box[0], dtype=np.uint8).copy()
        vQ = np.ascontiguousarray(vQ, dtype=np.float64)
        kw.setdefault("device", torch.cuda.current_device() if torch.cuda.is_available() else -1)
        # collective: "auto" = in-kernel all-reduce over NVLink peer memory when available, else NCCL;
        # "nccl" = host-driven segments + ncclAllReduce; "peer" = peer path required
        dist_mode = {"auto": 0, "nccl": 1, "peer": 2}[kw.pop("distMode", "auto")]
        self.settings = make_settings(**kw)
        self.settings.reserved_i[1] = dist_mode
        self._h = C.c_void_p()
        if presliced is not None:
            P_r, A_r, l_r, u_r, self.rows, self.cols = presliced
            self.m_local = int(A_r.shape[0])
            Pp, Pi, Pv = _csc_arrays(P_r)
            Ap, Ai, Av = _csc_arrays(A_r)
            l_r = np.ascontiguousarray(l_r, dtype=np.float64); u_r = np.ascontiguousarray(u_r, dtype=np.float64)
            _lib.check(lib.qpb200_dist_create(C.byref(self._h), self.rank, self.nranks, idbuf.ctypes.data_as(C.c_void_p),
                                              self.n, self.m_local, _p64(Pp), _p64(Pi), _pd(Pv), _p64(Ap), _p64(Ai), _pd(Av),
                                              _pd(vQ), _pd(l_r), _pd(u_r), C.byref(self.settings), 0))
        else:
            (Pp, Pi, Pv), (Ap, Ai, Av) = arrays if arrays is not None else (_csc_arrays(mP), _csc_arrays(mA))
            vL = np.ascontiguousarray(vL, dtype=np.float64); vU = np.ascontiguousarray(vU, dtype=np.float64)
            m = int(vL.shape[0])
            _lib.check(lib.qpb200_dist_create_full(C.byref(self._h), self.rank, self.nranks, idbuf.ctypes.data_as(C.c_void_p),
                                                   self.n, m, _p64(Pp), _p64(Pi), _pd(Pv), _p64(Ap), _p64(Ai), _pd(Av),
                                                   _pd(vQ), _pd(vL), _pd(vU), C.byref(self.settings), 0))
            r0, r1 = C.c_int64(), C.c_int64()
            _lib.check(lib.qpb200_dist_rows(self._h, C.byref(r0), C.byref(r1)))
            self.rows = (int(r0.value), int(r1.value))
            self.cols = None
            self.m_local = self.rows[1] - self.rows[0]
        QPB200DistSolver._comm_key = key
        self.info = None

    def solve(self, vX, want_zy: bool = False):
        """Collective.  ``vX`` (replicated start point) is overwritten with the solution on every rank."""
        info = Info()
        z = np.empty(self.m_local) if want_zy else None
        y = np.empty(self.m_local) if want_zy else None
        _lib.check(_lib.load().qpb200_dist_solve(self._h, _pd(vX), _pd(z) if want_zy else None,
                                                 _pd(y) if want_zy else None, C.byref(info)))
        self.info = info.as_dict()
        if want_zy:
            self.info["z_local"] = z
            self.info["y_local"] = y
        return ConvergenceFlag(info.conv_flag)

    def apply_bytes(self, which: int) -> int:
        return int(_lib.load().qpb200_apply_bytes(self._h, which))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.load().qpb200_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def csc_arrays_int64(mat):
    """``(colptr, rowval, nzval)`` of a scipy matrix as Int64/Int64/Float64, 0-based -- the three arrays a
    Julia ``SparseMatrixCSC{Float64,Int64}`` holds (scipy keeps int32 indices, Julia does not)."""
    return _csc_arrays(mat)


def solve_csc_arrays(n, m, P_arrays, q, A_arrays, l, u, vX, want_zy=False, **kw):
    """The exact call sequence of julia/QPB200.jl on raw CSC arrays: qpb200_create -> qpb200_solve ->
    qpb200_destroy, nothing else.  ``vX`` is mutated; returns ``(flag, info)``."""
    lib = _lib.load()
    Pp, Pi, Pv = P_arrays
    Ap, Ai, Av = A_arrays
    settings = make_settings(**kw)
    h = C.c_void_p()
    _lib.check(lib.qpb200_create(C.byref(h), n, m, _p64(Pp), _p64(Pi), _pd(Pv), _p64(Ap), _p64(Ai), _pd(Av), _pd(q), _pd(l),
                                 _pd(u), C.byref(settings), 0))
    try:
        info = Info()
        z = np.empty(m) if want_zy else None
        y = np.empty(m) if want_zy else None
        _lib.check(lib.qpb200_solve(h, _pd(vX), _pd(z) if want_zy else None, _pd(y) if want_zy else None, C.byref(info)))
    finally:
        lib.qpb200_destroy(h)
    d = info.as_dict()
    if want_zy:
        d["z"], d["y"] = z, y
    return ConvergenceFlag(info.conv_flag), d
