"""ctypes binding of ``libb200amg.so`` — the C-ABI in ``include/b200amg.h``.

This is the only route from the host mirror to the solve phase.  If the CUDA library is missing,
or no GPU is visible, the constructors raise: there is deliberately no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "_lib", "libb200amg.so")
_lib = None

KIND = {"none": 0, "gs": 1, "jacobi": 2, "sor": 3}
SWEEP = {"forward": 1, "backward": 2, "symmetric": 3}
SYMMETRY = {"hermitian": 0, "none": 1}
MEM_HOST, MEM_DEVICE = 0, 1
OP_A, OP_P, OP_R = 0, 1, 2

# every symbol include/b200amg.h declares (tests check the library exports all of them)
SYMBOLS = [
    "b200amg_last_error", "b200amg_version", "b200amg_device_count", "b200amg_create", "b200amg_add_level",
    "b200amg_set_coarse", "b200amg_set_coarse_callback", "b200amg_set_partition", "b200amg_finalize", "b200amg_destroy", "b200amg_solve",
    "b200amg_cycle", "b200amg_precond", "b200amg_smooth", "b200amg_apply", "b200amg_residual",
    "b200amg_coarse_solve", "b200amg_norm", "b200amg_pcg", "b200amg_smoother_create", "b200amg_smoother_apply",
    "b200amg_smoother_destroy", "b200amg_num_levels", "b200amg_level_info", "b200amg_launch_count",
    "b200amg_time_kernel", "b200amg_profile_cycle", "b200amg_device_vectors", "b200amg_set_option",
    "b200amg_residual_timings", "b200amg_get_stream", "b200amg_debug_gs_timeline", "b200amg_nccl_unique_id",
    "b200amg_partition_info", "b200amg_partition_plan", "b200amg_partition_plan_child", "b200amg_spgemm_begin", "b200amg_spgemm_fetch", "b200amg_spgemm_release",
    "b200amg_block_plan_check", "b200amg_solve_block", "b200amg_comm_stats", "b200amg_storage_info",
]


class CscDesc(C.Structure):
    """``b200amg_csc_t``"""

    _fields_ = [("m", C.c_int64), ("n", C.c_int64), ("colptr", C.c_void_p), ("rowval", C.c_void_p),
                ("nzval", C.c_void_p), ("index_bits", C.c_int32), ("index_base", C.c_int32),
                ("adjoint", C.c_int32), ("reserved", C.c_int32)]


class SmootherDesc(C.Structure):
    """``b200amg_smoother_t``"""

    _fields_ = [("kind", C.c_int32), ("sweep", C.c_int32), ("iter", C.c_int32), ("reserved", C.c_int32),
                ("omega", C.c_double)]


class B200AmgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200amg error {code}: {msg}")
        self.code = code


def build(force: bool = False) -> None:
    """Compile the CUDA library for sm_100a (nvcc cross-compiles without a GPU)."""
    srcdir = os.path.join(_HERE, "csrc", "device")
    srcs = [os.path.join(srcdir, f) for f in os.listdir(srcdir)] + [os.path.join(_HERE, "..", "include", "b200amg.h")]
    stale = (not os.path.exists(_LIBPATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIBPATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "device"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIBPATH):
            raise B200AmgError(-9, f"{_LIBPATH} is missing: build it with __graft_entry__.build() "
                                   "(the solve phase has no CPU fallback)")
        L = C.CDLL(_LIBPATH)
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        pcsc, psm = C.POINTER(CscDesc), C.POINTER(SmootherDesc)
        L.b200amg_last_error.restype = C.c_char_p
        L.b200amg_version.restype = i32
        L.b200amg_device_count.restype = i32
        sigs = {
            "b200amg_create": [C.POINTER(vp), i32],
            "b200amg_add_level": [vp, pcsc, pcsc, pcsc, psm, psm, i32],
            "b200amg_set_coarse": [vp, pcsc, i64, vp],
            "b200amg_set_coarse_callback": [vp, pcsc, i64, COARSE_FN, vp],
            "b200amg_set_partition": [vp, i32, i32, vp, i64],
            "b200amg_finalize": [vp],
            "b200amg_destroy": [vp],
            "b200amg_solve": [vp, vp, vp, i32, i32, dbl, dbl, i32, vp, i32, C.POINTER(i32), C.POINTER(i32), i32],
            "b200amg_solve_block": [vp, vp, vp, i64, i64, i32, i32, dbl, dbl, i32, vp, i32, C.POINTER(i32), C.POINTER(i32), i32],
            "b200amg_cycle": [vp, vp, vp, i32, i32],
            "b200amg_precond": [vp, vp, vp, i32, i32, i32],
            "b200amg_smooth": [vp, i32, i32, vp, vp, i32],
            "b200amg_apply": [vp, i32, i32, vp, vp, i32],
            "b200amg_residual": [vp, i32, vp, vp, vp, i32],
            "b200amg_coarse_solve": [vp, vp, vp, i32],
            "b200amg_norm": [vp, i64, vp, C.POINTER(dbl), i32],
            "b200amg_spgemm_begin": [i32, i64, i64, i64, vp, vp, vp, vp, vp, vp, C.POINTER(i64)],
            "b200amg_spgemm_fetch": [vp, vp, vp],
            "b200amg_spgemm_release": [],
            "b200amg_pcg": [vp, vp, vp, i32, i32, dbl, dbl, vp, i32, C.POINTER(i32), C.POINTER(i32), i32],
            "b200amg_smoother_create": [C.POINTER(vp), i32, pcsc, psm, i32],
            "b200amg_smoother_apply": [vp, vp, vp, i32],
            "b200amg_smoother_destroy": [vp],
            "b200amg_num_levels": [vp],
            "b200amg_level_info": [vp, i32, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)],
            "b200amg_time_kernel": [vp, i32, i32, i32, i32, i32, C.POINTER(dbl)],
            "b200amg_profile_cycle": [vp, i32, vp, i32],
            "b200amg_device_vectors": [vp, C.POINTER(vp), C.POINTER(vp)],
            "b200amg_set_option": [vp, i32, dbl],
            "b200amg_residual_timings": [vp, vp, i32, C.POINTER(i32)],
            "b200amg_get_stream": [vp, C.POINTER(vp)],
            "b200amg_nccl_unique_id": [vp, i64],
            "b200amg_partition_info": [vp] + [C.POINTER(i64)] * 8,
            "b200amg_partition_plan": [pcsc, pcsc, pcsc, i32, i32, vp, vp, vp, C.POINTER(i64), vp, vp, C.POINTER(i64), vp, vp, vp, i64],
            "b200amg_partition_plan_child": [pcsc, pcsc, pcsc, pcsc, vp, vp, i32, i32, vp, vp, vp, C.POINTER(i64), vp, vp, C.POINTER(i64),
                                             vp, vp, vp, i64],
            "b200amg_debug_gs_timeline": [vp, i32, i32, vp, i64, C.POINTER(i64)],
            "b200amg_block_plan_check": [pcsc, vp, vp, vp, vp, vp, vp, dbl, i32, i32, C.c_char_p, i64],
        }
        for name, args in sigs.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = i32
        L.b200amg_storage_info.argtypes = [vp, i32, vp, i32]
        L.b200amg_storage_info.restype = i32
        L.b200amg_comm_stats.argtypes = [vp, vp, i32]
        L.b200amg_comm_stats.restype = i32
        L.b200amg_launch_count.argtypes = [vp]
        L.b200amg_launch_count.restype = i64
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise B200AmgError(rc, lib().b200amg_last_error().decode(errors="replace"))


def device_count() -> int:
    return int(lib().b200amg_device_count())


def nccl_unique_id() -> bytes:
    """The 128-byte ncclUniqueId rank 0 creates and broadcasts (``MultiLevel.partition``)."""
    buf = C.create_string_buffer(128)
    _check(lib().b200amg_nccl_unique_id(C.cast(buf, C.c_void_p), 128))
    return buf.raw


def partition_plan(level, rank, world, parent_level=None, parent_plan=None):
    """Host-only partition plan of one ``Level`` (``b200amg_partition_plan``) as a dict of numpy arrays; with
    ``parent_level`` / ``parent_plan`` the plan of a level below a partitioned one (``b200amg_partition_plan_child``)."""
    keep = []
    a, p, r = csc_desc(level.A, keep), csc_desc(level.P, keep), csc_desc(level.R, keep)
    n = level.A.n
    cap = n + 8
    out = {"row_split": np.zeros(world + 1, np.int64), "coarse_split": np.zeros(world + 1, np.int64),
           "halo_cols": np.zeros(cap, np.int32), "recv_off": np.zeros(world + 1, np.int32),
           "send_idx": np.zeros(cap, np.int32), "send_off": np.zeros(world + 1, np.int32),
           "cx_lo": np.zeros(world, np.int64), "cx_hi": np.zeros(world, np.int64)}
    nhalo, nsend = C.c_int64(0), C.c_int64(0)
    tail = (rank, world, _ptr(out["row_split"]), _ptr(out["coarse_split"]), _ptr(out["halo_cols"]), C.byref(nhalo), _ptr(out["recv_off"]),
            _ptr(out["send_idx"]), C.byref(nsend), _ptr(out["send_off"]), _ptr(out["cx_lo"]), _ptr(out["cx_hi"]), C.c_int64(cap))
    if parent_level is None:
        _check(lib().b200amg_partition_plan(C.byref(a), C.byref(p), C.byref(r), *tail))
    else:
        pp = csc_desc(parent_level.P, keep)
        prs = np.ascontiguousarray(parent_plan["row_split"], dtype=np.int64)
        pcs = np.ascontiguousarray(parent_plan["coarse_split"], dtype=np.int64)
        _check(lib().b200amg_partition_plan_child(C.byref(a), C.byref(p), C.byref(r), C.byref(pp), _ptr(prs), _ptr(pcs), *tail))
    out["halo_cols"] = out["halo_cols"][: nhalo.value].copy()
    out["send_idx"] = out["send_idx"][: nsend.value].copy()
    return out


def spgemm(a, b, device=None):
    """``a * b`` on the device (``b200amg_spgemm_begin`` / ``_fetch``): the Galerkin products of the setup phase, same
    contract and bit-identical result as the host product ``_hostlib.spgemm`` (structural zeros kept)."""
    from ._hostlib import _csc

    if a.n != b.m:
        raise ValueError(f"DimensionMismatch: {a.shape} * {b.shape}")
    L = lib()
    arrs = [np.ascontiguousarray(v, dtype=t) for v, t in ((a.colptr, np.int32), (a.rowval, np.int32), (a.nzval, np.float64),
                                                          (b.colptr, np.int32), (b.rowval, np.int32), (b.nzval, np.float64))]
    nnz = C.c_int64(0)
    dev = _default_device() if device is None else device
    _check(L.b200amg_spgemm_begin(dev, C.c_int64(a.m), C.c_int64(a.n), C.c_int64(b.n), *[_ptr(v) for v in arrs], C.byref(nnz)))
    cp = np.empty(b.n + 1, np.int32)
    cj = np.empty(nnz.value, np.int32)
    cx = np.empty(nnz.value, np.float64)
    _check(L.b200amg_spgemm_fetch(_ptr(cp), _ptr(cj), _ptr(cx)))
    return _csc(a.m, b.n, cp, cj, cx)


def _ptr(a):
    """Pointer of a numpy array (host) or a raw integer device pointer / object with data_ptr()."""
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(int(a))


def _memkind(a):
    if isinstance(a, np.ndarray):
        return MEM_HOST
    if hasattr(a, "is_cuda"):
        if not a.is_cuda:
            raise TypeError("torch tensors passed to the engine must live on the GPU (or pass numpy arrays)")
        return MEM_DEVICE
    return MEM_DEVICE


def csc_desc(op, keep, julia_indices=False):
    """``b200amg_csc_t`` for a ``SparseMatrixCSC`` or a lazy ``Adjoint`` of one.  ``julia_indices``: hand the arrays over the way a
    Julia caller holds them — ``Int64``, 1-based (``SparseMatrixCSC{Float64,Int64}``, ``src/classical.jl:64-65``) — instead of
    this package's own int32 / 0-based storage."""
    adj = 0
    if hasattr(op, "parent"):
        op, adj = op.parent, 1
    keep.append(op)
    d = CscDesc()
    d.m, d.n = op.m, op.n
    colptr, rowval = op.colptr, op.rowval
    base = 0
    if julia_indices:
        colptr = np.ascontiguousarray(op.colptr, dtype=np.int64) + 1
        rowval = np.ascontiguousarray(op.rowval, dtype=np.int64) + 1
        keep += [colptr, rowval]
        base = 1
    d.colptr, d.rowval, d.nzval = colptr.ctypes.data, rowval.ctypes.data, op.nzval.ctypes.data
    d.index_bits = colptr.dtype.itemsize * 8
    d.index_base = base
    d.adjoint = adj
    return d


BLOCK_STATS = ("ok", "tiles", "stages", "steps", "lanes", "wavefronts", "mean_step_rows_x1000", "theta", "a", "b", "max_tile_rows",
               "max_tile_steps", "req_fwd", "req_bwd", "k_extent", "j_extent")


def block_plan_check(A, x=None, b=None, omega=1.0, sor=False, sweep=3, tile_rows=0, block_a=0, block_b=0, stage_nnz=0, stage_rows=0,
                     window=0, depth=0, verbose=0, emulate_pass=False):
    """Host-only: build + validate the blocked Gauss-Seidel plan of ``A`` (``b200amg_block_plan_check``) and, when ``x`` and ``b``
    are given, run the host emulation of the kernel's sweep (``emulate_pass``: of the pass sweep, pass_gs.cuh, on the same plan).  Returns ``(stats, message, new_of_old, x_out)``."""
    import numpy as np

    keep = []
    d = csc_desc(A, keep)
    params = np.array([tile_rows, block_a, block_b, stage_nnz, stage_rows, window, depth, verbose, int(bool(emulate_pass))], dtype=np.int64)
    stats = np.zeros(16, dtype=np.int64)
    perm = np.zeros(A.n, dtype=np.int32)
    msg = C.create_string_buffer(512)
    xo = None
    xp = bp = None
    if x is not None:
        xp = np.ascontiguousarray(x, dtype=np.float64)
        bp = np.ascontiguousarray(b, dtype=np.float64)
        xo = np.zeros(A.n)
    _check(lib().b200amg_block_plan_check(C.byref(d), params.ctypes.data, stats.ctypes.data, perm.ctypes.data,
                                          xp.ctypes.data if xp is not None else None, bp.ctypes.data if bp is not None else None,
                                          xo.ctypes.data if xo is not None else None, float(omega), int(bool(sor)), int(sweep), msg, 512))
    return dict(zip(BLOCK_STATS, stats.tolist())), msg.value.decode(errors="replace"), perm, xo


def smoother_desc(config):
    d = SmootherDesc()
    if config is None:
        d.kind = 0
        return d
    d.kind = KIND[config.kind]
    d.sweep = SWEEP[getattr(config, "sweep_name", "symmetric")]
    d.iter = int(config.iter)
    d.omega = float(getattr(config, "omega", 1.0))
    return d


COARSE_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double))


def _coarse_trampoline(cs):
    """The C callback behind ``b200amg_set_coarse_callback``: wraps the pinned staging vectors as numpy arrays and calls the
    host coarse solver ``cs(x, b)``; an exception becomes a non-zero status (B200AMG_ERR_CALLBACK at the entry point) and is
    kept on the function object for the caller to inspect."""
    def call(_user, n, ncols, xp, bp):
        try:
            shape = (n,) if ncols == 1 else (n, ncols)
            x = np.ctypeslib.as_array(xp, shape=(n * ncols,)).reshape(shape, order="F")
            b = np.ctypeslib.as_array(bp, shape=(n * ncols,)).reshape(shape, order="F")
            cs(x, b)
            return 0
        except BaseException as e:   # never let an exception cross the C boundary
            call.last_exception = e
            return 1
    call.last_exception = None
    return call


def _default_device():
    if "B200AMG_DEVICE" in os.environ:
        return int(os.environ["B200AMG_DEVICE"])
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    return 0


class DeviceHierarchy:
    """Device-resident hierarchy: the handle behind a host ``MultiLevel``."""

    def __init__(self, ml, device=None, partition=None, julia_indices=False):
        L = lib()
        self._h = C.c_void_p()
        self._coarse_cb = None   # keeps the ctypes callback of a host coarse solver alive as long as the handle
        self.device = _default_device() if device is None else device
        _check(L.b200amg_create(C.byref(self._h), self.device))
        keep = []
        try:
            if partition is not None:
                rank, world, uid = partition[:3]
                buf = (C.c_char * len(uid)).from_buffer_copy(uid)
                _check(L.b200amg_set_partition(self._h, rank, world, C.cast(buf, C.c_void_p), len(uid)))
                if len(partition) > 3 and partition[3] != 1:
                    _check(L.b200amg_set_option(self._h, 12, float(partition[3])))      # B200AMG_OPT_PART_LEVELS
            for lv in ml.levels:
                a, p, r = (csc_desc(lv.A, keep, julia_indices), csc_desc(lv.P, keep, julia_indices),
                           csc_desc(lv.R, keep, julia_indices))
                pre, post = smoother_desc(lv.presmoother.config), smoother_desc(lv.postsmoother.config)
                sym = SYMMETRY[lv.presmoother.symmetry_name]
                _check(L.b200amg_add_level(self._h, C.byref(a), C.byref(p), C.byref(r), C.byref(pre), C.byref(post), sym))
            fa = csc_desc(ml.final_A, keep, julia_indices)
            cs = ml.coarse_solver
            op = cs.dense_operator() if hasattr(cs, "dense_operator") else None
            if op is not None:
                inv = np.ascontiguousarray(np.asarray(op, dtype=np.float64).reshape(-1, order="F"))
                _check(L.b200amg_set_coarse(self._h, C.byref(fa), ml.final_A.n, _ptr(inv)))
            else:
                # any callable cs(x, b) (coarse_solver.jl:24-58): run on the host from inside the cycle
                self._coarse_cb = COARSE_FN(_coarse_trampoline(cs))
                _check(L.b200amg_set_coarse_callback(self._h, C.byref(fa), ml.final_A.n, self._coarse_cb, None))
            _check(L.b200amg_finalize(self._h))
        except Exception:
            L.b200amg_destroy(self._h)
            self._h = None
            raise
        self.n = ml.levels[0].A.n if ml.levels else ml.final_A.n
        self.nlevels = len(ml.levels) + 1

    def close(self):
        if self._h:
            lib().b200amg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- solve phase -----------------------------------------------------------------------
    def solve(self, x, b, cycle=0, maxiter=100, abstol=0.0, reltol=1.4901161193847656e-08, calculate_residual=True):
        cap = int(maxiter) + 2
        res = np.zeros(cap)
        nres, iters = C.c_int32(0), C.c_int32(0)
        _check(lib().b200amg_solve(self._h, _ptr(x), _ptr(b), cycle, maxiter, abstol, reltol, int(calculate_residual),
                                   _ptr(res), cap, C.byref(nres), C.byref(iters), _memkind(x)))
        return res[: nres.value].copy(), iters.value

    def solve_block(self, x, b, cycle=0, maxiter=100, abstol=0.0, reltol=1.4901161193847656e-08, calculate_residual=True):
        """``_solve!`` for n x m blocks: ``x`` / ``b`` Fortran-ordered (column-major) float64 arrays (numpy, host)."""
        n, m = x.shape
        cap = int(maxiter) + 2
        res = np.zeros(cap)
        nres, iters = C.c_int32(0), C.c_int32(0)
        _check(lib().b200amg_solve_block(self._h, _ptr(x), _ptr(b), m, x.strides[1] // 8, cycle, maxiter, abstol, reltol,
                                         int(calculate_residual), _ptr(res), cap, C.byref(nres), C.byref(iters), MEM_HOST))
        return res[: nres.value].copy(), iters.value

    def cycle(self, x, b, cycle=0):
        _check(lib().b200amg_cycle(self._h, _ptr(x), _ptr(b), cycle, _memkind(x)))
        return x

    def precond(self, x, b, cycle=0, init_zero=True):
        _check(lib().b200amg_precond(self._h, _ptr(x), _ptr(b), cycle, int(init_zero), _memkind(x)))
        return x

    def smooth(self, level, which, x, b):
        _check(lib().b200amg_smooth(self._h, level, which, _ptr(x), _ptr(b), _memkind(x)))
        return x

    def apply(self, level, op, y, x):
        _check(lib().b200amg_apply(self._h, level, op, _ptr(y), _ptr(x), _memkind(x)))
        return y

    def residual(self, level, r, b, x):
        _check(lib().b200amg_residual(self._h, level, _ptr(r), _ptr(b), _ptr(x), _memkind(x)))
        return r

    def coarse_solve(self, x, b):
        _check(lib().b200amg_coarse_solve(self._h, _ptr(x), _ptr(b), _memkind(x)))
        return x

    def norm(self, v, n=None):
        out = C.c_double(0)
        n = int(v.size if n is None and isinstance(v, np.ndarray) else (n if n is not None else v.numel()))
        _check(lib().b200amg_norm(self._h, n, _ptr(v), C.byref(out), _memkind(v)))
        return out.value

    def pcg(self, x, b, cycle=0, maxiter=None, abstol=0.0, reltol=1.4901161193847656e-08):
        maxiter = self.n if maxiter is None else int(maxiter)
        cap = maxiter + 2
        res = np.zeros(cap)
        nres, iters = C.c_int32(0), C.c_int32(0)
        _check(lib().b200amg_pcg(self._h, _ptr(x), _ptr(b), cycle, maxiter, abstol, reltol, _ptr(res), cap,
                                 C.byref(nres), C.byref(iters), _memkind(x)))
        return res[: nres.value].copy(), iters.value

    # -- introspection / measurement -----------------------------------------------------------
    def level_info(self, level):
        n, nnza, nnzp, wf = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        _check(lib().b200amg_level_info(self._h, level, C.byref(n), C.byref(nnza), C.byref(nnzp), C.byref(wf)))
        return {"n": n.value, "nnz_a": nnza.value, "nnz_p": nnzp.value, "wavefronts": wf.value}

    def launch_count(self):
        return int(lib().b200amg_launch_count(self._h))

    def storage_info(self, level):
        """Bytes per stored matrix value the bandwidth kernels read on ``level``: ``{"A": 4 | 8, "P": .., "R": ..}``."""
        out = np.zeros(3, dtype=np.int32)
        _check(lib().b200amg_storage_info(self._h, level, out.ctypes.data, 3))
        return {"A": int(out[0]), "P": int(out[1]), "R": int(out[2])}

    def comm_stats(self):
        out = np.zeros(4, dtype=np.int64)
        _check(lib().b200amg_comm_stats(self._h, out.ctypes.data, 4))
        return {"nccl_groups": int(out[0]), "peer_exchanges": int(out[1]), "peer_halo": bool(out[2]), "partitioned_levels": int(out[3])}

    def time_kernel(self, level, what, cycle=0, reps=20, flush_l2=False):
        ms = C.c_double(0)
        _check(lib().b200amg_time_kernel(self._h, level, what, cycle, reps, int(flush_l2), C.byref(ms)))
        return ms.value

    def profile_cycle(self, cycle=0):
        cap = 6 * self.nlevels
        ms = np.zeros(cap)
        _check(lib().b200amg_profile_cycle(self._h, cycle, _ptr(ms), cap))
        return ms.reshape(self.nlevels, 6)

    def partition_info(self):
        v = [C.c_int64() for _ in range(8)]
        _check(lib().b200amg_partition_info(self._h, *[C.byref(t) for t in v]))
        keys = ["row_lo", "row_hi", "nhalo", "nsend", "coarse_lo", "coarse_hi", "cx_lo", "cx_hi"]
        return {k: t.value for k, t in zip(keys, v)}

    def set_option(self, option, value):
        _check(lib().b200amg_set_option(self._h, int(option), float(value)))

    def residual_timings(self, cap=4096):
        ms = np.zeros(cap)
        n = C.c_int32(0)
        _check(lib().b200amg_residual_timings(self._h, _ptr(ms), cap, C.byref(n)))
        return ms[: n.value].copy()

    def gs_timeline(self, level, backward=False):
        cap = 8 * (self.level_info(level)["n"] + 8)
        out = np.zeros(cap, dtype=np.uint64)
        nt = C.c_int64(0)
        _check(lib().b200amg_debug_gs_timeline(self._h, level, int(backward), _ptr(out), cap, C.byref(nt)))
        return out[: 8 * nt.value].reshape(-1, 8).astype(np.int64)

    def stream(self):
        s = C.c_void_p()
        _check(lib().b200amg_get_stream(self._h, C.byref(s)))
        return s.value or 0

    def device_vectors(self):
        x, b = C.c_void_p(), C.c_void_p()
        _check(lib().b200amg_device_vectors(self._h, C.byref(x), C.byref(b)))
        return x.value, b.value


class DeviceSmoother:
    """Device-side cache of a standalone smoother (``setup_smoother`` / ``smooth!``)."""

    def __init__(self, A, config, symmetry_name, device=None):
        L = lib()
        self._s = C.c_void_p()
        keep = []
        a = csc_desc(A, keep)
        cfg = smoother_desc(config)
        self.n = A.n
        rc = L.b200amg_smoother_create(C.byref(self._s), _default_device() if device is None else device, C.byref(a),
                                       C.byref(cfg), SYMMETRY[symmetry_name])
        if rc == -3:
            from .smoother import SingularException

            msg = L.b200amg_last_error().decode()
            raise SingularException(int(msg[msg.index("(") + 1: msg.index(")")]))
        _check(rc)

    def apply(self, x, b):
        """``smooth!(x, s, b)``.  A 2-D ``x`` / ``b`` (n x m, the reference's block right-hand sides) is relaxed column by
        column, as ``smooth!`` itself does (``for col in 1:size(x, 2)``, ``src/smoother.jl:77,118,195``): the C entry point
        takes ONE vector of n doubles, and a C-contiguous n x m buffer is not m vectors laid end to end."""
        if np.ndim(x) == 2:
            if np.ndim(b) != 2 or np.shape(b) != np.shape(x):
                raise AssertionError("x and b must have the same shape")
            if np.shape(x)[0] != self.n:
                raise ValueError(f"x has {np.shape(x)[0]} rows, the smoother's matrix {self.n}")
            for j in range(np.shape(x)[1]):
                xj = np.ascontiguousarray(x[:, j], dtype=np.float64) if isinstance(x, np.ndarray) else x[:, j].contiguous()
                bj = np.ascontiguousarray(b[:, j], dtype=np.float64) if isinstance(b, np.ndarray) else b[:, j].contiguous()
                _check(lib().b200amg_smoother_apply(self._s, _ptr(xj), _ptr(bj), _memkind(xj)))
                x[:, j] = xj
            return x
        if np.ndim(x) != 1 or np.shape(x)[0] != self.n or np.shape(b) != np.shape(x):
            raise ValueError(f"x and b must be vectors of length {self.n} (or n x m blocks)")
        xd = x if not isinstance(x, np.ndarray) else np.ascontiguousarray(x, dtype=np.float64)
        bd = b if not isinstance(b, np.ndarray) else np.ascontiguousarray(b, dtype=np.float64)
        _check(lib().b200amg_smoother_apply(self._s, _ptr(xd), _ptr(bd), _memkind(xd)))
        if isinstance(x, np.ndarray) and xd is not x:
            x[...] = xd
        return x

    def close(self):
        if self._s:
            lib().b200amg_smoother_destroy(self._s)
            self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
