"""ctypes binding of libb200lp.so (include/b200lp.h) -- the same C ABI the Lisp CFFI shim binds.

The product path has no CPU fallback: if the CUDA library is missing, or there is no GPU, the
calls raise (`B200LibraryError` / `B200DeviceError`) instead of computing anything on the host.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200lp.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "b200lp.h")

MAX_DEVICES = 8

OK, UNBOUNDED, INFEASIBLE, ITERATION_LIMIT, ARTIFICIAL_STUCK, ARTIFICIAL_NONZERO = 0, 1, 2, 3, 4, 5
FEAS_SCALED, FEAS_REFERENCE = 0, 1
ERR_INVALID_ARG, ERR_CUDA, ERR_NCCL, ERR_NO_DEVICE, ERR_OUT_OF_MEMORY, ERR_INTERNAL = \
    -1, -2, -3, -4, -5, -6
ERR_PEER_TIMEOUT = -7
RULE_REFERENCE, RULE_BLAND = 0, 1


class B200LibraryError(RuntimeError):
    """libb200lp.so is missing or does not export the declared ABI."""


class B200DeviceError(RuntimeError):
    """A CUDA / NCCL / argument failure reported by the library (negative status)."""

    def __init__(self, code, where, detail=""):
        self.code = code
        super().__init__(f"{where}: {strerror(code)} ({code}) {detail}".strip())


class Opts(ctypes.Structure):
    _fields_ = [
        ("fp_tolerance_factor", ctypes.c_double),
        ("pivot_rule", ctypes.c_int32),
        ("writeback_full", ctypes.c_int32),
        ("max_iters", ctypes.c_int64),
        ("ndev", ctypes.c_int32),
        ("devices", ctypes.c_int32 * MAX_DEVICES),
        ("trace_capacity", ctypes.c_int32),
        ("poll_interval", ctypes.c_int32),
        ("time_kernels", ctypes.c_int32),
        ("pivot_variant", ctypes.c_int32),
        ("feas_mode", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 5),
    ]


class Result(ctypes.Structure):
    _fields_ = [
        ("status", ctypes.c_int32),
        ("n_devices", ctypes.c_int32),
        ("iterations", ctypes.c_int64),
        ("iterations_phase1", ctypes.c_int64),
        ("iterations_cleanup", ctypes.c_int64),
        ("objective", ctypes.c_double),
        ("ms_total", ctypes.c_double),
        ("ms_h2d", ctypes.c_double),
        ("ms_solve", ctypes.c_double),
        ("ms_d2h", ctypes.c_double),
        ("ms_pivot_kernel", ctypes.c_double),
        ("pivot_kernel_launches", ctypes.c_int64),
        ("kernel_launches", ctypes.c_int64),
        ("h2d_bytes", ctypes.c_int64),
        ("d2h_bytes", ctypes.c_int64),
        ("bytes_per_pivot", ctypes.c_int64),
        ("trace_len", ctypes.c_int32),
        ("exchange_mode", ctypes.c_int32),
        ("ms_look_kernel", ctypes.c_double),
        ("ms_exchange", ctypes.c_double),
        ("look_kernel_launches", ctypes.c_int64),
        ("loop_mode", ctypes.c_int32),
        ("look_ctas", ctypes.c_int32),
        ("ms_look_wait", ctypes.c_double),
        ("ms_look_ratio", ctypes.c_double),
        ("ms_look_push", ctypes.c_double),
        ("ms_look_peer_wait", ctypes.c_double),
        ("ms_look_row", ctypes.c_double),
        ("sm_clock_mhz", ctypes.c_double),
        ("ms_look_dbg", ctypes.c_double * 8),
        ("redundant_rows", ctypes.c_int64),
        ("look_cluster", ctypes.c_int32),
        ("reserved_r", ctypes.c_int32),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
_lp = ctypes.POINTER(ctypes.c_int64)
_vp = ctypes.c_void_p

# name -> (restype, argtypes); every symbol include/b200lp.h declares
SIGNATURES = {
    "b200lp_solve": (ctypes.c_int, [ctypes.POINTER(Opts), _dp, ctypes.c_int64, ctypes.c_int64,
                                    ctypes.c_int64, _ip, ctypes.c_int32, ctypes.POINTER(Result),
                                    _ip, _ip]),
    "b200lp_solve_two_phase": (ctypes.c_int, [ctypes.POINTER(Opts), _dp, ctypes.c_int64,
                                              ctypes.c_int64, _ip, _dp, ctypes.c_int64,
                                              ctypes.c_int64, ctypes.c_int64, _ip, ctypes.c_int32,
                                              ctypes.POINTER(Result)]),
    "b200lp_create": (ctypes.c_int, [ctypes.POINTER(Opts), ctypes.c_int64, ctypes.c_int64,
                                     ctypes.c_int32, ctypes.POINTER(_vp)]),
    "b200lp_destroy": (None, [_vp]),
    "b200lp_upload": (ctypes.c_int, [_vp, _dp, ctypes.c_int64, _ip]),
    "b200lp_download": (ctypes.c_int, [_vp, _dp, ctypes.c_int64, _ip]),
    "b200lp_download_solution": (ctypes.c_int, [_vp, _dp, _dp, _ip]),
    "b200lp_find_entering_column": (ctypes.c_int, [_vp, _lp]),
    "b200lp_find_pivoting_row": (ctypes.c_int, [_vp, ctypes.c_int64, _lp]),
    "b200lp_pivot": (ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_int64]),
    "b200lp_iterate": (ctypes.c_int, [_vp, ctypes.c_int64, ctypes.POINTER(Result), _ip, _ip]),
    "b200lp_set_time_kernels": (ctypes.c_int, [_vp, ctypes.c_int32]),
    "b200lp_comm_unique_id": (ctypes.c_int, [_vp]),
    "b200lp_create_sharded": (ctypes.c_int, [ctypes.POINTER(Opts), ctypes.c_int64, ctypes.c_int64,
                                             ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _vp,
                                             ctypes.POINTER(_vp)]),
    "b200lp_shard_rows": (ctypes.c_int, [_vp, _lp, _lp]),
    "b200lp_partition": (None, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, _lp, _lp]),
    "b200lp_shutdown": (None, []),
    "b200lp_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "b200lp_version": (ctypes.c_int, []),
    "b200lp_abi_sizes": (None, [_lp, _lp]),
    "b200lp_device_count": (ctypes.c_int, []),
    "b200lp_last_error": (ctypes.c_char_p, []),
    "b200lp_thresholds": (None, [ctypes.c_double, _dp, _dp, _dp]),
}

_lib = None


def lib():
    """Load libb200lp.so; fail loudly when it is absent (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200LibraryError(
                f"{LIB_PATH} not built -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C linear-programming_b200/csrc`")
        L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(L, name)
            except AttributeError as exc:
                raise B200LibraryError(f"{LIB_PATH} does not export {name}") from exc
            fn.restype = res
            fn.argtypes = args
        so, sr = ctypes.c_int64(), ctypes.c_int64()
        L.b200lp_abi_sizes(ctypes.byref(so), ctypes.byref(sr))
        if (so.value, sr.value) != (ctypes.sizeof(Opts), ctypes.sizeof(Result)):
            raise B200LibraryError(
                f"{LIB_PATH}: struct sizes {(so.value, sr.value)} != this binding's "
                f"{(ctypes.sizeof(Opts), ctypes.sizeof(Result))} -- rebuild the library")
        _lib = L
    return _lib


def strerror(code):
    return lib().b200lp_strerror(int(code)).decode()


def last_error():
    return lib().b200lp_last_error().decode()


def _check(code, where):
    if code < 0:
        raise B200DeviceError(code, where, last_error())
    return code


def make_opts(fp_tolerance=1024.0, pivot_rule=RULE_REFERENCE, max_iters=0, devices=None,
              writeback_full=False, trace_capacity=0, poll_interval=0, time_kernels=False,
              pivot_variant=0, feas_mode=FEAS_SCALED):
    o = Opts()
    o.fp_tolerance_factor = float(fp_tolerance)
    o.pivot_rule = int(pivot_rule)
    o.max_iters = int(max_iters)
    o.writeback_full = int(bool(writeback_full))
    o.trace_capacity = int(trace_capacity)
    o.poll_interval = int(poll_interval)
    o.time_kernels = int(bool(time_kernels))
    o.pivot_variant = int(pivot_variant)
    o.feas_mode = int(feas_mode)
    if devices is None:
        o.ndev = 0
    else:
        devices = list(devices)
        if not 1 <= len(devices) <= MAX_DEVICES:
            raise ValueError("1..8 devices")
        o.ndev = len(devices) if len(devices) > 1 else 0
        for k, d in enumerate(devices):
            o.devices[k] = int(d)
    return o


def thresholds(fp_tolerance=1024.0):
    e, p, f = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    lib().b200lp_thresholds(float(fp_tolerance), ctypes.byref(e), ctypes.byref(p), ctypes.byref(f))
    return e.value, p.value, f.value


def partition(m, nranks, rank):
    b, e = ctypes.c_int64(), ctypes.c_int64()
    lib().b200lp_partition(int(m), int(nranks), int(rank), ctypes.byref(b), ctypes.byref(e))
    return b.value, e.value


def device_count():
    return lib().b200lp_device_count()


def _tab_args(tab):
    if not (isinstance(tab, np.ndarray) and tab.dtype == np.float64 and tab.ndim == 2
            and tab.strides[1] == 8 and tab.strides[0] % 8 == 0 and tab.strides[0] >= tab.shape[1] * 8):
        raise ValueError("tableau must be a row-major float64 matrix")
    return tab.ctypes.data_as(_dp), tab.shape[0], tab.shape[1], tab.strides[0] // 8


def _basis_arg(basis, n):
    if not (isinstance(basis, np.ndarray) and basis.dtype == np.int32 and basis.shape == (n,)
            and basis.flags.c_contiguous):
        raise ValueError(f"basis must be a contiguous int32 vector of length {n}")
    return basis.ctypes.data_as(_ip)


def solve(tab, basis, is_max=True, opts=None):
    """b200lp_solve: n-solve-tableau on the GPU, in place on the caller's host buffers.

    Returns (status, Result, trace) where trace is a list of (entering col, leaving row)."""
    opts = opts or make_opts()
    p, R, C, ld = _tab_args(tab)
    bp = _basis_arg(basis, R - 1)
    res = Result()
    cap = max(int(opts.trace_capacity), 1)
    tj = np.zeros(cap, np.int32)
    tr = np.zeros(cap, np.int32)
    st = lib().b200lp_solve(ctypes.byref(opts), p, R, C, ld, bp, int(bool(is_max)),
                            ctypes.byref(res), tj.ctypes.data_as(_ip), tr.ctypes.data_as(_ip))
    _check(st, "b200lp_solve")
    n = res.trace_len
    return st, res, list(zip(tj[:n].tolist(), tr[:n].tolist()))


def solve_two_phase(art, art_basis, main, main_basis, is_max=True, opts=None):
    """b200lp_solve_two_phase: the list branch of n-solve-tableau. Returns (status, Result)."""
    opts = opts or make_opts()
    ap, R, C_art, ld_art = _tab_args(art)
    mp, R2, C, ld = _tab_args(main)
    if R != R2:
        raise ValueError("both tableaus must have the same number of rows")
    res = Result()
    st = lib().b200lp_solve_two_phase(ctypes.byref(opts), ap, C_art, ld_art,
                                      _basis_arg(art_basis, R - 1), mp, R, C, ld,
                                      _basis_arg(main_basis, R - 1), int(bool(is_max)),
                                      ctypes.byref(res))
    _check(st, "b200lp_solve_two_phase")
    return st, res


class DeviceTableau:
    """A tableau resident in HBM (b200lp_solver handle): the hot path one reference call at a time."""

    def __init__(self, R, C, is_max=True, opts=None, shard=None):
        """shard = (rank, nranks, unique_id_bytes) for one-process-per-GPU row-block sharding."""
        self.opts = opts or make_opts()
        self.R, self.C, self.is_max = int(R), int(C), bool(is_max)
        self._h = _vp()
        if shard is None:
            code = lib().b200lp_create(ctypes.byref(self.opts), self.R, self.C, int(self.is_max),
                                       ctypes.byref(self._h))
            self.rank, self.nranks = 0, 1
            where = "b200lp_create"
        else:
            rank, nranks, uid = shard
            buf = ctypes.create_string_buffer(bytes(uid), 128) if uid is not None else None
            code = lib().b200lp_create_sharded(ctypes.byref(self.opts), self.R, self.C,
                                               int(self.is_max), int(rank), int(nranks),
                                               ctypes.cast(buf, _vp) if buf is not None else None,
                                               ctypes.byref(self._h))
            self.rank, self.nranks = int(rank), int(nranks)
            where = "b200lp_create_sharded"
        _check(code, where)
        b, e = ctypes.c_int64(), ctypes.c_int64()
        _check(lib().b200lp_shard_rows(self._h, ctypes.byref(b), ctypes.byref(e)), "b200lp_shard_rows")
        self.row_begin, self.row_end = b.value, e.value
        self.local_rows = (self.row_end - self.row_begin) if shard is not None else self.R - 1

    def close(self):
        if getattr(self, "_h", None):
            lib().b200lp_destroy(self._h)
            self._h = _vp()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def upload(self, tab, basis):
        p, R, C, ld = _tab_args(tab)
        if (R, C) != (self.local_rows + 1, self.C):
            raise ValueError(f"expected a {(self.local_rows + 1, self.C)} block, got {(R, C)}")
        _check(lib().b200lp_upload(self._h, p, ld, _basis_arg(basis, self.local_rows)), "b200lp_upload")

    def download(self):
        tab = np.empty((self.local_rows + 1, self.C))
        basis = np.empty(self.local_rows, np.int32)
        _check(lib().b200lp_download(self._h, tab.ctypes.data_as(_dp), self.C,
                                     basis.ctypes.data_as(_ip)), "b200lp_download")
        return tab, basis

    def download_solution(self):
        rhs = np.empty(self.local_rows + 1)
        obj = np.empty(self.C)
        basis = np.empty(self.local_rows, np.int32)
        _check(lib().b200lp_download_solution(self._h, rhs.ctypes.data_as(_dp),
                                              obj.ctypes.data_as(_dp), basis.ctypes.data_as(_ip)),
               "b200lp_download_solution")
        return rhs, obj, basis

    def find_entering_column(self):
        j = ctypes.c_int64()
        _check(lib().b200lp_find_entering_column(self._h, ctypes.byref(j)), "b200lp_find_entering_column")
        return None if j.value < 0 else j.value

    def find_pivoting_row(self, j):
        r = ctypes.c_int64()
        _check(lib().b200lp_find_pivoting_row(self._h, int(j), ctypes.byref(r)), "b200lp_find_pivoting_row")
        return None if r.value < 0 else r.value

    def pivot(self, j, r):
        _check(lib().b200lp_pivot(self._h, int(j), int(r)), "b200lp_pivot")

    def set_time_kernels(self, on):
        _check(lib().b200lp_set_time_kernels(self._h, int(bool(on))), "b200lp_set_time_kernels")

    def iterate(self, max_iters=0):
        """Returns (status, Result, trace)."""
        res = Result()
        cap = max(int(self.opts.trace_capacity), 1)
        tj = np.zeros(cap, np.int32)
        tr = np.zeros(cap, np.int32)
        st = lib().b200lp_iterate(self._h, int(max_iters), ctypes.byref(res),
                                  tj.ctypes.data_as(_ip), tr.ctypes.data_as(_ip))
        _check(st, "b200lp_iterate")
        n = res.trace_len
        return st, res, list(zip(tj[:n].tolist(), tr[:n].tolist()))


def shutdown():
    """Free the idle handles pooled by the one-shot calls."""
    lib().b200lp_shutdown()


def comm_unique_id():
    buf = ctypes.create_string_buffer(128)
    _check(lib().b200lp_comm_unique_id(ctypes.cast(buf, _vp)), "b200lp_comm_unique_id")
    return buf.raw
