"""ctypes/numpy front end of the CHECKERS in oracle/ -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package lj_gpu_b200 never does.

Two back ends:
  * Oracle      -- oracle/liblj_oracle.so, the C restatement (lj_oracle.c);
  * RefProcess  -- oracle/_ref/libljref_d<rho>.so, the REAL reference cpu_ref/force_soa.cpp
                   compiled by oracle/Makefile from /root/reference (prebuilt on the GPU box).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liblj_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

CUTOFF_LENGTH = 3.0
SEARCH_LENGTH = 3.3
CL2 = CUTOFF_LENGTH * CUTOFF_LENGTH
SL2 = SEARCH_LENGTH * SEARCH_LENGTH
DT = 0.001

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile the C restatement (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(ORACLE_SO) or \
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "lj_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "liblj_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/cpu_ref"):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)


def ref_so(density: float) -> str:
    return os.path.join(REF_DIR, "libljref_d%.1f.so" % density)


def have_ref(density: float) -> bool:
    return os.path.exists(ref_so(density))


class Oracle:
    """The C restatement.  All arrays are numpy; `pointer` is int64."""

    def __init__(self):
        build()
        L = self.lib = C.CDLL(ORACLE_SO)
        L.ljo_init_fcc.restype = C.c_int64
        L.ljo_init_fcc.argtypes = [C.c_double, C.c_double, _f64p, C.c_int64, C.POINTER(C.c_int)]
        for name in ("ljo_makepair_brute", "ljo_makepair_cell"):
            f = getattr(L, name)
            f.restype = C.c_int64
            f.argtypes = [_f64p, C.c_int64, C.c_double, C.c_int, _i32p, _i64p, _i32p, C.c_int64]
        L.ljo_force_rows.restype = None
        L.ljo_force_rows.argtypes = [_f64p, _i64p, C.c_int64, _i32p, _i64p, _i32p, C.c_double, C.c_double,
                                     C.c_int, _f64p]
        for name in ("ljo_force_sorted", "ljo_force_gather", "ljo_force_gather_static"):
            f = getattr(L, name)
            f.restype = None
            f.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                          C.c_int64, C.c_double, C.c_double, _i32p, _i32p, _i64p, C.c_int]
        L.ljo_force_gather_ell.restype = None
        L.ljo_force_gather_ell.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                           C.c_int64, C.c_int64, C.c_double, C.c_double, _i32p,
                                           _i32p, C.c_int]
        L.ljo_rows_brute.restype = C.c_int64
        L.ljo_rows_brute.argtypes = [_f64p, C.c_int64, C.c_double, C.c_int, _i64p, C.c_int64, _i32p, _i64p,
                                     _i32p, C.c_int64]
        L.ljo_shuffle_rows.restype = None
        L.ljo_shuffle_rows.argtypes = [_i32p, _i32p, _i64p, C.c_int64, C.c_uint32]
        L.ljo_transpose_list.restype = C.c_int32
        L.ljo_transpose_list.argtypes = [_i32p, _i32p, _i64p, C.c_int64, _i32p, C.c_int64]
        L.ljo_num_threads.restype = C.c_int
        L.ljo_set_num_threads.argtypes = [C.c_int]

    # ---- generator ----
    def init_fcc(self, density: float, L: float = 50.0) -> np.ndarray:
        """-> q as float64 [pn, 3] (the reference's init(), cpu_ref/force_soa.cpp:115-140)."""
        s = 1.0 / (density * 0.25) ** (1.0 / 3.0)
        n = int(L / s) + 1
        q = np.empty((4 * n * n * n, 3), np.float64)
        cells = C.c_int(0)
        pn = self.lib.ljo_init_fcc(density, L, q.reshape(-1), q.shape[0], C.byref(cells))
        assert pn >= 0
        return np.ascontiguousarray(q[:pn])

    # ---- neighbour list ----
    def makepair(self, q: np.ndarray, search_len: float = SEARCH_LENGTH, full: bool = True,
                 brute: bool = False, cap: int | None = None):
        """-> (number_of_partners int32[pn], pointer int64[pn], sorted_list int32[P])"""
        q = np.ascontiguousarray(q, np.float64)
        pn = q.shape[0]
        if cap is None:
            cap = max(1024, int(pn) * 200)
        nop = np.zeros(pn, np.int32)
        ptr = np.zeros(pn, np.int64)
        lst = np.zeros(cap, np.int32)
        if brute:
            total = self.lib.ljo_makepair_brute(q.reshape(-1), pn, search_len * search_len,
                                                int(full), nop, ptr, lst, cap)
        else:
            total = self.lib.ljo_makepair_cell(q.reshape(-1), pn, search_len, int(full), nop, ptr,
                                               lst, cap)
        if total < 0:
            return self.makepair(q, search_len, full, brute, cap=-total)
        return nop, ptr, lst[:total].copy()

    def rows_brute(self, q: np.ndarray, rows, search_len: float = SEARCH_LENGTH, full: bool = True):
        """Brute-force rows of the SAMPLED particles `rows`, each against all of q (ascending j)
        -> (number_of_partners int32[n], pointer int64[n+1], list int32[total])."""
        q = np.ascontiguousarray(q, np.float64)
        rows = np.ascontiguousarray(rows, np.int64)
        n = len(rows)
        nop = np.zeros(n, np.int32)
        ptr = np.zeros(n + 1, np.int64)
        cap = max(1024, 200 * n)
        while True:
            lst = np.zeros(cap, np.int32)
            total = self.lib.ljo_rows_brute(q.reshape(-1), q.shape[0], search_len * search_len, int(full),
                                            rows, n, nop, ptr, lst, cap)
            if total >= 0:
                return nop, ptr, lst[:total].copy()
            cap = -total

    # ---- force ----
    @staticmethod
    def _strides(a: np.ndarray):
        """(component stride, element stride) in doubles for [pn,3], [pn,4] (AoS) or [3+,stride] (SoA)."""
        assert a.dtype == np.float64
        if a.shape[0] in (3, 4) and a.shape[1] > 4:  # SoA planes
            return a.strides[0] // 8, a.strides[1] // 8
        return a.strides[1] // 8, a.strides[0] // 8

    def _force(self, fn, q, p, pn, nop, ptr, lst, steps, dt, cl2):
        qc, qe = self._strides(q)
        pc, pe = self._strides(p)
        fn(q.ctypes.data, qc, qe, p.ctypes.data, pc, pe, pn, dt, cl2,
           np.ascontiguousarray(lst, np.int32), np.ascontiguousarray(nop, np.int32),
           np.ascontiguousarray(ptr, np.int64), steps)

    def force_sorted(self, q, p, nop, ptr, lst, steps=1, dt=DT, cl2=CL2, pn=None):
        """half list + Newton-3, in place on p (cpu_ref/force_soa.cpp:163-195)."""
        self._force(self.lib.ljo_force_sorted, q, p, len(nop) if pn is None else pn, nop, ptr, lst,
                    steps, dt, cl2)

    def force_gather(self, q, p, nop, ptr, lst, steps=1, dt=DT, cl2=CL2, pn=None, static_q=False):
        """full list gather, in place on p (cuda/kernel.cuh:36-65).  static_q=True: the caller
        states that q is the same for all `steps` (the reference benchmark); the per-step increment
        is evaluated once and accumulated `steps` times -- the same bits at 1/steps of the cost."""
        fn = self.lib.ljo_force_gather_static if static_q else self.lib.ljo_force_gather
        self._force(fn, q, p, len(nop) if pn is None else pn, nop, ptr, lst, steps, dt, cl2)

    def force_rows(self, q, rows, nop, ptr, lst, steps=1, dt=DT, cl2=CL2):
        """momenta [n,3] of the sampled particles `rows` after `steps` gather applications on
        their rows (nop, ptr, lst) as rows_brute() returns them."""
        q = np.ascontiguousarray(q, np.float64)
        rows = np.ascontiguousarray(rows, np.int64)
        out = np.zeros((len(rows), 3), np.float64)
        self.lib.ljo_force_rows(q.reshape(-1), rows, len(rows), np.ascontiguousarray(nop, np.int32),
                                np.ascontiguousarray(ptr, np.int64), np.ascontiguousarray(lst, np.int32),
                                dt, cl2, steps, out.reshape(-1))
        return out

    def force_gather_ell(self, q, p, nop, tlist, steps=1, dt=DT, cl2=CL2):
        qc, qe = self._strides(q)
        pc, pe = self._strides(p)
        self.lib.ljo_force_gather_ell(q.ctypes.data, qc, qe, p.ctypes.data, pc, pe, len(nop), dt,
                                      cl2, tlist, np.ascontiguousarray(nop, np.int32), steps)

    def shuffle_rows(self, lst, nop, ptr, seed=10):
        self.lib.ljo_shuffle_rows(lst, np.ascontiguousarray(nop, np.int32),
                                  np.ascontiguousarray(ptr, np.int64), len(nop), seed)

    def transpose_list(self, lst, nop, ptr):
        pn = len(nop)
        max_np = int(nop.max()) if pn else 0
        out = np.empty(max(1, max_np * pn), np.int32)
        r = self.lib.ljo_transpose_list(lst, nop, np.ascontiguousarray(ptr, np.int64), pn, out,
                                        out.size)
        assert r == max_np
        return out, max_np

    def num_threads(self) -> int:
        return int(self.lib.ljo_num_threads())

    def set_num_threads(self, n: int) -> None:
        self.lib.ljo_set_num_threads(int(n))


class Ref:
    """The real reference (cpu_ref/force_soa.cpp) loaded in THIS process.

    init() owns a function-static mt19937, so one instance per density per process; use
    oracle/ref_worker.py (a subprocess) when a fresh state is needed.
    """
    _loaded: dict = {}

    def __init__(self, density: float, L: float = 50.0):
        key = "%.1f" % density
        if key in Ref._loaded:
            raise RuntimeError("reference for density %s already initialised in this process" % key)
        path = ref_so(density)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        lib = self.lib = C.CDLL(path)
        Ref._loaded[key] = self
        lib.ljref_density.restype = C.c_double
        lib.ljref_set_L.argtypes = [C.c_double]
        for n in ("ljref_q", "ljref_p"):
            getattr(lib, n).restype = C.POINTER(C.c_double)
        for n in ("ljref_number_of_partners", "ljref_pointer", "ljref_sorted_list",
                  "ljref_i_particles", "ljref_j_particles"):
            getattr(lib, n).restype = C.POINTER(C.c_int)
        lib.ljref_force.argtypes = [C.c_int, C.c_int]
        lib.ljref_std_shuffle_rows.argtypes = [_i32p, _i32p, _i64p, C.c_int64, C.c_uint32]
        assert abs(lib.ljref_density() - density) < 1e-12
        lib.ljref_set_L(L)
        lib.ljref_init()
        self.pn = lib.ljref_particle_number()
        self.stride = lib.ljref_plane_stride()

    def _planes(self, ptr):
        return np.ctypeslib.as_array(ptr, shape=(4, self.stride))

    @property
    def q_soa(self):
        return self._planes(self.lib.ljref_q())

    @property
    def p_soa(self):
        return self._planes(self.lib.ljref_p())

    def q_xyz(self) -> np.ndarray:
        return np.ascontiguousarray(self.q_soa[:3, :self.pn].T)

    def p_xyz(self) -> np.ndarray:
        return np.ascontiguousarray(self.p_soa[:3, :self.pn].T)

    def makepair(self):
        """reference makepair()+sortpair(): half list -> (nop int32, pointer int64, list int32)."""
        self.lib.ljref_makepair()
        self.lib.ljref_sortpair()
        npairs = self.lib.ljref_number_of_pairs()
        nop = np.ctypeslib.as_array(self.lib.ljref_number_of_partners(), shape=(self.pn,)).astype(np.int32)
        ptr = np.ctypeslib.as_array(self.lib.ljref_pointer(), shape=(self.pn,)).astype(np.int64)
        lst = np.ctypeslib.as_array(self.lib.ljref_sorted_list(), shape=(npairs,)).astype(np.int32)
        return nop, ptr, lst

    def savepair(self):
        """reference savepair(): makepair() + write ./pair.dat (cpu_ref/force_soa.cpp:369-377)."""
        self.lib.ljref_savepair()

    def loadpair(self):
        """reference loadpair(): read ./pair.dat into its globals (cpu_ref/force_soa.cpp:360-367)."""
        self.lib.ljref_loadpair()

    def pair_arrays(self):
        npairs = self.lib.ljref_number_of_pairs()
        nop = np.ctypeslib.as_array(self.lib.ljref_number_of_partners(), shape=(self.pn,)).astype(np.int32)
        ip = np.ctypeslib.as_array(self.lib.ljref_i_particles(), shape=(npairs,)).astype(np.int32)
        jp = np.ctypeslib.as_array(self.lib.ljref_j_particles(), shape=(npairs,)).astype(np.int32)
        return nop, ip, jp

    def force(self, kind: str = "sorted", steps: int = 100):
        self.lib.ljref_force({"pair": 0, "sorted": 1, "next": 2, "intrin": 3}[kind], steps)

    def zero_p(self):
        self.lib.ljref_zero_p()

    def std_shuffle_rows(self, lst, nop, ptr, seed=10):
        self.lib.ljref_std_shuffle_rows(lst, np.ascontiguousarray(nop, np.int32),
                                        np.ascontiguousarray(ptr, np.int64), len(nop), seed)


def half_to_full(nop_h, ptr_h, lst_h):
    """Directed (full) list from a half list (i<j): rows ascending in j, like
    cuda/force_cuda.cu:138-162 builds it."""
    pn = len(nop_h)
    i_idx = np.repeat(np.arange(pn, dtype=np.int64), nop_h)
    j_idx = lst_h.astype(np.int64)
    src = np.concatenate([i_idx, j_idx])
    dst = np.concatenate([j_idx, i_idx])
    order = np.lexsort((dst, src))
    src, dst = src[order], dst[order]
    nop = np.bincount(src, minlength=pn).astype(np.int32)
    ptr = np.zeros(pn, np.int64)
    np.cumsum(nop[:-1], out=ptr[1:])
    return nop, ptr, dst.astype(np.int32)


def sort_rows(nop, ptr, lst):
    """Rows sorted ascending (canonical form for set-equality of lists)."""
    pn = len(nop)
    row = np.repeat(np.arange(pn, dtype=np.int64), nop)
    order = np.lexsort((lst, row))
    return lst[order]


def print_results_lines(p_xyz: np.ndarray):
    """The reference's print_results() format (cuda/force_cuda.cu:344-352)."""
    pn = p_xyz.shape[0]
    rows = list(range(5)) + list(range(pn - 5, pn))
    return ["%.10f %.10f %.10f" % tuple(p_xyz[i, :3]) for i in rows]
