"""Host-side mirror of the reference driver's interface, on top of the C ABI.

Names follow the reference (cuda/force_cuda.cu): init, makepair, random_shfl,
make_transposed_pairlist, measure, print_results, cuda_ptr.  Device arrays are torch CUDA
tensors (torch is only the allocator/stream plumbing here); every computation goes through
liblj_b200.so via lj_gpu_b200._capi.  Nothing in this module computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi as capi
from ._capi import (LJ_AOS_D3, LJ_AOS_D4, LJ_LIST_CSR, LJ_LIST_ELL, LJ_PREC_FP64, LJ_PREC_MIXED,
                    LJ_SOA_D, LJ_VARIANT_AUTO, LJ_VARIANT_NEWTON3, LJ_VARIANT_SUBWARP,
                    LJ_VARIANT_TILE_TMA)

# constants of the reference driver (cuda/force_cuda.cu:12-23,37-39)
DENSITY = 0.5
L_BOX = 50.0
DT = 0.001
CUTOFF_LENGTH = 3.0
SEARCH_LENGTH = 3.3
CL2 = CUTOFF_LENGTH * CUTOFF_LENGTH
LOOP = 100

LAYOUTS = {"aos3": LJ_AOS_D3, "aos4": LJ_AOS_D4, "soa": LJ_SOA_D, "f4": capi.LJ_AOS_F4, "f3": capi.LJ_AOS_F3}
VARIANTS = {"auto": LJ_VARIANT_AUTO, "subwarp": LJ_VARIANT_SUBWARP, "warp": LJ_VARIANT_SUBWARP,
            "thread": LJ_VARIANT_SUBWARP, "tile": LJ_VARIANT_TILE_TMA, "n3": LJ_VARIANT_NEWTON3,
            "cluster": capi.LJ_VARIANT_CLUSTER, "celltile": capi.LJ_VARIANT_CELLTILE}


class LJError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("lj_b200 status %d (%s): %s" % (
            status, capi.load().lj_status_string(status).decode(), message))
        self.status = status


def init_fcc(density: float = DENSITY, L: float = L_BOX) -> np.ndarray:
    """init() of the reference (cuda/force_cuda.cu:60-94): float64 [pn,3] on the host."""
    lib = capi.load()
    cells = C.c_int32(0)
    need = -lib.lj_init_fcc(density, L, None, 0, C.byref(cells))
    q = np.empty((max(need, 0), 3), np.float64)
    if need > 0:
        got = lib.lj_init_fcc(density, L, q.ctypes.data, need, C.byref(cells))
        assert got == need
    return q


def print_results(p_xyz: np.ndarray) -> str:
    """print_results() (cuda/force_cuda.cu:344-352): p[0..4] and p[pn-5..pn-1], %.10f."""
    pn = p_xyz.shape[0]
    rows = list(range(5)) + list(range(pn - 5, pn))
    return "".join("%.10f %.10f %.10f\n" % tuple(p_xyz[i, :3]) for i in rows)


# ---------------------------------------------------------------------- pair-list cache files
REF_N_STATIC = 400000           # cpu_ref/force_soa.cpp:11
REF_MAX_PAIRS = 30 * REF_N_STATIC  # cpu_ref/force_soa.cpp:12


def makepaircache(path: str, number_of_partners, pointer, sorted_list) -> None:
    """makepaircache() (cuda/force_cuda.cu:165-176): the text cache `.cache_pair_{all,half}.dat`."""
    nop, ptr, lst = (np.ascontiguousarray(x, np.int32) for x in (number_of_partners, pointer, sorted_list))
    rc = capi.load().lj_paircache_write_text(path.encode(), len(nop), len(lst), nop.ctypes.data,
                                             ptr.ctypes.data, lst.ctypes.data)
    if rc:
        raise LJError(rc, "lj_paircache_write_text(%s)" % path)


def loadpair(path: str, particle_number: int = -1):
    """loadpair() + check_loadedpair() (cuda/force_cuda.cu:183-227) -> (nop, pointer, sorted_list)."""
    lib = capi.load()
    pn, npairs = C.c_int64(0), C.c_int64(0)
    rc = lib.lj_paircache_read_text(path.encode(), particle_number, C.byref(pn), C.byref(npairs), None, None,
                                    0, None, 0)
    if rc:
        raise LJError(rc, "pair-list cache %s is missing or broken" % path)
    nop, ptr = np.empty(pn.value, np.int32), np.empty(pn.value, np.int32)
    lst = np.empty(npairs.value, np.int32)
    rc = lib.lj_paircache_read_text(path.encode(), particle_number, C.byref(pn), C.byref(npairs),
                                    nop.ctypes.data, ptr.ctypes.data, len(nop), lst.ctypes.data, len(lst))
    if rc:
        raise LJError(rc, "pair-list cache %s is broken" % path)
    return nop, ptr, lst


def savepair_dat(path: str, number_of_partners, i_particles, j_particles) -> None:
    """cpu_ref's binary pair.dat (savepair(), cpu_ref/force_soa.cpp:369-377)."""
    nop, ip, jp = (np.ascontiguousarray(x, np.int32) for x in (number_of_partners, i_particles, j_particles))
    rc = capi.load().lj_pairdat_write(path.encode(), REF_N_STATIC, REF_MAX_PAIRS, len(nop), len(ip),
                                      nop.ctypes.data, ip.ctypes.data, jp.ctypes.data)
    if rc:
        raise LJError(rc, "lj_pairdat_write(%s)" % path)


def loadpair_dat(path: str, particle_number: int):
    """cpu_ref's loadpair() (cpu_ref/force_soa.cpp:360-367) -> (nop, i_particles, j_particles)."""
    lib = capi.load()
    npairs = C.c_int64(0)
    rc = lib.lj_pairdat_read(path.encode(), REF_N_STATIC, REF_MAX_PAIRS, particle_number, C.byref(npairs),
                             None, None, None, 0)
    if rc:
        raise LJError(rc, "pair.dat %s is missing or broken" % path)
    nop = np.empty(particle_number, np.int32)
    ip, jp = np.empty(npairs.value, np.int32), np.empty(npairs.value, np.int32)
    rc = lib.lj_pairdat_read(path.encode(), REF_N_STATIC, REF_MAX_PAIRS, particle_number, C.byref(npairs),
                             nop.ctypes.data, ip.ctypes.data, jp.ctypes.data, len(ip))
    if rc:
        raise LJError(rc, "pair.dat %s is broken" % path)
    return nop, ip, jp


@dataclass
class PairList:
    """The reference's three list arrays on the device (+ optional ELL table)."""
    number_of_partners: "torch.Tensor"   # int32[pn]
    pointer: "torch.Tensor"              # int32[pn] or int64[pn]
    sorted_list: "torch.Tensor"          # int32[capacity]
    number_of_pairs: int
    max_partners: int
    half: bool = False
    transposed_list: "torch.Tensor | None" = None
    sorted_list2d: "torch.Tensor | None" = None   # row-major padded table [pn, ell_width]
    ell_width: int = 0
    token: int = 0                        # lj_list_mirror_token() of the mirror built with / for these arrays
    build_flags: dict | None = None      # tiles / clusters / sort_rows / per_particle of the build: rebuild() reuses them

    @property
    def pointer64(self) -> bool:
        import torch
        return self.pointer.dtype == torch.int64


class LJContext:
    """One context per GPU (lj_ctx).  All calls are asynchronous on `stream`
    (default: torch's current stream) unless stated."""

    def __init__(self, device: int = 0):
        import torch
        if not torch.cuda.is_available():
            raise LJError(capi.LJ_ERR_NO_DEVICE, "no CUDA device; lj_gpu_b200 has no CPU fallback")
        self.lib = capi.load()
        self.device = device
        torch.cuda.set_device(device)
        torch.cuda.init()
        h = C.c_void_p()
        rc = self.lib.lj_ctx_create(C.byref(h), device)
        if rc:
            raise LJError(rc, "lj_ctx_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.lj_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int):
        if rc:
            raise LJError(rc, self.lib.lj_last_error_string(self.h).decode())

    def _stream(self, stream=None) -> C.c_void_p:
        import torch
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        return C.c_void_p(s.cuda_stream)

    @property
    def launches(self) -> int:
        return int(self.lib.lj_launch_count(self.h))

    def kernel_timing(self, enable: bool = True) -> None:
        """Start (with empty sums) / stop the live CUDA-event timing of the dominant force kernel alone."""
        self._check(self.lib.lj_kernel_timing(self.h, 1 if enable else 0))

    def kernel_timing_read(self):
        """-> (summed milliseconds, launches) of the dominant force kernel since kernel_timing(True)."""
        ms, n = C.c_double(0.0), C.c_int64(0)
        self._check(self.lib.lj_kernel_timing_read(self.h, C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)

    def sync(self, stream=None):
        self._check(self.lib.lj_sync(self.h, self._stream(stream)))

    @staticmethod
    def _layout_of(q, layout):
        if layout is not None:
            return LAYOUTS[layout] if isinstance(layout, str) else layout
        if q.dim() == 2 and q.shape[1] == 3:
            import torch
            return capi.LJ_AOS_F3 if q.dtype == torch.float32 else LJ_AOS_D3
        if q.dim() == 2 and q.shape[1] == 4:
            import torch
            return capi.LJ_AOS_F4 if q.dtype == torch.float32 else LJ_AOS_D4
        raise ValueError("cannot infer layout; pass layout='aos3'|'aos4'|'soa'")

    @staticmethod
    def _pn_stride(q, lay):
        if lay == LJ_SOA_D:  # [3(+), stride]
            return q.shape[1], q.stride(0)
        return q.shape[0], 0

    # ------------------------------------------------------------------ list build
    def makepair(self, q, search_len: float = SEARCH_LENGTH, half: bool = False, layout=None,
                 pointer64: bool = False, sort_rows: bool = False, capacity: int | None = None,
                 rows=None, pn=None, out: PairList | None = None, clusters: bool = False,
                 per_particle: bool = False, tiles: bool = False, stream=None) -> PairList:
        """makepair() (cuda/force_cuda.cu:122-163) on the GPU.  Returns device arrays.
        clusters=True also builds the library-owned cluster pair list (LJ_LIST_CLUSTERS) that the
        "auto"/"cluster" force variants use for exactly these arrays; tiles=True the cell-tile mirror
        (LJ_LIST_TILES) of the "auto"/"celltile" variants; tiles="wide" sizes its tiles for the
        mixed-precision kernel (LJ_LIST_TILES_WIDE)."""
        import torch
        lay = self._layout_of(q, layout)
        n, stride = self._pn_stride(q, lay)
        if pn is not None:
            n = pn
        dev = q.device
        if out is None:
            nop = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
            ptr = torch.empty(max(n, 1), dtype=torch.int64 if pointer64 else torch.int32, device=dev)
            lst = torch.empty(capacity if capacity is not None else 0, dtype=torch.int32, device=dev)
        else:
            nop, ptr, lst = out.number_of_partners, out.pointer, out.sorted_list
            pointer64 = out.pointer64
        a = capi.LjListArgs()
        a.q, a.pn, a.layout, a.half, a.plane_stride = q.data_ptr(), n, lay, int(half), stride
        a.search_len = search_len
        a.number_of_partners, a.pointer = nop.data_ptr(), ptr.data_ptr()
        a.pointer64 = int(pointer64)
        a.flags = (capi.LJ_LIST_SORT_ROWS if sort_rows else 0) | (capi.LJ_LIST_CLUSTERS if clusters else 0) | \
            (capi.LJ_LIST_PER_PARTICLE_SEARCH if per_particle else 0) | (capi.LJ_LIST_TILES if tiles else 0) | (capi.LJ_LIST_TILES_WIDE if tiles == "wide" else 0)
        if rows is not None:
            a.row_begin, a.row_end = rows
        total = C.c_int64(0)
        st = self._stream(stream)
        for _ in range(2):
            a.sorted_list, a.capacity = (lst.data_ptr() if lst.numel() else None), lst.numel()
            rc = self.lib.lj_build_list(self.h, C.byref(a), C.byref(total), st)
            if rc == capi.LJ_ERR_CAPACITY and out is None and capacity is None:
                lst = torch.empty(total.value + total.value // 64 + 1024, dtype=torch.int32, device=dev)
                continue
            self._check(rc)
            break
        mx = C.c_int32(0)
        self._check(self.lib.lj_list_result(self.h, C.byref(total), C.byref(mx), st))
        return PairList(nop, ptr, lst, int(total.value), int(mx.value), half, None, None, 0,
                        int(self.lib.lj_list_mirror_token(self.h)),
                        dict(tiles=tiles, clusters=clusters, sort_rows=sort_rows, per_particle=per_particle,
                             rows=rows))

    def rebuild(self, q, pl: PairList, search_len: float = SEARCH_LENGTH, layout=None,
                sort_rows=None, rows=None, pn=None, clusters=None,
                per_particle=None, tiles=None, stream=None):
        """Asynchronous rebuild into existing arrays (no host sync, no reallocation).  Flags left at
        None are the ones the list was built with (a mirror built by makepair(tiles=True) is rebuilt
        with the list instead of being silently dropped)."""
        bf = pl.build_flags or {}
        sort_rows = bf.get("sort_rows", False) if sort_rows is None else sort_rows
        clusters = bf.get("clusters", False) if clusters is None else clusters
        per_particle = bf.get("per_particle", False) if per_particle is None else per_particle
        tiles = bf.get("tiles", False) if tiles is None else tiles
        rows = bf.get("rows") if rows is None else rows
        pl.build_flags = dict(tiles=tiles, clusters=clusters, sort_rows=sort_rows, per_particle=per_particle,
                              rows=rows)
        lay = self._layout_of(q, layout)
        n, stride = self._pn_stride(q, lay)
        if pn is not None:
            n = pn
        a = capi.LjListArgs()
        a.q, a.pn, a.layout, a.half, a.plane_stride = q.data_ptr(), n, lay, int(pl.half), stride
        a.search_len = search_len
        a.number_of_partners, a.pointer = pl.number_of_partners.data_ptr(), pl.pointer.data_ptr()
        a.sorted_list, a.capacity = pl.sorted_list.data_ptr(), pl.sorted_list.numel()
        a.pointer64 = int(pl.pointer64)
        a.flags = (capi.LJ_LIST_SORT_ROWS if sort_rows else 0) | (capi.LJ_LIST_CLUSTERS if clusters else 0) | \
            (capi.LJ_LIST_PER_PARTICLE_SEARCH if per_particle else 0) | (capi.LJ_LIST_TILES if tiles else 0) | (capi.LJ_LIST_TILES_WIDE if tiles == "wide" else 0)
        if rows is not None:
            a.row_begin, a.row_end = rows
        self._check(self.lib.lj_build_list(self.h, C.byref(a), None, self._stream(stream)))
        pl.token = int(self.lib.lj_list_mirror_token(self.h))

    def list_mirror(self, q, pl: PairList, search_len: float = SEARCH_LENGTH, layout=None, pn=None,
                    wide: bool = False, stream=None) -> int:
        """lj_list_mirror(): the cell-tile mirror of a list the CALLER supplies (loaded from a pair
        cache, shuffled, built elsewhere).  Returns the number of rows the mirror could not take
        (they are served by the per-row kernel on the caller's arrays)."""
        lay = self._layout_of(q, layout)
        n, stride = self._pn_stride(q, lay)
        if pn is not None:
            n = pn
        outside = C.c_int64(0)
        self._check(self.lib.lj_list_mirror(self.h, q.data_ptr(), n, lay, stride, search_len,
                                            pl.number_of_partners.data_ptr(), pl.pointer.data_ptr(),
                                            int(pl.pointer64), pl.sorted_list.data_ptr(), pl.sorted_list.numel(),
                                            capi.LJ_LIST_TILES_WIDE if wide else 0, C.byref(outside),
                                            self._stream(stream)))
        pl.token = int(self.lib.lj_list_mirror_token(self.h))
        return int(outside.value)

    def list_invalidate(self):
        self._check(self.lib.lj_list_invalidate(self.h))

    def list_result(self, stream=None):
        total, mx = C.c_int64(0), C.c_int32(0)
        self._check(self.lib.lj_list_result(self.h, C.byref(total), C.byref(mx), self._stream(stream)))
        return int(total.value), int(mx.value)

    def make_transposed_pairlist(self, pl: PairList, stream=None):
        """make_transposed_pairlist() (cuda/force_cuda.cu:229-240) on the GPU."""
        import torch
        pn = pl.number_of_partners.numel()
        tl = torch.empty(max(1, pl.max_partners * pn), dtype=torch.int32,
                         device=pl.sorted_list.device)
        mx = C.c_int32(0)
        self._check(self.lib.lj_build_ell(self.h, pl.sorted_list.data_ptr(),
                                          pl.number_of_partners.data_ptr(), pl.pointer.data_ptr(),
                                          int(pl.pointer64), pn, tl.data_ptr(), tl.numel(),
                                          C.byref(mx), self._stream(stream)))
        pl.transposed_list = tl
        return tl

    def make_sorted_list2d(self, pl: PairList, width: int | None = None, stream=None):
        """make_sorted_list2d() (cuda/force_cuda.cu:242-253) done right: row-major padded table
        [pn, width], zero padded; width defaults to max_partners, a narrower one raises LJError
        (the reference's fixed NUM_NEIGH = 60 lets rows overlap)."""
        import torch
        pn = pl.number_of_partners.numel()
        width = pl.max_partners if width is None else width
        t2 = torch.empty(max(1, width * pn), dtype=torch.int32, device=pl.sorted_list.device)
        mx = C.c_int32(0)
        self._check(self.lib.lj_build_ell_rows(self.h, pl.sorted_list.data_ptr(),
                                               pl.number_of_partners.data_ptr(), pl.pointer.data_ptr(),
                                               int(pl.pointer64), pn, width, t2.data_ptr(), t2.numel(),
                                               C.byref(mx), self._stream(stream)))
        pl.sorted_list2d, pl.ell_width = t2, width
        return t2

    def random_shfl(self, pl: PairList, seed: int = 10, stream=None):
        """random_shfl() in spirit (cuda/force_cuda.cu:255-263): per-row device permutation."""
        self._check(self.lib.lj_shuffle_rows(self.h, pl.sorted_list.data_ptr(),
                                             pl.number_of_partners.data_ptr(), pl.pointer.data_ptr(),
                                             int(pl.pointer64), pl.number_of_partners.numel(), seed,
                                             self._stream(stream)))

    def check_loadedpair(self, pl: PairList, stream=None):
        """check_loadedpair() (cuda/force_cuda.cu:183-201) on the device; raises LJError."""
        self._check(self.lib.lj_validate_list(self.h, pl.sorted_list.data_ptr(),
                                              pl.number_of_partners.data_ptr(), pl.pointer.data_ptr(),
                                              int(pl.pointer64), pl.number_of_partners.numel(),
                                              pl.number_of_pairs, self._stream(stream)))

    # ------------------------------------------------------------------ force
    def force_args(self, q, p, pl: PairList, dt: float = DT, cl2: float = CL2, layout=None,
                   ell: bool = False, variant="auto", group: int = 0, precision: str = "fp64",
                   threads_per_block: int = 0, rows=None, pn=None, list_scalar: int = 0,
                   ell_rows: bool = False) -> capi.LjForceArgs:
        lay = self._layout_of(q, layout)
        n, stride = self._pn_stride(q, lay)
        if pn is not None:
            n = pn
        a = capi.LjForceArgs()
        a.q, a.p, a.pn, a.dt, a.cl2 = q.data_ptr(), p.data_ptr(), n, dt, cl2
        if ell_rows:
            if pl.sorted_list2d is None:
                raise ValueError("call make_sorted_list2d first")
            a.list, a.pointer, a.list_layout = pl.sorted_list2d.data_ptr(), None, capi.LJ_LIST_ELL_ROWS
            a.ell_width = pl.ell_width
        elif ell:
            if pl.transposed_list is None:
                raise ValueError("call make_transposed_pairlist first")
            a.list, a.pointer, a.list_layout = pl.transposed_list.data_ptr(), None, LJ_LIST_ELL
        else:
            a.list, a.pointer, a.list_layout = pl.sorted_list.data_ptr(), pl.pointer.data_ptr(), LJ_LIST_CSR
            a.list_entries = pl.sorted_list.numel()
        a.number_of_partners = pl.number_of_partners.data_ptr()
        a.layout = lay
        a.mirror_token = pl.token
        v = VARIANTS[variant] if isinstance(variant, str) else variant
        if pl.half:
            v = LJ_VARIANT_NEWTON3   # CSR: group lanes per i; half ELL table: thread per i (memopt2/3_with_aar)
        if variant == "warp" and group == 0:
            group = 32
        if variant == "thread" and group == 0:
            group = 1
        a.variant, a.group = v, group
        a.precision = {"fp64": LJ_PREC_FP64, "mixed": LJ_PREC_MIXED}[precision]
        a.pointer64 = int(pl.pointer64)
        a.threads_per_block = threads_per_block
        a.list_scalar = int(list_scalar)
        a.plane_stride = stride
        if rows is not None:
            a.row_begin, a.row_end = rows
        return a

    # ------------------------------------------------------------------ six-array SoA
    def makepair_soa6(self, qx, qy, qz, search_len: float = SEARCH_LENGTH, half: bool = False,
                      tiles=False, stream=None) -> PairList:
        """makepair() for the OpenACC SoA program's six separate arrays
        (openacc/force_oacc_soa.cpp:17-22, 101-141)."""
        import torch
        n, dev = qx.numel(), qx.device
        nop = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        ptr = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        lst = torch.empty(0, dtype=torch.int32, device=dev)
        a = capi.LjListArgs()
        a.pn, a.half, a.search_len = n, int(half), search_len
        a.number_of_partners, a.pointer = nop.data_ptr(), ptr.data_ptr()
        a.flags = (capi.LJ_LIST_TILES if tiles else 0) | (capi.LJ_LIST_TILES_WIDE if tiles == "wide" else 0)
        total = C.c_int64(0)
        st = self._stream(stream)
        for _ in range(2):
            a.sorted_list, a.capacity = (lst.data_ptr() if lst.numel() else None), lst.numel()
            rc = self.lib.lj_build_list_soa6(self.h, qx.data_ptr(), qy.data_ptr(), qz.data_ptr(), C.byref(a),
                                             C.byref(total), st)
            if rc == capi.LJ_ERR_CAPACITY and lst.numel() == 0:
                lst = torch.empty(total.value + total.value // 64 + 1024, dtype=torch.int32, device=dev)
                continue
            self._check(rc)
            break
        mx = C.c_int32(0)
        self._check(self.lib.lj_list_result(self.h, C.byref(total), C.byref(mx), st))
        pl = PairList(nop, ptr, lst, int(total.value), int(mx.value), half)
        pl.token = int(self.lib.lj_list_mirror_token(self.h))
        return pl

    def force_loop_soa6(self, qx, qy, qz, px, py, pz, pl: PairList, loop: int = LOOP, dt: float = DT,
                        cl2: float = CL2, ell: bool = False, variant="auto", group: int = 0,
                        precision: str = "fp64", stream=None):
        """force_reactless / force_reactless_memopt of the OpenACC SoA program
        (openacc/force_oacc_soa.cpp:203-263) on its six arrays, LOOP times (measure(), :279-284)."""
        a = capi.LjForceArgs()
        a.pn, a.dt, a.cl2 = qx.numel(), dt, cl2
        if ell:
            if pl.transposed_list is None:
                raise ValueError("call make_transposed_pairlist first")
            a.list, a.pointer, a.list_layout = pl.transposed_list.data_ptr(), None, LJ_LIST_ELL
        else:
            a.list, a.pointer, a.list_layout = pl.sorted_list.data_ptr(), pl.pointer.data_ptr(), LJ_LIST_CSR
            a.list_entries = pl.sorted_list.numel()
        a.number_of_partners = pl.number_of_partners.data_ptr()
        a.mirror_token = pl.token
        v = VARIANTS[variant] if isinstance(variant, str) else variant
        if pl.half and not ell:
            v = LJ_VARIANT_NEWTON3
        a.variant, a.group = v, group
        a.precision = {"fp64": LJ_PREC_FP64, "mixed": LJ_PREC_MIXED}[precision]
        a.pointer64 = int(pl.pointer64)
        self._check(self.lib.lj_force_loop_soa6(self.h, qx.data_ptr(), qy.data_ptr(), qz.data_ptr(),
                                                px.data_ptr(), py.data_ptr(), pz.data_ptr(), C.byref(a), loop,
                                                self._stream(stream)))

    def force_step(self, q, p, pl: PairList, stream=None, part=None, **kw):
        """One kernel launch of measure() (cuda/force_cuda.cu:334): p += dt * F(q), in place.
        part="interior" | "boundary": lj_force_step_part on the cell-tile mirror (the tiles that do /
        do not depend only on positions inside the list's row range)."""
        a = self.force_args(q, p, pl, **kw)
        if part is None:
            self._check(self.lib.lj_force_step(self.h, C.byref(a), self._stream(stream)))
        else:
            code = {"interior": capi.LJ_PART_INTERIOR, "boundary": capi.LJ_PART_BOUNDARY, "all": capi.LJ_PART_ALL}[part]
            self._check(self.lib.lj_force_step_part(self.h, C.byref(a), code, self._stream(stream)))

    def force_loop(self, q, p, pl: PairList, loop: int = LOOP, use_graph: bool = False, stream=None,
                   **kw):
        """The LOOP x launch body of measure() (cuda/force_cuda.cu:333-335)."""
        a = self.force_args(q, p, pl, **kw)
        self._check(self.lib.lj_force_loop(self.h, C.byref(a), loop, int(use_graph),
                                           self._stream(stream)))

    # ------------------------------------------------------------------ MD step
    def drift(self, q, p, dt: float = DT, layout=None, pn=None, stream=None):
        """q += p*dt (unit mass): the drift of a symplectic Euler step whose kick is force_step."""
        lay = self._layout_of(q, layout)
        n, stride = self._pn_stride(q, lay)
        self._check(self.lib.lj_drift(self.h, q.data_ptr(), p.data_ptr(), n if pn is None else pn, lay,
                                      stride, dt, self._stream(stream)))

    def max_displacement2(self, q, q_ref, layout=None, pn=None, stream=None) -> float:
        lay = self._layout_of(q, layout)
        n, stride = self._pn_stride(q, lay)
        out = C.c_double(0.0)
        self._check(self.lib.lj_max_displacement2(self.h, q.data_ptr(), q_ref.data_ptr(),
                                                  n if pn is None else pn, lay, stride, C.byref(out),
                                                  self._stream(stream)))
        return out.value

    def energy(self, q, p, pl: PairList, stream=None, **kw):
        """-> (kinetic, potential) of the listed pairs within the cutoff."""
        a = self.force_args(q, p, pl, **kw)
        ke, pe = C.c_double(0.0), C.c_double(0.0)
        self._check(self.lib.lj_energy(self.h, C.byref(a), C.byref(ke), C.byref(pe), self._stream(stream)))
        return ke.value, pe.value

    def md_run(self, q, p, pl: PairList, steps: int, dt: float = DT, search_len: float = SEARCH_LENGTH,
               cutoff: float = CUTOFF_LENGTH, check_every: int = 1, **fkw):
        """`steps` symplectic-Euler steps (kick = force_step, drift) with a skin-triggered list
        rebuild: the list is rebuilt when a particle has moved more than (search - cutoff)/2 since
        the last build.  Returns the number of rebuilds."""
        q_ref = q.clone()
        limit2 = (0.5 * (search_len - cutoff)) ** 2
        rebuilds = 0
        for s in range(steps):
            self.force_step(q, p, pl, dt=dt, cl2=cutoff * cutoff, **fkw)
            self.drift(q, p, dt=dt, layout=fkw.get("layout"), pn=fkw.get("pn"))
            if (s + 1) % check_every == 0 and \
                    self.max_displacement2(q, q_ref, layout=fkw.get("layout"), pn=fkw.get("pn")) > limit2:
                self.rebuild(q, pl, search_len=search_len, layout=fkw.get("layout"), pn=fkw.get("pn"))
                pl.number_of_pairs, pl.max_partners = self.list_result()  # raises on capacity overflow
                q_ref.copy_(q)
                rebuilds += 1
        return rebuilds

    # ------------------------------------------------------------------ measure()
    def measure(self, q_host: np.ndarray, p_host: np.ndarray, layout=None, loop: int = LOOP,
                rebuild_every: int = 0, half: bool = False, variant="auto", group: int = 0,
                precision: str = "fp64", threads_per_block: int = 0, use_graph: bool = False,
                dt: float = DT, cl2: float = CL2, search_len: float = SEARCH_LENGTH,
                host_list=None, sort_rows: bool = False) -> capi.LjMeasureArgs:
        """measure() (cuda/force_cuda.cu:319-342) on HOST arrays, p_host updated in place.
        host_list = (number_of_partners, pointer, sorted_list) int32 numpy arrays to upload the
        caller's list like the reference does; None builds the list on the GPU."""
        assert q_host.dtype == np.float64 and p_host.dtype == np.float64
        assert q_host.flags.c_contiguous and p_host.flags.c_contiguous
        if layout is None:
            layout = {3: "aos3", 4: "aos4"}[q_host.shape[1]]
        lay = LAYOUTS[layout]
        m = capi.LjMeasureArgs()
        m.q_host, m.p_host = q_host.ctypes.data, p_host.ctypes.data
        if lay == LJ_SOA_D:
            m.pn, m.plane_stride = q_host.shape[1], q_host.shape[1]
        else:
            m.pn = q_host.shape[0]
        m.layout, m.half = lay, int(half)
        m.dt, m.cl2, m.search_len = dt, cl2, search_len
        m.loop, m.rebuild_every = loop, rebuild_every
        m.variant = VARIANTS[variant] if isinstance(variant, str) else variant
        if variant == "warp" and group == 0:
            group = 32
        if variant == "thread" and group == 0:
            group = 1
        m.group = group
        m.precision = {"fp64": LJ_PREC_FP64, "mixed": LJ_PREC_MIXED}[precision]
        m.threads_per_block, m.use_graph = threads_per_block, int(use_graph)
        m.list_flags = capi.LJ_LIST_SORT_ROWS if sort_rows else 0
        keep = None
        if host_list is not None:
            nop, ptr, lst = (np.ascontiguousarray(x, np.int32) for x in host_list)
            keep = (nop, ptr, lst)
            m.number_of_partners_host, m.pointer_host, m.list_host = \
                nop.ctypes.data, ptr.ctypes.data, lst.ctypes.data
            m.number_of_pairs_in = len(lst)
        self._check(self.lib.lj_measure(self.h, C.byref(m)))
        del keep
        return m

    # ------------------------------------------------------------------ cuda_ptr
    def cuda_ptr(self, dtype, count: int) -> "CudaPtr":
        return CudaPtr(self, np.dtype(dtype), count)

    # ------------------------------------------------------------------ multi-GPU helpers
    def ipc_tensor(self, shape, dtype):
        """A torch tensor on IPC-shareable memory (lj_ipc_alloc = plain cudaMalloc)."""
        import torch
        n = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        ptr = C.c_void_p()
        self._check(self.lib.lj_ipc_alloc(self.h, n, C.byref(ptr)))

        class _Mem:  # torch.as_tensor consumes the CUDA array interface
            __cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr.value, False),
                                        "version": 2}
        t = torch.as_tensor(_Mem(), device="cuda:%d" % self.device).view(dtype).view(*shape)
        t._lj_ipc_base = ptr.value
        return t

    def ipc_export(self, tensor) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self.lib.lj_ipc_export(self.h, tensor.data_ptr(), buf))
        return buf.raw

    def ipc_open(self, handle: bytes) -> int:
        out = C.c_void_p()
        self._check(self.lib.lj_ipc_open(self.h, handle, C.byref(out)))
        return out.value

    def ipc_close(self, ptr: int):
        self._check(self.lib.lj_ipc_close(self.h, ptr))

    def flag_set(self, flag_ptr: int, value: int, stream=None):
        self._check(self.lib.lj_flag_set(self.h, flag_ptr, value, self._stream(stream)))

    def flag_wait(self, flag_ptr: int, at_least: int, stream=None):
        self._check(self.lib.lj_flag_wait(self.h, flag_ptr, at_least, self._stream(stream)))

    def halo_pull_sync(self, segs, stream=None):
        """segs: [(dst_ptr, src_ptr, nbytes, wait_flag_ptr|0, wait_value, done_flag_ptr|0, done_value)], one or two."""
        arr = (capi.LjHaloSeg * len(segs))()
        for k, (d, sp, n, wf, wv, df, dv) in enumerate(segs):
            arr[k].local_dst, arr[k].peer_src, arr[k].bytes = d, sp, n
            arr[k].wait_flag, arr[k].wait_value = (wf or None), wv
            arr[k].done_flag, arr[k].done_value = (df or None), dv
        self._check(self.lib.lj_halo_pull_sync(self.h, arr, len(segs), self._stream(stream)))

    def halo_pull(self, dst_ptr: int, peer_ptr: int, nbytes: int, stream=None):
        self._check(self.lib.lj_halo_pull(self.h, dst_ptr, peer_ptr, nbytes, self._stream(stream)))


class CudaPtr:
    """cuda_ptr<T> (cuda/cuda_ptr.cuh:10-104): paired pinned-host / device buffer with the same
    method set (allocate in the constructor, host2dev, dev2host, set_val, operator[] -> host
    view, dev_ptr, deallocate), but stream-ordered and sized by the request."""

    def __init__(self, ctx: LJContext, dtype: np.dtype, count: int):
        self.ctx, self.dtype, self.size = ctx, dtype, count
        self.buf = capi.LjBuf()
        ctx._check(ctx.lib.lj_buf_allocate(ctx.h, count * dtype.itemsize, C.byref(self.buf),
                                           ctx._stream()))
        arr_t = (C.c_char * (count * dtype.itemsize))
        self.host = np.frombuffer(arr_t.from_address(self.buf.host), dtype=dtype, count=count) \
            if count else np.empty(0, dtype)

    @property
    def dev_ptr(self) -> int:
        return self.buf.dev or 0

    def __getitem__(self, i):
        return self.host[i]

    def __setitem__(self, i, v):
        self.host[i] = v

    def host2dev(self, beg: int = 0, count: int | None = None, stream=None):
        count = self.size - beg if count is None else count
        it = self.dtype.itemsize
        self.ctx._check(self.ctx.lib.lj_buf_host2dev(self.ctx.h, C.byref(self.buf), beg * it,
                                                     count * it, self.ctx._stream(stream)))

    def dev2host(self, beg: int = 0, count: int | None = None, stream=None):
        count = self.size - beg if count is None else count
        it = self.dtype.itemsize
        self.ctx._check(self.ctx.lib.lj_buf_dev2host(self.ctx.h, C.byref(self.buf), beg * it,
                                                     count * it, self.ctx._stream(stream)))

    def set_val(self, val, beg: int = 0, count: int | None = None, stream=None):
        assert self.dtype.itemsize == 4, "set_val is for 4-byte element types"
        count = self.size - beg if count is None else count
        bits = int(np.array([val], self.dtype).view(np.uint32)[0])
        self.ctx._check(self.ctx.lib.lj_buf_set_val32(self.ctx.h, C.byref(self.buf), beg, count, bits,
                                                      self.ctx._stream(stream)))

    def deallocate(self):
        if self.buf.dev or self.buf.host:
            self.host = None
            self.ctx._check(self.ctx.lib.lj_buf_deallocate(self.ctx.h, C.byref(self.buf),
                                                           self.ctx._stream()))
