"""ctypes binding of include/recfilter_b200.h (the C ABI drop-in boundary).

Mirrors the reference's operator surface at the level a host language needs:
``Plan(extents, dtype, scans, border)`` is RecFilter::define + add_filter + split
(/root/reference/lib/recfilter.cpp:192-392, lib/split.cpp:1850-2080), ``Plan.execute``
is realize on device buffers (lib/recfilter.cpp:984-989), ``Plan.realize`` the
host-buffer realize, ``Plan.profile`` RecFilter::profile (lib/recfilter.cpp:991-1016).

No CPU fallback lives here: missing library or device => RecFilterError.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from dataclasses import dataclass
from typing import Iterable, Sequence

import numpy as np

RF_MAX_DIMS = 4
RF_MAX_SCANS = 64
RF_MAX_ORDER = 32

DTYPES = {
    "f32": (0, np.float32), "i32": (2, np.int32), "u32": (3, np.uint32),
    "i16": (4, np.int16), "u16": (5, np.uint16), "i8": (6, np.int8), "u8": (7, np.uint8),
}
ENGINES = {"auto": 0, "generic": 1, "fused": 2, "twopass": 3}
_NP_TO_NAME = {np.dtype(v[1]): k for k, v in DTYPES.items()}


class RecFilterError(RuntimeError):
    pass


class _Scan(C.Structure):
    _fields_ = [("dim", C.c_int32), ("causal", C.c_int32), ("order", C.c_int32),
                ("coeff", C.c_float * (RF_MAX_ORDER + 1))]


class _Options(C.Structure):
    _fields_ = [("tile", C.c_int32 * RF_MAX_DIMS), ("honor_tile", C.c_int32), ("fuse_dims", C.c_int32),
                ("open_lo", C.c_int32), ("open_hi", C.c_int32), ("shard_dim", C.c_int32),
                ("engine", C.c_int32), ("epilogue", C.c_int32), ("epi_in", C.c_float), ("epi_out", C.c_float),
                ("reserved", C.c_int32 * 4)]


class _Desc(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("extent", C.c_int64 * RF_MAX_DIMS), ("dtype", C.c_int32),
                ("border", C.c_int32), ("nscans", C.c_int32), ("scans", _Scan * RF_MAX_SCANS),
                ("opt", _Options)]


@dataclass
class Scan:
    """One add_filter call: ``Scan(dim, causal, [b0, a1..ar])`` (lib/recfilter.cpp:264-287)."""
    dim: int
    causal: bool
    coeff: Sequence[float]


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "librecfilter_b200.so")


_lib = None


def lib():
    """Load librecfilter_b200.so; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RecFilterError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    L = C.CDLL(path)
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    L.rf_version.restype = C.c_char_p
    L.rf_last_error.restype = C.c_char_p
    L.rf_device_count.restype = i32
    L.rf_set_device.argtypes = [i32]
    L.rf_plan_create.argtypes = [C.POINTER(_Desc), C.POINTER(vp)]
    L.rf_plan_destroy.argtypes = [vp]
    L.rf_plan_destroy.restype = None
    L.rf_plan_workspace_bytes.argtypes = [vp]
    L.rf_plan_workspace_bytes.restype = sz
    L.rf_plan_num_launches.argtypes = [vp]
    L.rf_plan_describe.argtypes = [vp, C.c_char_p, sz]
    L.rf_plan_execute.argtypes = [vp, vp, vp, vp]
    L.rf_plan_check.argtypes = [vp]
    L.rf_plan_execute_host.argtypes = [vp, vp, vp]
    L.rf_plan_execute_host_batch.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp)]
    L.rf_plan_profile.argtypes = [vp, vp, vp, i32, C.POINTER(C.c_float)]
    L.rf_stencil_execute.argtypes = [i32, C.POINTER(C.c_int64), i32, i32, vp, C.c_float, vp, vp, vp, vp]
    L.rf_plan_shard_tail_bytes.argtypes = [vp]
    L.rf_plan_shard_tail_bytes.restype = sz
    L.rf_plan_stage1.argtypes = [vp, vp, vp, vp, vp]
    L.rf_plan_stage2.argtypes = [vp, vp, vp, vp, i32, i32, vp]
    L.rf_plan_shard_vectors.argtypes = [vp]
    L.rf_plan_shard_neighbors_suffice.argtypes = [vp]
    L.rf_plan_shard_resolve_lines.argtypes = [vp, vp, i32, C.c_int64, vp, vp]
    L.rf_plan_stage2_ext.argtypes = [vp, vp, vp, vp, vp]
    L.rf_plan_stage_timing.argtypes = [vp, i32]
    L.rf_plan_stage_times.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_long), i32]
    L.rf_clock_begin.argtypes = [vp, C.POINTER(vp)]
    L.rf_clock_end.argtypes = [vp, vp, C.POINTER(C.c_float)]
    L.rf_mgpu_create.argtypes = [C.POINTER(_Desc), i32, C.POINTER(vp)]
    L.rf_mgpu_destroy.argtypes = [vp]
    L.rf_mgpu_destroy.restype = None
    L.rf_mgpu_ngpus.argtypes = [vp]
    L.rf_mgpu_describe.argtypes = [vp, C.c_char_p, sz]
    L.rf_mgpu_execute_host.argtypes = [vp, vp, vp]
    L.rf_mgpu_profile.argtypes = [vp, vp, i32, C.POINTER(C.c_float)]
    L.rf_mgpu_last_error.restype = C.c_char_p
    L.rf_xchg_create.argtypes = [sz, i32, i32, C.POINTER(vp)]
    L.rf_xchg_destroy.argtypes = [vp]
    L.rf_xchg_destroy.restype = None
    L.rf_xchg_handle_bytes.restype = sz
    L.rf_xchg_ipc_handle.argtypes = [vp, vp]
    L.rf_xchg_open_peer.argtypes = [vp, i32, vp]
    L.rf_xchg_set_peer.argtypes = [vp, i32, vp]
    L.rf_xchg_put.argtypes = [vp, vp, sz, vp]
    L.rf_xchg_wait.argtypes = [vp, vp, C.POINTER(vp)]
    L.rf_xchg_put_part.argtypes = [vp, C.c_uint, vp, sz, sz, i32, vp]
    L.rf_xchg_wait_from.argtypes = [vp, C.c_uint, vp, C.POINTER(vp)]
    L.rf_xchg_check.argtypes = [vp]
    L.rf_xchg_last_error.restype = C.c_char_p
    L.rf_malloc.argtypes = [C.POINTER(vp), sz]
    L.rf_free.argtypes = [vp]
    L.rf_memcpy_h2d.argtypes = [vp, vp, sz]
    L.rf_memcpy_d2h.argtypes = [vp, vp, sz]
    L.rf_memset.argtypes = [vp, i32, sz]
    L.rf_malloc_host.argtypes = [C.POINTER(vp), sz]
    L.rf_free_host.argtypes = [vp]
    _lib = L
    return L


def exported_symbols() -> list[str]:
    """Every function name declared in include/recfilter_b200.h."""
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "recfilter_b200.h")
    text = open(hdr).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rf_[a-z0-9_]+)\s*\(", text)))


def _check(rc: int, what: str):
    if rc != 0:
        raise RecFilterError(f"{what} failed ({rc}): {lib().rf_last_error().decode()}")


def device_count() -> int:
    return int(lib().rf_device_count())


def _dtype_name(dtype) -> str:
    if isinstance(dtype, str):
        if dtype not in DTYPES:
            raise RecFilterError(f"unsupported dtype {dtype}")
        return dtype
    d = np.dtype(dtype)
    if d not in _NP_TO_NAME:
        raise RecFilterError(f"unsupported dtype {d}")
    return _NP_TO_NAME[d]


class Plan:
    """A compiled filter: scan list + extents -> launch plan on the current CUDA device.

    ``extents`` are given in the reference's order, dimension 0 (x) first and
    contiguous; numpy / torch arrays therefore have shape ``extents[::-1]``.
    """

    def __init__(self, extents: Sequence[int], dtype, scans: Iterable[Scan], border: str = "zero", *,
                 tile: Sequence[int] | int | None = None, honor_tile: bool = False, fuse_dims: int = -1,
                 shard_dim: int = -1, open_lo: bool = False, open_hi: bool = False, engine: str = "auto",
                 epilogue: tuple[float, float] | None = None):
        """``epilogue=(a_in, a_out)``: the result is ``a_out * filtered + a_in * input`` (the unsharp mask of
        apps/usm/unsharp_mask_optimized.cpp:61-66), fused into the filter's last store."""
        self._h = C.c_void_p()
        L = lib()
        d = self._describe(extents, dtype, scans, border, tile, honor_tile, fuse_dims, shard_dim, open_lo, open_hi, engine)
        if epilogue is not None:
            d.opt.epilogue, d.opt.epi_in, d.opt.epi_out = 1, float(epilogue[0]), float(epilogue[1])
        self.size = int(np.prod(self.extents)) if self.extents else 0
        _check(L.rf_plan_create(C.byref(d), C.byref(self._h)), "rf_plan_create")

    def _describe(self, extents, dtype, scans, border, tile, honor_tile, fuse_dims, shard_dim, open_lo, open_hi, engine):
        """Fill an rf_desc (shared with MultiGpuPlan)."""
        self.extents = tuple(int(e) for e in extents)
        self.dtype_name = _dtype_name(dtype)
        self.np_dtype = np.dtype(DTYPES[self.dtype_name][1])
        scans = list(scans)
        if len(self.extents) > RF_MAX_DIMS or len(self.extents) < 1:
            raise RecFilterError("1..4 dimensions supported")
        if len(scans) > RF_MAX_SCANS:
            raise RecFilterError(f"at most {RF_MAX_SCANS} scans")
        d = _Desc()
        d.ndim = len(self.extents)
        for i, e in enumerate(self.extents):
            d.extent[i] = e
        d.dtype = DTYPES[self.dtype_name][0]
        if border not in ("zero", "clamp"):
            raise RecFilterError("border must be 'zero' or 'clamp'")
        d.border = 1 if border == "clamp" else 0
        d.nscans = len(scans)
        for i, s in enumerate(scans):
            coeff = [float(c) for c in s.coeff]
            # lib/recfilter.cpp:274-278: needs a feed-forward and at least one feedback coefficient
            if len(coeff) < 2:
                raise RecFilterError("Cannot add scan without feed forward and feedback coefficients")
            if len(coeff) - 1 > RF_MAX_ORDER:
                raise RecFilterError(f"filter order above {RF_MAX_ORDER} is not supported")
            d.scans[i].dim = int(s.dim)
            d.scans[i].causal = 1 if s.causal else 0
            d.scans[i].order = len(coeff) - 1
            for k, c in enumerate(coeff):
                d.scans[i].coeff[k] = c
        if tile is not None:
            tl = [tile] * d.ndim if isinstance(tile, int) else list(tile)
            for i, t in enumerate(tl):
                d.opt.tile[i] = int(t)
        d.opt.honor_tile = 1 if honor_tile else 0
        d.opt.fuse_dims = int(fuse_dims)
        d.opt.shard_dim = int(shard_dim)
        if engine not in ENGINES:
            raise RecFilterError(f"engine must be one of {sorted(ENGINES)}")
        d.opt.engine = ENGINES[engine]
        d.opt.open_lo = 1 if open_lo else 0
        d.opt.open_hi = 1 if open_hi else 0
        return d

    # -- life cycle -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().rf_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- introspection ----------------------------------------------------------------------
    @property
    def workspace_bytes(self) -> int:
        return int(lib().rf_plan_workspace_bytes(self._h))

    @property
    def num_launches(self) -> int:
        return int(lib().rf_plan_num_launches(self._h))

    def describe(self) -> str:
        buf = C.create_string_buffer(8192)
        _check(lib().rf_plan_describe(self._h, buf, len(buf)), "rf_plan_describe")
        return buf.value.decode()

    @property
    def shard_tail_bytes(self) -> int:
        return int(lib().rf_plan_shard_tail_bytes(self._h))

    # -- execution --------------------------------------------------------------------------
    def execute_ptr(self, in_ptr: int, out_ptr: int, stream: int = 0):
        _check(lib().rf_plan_execute(self._h, C.c_void_p(in_ptr), C.c_void_p(out_ptr), C.c_void_p(stream)),
               "rf_plan_execute")

    def execute(self, src, dst=None, stream=None):
        """Run on torch CUDA tensors (device resident)."""
        import torch
        if dst is None:
            dst = torch.empty_like(src)
        self._check_tensor(src), self._check_tensor(dst)
        st = torch.cuda.current_stream().cuda_stream if stream is None else int(stream)
        self.execute_ptr(src.data_ptr(), dst.data_ptr(), st)
        return dst

    def check(self):
        """Synchronise and raise if a kernel of the plan flagged an internal failure (rf_plan_check)."""
        _check(lib().rf_plan_check(self._h), "rf_plan_check")

    def _check_tensor(self, t):
        if not t.is_cuda or not t.is_contiguous():
            raise RecFilterError("tensors must be contiguous CUDA tensors")
        if t.numel() != self.size or t.element_size() != self.np_dtype.itemsize:
            raise RecFilterError("tensor size / element size does not match the plan")

    def realize(self, array: np.ndarray) -> np.ndarray:
        """Host buffers in, host buffers out (H2D + kernels + D2H): RecFilter::realize."""
        a = np.ascontiguousarray(array, dtype=self.np_dtype)
        if a.size != self.size:
            raise RecFilterError(f"array has {a.size} samples, plan expects {self.size}")
        out = np.empty_like(a)
        _check(lib().rf_plan_execute_host(self._h, a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)),
               "rf_plan_execute_host")
        return out

    def realize_ptr(self, in_host_ptr: int, out_host_ptr: int):
        _check(lib().rf_plan_execute_host(self._h, C.c_void_p(in_host_ptr), C.c_void_p(out_host_ptr)),
               "rf_plan_execute_host")

    def realize_batch_ptr(self, in_host_ptrs, out_host_ptrs):
        """n images in host memory (pinned for overlap): pipelined H2D / filter / D2H, rf_plan_execute_host_batch."""
        n = len(in_host_ptrs)
        if n != len(out_host_ptrs):
            raise RecFilterError("realize_batch_ptr: input and output lists differ in length")
        ins = (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in in_host_ptrs])
        outs = (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in out_host_ptrs])
        _check(lib().rf_plan_execute_host_batch(self._h, n, ins, outs), "rf_plan_execute_host_batch")

    def realize_batch(self, arrays):
        """List of host arrays in, list of host arrays out (see realize_batch_ptr)."""
        ins = [np.ascontiguousarray(a, dtype=self.np_dtype) for a in arrays]
        for a in ins:
            if a.size != self.size:
                raise RecFilterError(f"array has {a.size} samples, plan expects {self.size}")
        outs = [np.empty_like(a) for a in ins]
        self.realize_batch_ptr([a.ctypes.data for a in ins], [o.ctypes.data for o in outs])
        return outs

    def profile(self, src, dst, iters: int) -> float:
        ms = C.c_float()
        _check(lib().rf_plan_profile(self._h, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), int(iters),
                                     C.byref(ms)), "rf_plan_profile")
        return float(ms.value)

    STAGES = ("tile_tails", "carry_chain", "cross_residual", "tile_final", "convert")

    def stage_timing(self, enable: bool):
        _check(lib().rf_plan_stage_timing(self._h, 1 if enable else 0), "rf_plan_stage_timing")

    def stage_times(self) -> dict:
        ms = (C.c_double * 5)()
        cnt = (C.c_long * 5)()
        _check(lib().rf_plan_stage_times(self._h, ms, cnt, 5), "rf_plan_stage_times")
        return {n: {"ms": float(ms[i]), "launches": int(cnt[i])} for i, n in enumerate(self.STAGES)}

    # -- sharded execution ------------------------------------------------------------------
    def stage1(self, src, dst, tails, stream=None):
        import torch
        st = torch.cuda.current_stream().cuda_stream if stream is None else int(stream)
        _check(lib().rf_plan_stage1(self._h, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()),
                                    C.c_void_p(tails.data_ptr()), C.c_void_p(st)), "rf_plan_stage1")

    @property
    def shard_vectors(self) -> int:
        return int(lib().rf_plan_shard_vectors(self._h))

    @property
    def shard_neighbors_suffice(self) -> bool:
        """True when the carries entering this strip follow from the tails of the two adjacent strips alone
        (rf_plan_shard_neighbors_suffice: whole-strip transition matrices below 1e-30)."""
        return bool(lib().rf_plan_shard_neighbors_suffice(self._h))

    def shard_resolve_lines(self, gathered, nshards: int, nlines: int, ext_all, stream=None):
        import torch
        st = torch.cuda.current_stream().cuda_stream if stream is None else int(stream)
        _check(lib().rf_plan_shard_resolve_lines(self._h, C.c_void_p(gathered.data_ptr()), int(nshards), int(nlines),
                                                 C.c_void_p(ext_all.data_ptr()), C.c_void_p(st)), "rf_plan_shard_resolve_lines")

    def stage2_ext(self, src, dst, ext, stream=None):
        import torch
        st = torch.cuda.current_stream().cuda_stream if stream is None else int(stream)
        _check(lib().rf_plan_stage2_ext(self._h, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()),
                                        C.c_void_p(ext.data_ptr()), C.c_void_p(st)), "rf_plan_stage2_ext")

    def stage2(self, src, dst, gathered, nshards: int, rank: int, stream=None):
        """gathered: a tensor, or the device pointer of the gathered tails (Exchange.wait)."""
        import torch
        st = torch.cuda.current_stream().cuda_stream if stream is None else int(stream)
        gp = gathered if isinstance(gathered, int) else gathered.data_ptr()
        _check(lib().rf_plan_stage2(self._h, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()),
                                    C.c_void_p(gp), int(nshards), int(rank), C.c_void_p(st)),
               "rf_plan_stage2")


class MultiGpuPlan(Plan):
    """One filter over `ngpus` GPUs of this process (rf_mgpu_*): the outermost dimension is cut into independent parts
    (no scans along it) or strips (one order-r tail exchange over NVLink per call).  Host arrays in, host arrays out."""

    def __init__(self, extents, dtype, scans, border: str = "zero", *, ngpus: int):
        self._h = C.c_void_p()              # (no single-GPU plan)
        self._m = C.c_void_p()
        d = self._describe(extents, dtype, scans, border, None, False, -1, -1, False, False, "auto")
        self.size = int(np.prod(self.extents)) if self.extents else 0
        rc = lib().rf_mgpu_create(C.byref(d), int(ngpus), C.byref(self._m))
        if rc != 0:
            raise RecFilterError(f"rf_mgpu_create failed ({rc}): {lib().rf_mgpu_last_error().decode()}")

    def describe(self) -> str:
        buf = C.create_string_buffer(8192)
        lib().rf_mgpu_describe(self._m, buf, len(buf))
        return buf.value.decode()

    def realize(self, array: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(array, dtype=self.np_dtype)
        if a.size != self.size:
            raise RecFilterError(f"array has {a.size} samples, plan expects {self.size}")
        out = np.empty_like(a)
        rc = lib().rf_mgpu_execute_host(self._m, a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RecFilterError(f"rf_mgpu_execute_host failed ({rc}): {lib().rf_mgpu_last_error().decode()}")
        return out

    def profile_host(self, array: np.ndarray, iters: int) -> float:
        a = np.ascontiguousarray(array, dtype=self.np_dtype)
        ms = C.c_float()
        rc = lib().rf_mgpu_profile(self._m, a.ctypes.data_as(C.c_void_p), int(iters), C.byref(ms))
        if rc != 0:
            raise RecFilterError(f"rf_mgpu_profile failed ({rc}): {lib().rf_mgpu_last_error().decode()}")
        return float(ms.value)

    def close(self):
        if getattr(self, "_m", None) is not None and self._m.value:
            lib().rf_mgpu_destroy(self._m)
            self._m = C.c_void_p()


class Exchange:
    """Peer-to-peer exchange window for strip tails (rf_xchg_*): one per rank, on the current CUDA device."""

    def __init__(self, bytes_per_rank: int, nranks: int, rank: int):
        self._h = C.c_void_p()
        self.nranks, self.rank = nranks, rank
        self._xcheck(lib().rf_xchg_create(int(bytes_per_rank), int(nranks), int(rank), C.byref(self._h)), "rf_xchg_create")

    @staticmethod
    def _xcheck(rc, what):
        if rc != 0:
            raise RecFilterError(f"{what} failed ({rc}): {lib().rf_xchg_last_error().decode()}")

    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(int(lib().rf_xchg_handle_bytes()))
        self._xcheck(lib().rf_xchg_ipc_handle(self._h, buf), "rf_xchg_ipc_handle")
        return buf.raw

    def open_peer(self, peer: int, handle: bytes):
        self._xcheck(lib().rf_xchg_open_peer(self._h, int(peer), C.create_string_buffer(handle, len(handle))), "rf_xchg_open_peer")

    def set_peer(self, peer: int, other: "Exchange"):
        self._xcheck(lib().rf_xchg_set_peer(self._h, int(peer), other._h), "rf_xchg_set_peer")

    def connect(self, group=None):
        """One process per GPU: pass the IPC handles around once (torch.distributed object all-gather)."""
        import torch.distributed as dist
        handles = [None] * self.nranks
        dist.all_gather_object(handles, self.ipc_handle(), group=group)
        for p, h in enumerate(handles):
            if p != self.rank:
                self.open_peer(p, h)
        dist.barrier(group=group)

    def put(self, tails, stream=None):
        import torch
        st = torch.cuda.current_stream().cuda_stream if stream is None else int(stream)
        self._xcheck(lib().rf_xchg_put(self._h, C.c_void_p(tails.data_ptr()), tails.numel() * tails.element_size(), C.c_void_p(st)),
                     "rf_xchg_put")

    def wait(self, stream=None) -> int:
        import torch
        st = torch.cuda.current_stream().cuda_stream if stream is None else int(stream)
        out = C.c_void_p()
        self._xcheck(lib().rf_xchg_wait(self._h, C.c_void_p(st), C.byref(out)), "rf_xchg_wait")
        return int(out.value)

    def put_part(self, peer_mask: int, tails, offset: int, nbytes: int, last: bool, stream=None):
        """One part of a step: bytes [offset, offset + nbytes) of `tails` into this rank's slot of the windows of the
        ranks in `peer_mask`; the part with last=True raises the arrival words (rf_xchg_put_part)."""
        import torch
        st = torch.cuda.current_stream().cuda_stream if stream is None else int(stream)
        self._xcheck(lib().rf_xchg_put_part(self._h, C.c_uint(peer_mask), C.c_void_p(tails.data_ptr() if tails is not None else None),
                                            C.c_size_t(offset), C.c_size_t(nbytes), 1 if last else 0, C.c_void_p(st)), "rf_xchg_put_part")

    def wait_from(self, from_mask: int, stream=None) -> int:
        import torch
        st = torch.cuda.current_stream().cuda_stream if stream is None else int(stream)
        out = C.c_void_p()
        self._xcheck(lib().rf_xchg_wait_from(self._h, C.c_uint(from_mask), C.c_void_p(st), C.byref(out)), "rf_xchg_wait_from")
        return int(out.value)

    def check(self):
        self._xcheck(lib().rf_xchg_check(self._h), "rf_xchg_check")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().rf_xchg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Tap(C.Structure):
    _fields_ = [("weight", C.c_float), ("source", C.c_int32), ("offset", C.c_int32 * 4), ("lo", C.c_int32 * 4),
                ("hi", C.c_int32 * 4)]


def stencil(src, taps, post_scale: float = 1.0, dst=None, src2=None):
    """Pointwise linear stencil of a torch CUDA tensor (rf_stencil_execute): taps are
    (weight, offsets[, lo, hi]) with offsets / clamps per dimension, dimension 0 = contiguous.
    out(x) = post_scale * sum_i w_i * src(clamp(x + off_i, lo_i, hi_i))."""
    import torch
    if dst is None:
        dst = torch.empty_like(src)
    nd = src.dim()
    ext = (C.c_int64 * 4)(*([int(e) for e in reversed(src.shape)] + [1] * (4 - nd)))
    arr = (_Tap * len(taps))()
    for i, t in enumerate(taps):
        w, off = t[0], t[1]
        lo = t[2] if len(t) > 2 else [None] * nd
        hi = t[3] if len(t) > 3 else [None] * nd
        arr[i].weight = float(w)
        arr[i].source = int(t[4]) if len(t) > 4 else 0
        for d in range(4):
            arr[i].offset[d] = int(off[d]) if d < nd else 0
            arr[i].lo[d] = int(lo[d]) if d < nd and lo[d] is not None else -2**31
            arr[i].hi[d] = int(hi[d]) if d < nd and hi[d] is not None else 2**31 - 1
    dt = {torch.float32: 0, torch.int32: 2}[src.dtype]
    _check(lib().rf_stencil_execute(nd, ext, dt, len(taps), C.cast(arr, C.c_void_p), float(post_scale),
                                    C.c_void_p(src.data_ptr()), C.c_void_p(src2.data_ptr() if src2 is not None else None),
                                    C.c_void_p(dst.data_ptr()),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)), "rf_stencil_execute")
    return dst
