"""Context / SRS handles: the object-level view of the C ABI (one Context per GPU)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import field
from ._lib import GM_ERR_LENGTH, GeminiError, check, lib


def _ptr(a) -> C.c_void_p:
    """Host pointer of a C-contiguous numpy array / torch CPU tensor, or a raw int address."""
    if a is None:
        return C.c_void_p(None)
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):  # torch tensor (pinned host or device memory)
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return C.c_void_p(a.data_ptr())
    raise TypeError(f"cannot take the address of {type(a)!r}")


def as_fr_array(scalars, montgomery: bool = True) -> np.ndarray:
    """Accept (n,4) uint64 limb arrays as they are; convert sequences of Python ints."""
    if isinstance(scalars, np.ndarray) and scalars.dtype == np.uint64:
        return np.ascontiguousarray(scalars.reshape(-1, 4))
    return field.fr_to_limbs(scalars, montgomery=montgomery)


class Srs:
    """Device-resident G1 bases (``CommitterKey::powers_of_g``, src/kzg/time.rs:24-27)."""

    def __init__(self, ctx: "Context", handle: int):
        self.ctx = ctx
        self._h = C.c_void_p(handle)

    def __len__(self) -> int:
        return int(lib.gm_srs_len(self._h))

    def precompute(self, expected_msm_len: int = 0) -> "Srs":
        """One-time table of 2^(c*w) multiples (gm_srs_precompute); MSMs over this SRS get faster."""
        check(lib.gm_srs_precompute(self.ctx._h, self._h, expected_msm_len))
        return self

    def precompute_info(self):
        c, w = C.c_int(0), C.c_int(0)
        check(lib.gm_srs_precompute_info(self._h, C.byref(c), C.byref(w)))
        return int(c.value), int(w.value)

    def read(self, offset: int = 0, n: Optional[int] = None) -> np.ndarray:
        n = len(self) - offset if n is None else n
        out = np.empty((n, 12), dtype=np.uint64)
        check(lib.gm_srs_read(self.ctx._h, self._h, offset, n, _ptr(out)))
        return out

    def points(self, offset: int = 0, n: Optional[int] = None):
        return field.g1_from_limbs(self.read(offset, n))

    def free(self) -> None:
        """Safe before or after ``Context.close()``: the handle keeps the C context alive (gm_srs_free)."""
        if self._h:
            lib.gm_srs_free(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One CUDA device + stream + scratch arena (``gm_ctx``)."""

    def __init__(self, device_id: int = 0):
        h = C.c_void_p()
        check(lib.gm_init(device_id, C.byref(h)))
        self._h = h
        self.device_id = device_id

    # -- lifetime ----------------------------------------------------------------------------
    def close(self) -> None:
        """gm_shutdown: drops this reference.  Handles made from the context (Srs, provers, streams) stay valid and
        may be freed afterwards in any order; calls that need the context itself raise GM_ERR_STATE."""
        if self._h:
            lib.gm_shutdown(self._h)
            self._h = C.c_void_p(None)

    # -- multi-GPU ---------------------------------------------------------------------------
    def comm_init(self, unique_id: bytes, rank: int, world: int) -> None:
        """gm_comm_init: join the NCCL communicator described by ``unique_id`` (128 bytes from ``comm_unique_id()`` on
        rank 0, shared out of band)."""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        check(lib.gm_comm_init(self._h, buf, rank, world))

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        check(lib.gm_comm_unique_id(buf))
        return bytes(buf)

    def comm_init_torch(self, group=None, device=None) -> None:
        """Bootstrap through an initialised torch.distributed group: rank 0 draws the id, a broadcast shares it."""
        import torch
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        raw = self.comm_unique_id() if rank == 0 else bytes(128)
        t = torch.tensor(list(raw), dtype=torch.uint8)
        if device is not None:
            t = t.to(device)
        dist.broadcast(t, src=0, group=group)
        self.comm_init(bytes(t.cpu().tolist()), rank, world)

    @property
    def comm_rank(self) -> int:
        return int(lib.gm_comm_rank(self._h))

    @property
    def comm_world(self) -> int:
        return int(lib.gm_comm_world(self._h))

    def comm_barrier(self) -> None:
        check(lib.gm_comm_barrier(self._h))

    def comm_allgather(self, row: np.ndarray) -> np.ndarray:
        """one small uint64 row per rank -> (world, len) uint64, identical on every rank (gm_comm_allgather)"""
        a = np.ascontiguousarray(row, dtype=np.uint64).reshape(-1)
        out = np.empty((self.comm_world, a.shape[0]), dtype=np.uint64)
        check(lib.gm_comm_allgather(self._h, _ptr(a), a.nbytes, _ptr(out)))
        return out

    def msm_sharded(self, srs: "Srs", scalars, base_offset: int = 0, bigint: bool = False, n: Optional[int] = None) -> np.ndarray:
        """This rank's shard of a multi-GPU msm_unchecked; returns the total over all ranks (gm_msm_g1_sharded)."""
        out = np.empty(18, dtype=np.uint64)
        if hasattr(scalars, "data_ptr"):
            cnt = scalars.numel() * scalars.element_size() // 32 if n is None else n
            fn = lib.gm_msm_g1_sharded_dev if scalars.is_cuda else lib.gm_msm_g1_sharded
            check(fn(self._h, srs._h, base_offset, _ptr(scalars), cnt, int(bigint), _ptr(out)))
            return out
        arr = as_fr_array(scalars, montgomery=not bigint)
        cnt = arr.shape[0] if n is None else n
        check(lib.gm_msm_g1_sharded(self._h, srs._h, base_offset, _ptr(arr), cnt, int(bigint), _ptr(out)))
        return out

    def msm_strided_dev(self, srs: "Srs", scalars_dev_ptr: int, n: int, stride: int, sharded: bool = False, base_offset: int = 0,
                        bigint: bool = False) -> np.ndarray:
        """term i uses the scalar at scalars_dev_ptr + 32 * i * stride (gm_msm_g1_strided_dev)"""
        out = np.empty(18, dtype=np.uint64)
        check(lib.gm_msm_g1_strided_dev(self._h, srs._h, base_offset, C.c_void_p(scalars_dev_ptr), n, stride, int(bigint), int(sharded), _ptr(out)))
        return out

    def srs_subsample(self, srs: "Srs", first: int, stride: int, count: int) -> "Srs":
        h = C.c_void_p()
        check(lib.gm_srs_subsample(self._h, srs._h, first, stride, count, C.byref(h)))
        return Srs(self, h.value)

    def msm_sharded_dev(self, srs: "Srs", scalars_dev_ptr: int, n: int, base_offset: int = 0, bigint: bool = False) -> np.ndarray:
        out = np.empty(18, dtype=np.uint64)
        check(lib.gm_msm_g1_sharded_dev(self._h, srs._h, base_offset, C.c_void_p(scalars_dev_ptr), n, int(bigint), _ptr(out)))
        return out

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(lib.gm_launch_count(self._h))

    def last_device_ms(self, phase: int = 0) -> float:
        return float(lib.gm_last_device_ms(self._h, phase))

    def synchronize(self) -> None:
        check(lib.gm_device_synchronize(self._h))

    def timer_start(self) -> None:
        check(lib.gm_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        check(lib.gm_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def l2_flush(self) -> None:
        check(lib.gm_l2_flush(self._h))

    # -- SRS ---------------------------------------------------------------------------------
    def srs_load(self, points) -> Srs:
        """points: sequence of oracle-style points ((x, y) | None), an (n,12) uint64 array of packed
        Montgomery x|y records, or an (n,104) uint8 array of arkworks ``Affine`` records."""
        if isinstance(points, np.ndarray) and points.dtype == np.uint8:
            arr = np.ascontiguousarray(points)
            n, stride, inf = arr.shape[0], arr.shape[1], 96
        else:
            arr = points if isinstance(points, np.ndarray) else field.g1_to_limbs(points)
            arr = np.ascontiguousarray(arr.reshape(-1, 12))
            n, stride, inf = arr.shape[0], 96, -1
        h = C.c_void_p()
        check(lib.gm_srs_load_g1(self._h, _ptr(arr), n, stride, inf, C.byref(h)))
        return Srs(self, h.value)

    def srs_generate(self, n: int, first_multiple: int = 1) -> Srs:
        h = C.c_void_p()
        check(lib.gm_srs_generate_g1(self._h, n, first_multiple, C.byref(h)))
        return Srs(self, h.value)

    def srs_setup(self, g, tau: int, n: int) -> Srs:
        """powers_of_g[i] = tau^i * g, i < n (CommitterKey::new, kzg/time.rs:49-72), computed on the device."""
        garr = field.g1_to_limbs([g])
        tarr = field.fr_to_limbs([tau])
        h = C.c_void_p()
        check(lib.gm_srs_setup_g1(self._h, _ptr(garr), _ptr(tarr), n, C.byref(h)))
        return Srs(self, h.value)

    def srs_fill(self, point, n: int) -> Srs:
        arr = field.g1_to_limbs([point])
        h = C.c_void_p()
        check(lib.gm_srs_fill_g1(self._h, _ptr(arr), n, C.byref(h)))
        return Srs(self, h.value)

    # -- MSM ---------------------------------------------------------------------------------
    def msm(self, srs: Srs, scalars, base_offset: int = 0, bigint: bool = False, n: Optional[int] = None) -> np.ndarray:
        """msm_unchecked / msm_bigint over ``srs[base_offset:]``; returns the 18-limb Jacobian result."""
        out = np.empty(18, dtype=np.uint64)
        if hasattr(scalars, "data_ptr"):  # torch tensor: pinned host or device memory
            cnt = scalars.numel() * scalars.element_size() // 32 if n is None else n
            fn = lib.gm_msm_g1_dev if scalars.is_cuda else lib.gm_msm_g1
            check(fn(self._h, srs._h, base_offset, _ptr(scalars), cnt, int(bigint), _ptr(out)))
            return out
        arr = as_fr_array(scalars, montgomery=not bigint)
        cnt = arr.shape[0] if n is None else n
        check(lib.gm_msm_g1(self._h, srs._h, base_offset, _ptr(arr), cnt, int(bigint), _ptr(out)))
        return out

    def msm_dev(self, srs: Srs, scalars_dev_ptr: int, n: int, base_offset: int = 0, bigint: bool = False) -> np.ndarray:
        out = np.empty(18, dtype=np.uint64)
        check(lib.gm_msm_g1_dev(self._h, srs._h, base_offset, C.c_void_p(scalars_dev_ptr), n, int(bigint), _ptr(out)))
        return out

    def msm_checked(self, srs: Srs, bases_len: int, scalars, base_offset: int = 0):
        """VariableBaseMSM::msm -> ("ok", jacobian) | ("err", min_len)."""
        arr = as_fr_array(scalars)
        out = np.empty(18, dtype=np.uint64)
        min_len = C.c_size_t(0)
        rc = lib.gm_msm_g1_checked(self._h, srs._h, base_offset, bases_len, _ptr(arr), arr.shape[0], _ptr(out), C.byref(min_len))
        if rc == GM_ERR_LENGTH:
            return ("err", int(min_len.value))
        check(rc)
        return ("ok", out)

    def msm_hostbases(self, points, scalars, bigint: bool = False) -> np.ndarray:
        if isinstance(points, np.ndarray) and points.dtype == np.uint8:
            parr, stride, inf = np.ascontiguousarray(points), points.shape[1], 96
        else:
            parr = points if isinstance(points, np.ndarray) else field.g1_to_limbs(points)
            parr, stride, inf = np.ascontiguousarray(parr.reshape(-1, 12)), 96, -1
        sarr = as_fr_array(scalars, montgomery=not bigint)
        n = min(parr.shape[0], sarr.shape[0])
        out = np.empty(18, dtype=np.uint64)
        check(lib.gm_msm_g1_hostbases(self._h, _ptr(parr), stride, inf, _ptr(sarr), n, int(bigint), _ptr(out)))
        return out

    def g1_sum(self, jacobians) -> np.ndarray:
        arr = np.ascontiguousarray(np.asarray(jacobians, dtype=np.uint64).reshape(-1, 18))
        out = np.empty(18, dtype=np.uint64)
        check(lib.gm_g1_sum(self._h, _ptr(arr), arr.shape[0], _ptr(out)))
        return out

    # -- Fr ----------------------------------------------------------------------------------
    def fr_fold(self, f, r) -> np.ndarray:
        arr = as_fr_array(f)
        rr = as_fr_array([r] if isinstance(r, int) else r)
        out = np.empty(((arr.shape[0] + 1) // 2, 4), dtype=np.uint64)
        check(lib.gm_fr_fold(self._h, _ptr(arr), arr.shape[0], _ptr(rr), _ptr(out)))
        return out

    def fr_fold_chain(self, f, challenges) -> list:
        arr = as_fr_array(f)
        ch = as_fr_array(challenges)
        k, n = ch.shape[0], arr.shape[0]
        tot = int(lib.gm_fr_fold_chain_len(n, k))
        out = np.empty((tot, 4), dtype=np.uint64)
        check(lib.gm_fr_fold_chain(self._h, _ptr(arr), n, _ptr(ch), k, _ptr(out)))
        levels, off = [], 0
        for _ in range(k):
            n = (n + 1) // 2
            levels.append(out[off:off + n])
            off += n
        return levels

    # -- raw device buffers ------------------------------------------------------------------
    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        check(lib.gm_dev_alloc(self._h, nbytes, C.byref(p)))
        return int(p.value)

    def dev_free(self, ptr: int) -> None:
        check(lib.gm_dev_free(self._h, C.c_void_p(ptr)))

    def dev_upload(self, ptr: int, arr) -> None:
        a = np.ascontiguousarray(arr) if isinstance(arr, np.ndarray) else arr
        nbytes = a.nbytes if isinstance(a, np.ndarray) else a.numel() * a.element_size()
        check(lib.gm_dev_upload(self._h, C.c_void_p(ptr), _ptr(a), nbytes))

    def dev_download(self, ptr: int, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes // 8, dtype=np.uint64)
        check(lib.gm_dev_download(self._h, _ptr(out), C.c_void_p(ptr), nbytes))
        return out

    def fr_random_dev(self, ptr: int, n: int, seed: int) -> None:
        check(lib.gm_fr_random_dev(self._h, C.c_void_p(ptr), n, seed))
