"""ctypes binding of include/mfb200.h.

Host-flavour methods take and return numpy arrays; ``*_dev`` methods take raw device pointers (ints, e.g.
``tensor.data_ptr()``) and a CUDA stream handle, and are asynchronous.  PyTorch is only plumbing here
(device memory, streams, torch.distributed); nothing in this module imports it.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
N, NC, NCP, L64 = 1470, 1471, 1472, 11
CT_BYTES = 92
CTR_CT = CT_BYTES * N
P = 0xFFFFFFFB
ENT_BYTES = 70
FLAT_CT_U64 = NC * L64
FLAT_SK_U64 = N * L64
PLANAR_U64 = L64 * NCP
ALGO_BYTES_PER_MAC = NC * 88


def resident_to_flat(res: np.ndarray) -> np.ndarray:
    """(k, 16192) u64 in the resident tile-planar layout -> (k, 1471, 11) flat ciphertexts (host-side decode)."""
    k = res.shape[0]
    t = res.reshape(k, NCP // 64, L64, 64)            # [ct][tile][row][lane]
    return np.ascontiguousarray(t.transpose(0, 1, 3, 2).reshape(k, NCP, L64)[:, :NC, :])


class MfbError(RuntimeError):
    pass


def library_path() -> Path:
    return PKG / "lib" / "libmfb200.so"


def build_library(verbose: bool = False) -> Path:
    """Compile the CUDA sources for sm_100a with nvcc (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", str(PKG / "csrc"), "-j4"], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise MfbError("building libmfb200.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return library_path()


_u8p, _u32p, _u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
_vp = C.c_void_p

_SIGS = {
    "mfb_ctx_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "mfb_ctx_destroy": (None, [_vp]),
    "mfb_ctx_warm": (C.c_int, [_vp]),
    "mfb_ctx_reserve": (C.c_int, [_vp, C.c_size_t, C.c_size_t]),
    "mfb_last_error": (C.c_char_p, []),
    "mfb_device_sm_count": (C.c_int, [_vp]),
    "mfb_launch_count": (C.c_uint64, [_vp]),
    "mfb_sync": (C.c_int, [_vp]),
    "mfb_profile_begin": (C.c_int, [_vp]),
    "mfb_profile_end": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "mfb_stream": (C.c_int, [_vp, _u8p, C.c_uint64, _u8p, C.c_size_t]),
    "mfb_stream_dev": (C.c_int, [_vp, _u8p, C.c_uint64, _vp, C.c_size_t, _vp]),
    "mfb_expand_dev": (C.c_int, [_vp, _u8p, C.c_uint64, _vp, C.c_size_t, _vp, _vp]),
    "mfb_lincomb_dev": (C.c_int, [_vp, _vp, _vp, C.c_size_t, _vp, _vp, _vp]),
    "mfb_lincomb2_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_size_t, _vp, _vp, _vp, _vp, _vp]),
    "mfb_region_lincomb2": (C.c_int, [_vp, _vp, C.c_size_t, _u32p, _u32p, C.c_size_t, _u64p, _u64p]),
    "mfb_lincomb_generic_dev": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, C.c_size_t, _vp, _vp]),
    "mfb_encrypt_generic_dev": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _u8p, C.c_uint64, _vp, C.c_int, _vp, _vp, C.c_int, C.c_int,
                                         C.c_size_t, _vp, _vp]),
    "mfb_lincomb": (C.c_int, [_vp, _u64p, _u32p, C.c_size_t, _u64p]),
    "mfb_region_create": (C.c_int, [_vp, _u8p, C.c_uint64, _u8p, C.c_size_t, C.POINTER(_vp)]),
    "mfb_region_create_async": (C.c_int, [_vp, _u8p, C.c_uint64, _u8p, C.c_size_t, _vp, C.POINTER(_vp)]),
    "mfb_region_destroy": (None, [_vp, _vp]),
    "mfb_region_lincomb": (C.c_int, [_vp, _vp, C.c_size_t, _u32p, C.c_size_t, _u64p]),
    "mfb_peer_create": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_vp), _u8p]),
    "mfb_peer_base": (_vp, [_vp]),
    "mfb_peer_connect": (C.c_int, [_vp, _vp, _u8p]),
    "mfb_peer_connect_local": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "mfb_peer_set_timeout": (C.c_int, [_vp, C.c_double]),
    "mfb_peer_status": (C.c_int, [_vp, _vp]),
    "mfb_peer_disconnect": (C.c_int, [_vp, _vp]),
    "mfb_peer_destroy": (None, [_vp, _vp]),
    "mfb_lincomb_peer_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_size_t, _vp, _vp, _vp]),
    "mfb_peer_allreduce_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "mfb_eval_poly_peer_dev": (C.c_int, [_vp, _vp, _u8p, C.c_uint64, _vp, _vp, _vp, C.c_size_t, _vp, _vp, _vp]),
    "mfb_ctx_device": (C.c_int, [_vp]),
    "mfb_region_cts": (_vp, [_vp]),
    "mfb_region_count": (C.c_size_t, [_vp]),
    "mfb_set_create": (C.c_int, [_vp, C.POINTER(C.c_int), C.c_int, C.POINTER(_vp)]),
    "mfb_set_destroy": (None, [_vp]),
    "mfb_set_size": (C.c_int, [_vp]),
    "mfb_set_last_error": (C.c_char_p, []),
    "mfb_set_region_create": (C.c_int, [_vp, _u8p, C.c_uint64, _u8p, C.c_size_t, C.POINTER(_vp)]),
    "mfb_set_region_destroy": (None, [_vp, _vp]),
    "mfb_set_encrypt_cb": (C.c_int, [_vp, _u8p, C.c_uint64, _u64p, _u64p, C.CFUNCTYPE(None, _vp, _vp, C.c_size_t), _vp, C.c_int,
                                    C.c_int, C.c_size_t, _u8p]),
    "mfb_set_eval_poly2": (C.c_int, [_vp, _u8p, C.c_uint64, _u8p, _u64p, _u64p, C.c_size_t, _u64p, _u64p]),
    "mfb_set_region_lincomb2": (C.c_int, [_vp, _vp, _u32p, _u32p, C.c_size_t, _u64p, _u64p]),
    "mfb_columns_split_dev": (C.c_int, [_vp, _vp, _vp, _vp]),
    "mfb_columns_carry_dev": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "mfb_eval_poly": (C.c_int, [_vp, _u8p, C.c_uint64, _u8p, _u64p, _u32p, C.c_size_t, _u64p]),
    "mfb_eval_poly_dev": (C.c_int, [_vp, _u8p, C.c_uint64, _vp, _vp, _vp, C.c_size_t, _vp, _vp, _vp]),
    "mfb_eval_poly2": (C.c_int, [_vp, _u8p, C.c_uint64, _u8p, _u64p, _u64p, C.c_size_t, _u64p, _u64p]),
    "mfb_eval_poly2_dev": (C.c_int, [_vp, _u8p, C.c_uint64, _vp, _vp, _vp, C.c_size_t, _vp, _vp, _vp, _vp, _vp]),
    "mfb_eval_poly2_begin_dev": (C.c_int, [_vp, _u8p, C.c_uint64, _vp, _vp, C.c_size_t, C.c_int, _vp]),
    "mfb_eval_poly2_end_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp, _vp, _vp, _vp, _vp]),
    "mfb_encrypt": (C.c_int, [_vp, _u8p, C.c_uint64, _u64p, _u64p, _u8p, C.c_int, C.c_int, C.c_size_t, _u8p]),
    "mfb_encrypt_cb": (C.c_int, [_vp, _u8p, C.c_uint64, _u64p, _u64p, C.CFUNCTYPE(None, _vp, _vp, C.c_size_t), _vp, C.c_int,
                                C.c_int, C.c_size_t, _u8p]),
    "mfb_encrypt_cb_segs": (C.c_int, [_vp, _u8p, C.c_uint64, _u64p, _u64p, C.CFUNCTYPE(None, _vp, _vp, C.c_size_t), _vp, C.c_int,
                                     C.c_int, C.c_size_t, _vp, C.c_int]),
    "mfb_encrypt_dev": (C.c_int, [_vp, _u8p, C.c_uint64, _vp, _vp, _vp, C.c_int, C.c_int, C.c_size_t, _vp, _vp]),
    "mfb_encrypt_generic_plan": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.POINTER(C.c_uint32)]),
    "mfb_decrypt": (C.c_int, [_vp, _u64p, _u64p, _u8p, C.c_size_t, _u64p, _u64p]),
    "mfb_decrypt_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_size_t, _vp, _vp, _vp]),
    "mfb_ssp_prover_polys": (C.c_int, [_vp, _u64p, C.c_size_t, C.c_size_t, _u64p, C.c_size_t, C.c_uint64, _u64p, _u64p, _u64p]),
    "mfb_ssp_create": (C.c_int, [_vp, _u64p, C.c_size_t, C.c_size_t, C.POINTER(_vp)]),
    "mfb_ssp_destroy": (None, [_vp, _vp]),
    "mfb_ssp_prover_polys_resident": (C.c_int, [_vp, _vp, _u64p, C.c_size_t, C.c_uint64, _u64p, _u64p, _u64p]),
    "mfb_ssp_prover_polys_resident_dev": (C.c_int, [_vp, _vp, _u64p, C.c_size_t, C.c_uint64, C.POINTER(_vp)]),
    "mfb_ssp_degree_bound": (C.c_size_t, [_vp]),
    "mfb_prove_resident": (C.c_int, [_vp, _vp, _vp, _vp, _u64p, C.c_size_t, C.c_uint64, _u64p, _u64p, _u64p, _u64p]),
    "mfb_prove_resident_bw": (C.c_int, [_vp, _vp, _vp, _vp, _u64p, C.c_size_t, C.c_uint64, _u8p, C.c_uint64, _u8p, C.c_size_t,
                                       _u64p, _u64p, _u64p, _u64p, _u64p]),
    "mfb_lincomb2_partials_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_size_t, C.c_int, _vp]),
    "mfb_lincomb_finish4_dev": (C.c_int, [_vp, _vp, _vp, C.c_size_t, _vp]),
    "mfb_peer_allreduce_lanes_dev": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_int, _vp, _vp, C.c_size_t, _vp]),
    "mfb_peer_finish4_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "mfb_peer_finish4_push_dev": (C.c_int, [_vp, _vp, _vp]),
    "mfb_peer_wait4_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "mfb_peer_push_lanes_dev": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_int, _vp]),
    "mfb_peer_wait_lanes_dev": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, C.c_size_t, _vp]),
    "mfb_ssp_prover_polys_resident_async": (C.c_int, [_vp, _vp, _u64p, C.c_size_t, C.c_uint64, _vp, C.POINTER(_vp)]),
    "mfb_b_w_dev": (C.c_int, [_vp, _u8p, C.c_uint64, _u8p, C.c_size_t, _u64p, C.c_size_t, C.c_uint64, _vp, _vp]),
    "mfb_set_encrypt_par": (C.c_int, [_vp, _u8p, C.c_uint64, _u64p, _u64p, C.CFUNCTYPE(None, _vp, _vp, C.c_size_t), _vp, C.c_int,
                                     C.c_int, C.c_size_t, _u8p, _vp, C.c_int]),
    "mfb_set_prove_resident_bw": (C.c_int, [_vp, _vp, _vp, _vp, _u64p, C.c_size_t, C.c_uint64, _u8p, C.c_uint64, _u8p, C.c_size_t,
                                           _u64p, _u64p, _u64p, _u64p, _u64p]),
    "mfb_set_prove_resident": (C.c_int, [_vp, _vp, _vp, _vp, _u64p, C.c_size_t, C.c_uint64, _u64p, _u64p, _u64p, _u64p]),
    "mfb_ssp_eval_resident": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_uint64, _u64p]),
    "mfb_ssp_eval": (C.c_int, [_vp, _u64p, C.c_size_t, C.c_size_t, C.c_uint64, _u64p]),
    "mfb_flat_to_planar_dev": (C.c_int, [_vp, _vp, C.c_int, C.c_size_t, _vp, _vp]),
    "mfb_flat_to_resident_dev": (C.c_int, [_vp, _vp, C.c_size_t, _vp, _vp]),
}

EXPORTS = tuple(_SIGS)
_lib = None


def load_library() -> C.CDLL:
    """Load lib/libmfb200.so.  There is no fallback: a missing library is an error."""
    global _lib
    if _lib is None:
        path = library_path()
        if not path.exists():
            raise MfbError(f"{path} is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `make -C c_lwe_snarks_b200/csrc`); there is no CPU fallback")
        lib = C.CDLL(str(path))
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _p8(a):
    return a.ctypes.data_as(_u8p)


def _p32(a):
    return a.ctypes.data_as(_u32p)


def _p64(a):
    return a.ctypes.data_as(_u64p)


def _seed(seed) -> np.ndarray:
    s = np.frombuffer(bytes(seed), dtype=np.uint8).copy()
    if s.size != 40:
        raise ValueError("seed is 40 bytes: nonce(8) || key(32)")
    return s


def _arr(a, dtype) -> np.ndarray:
    if isinstance(a, (bytes, bytearray)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=dtype)


class Region:
    """A CRS region expanded into HBM (planar layout) and kept resident."""

    def __init__(self, ctx: "Context", handle: int, count: int):
        self.ctx, self.handle, self.count = ctx, handle, count

    def lincomb(self, coeffs, first: int = 0, rop=None) -> np.ndarray:
        co = _arr(coeffs, np.uint32)
        r = np.zeros((NC, L64), np.uint64) if rop is None else _arr(rop, np.uint64).copy()
        self.ctx._ck(self.ctx.lib.mfb_region_lincomb(self.ctx.h, self.handle, first, _p32(co), co.size, _p64(r)))
        return r

    def lincomb2(self, coeffs0, coeffs1, first: int = 0, rop0=None, rop1=None):
        c0, c1 = _arr(coeffs0, np.uint32), _arr(coeffs1, np.uint32)
        if c0.size != c1.size:
            raise ValueError("coefficient vectors must have equal length")
        r0 = np.zeros((NC, L64), np.uint64) if rop0 is None else _arr(rop0, np.uint64).copy()
        r1 = np.zeros((NC, L64), np.uint64) if rop1 is None else _arr(rop1, np.uint64).copy()
        self.ctx._ck(self.ctx.lib.mfb_region_lincomb2(self.ctx.h, self.handle, first, _p32(c0), _p32(c1), c0.size, _p64(r0),
                                                      _p64(r1)))
        return r0, r1

    def close(self):
        if self.handle:
            self.ctx.lib.mfb_region_destroy(self.ctx.h, self.handle)
            self.handle = None


PEER_HANDLE_BYTES = 64


class PeerGroup:
    """This rank's end of a peer-memory exchange group (mfb_peer_*): a symmetric buffer the other ranks map over
    NVLink; lincomb_dev / eval_poly_dev run the sharded lincomb with the exchange fused into the finish kernel."""

    def __init__(self, ctx: "Context", world: int, rank: int):
        self.ctx, self.world, self.rank = ctx, world, rank
        h = _vp()
        handle = np.zeros(PEER_HANDLE_BYTES, np.uint8)
        ctx._ck(ctx.lib.mfb_peer_create(ctx.h, world, rank, C.byref(h), _p8(handle)))
        self.handle_, self.ipc_handle = h, handle

    @property
    def base(self) -> int:
        return int(self.ctx.lib.mfb_peer_base(self.handle_))

    def connect(self, handles):
        """handles: world x 64 bytes, rank order (every rank's ``ipc_handle``)."""
        hs = _arr(handles, np.uint8).reshape(-1)
        if hs.size != self.world * PEER_HANDLE_BYTES:
            raise ValueError("need one 64-byte handle per rank")
        self.ctx._ck(self.ctx.lib.mfb_peer_connect(self.ctx.h, self.handle_, _p8(hs)))

    def connect_local(self, bases):
        """Same-process group: bases[r] = PeerGroup.base of rank r."""
        arr = (_vp * self.world)(*[int(b) for b in bases])
        self.ctx._ck(self.ctx.lib.mfb_peer_connect_local(self.ctx.h, self.handle_, arr))

    def set_timeout(self, seconds: float):
        self.ctx._ck(self.ctx.lib.mfb_peer_set_timeout(self.handle_, float(seconds)))

    def check(self):
        self.ctx._ck(self.ctx.lib.mfb_peer_status(self.ctx.h, self.handle_))

    def lincomb_dev(self, cts_ptr: int, coeffs_ptr: int, d: int, rop_in_ptr, rop_out_ptr: int, stream: int = 0):
        self.ctx._ck(self.ctx.lib.mfb_lincomb_peer_dev(self.ctx.h, self.handle_, cts_ptr, coeffs_ptr, d, rop_in_ptr, rop_out_ptr,
                                                       stream))

    def allreduce_dev(self, partial_flat_ptr: int, rop_in_ptr, rop_out_ptr: int, stream: int = 0):
        self.ctx._ck(self.ctx.lib.mfb_peer_allreduce_dev(self.ctx.h, self.handle_, partial_flat_ptr, rop_in_ptr, rop_out_ptr, stream))

    def eval_poly_dev(self, seed, offset: int, c8_ptr: int, coeffs_ptr: int, idx_ptr, d: int, rop_in_ptr, rop_out_ptr: int,
                      stream: int = 0):
        self.ctx._ck(self.ctx.lib.mfb_eval_poly_peer_dev(self.ctx.h, self.handle_, _p8(_seed(seed)), offset, c8_ptr, coeffs_ptr,
                                                         idx_ptr, d, rop_in_ptr, rop_out_ptr, stream))

    def disconnect(self):
        if self.handle_:
            self.ctx._ck(self.ctx.lib.mfb_peer_disconnect(self.ctx.h, self.handle_))

    def close(self):
        if self.handle_:
            self.ctx.lib.mfb_peer_destroy(self.ctx.h, self.handle_)
            self.handle_ = None


class DeviceSet:
    """Several GPUs driven by this thread (mfb_set_*): the primary context plus one member per entry of `devices`."""

    def __init__(self, ctx: "Context", devices):
        self.ctx = ctx
        devs = (C.c_int * len(devices))(*devices)
        h = _vp()
        rc = ctx.lib.mfb_set_create(ctx.h, devs, len(devices), C.byref(h))
        if rc != 0:
            raise MfbError(f"mfb_set_create failed ({rc}): {ctx.lib.mfb_set_last_error().decode()}")
        self.h = h

    def _ck(self, rc):
        if rc != 0:
            raise MfbError(f"mfb_set call failed ({rc}): {self.ctx.lib.mfb_set_last_error().decode()}")

    @property
    def size(self) -> int:
        return int(self.ctx.lib.mfb_set_size(self.h))

    def encrypt_cb(self, seed, offset: int, sk_flat, msg, draw, ent_stride: int = ENT_BYTES, ent_nbytes: int = ENT_BYTES - 1):
        """Context.encrypt_cb with the pieces spread over the members (mfb_set_encrypt_cb)."""
        s, sk, m = _seed(seed), _arr(sk_flat, np.uint64), _arr(msg, np.uint64)
        if sk.size != FLAT_SK_U64:
            raise ValueError("sk_flat must be (1470, 11) uint64")
        out = np.zeros((m.size, CT_BYTES), np.uint8)
        fn_t = _SIGS["mfb_set_encrypt_cb"][1][5]

        def _draw(_user, dst, nbytes):
            data = bytes(draw(nbytes))
            if len(data) != nbytes:
                raise ValueError("entropy callback returned the wrong number of bytes")
            C.memmove(dst, data, nbytes)

        cb = fn_t(_draw)
        self._ck(self.ctx.lib.mfb_set_encrypt_cb(self.h, _p8(s), offset, _p64(sk), _p64(m), cb, None, ent_stride, ent_nbytes,
                                                m.size, _p8(out)))
        return out

    def encrypt_par(self, seed, offset: int, sk_flat, msg, draw, ent_stride: int = ENT_BYTES, ent_nbytes: int = ENT_BYTES - 1,
                    segments=None):
        """mfb_set_encrypt_par: one host thread per member, contiguous ranges; `draw(nbytes)` is called concurrently from the
        member threads.  segments = [(first, count), ...] partitions the record index space: the records are then returned
        as one array per segment (written by the library straight into them), else as one (count, 92) array."""
        s, sk, m = _seed(seed), _arr(sk_flat, np.uint64), _arr(msg, np.uint64)
        if sk.size != FLAT_SK_U64:
            raise ValueError("sk_flat must be (1470, 11) uint64")
        fn_t = _SIGS["mfb_set_encrypt_par"][1][5]

        def _draw(_user, dst, nbytes):
            data = bytes(draw(nbytes))
            if len(data) != nbytes:
                raise ValueError("entropy callback returned the wrong number of bytes")
            C.memmove(dst, data, nbytes)

        cb = fn_t(_draw)
        if segments is None:
            out = np.zeros((m.size, CT_BYTES), np.uint8)
            self._ck(self.ctx.lib.mfb_set_encrypt_par(self.h, _p8(s), offset, _p64(sk), _p64(m), cb, None, ent_stride, ent_nbytes,
                                                     m.size, _p8(out), None, 0))
            return out

        class Seg(C.Structure):
            _fields_ = [("first", C.c_size_t), ("count", C.c_size_t), ("dst", C.c_void_p)]

        outs = [np.zeros((cnt, CT_BYTES), np.uint8) for _, cnt in segments]
        segs = (Seg * len(segments))(*[Seg(f, c, o.ctypes.data) for (f, c), o in zip(segments, outs)])
        self._ck(self.ctx.lib.mfb_set_encrypt_par(self.h, _p8(s), offset, _p64(sk), _p64(m), cb, None, ent_stride, ent_nbytes,
                                                 m.size, None, C.cast(segs, C.c_void_p), len(segments)))
        return outs

    def eval_poly2(self, seed, offset: int, c8, coeffs0, coeffs1=None, rop0=None, rop1=None):
        """eval_poly (coeffs1 None) / eval_poly2 with nothing resident, sharded over the members."""
        s, rec = _seed(seed), _arr(c8, np.uint8)
        c0 = _arr(coeffs0, np.uint64)
        c1 = None if coeffs1 is None else _arr(coeffs1, np.uint64)
        r0 = np.zeros((NC, L64), np.uint64) if rop0 is None else _arr(rop0, np.uint64).copy()
        r1 = None if c1 is None else (np.zeros((NC, L64), np.uint64) if rop1 is None else _arr(rop1, np.uint64).copy())
        self._ck(self.ctx.lib.mfb_set_eval_poly2(self.h, _p8(s), offset, _p8(rec), _p64(c0), None if c1 is None else _p64(c1),
                                                c0.size, _p64(r0), None if r1 is None else _p64(r1)))
        return r0 if c1 is None else (r0, r1)

    def region(self, seed, offset: int, c8) -> "SetRegion":
        s, rec = _seed(seed), _arr(c8, np.uint8)
        count = rec.size // CT_BYTES
        h = _vp()
        self._ck(self.ctx.lib.mfb_set_region_create(self.h, _p8(s), offset, _p8(rec), count, C.byref(h)))
        return SetRegion(self, h, count)

    def close(self):
        if self.h:
            self.ctx.lib.mfb_set_destroy(self.h)
            self.h = None


class SetRegion:
    """A CRS region sharded by ciphertext index over the members of a DeviceSet."""

    def __init__(self, dset: DeviceSet, handle, count: int):
        self.dset, self.handle, self.count = dset, handle, count

    def lincomb(self, coeffs, rop=None) -> np.ndarray:
        co = _arr(coeffs, np.uint32)
        r = np.zeros((NC, L64), np.uint64) if rop is None else _arr(rop, np.uint64).copy()
        self.dset._ck(self.dset.ctx.lib.mfb_set_region_lincomb2(self.dset.h, self.handle, _p32(co), None, co.size, _p64(r), None))
        return r

    def lincomb2(self, coeffs0, coeffs1, rop0=None, rop1=None):
        c0, c1 = _arr(coeffs0, np.uint32), _arr(coeffs1, np.uint32)
        if c0.size != c1.size:
            raise ValueError("coefficient vectors must have equal length")
        r0 = np.zeros((NC, L64), np.uint64) if rop0 is None else _arr(rop0, np.uint64).copy()
        r1 = np.zeros((NC, L64), np.uint64) if rop1 is None else _arr(rop1, np.uint64).copy()
        self.dset._ck(self.dset.ctx.lib.mfb_set_region_lincomb2(self.dset.h, self.handle, _p32(c0), _p32(c1), c0.size, _p64(r0),
                                                               _p64(r1)))
        return r0, r1

    def close(self):
        if self.handle:
            self.dset.ctx.lib.mfb_set_region_destroy(self.dset.h, self.handle)
            self.handle = None


class ResidentSsp:
    """An SSP blob kept on the device (with the cached Newton inverse of rev(t))."""

    def __init__(self, ctx: "Context", handle, D: int, M: int):
        self.ctx, self.handle, self.D, self.M = ctx, handle, D, M

    def prover_polys(self, witness_limbs, delta: int):
        wl = _arr(witness_limbs, np.uint64)
        w, v, h = (np.zeros(self.D, np.uint64) for _ in range(3))
        self.ctx._ck(self.ctx.lib.mfb_ssp_prover_polys_resident(self.ctx.h, self.handle, _p64(wl), wl.size, delta, _p64(w),
                                                                _p64(v), _p64(h)))
        return w, v, h

    def eval(self, first: int, npoly: int, x: int) -> np.ndarray:
        out = np.zeros(npoly, np.uint64)
        self.ctx._ck(self.ctx.lib.mfb_ssp_eval_resident(self.ctx.h, self.handle, first, npoly, x, _p64(out)))
        return out

    def close(self):
        if self.handle:
            self.ctx.lib.mfb_ssp_destroy(self.ctx.h, self.handle)
            self.handle = None


class Context:
    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = _vp()
        rc = self.lib.mfb_ctx_create(C.byref(h), device)
        if rc != 0:
            raise MfbError(f"mfb_ctx_create({device}) failed ({rc}): {self.lib.mfb_last_error().decode()}")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.mfb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def _ck(self, rc: int):
        if rc != 0:
            raise MfbError(f"mfb call failed ({rc}): {self.lib.mfb_last_error().decode()}")

    @property
    def sm_count(self) -> int:
        return int(self.lib.mfb_device_sm_count(self.h))

    @property
    def launches(self) -> int:
        return int(self.lib.mfb_launch_count(self.h))

    def sync(self):
        self._ck(self.lib.mfb_sync(self.h))

    def profile_begin(self):
        self._ck(self.lib.mfb_profile_begin(self.h))

    def profile_end(self):
        """(summed ms of the dominant kernel, number of launches timed) since profile_begin."""
        ms, n = C.c_double(0), C.c_int(0)
        self._ck(self.lib.mfb_profile_end(self.h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    # ---------------------------------------------------------------- host flavour
    def stream(self, seed, offset: int, nbytes: int) -> np.ndarray:
        s, out = _seed(seed), np.zeros(nbytes, np.uint8)
        self._ck(self.lib.mfb_stream(self.h, _p8(s), offset, _p8(out), nbytes))
        return out

    def eval_poly(self, seed, offset: int, c8, coeffs, rop=None, idx=None) -> np.ndarray:
        s, rec, co = _seed(seed), _arr(c8, np.uint8), _arr(coeffs, np.uint64)
        ix = None if idx is None else _arr(idx, np.uint32)
        r = np.zeros((NC, L64), np.uint64) if rop is None else _arr(rop, np.uint64).copy()
        need = (int(ix.max()) + 1 if ix is not None and ix.size else co.size) * CT_BYTES
        if rec.size < need:
            raise ValueError("record array shorter than the ciphertext indices used")
        self._ck(self.lib.mfb_eval_poly(self.h, _p8(s), offset, _p8(rec), _p64(co), None if ix is None else _p32(ix),
                                        co.size, _p64(r)))
        return r

    def eval_poly2(self, seed, offset: int, c8, coeffs0, coeffs1, rop0=None, rop1=None):
        s, rec = _seed(seed), _arr(c8, np.uint8)
        c0, c1 = _arr(coeffs0, np.uint64), _arr(coeffs1, np.uint64)
        if c0.size != c1.size or rec.size < c0.size * CT_BYTES:
            raise ValueError("coefficient vectors must have equal length and one record per ciphertext")
        r0 = np.zeros((NC, L64), np.uint64) if rop0 is None else _arr(rop0, np.uint64).copy()
        r1 = np.zeros((NC, L64), np.uint64) if rop1 is None else _arr(rop1, np.uint64).copy()
        self._ck(self.lib.mfb_eval_poly2(self.h, _p8(s), offset, _p8(rec), _p64(c0), _p64(c1), c0.size, _p64(r0), _p64(r1)))
        return r0, r1

    def lincomb(self, cts_flat, coeffs, rop=None) -> np.ndarray:
        cts, co = _arr(cts_flat, np.uint64), _arr(coeffs, np.uint32)
        if cts.size != co.size * FLAT_CT_U64:
            raise ValueError("cts_flat must be (d, 1471, 11) uint64")
        r = np.zeros((NC, L64), np.uint64) if rop is None else _arr(rop, np.uint64).copy()
        self._ck(self.lib.mfb_lincomb(self.h, _p64(cts), _p32(co), co.size, _p64(r)))
        return r

    def region(self, seed, offset: int, c8) -> Region:
        s, rec = _seed(seed), _arr(c8, np.uint8)
        count = rec.size // CT_BYTES
        h = _vp()
        self._ck(self.lib.mfb_region_create(self.h, _p8(s), offset, _p8(rec), count, C.byref(h)))
        return Region(self, h, count)

    def encrypt(self, seed, offset: int, sk_flat, msg, ent, ent_stride: int = ENT_BYTES,
                ent_nbytes: int = ENT_BYTES - 1) -> np.ndarray:
        s, sk, m, e = _seed(seed), _arr(sk_flat, np.uint64), _arr(msg, np.uint64), _arr(ent, np.uint8)
        if sk.size != FLAT_SK_U64:
            raise ValueError("sk_flat must be (1470, 11) uint64")
        if e.size < m.size * ent_stride:
            raise ValueError("entropy buffer too short")
        out = np.zeros((m.size, CT_BYTES), np.uint8)
        self._ck(self.lib.mfb_encrypt(self.h, _p8(s), offset, _p64(sk), _p64(m), _p8(e), ent_stride, ent_nbytes,
                                      m.size, _p8(out)))
        return out

    def encrypt_cb(self, seed, offset: int, sk_flat, msg, draw, ent_stride: int = ENT_BYTES,
                   ent_nbytes: int = ENT_BYTES - 1) -> np.ndarray:
        """encrypt() with the entropy drawn through ``draw(nbytes) -> bytes`` piece by piece, in order, while the
        device encrypts the previous piece (mfb_encrypt_cb; what setup() uses)."""
        s, sk, m = _seed(seed), _arr(sk_flat, np.uint64), _arr(msg, np.uint64)
        if sk.size != FLAT_SK_U64:
            raise ValueError("sk_flat must be (1470, 11) uint64")
        out = np.zeros((m.size, CT_BYTES), np.uint8)
        fn_t = _SIGS["mfb_encrypt_cb"][1][5]

        def _draw(_user, dst, nbytes):
            data = bytes(draw(nbytes))
            if len(data) != nbytes:
                raise ValueError("entropy callback returned the wrong number of bytes")
            C.memmove(dst, data, nbytes)

        cb = fn_t(_draw)
        self._ck(self.lib.mfb_encrypt_cb(self.h, _p8(s), offset, _p64(sk), _p64(m), cb, None, ent_stride, ent_nbytes, m.size,
                                         _p8(out)))
        return out

    def encrypt_cb_segs(self, seed, offset: int, sk_flat, msg, draw, cuts, ent_stride: int = ENT_BYTES,
                        ent_nbytes: int = ENT_BYTES - 1):
        """encrypt_cb() with the records written straight into separate arrays (mfb_encrypt_cb_segs; setup() fills
        crs->s / as / t / v this way): ``cuts`` = record counts of consecutive segments; returns one array per segment."""
        s, sk, m = _seed(seed), _arr(sk_flat, np.uint64), _arr(msg, np.uint64)
        if sk.size != FLAT_SK_U64 or sum(cuts) != m.size:
            raise ValueError("sk_flat must be (1470, 11) uint64 and the segments must cover the messages")

        class Seg(C.Structure):
            _fields_ = [("first", C.c_size_t), ("count", C.c_size_t), ("dst", C.c_void_p)]

        outs = [np.zeros((c, CT_BYTES), np.uint8) for c in cuts]
        segs, first = (Seg * len(cuts))(), 0
        for i, (c, o) in enumerate(zip(cuts, outs)):
            segs[i] = Seg(first, c, o.ctypes.data)
            first += c
        fn_t = _SIGS["mfb_encrypt_cb_segs"][1][5]

        def _draw(_user, dst, nbytes):
            data = bytes(draw(nbytes))
            if len(data) != nbytes:
                raise ValueError("entropy callback returned the wrong number of bytes")
            C.memmove(dst, data, nbytes)

        cb = fn_t(_draw)
        self._ck(self.lib.mfb_encrypt_cb_segs(self.h, _p8(s), offset, _p64(sk), _p64(m), cb, None, ent_stride, ent_nbytes, m.size,
                                              C.cast(segs, C.c_void_p), len(cuts)))
        return outs

    def decrypt(self, sk_flat, cts_flat, b_neg=None, want_dot: bool = False):
        sk, cts = _arr(sk_flat, np.uint64), _arr(cts_flat, np.uint64)
        count = cts.size // FLAT_CT_U64
        neg = None if b_neg is None else _arr(np.asarray(b_neg, dtype=bool), np.uint8)
        m = np.zeros(count, np.uint64)
        dot = np.zeros((count, L64), np.uint64) if want_dot else None
        self._ck(self.lib.mfb_decrypt(self.h, _p64(sk), _p64(cts), None if neg is None else _p8(neg), count, _p64(m),
                                      None if dot is None else _p64(dot)))
        return (m, dot) if want_dot else m

    def device_set(self, devices) -> DeviceSet:
        """this context + one member per entry of `devices` (entries may repeat / equal this context's device)"""
        return DeviceSet(self, list(devices))

    def peer_group(self, world: int, rank: int) -> PeerGroup:
        return PeerGroup(self, world, rank)

    def ssp_prover_polys(self, ssp, D: int, M: int, witness_limbs, delta: int):
        """(w, v, h) coefficient arrays (D u64 each) of the prover's polynomial step."""
        blob, wl = _arr(ssp, np.uint8).view(np.uint64), _arr(witness_limbs, np.uint64)
        if blob.size < D * (M + 1):
            raise ValueError("SSP blob shorter than D*(M+1) coefficients")
        w, v, h = (np.zeros(D, np.uint64) for _ in range(3))
        self._ck(self.lib.mfb_ssp_prover_polys(self.h, _p64(blob), D, M, _p64(wl), wl.size, delta, _p64(w), _p64(v), _p64(h)))
        return w, v, h

    def ssp_resident(self, ssp, D: int, M: int) -> "ResidentSsp":
        blob = _arr(ssp, np.uint8).view(np.uint64)
        if blob.size < D * (M + 1):
            raise ValueError("SSP blob shorter than D*(M+1) coefficients")
        h = _vp()
        self._ck(self.lib.mfb_ssp_create(self.h, _p64(blob), D, M, C.byref(h)))
        return ResidentSsp(self, h, D, M)

    def ssp_eval(self, polys, D: int, x: int) -> np.ndarray:
        pl = _arr(polys, np.uint64).reshape(-1)
        npoly = pl.size // D
        out = np.zeros(npoly, np.uint64)
        self._ck(self.lib.mfb_ssp_eval(self.h, _p64(pl), D, npoly, x, _p64(out)))
        return out

    # ---------------------------------------------------------------- device flavour (raw pointers)
    def stream_dev(self, seed, offset: int, out_ptr: int, nbytes: int, stream: int = 0):
        self._ck(self.lib.mfb_stream_dev(self.h, _p8(_seed(seed)), offset, out_ptr, nbytes, stream))

    def expand_dev(self, seed, offset: int, c8_ptr: int, count: int, cts_ptr: int, stream: int = 0):
        self._ck(self.lib.mfb_expand_dev(self.h, _p8(_seed(seed)), offset, c8_ptr, count, cts_ptr, stream))

    def lincomb_dev(self, cts_ptr: int, coeffs_ptr: int, d: int, rop_in_ptr, rop_out_ptr: int, stream: int = 0):
        self._ck(self.lib.mfb_lincomb_dev(self.h, cts_ptr, coeffs_ptr, d, rop_in_ptr, rop_out_ptr, stream))

    def lincomb2_dev(self, cts_ptr: int, coeffs0_ptr: int, coeffs1_ptr: int, d: int, rop0_in, rop0_out: int, rop1_in,
                     rop1_out: int, stream: int = 0):
        self._ck(self.lib.mfb_lincomb2_dev(self.h, cts_ptr, coeffs0_ptr, coeffs1_ptr, d, rop0_in, rop0_out, rop1_in, rop1_out,
                                           stream))

    def encrypt_generic_dev(self, limbs64: int, n: int, ct_bytes: int, seed, offset: int, sk_ptr: int, sk_stride: int, msg_ptr: int,
                            ent_ptr: int, ent_stride: int, ent_nbytes: int, count: int, out_ptr: int, stream: int = 0):
        self._ck(self.lib.mfb_encrypt_generic_dev(self.h, limbs64, n, ct_bytes, _p8(_seed(seed)), offset, sk_ptr, sk_stride, msg_ptr,
                                                  ent_ptr, ent_stride, ent_nbytes, count, out_ptr, stream))

    def lincomb_generic_dev(self, limbs64: int, ncoords: int, cts_ptr: int, coeffs_ptr: int, d: int, out_ptr: int, stream: int = 0):
        self._ck(self.lib.mfb_lincomb_generic_dev(self.h, limbs64, ncoords, cts_ptr, coeffs_ptr, d, out_ptr, stream))

    def eval_poly_dev(self, seed, offset: int, c8_ptr: int, coeffs_ptr: int, idx_ptr, d: int, rop_in_ptr,
                      rop_out_ptr: int, stream: int = 0):
        self._ck(self.lib.mfb_eval_poly_dev(self.h, _p8(_seed(seed)), offset, c8_ptr, coeffs_ptr, idx_ptr, d, rop_in_ptr,
                                            rop_out_ptr, stream))

    def eval_poly2_dev(self, seed, offset: int, c8_ptr: int, coeffs0_ptr: int, coeffs1_ptr: int, d: int, rop0_in, rop0_out: int,
                       rop1_in, rop1_out: int, stream: int = 0):
        self._ck(self.lib.mfb_eval_poly2_dev(self.h, _p8(_seed(seed)), offset, c8_ptr, coeffs0_ptr, coeffs1_ptr, d, rop0_in,
                                             rop0_out, rop1_in, rop1_out, stream))

    def eval_poly2_begin_dev(self, seed, offset: int, coeffs0_ptr: int, coeffs1_ptr, d: int, records_from_host: bool, stream: int = 0):
        self._ck(self.lib.mfb_eval_poly2_begin_dev(self.h, _p8(_seed(seed)), offset, coeffs0_ptr, coeffs1_ptr, d,
                                                   1 if records_from_host else 0, stream))

    def eval_poly2_end_dev(self, c8_dev_ptr: int, c8_host_ptr, coeffs0_ptr: int, coeffs1_ptr, d: int, rop0_in, rop0_out: int,
                           rop1_in, rop1_out, stream: int = 0):
        self._ck(self.lib.mfb_eval_poly2_end_dev(self.h, c8_dev_ptr, c8_host_ptr, coeffs0_ptr, coeffs1_ptr, d, rop0_in, rop0_out,
                                                 rop1_in, rop1_out, stream))

    def columns_split_dev(self, flat_ptr: int, cols_ptr: int, stream: int = 0):
        self._ck(self.lib.mfb_columns_split_dev(self.h, flat_ptr, cols_ptr, stream))

    def columns_carry_dev(self, cols_ptr: int, c0: int, ncoord: int, flat_in_ptr, flat_out_ptr: int, stream: int = 0):
        self._ck(self.lib.mfb_columns_carry_dev(self.h, cols_ptr, c0, ncoord, flat_in_ptr, flat_out_ptr, stream))

    def encrypt_dev(self, seed, offset: int, sk_planar_ptr: int, msg_ptr: int, ent_ptr: int, ent_stride: int,
                    ent_nbytes: int, count: int, out_ptr: int, stream: int = 0):
        self._ck(self.lib.mfb_encrypt_dev(self.h, _p8(_seed(seed)), offset, sk_planar_ptr, msg_ptr, ent_ptr, ent_stride,
                                          ent_nbytes, count, out_ptr, stream))

    def decrypt_dev(self, sk_planar_ptr: int, cts_flat_ptr: int, b_neg_ptr, count: int, out_m_ptr: int, out_dot_ptr,
                    stream: int = 0):
        self._ck(self.lib.mfb_decrypt_dev(self.h, sk_planar_ptr, cts_flat_ptr, b_neg_ptr, count, out_m_ptr, out_dot_ptr,
                                          stream))

    def flat_to_resident_dev(self, flat_ptr: int, count: int, cts_ptr: int, stream: int = 0):
        self._ck(self.lib.mfb_flat_to_resident_dev(self.h, flat_ptr, count, cts_ptr, stream))

    def flat_to_planar_dev(self, flat_ptr: int, n: int, count: int, planar_ptr: int, stream: int = 0):
        self._ck(self.lib.mfb_flat_to_planar_dev(self.h, flat_ptr, n, count, planar_ptr, stream))
