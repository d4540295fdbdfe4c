"""Python view of the drop-in C layer (lib/libmangiafuoco_b200.so): the reference's setup / prover / verifier
(snark.h:44-51) with their own struct layouts, called through ctypes.  Used by bench.py for the prove-latency
line and by the tests; the work happens in C and on the GPU.
"""
from __future__ import annotations

import ctypes as C
import time
from pathlib import Path

import numpy as np

from .api import MfbError, load_library

PKG = Path(__file__).resolve().parent
N = 1470
CT_BYTES = 92


class Mpz(C.Structure):  # __mpz_struct of GMP 6 (include/compat/gmp.h)
    _fields_ = [("alloc", C.c_int), ("size", C.c_int), ("d", C.POINTER(C.c_uint64))]


CtT = Mpz * (N + 1)
SkT = Mpz * N


class Proof(C.Structure):  # snark.h:14-20
    _fields_ = [("h", CtT), ("hat_h", CtT), ("hat_v", CtT), ("v_w", CtT), ("b_w", CtT)]


class Vrs(C.Structure):  # snark.h:23-29
    _fields_ = [("alpha", C.c_uint64), ("beta", C.c_uint64), ("s", C.c_uint64), ("sk", SkT)]


class Crs(C.Structure):  # snark.h:31-37
    _fields_ = [("seed", C.c_uint8 * 40), ("s", C.c_void_p), ("as_", C.c_void_p), ("v", C.c_void_p), ("t", C.c_void_p)]


_host = None


def load_host():
    global _host
    if _host is None:
        load_library()
        path = PKG / "lib" / "libmangiafuoco_b200.so"
        if not path.exists():
            raise MfbError(f"{path} is not built (make -C c_lwe_snarks_b200/host)")
        _host = C.CDLL(str(path))
        _host.mf_set_instance.argtypes = [C.c_size_t, C.c_size_t]
        _host.mf_gpu_launches.restype = C.c_uint64
        _host.verifier.restype = C.c_bool
        _host.mf_gamma_d.restype = C.c_size_t
        _host.mf_gamma_d.argtypes = [C.c_size_t]
    return _host


class Snark:
    """One SSP instance of size (D, M): random_ssp -> setup -> prove -> verify, as test_snark.c:19-116 drives them."""

    def __init__(self, D: int, M: int):
        self.lib = load_host()
        self.gmp = C.CDLL("libgmp.so.10")
        self.D, self.M = D, M
        self.lib.mf_set_instance(D, M)
        self.crs, self.vrs, self.proof = Crs(), Vrs(), Proof()
        self.witness = Mpz()
        getattr(self.gmp, "__gmpz_init")(C.byref(self.witness))  # getattr: no class-private name mangling
        self.ssp = np.zeros(D * 8 * (M + 3), np.uint8)
        self._crs_live = self._proof_live = self._vrs_live = False

    def _ssp_ptr(self):
        return self.ssp.ctypes.data_as(C.POINTER(C.c_uint8))

    def random_ssp(self):
        self.lib.random_ssp(C.byref(self.witness), self._ssp_ptr())

    def setup(self) -> float:
        if self._vrs_live:  # a repeated setup replaces the previous CRS / verification key
            self.lib.key_clear(self.vrs.sk)
            self._vrs_live = False
        if self._crs_live:
            self.lib.crs_clear(C.byref(self.crs))
        self.lib.crs_init(C.byref(self.crs))
        self._crs_live = True
        t0 = time.perf_counter()
        self.lib.setup(C.byref(self.crs), C.byref(self.vrs), self._ssp_ptr())
        self._vrs_live = True
        return time.perf_counter() - t0

    def set_devices(self, n: int, spread: bool = True):
        """Shard the resident CRS regions over n GPUs driven by this thread (mf_set_devices); call before make_resident."""
        self.lib.mf_set_devices(int(n), 1 if spread else 0)

    def make_resident(self):
        """Keep the CRS regions s / as (expanded) and the SSP blob (with the cached inverse of rev(t)) in HBM."""
        self.lib.mf_crs_make_resident(C.byref(self.crs))
        self.lib.mf_ssp_make_resident(self._ssp_ptr())
        self._ssp_resident = True

    def prove(self) -> float:
        if self._proof_live:
            self.lib.proof_clear(C.byref(self.proof))
        self.lib.proof_init(C.byref(self.proof))
        self._proof_live = True
        t0 = time.perf_counter()
        self.lib.prover(C.byref(self.proof), C.byref(self.crs), self._ssp_ptr(), C.byref(self.witness))
        return time.perf_counter() - t0

    def verify(self):
        t0 = time.perf_counter()
        ok = bool(self.lib.verifier(self._ssp_ptr(), C.byref(self.vrs), C.byref(self.proof)))
        return ok, time.perf_counter() - t0

    def tamper(self):
        """Flip one bit of the proof (limb 0 of h's b coordinate)."""
        self.proof.h[N].d[0] ^= 1 << 40

    # persistence (mf_io.c)
    def save_crs(self, path: str):
        if self.lib.mf_crs_write(str(path).encode(), C.byref(self.crs)) != 0:
            raise OSError(f"mf_crs_write({path}) failed")

    def load_crs(self, path: str):
        if not self._crs_live:
            self.lib.crs_init(C.byref(self.crs))
            self._crs_live = True
        if self.lib.mf_crs_read(str(path).encode(), C.byref(self.crs)) != 0:
            raise OSError(f"mf_crs_read({path}) failed (wrong instance size or truncated file)")

    def save_proof(self, path: str):
        if self.lib.mf_proof_write(str(path).encode(), C.byref(self.proof)) != 0:
            raise OSError(f"mf_proof_write({path}) failed")

    def load_proof(self, path: str):
        if not self._proof_live:
            self.lib.proof_init(C.byref(self.proof))
            self._proof_live = True
        if self.lib.mf_proof_read(str(path).encode(), C.byref(self.proof)) != 0:
            raise OSError(f"mf_proof_read({path}) failed")

    def crs_records(self):
        """(seed, s, as, t, v) of the live CRS as numpy copies."""
        D, M = self.D, self.M
        grab = lambda p, n: np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n,)).copy()  # noqa: E731
        return (bytes(self.crs.seed), grab(self.crs.s, 92 * D), grab(self.crs.as_, 92 * D), grab(self.crs.t, 92),
                grab(self.crs.v, 92 * (M - 1)))

    def gpu_launches(self) -> int:
        return int(self.lib.mf_gpu_launches())

    def close(self):
        if getattr(self, "_ssp_resident", False):
            self.lib.mf_ssp_release(self._ssp_ptr())
            self._ssp_resident = False
        if self._proof_live:
            self.lib.proof_clear(C.byref(self.proof))
            self._proof_live = False
        if self._vrs_live:
            self.lib.key_clear(self.vrs.sk)
            self._vrs_live = False
        if self._crs_live:
            self.lib.crs_clear(C.byref(self.crs))
            self._crs_live = False
