"""TEST INFRASTRUCTURE — ctypes front-ends for the two CPU checkers.

* :class:`Oracle`     -> ``oracle/liboracle.so``  (our plain-C restatement, ``mf_oracle.c``)
* :class:`Reference`  -> ``oracle/_ref/libmfref_d<D>_m<M>.so``  (the unmodified reference sources
  compiled by ``oracle/Makefile`` + ``ref_shim.c``)

Both expose the same method names and the same flat numpy formats, so a test can run one body
against either.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import
this module; the product package never does.

Formats: a coordinate is 12 little-endian uint64 limbs, a ciphertext is ``(1471, 12)`` uint64, a
secret key ``(1470, 12)`` uint64, wire records are ``(k, 92)`` uint8, seeds are 40 bytes.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
N = 1470
NC = N + 1
LIMBS = 12
P = 0xFFFFFFFB
CT_BYTES = 92
CTR_CT = CT_BYTES * N
NOISE_BYTES = 69
SMUDGE_BYTES = 80

_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int32)
_dblp = C.POINTER(C.c_double)


def _p8(a):
    return a.ctypes.data_as(_u8p)


def _p64(a):
    return a.ctypes.data_as(_u64p)


def _p32(a):
    return a.ctypes.data_as(_i32p)


def _seed(seed) -> np.ndarray:
    s = np.frombuffer(bytes(seed), dtype=np.uint8).copy()
    assert s.size == 40, "seed is 40 bytes (8 nonce + 32 key)"
    return s


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray)):
        return np.frombuffer(bytes(a), dtype=np.uint8).copy()
    return np.ascontiguousarray(a, dtype=np.uint8)


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def build_oracle() -> Path:
    """Compile oracle/liboracle.so (gcc, a second or two)."""
    subprocess.run(["make", "-s", "-C", str(HERE), "oracle"], check=True)
    return HERE / "liboracle.so"


def build_reference(D: int = 256, M: int = 64, ref_src: str = "/root/reference/src") -> Path | None:
    """Compile the reference for instance size (D, M) if its sources are present; else None."""
    out = HERE / "_ref" / f"libmfref_d{D}_m{M}.so"
    if not Path(ref_src).is_dir():
        return out if out.exists() else None
    newest = max(p.stat().st_mtime for p in [HERE / "ref_shim.c", HERE / "Makefile"])
    if not out.exists() or out.stat().st_mtime < newest:
        subprocess.run(["make", "-s", "-C", str(HERE), "ref", f"D={D}", f"M={M}", f"REF_SRC={ref_src}"],
                       check=True)
    return out


def build_dropin() -> Path:
    """Compile ref_shim.c against the PRODUCT's drop-in headers/library (oracle/_dropin/libmfdropin.so)."""
    out = HERE / "_dropin" / "libmfdropin.so"
    lib = HERE.parent / "c_lwe_snarks_b200" / "lib" / "libmangiafuoco_b200.so"
    newest = max(p.stat().st_mtime for p in [HERE / "ref_shim.c", HERE / "Makefile", lib] if p.exists())
    if not out.exists() or out.stat().st_mtime < newest:
        subprocess.run(["make", "-s", "-C", str(HERE), "dropin"], check=True)
    return out


class Oracle:
    """Our CPU restatement.  Stateless; entropy is passed explicitly per call."""

    kind = "port"

    def __init__(self, path: str | os.PathLike | None = None):
        path = Path(path) if path else HERE / "liboracle.so"
        if not path.exists():
            build_oracle()
        self.lib = L = C.CDLL(str(path))
        L.orc_aes256_encrypt_block.argtypes = [_u8p, _u8p, _u8p]
        L.orc_stream.argtypes = [_u8p, C.c_uint64, _u8p, C.c_size_t]
        L.orc_urandomb.argtypes = [_u8p, C.c_uint64, C.c_size_t, _u64p, _i32p]
        L.orc_modq.argtypes = [_u64p, C.c_int, _u64p, _i32p]
        L.orc_ct_import.argtypes = [_u8p, C.c_uint64, _u8p, _u64p]
        L.orc_ct_export.argtypes = [_u64p, _u8p]
        L.orc_ct_mul_ui.argtypes = [_u64p, C.c_uint64, _u64p]
        L.orc_ct_addmul_ui.argtypes = [_u64p, _u64p, C.c_uint64]
        L.orc_ct_add.argtypes = [_u64p, _u64p, _u64p]
        L.orc_eval_poly.argtypes = [_u8p, C.c_uint64, _u8p, _u64p, C.c_size_t, _u64p]
        L.orc_ct_smudge.argtypes = [_u64p, _u8p]
        L.orc_ct_smudge.restype = C.c_int
        L.orc_key_gen.argtypes = [_u8p, _u64p]
        L.orc_encrypt.argtypes = [_u8p, C.c_uint64, _u64p, _u64p, C.c_size_t, _u8p, _u8p, _u64p]
        L.orc_dotp.argtypes = [_u64p, _u64p, C.c_size_t, _u64p]
        L.orc_decrypt.argtypes = [_u64p, _u64p, C.c_int]
        L.orc_decrypt.restype = C.c_uint64

    # -- AES / stream
    def aes256_block(self, key: bytes, block: bytes) -> bytes:
        k, b, o = _u8(key), _u8(block), np.zeros(16, np.uint8)
        self.lib.orc_aes256_encrypt_block(_p8(k), _p8(b), _p8(o))
        return o.tobytes()

    def stream(self, seed, offset: int, nbytes: int) -> np.ndarray:
        s, out = _seed(seed), np.zeros(nbytes, np.uint8)
        self.lib.orc_stream(_p8(s), offset, _p8(out), nbytes)
        return out

    def urandomb(self, seed, offset: int, nbits: int):
        s, out, siz = _seed(seed), np.zeros(LIMBS, np.uint64), np.zeros(1, np.int32)
        self.lib.orc_urandomb(_p8(s), offset, nbits, _p64(out), _p32(siz))
        return out, int(siz[0])

    def modq(self, limbs):
        x, out, siz = _u64(limbs), np.zeros(LIMBS, np.uint64), np.zeros(1, np.int32)
        self.lib.orc_modq(_p64(x), x.size, _p64(out), _p32(siz))
        return out, int(siz[0])

    # -- ciphertexts
    def ct_import(self, seed, offset: int, b92) -> np.ndarray:
        s, b, out = _seed(seed), _u8(b92), np.zeros((NC, LIMBS), np.uint64)
        self.lib.orc_ct_import(_p8(s), offset, _p8(b), _p64(out))
        return out

    def ct_export(self, ct) -> np.ndarray:
        c, out = _u64(ct), np.zeros(CT_BYTES, np.uint8)
        self.lib.orc_ct_export(_p64(c), _p8(out))
        return out

    def ct_mul_ui(self, a, b: int) -> np.ndarray:
        x, out = _u64(a), np.zeros((NC, LIMBS), np.uint64)
        self.lib.orc_ct_mul_ui(_p64(x), b, _p64(out))
        return out

    def ct_addmul_ui(self, rop, a, b: int) -> np.ndarray:
        r, x = _u64(rop).copy(), _u64(a)
        self.lib.orc_ct_addmul_ui(_p64(r), _p64(x), b)
        return r

    def ct_add(self, a, b) -> np.ndarray:
        x, y, out = _u64(a), _u64(b), np.zeros((NC, LIMBS), np.uint64)
        self.lib.orc_ct_add(_p64(x), _p64(y), _p64(out))
        return out

    def eval_poly(self, seed, offset: int, c8, coeffs, rop=None) -> np.ndarray:
        s, c, h = _seed(seed), _u8(c8), _u64(coeffs)
        d = h.size
        assert c.size == d * CT_BYTES
        r = np.zeros((NC, LIMBS), np.uint64) if rop is None else _u64(rop).copy()
        self.lib.orc_eval_poly(_p8(s), offset, _p8(c), _p64(h), d, _p64(r))
        return r

    def ct_smudge(self, ct, entropy81):
        c, e = _u64(ct).copy(), _u8(entropy81)
        assert e.size == SMUDGE_BYTES + 1
        neg = self.lib.orc_ct_smudge(_p64(c), _p8(e))
        return c, bool(neg)

    # -- keys / encrypt / decrypt
    def key_gen(self, entropy) -> np.ndarray:
        e, sk = _u8(entropy), np.zeros((N, LIMBS), np.uint64)
        assert e.size == N * CT_BYTES
        self.lib.orc_key_gen(_p8(e), _p64(sk))
        return sk

    def encrypt(self, seed, offset: int, sk, m, entropy, want_ct: bool = False):
        s, k, mm, e = _seed(seed), _u64(sk), _u64(m), _u8(entropy)
        cnt = mm.size
        assert e.size == cnt * (NOISE_BYTES + 1)
        out = np.zeros((cnt, CT_BYTES), np.uint8)
        cts = np.zeros((cnt, NC, LIMBS), np.uint64) if want_ct else None
        self.lib.orc_encrypt(_p8(s), offset, _p64(k), _p64(mm), cnt, _p8(e), _p8(out),
                             _p64(cts) if want_ct else None)
        return (out, cts) if want_ct else out

    def dotp(self, a, b) -> np.ndarray:
        x, y, out = _u64(a), _u64(b), np.zeros(LIMBS, np.uint64)
        self.lib.orc_dotp(_p64(x), _p64(y), x.shape[0], _p64(out))
        return out

    def decrypt(self, sk, ct, b_negative: bool = False) -> int:
        k, c = _u64(sk), _u64(ct)
        return int(self.lib.orc_decrypt(_p64(k), _p64(c), int(b_negative)))


class Reference:
    """The compiled reference.  Entropy-consuming calls read from the stream given to
    :meth:`set_entropy` (process-global inside the .so), in the reference's own call order."""

    kind = "reference"

    def __init__(self, D: int = 256, M: int = 64, path: str | os.PathLike | None = None):
        if path is None:
            path = build_reference(D, M)
        if path is None or not Path(path).exists():
            raise FileNotFoundError(f"reference build for D={D}, M={M} not available")
        self.lib = L = C.CDLL(str(path))
        self._bind(D, M)

    def _bind(self, D: int, M: int):
        L = self.lib
        self._entropy = None
        L.ref_param.argtypes = [C.c_char_p]
        L.ref_param.restype = C.c_uint64
        if hasattr(L, "ref_set_instance"):
            L.ref_set_instance.argtypes = [C.c_size_t, C.c_size_t]
            L.ref_set_instance(D, M)
        self.D, self.M = int(L.ref_param(b"D")), int(L.ref_param(b"M"))
        assert (self.D, self.M) == (D, M), "instance size of the loaded library differs from the request"
        L.ref_set_entropy.argtypes = [_u8p, C.c_size_t]
        L.ref_entropy_consumed.restype = C.c_size_t
        L.ref_entropy_calls.restype = C.c_uint64
        L.ref_stream.argtypes = [_u8p, C.c_uint64, _u8p, C.c_size_t]
        L.ref_stream_chunked.argtypes = [_u8p, C.c_uint64, _u8p, C.c_size_t, C.c_size_t]
        L.ref_urandomb.argtypes = [_u8p, C.c_uint64, C.c_size_t, _u64p, _i32p]
        L.ref_modq.argtypes = [_u64p, C.c_int, _u64p, _i32p]
        L.ref_ct_import.argtypes = [_u8p, C.c_uint64, _u8p, _u64p, _i32p]
        L.ref_ct_export.argtypes = [_u64p, _u8p]
        L.ref_eval_poly.argtypes = [_u8p, C.c_uint64, _u8p, _u64p, C.c_size_t, _u64p, _i32p]
        L.ref_time_eval_poly.argtypes = [_u8p, C.c_uint64, _u8p, _u64p, C.c_size_t, _u64p, _dblp, _dblp]
        L.ref_time_eval_poly.restype = C.c_double
        L.ref_ct_mul_ui.argtypes = [_u64p, C.c_uint64, _u64p, _i32p]
        L.ref_ct_addmul_ui.argtypes = [_u64p, _u64p, C.c_uint64, _i32p]
        L.ref_ct_add.argtypes = [_u64p, _u64p, _u64p, _i32p]
        L.ref_ct_smudge.argtypes = [_u64p, _i32p]
        L.ref_key_gen.argtypes = [_u64p]
        L.ref_encrypt.argtypes = [_u8p, C.c_uint64, _u64p, _u64p, C.c_size_t, _u8p, _u64p]
        L.ref_decrypt.argtypes = [_u64p, _u64p]
        L.ref_decrypt.restype = C.c_uint64
        L.ref_dotp.argtypes = [_u64p, _u64p, C.c_size_t, _u64p, _i32p]
        L.ref_random_ssp.argtypes = [_u8p, _u64p, C.c_size_t]
        L.ref_setup.argtypes = [_u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u64p, _u64p]
        L.ref_prover.argtypes = [_u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u64p, C.c_size_t, _u64p, _i32p]
        L.ref_verifier.argtypes = [_u8p, _u64p, _u64p, _u64p]
        L.ref_verifier.restype = C.c_int
        L.ref_benchmark_snark.argtypes = [_dblp]
        L.ref_benchmark_snark.restype = C.c_int

    def param(self, name: str) -> int:
        return int(self.lib.ref_param(name.encode()))

    # -- entropy
    def set_entropy(self, data) -> None:
        self._entropy = _u8(data)  # keep alive: the .so holds a raw pointer
        self.lib.ref_set_entropy(_p8(self._entropy), self._entropy.size)

    def clear_entropy(self) -> None:
        self._entropy = None
        self.lib.ref_set_entropy(None, 0)

    def entropy_consumed(self) -> int:
        return int(self.lib.ref_entropy_consumed())

    # -- AES / stream
    def stream(self, seed, offset: int, nbytes: int, chunk: int | None = None) -> np.ndarray:
        s, out = _seed(seed), np.zeros(nbytes, np.uint8)
        if chunk:
            self.lib.ref_stream_chunked(_p8(s), offset, _p8(out), nbytes, chunk)
        else:
            self.lib.ref_stream(_p8(s), offset, _p8(out), nbytes)
        return out

    def urandomb(self, seed, offset: int, nbits: int):
        s, out, siz = _seed(seed), np.zeros(LIMBS, np.uint64), np.zeros(1, np.int32)
        self.lib.ref_urandomb(_p8(s), offset, nbits, _p64(out), _p32(siz))
        return out, int(siz[0])

    def modq(self, limbs):
        x, out, siz = _u64(limbs), np.zeros(LIMBS, np.uint64), np.zeros(1, np.int32)
        self.lib.ref_modq(_p64(x), x.size, _p64(out), _p32(siz))
        return out, int(siz[0])

    # -- ciphertexts
    def ct_import(self, seed, offset: int, b92) -> np.ndarray:
        s, b, out = _seed(seed), _u8(b92), np.zeros((NC, LIMBS), np.uint64)
        self.lib.ref_ct_import(_p8(s), offset, _p8(b), _p64(out), None)
        return out

    def ct_export(self, ct) -> np.ndarray:
        c, out = _u64(ct), np.zeros(CT_BYTES, np.uint8)
        self.lib.ref_ct_export(_p64(c), _p8(out))
        return out

    def ct_mul_ui(self, a, b: int) -> np.ndarray:
        x, out = _u64(a), np.zeros((NC, LIMBS), np.uint64)
        self.lib.ref_ct_mul_ui(_p64(x), b, _p64(out), None)
        return out

    def ct_addmul_ui(self, rop, a, b: int) -> np.ndarray:
        r, x = _u64(rop).copy(), _u64(a)
        self.lib.ref_ct_addmul_ui(_p64(r), _p64(x), b, None)
        return r

    def ct_add(self, a, b) -> np.ndarray:
        x, y, out = _u64(a), _u64(b), np.zeros((NC, LIMBS), np.uint64)
        self.lib.ref_ct_add(_p64(x), _p64(y), _p64(out), None)
        return out

    def eval_poly(self, seed, offset: int, c8, coeffs, rop=None) -> np.ndarray:
        s, c, h = _seed(seed), _u8(c8), _u64(coeffs)
        d = h.size
        assert c.size == d * CT_BYTES
        r = np.zeros((NC, LIMBS), np.uint64) if rop is None else _u64(rop).copy()
        self.lib.ref_eval_poly(_p8(s), offset, _p8(c), _p64(h), d, _p64(r), None)
        return r

    def time_eval_poly(self, seed, offset: int, c8, coeffs, split: bool = False):
        """Seconds inside eval_poly (clock_gettime); with split=True also the per-ciphertext
        ct_import / ct_addmul_ui seconds."""
        s, c, h = _seed(seed), _u8(c8), _u64(coeffs)
        ti, ta = C.c_double(0), C.c_double(0)
        r = np.zeros((NC, LIMBS), np.uint64)
        t = self.lib.ref_time_eval_poly(_p8(s), offset, _p8(c), _p64(h), h.size, _p64(r),
                                        C.byref(ti) if split else None, C.byref(ta) if split else None)
        return (float(t), r, float(ti.value), float(ta.value)) if split else (float(t), r)

    def ct_smudge(self, ct):
        """Consumes 80 + 1 entropy bytes.  Returns (ct, negative?)."""
        c, siz = _u64(ct).copy(), np.zeros(NC, np.int32)
        self.lib.ref_ct_smudge(_p64(c), _p32(siz))
        return c, bool(siz[N] < 0)

    # -- keys / encrypt / decrypt
    def key_gen(self) -> np.ndarray:
        """Consumes 1470 * 92 entropy bytes."""
        sk = np.zeros((N, LIMBS), np.uint64)
        self.lib.ref_key_gen(_p64(sk))
        return sk

    def encrypt(self, seed, offset: int, sk, m, want_ct: bool = False):
        """Consumes (69 + 1) entropy bytes per message."""
        s, k, mm = _seed(seed), _u64(sk), _u64(m)
        cnt = mm.size
        out = np.zeros((cnt, CT_BYTES), np.uint8)
        cts = np.zeros((cnt, NC, LIMBS), np.uint64) if want_ct else None
        self.lib.ref_encrypt(_p8(s), offset, _p64(k), _p64(mm), cnt, _p8(out),
                             _p64(cts) if want_ct else None)
        return (out, cts) if want_ct else out

    def dotp(self, a, b) -> np.ndarray:
        x, y, out = _u64(a), _u64(b), np.zeros(LIMBS, np.uint64)
        self.lib.ref_dotp(_p64(x), _p64(y), x.shape[0], _p64(out), None)
        return out

    def decrypt(self, sk, ct) -> int:
        k, c = _u64(sk), _u64(ct)
        return int(self.lib.ref_decrypt(_p64(k), _p64(c)))

    # -- full SNARK (instance size fixed at build time: self.D, self.M)
    def ssp_size(self) -> int:
        return self.param("SSP_SIZE")

    def random_ssp(self):
        """Consumes M/8 + M*8*D entropy bytes.  Returns (ssp blob, witness limbs)."""
        ssp = np.zeros(self.ssp_size(), np.uint8)
        wl = np.zeros((self.M + 63) // 64, np.uint64)
        self.lib.ref_random_ssp(_p8(ssp), _p64(wl), wl.size)
        return ssp, wl

    def setup(self, ssp):
        """crs_init + setup.  Returns dict(seed, s, as_, v, t, alpha, beta, s_point, sk)."""
        D, M = self.D, self.M
        seed = np.zeros(40, np.uint8)
        cs, cas = np.zeros((D, CT_BYTES), np.uint8), np.zeros((D, CT_BYTES), np.uint8)
        cv, ctt = np.zeros((M, CT_BYTES), np.uint8), np.zeros(CT_BYTES, np.uint8)
        abs_ = np.zeros(3, np.uint64)
        sk = np.zeros((N, LIMBS), np.uint64)
        s = _u8(ssp)
        self.lib.ref_setup(_p8(s), _p8(seed), _p8(cs), _p8(cas), _p8(cv), _p8(ctt), _p64(abs_), _p64(sk))
        return dict(seed=seed, s=cs, as_=cas, v=cv, t=ctt, alpha=int(abs_[0]), beta=int(abs_[1]),
                    s_point=int(abs_[2]), sk=sk)

    def prover(self, ssp, crs: dict, witness_limbs):
        """Consumes 8 + 5*(80+1) entropy bytes.  Returns (proof (5,1471,12) uint64, siz (5,1471))."""
        s, wl = _u8(ssp), _u64(witness_limbs)
        proof = np.zeros((5, NC, LIMBS), np.uint64)
        siz = np.zeros((5, NC), np.int32)
        self.lib.ref_prover(_p8(s), _p8(_u8(crs["seed"])), _p8(_u8(crs["s"])), _p8(_u8(crs["as_"])),
                            _p8(_u8(crs["v"])), _p8(_u8(crs["t"])), _p64(wl), wl.size, _p64(proof),
                            _p32(siz))
        return proof, siz

    def verifier(self, ssp, crs: dict, proof) -> bool:
        s = _u8(ssp)
        abs_ = np.array([crs["alpha"], crs["beta"], crs["s_point"]], np.uint64)
        return bool(self.lib.ref_verifier(_p8(s), _p64(abs_), _p64(_u64(crs["sk"])), _p64(_u64(proof))))

    def benchmark_snark(self):
        """(setup_s, prover_s, verifier_s, accept) with OS entropy, as benchmark_snark.c:56-82."""
        secs = (C.c_double * 3)()
        ok = self.lib.ref_benchmark_snark(secs)
        return float(secs[0]), float(secs[1]), float(secs[2]), bool(ok)


class DropIn(Reference):
    """The PRODUCT behind the reference's own C interface: ref_shim.c compiled against include/mangiafuoco and
    linked with libmangiafuoco_b200.so (GPU kernels).  Same methods and formats as :class:`Reference`, so a
    parity test is `DropIn(...).f(x) == Reference(...).f(x)`.  The instance size is a run-time value here."""

    kind = "dropin"

    def __init__(self, D: int = 256, M: int = 64):
        path = build_dropin()
        self.lib = C.CDLL(str(path))
        self._bind(D, M)
        self.lib.ref_gpu_launches.restype = C.c_uint64
        if hasattr(self.lib, "ref_prover_resident"):
            self.lib.ref_prover_resident.argtypes = self.lib.ref_prover.argtypes

    def set_instance(self, D: int, M: int):
        self.lib.ref_set_instance(D, M)
        self.D, self.M = D, M

    def gpu_launches(self) -> int:
        return int(self.lib.ref_gpu_launches())

    def prover_resident(self, ssp, crs: dict, witness_limbs):
        """prover() after mf_crs_make_resident(): the s / as regions are expanded into HBM first."""
        s, wl = _u8(ssp), _u64(witness_limbs)
        proof = np.zeros((5, NC, LIMBS), np.uint64)
        siz = np.zeros((5, NC), np.int32)
        self.lib.ref_prover_resident(_p8(s), _p8(_u8(crs["seed"])), _p8(_u8(crs["s"])), _p8(_u8(crs["as_"])),
                                     _p8(_u8(crs["v"])), _p8(_u8(crs["t"])), _p64(wl), wl.size, _p64(proof), _p32(siz))
        return proof, siz
