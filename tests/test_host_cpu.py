"""CPU: host-side logic of the drop-in layer that needs no GPU — entropy routing, byte -> mpz sampling, modq,
the SSP generator / wire format, F_p[x] arithmetic, ct_export, ct_smudge — through the same flat-buffer shim the
GPU tests use (oracle/_dropin/libmfdropin.so = oracle/ref_shim.c over libmangiafuoco_b200.so)."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import sha, xof
from oracle.loader import CT_BYTES, N, P

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((Path(__file__).parent / "golden" / "vectors.json").read_text())


@pytest.fixture(scope="module")
def dropin():
    from oracle.loader import DropIn
    return DropIn(64, 16)


def hexs(a) -> str:
    return np.ascontiguousarray(a).tobytes().hex()


def test_random_ssp_golden(dropin):
    g = GOLD["snark_d64_m16"]
    dropin.set_instance(g["D"], g["M"])
    assert dropin.ssp_size() == g["D"] * 8 * (g["M"] + 3)
    dropin.set_entropy(xof("snark-entropy-d64-m16", g["entropy_bytes"]))
    ssp, wit = dropin.random_ssp()
    assert dropin.entropy_consumed() == g["M"] // 8 + g["M"] * 8 * g["D"]
    dropin.clear_entropy()
    assert sha(ssp) == g["ssp_sha"] and hexs(wit) == g["witness"]


def test_key_gen_matches_oracle(dropin, oracle):
    ent = xof("lwe-entropy", N * CT_BYTES)
    dropin.set_entropy(ent)
    sk = dropin.key_gen()
    assert dropin.entropy_consumed() == N * CT_BYTES
    dropin.clear_entropy()
    assert sha(sk) == GOLD["lwe"]["sk_sha"] and np.array_equal(sk, oracle.key_gen(ent))


def test_modq_golden(dropin):
    for m in GOLD["modq"]:
        x = np.frombuffer(bytes.fromhex(m["x"]), "<u8")
        out, siz = dropin.modq(x)
        assert hexs(out) == m["out"] and siz == m["siz"]


def test_smudge_and_export_match_oracle(dropin, oracle):
    ct = xof("host-ct", 1471 * 96).view("<u8").reshape(1471, 12).copy()
    ct[:, 11] = 0
    for k in range(4):
        e81 = xof(f"host-smudge{k}", 81)
        dropin.set_entropy(e81)
        got, neg = dropin.ct_smudge(ct)
        assert dropin.entropy_consumed() == 81
        dropin.clear_entropy()
        want, wneg = oracle.ct_smudge(ct, e81)
        assert np.array_equal(got, want) and neg == wneg
    assert np.array_equal(dropin.ct_export(ct), oracle.ct_export(ct))


def test_reference_ssp_program_against_the_dropin():
    """test_ssp.c of the reference (wire round-trip; t | v^2 - 1 via nmod_poly pow/rem), unmodified, on the drop-in."""
    exe = ROOT / "oracle" / "_ref" / "dropin_test_ssp"
    if not exe.exists():
        pytest.skip("built only where the reference sources are present")
    assert subprocess.run([str(exe)], timeout=300).returncode == 0


def test_python_binding_argument_checks():
    from c_lwe_snarks_b200.api import _seed
    with pytest.raises(ValueError):
        _seed(b"short")
    import c_lwe_snarks_b200 as m
    assert m.ALGO_BYTES_PER_MAC == 1471 * 88 and m.PLANAR_U64 * 8 == 129536 and m.CTR_CT == 135240 and m.P == P


def test_crs_file_roundtrip(tmp_path):
    """mf_crs_write / mf_crs_read: header + seed + records in stream order (s, as, t, v); size mismatch is refused."""
    import ctypes as C

    from c_lwe_snarks_b200.snark import Snark
    D, M = 32, 16
    a = Snark(D, M)
    a.lib.crs_init(C.byref(a.crs))
    a._crs_live = True
    for ptr, n, label in ((a.crs.s, 92 * D, "s"), (a.crs.as_, 92 * D, "as"), (a.crs.t, 92, "t"), (a.crs.v, 92 * (M - 1), "v")):
        C.memmove(ptr, xof("crs-" + label, n).tobytes(), n)
    path = tmp_path / "crs.mfuoco"
    a.save_crs(path)
    raw = path.read_bytes()
    seed, s_, as_, t_, v_ = a.crs_records()
    assert raw[:8] == b"MFUOCO1\0" and int.from_bytes(raw[8:16], "little") == D and int.from_bytes(raw[16:24], "little") == M
    assert raw[24:64] == seed and raw[64:] == s_.tobytes() + as_.tobytes() + t_.tobytes() + v_.tobytes()
    b = Snark(D, M)
    b.load_crs(path)
    assert b.crs_records()[0] == seed and all(np.array_equal(x, y) for x, y in zip(a.crs_records()[1:], b.crs_records()[1:]))
    c = Snark(2 * D, M)
    with pytest.raises(OSError):
        c.load_crs(path)
    a.lib.mf_set_instance(D, M)
    for sn in (a, b, c):
        sn.lib.mf_set_instance(sn.D, sn.M)
        sn.close()


def test_proof_file_roundtrip(tmp_path):
    import ctypes as C

    from c_lwe_snarks_b200.snark import N, Snark
    sn = Snark(32, 16)
    sn.lib.proof_init(C.byref(sn.proof))
    sn._proof_live = True
    imp = getattr(sn.gmp, "__gmpz_import")
    imp.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_size_t, C.c_int, C.c_size_t, C.c_void_p]
    neg = getattr(sn.gmp, "__gmpz_neg")
    cmp_ = getattr(sn.gmp, "__gmpz_cmp")
    elems = [sn.proof.h, sn.proof.hat_h, sn.proof.hat_v, sn.proof.v_w, sn.proof.b_w]
    for k, el in enumerate(elems):
        data = xof(f"proof-{k}", (N + 1) * 88)
        for i in range(0, N + 1, 97):
            imp(C.byref(el[i]), 88, -1, 1, -1, 0, data[i * 88:(i + 1) * 88].ctypes.data)
        imp(C.byref(el[N]), 88, -1, 1, -1, 0, data[N * 88:].ctypes.data)
    neg(C.byref(elems[3][N]), C.byref(elems[3][N]))  # a negative b, as ct_smudge can leave it
    path = tmp_path / "proof.bin"
    sn.save_proof(path)
    other = Snark(32, 16)
    other.load_proof(path)
    for a, b in zip(elems, [other.proof.h, other.proof.hat_h, other.proof.hat_v, other.proof.v_w, other.proof.b_w]):
        for i in list(range(0, N + 1, 97)) + [1, N]:
            assert cmp_(C.byref(a[i]), C.byref(b[i])) == 0
    assert other.proof.v_w[N].size < 0
    sn.close()
    other.close()


def test_gmp_impl_macros_of_the_reference_compile_and_work(tmp_path):
    """gmp-impl.h:15-29 (UNLIKELY, MPN_NORMALIZE, MPZ_NEWALLOC) and entropy.h:56 (mpz_entropy_init) are part of the
    reference's header surface: a program using them compiles against the drop-in headers, links and runs."""
    src = tmp_path / "macros.c"
    src.write_text(r'''
#include "gmp-impl.h"
#include "entropy.h"
#include <stdio.h>
int main(void) {
  mpz_t z;
  mpz_init(z);
  mp_ptr p = MPZ_NEWALLOC(z, 5);            /* grows the limb array */
  p[0] = 7; p[1] = 0; p[2] = 9; p[3] = 0; p[4] = 0;
  mp_size_t n = 5;
  MPN_NORMALIZE(p, n);                      /* strips the two zero high limbs */
  SIZ(z) = (int)n;
  if (UNLIKELY(n != 3)) return 1;
  if (ALLOC(z) < 5 || PTR(z) != p) return 2;
  if (MPZ_NEWALLOC(z, 2) != p) return 3;    /* no reallocation when it already fits */
  mpz_entropy_init();
  mpz_clear(z);
  puts("ok");
  return 0;
}
''')
    lib = ROOT / "c_lwe_snarks_b200" / "lib"
    exe = tmp_path / "macros"
    cmd = ["gcc", "-std=gnu11", "-O1", f"-I{ROOT / 'include' / 'mangiafuoco'}", f"-I{ROOT / 'include' / 'compat'}",
           f"-I{ROOT / 'include'}", str(src), f"-L{lib}", "-lmangiafuoco_b200", "-lmfb200", "-l:libgmp.so.10",
           f"-Wl,-rpath,{lib}", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == "ok", (r.returncode, r.stdout, r.stderr[-500:])


def test_generic_encrypt_tiling_plan():
    """Host logic of mfb_encrypt_generic_dev (BASELINE configs[4]; no device needed): the tiles cover the ciphertext, their
    (padded) keystream fits one 45 KB buffer at any stream alignment, the padded layout is taken exactly when it lowers
    the bank-conflict degree of the consumers' reads, its reciprocal is exact, and whole-round tiles never need more
    consumer rounds than the balanced split."""
    import ctypes as C
    from math import gcd

    from c_lwe_snarks_b200.api import load_library
    lib = load_library()
    KS_TILE_BYTES, CONSUMERS = 490 * 92, 128
    KS_BUF_BYTES = ((KS_TILE_BYTES + 30) // 16) * 16 + 16
    for ctb in range(32, 129, 4):
        for n in (1, 33, 65, 700, 1024, 1246, 1470, 1600, 2047, 4096):
            tile, ntiles, wb, inv = C.c_int(), C.c_int(), C.c_int(), C.c_uint32()
            assert lib.mfb_encrypt_generic_plan(n, ctb, C.byref(tile), C.byref(ntiles), C.byref(wb), C.byref(inv)) == 0
            tile, ntiles, wb, inv = tile.value, ntiles.value, wb.value, inv.value
            assert tile >= 1 and (ntiles - 1) * tile < n <= ntiles * tile
            words = ctb // 4
            want_pad = ctb % 16 == 0 and gcd(words + 4, 32) < gcd(words, 32)
            assert (wb != 0) == want_pad and (wb == 0 or wb == ctb // 16)
            slot = ctb + (16 if wb else 0)
            for delta in (0, 15):  # the tile's first coordinate may start anywhere inside an AES block
                nblk = (delta + tile * ctb + 15) // 16
                last_slot = (nblk - 1) + ((nblk - 1) // wb if wb else 0)
                assert 16 * (last_slot + 1) <= KS_BUF_BYTES
                assert (delta & ~3) + (tile - 1) * slot + ctb + 16 + 4 <= KS_BUF_BYTES  # the consumers' furthest read
            if wb:
                assert all((b * inv) >> 16 == b // wb for b in range(0, KS_BUF_BYTES // 16 + 1))
            max_tile = KS_TILE_BYTES // slot
            nt_bal = -(-n // max_tile)
            t_bal = -(-n // nt_bal)
            rounds = lambda tl, nt: (nt - 1) * -(-tl // CONSUMERS) + -(-(n - (nt - 1) * tl) // CONSUMERS)  # noqa: E731
            assert rounds(tile, ntiles) <= rounds(t_bal, nt_bal)
    bad = C.c_int()
    assert lib.mfb_encrypt_generic_plan(1470, 90, C.byref(bad), C.byref(bad), C.byref(bad), C.byref(C.c_uint32())) != 0
