"""CPU: pins the oracle (oracle/mf_oracle.c, our plain-C restatement) against
  (1) the FIPS-197 known answer and the golden vectors the COMPILED REFERENCE emitted (tests/golden/vectors.json),
  (2) the compiled reference itself, function by function, when oracle/_ref is present (it is built from the
      reference's own sources by oracle/Makefile and travels with the repository snapshot),
  (3) the small-case Python restatement of the protocol layer (oracle/snark_py.py) for the full SNARK.
"""
import json
from pathlib import Path

import numpy as np
import pytest

from conftest import SEED, sha, xof, xof_records, xof_scalars
from oracle.loader import CT_BYTES, CTR_CT, N, NOISE_BYTES, P

GOLD = json.loads((Path(__file__).parent / "golden" / "vectors.json").read_text())


def hexs(a) -> str:
    return np.ascontiguousarray(a).tobytes().hex()


# ------------------------------------------------------------------ (1) golden vectors
def test_aes_fips197(oracle):
    g = GOLD["aes_fips197_c3"]
    assert oracle.aes256_block(bytes.fromhex(g["key"]), bytes.fromhex(g["pt"])).hex() == g["ct"]


def test_probe_block0(oracle):
    # seed 00..27 -> keystream block 0
    assert oracle.stream(SEED, 0, 16).tobytes().hex() == "8477f45516027713a26a881ae67882bf"


def test_stream_golden(oracle):
    for s in GOLD["stream"]:
        assert hexs(oracle.stream(SEED, s["offset"], s["n"])) == s["hex"]
    assert sha(oracle.stream(SEED, 3 * CTR_CT, CTR_CT)) == GOLD["stream_ct3_sha"]


def test_urandomb_modq_golden(oracle):
    for u in GOLD["urandomb"]:
        limbs, siz = oracle.urandomb(SEED, u["offset"], u["nbits"])
        assert hexs(limbs) == u["limbs"] and siz == u["siz"]
    for m in GOLD["modq"]:
        x = np.frombuffer(bytes.fromhex(m["x"]), "<u8")
        out, siz = oracle.modq(x)
        assert hexs(out) == m["out"] and siz == m["siz"]


def test_modq_is_mod_2_704(oracle):
    # modq(2^735 + 2^704 + 5) == 5
    x = np.zeros(12, np.uint64)
    x[0], x[11] = 5, (1 << 31) | 1
    out, siz = oracle.modq(x)
    assert out[0] == 5 and not out[1:].any() and siz == 1


def test_ct_import_golden(oracle):
    b = xof_records("import-b", 1)[0]
    for name in ("ct_import_even", "ct_import_odd"):
        g = GOLD[name]
        ct = oracle.ct_import(SEED, g["offset"], b)
        assert sha(ct) == g["sha"] and hexs(ct[0]) == g["a0"] and hexs(ct[1470]) == g["b_limbs"]


def test_eval_poly_golden(oracle):
    d = 12
    c8, h = xof_records("eval-c8", d), xof_scalars("eval-h", d)
    h[3], h[4] = 0, P - 1
    acc = oracle.eval_poly(SEED, 3 * CTR_CT, c8, h)
    g = GOLD["eval_poly"]
    assert sha(acc) == g["sha"] and hexs(acc[0]) == g["c0"] and hexs(acc[777]) == g["c777"] and hexs(acc[1470]) == g["b"]
    assert sha(oracle.eval_poly(SEED, 3 * CTR_CT, c8, h, rop=acc)) == GOLD["eval_poly_accumulate"]["sha"]
    x = oracle.ct_import(SEED, 0, c8[0])
    y = oracle.ct_import(SEED, CTR_CT, c8[1])
    assert sha(oracle.ct_mul_ui(x, int(h[0]))) == GOLD["ct_ops"]["mul"]
    assert sha(oracle.ct_add(x, y)) == GOLD["ct_ops"]["add"]
    assert sha(oracle.ct_addmul_ui(oracle.ct_mul_ui(x, 7), y, int(h[1]))) == GOLD["ct_ops"]["addmul"]


def test_lwe_golden(oracle):
    g = GOLD["lwe"]
    cnt = g["count"]
    ent = xof("lwe-entropy", N * CT_BYTES + cnt * (NOISE_BYTES + 1))
    m = np.array(g["m"], np.uint64)
    sk = oracle.key_gen(ent[: N * CT_BYTES])
    assert sha(sk) == g["sk_sha"]
    recs, cts = oracle.encrypt(SEED, g["offset"], sk, m, ent[N * CT_BYTES:], want_ct=True)
    assert hexs(recs) == g["records"]
    assert [oracle.decrypt(sk, cts[k]) for k in range(cnt)] == g["m"]
    assert hexs(oracle.dotp(cts[0][:N], sk)) == g["dotp0"]
    for k, s in enumerate(GOLD["smudge"]):
        out, neg = oracle.ct_smudge(cts[0], xof(f"smudge{k}", 81))
        assert hexs(out[N]) == s["b"] and neg == s["negative"]
        if s["dec"] is not None:
            assert oracle.decrypt(sk, out) == s["dec"]


def test_full_snark_python_restatement_golden(oracle):
    """oracle/snark_py.py (protocol layer over the oracle primitives) reproduces the reference's D=64, M=16 run."""
    from oracle import snark_py as sp
    g = GOLD["snark_d64_m16"]
    D, M = g["D"], g["M"]
    ent = sp.Entropy(xof("snark-entropy-d64-m16", g["entropy_bytes"]))
    ssp, witness = sp.random_ssp(D, M, ent)
    assert sha(ssp) == g["ssp_sha"]
    crs = sp.setup(oracle, ssp, D, M, ent)
    assert hexs(crs["seed"]) == g["seed"] and (crs["alpha"], crs["beta"], crs["s_point"]) == (g["alpha"], g["beta"], g["s_point"])
    assert sha(crs["s"]) == g["crs_s_sha"] and sha(crs["as_"]) == g["crs_as_sha"] and hexs(crs["t"]) == g["crs_t"]
    assert sha(crs["v"][: M - 1]) == g["crs_v_sha"]
    proof, neg = sp.prover(oracle, ssp, crs, witness, D, M, ent)
    assert ent.pos == g["entropy_bytes"]
    for k in range(5):
        assert sha(proof[k]) == g["proof_sha"][k] and neg[k] == g["proof_negative"][k]
    assert sp.verifier(oracle, ssp, crs, proof, D, neg) == g["accept"]
    bad = proof.copy()
    bad[0, N, 0] ^= np.uint64(1 << 40)
    assert sp.verifier(oracle, ssp, crs, bad, D, neg) == g["accept_tampered"]


# ------------------------------------------------------------------ (2) beside the compiled reference
@pytest.mark.parametrize("off,n", [(0, 1), (5, 100), (135240 * 3 + 7, 4096), (2**40 + 9, 77)])
def test_stream_vs_reference(oracle, reference, off, n):
    assert np.array_equal(oracle.stream(SEED, off, n), reference.stream(SEED, off, n))
    assert np.array_equal(oracle.stream(SEED, off, n), reference.stream(SEED, off, n, chunk=7))


def test_ciphertext_ops_vs_reference(oracle, reference):
    d, off = 9, 11 * CTR_CT + 8
    c8, h = xof_records("r-c8", d), xof_scalars("r-h", d)
    x, y = oracle.ct_import(SEED, off, c8[0]), oracle.ct_import(SEED, off + CTR_CT, c8[1])
    assert np.array_equal(x, reference.ct_import(SEED, off, c8[0]))
    assert np.array_equal(oracle.ct_export(x), reference.ct_export(x))
    rop0 = oracle.ct_mul_ui(y, 12345)
    assert np.array_equal(rop0, reference.ct_mul_ui(y, 12345))
    assert np.array_equal(oracle.ct_add(x, y), reference.ct_add(x, y))
    assert np.array_equal(oracle.ct_addmul_ui(rop0, x, P - 1), reference.ct_addmul_ui(rop0, x, P - 1))
    assert np.array_equal(oracle.eval_poly(SEED, off, c8, h, rop=rop0), reference.eval_poly(SEED, off, c8, h, rop=rop0))


def test_encrypt_decrypt_vs_reference(oracle, reference):
    cnt, off = 3, 2 * CTR_CT * 256 + CTR_CT
    ent = xof("r-ent", N * CT_BYTES + cnt * 70 + 81)
    m = xof_scalars("r-m", cnt)
    reference.set_entropy(ent)
    sk = reference.key_gen()
    recs, cts = reference.encrypt(SEED, off, sk, m, want_ct=True)
    sm, neg = reference.ct_smudge(cts[1])
    reference.clear_entropy()
    assert np.array_equal(sk, oracle.key_gen(ent[: N * CT_BYTES]))
    o_recs, o_cts = oracle.encrypt(SEED, off, sk, m, ent[N * CT_BYTES: N * CT_BYTES + cnt * 70], want_ct=True)
    assert np.array_equal(recs, o_recs) and np.array_equal(cts, o_cts)
    o_sm, o_neg = oracle.ct_smudge(cts[1], ent[-81:])
    assert np.array_equal(sm, o_sm) and neg == o_neg
    for k in range(cnt):
        assert oracle.decrypt(sk, cts[k]) == reference.decrypt(sk, cts[k]) == int(m[k])
    assert np.array_equal(oracle.dotp(cts[2][:N], sk), reference.dotp(cts[2][:N], sk))
