"""BASELINE.json's full sizes (2^16 and 2^20 ciphertexts) through size-independent properties.

The oracle needs ~1 ms per ciphertext, so at these sizes it checks SAMPLES (single ciphertexts picked out with unit
scalar vectors, sparse lincombs, sampled encryption records) while whole-array agreement comes from properties of
exact arithmetic mod 2^704 that hold whatever the size:

  * two independent kernels agree: fused AES+MAC (k_evalpoly) == resident TMA lincomb (k_lincomb) over all D;
  * split invariance: the sum over [0, D) == sum over [0, a) accumulated into the sum over [a, D) (eval_poly
    accumulates into rop, lwe.c:176-186), for ragged a;
  * linearity in the scalars: L(h1) + L(h2) == L(h1 + h2) mod 2^704;
  * two-vector pass == two single passes;
  * encode -> lincomb -> decode (the SNARK's own correctness argument, snark.c:157-215): with b_i = Enc(m_i),
    Dec(sum h_i CT_i) == sum h_i m_i mod p — setup's kernel, the prover's kernel and the verifier's kernel in one
    chain against plain integer arithmetic.
"""
import numpy as np
import pytest

from conftest import SEED, xof, xof_records, xof_scalars

pytestmark = pytest.mark.gpu

N, NC, L64, CT_BYTES, CTR_CT, P = 1470, 1471, 11, 92, 92 * 1470, 0xFFFFFFFB
D16 = 1 << 16
OFF = 3 * CTR_CT + 8  # not block aligned: every tile starts mid-block


@pytest.fixture(scope="module")
def ctx():
    import c_lwe_snarks_b200 as m
    c = m.Context(0)
    yield c
    c.close()


def wide(flat11):
    out = np.zeros(flat11.shape[:-1] + (12,), np.uint64)
    out[..., :11] = flat11
    return out


def to_int(flat):  # (NC, 11) u64 -> list of python ints
    return [int.from_bytes(np.ascontiguousarray(row).tobytes(), "little") for row in flat]


def add704(x, y):
    m = (1 << 704) - 1
    out = np.zeros((NC, L64), np.uint64)
    for c, (a, b) in enumerate(zip(to_int(x), to_int(y))):
        out[c] = np.frombuffer(((a + b) & m).to_bytes(88, "little"), "<u8")
    return out


@pytest.fixture(scope="module")
def big(ctx):
    """D = 2^16 ciphertexts resident in HBM (8.49 GB) + the reference results of the properties' left-hand sides."""
    c8 = xof_records("full-c8", D16)
    h = xof_scalars("full-h", D16)
    reg = ctx.region(SEED, OFF, c8)
    full = reg.lincomb(h)
    yield {"c8": c8, "h": h, "reg": reg, "full": full}
    reg.close()


def test_2_16_resident_equals_fused(ctx, big):
    fused = ctx.eval_poly(SEED, OFF, big["c8"], big["h"])
    assert np.array_equal(fused, big["full"])
    assert fused.any()


@pytest.mark.parametrize("cut", [1, 4097, D16 // 2, D16 - 3])
def test_2_16_split_invariance(ctx, big, cut):
    reg, h = big["reg"], big["h"]
    lo = reg.lincomb(h[:cut])
    both = reg.lincomb(h[cut:], first=cut, rop=lo)  # accumulates into rop like eval_poly
    assert np.array_equal(both, big["full"])


def test_2_16_linearity(ctx, big):
    reg, h = big["reg"], big["h"]
    h1 = h >> np.uint64(1)
    h2 = h - h1
    assert np.array_equal(add704(reg.lincomb(h1), reg.lincomb(h2)), big["full"])


def test_2_16_two_vector_pass(ctx, big):
    reg, h = big["reg"], big["h"]
    g = np.roll(h, 12345)
    r0, r1 = reg.lincomb2(h, g)
    assert np.array_equal(r0, big["full"])
    assert np.array_equal(r1, reg.lincomb(g))
    f0, f1 = ctx.eval_poly2(SEED, OFF, big["c8"], h, g)
    assert np.array_equal(f0, r0) and np.array_equal(f1, r1)


@pytest.mark.parametrize("i", [0, 1, 777, D16 // 2 + 1, D16 - 1])
def test_2_16_unit_vectors_pick_oracle_ciphertexts(ctx, oracle, big, i):
    e = np.zeros(D16, np.uint64)
    e[i] = 1
    want = oracle.ct_import(SEED, OFF + i * CTR_CT, big["c8"][i])  # (1471, 12), limb 11 = the dead top 32 bits
    got = big["reg"].lincomb(e)
    assert np.array_equal(got, want[:, :11])


def test_2_16_sparse_lincomb_vs_oracle(ctx, oracle, big):
    idx = np.unique(np.frombuffer(xof("full-sparse", 4 * 60), "<u4") % D16)
    want = np.zeros((NC, 12), np.uint64)
    for i in idx:
        want = oracle.eval_poly(SEED, OFF + int(i) * CTR_CT, big["c8"][i:i + 1], big["h"][i:i + 1], rop=want)
    e = np.zeros(D16, np.uint64)
    e[idx] = big["h"][idx]
    assert np.array_equal(wide(big["reg"].lincomb(e)), want)
    got = ctx.eval_poly(SEED, OFF, big["c8"], big["h"][idx], idx=idx.astype(np.uint32))
    assert np.array_equal(wide(got), want)


def test_2_16_encrypt_lincomb_decrypt_roundtrip(ctx, oracle):
    """setup's encryptions (k_encrypt) -> the prover's lincomb (k_expand + k_lincomb) -> the verifier's decryption
    (k_decrypt) over 2^16 ciphertexts: equals sum h_i m_i mod p computed with plain integers.  No wrap mod 2^704:
    sum h_i (e_i p + m_i) < 2^16 * 2^32 * 2^552 * 2^32."""
    sk = oracle.key_gen(xof("full-sk", N * CT_BYTES))
    m = xof_scalars("full-m", D16)
    hh = xof_scalars("full-hh", D16)
    ent = xof("full-ent", D16 * 70)
    recs = ctx.encrypt(SEED, OFF, sk[:, :11], m, ent)
    # sampled records against the oracle's regev_encrypt2 + ct_export
    for i in (0, 1, 31337, D16 - 1):
        want = oracle.encrypt(SEED, OFF + i * CTR_CT, sk, m[i:i + 1], ent[70 * i:70 * i + 70])
        assert np.array_equal(recs[i], want[0])
    reg = ctx.region(SEED, OFF, recs)
    try:
        acc = reg.lincomb(hh)
    finally:
        reg.close()
    dec = ctx.decrypt(sk[:, :11], acc[None])
    want = sum(int(a) * int(b) for a, b in zip(hh, m)) % P
    assert int(dec[0]) == want
    assert int(dec[0]) == oracle.decrypt(sk, wide(acc))


def test_2_20_fused_split_invariance(ctx):
    """2^20 ciphertexts (BASELINE configs[3]) through the fused kernel: one call over [0, 2^20) == 16 calls over
    consecutive 2^16 slices, each accumulating into the previous result."""
    D = 1 << 20
    c8 = xof_records("full20-c8", D)
    h = xof_scalars("full20-h", D)
    one = ctx.eval_poly(SEED, OFF, c8, h)
    acc = None
    for k in range(16):
        s = slice(k * D16, (k + 1) * D16)
        acc = ctx.eval_poly(SEED, OFF + k * D16 * CTR_CT, c8[s], h[s], rop=acc)
    assert np.array_equal(acc, one)


def test_2_20_resident_single_gpu(ctx):
    """2^20 ciphertexts resident on ONE GPU (135.8 GB of the 180 GB): resident lincomb == fused eval_poly."""
    import torch
    free, _ = torch.cuda.mem_get_info(0)
    if free < 150e9:
        pytest.skip(f"needs ~136 GB of free HBM, {free / 1e9:.0f} GB available")
    D = 1 << 20
    c8 = xof_records("full20-c8", D)
    h = xof_scalars("full20-h", D)
    reg = ctx.region(SEED, OFF, c8)
    try:
        res = reg.lincomb(h)
    finally:
        reg.close()
    assert np.array_equal(res, ctx.eval_poly(SEED, OFF, c8, h))
