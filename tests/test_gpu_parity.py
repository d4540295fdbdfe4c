"""Parity of the CUDA path (through the C-ABI, include/mfb200.h) with the CPU oracle and the golden vectors
emitted by the compiled reference.  Bit-exact: everything here is unsigned integer / byte work.

Mirrors the reference's own tests where they exist: test_entropy.c (determinism, chunking independence, seek),
test_lwe.c::{test_import_export, test_eval, test_correctness, test_smudging}.
"""
import json
from pathlib import Path

import numpy as np
import pytest

from conftest import SEED, sha, xof, xof_records, xof_scalars

pytestmark = pytest.mark.gpu

GOLD = json.loads((Path(__file__).parent / "golden" / "vectors.json").read_text())
N, NC, L64, CT_BYTES, CTR_CT, P = 1470, 1471, 11, 92, 92 * 1470, 0xFFFFFFFB


@pytest.fixture(scope="module")
def ctx():
    import c_lwe_snarks_b200 as m
    c = m.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def torch():
    import torch as t
    assert t.cuda.is_available()
    return t


def wide(flat11: np.ndarray) -> np.ndarray:
    """(…, 11) u64 -> (…, 12) u64, the oracle's 736-bit width (limb 11 is zero after modq)."""
    out = np.zeros(flat11.shape[:-1] + (12,), np.uint64)
    out[..., :11] = flat11
    return out


def dev(torch, a: np.ndarray):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).cuda()


def resident_to_flat(res: np.ndarray) -> np.ndarray:
    """(k, 16192) u64, resident tile-planar layout -> (k, 1471, 11)"""
    from c_lwe_snarks_b200.api import resident_to_flat as f
    return f(res)


# ------------------------------------------------------------------------------------------- K2 stream
@pytest.mark.parametrize("off,n", [(0, 64), (8, 40), (135240, 100), (3, 1), (15, 2), (16, 16), (1, 4097),
                                   (2 * 135240 * 65536 + 135240 - 3, 50), (2**36 + 5, 33), (2**40 + 7, 1000)])
def test_stream_vs_oracle(ctx, oracle, off, n):
    assert np.array_equal(ctx.stream(SEED, off, n), oracle.stream(SEED, off, n))


def test_stream_golden(ctx):
    for s in GOLD["stream"]:
        assert ctx.stream(SEED, s["offset"], s["n"]).tobytes().hex() == s["hex"]
    assert sha(ctx.stream(SEED, 3 * CTR_CT, CTR_CT)) == GOLD["stream_ct3_sha"]


def test_stream_chunking_and_seek(ctx, oracle):
    # test_entropy.c:111-156: one bulk read == many small reads; seek == consume
    bulk = ctx.stream(SEED, 0, 92 * 1000)
    parts = np.concatenate([ctx.stream(SEED, 92 * i, 92) for i in range(0, 1000, 97)])
    assert np.array_equal(parts, np.concatenate([bulk[92 * i:92 * i + 92] for i in range(0, 1000, 97)]))
    assert np.array_equal(ctx.stream(SEED, 512, 64), bulk[512:576])
    other = bytes(range(7, 47))
    assert not np.array_equal(ctx.stream(other, 0, 64), bulk[:64])
    assert np.array_equal(ctx.stream(other, 5, 64), oracle.stream(other, 5, 64))


def test_stream_empty(ctx):
    assert ctx.stream(SEED, 12345, 0).size == 0


# ------------------------------------------------------------------------------------------- expand (ct_import)
@pytest.mark.parametrize("off", [0, 4 * CTR_CT, 5 * CTR_CT, 7, 2 * CTR_CT * 256 + CTR_CT + 2, 2**37 + 13])
def test_expand_vs_oracle(ctx, oracle, torch, off):
    k = 5
    c8 = xof_records(f"expand-{off}", k)
    c8[1, 88:] = 0xFF  # a record with junk in the dead top bytes must give the same live limbs
    d_c8 = dev(torch, c8)
    d_cts = torch.zeros(k * L64 * 1472 * 8, dtype=torch.uint8, device="cuda")
    ctx.expand_dev(SEED, off, d_c8.data_ptr(), k, d_cts.data_ptr())
    torch.cuda.synchronize()
    res = d_cts.cpu().numpy().view(np.uint64).reshape(k, L64 * 1472)
    assert not res.reshape(k, 23, L64, 64)[:, 22, :, 63].any()  # padding coordinate 1471
    got = resident_to_flat(res)
    for i in range(k):
        want = oracle.ct_import(SEED, off + i * CTR_CT, c8[i])
        assert np.array_equal(got[i], want[:, :11]), f"ciphertext {i}"


def test_expand_golden(ctx, torch):
    for name in ("ct_import_even", "ct_import_odd"):
        g = GOLD[name]
        b = np.frombuffer(bytes.fromhex(g["b"]), np.uint8)
        d_cts = torch.zeros(L64 * 1472 * 8, dtype=torch.uint8, device="cuda")
        ctx.expand_dev(SEED, g["offset"], dev(torch, b).data_ptr(), 1, d_cts.data_ptr())
        torch.cuda.synchronize()
        ct = wide(resident_to_flat(d_cts.cpu().numpy().view(np.uint64).reshape(1, L64 * 1472))[0])
        # the golden ciphertext keeps the dead limb 11 of every a_j; compare the live part and the literals
        assert ct[0, :11].tobytes().hex() == g["a0"][: 11 * 16]
        assert ct[1469, :11].tobytes().hex() == g["a1469"][: 11 * 16]
        assert ct[1470, :11].tobytes().hex() == g["b_limbs"][: 11 * 16]


# ------------------------------------------------------------------------------------------- K1 + fused eval_poly
def golden_eval_inputs():
    d = 12
    c8, h = xof_records("eval-c8", d), xof_scalars("eval-h", d)
    h[3] = 0
    h[4] = P - 1
    return d, c8, h, 3 * CTR_CT


def test_eval_poly_golden(ctx):
    d, c8, h, off = golden_eval_inputs()
    acc = ctx.eval_poly(SEED, off, c8, h)
    assert sha(wide(acc)) == GOLD["eval_poly"]["sha"]
    acc2 = ctx.eval_poly(SEED, off, c8, h, rop=acc)  # accumulates INTO rop (lwe.c:176-186)
    assert sha(wide(acc2)) == GOLD["eval_poly_accumulate"]["sha"]


def test_lincomb_resident_golden(ctx, torch):
    d, c8, h, off = golden_eval_inputs()
    d_cts = torch.zeros(d * L64 * 1472 * 8, dtype=torch.uint8, device="cuda")
    ctx.expand_dev(SEED, off, dev(torch, c8).data_ptr(), d, d_cts.data_ptr())
    d_h = dev(torch, h.astype(np.uint32))
    d_rop = torch.zeros(NC * L64 * 8, dtype=torch.uint8, device="cuda")
    ctx.lincomb_dev(d_cts.data_ptr(), d_h.data_ptr(), d, None, d_rop.data_ptr())
    torch.cuda.synchronize()
    acc = d_rop.cpu().numpy().view(np.uint64).reshape(NC, L64)
    assert sha(wide(acc)) == GOLD["eval_poly"]["sha"]
    ctx.lincomb_dev(d_cts.data_ptr(), d_h.data_ptr(), d, d_rop.data_ptr(), d_rop.data_ptr())  # in place
    torch.cuda.synchronize()
    assert sha(wide(d_rop.cpu().numpy().view(np.uint64).reshape(NC, L64))) == GOLD["eval_poly_accumulate"]["sha"]


@pytest.mark.parametrize("d,off", [(1, 0), (2, CTR_CT), (100, 0), (149, 5 * CTR_CT), (300, 11), (700, 2**36 + CTR_CT)])
def test_eval_poly_vs_oracle(ctx, oracle, d, off):
    # test_lwe.c::test_eval uses d = 100 all-ones coefficients; here random scalars in [0, p) with edge values
    c8, h = xof_records(f"ev-c8-{d}", d), xof_scalars(f"ev-h-{d}", d)
    h[0] = P - 1
    if d > 2:
        h[1], h[2] = 0, 1
    rop0 = xof(f"ev-rop-{d}", NC * 88).view("<u8").reshape(NC, L64)
    want = oracle.eval_poly(SEED, off, c8, h, rop=wide(rop0))
    got = ctx.eval_poly(SEED, off, c8, h, rop=rop0)
    assert np.array_equal(wide(got), want)


def test_eval_poly_all_ones(ctx, oracle):
    d = 100  # exactly test_lwe.c:105-181's shape
    c8 = xof_records("ones", d)
    h = np.ones(d, np.uint64)
    assert np.array_equal(wide(ctx.eval_poly(SEED, 0, c8, h)), oracle.eval_poly(SEED, 0, c8, h))


def test_eval_poly_empty_and_idx(ctx, oracle):
    rop0 = xof("rop-empty", NC * 88).view("<u8").reshape(NC, L64)
    assert np.array_equal(ctx.eval_poly(SEED, 0, np.zeros((0, 92), np.uint8), np.zeros(0, np.uint64), rop=rop0), rop0)
    # index list = the prover's b_w loop (snark.c:143-155): only set witness bits contribute
    M = 40
    c8 = xof_records("idx-c8", M)
    bits = xof("idx-bits", M) & 1
    bits[0] = 1
    full = np.where(bits == 1, 1, 0).astype(np.uint64)
    full[0] = 123456789  # delta on CT_t
    want = oracle.eval_poly(SEED, 2 * CTR_CT * 64, c8, full)
    idx = np.nonzero(full)[0].astype(np.uint32)
    got = ctx.eval_poly(SEED, 2 * CTR_CT * 64, c8, full[idx], idx=idx)
    assert np.array_equal(wide(got), want)


def test_lincomb_resident_vs_fused_and_oracle(ctx, oracle, torch):
    d, off = 2500, 7 * CTR_CT  # > nchunks: every CTA gets a ragged slice
    c8, h = xof_records("res-c8", d), xof_scalars("res-h", d)
    d_cts = torch.zeros(d * L64 * 1472 * 8, dtype=torch.uint8, device="cuda")
    ctx.expand_dev(SEED, off, dev(torch, c8).data_ptr(), d, d_cts.data_ptr())
    d_h = dev(torch, h.astype(np.uint32))
    d_rop = torch.zeros(NC * L64 * 8, dtype=torch.uint8, device="cuda")
    ctx.lincomb_dev(d_cts.data_ptr(), d_h.data_ptr(), d, None, d_rop.data_ptr())
    torch.cuda.synchronize()
    res = d_rop.cpu().numpy().view(np.uint64).reshape(NC, L64)
    fused = ctx.eval_poly(SEED, off, c8, h)
    assert np.array_equal(res, fused)
    want = oracle.eval_poly(SEED, off, c8[:400], h[:400])
    ctx.lincomb_dev(d_cts.data_ptr(), d_h.data_ptr(), 400, None, d_rop.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(wide(d_rop.cpu().numpy().view(np.uint64).reshape(NC, L64)), want)


@pytest.mark.parametrize("d,off", [(1, 0), (2, 8), (3, 16), (7, 3 * CTR_CT + 8), (60, 11), (147, 5), (149, 2 * CTR_CT), (333, 2**36 + 5 * CTR_CT),
                                   (1000, 77)])
def test_eval_poly2_vs_oracle(ctx, oracle, d, off):
    """two scalar vectors in one pass == two eval_poly calls == the oracle (prover pairs v_w/h and hat_v/hat_h)"""
    c8 = xof_records(f"ev2-c8-{d}", d)
    h0, h1 = xof_scalars(f"ev2-h0-{d}", d), xof_scalars(f"ev2-h1-{d}", d)
    h1[0] = 0
    rop0 = xof(f"ev2-rop-{d}", NC * 88).view("<u8").reshape(NC, L64)
    got0, got1 = ctx.eval_poly2(SEED, off, c8, h0, h1, rop0=rop0)
    assert np.array_equal(wide(got0), oracle.eval_poly(SEED, off, c8, h0, rop=wide(rop0)))
    assert np.array_equal(wide(got1), oracle.eval_poly(SEED, off, c8, h1))
    assert np.array_equal(got1, ctx.eval_poly(SEED, off, c8, h1))


@pytest.mark.parametrize("d,two", [(1, True), (75, True), (75, False), (500, True)])
def test_eval_poly2_in_two_halves_with_host_records(ctx, oracle, torch, d, two):
    """mfb_eval_poly2_begin_dev / _end_dev (what a device set queues per member): the AES + MAC kernel first, the wire records
    copied from PAGEABLE host memory on the second stream behind it — equal to the oracle, one or two scalar vectors"""
    off = 4 * CTR_CT + 12
    c8 = np.ascontiguousarray(xof_records(f"halves-c8-{d}", d))
    h0, h1 = xof_scalars(f"halves-h0-{d}", d), xof_scalars(f"halves-h1-{d}", d)
    st = torch.cuda.current_stream().cuda_stream
    d_h0, d_h1 = dev(torch, h0.astype(np.uint32)), dev(torch, h1.astype(np.uint32))
    d_c8 = torch.zeros(d * CT_BYTES + 16, dtype=torch.uint8, device="cuda")
    d_r0 = torch.zeros(1472 * L64, dtype=torch.int64, device="cuda")
    d_r1 = torch.zeros(1472 * L64, dtype=torch.int64, device="cuda")
    p1 = d_h1.data_ptr() if two else None
    for _ in range(2):  # twice: the second pair reuses the context's second stream and events
        ctx.eval_poly2_begin_dev(SEED, off, d_h0.data_ptr(), p1, d, True, st)
        ctx.eval_poly2_end_dev(d_c8.data_ptr(), c8.ctypes.data, d_h0.data_ptr(), p1, d, None, d_r0.data_ptr(), None,
                               d_r1.data_ptr() if two else None, st)
        torch.cuda.synchronize()
        got0 = d_r0.cpu().numpy().view(np.uint64)[: NC * L64].reshape(NC, L64)
        assert np.array_equal(wide(got0), oracle.eval_poly(SEED, off, c8, h0))
        if two:
            got1 = d_r1.cpu().numpy().view(np.uint64)[: NC * L64].reshape(NC, L64)
            assert np.array_equal(wide(got1), oracle.eval_poly(SEED, off, c8, h1))


def test_region(ctx, oracle):
    d, off = 64, 9 * CTR_CT
    c8, h = xof_records("reg-c8", d), xof_scalars("reg-h", d)
    reg = ctx.region(SEED, off, c8)
    try:
        assert np.array_equal(wide(reg.lincomb(h)), oracle.eval_poly(SEED, off, c8, h))
        part = reg.lincomb(h[10:30], first=10)
        assert np.array_equal(wide(part), oracle.eval_poly(SEED, off + 10 * CTR_CT, c8[10:30], h[10:30]))
    finally:
        reg.close()


def test_region_two_vectors(ctx, oracle):
    d, off = 300, 4 * CTR_CT
    c8 = xof_records("reg2-c8", d)
    h0, h1 = xof_scalars("reg2-h0", d), xof_scalars("reg2-h1", d)
    reg = ctx.region(SEED, off, c8)
    try:
        rop1 = xof("reg2-rop", NC * 88).view("<u8").reshape(NC, L64)
        r0, r1 = reg.lincomb2(h0, h1, rop1=rop1)
        assert np.array_equal(wide(r0), oracle.eval_poly(SEED, off, c8, h0))
        assert np.array_equal(wide(r1), oracle.eval_poly(SEED, off, c8, h1, rop=wide(rop1)))
        assert np.array_equal(reg.lincomb(h0), r0)  # the queues were re-armed
    finally:
        reg.close()


def test_ct_ops_golden(ctx, oracle):
    # ct_mul_ui / ct_add / ct_addmul_ui (lwe.c:131-157) as lincombs over host ciphertexts
    d, c8, h, _ = golden_eval_inputs()
    x = oracle.ct_import(SEED, 0, c8[0])[:, :11]
    y = oracle.ct_import(SEED, CTR_CT, c8[1])[:, :11]
    mul = ctx.lincomb(x[None], [int(h[0])])
    assert sha(wide(mul)) == GOLD["ct_ops"]["mul"]
    add = ctx.lincomb(np.stack([x, y]), [1, 1])
    assert sha(wide(add)) == GOLD["ct_ops"]["add"]
    addmul = ctx.lincomb(y[None], [int(h[1])], rop=ctx.lincomb(x[None], [7]))
    assert sha(wide(addmul)) == GOLD["ct_ops"]["addmul"]


# ------------------------------------------------------------------------------------------- K3 encrypt / K4 decrypt
def golden_lwe_inputs(oracle):
    cnt = 4
    ent = xof("lwe-entropy", N * CT_BYTES + cnt * 70)
    m = xof_scalars("lwe-m", cnt)
    m[0] = 0
    m[1] = P - 1
    sk = oracle.key_gen(ent[: N * CT_BYTES])
    return cnt, sk, m, ent[N * CT_BYTES:], GOLD["lwe"]["offset"]


def test_encrypt_golden(ctx, oracle):
    cnt, sk, m, ent, off = golden_lwe_inputs(oracle)
    assert sha(sk) == GOLD["lwe"]["sk_sha"]
    recs = ctx.encrypt(SEED, off, sk[:, :11], m, ent)
    assert recs.tobytes().hex() == GOLD["lwe"]["records"]


@pytest.mark.parametrize("cnt,off", [(1, 0), (3, CTR_CT), (200, 13), (301, 2**35 + 5 * CTR_CT)])
def test_encrypt_vs_oracle(ctx, oracle, cnt, off):
    sk = oracle.key_gen(xof(f"sk-{cnt}", N * CT_BYTES))
    m = xof_scalars(f"m-{cnt}", cnt)
    ent = xof(f"ent-{cnt}", cnt * 70)
    ent[:69] = 0xFF  # maximal noise
    want = oracle.encrypt(SEED, off, sk, m, ent)
    got = ctx.encrypt(SEED, off, sk[:, :11], m, ent)
    assert np.array_equal(got, want)


def test_encrypt_decrypt_roundtrip(ctx, oracle):
    # test_lwe.c::test_correctness: Dec(Enc(m)) = m
    cnt, off = 10, 4 * CTR_CT
    sk = oracle.key_gen(xof("sk-rt", N * CT_BYTES))
    m = xof_scalars("m-rt", cnt)
    recs = ctx.encrypt(SEED, off, sk[:, :11], m, xof("ent-rt", cnt * 70))
    cts = np.stack([oracle.ct_import(SEED, off + i * CTR_CT, recs[i])[:, :11] for i in range(cnt)])
    dec, dot = ctx.decrypt(sk[:, :11], cts, want_dot=True)
    assert [int(v) for v in dec] == [int(v) for v in m]
    for i in range(cnt):
        assert np.array_equal(wide(dot[i]), oracle.dotp(wide(cts[i][:N]), sk))
        assert int(dec[i]) == oracle.decrypt(sk, wide(cts[i]))


def test_decrypt_golden_dot(ctx, oracle):
    cnt, sk, m, ent, off = golden_lwe_inputs(oracle)
    recs = ctx.encrypt(SEED, off, sk[:, :11], m, ent)
    ct0 = oracle.ct_import(SEED, off, recs[0])[:, :11]
    dec, dot = ctx.decrypt(sk[:, :11], ct0[None], want_dot=True)
    assert wide(dot[0]).tobytes().hex() == GOLD["lwe"]["dotp0"]
    assert int(dec[0]) == GOLD["lwe"]["m"][0]


def test_decrypt_after_smudge_signs(ctx, oracle):
    # test_lwe.c::test_smudging + the negative-b case ct_smudge can produce (lwe.c:65-76)
    cnt, sk, m, ent, off = golden_lwe_inputs(oracle)
    recs = ctx.encrypt(SEED, off, sk[:, :11], m, ent)
    ct0 = oracle.ct_import(SEED, off, recs[0])
    cts, negs, want = [], [], []
    for k in range(6):
        out, neg = oracle.ct_smudge(ct0, xof(f"smudge{k}", 81))
        assert out[N].tobytes().hex() == GOLD["smudge"][k]["b"] and neg == GOLD["smudge"][k]["negative"]
        cts.append(out[:, :11])
        negs.append(neg)
        want.append(oracle.decrypt(sk, out, neg))
    got = ctx.decrypt(sk[:, :11], np.stack(cts), b_neg=negs)
    assert [int(v) for v in got] == want
    for k in range(6):
        if GOLD["smudge"][k]["dec"] is not None:
            assert int(got[k]) == GOLD["smudge"][k]["dec"]
    # a uniformly random b makes a negative result a 2^-32 event; force it with a tiny b and a negative smudge
    small = ct0.copy()
    small[N] = 0
    small[N, 0] = 5
    e81 = xof("smudge-neg", 81)
    e81[80] |= 1
    out, neg = oracle.ct_smudge(small, e81)
    assert neg
    e81[80] &= 0xFE
    out2, neg2 = oracle.ct_smudge(small, e81)
    assert not neg2
    got = ctx.decrypt(sk[:, :11], np.stack([out[:, :11], out2[:, :11], small[:, :11]]), b_neg=[True, False, False])
    assert [int(v) for v in got] == [oracle.decrypt(sk, out, True), oracle.decrypt(sk, out2, False),
                                     oracle.decrypt(sk, small, False)]


@pytest.mark.parametrize("cnt", [1, 5, 148 * 16, 148 * 16 + 1, 148 * 126 + 77, 40000])
def test_encrypt_cb_equals_encrypt(ctx, oracle, cnt):
    """mfb_encrypt_cb (entropy drawn piece by piece through a callback while the device works; setup() uses it) gives
    the records of mfb_encrypt on the same entropy, draws every byte exactly once and in order."""
    sk = oracle.key_gen(xof("sk-cb", N * CT_BYTES))
    m = xof_scalars(f"m-cb-{cnt}", cnt)
    ent = xof(f"ent-cb-{cnt}", cnt * 70)
    pos, calls = [0], []

    def draw(n):
        out = ent[pos[0]: pos[0] + n].tobytes()
        calls.append(n)
        pos[0] += n
        return out

    got = ctx.encrypt_cb(SEED, 7 * CTR_CT + 3, sk[:, :11], m, draw)
    assert pos[0] == cnt * 70 and all(n % 70 == 0 and n > 0 for n in calls)
    assert np.array_equal(got, ctx.encrypt(SEED, 7 * CTR_CT + 3, sk[:, :11], m, ent))
    k = cnt - 1
    want = oracle.encrypt(SEED, 7 * CTR_CT + 3 + k * CTR_CT, sk, m[k:k + 1], ent[70 * k: 70 * k + 70])
    assert np.array_equal(got[k], want[0])


@pytest.mark.parametrize("cuts", [(1,), (3, 0, 2), (148 * 16 - 1, 2, 148 * 110 + 5, 1, 63), (20000, 20000, 1, 63)])
def test_encrypt_cb_into_segments(ctx, oracle, cuts):
    """mfb_encrypt_cb_segs (setup(): records straight into crs->s / as / t / v, each piece travelling back while the next
    one is encrypted): the segments, concatenated, are the records of mfb_encrypt — cuts inside pieces, empty segments"""
    cnt = sum(cuts)
    sk = oracle.key_gen(xof("sk-cbs", N * CT_BYTES))
    m = xof_scalars(f"m-cbs-{cnt}", cnt)
    ent = xof(f"ent-cbs-{cnt}", cnt * 70)
    pos = [0]

    def draw(n):
        out = ent[pos[0]: pos[0] + n].tobytes()
        pos[0] += n
        return out

    outs = ctx.encrypt_cb_segs(SEED, 2 * CTR_CT + 9, sk[:, :11], m, draw, cuts)
    assert pos[0] == cnt * 70 and [o.shape[0] for o in outs] == list(cuts)
    assert np.array_equal(np.concatenate(outs), ctx.encrypt(SEED, 2 * CTR_CT + 9, sk[:, :11], m, ent))


def test_context_reserve_leaves_no_trace(oracle):
    """mfb_ctx_reserve (the drop-in's background warm-up: scratch at the instance's sizes, one-element dry runs that load
    every kernel) on a fresh context, then the real calls: results equal the oracle's, accumulators start from zero"""
    import c_lwe_snarks_b200 as m
    c = m.Context(0)
    try:
        c._ck(c.lib.mfb_ctx_reserve(c.h, 300, 20))
        c8, h = xof_records("resv-c8", 7), xof_scalars("resv-h", 7)
        assert np.array_equal(wide(c.eval_poly(SEED, 5, c8, h)), oracle.eval_poly(SEED, 5, c8, h))
        sk = oracle.key_gen(xof("sk-resv", N * CT_BYTES))
        msg, ent = xof_scalars("m-resv", 3), xof("e-resv", 3 * 70)
        assert np.array_equal(c.encrypt(SEED, 11, sk[:, :11], msg, ent), oracle.encrypt(SEED, 11, sk, msg, ent))
        assert c.lib.mfb_ctx_reserve(c.h, 0, 1) != 0
    finally:
        c.close()


def test_allocation_failure_is_reported_and_does_not_poison_later_calls(ctx, oracle):
    """A region that cannot be allocated (2^21 ciphertexts = 272 GB) fails with MFB_ENOMEM and leaves no stale CUDA
    error behind: the next calls work (mf_crs_make_resident relies on this to fall back to the fused path)."""
    import c_lwe_snarks_b200 as m
    big = np.zeros((1 << 21, 92), np.uint8)
    with pytest.raises(m.api.MfbError, match="cudaMalloc"):
        ctx.region(SEED, 0, big)
    c8, h = xof_records("after-oom-c8", 5), xof_scalars("after-oom-h", 5)
    assert np.array_equal(wide(ctx.eval_poly(SEED, 0, c8, h)), oracle.eval_poly(SEED, 0, c8, h))
    blob = np.arange(64 * 9, dtype=np.uint64)
    assert ctx.ssp_eval(blob, 64, 7).shape == (9,)
