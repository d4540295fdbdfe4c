"""TEST INFRASTRUCTURE — small-case Python restatement of the protocol layer (snark.c, ssp.c).

Built on :class:`oracle.loader.Oracle` primitives with plain Python integers for F_p[x]
(schoolbook; use only for D <= a few hundred).  Follows, step by step:

* ``random_ssp``  ssp.c:37-77      (entropy: M/8 bytes, then M draws of 8*D bytes)
* ``setup``       snark.c:35-48,57-115   (entropy: 40 seed; 8,8,8; 1470*92; (69,1) per encryption)
* ``prover``      snark.c:117-190  (entropy: 8 for delta; 5 x (80,1) smudging — v_w twice, b_w never)
* ``verifier``    snark.c:192-250

The entropy argument of each function is the byte string the reference would have pulled from
getrandom(2), in call order, so results can be compared bit for bit with oracle/_ref under
ref_shim.c's interposer and with tests/golden/vectors.json.
"""
from __future__ import annotations

import numpy as np

from .loader import CT_BYTES, CTR_CT, LIMBS, N, NC, NOISE_BYTES, P, SMUDGE_BYTES, Oracle


class Entropy:
    """Sequential reader over the injected entropy bytes (mirrors successive getrandom calls)."""

    def __init__(self, data):
        self.data = np.ascontiguousarray(data, dtype=np.uint8)
        self.pos = 0

    def take(self, n: int) -> np.ndarray:
        assert self.pos + n <= self.data.size, "entropy exhausted"
        out = self.data[self.pos:self.pos + n]
        self.pos += n
        return out

    def rand_modp(self) -> int:  # lwe.h:97-103
        return int(self.take(8).view("<u8")[0]) % P


# ---------------------------------------------------------------- F_p[x], coefficient lists (low first)

def poly_trim(a):
    while a and a[-1] == 0:
        a.pop()
    return a


def poly_add(a, b):
    n = max(len(a), len(b))
    return poly_trim([((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % P for i in range(n)])


def poly_sub(a, b):
    n = max(len(a), len(b))
    return poly_trim([((a[i] if i < len(a) else 0) - (b[i] if i < len(b) else 0)) % P for i in range(n)])


def poly_mul(a, b):
    if not a or not b:
        return []
    r = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                r[i + j] += x * y
    return poly_trim([v % P for v in r])


def poly_div(a, b):
    a = list(a)
    if len(a) < len(b):
        return []
    inv = pow(b[-1], P - 2, P)
    q = [0] * (len(a) - len(b) + 1)
    for i in range(len(q) - 1, -1, -1):
        c = a[i + len(b) - 1] * inv % P
        q[i] = c
        if c:
            for j, y in enumerate(b):
                a[i + j] = (a[i + j] - c * y) % P
    return poly_trim(q)


def poly_eval(a, x):
    acc = 0
    for c in reversed(a):
        acc = (acc * x + c) % P
    return acc


# ---------------------------------------------------------------- SSP wire format (ssp.h:6-9, ssp.c:18-34)

def ssp_size(D: int, M: int) -> int:
    return D * 8 * (M + 3)


def ssp_poly(ssp: np.ndarray, D: int, index: int):
    """index 0 = t, index i+1 = v_i; 8-byte little-endian coefficients reduced mod p on import."""
    raw = ssp[D * 8 * index: D * 8 * (index + 1)].view("<u8")
    return poly_trim([int(v) % P for v in raw])


def random_ssp(D: int, M: int, ent: Entropy):
    ssp = np.zeros(ssp_size(D, M), np.uint8)
    wbytes = ent.take(M // 8)  # mpz2_urandomb2(input, GAMMA_M): M/8 bytes, masked to M bits
    witness = int.from_bytes(wbytes.tobytes(), "little") & ((1 << M) - 1)

    def put(index, poly):
        buf = np.zeros(D, "<u8")
        buf[:len(poly)] = poly
        ssp[D * 8 * index: D * 8 * (index + 1)] = buf.view(np.uint8)

    t = []
    for i in range(M):
        raw = ent.take(8 * D).view("<u8")
        v = poly_trim([int(x) % P for x in raw])
        put(i + 1, v)
        if i == 0 or (witness >> (i - 1)) & 1:
            t = poly_add(t, v)
    t = poly_sub(t, [1])
    put(0, t)
    return ssp, witness


# ---------------------------------------------------------------- setup / prover / verifier

def setup(orc: Oracle, ssp: np.ndarray, D: int, M: int, ent: Entropy) -> dict:
    seed = ent.take(40).copy()
    alpha, beta, s = ent.rand_modp(), ent.rand_modp(), ent.rand_modp()
    sk = orc.key_gen(ent.take(N * CT_BYTES))
    msgs = []
    x = 1
    for _ in range(D):  # Enc(s^i)           snark.c:75-82
        msgs.append(x)
        x = x * s % P
    x = alpha
    for _ in range(D):  # Enc(alpha s^i)     snark.c:84-91
        msgs.append(x)
        x = x * s % P
    msgs.append(poly_eval(ssp_poly(ssp, D, 0), s) * beta % P)  # beta t(s)   snark.c:97-101
    for i in range(1, M):  # beta v_i(s)      snark.c:104-110
        msgs.append(poly_eval(ssp_poly(ssp, D, i + 1), s) * beta % P)
    count = len(msgs)
    # one rng, read sequentially from position 0: the stream order IS the region map snark.h:8-12
    recs = orc.encrypt(seed, 0, sk, np.array(msgs, np.uint64), ent.take(count * (NOISE_BYTES + 1)))
    v = np.zeros((M, CT_BYTES), np.uint8)
    v[: M - 1] = recs[2 * D + 1:]
    return dict(seed=seed, s=recs[:D], as_=recs[D:2 * D], t=recs[2 * D], v=v, alpha=alpha, beta=beta,
                s_point=s, sk=sk)


def prover(orc: Oracle, ssp: np.ndarray, crs: dict, witness: int, D: int, M: int, ent: Entropy):
    """Returns (proof (5, 1471, 12) uint64 in struct order h, hat_h, hat_v, v_w, b_w; negative[5])."""
    seed = crs["seed"]
    CTR_S, CTR_AS, CTR_BT, CTR_BV = 0, CTR_CT * D, 2 * CTR_CT * D, 2 * CTR_CT * D + CTR_CT
    t = ssp_poly(ssp, D, 0)
    delta = ent.rand_modp()
    w = poly_trim([c * delta % P for c in t])
    b_w = orc.ct_mul_ui(orc.ct_import(seed, CTR_BT, crs["t"]), delta)
    for i in range(1, M):
        if (witness >> (i - 1)) & 1:
            w = poly_add(w, ssp_poly(ssp, D, i + 1))
            b_w = orc.ct_add(b_w, orc.ct_import(seed, CTR_BV + (i - 1) * CTR_CT, crs["v"][i - 1]))

    def coeffs(poly):
        c = np.zeros(D, np.uint64)
        c[:len(poly)] = poly
        return c

    v_w = orc.eval_poly(seed, CTR_S, crs["s"], coeffs(w))
    w = poly_add(w, ssp_poly(ssp, D, 1))
    hat_v = orc.eval_poly(seed, CTR_AS, crs["as_"], coeffs(w))
    h = poly_div(poly_sub(poly_mul(w, w), [1]), t)
    h_ct = orc.eval_poly(seed, CTR_S, crs["s"], coeffs(h))
    hat_h = orc.eval_poly(seed, CTR_AS, crs["as_"], coeffs(h))
    neg = [False] * 5
    h_ct, neg[0] = orc.ct_smudge(h_ct, ent.take(SMUDGE_BYTES + 1))
    hat_h, neg[1] = orc.ct_smudge(hat_h, ent.take(SMUDGE_BYTES + 1))
    hat_v, neg[2] = orc.ct_smudge(hat_v, ent.take(SMUDGE_BYTES + 1))
    v_w, n1 = orc.ct_smudge(v_w, ent.take(SMUDGE_BYTES + 1))
    assert not n1, "negative b after first v_w smudge: second smudge on a negative value not restated"
    v_w, neg[3] = orc.ct_smudge(v_w, ent.take(SMUDGE_BYTES + 1))
    return np.stack([h_ct, hat_h, hat_v, v_w, b_w]), neg


def verifier(orc: Oracle, ssp: np.ndarray, crs: dict, proof: np.ndarray, D: int, negative=None) -> bool:
    negative = negative or [False] * 5
    alpha, beta, s, sk = crs["alpha"], crs["beta"], crs["s_point"], crs["sk"]
    t_s = poly_eval(ssp_poly(ssp, D, 0), s)
    h_s, hath_s, hatv_s, w_s, b_s = (orc.decrypt(sk, proof[k], negative[k]) for k in range(5))
    v_s = (poly_eval(ssp_poly(ssp, D, 1), s) + w_s) % P
    if h_s * alpha % P != hath_s:  # eq-pke
        return False
    if v_s * alpha % P != hatv_s:
        return False
    if (v_s * v_s - 1 - h_s * t_s) % P != 0:  # eq-div
        return False
    if w_s * beta % P != b_s:  # eq-lin
        return False
    # test-error (snark.c:238-241): SIZ(test) of a non-positive value is never >= 80 -> vacuous
    return True
