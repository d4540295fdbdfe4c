"""Generate tests/golden/vectors.json from the COMPILED REFERENCE (oracle/_ref/libmfref_*.so).

The reference ships no golden vectors or known-answer tests for this path, so
the fixtures are outputs of the reference's own sources, compiled by oracle/Makefile and run in the
build container under the deterministic entropy interposer of oracle/ref_shim.c.  Inputs are
SHAKE-256 expansions of the labels below (tests/conftest.py::xof), so nothing depends on a numpy RNG.
Large outputs are stored as SHA-256 digests of their little-endian bytes plus a few literal limbs.

Run from the repo root (needs /root/reference):   python tests/golden/make_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from conftest import SEED, sha, xof, xof_records, xof_scalars  # noqa: E402
from oracle.loader import CT_BYTES, CTR_CT, N, NOISE_BYTES, P, Reference  # noqa: E402


def hexs(a) -> str:
    return np.ascontiguousarray(a).tobytes().hex()


def main() -> None:
    ref = Reference(256, 64)
    g: dict = {"_comment": "emitted by tests/golden/make_golden.py from the compiled reference; do not edit"}

    # ---- AES / stream (aes.c:104-144, entropy.c:46-61)
    g["aes_fips197_c3"] = dict(key=bytes(range(32)).hex(), pt="00112233445566778899aabbccddeeff",
                               ct="8ea2b7ca516745bfeafc49904b496089")
    g["stream"] = []
    for off, n in [(0, 64), (8, 40), (135240, 100), (2 * 135240 * 65536 + 135240 - 3, 50), (2**36 + 5, 33)]:
        g["stream"].append(dict(offset=off, n=n, hex=hexs(ref.stream(SEED, off, n))))
    big = ref.stream(SEED, 3 * CTR_CT, CTR_CT)
    g["stream_ct3_sha"] = sha(big)

    # ---- mpz2_urandomb (entropy.c:11-26) and modq (lwe.h:108-118)
    g["urandomb"] = []
    for nbits in [64, 1, 5, 32, 40, 520, 512, 700, 736]:
        limbs, siz = ref.urandomb(SEED, 11, nbits)
        g["urandomb"].append(dict(offset=11, nbits=nbits, limbs=hexs(limbs), siz=siz))
    g["modq"] = []
    for k, nl in enumerate([5, 11, 12, 13, 14]):
        x = xof(f"modq{k}", 8 * nl).view("<u8")
        out, siz = ref.modq(x)
        g["modq"].append(dict(x=hexs(x), out=hexs(out), siz=siz))

    # ---- ct_import (lwe.c:122-126)
    b = xof_records("import-b", 1)[0]
    for name, off in [("ct_import_even", 4 * CTR_CT), ("ct_import_odd", 5 * CTR_CT)]:
        ct = ref.ct_import(SEED, off, b)
        g[name] = dict(offset=off, b=hexs(b), sha=sha(ct), a0=hexs(ct[0]), a1469=hexs(ct[1469]), b_limbs=hexs(ct[1470]))

    # ---- eval_poly (lwe.c:176-186) over 12 ciphertexts starting at an odd ciphertext index
    d = 12
    c8, h = xof_records("eval-c8", d), xof_scalars("eval-h", d)
    h[3] = 0
    h[4] = P - 1
    off = 3 * CTR_CT
    acc = ref.eval_poly(SEED, off, c8, h)
    g["eval_poly"] = dict(offset=off, d=d, sha=sha(acc), c0=hexs(acc[0]), c777=hexs(acc[777]), b=hexs(acc[1470]))
    acc2 = ref.eval_poly(SEED, off, c8, h, rop=acc)  # accumulates INTO rop
    g["eval_poly_accumulate"] = dict(sha=sha(acc2))

    # ---- ct_mul_ui / ct_add / ct_addmul_ui (lwe.c:131-157)
    x = ref.ct_import(SEED, 0, c8[0])
    y = ref.ct_import(SEED, CTR_CT, c8[1])
    g["ct_ops"] = dict(mul=sha(ref.ct_mul_ui(x, int(h[0]))), add=sha(ref.ct_add(x, y)),
                       addmul=sha(ref.ct_addmul_ui(ref.ct_mul_ui(x, 7), y, int(h[1]))))

    # ---- key_gen / regev_encrypt / regev_decrypt / ct_smudge (lwe.c:30-34,60-111)
    cnt = 4
    ent = xof("lwe-entropy", N * CT_BYTES + cnt * (NOISE_BYTES + 1))
    m = xof_scalars("lwe-m", cnt)
    m[0] = 0
    m[1] = P - 1
    ref.set_entropy(ent)
    sk = ref.key_gen()
    off = 2 * CTR_CT * 256 + CTR_CT  # CTR_BV of the debug instance: starts mid-AES-block
    recs, cts = ref.encrypt(SEED, off, sk, m, want_ct=True)
    assert ref.entropy_consumed() == ent.size
    dec = [ref.decrypt(sk, cts[k]) for k in range(cnt)]
    assert dec == [int(v) for v in m]
    g["lwe"] = dict(offset=off, count=cnt, sk_sha=sha(sk), m=[int(v) for v in m], records=hexs(recs),
                    dotp0=hexs(ref.dotp(cts[0][:N], sk)))
    sm = []
    for k in range(6):
        e81 = xof(f"smudge{k}", 81)
        ref.set_entropy(e81)
        out, neg = ref.ct_smudge(cts[0])
        sm.append(dict(b=hexs(out[N]), negative=neg, dec=ref.decrypt(sk, out) if not neg else None))
    g["smudge"] = sm
    ref.clear_entropy()

    # ---- full SNARK on the small instance D=64, M=16 (snark.c:35-250, ssp.c:37-77)
    rs = Reference(64, 16)
    D, M = rs.D, rs.M
    n_ent = (M // 8 + M * 8 * D) + 40 + 24 + N * CT_BYTES + (2 * D + M) * (NOISE_BYTES + 1) + 8 + 5 * 81
    ent = xof("snark-entropy-d64-m16", n_ent)
    rs.set_entropy(ent)
    ssp, wit = rs.random_ssp()
    crs = rs.setup(ssp)
    proof, siz = rs.prover(ssp, crs, wit)
    assert rs.entropy_consumed() == n_ent, (rs.entropy_consumed(), n_ent)
    ok = rs.verifier(ssp, crs, proof)
    bad = proof.copy()
    bad[0, N, 0] ^= np.uint64(1 << 40)
    ok_bad = rs.verifier(ssp, crs, bad)
    rs.clear_entropy()
    g["snark_d64_m16"] = dict(
        D=D, M=M, entropy_bytes=n_ent, ssp_sha=sha(ssp), witness=hexs(wit), seed=hexs(crs["seed"]),
        alpha=crs["alpha"], beta=crs["beta"], s_point=crs["s_point"], sk_sha=sha(crs["sk"]),
        crs_s_sha=sha(crs["s"]), crs_as_sha=sha(crs["as_"]), crs_v_sha=sha(crs["v"][: M - 1]),
        crs_t=hexs(crs["t"]), crs_s0=hexs(crs["s"][0]), proof_sha=[sha(proof[k]) for k in range(5)],
        proof_b=[hexs(proof[k][N]) for k in range(5)], proof_negative=[bool(siz[k][N] < 0) for k in range(5)],
        accept=ok, accept_tampered=ok_bad)

    out = Path(__file__).with_name("vectors.json")
    out.write_text(json.dumps(g, indent=1) + "\n")
    print(f"wrote {out} ({out.stat().st_size} bytes); snark accept={ok}, tampered accept={ok_bad}")


if __name__ == "__main__":
    main()
