"""The product behind the reference's OWN C interface (include/mangiafuoco, libmangiafuoco_b200.so).

Three kinds of evidence, all on the GPU:
  1. the full SNARK (random_ssp -> crs_init -> setup -> prover -> verifier) under injected entropy reproduces the
     golden vectors that the compiled reference emitted (tests/golden/make_golden.py): CRS, proof and accept bit;
  2. the same run beside the compiled reference itself (oracle/_ref/libmfref_*.so on the host CPU), larger instance;
  3. the reference's five test programs, UNMODIFIED, compiled against the drop-in headers and linked with the
     product (oracle/Makefile `dropin-tests`), exit 0.
"""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import SEED, sha, xof, xof_records, xof_scalars

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((Path(__file__).parent / "golden" / "vectors.json").read_text())
N, CT_BYTES, CTR_CT, P = 1470, 92, 92 * 1470, 0xFFFFFFFB


@pytest.fixture(scope="module")
def dropin():
    from oracle.loader import DropIn
    return DropIn(64, 16)


def hexs(a) -> str:
    return np.ascontiguousarray(a).tobytes().hex()


def snark_entropy_len(D, M):
    return (M // 8 + M * 8 * D) + 40 + 24 + N * CT_BYTES + (2 * D + M) * 70 + 8 + 5 * 81


def run_snark(lib, D, M, ent):
    lib.set_entropy(ent)
    ssp, wit = lib.random_ssp()
    crs = lib.setup(ssp)
    proof, siz = lib.prover(ssp, crs, wit)
    used = lib.entropy_consumed()
    ok = lib.verifier(ssp, crs, proof)
    bad = proof.copy()
    bad[0, N, 0] ^= np.uint64(1 << 40)
    ok_bad = lib.verifier(ssp, crs, bad)
    lib.clear_entropy()
    return dict(ssp=ssp, wit=wit, crs=crs, proof=proof, siz=siz, used=used, ok=ok, ok_bad=ok_bad)


def test_full_snark_golden(dropin):
    g = GOLD["snark_d64_m16"]
    D, M = g["D"], g["M"]
    dropin.set_instance(D, M)
    ent = xof("snark-entropy-d64-m16", g["entropy_bytes"])
    launches0 = dropin.gpu_launches()
    r = run_snark(dropin, D, M, ent)
    assert r["used"] == g["entropy_bytes"], "entropy draws differ from the reference's sequence"
    assert sha(r["ssp"]) == g["ssp_sha"] and hexs(r["wit"]) == g["witness"]
    crs = r["crs"]
    assert hexs(crs["seed"]) == g["seed"]
    assert (crs["alpha"], crs["beta"], crs["s_point"]) == (g["alpha"], g["beta"], g["s_point"])
    assert sha(crs["sk"]) == g["sk_sha"]
    assert hexs(crs["s"][0]) == g["crs_s0"]
    assert sha(crs["s"]) == g["crs_s_sha"] and sha(crs["as_"]) == g["crs_as_sha"]
    assert sha(crs["v"][: M - 1]) == g["crs_v_sha"] and hexs(crs["t"]) == g["crs_t"]
    for k in range(5):
        assert hexs(r["proof"][k][N]) == g["proof_b"][k], f"proof element {k}: b differs"
        assert sha(r["proof"][k]) == g["proof_sha"][k], f"proof element {k} differs"
        assert bool(r["siz"][k][N] < 0) == g["proof_negative"][k]
    assert r["ok"] == g["accept"] and r["ok_bad"] == g["accept_tampered"]
    assert r["ok"] and not r["ok_bad"]
    assert dropin.gpu_launches() > launches0, "no GPU kernel ran: the drop-in must not compute on the CPU"


def test_prover_with_resident_crs_is_identical(dropin):
    g = GOLD["snark_d64_m16"]
    D, M = g["D"], g["M"]
    dropin.set_instance(D, M)
    ent = xof("snark-entropy-d64-m16", g["entropy_bytes"])
    dropin.set_entropy(ent)
    ssp, wit = dropin.random_ssp()
    crs = dropin.setup(ssp)
    proof, _ = dropin.prover_resident(ssp, crs, wit)
    dropin.clear_entropy()
    for k in range(5):
        assert sha(proof[k]) == g["proof_sha"][k]


@pytest.mark.parametrize("n", [2, 3])
def test_prover_with_crs_sharded_over_a_device_set_is_identical(dropin, n):
    """mf_set_devices(n): the resident regions are sharded by ciphertext index over n members (here all on GPU 0, or one
    per GPU when the box has them) and every lincomb is combined over peer memory: the proof does not change."""
    import torch
    g = GOLD["snark_d64_m16"]
    D, M = g["D"], g["M"]
    dropin.set_instance(D, M)
    spread = 1 if torch.cuda.device_count() >= n else 0
    dropin.lib.mf_set_devices(n, spread)
    try:
        dropin.set_entropy(xof("snark-entropy-d64-m16", g["entropy_bytes"]))
        ssp, wit = dropin.random_ssp()
        crs = dropin.setup(ssp)
        proof, _ = dropin.prover_resident(ssp, crs, wit)
        dropin.clear_entropy()
    finally:
        dropin.lib.mf_set_devices(1, 1)
    for k in range(5):
        assert sha(proof[k]) == g["proof_sha"][k]


def test_prover_sharded_over_a_device_set_without_residency(dropin):
    """mf_set_devices(2) with a NON-resident CRS: both fused passes are sharded by ciphertext index (every member
    regenerates its a-vectors from AES) — same proof."""
    import torch
    g = GOLD["snark_d64_m16"]
    dropin.set_instance(g["D"], g["M"])
    dropin.lib.mf_set_devices(2, 1 if torch.cuda.device_count() >= 2 else 0)
    try:
        dropin.set_entropy(xof("snark-entropy-d64-m16", g["entropy_bytes"]))
        ssp, wit = dropin.random_ssp()
        crs = dropin.setup(ssp)
        proof, _ = dropin.prover(ssp, crs, wit)
        dropin.clear_entropy()
    finally:
        dropin.lib.mf_set_devices(1, 1)
    for k in range(5):
        assert sha(proof[k]) == g["proof_sha"][k]


def test_setup_over_a_device_set_is_identical(dropin):
    """mf_set_devices(2): setup()'s encryptions are spread over the members piece by piece; the CRS does not change."""
    import torch
    g = GOLD["snark_d64_m16"]
    dropin.set_instance(g["D"], g["M"])
    dropin.lib.mf_set_devices(2, 1 if torch.cuda.device_count() >= 2 else 0)
    try:
        dropin.set_entropy(xof("snark-entropy-d64-m16", g["entropy_bytes"]))
        ssp, wit = dropin.random_ssp()
        crs = dropin.setup(ssp)
        dropin.clear_entropy()
    finally:
        dropin.lib.mf_set_devices(1, 1)
    assert sha(crs["s"]) == g["crs_s_sha"] and sha(crs["as_"]) == g["crs_as_sha"]
    assert sha(crs["v"][: g["M"] - 1]) == g["crs_v_sha"] and hexs(crs["t"]) == g["crs_t"]


def test_prover_with_only_one_region_resident_is_identical(dropin, monkeypatch):
    """When only one of the two CRS regions fits in HBM (D = 2^20 on one GPU) mf_crs_make_resident keeps that one and
    prover() mixes a resident pass with a fused one; $MF_B200_ONE_REGION forces that path at a small size."""
    g = GOLD["snark_d64_m16"]
    dropin.set_instance(g["D"], g["M"])
    monkeypatch.setenv("MF_B200_ONE_REGION", "1")
    dropin.set_entropy(xof("snark-entropy-d64-m16", g["entropy_bytes"]))
    ssp, wit = dropin.random_ssp()
    crs = dropin.setup(ssp)
    proof, _ = dropin.prover_resident(ssp, crs, wit)
    dropin.clear_entropy()
    for k in range(5):
        assert sha(proof[k]) == g["proof_sha"][k]


def test_full_snark_beside_compiled_reference(dropin, reference):
    D, M = reference.D, reference.M  # 256, 64
    dropin.set_instance(D, M)
    ent = xof("snark-entropy-d256-m64", snark_entropy_len(D, M))
    a = run_snark(dropin, D, M, ent)
    b = run_snark(reference, D, M, ent)
    assert a["used"] == b["used"] == ent.size
    assert np.array_equal(a["ssp"], b["ssp"]) and np.array_equal(a["wit"], b["wit"])
    for key in ("seed", "s", "as_", "t", "sk"):
        assert np.array_equal(a["crs"][key], b["crs"][key]), key
    assert np.array_equal(a["crs"]["v"][: M - 1], b["crs"]["v"][: M - 1])
    assert (a["crs"]["alpha"], a["crs"]["beta"], a["crs"]["s_point"]) == (b["crs"]["alpha"], b["crs"]["beta"], b["crs"]["s_point"])
    assert np.array_equal(a["proof"], b["proof"]) and np.array_equal(a["siz"] < 0, b["siz"] < 0)
    assert a["ok"] and b["ok"] and not a["ok_bad"] and not b["ok_bad"]
    # cross-verification: each side accepts the other's proof
    assert reference.verifier(b["ssp"], b["crs"], a["proof"]) and dropin.verifier(a["ssp"], a["crs"], b["proof"])


def test_full_snark_config1_beside_compiled_reference(dropin):
    """BASELINE configs[0]: default LWE parameters, ~2^10-constraint SSP (D = 1024, M = 64), reference CPU path beside ours."""
    from oracle.loader import Reference
    try:
        ref = Reference(1024, 64)
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"compiled reference for D=1024 unavailable: {e}")
    D, M = 1024, 64
    dropin.set_instance(D, M)
    ent = xof("snark-entropy-d1024-m64", snark_entropy_len(D, M))
    a = run_snark(dropin, D, M, ent)
    b = run_snark(ref, D, M, ent)
    assert a["used"] == b["used"] == ent.size
    for key in ("seed", "s", "as_", "t", "sk"):
        assert np.array_equal(a["crs"][key], b["crs"][key]), key
    assert np.array_equal(a["crs"]["v"][: M - 1], b["crs"]["v"][: M - 1])
    assert np.array_equal(a["proof"], b["proof"]) and np.array_equal(a["siz"] < 0, b["siz"] < 0)
    assert a["ok"] and b["ok"] and not a["ok_bad"] and not b["ok_bad"]


def test_lwe_api_beside_compiled_reference(dropin, reference):
    """ct_import / eval_poly / ct_mul_ui / ct_add / ct_addmul_ui / encrypt / decrypt / dotp / smudge / urandomb / modq."""
    dropin.set_instance(256, 64)
    d, off = 40, 3 * CTR_CT + 8
    c8, h = xof_records("api-c8", d), xof_scalars("api-h", d)
    for lib in (dropin, reference):
        lib.clear_entropy()
    assert np.array_equal(dropin.stream(SEED, 11, 1000), reference.stream(SEED, 11, 1000))
    assert np.array_equal(dropin.stream(SEED, 0, 5000, chunk=92), reference.stream(SEED, 0, 5000, chunk=92))
    for nbits in (64, 1, 5, 32, 40, 520, 512, 700, 736, 751):
        assert dropin.urandomb(SEED, 11, nbits)[1] == reference.urandomb(SEED, 11, nbits)[1]
        assert np.array_equal(dropin.urandomb(SEED, 11, nbits)[0], reference.urandomb(SEED, 11, nbits)[0])
    for k, nl in enumerate([5, 11, 12, 13, 14]):
        x = xof(f"modq{k}", 8 * nl).view("<u8")
        assert np.array_equal(dropin.modq(x)[0], reference.modq(x)[0]) and dropin.modq(x)[1] == reference.modq(x)[1]
    x = reference.ct_import(SEED, off, c8[0])
    assert np.array_equal(dropin.ct_import(SEED, off, c8[0]), x)  # full 736-bit a_j
    y = reference.ct_import(SEED, off + CTR_CT, c8[1])
    assert np.array_equal(dropin.ct_export(x), reference.ct_export(x))
    assert np.array_equal(dropin.eval_poly(SEED, off, c8, h), reference.eval_poly(SEED, off, c8, h))
    assert np.array_equal(dropin.ct_mul_ui(x, int(h[0])), reference.ct_mul_ui(x, int(h[0])))
    assert np.array_equal(dropin.ct_add(x, y), reference.ct_add(x, y))
    assert np.array_equal(dropin.ct_addmul_ui(x, y, int(h[1])), reference.ct_addmul_ui(x, y, int(h[1])))
    cnt = 3
    ent = xof("api-ent", N * CT_BYTES + cnt * 70 + 81)
    m = xof_scalars("api-m", cnt)
    out = {}
    for lib in (dropin, reference):
        lib.set_entropy(ent)
        sk = lib.key_gen()
        recs, cts = lib.encrypt(SEED, off, sk, m, want_ct=True)
        dec = [lib.decrypt(sk, cts[i]) for i in range(cnt)]
        dot = lib.dotp(cts[0][:N], sk)
        sm, neg = lib.ct_smudge(cts[0])
        dec_sm = lib.decrypt(sk, sm)
        assert lib.entropy_consumed() == ent.size
        lib.clear_entropy()
        out[lib.kind] = (sk, recs, cts, dec, dot, sm, neg, dec_sm)
    a, b = out["dropin"], out["reference"]
    for i in (0, 1, 2, 4, 5):
        assert np.array_equal(a[i], b[i]), i
    assert a[3] == b[3] == [int(v) for v in m] and a[6] == b[6] and a[7] == b[7] == int(m[0])


@pytest.mark.parametrize("name", ["aes", "entropy", "ssp", "lwe", "snark"])
def test_reference_test_programs_against_the_dropin(name):
    exe = ROOT / "oracle" / "_ref" / f"dropin_test_{name}"
    if not exe.exists():
        pytest.skip("built only where the reference sources are present (oracle/Makefile dropin-tests)")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_reference_benchmark_snark_against_the_dropin():
    """benchmark_snark.c of the reference, unmodified (debug instance): prints setup/prover/verifier seconds in its own
    format and exits 0 iff the verifier accepts (benchmark_snark.c:94-96)."""
    exe = ROOT / "oracle" / "_ref" / "dropin_benchmark_snark_debug"
    if not exe.exists():
        pytest.skip("built only where the reference sources are present (oracle/Makefile dropin-tests)")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    names = [line.split("\t")[0] for line in r.stdout.splitlines() if "\t" in line]
    assert names == ["setup", "prover", "verifier"]


def test_python_snark_binding_roundtrip():
    """c_lwe_snarks_b200.snark.Snark (ctypes over the drop-in's own structs) with OS entropy: accept, then reject."""
    from c_lwe_snarks_b200.snark import Snark
    sn = Snark(128, 16)
    try:
        sn.random_ssp()
        sn.setup()
        sn.prove()
        ok, _ = sn.verify()
        assert ok
        sn.make_resident()
        sn.prove()
        assert sn.verify()[0]
        sn.tamper()
        assert not sn.verify()[0]
        assert sn.gpu_launches() > 0
    finally:
        sn.close()


def test_setup_over_a_resident_crs_drops_the_stale_regions():
    """setup() rewrites crs->s / crs->as in place: a resident copy made before it (mf_crs_make_resident) would be stale,
    and prover() finds regions by crs pointer.  setup() must drop it — the proof over the NEW records verifies."""
    import ctypes as C

    from c_lwe_snarks_b200.snark import Snark
    sn = Snark(128, 16)
    try:
        sn.random_ssp()
        sn.setup()
        sn.make_resident()
        sn.prove()
        assert sn.verify()[0]
        sn.lib.key_clear(sn.vrs.sk)
        sn.lib.setup(C.byref(sn.crs), C.byref(sn.vrs), sn._ssp_ptr())  # same crs object, new alpha / beta / s / sk / records
        sn.prove()  # must not use the regions expanded from the old records
        assert sn.verify()[0]
        sn.make_resident()
        sn.prove()
        assert sn.verify()[0]
    finally:
        sn.close()


def test_auto_resident_ssp_detects_a_regenerated_blob():
    """setup() / prover() keep the SSP blob they are given resident on the device by themselves.  A blob that is
    REGENERATED IN PLACE (same address, new contents) must not be served from the stale copy: the fingerprint check
    drops it and the new instance proves and verifies."""
    from c_lwe_snarks_b200.snark import Snark
    sn = Snark(128, 16)
    try:
        for _ in range(3):  # same buffer, three different instances
            sn.random_ssp()
            sn.setup()
            sn.prove()
            assert sn.verify()[0]
        sn.tamper()
        assert not sn.verify()[0]
    finally:
        sn.lib.mf_ssp_release(sn._ssp_ptr())
        sn.close()


def test_host_blob_path_without_auto_residency(dropin, monkeypatch):
    """$MF_B200_NO_AUTO_SSP: setup() and prover() stream the dense blob from host memory on every call (the path a blob
    too large for the device takes) — same CRS, same proof."""
    g = GOLD["snark_d64_m16"]
    dropin.set_instance(g["D"], g["M"])
    monkeypatch.setenv("MF_B200_NO_AUTO_SSP", "1")
    r = run_snark(dropin, g["D"], g["M"], xof("snark-entropy-d64-m16", g["entropy_bytes"]))
    assert sha(r["crs"]["s"]) == g["crs_s_sha"] and sha(r["crs"]["as_"]) == g["crs_as_sha"]
    for k in range(5):
        assert sha(r["proof"][k]) == g["proof_sha"][k]
    assert r["ok"] and not r["ok_bad"]
