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
