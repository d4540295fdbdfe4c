"""CRS / proof persistence (host/mf_io.c) on the GPU: a CRS written by mf_crs_write, re-read by
a FRESH PROCESS, made resident and proved there gives the golden proof — the one the compiled reference emitted for
the same entropy (tests/golden/make_golden.py) — and the proof file it writes is accepted after mf_proof_read."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import sha, xof

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((Path(__file__).parent / "golden" / "vectors.json").read_text())
N, LIMBS = 1470, 12


def proof_arrays(sn):
    """The five proof elements of a Snark object as (1471, 12) u64 magnitudes + the sign of each coordinate."""
    out, neg = np.zeros((5, N + 1, LIMBS), np.uint64), np.zeros((5, N + 1), bool)
    for k, el in enumerate((sn.proof.h, sn.proof.hat_h, sn.proof.hat_v, sn.proof.v_w, sn.proof.b_w)):
        for i in range(N + 1):
            n = el[i].size
            neg[k, i] = n < 0
            for j in range(abs(n)):
                out[k, i, j] = el[i].d[j]
    return out, neg


@pytest.mark.parametrize("devices", [1, 2])
def test_crs_written_reloaded_in_a_fresh_process_proves_to_the_golden_proof(tmp_path, devices):
    from c_lwe_snarks_b200.snark import Snark
    from oracle.loader import DropIn
    g = GOLD["snark_d64_m16"]
    D, M = g["D"], g["M"]
    shim = DropIn(D, M)  # only for its entropy hook (the library instance is shared with Snark)
    ent = xof("snark-entropy-d64-m16", g["entropy_bytes"])
    sn = Snark(D, M)
    try:
        shim.set_entropy(ent)
        sn.random_ssp()
        sn.setup()
        used = shim.entropy_consumed()
        shim.clear_entropy()
        assert used == g["entropy_bytes"] - 8 - 5 * 81  # what is left is the prover's: delta, then the smudging
        seed, s, as_, t, v = sn.crs_records()
        assert sha(s) == g["crs_s_sha"] and sha(as_) == g["crs_as_sha"]
        sn.save_crs(tmp_path / "crs.mfuoco")
        np.save(tmp_path / "ssp.npy", sn.ssp)
        wl = np.array([sn.witness.d[j] for j in range(abs(sn.witness.size))], np.uint64)
        np.save(tmp_path / "wit.npy", wl)
        ent[used:].tofile(tmp_path / "ent.bin")
        r = subprocess.run([sys.executable, str(ROOT / "tests" / "reload_prove_helper.py"), str(D), str(M),
                            str(tmp_path / "crs.mfuoco"), str(tmp_path / "ssp.npy"), str(tmp_path / "wit.npy"),
                            str(tmp_path / "ent.bin"), str(tmp_path / "proof.mfuoco"), str(devices)],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        sn.load_proof(tmp_path / "proof.mfuoco")
        arr, neg = proof_arrays(sn)
        for k in range(5):
            assert sha(arr[k]) == g["proof_sha"][k], f"proof element {k} differs from the golden proof"
            assert bool(neg[k, N]) == g["proof_negative"][k]
        ok, _ = sn.verify()  # this process still holds the verification key of the setup above
        assert ok
        sn.tamper()
        assert not sn.verify()[0]
    finally:
        shim.clear_entropy()
        sn.close()
