"""CPU: the C-ABI libraries load and export every symbol their headers declare; without a GPU the product
fails loudly instead of computing on the CPU."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared(header: Path, pattern: str):
    return sorted(set(re.findall(pattern, header.read_text())))


def test_mfb200_exports_every_declared_symbol():
    from c_lwe_snarks_b200.api import EXPORTS, library_path, load_library
    names = declared(ROOT / "include" / "mfb200.h", r"MFB_API [\w \*]*?\b(mfb_\w+)\(")
    assert len(names) >= 25
    lib = ctypes.CDLL(str(library_path()))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mfb200.h but not exported"
    assert set(EXPORTS) == set(names), "python binding and header disagree"
    load_library()


def test_dropin_exports_the_reference_interface():
    lib = ctypes.CDLL(str(ROOT / "c_lwe_snarks_b200" / "lib" / "libmangiafuoco_b200.so"))
    # snark.h:44-51, lwe.h:36-74 (+ the two externally linked but undeclared ct_addmul_ui / ct_zero), entropy.h, aes.h, ssp.h
    for n in ["crs_init", "crs_clear", "proof_init", "proof_clear", "setup", "prover", "verifier",
              "key_gen", "key_clear", "errdist_uniform", "ct_init", "ct_clear", "ct_export", "ct_import",
              "decompress_encryption", "regev_encrypt2", "mpz_add_dotp", "regev_decrypt", "ct_smudge", "ct_add",
              "ct_mul_ui", "ct_addmul_ui", "ct_zero", "eval_poly", "rng_init", "rng_clear", "rng_seek",
              "mpz2_urandomb", "mpz2_urandomb2", "mpz_entropy_init", "aesctr_init", "aesctr_prg", "aesctr_clear",
              "nmod_poly_import", "nmod_poly_export", "random_ssp"]:
        assert hasattr(lib, n), n
    # the additions of include/mangiafuoco/mangiafuoco_b200.h (INTEGRATION.md "What is new")
    for n in ["mf_set_instance", "mf_set_entropy_source", "mf_entropy", "mf_set_device", "mf_set_devices",
              "mf_crs_make_resident", "mf_crs_release", "mf_ssp_make_resident", "mf_ssp_release", "mf_crs_write",
              "mf_crs_read", "mf_proof_write", "mf_proof_read", "mf_gpu_launches"]:
        assert hasattr(lib, n), n


def test_no_cpu_fallback():
    import torch

    import c_lwe_snarks_b200 as m
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(m.MfbError, match="no CPU fallback"):
        m.Context(0)


def test_product_does_not_reference_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/: nothing under the product
    package imports, links, loads or even names it."""
    for p in (ROOT / "c_lwe_snarks_b200").rglob("*"):
        if p.is_file() and (p.suffix in {".py", ".c", ".h", ".cu", ".cuh"} or p.name == "Makefile"):
            assert "oracle" not in p.read_text().lower(), p
