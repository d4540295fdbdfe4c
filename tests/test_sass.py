"""CPU: properties of the compiled sm_100a code that the design relies on, read with cuobjdump (no GPU needed).

These are the claims DESIGN.md makes about the machine code — TMA bulk copies completing on mbarriers in K1, carry-chained
IMAD.WIDE in the MAC, REDUX in the encryption tail, no spills in the hot kernels — checked on the library the tests and
the bench actually load."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "c_lwe_snarks_b200" / "lib" / "libmfb200.so"

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")


@pytest.fixture(scope="module")
def usage():
    out = subprocess.run(["cuobjdump", "-res-usage", str(LIB)], capture_output=True, text=True, check=True).stdout
    res = {}
    for name, line in re.findall(r"Function (\S+):\n\s*(REG:.*)", out):
        res[name] = {k: int(v) for k, v in re.findall(r"(\w+)(?:\[\d+\])?:(\d+)", line)}
    assert res, "no kernels found in the library"
    return res


_SASS = None


def sass_of(pattern: str) -> str:
    global _SASS
    if _SASS is None:
        _SASS = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    out = _SASS
    chunks = re.split(r"\n\s*Function : ", out)
    sel = [c for c in chunks if re.match(pattern, c)]
    assert sel, f"no kernel matches {pattern}"
    return "\n".join(sel)


def test_built_for_sm_100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", str(LIB)], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"sm_(\d+\w?)", out))
    assert archs == {"100a"}, archs


def test_hot_kernels_do_not_spill(usage):
    hot = [n for n in usage if re.search(r"k_lincombILi[12]|k_evalpolyILi[12]|k_expand|k_lincomb_finish|k_peer_allreduce|"
                                         r"k_ntt_cols|k_ntt_local|k_decrypt|k_bcoord", n)]
    assert len(hot) >= 12
    for n in hot:
        assert usage[n]["STACK"] == 0 and usage[n]["LOCAL"] == 0, (n, usage[n])
    # 512-thread AES kernels: one CTA per SM needs <= 128 registers; the 640-thread k_encrypt starts at <= 102 and
    # re-balances its register file between producer and consumer warpgroups (setmaxnreg): no spills either
    for n, u in usage.items():
        if re.search(r"k_evalpolyILi[12]|k_expand|k_stream_bytes", n):
            assert u["REG"] <= 128, (n, u)
        if "k_encrypt" in n:
            assert u["REG"] <= 102 and u["STACK"] == 0 and u["LOCAL"] == 0, (n, u)


def test_k1_streams_with_tma_bulk_copies_on_mbarriers():
    s = sass_of(r"_ZN3mfb9k_lincombILi1")
    assert "UBLKCP" in s, "cp.async.bulk (TMA) missing from k_lincomb"
    assert "SYNCS" in s, "mbarrier instructions missing from k_lincomb"
    assert len(re.findall(r"IMAD\.WIDE\.U32(\.X)?", s)) >= 22, "704-bit MAC is not 22 IMAD.WIDE"


def test_aes_kernels_use_prmt_addressing_and_shared_tables():
    s = sass_of(r"_ZN3mfb8k_expand")
    assert s.count("PRMT") > 400 and s.count("LDS") > 400
    # the round key folded into the rotated half: fewer LOP3 than LDS (was 373 LOP3 against 471 LDS before the folding)
    assert s.count("LOP3") < 0.7 * s.count("LDS")


def test_encrypt_tail_uses_redux_and_shared_atomics():
    s = sass_of(r"_ZN3mfb9k_encrypt")
    assert "REDUX" in s and "ATOMS" in s
    assert "USETMAXREG.DEALLOC" in s and "USETMAXREG.TRY_ALLOC" in s, "producer / consumer register re-balancing missing"
    assert len(re.findall(r"IMAD\.WIDE\.U32", s)) >= 250, "low-half 22x22 product is 253 limb products"


def test_two_vector_fused_pass_keeps_the_full_tile_budget(usage):
    """k_evalpoly<2> (round 2): two CANONICAL accumulators in the register budget of one E/O accumulator — at most a few
    registers more than k_evalpoly<1>, no spills, and both carry chains of both vectors as IMAD.WIDE.U32.X"""
    r1 = next(u for n, u in usage.items() if "k_evalpolyILi1" in n)
    r2 = next(u for n, u in usage.items() if "k_evalpolyILi2" in n)
    assert r2["REG"] <= r1["REG"] + 8 and r2["REG"] <= 128 and r2["STACK"] == 0
    s1, s2 = sass_of(r"_ZN3mfb10k_evalpolyILi1"), sass_of(r"_ZN3mfb10k_evalpolyILi2")
    x1, x2 = len(re.findall(r"IMAD\.WIDE\.U32\.X", s1)), len(re.findall(r"IMAD\.WIDE\.U32\.X", s2))
    assert x1 >= 20 and x2 >= 2 * x1 - 4, (x1, x2)
    # the same AES per tile: the lookups (LDS) of the two kernels differ only by a handful
    assert abs(s1.count("LDS") - s2.count("LDS")) <= 8


def test_polynomial_step_has_the_fused_middle_kernel(usage):
    """forward local stages + pointwise product + inverse local stages of a product are ONE kernel (k_ntt_local_mul); the
    separate pointwise kernel is gone"""
    names = list(usage)
    assert any("k_ntt_local_mul" in n for n in names)
    assert not any("k_pointwise" in n for n in names)
    u = next(u for n, u in usage.items() if "k_ntt_local_mul" in n)
    assert u["STACK"] == 0 and u["REG"] <= 40 and u["SHARED"] <= 12288 + 1024  # one block + one twiddle table (+ the driver's 1 KB)


def test_generic_encrypt_kernels_do_not_spill(usage):
    gen = [(n, u) for n, u in usage.items() if "k_encrypt_gILi" in n]
    assert len(gen) == 9  # limb counts 4, 6, 8, 10, 11, 12, 13, 14, 16
    for n, u in gen:
        assert u["STACK"] == 0 and u["LOCAL"] == 0 and u["REG"] <= 102, (n, u)
