"""Shared fixtures.  `gpu` marks tests that need a B200; everything else runs on CPU only.

tests/ is the only place (with __graft_entry__.smoke() and bench.py's CPU legs) that may load the
checkers under oracle/.
"""
import hashlib
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than ~20 s on CPU")


def xof(label: str, n: int) -> np.ndarray:
    """Deterministic test bytes: SHAKE-256(label) — independent of numpy's RNG streams."""
    return np.frombuffer(hashlib.shake_256(label.encode()).digest(n), dtype=np.uint8).copy()


def xof_scalars(label: str, n: int, p: int = 0xFFFFFFFB) -> np.ndarray:
    """n scalars in [0, p) from 8-byte draws reduced mod p (as rand_modp, lwe.h:97-103)."""
    raw = xof(label, 8 * n).view("<u8")
    return (raw % np.uint64(p)).astype(np.uint64)


def xof_records(label: str, n: int) -> np.ndarray:
    """n wire records of 92 bytes whose top 4 bytes are zero (what ct_export emits)."""
    rec = xof(label, 92 * n).reshape(n, 92)
    rec[:, 88:] = 0
    return rec


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


SEED = bytes(range(40))  # the survey's probe seed: block 0 = 8477f45516027713a26a881ae67882bf


@pytest.fixture(scope="session")
def oracle():
    from oracle.loader import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The compiled reference (debug instance D=256, M=64); skipped when it cannot be had."""
    from oracle.loader import Reference
    try:
        return Reference(256, 64)
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"compiled reference unavailable: {e}")


@pytest.fixture(scope="session")
def reference_small():
    from oracle.loader import Reference
    try:
        return Reference(64, 16)
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"compiled reference unavailable: {e}")
