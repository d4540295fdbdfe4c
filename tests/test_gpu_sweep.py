"""BASELINE configs[4] (LWE parameter sweep): the generic-parameter lincomb kernel against plain integer arithmetic.
The reference implements only (n, log q) = (1470, 736), so other points have no reference output; q_eff = 2^(64 L)."""
import numpy as np
import pytest

from conftest import xof

pytestmark = pytest.mark.gpu


def to_tile_planar(flat: np.ndarray) -> np.ndarray:
    """(d, ncoords, L) u64 -> (d, T*L*64) in the layout mfb_lincomb_generic_dev documents."""
    d, nc, L = flat.shape
    T = (nc + 63) // 64
    pad = np.zeros((d, T * 64, L), np.uint64)
    pad[:, :nc] = flat
    return np.ascontiguousarray(pad.reshape(d, T, 64, L).transpose(0, 1, 3, 2)).reshape(d, T * L * 64)


def from_tile_planar(tp: np.ndarray, nc: int, L: int) -> np.ndarray:
    T = (nc + 63) // 64
    return np.ascontiguousarray(tp.reshape(T, L, 64).transpose(0, 2, 1)).reshape(T * 64, L)[:nc]


@pytest.mark.parametrize("L,nc,d", [(11, 1471, 50), (4, 65, 7), (8, 1025, 33), (12, 1471, 21), (16, 2048, 9), (13, 700, 130),
                                    (6, 64, 1), (10, 1300, 5), (14, 100, 257)])
def test_generic_lincomb_vs_python(L, nc, d):
    import torch

    import c_lwe_snarks_b200 as m
    ctx = m.Context(0)
    try:
        flat = xof(f"sweep-{L}-{nc}-{d}", d * nc * L * 8).view("<u8").reshape(d, nc, L).copy()
        h = (xof(f"sweep-h-{L}-{nc}-{d}", 4 * d).view("<u4")).astype(np.uint32)
        h[0] = 0xFFFFFFFF
        d_cts = torch.from_numpy(to_tile_planar(flat).view(np.int64)).cuda()
        d_h = torch.from_numpy(h.view(np.int32)).cuda()
        T = (nc + 63) // 64
        d_out = torch.zeros(T * L * 64, dtype=torch.int64, device="cuda")
        for _ in range(2):  # twice: the chunk queues must have been re-armed
            ctx.lincomb_generic_dev(L, nc, d_cts.data_ptr(), d_h.data_ptr(), d, d_out.data_ptr())
        torch.cuda.synchronize()
        got = from_tile_planar(d_out.cpu().numpy().view(np.uint64), nc, L)
        mod = 1 << (64 * L)
        for c in list(range(0, nc, max(1, nc // 40))) + [nc - 1]:
            want = sum(int(h[i]) * int.from_bytes(flat[i, c].tobytes(), "little") for i in range(d)) % mod
            assert int.from_bytes(got[c].tobytes(), "little") == want, f"coordinate {c}"
    finally:
        ctx.close()
