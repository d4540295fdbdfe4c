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


def sk_planar(sk_flat: np.ndarray, stride: int) -> np.ndarray:
    """(n, L) u64 -> row-planar [L][stride] (limb row j of coordinate c at j * stride + c)."""
    n, L = sk_flat.shape
    out = np.zeros((L, stride), np.uint64)
    out[:, :n] = sk_flat.T
    return out


@pytest.mark.parametrize("L,n,ctb,cnt", [(11, 1470, 92, 3), (8, 1024, 64, 5), (10, 1246, 80, 4), (12, 1470, 100, 3), (14, 1470, 112, 2),
                                         (16, 2047, 128, 2), (4, 65, 32, 9), (6, 700, 52, 300), (13, 33, 104, 150),
                                         (12, 1600, 96, 2), (8, 600, 64, 200), (16, 320, 128, 170)])
def test_generic_encrypt_vs_python(L, n, ctb, cnt, oracle):
    """b_k = (e_k p + <sk, a_k> + m_k) mod 2^(64 L) with a_k from the AES-CTR stream at coordinate width ctb = log q / 8 —
    against plain integers over the oracle's keystream (ragged tiles, mid-block stream offsets, maximal noise bytes; widths of
    64 / 96 / 128 / 32 bytes take the padded keystream layout of k_encrypt_g, the others the plain one)."""
    import torch

    import c_lwe_snarks_b200 as m
    from conftest import SEED, xof_scalars
    P = 0xFFFFFFFB
    off = 7 * 16 + 5  # not block aligned
    sk = xof(f"gsk-{L}-{n}", n * L * 8).view("<u8").reshape(n, L).copy()
    msg = xof_scalars(f"gm-{L}-{n}", cnt)
    ent_nb = min(8 * L, 69)
    ent = xof(f"ge-{L}-{n}", cnt * (ent_nb + 1)).reshape(cnt, ent_nb + 1).copy()
    ent[0, :ent_nb] = 0xFF  # maximal noise
    stride = (n + 63) // 64 * 64
    ctx = m.Context(0)
    try:
        d_sk = torch.from_numpy(sk_planar(sk, stride).view(np.int64).reshape(-1)).cuda()
        d_msg = torch.from_numpy(msg.view(np.int64)).cuda()
        d_ent = torch.from_numpy(ent.reshape(-1)).cuda()
        d_out = torch.zeros(cnt * ctb, dtype=torch.uint8, device="cuda")
        ctx.encrypt_generic_dev(L, n, ctb, SEED, off, d_sk.data_ptr(), stride, d_msg.data_ptr(), d_ent.data_ptr(), ent_nb + 1, ent_nb,
                                cnt, d_out.data_ptr())
        torch.cuda.synchronize()
        got = d_out.cpu().numpy().reshape(cnt, ctb)
    finally:
        ctx.close()
    mod = 1 << (64 * L)
    sk_int = [int.from_bytes(sk[j].tobytes(), "little") for j in range(n)]
    check = range(cnt) if cnt <= 10 else [0, 1, cnt // 2, cnt - 2, cnt - 1]
    for k in check:
        ks = oracle.stream(SEED, off + k * n * ctb, n * ctb).reshape(n, ctb)
        dot = sum(sk_int[j] * int.from_bytes(ks[j].tobytes(), "little") for j in range(n))
        e = int.from_bytes(ent[k, :ent_nb].tobytes(), "little")
        want = (e * P + dot + int(msg[k])) % mod
        assert int.from_bytes(got[k].tobytes(), "little") == want, f"ciphertext {k}"


def test_generic_encrypt_equals_the_reference_point_kernel(oracle):
    """(n, log q) = (1470, 736) through the generic kernel == the specialised k_encrypt (which is pinned to the reference)."""
    import torch

    import c_lwe_snarks_b200 as m
    from conftest import SEED, xof_scalars
    cnt, n, L, ctb = 148 * 3 + 7, 1470, 11, 92
    sk = oracle.key_gen(xof("sk-gen-eq", n * 92))[:, :11].copy()
    msg = xof_scalars("m-gen-eq", cnt)
    ent = xof("e-gen-eq", cnt * 70)
    ctx = m.Context(0)
    try:
        want = ctx.encrypt(SEED, 3 * 135240 + 8, sk, msg, ent)
        d_sk = torch.from_numpy(sk_planar(sk, 1472).view(np.int64).reshape(-1)).cuda()
        d_msg = torch.from_numpy(msg.view(np.int64)).cuda()
        d_ent = torch.from_numpy(ent).cuda()
        d_out = torch.zeros(cnt * ctb, dtype=torch.uint8, device="cuda")
        ctx.encrypt_generic_dev(L, n, ctb, SEED, 3 * 135240 + 8, d_sk.data_ptr(), 1472, d_msg.data_ptr(), d_ent.data_ptr(), 70, 69, cnt,
                                d_out.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy().reshape(cnt, ctb), want)
    finally:
        ctx.close()
