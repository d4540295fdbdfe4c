"""Wall time of the prover's polynomial step over a resident SSP (mfb_ssp_create + mfb_ssp_prover_polys_resident).
Usage: python tools/polys_time.py [log2d ...]"""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import c_lwe_snarks_b200 as m  # noqa: E402

P = 0xFFFFFFFB
ctx = m.Context(0)
for log2d in [int(a) for a in sys.argv[1:]] or [16, 20]:
    D, M = 1 << log2d, 64
    rng = np.random.Generator(np.random.PCG64(log2d))
    v = rng.integers(0, P, size=(M, D), dtype=np.uint64)
    wl = rng.integers(0, 1 << 63, size=1, dtype=np.uint64)
    t = v[0].copy()
    for i in range(1, M):
        if (int(wl[0]) >> (i - 1)) & 1:
            t = (t + v[i]) % np.uint64(P)
    t[0] = (t[0] + np.uint64(P - 1)) % np.uint64(P)   # t = v_0 + sum w_i v_i - 1 (ssp.c:37-77)
    blob = np.concatenate([t[None], v]).astype(np.uint64)
    res = ctx.ssp_resident(blob.view(np.uint8), D, M)
    res.prover_polys(wl, 12345)
    times = []
    for _ in range(5):
        t0 = time.perf_counter()
        res.prover_polys(wl, 12345)
        times.append(1e3 * (time.perf_counter() - t0))
    # the device-resident variant the fused prover pipeline uses (w | v | h stay on the device: no 24 D bytes of D2H)
    import ctypes as C
    ptr = C.c_void_p()
    wl_p = wl.ctypes.data_as(C.POINTER(C.c_uint64))
    dev = []
    for _ in range(6):
        t0 = time.perf_counter()
        ctx._ck(ctx.lib.mfb_ssp_prover_polys_resident_dev(ctx.h, res.handle, wl_p, wl.size, 12345, C.byref(ptr)))
        dev.append(1e3 * (time.perf_counter() - t0))
    print(json.dumps({"D": D, "M": M, "polys_resident_ms": [round(x, 3) for x in times],
                      "polys_resident_dev_ms": [round(x, 3) for x in dev[1:]]}), flush=True)
    res.close()
ctx.close()
