"""Device-side timings (CUDA events on the launching stream) of every kernel family at BASELINE sizes.
Usage: python tools/kernel_times.py [log2d]   -> one JSON line per measurement on stdout.
KT_ONCE=1: every kernel family exactly once, no host-call section (the command ncu captures)."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import c_lwe_snarks_b200 as m  # noqa: E402

SEED = bytes(range(40))
NC, NCP, L64, CT = 1471, 1472, 11, 92
CTR_CT = 92 * 1470
AES_BLOCKS_PER_CT = CTR_CT / 16


import os
ONCE = os.environ.get("KT_ONCE") == "1"


def timeit(fn, reps=5, warm=2):
    if ONCE:
        reps, warm = 1, 0
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    log2d = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    D = 1 << log2d
    ctx = m.Context(0)
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.Generator(np.random.PCG64(1))
    c8 = rng.integers(0, 256, size=(D, CT), dtype=np.uint8)
    c8[:, 88:] = 0
    h = rng.integers(0, m.P, size=D, dtype=np.uint64).astype(np.uint32)
    d_c8 = torch.from_numpy(c8.reshape(-1)).cuda()
    d_h = torch.from_numpy(h.view(np.int32)).cuda()
    d_h1 = torch.roll(d_h, 1)
    d_r0 = torch.zeros(NCP * L64, dtype=torch.int64, device="cuda")
    d_r1 = torch.zeros(NCP * L64, dtype=torch.int64, device="cuda")
    out = []

    def rec(name, ms, units, unit_name, **kw):
        line = {"kernel": name, "ms": round(ms, 4), unit_name + "_per_s": units / (ms * 1e-3), **kw}
        out.append(line)
        print(json.dumps(line), flush=True)

    ms = timeit(lambda: ctx.eval_poly_dev(SEED, 0, d_c8.data_ptr(), d_h.data_ptr(), None, D, None, d_r0.data_ptr(), st))
    rec("k_evalpoly<1> + finish", ms, D * AES_BLOCKS_PER_CT, "aes_blocks", D=D)
    ms = timeit(lambda: ctx.eval_poly2_dev(SEED, 0, d_c8.data_ptr(), d_h.data_ptr(), d_h1.data_ptr(), D, None, d_r0.data_ptr(),
                                           None, d_r1.data_ptr(), st))
    rec("k_evalpoly<2> + 2 finish", ms, D * AES_BLOCKS_PER_CT, "aes_blocks", D=D)

    # host-flavour call (what bench.py's e2e times): pageable and pinned inputs, against the device call + sync
    import time
    h64 = h.astype(np.uint64)
    rop = np.zeros((NC, L64), np.uint64)

    def host_call(c8a, ha, ra):
        ctx._ck(ctx.lib.mfb_eval_poly(ctx.h, m.api._p8(np.frombuffer(SEED, np.uint8).copy()), 0, m.api._p8(c8a), m.api._p64(ha),
                                      None, D, m.api._p64(ra)))

    def wall(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return 1e3 * (time.perf_counter() - t0) / reps

    if not ONCE:
        ms = wall(lambda: host_call(c8.reshape(-1), h64, rop))
        rec("mfb_eval_poly host call, pageable buffers", ms, D, "mac", D=D)
        p_c8, p_h, p_r = (torch.from_numpy(a.copy()).pin_memory().numpy()
                          for a in (c8.reshape(-1), h64, rop.view(np.uint64).reshape(-1)))
        ms = wall(lambda: host_call(p_c8, p_h, p_r))
        rec("mfb_eval_poly host call, pinned buffers", ms, D, "mac", D=D)
        ms = wall(lambda: (ctx.eval_poly_dev(SEED, 0, d_c8.data_ptr(), d_h.data_ptr(), None, D, None, d_r0.data_ptr(), st),
                           torch.cuda.synchronize()))
        rec("mfb_eval_poly_dev + synchronize (wall)", ms, D, "mac", D=D)
        h64b = np.roll(h64, 1)
        ms = wall(lambda: ctx.eval_poly2(SEED, 0, c8, h64, h64b))
        rec("mfb_eval_poly2 host call (python wrapper, pageable buffers)", ms, 2 * D, "mac", D=D)

    d_cts = torch.empty(D * NCP * L64, dtype=torch.int64, device="cuda")
    ms = timeit(lambda: ctx.expand_dev(SEED, 0, d_c8.data_ptr(), D, d_cts.data_ptr(), st), reps=3, warm=1)
    rec("k_expand", ms, D * AES_BLOCKS_PER_CT, "aes_blocks", D=D)
    ms = timeit(lambda: ctx.lincomb_dev(d_cts.data_ptr(), d_h.data_ptr(), D, None, d_r0.data_ptr(), st), reps=20, warm=3)
    rec("k_lincomb<1> + finish", ms, D * 129448, "bytes", D=D)
    ms = timeit(lambda: ctx.lincomb2_dev(d_cts.data_ptr(), d_h.data_ptr(), d_h1.data_ptr(), D, None, d_r0.data_ptr(), None,
                                         d_r1.data_ptr(), st), reps=20, warm=3)
    rec("k_lincomb<2> + 2 finish", ms, D * 129448, "bytes", D=D)
    del d_cts

    cnt = 2 * D + 64
    sk = rng.integers(0, 1 << 63, size=(L64, NCP), dtype=np.uint64)
    sk[:, 1470:] = 0
    d_sk = torch.from_numpy(sk.view(np.int64).reshape(-1)).cuda()
    d_msg = torch.from_numpy(rng.integers(0, m.P, size=cnt, dtype=np.uint64).view(np.int64)).cuda()
    d_ent = torch.from_numpy(rng.integers(0, 256, size=cnt * 70, dtype=np.uint8)).cuda()
    d_out = torch.empty(cnt * CT, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: ctx.encrypt_dev(SEED, 0, d_sk.data_ptr(), d_msg.data_ptr(), d_ent.data_ptr(), 70, 69, cnt, d_out.data_ptr(), st),
                reps=3, warm=1)
    rec("k_encrypt", ms, cnt * AES_BLOCKS_PER_CT, "aes_blocks", count=cnt)

    ncts = 5
    d_flat = torch.from_numpy(rng.integers(0, 1 << 63, size=ncts * NC * L64, dtype=np.uint64).view(np.int64)).cuda()
    d_m = torch.zeros(ncts, dtype=torch.int64, device="cuda")
    d_dot = torch.zeros(ncts * L64, dtype=torch.int64, device="cuda")
    ms = timeit(lambda: ctx.decrypt_dev(d_sk.data_ptr(), d_flat.data_ptr(), None, ncts, d_m.data_ptr(), d_dot.data_ptr(), st), reps=20)
    rec("k_decrypt (5 ciphertexts, the verifier's batch)", ms, ncts, "ciphertexts")
    ctx.close()


if __name__ == "__main__":
    main()
