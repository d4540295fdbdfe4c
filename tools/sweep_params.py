#!/usr/bin/env python
"""BASELINE configs[4]: lincomb AND encrypt throughput over LWE parameter points (n, log q).

Only (1470, 736) exists in the reference (lwe.h:119-121 `#error` otherwise), so the other points are synthetic shapes:
ciphertexts of n + 1 coordinates with q_eff = 2^(64 * floor(log q / 64)).
  lincomb  resident ciphertexts (random contents), D chosen so that the array is ~16 GB; bound: HBM — algorithmic bytes
           (n + 1) * 8 L per ciphertext-MAC against the measured copy peak.
  encrypt  a regenerated from AES-256-CTR in-kernel (log q / 8 stream bytes per coordinate); bound: the AES generator —
           AES blocks/s against the shared-memory-lookup bound of DESIGN.md §3 K2 (206 wavefronts per block: 4.5e10 blocks/s),
           plus the GB/s of a-vector bytes the kernel replaced.
Prints one JSON line per point and kernel; run on a B200:  python tools/sweep_params.py > profiles/sweep_params_r02.jsonl
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import c_lwe_snarks_b200 as m  # noqa: E402

POINTS = [(1024, 512), (1024, 640), (1246, 640), (1246, 736), (1470, 736), (1470, 800), (1470, 896), (1600, 768), (2047, 1024)]
if os.environ.get("SWEEP_POINTS"):  # e.g. SWEEP_POINTS=2047:1024,1470:896 SWEEP_KERNELS=encrypt (profiling one point under ncu)
    POINTS = [tuple(int(v) for v in pt.split(":")) for pt in os.environ["SWEEP_POINTS"].split(",")]
KERNELS = os.environ.get("SWEEP_KERNELS", "lincomb,encrypt").split(",")
AES_BOUND = 4.5e10  # blocks/s at 100 % of the LDS pipe, 1.965 GHz
SEED = bytes(range(40))


def timed(fn, reps):
    for _ in range(2):
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
    gb = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    ctx = m.Context(0)
    st = torch.cuda.current_stream().cuda_stream
    for n, logq in POINTS:
        L = logq // 64
        nc = n + 1
        T = (nc + 63) // 64
        ct_bytes = T * L * 64 * 8
        d = max(1024, int(gb * 1e9 // ct_bytes))
        if "lincomb" in KERNELS:
            cts = torch.randint(-2**62, 2**62, (d * T * L * 64,), dtype=torch.int64, device="cuda")
            h = torch.randint(0, 2**31 - 1, (d,), dtype=torch.int32, device="cuda")
            out = torch.zeros(T * L * 64, dtype=torch.int64, device="cuda")
            ms = timed(lambda: ctx.lincomb_generic_dev(L, nc, cts.data_ptr(), h.data_ptr(), d, out.data_ptr(), st), 10)
            algo = d * nc * L * 8  # live bytes: (n+1) coordinates x L limbs
            print(json.dumps({"kernel": "lincomb", "n": n, "logq": logq, "q_eff_bits": 64 * L, "ciphertexts": d,
                              "resident_GB": d * ct_bytes / 1e9, "ms": ms, "mac_per_s": d / (ms * 1e-3), "algorithmic_GBps": algo / ms / 1e6,
                              "frac_of_measured_copy_peak": algo / ms / 1e6 / peak}), flush=True)
            del cts
            torch.cuda.empty_cache()
        if "encrypt" not in KERNELS:
            continue
        # encrypt: ~256 ciphertexts per SM
        ctb = logq // 8
        cnt = 148 * 256
        stride = (n + 63) // 64 * 64
        sk = torch.randint(-2**62, 2**62, (L * stride,), dtype=torch.int64, device="cuda")
        msg = torch.randint(0, 2**31 - 1, (cnt,), dtype=torch.int64, device="cuda")
        nb = min(8 * L, 69)
        ent = torch.randint(0, 256, (cnt * (nb + 1),), dtype=torch.uint8, device="cuda")
        rec = torch.zeros(cnt * ctb, dtype=torch.uint8, device="cuda")
        ms = timed(lambda: ctx.encrypt_generic_dev(L, n, ctb, SEED, 0, sk.data_ptr(), stride, msg.data_ptr(), ent.data_ptr(), nb + 1, nb,
                                                   cnt, rec.data_ptr(), st), 3)
        blocks = cnt * n * ctb / 16
        print(json.dumps({"kernel": "encrypt", "n": n, "logq": logq, "q_eff_bits": 64 * L, "ciphertexts": cnt, "ms": ms,
                          "encryptions_per_s": cnt / (ms * 1e-3), "aes_blocks_per_s": blocks / (ms * 1e-3),
                          "frac_of_lds_bound": blocks / (ms * 1e-3) / AES_BOUND, "a_bytes_replaced_GBps": cnt * n * ctb / ms / 1e6,
                          "limb_products_per_coordinate": L * (2 * L + 1)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
