#!/usr/bin/env python
"""BASELINE configs[4]: lincomb throughput against the HBM roofline over LWE parameter points (n, log q).

Only (1470, 736) exists in the reference (lwe.h:119-121 `#error` otherwise), so the other points are synthetic shapes:
ciphertexts of n + 1 coordinates with q_eff = 2^(64 * floor(log q / 64)), random contents, D chosen so that the
resident array is ~4 GB.  Prints one JSON line per point; run on a B200:  python tools/sweep_params.py > profiles/sweep.jsonl
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import c_lwe_snarks_b200 as m  # noqa: E402

POINTS = [(1024, 512), (1024, 640), (1246, 640), (1246, 736), (1470, 736), (1470, 800), (1470, 896), (1600, 768), (2047, 1024)]


def main():
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    ctx = m.Context(0)
    for n, logq in POINTS:
        L = logq // 64
        nc = n + 1
        T = (nc + 63) // 64
        ct_bytes = T * L * 64 * 8
        d = max(1024, int(4e9 // ct_bytes))
        cts = torch.randint(-2**62, 2**62, (d * T * L * 64,), dtype=torch.int64, device="cuda")
        h = torch.randint(0, 2**31 - 1, (d,), dtype=torch.int32, device="cuda")
        out = torch.zeros(T * L * 64, dtype=torch.int64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            ctx.lincomb_generic_dev(L, nc, cts.data_ptr(), h.data_ptr(), d, out.data_ptr(), st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            ctx.lincomb_generic_dev(L, nc, cts.data_ptr(), h.data_ptr(), d, out.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        algo = d * nc * L * 8  # live bytes: (n+1) coordinates x L limbs
        print(json.dumps({"n": n, "logq": logq, "q_eff_bits": 64 * L, "ciphertexts": d, "resident_GB": d * ct_bytes / 1e9,
                          "ms": ms, "mac_per_s": d / (ms * 1e-3), "algorithmic_GBps": algo / ms / 1e6,
                          "frac_of_measured_copy_peak": algo / ms / 1e6 / peak}), flush=True)
        del cts
        torch.cuda.empty_cache()
    ctx.close()


if __name__ == "__main__":
    main()
