"""Full setup / prover / verifier through the drop-in C layer with the CRS regions sharded over the GPUs of one box
(mf_set_devices): BASELINE configs[3] — a 2^20-constraint SSP proved on 2/4/8 B200s — as ONE single-threaded program,
the shape of the reference's own test_snark / benchmark_snark.
Usage: python tools/snark_box.py [log2d] [M] [n_devices ...]   -> one JSON line per device count."""
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from c_lwe_snarks_b200.snark import Snark  # noqa: E402

log2d = int(sys.argv[1]) if len(sys.argv) > 1 else 20
M = int(sys.argv[2]) if len(sys.argv) > 2 else 64
counts = [int(a) for a in sys.argv[3:]] or [8]

sn = Snark(1 << log2d, M)
t0 = time.perf_counter()
sn.random_ssp()
t_ssp = time.perf_counter() - t0
t_setup_cold = sn.setup()   # includes CUDA initialisation and the first allocations
t_setup = sn.setup()
for n in counts:
    sn.set_devices(n)
    sn.setup()  # warm-up of the members' buffers
    t_setup_set = sn.setup()  # the encryptions spread over the n GPUs, one host thread per GPU (mfb_set_encrypt_par)
    t_nonres = float("nan")
    if not os.environ.get("SNARK_BOX_RESIDENT_ONLY"):  # (set when only the resident pipeline is to be profiled)
        sn.prove()  # nothing resident: both fused AES + MAC passes sharded over the n GPUs (mfb_set_eval_poly2)
        t_nonres = min(sn.prove() for _ in range(2))
    t0 = time.perf_counter()
    sn.make_resident()
    t_res = time.perf_counter() - t0
    proves = [sn.prove() for _ in range(4)]
    ok, t_ver = sn.verify()
    sn.tamper()
    bad, _ = sn.verify()
    print(json.dumps({"D": 1 << log2d, "M": M, "devices": n, "random_ssp_s": round(t_ssp, 3), "setup_cold_s": round(t_setup_cold, 3), "setup_s": round(t_setup, 4),
                      "setup_over_the_set_s": round(t_setup_set, 4), "prove_nothing_resident_ms": round(1e3 * t_nonres, 2), "make_resident_s": round(t_res, 3),
                      "prove_ms": [round(1e3 * t, 3) for t in proves],
                      "verify_ms": round(1e3 * t_ver, 3), "accept": ok, "tampered_accept": bad,
                      "api": "setup/prover/verifier (snark.h:44-51) via libmangiafuoco_b200.so, one host thread"}), flush=True)
    sn.lib.mf_crs_release(__import__("ctypes").byref(sn.crs))
    sn.lib.mf_ssp_release(sn._ssp_ptr())
    sn._ssp_resident = False
sn.close()
