"""Phase timings of setup / prover / verifier through the drop-in C layer (MF_B200_TRACE=1 prints the phases).
Usage: MF_B200_TRACE=1 python tools/snark_trace.py [log2d] [M]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from c_lwe_snarks_b200.snark import Snark  # noqa: E402

log2d = int(sys.argv[1]) if len(sys.argv) > 1 else 16
M = int(sys.argv[2]) if len(sys.argv) > 2 else 64
sn = Snark(1 << log2d, M)
sn.random_ssp()
print("setup (cold)", sn.setup(), file=sys.stderr)
print("setup", sn.setup(), file=sys.stderr)
print("setup", sn.setup(), file=sys.stderr)
print("prove (cold)", sn.prove(), file=sys.stderr)
print("prove", sn.prove(), file=sys.stderr)
import ctypes
import time
for label in ("make_resident (cold)", "make_resident"):
    t0 = time.perf_counter()
    sn.lib.mf_crs_make_resident(ctypes.byref(sn.crs))
    t1 = time.perf_counter()
    sn.lib.mf_ssp_make_resident(sn._ssp_ptr())
    t2 = time.perf_counter()
    sn._ssp_resident = True
    print(label, "crs %.4f ssp %.4f" % (t1 - t0, t2 - t1), file=sys.stderr)
print("prove_res (cold)", sn.prove(), file=sys.stderr)
print("prove_res", sn.prove(), file=sys.stderr)
print("verify", sn.verify(), file=sys.stderr)
print("verify", sn.verify(), file=sys.stderr)
sn.close()
