"""Helper PROCESS of tests/test_gpu_persistence.py: a fresh process that has seen nothing but files.

    python tests/reload_prove_helper.py D M crs_file ssp.npy witness.npy entropy.bin out_proof [devices]

Loads the CRS with mf_crs_read, makes it and the SSP blob resident, proves with the given entropy (the prover's
own draws: delta, then the smudging) and writes the proof with mf_proof_write."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    D, M = int(sys.argv[1]), int(sys.argv[2])
    crs_file, ssp_file, wit_file, ent_file, out_proof = sys.argv[3:8]
    devices = int(sys.argv[8]) if len(sys.argv) > 8 else 1
    from c_lwe_snarks_b200.snark import Snark
    sn = Snark(D, M)
    ent = np.fromfile(ent_file, dtype=np.uint8)
    pos = [0]

    @C.CFUNCTYPE(None, C.c_void_p, C.c_size_t, C.c_void_p)
    def draw(buf, n, _arg):
        if pos[0] + n > ent.size:
            raise SystemExit("helper: entropy exhausted")
        C.memmove(buf, ent.ctypes.data + pos[0], n)
        pos[0] += n

    sn.load_crs(crs_file)  # (crs_init draws a throw-away seed from the OS: before the hook is installed)
    sn.lib.mf_set_entropy_source(draw, None)
    sn.ssp[:] = np.load(ssp_file)
    wl = np.load(wit_file).astype(np.uint64)
    imp = getattr(sn.gmp, "__gmpz_import")
    imp.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_size_t, C.c_int, C.c_size_t, C.c_void_p]
    imp(C.byref(sn.witness), wl.size, -1, 8, 0, 0, wl.ctypes.data)
    if devices > 1:
        import torch
        sn.set_devices(devices, spread=torch.cuda.device_count() >= devices)
    sn.make_resident()
    sn.prove()
    sn.save_proof(out_proof)
    assert pos[0] == ent.size, f"the prover drew {pos[0]} of {ent.size} entropy bytes"
    sn.lib.mf_set_entropy_source(None, None)
    sn.close()
    print("reload_prove_helper ok")


if __name__ == "__main__":
    main()
