"""c_lwe_snarks_b200 — B200 (sm_100a) kernels for the hot path of the lattice SSP-SNARK of
mmaker/c-lwe-snarks ("mangiafuoco"): ciphertext linear combination, AES-CTR expansion of the
a-vectors, Regev encryption and batched decryption.

The product is the C-ABI library ``lib/libmfb200.so`` (``include/mfb200.h``) plus the C drop-in of the
reference's own headers (``host/``).  This package is the Python binding used by tests and bench.py;
it never falls back to a CPU implementation: importing :mod:`c_lwe_snarks_b200.api` without the built
library raises.
"""
from .api import (ALGO_BYTES_PER_MAC, CT_BYTES, CTR_CT, ENT_BYTES, FLAT_CT_U64, L64, N, NC, NCP, P, PLANAR_U64,
                  Context, DeviceSet, MfbError, PeerGroup, Region, ResidentSsp, SetRegion, build_library, library_path)

__all__ = ["Context", "Region", "ResidentSsp", "PeerGroup", "DeviceSet", "SetRegion", "MfbError", "build_library",
           "library_path", "N", "NC", "NCP", "L64", "P", "CT_BYTES", "CTR_CT", "ENT_BYTES", "FLAT_CT_U64", "PLANAR_U64",
           "ALGO_BYTES_PER_MAC"]
