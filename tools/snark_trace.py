import sys; sys.path.insert(0, '.')
from c_lwe_snarks_b200.snark import Snark
import time
sn = Snark(1 << 16, 64)
sn.random_ssp()
print("setup", sn.setup(), file=sys.stderr)
print("prove", sn.prove(), file=sys.stderr)
print("prove", sn.prove(), file=sys.stderr)
sn.make_resident()
print("prove_res", sn.prove(), file=sys.stderr)
print("prove_res", sn.prove(), file=sys.stderr)
print("verify", sn.verify(), file=sys.stderr)
sn.close()
