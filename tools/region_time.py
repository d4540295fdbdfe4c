import sys, time
sys.path.insert(0, '.')
import numpy as np
import c_lwe_snarks_b200 as m
ctx = m.Context(0)
D = 1 << 16
c8 = np.zeros((D, 92), np.uint8)
seed = bytes(range(40))
for i in range(4):
    t0 = time.perf_counter(); reg = ctx.region(seed, 0, c8); t1 = time.perf_counter(); reg.close(); t2 = time.perf_counter()
    print("create %.1f ms  destroy %.1f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)), flush=True)
ctx.close()
