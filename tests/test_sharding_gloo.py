"""CPU, world_size 2 (gloo): the multi-GPU sharding + exchange logic of c_lwe_snarks_b200/sharding.py.

The collectives and the bookkeeping are the production code; the three kernels are replaced by numpy/oracle
stand-ins that live HERE (test doubles, not a product fallback).  The result of the 2-rank run must equal the
oracle's single-process eval_poly over all ciphertexts, bit for bit.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

N, NC, NCP, L64, L32, CTR_CT = 1470, 1471, 1472, 11, 22, 92 * 1470
D_TOTAL = 13  # odd on purpose: ranks get 7 and 6 ciphertexts


class NumpyOps:
    """Stand-ins with the semantics of k_lincomb(+finish), k_columns_split and k_columns_carry."""

    def __init__(self, oracle, seed, stream_first, c8, h):
        self.oracle, self.seed, self.first, self.c8, self.h = oracle, seed, stream_first, c8, h

    def lincomb(self, cts, coeffs, d, out_flat):
        acc = self.oracle.eval_poly(self.seed, self.first * CTR_CT, self.c8, self.h)[:, :L64]  # (1471, 11)
        buf = np.zeros((NCP, L64), np.uint64)
        buf[:NC] = acc
        out_flat.numpy().view(np.uint64)[:] = buf.reshape(-1)

    def columns_split(self, flat, cols):
        v = flat.numpy().view(np.uint64).reshape(NCP, L64)
        c = np.zeros((NCP, L32), np.uint64)
        c[:, 0::2] = v & np.uint64(0xFFFFFFFF)
        c[:, 1::2] = v >> np.uint64(32)
        cols.numpy().view(np.uint64)[:] = c.reshape(-1)

    def columns_carry(self, cols_own, first_coord, ncoord, out_own_flat):
        c = cols_own.numpy().view(np.uint64).reshape(ncoord, L32)
        out = np.zeros((ncoord, L64), np.uint64)
        for t in range(ncoord):
            if first_coord + t >= NC:
                continue
            val = sum(int(c[t, l]) << (32 * l) for l in range(L32)) % (1 << 704)
            out[t] = np.frombuffer(val.to_bytes(88, "little"), "<u8")
        out_own_flat.numpy().view(np.uint64)[:] = out.reshape(-1)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from conftest import SEED, xof_records, xof_scalars
    from c_lwe_snarks_b200.sharding import ShardedLincomb, ShardPlan
    from oracle.loader import Oracle

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = ShardPlan(world, rank)
        first, count = plan.ct_range(D_TOTAL)
        c8, h = xof_records("shard-c8", D_TOTAL), xof_scalars("shard-h", D_TOTAL)
        ops = NumpyOps(Oracle(), SEED, first, c8[first:first + count], h[first:first + count])
        sl = ShardedLincomb(plan, ops, dist, lambda n: torch.zeros(n, dtype=torch.int64))
        res = sl.step(None, None, count).numpy().view(np.uint64).reshape(NCP, L64).copy()
        q.put((rank, first, count, res))
    finally:
        dist.destroy_process_group()


def _setup_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from conftest import SEED, xof, xof_scalars
    from c_lwe_snarks_b200.sharding import ShardedSetup, ShardPlan
    from oracle.loader import Oracle

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total = 7
        orc = Oracle()
        sk = orc.key_gen(xof("shard-sk", N * 92))
        msg, ent = xof_scalars("shard-msg", total), xof("shard-ent", total * 70)

        def encrypt(first, count):  # stand-in for DeviceOps.encrypt: the oracle on this rank's slice
            recs = orc.encrypt(SEED, first * CTR_CT, sk, msg[first:first + count], ent[first * 70:(first + count) * 70])
            return torch.from_numpy(recs.reshape(-1).copy())

        out = ShardedSetup(ShardPlan(world, rank), dist, lambda n: torch.zeros(n, dtype=torch.uint8)).run(total, encrypt)
        q.put((rank, out.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_setup_matches_single_process_oracle(oracle):
    import torch.multiprocessing as mp

    from conftest import SEED, xof, xof_scalars
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_setup_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    sk = oracle.key_gen(xof("shard-sk", N * 92))
    want = oracle.encrypt(SEED, 0, sk, xof_scalars("shard-msg", 7), xof("shard-ent", 7 * 70)).reshape(-1)
    for _, got in out:
        assert np.array_equal(got, want)


def test_shard_plan():
    from c_lwe_snarks_b200.sharding import ShardPlan
    for world in (1, 2, 4, 8):
        covered = []
        for r in range(world):
            f, c = ShardPlan(world, r).ct_range(1000003)
            covered.append((f, c))
            assert ShardPlan(world, r).coords_per_rank * world == NCP
        assert covered[0][0] == 0 and sum(c for _, c in covered) == 1000003
        assert all(covered[i][0] + covered[i][1] == covered[i + 1][0] for i in range(world - 1))
    # world sizes that do not divide 1472 shard the ciphertexts fine (the peer-memory exchange takes any world <= 16);
    # only the NCCL exchange, which scatters the coordinates evenly, refuses them
    assert [ShardPlan(3, r).ct_range(10) for r in range(3)] == [(0, 4), (4, 3), (7, 3)]
    with pytest.raises(ValueError):
        ShardPlan(3, 0).coords_per_rank
    with pytest.raises(ValueError):
        ShardPlan(2, 2)
    assert ShardPlan(8, 7).ct_range(5) == (5, 0)  # more ranks than ciphertexts: empty shard


@pytest.mark.timeout(300)
def test_two_rank_exchange_matches_single_process_oracle(oracle):
    import torch.multiprocessing as mp

    from conftest import SEED, xof_records, xof_scalars
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    c8, h = xof_records("shard-c8", D_TOTAL), xof_scalars("shard-h", D_TOTAL)
    want = oracle.eval_poly(SEED, 0, c8, h)[:, :L64]
    ranges = sorted((first, count) for _, first, count, _ in out)
    assert ranges == [(0, 7), (7, 6)]
    for rank, _, _, res in out:
        assert not res[NC:].any()
        assert np.array_equal(res[:NC], want), f"rank {rank}"


# ------------------------------------------------------------------------------------------- peer exchange, host side
class FakePeerGroup:
    """Test double of api.PeerGroup: a shared-memory "symmetric buffer" is replaced by a file per rank in a directory
    both ranks see; connect() records the handles, lincomb_dev publishes this rank's partial and adds everybody's."""

    def __init__(self, rank, world, tmpdir, partial):
        self.rank, self.world, self.dir, self.partial = rank, world, tmpdir, partial
        self.ipc_handle = np.full(64, rank + 1, np.uint8)
        self.log = []

    def connect(self, handles):
        self.log.append(("connect", bytes(handles)))

    def lincomb_dev(self, cts_ptr, coeffs_ptr, d, rop_in_ptr, rop_out_ptr, stream):
        import ctypes
        import time
        np.save(os.path.join(self.dir, f"p{self.rank}.npy"), self.partial)
        os.replace(os.path.join(self.dir, f"p{self.rank}.npy"), os.path.join(self.dir, f"done{self.rank}.npy"))
        acc = [0] * NC
        for r in range(self.world):
            path = os.path.join(self.dir, f"done{r}.npy")
            for _ in range(2000):
                if os.path.exists(path):
                    break
                time.sleep(0.01)
            part = np.load(path)
            for c in range(NC):
                acc[c] = (acc[c] + int.from_bytes(part[c].tobytes(), "little")) % (1 << 704)
        out = np.zeros((NCP, L64), np.uint64)
        for c in range(NC):
            out[c] = np.frombuffer(acc[c].to_bytes(88, "little"), "<u8")
        ctypes.memmove(rop_out_ptr, out.ctypes.data, out.nbytes)
        self.log.append(("lincomb", d))

    def check(self):
        self.log.append(("check",))

    def disconnect(self):
        self.log.append(("disconnect",))

    def close(self):
        self.log.append(("close",))


def _peer_worker(rank, world, port, tmpdir, q):
    import torch
    import torch.distributed as dist

    from conftest import SEED, xof_records, xof_scalars
    from c_lwe_snarks_b200.sharding import PeerShardedLincomb, ShardPlan
    from oracle.loader import Oracle

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = ShardPlan(world, rank)
        first, count = plan.ct_range(D_TOTAL)
        c8, h = xof_records("shard-c8", D_TOTAL), xof_scalars("shard-h", D_TOTAL)
        partial = Oracle().eval_poly(SEED, first * CTR_CT, c8[first:first + count], h[first:first + count])[:, :L64].copy()
        group = FakePeerGroup(rank, world, tmpdir, partial)
        peer = PeerShardedLincomb(plan, group, dist, lambda n: torch.zeros(n, dtype=torch.int64),
                                  lambda n: torch.zeros(n, dtype=torch.uint8))
        dummy = torch.zeros(1, dtype=torch.int64)
        res = peer.step(dummy, dummy, count).numpy().view(np.uint64).reshape(NCP, L64).copy()
        peer.check()
        peer.close()
        q.put((rank, res, group.log))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_peer_exchange_host_logic(oracle, tmp_path):
    """sharding.PeerShardedLincomb under gloo with a stand-in group: the 64-byte handles are gathered in rank order and
    handed to connect(); step() returns the all-rank sum; close() runs barrier / disconnect / barrier / close."""
    import torch.multiprocessing as mp

    from conftest import SEED, xof_records, xof_scalars
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29300 + os.getpid() % 150
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    c8, h = xof_records("shard-c8", D_TOTAL), xof_scalars("shard-h", D_TOTAL)
    want = oracle.eval_poly(SEED, 0, c8, h)[:, :L64]
    for rank, res, log in out:
        assert np.array_equal(res[:NC], want), f"rank {rank}"
        assert log[0] == ("connect", bytes([1] * 64 + [2] * 64))
        assert [e[0] for e in log] == ["connect", "lincomb", "check", "disconnect", "close"]
