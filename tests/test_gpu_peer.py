"""The peer-memory exchange (mfb_peer_*, k_lincomb_finish_peer): the sharded lincomb whose finish kernel pushes its
partial sum into every rank's symmetric buffer over NVLink, waits for the others' and adds them.

One GPU is enough to exercise the whole protocol: `world` ranks = `world` contexts in this process, each with its
own stream, connected with mfb_peer_connect_local; their finish kernels run concurrently and really wait for each
other's flags.  With two or more GPUs the same test runs with one rank per device (true peer access), and the
multi-process test drives the CUDA-IPC path under torch.distributed/NCCL beside the NCCL exchange.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import SEED, xof_records, xof_scalars

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
N, NC, NCP, L64, CT_BYTES, CTR_CT = 1470, 1471, 1472, 11, 92, 92 * 1470


def _flat(t):
    return t.cpu().numpy().view(np.uint64)[: NC * L64].reshape(NC, L64)


def _run_local_world(world, d_total, calls, devices, fused=False, split=False):
    import torch

    import c_lwe_snarks_b200 as m

    def ct_range(r):  # ShardPlan.ct_range without its "world divides 1472" rule (that one is the NCCL exchange's)
        base, extra = divmod(d_total, world)
        return r * base + min(r, extra), base + (1 if r < extra else 0)

    c8, h = xof_records(f"peer-c8-{d_total}", d_total), xof_scalars(f"peer-h-{d_total}", d_total)
    ranks = []
    for r in range(world):
        dev = devices[r % len(devices)]
        torch.cuda.set_device(dev)
        ctx = m.Context(dev)
        first, cnt = ct_range(r)
        st = torch.cuda.Stream(device=dev)
        with torch.cuda.device(dev):
            d_c8 = torch.from_numpy(c8[first:first + cnt].reshape(-1).copy()).cuda()
            d_h = torch.from_numpy(h[first:first + cnt].astype(np.uint32).view(np.int32)).cuda()
            d_cts = torch.zeros(max(cnt, 1) * NCP * L64, dtype=torch.int64, device="cuda")
            ctx.expand_dev(SEED, first * CTR_CT, d_c8.data_ptr(), cnt, d_cts.data_ptr(), 0)
            out = [torch.zeros(NCP * L64, dtype=torch.int64, device="cuda") for _ in range(2)]
            part = [torch.zeros(NCP * L64, dtype=torch.int64, device="cuda") for _ in range(2)]
            torch.cuda.synchronize()
        g = ctx.peer_group(world, r)
        g.set_timeout(5.0)
        st2 = torch.cuda.Stream(device=dev)
        ranks.append(dict(ctx=ctx, g=g, st=st, st2=st2, part=part, first=first, cnt=cnt, c8=d_c8, h=d_h, cts=d_cts, out=out, dev=dev))
    bases = [rk["g"].base for rk in ranks]
    if world > 1:
        for rk in ranks:
            torch.cuda.set_device(rk["dev"])
            rk["g"].connect_local(bases)
    # reference: one context over all the ciphertexts
    torch.cuda.set_device(devices[0])
    want1 = ranks[0]["ctx"].eval_poly(SEED, 0, c8, h)
    results = []
    try:
        for i in range(calls):
            # call i accumulates into the result of call i-1: rop_in is exercised and every call has a new answer
            for rk in ranks:
                torch.cuda.set_device(rk["dev"])
                prev = rk["out"][(i + 1) % 2] if i else None
                if split:  # plain lincomb on the main stream, the all-reduce kernel on the side stream
                    part = rk["part"][i % 2]
                    rk["ctx"].lincomb_dev(rk["cts"].data_ptr(), rk["h"].data_ptr(), rk["cnt"], None, part.data_ptr(),
                                          rk["st"].cuda_stream)
                    ev = torch.cuda.Event()
                    ev.record(rk["st"])
                    rk["st2"].wait_event(ev)
                    rk["g"].allreduce_dev(part.data_ptr(), None if prev is None else prev.data_ptr(),
                                          rk["out"][i % 2].data_ptr(), rk["st2"].cuda_stream)
                    ev2 = torch.cuda.Event()
                    ev2.record(rk["st2"])
                    rk["st"].wait_event(ev2)  # part[i % 2] is reused two calls later; keep it simple: wait
                elif fused:
                    rk["g"].eval_poly_dev(SEED, rk["first"] * CTR_CT, rk["c8"].data_ptr(), rk["h"].data_ptr(), None, rk["cnt"],
                                          None if prev is None else prev.data_ptr(), rk["out"][i % 2].data_ptr(),
                                          rk["st"].cuda_stream)
                else:
                    rk["g"].lincomb_dev(rk["cts"].data_ptr(), rk["h"].data_ptr(), rk["cnt"],
                                        None if prev is None else prev.data_ptr(), rk["out"][i % 2].data_ptr(),
                                        rk["st"].cuda_stream)
        for rk in ranks:
            torch.cuda.set_device(rk["dev"])
            rk["st"].synchronize()
            rk["st2"].synchronize()
            rk["g"].check()
            results.append(_flat(rk["out"][(calls - 1) % 2]))
    finally:
        for rk in ranks:
            torch.cuda.set_device(rk["dev"])
            rk["g"].disconnect()
        for rk in ranks:
            torch.cuda.set_device(rk["dev"])
            rk["g"].close()
            rk["ctx"].close()
        torch.cuda.set_device(devices[0])
    # expected: calls x the single-context sum, mod 2^704
    mask = (1 << 704) - 1
    want = np.zeros((NC, L64), np.uint64)
    for c in range(NC):
        v = int.from_bytes(want1[c].tobytes(), "little") * calls & mask
        want[c] = np.frombuffer(v.to_bytes(88, "little"), "<u8")
    return results, want


@pytest.mark.parametrize("world,d_total,calls", [(1, 37, 3), (2, 13, 1), (2, 300, 6), (3, 50, 5), (8, 1000, 7)])
def test_peer_exchange_ranks_in_one_process(world, d_total, calls):
    results, want = _run_local_world(world, d_total, calls, devices=[0])
    for r, got in enumerate(results):
        assert np.array_equal(got, want), f"rank {r}"


def test_peer_exchange_fused_eval_poly():
    results, want = _run_local_world(4, 64, 3, devices=[0], fused=True)
    for got in results:
        assert np.array_equal(got, want)


@pytest.mark.parametrize("world,d_total,calls", [(1, 20, 3), (2, 300, 9), (4, 100, 6)])
def test_peer_allreduce_on_a_side_stream(world, d_total, calls):
    # (at most 4 ranks x 2 streams here: a process gets 8 hardware queues, more streams would alias and serialise)
    results, want = _run_local_world(world, d_total, calls, devices=[0], split=True)
    for r, got in enumerate(results):
        assert np.array_equal(got, want), f"rank {r}"


def test_peer_exchange_empty_rank():
    # fewer ciphertexts than ranks: some ranks contribute a zero partial (d = 0) and still take part
    results, want = _run_local_world(4, 2, 2, devices=[0])
    for got in results:
        assert np.array_equal(got, want)


def test_peer_timeout_is_reported_not_hung():
    """a rank whose peer never calls gives up after the timeout and mfb_peer_status says who was missing"""
    import torch

    import c_lwe_snarks_b200 as m
    ctx0, ctx1 = m.Context(0), m.Context(0)
    g0, g1 = ctx0.peer_group(2, 0), ctx1.peer_group(2, 1)
    try:
        g0.set_timeout(0.2)
        bases = [g0.base, g1.base]
        g0.connect_local(bases)
        g1.connect_local(bases)
        out = torch.zeros(NCP * L64, dtype=torch.int64, device="cuda")
        g0.lincomb_dev(0, 0, 0, None, out.data_ptr(), 0)  # rank 1 never calls
        torch.cuda.synchronize()
        with pytest.raises(m.api.MfbError, match="rank 1 never delivered"):
            g0.check()
    finally:
        g0.disconnect()
        g1.disconnect()
        g0.close()
        g1.close()
        ctx0.close()
        ctx1.close()


def test_peer_exchange_one_rank_per_gpu_same_process():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    world = min(n, 8)
    results, want = _run_local_world(world, 500, 5, devices=list(range(world)))
    for r, got in enumerate(results):
        assert np.array_equal(got, want), f"rank {r}"


def _ipc_worker(rank, world, port, d_per_rank, calls, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import c_lwe_snarks_b200 as m
    from c_lwe_snarks_b200.sharding import DeviceOps, PeerShardedLincomb, ShardedLincomb, ShardPlan
    from conftest import SEED as S, xof_records as xr, xof_scalars as xs

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = m.Context(rank)
    plan = ShardPlan(world, rank)
    d_total = world * d_per_rank
    c8, h = xr(f"ipc-c8-{d_total}", d_total), xs(f"ipc-h-{d_total}", d_total)
    first, cnt = plan.ct_range(d_total)
    d_c8 = torch.from_numpy(c8[first:first + cnt].reshape(-1).copy()).cuda()
    d_h = torch.from_numpy(h[first:first + cnt].astype(np.uint32).view(np.int32)).cuda()
    d_cts = torch.zeros(cnt * NCP * L64, dtype=torch.int64, device="cuda")
    ctx.expand_dev(S, first * CTR_CT, d_c8.data_ptr(), cnt, d_cts.data_ptr(), 0)
    new_i64 = lambda n: torch.zeros(n, dtype=torch.int64, device="cuda")  # noqa: E731
    new_u8 = lambda n: torch.zeros(n, dtype=torch.uint8, device="cuda")  # noqa: E731
    group = ctx.peer_group(world, rank)
    group.set_timeout(10.0)
    peer = PeerShardedLincomb(plan, group, dist, new_i64, new_u8)
    nccl = ShardedLincomb(plan, DeviceOps(ctx, torch), dist, new_i64)
    ok = True
    for i in range(calls):
        a = peer.step(d_cts, d_h, cnt)
        b = nccl.step(d_cts, d_h, cnt)
        torch.cuda.synchronize()
        peer.check()
        ok = ok and bool(torch.equal(a[: NC * L64], b[: NC * L64]))
    f = peer.step_fused(S, first * CTR_CT, d_c8, d_h, cnt)
    torch.cuda.synchronize()
    ok = ok and bool(torch.equal(f[: NC * L64], b[: NC * L64]))
    # the pipelined schedule on the same group: lincomb on the main stream, all-reduce kernel on a side stream
    from c_lwe_snarks_b200.sharding import PipelinedPeerShardedLincomb
    pipe = PipelinedPeerShardedLincomb.__new__(PipelinedPeerShardedLincomb)
    pipe.torch, pipe.ctx, pipe.inner, pipe.group, pipe.results = torch, ctx, peer, group, peer.results
    pipe.partials = [new_i64(NCP * L64), new_i64(NCP * L64)]
    pipe.side, pipe.done, pipe.calls = torch.cuda.Stream(), [None, None], 0
    for i in range(5):
        p_res = pipe.submit(d_cts, d_h, cnt)
    pipe.drain()
    torch.cuda.synchronize()
    peer.check()
    ok = ok and bool(torch.equal(p_res[: NC * L64], b[: NC * L64]))
    res = a[: NC * L64].cpu().numpy().view(np.uint64).reshape(NC, L64).copy()
    peer.close()
    ctx.close()
    dist.destroy_process_group()
    q.put((rank, ok, res))


def test_peer_exchange_cuda_ipc_beside_nccl():
    """one process per GPU (torch.distributed, NCCL for the handle gather): the IPC peer exchange equals the NCCL
    reduce-scatter / all-gather exchange and a single-GPU eval_poly over all ciphertexts."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    import c_lwe_snarks_b200 as m
    world, d_per_rank, calls = min(n, 8), 200, 4
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctxm.Process(target=_ipc_worker, args=(r, world, port, d_per_rank, calls, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    d_total = world * d_per_rank
    c8, h = xof_records(f"ipc-c8-{d_total}", d_total), xof_scalars(f"ipc-h-{d_total}", d_total)
    ctx = m.Context(0)
    want = ctx.eval_poly(SEED, 0, c8, h)
    ctx.close()
    for rank, ok, res in got:
        assert ok, f"rank {rank}: peer exchange != NCCL exchange"
        assert np.array_equal(res, want), f"rank {rank}"


# ------------------------------------------------------------------------------------------- device sets
@pytest.mark.parametrize("nmembers,d", [(1, 50), (2, 101), (3, 64), (4, 1000), (4, 3)])
def test_device_set_region_lincomb_on_one_gpu(nmembers, d, oracle):
    """mfb_set_*: a region sharded over the members of a device set (here all on GPU 0) gives the single-context result,
    for one and two scalar vectors, accumulating into rop, call after call."""
    import c_lwe_snarks_b200 as m
    off = 5 * CTR_CT + 8
    c8 = xof_records(f"set-c8-{d}", d)
    h0, h1 = xof_scalars(f"set-h0-{d}", d), xof_scalars(f"set-h1-{d}", d)
    rop = xof_records("set-rop", NC)[:, :88].copy().view("<u8").reshape(NC, L64)
    ctx = m.Context(0)
    dset = ctx.device_set([0] * (nmembers - 1))
    try:
        assert dset.size == nmembers
        reg = dset.region(SEED, off, c8)
        try:
            want0 = ctx.eval_poly(SEED, off, c8, h0, rop=rop)
            want1 = ctx.eval_poly(SEED, off, c8, h1)
            for _ in range(3):  # epochs advance: both slot parities are reused
                r0, r1 = reg.lincomb2(h0, h1, rop0=rop)
                assert np.array_equal(r0, want0) and np.array_equal(r1, want1)
                assert np.array_equal(reg.lincomb(h1), want1)
        finally:
            reg.close()
    finally:
        dset.close()
        ctx.close()
    small = min(d, 40)
    w = oracle.eval_poly(SEED, off, c8[:small], h1[:small])
    ctx = m.Context(0)
    try:
        assert np.array_equal(ctx.eval_poly(SEED, off, c8[:small], h1[:small]), w[:, :11])
    finally:
        ctx.close()


def test_device_set_over_all_gpus():
    import torch

    import c_lwe_snarks_b200 as m
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    d, off = 4000, 2 * CTR_CT
    c8 = xof_records("setN-c8", d)
    h0, h1 = xof_scalars("setN-h0", d), xof_scalars("setN-h1", d)
    ctx = m.Context(0)
    dset = ctx.device_set(list(range(1, min(n, 8))))
    try:
        reg = dset.region(SEED, off, c8)
        try:
            for _ in range(3):
                r0, r1 = reg.lincomb2(h0, h1)
                assert np.array_equal(r0, ctx.eval_poly(SEED, off, c8, h0))
                assert np.array_equal(r1, ctx.eval_poly(SEED, off, c8, h1))
        finally:
            reg.close()
    finally:
        dset.close()
        ctx.close()


@pytest.mark.parametrize("nmembers,d", [(2, 9), (3, 64), (4, 301)])
def test_device_set_eval_poly_without_residency(nmembers, d):
    """mfb_set_eval_poly2: nothing resident, every member regenerates the a-vectors of its range from AES in-kernel."""
    import c_lwe_snarks_b200 as m
    off = 11 * CTR_CT + 3
    c8 = xof_records(f"setf-c8-{d}", d)
    h0, h1 = xof_scalars(f"setf-h0-{d}", d), xof_scalars(f"setf-h1-{d}", d)
    rop = xof_records("setf-rop", NC)[:, :88].copy().view("<u8").reshape(NC, L64)
    ctx = m.Context(0)
    dset = ctx.device_set([0] * (nmembers - 1))
    try:
        want0 = ctx.eval_poly(SEED, off, c8, h0)
        want1 = ctx.eval_poly(SEED, off, c8, h1, rop=rop)
        for _ in range(2):
            r0, r1 = dset.eval_poly2(SEED, off, c8, h0, h1, rop1=rop)
            assert np.array_equal(r0, want0) and np.array_equal(r1, want1)
            assert np.array_equal(dset.eval_poly2(SEED, off, c8, h1, rop0=rop), want1)
    finally:
        dset.close()
        ctx.close()


@pytest.mark.parametrize("nmembers,cnt", [(1, 7), (2, 148 * 16 + 5), (3, 148 * 16 + 148 * 110 * 2 + 33)])
def test_device_set_encrypt_cb(nmembers, cnt, oracle):
    """mfb_set_encrypt_cb: setup's encryptions with the pieces spread over the members == mfb_encrypt on the same entropy,
    drawn exactly once and in order."""
    import c_lwe_snarks_b200 as m
    from conftest import xof
    sk = oracle.key_gen(xof("sk-setenc", N * CT_BYTES))
    msg = xof_scalars(f"m-setenc-{cnt}", cnt)
    ent = xof(f"ent-setenc-{cnt}", cnt * 70)
    pos = [0]

    def draw(n):
        out = ent[pos[0]: pos[0] + n].tobytes()
        pos[0] += n
        return out

    ctx = m.Context(0)
    dset = ctx.device_set([0] * (nmembers - 1))
    try:
        got = dset.encrypt_cb(SEED, 5 * CTR_CT + 8, sk[:, :11], msg, draw)
        assert pos[0] == cnt * 70
        assert np.array_equal(got, ctx.encrypt(SEED, 5 * CTR_CT + 8, sk[:, :11], msg, ent))
    finally:
        dset.close()
        ctx.close()


@pytest.mark.parametrize("nmembers,cnt", [(1, 7), (2, 5), (3, 148 * 16 + 148 * 110 * 3 + 33), (4, 3)])
def test_device_set_encrypt_par(nmembers, cnt, oracle):
    """mfb_set_encrypt_par (what setup() uses with OS entropy): one host thread per member over contiguous ranges, the
    records written straight into the caller's segments.  With an entropy source that hands every caller the same
    bytes the result must equal mfb_encrypt on that entropy, whatever thread drew what."""
    import c_lwe_snarks_b200 as m
    from conftest import xof
    sk = oracle.key_gen(xof("sk-setenc", N * CT_BYTES))
    msg = xof_scalars(f"m-setpar-{cnt}", cnt)
    unit = xof("ent-setpar-unit", 70)
    ent = np.tile(unit, cnt)
    calls = []

    def draw(n):  # called concurrently from the member threads (ctypes takes the GIL for each call)
        assert n % 70 == 0
        calls.append(n)
        return np.tile(unit, n // 70).tobytes()

    ctx = m.Context(0)
    dset = ctx.device_set([0] * (nmembers - 1))
    try:
        want = ctx.encrypt(SEED, 5 * CTR_CT + 8, sk[:, :11], msg, ent)
        got = dset.encrypt_par(SEED, 5 * CTR_CT + 8, sk[:, :11], msg, draw)
        assert sum(calls) == cnt * 70
        assert np.array_equal(got, want)
        # segments: a ragged partition of the record index space, as setup() passes (s | as | t | v)
        cuts = sorted({0, cnt // 3, cnt // 3 + 1, (2 * cnt) // 3, cnt})
        segs = [(a, b - a) for a, b in zip(cuts[:-1], cuts[1:])]
        outs = dset.encrypt_par(SEED, 5 * CTR_CT + 8, sk[:, :11], msg, draw, segments=segs)
        assert np.array_equal(np.concatenate(outs), want)
    finally:
        dset.close()
        ctx.close()
