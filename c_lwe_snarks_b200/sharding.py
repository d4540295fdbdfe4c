"""Multi-GPU sharding of the prover lincomb by ciphertext index, one process per GPU.

Each rank holds a contiguous range of the ciphertexts and produces a canonical partial sum mod 2^704 with K1;
exact sums mod 2^704 are associative, so adding the partials in any order is bit-identical to the reference's
sequential fold.  The ONE exchange per proof element:

    partial (1472 x 11 u64)  --columns_split-->  1472 x 22 u64 columns, each < 2^32        (local kernel)
    reduce-scatter(SUM) by coordinate: rank r receives the summed columns of coordinates [r*1472/W, (r+1)*1472/W)
    columns_carry: carry-propagate + truncate to 704 bit for the owned coordinates        (local kernel)
    all-gather of the owned slices -> the full result on every rank

An elementwise integer sum of W <= 2^31 values below 2^32 cannot overflow int64, which is what makes a plain
NCCL SUM exact.  The messages are 259 KB in / 129 KB out: latency-bound, so NCCL's collectives are used as is.

The kernels are reached through an `ops` object (DeviceOps = the CUDA path); the collectives through
torch.distributed.  The CPU test-suite drives the same code with gloo and a stand-in `ops`.

PeerShardedLincomb is the same step with the exchange FUSED into the finish kernel over NVLink peer memory
(include/mfb200.h, mfb_peer_*): every rank's finish kernel pushes its partial into all ranks' symmetric buffers,
waits for the others' and adds them — two kernel launches per sharded lincomb, as on one GPU, and no collective
call on the data path.  torch.distributed is only used once, to hand the 64-byte CUDA IPC handles around.
"""
from __future__ import annotations

from dataclasses import dataclass

NCP, L64, L32 = 1472, 11, 22


@dataclass(frozen=True)
class ShardPlan:
    world: int
    rank: int

    def __post_init__(self):
        if self.world < 1 or not (0 <= self.rank < self.world):
            raise ValueError("bad rank / world size")

    def ct_range(self, d_total: int):
        """Contiguous slice [first, first + count) of the ciphertext indices; the remainder goes to the low ranks."""
        base, extra = divmod(d_total, self.world)
        first = self.rank * base + min(self.rank, extra)
        return first, base + (1 if self.rank < extra else 0)

    @property
    def coords_per_rank(self) -> int:
        """Coordinates a rank owns after the NCCL reduce-scatter; only that exchange needs an even split (the
        peer-memory exchange works for any world size <= 16)."""
        if NCP % self.world:
            raise ValueError(f"the NCCL exchange needs a world size that divides {NCP} (= 2^6 * 23): 1, 2, 4, 8, 16, 23, ...")
        return NCP // self.world

    @property
    def first_coord(self) -> int:
        return self.rank * self.coords_per_rank


class DeviceOps:
    """The CUDA kernels behind include/mfb200.h, on torch int64 CUDA tensors (raw pointers, current stream)."""

    def __init__(self, ctx, torch):
        self.ctx, self.torch = ctx, torch

    def _st(self):
        return self.torch.cuda.current_stream().cuda_stream

    def lincomb(self, cts, coeffs, d, out_flat):
        self.ctx.lincomb_dev(cts.data_ptr(), coeffs.data_ptr(), d, None, out_flat.data_ptr(), self._st())

    def columns_split(self, flat, cols):
        self.ctx.columns_split_dev(flat.data_ptr(), cols.data_ptr(), self._st())

    def encrypt(self, seed, sk_planar, msg, ent, first, count, out):
        """out[count*92] <- records of ciphertexts [first, first+count) (msg / ent are this rank's slices)."""
        self.ctx.encrypt_dev(seed, first * 135240, sk_planar.data_ptr(), msg.data_ptr(), ent.data_ptr(), 70, 69, count,
                             out.data_ptr(), self._st())

    def columns_carry(self, cols_own, first_coord, ncoord, out_own_flat):
        # the kernel indexes its output by absolute coordinate: bias the pointer so that it lands in the slice
        self.ctx.columns_carry_dev(cols_own.data_ptr(), first_coord, ncoord, None,
                                   out_own_flat.data_ptr() - first_coord * L64 * 8, self._st())


class ShardedLincomb:
    """Buffers + the per-step exchange for one rank.  `new_i64(n)` allocates a zeroed int64 tensor of n elements
    on the rank's device."""

    def __init__(self, plan: ShardPlan, ops, dist, new_i64):
        self.plan, self.ops, self.dist = plan, ops, dist
        own = plan.coords_per_rank
        self.partial = new_i64(NCP * L64)      # this rank's canonical partial sum, flat, padded to 1472 coordinates
        self.cols = new_i64(NCP * L32)
        self.cols_own = new_i64(own * L32)
        self.own_flat = new_i64(own * L64)
        self.result = new_i64(NCP * L64)       # full result, every rank

    def step(self, cts, coeffs, d_local):
        """result <- sum over ALL ranks' ciphertexts; returns the flat [1472][11] tensor (coordinate 1471 is padding)."""
        p = self.plan
        if p.world == 1:
            self.ops.lincomb(cts, coeffs, d_local, self.result)
            return self.result
        self.ops.lincomb(cts, coeffs, d_local, self.partial)
        return self.exchange()

    def exchange(self):
        """Combine the ranks' canonical partial sums (already in self.partial) into self.result on every rank."""
        p = self.plan
        if p.world == 1:
            self.result.copy_(self.partial)
            return self.result
        self.ops.columns_split(self.partial, self.cols)
        self.dist.reduce_scatter_tensor(self.cols_own, self.cols, op=self.dist.ReduceOp.SUM)
        self.own_flat.zero_()
        self.ops.columns_carry(self.cols_own, p.first_coord, p.coords_per_rank, self.own_flat)
        self.dist.all_gather_into_tensor(self.result, self.own_flat)
        return self.result


def gather_peer_handles(plan: ShardPlan, dist, my_handle: bytes, new_u8) -> bytes:
    """Every rank's 64-byte IPC handle, concatenated in rank order (one all-gather; also a barrier: a handle is only
    published after its buffer has been zeroed)."""
    import torch
    mine = new_u8(64)
    mine.copy_(torch.frombuffer(bytearray(my_handle), dtype=torch.uint8))
    if plan.world == 1:
        return bytes(mine.cpu().numpy().tobytes())
    out = new_u8(64 * plan.world)
    dist.all_gather_into_tensor(out, mine)
    return bytes(out.cpu().numpy().tobytes())


class PeerShardedLincomb:
    """Sharded lincomb with the exchange fused into the finish kernel over peer memory.  `group` is this rank's
    api.PeerGroup (or a test double with ipc_handle / connect / lincomb_dev / eval_poly_dev / check / disconnect /
    close).  Results alternate between two buffers so that call i+1 can start while the host still reads result i."""

    def __init__(self, plan: ShardPlan, group, dist, new_i64, new_u8, stream_of=lambda: 0):
        self.plan, self.group, self.dist, self.stream_of = plan, group, dist, stream_of
        handles = gather_peer_handles(plan, dist, bytes(group.ipc_handle), new_u8)
        if plan.world > 1:
            group.connect(handles)
        self.results = [new_i64(NCP * L64), new_i64(NCP * L64)]
        self.calls = 0

    def step(self, cts, coeffs, d_local, rop_in=None):
        """result <- rop_in + sum over ALL ranks' ciphertexts, on every rank; flat [1472][11] (1471 is padding)."""
        out = self.results[self.calls % 2]
        self.calls += 1
        self.group.lincomb_dev(cts.data_ptr(), coeffs.data_ptr(), d_local, None if rop_in is None else rop_in.data_ptr(),
                               out.data_ptr(), self.stream_of())
        return out

    def step_fused(self, seed, offset, c8, coeffs, d_local, rop_in=None):
        """the same with the rank's a-vectors regenerated by AES in-kernel (nothing resident)"""
        out = self.results[self.calls % 2]
        self.calls += 1
        self.group.eval_poly_dev(seed, offset, c8.data_ptr(), coeffs.data_ptr(), None, d_local,
                                 None if rop_in is None else rop_in.data_ptr(), out.data_ptr(), self.stream_of())
        return out

    def check(self):
        """raises if a peer never arrived (call after synchronising the stream)"""
        self.group.check()

    def close(self):
        if self.plan.world > 1:
            self.dist.barrier()  # nobody is still inside an exchange
        self.group.disconnect()
        if self.plan.world > 1:
            self.dist.barrier()  # every rank has unmapped its peers' buffers: they can be freed
        self.group.close()


class PipelinedPeerShardedLincomb:
    """Back-to-back sharded lincombs over peer memory: the plain lincomb (k_lincomb + k_lincomb_finish) runs on the
    caller's stream into one of two partial buffers; the exchange of call i — ONE 23-CTA kernel that pushes the partial
    to every rank, waits for the others' and adds them (mfb_peer_allreduce_dev) — runs on a side stream next to the
    lincomb kernel of call i+1.  Results land in self.results[i % 2]; call drain() before reading the last ones."""

    def __init__(self, plan: ShardPlan, ctx, group, dist, new_i64, new_u8, torch):
        self.torch, self.ctx = torch, ctx
        self.inner = PeerShardedLincomb(plan, group, dist, new_i64, new_u8, lambda: torch.cuda.current_stream().cuda_stream)
        self.group, self.results = group, self.inner.results
        self.partials = [new_i64(NCP * L64), new_i64(NCP * L64)]
        self.side = torch.cuda.Stream()
        self.done = [None, None]   # event: the exchange that last read partials[k] / wrote results[k] has finished
        self.calls = 0

    def submit(self, cts, coeffs, d_local):
        t = self.torch
        k = self.calls % 2
        main = t.cuda.current_stream()
        if self.done[k] is not None:
            main.wait_event(self.done[k])          # partials[k] is about to be overwritten
        self.ctx.lincomb_dev(cts.data_ptr(), coeffs.data_ptr(), d_local, None, self.partials[k].data_ptr(), main.cuda_stream)
        ready = t.cuda.Event()
        ready.record(main)
        self.side.wait_event(ready)
        self.group.allreduce_dev(self.partials[k].data_ptr(), None, self.results[k].data_ptr(), self.side.cuda_stream)
        ev = t.cuda.Event()
        ev.record(self.side)
        self.done[k] = ev
        self.calls += 1
        return self.results[k]

    def drain(self):
        main = self.torch.cuda.current_stream()
        for ev in self.done:
            if ev is not None:
                main.wait_event(ev)

    def check(self):
        self.inner.check()

    def close(self):
        self.inner.close()


class PipelinedShardedLincomb:
    """Back-to-back sharded lincombs (the prover runs several per proof) with the exchange of call i overlapped with the
    lincomb kernel of call i+1: the lincomb runs on the caller's stream, the exchange chain (columns_split,
    reduce-scatter, columns_carry, all-gather) on a side stream, two sets of exchange buffers alternate.
    Results land in self.results[i % 2]; call drain() before reading the last ones."""

    def __init__(self, plan: ShardPlan, ops, dist, new_i64, torch):
        self.torch = torch
        self.lanes = [ShardedLincomb(plan, ops, dist, new_i64) for _ in range(2)]
        self.results = [lane.result for lane in self.lanes]
        self.side = torch.cuda.Stream()
        self.done = [None, None]   # event: exchange of the call that last used lane k has finished
        self.calls = 0

    def submit(self, cts, coeffs, d_local):
        t = self.torch
        k = self.calls % 2
        lane = self.lanes[k]
        main = t.cuda.current_stream()
        if self.done[k] is not None:
            main.wait_event(self.done[k])          # lane.partial is about to be overwritten
        if lane.plan.world == 1:
            lane.ops.lincomb(cts, coeffs, d_local, lane.result)
        else:
            lane.ops.lincomb(cts, coeffs, d_local, lane.partial)
            ready = t.cuda.Event()
            ready.record(main)
            self.side.wait_event(ready)
            with t.cuda.stream(self.side):
                lane.exchange()
                ev = t.cuda.Event()
                ev.record(self.side)
            self.done[k] = ev
        self.calls += 1
        return self.results[k]

    def drain(self):
        main = self.torch.cuda.current_stream()
        for ev in self.done:
            if ev is not None:
                main.wait_event(ev)


class ShardedSetup:
    """CRS generation (setup's 2D + M Regev encryptions, snark.c:75-110) sharded by ciphertext index: rank r encrypts
    the contiguous range ShardPlan.ct_range(count) with the stream positioned at first * 135240 and its slice of the
    messages / entropy; outputs are disjoint 92-byte records, so the only communication is the gather of the records.
    `encrypt(first, count) -> uint8 tensor (count * 92)` is the kernel call (DeviceOps.encrypt or a test double)."""

    def __init__(self, plan: ShardPlan, dist, new_u8):
        self.plan, self.dist, self.new_u8 = plan, dist, new_u8

    def run(self, total: int, encrypt):
        p = self.plan
        first, count = p.ct_range(total)
        mine = encrypt(first, count)
        if p.world == 1:
            return mine
        # ranges differ by at most one ciphertext: pad to the longest, gather, cut the padding
        longest = (total + p.world - 1) // p.world
        buf = self.new_u8(longest * 92)
        buf[: count * 92] = mine
        out = self.new_u8(p.world * longest * 92)
        self.dist.all_gather_into_tensor(out, buf)
        parts = []
        for r in range(p.world):
            _, c = ShardPlan(p.world, r).ct_range(total)
            parts.append(out[r * longest * 92: r * longest * 92 + c * 92])
        import torch
        return torch.cat(parts)
