#!/usr/bin/env python
"""Benchmark of the prover's ciphertext linear combination (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--log2d 16]

One *step* = one pass of the hot path over one batch of synthetic input: rop = sum_i h_i * CT_i mod 2^704
over D = 2^16 Regev ciphertexts per GPU (n = 1470, log q = 736 / effective 704), i.e. one eval_poly
(lwe.c:176-186).  Unit = ciphertext-MAC.

  value     whole-job ciphertext-MAC/s with the ciphertexts RESIDENT in HBM (planar layout, expanded once
            from the CRS seed by the AES-CTR kernel before the timed region), K1 = k_lincomb + k_lincomb_finish;
            CUDA events on the launching stream, max over ranks.
  roofline  k_lincomb alone (events recorded around that launch inside the library, mfb_profile_*):
            algorithmic bytes = D * 129448 per launch against MEASURED_PEAKS.json's HBM copy bandwidth.
  e2e       the same eval_poly through the host-flavour C-ABI call mfb_eval_poly with HOST buffers (seed, 92-byte
            records, coefficients, accumulator), as the reference's eval_poly(rop, rng, c8, coeffs, d) is called:
            H2D of records + scalars + rop, AES-CTR regeneration of every a-vector in-kernel (nothing is
            resident), fused MAC, D2H of the result — all inside the timed region.
  cpu_baseline  the compiled reference (oracle/_ref, unmodified lwe.c/entropy.c/aes.c) on one host core, over a
            bounded prefix of the same ciphertexts.
  N > 1     weak scaling: every rank holds its own 2^16 ciphertexts (global D = N * 2^16); a step adds the
            exchange of SURVEY §8e.  --exchange p2p (default): fused into the finish kernel over NVLink peer memory
            (k_lincomb_finish_peer: push the partial to every rank, per-tile flags, local sum).  --exchange nccl:
            widen partial sums to u64 columns, NCCL reduce-scatter, carry-propagate + truncate on the owner,
            all-gather (also what runs if CUDA IPC is not available on the box).

`--impl reference` times the reference's own CPU eval_poly with every host core (one process per core over disjoint
ciphertext ranges positioned with rng_seek, partials folded with ct_add), each step a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "ciphertext-MAC/s"
SEED = bytes(range(40))
N, NC, L64, CT_BYTES, CTR_CT, P = 1470, 1471, 11, 92, 92 * 1470, 0xFFFFFFFB
ALGO_BYTES = NC * 88
PLANAR_BYTES = L64 * 1472 * 8


def synth_inputs(d: int, rank: int = 0):
    """Seeded synthetic instance: 92-byte b records (top 4 bytes zero) and scalars uniform in [0, p)."""
    rng = np.random.Generator(np.random.PCG64(20181018 + rank))
    c8 = rng.integers(0, 256, size=(d, CT_BYTES), dtype=np.uint8)
    c8[:, 88:] = 0
    h = (rng.integers(0, 2**63, size=d, dtype=np.uint64) % np.uint64(P)).astype(np.uint64)
    return c8, h


def workload_config(log2d: int) -> dict:
    """`config` — IDENTICAL in both arms (how each arm runs the workload is described under `config_details`)."""
    return {"workload": f"prover lincomb (eval_poly, lwe.c:176-186), D=2^{log2d} Regev ciphertexts per GPU, n=1470, logq=736 "
                        f"(eff. 704), p=2^32-5",
            "log2_ciphertexts_per_gpu": log2d, "n": 1470, "logq": 736, "q_eff_bits": 704, "p": P}


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms from just before the timed region to the end of the e2e legs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU legs
def ref_lib():
    """The compiled reference (oracle/_ref/libmfref_d256_m64.so); eval_poly does not depend on the instance size."""
    from oracle.loader import Reference
    return Reference(256, 64)


def cpu_baseline_single(sample: int):
    """The compiled reference's eval_poly on ONE core over the first `sample` ciphertexts of the workload."""
    ref = ref_lib()
    c8, h = synth_inputs(sample)
    secs, _rop, t_imp, t_mac = ref.time_eval_poly(SEED, 0, c8, h, split=True)
    return {"value": sample / secs, "unit": METRIC, "cores": 1, "kind": "reference",
            "sample": f"first {sample} of the 2^16 ciphertexts, eval_poly of oracle/_ref (unmodified lwe.c) in {secs:.2f} s; "
                      f"per ciphertext: ct_import (AES regen) {t_imp * 1e6:.0f} us, ct_addmul_ui (GMP) {t_mac * 1e6:.1f} us"}


def config1_reference():
    """BASELINE configs[0]: the reference's benchmark_snark (setup / prover / verifier seconds, benchmark_snark.c:56-82) at
    the default LWE parameters on a ~2^10-constraint random SSP (D = 1024, M = 64), single-threaded CPU, and the same
    three calls through the drop-in on the GPU."""
    from oracle.loader import Reference
    try:
        ref = Reference(1024, 64)
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)[:200]}
    s_ref, p_ref, v_ref, ok = ref.benchmark_snark()
    ours = snark_latency(10, 64)
    return {"D": 1024, "M": 64, "reference_cpu_ms": {"setup": 1e3 * s_ref, "prove": 1e3 * p_ref, "verify": 1e3 * v_ref, "accept": ok,
                                                       "cores": 1},
            "b200_ms": {"setup": ours["setup_ms"], "prove": ours["prove_ms"], "prove_resident": ours["prove_resident_ms"],
                        "verify": ours["verify_ms"], "accept": ours["accept"]}}


def _ref_worker(args):
    first, count, reps = args
    ref = ref_lib()
    c8, h = synth_inputs(first + count)
    out = []
    for _ in range(reps):
        t, rop = ref.time_eval_poly(SEED, first * CTR_CT, c8[first:first + count], h[first:first + count])
        out.append(t)
    return out, rop


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU eval_poly on all host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cores = max(1, min(cores, 256))
    steps, warm = args.steps, args.warmup
    # ~2 ms per ciphertext-MAC per core; keep the whole run near two minutes
    per_core = max(16, min(2048, int(120.0 / (steps + warm) / 2.1e-3)))
    total = per_core * cores
    ranges = [(k * per_core, per_core, steps + warm) for k in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_ref_worker, ranges)
    wall = time.perf_counter() - t0
    ref = ref_lib()
    # fold the partial sums as the reference would (ct_add), once; its cost is included in the step time below
    t1 = time.perf_counter()
    acc = res[0][1]
    for _, rop in res[1:]:
        acc = ref.ct_add(acc, rop)
    fold = time.perf_counter() - t1
    # a step ends when its slowest worker ends
    step_times = [max(res[k][0][s] for k in range(cores)) + fold for s in range(warm, warm + steps)]
    t = sum(step_times)
    value = total * steps / t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 * t / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(args.log2d),
            "config_details": {"path": "eval_poly of the unmodified reference (oracle/_ref): AES-CTR regeneration + GMP MAC on the "
                                       "host cores", "sample_ciphertexts_per_step": total},
            "cpu_baseline": {"value": value, "unit": METRIC, "cores": cores, "kind": "reference",
                             "sample": f"{total} ciphertexts per step ({per_core} per core x {cores} processes over disjoint "
                                       f"ranges via rng_seek, partials folded with ct_add); wall {wall:.1f} s"},
            "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def snark_latency(log2d: int, M: int):
    """BASELINE configs[2]: full setup + prove + verify of a 2^log2d-constraint SSP through the drop-in C interface
    (setup/prover/verifier of snark.h:44-51, OS entropy, random degenerate SSP of ssp.c:37-77).  Wall-clock of each
    call, host polynomial arithmetic included.  `prove_ms` regenerates every a-vector from AES (as the reference
    does); `prove_resident_ms` is the same call after mf_crs_make_resident (regions s / as kept in HBM)."""
    from c_lwe_snarks_b200.snark import Snark
    D = 1 << log2d
    sn = Snark(D, M)
    try:
        sn.random_ssp()
        t_setup_cold = sn.setup()  # first call: pinned staging buffers, device scratch
        t_setup = sn.setup()
        sn.prove()  # warm-up (scratch allocation)
        t_prove = min(sn.prove() for _ in range(3))
        ok, t_verify = sn.verify()
        t0 = time.perf_counter()
        sn.make_resident()
        t_res = time.perf_counter() - t0
        sn.prove()
        t_prove_res = min(sn.prove() for _ in range(3))
        ok2, _ = sn.verify()
        sn.tamper()
        bad, _ = sn.verify()
    finally:
        sn.close()
    return {"D": D, "M": M, "setup_ms": 1e3 * t_setup, "setup_first_call_ms": 1e3 * t_setup_cold, "prove_ms": 1e3 * t_prove,
            "verify_ms": 1e3 * t_verify,
            "make_resident_ms": 1e3 * t_res, "prove_resident_ms": 1e3 * t_prove_res, "accept": bool(ok and ok2),
            "tampered_accept": bool(bad), "entropy": "getrandom(2)", "api": "setup/prover/verifier (snark.h:44-51) via libmangiafuoco_b200.so"}


def default_instance():
    """The reference's OWN default instance (NDEBUG: D = 2^15, M = 21845, lwe.h:14-17) through its UNMODIFIED benchmark
    main (benchmark_snark.c:27-96), compiled against the drop-in headers and linked with the product (oracle/Makefile
    `dropin-tests`; built only where the reference sources are present).  The program knows nothing of this library's
    additions: setup() uploads the 5.7 GB SSP blob once and leaves it resident for prover() and verifier()."""
    exe = ROOT / "oracle" / "_ref" / "dropin_benchmark_snark_default"
    if not exe.exists():
        return {"unavailable": "oracle/_ref/dropin_benchmark_snark_default is built only where the reference sources are present"}
    t0 = time.perf_counter()
    try:
        r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600, env={**os.environ, "MF_B200_TRACE": "1"})
    except subprocess.TimeoutExpired:
        return {"error": "timeout"}
    wall = time.perf_counter() - t0
    out = {"D": 1 << 15, "M": 21845, "accept": r.returncode == 0, "program_wall_s": round(wall, 2),
           "program": "benchmark_snark.c of the reference, unmodified, over libmangiafuoco_b200.so"}
    for ln in r.stdout.splitlines():
        f = ln.split("\t")
        if len(f) == 2 and f[0] in ("setup", "prover", "verifier"):
            out[f[0] + "_ms"] = 1e3 * float(f[1])
    phases = {}
    for ln in r.stderr.splitlines():
        f = ln.split("\t")
        if len(f) == 2 and (f[0].startswith("setup.") or f[0].startswith("prover.")):
            try:
                phases[f[0]] = round(1e3 * float(f[1]), 3)
            except ValueError:
                pass
    out["phases_ms"] = phases
    return out


# ------------------------------------------------------------------------------------------ GPU arm
def snark_box_latency(log2d_total: int, M: int, n_dev: int):
    """BASELINE configs[3]: one single-threaded program (the drop-in's setup/prover/verifier) proving a 2^log2d_total-
    constraint SSP with the CRS regions sharded by ciphertext index over n_dev GPUs (mf_set_devices) and every lincomb
    combined over NVLink peer memory.  Runs after the timed sections as a SUBPROCESS of rank 0 (tools/snark_box.py: the
    drop-in aborts on errors, which must not take the bench line with it); the other ranks wait on the CPU.
    setup_ms = setup() with its encryptions spread over the n_dev GPUs (one host thread per GPU draws the entropy of its
    range); setup_one_gpu_ms = the same call before mf_set_devices."""
    import subprocess
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "snark_box.py"), str(log2d_total), str(M), str(n_dev)],
                       capture_output=True, text=True, timeout=600)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    if r.returncode != 0 or not lines:
        raise RuntimeError(f"tools/snark_box.py exited {r.returncode}: {(r.stderr or r.stdout)[-300:]}")
    j = json.loads(lines[-1])
    return {"D": j["D"], "M": j["M"], "devices": j["devices"], "setup_ms": 1e3 * j["setup_over_the_set_s"],
            "setup_one_gpu_ms": 1e3 * j["setup_s"],
            "make_resident_ms": 1e3 * j["make_resident_s"], "prove_nothing_resident_ms": j["prove_nothing_resident_ms"],
            "prove_resident_ms": min(j["prove_ms"][1:]), "verify_ms": j["verify_ms"], "accept": j["accept"],
            "tampered_accept": j["tampered_accept"], "api": j["api"]}


def strong_2e20_leg(ctx, torch, dist, world, rank, steps, peer, ppipe, sl, barrier, st):
    """BASELINE configs[3] / north_star: ONE lincomb over D = 2^20 ciphertexts (snark.c:157-174 at GAMMA_D = 2^20) split by
    ciphertext index over the N GPUs — rank r holds the contiguous range r of 2^20 / N ciphertexts resident (N = 1: the whole
    136 GB region on one B200).  Two timings per lincomb: `latency` = lincomb + exchange back to back, the exchange on the
    critical path (what one proof sees); `pipelined` = the exchange of lincomb i overlapped with lincomb i + 1 (throughput of
    a stream of proofs).  Parity: rank 0 recomputes the whole sum alone through the fused AES + MAC kernel."""
    Dt = 1 << 20
    if Dt % world:
        return {"skipped": f"2^20 is not divisible by {world} ranks"}
    Dg, first = Dt // world, rank * (Dt // world)
    c8g, hg = synth_inputs(Dt, 0)  # the same global instance on every rank
    try:
        d_c8 = torch.from_numpy(c8g[first: first + Dg].reshape(-1).copy()).cuda()
        d_h = torch.from_numpy(hg[first: first + Dg].astype(np.uint32).view(np.int32)).cuda()
        d_big = torch.empty(Dg * PLANAR_BYTES, dtype=torch.uint8, device="cuda")
    except RuntimeError as e:  # (another tenant on the GPU: report, do not fail the bench line)
        return {"skipped": f"allocation of {Dg * PLANAR_BYTES / 1e9:.1f} GB failed: {str(e)[:120]}"}
    ctx.expand_dev(SEED, first * CTR_CT, d_c8.data_ptr(), Dg, d_big.data_ptr(), st)
    torch.cuda.synchronize()

    def one(coeffs):
        if peer is not None:
            return peer.step(d_big, coeffs, Dg)
        return sl.step(d_big, coeffs, Dg)

    for _ in range(3):
        res = one(d_h)
    barrier()
    ctx.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        res = one(d_h)
    e1.record()
    barrier()
    k_ms, k_n = ctx.profile_end()
    lat_ms = e0.elapsed_time(e1) / steps
    got = res.cpu().numpy().view(np.uint64)[: NC * L64].reshape(NC, L64).copy()
    pipe_ms = lat_ms
    if ppipe is not None:
        for _ in range(3):
            ppipe.submit(d_big, d_h, Dg)
        ppipe.drain()
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for _ in range(steps):
            ppipe.submit(d_big, d_h, Dg)
        ppipe.drain()
        e3.record()
        barrier()
        pipe_ms = e2.elapsed_time(e3) / steps
    t_all = torch.tensor([lat_ms, pipe_ms, k_ms / max(1, k_n)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    lat_ms, pipe_ms, kern_ms = (float(x) for x in t_all.tolist())
    ok = None
    if rank == 0:  # independent of the exchange: the whole 2^20-ciphertext sum on this GPU alone, a regenerated from AES
        ok = bool(np.array_equal(ctx.eval_poly(SEED, 0, c8g, hg), got))
        if not ok:
            raise SystemExit("bench.py: strong_2e20: the sharded result differs from rank 0's own computation — numbers withheld")
    barrier()
    del d_big
    torch.cuda.empty_cache()
    return {"D_total": Dt, "ciphertexts_per_gpu": Dg, "resident_gb_per_gpu": Dg * PLANAR_BYTES / 1e9,
            "ms_per_lincomb_latency": lat_ms, "mac_per_s_latency": Dt / (lat_ms * 1e-3),
            "ms_per_lincomb_pipelined": pipe_ms, "mac_per_s_pipelined": Dt / (pipe_ms * 1e-3),
            "kernel_ms": kern_ms, "kernel_gbs": Dg * ALGO_BYTES / (kern_ms * 1e-3) / 1e9,
            "parity_independent_recompute": ok, "scaling": "strong",
            "note": "latency: k_lincomb + finish (N > 1: fused with the peer-memory exchange) back to back, max over ranks; "
                    "pipelined: exchange on a side stream under the next lincomb; kernel_*: k_lincomb alone (CUDA events in the library)"}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    import c_lwe_snarks_b200 as m

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = m.Context(local)
    D = 1 << args.log2d
    steps, warm = args.steps, max(3, args.warmup)
    st = torch.cuda.current_stream().cuda_stream

    # ---- synthetic instance, made resident (untimed): expand the rank's ciphertext range from the seed
    c8, h = synth_inputs(D, rank)
    stream_off = rank * D * CTR_CT  # rank r owns ciphertext indices [r*D, (r+1)*D) of the global stream
    d_c8 = torch.from_numpy(c8.reshape(-1)).cuda()
    d_h = torch.from_numpy(h.astype(np.uint32).view(np.int32)).cuda()
    d_cts = torch.empty(D * PLANAR_BYTES, dtype=torch.uint8, device="cuda")
    ctx.expand_dev(SEED, stream_off, d_c8.data_ptr(), D, d_cts.data_ptr(), st)
    from c_lwe_snarks_b200.sharding import DeviceOps, PipelinedShardedLincomb, ShardedLincomb, ShardPlan
    plan = ShardPlan(world, rank)
    assert plan.ct_range(world * D) == (rank * D, D)
    sl = ShardedLincomb(plan, DeviceOps(ctx, torch), dist, lambda n: torch.zeros(n, dtype=torch.int64, device="cuda"))
    d_rop = sl.result
    torch.cuda.synchronize()

    # the timed steps are independent eval_polys (a proof runs several): for N > 1 the exchange of step i overlaps
    # the lincomb kernel of step i+1 (side stream, alternating exchange buffers)
    new_i64 = lambda n: torch.zeros(n, dtype=torch.int64, device="cuda")  # noqa: E731
    new_u8 = lambda n: torch.zeros(n, dtype=torch.uint8, device="cuda")  # noqa: E731
    pipe = peer = ppipe = None
    exchange = "none"
    if world > 1 and args.exchange == "p2p":
        # the exchange fused into the finish kernel over peer memory; every rank must agree on the outcome of the
        # IPC hand-shake, so a rank that cannot map its peers makes ALL ranks take the NCCL exchange
        from c_lwe_snarks_b200.sharding import PipelinedPeerShardedLincomb
        ok, why = 1, ""
        try:
            group = ctx.peer_group(world, rank)
            ppipe = PipelinedPeerShardedLincomb(plan, ctx, group, dist, new_i64, new_u8, torch)
            peer = ppipe.inner
        except Exception as e:  # noqa: BLE001
            ok, why, peer, ppipe = 0, str(e), None, None
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if why:
                print(f"bench.py: rank {rank}: peer-memory exchange unavailable ({why}); using the NCCL exchange", file=sys.stderr)
            peer = ppipe = None
        else:
            exchange = ("over NVLink peer memory (CUDA IPC), no collective call on the data path: ONE 23-CTA kernel per step "
                        "on a side stream pushes the rank's partial sum into every rank's symmetric buffer, waits on per-tile "
                        "flags and adds the ranks' tiles, next to the following step's lincomb kernel")
    if world > 1 and peer is None:
        pipe = PipelinedShardedLincomb(plan, DeviceOps(ctx, torch), dist, new_i64, torch)
        exchange = ("u64-column reduce-scatter + carry + all-gather (NCCL) on a side stream, overlapped with the next "
                    "step's lincomb kernel")

    def step():
        if ppipe is not None:
            return ppipe.submit(d_cts, d_h, D)
        if pipe is None:
            return sl.step(d_cts, d_h, D)
        return pipe.submit(d_cts, d_h, D)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ctx.launches
    ctx.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        step()
    if pipe is not None:
        pipe.drain()
    if ppipe is not None:
        ppipe.drain()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if pipe is not None:
        d_rop = pipe.results[(pipe.calls - 1) % 2]
    if ppipe is not None:
        ppipe.check()
        d_rop = ppipe.results[(ppipe.calls - 1) % 2]
    k_ms, k_n = ctx.profile_end()
    launches = ctx.launches - l0  # our kernels only (NCCL's and torch's zero_ are not counted)
    t_all = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    ms = float(t_all.item())
    result = d_rop.cpu().numpy().view(np.uint64)[: NC * L64].reshape(NC, L64).copy()

    def sync_step(coeffs, cts=None, d_local=None):
        """one sharded lincomb with the exchange ON the critical path (no pipelining across steps)"""
        cts = d_cts if cts is None else cts
        d_local = D if d_local is None else d_local
        if peer is not None:
            return peer.step(cts, coeffs, d_local)   # k_lincomb + the finish kernel fused with the peer exchange
        return sl.step(cts, coeffs, d_local)           # one GPU, or the NCCL exchange

    # ---- parity outside the timed regions (1): a 64-ciphertext sample against the CPU oracle, through the SAME sharded
    # resident path (scalars zero outside the sample: every rank contributes a run of 64 / N ciphertexts)
    n_s = max(1, 64 // world)
    run_at = lambda r: (7919 * (r + 1)) % (D - n_s)  # noqa: E731
    h_sp = np.zeros(D, np.uint32)
    h_sp[run_at(rank): run_at(rank) + n_s] = h[run_at(rank): run_at(rank) + n_s].astype(np.uint32)
    d_hs = torch.from_numpy(h_sp.view(np.int32)).cuda()
    got_s = sync_step(d_hs)
    torch.cuda.synchronize()
    got_s = got_s.cpu().numpy().view(np.uint64)[: NC * L64].reshape(NC, L64).copy()
    oracle_ok = multi_ok = None
    if rank == 0:
        from oracle.loader import Oracle  # the CPU restatement, here only as the checker
        orc = Oracle()
        want = None
        for r in range(world):
            c8_r, h_r = (c8, h) if r == rank else synth_inputs(D, r)
            j = run_at(r)
            want = orc.eval_poly(SEED, (r * D + j) * CTR_CT, c8_r[j: j + n_s], h_r[j: j + n_s], rop=want)
        oracle_ok = bool(np.array_equal(got_s, want[:, :L64]))
        if not oracle_ok:
            raise SystemExit("bench.py: the sharded resident lincomb differs from the CPU oracle on the 64-ciphertext sample — numbers withheld")
        # (2) N > 1: rank 0 ALONE recomputes the global sum of the timed step — every rank's ciphertext range through the
        # fused AES + MAC kernel on this one GPU, no exchange involved — and compares it bit for bit with the exchanged result
        if world > 1:
            acc = np.zeros((NC, L64), np.uint64)
            for r in range(world):
                c8_r, h_r = (c8, h) if r == rank else synth_inputs(D, r)
                acc = ctx.eval_poly(SEED, r * D * CTR_CT, c8_r, h_r, rop=acc)
            multi_ok = bool(np.array_equal(acc, result))
            if not multi_ok:
                raise SystemExit("bench.py: the exchanged result differs from rank 0's own computation of the global sum — numbers withheld")
    barrier()

    # ---- informational: the two-scalar-vector pass the prover uses over a resident region (k_lincomb<2>)
    d_r0 = torch.zeros(1472 * L64, dtype=torch.int64, device="cuda")
    d_r1 = torch.zeros(1472 * L64, dtype=torch.int64, device="cuda")
    d_h1 = torch.roll(d_h, 1)
    for _ in range(3):
        ctx.lincomb2_dev(d_cts.data_ptr(), d_h.data_ptr(), d_h1.data_ptr(), D, None, d_r0.data_ptr(), None, d_r1.data_ptr(), st)
    torch.cuda.synchronize()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(steps):
        ctx.lincomb2_dev(d_cts.data_ptr(), d_h.data_ptr(), d_h1.data_ptr(), D, None, d_r0.data_ptr(), None, d_r1.data_ptr(), st)
    e3.record()
    torch.cuda.synchronize()
    pair_ms = e2.elapsed_time(e3) / steps
    pair_ok = bool(np.array_equal(d_r0.cpu().numpy().view(np.uint64)[: NC * L64].reshape(NC, L64), result)) if world == 1 else None

    # ---- e2e: host buffers through mfb_eval_poly (fused AES regeneration), every rank its own range
    e2e_steps = max(2, min(steps, 5))
    pin = lambda a: torch.from_numpy(a).pin_memory()  # noqa: E731
    p_c8, p_h = pin(c8.reshape(-1).copy()), pin(h.copy())
    p_rop = torch.zeros(NC * L64, dtype=torch.int64).pin_memory()
    np_c8, np_h, np_rop = p_c8.numpy(), p_h.numpy(), p_rop.numpy().view(np.uint64).reshape(NC, L64)

    d_c8_e = torch.empty(D * CT_BYTES, dtype=torch.uint8, device="cuda")
    d_h_e = torch.empty(D, dtype=torch.int32, device="cuda")
    p_h32 = pin(h.astype(np.uint32).view(np.int32).copy())
    seed_arr = np.frombuffer(SEED, np.uint8).copy()

    def e2e_step():
        if world == 1:
            np_rop[:] = 0
            ctx._ck(ctx.lib.mfb_eval_poly(ctx.h, m.api._p8(seed_arr), stream_off, m.api._p8(np_c8), m.api._p64(np_h), None, D,
                                          m.api._p64(np_rop)))
        else:
            # every rank: pinned host records + scalars -> device, fused AES + MAC over its own ciphertext range,
            # the same exchange as the resident path, result back to pinned host memory
            d_c8_e.copy_(p_c8, non_blocking=True)
            d_h_e.copy_(p_h32, non_blocking=True)
            if peer is not None:
                res = peer.step_fused(SEED, stream_off, d_c8_e, d_h_e, D)
            else:
                ctx.eval_poly_dev(SEED, stream_off, d_c8_e.data_ptr(), d_h_e.data_ptr(), None, D, None, sl.partial.data_ptr(), st)
                res = sl.exchange()
            p_rop.copy_(res[: NC * L64], non_blocking=True)
            torch.cuda.current_stream().synchronize()

    e2e_step()
    fused_result = np_rop.copy()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t_all = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    e2e_s = float(t_all.item())
    # ---- e2e with the CRS region resident (mf_crs_make_resident / mfb_region_*): host scalars in, host result out
    reg = ctx.region(SEED, stream_off, c8) if world == 1 else None
    co32 = torch.from_numpy(h.astype(np.uint32).view(np.int32).copy()).pin_memory().numpy().view(np.uint32)

    d_co_r = torch.empty(D, dtype=torch.int32, device="cuda")
    p_co32 = torch.from_numpy(co32.view(np.int32))

    def res_step():
        if world == 1:
            np_rop[:] = 0
            ctx._ck(ctx.lib.mfb_region_lincomb(ctx.h, reg.handle, 0, m.api._p32(co32), D, m.api._p64(np_rop)))
        else:
            d_co_r.copy_(p_co32, non_blocking=True)
            res = peer.step(d_cts, d_co_r, D) if peer is not None else sl.step(d_cts, d_co_r, D)
            p_rop.copy_(res[: NC * L64], non_blocking=True)
            torch.cuda.current_stream().synchronize()

    res_step()
    resident_result = np_rop.copy()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        res_step()
    barrier()
    res_s = time.perf_counter() - t0
    if reg is not None:
        reg.close()
    t_all = torch.tensor([res_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    res_s = float(t_all.item())
    if rank == 0 and not np.array_equal(resident_result, fused_result):
        raise SystemExit("bench.py: region lincomb and fused eval_poly disagree — numbers withheld")
    clocks = sampler.stop() if sampler else None

    # ---- sanity outside the timed regions: resident path == fused path (two independent kernels, same exchange)
    check = bool(np.array_equal(result, fused_result)) if rank == 0 else None
    if rank == 0 and not check:
        raise SystemExit("bench.py: resident lincomb and fused eval_poly disagree — numbers withheld")

    parity = {"resident_equals_fused": check, "oracle_sample_64_ciphertexts": oracle_ok,
              "multi_gpu_independent_recompute": multi_ok,
              "what": "resident lincomb == fused AES+MAC eval_poly (two independent kernels); a 64-ciphertext sample through the "
                      "sharded resident path == the CPU oracle; N > 1: the exchanged result == rank 0 recomputing the global sum "
                      "alone (mfb_eval_poly over every rank's range, no exchange)"}

    # ---- N > 1: the full drop-in SNARK from ONE host thread over the N GPUs (a subprocess of rank 0): the north-star
    # instance (D = 2^20) and the weak-scaling instance (D = N * 2^16: per-GPU work as in the N = 1 `snark` leg).  The other
    # ranks free their big buffers and wait ON THE CPU (a c10d store key): an NCCL barrier would keep a kernel spinning
    # on their GPUs, and rank 0's member contexts on those GPUs would be time-sliced against it.  (Runs before the
    # strong-scaling leg: memory another process has just freed is scrubbed by the driver when it is allocated again,
    # which would sit in make_resident_ms.)
    if reg is not None:
        reg = None
    del d_cts, d_c8_e, d_h_e, d_co_r
    torch.cuda.empty_cache()
    boxes = {}
    if world > 1 and not args.no_snark:
        from datetime import timedelta
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            for key, lg in (("snark_box", 20), ("snark_box_weak", args.log2d + (world - 1).bit_length())):
                try:
                    boxes[key] = snark_box_latency(lg, 64, world)
                except Exception as e:  # noqa: BLE001  (reported, the bench line stands)
                    boxes[key] = {"error": str(e)[:300]}
            store.set("snark_box_done", "1")
        else:
            store.wait(["snark_box_done"], timedelta(seconds=1500))

    # ---- BASELINE configs[3]: the 2^20-ciphertext lincomb split by ciphertext index over the N GPUs (strong scaling)
    strong = None
    if not args.no_strong:
        strong = strong_2e20_leg(ctx, torch, dist, world, rank, steps, peer, ppipe, sl, barrier, st)

    if rank == 0:
        peak, peak_src = peaks()
        per_launch_ms = k_ms / max(1, k_n)
        achieved = D * ALGO_BYTES / (per_launch_ms * 1e-3) / 1e9
        traffic = None
        tf = ROOT / "profiles" / "k_lincomb_traffic.json"
        if tf.exists():
            try:
                traffic = json.loads(tf.read_text()).get("dram_bytes_per_launch")
            except (ValueError, OSError):
                traffic = None
        line = {
            "metric": METRIC, "value": world * D * steps / (ms * 1e-3), "unit": METRIC, "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": workload_config(args.log2d),
            "config_details": {"path": f"ciphertexts resident in HBM ({D * PLANAR_BYTES / 1e9:.2f} GB tile-planar per GPU), "
                                       f"k_lincomb + finish",
                               "ciphertexts_per_gpu": D, "l2": "inputs (8.49 GB per step) exceed the 126 MB L2; no flush needed",
                               "exchange": exchange},
            "parity_check": parity,
            "roofline": {"bound": "hbm", "kernel": "k_lincomb", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": D * ALGO_BYTES, "kernel_ms": per_launch_ms, "launches_timed": k_n},
            "e2e": {"value": world * D * e2e_steps / e2e_s, "unit": METRIC,
                    "h2d_bytes_per_step": D * (CT_BYTES + 4) + NC * 88, "d2h_bytes_per_step": NC * 88,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps,
                    "path": "mfb_eval_poly (host buffers; a regenerated by AES-256-CTR in-kernel, nothing resident)"},
            "e2e_resident": {"value": world * D * steps / res_s, "unit": METRIC, "h2d_bytes_per_step": D * 4 + NC * 88,
                             "d2h_bytes_per_step": NC * 88, "ms_per_step": 1e3 * res_s / steps, "steps": steps,
                             "path": "mfb_region_lincomb (host scalars -> host result; the CRS region was expanded into HBM "
                                     "once by mfb_region_create, as mf_crs_make_resident does for prover())"},
            "two_vector_pass": {"ms_per_pass": pair_ms, "mac_per_s": 2 * D / (pair_ms * 1e-3), "matches_single": pair_ok,
                                "note": "k_lincomb<2>: two scalar vectors over the same resident ciphertexts in one pass "
                                        "(the prover pairs v_w/h and hat_v/hat_h); per-GPU, no exchange"},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if world > 1:
            line["parity_multi_gpu"] = bool(multi_ok)
        if strong is not None:
            peak_s = peak
            strong["roofline_frac"] = strong["kernel_gbs"] / peak_s
            line["strong_2e20"] = strong
        if world == 1 and not args.no_snark:
            line["snark"] = snark_latency(args.log2d, 64)
        if world == 1 and not args.no_snark and not args.no_default_instance:
            line["default_instance"] = default_instance()
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_single(args.cpu_sample)
            line["config1_reference_cpu"] = config1_reference()
        line.update(boxes)
        emit(line)
    if peer is not None:
        peer.check()
        peer.close()
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else any library prints (NCCL's version banner, torchrun
    notices) was redirected to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2d", type=int, default=16, help="ciphertexts per GPU = 2^log2d (BASELINE configs[1]: 16)")
    ap.add_argument("--cpu-sample", type=int, default=6000, help="ciphertexts timed by the 1-core cpu_baseline leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-snark", action="store_true", help="skip the full setup/prove/verify latency leg")
    ap.add_argument("--no-default-instance", action="store_true", help="skip the reference's default instance (D=2^15, M=21845) leg")
    ap.add_argument("--no-strong", action="store_true", help="skip the 2^20-ciphertext strong-scaling leg")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: p2p = exchange fused into the finish kernel over NVLink peer memory (CUDA IPC); "
                         "nccl = u64-column reduce-scatter + carry + all-gather")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
