// K1 — prover ciphertext linear combination over ciphertexts RESIDENT in HBM.
//
// Replaces the inner loop of eval_poly (lwe.c:176-186): for every coordinate c,
//   rop[c] = (rop[c] + sum_i coeff_i * CT_i[c]) mod 2^704           (ct_addmul_ui lwe.c:141-149 + modq lwe.h:108-118)
// Exact sums mod 2^704 are associative, so splitting the i-range over CTAs (and GPUs) and adding the
// canonical partial sums afterwards is bit-identical to the reference's sequential fold.
//
// HBM-bound: 129 536 B per ciphertext (129 448 B algorithmic), 22 IMAD.WIDE per coordinate.  The kernel is a
// streaming reduction built on the Blackwell copy engine:
//   * resident layout = tile-planar (mfb_common.cuh): (ciphertext, 64-coordinate tile) is one contiguous 5632 B
//     block, moved global -> shared by ONE `cp.async.bulk` (TMA, SASS UBLKCP) that completes on an mbarrier;
//   * CTA = 64 threads bound to one coordinate tile, a ring of LC_STAGES x LC_G blocks in shared memory; thread 0
//     is the producer (expect_tx + bulk copies), both warps consume (11 conflict-free LDS.64 per ciphertext) and
//     release the slot through a second mbarrier;
//   * work distribution is DYNAMIC: every tile has a queue of chunks of the ciphertext range (an atomic counter);
//     CTAs of that tile pull chunks until the queue is empty.  All CTAs are co-resident (grid = 23 tiles x
//     floor(SMs * occupancy / 23) slots), so there is no wave quantisation and SMs that run faster take more
//     chunks — measured 1.13 ms for D = 2^16 (7.5 TB/s) against 1.49 ms for the static LDG version;
//   * each thread owns one coordinate and keeps the 704-bit accumulator pair (E, O) in registers for all the
//     chunks it pulls; at the end the CTA writes one canonical partial sum; k_lincomb_finish adds the partials,
//     adds the incoming rop and resets the queues.
#include "mfb_common.cuh"

namespace mfb {

constexpr int LC_TILE = RT_TILE;  // 64 threads per CTA = coordinates per tile
// ring depth: 2 stages for one scalar vector (9 CTAs per SM: 203 KB of copies in flight per SM); the two-vector
// kernel holds two accumulators (157 registers: 6 CTAs per SM), so it gets 3 stages to keep as many bytes in flight
template <int NVEC> struct LcCfg { static constexpr int STAGES = NVEC == 2 ? 3 : 2; };
constexpr int LC_G = 2;           // ciphertext blocks per stage
constexpr int LC_TILE_BYTES = RT_TILE_U64 * 8;  // 5632
template <int NVEC> constexpr int lc_smem_bytes() { return LcCfg<NVEC>::STAGES * LC_G * LC_TILE_BYTES; }
constexpr int LC_CTR_STRIDE = 32;  // one queue counter per 128-byte line

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// grid (23 tiles, nslots).  queue[tile * LC_CTR_STRIDE] = next chunk of that tile (zero on entry; reset by finish).
// partial[slot] is row-planar: limb row j, coordinate c at j*1472 + c.
// NVEC scalar vectors share one pass over the ciphertexts (coeffs[v], partial + v * partial_stride): the prover pairs
// (v_w, h) over the s region and (hat_v, hat_h) over the as region, halving its HBM traffic.
template <int NVEC>
__global__ void __launch_bounds__(LC_TILE)
k_lincomb(const uint64_t *__restrict__ cts, const uint32_t *__restrict__ coeffs0, const uint32_t *__restrict__ coeffs1,
          size_t d, uint32_t chunk_len, unsigned int *__restrict__ queue, uint64_t *__restrict__ partial,
          size_t partial_stride) {
  constexpr int LC_STAGES = LcCfg<NVEC>::STAGES;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * LC_STAGES];  // [0, S): full (tx), [S, 2S): empty (2 warps)
  __shared__ uint32_t meta_first[LC_STAGES];             // first ciphertext of the stage (d < 2^32 per GPU)
  __shared__ uint32_t meta_n[LC_STAGES];                 // ciphertexts in the stage; 0 = queue drained
  const int tile = blockIdx.x;
  const uint32_t nchunks = (uint32_t)((d + chunk_len - 1) / chunk_len);
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t bbase = (uint32_t)__cvta_generic_to_shared(bars);
  if (threadIdx.x == 0) {
    for (int s = 0; s < LC_STAGES; s++) {
      mbar_init(bbase + 8 * s, 1);
      mbar_init(bbase + 8 * (LC_STAGES + s), LC_TILE / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // producer state (thread 0)
  size_t cur = 0, end = 0;
  bool drained = false;
  auto issue = [&](int s) {
    if (cur == end && !drained) {
      const uint32_t ch = atomicAdd(queue + tile * LC_CTR_STRIDE, 1u);
      if (ch < nchunks) {
        cur = (size_t)ch * chunk_len;
        end = cur + chunk_len < d ? cur + chunk_len : d;
      } else {
        drained = true;
      }
    }
    if (drained) {  // end marker: consumers see n == 0
      meta_n[s] = 0;
      mbar_arrive(bbase + 8 * s);
      return;
    }
    const int n = (int)(end - cur < (size_t)LC_G ? end - cur : (size_t)LC_G);
    meta_first[s] = (uint32_t)cur;
    meta_n[s] = n;
    mbar_expect_tx(bbase + 8 * s, n * LC_TILE_BYTES);
    for (int g = 0; g < n; g++)
      bulk_g2s(sbase + (s * LC_G + g) * LC_TILE_BYTES, cts + (cur + g) * PLANAR_U64 + (size_t)tile * RT_TILE_U64,
               LC_TILE_BYTES, bbase + 8 * s);
    cur += n;
  };
  if (threadIdx.x == 0)
    for (int s = 0; s < LC_STAGES; s++) issue(s);

  Acc704 acc[NVEC];
#pragma unroll
  for (int v = 0; v < NVEC; v++) acc_zero(acc[v]);
  for (uint32_t it = 0;; it++) {
    const int s = (int)(it % LC_STAGES);
    const uint32_t ph = (it / LC_STAGES) & 1;
    mbar_wait(bbase + 8 * s, ph);
    const int n = (int)meta_n[s];
    if (n == 0) break;
    const size_t first = meta_first[s];
    for (int g = 0; g < n; g++) {
      const uint64_t *sp = reinterpret_cast<const uint64_t *>(smem + (s * LC_G + g) * LC_TILE_BYTES) + threadIdx.x;
      uint32_t a[22];
#pragma unroll
      for (int j = 0; j < L64; j++) {
        const uint64_t v = sp[j * RT_TILE];
        a[2 * j] = (uint32_t)v;
        a[2 * j + 1] = (uint32_t)(v >> 32);
      }
      acc_mad(acc[0], a, __ldg(coeffs0 + first + g));
      if constexpr (NVEC == 2) acc_mad(acc[1], a, __ldg(coeffs1 + first + g));
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bbase + 8 * (LC_STAGES + s));  // this warp is done with slot s
    if (threadIdx.x == 0) {
      mbar_wait(bbase + 8 * (LC_STAGES + s), ph);  // every warp released it
      issue(s);
    }
  }

#pragma unroll
  for (int v = 0; v < NVEC; v++) {
    uint32_t r[22];
    acc_fold(acc[v], r);
    uint64_t *out = partial + v * partial_stride + (size_t)blockIdx.y * PLANAR_U64 + tile * LC_TILE + threadIdx.x;
#pragma unroll
    for (int j = 0; j < L64; j++) out[(size_t)j * NCP] = (uint64_t)r[2 * j] | (uint64_t)r[2 * j + 1] << 32;
  }
}

// rop[c] = (rop_in[c] + sum_k partial[k][c]) mod 2^704.  rop is "flat": [1471][11] u64, coordinate-major.
// CTA = 64 coordinates x FIN_SLICES slices of the partial index: each thread adds every FIN_SLICES-th partial
// (loads of successive partials are independent, so several are in flight), the slices meet in shared memory.
constexpr int FIN_SLICES = 8;

// Sum of the row-planar partials for coordinate c; the result is in r for the slice-0 threads.  Contains two
// __syncthreads (the second protects `sm` for reuse by the caller).
__device__ __forceinline__ void fin_sum_partials(const uint64_t *__restrict__ partial, int nparts,
                                                 uint32_t (*sm)[22][64], int c, int cl, int slice, uint32_t (&r)[22]) {
#pragma unroll
  for (int j = 0; j < 22; j++) r[j] = 0;
#pragma unroll 2
  for (int k = slice; k < nparts; k += FIN_SLICES) {
    uint32_t b[22];
#pragma unroll
    for (int j = 0; j < L64; j++) {
      const uint64_t v = partial[(size_t)k * PLANAR_U64 + (size_t)j * NCP + c];
      b[2 * j] = (uint32_t)v;
      b[2 * j + 1] = (uint32_t)(v >> 32);
    }
    add704(r, b);
  }
  if (slice > 0) {
#pragma unroll
    for (int j = 0; j < 22; j++) sm[slice - 1][j][cl] = r[j];
  }
  __syncthreads();
  if (slice == 0) {
    for (int s2 = 0; s2 < FIN_SLICES - 1; s2++) {
      uint32_t b[22];
#pragma unroll
      for (int j = 0; j < 22; j++) b[j] = sm[s2][j][cl];
      add704(r, b);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void fin_add_flat(uint32_t (&r)[22], const uint64_t *flat, int c) {
  uint32_t b[22];
#pragma unroll
  for (int j = 0; j < L64; j++) {
    const uint64_t v = flat[(size_t)c * L64 + j];
    b[2 * j] = (uint32_t)v;
    b[2 * j + 1] = (uint32_t)(v >> 32);
  }
  add704(r, b);
}

// One launch finishes up to PEER_LANES accumulators: grid (23 tiles, lanes); lane l adds the partials at
// partial + l * lane_stride and reads / writes its flat ciphertext at rop + l * rop_stride.  (The prover's two
// two-vector passes leave four sets of partials: one finish launch instead of four.)  queue / queue2: chunk queues
// to re-arm (one per pass that fed this finish), nullable.
__global__ void __launch_bounds__(64 * FIN_SLICES)
k_lincomb_finish(const uint64_t *__restrict__ partial, size_t lane_stride, int nparts, const uint64_t *rop_in,
                 uint64_t *rop_out, size_t rop_stride, unsigned int *queue, unsigned int *queue2) {
  __shared__ uint32_t sm[FIN_SLICES - 1][22][64];
  const int lane = blockIdx.y;
  if (threadIdx.x == 0 && lane == 0) {  // one CTA per tile: re-arm its chunk queues
    if (queue) queue[blockIdx.x * LC_CTR_STRIDE] = 0;
    if (queue2) queue2[blockIdx.x * LC_CTR_STRIDE] = 0;
  }
  partial += (size_t)lane * lane_stride;
  if (rop_in) rop_in += (size_t)lane * rop_stride;
  rop_out += (size_t)lane * rop_stride;
  const int cl = threadIdx.x & 63, slice = threadIdx.x >> 6;
  const int c = blockIdx.x * 64 + cl;
  uint32_t r[22];
  fin_sum_partials(partial, nparts, sm, c, cl, slice, r);
  if (slice == 0 && c < NC) {
    if (rop_in) fin_add_flat(r, rop_in, c);
#pragma unroll
    for (int j = 0; j < L64; j++) rop_out[(size_t)c * L64 + j] = (uint64_t)r[2 * j] | (uint64_t)r[2 * j + 1] << 32;
  }
}

// --- multi-GPU: the finish step FUSED with the exchange over NVLink peer memory -----------------
// Every rank owns a "symmetric" buffer that its peers map (CUDA IPC between processes, plain peer access inside
// one); mfb_common.cuh: PEER_*.  For the call with sequence number `epoch` (parity q = epoch & 1), CTA (tile, lane)
//   1. adds this rank's row-planar partials for its 64 coordinates (as k_lincomb_finish does);
//   2. PUSHES the 5632-byte flat tile into slot [q][rank][lane] of EVERY rank's buffer (coalesced 16-byte stores that
//      travel over NVLink / NVSwitch; its own copy is a local store), fences at system scope and then releases
//      flag [q][rank][lane][tile] = epoch in every rank's buffer;
//   3. waits until the flags [q][s][lane][tile] of its OWN buffer carry `epoch` for every source rank s (acquire
//      loads of local memory, written remotely), and
//   4. adds the `world` tiles now sitting in local HBM/L2 (exact mod 2^704: the order does not matter), adds the
//      incoming rop and writes the result — every rank ends with the full sum, with one NVLink store latency
//      on the critical path instead of the four launches of split -> reduce-scatter -> carry -> all-gather.
// Two parities suffice: a rank can only push call e+2 after it has seen every peer's flags of call e+1, which a
// peer raises in the kernel that follows (in stream order) its own reads of call e.  All lanes of one launch share
// the epoch: every rank must pass the same number of lanes.
struct PeerTable {
  uint8_t *base[PEER_MAX];
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ bool peer_wait_u32(const uint32_t *f, uint32_t want, uint64_t timeout_ns) {
  const uint64_t t0 = globaltimer_ns();
  while (ld_acquire_sys_u32(f) != want) {
    if (globaltimer_ns() - t0 > timeout_ns) return false;
    __nanosleep(40);
  }
  return true;
}

__global__ void __launch_bounds__(64 * FIN_SLICES)
k_lincomb_finish_peer(const uint64_t *__restrict__ partial, size_t lane_stride, int nparts, const uint64_t *rop_in,
                      uint64_t *rop_out, size_t rop_stride, unsigned int *queue, unsigned int *queue2,
                      const __grid_constant__ PeerTable peers, int world, int rank, uint32_t epoch, uint64_t timeout_ns,
                      int *status, int push_only) {
  __shared__ uint32_t sm[FIN_SLICES - 1][22][64];
  __shared__ __align__(16) uint64_t stage[RT_TILE * L64];  // the flat tile: [64][11] u64
  const int tile = blockIdx.x, lane = blockIdx.y;
  if (threadIdx.x == 0 && lane == 0) {
    if (queue) queue[tile * LC_CTR_STRIDE] = 0;
    if (queue2) queue2[tile * LC_CTR_STRIDE] = 0;
  }
  partial += (size_t)lane * lane_stride;
  if (rop_in) rop_in += (size_t)lane * rop_stride;
  rop_out += (size_t)lane * rop_stride;
  const int cl = threadIdx.x & 63, slice = threadIdx.x >> 6;
  const int c = tile * 64 + cl;
  const uint32_t q = epoch & 1u;
  uint32_t r[22];
  fin_sum_partials(partial, nparts, sm, c, cl, slice, r);
  if (slice == 0) {
#pragma unroll
    for (int j = 0; j < L64; j++) stage[cl * L64 + j] = c < NC ? ((uint64_t)r[2 * j] | (uint64_t)r[2 * j + 1] << 32) : 0;
  }
  __syncthreads();
  // 2. push: 352 threads x one uint4 per destination rank
  constexpr int TILE_FLAT_BYTES = RT_TILE * L64 * 8;  // 5632
  if (threadIdx.x < TILE_FLAT_BYTES / 16) {
    const uint4 v = reinterpret_cast<const uint4 *>(stage)[threadIdx.x];
    for (int i = 0; i < world; i++) {
      const int p = (rank + 1 + i) % world;  // remote destinations first, the local copy last
      uint8_t *dst = peers.base[p] + peer_slot_offset(q, world, rank, lane) + (size_t)tile * TILE_FLAT_BYTES + 16 * threadIdx.x;
      *reinterpret_cast<uint4 *>(dst) = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world)
    st_release_sys_u32(reinterpret_cast<uint32_t *>(peers.base[threadIdx.x] + peer_flag_offset(q, world, rank, lane, tile)), epoch);
  if (push_only) return;  // steps 3 and 4 are a separate, small kernel (k_peer_allreduce, wait_only): this one never blocks
  // 3. wait for every source rank's tile
  if ((int)threadIdx.x < world) {
    const uint32_t *f = reinterpret_cast<const uint32_t *>(peers.base[rank] + peer_flag_offset(q, world, (int)threadIdx.x, lane, tile));
    // a peer that never arrives is reported instead of hanging the GPU
    if (!peer_wait_u32(f, epoch, timeout_ns)) *reinterpret_cast<volatile int *>(status) = 1 + (int)threadIdx.x;
  }
  __syncthreads();
  // 4. add the world tiles (L2 loads: the lines were written by remote stores, never cached in this SM's L1)
#pragma unroll
  for (int j = 0; j < 22; j++) r[j] = 0;
  for (int src = slice; src < world; src += FIN_SLICES) {
    const uint64_t *t = reinterpret_cast<const uint64_t *>(peers.base[rank] + peer_slot_offset(q, world, src, lane)) + (size_t)c * L64;
    uint32_t b[22];
#pragma unroll
    for (int j = 0; j < L64; j++) {
      const uint64_t v = __ldcg(t + j);
      b[2 * j] = (uint32_t)v;
      b[2 * j + 1] = (uint32_t)(v >> 32);
    }
    add704(r, b);
  }
  if (slice > 0) {
#pragma unroll
    for (int j = 0; j < 22; j++) sm[slice - 1][j][cl] = r[j];
  }
  __syncthreads();
  if (slice == 0 && c < NC) {
    const int nsl = world < FIN_SLICES ? world : FIN_SLICES;
    for (int s2 = 0; s2 < nsl - 1; s2++) {
      uint32_t b[22];
#pragma unroll
      for (int j = 0; j < 22; j++) b[j] = sm[s2][j][cl];
      add704(r, b);
    }
    if (rop_in) fin_add_flat(r, rop_in, c);
#pragma unroll
    for (int j = 0; j < L64; j++) rop_out[(size_t)c * L64 + j] = (uint64_t)r[2 * j] | (uint64_t)r[2 * j + 1] << 32;
  }
}

// The exchange alone (mfb_peer_allreduce_dev): this rank's contribution is already a flat ciphertext (per lane, at
// flat_partial + lane * in_stride).  64 threads per (tile, lane) and no shared memory, so that the kernel is co-resident
// with a running k_lincomb (which fills the SMs' shared memory): on a side stream it overlaps the NEXT lincomb kernel.
__global__ void __launch_bounds__(RT_TILE)
k_peer_allreduce(const uint64_t *__restrict__ flat_partial, size_t in_stride, const uint64_t *rop_in, uint64_t *rop_out,
                 size_t rop_stride, const __grid_constant__ PeerTable peers, int world, int rank, uint32_t epoch,
                 uint64_t timeout_ns, int *status, int wait_only, int push_only) {
  const int tile = blockIdx.x, lane = blockIdx.y;
  flat_partial += (size_t)lane * in_stride;
  if (rop_in) rop_in += (size_t)lane * rop_stride;
  rop_out += (size_t)lane * rop_stride;
  const int c = tile * RT_TILE + threadIdx.x;
  const uint32_t q = epoch & 1u;
  constexpr int TILE_FLAT_U64 = RT_TILE * L64;  // 704 u64 = 352 x 16 bytes
  // push: the tile as 16-byte pieces (the flat input ends at coordinate 1470: the padding coordinate is sent as zeros);
  // wait_only: the tiles were pushed (and the flags raised) by k_lincomb_finish_peer(push_only) of the same epoch
  for (int i = threadIdx.x; i < (wait_only ? 0 : TILE_FLAT_U64 / 2); i += RT_TILE) {
    const size_t e = (size_t)tile * TILE_FLAT_U64 + 2 * (size_t)i;
    uint4 v;
    const uint64_t lo = e < (size_t)NC * L64 ? flat_partial[e] : 0, hi = e + 1 < (size_t)NC * L64 ? flat_partial[e + 1] : 0;
    v.x = (uint32_t)lo;
    v.y = (uint32_t)(lo >> 32);
    v.z = (uint32_t)hi;
    v.w = (uint32_t)(hi >> 32);
    for (int k = 0; k < world; k++) {
      const int p = (rank + 1 + k) % world;
      *reinterpret_cast<uint4 *>(peers.base[p] + peer_slot_offset(q, world, rank, lane) + e * 8) = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world && !wait_only)
    st_release_sys_u32(reinterpret_cast<uint32_t *>(peers.base[threadIdx.x] + peer_flag_offset(q, world, rank, lane, tile)), epoch);
  if (push_only) return;  // the waiting half is a second launch of this kernel (wait_only)
  if ((int)threadIdx.x < world) {
    const uint32_t *f = reinterpret_cast<const uint32_t *>(peers.base[rank] + peer_flag_offset(q, world, (int)threadIdx.x, lane, tile));
    if (!peer_wait_u32(f, epoch, timeout_ns)) *reinterpret_cast<volatile int *>(status) = 1 + (int)threadIdx.x;
  }
  __syncthreads();
  uint32_t r[22];
#pragma unroll
  for (int j = 0; j < 22; j++) r[j] = 0;
  for (int src = 0; src < world; src++) {
    const uint64_t *t = reinterpret_cast<const uint64_t *>(peers.base[rank] + peer_slot_offset(q, world, src, lane)) + (size_t)c * L64;
    uint32_t b[22];
#pragma unroll
    for (int j = 0; j < L64; j++) {
      const uint64_t v = __ldcg(t + j);
      b[2 * j] = (uint32_t)v;
      b[2 * j + 1] = (uint32_t)(v >> 32);
    }
    add704(r, b);
  }
  if (c < NC) {
    if (rop_in) fin_add_flat(r, rop_in, c);
#pragma unroll
    for (int j = 0; j < L64; j++) rop_out[(size_t)c * L64 + j] = (uint64_t)r[2 * j] | (uint64_t)r[2 * j + 1] << 32;
  }
}

// --- multi-GPU exchange helpers --------------------------------------------------
// A rank's canonical partial sum is widened to 22 u64 columns per coordinate (each < 2^32) so that
// an elementwise integer sum over <= 2^32 ranks cannot overflow; after the reduce the owner
// carry-propagates and truncates to 704 bit.
// cols layout: [c][22] u64, c in [0, NCP) (padding coordinate included so the buffer splits evenly).
__global__ void k_columns_split(const uint64_t *__restrict__ flat, uint64_t *__restrict__ cols) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (c, limb)
  if (idx >= NCP * L32) return;
  const int c = idx / L32, l = idx % L32;
  uint64_t v = 0;
  if (c < NC) {
    const uint64_t w = flat[(size_t)c * L64 + (l >> 1)];
    v = (l & 1) ? (w >> 32) : (w & 0xffffffffu);
  }
  cols[idx] = v;
}

// flat[c] = (flat_in[c] + sum_l cols[c][l] << 32l) mod 2^704, for coordinates [c0, c0 + ncoord)
__global__ void k_columns_carry(const uint64_t *__restrict__ cols, int c0, int ncoord,
                                const uint64_t *flat_in, uint64_t *flat_out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncoord) return;
  const int c = c0 + t;
  if (c >= NC) return;
  uint32_t r[22];
  uint64_t carry = 0;
#pragma unroll
  for (int l = 0; l < L32; l++) {
    const uint64_t s = cols[(size_t)t * L32 + l] + carry;  // cols < 2^32 * ranks, carry < 2^33: no overflow
    r[l] = (uint32_t)s;
    carry = s >> 32;
  }
  if (flat_in) {
    uint32_t b[22];
#pragma unroll
    for (int j = 0; j < L64; j++) {
      const uint64_t v = flat_in[(size_t)c * L64 + j];
      b[2 * j] = (uint32_t)v;
      b[2 * j + 1] = (uint32_t)(v >> 32);
    }
    add704(r, b);
  }
#pragma unroll
  for (int j = 0; j < L64; j++) flat_out[(size_t)c * L64 + j] = (uint64_t)r[2 * j] | (uint64_t)r[2 * j + 1] << 32;
}

// ---------------------------------------------------------------------------------------------
// number of CTA slots per tile such that every CTA of the grid is resident at once
template <int NVEC>
static int nslots_for(size_t d, int sm_count) {
  static int occ = 0;
  if (!occ) {
    cudaFuncSetAttribute(k_lincomb<NVEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, lc_smem_bytes<NVEC>());
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_lincomb<NVEC>, LC_TILE, lc_smem_bytes<NVEC>()) != cudaSuccess || occ < 1)
      occ = 6;
  }
  size_t n = (size_t)sm_count * occ / RT_NTILES;
  const size_t chunks = (d + 3) / 4;
  if (n > chunks) n = chunks;
  return (int)(n ? n : 1);
}
int lincomb_nslots(size_t d, int sm_count, int nvec) {
  return nvec == 2 ? nslots_for<2>(d, sm_count) : nslots_for<1>(d, sm_count);
}

// the opt-in to > 48 KB of dynamic shared memory is per device and per kernel: done once, not on every launch (a launch
// sequence over 8 GPUs from one host thread is latency-critical)
template <int NVEC>
static cudaError_t lincomb_smem_optin() {
  static bool done[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(k_lincomb<NVEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, lc_smem_bytes<NVEC>());
  if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
  return e;
}

typedef void (*mark_fn)(void *, int, cudaStream_t);
// launches the main kernel only; *nslots_inout returns the number of partial sums written per vector.
// coeffs1 == nullptr: one scalar vector; else two (partials of vector 1 start at partial_ws + nslots * PLANAR_U64).
cudaError_t launch_lincomb_partials(const uint64_t *cts, const uint32_t *coeffs0, const uint32_t *coeffs1, size_t d,
                                    uint64_t *partial_ws, unsigned int *queue, int *nslots_inout, cudaStream_t st,
                                    mark_fn mark, void *mark_arg) {
  int nslots = d ? *nslots_inout : 0;
  if (nslots > 0) {
    if (d >> 32) return cudaErrorInvalidValue;  // stage metadata holds 32-bit ciphertext indices
    const uint32_t chunk_len = 4;  // ciphertexts per queue entry = one full ring (2 stages x 2 blocks)
    dim3 grid(RT_NTILES, nslots);
    cudaError_t e;
    if (coeffs1) {
      if ((e = lincomb_smem_optin<2>()) != cudaSuccess) return e;
      if (mark) mark(mark_arg, 0, st);
      k_lincomb<2><<<grid, LC_TILE, lc_smem_bytes<2>(), st>>>(cts, coeffs0, coeffs1, d, chunk_len, queue, partial_ws,
                                                          (size_t)nslots * PLANAR_U64);
    } else {
      if ((e = lincomb_smem_optin<1>()) != cudaSuccess) return e;
      if (mark) mark(mark_arg, 0, st);
      k_lincomb<1><<<grid, LC_TILE, lc_smem_bytes<1>(), st>>>(cts, coeffs0, nullptr, d, chunk_len, queue, partial_ws, 0);
    }
    if (mark) mark(mark_arg, 1, st);
  }
  *nslots_inout = nslots;
  return cudaGetLastError();
}

// finish `lanes` accumulators in one launch: lane l = partials at partial_ws + l * lane_stride (u64), flat ciphertext at
// rop + l * rop_stride (u64)
cudaError_t launch_lincomb_finish(const uint64_t *partial_ws, size_t lane_stride, int lanes, int nparts, const uint64_t *rop_in,
                                  uint64_t *rop_out, size_t rop_stride, unsigned int *queue, unsigned int *queue2,
                                  cudaStream_t st) {
  if (lanes < 1 || lanes > PEER_LANES) return cudaErrorInvalidValue;
  k_lincomb_finish<<<dim3(RT_NTILES, lanes), 64 * FIN_SLICES, 0, st>>>(partial_ws, lane_stride, nparts, rop_in, rop_out, rop_stride,
                                                                     queue, queue2);
  return cudaGetLastError();
}

// mode 0: flat_partial != nullptr: the exchange alone (k_peer_allreduce; lane l contributes flat_partial + l * lane_stride),
//         else the finish of the row-planar partials fused with the exchange (one kernel that pushes, waits and adds);
// mode 1: push only (never blocks): the finish of the row-planar partials + push, or (flat_partial) the push of flat
//         ciphertexts;  mode 2: wait + add only (a small kernel: 64 threads per tile and lane) — the two halves of one
//         exchange (same epoch) when the ranks' kernels cannot be assumed to run concurrently (members of a device set
//         sharing a GPU; a profiler or debugger that serialises launches)
cudaError_t launch_lincomb_finish_peer(const uint64_t *partial_ws, size_t lane_stride, int lanes, int nparts,
                                       const uint64_t *flat_partial, const uint64_t *rop_in, uint64_t *rop_out, size_t rop_stride,
                                       unsigned int *queue, unsigned int *queue2, uint8_t *const *bases, int world, int rank,
                                       uint32_t epoch, uint64_t timeout_ns, int *status, int mode, cudaStream_t st) {
  if (lanes < 1 || lanes > PEER_LANES) return cudaErrorInvalidValue;
  PeerTable t = {};
  for (int i = 0; i < world; i++) t.base[i] = bases[i];
  if (mode == 2)
    k_peer_allreduce<<<dim3(RT_NTILES, lanes), RT_TILE, 0, st>>>(rop_out, 0, rop_in, rop_out, rop_stride, t, world, rank, epoch, timeout_ns,
                                                                status, 1, 0);
  else if (flat_partial)
    k_peer_allreduce<<<dim3(RT_NTILES, lanes), RT_TILE, 0, st>>>(flat_partial, lane_stride, rop_in, rop_out, rop_stride, t, world, rank,
                                                                epoch, timeout_ns, status, 0, mode == 1);
  else
    k_lincomb_finish_peer<<<dim3(RT_NTILES, lanes), 64 * FIN_SLICES, 0, st>>>(partial_ws, lane_stride, nparts, rop_in, rop_out, rop_stride,
                                                                            queue, queue2, t, world, rank, epoch, timeout_ns, status,
                                                                            mode == 1);
  return cudaGetLastError();
}

cudaError_t probe_kernel_image() {
  cudaFuncAttributes at;
  return cudaFuncGetAttributes(&at, k_columns_split);  // cudaErrorInvalidDeviceFunction when no sm_100a image loads
}

cudaError_t launch_columns_split(const uint64_t *flat, uint64_t *cols, cudaStream_t st) {
  k_columns_split<<<(NCP * L32 + 255) / 256, 256, 0, st>>>(flat, cols);
  return cudaGetLastError();
}
cudaError_t launch_columns_carry(const uint64_t *cols, int c0, int ncoord, const uint64_t *flat_in,
                                 uint64_t *flat_out, cudaStream_t st) {
  k_columns_carry<<<(ncoord + 127) / 128, 128, 0, st>>>(cols, c0, ncoord, flat_in, flat_out);
  return cudaGetLastError();
}

}  // namespace mfb
