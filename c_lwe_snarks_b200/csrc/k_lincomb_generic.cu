// K1 for other LWE parameter points (BASELINE configs[4]: n, log q sweep).
//
// The reference implements exactly one parameter set (n = 1470, log q = 736 with q_eff = 2^704; any other GAMMA_LOGQ
// is `#error "Not implemented"`, lwe.h:119-121), so these points have no reference output: they are checked against
// plain integer arithmetic with q_eff = 2^(64 * L64) and exist to measure lincomb throughput against the HBM roofline
// as the ciphertext shape changes.  Same design as k_lincomb.cu (tile-planar layout, TMA bulk ring, dynamic per-tile
// chunk queues, carry-chain MAC), with the limb count a template parameter and the coordinate count a run-time value:
//   u64 index of (ct i, limb row j, coordinate c) = i*ntiles*64*L64 + (c/64)*64*L64 + j*64 + c%64,  ntiles = ceil((n+1)/64)
#include "mfb_common.cuh"

namespace mfb {

__device__ __forceinline__ void g_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void g_mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void g_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void g_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void g_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ring: GS stages x GG blocks (ciphertext tiles of L * 512 bytes) per stage, ~11 KB per stage as in K1: short limb
// vectors get more blocks per stage so that as many bytes stay in flight
constexpr int GS = 2;
template <int L> struct GgOf { static constexpr int value = 11264 / (L * 512) >= 4 ? 4 : 2; };

// grid (ntiles, nslots); queue[tile*32]; partial[slot][tile][L][64] (same tile-planar shape as one ciphertext)
template <int L>
__global__ void __launch_bounds__(64)
k_lincomb_g(const uint64_t *__restrict__ cts, const uint32_t *__restrict__ coeffs, size_t d, uint32_t chunk_len,
            unsigned int *__restrict__ queue, uint64_t *__restrict__ partial) {
  constexpr int TB = L * 64 * 8;  // tile bytes
  constexpr int GG = GgOf<L>::value;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * GS];
  __shared__ uint32_t meta_first[GS], meta_n[GS];
  const int tile = blockIdx.x;
  const size_t ct_u64 = (size_t)gridDim.x * L * 64;
  const uint32_t nchunks = (uint32_t)((d + chunk_len - 1) / chunk_len);
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem), bbase = (uint32_t)__cvta_generic_to_shared(bars);
  if (threadIdx.x == 0) {
    for (int s = 0; s < GS; s++) { g_mbar_init(bbase + 8 * s, 1); g_mbar_init(bbase + 8 * (GS + s), 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  size_t cur = 0, end = 0;
  bool drained = false;
  auto issue = [&](int s) {
    if (cur == end && !drained) {
      const uint32_t ch = atomicAdd(queue + tile * 32, 1u);
      if (ch < nchunks) { cur = (size_t)ch * chunk_len; end = cur + chunk_len < d ? cur + chunk_len : d; }
      else drained = true;
    }
    if (drained) { meta_n[s] = 0; g_mbar_arrive(bbase + 8 * s); return; }
    const int n = (int)(end - cur < (size_t)GG ? end - cur : (size_t)GG);
    meta_first[s] = (uint32_t)cur; meta_n[s] = n;
    g_mbar_expect_tx(bbase + 8 * s, n * TB);
    for (int g = 0; g < n; g++)
      g_bulk_g2s(sbase + (s * GG + g) * TB, cts + (cur + g) * ct_u64 + (size_t)tile * L * 64, TB, bbase + 8 * s);
    cur += n;
  };
  if (threadIdx.x == 0)
    for (int s = 0; s < GS; s++) issue(s);
  AccN<2 * L> acc;
  acc_zero(acc);
  for (uint32_t it = 0;; it++) {
    const int s = (int)(it % GS);
    const uint32_t ph = (it / GS) & 1;
    g_mbar_wait(bbase + 8 * s, ph);
    const int n = (int)meta_n[s];
    if (n == 0) break;
    const size_t first = meta_first[s];
    for (int g = 0; g < n; g++) {
      const uint64_t *sp = reinterpret_cast<const uint64_t *>(smem + (s * GG + g) * TB) + threadIdx.x;
      uint32_t a[2 * L];
#pragma unroll
      for (int j = 0; j < L; j++) { const uint64_t v = sp[j * 64]; a[2 * j] = (uint32_t)v; a[2 * j + 1] = (uint32_t)(v >> 32); }
      acc_mad(acc, a, __ldg(coeffs + first + g));
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) g_mbar_arrive(bbase + 8 * (GS + s));
    if (threadIdx.x == 0) { g_mbar_wait(bbase + 8 * (GS + s), ph); issue(s); }
  }
  // fold E + (O << 32) and write the partial
  uint32_t r[2 * L];
  acc_fold(acc, r);
  uint64_t *out = partial + (size_t)blockIdx.y * ct_u64 + (size_t)tile * L * 64 + threadIdx.x;
#pragma unroll
  for (int j = 0; j < L; j++) out[j * 64] = (uint64_t)r[2 * j] | (uint64_t)r[2 * j + 1] << 32;
}

// out (tile-planar, one ciphertext shape) = sum of nparts partials mod 2^(64 L); also re-arms the queues.
// CTA = 64 coordinates x GF_SLICES slices of the partial index (independent loads in flight), met in shared memory.
template <int L> struct GfSlices { static constexpr int value = L > 12 ? 4 : 8; };  // (static shared memory <= 48 KB)
template <int L>
__device__ __forceinline__ void addL(uint64_t (&r)[L], const uint64_t (&b)[L]) {
  unsigned long long carry = 0;
#pragma unroll
  for (int j = 0; j < L; j++) {
    const uint64_t s1 = r[j] + b[j];
    const uint64_t c1 = s1 < b[j];
    const uint64_t s2 = s1 + carry;
    const uint64_t c2 = s2 < carry;
    r[j] = s2;
    carry = c1 + c2;
  }
}
template <int L>
__global__ void __launch_bounds__(64 * GfSlices<L>::value) k_finish_g(const uint64_t *__restrict__ partial, int nparts, uint64_t *out, unsigned int *queue) {
  constexpr int GF_SLICES = GfSlices<L>::value;
  __shared__ uint64_t sm[GF_SLICES - 1][L][64];
  const int tile = blockIdx.x, cl = threadIdx.x & 63, slice = threadIdx.x >> 6;
  const size_t ct_u64 = (size_t)gridDim.x * L * 64;
  if (threadIdx.x == 0) queue[tile * 32] = 0;
  uint64_t r[L];
#pragma unroll
  for (int j = 0; j < L; j++) r[j] = 0;
  for (int k = slice; k < nparts; k += GF_SLICES) {
    const uint64_t *p = partial + (size_t)k * ct_u64 + (size_t)tile * L * 64 + cl;
    uint64_t b[L];
#pragma unroll
    for (int j = 0; j < L; j++) b[j] = p[j * 64];
    addL<L>(r, b);
  }
  if (slice > 0) {
#pragma unroll
    for (int j = 0; j < L; j++) sm[slice - 1][j][cl] = r[j];
  }
  __syncthreads();
  if (slice == 0) {
    for (int s2 = 0; s2 < GF_SLICES - 1; s2++) {
      uint64_t b[L];
#pragma unroll
      for (int j = 0; j < L; j++) b[j] = sm[s2][j][cl];
      addL<L>(r, b);
    }
    uint64_t *o = out + (size_t)tile * L * 64 + cl;
#pragma unroll
    for (int j = 0; j < L; j++) o[j * 64] = r[j];
  }
}

template <int L>
static cudaError_t run_generic(const uint64_t *cts, const uint32_t *coeffs, size_t d, int ntiles, uint64_t *out, uint64_t *partial_ws,
                               size_t partial_cap_u64, unsigned int *queue, int sm_count, cudaStream_t st) {
  constexpr int GG = GgOf<L>::value;
  const int smem = GS * GG * L * 64 * 8;
  cudaError_t e = cudaFuncSetAttribute(k_lincomb_g<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_lincomb_g<L>, 64, smem) != cudaSuccess || occ < 1) occ = 4;
  size_t nslots = (size_t)sm_count * occ / ntiles;
  const size_t chunks = (d + GS * GG - 1) / (GS * GG);
  if (nslots > chunks) nslots = chunks;
  const size_t ct_u64 = (size_t)ntiles * L * 64;
  if (nslots * ct_u64 > partial_cap_u64) nslots = partial_cap_u64 / ct_u64;
  if (nslots < 1) nslots = 1;
  k_lincomb_g<L><<<dim3(ntiles, (unsigned)nslots), 64, smem, st>>>(cts, coeffs, d, GS * GG, queue, partial_ws);
  k_finish_g<L><<<ntiles, 64 * GfSlices<L>::value, 0, st>>>(partial_ws, (int)nslots, out, queue);
  return cudaGetLastError();
}

cudaError_t launch_lincomb_generic(int limbs64, const uint64_t *cts, const uint32_t *coeffs, size_t d, int ntiles, uint64_t *out,
                                   uint64_t *partial_ws, size_t partial_cap_u64, unsigned int *queue, int sm_count, cudaStream_t st) {
  switch (limbs64) {
#define MFB_CASE(Lv) case Lv: return run_generic<Lv>(cts, coeffs, d, ntiles, out, partial_ws, partial_cap_u64, queue, sm_count, st);
    MFB_CASE(4) MFB_CASE(6) MFB_CASE(8) MFB_CASE(10) MFB_CASE(11) MFB_CASE(12) MFB_CASE(13) MFB_CASE(14) MFB_CASE(16)
#undef MFB_CASE
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace mfb
