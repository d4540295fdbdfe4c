// Development harness: K1 (resident lincomb) variants, timed with CUDA events on synthetic data.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I.. tune_k1.cu -o ../../lib/tune_k1
//   run:   tune_k1 [log2d=16]
// Not part of the product library; variants that win are moved into k_lincomb.cu.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../mfb_common.cuh"

using namespace mfb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
// logical value of limb row j of coordinate c of ciphertext i
__device__ __forceinline__ uint64_t val(size_t i, int j, int c) { return c >= NC ? 0 : mix((i * 11 + j) * 1472 + c + 0x9e3779b97f4a7c15ULL); }

__global__ void fill_planar(uint64_t *p, size_t d) {
  size_t total = d * PLANAR_U64;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
    size_t i = k / PLANAR_U64; int r = (int)(k % PLANAR_U64); p[k] = val(i, r / NCP, r % NCP);
  }
}
// tile-planar: [i][tile 23][row 11][64]
__global__ void fill_tiled(uint64_t *p, size_t d) {
  size_t total = d * PLANAR_U64;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
    size_t i = k / PLANAR_U64; int r = (int)(k % PLANAR_U64); int tile = r / 704, rr = r % 704; p[k] = val(i, rr / 64, tile * 64 + rr % 64);
  }
}
__global__ void fill_h(uint32_t *h, size_t d) {
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < d; k += (size_t)gridDim.x * blockDim.x) h[k] = (uint32_t)mix(k + 77);
}

// ---------------------------------------------------------------------------- V0: LDG planar (the shipped kernel's shape)
template <int UNROLL>
__global__ void __launch_bounds__(64) k_ldg(const uint64_t *__restrict__ cts, const uint32_t *__restrict__ coeffs, size_t d,
                                            size_t chunk_len, uint64_t *__restrict__ partial) {
  __shared__ uint32_t hs[512];
  const int c = blockIdx.x * 64 + threadIdx.x;
  const size_t i0 = (size_t)blockIdx.y * chunk_len;
  const size_t i1 = i0 + chunk_len < d ? i0 + chunk_len : d;
  Acc704 acc; acc_zero(acc);
  for (size_t base = i0; base < i1; base += 512) {
    const int n = (int)(i1 - base < 512 ? i1 - base : 512);
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += 64) hs[k] = coeffs[base + k];
    __syncthreads();
    const uint64_t *p = cts + base * PLANAR_U64 + c;
    int k = 0;
    for (; k + UNROLL <= n; k += UNROLL) {
      uint64_t v[UNROLL][L64];
#pragma unroll
      for (int u = 0; u < UNROLL; u++)
#pragma unroll
        for (int j = 0; j < L64; j++) v[u][j] = __ldcs(p + (size_t)(k + u) * PLANAR_U64 + (size_t)j * NCP);
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        uint32_t a[22];
#pragma unroll
        for (int j = 0; j < L64; j++) { a[2 * j] = (uint32_t)v[u][j]; a[2 * j + 1] = (uint32_t)(v[u][j] >> 32); }
        acc_mad(acc, a, hs[k + u]);
      }
    }
    for (; k < n; k++) {
      uint32_t a[22];
#pragma unroll
      for (int j = 0; j < L64; j++) { const uint64_t v = __ldcs(p + (size_t)k * PLANAR_U64 + (size_t)j * NCP); a[2 * j] = (uint32_t)v; a[2 * j + 1] = (uint32_t)(v >> 32); }
      acc_mad(acc, a, hs[k]);
    }
  }
  uint32_t r[22]; acc_fold(acc, r);
  uint64_t *out = partial + (size_t)blockIdx.y * PLANAR_U64 + c;
#pragma unroll
  for (int j = 0; j < L64; j++) out[(size_t)j * NCP] = (uint64_t)r[2 * j] | (uint64_t)r[2 * j + 1] << 32;
}

// ---------------------------------------------------------------------------- V4: TMA bulk ring, tile-planar layout
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int TILE_BYTES = 11 * 64 * 8;  // 5632
template <int STAGES, int G>  // G ciphertexts per stage
__global__ void __launch_bounds__(64) k_tma(const uint64_t *__restrict__ cts, const uint32_t *__restrict__ coeffs, size_t d,
                                            size_t chunk_len, uint64_t *__restrict__ partial) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * STAGES];
  const int tile = blockIdx.x;
  const size_t i0 = (size_t)blockIdx.y * chunk_len;
  const size_t i1 = i0 + chunk_len < d ? i0 + chunk_len : d;
  const size_t nitems = i1 > i0 ? (i1 - i0 + G - 1) / G : 0;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t bbase = (uint32_t)__cvta_generic_to_shared(bars);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(bbase + 8 * s, 1); mbar_init(bbase + 8 * (STAGES + s), 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](size_t it) {
    const int s = (int)(it % STAGES);
    const size_t first = i0 + it * G;
    const int n = (int)(i1 - first < (size_t)G ? i1 - first : (size_t)G);
    mbar_expect_tx(bbase + 8 * s, n * TILE_BYTES);
    for (int g = 0; g < n; g++)
      bulk_g2s(sbase + (s * G + g) * TILE_BYTES, cts + (first + g) * PLANAR_U64 + (size_t)tile * 704, TILE_BYTES, bbase + 8 * s);
  };
  if (threadIdx.x == 0)
    for (size_t it = 0; it < nitems && it < STAGES; it++) issue(it);
  Acc704 acc; acc_zero(acc);
  for (size_t it = 0; it < nitems; it++) {
    const int s = (int)(it % STAGES);
    const uint32_t ph = (uint32_t)((it / STAGES) & 1);
    mbar_wait(bbase + 8 * s, ph);
    const size_t first = i0 + it * G;
    const int n = (int)(i1 - first < (size_t)G ? i1 - first : (size_t)G);
    for (int g = 0; g < n; g++) {
      const uint64_t *sp = reinterpret_cast<const uint64_t *>(smem + (s * G + g) * TILE_BYTES) + threadIdx.x;
      uint32_t a[22];
#pragma unroll
      for (int j = 0; j < L64; j++) { const uint64_t v = sp[j * 64]; a[2 * j] = (uint32_t)v; a[2 * j + 1] = (uint32_t)(v >> 32); }
      acc_mad(acc, a, __ldg(coeffs + first + g));
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bbase + 8 * (STAGES + s));  // this warp is done with stage s
    if (threadIdx.x == 0 && it + STAGES < nitems) {
      mbar_wait(bbase + 8 * (STAGES + s), ph);  // both warps released the slot
      issue(it + STAGES);
    }
  }
  uint32_t r[22]; acc_fold(acc, r);
  uint64_t *out = partial + (size_t)blockIdx.y * PLANAR_U64 + tile * 64 + threadIdx.x;
#pragma unroll
  for (int j = 0; j < L64; j++) out[(size_t)j * NCP] = (uint64_t)r[2 * j] | (uint64_t)r[2 * j + 1] << 32;
}

// ---------------------------------------------------------------------------- V5: TMA ring + dynamic chunk queue per tile
// grid (23 tiles, nslots).  counters[tile*32] = next chunk of that tile (zeroed before launch).
template <int STAGES, int G>
__global__ void __launch_bounds__(64) k_tma_dyn(const uint64_t *__restrict__ cts, const uint32_t *__restrict__ coeffs, size_t d,
                                                uint32_t chunk_len, unsigned int *__restrict__ counters,
                                                uint64_t *__restrict__ partial) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * STAGES];
  __shared__ uint32_t meta_first[STAGES];  // low 32 bits of the first ciphertext index of the stage (d < 2^32)
  __shared__ uint32_t meta_n[STAGES];
  const int tile = blockIdx.x;
  const uint32_t nchunks = (uint32_t)((d + chunk_len - 1) / chunk_len);
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t bbase = (uint32_t)__cvta_generic_to_shared(bars);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(bbase + 8 * s, 1); mbar_init(bbase + 8 * (STAGES + s), 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // producer state (thread 0 only)
  size_t cur = 0, end = 0;
  bool drained = false;
  auto issue = [&](int s) {  // fill stage s with the next <= G ciphertexts of this tile's queue, or publish the end marker
    if (cur == end && !drained) {
      const uint32_t ch = atomicAdd(counters + tile * 32, 1u);
      if (ch < nchunks) { cur = (size_t)ch * chunk_len; end = cur + chunk_len < d ? cur + chunk_len : d; }
      else drained = true;
    }
    if (drained) { meta_n[s] = 0; mbar_arrive(bbase + 8 * s); return; }
    const int n = (int)(end - cur < (size_t)G ? end - cur : (size_t)G);
    meta_first[s] = (uint32_t)cur; meta_n[s] = n;
    mbar_expect_tx(bbase + 8 * s, n * TILE_BYTES);
    for (int g = 0; g < n; g++)
      bulk_g2s(sbase + (s * G + g) * TILE_BYTES, cts + (cur + g) * PLANAR_U64 + (size_t)tile * 704, TILE_BYTES, bbase + 8 * s);
    cur += n;
  };
  if (threadIdx.x == 0)
    for (int s = 0; s < STAGES; s++) issue(s);
  Acc704 acc; acc_zero(acc);
  for (uint32_t it = 0;; it++) {
    const int s = (int)(it % STAGES);
    const uint32_t ph = (it / STAGES) & 1;
    mbar_wait(bbase + 8 * s, ph);
    const int n = (int)meta_n[s];
    if (n == 0) break;
    const size_t first = meta_first[s];
    for (int g = 0; g < n; g++) {
      const uint64_t *sp = reinterpret_cast<const uint64_t *>(smem + (s * G + g) * TILE_BYTES) + threadIdx.x;
      uint32_t a[22];
#pragma unroll
      for (int j = 0; j < L64; j++) { const uint64_t v = sp[j * 64]; a[2 * j] = (uint32_t)v; a[2 * j + 1] = (uint32_t)(v >> 32); }
      acc_mad(acc, a, __ldg(coeffs + first + g));
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bbase + 8 * (STAGES + s));
    if (threadIdx.x == 0) {
      mbar_wait(bbase + 8 * (STAGES + s), ph);
      issue(s);
    }
  }
  uint32_t r[22]; acc_fold(acc, r);
  uint64_t *out = partial + (size_t)blockIdx.y * PLANAR_U64 + tile * 64 + threadIdx.x;
#pragma unroll
  for (int j = 0; j < L64; j++) out[(size_t)j * NCP] = (uint64_t)r[2 * j] | (uint64_t)r[2 * j + 1] << 32;
}

// ---------------------------------------------------------------------------- read-only bandwidth reference
__global__ void k_readsum(const uint4 *__restrict__ p, size_t n16, unsigned long long *out) {
  uint64_t s = 0;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {
    uint4 a = __ldcs(p + i), b = __ldcs(p + i + stride), c = __ldcs(p + i + 2 * stride), e = __ldcs(p + i + 3 * stride);
    s += a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ e.x ^ e.y ^ e.z ^ e.w;
  }
  for (; i < n16; i += stride) { uint4 a = __ldcs(p + i); s += a.x ^ a.y ^ a.z ^ a.w; }
  if (s == 0x1234567) atomicAdd(out, s);
}

__global__ void k_hash(const uint64_t *partial, int nparts, unsigned long long *out) {
  // order-independent digest of sum of partials' low limbs (sanity only; exact finish is the library's job)
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= NC) return;
  uint64_t s = 0;
  for (int k = 0; k < nparts; k++) s += partial[(size_t)k * PLANAR_U64 + c];  // row 0 = low 64 bits: exact mod 2^64
  atomicAdd(out, mix(s + c));
}

template <class F>
static float time_it(F f, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; r++) f();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGetLastError());
  return ms / reps;
}

int main(int argc, char **argv) {
  const int log2d = argc > 1 ? atoi(argv[1]) : 16;
  const size_t d = (size_t)1 << log2d;
  const double algo = (double)d * 129448.0;
  uint64_t *planar, *tiled, *partial; uint32_t *h; unsigned long long *dig;
  CK(cudaMalloc(&planar, d * PLANAR_U64 * 8)); CK(cudaMalloc(&tiled, d * PLANAR_U64 * 8));
  CK(cudaMalloc(&partial, 512 * PLANAR_U64 * 8)); CK(cudaMalloc(&h, d * 4)); CK(cudaMalloc(&dig, 8));
  fill_planar<<<148 * 8, 256>>>(planar, d); fill_tiled<<<148 * 8, 256>>>(tiled, d); fill_h<<<148, 256>>>(h, d);
  CK(cudaDeviceSynchronize());
  auto digest = [&](int nparts) { unsigned long long z = 0; CK(cudaMemcpy(dig, &z, 8, cudaMemcpyHostToDevice)); k_hash<<<(NC + 127) / 128, 128>>>(partial, nparts, dig); CK(cudaMemcpy(&z, dig, 8, cudaMemcpyDeviceToHost)); return z; };
  auto report = [&](const char *name, float ms, int nparts) { printf("%-34s %8.4f ms  %8.1f GB/s  frac %.3f  digest %016llx\n", name, ms, algo / ms / 1e6, algo / ms / 1e6 / 6539.9, digest(nparts)); fflush(stdout); };

  { float ms = time_it([&] { k_readsum<<<148 * 16, 256>>>((const uint4 *)planar, d * PLANAR_U64 / 2, dig); }, 10);
    printf("%-34s %8.4f ms  %8.1f GB/s (raw bytes)\n", "read-only LDG.128 sum", ms, d * PLANAR_U64 * 8.0 / ms / 1e6); }

  for (int nch : {65, 77}) {
    const size_t cl = (d + nch - 1) / nch; dim3 g(23, (unsigned)((d + cl - 1) / cl));
    char nm[64];
    snprintf(nm, 64, "ldg U1 nchunks=%d", nch); report(nm, time_it([&] { k_ldg<1><<<g, 64>>>(planar, h, d, cl, partial); }, 10), g.y);
    snprintf(nm, 64, "ldg U2 nchunks=%d", nch); report(nm, time_it([&] { k_ldg<2><<<g, 64>>>(planar, h, d, cl, partial); }, 10), g.y);
    snprintf(nm, 64, "ldg U4 nchunks=%d", nch); report(nm, time_it([&] { k_ldg<4><<<g, 64>>>(planar, h, d, cl, partial); }, 10), g.y);
  }
#define SWEEP_TMA(ST, GG)                                                                                 \
  {                                                                                                       \
    const int sm = ST * GG * TILE_BYTES;                                                                  \
    CK(cudaFuncSetAttribute(k_tma<ST, GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));            \
    float best = 1e9; int bestn = 0;                                                                      \
    for (int nch : {13, 16, 19, 22, 26, 29, 32, 35, 38, 42, 45, 48, 51, 58, 64, 71, 77, 84, 90, 96, 103, 116, 129, 154, 180, 206}) { \
      const size_t cl = (d + nch - 1) / nch; dim3 g(23, (unsigned)((d + cl - 1) / cl));                      \
      float ms = time_it([&] { k_tma<ST, GG><<<g, 64, sm>>>(tiled, h, d, cl, partial); }, 8);             \
      if (ms < best) { best = ms; bestn = nch; }                                                          \
      if (verbose) printf("   tma S%d G%d nch %3d: %.4f ms\n", ST, GG, nch, ms);                        \
    }                                                                                                     \
    printf("tma S%d G%d (%3d KB/CTA): best %.4f ms at nchunks=%d  -> %.1f GB/s frac %.3f\n", ST, GG, sm / 1024, best, bestn, \
           algo / best / 1e6, algo / best / 1e6 / 6539.9); fflush(stdout);                                \
  }
  const bool verbose = argc > 2;
  SWEEP_TMA(2, 4) SWEEP_TMA(2, 8) SWEEP_TMA(3, 1)
  unsigned int *counters; CK(cudaMalloc(&counters, 23 * 32 * 4));
#define SWEEP_DYN(ST, GG)                                                                                 \
  {                                                                                                       \
    const int sm = ST * GG * TILE_BYTES;                                                                  \
    CK(cudaFuncSetAttribute(k_tma_dyn<ST, GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));        \
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_tma_dyn<ST, GG>, 64, sm));      \
    for (uint32_t cl : {GG * 2u, GG * 4u, GG * 16u}) {                                                    \
      float best = 1e9; int bestn = 0;                                                                    \
      for (int nslots : {148 * occ / 23, 148 * occ / 23 + 1, 148 * occ / 46, 148 * occ / 23 * 2}) {         \
        if (nslots < 1 || nslots > 500) continue;                                                         \
        dim3 g(23, nslots);                                                                               \
        float ms = time_it([&] { cudaMemsetAsync(counters, 0, 23 * 32 * 4); k_tma_dyn<ST, GG><<<g, 64, sm>>>(tiled, h, d, cl, counters, partial); }, 8); \
        if (ms < best) { best = ms; bestn = nslots; }                                                     \
        if (verbose) printf("   dyn S%d G%d cl %u nslots %3d: %.4f ms  digest %016llx\n", ST, GG, cl, nslots, ms, digest(nslots)); \
      }                                                                                                   \
      printf("dyn S%d G%d occ %2d chunk %3u: best %.4f ms at nslots=%d -> %.1f GB/s frac %.3f\n", ST, GG, occ, cl, best, bestn, \
             algo / best / 1e6, algo / best / 1e6 / 6539.9); fflush(stdout);                              \
    }                                                                                                     \
  }
  SWEEP_DYN(2, 4) SWEEP_DYN(3, 4) SWEEP_DYN(2, 8) SWEEP_DYN(3, 2) SWEEP_DYN(4, 2) SWEEP_DYN(3, 1) SWEEP_DYN(4, 1) SWEEP_DYN(6, 1) SWEEP_DYN(2, 2)
  return 0;
}
