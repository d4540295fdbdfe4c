// Development harness: AES-256-CTR keystream generation variants (shared-memory T-tables), blocks/s.
//   tune_aes            -> each variant fills `nblk`-block tiles into shared memory, ITEMS tiles per CTA, 148 CTAs
#include <cstdio>
#include <cstdlib>
#include "../aes256.cuh"
using namespace mfb;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void sts128(uint32_t addr, const AesState &v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.w0), "r"(v.w1), "r"(v.w2), "r"(v.w3) : "memory");
}

// VARIANT 0: shipped (aes256_ctr_block), 1: plain on AesLut<2>, 2: cached TABS=2, 3: plain TABS=4, 4: cached TABS=4
template <int VARIANT, int NT, int AT, int FM = 0>
__global__ void __launch_bounds__(NT, 1) k_aes(const __grid_constant__ AesKey key, const uint32_t *__restrict__ t0g, int nblk,
                                              int items, unsigned long long *digest, uint32_t m8, uint32_t m16, uint32_t m24,
                                              unsigned long long tex) {
  extern __shared__ __align__(16) uint8_t dyn[];
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(dyn);
  const uint32_t tabA = (s0 + 0xffffu) & ~0xffffu;
  constexpr bool four = (VARIANT == 3 || VARIANT == 4);
  const uint32_t tabB = tabA + 0x10000;
  const uint32_t buf = s0;  // pad region below the tables
  aes_tables_init(dyn + (tabA - s0), t0g, threadIdx.x, NT);
  if (four) aes_tables_init_b(dyn + (tabB - s0), t0g, threadIdx.x, NT);
  __syncthreads();
  AesLut<four ? 4 : 2, FM> L;
  L.lbA = tabA | ((threadIdx.x & 31) << 2);
  L.lbB = tabB | ((threadIdx.x & 31) << 2);
  L.m8 = m8; L.m16 = m16; L.m24 = m24;
  L.tex = tex;
  AesCtrCache cache;
  cache.window = ~0ull;
  uint32_t x = 0;
  for (int it = 0; it < items; it++) {
    const uint64_t first = ((uint64_t)blockIdx.x * items + it) * 8453ull + 12345;
    if constexpr (VARIANT == 5) {  // two blocks per thread in lockstep
      for (int b = threadIdx.x; b < nblk; b += 2 * AT) {
        AesState va, vb;
        if (b + AT < nblk) {
          aes256_ctr_block_cached_x2(L, key, first + b, first + b + AT, cache, va, vb);
          sts128(buf + 16u * b, va);
          sts128(buf + 16u * (b + AT), vb);
        } else {
          sts128(buf + 16u * b, aes256_ctr_block_cached(L, key, first + b, cache));
        }
      }
    } else
    if (threadIdx.x < AT)
    for (int b = threadIdx.x; b < nblk; b += AT) {
      AesState v;
      if constexpr (VARIANT == 0) v = aes256_ctr_block(key, first + b, L.lbA);
      else if constexpr (VARIANT == 1 || VARIANT == 3) v = aes256_ctr_block_plain(L, key, first + b);
      else v = aes256_ctr_block_cached(L, key, first + b, cache);
      sts128(buf + 16u * b, v);
    }
    __syncthreads();
    // consume: fold the buffer (one word per thread per 2 KB) so the stores cannot be dropped
    for (int w = threadIdx.x; w < nblk * 4; w += NT) x ^= lds32(buf + 4 * w) * (uint32_t)(w + 1 + it);
    __syncthreads();
  }
  atomicXor(digest, (unsigned long long)x * 0x9e3779b97f4a7c15ull + blockIdx.x);
}

static unsigned long long g_tex = 0;
template <int VARIANT, int NT, int AT = NT, int FM = 0>
static void run(const char *name, const AesKey &key, const uint32_t *t0, int nblk, int items, unsigned long long *dig) {
  const int smem = (VARIANT == 3 || VARIANT == 4) ? 0x30000 : 0x20000;
  CK(cudaFuncSetAttribute(k_aes<VARIANT, NT, AT, FM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  unsigned long long z = 0, out = 0;
  k_aes<VARIANT, NT, AT, FM><<<148, NT, smem>>>(key, t0, nblk, items, dig, 1u << 8, 1u << 16, 1u << 24, g_tex);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(dig, &z, 8, cudaMemcpyHostToDevice));
  CK(cudaEventRecord(e0));
  k_aes<VARIANT, NT, AT, FM><<<148, NT, smem>>>(key, t0, nblk, items, dig, 1u << 8, 1u << 16, 1u << 24, g_tex);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaMemcpy(&out, dig, 8, cudaMemcpyDeviceToHost));
  const double blocks = 148.0 * items * nblk;
  printf("%-26s FM=%2d NT=%4d AT=%4d nblk=%4d: %8.3f ms  %.3e blocks/s  (%.2f cyc/blk/SM @1.965GHz)  digest %016llx\n", name, FM, NT, AT, nblk, ms,
         blocks / ms * 1e3, 1.965e9 * 148 / (blocks / ms * 1e3), out);
  fflush(stdout);
}

int main() {
  uint8_t seed[40]; for (int i = 0; i < 40; i++) seed[i] = (uint8_t)i;
  AesKey key; aes_host::expand(seed, &key);
  uint32_t t0h[256]; aes_host::t0_table(t0h);
  uint32_t *t0; unsigned long long *dig;
  CK(cudaMalloc(&t0, 1024)); CK(cudaMemcpy(t0, t0h, 1024, cudaMemcpyHostToDevice)); CK(cudaMalloc(&dig, 8));
  {
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = t0;
    rd.res.linear.desc = cudaCreateChannelDesc<unsigned int>();
    rd.res.linear.sizeInBytes = 1024;
    cudaTextureDesc td = {};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    g_tex = (unsigned long long)tex;
  }
  const int items = 600;
  run<2, 512>("ctr-cached 2-table", key, t0, 2818, items, dig);   // shipped
  run<2, 512, 512, 0x100>("2-table, 1 lookup/round via TEX", key, t0, 2818, items, dig);
  run<2, 512, 512, 0x200>("2-table, 2 lookups/round via TEX", key, t0, 2818, items, dig);
  run<2, 512, 512, 0x300>("2-table, 3 lookups/round via TEX", key, t0, 2818, items, dig);
  run<2, 512, 512, 0x400>("2-table, 4 lookups/round via TEX", key, t0, 2818, items, dig);
  run<2, 512>("ctr-cached 2-table", key, t0, 3072, items, dig);
  run<2, 512, 512, 0x200>("2-table, 2 lookups/round via TEX", key, t0, 3072, items, dig);
  return 0;
}
