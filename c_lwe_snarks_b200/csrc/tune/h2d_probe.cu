// Development probe (not part of the library): how fast can a multi-GB PAGEABLE host buffer reach the device?
//   a) cudaMemcpy from pageable memory (the driver's own staging)
//   b) memcpy by T threads into two pinned bounce buffers + cudaMemcpyAsync (what h2d_pieces does), T = 1..32
//   c) cudaHostRegister the buffer in place + one cudaMemcpyAsync
// Usage: h2d_probe [GB]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main(int argc, char **argv) {
  const size_t GB = argc > 1 ? atol(argv[1]) : 4;
  const size_t N = GB << 30;
  uint8_t *src = (uint8_t *)malloc(N);
  for (size_t i = 0; i < N; i += 4096) src[i] = (uint8_t)i;  // touch every page
  memset(src, 1, N);
  void *dst;
  CK(cudaMalloc(&dst, N));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  printf("host threads: %u\n", std::thread::hardware_concurrency());
  double t0 = now();
  CK(cudaMemcpy(dst, src, N, cudaMemcpyHostToDevice));
  printf("a) cudaMemcpy pageable: %.2f GB/s\n", N / (now() - t0) / 1e9);
  const size_t B = (size_t)64 << 20;
  uint8_t *bb[2];
  cudaEvent_t ev[2];
  for (int k = 0; k < 2; k++) { CK(cudaHostAlloc((void **)&bb[k], B, cudaHostAllocDefault)); CK(cudaEventCreate(&ev[k])); }
  for (unsigned T : {1u, 4u, 8u, 16u, 32u}) {
    t0 = now();
    int k = 0;
    bool used[2] = {false, false};
    for (size_t off = 0; off < N; off += B, k ^= 1) {
      const size_t n = N - off < B ? N - off : B;
      if (used[k]) CK(cudaEventSynchronize(ev[k]));
      std::vector<std::thread> pool;
      const size_t per = (n + T - 1) / T;
      for (unsigned t = 0; t < T; t++) {
        const size_t lo = t * per, hi = lo + per < n ? lo + per : n;
        if (lo < hi) pool.emplace_back([=] { memcpy(bb[k] + lo, src + off + lo, hi - lo); });
      }
      for (auto &th : pool) th.join();
      CK(cudaMemcpyAsync((uint8_t *)dst + off, bb[k], n, cudaMemcpyHostToDevice, st));
      CK(cudaEventRecord(ev[k], st));
      used[k] = true;
    }
    CK(cudaStreamSynchronize(st));
    printf("b) bounce, %2u threads: %.2f GB/s\n", T, N / (now() - t0) / 1e9);
  }
  t0 = now();
  CK(cudaHostRegister(src, N, cudaHostRegisterDefault));
  const double treg = now() - t0;
  t0 = now();
  CK(cudaMemcpyAsync(dst, src, N, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  const double tcp = now() - t0;
  printf("c) cudaHostRegister %.3f s (%.2f GB/s) + copy %.3f s (%.2f GB/s) = %.2f GB/s overall\n", treg, N / treg / 1e9, tcp, N / tcp / 1e9,
         N / (treg + tcp) / 1e9);
  t0 = now();
  CK(cudaHostUnregister(src));
  printf("   unregister %.3f s\n", now() - t0);
  return 0;
}
