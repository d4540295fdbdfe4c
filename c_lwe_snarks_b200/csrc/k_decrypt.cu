// K4 — batched Regev decryption (regev_decrypt lwe.c:105-111, mpz_dotp lwe.h:57-61) and the verifier's
// extra "test-error" dot product (snark.c:238).
//
// For ciphertext k:  dot = (sum_j ct[k][j] * sk[j]) mod 2^704   (mpz_add_dotp lwe.c:20-28 + modq)
//                    m   = (b - dot) floor-mod p                (b may be negative after ct_smudge: its sign is
//                                                                passed beside the magnitude)
// One CTA per ciphertext: 0.65 MB of input for a whole proof, latency-bound.  Ciphertexts arrive "flat"
// ([1471][11] u64, coordinate-major, as the proof holds them), sk planar.
#include "mfb_common.cuh"

namespace mfb {

constexpr int DEC_THREADS = 512;

__global__ void __launch_bounds__(DEC_THREADS)
k_decrypt(const uint64_t *__restrict__ sk, const uint64_t *__restrict__ cts_flat, const uint8_t *__restrict__ b_neg,
          size_t count, uint64_t *__restrict__ out_m, uint64_t *__restrict__ out_dot) {
  __shared__ uint32_t red[DEC_THREADS / 32][44];
  __shared__ unsigned long long cols[22];
  const size_t k = blockIdx.x;
  if (k >= count) return;
  const uint64_t *ct = cts_flat + k * (size_t)NC * L64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  Acc704 acc;
  acc_zero(acc);
  for (int c = threadIdx.x; c < N; c += DEC_THREADS) {
    uint32_t a[22], b[22];
#pragma unroll
    for (int j = 0; j < L64; j++) {
      const uint64_t v = ct[(size_t)c * L64 + j];
      a[2 * j] = (uint32_t)v;
      a[2 * j + 1] = (uint32_t)(v >> 32);
      const uint64_t w = __ldg(sk + (size_t)j * NCP + c);
      b[2 * j] = (uint32_t)w;
      b[2 * j + 1] = (uint32_t)(w >> 32);
    }
    acc_mul(acc, a, b);
  }
  uint32_t r[22];
  acc_fold(acc, r);
#pragma unroll
  for (int l = 0; l < 22; l++) {
    const uint32_t lo = __reduce_add_sync(0xffffffffu, r[l] & 0xffffu);
    const uint32_t hi = __reduce_add_sync(0xffffffffu, r[l] >> 16);
    if (lane == 0) {
      red[warp][2 * l] = lo;
      red[warp][2 * l + 1] = hi;
    }
  }
  __syncthreads();
  if (threadIdx.x < 22) {
    unsigned long long lo = 0, hi = 0;
    for (int w = 0; w < DEC_THREADS / 32; w++) {
      lo += red[w][2 * threadIdx.x];
      hi += red[w][2 * threadIdx.x + 1];
    }
    cols[threadIdx.x] = lo + (hi << 16);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t dot[22], bb[22];
    uint64_t carry = 0;
    for (int l = 0; l < 22; l++) {
      const uint64_t t = cols[l] + carry;
      dot[l] = (uint32_t)t;
      carry = t >> 32;
    }
    for (int j = 0; j < L64; j++) {
      const uint64_t v = ct[(size_t)N * L64 + j];
      bb[2 * j] = (uint32_t)v;
      bb[2 * j + 1] = (uint32_t)(v >> 32);
      if (out_dot) out_dot[k * L64 + j] = (uint64_t)dot[2 * j] | (uint64_t)dot[2 * j + 1] << 32;
    }
    // t = b - dot as sign + magnitude (23 limbs: |b| + dot can reach 2^704)
    uint32_t mag[23];
    bool neg;
    if (b_neg && b_neg[k]) {  // -( |b| + dot )
      uint64_t c = 0;
      for (int l = 0; l < 22; l++) {
        c += (uint64_t)bb[l] + dot[l];
        mag[l] = (uint32_t)c;
        c >>= 32;
      }
      mag[22] = (uint32_t)c;
      neg = true;
    } else {
      uint32_t d1[22];
      const uint32_t borrow = sub704(d1, bb, dot);
      if (borrow) sub704(d1, dot, bb);
      for (int l = 0; l < 22; l++) mag[l] = d1[l];
      mag[22] = 0;
      neg = borrow != 0;
    }
    uint64_t rem = 0;
    for (int l = 22; l >= 0; l--) rem = ((rem << 32) | mag[l]) % P;
    out_m[k] = (neg && rem) ? P - rem : rem;  // mpz_mod_ui: non-negative residue
  }
}

cudaError_t launch_decrypt(const uint64_t *sk, const uint64_t *cts_flat, const uint8_t *b_neg, size_t count,
                           uint64_t *out_m, uint64_t *out_dot, cudaStream_t st) {
  if (count == 0) return cudaSuccess;
  k_decrypt<<<(unsigned)count, DEC_THREADS, 0, st>>>(sk, cts_flat, b_neg, count, out_m, out_dot);
  return cudaGetLastError();
}

// flat [count][n][11] u64 (coordinate-major) -> row-planar [count][11][1472] (tiled = 0: secret keys, n = 1470)
// or the resident tile-planar layout (tiled = 1: host-supplied ciphertexts, n = 1471); missing coordinates are 0
__global__ void k_flat_to_planar(const uint64_t *__restrict__ flat, int n, size_t count, uint64_t *__restrict__ out,
                                 int tiled) {
  const size_t total = count * PLANAR_U64;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t k = idx / PLANAR_U64;
    const int rem = (int)(idx % PLANAR_U64);
    int j, c;
    if (tiled) {
      const int tile = rem / RT_TILE_U64, rr = rem % RT_TILE_U64;
      j = rr / RT_TILE;
      c = tile * RT_TILE + rr % RT_TILE;
    } else {
      j = rem / NCP;
      c = rem % NCP;
    }
    out[idx] = c < n ? flat[(k * n + c) * L64 + j] : 0;
  }
}

cudaError_t launch_flat_to_planar(const uint64_t *flat, int n, size_t count, uint64_t *out, int tiled, cudaStream_t st) {
  if (count == 0) return cudaSuccess;
  size_t blocks = (count * PLANAR_U64 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_flat_to_planar<<<(unsigned)blocks, 256, 0, st>>>(flat, n, count, out, tiled);
  return cudaGetLastError();
}

}  // namespace mfb
