// K1 — prover ciphertext linear combination over ciphertexts RESIDENT in HBM.
//
// Replaces the inner loop of eval_poly (lwe.c:176-186): for every coordinate c,
//   rop[c] = (rop[c] + sum_i coeff_i * CT_i[c]) mod 2^704           (ct_addmul_ui lwe.c:141-149 + modq lwe.h:108-118)
// Exact sums mod 2^704 are associative, so splitting the i-range over CTAs (and GPUs) and adding the
// canonical partial sums afterwards is bit-identical to the reference's sequential fold.
//
// Data layout: the planar layout of mfb_common.cuh — one warp-wide LDG.64 per limb row reads 256
// contiguous bytes; nothing is read twice, nothing is staged.  HBM-bound: 129 536 B per ciphertext
// (129 448 B algorithmic), 22 IMAD.WIDE per coordinate.
//
// Grid: (NCP / TILE coordinate tiles) x (nchunks slices of the ciphertext index range).
// Each thread owns one coordinate, keeps the 704-bit accumulator pair (E, O) in registers for its
// whole slice, folds it once and writes a canonical partial sum; k_lincomb_finish adds the partials.
#include "mfb_common.cuh"

namespace mfb {

constexpr int LC_TILE = 64;    // threads per CTA = coordinates per tile; 1472 = 23 * 64
constexpr int LC_HSTAGE = 512; // scalars staged in shared memory per refill

template <int UNROLL>
__global__ void __launch_bounds__(LC_TILE) k_lincomb(const uint64_t *__restrict__ cts,
                                                      const uint32_t *__restrict__ coeffs, size_t d,
                                                      size_t chunk_len, uint64_t *__restrict__ partial) {
  __shared__ uint32_t hs[LC_HSTAGE];
  const int c = blockIdx.x * LC_TILE + threadIdx.x;
  const size_t i0 = (size_t)blockIdx.y * chunk_len;
  const size_t i1 = i0 + chunk_len < d ? i0 + chunk_len : d;

  Acc704 acc;
  acc_zero(acc);

  for (size_t base = i0; base < i1; base += LC_HSTAGE) {
    const int n = (int)(i1 - base < (size_t)LC_HSTAGE ? i1 - base : (size_t)LC_HSTAGE);
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += LC_TILE) hs[k] = coeffs[base + k];
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
        for (int j = 0; j < L64; j++) {
          a[2 * j] = (uint32_t)v[u][j];
          a[2 * j + 1] = (uint32_t)(v[u][j] >> 32);
        }
        acc_mad(acc, a, hs[k + u]);
      }
    }
    for (; k < n; k++) {
      uint32_t a[22];
#pragma unroll
      for (int j = 0; j < L64; j++) {
        const uint64_t v = __ldcs(p + (size_t)k * PLANAR_U64 + (size_t)j * NCP);
        a[2 * j] = (uint32_t)v;
        a[2 * j + 1] = (uint32_t)(v >> 32);
      }
      acc_mad(acc, a, hs[k]);
    }
  }

  uint32_t r[22];
  acc_fold(acc, r);
  uint64_t *out = partial + (size_t)blockIdx.y * PLANAR_U64 + c;
#pragma unroll
  for (int j = 0; j < L64; j++) out[(size_t)j * NCP] = (uint64_t)r[2 * j] | (uint64_t)r[2 * j + 1] << 32;
}

// rop[c] = (rop_in[c] + sum_k partial[k][c]) mod 2^704.  rop is "flat": [1471][11] u64, coordinate-major.
// CTA = 64 coordinates x FIN_SLICES slices of the partial index: each thread adds every FIN_SLICES-th partial
// (loads of successive partials are independent, so several are in flight), the slices meet in shared memory.
constexpr int FIN_SLICES = 8;
__global__ void __launch_bounds__(64 * FIN_SLICES)
k_lincomb_finish(const uint64_t *__restrict__ partial, int nparts, const uint64_t *rop_in, uint64_t *rop_out) {
  __shared__ uint32_t sm[FIN_SLICES - 1][22][64];
  const int cl = threadIdx.x & 63, slice = threadIdx.x >> 6;
  const int c = blockIdx.x * 64 + cl;
  uint32_t r[22];
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
  if (slice == 0 && c < NC) {
    for (int s2 = 0; s2 < FIN_SLICES - 1; s2++) {
      uint32_t b[22];
#pragma unroll
      for (int j = 0; j < 22; j++) b[j] = sm[s2][j][cl];
      add704(r, b);
    }
    if (rop_in) {
      uint32_t b[22];
#pragma unroll
      for (int j = 0; j < L64; j++) {
        const uint64_t v = rop_in[(size_t)c * L64 + j];
        b[2 * j] = (uint32_t)v;
        b[2 * j + 1] = (uint32_t)(v >> 32);
      }
      add704(r, b);
    }
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
int lincomb_nchunks(size_t d, int sm_count) {
  // ~2 waves of resident CTAs: 23 tiles x nchunks CTAs of 64 threads; aim for ~10 CTAs (20 warps) per SM
  size_t target = ((size_t)sm_count * 10 + 22) / 23;
  if (target < 1) target = 1;
  size_t n = d < target ? d : target;
  return (int)(n ? n : 1);
}

typedef void (*mark_fn)(void *, int, cudaStream_t);
// launches the main kernel only; *nchunks_inout returns the number of partial sums written
cudaError_t launch_lincomb_partials(const uint64_t *cts, const uint32_t *coeffs, size_t d, uint64_t *partial_ws,
                                    int *nchunks_inout, cudaStream_t st, mark_fn mark, void *mark_arg) {
  int nchunks = d ? *nchunks_inout : 0;
  if (nchunks > 0) {
    const size_t chunk_len = (d + nchunks - 1) / nchunks;
    nchunks = (int)((d + chunk_len - 1) / chunk_len);
    dim3 grid(NCP / LC_TILE, nchunks);
    if (mark) mark(mark_arg, 0, st);
    k_lincomb<2><<<grid, LC_TILE, 0, st>>>(cts, coeffs, d, chunk_len, partial_ws);
    if (mark) mark(mark_arg, 1, st);
  }
  *nchunks_inout = nchunks;
  return cudaGetLastError();
}

cudaError_t launch_lincomb_finish(const uint64_t *partial_ws, int nparts, const uint64_t *rop_in, uint64_t *rop_out,
                                  cudaStream_t st) {
  k_lincomb_finish<<<NCP / 64, 64 * FIN_SLICES, 0, st>>>(partial_ws, nparts, rop_in, rop_out);
  return cudaGetLastError();
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
