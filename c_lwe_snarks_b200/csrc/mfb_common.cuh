// Shared definitions for the sm_100a kernels of the SSP-SNARK hot path.
//
// Parameter set = the reference's only one (lwe.h:23-31): n = 1470, log q = 736 with the *effective*
// modulus 2^704 that modq() (lwe.h:108-118) implements, p = 2^32 - 5.  All arithmetic is unsigned
// multi-limb integer; "reduction mod q" is truncation to 22 x 32-bit limbs.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "mfb_chains.cuh"

namespace mfb {

constexpr int N = 1470;            // GAMMA_N (lwe.h:23)
constexpr int NC = N + 1;          // coordinates per ciphertext: a_0..a_1469, b
constexpr int NCP = 1472;          // coordinates padded to a multiple of 64 (23 x 64) for the planar layout
constexpr int L32 = 22;            // live 32-bit limbs per coordinate (704 bit)
constexpr int L64 = 11;            // live 64-bit limbs per coordinate
constexpr int CT_BYTES = 92;       // wire width of one coordinate (LOGQ_BYTES, lwe.h:29)
constexpr int CTR_CT = CT_BYTES * N;  // stream bytes per ciphertext (snark.h:8) = 135240
constexpr uint32_t P = 0xfffffffbu;   // GAMMA_P (lwe.h:25)
constexpr int NOISE_BYTES = 69;    // (GAMMA_LOG_SIGMA + 3) / 8 (lwe.c:62, entropy.c:32)
constexpr int ENT_BYTES = 70;      // noise bytes + the sign byte that is drawn and discarded (lwe.c:87)

// Resident ("tile-planar") layout of a ciphertext array in HBM.  The 1472 coordinates of a ciphertext are cut
// into 23 tiles of 64; a tile is 11 limb rows of 64 x u64, stored contiguously (5632 B):
//   u64 index of (ct i, 64-bit limb row j, coordinate c) = i*16192 + (c / 64)*704 + j*64 + (c % 64)
// so that (ct, tile) is ONE contiguous, 16-byte-aligned block for a TMA bulk copy, and inside it a warp reads a
// limb row as 256 contiguous bytes (bank-conflict free LDS.64).  Coordinate 1470 is b, 1471 is zero padding.
constexpr int RT_TILE = 64;                       // coordinates per tile
constexpr int RT_NTILES = NCP / RT_TILE;          // 23
constexpr int RT_TILE_U64 = L64 * RT_TILE;        // 704 u64 = 5632 B
constexpr size_t PLANAR_U64 = (size_t)L64 * NCP;  // 16192 u64 = 129536 B per ciphertext
__host__ __device__ __forceinline__ size_t resident_index(int c, int j) {
  return (size_t)(c / RT_TILE) * RT_TILE_U64 + (size_t)j * RT_TILE + (c % RT_TILE);
}
// "Row-planar" layout (secret keys, partial sums): limb row j, coordinate c at j*1472 + c.

// Symmetric exchange buffer of one rank for the peer-memory finish (k_lincomb_finish_peer): for each call parity q,
// each source rank s and each LANE (one exchange call combines up to PEER_LANES ciphertexts at once — the prover's
// four accumulators), a flat [1472][11] u64 slot, followed by the arrival flags [q][s][lane][23 tiles] (u32).
constexpr int PEER_MAX = 16;
constexpr int PEER_LANES = 4;
constexpr size_t PEER_SLOT_BYTES = PLANAR_U64 * 8;  // 129536
__host__ __device__ __forceinline__ size_t peer_slot_offset(uint32_t q, int world, int src, int lane) {
  return (((size_t)q * world + src) * PEER_LANES + lane) * PEER_SLOT_BYTES;
}
__host__ __device__ __forceinline__ size_t peer_flag_offset(uint32_t q, int world, int src, int lane, int tile) {
  return (size_t)2 * world * PEER_LANES * PEER_SLOT_BYTES +
         ((((size_t)q * world + src) * PEER_LANES + lane) * RT_NTILES + tile) * 4;
}
__host__ __device__ __forceinline__ size_t peer_buffer_bytes(int world) {
  return (size_t)2 * world * PEER_LANES * PEER_SLOT_BYTES + (size_t)2 * world * PEER_LANES * RT_NTILES * 4;
}

// ---------------------------------------------------------------------------------------------
// 704-bit multiply-accumulate with in-thread carry chains.
//
// acc (mod 2^704) is kept as two interleaved accumulators so that every product lands on a 64-bit
// boundary of one of them and a whole row is one carry chain:
//   E[0..21]  holds sum of  s*a[2k]   << 64k      (k = 0..10)
//   O[0..20]  holds sum of  s*a[2k+1] << 64k      (k = 0..10, the last one low half only),
//             worth O << 32 in the final value.
// value = (E + (O << 32)) mod 2^704.  Carries out of the top are dropped: that IS modq.
// The lo/hi pairs below are fused by ptxas into IMAD.WIDE.U32 with carry-in/-out (one per limb).
// ---------------------------------------------------------------------------------------------
// NL = number of 32-bit limbs (22 for the reference's parameter set; other even values for the (n, log q) sweep).
template <int NL>
struct AccN {
  uint32_t E[NL];
  uint32_t O[NL - 1];
};
using Acc704 = AccN<22>;

template <int NL>
__device__ __forceinline__ void acc_zero(AccN<NL> &x) {
#pragma unroll
  for (int i = 0; i < NL; i++) x.E[i] = 0;
#pragma unroll
  for (int i = 0; i < NL - 1; i++) x.O[i] = 0;
}

// The chains below are ONE inline-asm statement each (mfb_chains.cuh, generated): the CC flag that links the limbs
// never has to survive between statements.
// acc += s * a   (a = NL limbs, s < 2^32), mod 2^(32 NL)
template <int NL>
__device__ __forceinline__ void acc_mad(AccN<NL> &x, const uint32_t (&a)[NL], uint32_t s) {
  static_assert(NL % 2 == 0, "even limb count");
  MadChain<NL / 2, false>::template run<0, 0>(x.E, a, s);     // even limbs: E[0..NL-1]
  MadChain<NL / 2 - 1, true>::template run<0, 1>(x.O, a, s);  // odd limbs: O[0..NL-3], low half of a[NL-1] * s into O[NL-2]
}

// r += s * a on a CANONICAL accumulator (NL limbs, value = r): the even chain adds a[2i] * s on the register pairs
// (r[2i], r[2i+1]), the odd chain a[2i+1] * s on (r[2i+1], r[2i+2]) — half the registers of the E/O form; the odd chain's
// pairs are not 64-bit aligned, so it costs a few more instructions.  For kernels that need two accumulators in the
// register budget of one (k_evalpoly<2>).
template <int NL>
__device__ __forceinline__ void acc_mad_canon(uint32_t (&r)[NL], const uint32_t (&a)[NL], uint32_t s) {
  static_assert(NL % 2 == 0, "even limb count");
  MadChain<NL / 2, false>::template run<0, 0>(r, a, s);
  MadChain<NL / 2 - 1, true>::template run<1, 1>(r, a, s);
}

// acc += (a * b) mod 2^(32 NL) for two NL-limb operands (schoolbook low half: NL (NL + 1) / 2 limb products, 253 for 22).
// Row KB adds a[0..NL-1-KB] * b[KB] at limb offset KB as two carry chains (l even, l odd).  A product at
// limb position pos = l + KB goes to E[pos], E[pos+1] when pos is even and to O[pos-1], O[pos] when
// pos is odd; along a chain pos keeps its parity and advances by 2, so consecutive products occupy
// consecutive 64-bit slots of one accumulator and the carry flag links them.  One chain of every row
// ends at pos NL-2 (E[NL-2], E[NL-1]), the other at pos NL-1 (low half into O[NL-2]): both reach the top, where
// the carry is dropped (that is modq).
template <int NL, int KB, int Lx>
__device__ __forceinline__ void acc_mul_chain(AccN<NL> &x, const uint32_t (&a)[NL], uint32_t s) {
  constexpr int pos = Lx + KB;  // limb position of the chain's first product
  if constexpr (pos <= NL - 1) {
    if constexpr ((pos & 1) == 0)
      MadChain<(NL - 2 - pos) / 2 + 1, false>::template run<pos, Lx>(x.E, a, s);
    else
      MadChain<(NL - 1 - pos) / 2, true>::template run<pos - 1, Lx>(x.O, a, s);
  }
}

template <int NL, int KB>
__device__ __forceinline__ void acc_mul_rows(AccN<NL> &x, const uint32_t (&a)[NL], const uint32_t (&b)[NL]) {
  if constexpr (KB <= NL - 1) {
    acc_mul_chain<NL, KB, 0>(x, a, b[KB]);
    acc_mul_chain<NL, KB, 1>(x, a, b[KB]);
    acc_mul_rows<NL, KB + 1>(x, a, b);
  }
}

template <int NL>
__device__ __forceinline__ void acc_mul(AccN<NL> &x, const uint32_t (&a)[NL], const uint32_t (&b)[NL]) {
  acc_mul_rows<NL, 0>(x, a, b);
}

// r[0..NL-1] = (E + (O << 32)) mod 2^(32 NL)
template <int NL>
__device__ __forceinline__ void acc_fold(const AccN<NL> &x, uint32_t (&r)[NL]) {
  r[0] = x.E[0];
  AddChain<NL - 1>::template run<1, 1, 0>(r, x.E, x.O);
}

// r += b (22 limbs), mod 2^704
__device__ __forceinline__ void add704(uint32_t (&r)[22], const uint32_t (&b)[22]) { AccChain<22>::run(r, b); }

// r = a - b (22 limbs) mod 2^704; returns 1 when a < b (borrow out of the top)
__device__ __forceinline__ uint32_t sub704(uint32_t (&r)[22], const uint32_t (&a)[22], const uint32_t (&b)[22]) {
  return SubChain<22>::run<0, 0, 0>(r, a, b);
}

#define MFB_CUDA_TRY(expr)                         \
  do {                                             \
    cudaError_t _e = (expr);                       \
    if (_e != cudaSuccess) return mfb::fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

int fail(cudaError_t e, const char *what, const char *file, int line);  // mfb_capi.cu

}  // namespace mfb
