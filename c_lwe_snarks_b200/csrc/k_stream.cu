// K2 — AES-256-CTR expansion of the a-vectors, standalone and fused into its consumers.
//
//   k_stream_bytes   raw keystream bytes                          (aesctr_prg aes.c:104-144, rng_seek entropy.c:46-56)
//   k_expand         AES -> resident (tile-planar) ciphertext array     (ct_import lwe.c:122-126, mpz2_urandomb entropy.c:11-26)
//   k_evalpoly       AES -> 704-bit MAC, a never touches HBM       (eval_poly lwe.c:176-186)
//   k_encrypt        AES -> <a, sk> + e*p + m -> 92-byte record    (regev_encrypt2 lwe.c:78-97 + ct_export :115-119)
//
// Stream geometry (snark.h:8-12, entropy.h:62-66): coordinate j of the ciphertext whose a-vector starts at stream
// byte `off` is the little-endian integer of the 92 bytes at off + 92*j; only its low 88 bytes survive modq.
//
// CTA shape: 512 threads, one CTA per SM.  Shared memory: 64 KB bank-replicated T-tables at a 64 KB-aligned
// address (aes256.cuh) and two keystream buffers of one coordinate tile each (double buffered, one
// __syncthreads per work item): the MAC/store phase of item t and the AES phase of item t+1 overlap
// across warps.  The bound is the shared-memory lookup rate: 224 conflict-free LDS per AES block.
#include "aes256.cuh"
#include "mfb_common.cuh"

namespace mfb {

constexpr int KS_THREADS = 512;
constexpr int KS_TILE = 490;                            // coordinates per tile; 1470 = 3 * 490
constexpr int KS_NTILES = N / KS_TILE;                  // 3
constexpr int KS_TILE_BYTES = KS_TILE * CT_BYTES;       // 45080
constexpr int KS_MAX_BLK = (KS_TILE_BYTES + 15 + 15) / 16;  // 2819 blocks cover a tile at any byte alignment
constexpr int KS_BUF_BYTES = KS_MAX_BLK * 16 + 16;      // + one word of slack for the funnel read
constexpr int KS_SMEM_BYTES = 0x20000 + KS_BUF_BYTES;   // [pad/buffer A | 64 KB tables | buffer B]

struct KsSmem {
  uint32_t tab;   // shared address of the tables (64 KB aligned)
  uint32_t buf[2];
  AesLut<2> lut;  // tab | 4*lane
};

__device__ __forceinline__ KsSmem ks_smem_setup(uint8_t *dyn, const uint32_t *__restrict__ t0_global) {
  KsSmem s;
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(dyn);
  s.tab = (s0 + 0xffffu) & ~0xffffu;
  // buffer A lives in the alignment pad below the tables when it fits there, else above buffer B
  if (s.tab - s0 < (uint32_t)KS_BUF_BYTES) __trap();  // dynamic smem starts within 20 KB of the window base
  s.buf[0] = s0;
  s.buf[1] = s.tab + AES_TAB_BYTES;
  s.lut.lbA = s.tab | ((threadIdx.x & 31) << 2);
  s.lut.lbB = 0;
  s.lut.m8 = s.lut.m16 = s.lut.m24 = 0;  // only read by the FMA-addressing variants
  s.lut.tex = 0;                         // only read by the texture variants
  aes_tables_init(dyn + (s.tab - s0), t0_global, threadIdx.x, blockDim.x);
  return s;
}

__device__ __forceinline__ void sts128(uint32_t addr, const AesState &v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.w0), "r"(v.w1), "r"(v.w2), "r"(v.w3)
               : "memory");
}

// keystream blocks [first, first + nblk) -> shared buffer (counter-mode cached AES, aes256.cuh)
__device__ __forceinline__ void ks_fill(const AesKey &key, uint64_t first, int nblk, uint32_t buf, const AesLut<2> &lut,
                                        AesCtrCache &cache) {
  for (int b = threadIdx.x; b < nblk; b += KS_THREADS)
    sts128(buf + 16u * b, aes256_ctr_block_cached<2>(lut, key, first + b, cache));
}

// the 22 live limbs of the coordinate that starts at byte `pos` of the shared buffer
__device__ __forceinline__ void ks_read_coord(uint32_t buf, uint32_t pos, uint32_t (&a)[22]) {
  const uint32_t w = buf + (pos & ~3u);
  const uint32_t sh = (pos & 3u) * 8;
  uint32_t lo = lds32(w);
#pragma unroll
  for (int l = 0; l < 22; l++) {
    const uint32_t hi = lds32(w + 4 * (l + 1));
    a[l] = __funnelshift_r(lo, hi, sh);
    lo = hi;
  }
}

struct TileGeom {
  uint64_t first;  // first AES block
  int nblk;
  uint32_t delta;  // byte offset of the tile's first coordinate inside block `first`
};
__device__ __forceinline__ TileGeom tile_geom(uint64_t ct_off, int tile) {
  const uint64_t off = ct_off + (uint64_t)tile * KS_TILE_BYTES;
  TileGeom g;
  g.first = off >> 4;
  g.delta = (uint32_t)(off & 15);
  g.nblk = (int)((g.delta + KS_TILE_BYTES + 15) >> 4);
  return g;
}

// ------------------------------------------------------------------------------------ raw bytes
// out[0..nbytes) = stream bytes [offset, offset + nbytes).  One AES block per thread iteration.
__global__ void __launch_bounds__(KS_THREADS, 1)
k_stream_bytes(const __grid_constant__ AesKey key, const uint32_t *__restrict__ t0_global, uint64_t offset,
               uint8_t *__restrict__ out, size_t nbytes) {
  extern __shared__ __align__(16) uint8_t dyn[];
  const KsSmem s = ks_smem_setup(dyn, t0_global);
  AesCtrCache cache;
  cache.window = ~0ull;
  __syncthreads();
  const uint64_t first = offset >> 4;
  const uint64_t end = offset + nbytes;
  const uint64_t nblk = ((end + 15) >> 4) - first;
  for (uint64_t b = (uint64_t)blockIdx.x * KS_THREADS + threadIdx.x; b < nblk; b += (uint64_t)gridDim.x * KS_THREADS) {
    const AesState v = aes256_ctr_block_cached<2>(s.lut, key, first + b, cache);
    const uint64_t p0 = (first + b) << 4;  // stream position of this block's byte 0
    const uint32_t w[4] = {v.w0, v.w1, v.w2, v.w3};
    if (p0 >= offset && p0 + 16 <= end && (((uintptr_t)(out + (p0 - offset))) & 15) == 0) {
      *reinterpret_cast<uint4 *>(out + (p0 - offset)) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const uint64_t p = p0 + i;
        if (p >= offset && p < end) out[p - offset] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
      }
    }
  }
}

// ------------------------------------------------------------------------------------ expand
// cts[k] (resident tile-planar layout) <- a-vector of the ciphertext at stream offset + k*CTR_CT, b from the wire record c8[k].
__global__ void __launch_bounds__(KS_THREADS, 1)
k_expand(const __grid_constant__ AesKey key, const uint32_t *__restrict__ t0_global, uint64_t offset,
         const uint8_t *__restrict__ c8, size_t count, uint64_t *__restrict__ cts) {
  extern __shared__ __align__(16) uint8_t dyn[];
  const KsSmem s = ks_smem_setup(dyn, t0_global);
  AesCtrCache cache;
  cache.window = ~0ull;
  // a CTA takes a CONTIGUOUS range of the (ciphertext, tile) items: consecutive items are consecutive in the
  // stream, so the counter-mode cache (one refill per 65536-block window) is refilled every ~23 items only
  const size_t total = count * KS_NTILES;
  const size_t per = total / gridDim.x, rem = total % gridDim.x;
  size_t it = blockIdx.x * per + (blockIdx.x < rem ? blockIdx.x : rem);
  const size_t nitems = it + per + (blockIdx.x < rem ? 1 : 0);  // end of this CTA's range
  int ph = 0;
  __syncthreads();  // tables ready
  if (it < nitems) {
    const TileGeom g = tile_geom(offset + (it / KS_NTILES) * (uint64_t)CTR_CT, (int)(it % KS_NTILES));
    ks_fill(key, g.first, g.nblk, s.buf[0], s.lut, cache);
  }
  __syncthreads();
  for (; it < nitems; it += 1, ph ^= 1) {
    const size_t k = it / KS_NTILES;
    const int tile = (int)(it % KS_NTILES);
    const TileGeom g = tile_geom(offset + k * (uint64_t)CTR_CT, tile);
    uint64_t *dst = cts + k * PLANAR_U64;
    if (threadIdx.x < KS_TILE) {
      uint32_t a[22];
      ks_read_coord((ph ? s.buf[1] : s.buf[0]), g.delta + CT_BYTES * threadIdx.x, a);
      const int c = tile * KS_TILE + threadIdx.x;
#pragma unroll
      for (int j = 0; j < L64; j++) dst[resident_index(c, j)] = (uint64_t)a[2 * j] | (uint64_t)a[2 * j + 1] << 32;
    } else if (tile == 0 && threadIdx.x < KS_TILE + L64) {
      // b coordinate: low 88 bytes of the wire record (ct_import lwe.c:125); padding coordinate: 0
      const int j = threadIdx.x - KS_TILE;
      const uint8_t *rec = c8 + k * CT_BYTES + 8 * j;
      uint64_t v = 0;
#pragma unroll
      for (int i = 0; i < 8; i++) v |= (uint64_t)rec[i] << (8 * i);
      dst[resident_index(N, j)] = v;
      dst[resident_index(N + 1, j)] = 0;
    }
    const size_t nx = it + 1;
    if (nx < nitems) {
      const TileGeom gn = tile_geom(offset + (nx / KS_NTILES) * (uint64_t)CTR_CT, (int)(nx % KS_NTILES));
      ks_fill(key, gn.first, gn.nblk, (ph ? s.buf[0] : s.buf[1]), s.lut, cache);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------ fused eval_poly
// partial[chunk] (planar, canonical) = sum over this CTA's ciphertexts of coeff * CT, for the CTA's tile.
// Work item m of the list: ciphertext index idx[m] (or m), scalar coeffs[m].  The grid is `ncta` CTAs (one per SM);
// CTA b works on tile b % NTILES as chunk b / NTILES of that tile, so the first ncta % NTILES tiles have one chunk
// more than the others; a tile's chunks split [0, d) into CONTIGUOUS, balanced ranges (consecutive ciphertexts are
// 8452.5 blocks apart: the counter-mode cache is refilled every ~8 items instead of every item).  The b coordinate
// (1470) comes from the wire records, not from the stream: k_bcoord below.
//
// Pipeline: THREE keystream buffers and two mbarriers per buffer (full: every warp has stored its blocks; empty:
// every warp has read its coordinates) instead of a CTA-wide __syncthreads per item.  A tile is 2818 blocks for 512
// threads — 5.5 per thread — and the assignment rotates by half a CTA every item, so that each warp gets 5 and 6
// blocks alternately; with the AES of item t+2 running while item t is consumed, a warp that is ahead keeps
// working instead of waiting for the slowest warp of every item (that wait cost 6/5.5 of the time).
// The warps are STAGGERED: half of the warps (two per scheduler) consume item t and then generate item t+2, the others
// generate item t+2 first and consume item t afterwards, so that at any time half of the warps are in the MAC phase (FMA pipe: IMAD.WIDE
// carry chains) while the other half is in the AES phase (ALU + LSU pipes) instead of all of them queueing on one.
constexpr int KS_NBUF = 3;
constexpr int KS3_SMEM_BYTES = 0x20000 + 2 * KS_BUF_BYTES;  // [pad: buffer 0 | 64 KB tables | buffers 1, 2]

__device__ __forceinline__ void ksb_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void ksb_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ksb_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}

// blocks [first, first + nblk) -> buffer; thread -> block assignment rotated by `rot` (a multiple of 32)
__device__ __forceinline__ void ks_fill_rot(const AesKey &key, uint64_t first, int nblk, uint32_t buf, const AesLut<2> &lut,
                                            AesCtrCache &cache, int rot) {
  for (int b = (threadIdx.x + rot) & (KS_THREADS - 1); b < nblk; b += KS_THREADS)
    sts128(buf + 16u * b, aes256_ctr_block_cached<2>(lut, key, first + b, cache));
}

template <int NVEC>
__global__ void __launch_bounds__(KS_THREADS, 1)
k_evalpoly(const __grid_constant__ AesKey key, const uint32_t *__restrict__ t0_global, uint64_t offset,
           const uint32_t *__restrict__ coeffs0, const uint32_t *__restrict__ coeffs1,
           const uint32_t *__restrict__ idx, size_t d, int nparts, uint64_t *__restrict__ partial0,
           uint64_t *__restrict__ partial1) {
  // Tiles of 490 coordinates, thread t < 490 owns coordinate t.  NVEC = 1: one E/O accumulator (43 registers, 22
  // IMAD.WIDE.X per MAC).  NVEC = 2: TWO scalar vectors over the same keystream tile; the two accumulators are kept
  // CANONICAL (22 limbs each, 44 registers together — what one E/O accumulator takes), so the kernel keeps the
  // full-size tile, the 5.5 blocks per thread and the register budget of NVEC = 1; a canonical MAC is the even chain
  // on the register pairs (r[2i], r[2i+1]) plus the odd chain on (r[2i+1], r[2i+2]) (acc_mad_canon), a few more
  // instructions than the E/O form in a kernel whose MAC phase is 2-3 % of its instructions.
  constexpr int TILE = KS_TILE;
  constexpr int NTILES = KS_NTILES;
  constexpr int TILE_BYTES = KS_TILE_BYTES;
  extern __shared__ __align__(16) uint8_t dyn[];
  __shared__ __align__(8) uint64_t bars[2 * KS_NBUF];
  KsSmem s = ks_smem_setup(dyn, t0_global);
  auto buf_of = [&](int b) { return b == 0 ? s.buf[0] : s.buf[1] + (uint32_t)(b - 1) * (uint32_t)KS_BUF_BYTES; };
  const uint32_t bbase = (uint32_t)__cvta_generic_to_shared(bars);
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2 * KS_NBUF; b++) ksb_init(bbase + 8 * b, KS_THREADS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  AesCtrCache cache;
  cache.window = ~0ull;
  const int tile = blockIdx.x % NTILES;
  const int chunk = blockIdx.x / NTILES;
  const int lane = threadIdx.x & 31;
  const int lc = threadIdx.x;
  const bool is_mac = threadIdx.x < TILE;

  auto geom = [&](size_t m) {
    const size_t k = idx ? idx[m] : m;
    const uint64_t off = offset + k * (uint64_t)CTR_CT + (uint64_t)tile * TILE_BYTES;
    TileGeom g;
    g.first = off >> 4;
    g.delta = (uint32_t)(off & 15);
    g.nblk = (int)((g.delta + TILE_BYTES + 15) >> 4);
    return g;
  };
  // this tile's chunks: tiles below gridDim.x % NTILES have one more; chunk c takes a balanced contiguous range
  const int nch = (int)(gridDim.x / NTILES) + (tile < (int)(gridDim.x % NTILES) ? 1 : 0);
  const size_t per = d / nch, rem = d % nch;
  const size_t m0 = (size_t)chunk * per + ((size_t)chunk < rem ? (size_t)chunk : rem);
  const size_t nitems = per + ((size_t)chunk < rem ? 1 : 0);
  auto fill = [&](size_t t) {  // item t of this CTA -> buffer t % 3
    const int b = (int)(t % KS_NBUF);
    if (t >= KS_NBUF) ksb_wait(bbase + 8 * (KS_NBUF + b), (uint32_t)((t / KS_NBUF - 1) & 1));  // its previous reader is done
    const TileGeom g = geom(m0 + t);
    // 2818 blocks = 5.5 per thread: rotate by half a CTA (5, 6, 5, 6, ...), every warp averages the same work
    ks_fill_rot(key, g.first, g.nblk, buf_of(b), s.lut, cache, (int)((t & 1) * (KS_THREADS / 2)));
    __syncwarp();
    if (lane == 0) ksb_arrive(bbase + 8 * b);
  };

  Acc704 acc;            // NVEC = 1
  uint32_t r0[22], r1[22];  // NVEC = 2
  if constexpr (NVEC == 1) {
    acc_zero(acc);
  } else {
#pragma unroll
    for (int l = 0; l < 22; l++) r0[l] = r1[l] = 0;
  }
  __syncthreads();  // tables and barriers ready
  // warp w runs on scheduler w % 4: warps 4-7 and 12-15 are the late ones, so every scheduler has two of each kind
  const bool late = (threadIdx.x >> 7) & 1;  // late warps: AES of item t+2 before the MAC of item t
  if (nitems > 0) fill(0);
  if (nitems > 1) fill(1);
  for (size_t t = 0; t < nitems; t++) {
    const int b = (int)(t % KS_NBUF);
    const size_t m = m0 + t;
    if (late && t + 2 < nitems) fill(t + 2);
    ksb_wait(bbase + 8 * b, (uint32_t)((t / KS_NBUF) & 1));  // every warp has stored its blocks of item t
    if (is_mac) {
      const TileGeom g = geom(m);
      uint32_t a[22];
      ks_read_coord(buf_of(b), g.delta + CT_BYTES * lc, a);
      if constexpr (NVEC == 1) {
        acc_mad(acc, a, coeffs0[m]);
      } else {
        acc_mad_canon(r0, a, coeffs0[m]);
        acc_mad_canon(r1, a, coeffs1[m]);
      }
    }
    __syncwarp();
    if (lane == 0) ksb_arrive(bbase + 8 * (KS_NBUF + b));
    if (!late && t + 2 < nitems) fill(t + 2);
  }

  if (is_mac) {
    // the finish kernel adds `nparts` partials for every coordinate: a tile with one chunk fewer zeroes the last slot
    const int c = tile * TILE + lc;
    if (chunk == nch - 1 && nch < nparts) {
#pragma unroll
      for (int v = 0; v < NVEC; v++) {
        uint64_t *z = (v ? partial1 : partial0) + (size_t)nch * PLANAR_U64;
#pragma unroll
        for (int j = 0; j < L64; j++) z[(size_t)j * NCP + c] = 0;
      }
    }
    if constexpr (NVEC == 1) acc_fold(acc, r0);
    uint64_t *out = partial0 + (size_t)chunk * PLANAR_U64;
#pragma unroll
    for (int j = 0; j < L64; j++) out[(size_t)j * NCP + c] = (uint64_t)r0[2 * j] | (uint64_t)r0[2 * j + 1] << 32;
    if constexpr (NVEC == 2) {
      out = partial1 + (size_t)chunk * PLANAR_U64;
#pragma unroll
      for (int j = 0; j < L64; j++) out[(size_t)j * NCP + c] = (uint64_t)r1[2 * j] | (uint64_t)r1[2 * j + 1] << 32;
    }
  }
}

// The b coordinate (1470) of the same partial sums, straight from the 92-byte wire records (ct_import lwe.c:125): this is
// the ONLY consumer of the records, so the host flavour copies them to the device on a second stream WHILE k_evalpoly
// (which needs only the seed and the scalars) is already running.  CTA k (one per partial slot) takes a balanced
// contiguous range of the work items, every thread MACs its records, the CTA reduces (REDUX on 16-bit halves, then a
// 22-step carry chain) and writes coordinate 1470 — and the zero padding coordinate 1471 — of partial slot k.
constexpr int KB_THREADS = 256;
template <int NVEC>
__global__ void __launch_bounds__(KB_THREADS)
k_bcoord(const uint8_t *__restrict__ c8, const uint32_t *__restrict__ coeffs0, const uint32_t *__restrict__ coeffs1,
         const uint32_t *__restrict__ idx, size_t d, uint64_t *__restrict__ partial0, uint64_t *__restrict__ partial1) {
  __shared__ uint32_t red[NVEC][KB_THREADS / 32][2 * L32];
  __shared__ unsigned long long cols[NVEC][L32];
  const size_t per = d / gridDim.x, rem = d % gridDim.x;
  const size_t m0 = blockIdx.x * per + (blockIdx.x < rem ? blockIdx.x : rem);
  const size_t m1 = m0 + per + (blockIdx.x < rem ? 1 : 0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Acc704 acc[NVEC];
#pragma unroll
  for (int v = 0; v < NVEC; v++) acc_zero(acc[v]);
  for (size_t m = m0 + threadIdx.x; m < m1; m += KB_THREADS) {
    const size_t k = idx ? idx[m] : m;
    const uint32_t *rec = reinterpret_cast<const uint32_t *>(c8 + k * CT_BYTES);  // 92 % 4 == 0
    uint32_t a[22];
#pragma unroll
    for (int l = 0; l < 22; l++) a[l] = rec[l];
    acc_mad(acc[0], a, coeffs0[m]);
    if constexpr (NVEC == 2) acc_mad(acc[1], a, coeffs1[m]);
  }
#pragma unroll
  for (int v = 0; v < NVEC; v++) {
    uint32_t r[22];
    acc_fold(acc[v], r);
#pragma unroll
    for (int l = 0; l < 22; l++) {
      const uint32_t lo = __reduce_add_sync(0xffffffffu, r[l] & 0xffffu);
      const uint32_t hi = __reduce_add_sync(0xffffffffu, r[l] >> 16);
      if (lane == 0) {
        red[v][warp][2 * l] = lo;
        red[v][warp][2 * l + 1] = hi;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < NVEC * L32) {
    const int v = threadIdx.x / L32, l = threadIdx.x % L32;
    unsigned long long lo = 0, hi = 0;
    for (int w = 0; w < KB_THREADS / 32; w++) {
      lo += red[v][w][2 * l];
      hi += red[v][w][2 * l + 1];
    }
    cols[v][l] = lo + (hi << 16);  // < 2^42
  }
  __syncthreads();
  if (threadIdx.x < NVEC) {
    const int v = threadIdx.x;
    uint64_t *out = (v ? partial1 : partial0) + (size_t)blockIdx.x * PLANAR_U64;
    uint64_t carry = 0;
    uint32_t prev = 0;
    for (int l = 0; l < L32; l++) {
      const uint64_t t = cols[v][l] + carry;
      carry = t >> 32;
      if (l & 1) out[(size_t)(l >> 1) * NCP + N] = (uint64_t)prev | (uint64_t)(uint32_t)t << 32;
      else prev = (uint32_t)t;
    }
    for (int j = 0; j < L64; j++) out[(size_t)j * NCP + N + 1] = 0;
  }
}

// ------------------------------------------------------------------------------------ Regev encryption
// b_k = (e_k * p + <sk, a_k> + m_k) mod 2^704, written as the 92-byte wire record.  sk is row-planar [11][1472].
// e_k = little-endian integer of ent[k*ent_stride .. +ent_nbytes) (69 noise bytes of errdist_uniform lwe.c:60-63;
// the sign byte that follows is consumed by the host protocol and never used, lwe.c:86-87).
//
// A CTA takes a contiguous range of whole ciphertexts; its work items are their 3 coordinate tiles in order, on the
// three-buffer mbarrier pipeline of k_evalpoly — but WARP-SPECIALISED: 16 producer warps only generate keystream
// (ALU + LSU pipes) and 4 consumer warps, one per scheduler, only multiply: per coordinate a 22 x 22-limb low product
// is 253 IMAD.WIDE carry chains, i.e. a long DEPENDENT sequence (one carry flag), so a warp inside it issues about one
// instruction every 8 cycles.  Measured on the unspecialised version: the keystream rate is proportional to the warp
// time spent generating it, so every cycle a producer warp spent inside the product was lost (34.4 ms for 131 136
// ciphertexts against 28.6 ms of pure keystream).  The consumers take ~4 coordinates per thread and tile and sum all
// their coordinates into ONE accumulator (the dot product adds them anyway); they are about half busy.
//   producers: wait empty[t % 3] -> AES of item t -> arrive full[t % 3]                 (16 arrivals)
//   consumers: wait full[t % 3] -> acc += a_c * sk_c for their coordinates -> arrive empty (4 arrivals)
// After a ciphertext's third tile the consumer warps reduce (REDUX on 16-bit halves, 32-bit shared atomics, a counter);
// the last one adds e*p + m, propagates the carries (one uniform 22-step chain) and writes the record.
constexpr int KE_PRODUCERS = KS_THREADS;        // 512 threads = 16 warps
constexpr int KE_CONSUMERS = 128;               // 4 warps
constexpr int KE_THREADS = KE_PRODUCERS + KE_CONSUMERS;
// register budget per thread after the re-balancing: 640 threads x 96 = 61440 registers at launch;
// 512 x 80 + 128 x 160 = 61440
constexpr int KE_PRODUCER_REGS = 80;
constexpr int KE_CONSUMER_REGS = 160;

__global__ void __launch_bounds__(KE_THREADS, 1)
k_encrypt(const __grid_constant__ AesKey key, const uint32_t *__restrict__ t0_global, uint64_t offset,
          const uint64_t *__restrict__ sk, const uint64_t *__restrict__ msg, const uint8_t *__restrict__ ent,
          int ent_stride, int ent_nbytes, size_t count, uint8_t *__restrict__ out_c8) {
  extern __shared__ __align__(16) uint8_t dyn[];
  __shared__ __align__(8) uint64_t bars[2 * KS_NBUF];
  __shared__ uint32_t cols[2][2 * L32];  // [ciphertext parity][limb l: low-half sum, high-half sum]
  __shared__ uint32_t arrived[2];
  KsSmem s = ks_smem_setup(dyn, t0_global);
  auto buf_of = [&](int b) { return b == 0 ? s.buf[0] : s.buf[1] + (uint32_t)(b - 1) * (uint32_t)KS_BUF_BYTES; };
  const uint32_t bbase = (uint32_t)__cvta_generic_to_shared(bars);
  if (threadIdx.x == 0) {
    for (int b = 0; b < KS_NBUF; b++) {
      ksb_init(bbase + 8 * b, KE_PRODUCERS / 32);              // full: every producer warp has stored its blocks
      ksb_init(bbase + 8 * (KS_NBUF + b), KE_CONSUMERS / 32);  // empty: every consumer warp has read its coordinates
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    arrived[0] = arrived[1] = 0;
  }
  if (threadIdx.x < 4 * L32) cols[threadIdx.x / (2 * L32)][threadIdx.x % (2 * L32)] = 0;
  const int lane = threadIdx.x & 31;
  // contiguous, balanced range of ciphertexts per CTA (keeps the counter-mode cache warm across items)
  const size_t per = count / gridDim.x, rem = count % gridDim.x;
  const size_t k0 = blockIdx.x * per + (blockIdx.x < rem ? blockIdx.x : rem);
  const size_t nct = per + (blockIdx.x < rem ? 1 : 0);
  const size_t nitems = nct * KS_NTILES;
  __syncthreads();  // tables, barriers, columns ready

  if (threadIdx.x < KE_PRODUCERS) {
    // ---------------------------------------------------------------- producers: keystream only
    // The CTA's register file is re-balanced between the two roles (setmaxnreg, warpgroup-wide): the 16 producer warps
    // give registers back, the consumer warpgroup takes them — its 22 x 22-limb product needs the accumulator (43), one
    // operand (22) and the other operand's limbs in flight at once, which does not fit the 96 registers a 640-thread
    // CTA starts with.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(KE_PRODUCER_REGS));
    AesCtrCache cache;
    cache.window = ~0ull;
    for (size_t t = 0; t < nitems; t++) {
      const int b = (int)(t % KS_NBUF);
      if (t >= KS_NBUF) ksb_wait(bbase + 8 * (KS_NBUF + b), (uint32_t)((t / KS_NBUF - 1) & 1));  // its previous readers are done
      const TileGeom g = tile_geom(offset + (k0 + t / KS_NTILES) * (uint64_t)CTR_CT, (int)(t % KS_NTILES));
      ks_fill_rot(key, g.first, g.nblk, buf_of(b), s.lut, cache, (int)((t & 1) * (KS_THREADS / 2)));
      __syncwarp();
      if (lane == 0) ksb_arrive(bbase + 8 * b);
    }
    return;
  }

  // ------------------------------------------------------------------ consumers: <a, sk> and the record
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(KE_CONSUMER_REGS));
  const int ct_id = threadIdx.x - KE_PRODUCERS;  // 0..127
  Acc704 acc;
  acc_zero(acc);
  for (size_t t = 0; t < nitems; t++) {
    const int b = (int)(t % KS_NBUF);
    const int tile = (int)(t % KS_NTILES);
    const size_t k = k0 + t / KS_NTILES;
    ksb_wait(bbase + 8 * b, (uint32_t)((t / KS_NBUF) & 1));
    const TileGeom g = tile_geom(offset + k * (uint64_t)CTR_CT, tile);
    for (int lc = ct_id; lc < KS_TILE; lc += KE_CONSUMERS) {
      uint32_t a[22];
      const int c = tile * KS_TILE + lc;
      // (all 22 key limbs up front: loading them row by row as the product needs them measured 6 % slower)
      uint32_t w[22];
#pragma unroll
      for (int j = 0; j < L64; j++) {
        const uint64_t v = __ldg(sk + (size_t)j * NCP + c);
        w[2 * j] = (uint32_t)v;
        w[2 * j + 1] = (uint32_t)(v >> 32);
      }
      ks_read_coord(buf_of(b), g.delta + CT_BYTES * lc, a);
      acc_mul(acc, a, w);
    }
    __syncwarp();
    if (lane == 0) ksb_arrive(bbase + 8 * (KS_NBUF + b));

    if (tile == KS_NTILES - 1) {
      const int slot = (int)((t / KS_NTILES) & 1);
      uint32_t r[22];
      acc_fold(acc, r);
      acc_zero(acc);
      uint32_t mylo = 0, myhi = 0;
#pragma unroll
      for (int l = 0; l < 22; l++) {
        const uint32_t lo = __reduce_add_sync(0xffffffffu, r[l] & 0xffffu);
        const uint32_t hi = __reduce_add_sync(0xffffffffu, r[l] >> 16);
        if (lane == l) {
          mylo = lo;
          myhi = hi;
        }
      }
      if (lane < L32) {
        atomicAdd(&cols[slot][2 * lane], mylo);  // < 128 * 2^16: no overflow
        atomicAdd(&cols[slot][2 * lane + 1], myhi);
      }
      __syncwarp();
      uint32_t prev = 0;
      if (lane == 0) {
        __threadfence_block();
        prev = atomicAdd(&arrived[slot], 1u);
      }
      prev = __shfl_sync(0xffffffffu, prev, 0);
      if (prev == KE_CONSUMERS / 32 - 1) {  // this warp is the last one: all column sums are in
        __threadfence_block();
        unsigned long long col = 0;
        uint32_t el = 0;
        if (lane < L32) {
          volatile uint32_t *cv = cols[slot];
          col = (unsigned long long)cv[2 * lane] + ((unsigned long long)cv[2 * lane + 1] << 16);
          cv[2 * lane] = 0;  // re-armed for the ciphertext after the next
          cv[2 * lane + 1] = 0;
          const uint8_t *e8 = ent + k * (size_t)ent_stride;
          for (int i = 0; i < 4; i++) {
            const int byte = 4 * lane + i;
            if (byte < ent_nbytes) el |= (uint32_t)e8[byte] << (8 * i);
          }
        }
        if (lane == 0) *(volatile uint32_t *)&arrived[slot] = 0;
        // b = sum_l col_l << 32l  +  e * p  +  m      (mod 2^704), the same 22-step chain in every lane
        const uint64_t mm = msg[k];
        uint64_t carry = 0;  // < 2^34 throughout
        uint32_t myb = 0;
#pragma unroll
        for (int l = 0; l < L32; l++) {
          uint64_t tt = __shfl_sync(0xffffffffu, col, l) + carry;  // < 2^44
          if (l == 0) tt += mm & 0xffffffffu;
          if (l == 1) tt += mm >> 32;
          const uint64_t prod = (uint64_t)__shfl_sync(0xffffffffu, el, l) * P;
          const uint64_t sum = tt + prod;
          const uint64_t c_out = sum < prod ? 1 : 0;
          carry = (sum >> 32) + (c_out << 32);
          if (lane == l) myb = (uint32_t)sum;
        }
        // lanes 0..21: the 88 live bytes; lane 22: bytes 88..91 of the record (zero)
        if (lane <= L32) reinterpret_cast<uint32_t *>(out_c8 + k * CT_BYTES)[lane] = myb;  // 92 % 4 == 0
      }
    }
  }
}

// ------------------------------------------------------------------------------------ Regev encryption, other (n, log q)
// BASELINE configs[4]: the reference implements ONE parameter set (lwe.h:23-31; any other GAMMA_LOGQ is `#error "Not
// implemented"`, lwe.h:119-121), so this kernel has no reference output to match: it is the design of k_encrypt with the
// limb count a template parameter and n, the coordinate width and the tiling run-time values, checked against plain
// integer arithmetic:
//   coordinate j of ciphertext k = little-endian integer of the ctb = log q / 8 stream bytes at offset + (k n + j) ctb,
//   reduced mod q_eff = 2^(64 L), L = floor(log q / 64) (what lwe.h:108-118 does for 736: mask, then drop the top limb);
//   b_k = (e_k p + <sk, a_k> + m_k) mod q_eff, written as a ctb-byte record.
// A ciphertext is cut into `ntiles` tiles of at most `tile` coordinates (tile * ctb <= the 45 KB keystream buffer).
//
// Bank conflicts of the consumers' reads: lane L reads word l of coordinate c + L, i.e. words ctb/4 apart.  For the
// reference's width (92 bytes = 23 words) and for 100 bytes (25 words) the stride is odd and the 32 lanes hit 32 banks; for
// log q = 512 / 768 / 1024 (16 / 24 / 32 words) they hit 2 / 4 / ONE bank(s): 33 reads of 32 wavefronts each per coordinate
// round, on the LSU pipe the AES lookups are bound by (measured: 0.49 of the lookup bound at log q = 1024).  PADDED layout
// (pad_wb = ctb / 16 > 0, chosen by the launcher when it lowers the conflict degree): the producers leave one 16-byte
// block free after every ctb bytes of keystream (block b goes to slot b + b / pad_wb), so the coordinate stride becomes
// ctb/4 + 4 words — 20 / 28 / 36: 4-way instead of 16 / 8 / 32-way; a 16-byte pad keeps the producers' STS.128 aligned.
template <int L>
__global__ void __launch_bounds__(KE_THREADS, 1)
k_encrypt_g(const __grid_constant__ AesKey key, const uint32_t *__restrict__ t0_global, uint64_t offset,
            const uint64_t *__restrict__ sk, int sk_stride, const uint64_t *__restrict__ msg, const uint8_t *__restrict__ ent,
            int ent_stride, int ent_nbytes, size_t count, int n, int ctb, int tile, int ntiles, int pad_wb, uint32_t pad_inv,
            uint8_t *__restrict__ out_c8) {
  constexpr int NL = 2 * L;
  extern __shared__ __align__(16) uint8_t dyn[];
  __shared__ __align__(8) uint64_t bars[2 * KS_NBUF];
  __shared__ uint32_t cols[2][2 * NL];
  __shared__ uint32_t arrived[2];
  KsSmem s = ks_smem_setup(dyn, t0_global);
  auto buf_of = [&](int b) { return b == 0 ? s.buf[0] : s.buf[1] + (uint32_t)(b - 1) * (uint32_t)KS_BUF_BYTES; };
  const uint32_t bbase = (uint32_t)__cvta_generic_to_shared(bars);
  if (threadIdx.x == 0) {
    for (int b = 0; b < KS_NBUF; b++) {
      ksb_init(bbase + 8 * b, KE_PRODUCERS / 32);
      ksb_init(bbase + 8 * (KS_NBUF + b), KE_CONSUMERS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    arrived[0] = arrived[1] = 0;
  }
  if (threadIdx.x < 4 * NL) cols[threadIdx.x / (2 * NL)][threadIdx.x % (2 * NL)] = 0;
  const int lane = threadIdx.x & 31;
  const size_t per = count / gridDim.x, rem = count % gridDim.x;
  const size_t k0 = blockIdx.x * per + (blockIdx.x < rem ? blockIdx.x : rem);
  const size_t nct = per + (blockIdx.x < rem ? 1 : 0);
  const size_t nitems = nct * (size_t)ntiles;
  const uint64_t ct_stream = (uint64_t)n * (uint64_t)ctb;  // stream bytes per ciphertext
  auto geom = [&](size_t t) {
    const size_t k = k0 + t / ntiles;
    const int tl = (int)(t % ntiles);
    const uint64_t off = offset + k * ct_stream + (uint64_t)tl * tile * ctb;
    const int nco = n - tl * tile < tile ? n - tl * tile : tile;
    TileGeom g;
    g.first = off >> 4;
    g.delta = (uint32_t)(off & 15);
    g.nblk = (int)((g.delta + (uint32_t)nco * ctb + 15) >> 4);
    return g;
  };
  __syncthreads();

  if (threadIdx.x < KE_PRODUCERS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(KE_PRODUCER_REGS));
    AesCtrCache cache;
    cache.window = ~0ull;
    for (size_t t = 0; t < nitems; t++) {
      const int b = (int)(t % KS_NBUF);
      if (t >= KS_NBUF) ksb_wait(bbase + 8 * (KS_NBUF + b), (uint32_t)((t / KS_NBUF - 1) & 1));
      const TileGeom g = geom(t);
      const int rot = (int)((t & 1) * (KS_THREADS / 2));
      if (pad_wb == 0) {
        ks_fill_rot(key, g.first, g.nblk, buf_of(b), s.lut, cache, rot);
      } else {  // padded layout: block bl -> slot bl + bl / pad_wb (pad_inv: the launcher's exact reciprocal)
        const uint32_t bufb = buf_of(b);
        for (int bl = (threadIdx.x + rot) & (KS_THREADS - 1); bl < g.nblk; bl += KS_THREADS)
          sts128(bufb + 16u * ((uint32_t)bl + (((uint32_t)bl * pad_inv) >> 16)),
                 aes256_ctr_block_cached<2>(s.lut, key, g.first + bl, cache));
      }
      __syncwarp();
      if (lane == 0) ksb_arrive(bbase + 8 * b);
    }
    return;
  }

  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(KE_CONSUMER_REGS));
  const int ct_id = threadIdx.x - KE_PRODUCERS;
  AccN<NL> acc;
  acc_zero(acc);
  for (size_t t = 0; t < nitems; t++) {
    const int b = (int)(t % KS_NBUF);
    const int tl = (int)(t % ntiles);
    const size_t k = k0 + t / ntiles;
    ksb_wait(bbase + 8 * b, (uint32_t)((t / KS_NBUF) & 1));
    const TileGeom g = geom(t);
    const int nco = n - tl * tile < tile ? n - tl * tile : tile;
    for (int lc = ct_id; lc < nco; lc += KE_CONSUMERS) {
      uint32_t a[NL], w[NL];
      const int c = tl * tile + lc;
#pragma unroll
      for (int j = 0; j < L; j++) {
        const uint64_t v = __ldg(sk + (size_t)j * sk_stride + c);
        w[2 * j] = (uint32_t)v;
        w[2 * j + 1] = (uint32_t)(v >> 32);
      }
      {  // the NL live limbs of the coordinate that starts at byte pos of the keystream buffer
        const uint32_t pos = g.delta + (uint32_t)ctb * lc;
        const uint32_t sh = (pos & 3u) * 8;
        if (pad_wb == 0) {
          const uint32_t wa = buf_of(b) + (pos & ~3u);
          uint32_t lo = lds32(wa);
#pragma unroll
          for (int l = 0; l < NL; l++) {
            const uint32_t hi = lds32(wa + 4 * (l + 1));
            a[l] = __funnelshift_r(lo, hi, sh);
            lo = hi;
          }
        } else {
          // padded layout: the coordinate starts in window lc (16 lc pad bytes before it); its word l lies in the next
          // window — one more pad block away — once (delta & ~3) + 4 l reaches ctb
          const uint32_t wa = buf_of(b) + (pos & ~3u) + 16u * (uint32_t)lc;
          const uint32_t room = (uint32_t)ctb - (g.delta & ~3u);  // bytes of this window from the coordinate's first word on
          uint32_t lo = lds32(wa);
#pragma unroll
          for (int l = 0; l < NL; l++) {
            const uint32_t hi = lds32(wa + 4 * (l + 1) + (4u * (l + 1) >= room ? 16u : 0u));
            a[l] = __funnelshift_r(lo, hi, sh);
            lo = hi;
          }
        }
      }
      acc_mul(acc, a, w);
    }
    __syncwarp();
    if (lane == 0) ksb_arrive(bbase + 8 * (KS_NBUF + b));

    if (tl == ntiles - 1) {
      const int slot = (int)((t / ntiles) & 1);
      uint32_t r[NL];
      acc_fold(acc, r);
      acc_zero(acc);
      uint32_t mylo = 0, myhi = 0;
#pragma unroll
      for (int l = 0; l < NL; l++) {
        const uint32_t lo = __reduce_add_sync(0xffffffffu, r[l] & 0xffffu);
        const uint32_t hi = __reduce_add_sync(0xffffffffu, r[l] >> 16);
        if (lane == l) {
          mylo = lo;
          myhi = hi;
        }
      }
      if (lane < NL) {
        atomicAdd(&cols[slot][2 * lane], mylo);
        atomicAdd(&cols[slot][2 * lane + 1], myhi);
      }
      __syncwarp();
      uint32_t prev = 0;
      if (lane == 0) {
        __threadfence_block();
        prev = atomicAdd(&arrived[slot], 1u);
      }
      prev = __shfl_sync(0xffffffffu, prev, 0);
      if (prev == KE_CONSUMERS / 32 - 1) {
        __threadfence_block();
        unsigned long long col = 0;
        uint32_t el = 0;
        if (lane < NL) {
          volatile uint32_t *cv = cols[slot];
          col = (unsigned long long)cv[2 * lane] + ((unsigned long long)cv[2 * lane + 1] << 16);
          cv[2 * lane] = 0;
          cv[2 * lane + 1] = 0;
          const uint8_t *e8 = ent + k * (size_t)ent_stride;
          for (int i = 0; i < 4; i++) {
            const int byte = 4 * lane + i;
            if (byte < ent_nbytes) el |= (uint32_t)e8[byte] << (8 * i);
          }
        }
        if (lane == 0) *(volatile uint32_t *)&arrived[slot] = 0;
        const uint64_t mm = msg[k];
        uint64_t carry = 0;
        uint32_t myb = 0;
#pragma unroll
        for (int l = 0; l < NL; l++) {
          uint64_t tt = __shfl_sync(0xffffffffu, col, l) + carry;
          if (l == 0) tt += mm & 0xffffffffu;
          if (l == 1) tt += mm >> 32;
          const uint64_t prod = (uint64_t)__shfl_sync(0xffffffffu, el, l) * P;
          const uint64_t sum = tt + prod;
          const uint64_t c_out = sum < prod ? 1 : 0;
          carry = (sum >> 32) + (c_out << 32);
          if (lane == l) myb = (uint32_t)sum;
        }
        // the record: NL live words, then zero up to ctb bytes (ctb is a multiple of 4, ctb / 4 <= 32)
        if (lane < ctb / 4) reinterpret_cast<uint32_t *>(out_c8 + k * (size_t)ctb)[lane] = lane < NL ? myb : 0u;
      }
    }
  }
}

// ------------------------------------------------------------------------------------ launchers
static cudaError_t ks_attr(const void *fn) {
  return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, KS_SMEM_BYTES);
}

cudaError_t launch_stream_bytes(const AesKey &key, const uint32_t *t0, uint64_t offset, uint8_t *out, size_t nbytes,
                                int sm_count, cudaStream_t st) {
  if (nbytes == 0) return cudaSuccess;
  cudaError_t e = ks_attr((const void *)k_stream_bytes);
  if (e != cudaSuccess) return e;
  const size_t nblk = (nbytes + 31) / 16;
  size_t grid = (nblk + KS_THREADS - 1) / KS_THREADS;
  if (grid > (size_t)sm_count) grid = sm_count;
  k_stream_bytes<<<(unsigned)grid, KS_THREADS, KS_SMEM_BYTES, st>>>(key, t0, offset, out, nbytes);
  return cudaGetLastError();
}

cudaError_t launch_expand(const AesKey &key, const uint32_t *t0, uint64_t offset, const uint8_t *c8, size_t count,
                          uint64_t *cts, int sm_count, cudaStream_t st) {
  if (count == 0) return cudaSuccess;
  cudaError_t e = ks_attr((const void *)k_expand);
  if (e != cudaSuccess) return e;
  size_t grid = count * KS_NTILES;
  if (grid > (size_t)sm_count) grid = sm_count;
  k_expand<<<(unsigned)grid, KS_THREADS, KS_SMEM_BYTES, st>>>(key, t0, offset, c8, count, cts);
  return cudaGetLastError();
}

// CTAs of the fused eval_poly grid: one per SM, at most one per (tile, ciphertext); *nparts = partial sums per coordinate
static int evalpoly_plan(size_t d, int sm_count, int ntiles, int max_parts, int *nparts) {
  size_t ncta = (size_t)sm_count;
  if (ncta > (size_t)ntiles * d) ncta = (size_t)ntiles * d;
  if (ncta > (size_t)ntiles * max_parts) ncta = (size_t)ntiles * max_parts;
  if (ncta < (size_t)ntiles) ncta = ntiles;
  *nparts = (int)((ncta + ntiles - 1) / ntiles);
  return (int)ncta;
}
constexpr int EVALPOLY_MAX_PARTS = 128;  // mfb_capi.cu: MAX_CHUNKS = 256 row-planar slots, two vectors

int evalpoly_nchunks(size_t d, int sm_count) {
  int nparts;
  evalpoly_plan(d, sm_count, KS_NTILES, EVALPOLY_MAX_PARTS, &nparts);
  return nparts;
}

// writes `nchunks` (= evalpoly_nchunks) canonical planar partial sums to partial_ws
// (the a coordinates: needs no wire records)
cudaError_t launch_evalpoly_partials(const AesKey &key, const uint32_t *t0, uint64_t offset, const uint32_t *coeffs,
                                     const uint32_t *idx, size_t d, int nchunks, int sm_count, uint64_t *partial_ws,
                                     cudaStream_t st) {
  if (d == 0 || nchunks == 0) return cudaSuccess;
  int nparts;
  const int ncta = evalpoly_plan(d, sm_count, KS_NTILES, nchunks, &nparts);
  if (nparts != nchunks) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute((const void *)k_evalpoly<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, KS3_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  k_evalpoly<1><<<ncta, KS_THREADS, KS3_SMEM_BYTES, st>>>(key, t0, offset, coeffs, nullptr, idx, d, nparts, partial_ws, nullptr);
  return cudaGetLastError();
}

// the b coordinate of the same `nchunks` partial sums (coeffs1 / partial1 = nullptr: one scalar vector)
cudaError_t launch_bcoord_partials(const uint8_t *c8, const uint32_t *coeffs0, const uint32_t *coeffs1, const uint32_t *idx,
                                   size_t d, int nchunks, uint64_t *partial0, uint64_t *partial1, cudaStream_t st) {
  if (d == 0 || nchunks == 0) return cudaSuccess;
  if (coeffs1)
    k_bcoord<2><<<nchunks, KB_THREADS, 0, st>>>(c8, coeffs0, coeffs1, idx, d, partial0, partial1);
  else
    k_bcoord<1><<<nchunks, KB_THREADS, 0, st>>>(c8, coeffs0, nullptr, idx, d, partial0, nullptr);
  return cudaGetLastError();
}

int evalpoly2_nchunks(size_t d, int sm_count) {
  int nparts;
  evalpoly_plan(d, sm_count, KS_NTILES, EVALPOLY_MAX_PARTS, &nparts);
  return nparts;
}

// two scalar vectors in one pass: writes nchunks partial sums to partial0 and nchunks to partial1
cudaError_t launch_evalpoly2_partials(const AesKey &key, const uint32_t *t0, uint64_t offset,
                                      const uint32_t *coeffs0, const uint32_t *coeffs1, size_t d, int nchunks, int sm_count,
                                      uint64_t *partial0, uint64_t *partial1, cudaStream_t st) {
  if (d == 0 || nchunks == 0) return cudaSuccess;
  int nparts;
  const int ncta = evalpoly_plan(d, sm_count, KS_NTILES, nchunks, &nparts);
  if (nparts != nchunks) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute((const void *)k_evalpoly<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, KS3_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  k_evalpoly<2><<<ncta, KS_THREADS, KS3_SMEM_BYTES, st>>>(key, t0, offset, coeffs0, coeffs1, nullptr, d, nparts, partial0,
                                                            partial1);
  return cudaGetLastError();
}

cudaError_t launch_encrypt(const AesKey &key, const uint32_t *t0, uint64_t offset, const uint64_t *sk,
                           const uint64_t *msg, const uint8_t *ent, int ent_stride, int ent_nbytes, size_t count,
                           uint8_t *out_c8, int sm_count, cudaStream_t st) {
  if (count == 0) return cudaSuccess;
  size_t grid = count < (size_t)sm_count ? count : (size_t)sm_count;
  cudaError_t e = cudaFuncSetAttribute((const void *)k_encrypt, cudaFuncAttributeMaxDynamicSharedMemorySize, KS3_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  k_encrypt<<<(unsigned)grid, KE_THREADS, KS3_SMEM_BYTES, st>>>(key, t0, offset, sk, msg, ent, ent_stride, ent_nbytes,
                                                               count, out_c8);
  return cudaGetLastError();
}

// Tiling of k_encrypt_g for a coordinate width of ctb bytes (host logic; mfb_encrypt_generic_plan exposes it to the tests).
//  * padded keystream layout (see the kernel): when the width is a whole number of AES blocks and one pad block per
//    coordinate lowers the bank-conflict degree gcd(words per coordinate, 32) of the consumers' reads;
//  * tiles: at most the coordinates whose (padded) keystream fits one buffer, balanced — or, when that needs fewer consumer
//    rounds per ciphertext, tiles of a whole number of rounds: the 128 consumer threads walk a tile in rounds of 128
//    coordinates and a partly filled round costs a full one (log q = 1024 padded: 7 tiles of 293 = 21 rounds against 8
//    tiles of 256 = 16).
void encrypt_generic_plan(int n, int ctb, int *tile_out, int *ntiles_out, int *pad_wb_out, uint32_t *pad_inv_out) {
  auto gcd32 = [](int v) { int g = 32; while (v % g) g >>= 1; return g; };
  int pad_wb = 0;
  uint32_t pad_inv = 0;
  if (ctb % 16 == 0 && gcd32(ctb / 4 + 4) < gcd32(ctb / 4)) {
    pad_wb = ctb / 16;
    pad_inv = 65536u / (uint32_t)pad_wb + 1;
    for (uint32_t bl = 0; bl <= (uint32_t)KS_MAX_BLK; bl++)  // the reciprocal must be exact for every block of a tile
      if (((bl * pad_inv) >> 16) != bl / (uint32_t)pad_wb) pad_wb = 0;
    if (!pad_wb) pad_inv = 0;
  }
  const int max_tile = KS_TILE_BYTES / (ctb + (pad_wb ? 16 : 0));
  int ntiles = (n + max_tile - 1) / max_tile;
  int tile = (n + ntiles - 1) / ntiles;  // balanced
  const int t128 = max_tile / KE_CONSUMERS * KE_CONSUMERS;
  if (t128 >= KE_CONSUMERS) {
    auto rounds = [&](int tl, int nt) {
      const int last = n - (nt - 1) * tl;
      return (nt - 1) * ((tl + KE_CONSUMERS - 1) / KE_CONSUMERS) + (last + KE_CONSUMERS - 1) / KE_CONSUMERS;
    };
    const int nt128 = (n + t128 - 1) / t128;
    if (rounds(t128, nt128) < rounds(tile, ntiles)) {
      tile = t128;
      ntiles = nt128;
    }
  }
  *tile_out = tile;
  *ntiles_out = ntiles;
  *pad_wb_out = pad_wb;
  *pad_inv_out = pad_inv;
}

template <int L>
static cudaError_t run_encrypt_g(const AesKey &key, const uint32_t *t0, uint64_t offset, const uint64_t *sk, int sk_stride,
                                 const uint64_t *msg, const uint8_t *ent, int ent_stride, int ent_nbytes, size_t count, int n, int ctb,
                                 uint8_t *out_c8, int sm_count, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute((const void *)k_encrypt_g<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, KS3_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  int tile, ntiles, pad_wb;
  uint32_t pad_inv;
  encrypt_generic_plan(n, ctb, &tile, &ntiles, &pad_wb, &pad_inv);
  const size_t grid = count < (size_t)sm_count ? count : (size_t)sm_count;
  k_encrypt_g<L><<<(unsigned)grid, KE_THREADS, KS3_SMEM_BYTES, st>>>(key, t0, offset, sk, sk_stride, msg, ent, ent_stride, ent_nbytes,
                                                                      count, n, ctb, tile, ntiles, pad_wb, pad_inv, out_c8);
  return cudaGetLastError();
}

cudaError_t launch_encrypt_generic(int limbs64, const AesKey &key, const uint32_t *t0, uint64_t offset, const uint64_t *sk,
                                   int sk_stride, const uint64_t *msg, const uint8_t *ent, int ent_stride, int ent_nbytes,
                                   size_t count, int n, int ctb, uint8_t *out_c8, int sm_count, cudaStream_t st) {
  if (count == 0) return cudaSuccess;
  switch (limbs64) {
#define MFB_CASE(Lv) \
  case Lv: return run_encrypt_g<Lv>(key, t0, offset, sk, sk_stride, msg, ent, ent_stride, ent_nbytes, count, n, ctb, out_c8, sm_count, st);
    MFB_CASE(4) MFB_CASE(6) MFB_CASE(8) MFB_CASE(10) MFB_CASE(11) MFB_CASE(12) MFB_CASE(13) MFB_CASE(14) MFB_CASE(16)
#undef MFB_CASE
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace mfb
