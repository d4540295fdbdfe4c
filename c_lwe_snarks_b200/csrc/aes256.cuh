// AES-256-CTR keystream for the a-vectors (replaces aes.c:49-144 + entropy.c:46-61 on the device).
//
// Stream definition (aes.c:122-133, entropy.c:58-61): with seed = nonce(8 B) || key(32 B),
//   keystream block k = AES256_Enc(key, nonce || LE64(k)),   stream byte p = byte p%16 of block p/16.
//
// Device formulation: FIPS-197 rounds through one 32-bit "T-table" in a little-endian column
// convention (column word = bytes 4c..4c+3 of the block, byte r = row r):
//   T0[x] = 2S | S<<8 | S<<16 | 3S<<24          (S = sbox[x]; contribution of a row-0 byte to its column)
//   T2[x] = rotl(T0[x], 16)                     (row-2 byte);   rows 1 and 3 are rotl8 of T0 / T2.
// Both tables sit in shared memory REPLICATED ACROSS THE 32 BANKS: entry x occupies 256 B,
//   [x*256 + 4*lane]       = T0[x]     and     [x*256 + 128 + 4*lane] = T2[x],
// so lane L only ever touches bank L: every lookup is conflict-free whatever the data
// (64 KB per CTA).  A lookup address is one PRMT: (byte << 8) | (lane << 2).
#pragma once
#include <cstdint>
#include <cstring>

namespace mfb {

constexpr int AES_TAB_BYTES = 256 * 256;  // 64 KB

struct AesKey {
  uint32_t rk[60];   // 15 round keys x 4 little-endian column words
  uint32_t nonce[2]; // block bytes 0..7
  uint32_t rkr[60];  // rotr8(rk[i]): lets the round key ride along inside the rotated half of a column (aes_round)
};

// ----------------------------------------------------------------------------------- host side
namespace aes_host {

inline uint8_t gmul(uint8_t a, uint8_t b) {
  uint8_t r = 0;
  while (b) {
    if (b & 1) r ^= a;
    a = (uint8_t)((a << 1) ^ ((a & 0x80) ? 0x1b : 0));
    b >>= 1;
  }
  return r;
}

inline const uint8_t *sbox() {
  static uint8_t S[256];
  static bool ready = false;
  if (!ready) {
    // multiplicative inverse via the generator 3, then the FIPS-197 affine map
    uint8_t p = 1, q = 1;
    do {
      p = (uint8_t)(p ^ (p << 1) ^ ((p & 0x80) ? 0x1b : 0));  // p *= 3
      q ^= (uint8_t)(q << 1);                                  // q /= 3
      q ^= (uint8_t)(q << 2);
      q ^= (uint8_t)(q << 4);
      if (q & 0x80) q ^= 0x09;
      uint8_t x = (uint8_t)(q ^ (q << 1 | q >> 7) ^ (q << 2 | q >> 6) ^ (q << 3 | q >> 5) ^ (q << 4 | q >> 4));
      S[p] = (uint8_t)(x ^ 0x63);
    } while (p != 1);
    S[0] = 0x63;
    ready = true;
  }
  return S;
}

// seed = nonce(8) || key(32)  (entropy.c:58-61)
inline void expand(const uint8_t seed[40], AesKey *out) {
  const uint8_t *S = sbox();
  const uint8_t *key = seed + 8;
  uint8_t w[60][4];
  std::memcpy(w, key, 32);
  uint8_t rcon = 1;
  for (int i = 8; i < 60; i++) {
    uint8_t t[4] = {w[i - 1][0], w[i - 1][1], w[i - 1][2], w[i - 1][3]};
    if (i % 8 == 0) {
      uint8_t t0 = t[0];
      t[0] = (uint8_t)(S[t[1]] ^ rcon);
      t[1] = S[t[2]];
      t[2] = S[t[3]];
      t[3] = S[t0];
      rcon = gmul(rcon, 2);
    } else if (i % 8 == 4) {
      for (int k = 0; k < 4; k++) t[k] = S[t[k]];
    }
    for (int k = 0; k < 4; k++) w[i][k] = (uint8_t)(w[i - 8][k] ^ t[k]);
  }
  for (int i = 0; i < 60; i++)
    out->rk[i] = (uint32_t)w[i][0] | (uint32_t)w[i][1] << 8 | (uint32_t)w[i][2] << 16 | (uint32_t)w[i][3] << 24;
  for (int i = 0; i < 60; i++) out->rkr[i] = out->rk[i] >> 8 | out->rk[i] << 24;
  std::memcpy(out->nonce, seed, 8);
}

inline void t0_table(uint32_t T0[256]) {
  const uint8_t *S = sbox();
  for (int x = 0; x < 256; x++) {
    uint32_t s = S[x], s2 = gmul((uint8_t)s, 2), s3 = s2 ^ s;
    T0[x] = s2 | s << 8 | s << 16 | s3 << 24;
  }
}

}  // namespace aes_host

// ----------------------------------------------------------------------------------- device side
#ifdef __CUDACC__

// Fill the bank-replicated tables of this CTA from the 1 KB global T0 table.  `tab` is 64 KB, 256-B aligned.
__device__ __forceinline__ void aes_tables_init(uint8_t *tab, const uint32_t *__restrict__ t0_global, int tid,
                                                int nthreads) {
  for (int w = tid; w < 256 * 64; w += nthreads) {
    const int x = w >> 6, slot = w & 63;
    uint32_t t = __ldg(t0_global + x);
    if (slot >= 32) t = __byte_perm(t, 0, 0x1032);  // rotl 16
    reinterpret_cast<uint32_t *>(tab)[w] = t;
  }
}

// Lookup addressing.  The tables live at a 64 KB-ALIGNED address of the CTA's shared window, so the
// full address  tab | (byte << 8) | (lane << 2)  is composed by ONE PRMT from
//   lanebase = tab | 4*lane        (bytes: [0] = 4*lane, [1] = 0, [2..3] = tab >> 16)
// and the state word; T2 is reached with the LDS immediate offset +128.  No adds on the lookup path.
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32_t2(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1+128];" : "=r"(v) : "r"(addr));
  return v;
}
template <int K>
__device__ __forceinline__ uint32_t aes_addr(uint32_t w, uint32_t lanebase) {
  // __byte_perm(x, y, s): nibble i of s selects result byte i from {x.b0..b3 = 0..3, y.b0..b3 = 4..7}
  return __byte_perm(lanebase, w, 0x3200 | ((4 + K) << 4));
}
__device__ __forceinline__ uint32_t rotl8(uint32_t x) { return __byte_perm(x, 0, 0x2103); }

struct AesState {
  uint32_t w0, w1, w2, w3;
};

// One full AES-256 encryption of block (nonce || LE64(ctr)); lb = lanebase (see aes_addr).
__device__ __forceinline__ AesState aes256_ctr_block(const AesKey &k, uint64_t ctr, uint32_t lb) {
  uint32_t s0 = k.nonce[0] ^ k.rk[0];
  uint32_t s1 = k.nonce[1] ^ k.rk[1];
  uint32_t s2 = (uint32_t)ctr ^ k.rk[2];
  uint32_t s3 = (uint32_t)(ctr >> 32) ^ k.rk[3];
#pragma unroll
  for (int r = 1; r < 14; r++) {
    // column c: T0[b0(s_c)] ^ rotl8(T0[b1(s_{c+1})]) ^ T2[b2(s_{c+2})] ^ rotl8(T2[b3(s_{c+3})]) ^ rk
    uint32_t a0 = lds32(aes_addr<0>(s0, lb)), a1 = lds32(aes_addr<1>(s1, lb));
    uint32_t a2 = lds32_t2(aes_addr<2>(s2, lb)), a3 = lds32_t2(aes_addr<3>(s3, lb));
    uint32_t b0 = lds32(aes_addr<0>(s1, lb)), b1 = lds32(aes_addr<1>(s2, lb));
    uint32_t b2 = lds32_t2(aes_addr<2>(s3, lb)), b3 = lds32_t2(aes_addr<3>(s0, lb));
    uint32_t c0 = lds32(aes_addr<0>(s2, lb)), c1 = lds32(aes_addr<1>(s3, lb));
    uint32_t c2 = lds32_t2(aes_addr<2>(s0, lb)), c3 = lds32_t2(aes_addr<3>(s1, lb));
    uint32_t d0 = lds32(aes_addr<0>(s3, lb)), d1 = lds32(aes_addr<1>(s0, lb));
    uint32_t d2 = lds32_t2(aes_addr<2>(s1, lb)), d3 = lds32_t2(aes_addr<3>(s2, lb));
    s0 = a0 ^ a2 ^ k.rk[4 * r + 0] ^ rotl8(a1 ^ a3);
    s1 = b0 ^ b2 ^ k.rk[4 * r + 1] ^ rotl8(b1 ^ b3);
    s2 = c0 ^ c2 ^ k.rk[4 * r + 2] ^ rotl8(c1 ^ c3);
    s3 = d0 ^ d2 ^ k.rk[4 * r + 3] ^ rotl8(d1 ^ d3);
  }
  // last round: SubBytes + ShiftRows only.  S sits in byte 0 of T2 (=S,3S,2S,S), byte 1 and 2 of
  // T0 (=2S,S,S,3S) and byte 3 of T2, i.e. already at the byte position where it is needed.
  AesState o;
  {
    uint32_t a0 = lds32_t2(aes_addr<0>(s0, lb)), a1 = lds32(aes_addr<1>(s1, lb));
    uint32_t a2 = lds32(aes_addr<2>(s2, lb)), a3 = lds32_t2(aes_addr<3>(s3, lb));
    uint32_t b0 = lds32_t2(aes_addr<0>(s1, lb)), b1 = lds32(aes_addr<1>(s2, lb));
    uint32_t b2 = lds32(aes_addr<2>(s3, lb)), b3 = lds32_t2(aes_addr<3>(s0, lb));
    uint32_t c0 = lds32_t2(aes_addr<0>(s2, lb)), c1 = lds32(aes_addr<1>(s3, lb));
    uint32_t c2 = lds32(aes_addr<2>(s0, lb)), c3 = lds32_t2(aes_addr<3>(s1, lb));
    uint32_t d0 = lds32_t2(aes_addr<0>(s3, lb)), d1 = lds32(aes_addr<1>(s0, lb));
    uint32_t d2 = lds32(aes_addr<2>(s1, lb)), d3 = lds32_t2(aes_addr<3>(s2, lb));
    o.w0 = ((a0 & 0x000000ffu) | (a1 & 0x0000ff00u) | (a2 & 0x00ff0000u) | (a3 & 0xff000000u)) ^ k.rk[56];
    o.w1 = ((b0 & 0x000000ffu) | (b1 & 0x0000ff00u) | (b2 & 0x00ff0000u) | (b3 & 0xff000000u)) ^ k.rk[57];
    o.w2 = ((c0 & 0x000000ffu) | (c1 & 0x0000ff00u) | (c2 & 0x00ff0000u) | (c3 & 0xff000000u)) ^ k.rk[58];
    o.w3 = ((d0 & 0x000000ffu) | (d1 & 0x0000ff00u) | (d2 & 0x00ff0000u) | (d3 & 0xff000000u)) ^ k.rk[59];
  }
  return o;
}

// ---------------------------------------------------------------------------------------------
// Counter-mode specialisation.  Within a window of 65536 consecutive blocks only the two low counter bytes
// change, i.e. bytes 8 and 9 of the AES input block.  After round 1 (ShiftRows moves row r of column 2 into
// column 2 - r) columns 0 and 3 are constant in the window and columns 1 and 2 need ONE lookup each; in round 2
// every output column takes two of its four bytes from the constant columns.  A per-thread cache keyed on
// ctr >> 16 holds those constants: a block then costs 2 + 8 + 12*16 = 202 lookups instead of 224.
//
// TABS = 2: tables T0 | T2 in one 64 KB region (rows 1 / 3 by rotl8);  TABS = 4: a second 64 KB region holds
// T1 | T3 (lbB), no rotations.
struct AesCtrCache {
  uint64_t window;        // ctr >> 16 the constants were computed for (init: ~0)
  uint32_t k1p, k2p;      // round-1 columns 1, 2 without their counter-dependent lookup
  uint32_t q0, q1, q2, q3;  // round-2 columns: contributions of the constant round-1 columns 0 and 3 (+ round key)
};

// FM: bit K set = the address of a byte-K lookup is computed on the FMA pipe (idle otherwise) instead of by the
// ALU-pipe PRMT: byte = mul.hi(w * 2^(8*(3-K)), 2^8), address = byte * 2^8 + lanebase.  The multipliers are run-time
// values (m8, m16, m24 come from kernel parameters) so that ptxas keeps IMADs and does not strength-reduce to shifts.
// FM bits 8..: number of row-0 lookups per full round (a0, b0, c0, d0 in this order) that go through the TEXTURE path
// (tex1Dfetch on the 1 KB global T0 table, `tex`) instead of shared memory — a development variant (tune_aes.cu).
template <int TABS, int FM = 0>
struct AesLut {
  uint32_t lbA, lbB;
  uint32_t m8, m16, m24;
  unsigned long long tex;  // cudaTextureObject_t of the global T0 table (texture variants only)
  template <int K> __device__ __forceinline__ uint32_t addr(uint32_t w, uint32_t lb) const {
    if constexpr ((FM >> K) & 1) {
      uint32_t b, a;
      if constexpr (K == 3) {
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(b) : "r"(w), "r"(m8));
      } else {
        uint32_t t;
        const uint32_t mm = K == 2 ? m8 : K == 1 ? m16 : m24;
        asm("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(w), "r"(mm));
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(b) : "r"(t), "r"(m8));
      }
      asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(b), "r"(m8), "r"(lb));
      return a;
    } else {
      return aes_addr<K>(w, lb);
    }
  }
  // Tn[byte K of w]
  template <int K> __device__ __forceinline__ uint32_t t0(uint32_t w) const { return lds32(addr<K>(w, lbA)); }
  template <int K> __device__ __forceinline__ uint32_t t2(uint32_t w) const { return lds32_t2(addr<K>(w, lbA)); }
  template <int K> __device__ __forceinline__ uint32_t t1(uint32_t w) const {
    if constexpr (TABS == 4) return lds32(addr<K>(w, lbB));
    else return rotl8(lds32(addr<K>(w, lbA)));
  }
  template <int K> __device__ __forceinline__ uint32_t t3(uint32_t w) const {
    if constexpr (TABS == 4) return lds32_t2(addr<K>(w, lbB));
    else return rotl8(lds32_t2(addr<K>(w, lbA)));
  }
  // raw loads (no rotation) from region A, for the formulations that rotate after XOR-combining
  template <int K> __device__ __forceinline__ uint32_t a0(uint32_t w) const { return lds32(addr<K>(w, lbA)); }
  template <int K> __device__ __forceinline__ uint32_t a2(uint32_t w) const { return lds32_t2(addr<K>(w, lbA)); }
};

// kr0..kr3 = rotr8 of the round key words (AesKey::rkr): in the two-table form a column is
//   T0[.] ^ T2[.] ^ rotl8(T0[.] ^ T2[.]) ^ k  =  LOP3(a0, a2, rotl8(LOP3(a1, a3, rotr8(k))))
// — two 3-input LOP3 and one PRMT instead of three LOP3 and one PRMT (the ALU pipe is this kernel's bound).
template <int TABS, int FM>
__device__ __forceinline__ void aes_round(const AesLut<TABS, FM> &L, uint32_t &s0, uint32_t &s1, uint32_t &s2, uint32_t &s3,
                                          uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3, uint32_t kr0, uint32_t kr1,
                                          uint32_t kr2, uint32_t kr3) {
  if constexpr (TABS == 4) {
    const uint32_t a0 = L.template t0<0>(s0), a1 = L.template t1<1>(s1), a2 = L.template t2<2>(s2), a3 = L.template t3<3>(s3);
    const uint32_t b0 = L.template t0<0>(s1), b1 = L.template t1<1>(s2), b2 = L.template t2<2>(s3), b3 = L.template t3<3>(s0);
    const uint32_t c0 = L.template t0<0>(s2), c1 = L.template t1<1>(s3), c2 = L.template t2<2>(s0), c3 = L.template t3<3>(s1);
    const uint32_t d0 = L.template t0<0>(s3), d1 = L.template t1<1>(s0), d2 = L.template t2<2>(s1), d3 = L.template t3<3>(s2);
    s0 = a0 ^ a1 ^ a2 ^ a3 ^ k0;
    s1 = b0 ^ b1 ^ b2 ^ b3 ^ k1;
    s2 = c0 ^ c1 ^ c2 ^ c3 ^ k2;
    s3 = d0 ^ d1 ^ d2 ^ d3 ^ k3;
  } else {
    constexpr int TX = FM >> 8;
    auto row0 = [&](uint32_t w, bool via_tex) -> uint32_t {
      if (via_tex) return tex1Dfetch<uint32_t>((cudaTextureObject_t)L.tex, (int)(w & 0xffu));
      return L.template a0<0>(w);
    };
    const uint32_t a0 = row0(s0, TX >= 1), a1 = L.template a0<1>(s1);
    const uint32_t a2 = L.template a2<2>(s2), a3 = L.template a2<3>(s3);
    const uint32_t b0 = row0(s1, TX >= 2), b1 = L.template a0<1>(s2);
    const uint32_t b2 = L.template a2<2>(s3), b3 = L.template a2<3>(s0);
    const uint32_t c0 = row0(s2, TX >= 3), c1 = L.template a0<1>(s3);
    const uint32_t c2 = L.template a2<2>(s0), c3 = L.template a2<3>(s1);
    const uint32_t d0 = row0(s3, TX >= 4), d1 = L.template a0<1>(s0);
    const uint32_t d2 = L.template a2<2>(s1), d3 = L.template a2<3>(s2);
    s0 = a0 ^ a2 ^ rotl8(a1 ^ a3 ^ kr0);
    s1 = b0 ^ b2 ^ rotl8(b1 ^ b3 ^ kr1);
    s2 = c0 ^ c2 ^ rotl8(c1 ^ c3 ^ kr2);
    s3 = d0 ^ d2 ^ rotl8(d1 ^ d3 ^ kr3);
  }
}

template <int TABS, int FM>
__device__ __forceinline__ AesState aes_last_round(const AesLut<TABS, FM> &L, const AesKey &k, uint32_t s0, uint32_t s1,
                                                   uint32_t s2, uint32_t s3) {
  // S sits in byte 0 and 3 of T2 (= S,3S,2S,S) and in byte 1 and 2 of T0 (= 2S,S,S,3S): already in position
  const uint32_t a0 = L.template a2<0>(s0), a1 = L.template a0<1>(s1);
  const uint32_t a2 = L.template a0<2>(s2), a3 = L.template a2<3>(s3);
  const uint32_t b0 = L.template a2<0>(s1), b1 = L.template a0<1>(s2);
  const uint32_t b2 = L.template a0<2>(s3), b3 = L.template a2<3>(s0);
  const uint32_t c0 = L.template a2<0>(s2), c1 = L.template a0<1>(s3);
  const uint32_t c2 = L.template a0<2>(s0), c3 = L.template a2<3>(s1);
  const uint32_t d0 = L.template a2<0>(s3), d1 = L.template a0<1>(s0);
  const uint32_t d2 = L.template a0<2>(s1), d3 = L.template a2<3>(s2);
  AesState o;
  o.w0 = ((a0 & 0x000000ffu) | (a1 & 0x0000ff00u) | (a2 & 0x00ff0000u) | (a3 & 0xff000000u)) ^ k.rk[56];
  o.w1 = ((b0 & 0x000000ffu) | (b1 & 0x0000ff00u) | (b2 & 0x00ff0000u) | (b3 & 0xff000000u)) ^ k.rk[57];
  o.w2 = ((c0 & 0x000000ffu) | (c1 & 0x0000ff00u) | (c2 & 0x00ff0000u) | (c3 & 0xff000000u)) ^ k.rk[58];
  o.w3 = ((d0 & 0x000000ffu) | (d1 & 0x0000ff00u) | (d2 & 0x00ff0000u) | (d3 & 0xff000000u)) ^ k.rk[59];
  return o;
}

template <int TABS, int FM>
__device__ __forceinline__ void aes_ctr_cache_fill(const AesLut<TABS, FM> &L, const AesKey &k, uint64_t ctr, AesCtrCache &c) {
  const uint32_t s0 = k.nonce[0] ^ k.rk[0], s1 = k.nonce[1] ^ k.rk[1];
  const uint32_t s2 = (uint32_t)ctr ^ k.rk[2], s3 = (uint32_t)(ctr >> 32) ^ k.rk[3];
  // round 1: columns 0 and 3 are constant in the window (they see counter bytes 2 and 3 only)
  const uint32_t r0 = L.template t0<0>(s0) ^ L.template t1<1>(s1) ^ L.template t2<2>(s2) ^ L.template t3<3>(s3) ^ k.rk[4];
  const uint32_t r3 = L.template t0<0>(s3) ^ L.template t1<1>(s0) ^ L.template t2<2>(s1) ^ L.template t3<3>(s2) ^ k.rk[7];
  c.k1p = L.template t0<0>(s1) ^ L.template t2<2>(s3) ^ L.template t3<3>(s0) ^ k.rk[5];  // + T1[b1(s2)]
  c.k2p = L.template t1<1>(s3) ^ L.template t2<2>(s0) ^ L.template t3<3>(s1) ^ k.rk[6];  // + T0[b0(s2)]
  // round 2: the part of each column that comes from r0 and r3
  c.q0 = L.template t0<0>(r0) ^ L.template t3<3>(r3) ^ k.rk[8];
  c.q1 = L.template t2<2>(r3) ^ L.template t3<3>(r0) ^ k.rk[9];
  c.q2 = L.template t1<1>(r3) ^ L.template t2<2>(r0) ^ k.rk[10];
  c.q3 = L.template t0<0>(r3) ^ L.template t1<1>(r0) ^ k.rk[11];
  c.window = ctr >> 16;
}

// One AES-256 encryption of (nonce || LE64(ctr)) with the counter-mode cache.
template <int TABS, int FM>
__device__ __forceinline__ AesState aes256_ctr_block_cached(const AesLut<TABS, FM> &L, const AesKey &k, uint64_t ctr,
                                                            AesCtrCache &c) {
  if ((ctr >> 16) != c.window) aes_ctr_cache_fill<TABS, FM>(L, k, ctr, c);
  const uint32_t s2 = (uint32_t)ctr ^ k.rk[2];
  // round 1, columns 1 and 2
  const uint32_t r1 = c.k1p ^ L.template t1<1>(s2);
  const uint32_t r2 = c.k2p ^ L.template t0<0>(s2);
  // round 2
  uint32_t s0 = c.q0 ^ L.template t1<1>(r1) ^ L.template t2<2>(r2);
  uint32_t s1 = c.q1 ^ L.template t0<0>(r1) ^ L.template t1<1>(r2);
  uint32_t t2v = c.q2 ^ L.template t0<0>(r2) ^ L.template t3<3>(r1);
  uint32_t s3 = c.q3 ^ L.template t2<2>(r1) ^ L.template t3<3>(r2);
  uint32_t s2v = t2v;
#pragma unroll
  for (int r = 3; r < 14; r++)
    aes_round<TABS, FM>(L, s0, s1, s2v, s3, k.rk[4 * r], k.rk[4 * r + 1], k.rk[4 * r + 2], k.rk[4 * r + 3], k.rkr[4 * r],
                        k.rkr[4 * r + 1], k.rkr[4 * r + 2], k.rkr[4 * r + 3]);
  return aes_last_round<TABS, FM>(L, k, s0, s1, s2v, s3);
}

// Two blocks of the SAME 65536-block window in lockstep (ctr_b = ctr_a + stride, both keyed on the same cache): the
// rounds of the two independent blocks interleave in one instruction stream, doubling the lookups in flight per warp.
// Falls back to two single calls when the blocks straddle a window boundary.
template <int TABS, int FM>
__device__ __forceinline__ void aes256_ctr_block_cached_x2(const AesLut<TABS, FM> &L, const AesKey &k, uint64_t ctr_a,
                                                           uint64_t ctr_b, AesCtrCache &c, AesState &out_a, AesState &out_b) {
  if ((ctr_a >> 16) != (ctr_b >> 16)) {
    out_a = aes256_ctr_block_cached<TABS, FM>(L, k, ctr_a, c);
    out_b = aes256_ctr_block_cached<TABS, FM>(L, k, ctr_b, c);
    return;
  }
  if ((ctr_a >> 16) != c.window) aes_ctr_cache_fill<TABS, FM>(L, k, ctr_a, c);
  const uint32_t xa = (uint32_t)ctr_a ^ k.rk[2], xb = (uint32_t)ctr_b ^ k.rk[2];
  const uint32_t ra1 = c.k1p ^ L.template t1<1>(xa), rb1 = c.k1p ^ L.template t1<1>(xb);
  const uint32_t ra2 = c.k2p ^ L.template t0<0>(xa), rb2 = c.k2p ^ L.template t0<0>(xb);
  uint32_t a0 = c.q0 ^ L.template t1<1>(ra1) ^ L.template t2<2>(ra2), b0 = c.q0 ^ L.template t1<1>(rb1) ^ L.template t2<2>(rb2);
  uint32_t a1 = c.q1 ^ L.template t0<0>(ra1) ^ L.template t1<1>(ra2), b1 = c.q1 ^ L.template t0<0>(rb1) ^ L.template t1<1>(rb2);
  uint32_t a2 = c.q2 ^ L.template t0<0>(ra2) ^ L.template t3<3>(ra1), b2 = c.q2 ^ L.template t0<0>(rb2) ^ L.template t3<3>(rb1);
  uint32_t a3 = c.q3 ^ L.template t2<2>(ra1) ^ L.template t3<3>(ra2), b3 = c.q3 ^ L.template t2<2>(rb1) ^ L.template t3<3>(rb2);
#pragma unroll
  for (int r = 3; r < 14; r++) {
    aes_round<TABS, FM>(L, a0, a1, a2, a3, k.rk[4 * r], k.rk[4 * r + 1], k.rk[4 * r + 2], k.rk[4 * r + 3], k.rkr[4 * r],
                        k.rkr[4 * r + 1], k.rkr[4 * r + 2], k.rkr[4 * r + 3]);
    aes_round<TABS, FM>(L, b0, b1, b2, b3, k.rk[4 * r], k.rk[4 * r + 1], k.rk[4 * r + 2], k.rk[4 * r + 3], k.rkr[4 * r],
                        k.rkr[4 * r + 1], k.rkr[4 * r + 2], k.rkr[4 * r + 3]);
  }
  out_a = aes_last_round<TABS, FM>(L, k, a0, a1, a2, a3);
  out_b = aes_last_round<TABS, FM>(L, k, b0, b1, b2, b3);
}

// Plain variant (no cache) on the same table abstraction.
template <int TABS, int FM>
__device__ __forceinline__ AesState aes256_ctr_block_plain(const AesLut<TABS, FM> &L, const AesKey &k, uint64_t ctr) {
  uint32_t s0 = k.nonce[0] ^ k.rk[0], s1 = k.nonce[1] ^ k.rk[1];
  uint32_t s2 = (uint32_t)ctr ^ k.rk[2], s3 = (uint32_t)(ctr >> 32) ^ k.rk[3];
#pragma unroll
  for (int r = 1; r < 14; r++)
    aes_round<TABS, FM>(L, s0, s1, s2, s3, k.rk[4 * r], k.rk[4 * r + 1], k.rk[4 * r + 2], k.rk[4 * r + 3], k.rkr[4 * r],
                        k.rkr[4 * r + 1], k.rkr[4 * r + 2], k.rkr[4 * r + 3]);
  return aes_last_round<TABS, FM>(L, k, s0, s1, s2, s3);
}

// Fill region B (T1 | T3, 64 KB) for TABS = 4 from the global T0 table.
__device__ __forceinline__ void aes_tables_init_b(uint8_t *tabB, const uint32_t *__restrict__ t0_global, int tid, int nthreads) {
  for (int w = tid; w < 256 * 64; w += nthreads) {
    const int x = w >> 6, slot = w & 63;
    uint32_t t = __ldg(t0_global + x);
    t = (slot >= 32) ? __byte_perm(t, 0, 0x0321) : __byte_perm(t, 0, 0x2103);  // T3 = rotl 24, T1 = rotl 8
    reinterpret_cast<uint32_t *>(tabB)[w] = t;
  }
}

#endif  // __CUDACC__
}  // namespace mfb
