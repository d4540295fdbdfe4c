// F_p[x] arithmetic for the prover's polynomial step and setup's evaluations.
//
//   prover (snark.c:138-169):  w = delta*t + sum_{witness bit i-1} v_i,  v = w + v_0,  h = (v^2 - 1) / t
//   setup  (snark.c:93-110):   v_i(s) for every SSP polynomial (Horner in the reference)
//
// p = 2^32 - 5 has 2-adicity 1 (p - 1 = 2*5*429496729), so F_p has no useful NTT.  Products are computed over the
// integers through three 30-bit NTT primes and Garner CRT (coefficients < D * (p-1)^2 < P1*P2*P3 ~ 2^86 for D <= 2^21), then
// reduced mod p; division is Newton inversion of the reversed divisor.  Results are canonical residues, hence
// bit-identical to FLINT's (and to the host stand-in c_lwe_snarks_b200/host/nmod_poly.c).
//
// NTT: radix-2, Montgomery arithmetic.  Forward = DIF (natural in, bit-reversed out), inverse = DIT (bit-reversed in,
// natural out), so no permutation pass is needed around the pointwise product.  A transform of n > NTT_B points is TWO
// launches ("four-step" form): k_ntt_cols does all the stages whose butterflies span NTT_B elements or more — for a
// tile of columns held in shared memory: a length-(n/NTT_B) transform down every column, then the column factors —
// and k_ntt_local the last log2(NTT_B) stages on contiguous blocks, also in shared memory; the data is read and
// written twice per transform instead of once per stage; both kernels do two radix-2 stages per pass in registers
// (radix-4 steps).  The prover's polynomial step at D = 2^20 went from 1.68 ms to 0.81 ms, at 2^18 from 0.61 to 0.33 ms.
// All three primes share a launch (grid.y).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mfb_common.cuh"

namespace mfb {

constexpr int NPR = 3;
__constant__ uint32_t c_P[NPR] = {998244353u, 469762049u, 167772161u};
__constant__ uint32_t c_PINV[NPR];  // -P^{-1} mod 2^32
__constant__ uint32_t c_R2[NPR];    // 2^64 mod P (to Montgomery form)
static const uint32_t h_P[NPR] = {998244353u, 469762049u, 167772161u};
static const uint32_t h_G[NPR] = {3u, 3u, 3u};

constexpr int NTT_LOG_B = 11;
constexpr int NTT_B = 1 << NTT_LOG_B;  // elements per CTA in the shared-memory kernel

__device__ __forceinline__ uint32_t mont_mul(uint32_t a, uint32_t b, uint32_t P, uint32_t pinv) {
  const uint64_t t = (uint64_t)a * b;
  const uint32_t m = (uint32_t)t * pinv;
  const uint32_t r = (uint32_t)((t + (uint64_t)m * P) >> 32);  // < 2P
  return r >= P ? r - P : r;
}
__device__ __forceinline__ uint32_t add_mod(uint32_t a, uint32_t b, uint32_t P) {
  const uint32_t s = a + b;
  return s >= P ? s - P : s;
}
__device__ __forceinline__ uint32_t sub_mod(uint32_t a, uint32_t b, uint32_t P) { return a >= b ? a - b : a + P - b; }

// ---- host-side modular helpers (table construction) ------------------------------------------------------
static uint32_t h_mulmod(uint32_t a, uint32_t b, uint32_t m) { return (uint32_t)((uint64_t)a * b % m); }
static uint32_t h_powmod(uint32_t a, uint64_t e, uint32_t m) {
  uint32_t r = 1;
  while (e) {
    if (e & 1) r = h_mulmod(r, a, m);
    a = h_mulmod(a, a, m);
    e >>= 1;
  }
  return r;
}

// tw[k][j] = w_k^j in Montgomery form, j < nmax/2, w_k a primitive nmax-th root of unity mod P_k (inverse table: w^-j)
__global__ void k_twiddle_fill(uint32_t *tw, uint32_t *twi, uint32_t nmax_half, uint32_t w0, uint32_t w1, uint32_t w2,
                               uint32_t wi0, uint32_t wi1, uint32_t wi2) {
  const int k = blockIdx.y;
  const uint32_t P = c_P[k], pinv = c_PINV[k], r2 = c_R2[k];
  const uint32_t w = k == 0 ? w0 : k == 1 ? w1 : w2, wi = k == 0 ? wi0 : k == 1 ? wi1 : wi2;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nmax_half; j += gridDim.x * blockDim.x) {
    // w^j by square-and-multiply in Montgomery form
    uint32_t base = mont_mul(w, r2, P, pinv), basei = mont_mul(wi, r2, P, pinv);
    uint32_t acc = mont_mul(1, r2, P, pinv), acci = acc;
    for (uint32_t e = j; e; e >>= 1) {
      if (e & 1) {
        acc = mont_mul(acc, base, P, pinv);
        acci = mont_mul(acci, basei, P, pinv);
      }
      base = mont_mul(base, base, P, pinv);
      basei = mont_mul(basei, basei, P, pinv);
    }
    tw[(size_t)k * nmax_half + j] = acc;
    twi[(size_t)k * nmax_half + j] = acci;
  }
}

// tw_loc[k][e] = w_NTT_B^e = tw[k][e * nmax / NTT_B] for e < NTT_B / 2 (and the inverse table likewise)
__global__ void k_twiddle_compact(const uint32_t *__restrict__ tw, const uint32_t *__restrict__ twi, uint32_t nmax_half,
                                  uint32_t *tw_loc, uint32_t *twi_loc) {
  const int k = blockIdx.y;
  const uint32_t step = nmax_half / (NTT_B / 2);
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < NTT_B / 2; e += gridDim.x * blockDim.x) {
    tw_loc[(size_t)k * (NTT_B / 2) + e] = tw[(size_t)k * nmax_half + (size_t)e * step];
    twi_loc[(size_t)k * (NTT_B / 2) + e] = twi[(size_t)k * nmax_half + (size_t)e * step];
  }
}

// all stages with span < NTT_B inside shared memory; one CTA per contiguous block of blk = min(n, NTT_B) elements.
// The twiddles of these stages are the blk/2 powers of w_blk: staged once per CTA in shared memory (twl[e] =
// w_blk^e = tw[e * nmax / blk]); stage `half` reads twl[j * (blk/2/half)] — no strided global loads in the loop.
// Two radix-2 stages at a time in registers (a radix-4 step on elements i, i + h/2, i + h, i + 3h/2): half the
// barriers and half the shared-memory traffic of a stage-by-stage loop; an odd stage count leaves one radix-2 stage.
constexpr int NTTL_T = NTT_B / 4;  // threads: one radix-4 step each per pass
// twloc: the COMPACT table of the NTT_B / 2 powers of w_NTT_B per prime (k_twiddle_compact): contiguous, so that the
// 3072 CTAs of a 2^21-point transform read 4 KB each in full lines instead of gathering 1024 sectors from the big table
// (that gather was two thirds of this kernel's L2 traffic).
// the local stages of one block held in shared memory (s: blk elements, twl: the blk/2 powers of w_blk or w_blk^-1);
// ends with a __syncthreads
template <bool INVERSE>
__device__ __forceinline__ void ntt_block_stages(uint32_t *s, const uint32_t *twl, uint32_t blk, uint32_t P, uint32_t pinv) {
  const uint32_t bh = blk >> 1;
  auto radix2 = [&](uint32_t half) {  // one plain stage (half = 1 when the stage count is odd)
    const uint32_t ts = bh / half;
    for (uint32_t t = threadIdx.x; t < bh; t += NTTL_T) {
      const uint32_t j = t & (half - 1), i = ((t - j) << 1) + j;
      const uint32_t a = s[i], b = s[i + half];
      if (!INVERSE) {
        s[i] = add_mod(a, b, P);
        s[i + half] = mont_mul(sub_mod(a, b, P), twl[j * ts], P, pinv);
      } else {
        const uint32_t bw = mont_mul(b, twl[j * ts], P, pinv);
        s[i] = add_mod(a, bw, P);
        s[i + half] = sub_mod(a, bw, P);
      }
    }
    __syncthreads();
  };
  auto radix4 = [&](uint32_t h) {  // the stages with half = h and half = h/2 (DIF: in this order; DIT: h/2 then h)
    const uint32_t q4 = h >> 1, ts = bh / h;
    for (uint32_t t = threadIdx.x; t < (blk >> 2); t += NTTL_T) {
      const uint32_t j = t & (q4 - 1), i = ((t - j) << 2) + j;
      const uint32_t t1 = twl[j * ts], t2 = twl[(j + q4) * ts], t3 = twl[2 * j * ts];
      uint32_t a0 = s[i], a1 = s[i + q4], a2 = s[i + h], a3 = s[i + h + q4];
      if (!INVERSE) {
        const uint32_t b0 = add_mod(a0, a2, P), b2 = mont_mul(sub_mod(a0, a2, P), t1, P, pinv);
        const uint32_t b1 = add_mod(a1, a3, P), b3 = mont_mul(sub_mod(a1, a3, P), t2, P, pinv);
        a0 = add_mod(b0, b1, P);
        a1 = mont_mul(sub_mod(b0, b1, P), t3, P, pinv);
        a2 = add_mod(b2, b3, P);
        a3 = mont_mul(sub_mod(b2, b3, P), t3, P, pinv);
      } else {
        const uint32_t u1 = mont_mul(a1, t3, P, pinv), u3 = mont_mul(a3, t3, P, pinv);
        const uint32_t b0 = add_mod(a0, u1, P), b1 = sub_mod(a0, u1, P);
        const uint32_t b2 = add_mod(a2, u3, P), b3 = sub_mod(a2, u3, P);
        const uint32_t v2 = mont_mul(b2, t1, P, pinv), v3 = mont_mul(b3, t2, P, pinv);
        a0 = add_mod(b0, v2, P);
        a2 = sub_mod(b0, v2, P);
        a1 = add_mod(b1, v3, P);
        a3 = sub_mod(b1, v3, P);
      }
      s[i] = a0;
      s[i + q4] = a1;
      s[i + h] = a2;
      s[i + h + q4] = a3;
    }
    __syncthreads();
  };
  if (!INVERSE) {
    uint32_t h = bh;
    for (; h >= 2; h >>= 2) radix4(h);
    if (h == 1) radix2(1);
  } else {
    int lg = 0;
    while ((2u << lg) <= bh) lg++;          // bh = 2^lg: stages half = 1, 2, ..., 2^lg
    uint32_t h = 2;
    if (((lg + 1) & 1) && bh >= 1) {         // odd stage count: the first one alone
      radix2(1);
      h = 4;
    }
    for (; h <= bh; h <<= 2) radix4(h);
  }
}

template <bool INVERSE>
__global__ void __launch_bounds__(NTTL_T) k_ntt_local(uint32_t *x, uint32_t n, const uint32_t *__restrict__ twloc,
                                                      uint32_t sc0, uint32_t sc1, uint32_t sc2) {
  __shared__ uint32_t s[NTT_B];
  __shared__ uint32_t twl[NTT_B / 2];
  const int k = blockIdx.y;
  const uint32_t P = c_P[k], pinv = c_PINV[k];
  const uint32_t scale = k == 0 ? sc0 : k == 1 ? sc1 : sc2;  // Montgomery n^-1 when this is the last kernel of an inverse
  const uint32_t blk = n < (uint32_t)NTT_B ? n : (uint32_t)NTT_B;
  const uint32_t bh = blk >> 1;
  uint32_t *xk = x + (size_t)k * n + (size_t)blockIdx.x * blk;
  const uint32_t *twk = twloc + (size_t)k * (NTT_B / 2);
  for (uint32_t i = threadIdx.x; i < blk; i += NTTL_T) s[i] = xk[i];
  for (uint32_t e = threadIdx.x; e < bh; e += NTTL_T) twl[e] = twk[(size_t)e * ((NTT_B / 2) / bh)];  // w_blk^e = w_NTT_B^(e NTT_B / blk)
  __syncthreads();
  ntt_block_stages<INVERSE>(s, twl, blk, P, pinv);
  if (!INVERSE) {
    for (uint32_t i = threadIdx.x; i < blk; i += NTTL_T) xk[i] = s[i];
  } else {
    for (uint32_t i = threadIdx.x; i < blk; i += NTTL_T) xk[i] = scale ? mont_mul(s[i], scale, P, pinv) : s[i];
  }
}

// The middle of a product in ONE kernel: the forward transform's local stages, the pointwise product and the inverse
// transform's local stages all work on the same contiguous blocks of NTT_B elements (DIF leaves a block bit-reversed,
// DIT takes it bit-reversed), so a block makes one trip through shared memory instead of three kernels and three trips
// through L2:  x_block <- iNTT_local( NTT_local(x_block) .* (b ? b_block : NTT_local(x_block)) ).  b = the other operand's
// forward transform (nullptr: a square).  Same operations in the same order as the three kernels: identical results.
__global__ void __launch_bounds__(NTTL_T) k_ntt_local_mul(uint32_t *x, const uint32_t *__restrict__ b, uint32_t n,
                                                          const uint32_t *__restrict__ twloc, const uint32_t *__restrict__ twiloc,
                                                          uint32_t sc0, uint32_t sc1, uint32_t sc2) {
  __shared__ uint32_t s[NTT_B];
  __shared__ uint32_t twl[NTT_B / 2];
  const int k = blockIdx.y;
  const uint32_t P = c_P[k], pinv = c_PINV[k];
  const uint32_t scale = k == 0 ? sc0 : k == 1 ? sc1 : sc2;
  const uint32_t blk = n < (uint32_t)NTT_B ? n : (uint32_t)NTT_B;
  const uint32_t bh = blk >> 1;
  const size_t base = (size_t)k * n + (size_t)blockIdx.x * blk;
  uint32_t *xk = x + base;
  const uint32_t tstep = (NTT_B / 2) / bh;
  for (uint32_t i = threadIdx.x; i < blk; i += NTTL_T) s[i] = xk[i];
  for (uint32_t e = threadIdx.x; e < bh; e += NTTL_T) twl[e] = twloc[(size_t)k * (NTT_B / 2) + (size_t)e * tstep];
  __syncthreads();
  ntt_block_stages<false>(s, twl, blk, P, pinv);
  for (uint32_t i = threadIdx.x; i < blk; i += NTTL_T) {
    const uint32_t v = s[i];
    s[i] = mont_mul(v, b ? b[base + i] : v, P, pinv);
  }
  for (uint32_t e = threadIdx.x; e < bh; e += NTTL_T) twl[e] = twiloc[(size_t)k * (NTT_B / 2) + (size_t)e * tstep];
  __syncthreads();
  ntt_block_stages<true>(s, twl, blk, P, pinv);
  for (uint32_t i = threadIdx.x; i < blk; i += NTTL_T) xk[i] = scale ? mont_mul(s[i], scale, P, pinv) : s[i];
}

// The stages with span >= NTT_B as ONE kernel ("four-step" form).  View x as m = n / NTT_B rows of c = NTT_B columns
// (element i = l + c*k).  The first log2(m) DIF stages of the size-n transform are, for every column l, a plain
// length-m DIF transform over the rows followed by a multiplication of row k' (output index p = bitrev(k')) by
// w_n^(l*p); k_ntt_local then transforms the rows.  A CTA keeps m x C elements (C consecutive columns) in shared
// memory, so x is read and written ONCE instead of once per stage, the row twiddles (m/2 powers of w_m) sit in
// shared memory and the column factors are generated by repeated multiplication from two table entries per thread.
// The inverse is the mirror image (multiply by w_n^-(l*p), DIT over the rows, scale by 1/n).
// The results are identical — value and position — to the stage-by-stage radix-2 kernels this replaces.
constexpr int NTTC_T = 512;
__device__ __forceinline__ uint32_t ntt_bitrev(uint32_t v, int bits) { return bits ? __brev(v) >> (32 - bits) : 0; }
// w_n^e for e < n from the table of w_nmax^j, j < nmax/2 (w^(n/2) = -1)
__device__ __forceinline__ uint32_t ntt_root_pow(const uint32_t *__restrict__ twk, uint32_t e, uint32_t n, uint32_t nmax_half,
                                                 uint32_t P) {
  const uint32_t step = 2 * nmax_half / n;
  const uint32_t h = n >> 1;
  const uint32_t v = twk[(size_t)(e & (h - 1)) * step];
  return (e & h) ? (v ? P - v : 0) : v;
}

template <bool INVERSE>
__global__ void __launch_bounds__(NTTC_T) k_ntt_cols(uint32_t *x, uint32_t n, uint32_t m, int logm, int logC,
                                                     const uint32_t *__restrict__ tw, uint32_t nmax_half, uint32_t sc0,
                                                     uint32_t sc1, uint32_t sc2) {
  extern __shared__ uint32_t sh[];
  uint32_t *s = sh;              // [m][C]
  uint32_t *twm = sh + ((size_t)m << logC);  // [m/2]: w_m^e
  const int k = blockIdx.y;
  const uint32_t P = c_P[k], pinv = c_PINV[k];
  const uint32_t scale = k == 0 ? sc0 : k == 1 ? sc1 : sc2;
  const uint32_t C = 1u << logC, Cm = C - 1;    // columns per tile: a power of two, so t / C and t % C are a shift and a mask
  const uint32_t c = n / m;                     // columns of the whole array (= NTT_B)
  const uint32_t l0 = blockIdx.x * C;           // first column of this tile
  uint32_t *xk = x + (size_t)k * n;
  const uint32_t *twk = tw + (size_t)k * nmax_half;
  const uint32_t mh = m >> 1;
  for (uint32_t t = threadIdx.x; t < m * C; t += NTTC_T) s[t] = xk[(size_t)(t >> logC) * c + l0 + (t & Cm)];
  for (uint32_t e = threadIdx.x; e < mh; e += NTTC_T) twm[e] = twk[(size_t)e * (nmax_half / mh)];
  __syncthreads();

  // column factors: thread (lc, g) walks output indices p = g, g + G, g + 2G, ... of column l0 + lc
  auto col_factors = [&]() {
    const uint32_t G = NTTC_T >> logC ? NTTC_T >> logC : 1;  // rows covered per sweep (C <= NTTC_T: see the launcher)
    const uint32_t lc = threadIdx.x & Cm, g = threadIdx.x >> logC;
    if (g >= G || g >= m) return;
    const uint32_t l = l0 + lc;
    uint32_t f = ntt_root_pow(twk, (uint32_t)(((uint64_t)l * g) & (n - 1)), n, nmax_half, P);       // w^(l*g)
    const uint32_t r = ntt_root_pow(twk, (uint32_t)(((uint64_t)l * G) & (n - 1)), n, nmax_half, P);  // w^(l*G)
    for (uint32_t pp = g; pp < m; pp += G) {
      uint32_t *e = s + (ntt_bitrev(pp, logm) << logC) + lc;
      *e = mont_mul(*e, f, P, pinv);
      f = mont_mul(f, r, P, pinv);
    }
  };

  // row transforms: two radix-2 stages per pass in registers (rows i, i + h/2, i + h, i + 3h/2 of one column), as in
  // k_ntt_local; an odd stage count leaves one plain stage (half = 1)
  auto radix2 = [&](uint32_t half) {
    const uint32_t ts = mh / half;
    for (uint32_t t = threadIdx.x; t < mh * C; t += NTTC_T) {
      const uint32_t bt = t >> logC, lc = t & Cm;
      const uint32_t j = bt & (half - 1), i = ((bt - j) << 1) + j;
      uint32_t *pa = s + (i << logC) + lc, *pb = pa + (half << logC);
      const uint32_t a = *pa, b = *pb;
      if (!INVERSE) {
        *pa = add_mod(a, b, P);
        *pb = mont_mul(sub_mod(a, b, P), twm[j * ts], P, pinv);
      } else {
        const uint32_t bw = mont_mul(b, twm[j * ts], P, pinv);
        *pa = add_mod(a, bw, P);
        *pb = sub_mod(a, bw, P);
      }
    }
    __syncthreads();
  };
  auto radix4 = [&](uint32_t h) {
    const uint32_t q4 = h >> 1, ts = mh / h;
    for (uint32_t t = threadIdx.x; t < (m >> 2) * C; t += NTTC_T) {
      const uint32_t bt = t >> logC, lc = t & Cm;
      const uint32_t j = bt & (q4 - 1), i = ((bt - j) << 2) + j;
      const uint32_t t1 = twm[j * ts], t2 = twm[(j + q4) * ts], t3 = twm[2 * j * ts];
      uint32_t *p0 = s + (i << logC) + lc, *p1 = p0 + (q4 << logC), *p2 = p0 + (h << logC), *p3 = p2 + (q4 << logC);
      uint32_t a0 = *p0, a1 = *p1, a2 = *p2, a3 = *p3;
      if (!INVERSE) {
        const uint32_t b0 = add_mod(a0, a2, P), b2 = mont_mul(sub_mod(a0, a2, P), t1, P, pinv);
        const uint32_t b1 = add_mod(a1, a3, P), b3 = mont_mul(sub_mod(a1, a3, P), t2, P, pinv);
        a0 = add_mod(b0, b1, P);
        a1 = mont_mul(sub_mod(b0, b1, P), t3, P, pinv);
        a2 = add_mod(b2, b3, P);
        a3 = mont_mul(sub_mod(b2, b3, P), t3, P, pinv);
      } else {
        const uint32_t u1 = mont_mul(a1, t3, P, pinv), u3 = mont_mul(a3, t3, P, pinv);
        const uint32_t b0 = add_mod(a0, u1, P), b1 = sub_mod(a0, u1, P);
        const uint32_t b2 = add_mod(a2, u3, P), b3 = sub_mod(a2, u3, P);
        const uint32_t v2 = mont_mul(b2, t1, P, pinv), v3 = mont_mul(b3, t2, P, pinv);
        a0 = add_mod(b0, v2, P);
        a2 = sub_mod(b0, v2, P);
        a1 = add_mod(b1, v3, P);
        a3 = sub_mod(b1, v3, P);
      }
      *p0 = a0;
      *p1 = a1;
      *p2 = a2;
      *p3 = a3;
    }
    __syncthreads();
  };

  if (!INVERSE) {
    uint32_t h = mh;
    for (; h >= 2; h >>= 2) radix4(h);
    if (h == 1) radix2(1);
    col_factors();
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < m * C; t += NTTC_T) xk[(size_t)(t >> logC) * c + l0 + (t & Cm)] = s[t];
  } else {
    col_factors();  // tw is the inverse table here: w_n^-(l*p)
    __syncthreads();
    uint32_t h = 2;
    if (logm & 1) {  // odd stage count: the first one alone
      radix2(1);
      h = 4;
    }
    for (; h <= mh; h <<= 2) radix4(h);
    for (uint32_t t = threadIdx.x; t < m * C; t += NTTC_T)
      xk[(size_t)(t >> logC) * c + l0 + (t & Cm)] = mont_mul(s[t], scale, P, pinv);
  }
}

// lift: out[k][i] = Montgomery(in[i] mod P_k) for i < len (in: canonical residues mod p as u32), 0 for len <= i < n.
// reverse != 0 reads in[src_len - 1 - i] (polynomial reversal).
__global__ void k_lift(const uint32_t *__restrict__ in, uint32_t len, uint32_t src_len, int reverse, uint32_t *out, uint32_t n) {
  const int k = blockIdx.y;
  const uint32_t P = c_P[k], pinv = c_PINV[k], r2 = c_R2[k];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t v = 0;
    if (i < len) {
      v = reverse ? in[src_len - 1 - i] : in[i];
      v = mont_mul(v % P, r2, P, pinv);
    }
    out[(size_t)k * n + i] = v;
  }
}

struct CrtConst {
  uint32_t inv_p1_p2;    // P1^-1 mod P2, Montgomery form mod P2
  uint32_t inv_p1p2_p3;  // (P1 P2)^-1 mod P3, Montgomery form mod P3
  uint32_t p1_mod_p3;
  uint32_t p1_mod_p, p1p2_mod_p;
};
constexpr uint32_t FP = P;  // 2^32 - 5

// Garner: x = x1 + x2*P1 + x3*P1*P2 (the exact integer coefficient), reduced mod p.  res: [3][n] in Montgomery form.
// out[i] for i < len; mode 0: plain; mode 1: out[i] = (2*[i==0] - x) mod p  (Newton's 2 - f*g);
// `lo` skips the first coefficients: out[i - lo] = coefficient i (used to take the upper half of a product).
__global__ void k_crt(const uint32_t *__restrict__ res, uint32_t n, uint32_t lo, uint32_t len, CrtConst c, int mode,
                      uint32_t *out) {
  const uint32_t P1 = c_P[0], P2 = c_P[1], P3 = c_P[2];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
    const uint32_t idx = lo + i;
    // leave Montgomery form: mont_mul(x, 1)
    const uint32_t x1 = mont_mul(res[idx], 1, P1, c_PINV[0]);
    const uint32_t y2 = mont_mul(res[(size_t)n + idx], 1, P2, c_PINV[1]);
    const uint32_t y3 = mont_mul(res[2 * (size_t)n + idx], 1, P3, c_PINV[2]);
    const uint32_t x2 = mont_mul(sub_mod(y2, x1 % P2, P2), c.inv_p1_p2, P2, c_PINV[1]);  // (y2 - x1) / P1 mod P2
    const uint32_t part = (uint32_t)(((uint64_t)x2 * c.p1_mod_p3 + x1 % P3) % P3);         // x1 + x2 P1 mod P3
    const uint32_t x3 = mont_mul(sub_mod(y3, part, P3), c.inv_p1p2_p3, P3, c_PINV[2]);
    // value mod p: three terms, each < 2^64, summed in 128-bit-free form by reducing as we go
    uint64_t v = x1 % FP;
    v += (uint64_t)(x2 % FP) * c.p1_mod_p % FP;
    v += (uint64_t)(x3 % FP) * c.p1p2_mod_p % FP;
    uint32_t r = (uint32_t)(v % FP);
    if (mode == 1) {
      r = r ? FP - r : 0;
      if (idx == 0) r = (uint32_t)(((uint64_t)r + 2) % FP);
    }
    out[i] = r;
  }
}

// ---- prover-specific element-wise kernels ------------------------------------------------------------------
// w[c] = (delta * t[c] + sum_k sel[k][c]) mod p ; sel are u64 wire coefficients (reduced mod p on import, ssp.c:28-34)
__global__ void k_ssp_accumulate(const uint64_t *__restrict__ t, uint64_t delta, const uint64_t *__restrict__ sel,
                                 uint32_t nsel, uint32_t D, int first_batch, uint32_t *w) {
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < D; c += gridDim.x * blockDim.x) {
    uint64_t acc = first_batch ? (t[c] % FP) * (delta % FP) % FP : w[c];
    for (uint32_t k = 0; k < nsel; k++) {
      acc += sel[(size_t)k * D + c] % FP;
      if (acc >= ((uint64_t)1 << 63)) acc %= FP;
    }
    w[c] = (uint32_t)(acc % FP);
  }
}
// resident SSP (u32 residues, polynomial k at blob[k*D .. (k+1)*D)): w = delta*t + sum of the selected polynomials, v = w + v_0
// hdr = [number of selected polynomials, delta (< p), their indices ...]: everything that changes from proof to proof sits
// in device memory, so that the launch parameters are constants of the instance (the step is replayed as a CUDA graph)
__global__ void k_ssp_accumulate_res(const uint32_t *__restrict__ blob, uint32_t D, const uint32_t *__restrict__ hdr,
                                     uint32_t *w, uint32_t *v) {
  const uint32_t nsel = hdr[0];
  const uint64_t delta = hdr[1];
  const uint32_t *__restrict__ sel = hdr + 2;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < D; c += gridDim.x * blockDim.x) {
    uint64_t acc = (uint64_t)blob[c] * (delta % FP) % FP;
    for (uint32_t k = 0; k < nsel; k++) acc += blob[(size_t)sel[k] * D + c];  // nsel < 2^31 terms below 2^32
    const uint32_t wc = (uint32_t)(acc % FP);
    w[c] = wc;
    v[c] = (uint32_t)(((uint64_t)wc + blob[(size_t)D + c]) % FP);
  }
}
__global__ void k_add_u64poly(const uint32_t *__restrict__ a, const uint64_t *__restrict__ b, uint32_t D, uint32_t *out) {
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < D; c += gridDim.x * blockDim.x)
    out[c] = (uint32_t)(((uint64_t)a[c] + b[c] % FP) % FP);
}
__global__ void k_reduce_u64poly(const uint64_t *__restrict__ b, uint32_t D, uint32_t *out) {
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < D; c += gridDim.x * blockDim.x) out[c] = (uint32_t)(b[c] % FP);
}
// *len = 1 + index of the highest non-zero coefficient (0 for the zero polynomial); *len must be 0 on entry
__global__ void k_poly_length(const uint32_t *__restrict__ a, uint32_t n, uint32_t *len) {
  uint32_t best = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (a[i]) best = i + 1;
  best = __reduce_max_sync(0xffffffffu, best);
  if ((threadIdx.x & 31) == 0 && best) atomicMax(len, best);
}
__global__ void k_sub_one(uint32_t *a) {
  if (threadIdx.x == 0 && blockIdx.x == 0) a[0] = a[0] ? a[0] - 1 : FP - 1;
}
__global__ void k_set_inv0(const uint32_t *f, uint32_t *g) {  // g[0] = f[0]^(p-2) mod p
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    uint64_t base = f[0], acc = 1;
    for (uint32_t e = FP - 2; e; e >>= 1) {
      if (e & 1) acc = acc * base % FP;
      base = base * base % FP;
    }
    g[0] = (uint32_t)acc;
  }
}
// out[i] = in[len - 1 - i] for i < len, zero up to n (reverse + pad), u32 -> u32
__global__ void k_reverse(const uint32_t *__restrict__ in, uint32_t len, uint32_t *out, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = i < len ? in[len - 1 - i] : 0;
}
__global__ void k_widen(const uint32_t *__restrict__ in, uint32_t n, uint64_t *out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[i];
}

// values[q] = poly_q(x) mod p for q < npoly; polys are [npoly][D] u64 wire coefficients; pw[i] = x^i mod p
__global__ void k_powers(uint64_t x, uint32_t D, uint32_t *pw) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < D; i += gridDim.x * blockDim.x) {
    uint64_t base = x % FP, acc = 1;
    for (uint32_t e = i; e; e >>= 1) {
      if (e & 1) acc = acc * base % FP;
      base = base * base % FP;
    }
    pw[i] = (uint32_t)acc;
  }
}
__global__ void __launch_bounds__(256) k_eval(const uint64_t *__restrict__ polys, uint32_t D, const uint32_t *__restrict__ pw,
                                              uint64_t *values) {
  __shared__ unsigned long long part[8];
  const uint64_t *p = polys + (size_t)blockIdx.x * D;
  unsigned long long acc = 0;  // D * p < 2^54 for D <= 2^22
  for (uint32_t i = threadIdx.x; i < D; i += 256) acc += (p[i] % FP) * pw[i] % FP;
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s = 0;
    for (int w = 0; w < 8; w++) s += part[w];
    values[blockIdx.x] = s % FP;
  }
}

// the same over the u32 residues of a resident SSP blob
__global__ void __launch_bounds__(256) k_eval32(const uint32_t *__restrict__ polys, uint32_t D, const uint32_t *__restrict__ pw,
                                                uint64_t *values) {
  __shared__ unsigned long long part[8];
  const uint32_t *p = polys + (size_t)blockIdx.x * D;
  unsigned long long acc = 0;
  for (uint32_t i = threadIdx.x; i < D; i += 256) acc += (unsigned long long)p[i] * pw[i] % FP;
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s = 0;
    for (int w = 0; w < 8; w++) s += part[w];
    values[blockIdx.x] = s % FP;
  }
}

// =============================================================================================== host driver
struct PolyEngine {
  uint32_t nmax = 0;  // twiddle tables cover transforms up to this size
  uint32_t *tw = nullptr, *twi = nullptr;
  uint32_t *tw_loc = nullptr, *twi_loc = nullptr;  // [3][NTT_B / 2]: powers of w_NTT_B, contiguous (k_ntt_local)
  uint32_t *fa = nullptr, *fb = nullptr;  // [3][nmax] work arrays
  uint32_t *t1 = nullptr, *t2 = nullptr, *t3 = nullptr;  // coefficient scratch (u32, nmax each)
  uint32_t *d_len = nullptr;
  CrtConst crt;
  bool consts_ready = false;
  uint64_t launches = 0;
};

static inline unsigned gridfor(uint32_t n, int threads = 256) {
  unsigned g = (n + threads - 1) / threads;
  return g > 1184 ? 1184 : (g ? g : 1);
}

static cudaError_t engine_consts(PolyEngine &E) {
  if (E.consts_ready) return cudaSuccess;
  uint32_t pinv[NPR], r2[NPR];
  for (int k = 0; k < NPR; k++) {
    uint32_t inv = 1;
    for (int i = 0; i < 5; i++) inv *= 2 - h_P[k] * inv;  // P^-1 mod 2^32 (Newton)
    pinv[k] = (uint32_t)(0u - inv);
    r2[k] = (uint32_t)(((unsigned __int128)1 << 64) % h_P[k]);
  }
  cudaError_t e;
  if ((e = cudaMemcpyToSymbol(c_PINV, pinv, sizeof(pinv))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_R2, r2, sizeof(r2))) != cudaSuccess) return e;
  const uint64_t P1 = h_P[0], P2 = h_P[1], P3 = h_P[2];
  auto to_mont = [](uint32_t v, uint32_t m) { return (uint32_t)(((uint64_t)v << 32) % m); };
  E.crt.inv_p1_p2 = to_mont(h_powmod((uint32_t)(P1 % P2), P2 - 2, (uint32_t)P2), (uint32_t)P2);
  E.crt.inv_p1p2_p3 = to_mont(h_powmod((uint32_t)(P1 * P2 % P3), P3 - 2, (uint32_t)P3), (uint32_t)P3);
  E.crt.p1_mod_p3 = (uint32_t)(P1 % P3);
  E.crt.p1_mod_p = (uint32_t)(P1 % FP);
  E.crt.p1p2_mod_p = (uint32_t)((unsigned __int128)P1 * P2 % FP);
  E.consts_ready = true;
  return cudaSuccess;
}

static void engine_free(PolyEngine &E) {
  cudaFree(E.tw); cudaFree(E.twi); cudaFree(E.tw_loc); cudaFree(E.twi_loc); cudaFree(E.fa); cudaFree(E.fb); cudaFree(E.t1); cudaFree(E.t2); cudaFree(E.t3);
  cudaFree(E.d_len);
  E = PolyEngine();
}

static cudaError_t engine_reserve(PolyEngine &E, uint32_t n, cudaStream_t st) {
  cudaError_t e = engine_consts(E);
  if (e != cudaSuccess) return e;
  if (n < (uint32_t)NTT_B) n = NTT_B;  // the compact table of k_ntt_local holds the powers of w_NTT_B
  if (n <= E.nmax) return cudaSuccess;
  if (n > (1u << 23)) return cudaErrorInvalidValue;  // 2-adicity of the NTT primes
  const bool had_consts = E.consts_ready;
  const CrtConst crt = E.crt;
  engine_free(E);
  E.consts_ready = had_consts;
  E.crt = crt;
  const size_t half = n / 2;
  if ((e = cudaMalloc(&E.tw, NPR * half * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&E.twi, NPR * half * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&E.tw_loc, (size_t)NPR * (NTT_B / 2) * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&E.twi_loc, (size_t)NPR * (NTT_B / 2) * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&E.fa, (size_t)NPR * n * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&E.fb, (size_t)NPR * n * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&E.t1, (size_t)n * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&E.t2, (size_t)n * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&E.t3, (size_t)n * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&E.d_len, 16)) != cudaSuccess) return e;
  uint32_t w[NPR], wi[NPR];
  for (int k = 0; k < NPR; k++) {
    w[k] = h_powmod(h_G[k], (h_P[k] - 1) / n, h_P[k]);
    wi[k] = h_powmod(w[k], h_P[k] - 2, h_P[k]);
  }
  dim3 g(gridfor((uint32_t)half), NPR);
  k_twiddle_fill<<<g, 256, 0, st>>>(E.tw, E.twi, (uint32_t)half, w[0], w[1], w[2], wi[0], wi[1], wi[2]);
  E.launches++;
  k_twiddle_compact<<<dim3(4, NPR), 256, 0, st>>>(E.tw, E.twi, (uint32_t)half, E.tw_loc, E.twi_loc);
  E.launches++;
  E.nmax = n;
  return cudaGetLastError();
}

// in-place transforms of x[3][n]: n <= NTT_B: one shared-memory kernel; else k_ntt_cols + k_ntt_local (two launches)
static void ntt_cols_geometry(uint32_t n, uint32_t *m, int *logm, uint32_t *C, size_t *smem) {
  *m = n / (uint32_t)NTT_B;
  *logm = 0;
  while ((1u << *logm) < *m) (*logm)++;
  uint32_t c = 16384 / *m;          // ~16 K elements (64 KB) per tile, at least 16 columns (64-byte rows)
  if (c < 16) c = 16;
  if (c > (uint32_t)NTTC_T) c = NTTC_T;
  *C = c;
  *smem = ((size_t)*m * c + *m / 2 + 1) * 4;
}
template <bool INVERSE>
static cudaError_t launch_ntt_cols(uint32_t *x, uint32_t n, const uint32_t *tw, uint32_t nh, const uint32_t sc[NPR], cudaStream_t st) {
  uint32_t m, C;
  int logm;
  size_t smem;
  ntt_cols_geometry(n, &m, &logm, &C, &smem);
  if (smem > 48 * 1024) {  // (per device and per kernel; the call is cheap, so no caching)
    cudaError_t e = cudaFuncSetAttribute(k_ntt_cols<INVERSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  int logC = 0;
  while ((1u << logC) < C) logC++;
  k_ntt_cols<INVERSE><<<dim3(NTT_B / C, NPR), NTTC_T, smem, st>>>(x, n, m, logm, logC, tw, nh, sc[0], sc[1], sc[2]);
  return cudaGetLastError();
}
static cudaError_t ntt_forward(PolyEngine &E, uint32_t *x, uint32_t n, cudaStream_t st) {
  const uint32_t nh = E.nmax / 2;
  const uint32_t zero[NPR] = {0, 0, 0};
  if (n > (uint32_t)NTT_B) {
    cudaError_t e = launch_ntt_cols<false>(x, n, E.tw, nh, zero, st);
    if (e != cudaSuccess) return e;
    E.launches++;
  }
  const uint32_t blk = n < (uint32_t)NTT_B ? n : (uint32_t)NTT_B;
  k_ntt_local<false><<<dim3(n / blk, NPR), NTTL_T, 0, st>>>(x, n, E.tw_loc, 0, 0, 0);
  E.launches++;
  return cudaGetLastError();
}
// x <- iNTT( NTT(x) .* (bhat ? bhat : NTT(x)) ): the column stages as their own launches (n > NTT_B), everything between
// them — local forward stages, pointwise product, local inverse stages — as one launch (k_ntt_local_mul)
static cudaError_t ntt_product_inplace(PolyEngine &E, uint32_t *x, const uint32_t *bhat, uint32_t n, cudaStream_t st) {
  const uint32_t nh = E.nmax / 2;
  const uint32_t zero[NPR] = {0, 0, 0};
  uint32_t sc[NPR];
  for (int k = 0; k < NPR; k++) {
    const uint32_t ninv = h_powmod(n % h_P[k], h_P[k] - 2, h_P[k]);
    sc[k] = (uint32_t)(((uint64_t)ninv << 32) % h_P[k]);
  }
  const bool cols = n > (uint32_t)NTT_B;
  cudaError_t e;
  if (cols) {
    if ((e = launch_ntt_cols<false>(x, n, E.tw, nh, zero, st)) != cudaSuccess) return e;
    E.launches++;
  }
  const uint32_t blk = n < (uint32_t)NTT_B ? n : (uint32_t)NTT_B;
  k_ntt_local_mul<<<dim3(n / blk, NPR), NTTL_T, 0, st>>>(x, bhat, n, E.tw_loc, E.twi_loc, cols ? 0 : sc[0], cols ? 0 : sc[1],
                                                          cols ? 0 : sc[2]);
  E.launches++;
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (cols) {
    if ((e = launch_ntt_cols<true>(x, n, E.twi, nh, sc, st)) != cudaSuccess) return e;
    E.launches++;
  }
  return cudaGetLastError();
}

// out[0..out_len) = coefficients [lo, lo+out_len) of a*b mod p  (mode 1: of 2 - a*b).
// a: la coefficients (read reversed from a source of a_src_len coefficients when a_rev), b likewise.
static cudaError_t poly_mul(PolyEngine &E, const uint32_t *a, uint32_t la, uint32_t a_src_len, int a_rev, const uint32_t *b,
                            uint32_t lb, uint32_t b_src_len, int b_rev, uint32_t lo, uint32_t out_len, int mode, uint32_t *out,
                            cudaStream_t st) {
  if (la == 0 || lb == 0 || out_len == 0) return cudaSuccess;
  uint32_t n = 1;
  while (n < la + lb - 1) n <<= 1;
  if (n < 2) n = 2;
  if (n > E.nmax) return cudaErrorInvalidValue;
  const bool square = (a == b && la == lb && a_rev == b_rev && a_src_len == b_src_len);
  k_lift<<<dim3(gridfor(n), NPR), 256, 0, st>>>(a, la, a_src_len, a_rev, E.fa, n);
  E.launches++;
  cudaError_t ne;
  if (!square) {
    k_lift<<<dim3(gridfor(n), NPR), 256, 0, st>>>(b, lb, b_src_len, b_rev, E.fb, n);
    E.launches++;
    if ((ne = ntt_forward(E, E.fb, n, st)) != cudaSuccess) return ne;
  }
  if ((ne = ntt_product_inplace(E, E.fa, square ? nullptr : E.fb, n, st)) != cudaSuccess) return ne;
  k_crt<<<gridfor(out_len), 256, 0, st>>>(E.fa, n, lo, out_len, E.crt, mode, out);
  E.launches++;
  return cudaGetLastError();
}

// The same with the second operand given as its forward transform of n points (bhat[3][n], Montgomery form): an operand
// that does not change between calls (the cached rev(t)^-1 of a resident SSP) is lifted and transformed once.
static cudaError_t poly_mul_bhat(PolyEngine &E, const uint32_t *a, uint32_t la, uint32_t a_src_len, int a_rev, const uint32_t *bhat,
                                 uint32_t n, uint32_t lo, uint32_t out_len, uint32_t *out, cudaStream_t st) {
  if (n > E.nmax) return cudaErrorInvalidValue;
  k_lift<<<dim3(gridfor(n), NPR), 256, 0, st>>>(a, la, a_src_len, a_rev, E.fa, n);
  E.launches++;
  cudaError_t ne = ntt_product_inplace(E, E.fa, bhat, n, st);
  if (ne != cudaSuccess) return ne;
  k_crt<<<gridfor(out_len), 256, 0, st>>>(E.fa, n, lo, out_len, E.crt, 0, out);
  E.launches++;
  return cudaGetLastError();
}

// g[0..m) = f^-1 mod x^m for f of lf coefficients (f[0] != 0), Newton: g <- g * (2 - f g) mod x^2k
static cudaError_t poly_inv_series(PolyEngine &E, const uint32_t *f, uint32_t lf, uint32_t m, uint32_t *g, cudaStream_t st) {
  k_set_inv0<<<1, 32, 0, st>>>(f, g);
  E.launches++;
  for (uint32_t k = 1; k < m;) {
    const uint32_t k2 = 2 * k < m ? 2 * k : m;
    const uint32_t lfk = lf < k2 ? lf : k2;
    // t1 = 2 - f*g mod x^k2   (first k coefficients are 1, 0, 0, ... by construction; all are recomputed)
    cudaError_t e = poly_mul(E, f, lfk, lfk, 0, g, k, k, 0, 0, k2, 1, E.t1, st);
    if (e != cudaSuccess) return e;
    // g[k..k2) = (g * t1)[k..k2)
    uint32_t l1 = lfk + k - 1;
    if (l1 > k2) l1 = k2;
    e = poly_mul(E, g, k, k, 0, E.t1, l1, l1, 0, k, k2 - k, 0, g + k, st);
    if (e != cudaSuccess) return e;
    k = k2;
  }
  return cudaGetLastError();
}

// q_out[0..out_len) = rev( rev(a)[0..lq) * binv[0..lq) mod x^lq ) for binv = rev(b)^-1 mod x^(>= lq); uses E.t3
static cudaError_t poly_quot_from_inverse(PolyEngine &E, const uint32_t *a, uint32_t la, uint32_t lq, const uint32_t *binv,
                                          uint32_t *q_out, uint32_t out_len, cudaStream_t st) {
  cudaError_t e = poly_mul(E, a, lq, la, 1, binv, lq, lq, 0, 0, lq, 0, E.t3, st);
  if (e != cudaSuccess) return e;
  k_reverse<<<gridfor(out_len), 256, 0, st>>>(E.t3, lq, q_out, out_len);
  E.launches++;
  return cudaGetLastError();
}

// binv[0..prec) = rev(b)^-1 mod x^prec (Newton); uses E.t1, E.t3
static cudaError_t poly_rev_inverse(PolyEngine &E, const uint32_t *b, uint32_t lb, uint32_t prec, uint32_t *binv,
                                    cudaStream_t st) {
  k_reverse<<<gridfor(lb), 256, 0, st>>>(b, lb, E.t3, lb);
  E.launches++;
  return poly_inv_series(E, E.t3, lb < prec ? lb : prec, prec, binv, st);
}

// q = a / b (Euclidean quotient): q = rev( rev(a) * rev(b)^-1 mod x^lq ), lq = la - lb + 1; writes q_out[0..out_len)
// (zero padded / truncated).  Uses E.t2 (inverse series) and E.t3 (reversed quotient).
static cudaError_t poly_div(PolyEngine &E, const uint32_t *a, uint32_t la, const uint32_t *b, uint32_t lb, uint32_t *q_out,
                            uint32_t out_len, cudaStream_t st) {
  if (lb == 0) return cudaErrorInvalidValue;
  if (la < lb) return cudaMemsetAsync(q_out, 0, (size_t)out_len * 4, st);
  const uint32_t lq = la - lb + 1;
  cudaError_t e = poly_rev_inverse(E, b, lb, lq, E.t2, st);
  if (e != cudaSuccess) return e;
  return poly_quot_from_inverse(E, a, la, lq, E.t2, q_out, out_len, st);
}

}  // namespace mfb

// ================================================================================================ C-ABI part
#include "../../include/mfb200.h"

namespace mfb {
PolyEngine *poly_engine_of(mfb_ctx *ctx);         // mfb_capi.cu
cudaStream_t ctx_stream_of(mfb_ctx *ctx);
int ctx_scratch(mfb_ctx *ctx, int slot, size_t bytes, void **out);
int ctx_fail(cudaError_t e, const char *what, const char *file, int line);
void ctx_count_launches(mfb_ctx *ctx, uint64_t n);
int ctx_enter(mfb_ctx *ctx);
int ctx_bad_arg(const char *msg);
int ctx_h2d_pieces(mfb_ctx *ctx, void *dst_dev, const void *const *src, size_t piece_bytes, size_t npieces, cudaStream_t st);
int ctx_h2d(mfb_ctx *ctx, void *dst_dev, const void *src, size_t bytes, cudaStream_t st);
void ctx_trace(const char *label);
int ctx_h2d_narrow_u64(mfb_ctx *ctx, uint32_t *dst_dev, const uint64_t *src, size_t count, uint64_t limit, int *narrow_ok,
                       cudaStream_t st);

PolyEngine *poly_engine_new() { return new PolyEngine(); }
void poly_engine_delete(PolyEngine *e) {
  if (e) {
    engine_free(*e);
    delete e;
  }
}
}  // namespace mfb

using namespace mfb;

#define PTRY(expr)                                                          \
  do {                                                                      \
    cudaError_t e_ = (expr);                                                \
    if (e_ != cudaSuccess) return ctx_fail(e_, #expr, __FILE__, __LINE__);  \
  } while (0)

// A dense SSP blob kept on the device as u32 residues, with the Newton inverse of rev(t) cached: it depends on the
// instance only, so every proof after the first divides with one multiplication.
struct mfb_ssp {
  uint32_t *blob = nullptr;  // [(M + 1)][D]: t, v_0, ..., v_{M-1}
  size_t D = 0, M = 0;
  uint32_t lt = 0;           // normalised length of t
  uint32_t *binv = nullptr;  // rev(t)^-1 mod x^binv_len
  uint32_t binv_len = 0, binv_cap = 0;
  // forward transform (bhat_n points, three primes) of rev(t)^-1 mod x^bhat_lq: what the device-resident pipeline divides with
  uint32_t *bhat = nullptr;
  uint32_t bhat_n = 0, bhat_lq = 0;
  // pinned staging of the witness-selected polynomial indices (so that their upload is a true asynchronous copy) and the
  // event after which the staging may be overwritten
  uint32_t *sel_pin = nullptr;
  cudaEvent_t sel_free = nullptr;
  bool sel_used = false;
  // the polynomial step as an instantiated CUDA graph (one launch instead of ~17), valid for the pointers in graph_key
  cudaGraphExec_t gexec = nullptr;
  const void *graph_key[8] = {};
  uint64_t graph_kernels = 0;
  bool graph_off = false;  // capture failed once: plain launches from then on
};

// Host-blob path: v (D residues, device) and t (lt residues, stable device pointer) -> h = (v^2 - 1) / t truncated to D,
// then w | v | h back to the host as u64.  d_w, d_v, d_h are consecutive D-element u32 arrays.
static int polys_tail(mfb_ctx *ctx, PolyEngine &E, cudaStream_t st, uint32_t Du, uint32_t n, uint32_t *d_w, uint32_t *d_a,
                      uint64_t *d_wide, const uint32_t *t32, uint32_t lt, uint64_t *w_out, uint64_t *v_out,
                      uint64_t *h_out) {
  uint32_t *d_v = d_w + Du, *d_h = d_w + 2 * Du;
  uint32_t lv = 0;
  PTRY(cudaMemsetAsync(E.d_len, 0, 16, st));
  k_poly_length<<<gridfor(Du), 256, 0, st>>>(d_v, Du, E.d_len);
  E.launches++;
  PTRY(cudaMemcpyAsync(&lv, E.d_len, 4, cudaMemcpyDeviceToHost, st));
  PTRY(cudaStreamSynchronize(st));

  // a = v^2 - 1
  uint32_t la;
  if (lv == 0) {  // v = 0: a = -1
    PTRY(cudaMemsetAsync(d_a, 0, (size_t)n * 4, st));
    la = 1;
  } else {
    la = 2 * lv - 1;
    PTRY(poly_mul(E, d_v, lv, lv, 0, d_v, lv, lv, 0, 0, la, 0, d_a, st));
  }
  k_sub_one<<<1, 32, 0, st>>>(d_a);
  E.launches++;
  if (la == 1) {  // the constant may have become 0
    uint32_t a0;
    PTRY(cudaMemcpyAsync(&a0, d_a, 4, cudaMemcpyDeviceToHost, st));
    PTRY(cudaStreamSynchronize(st));
    if (a0 == 0) la = 0;
  }
  // h = a / t.  When t has (much) lower degree than D the quotient is longer than D and the transforms of the
  // Newton iteration / of the final product need up to 2*lq points: grow the engine for this call if necessary.
  if (la >= lt) {
    const uint64_t lq64 = (uint64_t)la - lt + 1;
    uint64_t need = 2 * lq64 > lq64 + lq64 / 2 + lt + 8 ? 2 * lq64 : lq64 + lq64 / 2 + lt + 8;
    uint32_t nn = n;
    while (nn < need) nn <<= 1;
    PTRY(engine_reserve(E, nn, st));
  }
  if (la < lt) {
    PTRY(cudaMemsetAsync(d_h, 0, (size_t)Du * 4, st));
  } else {
    PTRY(poly_div(E, d_a, la, t32, lt, d_h, Du, st));
  }
  // back to the host as u64 coefficient arrays (what nmod_poly / eval_poly consume)
  k_widen<<<gridfor(3 * Du), 256, 0, st>>>(d_w, 3 * Du, d_wide);
  E.launches++;
  PTRY(cudaMemcpyAsync(w_out, d_wide, (size_t)Du * 8, cudaMemcpyDeviceToHost, st));
  PTRY(cudaMemcpyAsync(v_out, d_wide + Du, (size_t)Du * 8, cudaMemcpyDeviceToHost, st));
  PTRY(cudaMemcpyAsync(h_out, d_wide + 2 * (size_t)Du, (size_t)Du * 8, cudaMemcpyDeviceToHost, st));
  PTRY(cudaStreamSynchronize(st));
  (void)ctx;
  return MFB_OK;
}

extern "C" int mfb_ssp_prover_polys(mfb_ctx *ctx, const uint64_t *ssp, size_t D, size_t M, const uint64_t *witness_limbs,
                                    size_t nlimbs, uint64_t delta, uint64_t *w_out, uint64_t *v_out, uint64_t *h_out) {
  int rc = ctx_enter(ctx);
  if (rc) return rc;
  if (!ssp || !witness_limbs || !w_out || !v_out || !h_out) return ctx_bad_arg("mfb_ssp_prover_polys: null pointer");
  if (D < 1 || D > (1u << 21) || M < 1) return ctx_bad_arg("mfb_ssp_prover_polys: need 1 <= D <= 2^21, M >= 1");
  PolyEngine &E = *poly_engine_of(ctx);
  cudaStream_t st = ctx_stream_of(ctx);
  const uint64_t L0 = E.launches;
  uint32_t n = 2;
  while (n < 2 * D) n <<= 1;
  PTRY(engine_reserve(E, n, st));

  // selected polynomials: t, v_0, and v_i for the set witness bits (bit i-1 <-> v_i, i = 1..M-1)
  const uint32_t Du = (uint32_t)D;
  // polynomials staged per accumulate launch: up to 256 MB of wire coefficients
  size_t BATCH = ((size_t)256 << 20) / (D * 8);
  if (BATCH < 1) BATCH = 1;
  if (BATCH > M) BATCH = M;
  void *d_t, *d_v0, *d_sel, *d_w, *d_a, *d_wide;
  if ((rc = ctx_scratch(ctx, 0, D * 8, &d_t))) return rc;
  if ((rc = ctx_scratch(ctx, 1, D * 8, &d_v0))) return rc;
  if ((rc = ctx_scratch(ctx, 3, BATCH * D * 8, &d_sel))) return rc;
  if ((rc = ctx_scratch(ctx, 4, (size_t)n * 4, &d_a))) return rc;  // v^2 - 1
  if ((rc = ctx_scratch(ctx, 5, D * 4 * 3, &d_w))) return rc;      // w | v | h
  if ((rc = ctx_scratch(ctx, 6, D * 8 * 3, &d_wide))) return rc;
  uint32_t *d_v = (uint32_t *)d_w + D;
  PTRY(cudaMemcpyAsync(d_t, ssp, D * 8, cudaMemcpyHostToDevice, st));
  PTRY(cudaMemcpyAsync(d_v0, ssp + D, D * 8, cudaMemcpyHostToDevice, st));
  std::vector<const void *> pieces;
  for (size_t i = 1; i < M; i++)
    if ((i - 1) / 64 < nlimbs && (witness_limbs[(i - 1) / 64] >> ((i - 1) % 64) & 1)) pieces.push_back(ssp + (i + 1) * D);
  int first = 1;
  size_t done = 0;
  do {
    const size_t cnt = pieces.size() - done < BATCH ? pieces.size() - done : BATCH;
    // (the staging area is reused in stream order: the previous accumulate has read it before the next copy lands)
    if (cnt && (rc = ctx_h2d_pieces(ctx, d_sel, pieces.data() + done, D * 8, cnt, st))) return rc;
    k_ssp_accumulate<<<gridfor(Du), 256, 0, st>>>((const uint64_t *)d_t, delta, (const uint64_t *)d_sel, (uint32_t)cnt, Du, first,
                                                   (uint32_t *)d_w);
    E.launches++;
    PTRY(cudaGetLastError());
    first = 0;
    done += cnt;
  } while (done < pieces.size());
  k_add_u64poly<<<gridfor(Du), 256, 0, st>>>((const uint32_t *)d_w, (const uint64_t *)d_v0, Du, d_v);
  E.launches++;

  // t as u32 residues in a buffer that survives the division (the staging area is free again), and its length
  uint32_t *t32 = (uint32_t *)d_sel;
  uint32_t lt = 0;
  PTRY(cudaStreamSynchronize(st));  // the last accumulate launch still reads d_sel
  k_reduce_u64poly<<<gridfor(Du), 256, 0, st>>>((const uint64_t *)d_t, Du, t32);
  PTRY(cudaMemsetAsync(E.d_len, 0, 16, st));
  k_poly_length<<<gridfor(Du), 256, 0, st>>>(t32, Du, E.d_len);
  E.launches += 2;
  PTRY(cudaMemcpyAsync(&lt, E.d_len, 4, cudaMemcpyDeviceToHost, st));
  PTRY(cudaStreamSynchronize(st));
  if (lt == 0) return ctx_bad_arg("mfb_ssp_prover_polys: t(x) is the zero polynomial");
  rc = polys_tail(ctx, E, st, Du, n, (uint32_t *)d_w, (uint32_t *)d_a, (uint64_t *)d_wide, t32, lt, w_out, v_out, h_out);
  ctx_count_launches(ctx, E.launches - L0);
  return rc;
}

extern "C" int mfb_ssp_create(mfb_ctx *ctx, const uint64_t *ssp, size_t D, size_t M, mfb_ssp **out) {
  int rc = ctx_enter(ctx);
  if (rc) return rc;
  if (!ssp || !out) return ctx_bad_arg("mfb_ssp_create: null pointer");
  *out = nullptr;
  if (D < 1 || D > (1u << 21) || M < 1) return ctx_bad_arg("mfb_ssp_create: need 1 <= D <= 2^21, M >= 1");
  PolyEngine &E = *poly_engine_of(ctx);
  cudaStream_t st = ctx_stream_of(ctx);
  const uint64_t L0 = E.launches;
  uint32_t n = 2;
  while (n < 2 * D) n <<= 1;
  ctx_trace("ssp_create: enter");
  // (twice the product size: the Newton inversion of rev(t) at the first proof needs transforms of up to 2.5 D + 8 points —
  // reserving that now saves it from freeing and re-allocating the engine's ten buffers inside the first prover() call)
  PTRY(engine_reserve(E, n <= (1u << 22) ? 2 * n : n, st));
  ctx_trace("ssp_create: transform engine reserved");
  mfb_ssp *h = new mfb_ssp();
  h->D = D;
  h->M = M;
  const size_t total = (M + 1) * D;
  cudaError_t e = cudaMalloc(&h->blob, total * 4);
  ctx_trace("ssp_create: blob allocated");
  if (e != cudaSuccess) {
    delete h;
    return ctx_fail(e, "cudaMalloc of the resident SSP", __FILE__, __LINE__);
  }
  // The reference's blob stores residues < p in 8 bytes each (ssp.h:6-9): the host threads that pack the pinned bounce
  // buffers narrow them to u32 on the way, so half the bytes cross PCIe and no device pass is needed.  A blob with a value
  // >= p somewhere (legal for this API: coefficients are reduced mod p) takes the full-width path below from the start.
  int narrow_ok = 0;
  if (total >= ((size_t)1 << 20)) {
    rc = ctx_h2d_narrow_u64(ctx, h->blob, ssp, total, (uint64_t)FP, &narrow_ok, st);
    if (rc != MFB_OK) {
      cudaStreamSynchronize(st);
      cudaFree(h->blob);
      delete h;
      return rc;
    }
  }
  ctx_trace("ssp_create: narrow upload queued");
  const size_t STAGE = (size_t)32 << 20;  // u64 coefficients per upload batch (256 MB)
  void *d_stage = nullptr;
  if (!narrow_ok) rc = ctx_scratch(ctx, 3, (total < STAGE ? total : STAGE) * 8, &d_stage);
  for (size_t o = 0; !narrow_ok && rc == MFB_OK && o < total; o += STAGE) {
    const size_t cnt = total - o < STAGE ? total - o : STAGE;
    if ((rc = ctx_h2d(ctx, d_stage, ssp + o, cnt * 8, st)) != MFB_OK) break;
    // (no synchronisation between batches: the staging area is reused in stream order, and the host packs the next
    // batch into the pinned bounce buffers while this one is on the DMA engine)
    k_reduce_u64poly<<<gridfor((uint32_t)cnt), 256, 0, st>>>((const uint64_t *)d_stage, (uint32_t)cnt, h->blob + o);
    E.launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) break;
  }
  if (rc == MFB_OK && e == cudaSuccess) {
    e = cudaMemsetAsync(E.d_len, 0, 16, st);
    k_poly_length<<<gridfor((uint32_t)D), 256, 0, st>>>(h->blob, (uint32_t)D, E.d_len);
    E.launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h->lt, E.d_len, 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  }
  ctx_trace("ssp_create: upload complete");
  ctx_count_launches(ctx, E.launches - L0);
  if (rc != MFB_OK || e != cudaSuccess) {
    cudaFree(h->blob);
    delete h;
    return rc != MFB_OK ? rc : ctx_fail(e, "resident SSP upload", __FILE__, __LINE__);
  }
  *out = h;
  return MFB_OK;
}

extern "C" void mfb_ssp_destroy(mfb_ctx *ctx, mfb_ssp *h) {
  if (!h) return;
  if (ctx) ctx_enter(ctx);
  cudaFree(h->blob);
  cudaFree(h->binv);
  cudaFree(h->bhat);
  if (h->sel_pin) cudaFreeHost(h->sel_pin);
  if (h->sel_free) cudaEventDestroy(h->sel_free);
  if (h->gexec) cudaGraphExecDestroy(h->gexec);
  delete h;
}

// The polynomial step over a RESIDENT blob, queued on `st` without any host round trip: *wvh = w | v | h as three
// consecutive arrays of D u32 residues in the context's scratch.
//   w = delta t + sum of the selected v_i, v = w + v_0            one kernel over the blob
//   a = v^2 - 1                                                  one product of 2D - 1 coefficients
//   h = a / t = rev( rev(a) * rev(t)^-1 mod x^lq ), lq = 2D - lt one product against the CACHED TRANSFORM of rev(t)^-1
// a is taken with its nominal length 2D - 1 (leading zeros when v has lower degree: the Euclidean quotient does not
// change, its top coefficients are zero), so that no length has to be read back from the device and rev(t)^-1 and its
// transform depend on the instance only.
static int polys_resident_queue(mfb_ctx *ctx, mfb_ssp *h, const uint64_t *witness_limbs, size_t nlimbs, uint64_t delta,
                                cudaStream_t st, uint32_t **wvh) {
  if (h->lt == 0) return ctx_bad_arg("mfb_ssp_prover_polys_resident: t(x) is the zero polynomial");
  PolyEngine &E = *poly_engine_of(ctx);
  const size_t D = h->D, M = h->M;
  const uint32_t Du = (uint32_t)D, la = 2 * Du - 1, lq = la - h->lt + 1;
  uint32_t n = 2, n2 = 2;
  while (n < 2 * D) n <<= 1;
  while (n2 < 2 * lq - 1) n2 <<= 1;
  {  // transforms of the Newton iteration / of the quotient product need up to 2 lq points
    const uint64_t need = 2 * (uint64_t)lq > (uint64_t)lq + lq / 2 + h->lt + 8 ? 2 * (uint64_t)lq : (uint64_t)lq + lq / 2 + h->lt + 8;
    uint32_t nn = n > n2 ? n : n2;
    while (nn < need) nn <<= 1;
    PTRY(engine_reserve(E, nn, st));
  }
  int rc;
  void *d_idx, *d_w, *d_a;
  if ((rc = ctx_scratch(ctx, 0, (M + 2) * 4, &d_idx))) return rc;
  if ((rc = ctx_scratch(ctx, 4, (size_t)n * 4, &d_a))) return rc;
  if ((rc = ctx_scratch(ctx, 5, D * 4 * 3, &d_w))) return rc;
  uint32_t *dw = (uint32_t *)d_w, *dv = dw + D, *dh = dw + 2 * D;
  if (h->bhat_lq != lq || h->bhat_n != n2) {  // first proof over this blob: Newton inverse of rev(t), then its transform
    if (h->binv_cap < lq) {
      if (h->binv) PTRY(cudaFree(h->binv));
      h->binv = nullptr;
      h->binv_cap = h->binv_len = 0;
      PTRY(cudaMalloc(&h->binv, (size_t)lq * 4));
      h->binv_cap = lq;
    }
    if (h->binv_len < lq) {
      PTRY(poly_rev_inverse(E, h->blob, h->lt, h->binv_cap, h->binv, st));
      h->binv_len = h->binv_cap;
    }
    if (h->bhat) PTRY(cudaFree(h->bhat));
    h->bhat = nullptr;
    h->bhat_n = h->bhat_lq = 0;
    PTRY(cudaMalloc(&h->bhat, (size_t)NPR * n2 * 4));
    k_lift<<<dim3(gridfor(n2), NPR), 256, 0, st>>>(h->binv, lq, lq, 0, h->bhat, n2);
    E.launches++;
    PTRY(ntt_forward(E, h->bhat, n2, st));
    h->bhat_n = n2;
    h->bhat_lq = lq;
  }
  // witness bit i-1 selects v_i = polynomial i+1 of the blob; header + index list through pinned staging
  const size_t hdr_bytes = (M + 2) * 4;
  if (!h->sel_pin) PTRY(cudaHostAlloc((void **)&h->sel_pin, hdr_bytes, cudaHostAllocDefault));
  if (!h->sel_free) PTRY(cudaEventCreateWithFlags(&h->sel_free, cudaEventDisableTiming));
  if (h->sel_used) PTRY(cudaEventSynchronize(h->sel_free));  // the previous proof's upload has read the staging
  uint32_t nsel = 0;
  for (size_t i = 1; i < M; i++)
    if ((i - 1) / 64 < nlimbs && (witness_limbs[(i - 1) / 64] >> ((i - 1) % 64) & 1)) h->sel_pin[2 + nsel++] = (uint32_t)(i + 1);
  h->sel_pin[0] = nsel;
  h->sel_pin[1] = (uint32_t)(delta % FP);
  // the step itself: upload, w | v, a = v^2 - 1, h = a / t.  Every launch parameter is a constant of (instance, buffers).
  auto body = [&](cudaStream_t s2) -> int {
    PTRY(cudaMemcpyAsync(d_idx, h->sel_pin, hdr_bytes, cudaMemcpyHostToDevice, s2));
    k_ssp_accumulate_res<<<gridfor(Du), 256, 0, s2>>>(h->blob, Du, (const uint32_t *)d_idx, dw, dv);
    E.launches++;
    PTRY(cudaGetLastError());
    PTRY(poly_mul(E, dv, Du, Du, 0, dv, Du, Du, 0, 0, la, 0, (uint32_t *)d_a, s2));
    k_sub_one<<<1, 32, 0, s2>>>((uint32_t *)d_a);
    E.launches++;
    PTRY(poly_mul_bhat(E, (const uint32_t *)d_a, lq, la, 1, h->bhat, n2, 0, lq, E.t3, s2));
    k_reverse<<<gridfor(Du), 256, 0, s2>>>(E.t3, lq, dh, Du);
    E.launches++;
    PTRY(cudaGetLastError());
    return MFB_OK;
  };
  bool launched = false;
  if (st != nullptr && !h->graph_off && !getenv("MFB_NO_GRAPHS")) {
    const void *key[8] = {d_idx, d_a, d_w, E.fa, E.t3, h->bhat, h->sel_pin, (const void *)(uintptr_t)E.nmax};
    if (!h->gexec || memcmp(key, h->graph_key, sizeof(key)) != 0) {  // (re)capture: first replayable proof, or a buffer moved
      if (h->gexec) cudaGraphExecDestroy(h->gexec);
      h->gexec = nullptr;
      cudaGraph_t g = nullptr;
      const uint64_t l0 = E.launches;
      bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      if (ok) {
        const int brc = body(st);
        ok = cudaStreamEndCapture(st, &g) == cudaSuccess && brc == MFB_OK && g != nullptr;
      }
      h->graph_kernels = E.launches - l0;
      E.launches = l0;  // (captured, not run)
      if (ok) ok = cudaGraphInstantiate(&h->gexec, g, 0) == cudaSuccess;
      if (g) cudaGraphDestroy(g);
      if (!ok) {
        cudaGetLastError();
        h->gexec = nullptr;
        h->graph_off = true;
      } else {
        memcpy(h->graph_key, key, sizeof(key));
      }
    }
    if (h->gexec) {
      PTRY(cudaGraphLaunch(h->gexec, st));
      E.launches += h->graph_kernels;
      launched = true;
    }
  }
  if (!launched && (rc = body(st))) return rc;
  PTRY(cudaEventRecord(h->sel_free, st));
  h->sel_used = true;
  *wvh = dw;
  return MFB_OK;
}

extern "C" int mfb_ssp_prover_polys_resident(mfb_ctx *ctx, mfb_ssp *h, const uint64_t *witness_limbs, size_t nlimbs,
                                             uint64_t delta, uint64_t *w_out, uint64_t *v_out, uint64_t *h_out) {
  int rc = ctx_enter(ctx);
  if (rc) return rc;
  if (!h || !witness_limbs || !w_out || !v_out || !h_out) return ctx_bad_arg("mfb_ssp_prover_polys_resident: null pointer");
  PolyEngine &E = *poly_engine_of(ctx);
  cudaStream_t st = ctx_stream_of(ctx);
  const uint64_t L0 = E.launches;
  const size_t D = h->D;
  void *d_wide;
  if ((rc = ctx_scratch(ctx, 6, D * 8 * 3, &d_wide))) return rc;
  uint32_t *d_w = nullptr;
  if ((rc = polys_resident_queue(ctx, h, witness_limbs, nlimbs, delta, st, &d_w))) return rc;
  // back to the host as u64 coefficient arrays (what nmod_poly / eval_poly consume)
  k_widen<<<gridfor(3 * (uint32_t)D), 256, 0, st>>>(d_w, 3 * (uint32_t)D, (uint64_t *)d_wide);
  E.launches++;
  PTRY(cudaMemcpyAsync(w_out, d_wide, D * 8, cudaMemcpyDeviceToHost, st));
  PTRY(cudaMemcpyAsync(v_out, (uint64_t *)d_wide + D, D * 8, cudaMemcpyDeviceToHost, st));
  PTRY(cudaMemcpyAsync(h_out, (uint64_t *)d_wide + 2 * D, D * 8, cudaMemcpyDeviceToHost, st));
  PTRY(cudaStreamSynchronize(st));
  ctx_count_launches(ctx, E.launches - L0);
  return MFB_OK;
}

// The same with the results left ON THE DEVICE, queued on `stream` WITHOUT waiting: *wvh_dev = three consecutive arrays of
// D u32 residues (w, v, h) in the context's scratch, valid (in stream order) until the next polynomial / encrypt /
// decrypt call on this context.
extern "C" int mfb_ssp_prover_polys_resident_async(mfb_ctx *ctx, mfb_ssp *h, const uint64_t *witness_limbs, size_t nlimbs,
                                                   uint64_t delta, void *stream, const uint32_t **wvh_dev) {
  int rc = ctx_enter(ctx);
  if (rc) return rc;
  if (!h || !witness_limbs || !wvh_dev) return ctx_bad_arg("mfb_ssp_prover_polys_resident_async: null pointer");
  PolyEngine &E = *poly_engine_of(ctx);
  const uint64_t L0 = E.launches;
  uint32_t *d_w = nullptr;
  rc = polys_resident_queue(ctx, h, witness_limbs, nlimbs, delta, (cudaStream_t)stream, &d_w);
  ctx_count_launches(ctx, E.launches - L0);
  *wvh_dev = d_w;
  return rc;
}

// ... on the context's own stream, which is idle on return
extern "C" int mfb_ssp_prover_polys_resident_dev(mfb_ctx *ctx, mfb_ssp *h, const uint64_t *witness_limbs, size_t nlimbs,
                                                 uint64_t delta, const uint32_t **wvh_dev) {
  int rc = mfb_ssp_prover_polys_resident_async(ctx, h, witness_limbs, nlimbs, delta, ctx ? ctx_stream_of(ctx) : nullptr, wvh_dev);
  if (rc) return rc;
  PTRY(cudaStreamSynchronize(ctx_stream_of(ctx)));
  return MFB_OK;
}

extern "C" size_t mfb_ssp_degree_bound(const mfb_ssp *h) { return h ? h->D : 0; }

extern "C" int mfb_ssp_eval(mfb_ctx *ctx, const uint64_t *polys, size_t D, size_t npoly, uint64_t x, uint64_t *values) {
  int rc = ctx_enter(ctx);
  if (rc) return rc;
  if (npoly == 0) return MFB_OK;
  if (!polys || !values) return ctx_bad_arg("mfb_ssp_eval: null pointer");
  if (D < 1 || D > (1u << 22)) return ctx_bad_arg("mfb_ssp_eval: need 1 <= D <= 2^22");
  PolyEngine &E = *poly_engine_of(ctx);
  cudaStream_t st = ctx_stream_of(ctx);
  const uint64_t L0 = E.launches;
  const size_t BATCH = (size_t)256 << 20 >> 3;  // coefficients staged per batch (256 MB)
  size_t per = BATCH / D ? BATCH / D : 1;
  if (per > npoly) per = npoly;
  void *d_pw, *d_p, *d_val;
  if ((rc = ctx_scratch(ctx, 0, D * 4, &d_pw))) return rc;
  if ((rc = ctx_scratch(ctx, 3, per * D * 8, &d_p))) return rc;
  if ((rc = ctx_scratch(ctx, 1, npoly * 8, &d_val))) return rc;
  k_powers<<<gridfor((uint32_t)D), 256, 0, st>>>(x, (uint32_t)D, (uint32_t *)d_pw);
  E.launches++;
  for (size_t q = 0; q < npoly; q += per) {
    const size_t cnt = npoly - q < per ? npoly - q : per;
    if ((rc = ctx_h2d(ctx, d_p, polys + q * D, cnt * D * 8, st)) != MFB_OK) return rc;
    k_eval<<<(unsigned)cnt, 256, 0, st>>>((const uint64_t *)d_p, (uint32_t)D, (const uint32_t *)d_pw, (uint64_t *)d_val + q);
    E.launches++;
    PTRY(cudaGetLastError());
    // (the staging buffer is reused in stream order: no synchronisation between batches)
  }
  PTRY(cudaMemcpyAsync(values, d_val, npoly * 8, cudaMemcpyDeviceToHost, st));
  PTRY(cudaStreamSynchronize(st));
  ctx_count_launches(ctx, E.launches - L0);
  return MFB_OK;
}

// values[q] = poly_{first+q}(x) for the polynomials first .. first+npoly of a RESIDENT blob ([t, v_0, ..., v_{M-1}]):
// the verifier's t(s), v_0(s) (snark.c:197-201, 214-215) without shipping 16 D bytes per proof check.
extern "C" int mfb_ssp_eval_resident(mfb_ctx *ctx, const mfb_ssp *h, size_t first, size_t npoly, uint64_t x, uint64_t *values) {
  int rc = ctx_enter(ctx);
  if (rc) return rc;
  if (npoly == 0) return MFB_OK;
  if (!h || !values) return ctx_bad_arg("mfb_ssp_eval_resident: null pointer");
  if (first > h->M + 1 || npoly > h->M + 1 - first) return ctx_bad_arg("mfb_ssp_eval_resident: polynomial range exceeds the blob");
  PolyEngine &E = *poly_engine_of(ctx);
  cudaStream_t st = ctx_stream_of(ctx);
  const uint64_t L0 = E.launches;
  const size_t D = h->D;
  void *d_pw, *d_val;
  if ((rc = ctx_scratch(ctx, 0, D * 4, &d_pw))) return rc;
  if ((rc = ctx_scratch(ctx, 1, npoly * 8, &d_val))) return rc;
  k_powers<<<gridfor((uint32_t)D), 256, 0, st>>>(x, (uint32_t)D, (uint32_t *)d_pw);
  k_eval32<<<(unsigned)npoly, 256, 0, st>>>(h->blob + first * D, (uint32_t)D, (const uint32_t *)d_pw, (uint64_t *)d_val);
  E.launches += 2;
  PTRY(cudaGetLastError());
  PTRY(cudaMemcpyAsync(values, d_val, npoly * 8, cudaMemcpyDeviceToHost, st));
  PTRY(cudaStreamSynchronize(st));
  ctx_count_launches(ctx, E.launches - L0);
  return MFB_OK;
}
