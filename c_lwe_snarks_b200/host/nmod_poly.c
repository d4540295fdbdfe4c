/* Dense polynomials over Z/nZ for the SSP SNARK host code (n = p = 2^32 - 5 in practice).
 *
 * Implements the declarations of include/compat/flint/nmod_poly.h.  FLINT is an external,
 * un-vendored dependency of the reference (configure.ac:22-23) and is absent from this
 * image; the call sites this serves are snark.c:93-110,122-181,197-215, ssp.c:18-77 and
 * lwe.c:183.  Multiplication uses three 31-bit NTT primes + Garner CRT (p - 1 = 2*5*429496729
 * has 2-adicity 1, so F_p itself has no useful NTT); division uses Newton inversion of the
 * reversed divisor.  All results are canonical residues in [0, n).
 */
#include <flint/nmod_poly.h>

#include <assert.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ---------------------------------------------------------------- scalar helpers */

static inline uint64_t n_addmod(uint64_t a, uint64_t b, uint64_t n) {
  uint64_t s = a + b;
  return (s < a || s >= n) ? s - n : s;
}
static inline uint64_t n_submod(uint64_t a, uint64_t b, uint64_t n) {
  return a >= b ? a - b : a + (n - b);
}
static inline uint64_t n_mulmod(uint64_t a, uint64_t b, uint64_t n) {
  return (uint64_t)((u128)a * b % n);
}
static uint64_t n_invmod(uint64_t a, uint64_t n) {
  /* extended Euclid on signed 128-bit cofactors */
  __int128 t = 0, newt = 1;
  uint64_t r = n, newr = a % n;
  while (newr) {
    uint64_t q = r / newr;
    __int128 tt = t - (__int128)q * newt;
    t = newt;
    newt = tt;
    uint64_t rr = r - q * newr;
    r = newr;
    newr = rr;
  }
  assert(r == 1 && "leading coefficient not invertible");
  if (t < 0) t += n;
  return (uint64_t)t;
}

/* ---------------------------------------------------------------- memory management */

void nmod_poly_init(nmod_poly_t poly, mp_limb_t n) {
  poly->coeffs = NULL;
  poly->alloc = 0;
  poly->length = 0;
  poly->mod.n = n;
  poly->mod.ninv = 0; /* FLINT's precomputed inverse; unused here */
  poly->mod.norm = n ? (flint_bitcnt_t)__builtin_clzl(n) : 0;
}

void nmod_poly_clear(nmod_poly_t poly) {
  free(poly->coeffs);
  poly->coeffs = NULL;
  poly->alloc = poly->length = 0;
}

void nmod_poly_fit_length(nmod_poly_t poly, slong alloc) {
  if (alloc <= poly->alloc) return;
  if (alloc < 2 * poly->alloc) alloc = 2 * poly->alloc;
  poly->coeffs = (mp_ptr)realloc(poly->coeffs, (size_t)alloc * sizeof(mp_limb_t));
  assert(poly->coeffs);
  poly->alloc = alloc;
}

static void poly_normalise(nmod_poly_t poly) {
  while (poly->length > 0 && poly->coeffs[poly->length - 1] == 0) poly->length--;
}

void nmod_poly_zero(nmod_poly_t poly) { poly->length = 0; }

void nmod_poly_set(nmod_poly_t a, const nmod_poly_t b) {
  if (a == b) return;
  nmod_poly_fit_length(a, b->length);
  if (b->length) memcpy(a->coeffs, b->coeffs, (size_t)b->length * sizeof(mp_limb_t));
  a->length = b->length;
}

void nmod_poly_set_coeff_ui(nmod_poly_t poly, slong j, ulong c) {
  if (c >= poly->mod.n) c %= poly->mod.n;
  nmod_poly_fit_length(poly, j + 1);
  if (j + 1 < poly->length) {
    poly->coeffs[j] = c;
  } else if (j + 1 == poly->length) {
    poly->coeffs[j] = c;
    poly_normalise(poly);
  } else {
    if (c == 0) return;
    for (slong i = poly->length; i < j; i++) poly->coeffs[i] = 0;
    poly->coeffs[j] = c;
    poly->length = j + 1;
  }
}

int nmod_poly_equal(const nmod_poly_t a, const nmod_poly_t b) {
  if (a->length != b->length) return 0;
  return a->length == 0 ||
         memcmp(a->coeffs, b->coeffs, (size_t)a->length * sizeof(mp_limb_t)) == 0;
}

/* ---------------------------------------------------------------- linear operations */

void nmod_poly_add(nmod_poly_t res, const nmod_poly_t a, const nmod_poly_t b) {
  const uint64_t n = res->mod.n;
  slong la = a->length, lb = b->length, lmax = la > lb ? la : lb, lmin = la < lb ? la : lb;
  nmod_poly_fit_length(res, lmax);
  for (slong i = 0; i < lmin; i++) res->coeffs[i] = n_addmod(a->coeffs[i], b->coeffs[i], n);
  if (la > lb && res != a) memcpy(res->coeffs + lmin, a->coeffs + lmin, (size_t)(la - lmin) * 8);
  if (lb > la && res != b) memcpy(res->coeffs + lmin, b->coeffs + lmin, (size_t)(lb - lmin) * 8);
  res->length = lmax;
  poly_normalise(res);
}

void nmod_poly_sub(nmod_poly_t res, const nmod_poly_t a, const nmod_poly_t b) {
  const uint64_t n = res->mod.n;
  slong la = a->length, lb = b->length, lmax = la > lb ? la : lb, lmin = la < lb ? la : lb;
  nmod_poly_fit_length(res, lmax);
  for (slong i = 0; i < lmin; i++) res->coeffs[i] = n_submod(a->coeffs[i], b->coeffs[i], n);
  if (la > lb && res != a) memcpy(res->coeffs + lmin, a->coeffs + lmin, (size_t)(la - lmin) * 8);
  for (slong i = lmin; i < lb; i++) res->coeffs[i] = n_submod(0, b->coeffs[i], n);
  res->length = lmax;
  poly_normalise(res);
}

void nmod_poly_scalar_mul_nmod(nmod_poly_t res, const nmod_poly_t a, mp_limb_t c) {
  const uint64_t n = res->mod.n;
  if (c >= n) c %= n;
  nmod_poly_fit_length(res, a->length);
  for (slong i = 0; i < a->length; i++) res->coeffs[i] = n_mulmod(a->coeffs[i], c, n);
  res->length = a->length;
  poly_normalise(res);
}

mp_limb_t nmod_poly_evaluate_nmod(const nmod_poly_t poly, mp_limb_t c) {
  const uint64_t n = poly->mod.n;
  uint64_t acc = 0;
  if (c >= n) c %= n;
  for (slong i = poly->length - 1; i >= 0; i--)
    acc = n_addmod(n_mulmod(acc, c, n), poly->coeffs[i], n);
  return acc;
}

/* ---------------------------------------------------------------- multiplication */

#define NTT_PRIMES 3
static const uint32_t ntt_p[NTT_PRIMES] = {998244353u, 469762049u, 167772161u};
static const uint32_t ntt_g[NTT_PRIMES] = {3u, 3u, 3u};

static inline uint32_t mul32(uint32_t a, uint32_t b, uint32_t m) {
  return (uint32_t)((uint64_t)a * b % m);
}
static uint32_t pow32(uint32_t a, uint64_t e, uint32_t m) {
  uint32_t r = 1;
  while (e) {
    if (e & 1) r = mul32(r, a, m);
    a = mul32(a, a, m);
    e >>= 1;
  }
  return r;
}

/* in-place iterative radix-2 transform of length len = 2^k over Z/m */
static void ntt32(uint32_t *a, size_t len, int inverse, uint32_t m, uint32_t g) {
  for (size_t i = 1, j = 0; i < len; i++) {
    size_t bit = len >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      uint32_t t = a[i];
      a[i] = a[j];
      a[j] = t;
    }
  }
  uint32_t *tw = (uint32_t *)malloc((len / 2 + 1) * sizeof(uint32_t));
  assert(tw);
  for (size_t h = 1; h < len; h <<= 1) {
    uint32_t w = pow32(g, (m - 1) / (2 * h), m);
    if (inverse) w = pow32(w, m - 2, m);
    tw[0] = 1;
    for (size_t k = 1; k < h; k++) tw[k] = mul32(tw[k - 1], w, m);
    for (size_t i = 0; i < len; i += 2 * h) {
      for (size_t k = 0; k < h; k++) {
        uint32_t u = a[i + k];
        uint32_t v = mul32(a[i + k + h], tw[k], m);
        uint32_t s = u + v;
        a[i + k] = s >= m ? s - m : s;
        a[i + k + h] = u >= v ? u - v : u + m - v;
      }
    }
  }
  free(tw);
  if (inverse) {
    uint32_t ninv = pow32((uint32_t)(len % m), m - 2, m);
    for (size_t i = 0; i < len; i++) a[i] = mul32(a[i], ninv, m);
  }
}

static void mul_schoolbook(uint64_t *r, const uint64_t *a, slong la, const uint64_t *b, slong lb,
                           uint64_t n) {
  for (slong i = 0; i < la + lb - 1; i++) r[i] = 0;
  if (n <= 0xffffffffULL) {
    for (slong i = 0; i < la; i++) {
      uint64_t ai = a[i];
      if (!ai) continue;
      for (slong j = 0; j < lb; j++) r[i + j] = (uint64_t)(((u128)ai * b[j] + r[i + j]) % n);
    }
  } else {
    for (slong i = 0; i < la; i++)
      for (slong j = 0; j < lb; j++) r[i + j] = n_addmod(r[i + j], n_mulmod(a[i], b[j], n), n);
  }
}

/* r[0 .. la+lb-1) = a * b mod n; r must not alias a or b */
static void mul_raw(uint64_t *r, const uint64_t *a, slong la, const uint64_t *b, slong lb,
                    uint64_t n) {
  if (la == 0 || lb == 0) return;
  slong lmin = la < lb ? la : lb;
  /* exact integer product coefficients are < lmin * (n-1)^2; the CRT modulus is ~7.87e25 */
  const long double crt_cap = 7.8e25L;
  int ntt_ok = n <= 0xffffffffULL &&
               (long double)lmin * (long double)(n - 1) * (long double)(n - 1) < crt_cap &&
               (la + lb - 1) <= (1L << 23);
  if (lmin < 32 || !ntt_ok) {
    mul_schoolbook(r, a, la, b, lb, n);
    return;
  }
  size_t len = 1;
  while ((slong)len < la + lb - 1) len <<= 1;
  uint32_t *res[NTT_PRIMES];
  uint32_t *fa = (uint32_t *)malloc(len * sizeof(uint32_t));
  uint32_t *fb = (uint32_t *)malloc(len * sizeof(uint32_t));
  assert(fa && fb);
  for (int k = 0; k < NTT_PRIMES; k++) {
    const uint32_t m = ntt_p[k];
    for (size_t i = 0; i < len; i++) fa[i] = (slong)i < la ? (uint32_t)(a[i] % m) : 0;
    ntt32(fa, len, 0, m, ntt_g[k]);
    if (a == b && la == lb) {
      for (size_t i = 0; i < len; i++) fa[i] = mul32(fa[i], fa[i], m);
    } else {
      for (size_t i = 0; i < len; i++) fb[i] = (slong)i < lb ? (uint32_t)(b[i] % m) : 0;
      ntt32(fb, len, 0, m, ntt_g[k]);
      for (size_t i = 0; i < len; i++) fa[i] = mul32(fa[i], fb[i], m);
    }
    ntt32(fa, len, 1, m, ntt_g[k]);
    res[k] = (uint32_t *)malloc((size_t)(la + lb - 1) * sizeof(uint32_t));
    assert(res[k]);
    memcpy(res[k], fa, (size_t)(la + lb - 1) * sizeof(uint32_t));
  }
  free(fa);
  free(fb);
  /* Garner: x = x1 + x2*P1 + x3*P1*P2 */
  const uint64_t P1 = ntt_p[0], P2 = ntt_p[1], P3 = ntt_p[2];
  const uint32_t inv_p1_p2 = pow32((uint32_t)(P1 % P2), P2 - 2, (uint32_t)P2);
  const uint32_t inv_p1p2_p3 = pow32((uint32_t)((u128)P1 * P2 % P3), P3 - 2, (uint32_t)P3);
  const uint64_t p1_mod_n = P1 % n;
  const uint64_t p1p2_mod_n = (uint64_t)((u128)P1 * P2 % n);
  for (slong i = 0; i < la + lb - 1; i++) {
    uint64_t x1 = res[0][i];
    uint64_t x2 = (uint64_t)mul32((uint32_t)((res[1][i] + P2 - x1 % P2) % P2), inv_p1_p2, (uint32_t)P2);
    uint64_t partial = (x1 + x2 * P1) % P3; /* x1 + x2*P1 < 2^60 */
    uint64_t x3 = (uint64_t)mul32((uint32_t)((res[2][i] + P3 - partial) % P3), inv_p1p2_p3, (uint32_t)P3);
    u128 v = (u128)x1 % n + (u128)(x2 % n) * p1_mod_n + (u128)(x3 % n) * p1p2_mod_n;
    r[i] = (uint64_t)(v % n);
  }
  for (int k = 0; k < NTT_PRIMES; k++) free(res[k]);
}

void nmod_poly_mul(nmod_poly_t res, const nmod_poly_t a, const nmod_poly_t b) {
  if (a->length == 0 || b->length == 0) {
    res->length = 0;
    return;
  }
  slong lr = a->length + b->length - 1;
  uint64_t *tmp = (uint64_t *)malloc((size_t)lr * sizeof(uint64_t));
  assert(tmp);
  mul_raw(tmp, a->coeffs, a->length, b->coeffs, b->length, res->mod.n);
  nmod_poly_fit_length(res, lr);
  memcpy(res->coeffs, tmp, (size_t)lr * sizeof(uint64_t));
  free(tmp);
  res->length = lr;
  poly_normalise(res);
}

void nmod_poly_pow(nmod_poly_t res, const nmod_poly_t a, ulong e) {
  nmod_poly_t base, acc;
  nmod_poly_init(base, a->mod.n);
  nmod_poly_init(acc, a->mod.n);
  nmod_poly_set(base, a);
  nmod_poly_set_coeff_ui(acc, 0, 1);
  while (e) {
    if (e & 1) nmod_poly_mul(acc, acc, base);
    e >>= 1;
    if (e) nmod_poly_mul(base, base, base);
  }
  nmod_poly_set(res, acc);
  nmod_poly_clear(base);
  nmod_poly_clear(acc);
}

/* ---------------------------------------------------------------- division */

/* g[0..m) = f^{-1} mod x^m, f[0] invertible; f has lf coefficients (read as zero beyond) */
static void inv_series(uint64_t *g, const uint64_t *f, slong lf, slong m, uint64_t n) {
  uint64_t *t1 = (uint64_t *)malloc((size_t)(4 * m + 8) * sizeof(uint64_t));
  uint64_t *t2 = (uint64_t *)malloc((size_t)(4 * m + 8) * sizeof(uint64_t));
  assert(t1 && t2);
  g[0] = n_invmod(f[0], n);
  slong k = 1;
  while (k < m) {
    slong k2 = 2 * k < m ? 2 * k : m;
    slong lfk = lf < k2 ? lf : k2;
    /* t1 = f * g mod x^k2 */
    mul_raw(t1, f, lfk, g, k, n);
    slong l1 = lfk + k - 1;
    if (l1 > k2) l1 = k2;
    /* t1 = 2 - t1 */
    for (slong i = 0; i < l1; i++) t1[i] = n_submod(0, t1[i], n);
    t1[0] = n_addmod(t1[0], 2 % n, n);
    /* g = g * t1 mod x^k2 */
    mul_raw(t2, g, k, t1, l1, n);
    slong l2 = k + l1 - 1;
    for (slong i = k; i < k2; i++) g[i] = i < l2 ? t2[i] : 0;
    k = k2;
  }
  free(t1);
  free(t2);
}

void nmod_poly_divrem(nmod_poly_t q, nmod_poly_t r, const nmod_poly_t a, const nmod_poly_t b) {
  const uint64_t n = a->mod.n;
  const slong la = a->length, lb = b->length;
  assert(lb > 0 && "division by zero polynomial");
  if (la < lb) {
    if (r) nmod_poly_set(r, a);
    if (q) q->length = 0;
    return;
  }
  const slong lq = la - lb + 1;
  uint64_t *qq = (uint64_t *)malloc((size_t)lq * sizeof(uint64_t));
  assert(qq);
  if (lq < 32 || lb < 32) {
    /* schoolbook long division on a scratch copy of a */
    uint64_t *rem = (uint64_t *)malloc((size_t)la * sizeof(uint64_t));
    assert(rem);
    memcpy(rem, a->coeffs, (size_t)la * sizeof(uint64_t));
    uint64_t linv = n_invmod(b->coeffs[lb - 1], n);
    for (slong i = lq - 1; i >= 0; i--) {
      uint64_t c = n_mulmod(rem[i + lb - 1], linv, n);
      qq[i] = c;
      if (c)
        for (slong j = 0; j < lb; j++)
          rem[i + j] = n_submod(rem[i + j], n_mulmod(c, b->coeffs[j], n), n);
    }
    if (r) {
      nmod_poly_fit_length(r, lb);
      memcpy(r->coeffs, rem, (size_t)(lb - 1) * sizeof(uint64_t));
      r->length = lb - 1;
      poly_normalise(r);
    }
    free(rem);
  } else {
    /* q = rev( rev(a) * rev(b)^{-1} mod x^lq ) */
    slong lbr = lb < lq ? lb : lq;
    uint64_t *brev = (uint64_t *)malloc((size_t)lbr * sizeof(uint64_t));
    uint64_t *arev = (uint64_t *)malloc((size_t)lq * sizeof(uint64_t));
    uint64_t *binv = (uint64_t *)malloc((size_t)lq * sizeof(uint64_t));
    uint64_t *prod = (uint64_t *)malloc((size_t)(2 * lq) * sizeof(uint64_t));
    assert(brev && arev && binv && prod);
    for (slong i = 0; i < lbr; i++) brev[i] = b->coeffs[lb - 1 - i];
    for (slong i = 0; i < lq; i++) arev[i] = a->coeffs[la - 1 - i];
    inv_series(binv, brev, lbr, lq, n);
    mul_raw(prod, arev, lq, binv, lq, n);
    for (slong i = 0; i < lq; i++) qq[lq - 1 - i] = prod[i];
    free(brev);
    free(arev);
    free(binv);
    free(prod);
    if (r) {
      /* r = a - q*b, only the low lb-1 coefficients can be non-zero */
      uint64_t *qb = (uint64_t *)malloc((size_t)(lq + lb) * sizeof(uint64_t));
      assert(qb);
      mul_raw(qb, qq, lq, b->coeffs, lb, n);
      nmod_poly_fit_length(r, lb);
      for (slong i = 0; i < lb - 1; i++) {
        uint64_t ai = a->coeffs[i];
        r->coeffs[i] = n_submod(ai, qb[i], n);
      }
      r->length = lb - 1;
      poly_normalise(r);
      free(qb);
    }
  }
  if (q) {
    nmod_poly_fit_length(q, lq);
    memcpy(q->coeffs, qq, (size_t)lq * sizeof(uint64_t));
    q->length = lq;
    poly_normalise(q);
  }
  free(qq);
}

void nmod_poly_div(nmod_poly_t q, const nmod_poly_t a, const nmod_poly_t b) {
  nmod_poly_divrem(q, NULL, a, b);
}

void nmod_poly_rem(nmod_poly_t r, const nmod_poly_t a, const nmod_poly_t b) {
  nmod_poly_t rr;
  nmod_poly_init(rr, a->mod.n);
  nmod_poly_divrem(NULL, rr, a, b);
  nmod_poly_set(r, rr);
  nmod_poly_clear(rr);
}
