/* lwe.h of the reference: Regev scheme over q_eff = 2^704, p = 2^32 - 5, n = 1470 — host adapters.
 *
 * Every function keeps the reference's signature and effect on its mpz_t arguments; the arithmetic runs in
 * the kernels behind include/mfb200.h.  What stays on the host is conversion between mpz_t and flat limb
 * arrays, entropy draws, and ct_smudge (one 704-bit addition; SURVEY.md §8 a14 keeps it in C).
 */
#include "mf_internal.h"

#define FLAT_CT (MFB_FLAT_CT_U64)

void key_gen(sk_t sk) { /* lwe.c:30-34: 1470 x 92 OS-entropy bytes, little-endian */
  mf_gpu_prefetch();
  mpz_initv(sk, GAMMA_N);
  mpz2_urandombv2(sk, GAMMA_LOGQ, GAMMA_N);
}

void key_clear(sk_t sk) { mpz_clearv(sk, GAMMA_N); }

void ct_init(ct_t ct) { mpz_initv(ct, GAMMA_N + 1); }

void ct_clear(ct_t ct) { mpz_clearv(ct, GAMMA_N + 1); }

void ct_zero(ct_t rop) { /* lwe.c:160-165 */
  for (size_t i = 0; i <= GAMMA_N; i++) mpz_set_ui(rop[i], 0);
}

/* lwe.c:51-58: one entropy byte; bit 0 set -> negate */
static void randomsgn(mpz_t dst, const mpz_t src) {
  uint8_t sign;
  mf_entropy(&sign, 1);
  if (sign & 0x01) mpz_neg(dst, src);
}

void errdist_uniform(mpz_t e) { mpz2_urandomb2(e, GAMMA_LOG_SIGMA + 3); } /* lwe.c:60-63 */

void ct_smudge(ct_t ct) { /* lwe.c:65-76 */
  mpz_t smudging;
  mpz_init(smudging);
  mpz2_urandomb2(smudging, GAMMA_LOG_SMUDGING);
  randomsgn(smudging, smudging);
  mpz_mul_ui(smudging, smudging, GAMMA_P);
  mpz_add(ct[GAMMA_N], ct[GAMMA_N], smudging);
  modq(ct[GAMMA_N]);
  mpz_clear(smudging);
}

void ct_export(uint8_t *buf, ct_t ct) { /* lwe.c:115-119 */
  memset(buf, 0, CT_BYTES);
  if (mpz_sizeinbase(ct[GAMMA_N], 2) > 8 * CT_BYTES) {
    fprintf(stderr, "mangiafuoco_b200: ct_export: b does not fit %lu bytes\n", CT_BYTES);
    abort();
  }
  mpz_export(buf, NULL, -1, sizeof(uint8_t), -1, 0, ct[GAMMA_N]);
}

/* a_j <- 1470 consecutive 736-bit stream draws (mpz2_urandommv entropy.h:62-66): one keystream read */
static void draw_a(ct_t ct, rng_t rng) {
  uint8_t *ks = malloc(CTR_CT);
  if (!ks) mf_die("malloc");
  rng_gen(rng, ks, CTR_CT);
  for (size_t j = 0; j < GAMMA_N; j++) mf_bytes_to_mpz(ct[j], ks + j * CT_BYTES, CT_BYTES);
  free(ks);
}

void ct_import(ct_t ct, rng_t rng, uint8_t *buf) { /* lwe.c:122-126 */
  draw_a(ct, rng);
  mf_bytes_to_mpz(ct[GAMMA_N], buf, LOGQ_BYTES);
}

void decompress_encryption(ct_t c, rng_t rng, mpz_t b) { /* lwe.c:99-103 */
  draw_a(c, rng);
  mpz_set(c[GAMMA_N], b);
}

static uint64_t *sk_flat_new(sk_t sk) {
  uint64_t *f = malloc(MFB_FLAT_SK_U64 * 8);
  if (!f) mf_die("malloc");
  for (size_t i = 0; i < GAMMA_N; i++)
    if (mf_to_flat(f + i * MF_LIMBS, sk[i])) {
      fprintf(stderr, "mangiafuoco_b200: negative secret-key coordinate %zu\n", i);
      abort();
    }
  return f;
}

void regev_encrypt2(ct_t c, rng_t rs, sk_t sk, mpz_t m, void (*chi)(mpz_t)) { /* lwe.c:78-97 */
  mpz_t e;
  mpz_init(e);
  (*chi)(e);
  /* c[n] = e*p happens in the kernel; the sign draw comes after it and never reaches c (lwe.c:86-87) */
  uint64_t e_flat[MF_LIMBS];
  {
    mpz_t t;
    mpz_init(t);
    mpz_fdiv_r_2exp(t, e, 64 * MF_LIMBS); /* e mod 2^704, also right for a signed chi */
    mf_to_flat(e_flat, t);
    mpz_clear(t);
  }
  randomsgn(e, e);

  const uint64_t pos = mf_rng_pos(rs);
  draw_a(c, rs); /* the caller's ciphertext holds the full 736-bit a_j, as in the reference */

  mpz_t mm;
  mpz_init(mm);
  mpz_fdiv_r_2exp(mm, m, 64);
  if (mpz_cmp(mm, m)) {
    fprintf(stderr, "mangiafuoco_b200: regev_encrypt2: message does not fit 64 bits (the reference asserts m < p)\n");
    abort();
  }
  uint64_t msg = mpz_get_ui(mm);
  mpz_clear(mm);

  uint64_t *skf = sk_flat_new(sk);
  uint8_t rec[CT_BYTES];
  MF_GPU(mfb_encrypt(mf_gpu(), mf_rng_seed(rs), pos, skf, &msg, (const uint8_t *)e_flat, 8 * MF_LIMBS, 8 * MF_LIMBS, 1, rec));
  explicit_bzero(skf, MFB_FLAT_SK_U64 * 8); /* the flat copy of the secret key does not outlive the call */
  free(skf);
  mf_bytes_to_mpz(c[GAMMA_N], rec, CT_BYTES);
  mpz_clear(e);
}

/* lwe.c:20-28: rop += sum a_i * b_i, then modq.  For rop >= 0 that is (rop + <a, b>) mod 2^704. */
void mpz_add_dotp(mpz_t rop, mpz_t a[], mpz_t b[], size_t len) {
  if (len > GAMMA_N) {
    fprintf(stderr, "mangiafuoco_b200: mpz_add_dotp: len %zu > %d\n", len, GAMMA_N);
    abort();
  }
  uint64_t *ct = calloc(FLAT_CT, 8), *sk = calloc(MFB_FLAT_SK_U64, 8);
  if (!ct || !sk) mf_die("malloc");
  for (size_t i = 0; i < len; i++) {
    /* (-x)(-y) = xy; a single negative factor flips the product: fold signs into the a side mod 2^704 */
    int na = mf_to_flat(ct + i * MF_LIMBS, a[i]), nb = mf_to_flat(sk + i * MF_LIMBS, b[i]);
    if (na ^ nb) { /* two's complement of the a operand */
      uint64_t *x = ct + i * MF_LIMBS, carry = 1;
      for (int l = 0; l < MF_LIMBS; l++) {
        x[l] = ~x[l] + carry;
        carry = carry && x[l] == 0;
      }
    }
  }
  uint64_t m_unused, dot[MF_LIMBS];
  MF_GPU(mfb_decrypt(mf_gpu(), sk, ct, NULL, 1, &m_unused, dot));
  mpz_t d;
  mpz_init(d);
  mf_from_flat(d, dot);
  mpz_add(rop, rop, d);
  if (SIZ(rop) >= 0) mpz_fdiv_r_2exp(rop, rop, 64 * MF_LIMBS); /* == modq for a non-negative value */
  mpz_clear(d);
  free(ct);
  free(sk);
}

void regev_decrypt(mpz_t m, sk_t sk, ct_t ct) { /* lwe.c:105-111 */
  uint64_t *cf = malloc(FLAT_CT * 8);
  if (!cf) mf_die("malloc");
  for (size_t i = 0; i < GAMMA_N; i++)
    if (mf_to_flat(cf + i * MF_LIMBS, ct[i])) {
      fprintf(stderr, "mangiafuoco_b200: regev_decrypt: negative a coordinate %zu\n", i);
      abort();
    }
  uint8_t neg = (uint8_t)mf_to_flat(cf + (size_t)GAMMA_N * MF_LIMBS, ct[GAMMA_N]);
  if (mpz_sizeinbase(ct[GAMMA_N], 2) > 64 * MF_LIMBS) {
    fprintf(stderr, "mangiafuoco_b200: regev_decrypt: |b| >= 2^704 is outside the supported range\n");
    abort();
  }
  uint64_t *skf = sk_flat_new(sk);
  uint64_t res;
  MF_GPU(mfb_decrypt(mf_gpu(), skf, cf, &neg, 1, &res, NULL));
  mpz_set_ui(m, res);
  explicit_bzero(skf, MFB_FLAT_SK_U64 * 8); /* the flat copy of the secret key does not outlive the call */
  free(skf);
  free(cf);
}

/* rop <- (init + sum_k coeff_k * ct_k) mod 2^704 over host ciphertexts, through the resident-lincomb kernel */
static void host_lincomb(ct_t rop, int use_rop, ct_t *cts, const uint32_t *coeffs, size_t d, const char *who) {
  uint64_t *flat = malloc(d * FLAT_CT * 8), *acc = calloc(FLAT_CT, 8);
  if (!flat || !acc) mf_die("malloc");
  for (size_t k = 0; k < d; k++) mf_ct_to_flat(flat + k * FLAT_CT, cts[k], who);
  if (use_rop) mf_ct_to_flat(acc, rop, who);
  MF_GPU(mfb_lincomb(mf_gpu(), flat, coeffs, d, acc));
  mf_ct_from_flat(rop, acc);
  free(flat);
  free(acc);
}

static uint32_t scalar32(uint64_t b, const char *who) {
  if (b >> 32) {
    fprintf(stderr, "mangiafuoco_b200: %s: scalar %lu >= 2^32 (the reference asserts b < p)\n", who, (unsigned long)b);
    abort();
  }
  return (uint32_t)b;
}

void ct_mul_ui(ct_t rop, ct_t a, uint64_t b) { /* lwe.c:131-139 */
  uint32_t co = scalar32(b, "ct_mul_ui");
  ct_t *src = (ct_t *)a;
  host_lincomb(rop, 0, src, &co, 1, "ct_mul_ui");
}

void ct_addmul_ui(ct_t rop, ct_t a, uint64_t b) { /* lwe.c:141-149 */
  uint32_t co = scalar32(b, "ct_addmul_ui");
  ct_t *src = (ct_t *)a;
  host_lincomb(rop, 1, src, &co, 1, "ct_addmul_ui");
}

void ct_add(ct_t rop, ct_t a, ct_t b) { /* lwe.c:151-157 */
  const uint32_t one = 1;
  uint64_t *flat = malloc(2 * FLAT_CT * 8), *acc = calloc(FLAT_CT, 8);
  if (!flat || !acc) mf_die("malloc");
  mf_ct_to_flat(flat, a, "ct_add");
  mf_ct_to_flat(flat + FLAT_CT, b, "ct_add");
  const uint32_t co[2] = {one, one};
  MF_GPU(mfb_lincomb(mf_gpu(), flat, co, 2, acc));
  mf_ct_from_flat(rop, acc);
  free(flat);
  free(acc);
}

void eval_poly(ct_t rop, rng_t rng, uint8_t (*c8)[CT_BYTES], nmod_poly_t p, size_t d) { /* lwe.c:176-186 */
  uint64_t *acc = malloc(FLAT_CT * 8), *co = malloc((d ? d : 1) * 8);
  if (!acc || !co) mf_die("malloc");
  mf_ct_to_flat(acc, rop, "eval_poly"); /* accumulates INTO rop, as the reference does */
  for (size_t i = 0; i < d; i++) co[i] = nmod_poly_get_coeff_ui(p, (slong)i);
  MF_GPU(mfb_eval_poly(mf_gpu(), mf_rng_seed(rng), mf_rng_pos(rng), (const uint8_t *)c8, co, NULL, d, acc));
  mf_rng_advance(rng, (uint64_t)d * CTR_CT); /* ct_import would have consumed 1470*92 bytes per ciphertext */
  mf_ct_from_flat(rop, acc);
  free(acc);
  free(co);
}
