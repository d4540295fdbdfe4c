/* TEST INFRASTRUCTURE — flat-buffer shim around the UNMODIFIED reference sources.
 *
 * Compiled together with /root/reference/src/{aes,entropy,lwe,ssp,snark}.c (where they lie)
 * into oracle/_ref/libmfref_*.so by oracle/Makefile.  It exists so that tests/ and
 * bench.py's cpu_baseline / `--impl reference` leg can drive the real reference from
 * Python (ctypes) with plain byte/limb arrays instead of mpz_t objects, and with a
 * deterministic entropy source.  Nothing in the product path may load this library.
 *
 * Determinism shims:
 *   - the reference sources are compiled with -Dgetrandom=ref_getrandom, so every
 *     getrandom(2) call site (lwe.h:101, lwe.c:54, entropy.c:36, snark.c:40, ssp.c:56,62)
 *     reads from the byte stream installed with ref_set_entropy();
 *   - GMP's allocator is replaced by a zero-filling one so that the 7 bits of the noise
 *     sample that entropy.c:34-40 leaves uninitialised are 0.
 *
 * Limb format used by every function here: little-endian uint64 limbs, REF_LIMBS (=12)
 * per coordinate, i.e. the reference's 736-bit width; a ciphertext is (GAMMA_N+1) coords.
 * `siz` outputs carry GMP's signed limb count so tests can see signs / normalisation.
 */
#include "config.h"

#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>

#include "lwe.h"
#include "snark.h"
#include "ssp.h"

#define REF_LIMBS 12
#define REF_NC (GAMMA_N + 1)

void ct_addmul_ui(ct_t rop, ct_t a, uint64_t b); /* lwe.c:141, not in lwe.h */
void ct_zero(ct_t rop);                          /* lwe.c:160, not in lwe.h */

/* ------------------------------------------------------------------ entropy interposer */

static const uint8_t *g_ent = NULL;
static size_t g_ent_len = 0, g_ent_pos = 0;
static uint64_t g_ent_calls = 0;

#ifdef MFB200_DROPIN
/* The same shim over the PRODUCT's drop-in library (libmangiafuoco_b200.so): entropy goes through the
 * library's own hook instead of a compile-time getrandom interposer, the instance size is a run-time value. */
static void dropin_entropy(void *buf, size_t len, void *arg);
void ref_set_instance(size_t D, size_t M) { mf_set_instance(D, M); }
uint64_t ref_gpu_launches(void) { return mf_gpu_launches(); }
#endif

void ref_set_entropy(const uint8_t *buf, size_t len) {
  g_ent = buf;
  g_ent_len = len;
  g_ent_pos = 0;
  g_ent_calls = 0;
#ifdef MFB200_DROPIN
  mf_set_entropy_source(buf ? dropin_entropy : NULL, NULL);
#endif
}
size_t ref_entropy_consumed(void) { return g_ent_pos; }
uint64_t ref_entropy_calls(void) { return g_ent_calls; }

ssize_t ref_getrandom(void *buf, size_t len, unsigned int flags) {
  g_ent_calls++;
  if (!g_ent) return syscall(SYS_getrandom, buf, len, flags);
  if (g_ent_pos + len > g_ent_len) {
    fprintf(stderr, "ref_shim: entropy stream exhausted (%zu + %zu > %zu)\n", g_ent_pos, len,
            g_ent_len);
    abort();
  }
  memcpy(buf, g_ent + g_ent_pos, len);
  g_ent_pos += len;
  return (ssize_t)len;
}

#ifdef MFB200_DROPIN
static void dropin_entropy(void *buf, size_t len, void *arg) {
  (void)arg;
  ref_getrandom(buf, len, 0);
}
#endif

/* ------------------------------------------------------------------ zeroing allocator */

static void *z_alloc(size_t n) {
  void *p = calloc(1, n);
  if (!p) abort();
  return p;
}
static void *z_realloc(void *p, size_t old, size_t n) {
  void *q = realloc(p, n);
  if (!q) abort();
  if (n > old) memset((uint8_t *)q + old, 0, n - old);
  return q;
}
static void z_free(void *p, size_t n) {
  (void)n;
  free(p);
}
__attribute__((constructor)) static void ref_shim_init(void) {
#ifndef MFB200_DROPIN
  mp_set_memory_functions(z_alloc, z_realloc, z_free);
#else
  (void)z_alloc; (void)z_realloc; (void)z_free;
#endif
}

/* ------------------------------------------------------------------ parameters */

uint64_t ref_param(const char *name) {
  if (!strcmp(name, "D")) return GAMMA_D;
  if (!strcmp(name, "M")) return GAMMA_M;
  if (!strcmp(name, "N")) return GAMMA_N;
  if (!strcmp(name, "LOGQ")) return GAMMA_LOGQ;
  if (!strcmp(name, "P")) return GAMMA_P;
  if (!strcmp(name, "CT_BYTES")) return CT_BYTES;
  if (!strcmp(name, "CTR_CT")) return CTR_CT;
  if (!strcmp(name, "CTR_S")) return CTR_S;
  if (!strcmp(name, "CTR_AS")) return CTR_AS;
  if (!strcmp(name, "CTR_BT")) return CTR_BT;
  if (!strcmp(name, "CTR_BV")) return CTR_BV;
  if (!strcmp(name, "SSP_SIZE")) return SSP_SIZE;
  if (!strcmp(name, "LOG_SMUDGING")) return GAMMA_LOG_SMUDGING;
  if (!strcmp(name, "LOG_SIGMA")) return GAMMA_LOG_SIGMA;
  return ~0ULL;
}

/* ------------------------------------------------------------------ mpz <-> flat limbs */

static void limbs_from_mpz(uint64_t *out, int32_t *siz, mpz_t z) {
  int n = abs(SIZ(z));
  memset(out, 0, REF_LIMBS * sizeof(uint64_t));
  if (n > REF_LIMBS) {
    fprintf(stderr, "ref_shim: mpz with %d limbs does not fit %d\n", n, REF_LIMBS);
    abort();
  }
  memcpy(out, PTR(z), (size_t)n * sizeof(uint64_t));
  if (siz) *siz = SIZ(z);
}
static void mpz_from_limbs(mpz_t z, const uint64_t *in, int nlimbs) {
  mpz_import(z, (size_t)nlimbs, -1, sizeof(uint64_t), 0, 0, in);
}
static void ct_to_flat(uint64_t *out, int32_t *siz, ct_t ct) {
  for (size_t i = 0; i < REF_NC; i++)
    limbs_from_mpz(out + i * REF_LIMBS, siz ? siz + i : NULL, ct[i]);
}
static void ct_from_flat(ct_t ct, const uint64_t *in) {
  for (size_t i = 0; i < REF_NC; i++) mpz_from_limbs(ct[i], in + i * REF_LIMBS, REF_LIMBS);
}
static void sk_from_flat(sk_t sk, const uint64_t *in) {
  mpz_initv(sk, GAMMA_N);
  for (size_t i = 0; i < GAMMA_N; i++) mpz_from_limbs(sk[i], in + i * REF_LIMBS, REF_LIMBS);
}

/* ------------------------------------------------------------------ L1/L2: AES-CTR stream */

/* aes.c:104-144 via entropy.c:46-61: `nbytes` of the stream of `seed` starting at byte `offset` */
void ref_stream(const uint8_t seed[40], uint64_t offset, uint8_t *out, size_t nbytes) {
  rng_t rng;
  rng_init(rng, (uint8_t *)seed);
  rng_seek(rng, offset);
  rng_gen(rng, out, nbytes);
  rng_clear(rng);
}

/* same bytes but read in `chunk`-byte calls (chunking independence, test_entropy.c:111-137) */
void ref_stream_chunked(const uint8_t seed[40], uint64_t offset, uint8_t *out, size_t nbytes,
                        size_t chunk) {
  rng_t rng;
  rng_init(rng, (uint8_t *)seed);
  rng_seek(rng, offset);
  for (size_t done = 0; done < nbytes; done += chunk)
    rng_gen(rng, out + done, nbytes - done < chunk ? nbytes - done : chunk);
  rng_clear(rng);
}

/* entropy.c:11-26: one PRG-backed draw of nbits at stream offset */
void ref_urandomb(const uint8_t seed[40], uint64_t offset, size_t nbits, uint64_t *out_limbs,
                  int32_t *siz) {
  rng_t rng;
  mpz_t a;
  mpz_init(a);
  rng_init(rng, (uint8_t *)seed);
  rng_seek(rng, offset);
  mpz2_urandomb(a, rng, nbits);
  limbs_from_mpz(out_limbs, siz, a);
  mpz_clear(a);
  rng_clear(rng);
}

/* lwe.h:108-118 on an arbitrary non-negative input of `nlimbs` limbs */
void ref_modq(const uint64_t *in, int nlimbs, uint64_t *out_limbs, int32_t *siz) {
  mpz_t a;
  mpz_init(a);
  mpz_from_limbs(a, in, nlimbs);
  modq(a);
  limbs_from_mpz(out_limbs, siz, a);
  mpz_clear(a);
}

/* ------------------------------------------------------------------ L3: ciphertext ops */

/* lwe.c:122-126 */
void ref_ct_import(const uint8_t seed[40], uint64_t offset, const uint8_t *b92, uint64_t *out,
                   int32_t *siz) {
  rng_t rng;
  ct_t ct;
  ct_init(ct);
  rng_init(rng, (uint8_t *)seed);
  rng_seek(rng, offset);
  ct_import(ct, rng, (uint8_t *)b92);
  ct_to_flat(out, siz, ct);
  ct_clear(ct);
  rng_clear(rng);
}

/* lwe.c:115-119 */
void ref_ct_export(const uint64_t *ct_flat, uint8_t *b92) {
  ct_t ct;
  ct_init(ct);
  ct_from_flat(ct, ct_flat);
  ct_export(b92, ct);
  ct_clear(ct);
}

/* lwe.c:176-186: rop (in/out, flat) += sum_i coeffs[i] * CT_i; stream positioned at `offset` */
void ref_eval_poly(const uint8_t seed[40], uint64_t offset, const uint8_t *c8, const uint64_t *coeffs,
                   size_t d, uint64_t *rop_flat, int32_t *siz) {
  rng_t rng;
  ct_t rop;
  nmod_poly_t p;
  ct_init(rop);
  ct_from_flat(rop, rop_flat);
  nmod_poly_init(p, GAMMA_P);
  for (size_t i = 0; i < d; i++) nmod_poly_set_coeff_ui(p, (slong)i, coeffs[i]);
  rng_init(rng, (uint8_t *)seed);
  rng_seek(rng, offset);
  eval_poly(rop, rng, (uint8_t(*)[CT_BYTES])c8, p, d);
  ct_to_flat(rop_flat, siz, rop);
  nmod_poly_clear(p);
  ct_clear(rop);
  rng_clear(rng);
}

/* Timed variant for the CPU baseline: returns seconds spent inside eval_poly only and, when
 * the pointers are non-NULL, the split the survey quotes (ct_import vs ct_addmul_ui,
 * lwe.c:182-183) measured by running the two halves of the loop body separately. */
double ref_time_eval_poly(const uint8_t seed[40], uint64_t offset, const uint8_t *c8,
                          const uint64_t *coeffs, size_t d, uint64_t *rop_flat,
                          double *t_import, double *t_addmul) {
  struct timespec t0, t1;
  rng_t rng;
  ct_t rop, ct;
  nmod_poly_t p;
  ct_init(rop);
  ct_init(ct);
  nmod_poly_init(p, GAMMA_P);
  for (size_t i = 0; i < d; i++) nmod_poly_set_coeff_ui(p, (slong)i, coeffs[i]);
  rng_init(rng, (uint8_t *)seed);
  rng_seek(rng, offset);
  clock_gettime(CLOCK_MONOTONIC, &t0);
  eval_poly(rop, rng, (uint8_t(*)[CT_BYTES])c8, p, d);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  double total = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  if (rop_flat) ct_to_flat(rop_flat, NULL, rop);
  if (t_import && t_addmul) {
    size_t dd = d < 64 ? d : 64;
    rng_seek(rng, offset);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (size_t i = 0; i < dd; i++) ct_import(ct, rng, (uint8_t *)c8 + i * CT_BYTES);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *t_import = ((double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec)) / (double)dd;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (size_t i = 0; i < dd; i++) ct_addmul_ui(rop, ct, nmod_poly_get_coeff_ui(p, (slong)i));
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *t_addmul = ((double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec)) / (double)dd;
  }
  nmod_poly_clear(p);
  ct_clear(rop);
  ct_clear(ct);
  rng_clear(rng);
  return total;
}

/* lwe.c:131-157 on flat ciphertexts */
void ref_ct_mul_ui(const uint64_t *a, uint64_t b, uint64_t *out, int32_t *siz) {
  ct_t x;
  ct_init(x);
  ct_from_flat(x, a);
  ct_mul_ui(x, x, b);
  ct_to_flat(out, siz, x);
  ct_clear(x);
}
void ref_ct_addmul_ui(uint64_t *rop_flat, const uint64_t *a, uint64_t b, int32_t *siz) {
  ct_t r, x;
  ct_init(r);
  ct_init(x);
  ct_from_flat(r, rop_flat);
  ct_from_flat(x, a);
  ct_addmul_ui(r, x, b);
  ct_to_flat(rop_flat, siz, r);
  ct_clear(r);
  ct_clear(x);
}
void ref_ct_add(const uint64_t *a, const uint64_t *b, uint64_t *out, int32_t *siz) {
  ct_t x, y;
  ct_init(x);
  ct_init(y);
  ct_from_flat(x, a);
  ct_from_flat(y, b);
  ct_add(x, x, y);
  ct_to_flat(out, siz, x);
  ct_clear(x);
  ct_clear(y);
}
/* lwe.c:65-76; consumes 80 + 1 entropy bytes */
void ref_ct_smudge(uint64_t *ct_flat, int32_t *siz) {
  ct_t x;
  ct_init(x);
  ct_from_flat(x, ct_flat);
  ct_smudge(x);
  ct_to_flat(ct_flat, siz, x);
  ct_clear(x);
}

/* lwe.c:30-34; consumes GAMMA_N * 92 entropy bytes */
void ref_key_gen(uint64_t *sk_flat) {
  sk_t sk;
  key_gen(sk);
  for (size_t i = 0; i < GAMMA_N; i++) limbs_from_mpz(sk_flat + i * REF_LIMBS, NULL, sk[i]);
  key_clear(sk);
}

/* lwe.c:78-97 + ct_export, `count` consecutive encryptions from stream `offset`;
 * consumes (69 + 1) entropy bytes per encryption.  out_ct_flat may be NULL. */
void ref_encrypt(const uint8_t seed[40], uint64_t offset, const uint64_t *sk_flat,
                 const uint64_t *m, size_t count, uint8_t *out_b92, uint64_t *out_ct_flat) {
  rng_t rng;
  sk_t sk;
  ct_t ct;
  mpz_t mm;
  mpz_init(mm);
  ct_init(ct);
  sk_from_flat(sk, sk_flat);
  rng_init(rng, (uint8_t *)seed);
  rng_seek(rng, offset);
  for (size_t i = 0; i < count; i++) {
    mpz_set_ui(mm, m[i]);
    regev_encrypt(ct, rng, sk, mm);
    ct_export(out_b92 + i * CT_BYTES, ct);
    if (out_ct_flat) ct_to_flat(out_ct_flat + i * REF_NC * REF_LIMBS, NULL, ct);
  }
  key_clear(sk);
  ct_clear(ct);
  mpz_clear(mm);
  rng_clear(rng);
}

/* lwe.c:105-111 */
uint64_t ref_decrypt(const uint64_t *sk_flat, const uint64_t *ct_flat) {
  sk_t sk;
  ct_t ct;
  mpz_t m;
  mpz_init(m);
  ct_init(ct);
  sk_from_flat(sk, sk_flat);
  ct_from_flat(ct, ct_flat);
  regev_decrypt(m, sk, ct);
  uint64_t r = mpz_get_ui(m);
  key_clear(sk);
  ct_clear(ct);
  mpz_clear(m);
  return r;
}

/* lwe.h:57-61 / lwe.c:20-28: rop = <a, b> mod 2^704 over `len` coordinates */
void ref_dotp(const uint64_t *a_flat, const uint64_t *b_flat, size_t len, uint64_t *out_limbs,
              int32_t *siz) {
  mpz_t *a = malloc(len * sizeof(mpz_t)), *b = malloc(len * sizeof(mpz_t));
  mpz_t r;
  mpz_init(r);
  for (size_t i = 0; i < len; i++) {
    mpz_init(a[i]);
    mpz_init(b[i]);
    mpz_from_limbs(a[i], a_flat + i * REF_LIMBS, REF_LIMBS);
    mpz_from_limbs(b[i], b_flat + i * REF_LIMBS, REF_LIMBS);
  }
  mpz_dotp(r, a, b, len);
  limbs_from_mpz(out_limbs, siz, r);
  for (size_t i = 0; i < len; i++) {
    mpz_clear(a[i]);
    mpz_clear(b[i]);
  }
  free(a);
  free(b);
  mpz_clear(r);
}

/* ------------------------------------------------------------------ L4/L5: full SNARK */

/* ssp.c:37-77; consumes M/8 + M * 8*D entropy bytes.  witness_bits: ceil(M/64) limbs out */
void ref_random_ssp(uint8_t *ssp, uint64_t *witness_limbs, size_t witness_nlimbs) {
  mpz_t w;
  mpz_init(w);
  random_ssp(w, ssp);
  memset(witness_limbs, 0, witness_nlimbs * sizeof(uint64_t));
  memcpy(witness_limbs, PTR(w), (size_t)abs(SIZ(w)) * sizeof(uint64_t));
  mpz_clear(w);
}

/* crs_init + setup (snark.c:35-48,57-115).  Entropy order: 40 (seed) then 8,8,8, then
 * N*92 (sk), then (69,1) per encryption.  CRS is returned in wire form. */
void ref_setup(const uint8_t *ssp, uint8_t seed_out[40], uint8_t *crs_s, uint8_t *crs_as,
               uint8_t *crs_v, uint8_t *crs_tb, uint64_t abs_out[3], uint64_t *sk_flat) {
  crs_t crs;
  vrs_t vrs;
  crs_init(crs);
  setup(crs, vrs, (ssp_t)ssp);
  memcpy(seed_out, crs->seed, 40);
  memcpy(crs_s, crs->s, CT_BYTES * GAMMA_D);
  memcpy(crs_as, crs->as, CT_BYTES * GAMMA_D);
  memcpy(crs_v, crs->v, CT_BYTES * (GAMMA_M - 1));
  memcpy(crs_tb, crs->t, CT_BYTES);
  abs_out[0] = vrs->alpha;
  abs_out[1] = vrs->beta;
  abs_out[2] = vrs->s;
  for (size_t i = 0; i < GAMMA_N; i++) limbs_from_mpz(sk_flat + i * REF_LIMBS, NULL, vrs->sk[i]);
  key_clear(vrs->sk);
  crs_clear(crs);
}

static void crs_from_wire(crs_t crs, const uint8_t seed[40], const uint8_t *crs_s,
                          const uint8_t *crs_as, const uint8_t *crs_v, const uint8_t *crs_tb) {
  memcpy(crs->seed, seed, 40);
  crs->s = (uint8_t(*)[CT_BYTES])crs_s;
  crs->as = (uint8_t(*)[CT_BYTES])crs_as;
  crs->v = (uint8_t(*)[CT_BYTES])crs_v;
  crs->t = (uint8_t *)crs_tb;
}

/* snark.c:117-190.  Entropy: 8 (delta) then 5 x (80, 1) smudging.  proof_flat: 5 flat
 * ciphertexts in struct order h, hat_h, hat_v, v_w, b_w (snark.h:14-20). */
void ref_prover(const uint8_t *ssp, const uint8_t seed[40], const uint8_t *crs_s,
                const uint8_t *crs_as, const uint8_t *crs_v, const uint8_t *crs_tb,
                const uint64_t *witness_limbs, size_t witness_nlimbs, uint64_t *proof_flat,
                int32_t *siz) {
  crs_t crs;
  proof_t pi;
  mpz_t w;
  mpz_init(w);
  mpz_from_limbs(w, witness_limbs, (int)witness_nlimbs);
  crs_from_wire(crs, seed, crs_s, crs_as, crs_v, crs_tb);
  proof_init(pi);
  prover(pi, crs, (ssp_t)ssp, w);
  const size_t stride = REF_NC * REF_LIMBS;
  ct_to_flat(proof_flat + 0 * stride, siz ? siz + 0 * REF_NC : NULL, pi->h);
  ct_to_flat(proof_flat + 1 * stride, siz ? siz + 1 * REF_NC : NULL, pi->hat_h);
  ct_to_flat(proof_flat + 2 * stride, siz ? siz + 2 * REF_NC : NULL, pi->hat_v);
  ct_to_flat(proof_flat + 3 * stride, siz ? siz + 3 * REF_NC : NULL, pi->v_w);
  ct_to_flat(proof_flat + 4 * stride, siz ? siz + 4 * REF_NC : NULL, pi->b_w);
  proof_clear(pi);
  mpz_clear(w);
}

#ifdef MFB200_DROPIN
/* product-only: the same prover call with the two big CRS regions made resident in HBM first */
void ref_prover_resident(const uint8_t *ssp, const uint8_t seed[40], const uint8_t *crs_s,
                         const uint8_t *crs_as, const uint8_t *crs_v, const uint8_t *crs_tb,
                         const uint64_t *witness_limbs, size_t witness_nlimbs, uint64_t *proof_flat,
                         int32_t *siz) {
  crs_t crs;
  proof_t pi;
  mpz_t w;
  mpz_init(w);
  mpz_from_limbs(w, witness_limbs, (int)witness_nlimbs);
  crs_from_wire(crs, seed, crs_s, crs_as, crs_v, crs_tb);
  mf_crs_make_resident(crs);
  mf_ssp_make_resident((ssp_t)ssp);
  proof_init(pi);
  prover(pi, crs, (ssp_t)ssp, w);
  mf_ssp_release((ssp_t)ssp);
  mf_crs_release(crs);
  const size_t stride = REF_NC * REF_LIMBS;
  ct_to_flat(proof_flat + 0 * stride, siz ? siz + 0 * REF_NC : NULL, pi->h);
  ct_to_flat(proof_flat + 1 * stride, siz ? siz + 1 * REF_NC : NULL, pi->hat_h);
  ct_to_flat(proof_flat + 2 * stride, siz ? siz + 2 * REF_NC : NULL, pi->hat_v);
  ct_to_flat(proof_flat + 3 * stride, siz ? siz + 3 * REF_NC : NULL, pi->v_w);
  ct_to_flat(proof_flat + 4 * stride, siz ? siz + 4 * REF_NC : NULL, pi->b_w);
  proof_clear(pi);
  mpz_clear(w);
}
#endif

/* snark.c:192-250 */
int ref_verifier(const uint8_t *ssp, const uint64_t abs_in[3], const uint64_t *sk_flat,
                 const uint64_t *proof_flat) {
  vrs_t vrs;
  proof_t pi;
  vrs->alpha = abs_in[0];
  vrs->beta = abs_in[1];
  vrs->s = abs_in[2];
  sk_from_flat(vrs->sk, sk_flat);
  proof_init(pi);
  const size_t stride = REF_NC * REF_LIMBS;
  ct_from_flat(pi->h, proof_flat + 0 * stride);
  ct_from_flat(pi->hat_h, proof_flat + 1 * stride);
  ct_from_flat(pi->hat_v, proof_flat + 2 * stride);
  ct_from_flat(pi->v_w, proof_flat + 3 * stride);
  ct_from_flat(pi->b_w, proof_flat + 4 * stride);
  bool ok = verifier((ssp_t)ssp, vrs, pi);
  proof_clear(pi);
  key_clear(vrs->sk);
  return ok ? 1 : 0;
}

/* wall-clock of the three protocol phases as benchmark_snark.c:56-82 reports them */
int ref_benchmark_snark(double secs[3]) {
  struct timespec t0, t1;
  ssp_t ssp = calloc(1, SSP_SIZE);
  mpz_t witness;
  mpz_init(witness);
  random_ssp(witness, ssp);
  crs_t crs;
  crs_init(crs);
  vrs_t vrs;
  proof_t pi;
  proof_init(pi);
#define LAP(i, stmt)                                                                    \
  clock_gettime(CLOCK_MONOTONIC, &t0);                                                  \
  stmt;                                                                                 \
  clock_gettime(CLOCK_MONOTONIC, &t1);                                                  \
  secs[i] = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec)
  LAP(0, setup(crs, vrs, ssp));
  LAP(1, prover(pi, crs, ssp, witness));
  bool ok;
  LAP(2, ok = verifier(ssp, vrs, pi));
#undef LAP
  proof_clear(pi);
  key_clear(vrs->sk);
  crs_clear(crs);
  mpz_clear(witness);
  free(ssp);
  return ok ? 1 : 0;
}
