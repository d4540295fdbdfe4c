/* mangiafuoco_b200.h — source-level drop-in for the public C interface of mmaker/c-lwe-snarks.
 *
 * The reference's "plugin interface" is its five headers (aes.h, entropy.h, lwe.h, ssp.h, snark.h); programs
 * written against them (its test_*.c and benchmark_*.c mains) compile unchanged against the forwarding
 * headers of the same names in this directory and link against libmangiafuoco_b200.so, whose inner loops
 * run on the GPU through the C-ABI of include/mfb200.h.  Names, argument meaning, ownership and (absent)
 * error reporting are the reference's; the citations give the reference definition each item mirrors
 * (paths under the reference's src/).
 *
 * Differences a caller can observe, all deliberate:
 *   - GAMMA_D / GAMMA_M (lwe.h:14-21) are run-time values: they default to the reference's compile-time
 *     choice (by NDEBUG) and can be changed with mf_set_instance() before any object is created.
 *   - every getrandom(2) the reference issues goes through mf_set_entropy_source(), default getrandom(2);
 *     the call sequence and sizes are the reference's (SURVEY.md §7 hard part 3).
 *   - the 7 noise bits that errdist_uniform leaves uninitialised (entropy.c:34-40 via lwe.c:60-63) are 0.
 *   - failures of the device layer are fatal (message on stderr, abort()): there is no CPU fallback.
 */
#ifndef MANGIAFUOCO_B200_H
#define MANGIAFUOCO_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <strings.h>
#include <sys/types.h>
#include <unistd.h>

#include <flint/nmod_poly.h>
#include <gmp.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- mpz internals the reference exposes through gmp-impl.h:8-10 ------------------------------------- */
#define SIZ(x) ((x)->_mp_size)
#define PTR(x) ((x)->_mp_d)
#define ALLOC(x) ((x)->_mp_alloc)
#define BITS_TO_LIMBS(n) (((n) + (GMP_NUMB_BITS - 1)) / GMP_NUMB_BITS)
/* gmp-impl.h:15-29.  As in the reference, MPN_NORMALIZE does not stop at NLIMBS == 0: only use it on a non-zero
 * limb array (the library's own code normalises with a bounded loop instead). */
#define UNLIKELY(cond) __GMP_UNLIKELY(cond)
#define MPN_NORMALIZE(DST, NLIMBS)         \
  do {                                     \
    while (1) {                            \
      if ((DST)[(NLIMBS)-1] != 0) break;   \
      (NLIMBS)--;                          \
    }                                      \
  } while (0)
#define MPZ_NEWALLOC(z, n) (UNLIKELY((n) > ALLOC(z)) ? (mp_ptr)_mpz_realloc(z, n) : PTR(z))

/* ---- parameters (lwe.h:14-31) ------------------------------------------------------------------------- */
#define GAMMA_N 1470
#define GAMMA_LOGQ 736
#define GAMMA_P 0xfffffffbUL
#define GAMMA_LU 10
#define GAMMA_LOG_SMUDGING 640
#define GAMMA_LOG_SIGMA 556
#define LOGQ_BYTES 92UL
#define LOGP_BYTES 4
#define CT_BYTES (LOGQ_BYTES)

#ifdef NDEBUG
#define MF_DEFAULT_D (1UL << 15)
#define MF_DEFAULT_M (21845UL)
#else
#define MF_DEFAULT_D (1UL << 8)
#define MF_DEFAULT_M (1UL << 6)
#endif
/* instance size: degree bound D and number of SSP polynomials M (M must be divisible by 8 only for random_ssp's
 * witness draw, as in the reference) */
void mf_set_instance(size_t D, size_t M);
size_t mf_gamma_d(size_t compile_time_default);
size_t mf_gamma_m(size_t compile_time_default);
#define GAMMA_D (mf_gamma_d(MF_DEFAULT_D))
#define GAMMA_M (mf_gamma_m(MF_DEFAULT_M))

/* ---- entropy injection ---------------------------------------------------------------------------------- */
/* Replacement for the reference's direct getrandom(2) calls (lwe.h:101, lwe.c:54, entropy.c:36, snark.c:40,
 * ssp.c:43,56,62).  fn must fill `len` bytes; NULL restores getrandom(2). */
typedef void (*mf_entropy_fn)(void *buf, size_t len, void *arg);
void mf_set_entropy_source(mf_entropy_fn fn, void *arg);
void mf_entropy(void *buf, size_t len);
/* device used by the library (default: $MF_B200_DEVICE or 0); call before the first GPU-backed function */
void mf_set_device(int device);

/* ---- aes.h:21-40 -------------------------------------------------------------------------------------- */
typedef struct mf_aes_key aes_key_t; /* opaque: the 32-byte key and a read-ahead window of keystream */
struct aesctr {
  uint64_t nonce;
  aes_key_t *key;
  uint64_t ctr;      /* next AES block to be generated (aes.c:122-133) */
  uint8_t remb[16];
  size_t rem;        /* bytes of block ctr-1 not handed out yet: stream position = 16*ctr - rem */
};
#define CTR(x) ((*(x))->ctr)
#define REM(x) ((*(x))->rem)
typedef struct aesctr *aesctr_ptr;
typedef struct aesctr aesctr_t[1];

void aesctr_init(aesctr_ptr stream, const uint8_t *key, const uint64_t nonce);
void aesctr_prg(aesctr_ptr stream, void *outbuf, size_t count);
void aesctr_clear(aesctr_ptr stream);

/* ---- entropy.h:35-72 ------------------------------------------------------------------------------------ */
#ifndef GRND_NONBLOCK
#define GRND_NONBLOCK 0x0001
#endif
/* the reference's programs call getrandom() directly for their own seeds: route it through the hook too */
static inline ssize_t mf_getrandom(void *buffer, size_t length, unsigned int flags) {
  (void)flags;
  mf_entropy(buffer, length);
  return (ssize_t)length;
}
#ifndef MF_KEEP_LIBC_GETRANDOM
#define getrandom mf_getrandom
#endif

typedef uint8_t rseed_t[32 + 8]; /* nonce(8) || key(32), entropy.c:58-61 */
typedef aesctr_t rng_t[1];

void rng_init(rng_t rs, uint8_t *rseed);
void rng_clear(rng_t rs);
void rng_seek(rng_t prg, size_t count);
static inline void rng_gen(rng_t prg, void *out, size_t count) { aesctr_prg((aesctr_ptr)prg, out, count); }
#define RNG_INIT(rs)               \
  do {                             \
    rseed_t rseed_;                \
    mf_entropy(rseed_, sizeof(rseed_t)); \
    rng_init(rs, rseed_);          \
  } while (0)

/* entropy.h:56 declares it, no reference source defines or calls it: kept for link compatibility, does nothing */
void mpz_entropy_init(void);
void mpz2_urandomb(mpz_ptr rop, rng_t prg, size_t nbits);
void mpz2_urandomb2(mpz_ptr rop, size_t nbits);
#define mpz2_urandommv(vs, rng, bits, len)                                   \
  do {                                                                       \
    for (size_t i_ = 0; i_ < (len); i_++) mpz2_urandomb((vs)[i_], rng, bits); \
  } while (0)
#define mpz2_urandombv2(vs, bits, len)                                   \
  do {                                                                   \
    for (size_t i_ = 0; i_ < (len); i_++) mpz2_urandomb2((vs)[i_], bits); \
  } while (0)

/* ---- lwe.h:33-118 ----------------------------------------------------------------------------------------- */
typedef mpz_t sk_t[GAMMA_N];
typedef mpz_t ct_t[GAMMA_N + 1];

void key_gen(sk_t sk);
void key_clear(sk_t sk);
void errdist_uniform(mpz_t e);
void ct_init(ct_t ct);
void ct_clear(ct_t ct);
void ct_zero(ct_t rop);
void ct_export(uint8_t *buf, ct_t ct);
void ct_import(ct_t ct, rng_t rng, uint8_t *buf);
void decompress_encryption(ct_t c, rng_t rs, mpz_t b);
void regev_encrypt2(ct_t c, rng_t rs, sk_t sk, mpz_t m, void (*chi)(mpz_t));
void mpz_add_dotp(mpz_t rop, mpz_t a[], mpz_t b[], size_t len);
void regev_decrypt(mpz_t m, sk_t sk, ct_t ct);
void ct_smudge(ct_t ct);
void ct_add(ct_t rop, ct_t a, ct_t b);
void ct_mul_ui(ct_t rop, ct_t a, uint64_t b);
void ct_addmul_ui(ct_t rop, ct_t a, uint64_t b);
void eval_poly(ct_t rop, rng_t rng, uint8_t (*c8)[CT_BYTES], nmod_poly_t coeffs, size_t d);

static inline void mpz_dotp(mpz_t rop, mpz_t a[], mpz_t b[], size_t len) {
  mpz_set_ui(rop, 0);
  mpz_add_dotp(rop, a, b, len);
}
static inline void regev_encrypt(ct_t c, rng_t rs, sk_t sk, mpz_t m) { regev_encrypt2(c, rs, sk, m, errdist_uniform); }
static inline uint64_t rand_modp(void) {
  uint64_t rop;
  mf_entropy(&rop, sizeof(rop));
  return rop % GAMMA_P;
}
#define mpz_initv(vs, len)                                              \
  do {                                                                  \
    for (size_t i_ = 0; i_ < (len); i_++) mpz_init2((vs)[i_], GAMMA_LOGQ); \
  } while (0)
#define mpz_clearv(vs, len)                                   \
  do {                                                        \
    for (size_t i_ = 0; i_ < (len); i_++) mpz_clear((vs)[i_]); \
  } while (0)
#define ct_clearv(vs, len)                                   \
  do {                                                       \
    for (size_t i_ = 0; i_ < (len); i_++) ct_clear((vs)[i_]); \
  } while (0)

/* lwe.h:108-118.  For a >= 0 this is a mod 2^704, NOT mod 2^736: the reference masks limb 11 and then
 * truncates the size to 11 limbs, which drops it.  Negative values are left alone, as under NDEBUG there. */
void modq(mpz_t a);

/* ---- ssp.h:6-14 ------------------------------------------------------------------------------------------- */
#define SSP_SIZE (GAMMA_D * 8 * (GAMMA_M + 3))
#define ssp_t_offset 0
#define ssp_v_offset(i) (GAMMA_D * 8 * ((i) + 1))
void nmod_poly_import(nmod_poly_t *pp, void *buf, size_t degree);
void nmod_poly_export(void *buf, nmod_poly_t *pp, size_t degree);
void random_ssp(mpz_t input, uint8_t *circuit);

/* ---- snark.h:6-51 ----------------------------------------------------------------------------------------- */
#define CRS_SIZE (CT_BYTES * (2 * GAMMA_D + GAMMA_M + 1 + 2))
#define CTR_CT (CT_BYTES * GAMMA_N)
#define CTR_S 0
#define CTR_AS (CTR_CT * GAMMA_D)
#define CTR_BT (2 * CTR_CT * GAMMA_D)
#define CTR_BV (2 * CTR_CT * GAMMA_D + CTR_CT)

struct proof {
  ct_t h;
  ct_t hat_h;
  ct_t hat_v;
  ct_t v_w;
  ct_t b_w;
};
struct vrs {
  uint64_t alpha;
  uint64_t beta;
  uint64_t s;
  sk_t sk;
};
struct crs {
  rseed_t seed;
  uint8_t (*s)[CT_BYTES];
  uint8_t (*as)[CT_BYTES];
  uint8_t (*v)[CT_BYTES];
  uint8_t *t;
};
typedef uint8_t *ssp_t;
typedef struct crs crs_t[1];
typedef struct proof proof_t[1];
typedef struct vrs vrs_t[1];

void crs_init(crs_t crs);
void crs_clear(crs_t crs);
void proof_init(proof_t pi);
void proof_clear(proof_t pi);
void setup(crs_t crs, vrs_t vrs, ssp_t ssp);
void prover(proof_t pi, crs_t crs, ssp_t ssp, mpz_t witness);
bool verifier(ssp_t ssp, vrs_t vrs, proof_t pi);

/* ---- additions (not in the reference) ------------------------------------------------------------------ */
/* Keep the two big CRS regions (s, as) expanded in HBM across prover() calls for this crs: the AES
 * regeneration of the a-vectors is then paid once instead of per proof.  mf_crs_release frees them.
 * The resident copy is a snapshot of crs->seed / crs->s / crs->as at the time of the call: call it AFTER setup()
 * or mf_crs_read().  Both of those (and crs_clear) drop a resident copy of the crs they overwrite, so a stale
 * copy is never used; a caller that rewrites the record arrays by hand must call mf_crs_release itself. */
void mf_crs_make_resident(crs_t crs);
void mf_crs_release(crs_t crs);
/* Number of GPUs the resident regions are sharded over (by ciphertext index), all driven by the calling thread:
 * n devices starting at the library's device (mf_set_device / $MF_B200_DEVICE), default $MF_B200_DEVICES or 1.
 * With n > 1 mf_crs_make_resident spreads the two regions over the GPUs' HBM (D = 2^20 needs 2 x 136 GB) and
 * prover() runs every lincomb on all of them at once, combining the partial sums over NVLink peer memory; results
 * are bit-identical.  Call before mf_crs_make_resident.  `spread` = 0 puts all members on the library's device
 * (testing on a one-GPU box). */
void mf_set_devices(int n, int spread);
/* The same for the SSP instance: the dense blob is uploaded once (as u32 residues) and the Newton inverse that the
 * prover's division h = (v^2 - 1)/t needs is cached with it; prover() then only ships the witness.
 * setup() and prover() do this BY THEMSELVES for the blob they are given (unless $MF_B200_NO_AUTO_SSP is set): the
 * reference's own programs, which know nothing of this call, then move the blob over PCIe once instead of in every
 * call.  Staleness rule: a resident copy is keyed on the blob's address and instance size and carries a fingerprint
 * (FNV-1a over 66 samples of 64 bytes spread over the blob) that is re-checked on every use — a blob that was
 * regenerated, or freed and re-allocated at the same address, is detected, the copy is dropped and the host blob is
 * used.  An in-place edit that misses every sample is NOT detected: call mf_ssp_release(ssp) after such an edit. */
void mf_ssp_make_resident(ssp_t ssp);
void mf_ssp_release(ssp_t ssp);
/* Persistence (the reference only sketches a "crs.mfuoco" mmap, benchmark_snark.c:23-24): the CRS file is a 24-byte
 * header (magic, D, M) + seed + the 92-byte records in stream order s, as, t, v; a proof file holds the five
 * ciphertexts as 88-byte magnitudes plus the sign of each b.  Return 0, or -1 with errno set. */
int mf_crs_write(const char *path, crs_t crs);
int mf_crs_read(const char *path, crs_t crs);
int mf_proof_write(const char *path, proof_t pi);
int mf_proof_read(const char *path, proof_t pi);
/* number of GPU kernels launched by the library so far */
uint64_t mf_gpu_launches(void);

#ifdef __cplusplus
}
#endif

/* Programs compiled against the reference pick their instance size at compile time (NDEBUG); record that
 * default in the library before main() so that library-side code (setup, prover, ...) agrees with the caller. */
#ifndef MF_NO_AUTO_INSTANCE
__attribute__((constructor)) static void mf_register_default_instance_(void) {
  (void)mf_gamma_d(MF_DEFAULT_D);
  (void)mf_gamma_m(MF_DEFAULT_M);
}
#endif

#endif /* MANGIAFUOCO_B200_H */
