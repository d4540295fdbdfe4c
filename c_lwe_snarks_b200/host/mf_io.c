/* CRS and proof persistence.
 *
 * The reference only sketches this (commented-out mmap of "crs.mfuoco", benchmark_snark.c:23-24,34-53; CRS_SIZE,
 * snark.h:6).  The CRS file keeps the wire layout the reference already uses in memory — seed, then 92-byte b records
 * in STREAM order (snark.h:8-12): s[0..D), as[0..D), t, v[0..M-1) — behind a small header, so that a prover can map
 * it and hand the record arrays to eval_poly / mfb_region_create unchanged.  A proof file holds the five ciphertexts
 * as 1471 x 88-byte little-endian magnitudes (every coordinate is < 2^704 after modq) plus the sign of each b
 * coordinate (ct_smudge can leave it negative, lwe.c:65-76).
 */
#include "mf_internal.h"

#include <errno.h>

static const char CRS_MAGIC[8] = {'M', 'F', 'U', 'O', 'C', 'O', '1', 0};
static const char PROOF_MAGIC[8] = {'M', 'F', 'P', 'R', 'O', 'O', 'F', '1'};

static int put(FILE *f, const void *p, size_t n) { return fwrite(p, 1, n, f) == n ? 0 : -1; }
static int get(FILE *f, void *p, size_t n) { return fread(p, 1, n, f) == n ? 0 : -1; }

int mf_crs_write(const char *path, crs_t crs) {
  const uint64_t D = GAMMA_D, M = GAMMA_M;
  FILE *f = fopen(path, "wb");
  if (!f) return -1;
  int rc = put(f, CRS_MAGIC, 8) | put(f, &D, 8) | put(f, &M, 8) | put(f, crs->seed, sizeof(rseed_t)) |
           put(f, crs->s, CT_BYTES * D) | put(f, crs->as, CT_BYTES * D) | put(f, crs->t, CT_BYTES) |
           put(f, crs->v, CT_BYTES * (M - 1));
  if (fclose(f) != 0) rc = -1;
  return rc ? -1 : 0;
}

/* crs must have been crs_init'ed for the current instance size; fails (-1, errno = EINVAL) on a size mismatch */
int mf_crs_read(const char *path, crs_t crs) {
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  char magic[8];
  uint64_t D = 0, M = 0;
  int rc = get(f, magic, 8) | get(f, &D, 8) | get(f, &M, 8);
  if (!rc && (memcmp(magic, CRS_MAGIC, 8) || D != GAMMA_D || M != GAMMA_M)) {
    errno = EINVAL;
    rc = -1;
  }
  if (!rc)
    rc = get(f, crs->seed, sizeof(rseed_t)) | get(f, crs->s, CT_BYTES * D) | get(f, crs->as, CT_BYTES * D) |
         get(f, crs->t, CT_BYTES) | get(f, crs->v, CT_BYTES * (M - 1));
  fclose(f);
  if (!rc) mf_crs_release(crs); /* any resident copy belongs to the old contents */
  return rc ? -1 : 0;
}

int mf_proof_write(const char *path, proof_t pi) {
  mpz_t *el[5] = {pi->h, pi->hat_h, pi->hat_v, pi->v_w, pi->b_w};
  FILE *f = fopen(path, "wb");
  if (!f) return -1;
  int rc = put(f, PROOF_MAGIC, 8);
  uint8_t *buf = malloc((size_t)(GAMMA_N + 1) * 88);
  if (!buf) mf_die("malloc");
  for (int k = 0; k < 5 && !rc; k++) {
    uint64_t neg = 0, limbs[MF_LIMBS];
    for (size_t i = 0; i <= GAMMA_N; i++) {
      if (mpz_sizeinbase(el[k][i], 2) > 704) {
        errno = ERANGE;
        rc = -1;
        break;
      }
      const int n = mf_to_flat(limbs, el[k][i]);
      if (n && i != GAMMA_N) {
        errno = ERANGE;
        rc = -1;
        break;
      }
      if (n) neg = 1;
      memcpy(buf + i * 88, limbs, 88);
    }
    if (!rc) rc = put(f, &neg, 8) | put(f, buf, (size_t)(GAMMA_N + 1) * 88);
  }
  free(buf);
  if (fclose(f) != 0) rc = -1;
  return rc ? -1 : 0;
}

int mf_proof_read(const char *path, proof_t pi) {
  mpz_t *el[5] = {pi->h, pi->hat_h, pi->hat_v, pi->v_w, pi->b_w};
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  char magic[8];
  int rc = get(f, magic, 8);
  if (!rc && memcmp(magic, PROOF_MAGIC, 8)) {
    errno = EINVAL;
    rc = -1;
  }
  uint8_t *buf = malloc((size_t)(GAMMA_N + 1) * 88);
  if (!buf) mf_die("malloc");
  for (int k = 0; k < 5 && !rc; k++) {
    uint64_t neg = 0;
    rc = get(f, &neg, 8) | get(f, buf, (size_t)(GAMMA_N + 1) * 88);
    if (rc) break;
    for (size_t i = 0; i <= GAMMA_N; i++) mf_bytes_to_mpz(el[k][i], buf + i * 88, 88);
    if (neg) mpz_neg(el[k][GAMMA_N], el[k][GAMMA_N]);
  }
  free(buf);
  fclose(f);
  return rc ? -1 : 0;
}
