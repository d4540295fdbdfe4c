/* ssp.h of the reference: dense SSP wire format and the random (degenerate) instance generator.
 * Host-only: this is the input side of the hot path (SURVEY.md §2 marks it out of scope for kernels). */
#include "mf_internal.h"

/* wire format (ssp.h:6-9): polynomial k occupies bytes [8*D*k, 8*D*(k+1)), 8-byte little-endian coefficients,
 * k = 0 is t, k = i + 1 is v_i */
void nmod_poly_export(void *buf, nmod_poly_t *pp, size_t degree) { /* ssp.c:18-26 */
  uint64_t *out = buf;
  for (size_t i = 0; i < degree; i++) out[i] = nmod_poly_get_coeff_ui(*pp, (slong)i);
}

void nmod_poly_import(nmod_poly_t *pp, void *buf, size_t degree) { /* ssp.c:28-34: coefficients reduced mod p */
  const uint64_t *in = buf;
  for (size_t i = 0; i < degree; i++) nmod_poly_set_coeff_ui(*pp, (slong)i, in[i]);
}

/* ssp.c:37-77.  Entropy, in order: M/8 bytes for the witness, then M draws of 8*D bytes (v_0 .. v_{M-1}).
 * t := v_0 + sum_{i>=1, w_{i-1}=1} v_i - 1, so that t | (v^2 - 1) with v = v_0 + sum w_i v_i. */
void random_ssp(mpz_t input, uint8_t *circuit) {
  mf_gpu_prefetch();
  const size_t D = GAMMA_D, M = GAMMA_M;
  uint8_t *buf = malloc(8 * D);
  if (!buf) mf_die("malloc");
  mpz2_urandomb2(input, M);

  nmod_poly_t v, t, one;
  nmod_poly_init(v, GAMMA_P);
  nmod_poly_init(t, GAMMA_P);
  nmod_poly_init(one, GAMMA_P);
  nmod_poly_set_coeff_ui(one, 0, 1);
  for (size_t i = 0; i < M; i++) {
    mf_entropy(buf, 8 * D);
    nmod_poly_import(&v, buf, D);
    nmod_poly_export(circuit + ssp_v_offset(i), &v, D);
    if (i == 0 || mpz_tstbit(input, i - 1)) nmod_poly_add(t, t, v);
  }
  nmod_poly_sub(t, t, one);
  nmod_poly_export(circuit + ssp_t_offset, &t, D);
  nmod_poly_clear(v);
  nmod_poly_clear(t);
  nmod_poly_clear(one);
  free(buf);
}
