/* Stand-in for the part of FLINT's <flint/nmod_poly.h> that the SSP SNARK host code uses.
 *
 * FLINT is not installed in this image.  The struct layouts below are FLINT 2.x's
 * (nmod_t = {n, ninv, norm}; nmod_poly_struct = {coeffs, alloc, length, mod}) so that
 * objects are interchangeable with a real FLINT build; the functions are implemented in
 * c_lwe_snarks_b200/host/nmod_poly.c (3-prime NTT multiplication + Newton division).
 * Results over F_p are canonical residues, so any correct implementation is
 * bit-identical to FLINT's.
 */
#ifndef MFB200_COMPAT_FLINT_NMOD_POLY_H
#define MFB200_COMPAT_FLINT_NMOD_POLY_H

/* FLINT's own headers pull these in; programs written against it (the reference's tests) rely on that */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <gmp.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef long slong;
typedef unsigned long ulong;
typedef ulong flint_bitcnt_t;

typedef struct {
  mp_limb_t n;
  mp_limb_t ninv;
  flint_bitcnt_t norm;
} nmod_t;

typedef struct {
  mp_ptr coeffs;
  slong alloc;
  slong length;
  nmod_t mod;
} nmod_poly_struct;

typedef nmod_poly_struct nmod_poly_t[1];

void nmod_poly_init(nmod_poly_t poly, mp_limb_t n);
void nmod_poly_clear(nmod_poly_t poly);
void nmod_poly_fit_length(nmod_poly_t poly, slong alloc);
void nmod_poly_zero(nmod_poly_t poly);
void nmod_poly_set(nmod_poly_t a, const nmod_poly_t b);
void nmod_poly_set_coeff_ui(nmod_poly_t poly, slong j, ulong c);
void nmod_poly_add(nmod_poly_t res, const nmod_poly_t a, const nmod_poly_t b);
void nmod_poly_sub(nmod_poly_t res, const nmod_poly_t a, const nmod_poly_t b);
void nmod_poly_scalar_mul_nmod(nmod_poly_t res, const nmod_poly_t a, mp_limb_t c);
void nmod_poly_mul(nmod_poly_t res, const nmod_poly_t a, const nmod_poly_t b);
void nmod_poly_pow(nmod_poly_t res, const nmod_poly_t a, ulong e);
void nmod_poly_divrem(nmod_poly_t q, nmod_poly_t r, const nmod_poly_t a, const nmod_poly_t b);
void nmod_poly_div(nmod_poly_t q, const nmod_poly_t a, const nmod_poly_t b);
void nmod_poly_rem(nmod_poly_t r, const nmod_poly_t a, const nmod_poly_t b);
mp_limb_t nmod_poly_evaluate_nmod(const nmod_poly_t poly, mp_limb_t c);
int nmod_poly_equal(const nmod_poly_t a, const nmod_poly_t b);

static inline slong nmod_poly_length(const nmod_poly_t poly) { return poly->length; }
static inline slong nmod_poly_degree(const nmod_poly_t poly) { return poly->length - 1; }
static inline mp_limb_t nmod_poly_modulus(const nmod_poly_t poly) { return poly->mod.n; }
static inline int nmod_poly_is_zero(const nmod_poly_t poly) { return poly->length == 0; }
static inline ulong nmod_poly_get_coeff_ui(const nmod_poly_t poly, slong j) {
  return (j >= poly->length) ? 0 : poly->coeffs[j];
}

#ifdef __cplusplus
}
#endif
#endif /* MFB200_COMPAT_FLINT_NMOD_POLY_H */
