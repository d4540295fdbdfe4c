/* Internal helpers of the host drop-in layer (not installed). */
#ifndef MF_INTERNAL_H
#define MF_INTERNAL_H

#define MF_NO_AUTO_INSTANCE
#include "mangiafuoco_b200.h"
#include "mfb200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MF_LIMBS 11 /* 64-bit limbs of a value mod 2^704 */

/* the process-wide device context; creates it on first use, aborts loudly when there is no usable GPU */
mfb_ctx *mf_gpu(void);
void mf_die(const char *what) __attribute__((noreturn));
/* start creating the device context in a background thread (no-op when it exists or is under way); never fails loudly */
void mf_gpu_prefetch(void);
#define MF_GPU(call)                                                             \
  do {                                                                           \
    if ((call) != MFB_OK) mf_die(#call);                                         \
  } while (0)

/* is an entropy hook installed (mf_set_entropy_source)?  Then draws must happen serially, in the reference's order. */
int mf_entropy_hooked(void);

/* stream position of an rng (aes.c bookkeeping: 16*ctr - rem) and the seed it was built from */
uint64_t mf_rng_pos(rng_t rng);
void mf_rng_advance(rng_t rng, uint64_t nbytes);
const uint8_t *mf_rng_seed(rng_t rng);

/* mpz <-> flat little-endian limbs mod 2^704.  to_flat returns 1 when z is negative (magnitude stored). */
int mf_to_flat(uint64_t out[MF_LIMBS], mpz_srcptr z);
void mf_from_flat(mpz_ptr z, const uint64_t in[MF_LIMBS]);
void mf_ct_to_flat(uint64_t *out, ct_t ct, const char *who); /* aborts on a negative coordinate */
void mf_ct_from_flat(ct_t ct, const uint64_t *in);
void mf_bytes_to_mpz(mpz_ptr z, const uint8_t *bytes, size_t n);


/* Phase timing in the reference's timeit.h format ("name\tseconds" lines on stderr) when $MF_B200_TRACE is set. */
double mf_now(void);
void mf_trace(const char *name, double t0);

#endif
