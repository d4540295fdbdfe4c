/* snark.h of the reference: setup / prover / verifier orchestration over the GPU kernels.
 *
 * The protocol steps, their order, the stream regions (snark.h:8-12) and the order of entropy draws are the
 * reference's; what changes is that each loop over ciphertexts is ONE batched device call:
 *   setup     2D + M Regev encryptions            -> one mfb_encrypt_cb (entropy drawn piece by piece while the
 *                                                    device encrypts the previous piece)          (snark.c:75-110)
 *   prover    b_w = delta*CT_t + sum_{w_i} CT_v_i -> one mfb_eval_poly with an index list (snark.c:143-155)
 *             v_w, hat_v, h, hat_h                 -> two two-vector fused passes (mfb_eval_poly2), or — after
 *                                                    mf_crs_make_resident + mf_ssp_make_resident — ONE device
 *                                                    pipeline (mfb_prove_resident; mfb_set_prove_resident when the
 *                                                    regions are sharded over several GPUs)       (snark.c:157-174)
 *   verifier  5 decryptions + the test-error dot  -> one mfb_decrypt   (snark.c:204-208, 238)
 * Polynomial arithmetic over F_p (FLINT in the reference) runs on the device (k_poly.cu).
 */
#include "mf_internal.h"

#define FLAT_CT (MFB_FLAT_CT_U64)

void proof_init(proof_t pi) {
  mf_gpu_prefetch();
  ct_init(pi->h);
  ct_init(pi->hat_h);
  ct_init(pi->hat_v);
  ct_init(pi->v_w);
  ct_init(pi->b_w);
}

void proof_clear(proof_t pi) {
  ct_clear(pi->h);
  ct_clear(pi->hat_h);
  ct_clear(pi->hat_v);
  ct_clear(pi->v_w);
  ct_clear(pi->b_w);
}

/* ------------------------------------------------------------------ resident CRS regions (addition) */
struct resident {
  struct crs *owner;
  mfb_region *s, *as;          /* one GPU */
  mfb_set_region *ms, *mas;    /* sharded over the device set */
  size_t d;
  struct resident *next;
};
static struct resident *g_resident = NULL;

/* device set (mf_set_devices): created on first use, lives as long as the process */
static int g_ndev = 0, g_spread = 1;
static mfb_set *g_set = NULL;

void mf_set_devices(int n, int spread) {
  if (g_set) {
    if (g_resident) {
      fprintf(stderr, "mangiafuoco_b200: mf_set_devices: release the resident CRS regions first\n");
      abort();
    }
    mfb_set_destroy(g_set);
    g_set = NULL;
  }
  g_ndev = n < 1 ? 1 : n;
  g_spread = spread;
}

static mfb_set *device_set(void) {
  if (g_ndev == 0) {
    const char *e = getenv("MF_B200_DEVICES");
    g_ndev = e ? atoi(e) : 1;
    if (g_ndev < 1) g_ndev = 1;
  }
  if (g_ndev == 1) return NULL;
  if (!g_set) {
    int devs[MFB_PEER_MAX];
    if (g_ndev > MFB_PEER_MAX) g_ndev = MFB_PEER_MAX;
    const int base = mfb_ctx_device(mf_gpu());
    for (int i = 1; i < g_ndev; i++) devs[i - 1] = g_spread ? base + i : base;
    if (mfb_set_create(mf_gpu(), devs, g_ndev - 1, &g_set) != MFB_OK) {
      fprintf(stderr, "mangiafuoco_b200: mfb_set_create(%d devices) failed: %s\n", g_ndev, mfb_set_last_error());
      abort();
    }
  }
  return g_set;
}

static struct resident *resident_find(struct crs *crs) {
  for (struct resident *r = g_resident; r; r = r->next)
    if (r->owner == crs) return r;
  return NULL;
}

void mf_crs_release(crs_t crs) {
  struct resident **pp = &g_resident;
  while (*pp) {
    struct resident *r = *pp;
    if (r->owner == crs) {
      mfb_region_destroy(mf_gpu(), r->s);
      mfb_region_destroy(mf_gpu(), r->as);
      mfb_set_region_destroy(g_set, r->ms);
      mfb_set_region_destroy(g_set, r->mas);
      *pp = r->next;
      free(r);
    } else {
      pp = &r->next;
    }
  }
}

void mf_crs_make_resident(crs_t crs) {
  double t0 = mf_now();
  mf_crs_release(crs);
  mf_trace("make_resident.release_previous", t0);
  t0 = mf_now();
  struct resident *r = calloc(1, sizeof(*r));
  if (!r) mf_die("malloc");
  r->owner = crs;
  r->d = GAMMA_D;
  /* 2 x D x 129 536 B of HBM; when that does not fit (D = 2^20 needs 272 GB on one GPU) the regions simply stay
   * non-resident and prover() keeps regenerating a from AES in-kernel — slower, same result */
  mfb_set *set = device_set();
  if (set) { /* sharded by ciphertext index over the set's GPUs */
    int rc = mfb_set_region_create(set, crs->seed, CTR_S, (const uint8_t *)crs->s, r->d, &r->ms);
    mf_trace("make_resident.set_region_s", t0);
    t0 = mf_now();
    if (rc == MFB_OK) rc = mfb_set_region_create(set, crs->seed, CTR_AS, (const uint8_t *)crs->as, r->d, &r->mas);
    mf_trace("make_resident.set_region_as", t0);
    if (rc != MFB_OK) {
      fprintf(stderr, "mangiafuoco_b200: mf_crs_make_resident over %d devices: %s; the CRS stays non-resident\n",
              mfb_set_size(set), mfb_set_last_error());
      mfb_set_region_destroy(set, r->ms);
      mfb_set_region_destroy(set, r->mas);
      free(r);
      return;
    }
    r->next = g_resident;
    g_resident = r;
    return;
  }
  int rc = mfb_region_create(mf_gpu(), crs->seed, CTR_S, (const uint8_t *)crs->s, r->d, &r->s);
  mf_trace("make_resident.region_s", t0);
  t0 = mf_now();
  if (rc == MFB_OK && getenv("MF_B200_ONE_REGION")) {
    /* test hook: behave as if the second region did not fit, so that the mixed resident / fused prover path (what
     * D = 2^20 takes on one GPU) can be checked at small sizes */
    r->as = NULL;
  } else if (rc == MFB_OK) {
    rc = mfb_region_create(mf_gpu(), crs->seed, CTR_AS, (const uint8_t *)crs->as, r->d, &r->as);
    mf_trace("make_resident.region_as", t0);
    if (rc == MFB_ENOMEM) { /* one region fits, two do not (D = 2^20 on one GPU): keep the first, fuse the other */
      fprintf(stderr, "mangiafuoco_b200: mf_crs_make_resident: %s; only the s region is resident\n", mfb_last_error());
      r->as = NULL;
      rc = MFB_OK;
    }
  }
  if (rc != MFB_OK) {
    if (rc != MFB_ENOMEM) mf_die("mfb_region_create");
    fprintf(stderr, "mangiafuoco_b200: mf_crs_make_resident: %s; the CRS stays non-resident\n", mfb_last_error());
    mfb_region_destroy(mf_gpu(), r->s);
    free(r);
    return;
  }
  r->next = g_resident;
  g_resident = r;
}

/* ------------------------------------------------------------------ resident SSP blobs (addition) */
/* A resident copy is keyed on the blob's address and instance size, and carries a FINGERPRINT of the blob (FNV-1a over
 * 66 samples of 64 bytes spread over it).  Every use re-computes the fingerprint (a few microseconds): a blob that was
 * rewritten, or freed and re-allocated at the same address, no longer matches, the stale copy is dropped and the caller
 * falls back to the host blob.  (A modification that misses every sample is not detected: callers that edit single
 * coefficients in place must call mf_ssp_release themselves — documented in the header.) */
struct resident_ssp {
  const uint8_t *owner;
  mfb_ssp *h;
  size_t d, m;
  uint64_t fp;
  struct resident_ssp *next;
};
static struct resident_ssp *g_resident_ssp = NULL;

static uint64_t ssp_fingerprint(const uint8_t *blob, size_t D, size_t M) {
  const size_t bytes = 8 * D * (M + 1), chunk = 64;
  uint64_t hsh = 0xcbf29ce484222325ull;
  const size_t nsamp = 66;
  for (size_t k = 0; k < nsamp; k++) {
    size_t off = bytes <= chunk ? 0 : (size_t)((unsigned __int128)(bytes - chunk) * k / (nsamp - 1));
    off &= ~(size_t)7;
    const size_t n = bytes < chunk ? bytes : chunk;
    for (size_t i = 0; i < n; i++) hsh = (hsh ^ blob[off + i]) * 0x100000001b3ull;
  }
  return hsh ^ (uint64_t)D ^ ((uint64_t)M << 32);
}

void mf_ssp_release(ssp_t ssp) {
  struct resident_ssp **pp = &g_resident_ssp;
  while (*pp) {
    struct resident_ssp *r = *pp;
    if (r->owner == ssp) {
      mfb_ssp_destroy(mf_gpu(), r->h);
      *pp = r->next;
      free(r);
    } else {
      pp = &r->next;
    }
  }
}

static mfb_ssp *ssp_make_resident(ssp_t ssp, int quiet) {
  mf_ssp_release(ssp);
  struct resident_ssp *r = calloc(1, sizeof(*r));
  if (!r) mf_die("malloc");
  r->owner = ssp;
  r->d = GAMMA_D;
  r->m = GAMMA_M;
  r->fp = ssp_fingerprint(ssp, r->d, r->m);
  const int rc = mfb_ssp_create(mf_gpu(), (const uint64_t *)ssp, r->d, r->m, &r->h);
  if (rc != MFB_OK) {
    if (!quiet) fprintf(stderr, "mangiafuoco_b200: mf_ssp_make_resident: %s; the SSP stays on the host\n", mfb_last_error());
    free(r);
    return NULL;
  }
  r->next = g_resident_ssp;
  g_resident_ssp = r;
  { /* the first polynomial step over a resident blob allocates its workspaces and captures its CUDA graph (4-70 ms,
     * measured): pay that here, where the 8 D (M + 1)-byte upload dwarfs it, not inside the first prover() call */
    const uint64_t no_bits = 0;
    const uint32_t *wvh = NULL;
    if (!getenv("MF_B200_NO_RESERVE")) (void)mfb_ssp_prover_polys_resident_dev(mf_gpu(), r->h, &no_bits, 1, 0, &wvh);
  }
  return r->h;
}

static mfb_ssp *resident_ssp_find(ssp_t ssp);
void mf_ssp_make_resident(ssp_t ssp) {
  double t0 = mf_now();
  if (!resident_ssp_find(ssp)) (void)ssp_make_resident(ssp, 0); /* (already resident and unchanged: nothing to do) */
  mf_trace("ssp_make_resident", t0);
}

static mfb_ssp *resident_ssp_find(ssp_t ssp) {
  for (struct resident_ssp *r = g_resident_ssp; r; r = r->next)
    if (r->owner == ssp && r->d == GAMMA_D && r->m == GAMMA_M) {
      if (ssp_fingerprint(ssp, r->d, r->m) == r->fp) return r->h;
      mf_ssp_release(ssp); /* the blob changed under the resident copy: drop it, use the host blob */
      return NULL;
    }
  return NULL;
}

/* setup(), prover() and verifier() all read the same dense blob (8 D (M + 3) bytes: 5.7 GB at the reference's default
 * instance).  Unless $MF_B200_NO_AUTO_SSP is set, the first of them to see a blob uploads it ONCE (as u32 residues) and
 * the others find it resident — the reference's own programs, which cannot call mf_ssp_make_resident, then pay the PCIe
 * transfer once instead of in every call.  Returns NULL when disabled or when the blob does not fit. */
static mfb_ssp *resident_ssp_auto(ssp_t ssp) {
  mfb_ssp *h = resident_ssp_find(ssp);
  if (h) return h;
  if (getenv("MF_B200_NO_AUTO_SSP")) return NULL;
  return ssp_make_resident(ssp, 1);
}

void crs_init(crs_t crs) { /* snark.c:35-48 */
  mf_gpu_prefetch();
  mf_entropy(crs->seed, sizeof(rseed_t));
  crs->s = malloc(CT_BYTES * GAMMA_D);
  crs->as = malloc(CT_BYTES * GAMMA_D);
  crs->v = malloc(CT_BYTES * GAMMA_M);
  crs->t = malloc(CT_BYTES);
  if (!crs->s || !crs->as || !crs->v || !crs->t) perror("Error allocating memory");
}

void crs_clear(crs_t crs) {
  mf_crs_release(crs);
  free(crs->s);
  free(crs->as);
  free(crs->v);
  free(crs->t);
}

/* ------------------------------------------------------------------ setup (snark.c:57-115) */
static void draw_entropy(void *user, uint8_t *dst, size_t nbytes) {
  (void)user;
  mf_entropy(dst, nbytes);
}

/* out[i] = x0 * r^i mod p for i < n, as 8 interleaved chains (chain c starts at c * n/8 with x0 * r^(c n/8)) */
static uint64_t powmod_p(uint64_t b, uint64_t e) {
  uint64_t acc = 1;
  for (b %= GAMMA_P; e; e >>= 1, b = b * b % GAMMA_P)
    if (e & 1) acc = acc * b % GAMMA_P;
  return acc;
}
static void geometric(uint64_t *out, size_t n, uint64_t x0, uint64_t r) {
  enum { CH = 8 };
  const size_t per = n / CH;
  uint64_t x[CH];
  for (int c = 0; c < CH; c++) x[c] = x0 % GAMMA_P * powmod_p(r, (uint64_t)c * per) % GAMMA_P;
  for (size_t i = 0; i < per; i++)
    for (int c = 0; c < CH; c++) {
      out[c * per + i] = x[c];
      x[c] = x[c] * r % GAMMA_P;
    }
  uint64_t y = per ? out[CH * per - 1] * r % GAMMA_P : x0 % GAMMA_P; /* the n % 8 elements after the last chain */
  for (size_t i = CH * per; i < n; i++, y = y * r % GAMMA_P) out[i] = y;
}

void setup(crs_t crs, vrs_t vrs, ssp_t ssp) {
  const size_t D = GAMMA_D, M = GAMMA_M, count = 2 * D + M;
  double t0 = mf_now();
  /* the records are about to be rewritten: a resident copy of this crs (mf_crs_make_resident) would be stale, and
   * prover() finds regions by crs pointer — drop it, as mf_crs_read() does */
  mf_crs_release(crs);
  vrs->alpha = rand_modp();
  vrs->beta = rand_modp();
  vrs->s = rand_modp();
  key_gen(vrs->sk);
  mf_trace("setup.key_gen", t0);
  t0 = mf_now();

  /* plaintexts in stream order (snark.h:8-12): s^i, alpha*s^i, beta*t(s), beta*v_i(s) for i = 1..M-1.  The two
   * geometric sequences are cut into 8 independent chains each (x_{i+1} = x_i * s is one dependent multiply + reduce per
   * element: 2^21 of them in a row would cost ~10 ms at D = 2^20) */
  uint64_t *msg = malloc(count * 8);
  if (!msg) mf_die("malloc");
  geometric(msg, D, 1, vrs->s);
  geometric(msg + D, D, vrs->alpha, vrs->s);
  /* t(s) and v_i(s): the reference's M Horner passes (snark.c:97-110) as one batched device evaluation over the
   * dense blob [t, v_0, ..., v_{M-1}] (ssp.h:6-9); v_0(s) is computed and not used, as its slot is skipped there */
  uint64_t *vals = malloc((M + 1) * 8);
  if (!vals) mf_die("malloc");
  mfb_ssp *rssp = resident_ssp_auto(ssp); /* uploads the blob once; prover() and verifier() then find it on the device */
  mf_trace("setup.ssp_resident", t0);
  if (rssp)
    MF_GPU(mfb_ssp_eval_resident(mf_gpu(), rssp, 0, M + 1, vrs->s, vals));
  else
    MF_GPU(mfb_ssp_eval(mf_gpu(), (const uint64_t *)ssp, D, M + 1, vrs->s, vals));
  msg[2 * D] = (vals[0] * vrs->beta) % GAMMA_P;
  for (size_t i = 1; i < M; i++) msg[2 * D + i] = (vals[i + 1] * vrs->beta) % GAMMA_P;
  free(vals);
  mf_trace("setup.plaintexts+evaluations", t0);
  t0 = mf_now();

  /* per encryption the reference draws 69 noise bytes, then 1 sign byte (lwe.c:85-87): 70 bytes each, in order.
   * The draws go through a callback piece by piece, so that getrandom(2) for the next piece runs while the device
   * encrypts the previous one.  Record k of the call is ciphertext k of the stream: s, as, t, v. */
  uint64_t *skf = malloc(MFB_FLAT_SK_U64 * 8);
  if (!skf) mf_die("malloc");
  for (size_t i = 0; i < GAMMA_N; i++) mf_to_flat(skf + i * MF_LIMBS, vrs->sk[i]);
  mfb_set *set = device_set();
  if (set && !mf_entropy_hooked()) {
    /* OS entropy has no order to preserve: every GPU of the set takes a contiguous range of the ciphertexts, driven by
     * its own host thread that draws that range's entropy, and writes its records straight into the CRS arrays */
    const mfb_c8_segment segs[4] = {{0, D, (uint8_t *)crs->s}, {D, D, (uint8_t *)crs->as}, {2 * D, 1, crs->t},
                                    {2 * D + 1, M - 1, (uint8_t *)crs->v}};
    if (mfb_set_encrypt_par(set, crs->seed, 0, skf, msg, draw_entropy, NULL, MFB_ENT_BYTES, MFB_ENT_BYTES - 1, count, NULL, segs, 4) !=
        MFB_OK) {
      fprintf(stderr, "mangiafuoco_b200: mfb_set_encrypt_par failed: %s\n", mfb_set_last_error());
      abort();
    }
    mf_trace("setup.entropy+encrypt (one thread per GPU)", t0);
  } else if (set) { /* hooked entropy: drawn by this thread in the reference's order, the pieces spread over the GPUs */
    uint8_t *recs = malloc(count * CT_BYTES);
    if (!recs) mf_die("malloc");
    if (mfb_set_encrypt_cb(set, crs->seed, 0, skf, msg, draw_entropy, NULL, MFB_ENT_BYTES, MFB_ENT_BYTES - 1, count, recs) != MFB_OK) {
      fprintf(stderr, "mangiafuoco_b200: mfb_set_encrypt_cb failed: %s\n", mfb_set_last_error());
      abort();
    }
    mf_trace("setup.entropy+encrypt", t0);
    memcpy(crs->s, recs, D * CT_BYTES);
    memcpy(crs->as, recs + D * CT_BYTES, D * CT_BYTES);
    memcpy(crs->t, recs + 2 * D * CT_BYTES, CT_BYTES);
    memcpy(crs->v, recs + (2 * D + 1) * CT_BYTES, (M - 1) * CT_BYTES);
    free(recs);
  } else {
    /* one GPU: the records of a piece come back on a second stream while the next piece is encrypted, straight into the
     * CRS arrays */
    const mfb_c8_segment segs[4] = {{0, D, (uint8_t *)crs->s}, {D, D, (uint8_t *)crs->as}, {2 * D, 1, crs->t},
                                    {2 * D + 1, M - 1, (uint8_t *)crs->v}};
    MF_GPU(mfb_encrypt_cb_segs(mf_gpu(), crs->seed, 0, skf, msg, draw_entropy, NULL, MFB_ENT_BYTES, MFB_ENT_BYTES - 1, count, segs,
                               M > 1 ? 4 : 3));
    mf_trace("setup.entropy+encrypt", t0);
  }
  explicit_bzero(skf, MFB_FLAT_SK_U64 * 8); /* the flat copy of the secret key does not outlive the call */
  free(skf);
  free(msg);
}

/* ------------------------------------------------------------------ prover (snark.c:117-190) */
static void lincomb_pair_resident(ct_t rop0, ct_t rop1, mfb_region *reg, mfb_set_region *mreg, const uint64_t *poly0,
                                  const uint64_t *poly1) {
  const size_t D = GAMMA_D;
  uint64_t *acc = malloc(2 * FLAT_CT * 8);
  uint32_t *co = malloc(2 * D * 4);
  if (!acc || !co) mf_die("malloc");
  mf_ct_to_flat(acc, rop0, "prover");
  mf_ct_to_flat(acc + FLAT_CT, rop1, "prover");
  for (size_t i = 0; i < D; i++) {
    co[i] = (uint32_t)poly0[i];
    co[D + i] = (uint32_t)poly1[i];
  }
  if (mreg) {
    if (mfb_set_region_lincomb2(g_set, mreg, co, co + D, D, acc, acc + FLAT_CT) != MFB_OK) {
      fprintf(stderr, "mangiafuoco_b200: mfb_set_region_lincomb2 failed: %s\n", mfb_set_last_error());
      abort();
    }
  } else {
    MF_GPU(mfb_region_lincomb2(mf_gpu(), reg, 0, co, co + D, D, acc, acc + FLAT_CT));
  }
  mf_ct_from_flat(rop0, acc);
  mf_ct_from_flat(rop1, acc + FLAT_CT);
  free(acc);
  free(co);
}

static void lincomb_pair(ct_t rop0, ct_t rop1, crs_t crs, int which_as, const uint64_t *poly0, const uint64_t *poly1) {
  uint64_t *acc = malloc(2 * FLAT_CT * 8);
  if (!acc) mf_die("malloc");
  mf_ct_to_flat(acc, rop0, "prover");
  mf_ct_to_flat(acc + FLAT_CT, rop1, "prover");
  mfb_set *set = device_set();
  if (set) { /* sharded by ciphertext index: every GPU regenerates the a-vectors of its range */
    if (mfb_set_eval_poly2(set, crs->seed, which_as ? CTR_AS : CTR_S, (const uint8_t *)(which_as ? crs->as : crs->s), poly0, poly1,
                           GAMMA_D, acc, acc + FLAT_CT) != MFB_OK) {
      fprintf(stderr, "mangiafuoco_b200: mfb_set_eval_poly2 failed: %s\n", mfb_set_last_error());
      abort();
    }
  } else {
    MF_GPU(mfb_eval_poly2(mf_gpu(), crs->seed, which_as ? CTR_AS : CTR_S, (const uint8_t *)(which_as ? crs->as : crs->s), poly0,
                          poly1, GAMMA_D, acc, acc + FLAT_CT));
  }
  mf_ct_from_flat(rop0, acc);
  mf_ct_from_flat(rop1, acc + FLAT_CT);
  free(acc);
}

void prover(proof_t pi, crs_t crs, ssp_t ssp, mpz_t witness) {
  const size_t D = GAMMA_D, M = GAMMA_M;
  const uint64_t delta = rand_modp();
  if (SIZ(witness) < 0) {
    fprintf(stderr, "mangiafuoco_b200: prover: negative witness\n");
    abort();
  }

  double t0 = mf_now();
  mfb_ssp *rssp = resident_ssp_auto(ssp);
  struct resident *res = resident_find(crs);
  const int all_resident = rssp && res && res->d == D && (res->ms || (res->s && res->as));

  /* polynomial step (snark.c:138-169) on the device: w = delta*t + sum_{w_i} v_i, v = w + v_0 (l_u = 0),
   * h = (v^2 - 1) / t  — FLINT's scalar_mul / add / pow / div in the reference.  With the SSP and both CRS regions
   * resident the coefficients never leave the device(s): see the fused pipeline below. */
  uint64_t *pw = NULL, *pv = NULL, *ph = NULL;
  if (!all_resident) {
    pw = malloc(3 * D * 8);
    if (!pw) mf_die("malloc");
    pv = pw + D;
    ph = pw + 2 * D;
    if (rssp)
      MF_GPU(mfb_ssp_prover_polys_resident(mf_gpu(), rssp, PTR(witness), (size_t)SIZ(witness), delta, pw, pv, ph));
    else
      MF_GPU(mfb_ssp_prover_polys(mf_gpu(), (const uint64_t *)ssp, D, M, PTR(witness), (size_t)SIZ(witness), delta, pw, pv, ph));
  }
  mf_trace("prover.polys", t0);
  t0 = mf_now();

  /* b_w = delta * CT_t + sum_{witness bit i-1} CT_v[i-1]: ciphertext k of the region at CTR_BT is t for k = 0
   * and v[k-1] after it.  The reference regenerates every a-vector to advance its stream; only the selected
   * ones contribute, so only those are expanded here. */
  uint8_t *recs = malloc(M * CT_BYTES);
  uint64_t *co = malloc(M * 8);
  uint32_t *idx = malloc(M * 4);
  if (!recs || !co || !idx) mf_die("malloc");
  memcpy(recs, crs->t, CT_BYTES);
  memcpy(recs + CT_BYTES, crs->v, (M - 1) * CT_BYTES);
  size_t nsel = 0;
  co[nsel] = delta;
  idx[nsel++] = 0;
  for (size_t i = 1; i < M; i++) {
    if (mpz_tstbit(witness, i - 1)) {
      co[nsel] = 1;
      idx[nsel++] = (uint32_t)i;
    }
  }
  const int b_w_in_pipeline = all_resident; /* everything resident: b_w rides in the device pipeline below */
  if (!b_w_in_pipeline) {
    uint64_t *acc = calloc(FLAT_CT, 8); /* ct_import overwrites pi->b_w: start from zero */
    if (!acc) mf_die("malloc");
    MF_GPU(mfb_eval_poly(mf_gpu(), crs->seed, CTR_BT, recs, co, idx, nsel, acc));
    mf_ct_from_flat(pi->b_w, acc);
    free(acc);
  }
  free(co);
  free(idx);

  mf_trace("prover.b_w", t0);
  t0 = mf_now();
  if (all_resident) {
    /* SSP and regions resident: ONE device pipeline — polynomial step, then both two-vector passes with the
     * coefficients read where they were computed (sharded regions: every GPU fetches its slices over NVLink);
     * only the witness bits and the four accumulators cross PCIe */
    static uint64_t *acc = NULL; /* kept across proofs: 647 KB would be a fresh mmap + page faults every call */
    if (!acc && !(acc = malloc(5 * FLAT_CT * 8))) mf_die("malloc");
    { /* eval_poly accumulates into rop (lwe.c:176-186) — but a proof normally starts from proof_init: all zero */
      mpz_t *el[4] = {pi->v_w, pi->h, pi->hat_v, pi->hat_h};
      int zero = 1;
      for (int k = 0; k < 4 && zero; k++)
        for (size_t i = 0; i <= GAMMA_N; i++)
          if (SIZ(el[k][i]) != 0) {
            zero = 0;
            break;
          }
      if (zero) {
        memset(acc, 0, 4 * FLAT_CT * 8);
      } else {
        for (int k = 0; k < 4; k++) mf_ct_to_flat(acc + k * FLAT_CT, el[k], "prover");
      }
    }
    if (res->ms) {
      if (mfb_set_prove_resident_bw(g_set, rssp, res->ms, res->mas, PTR(witness), (size_t)SIZ(witness), delta, crs->seed, CTR_BT, recs,
                                    M, acc, acc + FLAT_CT, acc + 2 * FLAT_CT, acc + 3 * FLAT_CT, acc + 4 * FLAT_CT) != MFB_OK) {
        fprintf(stderr, "mangiafuoco_b200: mfb_set_prove_resident_bw failed: %s\n", mfb_set_last_error());
        abort();
      }
    } else {
      MF_GPU(mfb_prove_resident_bw(mf_gpu(), rssp, res->s, res->as, PTR(witness), (size_t)SIZ(witness), delta, crs->seed, CTR_BT,
                                   recs, M, acc, acc + FLAT_CT, acc + 2 * FLAT_CT, acc + 3 * FLAT_CT, acc + 4 * FLAT_CT));
    }
    mf_ct_from_flat(pi->b_w, acc + 4 * FLAT_CT);
    mf_ct_from_flat(pi->v_w, acc);
    mf_ct_from_flat(pi->h, acc + FLAT_CT);
    mf_ct_from_flat(pi->hat_v, acc + 2 * FLAT_CT);
    mf_ct_from_flat(pi->hat_h, acc + 3 * FLAT_CT);
  } else {
    /* per region: resident in HBM -> one pass at the HBM roofline, else a regenerated from AES in-kernel; two scalar
     * vectors per pass either way */
    struct resident *r = (res && res->d == D) ? res : NULL;
    if (r && (r->s || r->ms))
      lincomb_pair_resident(pi->v_w, pi->h, r->s, r->ms, pw, ph);
    else
      lincomb_pair(pi->v_w, pi->h, crs, 0, pw, ph);
    if (r && (r->as || r->mas))
      lincomb_pair_resident(pi->hat_v, pi->hat_h, r->as, r->mas, pv, ph);
    else
      lincomb_pair(pi->hat_v, pi->hat_h, crs, 1, pv, ph);
  }
  free(pw);
  free(recs);
  mf_trace(all_resident ? "prover.polys+lincombs (device pipeline)" : "prover.lincombs", t0);
  t0 = mf_now();

  /* smudging, in the reference's order: v_w twice, b_w never (snark.c:185-189) */
  ct_smudge(pi->h);
  ct_smudge(pi->hat_h);
  ct_smudge(pi->hat_v);
  ct_smudge(pi->v_w);
  ct_smudge(pi->v_w);
  mf_trace("prover.smudge", t0);
}

/* ------------------------------------------------------------------ verifier (snark.c:192-250) */
bool verifier(ssp_t ssp, vrs_t vrs, proof_t pi) {
  const size_t D = GAMMA_D;
  const uint64_t p = GAMMA_P;
  /* t(s) and v_0(s): the first two polynomials of the blob, one batched device evaluation (snark.c:197-201,214-215) */
  uint64_t ts_v0s[2];
  mfb_ssp *rssp = resident_ssp_find(ssp);
  if (rssp)
    MF_GPU(mfb_ssp_eval_resident(mf_gpu(), rssp, 0, 2, vrs->s, ts_v0s));
  else
    MF_GPU(mfb_ssp_eval(mf_gpu(), (const uint64_t *)ssp, D, 2, vrs->s, ts_v0s));
  const uint64_t t_s = ts_v0s[0], v0_s = ts_v0s[1];

  /* one batch: decrypt h, hat_h, hat_v, v_w, b_w; the dot product of b_w is also the test-error input */
  mpz_t *elems[5] = {pi->h, pi->hat_h, pi->hat_v, pi->v_w, pi->b_w};
  uint64_t *cts = malloc(5 * FLAT_CT * 8), *skf = malloc(MFB_FLAT_SK_U64 * 8);
  if (!cts || !skf) mf_die("malloc");
  uint8_t neg[5];
  for (int k = 0; k < 5; k++) {
    for (size_t i = 0; i < GAMMA_N; i++)
      if (mf_to_flat(cts + k * FLAT_CT + i * MF_LIMBS, elems[k][i])) {
        fprintf(stderr, "mangiafuoco_b200: verifier: negative a coordinate in proof element %d\n", k);
        abort();
      }
    neg[k] = (uint8_t)mf_to_flat(cts + k * FLAT_CT + (size_t)GAMMA_N * MF_LIMBS, elems[k][GAMMA_N]);
  }
  for (size_t i = 0; i < GAMMA_N; i++) mf_to_flat(skf + i * MF_LIMBS, vrs->sk[i]);
  uint64_t dec[5], dots[5 * MF_LIMBS];
  MF_GPU(mfb_decrypt(mf_gpu(), skf, cts, neg, 5, dec, dots));
  free(cts);
  explicit_bzero(skf, MFB_FLAT_SK_U64 * 8); /* the flat copy of the secret key does not outlive the call */
  free(skf);
  const uint64_t h_s = dec[0], hath_s = dec[1], hatv_s = dec[2], w_s = dec[3], b_s = dec[4];
  const uint64_t v_s = (v0_s + w_s) % p;

  bool result = false;
  /* eq-pke */
  if ((unsigned __int128)h_s * vrs->alpha % p != hath_s) goto end;
  if ((unsigned __int128)v_s * vrs->alpha % p != hatv_s) goto end;
  /* eq-div: v_s^2 - 1 - h_s*t_s == 0 mod p */
  {
    const uint64_t lhs = (uint64_t)((unsigned __int128)v_s * v_s % p);
    const uint64_t rhs = (uint64_t)(((unsigned __int128)h_s * t_s + 1) % p);
    if (lhs != rhs) goto end;
  }
  /* eq-lin */
  if ((unsigned __int128)w_s * vrs->beta % p != b_s) goto end;
  /* test-error (snark.c:237-241): -<b_w, sk> ceil-divided by p; the size comparison is the reference's */
  {
    mpz_t test;
    mpz_init(test);
    mf_from_flat(test, dots + 4 * MF_LIMBS);
    mpz_neg(test, test);
    mpz_cdiv_q_ui(test, test, GAMMA_P);
    const bool too_big = SIZ(test) >= GAMMA_LOG_SMUDGING / 8;
    mpz_clear(test);
    if (too_big) goto end;
  }
  result = true;
end:
  return result;
}
