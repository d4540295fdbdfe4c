/* mfb200.h — C-ABI of the B200 (sm_100a) kernels behind the SSP-SNARK hot path of mmaker/c-lwe-snarks.
 *
 * This is the device boundary: plain pointers and sizes, C linkage, no torch / GMP / FLINT types.
 * The reference has no FFI layer — its "plugin interface" is its own C headers — so each entry point
 * below names the reference function (path:line under the reference's src/) whose inner loop it
 * replaces; the drop-in C layer that keeps the reference's mpz_t-typed signatures (include/mangiafuoco/
 * lwe.h, snark.h, entropy.h) is a thin adapter over these calls.
 *
 * Formats (all little-endian, all unsigned):
 *   seed      40 bytes: nonce(8) || AES-256 key(32)                         entropy.h:35, entropy.c:58-61
 *   record    92 bytes: the b coordinate as ct_export writes it            lwe.c:115-119 (top 4 bytes are 0)
 *   flat ct   MFB_NC (1471) coordinates x MFB_L64 (11) uint64 limbs, coordinate-major: the value of every
 *             coordinate mod 2^704 — the modulus modq() really implements    lwe.h:108-118
 *   flat sk   MFB_N (1470) coordinates x 11 uint64 limbs (bits >= 704 of a key never reach any output)
 *   resident  (HBM) "tile-planar" layout of a ciphertext array, MFB_PLANAR_U64 u64 per ciphertext: the 1472
 *             coordinates (1470 = b, 1471 = zero padding) are cut into 23 tiles of 64; (ct i, 64-bit limb row j,
 *             coordinate c) sits at u64 index i*16192 + (c/64)*704 + j*64 + c%64, so that (ct, tile) is one
 *             contiguous 5632-byte block for a TMA bulk copy
 *   row-planar  secret keys on the device: limb row j, coordinate c at j*1472 + c
 *   scalars   uint32 < p = 2^32 - 5 (nmod_poly coefficients; any uint32 is computed exactly)
 *
 * Every function returns 0 on success or a negative MFB_E* code; mfb_last_error() gives the text.
 * There is no CPU fallback: without a CUDA device mfb_ctx_create fails.
 *
 * Two flavours:
 *   host flavour  (no suffix)  host pointers in, host pointers out; copies + kernels on the context's own
 *                              stream; returns after the result is in host memory.
 *   _dev flavour               device pointers, asynchronous on the caller's stream (a cudaStream_t passed
 *                              as void*; NULL = the legacy default stream); no allocation, no sync.
 */
#ifndef MFB200_H
#define MFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFB_N 1470
#define MFB_NC 1471
#define MFB_NCP 1472
#define MFB_L64 11
#define MFB_CT_BYTES 92
#define MFB_CTR_CT (MFB_CT_BYTES * MFB_N) /* 135240 stream bytes per ciphertext, snark.h:8 */
#define MFB_P 0xfffffffbu
#define MFB_FLAT_CT_U64 ((size_t)MFB_NC * MFB_L64)   /* 16181 */
#define MFB_FLAT_SK_U64 ((size_t)MFB_N * MFB_L64)    /* 16170 */
#define MFB_PLANAR_U64 ((size_t)MFB_L64 * MFB_NCP)   /* 16192 u64 = 129536 B per resident ciphertext */
#define MFB_ENT_BYTES 70 /* per encryption: 69 noise bytes (lwe.c:62) + 1 sign byte (lwe.c:54,87) */
#define MFB_ALGO_BYTES_PER_MAC 129448 /* 1471 x 88: algorithmic HBM bytes of one ciphertext-MAC */

#define MFB_OK 0
#define MFB_ECUDA (-1)   /* a CUDA call failed */
#define MFB_EARG (-2)    /* bad argument */
#define MFB_ENODEV (-3)  /* no usable CUDA device */
#define MFB_ENOMEM (-4)

#if defined(__GNUC__)
#define MFB_API __attribute__((visibility("default")))
#else
#define MFB_API
#endif

typedef struct mfb_ctx mfb_ctx;
typedef struct mfb_ssp mfb_ssp;       /* an SSP blob kept on the device (mfb_ssp_create) */
/* entropy callback of mfb_encrypt_cb / mfb_set_encrypt_cb: fill dst[0..nbytes) */
typedef void (*mfb_entropy_fn)(void *user, uint8_t *dst, size_t nbytes);

/* ---- context ------------------------------------------------------------------------------ */
MFB_API int mfb_ctx_create(mfb_ctx **out, int device);
MFB_API void mfb_ctx_destroy(mfb_ctx *ctx);
/* optional: allocate the pinned staging buffers now instead of inside the first call that needs them */
MFB_API int mfb_ctx_warm(mfb_ctx *ctx);
/* optional: pay the cold-start costs of an instance of D constraints and M SSP polynomials now — the scratch buffers of
 * setup / prover / verifier at their final sizes, the second stream, events, pinned entropy buffers, and the first launch
 * of every kernel of the path (one-element dry runs).  The drop-in calls it from its background warm-up thread, so that a
 * one-shot program that times its only setup() and prover() calls (the reference's benchmark_snark) measures the work,
 * not allocation and module loading. */
MFB_API int mfb_ctx_reserve(mfb_ctx *ctx, size_t D, size_t M);
MFB_API const char *mfb_last_error(void);
MFB_API int mfb_device_sm_count(mfb_ctx *ctx);
MFB_API int mfb_ctx_device(mfb_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
MFB_API uint64_t mfb_launch_count(mfb_ctx *ctx);
MFB_API int mfb_sync(mfb_ctx *ctx);
/* Per-kernel timing for the roofline: between begin and end every lincomb / eval_poly call brackets its
 * dominant kernel (k_lincomb / k_evalpoly) with CUDA events on the launching stream; end waits for them and
 * returns the summed duration and the number of launches timed (at most 1024). */
MFB_API int mfb_profile_begin(mfb_ctx *ctx);
MFB_API int mfb_profile_end(mfb_ctx *ctx, double *sum_ms, int *count);

/* ---- K2: AES-256-CTR stream -------------------------------------------------------------- */
/* out[0..nbytes) = stream bytes [offset, offset+nbytes) of `seed`.
 * Replaces rng_init + rng_seek + rng_gen / aesctr_prg (entropy.c:46-61, aes.c:104-144). */
MFB_API int mfb_stream(mfb_ctx *ctx, const uint8_t seed[40], uint64_t offset, uint8_t *out, size_t nbytes);
MFB_API int mfb_stream_dev(mfb_ctx *ctx, const uint8_t seed[40], uint64_t offset, uint8_t *out_dev, size_t nbytes,
                   void *stream);

/* cts_dev[k] (resident layout) <- ct_import(rng at offset + k*MFB_CTR_CT, c8[k]) for k < count (lwe.c:122-126):
 * the a-vector from the stream, b from the 92-byte record.  This is how a CRS region is made resident. */
MFB_API int mfb_expand_dev(mfb_ctx *ctx, const uint8_t seed[40], uint64_t offset, const uint8_t *c8_dev, size_t count,
                   uint64_t *cts_dev, void *stream);

/* ---- K1: ciphertext linear combination over resident ciphertexts -------------------------- */
/* rop = (rop_in + sum_{i<d} coeffs[i] * cts[i]) mod 2^704, coordinate-wise: the loop of eval_poly
 * (lwe.c:176-186) = d x ct_addmul_ui (lwe.c:141-149) with modq (lwe.h:108-118); also ct_mul_ui and ct_add
 * (lwe.c:131-157) as the d = 1, 2 cases.  cts_dev is in the resident layout, d < 2^32 per call; rop_in may be
 * NULL (zero) or equal to rop_out.  One call at a time per context (it owns the partial-sum workspace). */
MFB_API int mfb_lincomb_dev(mfb_ctx *ctx, const uint64_t *cts_dev, const uint32_t *coeffs_dev, size_t d,
                    const uint64_t *rop_in_dev, uint64_t *rop_out_dev, void *stream);
/* two scalar vectors in one pass over the same resident ciphertexts (half the HBM traffic of two calls) */
MFB_API int mfb_lincomb2_dev(mfb_ctx *ctx, const uint64_t *cts_dev, const uint32_t *coeffs0_dev, const uint32_t *coeffs1_dev,
                     size_t d, const uint64_t *rop0_in_dev, uint64_t *rop0_out_dev, const uint64_t *rop1_in_dev,
                     uint64_t *rop1_out_dev, void *stream);
/* Building blocks of the prover pipeline (snark.c:157-174: two regions, two scalar vectors each): the two two-vector
 * passes leave their partial sums in the context's workspace (pass = 0, 1; both over equally many ciphertexts), then ONE
 * finish launch produces the four flat accumulators rop_out4 + k * rop_stride_u64, k = 0..3 in the order (pass 0 vector
 * 0, pass 0 vector 1, pass 1 vector 0, pass 1 vector 1); rop_in4 (same layout) may be NULL (zero).  mfb_peer_finish4_dev
 * is that finish fused with the peer-memory exchange of all four (see below): one kernel per rank. */
MFB_API int mfb_lincomb2_partials_dev(mfb_ctx *ctx, const uint64_t *cts_dev, const uint32_t *coeffs0_dev, const uint32_t *coeffs1_dev,
                              size_t d, int pass, void *stream);
MFB_API int mfb_lincomb_finish4_dev(mfb_ctx *ctx, const uint64_t *rop_in4_dev, uint64_t *rop_out4_dev, size_t rop_stride_u64,
                            void *stream);
/* host flavour over flat host ciphertexts (d small: ct_add / ct_mul_ui / ct_addmul_ui on ct_t objects) */
MFB_API int mfb_lincomb(mfb_ctx *ctx, const uint64_t *cts_flat, const uint32_t *coeffs, size_t d, uint64_t *rop_flat_inout);

/* Resident CRS region: expand once, reuse for every proof.  handle is owned by the context. */
typedef struct mfb_region mfb_region;
MFB_API int mfb_region_create(mfb_ctx *ctx, const uint8_t seed[40], uint64_t offset, const uint8_t *c8, size_t count,
                      mfb_region **out);
/* the same without waiting: the expansion is queued on `stream`; synchronise it before the region is used or another
 * region is created on this context (device sets expand their members' shards concurrently this way) */
MFB_API int mfb_region_create_async(mfb_ctx *ctx, const uint8_t seed[40], uint64_t offset, const uint8_t *c8, size_t count,
                            void *stream, mfb_region **out);
MFB_API void mfb_region_destroy(mfb_ctx *ctx, mfb_region *r);
/* rop += sum_i coeffs[i] * region[first + i], i < d; coeffs/rop in host memory */
MFB_API int mfb_region_lincomb(mfb_ctx *ctx, const mfb_region *r, size_t first, const uint32_t *coeffs, size_t d,
                       uint64_t *rop_flat_inout);

MFB_API int mfb_region_lincomb2(mfb_ctx *ctx, const mfb_region *r, size_t first, const uint32_t *coeffs0,
                        const uint32_t *coeffs1, size_t d, uint64_t *rop0_flat_inout, uint64_t *rop1_flat_inout);
/* device pointer of the region's resident ciphertext array (for the _dev entry points) */
MFB_API const void *mfb_region_cts(const mfb_region *r);
MFB_API size_t mfb_region_count(const mfb_region *r);


/* Other LWE parameter points (BASELINE configs[4]); the reference implements only (1470, 736), so this entry point
 * has no reference counterpart: out = sum_i coeffs[i] * cts[i] mod 2^(64*limbs64), coordinate-wise, for ciphertexts
 * of ncoords coordinates (= n + 1) in the tile-planar layout with limbs64 rows per tile:
 * u64 index (ct i, row j, coordinate c) = i*T*64*L + (c/64)*64*L + j*64 + c%64, T = ceil(ncoords/64), L = limbs64.
 * out has the shape of one ciphertext.  limbs64 in {4,6,8,10,11,12,13,14,16}. */
MFB_API int mfb_lincomb_generic_dev(mfb_ctx *ctx, int limbs64, int ncoords, const uint64_t *cts_dev, const uint32_t *coeffs_dev,
                            size_t d, uint64_t *out_dev, void *stream);

/* Regev encryption at other parameter points (same design as mfb_encrypt_dev; no reference counterpart either):
 * coordinate j of ciphertext k = the ct_bytes = log q / 8 stream bytes at offset + (k n + j) ct_bytes, reduced mod
 * q_eff = 2^(64 limbs64); out record k (ct_bytes bytes) = (e_k p + <sk, a_k> + msg[k]) mod q_eff, p = 2^32 - 5.
 * sk_planar_dev: limb row j of coordinate c at j * sk_stride + c.  e_k as in mfb_encrypt_dev (ent_nbytes <= 8 limbs64). */
MFB_API int mfb_encrypt_generic_dev(mfb_ctx *ctx, int limbs64, int n, int ct_bytes, const uint8_t seed[40], uint64_t offset,
                            const uint64_t *sk_planar_dev, int sk_stride, const uint64_t *msg_dev, const uint8_t *ent_dev,
                            int ent_stride, int ent_nbytes, size_t count, uint8_t *out_c8_dev, void *stream);
/* The tiling mfb_encrypt_generic_dev uses for (n, ct_bytes) — host logic, no device needed (the CPU tests check its
 * invariants): coordinates per tile, tiles per ciphertext, and the padded keystream layout (pad_blocks = ct_bytes / 16 when
 * one free 16-byte slot per coordinate lowers the bank-conflict degree of the consumers' reads, else 0; pad_reciprocal =
 * the multiplier with (b * pad_reciprocal) >> 16 == b / pad_blocks for every block index of a tile). */
MFB_API int mfb_encrypt_generic_plan(int n, int ct_bytes, int *tile, int *ntiles, int *pad_blocks, uint32_t *pad_reciprocal);

/* ---- multi-GPU exchange helpers (one process per GPU; the collective itself is NCCL) ------ */
/* cols_dev[1472][22] u64 <- the 32-bit limbs of flat_dev, widened, so that an elementwise integer sum over
 * ranks is exact; carry: flat_out[c] = (flat_in[c] + sum_l cols[c - c0][l] << 32l) mod 2^704 for the
 * ncoord coordinates this rank owns after the reduce-scatter. */
MFB_API int mfb_columns_split_dev(mfb_ctx *ctx, const uint64_t *flat_dev, uint64_t *cols_dev, void *stream);
MFB_API int mfb_columns_carry_dev(mfb_ctx *ctx, const uint64_t *cols_dev, int c0, int ncoord, const uint64_t *flat_in_dev,
                          uint64_t *flat_out_dev, void *stream);

/* ---- multi-GPU: sharded lincomb with the exchange FUSED into the finish kernel over peer memory ----------
 * No reference counterpart (the reference is single-threaded); SURVEY.md §8e.  The prover's sum over ciphertext
 * indices is sharded across GPUs; every rank owns a symmetric exchange buffer that the others map over NVLink
 * (CUDA IPC between processes: mfb_peer_connect; plain peer access inside one process: mfb_peer_connect_local).
 * mfb_lincomb_peer_dev = mfb_lincomb_dev over this rank's ciphertexts, except that the finish kernel pushes the
 * rank's partial sum into every rank's buffer, waits for the other ranks' and adds them: rop_out on EVERY rank =
 * (rop_in + sum over all ranks and their ciphertexts) mod 2^704, bit-identical to the single-GPU result, in the
 * same two kernel launches as the single-GPU call.  Like any collective, every rank of the group must make the
 * same sequence of *_peer_dev calls (one stream per group).  A rank that waits longer than the group's timeout
 * (default 20 s) gives up; mfb_peer_status then returns MFB_EPEER.
 *   create      allocates the buffer, returns the 64-byte CUDA IPC handle to hand to the other ranks
 *   connect     handles = world x 64 bytes in rank order (e.g. gathered with torch.distributed)
 *   disconnect  unmaps the peers' buffers; call on every rank (and synchronise the ranks) BEFORE any destroy */
#define MFB_EPEER (-5) /* a peer rank did not arrive in time */
#define MFB_PEER_HANDLE_BYTES 64
#define MFB_PEER_MAX 16
typedef struct mfb_peer_group mfb_peer_group;
MFB_API int mfb_peer_create(mfb_ctx *ctx, int world, int rank, mfb_peer_group **out, uint8_t handle_out[MFB_PEER_HANDLE_BYTES]);
MFB_API void *mfb_peer_base(mfb_peer_group *g);
MFB_API int mfb_peer_connect(mfb_ctx *ctx, mfb_peer_group *g, const uint8_t *handles);
MFB_API int mfb_peer_connect_local(mfb_ctx *ctx, mfb_peer_group *g, void *const *bases);
MFB_API int mfb_peer_set_timeout(mfb_peer_group *g, double seconds);
MFB_API int mfb_peer_status(mfb_ctx *ctx, mfb_peer_group *g);
MFB_API int mfb_peer_disconnect(mfb_ctx *ctx, mfb_peer_group *g);
MFB_API void mfb_peer_destroy(mfb_ctx *ctx, mfb_peer_group *g);
MFB_API int mfb_lincomb_peer_dev(mfb_ctx *ctx, mfb_peer_group *g, const uint64_t *cts_dev, const uint32_t *coeffs_dev, size_t d,
                         const uint64_t *rop_in_dev, uint64_t *rop_out_dev, void *stream);
/* The exchange alone: rop_out on every rank = (rop_in + sum over ranks of partial_flat) mod 2^704 for one flat
 * ciphertext per rank — one 23-CTA kernel (push, flags, wait, add).  For back-to-back lincombs run mfb_lincomb_dev on
 * the main stream and this on a side stream: it is co-resident with the next lincomb kernel, so the exchange costs
 * the main stream nothing (sharding.PipelinedPeerShardedLincomb). */
MFB_API int mfb_peer_allreduce_dev(mfb_ctx *ctx, mfb_peer_group *g, const uint64_t *partial_flat_dev, const uint64_t *rop_in_dev,
                           uint64_t *rop_out_dev, void *stream);
/* Up to 4 flat ciphertexts per rank in ONE exchange kernel ("lanes": partial_flat_dev + l * in_stride_u64 in, rop +
 * l * rop_stride_u64 in / out, l < lanes); every rank must pass the same number of lanes. */
MFB_API int mfb_peer_allreduce_lanes_dev(mfb_ctx *ctx, mfb_peer_group *g, const uint64_t *partial_flat_dev, size_t in_stride_u64, int lanes,
                                 const uint64_t *rop_in_dev, uint64_t *rop_out_dev, size_t rop_stride_u64, void *stream);
MFB_API int mfb_peer_finish4_dev(mfb_ctx *ctx, mfb_peer_group *g, const uint64_t *rop_in4_dev, uint64_t *rop_out4_dev,
                         size_t rop_stride_u64, void *stream);
/* The same exchange as TWO kernels: push (finishes the four partial sums, pushes them to every rank, raises the flags —
 * never blocks) and wait (a small kernel: waits for every rank's tiles and adds them).  For ranks whose kernels cannot
 * be assumed co-resident (several members of a device set sharing one GPU): enqueue every rank's push before any wait. */
MFB_API int mfb_peer_finish4_push_dev(mfb_ctx *ctx, mfb_peer_group *g, void *stream);
MFB_API int mfb_peer_wait4_dev(mfb_ctx *ctx, mfb_peer_group *g, const uint64_t *rop_in4_dev, uint64_t *rop_out4_dev,
                       size_t rop_stride_u64, void *stream);
/* ... and for flat ciphertexts (mfb_peer_allreduce_lanes_dev in two launches) */
MFB_API int mfb_peer_push_lanes_dev(mfb_ctx *ctx, mfb_peer_group *g, const uint64_t *partial_flat_dev, size_t in_stride_u64, int lanes,
                            void *stream);
MFB_API int mfb_peer_wait_lanes_dev(mfb_ctx *ctx, mfb_peer_group *g, int lanes, const uint64_t *rop_in_dev, uint64_t *rop_out_dev,
                            size_t rop_stride_u64, void *stream);
/* the same for the fused AES + MAC path (mfb_eval_poly_dev over this rank's ciphertexts) */
MFB_API int mfb_eval_poly_peer_dev(mfb_ctx *ctx, mfb_peer_group *g, const uint8_t seed[40], uint64_t offset, const uint8_t *c8_dev,
                           const uint32_t *coeffs_dev, const uint32_t *idx_dev, size_t d, const uint64_t *rop_in_dev,
                           uint64_t *rop_out_dev, void *stream);

/* ---- device sets: several GPUs driven by ONE host thread ----------------------------------------------------
 * The shape in which the reference's single-threaded prover() (snark.c:117-190) uses a whole NVSwitch box: a set is
 * the primary context plus one context per further device, joined in a same-process peer exchange group; a set
 * region is a CRS region sharded by ciphertext index over the members' HBM (D = 2^20: 2 x 136 GB over 8 GPUs).
 * mfb_set_region_lincomb2 = mfb_region_lincomb2 over the sharded region: every member runs the two-vector lincomb
 * kernel over its shard, one peer all-reduce kernel per vector combines them; host buffers in and out.
 * devices[] may repeat and may include the primary's device (members then share a GPU). */
typedef struct mfb_set mfb_set;
typedef struct mfb_set_region mfb_set_region;
MFB_API int mfb_set_create(mfb_ctx *primary, const int *devices, int ndev, mfb_set **out);
MFB_API void mfb_set_destroy(mfb_set *s);
MFB_API int mfb_set_size(const mfb_set *s);
MFB_API const char *mfb_set_last_error(void);
MFB_API int mfb_set_region_create(mfb_set *s, const uint8_t seed[40], uint64_t offset, const uint8_t *c8, size_t count,
                          mfb_set_region **out);
MFB_API void mfb_set_region_destroy(mfb_set *s, mfb_set_region *r);
MFB_API int mfb_set_region_lincomb2(mfb_set *s, const mfb_set_region *r, const uint32_t *coeffs0, const uint32_t *coeffs1, size_t d,
                            uint64_t *rop0_flat_inout, uint64_t *rop1_flat_inout);
/* mfb_encrypt_cb over a device set: the entropy is drawn by the calling thread piece by piece, in order; piece k is
 * encrypted by member k mod size (setup()'s 2D+M encryptions on a whole box: entropy-bound instead of AES-bound). */
MFB_API int mfb_set_encrypt_cb(mfb_set *s, const uint8_t seed[40], uint64_t offset, const uint64_t *sk_flat, const uint64_t *msg,
                       mfb_entropy_fn draw, void *user, int ent_stride, int ent_nbytes, size_t count, uint8_t *out_c8);
/* The same when the entropy source has no order to preserve (the OS): member i takes the contiguous range i of the
 * ciphertexts and is driven by its own host thread, which draws that range's entropy — draw is called CONCURRENTLY and
 * must be thread-safe — while its device encrypts: entropy, upload, AES and download all scale with the members.  The
 * records go to out_c8, or (out_c8 == NULL) to nsegs segments of the record index space: record k of segment g goes to
 * g.dst + (k - g.first) * 92 (setup(): crs->s, crs->as, crs->t, crs->v without an intermediate copy). */
typedef struct mfb_c8_segment {
  size_t first, count;
  uint8_t *dst;
} mfb_c8_segment;
MFB_API int mfb_set_encrypt_par(mfb_set *s, const uint8_t seed[40], uint64_t offset, const uint64_t *sk_flat, const uint64_t *msg,
                        mfb_entropy_fn draw, void *user, int ent_stride, int ent_nbytes, size_t count, uint8_t *out_c8,
                        const mfb_c8_segment *segs, int nsegs);
/* eval_poly / eval_poly2 with nothing resident, sharded: every member regenerates the a-vectors of its contiguous
 * ciphertext range from AES in-kernel; coeffs1 / rop1 may both be NULL. */
MFB_API int mfb_set_eval_poly2(mfb_set *s, const uint8_t seed[40], uint64_t offset, const uint8_t *c8, const uint64_t *coeffs0,
                       const uint64_t *coeffs1, size_t d, uint64_t *rop0_flat_inout, uint64_t *rop1_flat_inout);
/* mfb_prove_resident[_bw] over sharded regions: the polynomial step is queued on the primary (no host round trip), the
 * other members wait for it on the device, fetch their slices of w, v, h over NVLink, run both two-vector passes over
 * their shards and ONE finish + exchange kernel each for the four accumulators; b_w is computed meanwhile by a member
 * that would otherwise wait for the polynomial step.  One synchronisation, one pinned copy back.
 * A failed set call poisons the set (later calls return MFB_EPEER): destroy it and create a new one. */
MFB_API int mfb_set_prove_resident(mfb_set *s, mfb_ssp *ssp, const mfb_set_region *reg_s, const mfb_set_region *reg_as,
                           const uint64_t *witness_limbs, size_t nlimbs, uint64_t delta, uint64_t *v_w_flat_inout,
                           uint64_t *h_flat_inout, uint64_t *hat_v_flat_inout, uint64_t *hat_h_flat_inout);
MFB_API int mfb_set_prove_resident_bw(mfb_set *s, mfb_ssp *ssp, const mfb_set_region *reg_s, const mfb_set_region *reg_as,
                              const uint64_t *witness_limbs, size_t nlimbs, uint64_t delta, const uint8_t seed[40],
                              uint64_t bt_offset, const uint8_t *bt_recs, size_t M, uint64_t *v_w_flat_inout,
                              uint64_t *h_flat_inout, uint64_t *hat_v_flat_inout, uint64_t *hat_h_flat_inout,
                              uint64_t *b_w_flat_out);

/* ---- K2+K1 fused: eval_poly with a regenerated in-kernel ----------------------------------- */
/* rop += sum_{m<d} coeffs[m] * CT_{k(m)},  k(m) = idx ? idx[m] : m, where CT_k = ct_import(stream at
 * offset + k*MFB_CTR_CT, c8[k]).  Exactly eval_poly (lwe.c:176-186) with the rng positioned at `offset`;
 * with idx it is also the prover's b_w loop (snark.c:143-155), which skips unset witness bits.
 * Coefficients are uint64 as nmod_poly stores them; they must be < 2^32. */
MFB_API int mfb_eval_poly(mfb_ctx *ctx, const uint8_t seed[40], uint64_t offset, const uint8_t *c8, const uint64_t *coeffs,
                  const uint32_t *idx, size_t d, uint64_t *rop_flat_inout);
MFB_API int mfb_eval_poly_dev(mfb_ctx *ctx, const uint8_t seed[40], uint64_t offset, const uint8_t *c8_dev,
                      const uint32_t *coeffs_dev, const uint32_t *idx_dev, size_t d, const uint64_t *rop_in_dev,
                      uint64_t *rop_out_dev, void *stream);

/* Two scalar vectors over the SAME ciphertexts in one pass (half the AES work of two mfb_eval_poly calls):
 * rop0 += sum coeffs0[m] * CT_m, rop1 += sum coeffs1[m] * CT_m.  The prover pairs (v_w, h) over the s region and
 * (hat_v, hat_h) over the as region (snark.c:157-174). */
MFB_API int mfb_eval_poly2(mfb_ctx *ctx, const uint8_t seed[40], uint64_t offset, const uint8_t *c8, const uint64_t *coeffs0,
                   const uint64_t *coeffs1, size_t d, uint64_t *rop0_flat_inout, uint64_t *rop1_flat_inout);
MFB_API int mfb_eval_poly2_dev(mfb_ctx *ctx, const uint8_t seed[40], uint64_t offset, const uint8_t *c8_dev,
                       const uint32_t *coeffs0_dev, const uint32_t *coeffs1_dev, size_t d, const uint64_t *rop0_in_dev,
                       uint64_t *rop0_out_dev, const uint64_t *rop1_in_dev, uint64_t *rop1_out_dev, void *stream);
/* mfb_eval_poly2_dev in two halves, for callers that drive SEVERAL contexts from one thread (device sets): `begin` queues
 * the AES + MAC kernel (it needs the seed and the scalars only), `end` the b-coordinate kernel — the only consumer of the
 * wire records — and the finish.  With c8_host != NULL (and records_from_host != 0 in `begin`) `end` copies the d records
 * from host memory into c8_dev on the context's second stream, beside the running AES kernel; staging pageable memory
 * blocks the calling thread, so queue every context's `begin` before the first `end`.  coeffs1 / rop1 may be NULL (one
 * scalar vector).  The same (d, coeffs) must be passed to both halves; one begin/end pair at a time per context. */
MFB_API int mfb_eval_poly2_begin_dev(mfb_ctx *ctx, const uint8_t seed[40], uint64_t offset, const uint32_t *coeffs0_dev,
                             const uint32_t *coeffs1_dev, size_t d, int records_from_host, void *stream);
MFB_API int mfb_eval_poly2_end_dev(mfb_ctx *ctx, uint8_t *c8_dev, const uint8_t *c8_host, const uint32_t *coeffs0_dev,
                           const uint32_t *coeffs1_dev, size_t d, const uint64_t *rop0_in_dev, uint64_t *rop0_out_dev,
                           const uint64_t *rop1_in_dev, uint64_t *rop1_out_dev, void *stream);

/* ---- K3+K5: Regev encryption -------------------------------------------------------------- */
/* out_c8[k] = ct_export(regev_encrypt2(rng at offset + k*MFB_CTR_CT, sk, msg[k], e_k)) for k < count
 * (lwe.c:78-97, 115-119): b = (e*p + <sk, a> + m) mod 2^704.  e_k is the little-endian integer of
 * ent[k*ent_stride .. +ent_nbytes): pass the raw entropy of the reference's two getrandom calls per
 * encryption (stride 70, nbytes 69 — errdist_uniform lwe.c:60-63; the sign byte is drawn and unused,
 * lwe.c:86-87) or an explicit noise value for a custom chi (stride = nbytes <= 88). */
MFB_API int mfb_encrypt(mfb_ctx *ctx, const uint8_t seed[40], uint64_t offset, const uint64_t *sk_flat, const uint64_t *msg,
                const uint8_t *ent, int ent_stride, int ent_nbytes, size_t count, uint8_t *out_c8);
/* The same with the entropy DRAWN THROUGH A CALLBACK (the shape of the reference's own injection point, the `chi`
 * argument of regev_encrypt2, lwe.c:78): `draw(user, dst, nbytes)` is called from the calling thread for consecutive
 * pieces of the count*ent_stride entropy bytes, in order, exactly once each — so a getrandom-backed callback consumes
 * the OS entropy in the reference's order (per encryption: 69 noise bytes, then 1 sign byte) — while the device
 * encrypts the previous piece: setup()'s 2D+M draws (tens of ms of getrandom) hide behind the kernel. */
MFB_API int mfb_encrypt_cb(mfb_ctx *ctx, const uint8_t seed[40], uint64_t offset, const uint64_t *sk_flat, const uint64_t *msg,
                   mfb_entropy_fn draw, void *user, int ent_stride, int ent_nbytes, size_t count, uint8_t *out_c8);
/* The same with the records written straight to nsegs segments of the record index space (record k of segment g goes to
 * g.dst + (k - g.first) * 92): setup() fills crs->s, crs->as, crs->t, crs->v without an intermediate array.  In both
 * forms the records of a piece travel back on a second stream while the next piece is being encrypted. */
MFB_API int mfb_encrypt_cb_segs(mfb_ctx *ctx, const uint8_t seed[40], uint64_t offset, const uint64_t *sk_flat, const uint64_t *msg,
                        mfb_entropy_fn draw, void *user, int ent_stride, int ent_nbytes, size_t count, const mfb_c8_segment *segs,
                        int nsegs);
MFB_API int mfb_encrypt_dev(mfb_ctx *ctx, const uint8_t seed[40], uint64_t offset, const uint64_t *sk_planar_dev,
                    const uint64_t *msg_dev, const uint8_t *ent_dev, int ent_stride, int ent_nbytes, size_t count,
                    uint8_t *out_c8_dev, void *stream);

/* ---- K4: batched decryption ---------------------------------------------------------------- */
/* out_m[k] = regev_decrypt(sk, ct_k) (lwe.c:105-111) = (b - (<a, sk> mod 2^704)) floor-mod p, where the b
 * coordinate of ct_k is -(its stored magnitude) when b_neg[k] != 0 (ct_smudge can leave it negative,
 * lwe.c:65-76).  out_dot (nullable) receives <a, sk> mod 2^704 as 11 u64 (mpz_dotp, lwe.h:57-61). */
MFB_API int mfb_decrypt(mfb_ctx *ctx, const uint64_t *sk_flat, const uint64_t *cts_flat, const uint8_t *b_neg, size_t count,
                uint64_t *out_m, uint64_t *out_dot);
MFB_API int mfb_decrypt_dev(mfb_ctx *ctx, const uint64_t *sk_planar_dev, const uint64_t *cts_flat_dev,
                    const uint8_t *b_neg_dev, size_t count, uint64_t *out_m_dev, uint64_t *out_dot_dev, void *stream);

/* ---- F_p[x] steps of prover and setup (p = 2^32 - 5) ------------------------------------------- */
/* The prover's polynomial step (snark.c:138-169; FLINT's nmod_poly add / scalar_mul / pow / div in the reference):
 *   w = delta*t + sum_{witness bit i-1 set} v_i,   v = w + v_0,   h = (v^2 - 1) / t   (Euclidean quotient)
 * ssp is the dense wire blob of ssp.h:6-9 in host memory: polynomial k at ssp[k*D .. (k+1)*D) as u64 coefficients
 * (k = 0: t, k = i+1: v_i; reduced mod p on import as nmod_poly_import does).  witness bit i-1 selects v_i, i < M.
 * Outputs: D canonical coefficients each (h truncated to D, as eval_poly reads it).  Multiplication = 3-prime NTT
 * + CRT on the device, division = Newton inversion; results are canonical residues, identical to FLINT's. */
MFB_API int mfb_ssp_prover_polys(mfb_ctx *ctx, const uint64_t *ssp, size_t D, size_t M, const uint64_t *witness_limbs,
                         size_t nlimbs, uint64_t delta, uint64_t *w_out, uint64_t *v_out, uint64_t *h_out);
/* The same with the SSP blob kept on the device (u32 residues, (M+1)*D*4 bytes) and rev(t)^-1 cached in the handle:
 * after the first proof the division is one multiplication and nothing but the witness crosses PCIe. */
MFB_API int mfb_ssp_create(mfb_ctx *ctx, const uint64_t *ssp, size_t D, size_t M, mfb_ssp **out);
MFB_API void mfb_ssp_destroy(mfb_ctx *ctx, mfb_ssp *h);
MFB_API int mfb_ssp_prover_polys_resident(mfb_ctx *ctx, mfb_ssp *h, const uint64_t *witness_limbs, size_t nlimbs,
                                  uint64_t delta, uint64_t *w_out, uint64_t *v_out, uint64_t *h_out);
/* The same with the results left on the device: *wvh_dev = three consecutive arrays of D uint32 residues (w, v, h),
 * valid until the next polynomial / encrypt / decrypt call on this context. */
MFB_API int mfb_ssp_prover_polys_resident_dev(mfb_ctx *ctx, mfb_ssp *h, const uint64_t *witness_limbs, size_t nlimbs,
                                      uint64_t delta, const uint32_t **wvh_dev);
/* The same queued on `stream` WITHOUT waiting (no host round trip at all: the selection indices travel through pinned
 * staging, the quotient uses the cached transform of rev(t)^-1): *wvh_dev is valid in stream order. */
MFB_API int mfb_ssp_prover_polys_resident_async(mfb_ctx *ctx, mfb_ssp *h, const uint64_t *witness_limbs, size_t nlimbs,
                                        uint64_t delta, void *stream, const uint32_t **wvh_dev);
MFB_API size_t mfb_ssp_degree_bound(const mfb_ssp *h);
/* values[q] = poly_{first+q}(x) mod p over a resident blob [t, v_0, ..., v_{M-1}] (polynomial 0 = t): setup's and the
 * verifier's evaluations (snark.c:97-110, 197-201, 214-215) without shipping the coefficients again. */
MFB_API int mfb_ssp_eval_resident(mfb_ctx *ctx, const mfb_ssp *h, size_t first, size_t npoly, uint64_t x, uint64_t *values);
/* The prover's main pipeline with everything resident (snark.c:138-174) in ONE call, nothing but the witness bits and
 * the four accumulators crossing PCIe: polynomial step on the device, then
 *   (v_w, h) += sum (w_i, h_i) * CT_i over the s region,   (hat_v, hat_h) += sum (v_i, h_i) * CT_i over the as region
 * as two two-vector passes; regions of D = the SSP's degree bound ciphertexts.  Flat accumulators in and out. */
MFB_API int mfb_prove_resident(mfb_ctx *ctx, mfb_ssp *ssp, const mfb_region *reg_s, const mfb_region *reg_as,
                       const uint64_t *witness_limbs, size_t nlimbs, uint64_t delta, uint64_t *v_w_flat_inout,
                       uint64_t *h_flat_inout, uint64_t *hat_v_flat_inout, uint64_t *hat_h_flat_inout);
/* The same including b_w = delta * CT_t + sum_{witness bit i-1} CT_v[i-1] (snark.c:143-155; bt_recs = the t record followed
 * by the M-1 v records, region at stream offset bt_offset): the whole proof before smudging in one call, one
 * synchronisation.  b_w_flat_out is written (ct_import semantics), not accumulated; NULL skips it. */
MFB_API int mfb_prove_resident_bw(mfb_ctx *ctx, mfb_ssp *ssp, const mfb_region *reg_s, const mfb_region *reg_as,
                          const uint64_t *witness_limbs, size_t nlimbs, uint64_t delta, const uint8_t seed[40], uint64_t bt_offset,
                          const uint8_t *bt_recs, size_t M, uint64_t *v_w_flat_inout, uint64_t *h_flat_inout,
                          uint64_t *hat_v_flat_inout, uint64_t *hat_h_flat_inout, uint64_t *b_w_flat_out);
/* b_w alone, queued on `stream`: b_w_flat_out_dev (device) = delta * CT_t + sum_{witness bit i-1} CT_v[i-1]; bt_recs
 * are host records (t, then the M-1 v records). */
MFB_API int mfb_b_w_dev(mfb_ctx *ctx, const uint8_t seed[40], uint64_t bt_offset, const uint8_t *bt_recs, size_t M,
                const uint64_t *witness_limbs, size_t nlimbs, uint64_t delta, uint64_t *b_w_flat_out_dev, void *stream);
/* values[q] = poly_q(x) mod p for npoly polynomials of D u64 coefficients each (setup's nmod_poly_evaluate_nmod
 * calls, snark.c:97-110) */
MFB_API int mfb_ssp_eval(mfb_ctx *ctx, const uint64_t *polys, size_t D, size_t npoly, uint64_t x, uint64_t *values);

/* flat [count][n][11] -> row-planar [count][11][1472] on the device (secret keys: n = 1470) */
MFB_API int mfb_flat_to_planar_dev(mfb_ctx *ctx, const uint64_t *flat_dev, int n, size_t count, uint64_t *planar_dev,
                           void *stream);
/* flat ciphertexts [count][1471][11] -> the resident tile-planar layout that mfb_lincomb_dev streams */
MFB_API int mfb_flat_to_resident_dev(mfb_ctx *ctx, const uint64_t *flat_dev, size_t count, uint64_t *cts_dev,
                             void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MFB200_H */
