/* aes.h / entropy.h of the reference over the GPU keystream kernel.
 *
 * The stream is a pure function of (seed, byte position) — block k = AES256(key, nonce || LE64(k)), aes.c:122-133 —
 * so an rng is just a cursor.  The cursor keeps the reference's (ctr, rem) bookkeeping (aes.c:104-144,
 * entropy.c:46-56: position = 16*ctr - rem) so that CTR()/REM() read the same values, and a read-ahead window
 * so that the reference's habit of drawing 92 bytes at a time (entropy.h:62-66) costs one kernel launch per
 * window, not per draw.  No AES runs on the host.
 */
#include "mf_internal.h"

#define MF_WINDOW (1u << 20)          /* read-ahead window for small sequential reads */
#define MF_DIRECT (MF_WINDOW / 2)     /* reads at least this long bypass the window */

struct mf_aes_key {
  uint8_t seed[40]; /* nonce(8) || key(32) as mfb200.h wants it */
  uint8_t *win;
  uint64_t win_pos;
  size_t win_len;
};

void aesctr_init(aesctr_ptr stream, const uint8_t *key, const uint64_t nonce) {
  stream->rem = 0;
  stream->ctr = 0;
  stream->nonce = nonce;
  memset(stream->remb, 0, sizeof(stream->remb));
  stream->key = calloc(1, sizeof(struct mf_aes_key));
  if (!stream->key) {
    perror("Failed malloc");
    return;
  }
  memcpy(stream->key->seed, &nonce, 8);
  memcpy(stream->key->seed + 8, key, 32);
}

void aesctr_clear(aesctr_ptr stream) {
  if (stream->key) {
    free(stream->key->win);
    memset(stream->key, 0, sizeof(struct mf_aes_key));
    free(stream->key);
    stream->key = NULL;
  }
}

static inline uint64_t pos_of(aesctr_ptr s) { return s->ctr * 16 - s->rem; }
static inline void set_pos(aesctr_ptr s, uint64_t pos) {
  s->ctr = (pos + 15) / 16;
  s->rem = (size_t)(s->ctr * 16 - pos);
}

void aesctr_prg(aesctr_ptr stream, void *outbuf, size_t count) {
  struct mf_aes_key *k = stream->key;
  uint8_t *out = outbuf;
  uint64_t pos = pos_of(stream);
  set_pos(stream, pos + count);
  while (count) {
    if (k->win && pos >= k->win_pos && pos < k->win_pos + k->win_len) {
      size_t take = (size_t)(k->win_pos + k->win_len - pos);
      if (take > count) take = count;
      memcpy(out, k->win + (pos - k->win_pos), take);
      out += take;
      pos += take;
      count -= take;
      continue;
    }
    if (count >= MF_DIRECT) {
      MF_GPU(mfb_stream(mf_gpu(), k->seed, pos, out, count));
      return;
    }
    if (!k->win && !(k->win = malloc(MF_WINDOW))) mf_die("malloc");
    MF_GPU(mfb_stream(mf_gpu(), k->seed, pos, k->win, MF_WINDOW));
    k->win_pos = pos;
    k->win_len = MF_WINDOW;
  }
}

void rng_init(rng_t rs, uint8_t *rseed) {
  mf_gpu_prefetch();
  uint64_t nonce;
  memcpy(&nonce, rseed, 8);
  aesctr_init((aesctr_ptr)rs, rseed + 8, nonce);
}

void rng_clear(rng_t rng) { aesctr_clear((aesctr_ptr)rng); }

/* entropy.c:46-56: ctr = count/16, then count%16 bytes are sunk */
void rng_seek(rng_t prg, size_t count) { set_pos((aesctr_ptr)prg, count); }

uint64_t mf_rng_pos(rng_t rng) { return pos_of((aesctr_ptr)rng); }
void mf_rng_advance(rng_t rng, uint64_t nbytes) { set_pos((aesctr_ptr)rng, pos_of((aesctr_ptr)rng) + nbytes); }
const uint8_t *mf_rng_seed(rng_t rng) { return ((aesctr_ptr)rng)->key->seed; }

/* entropy.c:11-26 / 28-43: nbits/8 bytes -> little-endian limbs, top limb masked to nbits, normalised.
 * The destination limbs are cleared first, so bits between 8*(nbits/8) and nbits are 0 (the reference leaves
 * whatever the allocation held there). */
static void set_from_bytes(mpz_ptr rop, const uint8_t *bytes, size_t nbits) {
  const size_t limbs = BITS_TO_LIMBS(nbits), nbytes = nbits / 8;
  if (limbs == 0) {
    SIZ(rop) = 0;
    return;
  }
  mp_ptr rp = (size_t)ALLOC(rop) < limbs ? (mp_ptr)_mpz_realloc(rop, (mp_size_t)limbs) : PTR(rop);
  memset(rp, 0, limbs * sizeof(mp_limb_t));
  memcpy(rp, bytes, nbytes);
  rp[limbs - 1] &= (0xFFFFFFFFFFFFFFFFUL >> (limbs * 64 - nbits));
  size_t n = limbs;
  while (n > 0 && rp[n - 1] == 0) n--;
  SIZ(rop) = (int)n;
}

void mpz2_urandomb(mpz_ptr rop, rng_t prg, size_t nbits) {
  uint8_t buf[512];
  const size_t nbytes = nbits / 8;
  uint8_t *b = nbytes <= sizeof(buf) ? buf : malloc(nbytes);
  if (!b) mf_die("malloc");
  aesctr_prg((aesctr_ptr)prg, b, nbytes);
  set_from_bytes(rop, b, nbits);
  if (b != buf) free(b);
}

void mpz2_urandomb2(mpz_ptr rop, size_t nbits) {
  uint8_t buf[512];
  const size_t nbytes = nbits / 8;
  uint8_t *b = nbytes <= sizeof(buf) ? buf : malloc(nbytes);
  if (!b) mf_die("malloc");
  mf_entropy(b, nbytes);
  set_from_bytes(rop, b, nbits);
  if (b != buf) free(b);
}

/* entropy.h:56 — declared by the reference, defined and called nowhere in it */
void mpz_entropy_init(void) {}
