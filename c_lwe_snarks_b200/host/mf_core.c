/* Device context, instance size, entropy source and mpz <-> limb conversions for the drop-in layer. */
#include "mf_internal.h"

#include <pthread.h>
#include <sys/syscall.h>
#include <unistd.h>

/* ------------------------------------------------------------------ instance size (lwe.h:14-21 made run-time) */
static size_t g_d = 0, g_m = 0;       /* explicit (mf_set_instance) */
static size_t g_def_d = 0, g_def_m = 0; /* first compile-time default registered by a caller */

void mf_set_instance(size_t D, size_t M) {
  __atomic_store_n(&g_d, D, __ATOMIC_RELAXED);
  __atomic_store_n(&g_m, M, __ATOMIC_RELAXED);
}
/* (the warm-up thread peeks at these while the program's thread may be registering them: relaxed atomics) */
#define LD(x) __atomic_load_n(&(x), __ATOMIC_RELAXED)
#define ST(x, v) __atomic_store_n(&(x), (v), __ATOMIC_RELAXED)
size_t mf_gamma_d(size_t def) {
  if (LD(g_d)) return LD(g_d);
  if (!LD(g_def_d)) ST(g_def_d, def);
  return LD(g_def_d);
}
size_t mf_gamma_m(size_t def) {
  if (LD(g_m)) return LD(g_m);
  if (!LD(g_def_m)) ST(g_def_m, def);
  return LD(g_def_m);
}
/* the instance size if a caller has fixed it already (explicitly or by evaluating GAMMA_D / GAMMA_M), WITHOUT registering
 * a default: the library's own background thread must not decide the instance size for the program */
static int instance_peek(size_t *d, size_t *m) {
  *d = LD(g_d) ? LD(g_d) : LD(g_def_d);
  *m = LD(g_m) ? LD(g_m) : LD(g_def_m);
  return *d != 0 && *m != 0;
}

/* ------------------------------------------------------------------ entropy */
static mf_entropy_fn g_ent_fn = NULL;
static void *g_ent_arg = NULL;

void mf_set_entropy_source(mf_entropy_fn fn, void *arg) {
  g_ent_fn = fn;
  g_ent_arg = arg;
}

int mf_entropy_hooked(void) { return g_ent_fn != NULL; }

void mf_entropy(void *buf, size_t len) {
  if (g_ent_fn) {
    g_ent_fn(buf, len, g_ent_arg);
    return;
  }
  uint8_t *p = buf;
  while (len) { /* getrandom(2) hands out at most 32 MiB - 1 per call and may be interrupted */
    long r = syscall(SYS_getrandom, p, len, 0);
    if (r <= 0) mf_die("getrandom(2)");
    p += r;
    len -= (size_t)r;
  }
}

/* ------------------------------------------------------------------ device context */
static mfb_ctx *g_ctx = NULL;
static int g_device = -1;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
/* CUDA initialisation + the first pinned allocations take a few hundred ms.  The reference's programs call setup() as
 * their first GPU-backed function and time it (benchmark_snark.c:56-63), so the context is created IN THE BACKGROUND
 * as soon as the program touches the library at all (RNG_INIT / crs_init / random_ssp / key_gen ... all of which run
 * host-only code first): by the time a kernel is needed the context usually exists.  A failure there is silent —
 * mf_gpu() then retries in the foreground and reports. */
static pthread_t g_warm_thread;
static int g_warm_state = 0; /* 0 = not started, 1 = thread running / to be joined, 2 = joined */
static mfb_ctx *g_warm_ctx = NULL;

void mf_die(const char *what) {
  fprintf(stderr, "mangiafuoco_b200: %s failed: %s\n", what, mfb_last_error());
  abort();
}

void mf_set_device(int device) { g_device = device; }

static int wanted_device(void) {
  if (g_device >= 0) return g_device;
  const char *e = getenv("MF_B200_DEVICE");
  return e ? atoi(e) : 0;
}

static void *warm_main(void *arg) {
  (void)arg;
  mfb_ctx *c = NULL;
  if (mfb_ctx_create(&c, wanted_device()) == MFB_OK) {
    /* pinned staging buffers, events; then — if the program has fixed its instance size by now (it usually has: context
     * creation takes a few hundred ms) — the cold-start costs of THAT size: scratch buffers at their final sizes, the
     * second stream, the first launch of every kernel.  A one-shot program times its only setup() / prover() calls. */
    size_t d = 0, m = 0;
    mfb_ctx_warm(c);
    if (!getenv("MF_B200_NO_RESERVE") && instance_peek(&d, &m)) (void)mfb_ctx_reserve(c, d, m);
    g_warm_ctx = c;
  }
  return NULL;
}

void mf_gpu_prefetch(void) {
  if (g_ctx || g_warm_state) return;
  pthread_mutex_lock(&g_lock);
  if (!g_ctx && g_warm_state == 0 && !getenv("MF_B200_NO_PREFETCH")) {
    if (pthread_create(&g_warm_thread, NULL, warm_main, NULL) == 0) g_warm_state = 1;
  }
  pthread_mutex_unlock(&g_lock);
}

mfb_ctx *mf_gpu(void) {
  if (g_ctx) return g_ctx;
  pthread_mutex_lock(&g_lock);
  if (!g_ctx && g_warm_state == 1) {
    pthread_join(g_warm_thread, NULL);
    g_warm_state = 2;
    if (g_warm_ctx && mfb_ctx_device(g_warm_ctx) != wanted_device()) { /* mf_set_device() came after the prefetch */
      mfb_ctx_destroy(g_warm_ctx);
      g_warm_ctx = NULL;
    }
    g_ctx = g_warm_ctx;
  }
  if (!g_ctx) {
    mfb_ctx *c = NULL;
    if (mfb_ctx_create(&c, wanted_device()) != MFB_OK) mf_die("mfb_ctx_create (no CPU fallback exists)");
    g_ctx = c;
  }
  pthread_mutex_unlock(&g_lock);
  return g_ctx;
}

uint64_t mf_gpu_launches(void) { return g_ctx ? mfb_launch_count(g_ctx) : 0; }

/* ------------------------------------------------------------------ tracing (timeit.h:4-19 analogue) */
#include <time.h>
double mf_now(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}
void mf_trace(const char *name, double t0) {
  static int on = -1;
  if (on < 0) on = getenv("MF_B200_TRACE") != NULL;
  if (on) fprintf(stderr, "%s\t%.6f\n", name, mf_now() - t0);
  if (on && getenv("MF_B200_TRACE_WALL")) { /* wall-clock time of the trace point (to line it up with nvidia-smi samples) */
    struct timespec w;
    clock_gettime(CLOCK_REALTIME, &w);
    fprintf(stderr, "  at\t%.3f\n", (double)w.tv_sec + 1e-9 * (double)w.tv_nsec);
  }
}

/* ------------------------------------------------------------------ conversions */
int mf_to_flat(uint64_t out[MF_LIMBS], mpz_srcptr z) {
  int n = abs(SIZ(z));
  if (n > MF_LIMBS) n = MF_LIMBS; /* value mod 2^704 */
  memcpy(out, PTR(z), (size_t)n * 8);
  memset(out + n, 0, (size_t)(MF_LIMBS - n) * 8);
  return SIZ(z) < 0;
}

/* z <- the non-negative integer of 11 little-endian limbs: written straight into the limb array (what mpz_import
 * computes, without its per-call overhead — a proof is 5 x 1471 of these) */
void mf_from_flat(mpz_ptr z, const uint64_t in[MF_LIMBS]) {
  if (ALLOC(z) < MF_LIMBS) _mpz_realloc(z, MF_LIMBS);
  int n = MF_LIMBS;
  while (n > 0 && in[n - 1] == 0) n--;
  memcpy(PTR(z), in, (size_t)n * 8);
  SIZ(z) = n;
}

void mf_ct_to_flat(uint64_t *out, ct_t ct, const char *who) {
  for (size_t i = 0; i <= GAMMA_N; i++)
    if (mf_to_flat(out + i * MF_LIMBS, ct[i])) {
      fprintf(stderr, "mangiafuoco_b200: %s: coordinate %zu is negative (the reference asserts SIZ >= 0 here)\n", who, i);
      abort();
    }
}

void mf_ct_from_flat(ct_t ct, const uint64_t *in) {
  for (size_t i = 0; i <= GAMMA_N; i++) mf_from_flat(ct[i], in + i * MF_LIMBS);
}

void mf_bytes_to_mpz(mpz_ptr z, const uint8_t *bytes, size_t n) { mpz_import(z, n, -1, 1, -1, 0, bytes); }

/* lwe.h:108-118: SIZ > 11 -> keep limbs 0..10 (limb 11 is masked and then dropped), normalise */
void modq(mpz_t a) {
  const int pos = GAMMA_LOGQ / 64;
  if (SIZ(a) > pos) {
    int n = pos;
    while (n > 0 && PTR(a)[n - 1] == 0) n--;
    SIZ(a) = n;
  }
}
