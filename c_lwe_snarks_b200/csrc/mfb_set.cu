// Device sets: several B200s driven by ONE host thread (the reference is a single-threaded C program, so this is the
// shape in which its prover() can use a whole NVSwitch box without becoming a multi-process job).
//
// A set is a primary context plus one context per further device, joined in a peer-memory exchange group
// (same-process peer access, mfb_peer_connect_local).  A set region is a CRS region sharded by ciphertext index
// (SURVEY.md §8e): member i holds a contiguous range resident in its HBM.  A lincomb over the region runs
// k_lincomb<2> on every member over its range (launches are asynchronous, the members work in parallel) and then
// one k_peer_allreduce per scalar vector; the result is read from the primary.  Built on the public C-ABI only.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/mfb200.h"

namespace {

struct Member {
  mfb_ctx *ctx = nullptr;
  bool owned = false;
  int device = 0;
  cudaStream_t stream = nullptr;
  mfb_peer_group *group = nullptr;
  // device scratch of the lincomb: coefficients, the member's flat partial sums, the reduced results
  uint32_t *co = nullptr;
  size_t co_cap = 0;
  uint64_t *part = nullptr;  // 4 x flat
  uint64_t *res = nullptr;   // 4 x flat
  uint8_t *c8 = nullptr;     // wire records of the member's range (mfb_set_eval_poly2)
  size_t c8_cap = 0;
  // mfb_set_encrypt_cb: the secret key (flat + row-planar), per-piece inputs, the member's records, pinned entropy
  uint64_t *skf = nullptr, *skp = nullptr;
  uint8_t *enc_in = nullptr, *enc_out = nullptr, *ent_pin = nullptr;
  size_t enc_in_cap = 0, enc_out_cap = 0, ent_pin_cap = 0;
  cudaEvent_t ent_free = nullptr;
  bool ent_used = false;
};

thread_local char g_set_err[256] = "";
int set_fail(int code, const char *msg) {
  snprintf(g_set_err, sizeof(g_set_err), "%s", msg);
  return code;
}

}  // namespace

struct mfb_set {
  std::vector<Member> m;
  uint64_t *acc_pin = nullptr;  // pinned staging of the four flat accumulators of mfb_set_prove_resident
};

struct mfb_set_region {
  std::vector<mfb_region *> shard;   // one per member (null when its range is empty)
  std::vector<size_t> first, count;  // ciphertext range of every member
  size_t total = 0;
};

#define SET_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != MFB_OK) return _r; \
  } while (0)
#define SET_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      snprintf(g_set_err, sizeof(g_set_err), "%s -> %s", #expr, cudaGetErrorString(_e));      \
      return MFB_ECUDA;                                                                       \
    }                                                                                         \
  } while (0)

extern "C" {

MFB_API const char *mfb_set_last_error(void) { return g_set_err[0] ? g_set_err : mfb_last_error(); }

MFB_API void mfb_set_destroy(mfb_set *s) {
  if (!s) return;
  for (auto &mb : s->m) {
    if (!mb.ctx) continue;
    cudaSetDevice(mb.device);
    cudaDeviceSynchronize();
  }
  for (auto &mb : s->m)
    if (mb.ctx && mb.group) mfb_peer_disconnect(mb.ctx, mb.group);
  for (auto &mb : s->m) {
    if (!mb.ctx) continue;
    cudaSetDevice(mb.device);
    if (mb.group) mfb_peer_destroy(mb.ctx, mb.group);
    if (mb.co) cudaFree(mb.co);
    if (mb.c8) cudaFree(mb.c8);
    if (mb.skf) cudaFree(mb.skf);
    if (mb.skp) cudaFree(mb.skp);
    if (mb.enc_in) cudaFree(mb.enc_in);
    if (mb.enc_out) cudaFree(mb.enc_out);
    if (mb.ent_pin) cudaFreeHost(mb.ent_pin);
    if (mb.ent_free) cudaEventDestroy(mb.ent_free);
    if (mb.part) cudaFree(mb.part);
    if (mb.res) cudaFree(mb.res);
    if (mb.stream) cudaStreamDestroy(mb.stream);
    if (mb.owned) mfb_ctx_destroy(mb.ctx);
  }
  if (!s->m.empty()) cudaSetDevice(s->m[0].device);
  if (s->acc_pin) cudaFreeHost(s->acc_pin);
  delete s;
}

// primary: an existing context (stays owned by the caller) = member 0; devices[0..ndev): one further member each
// (a device may repeat, and may be the primary's: the members then share that GPU — how the tests run on one GPU).
MFB_API int mfb_set_create(mfb_ctx *primary, const int *devices, int ndev, mfb_set **out) {
  g_set_err[0] = 0;
  if (!primary || !out || (ndev && !devices) || ndev < 0 || ndev + 1 > MFB_PEER_MAX)
    return set_fail(MFB_EARG, "mfb_set_create: bad argument (at most 16 members)");
  *out = nullptr;
  mfb_set *s = new (std::nothrow) mfb_set();
  if (!s) return set_fail(MFB_ENOMEM, "out of host memory");
  const int world = ndev + 1;
  s->m.resize(world);
  int rc = MFB_OK;
  for (int i = 0; i < world && rc == MFB_OK; i++) {
    Member &mb = s->m[i];
    mb.device = i == 0 ? mfb_ctx_device(primary) : devices[i - 1];
    if (i == 0) {
      mb.ctx = primary;
    } else {
      rc = mfb_ctx_create(&mb.ctx, mb.device);
      mb.owned = rc == MFB_OK;
    }
    if (rc != MFB_OK) break;
    cudaError_t e = cudaSetDevice(mb.device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&mb.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void **)&mb.part, 4 * MFB_PLANAR_U64 * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&mb.res, 4 * MFB_PLANAR_U64 * 8);
    if (e != cudaSuccess) {
      snprintf(g_set_err, sizeof(g_set_err), "mfb_set_create: device %d: %s", mb.device, cudaGetErrorString(e));
      rc = MFB_ECUDA;
      break;
    }
    uint8_t handle[MFB_PEER_HANDLE_BYTES];
    rc = mfb_peer_create(mb.ctx, world, i, &mb.group, handle);
  }
  if (rc == MFB_OK && world > 1) {
    void *bases[MFB_PEER_MAX];
    for (int i = 0; i < world; i++) bases[i] = mfb_peer_base(s->m[i].group);
    for (int i = 0; i < world && rc == MFB_OK; i++) rc = mfb_peer_connect_local(s->m[i].ctx, s->m[i].group, bases);
  }
  if (rc != MFB_OK) {
    char keep[256];
    snprintf(keep, sizeof(keep), "%s", mfb_set_last_error());
    mfb_set_destroy(s);
    snprintf(g_set_err, sizeof(g_set_err), "%s", keep);
    return rc;
  }
  cudaSetDevice(s->m[0].device);
  *out = s;
  return MFB_OK;
}

MFB_API int mfb_set_size(const mfb_set *s) { return s ? (int)s->m.size() : 0; }

MFB_API void mfb_set_region_destroy(mfb_set *s, mfb_set_region *r) {
  if (!r) return;
  if (s)
    for (size_t i = 0; i < r->shard.size() && i < s->m.size(); i++)
      if (r->shard[i]) mfb_region_destroy(s->m[i].ctx, r->shard[i]);
  if (s && !s->m.empty()) cudaSetDevice(s->m[0].device);
  delete r;
}

// The region of `count` ciphertexts at stream offset `offset` (records c8), sharded contiguously over the members.
MFB_API int mfb_set_region_create(mfb_set *s, const uint8_t seed[40], uint64_t offset, const uint8_t *c8, size_t count,
                                  mfb_set_region **out) {
  g_set_err[0] = 0;
  if (!s || !seed || !out || (count && !c8)) return set_fail(MFB_EARG, "mfb_set_region_create: null pointer");
  *out = nullptr;
  mfb_set_region *r = new (std::nothrow) mfb_set_region();
  if (!r) return set_fail(MFB_ENOMEM, "out of host memory");
  const size_t world = s->m.size();
  r->shard.assign(world, nullptr);
  r->first.assign(world, 0);
  r->count.assign(world, 0);
  r->total = count;
  const size_t base = count / world, extra = count % world;
  int rc = MFB_OK;
  for (size_t i = 0; i < world && rc == MFB_OK; i++) {
    r->first[i] = i * base + (i < extra ? i : extra);
    r->count[i] = base + (i < extra ? 1 : 0);
    // queued on the member's stream: the members' AES expansions run concurrently
    cudaSetDevice(s->m[i].device);
    rc = mfb_region_create_async(s->m[i].ctx, seed, offset + r->first[i] * (uint64_t)MFB_CTR_CT, c8 + r->first[i] * MFB_CT_BYTES,
                                 r->count[i], s->m[i].stream, &r->shard[i]);
  }
  for (size_t i = 0; i < world; i++) {
    cudaSetDevice(s->m[i].device);
    const cudaError_t e = cudaStreamSynchronize(s->m[i].stream);
    if (e != cudaSuccess && rc == MFB_OK) {
      snprintf(g_set_err, sizeof(g_set_err), "mfb_set_region_create: device %d: %s", s->m[i].device, cudaGetErrorString(e));
      rc = MFB_ECUDA;
    }
  }
  if (rc != MFB_OK) {
    char keep[256];
    snprintf(keep, sizeof(keep), "%s", mfb_set_last_error());
    mfb_set_region_destroy(s, r);
    snprintf(g_set_err, sizeof(g_set_err), "%s", keep);
    return rc;
  }
  cudaSetDevice(s->m[0].device);
  *out = r;
  return MFB_OK;
}

// rop0 += sum coeffs0[i] * CT_i, rop1 += sum coeffs1[i] * CT_i over the WHOLE region (d = its ciphertext count);
// coeffs1 / rop1 may both be NULL for a single scalar vector.  Host buffers in, host buffers out, synchronous.
MFB_API int mfb_set_region_lincomb2(mfb_set *s, const mfb_set_region *r, const uint32_t *coeffs0, const uint32_t *coeffs1,
                                    size_t d, uint64_t *rop0_flat_inout, uint64_t *rop1_flat_inout) {
  g_set_err[0] = 0;
  if (!s || !r || !rop0_flat_inout || (d && !coeffs0) || ((coeffs1 == nullptr) != (rop1_flat_inout == nullptr)))
    return set_fail(MFB_EARG, "mfb_set_region_lincomb2: null pointer");
  if (d != r->total) return set_fail(MFB_EARG, "mfb_set_region_lincomb2: d must be the region's ciphertext count");
  const size_t world = s->m.size();
  const int nvec = coeffs1 ? 2 : 1;
  const size_t FLAT = MFB_FLAT_CT_U64;
  // 1. every member: coefficients to the device, lincomb over its shard into its flat partial(s)
  for (size_t i = 0; i < world; i++) {
    Member &mb = s->m[i];
    SET_CUDA(cudaSetDevice(mb.device));
    const size_t cnt = r->count[i], first = r->first[i];
    if (mb.co_cap < 2 * cnt + 4) {
      if (mb.co) SET_CUDA(cudaFree(mb.co));
      mb.co = nullptr;
      mb.co_cap = 0;
      SET_CUDA(cudaMalloc((void **)&mb.co, (2 * cnt + 4) * 4));
      mb.co_cap = 2 * cnt + 4;
    }
    uint32_t *c0 = mb.co, *c1 = mb.co + cnt;
    if (cnt) {
      SET_CUDA(cudaMemcpyAsync(c0, coeffs0 + first, cnt * 4, cudaMemcpyHostToDevice, mb.stream));
      if (nvec == 2) SET_CUDA(cudaMemcpyAsync(c1, coeffs1 + first, cnt * 4, cudaMemcpyHostToDevice, mb.stream));
    }
    if (i == 0) {  // the incoming accumulators join the sum on the primary
      SET_CUDA(cudaMemcpyAsync(mb.res, rop0_flat_inout, FLAT * 8, cudaMemcpyHostToDevice, mb.stream));
      if (nvec == 2) SET_CUDA(cudaMemcpyAsync(mb.res + MFB_PLANAR_U64, rop1_flat_inout, FLAT * 8, cudaMemcpyHostToDevice, mb.stream));
    }
    const uint64_t *cts = (const uint64_t *)mfb_region_cts(r->shard[i]);
    if (nvec == 2)
      SET_TRY(mfb_lincomb2_dev(mb.ctx, cts, c0, c1, cnt, nullptr, mb.part, nullptr, mb.part + MFB_PLANAR_U64, mb.stream));
    else
      SET_TRY(mfb_lincomb_dev(mb.ctx, cts, c0, cnt, nullptr, mb.part, mb.stream));
  }
  // 2. every member: one all-reduce kernel per vector (push to all members, wait, add); only the primary's result
  //    (which also adds the incoming accumulator) is read back
  for (int v = 0; v < nvec; v++)
    for (size_t i = 0; i < world; i++) {
      Member &mb = s->m[i];
      SET_CUDA(cudaSetDevice(mb.device));
      uint64_t *res = mb.res + (size_t)v * MFB_PLANAR_U64;
      SET_TRY(mfb_peer_allreduce_dev(mb.ctx, mb.group, mb.part + (size_t)v * MFB_PLANAR_U64, i == 0 ? res : nullptr, res, mb.stream));
    }
  Member &p = s->m[0];
  SET_CUDA(cudaSetDevice(p.device));
  SET_CUDA(cudaMemcpyAsync(rop0_flat_inout, p.res, FLAT * 8, cudaMemcpyDeviceToHost, p.stream));
  if (nvec == 2) SET_CUDA(cudaMemcpyAsync(rop1_flat_inout, p.res + MFB_PLANAR_U64, FLAT * 8, cudaMemcpyDeviceToHost, p.stream));
  int rc = MFB_OK;
  for (size_t i = 0; i < world; i++) {
    Member &mb = s->m[i];
    SET_CUDA(cudaSetDevice(mb.device));
    SET_CUDA(cudaStreamSynchronize(mb.stream));
    const int st = mfb_peer_status(mb.ctx, mb.group);
    if (st != MFB_OK) rc = st;
  }
  SET_CUDA(cudaSetDevice(p.device));
  return rc;
}

// eval_poly (one or two scalar vectors) with NOTHING resident, sharded by ciphertext index: member i regenerates the
// a-vectors of its contiguous range from AES in-kernel (mfb_eval_poly2_dev at its stream offset), one peer all-reduce
// kernel per vector combines the partial sums.  Host buffers in and out; coefficients as uint64 (< 2^32).
MFB_API int mfb_set_eval_poly2(mfb_set *s, const uint8_t seed[40], uint64_t offset, const uint8_t *c8, const uint64_t *coeffs0,
                               const uint64_t *coeffs1, size_t d, uint64_t *rop0_flat_inout, uint64_t *rop1_flat_inout) {
  g_set_err[0] = 0;
  if (!s || !seed || !rop0_flat_inout || (d && (!c8 || !coeffs0)) || ((coeffs1 == nullptr) != (rop1_flat_inout == nullptr)))
    return set_fail(MFB_EARG, "mfb_set_eval_poly2: null pointer");
  const size_t world = s->m.size(), FLAT = MFB_FLAT_CT_U64;
  const int nvec = coeffs1 ? 2 : 1;
  std::vector<uint32_t> co32((size_t)nvec * d + 1);
  for (int v = 0; v < nvec; v++) {
    const uint64_t *src = v ? coeffs1 : coeffs0;
    for (size_t i = 0; i < d; i++) {
      if (src[i] >> 32) return set_fail(MFB_EARG, "mfb_set_eval_poly2: a coefficient does not fit 32 bits");
      co32[(size_t)v * d + i] = (uint32_t)src[i];
    }
  }
  const size_t base = d / world, extra = d % world;
  for (size_t i = 0; i < world; i++) {
    Member &mb = s->m[i];
    SET_CUDA(cudaSetDevice(mb.device));
    const size_t first = i * base + (i < extra ? i : extra), cnt = base + (i < extra ? 1 : 0);
    if (mb.co_cap < 2 * cnt + 4) {
      if (mb.co) SET_CUDA(cudaFree(mb.co));
      mb.co = nullptr;
      mb.co_cap = 0;
      SET_CUDA(cudaMalloc((void **)&mb.co, (2 * cnt + 4) * 4));
      mb.co_cap = 2 * cnt + 4;
    }
    if (mb.c8_cap < cnt * MFB_CT_BYTES + 16) {
      if (mb.c8) SET_CUDA(cudaFree(mb.c8));
      mb.c8 = nullptr;
      mb.c8_cap = 0;
      SET_CUDA(cudaMalloc((void **)&mb.c8, cnt * MFB_CT_BYTES + 16));
      mb.c8_cap = cnt * MFB_CT_BYTES + 16;
    }
    uint32_t *c0 = mb.co, *c1 = mb.co + cnt;
    if (cnt) {
      SET_CUDA(cudaMemcpyAsync(c0, co32.data() + first, cnt * 4, cudaMemcpyHostToDevice, mb.stream));
      if (nvec == 2) SET_CUDA(cudaMemcpyAsync(c1, co32.data() + d + first, cnt * 4, cudaMemcpyHostToDevice, mb.stream));
      SET_CUDA(cudaMemcpyAsync(mb.c8, c8 + first * MFB_CT_BYTES, cnt * MFB_CT_BYTES, cudaMemcpyHostToDevice, mb.stream));
    }
    if (i == 0) {
      SET_CUDA(cudaMemcpyAsync(mb.res, rop0_flat_inout, FLAT * 8, cudaMemcpyHostToDevice, mb.stream));
      if (nvec == 2) SET_CUDA(cudaMemcpyAsync(mb.res + MFB_PLANAR_U64, rop1_flat_inout, FLAT * 8, cudaMemcpyHostToDevice, mb.stream));
    }
    const uint64_t off = offset + first * (uint64_t)MFB_CTR_CT;
    if (nvec == 2)
      SET_TRY(mfb_eval_poly2_dev(mb.ctx, seed, off, mb.c8, c0, c1, cnt, nullptr, mb.part, nullptr, mb.part + MFB_PLANAR_U64, mb.stream));
    else
      SET_TRY(mfb_eval_poly_dev(mb.ctx, seed, off, mb.c8, c0, nullptr, cnt, nullptr, mb.part, mb.stream));
  }
  for (int v = 0; v < nvec; v++)
    for (size_t i = 0; i < world; i++) {
      Member &mb = s->m[i];
      SET_CUDA(cudaSetDevice(mb.device));
      uint64_t *res = mb.res + (size_t)v * MFB_PLANAR_U64;
      SET_TRY(mfb_peer_allreduce_dev(mb.ctx, mb.group, mb.part + (size_t)v * MFB_PLANAR_U64, i == 0 ? res : nullptr, res, mb.stream));
    }
  Member &p = s->m[0];
  SET_CUDA(cudaSetDevice(p.device));
  SET_CUDA(cudaMemcpyAsync(rop0_flat_inout, p.res, FLAT * 8, cudaMemcpyDeviceToHost, p.stream));
  if (nvec == 2) SET_CUDA(cudaMemcpyAsync(rop1_flat_inout, p.res + MFB_PLANAR_U64, FLAT * 8, cudaMemcpyDeviceToHost, p.stream));
  int rc = MFB_OK;
  for (size_t i = 0; i < world; i++) {
    Member &mb = s->m[i];
    SET_CUDA(cudaSetDevice(mb.device));
    SET_CUDA(cudaStreamSynchronize(mb.stream));  // (co32 is read by the queued copies until here)
    const int st = mfb_peer_status(mb.ctx, mb.group);
    if (st != MFB_OK) rc = st;
  }
  SET_CUDA(cudaSetDevice(p.device));
  return rc;
}

// mfb_encrypt_cb over a device set: the entropy is still drawn by the calling thread piece by piece and in order
// (a hooked, deterministic source sees the reference's sequence), but piece k is encrypted by member k mod size, so
// the members work on different pieces at the same time and the call becomes entropy-bound instead of AES-bound.
static int grow(void **p, size_t *cap, size_t need, bool pinned) {
  if (*cap >= need) return MFB_OK;
  if (*p) {
    if (pinned) cudaFreeHost(*p);
    else cudaFree(*p);
  }
  *p = nullptr;
  *cap = 0;
  const cudaError_t e = pinned ? cudaHostAlloc(p, need, cudaHostAllocDefault) : cudaMalloc(p, need);
  if (e != cudaSuccess) {
    cudaGetLastError();
    snprintf(g_set_err, sizeof(g_set_err), "mfb_set_encrypt_cb: allocation of %zu bytes failed: %s", need, cudaGetErrorString(e));
    return MFB_ENOMEM;
  }
  *cap = need;
  return MFB_OK;
}

MFB_API int mfb_set_encrypt_cb(mfb_set *s, const uint8_t seed[40], uint64_t offset, const uint64_t *sk_flat, const uint64_t *msg,
                               mfb_entropy_fn draw, void *user, int ent_stride, int ent_nbytes, size_t count, uint8_t *out_c8) {
  g_set_err[0] = 0;
  if (count == 0) return MFB_OK;
  if (!s || !seed || !sk_flat || !msg || !draw || !out_c8) return set_fail(MFB_EARG, "mfb_set_encrypt_cb: null pointer");
  if (ent_nbytes < 0 || ent_nbytes > 88 || ent_stride < ent_nbytes || ent_stride <= 0)
    return set_fail(MFB_EARG, "mfb_set_encrypt_cb: need 0 <= ent_nbytes <= 88 and ent_stride >= max(1, ent_nbytes)");
  const size_t world = s->m.size();
  const size_t piece = (size_t)mfb_device_sm_count(s->m[0].ctx) * 110, first_piece = (size_t)mfb_device_sm_count(s->m[0].ctx) * 16;
  const size_t in_per = (size_t)ent_stride + 8;  // per ciphertext in a member's input buffer: entropy, then the message
  const size_t in_slot = piece * in_per + 8;     // (+ 8: the messages start at the next multiple of 8 bytes)
  // pieces: [first_piece, piece, piece, ...]; piece k goes to member k % world, at slot k / world of its buffers
  std::vector<size_t> start, len;
  for (size_t done = 0; done < count;) {
    size_t cnt = start.empty() ? first_piece : piece;
    if (cnt > count - done) cnt = count - done;
    start.push_back(done);
    len.push_back(cnt);
    done += cnt;
  }
  const size_t npieces = start.size(), slots = (npieces + world - 1) / world;
  for (size_t i = 0; i < world && i < npieces; i++) {
    Member &mb = s->m[i];
    SET_CUDA(cudaSetDevice(mb.device));
    size_t cap;
    cap = mb.skf ? MFB_PLANAR_U64 * 8 : 0;
    SET_TRY(grow((void **)&mb.skf, &cap, MFB_PLANAR_U64 * 8, false));
    cap = mb.skp ? MFB_PLANAR_U64 * 8 : 0;
    SET_TRY(grow((void **)&mb.skp, &cap, MFB_PLANAR_U64 * 8, false));
    SET_TRY(grow((void **)&mb.enc_in, &mb.enc_in_cap, slots * in_slot, false));
    SET_TRY(grow((void **)&mb.enc_out, &mb.enc_out_cap, slots * piece * MFB_CT_BYTES, false));
    SET_TRY(grow((void **)&mb.ent_pin, &mb.ent_pin_cap, piece * (size_t)ent_stride, true));
    if (!mb.ent_free) SET_CUDA(cudaEventCreateWithFlags(&mb.ent_free, cudaEventDisableTiming));
    mb.ent_used = false;
    SET_CUDA(cudaMemcpyAsync(mb.skf, sk_flat, MFB_FLAT_SK_U64 * 8, cudaMemcpyHostToDevice, mb.stream));
    SET_TRY(mfb_flat_to_planar_dev(mb.ctx, mb.skf, MFB_N, 1, mb.skp, mb.stream));
  }
  for (size_t k = 0; k < npieces; k++) {
    Member &mb = s->m[k % world];
    SET_CUDA(cudaSetDevice(mb.device));
    const size_t slot = k / world, cnt = len[k];
    uint8_t *d_ent = mb.enc_in + slot * in_slot;
    uint64_t *d_msg = (uint64_t *)(d_ent + piece * (size_t)ent_stride + ((8 - (piece * (size_t)ent_stride) % 8) % 8));
    if (mb.ent_used) SET_CUDA(cudaEventSynchronize(mb.ent_free));  // the copy that last read this pinned buffer is done
    draw(user, mb.ent_pin, cnt * (size_t)ent_stride);
    SET_CUDA(cudaMemcpyAsync(d_ent, mb.ent_pin, cnt * (size_t)ent_stride, cudaMemcpyHostToDevice, mb.stream));
    SET_CUDA(cudaEventRecord(mb.ent_free, mb.stream));
    mb.ent_used = true;
    SET_CUDA(cudaMemcpyAsync(d_msg, msg + start[k], cnt * 8, cudaMemcpyHostToDevice, mb.stream));
    SET_TRY(mfb_encrypt_dev(mb.ctx, seed, offset + start[k] * (uint64_t)MFB_CTR_CT, mb.skp, d_msg, d_ent, ent_stride, ent_nbytes, cnt,
                            mb.enc_out + slot * piece * MFB_CT_BYTES, mb.stream));
  }
  int rc = MFB_OK;
  for (size_t k = 0; k < npieces; k++) {  // records back, piece by piece (each copy waits for its member's stream)
    Member &mb = s->m[k % world];
    SET_CUDA(cudaSetDevice(mb.device));
    SET_CUDA(cudaMemcpyAsync(out_c8 + start[k] * MFB_CT_BYTES, mb.enc_out + (k / world) * piece * MFB_CT_BYTES, len[k] * MFB_CT_BYTES,
                             cudaMemcpyDeviceToHost, mb.stream));
  }
  for (size_t i = 0; i < world && i < npieces; i++) {
    Member &mb = s->m[i];
    SET_CUDA(cudaSetDevice(mb.device));
    SET_CUDA(cudaStreamSynchronize(mb.stream));
    memset(mb.ent_pin, 0, mb.ent_pin_cap);  // the noise is secret
    SET_CUDA(cudaMemsetAsync(mb.enc_in, 0, mb.enc_in_cap, mb.stream));
  }
  SET_CUDA(cudaSetDevice(s->m[0].device));
  return rc;
}

// mfb_prove_resident over sharded regions (see mfb200.h)
MFB_API int mfb_set_prove_resident(mfb_set *s, mfb_ssp *ssp, const mfb_set_region *rs, const mfb_set_region *ras,
                                   const uint64_t *witness_limbs, size_t nlimbs, uint64_t delta, uint64_t *v_w_flat_inout,
                                   uint64_t *h_flat_inout, uint64_t *hat_v_flat_inout, uint64_t *hat_h_flat_inout) {
  g_set_err[0] = 0;
  if (!s || !ssp || !rs || !ras || !witness_limbs || !v_w_flat_inout || !h_flat_inout || !hat_v_flat_inout || !hat_h_flat_inout)
    return set_fail(MFB_EARG, "mfb_set_prove_resident: null pointer");
  const size_t D = mfb_ssp_degree_bound(ssp), world = s->m.size(), FLAT = MFB_FLAT_CT_U64;
  if (rs->total != D || ras->total != D) return set_fail(MFB_EARG, "mfb_set_prove_resident: the regions must hold D ciphertexts");
  for (size_t i = 0; i < world; i++)
    if (rs->first[i] != ras->first[i] || rs->count[i] != ras->count[i])
      return set_fail(MFB_EARG, "mfb_set_prove_resident: the two regions are sharded differently");
  Member &p = s->m[0];
  SET_CUDA(cudaSetDevice(p.device));
  const uint32_t *wvh = nullptr;
  SET_TRY(mfb_ssp_prover_polys_resident_dev(p.ctx, ssp, witness_limbs, nlimbs, delta, &wvh));  // stream idle on return
  uint64_t *host[4] = {v_w_flat_inout, h_flat_inout, hat_v_flat_inout, hat_h_flat_inout};
  // the four accumulators travel as ONE pinned copy each way (slots of MFB_PLANAR_U64, the stride of p.res)
  if (!s->acc_pin) SET_CUDA(cudaHostAlloc((void **)&s->acc_pin, 4 * MFB_PLANAR_U64 * 8, cudaHostAllocDefault));
  for (int k = 0; k < 4; k++) memcpy(s->acc_pin + (size_t)k * MFB_PLANAR_U64, host[k], FLAT * 8);
  SET_CUDA(cudaMemcpyAsync(p.res, s->acc_pin, 4 * MFB_PLANAR_U64 * 8, cudaMemcpyHostToDevice, p.stream));
  // every member: its slices of w, v, h over NVLink, then both two-vector passes over its shards
  for (size_t i = 0; i < world; i++) {
    Member &mb = s->m[i];
    SET_CUDA(cudaSetDevice(mb.device));
    const size_t cnt = rs->count[i], first = rs->first[i];
    if (mb.co_cap < 3 * cnt + 4) {
      if (mb.co) SET_CUDA(cudaFree(mb.co));
      mb.co = nullptr;
      mb.co_cap = 0;
      SET_CUDA(cudaMalloc((void **)&mb.co, (3 * cnt + 4) * 4));
      mb.co_cap = 3 * cnt + 4;
    }
    uint32_t *cw = mb.co, *cv = mb.co + cnt, *ch = mb.co + 2 * cnt;
    if (cnt) {
      SET_CUDA(cudaMemcpyPeerAsync(cw, mb.device, wvh + first, p.device, cnt * 4, mb.stream));
      SET_CUDA(cudaMemcpyPeerAsync(cv, mb.device, wvh + D + first, p.device, cnt * 4, mb.stream));
      SET_CUDA(cudaMemcpyPeerAsync(ch, mb.device, wvh + 2 * D + first, p.device, cnt * 4, mb.stream));
    }
    uint64_t *pt = mb.part;
    SET_TRY(mfb_lincomb2_dev(mb.ctx, (const uint64_t *)mfb_region_cts(rs->shard[i]), cw, ch, cnt, nullptr, pt, nullptr,
                             pt + MFB_PLANAR_U64, mb.stream));
    SET_TRY(mfb_lincomb2_dev(mb.ctx, (const uint64_t *)mfb_region_cts(ras->shard[i]), cv, ch, cnt, nullptr, pt + 2 * MFB_PLANAR_U64,
                             nullptr, pt + 3 * MFB_PLANAR_U64, mb.stream));
  }
  for (int v = 0; v < 4; v++)
    for (size_t i = 0; i < world; i++) {
      Member &mb = s->m[i];
      SET_CUDA(cudaSetDevice(mb.device));
      uint64_t *res = mb.res + (size_t)v * MFB_PLANAR_U64;
      SET_TRY(mfb_peer_allreduce_dev(mb.ctx, mb.group, mb.part + (size_t)v * MFB_PLANAR_U64, i == 0 ? res : nullptr, res, mb.stream));
    }
  SET_CUDA(cudaSetDevice(p.device));
  SET_CUDA(cudaMemcpyAsync(s->acc_pin, p.res, 4 * MFB_PLANAR_U64 * 8, cudaMemcpyDeviceToHost, p.stream));
  int rc = MFB_OK;
  for (size_t i = 0; i < world; i++) {
    Member &mb = s->m[i];
    SET_CUDA(cudaSetDevice(mb.device));
    SET_CUDA(cudaStreamSynchronize(mb.stream));
    const int st = mfb_peer_status(mb.ctx, mb.group);
    if (st != MFB_OK) rc = st;
  }
  SET_CUDA(cudaSetDevice(p.device));
  for (int k = 0; k < 4; k++) memcpy(host[k], s->acc_pin + (size_t)k * MFB_PLANAR_U64, FLAT * 8);
  return rc;
}

}  // extern "C"
