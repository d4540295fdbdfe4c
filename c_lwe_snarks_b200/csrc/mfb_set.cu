// Device sets: several B200s driven by ONE host thread (the reference is a single-threaded C program, so this is the
// shape in which its prover() can use a whole NVSwitch box without becoming a multi-process job).
//
// A set is a primary context plus one context per further device, joined in a peer-memory exchange group
// (same-process peer access, mfb_peer_connect_local).  A set region is a CRS region sharded by ciphertext index
// (SURVEY.md §8e): member i holds a contiguous range resident in its HBM.  A lincomb over the region runs
// k_lincomb<2> on every member over its range (launches are asynchronous, the members work in parallel) and then
// ONE exchange kernel per member combines all the accumulators of the call (lanes); the result is read from the
// primary.  Built on the public C-ABI only.
//
// Failure handling: any error inside a set call leaves the members' exchange sequence numbers possibly out of step, so
// the set is marked POISONED (every later call fails with MFB_EPEER until it is destroyed and created again); results
// travel through pinned staging and reach the caller's buffers only when every member finished cleanly.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/mfb200.h"

namespace {

constexpr size_t SLOT = MFB_PLANAR_U64;  // stride (u64) between the flat accumulators of a member's part / res arrays
constexpr int NACC = 5;                  // v_w, h, hat_v, hat_h, b_w

struct Member {
  mfb_ctx *ctx = nullptr;
  bool owned = false;
  int device = 0;
  cudaStream_t stream = nullptr;
  mfb_peer_group *group = nullptr;
  // device scratch of the lincomb: coefficients, the member's flat partial sums, the reduced results
  uint32_t *co = nullptr;
  size_t co_cap = 0;
  uint64_t *part = nullptr;  // NACC x flat
  uint64_t *res = nullptr;   // NACC x flat
  uint8_t *c8 = nullptr;     // wire records of the member's range (mfb_set_eval_poly2)
  size_t c8_cap = 0;
  // encryption: the secret key (flat + row-planar), entropy / messages / records of the member's range, pinned staging
  uint64_t *skf = nullptr, *skp = nullptr;
  uint8_t *enc_ent = nullptr, *enc_out = nullptr;
  uint64_t *enc_msg = nullptr;
  size_t enc_ent_cap = 0, enc_out_cap = 0, enc_msg_cap = 0;
  uint8_t *ent_pin[2] = {nullptr, nullptr};
  size_t ent_pin_cap = 0;
  cudaEvent_t ent_free[2] = {nullptr, nullptr};
};

thread_local char g_set_err[256] = "";
int set_fail(int code, const char *msg) {
  snprintf(g_set_err, sizeof(g_set_err), "%s", msg);
  return code;
}

}  // namespace

struct mfb_set {
  std::vector<Member> m;
  uint64_t *acc_pin = nullptr;  // pinned staging of the flat accumulators (NACC slots of SLOT u64)
  cudaEvent_t ev_polys = nullptr;
  bool poisoned = false;
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
      cudaGetLastError();                                                                     \
      return MFB_ECUDA;                                                                       \
    }                                                                                         \
  } while (0)

namespace {

int check_usable(mfb_set *s, const char *who) {
  if (!s) return set_fail(MFB_EARG, "null set");
  if (s->poisoned) {
    snprintf(g_set_err, sizeof(g_set_err), "%s: an earlier call on this device set failed; destroy it and create a new one", who);
    return MFB_EPEER;
  }
  return MFB_OK;
}

// Ends a set call: waits for every member, collects the peer status words.  rc != MFB_OK (the body failed) or a member
// that did not finish cleanly poisons the set; the error text of the first failure is kept.
int finish_call(mfb_set *s, int rc) {
  char keep[256];
  snprintf(keep, sizeof(keep), "%s", rc != MFB_OK ? (g_set_err[0] ? g_set_err : mfb_last_error()) : "");
  for (auto &mb : s->m) {
    if (!mb.ctx) continue;
    if (cudaSetDevice(mb.device) != cudaSuccess || cudaStreamSynchronize(mb.stream) != cudaSuccess) {
      if (rc == MFB_OK) {
        snprintf(keep, sizeof(keep), "device %d: %s", mb.device, cudaGetErrorString(cudaGetLastError()));
        rc = MFB_ECUDA;
      }
      cudaGetLastError();
      continue;
    }
    if (mb.group) {
      const int st = mfb_peer_status(mb.ctx, mb.group);
      if (st != MFB_OK && rc == MFB_OK) {
        snprintf(keep, sizeof(keep), "%s", mfb_last_error());
        rc = st;
      }
    }
  }
  cudaSetDevice(s->m[0].device);
  if (rc != MFB_OK) {
    s->poisoned = true;
    snprintf(g_set_err, sizeof(g_set_err), "%s", keep);
  }
  return rc;
}

int grow(void **p, size_t *cap, size_t need, bool pinned) {
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
    snprintf(g_set_err, sizeof(g_set_err), "device set: allocation of %zu bytes failed: %s", need, cudaGetErrorString(e));
    return MFB_ENOMEM;
  }
  *cap = need;
  return MFB_OK;
}

int ensure_acc_pin(mfb_set *s) {
  if (!s->acc_pin) SET_CUDA(cudaHostAlloc((void **)&s->acc_pin, NACC * SLOT * 8, cudaHostAllocDefault));
  return MFB_OK;
}

// flat accumulators of the caller -> pinned staging slots; returns whether any limb is non-zero
bool stage_in(mfb_set *s, uint64_t *const *host, int n) {
  uint64_t any = 0;
  for (int k = 0; k < n; k++)
    for (size_t i = 0; i < MFB_FLAT_CT_U64; i++) any |= (s->acc_pin[(size_t)k * SLOT + i] = host[k][i]);
  return any != 0;
}
void stage_out(mfb_set *s, uint64_t *const *host, int n) {
  for (int k = 0; k < n; k++) memcpy(host[k], s->acc_pin + (size_t)k * SLOT, MFB_FLAT_CT_U64 * 8);
}

// contiguous, balanced split of [0, total) over the members (the remainder goes to the low ranks)
void split_range(size_t total, size_t world, size_t i, size_t *first, size_t *count) {
  const size_t base = total / world, extra = total % world;
  *first = i * base + (i < extra ? i : extra);
  *count = base + (i < extra ? 1 : 0);
}

void scrub_member_secrets(Member &mb) {
  cudaSetDevice(mb.device);
  if (mb.skf) cudaMemsetAsync(mb.skf, 0, MFB_PLANAR_U64 * 8, mb.stream);
  if (mb.skp) cudaMemsetAsync(mb.skp, 0, MFB_PLANAR_U64 * 8, mb.stream);
  if (mb.enc_ent) cudaMemsetAsync(mb.enc_ent, 0, mb.enc_ent_cap, mb.stream);
  cudaStreamSynchronize(mb.stream);
  for (int k = 0; k < 2; k++)
    if (mb.ent_pin[k]) memset(mb.ent_pin[k], 0, mb.ent_pin_cap);
  cudaGetLastError();
}

}  // namespace

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
    if (mb.stream) scrub_member_secrets(mb);  // keys and noise do not outlive the set in device or pinned memory
    cudaSetDevice(mb.device);
    if (mb.group) mfb_peer_destroy(mb.ctx, mb.group);
    if (mb.co) cudaFree(mb.co);
    if (mb.c8) cudaFree(mb.c8);
    if (mb.skf) cudaFree(mb.skf);
    if (mb.skp) cudaFree(mb.skp);
    if (mb.enc_ent) cudaFree(mb.enc_ent);
    if (mb.enc_msg) cudaFree(mb.enc_msg);
    if (mb.enc_out) cudaFree(mb.enc_out);
    for (int k = 0; k < 2; k++) {
      if (mb.ent_pin[k]) cudaFreeHost(mb.ent_pin[k]);
      if (mb.ent_free[k]) cudaEventDestroy(mb.ent_free[k]);
    }
    if (mb.part) cudaFree(mb.part);
    if (mb.res) cudaFree(mb.res);
    if (mb.stream) cudaStreamDestroy(mb.stream);
    if (mb.owned) mfb_ctx_destroy(mb.ctx);
  }
  if (!s->m.empty()) cudaSetDevice(s->m[0].device);
  if (s->acc_pin) cudaFreeHost(s->acc_pin);
  if (s->ev_polys) cudaEventDestroy(s->ev_polys);
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
    if (e == cudaSuccess) e = cudaMalloc((void **)&mb.part, NACC * SLOT * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&mb.res, NACC * SLOT * 8);
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
  if (rc == MFB_OK) {
    cudaSetDevice(s->m[0].device);
    if (cudaEventCreateWithFlags(&s->ev_polys, cudaEventDisableTiming) != cudaSuccess) {
      snprintf(g_set_err, sizeof(g_set_err), "mfb_set_create: cudaEventCreate failed");
      rc = MFB_ECUDA;
    }
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
  SET_TRY(check_usable(s, "mfb_set_region_create"));
  if (!seed || !out || (count && !c8)) return set_fail(MFB_EARG, "mfb_set_region_create: null pointer");
  *out = nullptr;
  mfb_set_region *r = new (std::nothrow) mfb_set_region();
  if (!r) return set_fail(MFB_ENOMEM, "out of host memory");
  const size_t world = s->m.size();
  r->shard.assign(world, nullptr);
  r->first.assign(world, 0);
  r->count.assign(world, 0);
  r->total = count;
  // one host thread per member: the device allocation (synchronous, milliseconds for tens of GB) and the staging of the
  // member's wire records from pageable memory (which blocks the calling thread) run side by side, like the expansions
  for (size_t i = 0; i < world; i++) split_range(count, world, i, &r->first[i], &r->count[i]);
  std::vector<int> rcs(world, MFB_OK);
  std::vector<std::string> errs(world);
  auto work = [&](size_t i) {
    g_set_err[0] = 0;
    Member &mb = s->m[i];
    int rc1 = MFB_OK;
    if (cudaSetDevice(mb.device) != cudaSuccess) {
      rc1 = set_fail(MFB_ECUDA, "mfb_set_region_create: cudaSetDevice failed");
      cudaGetLastError();
    }
    if (rc1 == MFB_OK)
      rc1 = mfb_region_create_async(mb.ctx, seed, offset + r->first[i] * (uint64_t)MFB_CTR_CT, c8 + r->first[i] * MFB_CT_BYTES,
                                    r->count[i], mb.stream, &r->shard[i]);
    if (rc1 == MFB_OK) {
      const cudaError_t e = cudaStreamSynchronize(mb.stream);
      if (e != cudaSuccess) {
        snprintf(g_set_err, sizeof(g_set_err), "mfb_set_region_create: device %d: %s", mb.device, cudaGetErrorString(e));
        cudaGetLastError();
        rc1 = MFB_ECUDA;
      }
    }
    if (rc1 != MFB_OK) errs[i] = mfb_set_last_error();  // (thread-local: copy it out before the thread ends)
    rcs[i] = rc1;
  };
  {
    std::vector<std::thread> pool;
    for (size_t i = 1; i < world; i++) pool.emplace_back(work, i);
    work(0);
    for (auto &t : pool) t.join();
  }
  int rc = MFB_OK;
  for (size_t i = 0; i < world && rc == MFB_OK; i++)
    if (rcs[i] != MFB_OK) {
      snprintf(g_set_err, sizeof(g_set_err), "%s", errs[i].c_str());
      rc = rcs[i];
    }
  if (rc != MFB_OK) {  // (no exchange was started: the set itself stays usable)
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

// Second phase of a set call: the members' exchange.  Every member first PUSHES its flat partial sums to all members (a
// kernel that never waits), then a small kernel per member waits for the others' tiles and adds them.  No kernel of a
// set call ever waits for work that is not yet in some stream: members may share a GPU, and a profiler that serialises
// launches (ncu) can capture the path.
static int exchange_all(mfb_set *s, int nvec, bool any) {
  for (size_t i = 0; i < s->m.size(); i++) {
    Member &mb = s->m[i];
    SET_CUDA(cudaSetDevice(mb.device));
    SET_TRY(mfb_peer_push_lanes_dev(mb.ctx, mb.group, mb.part, SLOT, nvec, mb.stream));
  }
  for (size_t i = 0; i < s->m.size(); i++) {
    Member &mb = s->m[i];
    SET_CUDA(cudaSetDevice(mb.device));
    SET_TRY(mfb_peer_wait_lanes_dev(mb.ctx, mb.group, nvec, i == 0 && any ? mb.res : nullptr, mb.res, SLOT, mb.stream));
  }
  return MFB_OK;
}

static int region_lincomb2_body(mfb_set *s, const mfb_set_region *r, const uint32_t *coeffs0, const uint32_t *coeffs1, int nvec,
                                bool any) {
  const size_t world = s->m.size();
  // every member: coefficients to the device, lincomb over its shard into its flat partial(s), then ONE exchange kernel
  // for all the vectors of the call; only the primary's result (which also adds the incoming accumulators) is read back
  for (size_t i = 0; i < world; i++) {
    Member &mb = s->m[i];
    SET_CUDA(cudaSetDevice(mb.device));
    const size_t cnt = r->count[i], first = r->first[i];
    {
      size_t cap = mb.co_cap * 4;
      SET_TRY(grow((void **)&mb.co, &cap, (3 * cnt + 4) * 4, false));
      mb.co_cap = cap / 4;
    }
    uint32_t *c0 = mb.co, *c1 = mb.co + cnt;
    if (cnt) {
      SET_CUDA(cudaMemcpyAsync(c0, coeffs0 + first, cnt * 4, cudaMemcpyHostToDevice, mb.stream));
      if (nvec == 2) SET_CUDA(cudaMemcpyAsync(c1, coeffs1 + first, cnt * 4, cudaMemcpyHostToDevice, mb.stream));
    }
    if (i == 0 && any) SET_CUDA(cudaMemcpyAsync(mb.res, s->acc_pin, (size_t)nvec * SLOT * 8, cudaMemcpyHostToDevice, mb.stream));
    const uint64_t *cts = (const uint64_t *)mfb_region_cts(r->shard[i]);
    if (nvec == 2)
      SET_TRY(mfb_lincomb2_dev(mb.ctx, cts, c0, c1, cnt, nullptr, mb.part, nullptr, mb.part + SLOT, mb.stream));
    else
      SET_TRY(mfb_lincomb_dev(mb.ctx, cts, c0, cnt, nullptr, mb.part, mb.stream));
  }
  SET_TRY(exchange_all(s, nvec, any));
  Member &p = s->m[0];
  SET_CUDA(cudaSetDevice(p.device));
  SET_CUDA(cudaMemcpyAsync(s->acc_pin, p.res, (size_t)nvec * SLOT * 8, cudaMemcpyDeviceToHost, p.stream));
  return MFB_OK;
}

// rop0 += sum coeffs0[i] * CT_i, rop1 += sum coeffs1[i] * CT_i over the WHOLE region (d = its ciphertext count);
// coeffs1 / rop1 may both be NULL for a single scalar vector.  Host buffers in, host buffers out, synchronous.
MFB_API int mfb_set_region_lincomb2(mfb_set *s, const mfb_set_region *r, const uint32_t *coeffs0, const uint32_t *coeffs1,
                                    size_t d, uint64_t *rop0_flat_inout, uint64_t *rop1_flat_inout) {
  g_set_err[0] = 0;
  SET_TRY(check_usable(s, "mfb_set_region_lincomb2"));
  if (!r || !rop0_flat_inout || (d && !coeffs0) || ((coeffs1 == nullptr) != (rop1_flat_inout == nullptr)))
    return set_fail(MFB_EARG, "mfb_set_region_lincomb2: null pointer");
  if (d != r->total) return set_fail(MFB_EARG, "mfb_set_region_lincomb2: d must be the region's ciphertext count");
  const int nvec = coeffs1 ? 2 : 1;
  SET_TRY(ensure_acc_pin(s));
  uint64_t *host[2] = {rop0_flat_inout, rop1_flat_inout};
  const bool any = stage_in(s, host, nvec);
  const int rc = finish_call(s, region_lincomb2_body(s, r, coeffs0, coeffs1, nvec, any));
  if (rc == MFB_OK) stage_out(s, host, nvec);
  return rc;
}

static int eval_poly2_body(mfb_set *s, const uint8_t seed[40], uint64_t offset, const uint8_t *c8, const uint32_t *co32, size_t d,
                           int nvec, bool any) {
  const size_t world = s->m.size();
  // 1. every member: its scalars, then the AES + MAC kernel over its range — it needs nothing else
  for (size_t i = 0; i < world; i++) {
    Member &mb = s->m[i];
    SET_CUDA(cudaSetDevice(mb.device));
    size_t first, cnt;
    split_range(d, world, i, &first, &cnt);
    {
      size_t cap = mb.co_cap * 4;
      SET_TRY(grow((void **)&mb.co, &cap, (3 * cnt + 4) * 4, false));
      mb.co_cap = cap / 4;
    }
    SET_TRY(grow((void **)&mb.c8, &mb.c8_cap, cnt * MFB_CT_BYTES + 16, false));
    uint32_t *c0 = mb.co, *c1 = nvec == 2 ? mb.co + cnt : nullptr;
    if (cnt) {
      SET_CUDA(cudaMemcpyAsync(c0, co32 + first, cnt * 4, cudaMemcpyHostToDevice, mb.stream));
      if (nvec == 2) SET_CUDA(cudaMemcpyAsync(c1, co32 + d + first, cnt * 4, cudaMemcpyHostToDevice, mb.stream));
    }
    if (i == 0 && any) SET_CUDA(cudaMemcpyAsync(mb.res, s->acc_pin, (size_t)nvec * SLOT * 8, cudaMemcpyHostToDevice, mb.stream));
    SET_TRY(mfb_eval_poly2_begin_dev(mb.ctx, seed, offset + first * (uint64_t)MFB_CTR_CT, c0, c1, cnt, 1, mb.stream));
  }
  // 2. the wire records (pageable host memory: staging them blocks this thread) follow on every member's second stream
  //    while all the AES kernels are already running; then the b coordinate and the finish
  for (size_t i = 0; i < world; i++) {
    Member &mb = s->m[i];
    SET_CUDA(cudaSetDevice(mb.device));
    size_t first, cnt;
    split_range(d, world, i, &first, &cnt);
    uint32_t *c0 = mb.co, *c1 = nvec == 2 ? mb.co + cnt : nullptr;
    SET_TRY(mfb_eval_poly2_end_dev(mb.ctx, mb.c8, c8 + first * MFB_CT_BYTES, c0, c1, cnt, nullptr, mb.part, nullptr,
                                   nvec == 2 ? mb.part + SLOT : nullptr, mb.stream));
  }
  SET_TRY(exchange_all(s, nvec, any));
  Member &p = s->m[0];
  SET_CUDA(cudaSetDevice(p.device));
  SET_CUDA(cudaMemcpyAsync(s->acc_pin, p.res, (size_t)nvec * SLOT * 8, cudaMemcpyDeviceToHost, p.stream));
  return MFB_OK;
}

// eval_poly (one or two scalar vectors) with NOTHING resident, sharded by ciphertext index: member i regenerates the
// a-vectors of its contiguous range from AES in-kernel (mfb_eval_poly2_dev at its stream offset), one exchange kernel
// per member combines the partial sums.  Host buffers in and out; coefficients as uint64 (< 2^32).
MFB_API int mfb_set_eval_poly2(mfb_set *s, const uint8_t seed[40], uint64_t offset, const uint8_t *c8, const uint64_t *coeffs0,
                               const uint64_t *coeffs1, size_t d, uint64_t *rop0_flat_inout, uint64_t *rop1_flat_inout) {
  g_set_err[0] = 0;
  SET_TRY(check_usable(s, "mfb_set_eval_poly2"));
  if (!seed || !rop0_flat_inout || (d && (!c8 || !coeffs0)) || ((coeffs1 == nullptr) != (rop1_flat_inout == nullptr)))
    return set_fail(MFB_EARG, "mfb_set_eval_poly2: null pointer");
  const int nvec = coeffs1 ? 2 : 1;
  std::vector<uint32_t> co32((size_t)nvec * d + 1);
  for (int v = 0; v < nvec; v++) {
    const uint64_t *src = v ? coeffs1 : coeffs0;
    for (size_t i = 0; i < d; i++) {
      if (src[i] >> 32) return set_fail(MFB_EARG, "mfb_set_eval_poly2: a coefficient does not fit 32 bits");
      co32[(size_t)v * d + i] = (uint32_t)src[i];
    }
  }
  SET_TRY(ensure_acc_pin(s));
  uint64_t *host[2] = {rop0_flat_inout, rop1_flat_inout};
  const bool any = stage_in(s, host, nvec);
  // (finish_call waits for every member: co32 is read by the queued copies until then)
  const int rc = finish_call(s, eval_poly2_body(s, seed, offset, c8, co32.data(), d, nvec, any));
  if (rc == MFB_OK) stage_out(s, host, nvec);
  return rc;
}

// ---------------------------------------------------------------------------------------------------- encryption
// One member's share of an encryption call: ciphertexts [first, first + cnt) of the call, cut into pieces; the entropy of
// piece k+1 is drawn into pinned memory while the device encrypts piece k.  Everything is queued on the member's stream;
// the records stay on the device (mb.enc_out) until collect_records().
static int member_encrypt(Member &mb, const uint8_t seed[40], uint64_t offset, const uint64_t *sk_flat, const uint64_t *msg,
                          mfb_entropy_fn draw, void *user, int ent_stride, int ent_nbytes, size_t first, size_t cnt, size_t piece,
                          size_t first_piece) {
  SET_CUDA(cudaSetDevice(mb.device));
  if (cnt == 0) return MFB_OK;
  size_t cap;
  cap = mb.skf ? MFB_PLANAR_U64 * 8 : 0;
  SET_TRY(grow((void **)&mb.skf, &cap, MFB_PLANAR_U64 * 8, false));
  cap = mb.skp ? MFB_PLANAR_U64 * 8 : 0;
  SET_TRY(grow((void **)&mb.skp, &cap, MFB_PLANAR_U64 * 8, false));
  SET_TRY(grow((void **)&mb.enc_ent, &mb.enc_ent_cap, cnt * (size_t)ent_stride, false));
  cap = mb.enc_msg_cap;
  SET_TRY(grow((void **)&mb.enc_msg, &cap, cnt * 8, false));
  mb.enc_msg_cap = cap;
  SET_TRY(grow((void **)&mb.enc_out, &mb.enc_out_cap, cnt * MFB_CT_BYTES, false));
  const size_t pin_need = piece * (size_t)ent_stride;
  if (mb.ent_pin_cap < pin_need) {
    for (int k = 0; k < 2; k++) {
      size_t c2 = mb.ent_pin_cap;
      SET_TRY(grow((void **)&mb.ent_pin[k], &c2, pin_need, true));
    }
    mb.ent_pin_cap = pin_need;
  }
  for (int k = 0; k < 2; k++)
    if (!mb.ent_free[k]) SET_CUDA(cudaEventCreateWithFlags(&mb.ent_free[k], cudaEventDisableTiming));
  SET_CUDA(cudaMemcpyAsync(mb.skf, sk_flat, MFB_FLAT_SK_U64 * 8, cudaMemcpyHostToDevice, mb.stream));
  SET_TRY(mfb_flat_to_planar_dev(mb.ctx, mb.skf, MFB_N, 1, mb.skp, mb.stream));
  SET_CUDA(cudaMemcpyAsync(mb.enc_msg, msg + first, cnt * 8, cudaMemcpyHostToDevice, mb.stream));  // all messages up front
  bool used[2] = {false, false};
  int k = 0;
  for (size_t done = 0; done < cnt; k ^= 1) {
    size_t n = done == 0 ? first_piece : piece;
    if (n > cnt - done) n = cnt - done;
    if (used[k]) SET_CUDA(cudaEventSynchronize(mb.ent_free[k]));  // the copy that last read this pinned buffer is done
    draw(user, mb.ent_pin[k], n * (size_t)ent_stride);
    uint8_t *d_e = mb.enc_ent + done * (size_t)ent_stride;
    SET_CUDA(cudaMemcpyAsync(d_e, mb.ent_pin[k], n * (size_t)ent_stride, cudaMemcpyHostToDevice, mb.stream));
    SET_CUDA(cudaEventRecord(mb.ent_free[k], mb.stream));
    used[k] = true;
    SET_TRY(mfb_encrypt_dev(mb.ctx, seed, offset + (first + done) * (uint64_t)MFB_CTR_CT, mb.skp, mb.enc_msg + done, d_e, ent_stride,
                            ent_nbytes, n, mb.enc_out + done * MFB_CT_BYTES, mb.stream));
    done += n;
  }
  return MFB_OK;
}

// records [first, first + cnt) of the call, from the member's device buffer to their destinations: out_c8 (contiguous)
// or the caller's segments (record k of segment g goes to g.dst + (k - g.first) * 92)
static int collect_records(Member &mb, size_t first, size_t cnt, uint8_t *out_c8, const mfb_c8_segment *segs, int nsegs) {
  SET_CUDA(cudaSetDevice(mb.device));
  if (cnt == 0) return MFB_OK;
  if (out_c8) {
    SET_CUDA(cudaMemcpyAsync(out_c8 + first * MFB_CT_BYTES, mb.enc_out, cnt * MFB_CT_BYTES, cudaMemcpyDeviceToHost, mb.stream));
    return MFB_OK;
  }
  for (int g = 0; g < nsegs; g++) {
    const size_t lo = segs[g].first > first ? segs[g].first : first;
    const size_t hi_s = segs[g].first + segs[g].count, hi_m = first + cnt, hi = hi_s < hi_m ? hi_s : hi_m;
    if (lo >= hi) continue;
    SET_CUDA(cudaMemcpyAsync(segs[g].dst + (lo - segs[g].first) * MFB_CT_BYTES, mb.enc_out + (lo - first) * MFB_CT_BYTES,
                             (hi - lo) * MFB_CT_BYTES, cudaMemcpyDeviceToHost, mb.stream));
  }
  return MFB_OK;
}

static int encrypt_args_ok(const char *who, const void *seed, const void *sk, const void *msg, mfb_entropy_fn draw, int ent_stride,
                           int ent_nbytes) {
  if (!seed || !sk || !msg || !draw) {
    snprintf(g_set_err, sizeof(g_set_err), "%s: null pointer", who);
    return MFB_EARG;
  }
  if (ent_nbytes < 0 || ent_nbytes > 88 || ent_stride < ent_nbytes || ent_stride <= 0) {
    snprintf(g_set_err, sizeof(g_set_err), "%s: need 0 <= ent_nbytes <= 88 and ent_stride >= max(1, ent_nbytes)", who);
    return MFB_EARG;
  }
  return MFB_OK;
}

// mfb_encrypt_cb over a device set, entropy in the REFERENCE'S ORDER: the calling thread draws it piece by piece and in
// order (a hooked, deterministic source sees the reference's sequence); piece k is encrypted by member k mod size, so
// the members work on different pieces at the same time and the call is entropy-bound instead of AES-bound.
MFB_API int mfb_set_encrypt_cb(mfb_set *s, const uint8_t seed[40], uint64_t offset, const uint64_t *sk_flat, const uint64_t *msg,
                               mfb_entropy_fn draw, void *user, int ent_stride, int ent_nbytes, size_t count, uint8_t *out_c8) {
  g_set_err[0] = 0;
  if (count == 0) return MFB_OK;
  SET_TRY(check_usable(s, "mfb_set_encrypt_cb"));
  SET_TRY(encrypt_args_ok("mfb_set_encrypt_cb", seed, sk_flat, msg, draw, ent_stride, ent_nbytes));
  if (!out_c8) return set_fail(MFB_EARG, "mfb_set_encrypt_cb: null pointer");
  const size_t world = s->m.size();
  const size_t piece = (size_t)mfb_device_sm_count(s->m[0].ctx) * 110, first_piece = (size_t)mfb_device_sm_count(s->m[0].ctx) * 16;
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
  auto body = [&]() -> int {
    for (size_t i = 0; i < world && i < npieces; i++) {
      Member &mb = s->m[i];
      SET_CUDA(cudaSetDevice(mb.device));
      size_t cap;
      cap = mb.skf ? MFB_PLANAR_U64 * 8 : 0;
      SET_TRY(grow((void **)&mb.skf, &cap, MFB_PLANAR_U64 * 8, false));
      cap = mb.skp ? MFB_PLANAR_U64 * 8 : 0;
      SET_TRY(grow((void **)&mb.skp, &cap, MFB_PLANAR_U64 * 8, false));
      SET_TRY(grow((void **)&mb.enc_ent, &mb.enc_ent_cap, slots * piece * (size_t)ent_stride, false));
      cap = mb.enc_msg_cap;
      SET_TRY(grow((void **)&mb.enc_msg, &cap, slots * piece * 8, false));  // separate buffer: always 8-byte aligned
      mb.enc_msg_cap = cap;
      SET_TRY(grow((void **)&mb.enc_out, &mb.enc_out_cap, slots * piece * MFB_CT_BYTES, false));
      if (mb.ent_pin_cap < piece * (size_t)ent_stride) {
        for (int k = 0; k < 2; k++) {
          size_t c2 = mb.ent_pin_cap;
          SET_TRY(grow((void **)&mb.ent_pin[k], &c2, piece * (size_t)ent_stride, true));
        }
        mb.ent_pin_cap = piece * (size_t)ent_stride;
      }
      for (int k = 0; k < 2; k++)
        if (!mb.ent_free[k]) SET_CUDA(cudaEventCreateWithFlags(&mb.ent_free[k], cudaEventDisableTiming));
      SET_CUDA(cudaMemcpyAsync(mb.skf, sk_flat, MFB_FLAT_SK_U64 * 8, cudaMemcpyHostToDevice, mb.stream));
      SET_TRY(mfb_flat_to_planar_dev(mb.ctx, mb.skf, MFB_N, 1, mb.skp, mb.stream));
      // the member's messages, all its pieces up front (not one pageable copy per piece between the kernels)
      for (size_t k = i; k < npieces; k += world)
        SET_CUDA(cudaMemcpyAsync(mb.enc_msg + (k / world) * piece, msg + start[k], len[k] * 8, cudaMemcpyHostToDevice, mb.stream));
    }
    std::vector<int> uses(world, 0);
    for (size_t k = 0; k < npieces; k++) {
      Member &mb = s->m[k % world];
      SET_CUDA(cudaSetDevice(mb.device));
      const size_t slot = k / world, cnt = len[k];
      const int b = uses[k % world] & 1;
      if (uses[k % world] >= 2) SET_CUDA(cudaEventSynchronize(mb.ent_free[b]));  // the copy that last read this buffer is done
      uses[k % world]++;
      draw(user, mb.ent_pin[b], cnt * (size_t)ent_stride);
      uint8_t *d_ent = mb.enc_ent + slot * piece * (size_t)ent_stride;
      SET_CUDA(cudaMemcpyAsync(d_ent, mb.ent_pin[b], cnt * (size_t)ent_stride, cudaMemcpyHostToDevice, mb.stream));
      SET_CUDA(cudaEventRecord(mb.ent_free[b], mb.stream));
      SET_TRY(mfb_encrypt_dev(mb.ctx, seed, offset + start[k] * (uint64_t)MFB_CTR_CT, mb.skp, mb.enc_msg + slot * piece, d_ent, ent_stride,
                              ent_nbytes, cnt, mb.enc_out + slot * piece * MFB_CT_BYTES, mb.stream));
    }
    for (size_t k = 0; k < npieces; k++) {  // records back, piece by piece (each copy waits for its member's stream)
      Member &mb = s->m[k % world];
      SET_CUDA(cudaSetDevice(mb.device));
      SET_CUDA(cudaMemcpyAsync(out_c8 + start[k] * MFB_CT_BYTES, mb.enc_out + (k / world) * piece * MFB_CT_BYTES, len[k] * MFB_CT_BYTES,
                               cudaMemcpyDeviceToHost, mb.stream));
    }
    return MFB_OK;
  };
  int rc = body();
  // (no exchange kernels here: a failure does not desynchronise the members, the set stays usable)
  char keep[256];
  snprintf(keep, sizeof(keep), "%s", rc != MFB_OK ? mfb_set_last_error() : "");
  for (size_t i = 0; i < world; i++) {
    Member &mb = s->m[i];
    cudaSetDevice(mb.device);
    const cudaError_t e = cudaStreamSynchronize(mb.stream);
    if (e != cudaSuccess && rc == MFB_OK) {
      snprintf(keep, sizeof(keep), "mfb_set_encrypt_cb: device %d: %s", mb.device, cudaGetErrorString(e));
      rc = MFB_ECUDA;
    }
    scrub_member_secrets(mb);  // the key and the noise are secret, on error paths too
  }
  cudaSetDevice(s->m[0].device);
  if (rc != MFB_OK) snprintf(g_set_err, sizeof(g_set_err), "%s", keep);
  return rc;
}

// The same for an entropy source WITHOUT an order to preserve (the OS: getrandom(2)): member i takes the contiguous range
// i of the ciphertexts and is driven by ITS OWN host thread, which draws that range's entropy (draw is called
// concurrently from the member threads and must be thread-safe) while its device encrypts — entropy, upload, AES and
// download all scale with the number of members.  The records go straight to their destinations: out_c8, or nsegs
// segments of the record index space (setup(): crs->s, crs->as, crs->t, crs->v).
MFB_API int mfb_set_encrypt_par(mfb_set *s, const uint8_t seed[40], uint64_t offset, const uint64_t *sk_flat, const uint64_t *msg,
                                mfb_entropy_fn draw, void *user, int ent_stride, int ent_nbytes, size_t count, uint8_t *out_c8,
                                const mfb_c8_segment *segs, int nsegs) {
  g_set_err[0] = 0;
  if (count == 0) return MFB_OK;
  SET_TRY(check_usable(s, "mfb_set_encrypt_par"));
  SET_TRY(encrypt_args_ok("mfb_set_encrypt_par", seed, sk_flat, msg, draw, ent_stride, ent_nbytes));
  if ((out_c8 == nullptr) == (segs == nullptr) || (segs && nsegs < 1))
    return set_fail(MFB_EARG, "mfb_set_encrypt_par: give either out_c8 or segments");
  const size_t world = s->m.size();
  const size_t piece = (size_t)mfb_device_sm_count(s->m[0].ctx) * 110, first_piece = (size_t)mfb_device_sm_count(s->m[0].ctx) * 16;
  std::vector<int> rcs(world, MFB_OK);
  std::vector<std::string> errs(world);
  auto work = [&](size_t i) {
    g_set_err[0] = 0;
    Member &mb = s->m[i];
    size_t first, cnt;
    split_range(count, world, i, &first, &cnt);
    int rc = member_encrypt(mb, seed, offset, sk_flat, msg, draw, user, ent_stride, ent_nbytes, first, cnt, piece, first_piece);
    if (rc == MFB_OK) rc = collect_records(mb, first, cnt, out_c8, segs, nsegs);
    if (rc == MFB_OK && cudaStreamSynchronize(mb.stream) != cudaSuccess) {
      snprintf(g_set_err, sizeof(g_set_err), "mfb_set_encrypt_par: device %d: %s", mb.device, cudaGetErrorString(cudaGetLastError()));
      rc = MFB_ECUDA;
    }
    if (rc != MFB_OK) errs[i] = mfb_set_last_error();  // (thread-local: copy it out before the thread ends)
    scrub_member_secrets(mb);
    rcs[i] = rc;
  };
  std::vector<std::thread> pool;
  for (size_t i = 1; i < world; i++) pool.emplace_back(work, i);
  work(0);
  for (auto &t : pool) t.join();
  cudaSetDevice(s->m[0].device);
  for (size_t i = 0; i < world; i++)
    if (rcs[i] != MFB_OK) {
      snprintf(g_set_err, sizeof(g_set_err), "%s", errs[i].c_str());
      return rcs[i];
    }
  return MFB_OK;
}

// ---------------------------------------------------------------------------------------------------- prover pipeline
static int prove_body(mfb_set *s, mfb_ssp *ssp, const mfb_set_region *rs, const mfb_set_region *ras, const uint64_t *witness_limbs,
                      size_t nlimbs, uint64_t delta, const uint8_t *seed, uint64_t bt_offset, const uint8_t *bt_recs, size_t M,
                      uint64_t *const *host, bool want_bw) {
  const size_t D = mfb_ssp_degree_bound(ssp), world = s->m.size();
  Member &p = s->m[0];
  // 1. the polynomial step on the primary FIRST (it needs only the witness): queued without a host round trip; everything
  //    the host does from here on runs beside it, and the other members wait for it ON THE DEVICE
  SET_CUDA(cudaSetDevice(p.device));
  const uint32_t *wvh = nullptr;
  SET_TRY(mfb_ssp_prover_polys_resident_async(p.ctx, ssp, witness_limbs, nlimbs, delta, p.stream, &wvh));
  SET_CUDA(cudaEventRecord(s->ev_polys, p.stream));
  // 2. b_w (a few selected ciphertexts regenerated from AES) on a member that would otherwise wait for the polynomial step
  Member &bm = s->m[world > 1 ? 1 : 0];
  if (want_bw) {
    SET_CUDA(cudaSetDevice(bm.device));
    SET_TRY(mfb_b_w_dev(bm.ctx, seed, bt_offset, bt_recs, M, witness_limbs, nlimbs, delta, bm.res + 4 * SLOT, bm.stream));
  }
  // 3. every member: its slices of w, v, h over NVLink (one strided copy), both two-vector passes over its shards, then ONE
  //    kernel that finishes its four partial sums and pushes them to every member (it never waits)
  for (size_t i = 0; i < world; i++) {
    Member &mb = s->m[i];
    SET_CUDA(cudaSetDevice(mb.device));
    const size_t cnt = rs->count[i], first = rs->first[i];
    const uint32_t *cw, *cv, *ch;
    if (i == 0) {
      cw = wvh + first;
      cv = wvh + D + first;
      ch = wvh + 2 * D + first;
    } else {
      size_t cap = mb.co_cap * 4;
      SET_TRY(grow((void **)&mb.co, &cap, (3 * cnt + 4) * 4, false));
      mb.co_cap = cap / 4;
      SET_CUDA(cudaStreamWaitEvent(mb.stream, s->ev_polys, 0));
      if (cnt)  // rows w, v, h of the primary's [3][D] array -> the member's [3][cnt]
        SET_CUDA(cudaMemcpy2DAsync(mb.co, cnt * 4, wvh + first, D * 4, cnt * 4, 3, cudaMemcpyDefault, mb.stream));
      cw = mb.co;
      cv = mb.co + cnt;
      ch = mb.co + 2 * cnt;
    }
    SET_TRY(mfb_lincomb2_partials_dev(mb.ctx, (const uint64_t *)mfb_region_cts(rs->shard[i]), cw, ch, cnt, 0, mb.stream));
    SET_TRY(mfb_lincomb2_partials_dev(mb.ctx, (const uint64_t *)mfb_region_cts(ras->shard[i]), cv, ch, cnt, 1, mb.stream));
    SET_TRY(mfb_peer_finish4_push_dev(mb.ctx, mb.group, mb.stream));
  }
  // 4. the incoming accumulators (usually all zero: a proof starts from proof_init) — staged while the devices work
  const bool any = stage_in(s, host, 4);
  SET_CUDA(cudaSetDevice(p.device));
  if (any) SET_CUDA(cudaMemcpyAsync(p.res, s->acc_pin, 4 * SLOT * 8, cudaMemcpyHostToDevice, p.stream));
  // 5. the waiting halves of the exchange, after every member's pushes are queued
  for (size_t i = 0; i < world; i++) {
    Member &mb = s->m[i];
    SET_CUDA(cudaSetDevice(mb.device));
    SET_TRY(mfb_peer_wait4_dev(mb.ctx, mb.group, i == 0 && any ? mb.res : nullptr, mb.res, SLOT, mb.stream));
  }
  SET_CUDA(cudaSetDevice(p.device));
  SET_CUDA(cudaMemcpyAsync(s->acc_pin, p.res, 4 * SLOT * 8, cudaMemcpyDeviceToHost, p.stream));
  if (want_bw) {
    SET_CUDA(cudaSetDevice(bm.device));
    SET_CUDA(cudaMemcpyAsync(s->acc_pin + 4 * SLOT, bm.res + 4 * SLOT, MFB_FLAT_CT_U64 * 8, cudaMemcpyDeviceToHost, bm.stream));
  }
  return MFB_OK;
}

// mfb_prove_resident_bw over sharded regions (see mfb200.h)
MFB_API int mfb_set_prove_resident_bw(mfb_set *s, mfb_ssp *ssp, const mfb_set_region *rs, const mfb_set_region *ras,
                                      const uint64_t *witness_limbs, size_t nlimbs, uint64_t delta, const uint8_t seed[40],
                                      uint64_t bt_offset, const uint8_t *bt_recs, size_t M, uint64_t *v_w_flat_inout,
                                      uint64_t *h_flat_inout, uint64_t *hat_v_flat_inout, uint64_t *hat_h_flat_inout,
                                      uint64_t *b_w_flat_out) {
  g_set_err[0] = 0;
  SET_TRY(check_usable(s, "mfb_set_prove_resident"));
  if (!ssp || !rs || !ras || !witness_limbs || !v_w_flat_inout || !h_flat_inout || !hat_v_flat_inout || !hat_h_flat_inout)
    return set_fail(MFB_EARG, "mfb_set_prove_resident: null pointer");
  if (b_w_flat_out && (!seed || !bt_recs || M < 1)) return set_fail(MFB_EARG, "mfb_set_prove_resident_bw: b_w needs the seed and the t | v records");
  const size_t D = mfb_ssp_degree_bound(ssp), world = s->m.size();
  if (rs->total != D || ras->total != D) return set_fail(MFB_EARG, "mfb_set_prove_resident: the regions must hold D ciphertexts");
  for (size_t i = 0; i < world; i++)
    if (rs->first[i] != ras->first[i] || rs->count[i] != ras->count[i])
      return set_fail(MFB_EARG, "mfb_set_prove_resident: the two regions are sharded differently");
  SET_TRY(ensure_acc_pin(s));
  uint64_t *host[NACC] = {v_w_flat_inout, h_flat_inout, hat_v_flat_inout, hat_h_flat_inout, b_w_flat_out};
  const int rc = finish_call(s, prove_body(s, ssp, rs, ras, witness_limbs, nlimbs, delta, seed, bt_offset, bt_recs, M, host,
                                           b_w_flat_out != nullptr));
  if (rc == MFB_OK) stage_out(s, host, b_w_flat_out ? NACC : 4);
  return rc;
}

MFB_API int mfb_set_prove_resident(mfb_set *s, mfb_ssp *ssp, const mfb_set_region *rs, const mfb_set_region *ras,
                                   const uint64_t *witness_limbs, size_t nlimbs, uint64_t delta, uint64_t *v_w_flat_inout,
                                   uint64_t *h_flat_inout, uint64_t *hat_v_flat_inout, uint64_t *hat_h_flat_inout) {
  return mfb_set_prove_resident_bw(s, ssp, rs, ras, witness_limbs, nlimbs, delta, nullptr, 0, nullptr, 0, v_w_flat_inout, h_flat_inout,
                                   hat_v_flat_inout, hat_h_flat_inout, nullptr);
}

}  // extern "C"
