// mdc_runtime.cpp -- host runtime above the kernels, behind the same C ABI: streaming of a host-resident ensemble
// through the device in row slabs (upload || analyse || download on three host threads, no interpreter in the loop),
// and column sharding over ranks with the observation halo carried by NCCL (grouped ncclSend / ncclRecv).
//
// What the reference has at this place: nothing -- LETKF<Tag>::Analyse (LETKF.hpp:63-119) walks one in-memory
// ensemble on one core.  The drop-in LETKF<CudaBackendTag>::Analyse (metada_b200/host/algorithms/LETKF.hpp) calls
// mdc_stream_analyse with the members' host pointers (State::getDataPtr, State.hpp:229-242), so an ensemble larger
// than the device (C5: 86 GB) and a job spread over the GPUs of a box go through the same call.
//
// Everything here is plain C++ on top of the library's own C entry points (mdc_ens_*, mdc_obs_*, mdc_letkf_analyse);
// NCCL is bound at run time (dlopen of libnccl.so.2), so the library loads where NCCL is absent and
// mdc_comm_init reports that.
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/metada_cuda_c_api.h"

namespace {

// ---- NCCL, bound at run time
struct NcclId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, NcclComm, void*) = nullptr;
  int (*Recv)(void*, size_t, int, int, NcclComm, void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string err;
  bool load() {
    if (h) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) { err = std::string("NCCL not found: ") + dlerror(); return false; }
#define MDC_SYM(field, sym)                                             \
  field = reinterpret_cast<decltype(field)>(dlsym(h, sym));             \
  if (!field) { err = std::string("NCCL symbol missing: ") + sym; h = nullptr; return false; }
    MDC_SYM(GetUniqueId, "ncclGetUniqueId") MDC_SYM(CommInitRank, "ncclCommInitRank") MDC_SYM(CommDestroy, "ncclCommDestroy")
    MDC_SYM(GroupStart, "ncclGroupStart") MDC_SYM(GroupEnd, "ncclGroupEnd") MDC_SYM(Send, "ncclSend") MDC_SYM(Recv, "ncclRecv")
    MDC_SYM(AllReduce, "ncclAllReduce") MDC_SYM(GetErrorString, "ncclGetErrorString")
#undef MDC_SYM
    return true;
  }
};
NcclApi g_nccl;
constexpr int kNcclFloat64 = 8, kNcclMax = 2;   // ncclDataType_t / ncclRedOp_t values (nccl.h)

// ---- row slabs (the same arithmetic as metada_b200/parallel.py, which the gloo tests pin)
inline void slab_bounds(int gny, int rank, int world, int& y0, int& y1) {
  y0 = (int)(((long long)gny * rank) / world);
  y1 = (int)(((long long)gny * (rank + 1)) / world);
}
inline int owner_of_row(int y, int gny, int world) {
  y = std::min(std::max(y, 0), gny - 1);
  int r = (int)(((long long)y * world) / gny);
  int lo, hi;
  slab_bounds(gny, r, world, lo, hi);
  if (y < lo) --r; else if (y >= hi) ++r;
  return r;
}
// rows [lo, hi) of src's own observations that dst's columns can reach; false if none
inline bool halo_range(int gny, int world, int reach, int src, int dst, int& lo, int& hi) {
  int d0, d1, s0, s1;
  slab_bounds(gny, dst, world, d0, d1);
  slab_bounds(gny, src, world, s0, s1);
  lo = std::max(src > 0 ? s0 : -(1 << 30), d0 - reach);          // (edge slabs also own the out-of-grid observations)
  hi = std::min(src < world - 1 ? s1 : (1 << 30), d1 + reach);
  return lo < hi;
}

template <typename T>
class Channel {   // unbounded queue between two pipeline stages; close() wakes the consumer with "no more"
 public:
  void put(T v) { { std::lock_guard<std::mutex> l(m_); q_.push_back(v); } cv_.notify_one(); }
  void close() { { std::lock_guard<std::mutex> l(m_); closed_ = true; } cv_.notify_all(); }
  bool get(T& out) {
    std::unique_lock<std::mutex> l(m_);
    cv_.wait(l, [&] { return !q_.empty() || closed_; });
    if (q_.empty()) return false;
    out = q_.front();
    q_.pop_front();
    return true;
  }
 private:
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<T> q_;
  bool closed_ = false;
};

struct ObsView {
  int64_t P;
  const int32_t *x, *y, *z;
  const double *val, *err;
  const uint8_t* valid;
};

}  // namespace

struct mdc_stream {
  mdc_stream_config cfg{};
  int device = 0;
  int reach = 0, nslab = 0, nslots = 0, depth = 0;
  std::vector<std::pair<int, int>> bounds;      // slab row ranges
  std::vector<std::vector<int>> neigh;          // slabs within reach of a slab
  std::vector<mdc_ctx*> ctxs;                   // one context (stream) per slot
  std::vector<mdc_ens*> ens;                    // one ensemble store per slot
  std::vector<mdc_obs*> obs;                    // one observation store per slab (on the context of slot s % nslots)
  void* pool = nullptr;                         // device buffer for the halo rows between slabs
  int64_t pool_bytes = 0;
  // sharding
  NcclComm comm = nullptr;
  int rank = 0, world = 1;
  struct Strip { mdc_ens* ens = nullptr; mdc_obs* obs = nullptr; int r0 = 0, r1 = 0; };
  std::map<int, Strip> strips;                  // 0: top, 1: bottom edge strip
  void* xbuf = nullptr;                         // device buffer for the rows sent to / received from other ranks
  int64_t xbuf_bytes = 0;
  double last_edge_ms = 0.0, last_stream_ms = 0.0;
  int64_t last_halo_rows = 0;
  char err[512] = {0};
};

namespace {

int fail(mdc_stream* s, int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(s->err, sizeof(s->err), fmt, ap);
  va_end(ap);
  return code;
}
#define MDC_RT(s, ctx, call)                                                                  \
  do {                                                                                        \
    int rc_ = (call);                                                                         \
    if (rc_) return fail((s), rc_, "%s: %s", #call, (ctx) ? mdc_last_error(ctx) : "failed");  \
  } while (0)

int grow_dev(mdc_stream* s, void** buf, int64_t* have, int64_t need) {
  if (*have >= need) return MDC_OK;
  if (*buf) mdc_dev_free(s->ctxs[0], *buf);
  *buf = nullptr;
  MDC_RT(s, s->ctxs[0], mdc_dev_malloc(s->ctxs[0], std::max<int64_t>(need, 8), buf));
  *have = need;
  return MDC_OK;
}

// observations (by index list) into a store, created on first use
int fill_obs(mdc_stream* s, mdc_ctx* ctx, mdc_obs** store, const ObsView& o, const std::vector<int64_t>& idx) {
  const size_t n = idx.size();
  std::vector<int32_t> x(n), y(n), z(n);
  std::vector<double> v(n), e(n);
  std::vector<uint8_t> ok(n);
  for (size_t i = 0; i < n; ++i) {
    const int64_t a = idx[i];
    x[i] = o.x[a]; y[i] = o.y[a]; z[i] = o.z ? o.z[a] : 0; v[i] = o.val[a]; e[i] = o.err[a]; ok[i] = o.valid ? o.valid[a] : 1;
  }
  if (!*store) MDC_RT(s, ctx, mdc_obs_create(ctx, (int64_t)n, x.data(), y.data(), z.data(), v.data(), e.data(), ok.data(), idx.data(), store));
  else MDC_RT(s, ctx, mdc_obs_assign(*store, (int64_t)n, x.data(), y.data(), z.data(), v.data(), e.data(), ok.data(), idx.data()));
  return MDC_OK;
}

// ---- multi-rank prologue: H on this rank's edge strips (from the HOST members), rows packed per destination rank,
// grouped NCCL send / recv; returns the received rows as (device pointer, rows) above / below this rank's range
int exchange_halo(mdc_stream* s, double* const* members, int host_row0, int host_ny, const ObsView& o,
                  std::vector<std::pair<const double*, int64_t>>& ext_top, std::vector<std::pair<const double*, int64_t>>& ext_bottom) {
  const int gny = s->cfg.gny, W = s->world, me = s->rank, rd = s->cfg.k + 8;
  mdc_ctx* ctx = s->ctxs[0];
  int y0, y1;
  slab_bounds(gny, me, W, y0, y1);
  // observations owned by this rank, and how many rows every (src -> dst) message carries: every rank knows the
  // global observation set, so the counts need no exchange
  std::vector<int64_t> own;
  std::vector<std::vector<int64_t>> cnt(W, std::vector<int64_t>(W, 0));
  for (int64_t a = 0; a < o.P; ++a) {
    const int src = owner_of_row(o.y[a], gny, W);
    if (src == me) own.push_back(a);
    for (int dst = 0; dst < W; ++dst) {
      int lo, hi;
      if (dst != src && halo_range(gny, W, s->reach, src, dst, lo, hi) && o.y[a] >= lo && o.y[a] < hi) ++cnt[src][dst];
    }
  }
  int64_t send_rows = 0, recv_rows = 0;
  for (int r = 0; r < W; ++r) { send_rows += cnt[me][r]; recv_rows += cnt[r][me]; }
  if (int rc = grow_dev(s, &s->xbuf, &s->xbuf_bytes, (send_rows + recv_rows + 1) * rd * 8)) return rc;
  double* sendbuf = static_cast<double*>(s->xbuf);
  double* recvbuf = sendbuf + send_rows * rd;
  std::vector<int64_t> send_off(W, 0), recv_off(W, 0);
  { int64_t a = 0, b = 0; for (int r = 0; r < W; ++r) { send_off[r] = a; a += cnt[me][r]; recv_off[r] = b; b += cnt[r][me]; } }
  // edge strips: side 0 = towards lower ranks, 1 = towards higher ranks
  for (int side = 0; side < 2; ++side) {
    int lo = 1 << 30, hi = -(1 << 30);
    std::vector<int> dsts;
    for (int d = 0; d < W; ++d) {
      int l, h;
      if (d != me && (side == 0 ? d < me : d > me) && halo_range(gny, W, s->reach, me, d, l, h)) { dsts.push_back(d); lo = std::min(lo, l); hi = std::max(hi, h); }
    }
    if (dsts.empty()) continue;
    const int r0 = std::max(lo, y0), r1 = std::min(hi, y1);
    if (r0 >= r1) continue;
    const int halo = r1 < gny ? 1 : 0;
    mdc_stream::Strip& st = s->strips[side];
    if (st.ens && (st.r0 != r0 || st.r1 != r1)) { mdc_ens_destroy(st.ens); st.ens = nullptr; }
    if (!st.ens) {
      MDC_RT(s, ctx, mdc_ens_create(ctx, s->cfg.gnx, (r1 - r0) + halo, s->cfg.nz, s->cfg.k, &st.ens));
      MDC_RT(s, ctx, mdc_ens_set_domain(st.ens, 0, r0, s->cfg.gnx, gny, s->cfg.gnx, r1 - r0));
      st.r0 = r0; st.r1 = r1;
    }
    MDC_RT(s, ctx, mdc_ens_upload_members_rows(st.ens, 0, s->cfg.k, members, host_ny, r0 - host_row0));
    std::vector<int64_t> idx;
    for (int64_t a : own) if (o.y[a] >= lo && o.y[a] < hi) idx.push_back(a);
    if (int rc = fill_obs(s, ctx, &st.obs, o, idx)) return rc;
    MDC_RT(s, ctx, mdc_hx_idw4(st.ens, st.obs));
    for (int d : dsts) {
      int l, h;
      halo_range(gny, W, s->reach, me, d, l, h);
      int64_t got = 0;
      if (cnt[me][d] > 0) MDC_RT(s, ctx, mdc_obs_pack_rows(st.obs, l, h, sendbuf + send_off[d] * rd, cnt[me][d], &got));
      if (got != cnt[me][d]) return fail(s, MDC_ERR_INVALID, "halo: packed %lld rows for rank %d, expected %lld", (long long)got, d, (long long)cnt[me][d]);
    }
  }
  MDC_RT(s, ctx, mdc_ctx_sync(ctx));
  void* stream = mdc_ctx_stream(ctx);
  int rc = g_nccl.GroupStart();
  for (int r = 0; r < W && !rc; ++r) {
    if (r == me) continue;
    if (cnt[me][r] > 0) rc = g_nccl.Send(sendbuf + send_off[r] * rd, (size_t)(cnt[me][r] * rd), kNcclFloat64, r, s->comm, stream);
    if (!rc && cnt[r][me] > 0) rc = g_nccl.Recv(recvbuf + recv_off[r] * rd, (size_t)(cnt[r][me] * rd), kNcclFloat64, r, s->comm, stream);
  }
  if (!rc) rc = g_nccl.GroupEnd();
  if (rc) return fail(s, MDC_ERR_CUDA, "NCCL halo exchange: %s", g_nccl.GetErrorString(rc));
  MDC_RT(s, ctx, mdc_ctx_sync(ctx));
  for (int r = 0; r < W; ++r)
    if (r != me && cnt[r][me] > 0) (r < me ? ext_top : ext_bottom).push_back({recvbuf + recv_off[r] * rd, cnt[r][me]});
  s->last_halo_rows = recv_rows;
  return MDC_OK;
}

// ---- the slab pipeline
int run_pipeline(mdc_stream* s, double* const* members, int host_row0, int host_ny, const ObsView& o,
                 const mdc_letkf_params* params, const std::vector<std::pair<const double*, int64_t>>& ext_top,
                 const std::vector<std::pair<const double*, int64_t>>& ext_bottom, mdc_letkf_stats* total) {
  const int S = s->nslab, K = s->nslots, R = s->reach, D = s->depth, gny = s->cfg.gny, rd = s->cfg.k + 8;
  mdc_letkf_params prm = *params;
  if (S > 1) prm.sm_reserve = s->cfg.sm_reserve;
  // observations of every slab; halo rows every slab sends to the slabs within its reach
  std::vector<std::vector<int64_t>> own(S);
  {
    // slab of a row by bisection over the slab starts; the grid's first / last slab also own the out-of-grid rows
    const int glo = s->bounds[0].first > 0 ? s->bounds[0].first : -(1 << 30);
    const int ghi = s->bounds[S - 1].second < gny ? s->bounds[S - 1].second : (1 << 30);
    for (int64_t a = 0; a < o.P; ++a) {
      const int y = o.y[a];
      if (y < glo || y >= ghi) continue;
      int lo = 0, hi = S - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s->bounds[mid].first <= y) lo = mid; else hi = mid - 1;
      }
      own[lo].push_back(a);
    }
  }
  struct Halo { int dst, lo, hi; int64_t n; double* buf; };
  std::vector<std::vector<Halo>> halo(S);
  int64_t need = 0;
  for (int sl = 0; sl < S; ++sl)
    for (int dst : s->neigh[sl]) {
      const int lo = s->bounds[dst].first - R, hi = s->bounds[dst].second + R;
      int64_t n = 0;
      for (int64_t a : own[sl]) if (o.y[a] >= lo && o.y[a] < hi) ++n;
      halo[sl].push_back({dst, lo, hi, n, nullptr});
      need += n * rd * 8;
    }
  if (int rc = grow_dev(s, &s->pool, &s->pool_bytes, need)) return rc;
  {
    double* p = static_cast<double*>(s->pool);
    for (auto& hs : halo) for (auto& h : hs) { h.buf = p; p += h.n * rd; }
  }
  // every slab's observations go to the device before the member traffic starts (a small copy issued later queues
  // on the copy engine behind ~190 MB member batches)
  for (int sl = 0; sl < S; ++sl)
    if (int rc = fill_obs(s, s->ctxs[sl % K], &s->obs[sl], o, own[sl])) return rc;
  int max_rows = 0;
  for (auto& b : s->bounds) max_rows = std::max(max_rows, b.second - b.first);
  for (int w = 0; w < K; ++w)
    if (!s->ens[w]) MDC_RT(s, s->ctxs[w], mdc_ens_create(s->ctxs[w], s->cfg.gnx, max_rows + 1, s->cfg.nz, s->cfg.k, &s->ens[w]));

  Channel<int> free_slots, q_up, q_down;
  for (int w = 0; w < K; ++w) free_slots.put(w);
  std::mutex err_m;
  std::string err_msg;
  std::atomic<int> err_code{0};
  auto set_err = [&](int rc, const std::string& what, mdc_ctx* ctx) {
    std::lock_guard<std::mutex> l(err_m);
    if (!err_code) { err_code = rc ? rc : MDC_ERR_INVALID; err_msg = what + ": " + (ctx ? mdc_last_error(ctx) : ""); }
    q_up.close(); q_down.close(); free_slots.close();
  };
  std::vector<mdc_letkf_stats> stats(S);
  std::vector<int> slot_of(S, -1);

  std::thread uploader([&] {
    for (int sl = 0; sl < S && !err_code; ++sl) {
      int w;
      if (!free_slots.get(w)) return;
      if (w != sl % K) { set_err(MDC_ERR_INVALID, "pipeline slots out of order", nullptr); return; }
      const int y0 = s->bounds[sl].first, y1 = s->bounds[sl].second;
      mdc_ens* e = s->ens[w];
      int rc = mdc_ens_set_rows(e, (y1 - y0) + (y1 < gny ? 1 : 0));
      if (!rc) rc = mdc_ens_set_domain(e, 0, y0, s->cfg.gnx, gny, s->cfg.gnx, y1 - y0);
      if (!rc) rc = mdc_ens_upload_members_rows(e, 0, s->cfg.k, members, host_ny, y0 - host_row0);
      if (!rc) rc = mdc_ctx_sync(s->ctxs[w]);
      if (rc) { set_err(rc, "slab upload", s->ctxs[w]); return; }
      q_up.put(sl);
    }
    q_up.close();
  });
  std::thread compute([&] {
    int loaded = -1;
    for (int sl = 0; sl < S && !err_code; ++sl) {
      while (loaded < std::min(sl + D, S - 1)) {     // slabs sl+1 .. sl+D must have gone through H: their rows feed sl
        int got;
        if (!q_up.get(got)) return;
        loaded = got;
        const int w = loaded % K;
        slot_of[loaded] = w;
        int rc = mdc_hx_idw4(s->ens[w], s->obs[loaded]);
        for (auto& h : halo[loaded]) {
          if (rc || h.n == 0) continue;
          int64_t n = 0;
          rc = mdc_obs_pack_rows(s->obs[loaded], h.lo, h.hi, h.buf, h.n, &n);
          if (!rc && n != h.n) { set_err(MDC_ERR_INVALID, "slab halo row count", nullptr); return; }
        }
        if (rc) { set_err(rc, "slab H / pack", s->ctxs[w]); return; }
      }
      const int w = slot_of[sl];
      mdc_obs* ob = s->obs[sl];
      int rc = 0;
      for (int src : s->neigh[sl])
        for (auto& h : halo[src])
          if (!rc && h.dst == sl && h.n > 0) rc = mdc_obs_append_rows(ob, h.buf, h.n);
      if (s->bounds[sl].first - R < s->cfg.row0)      // rows from other ranks: supersets are harmless
        for (auto& e : ext_top) if (!rc) rc = mdc_obs_append_rows(ob, e.first, e.second);
      if (s->bounds[sl].second + R > s->cfg.row1)
        for (auto& e : ext_bottom) if (!rc) rc = mdc_obs_append_rows(ob, e.first, e.second);
      if (!rc) rc = mdc_letkf_analyse(s->ens[w], ob, &prm, &stats[sl]);
      if (rc) { set_err(rc, "slab analysis", s->ctxs[w]); return; }
      q_down.put(sl);
    }
    q_down.close();
  });
  std::thread downloader([&] {
    int sl;
    while (!err_code && q_down.get(sl)) {
      const int w = sl % K, y0 = s->bounds[sl].first, y1 = s->bounds[sl].second;
      const int rc = mdc_ens_download_members_rows(s->ens[w], 0, s->cfg.k, members, host_ny, y0 - host_row0, y1 - y0);
      if (rc) { set_err(rc, "slab download", s->ctxs[w]); return; }
      free_slots.put(w);
    }
  });
  uploader.join();
  compute.join();
  downloader.join();
  if (err_code) return fail(s, err_code, "%s", err_msg.c_str());
  std::memset(total, 0, sizeof(*total));
  for (const auto& st : stats) {
    total->columns += st.columns; total->sum_local_obs += st.sum_local_obs; total->sum_sweeps += st.sum_sweeps;
    total->numeric_failures += st.numeric_failures; total->redo_transforms += st.redo_transforms;
    total->small_transforms += st.small_transforms;
    total->max_local_obs = std::max(total->max_local_obs, st.max_local_obs);
    total->max_sweeps = std::max(total->max_sweeps, st.max_sweeps);
    total->ms_hx += st.ms_hx; total->ms_index += st.ms_index; total->ms_columns += st.ms_columns; total->ms_total += st.ms_total;
  }
  return MDC_OK;
}

}  // namespace

extern "C" {

int mdc_stream_create(int device, const mdc_stream_config* cfg, mdc_stream** out) {
  if (!cfg || !out) return MDC_ERR_INVALID;
  *out = nullptr;
  if (cfg->gnx <= 0 || cfg->gny <= 0 || cfg->nz <= 0 || cfg->k <= 0 || cfg->row0 < 0 || cfg->row1 > cfg->gny || cfg->row0 >= cfg->row1 ||
      !(cfg->radius >= 0.0))
    return MDC_ERR_INVALID;
  mdc_stream* s = new mdc_stream;
  s->cfg = *cfg;
  s->device = device;
  s->reach = (int)std::floor(cfg->radius);
  const int rows = cfg->row1 - cfg->row0;
  // slab height: as asked, else >= 24 slabs per rank (fill and drain of the three-stage pipeline stay below ~10 %)
  // but not lower than 8 rows (per-slab launch and index costs)
  int slab_rows = cfg->slab_rows > 0 ? cfg->slab_rows : std::max(8, std::min(32, rows / 24));
  s->nslab = std::max(1, (rows + slab_rows - 1) / slab_rows);
  for (int i = 0; i < s->nslab; ++i)
    s->bounds.push_back({cfg->row0 + (int)(((long long)rows * i) / s->nslab), cfg->row0 + (int)(((long long)rows * (i + 1)) / s->nslab)});
  s->neigh.resize(s->nslab);
  for (int a = 0; a < s->nslab; ++a)
    for (int b = 0; b < s->nslab; ++b)
      if (a != b && s->bounds[b].first < s->bounds[a].second + s->reach && s->bounds[b].second > s->bounds[a].first - s->reach) {
        s->neigh[a].push_back(b);
        s->depth = std::max(s->depth, std::abs(a - b));
      }
  s->nslots = s->nslab > 1 ? std::max(std::max(3, cfg->slots), s->depth + 2) : 1;
  s->ctxs.assign(s->nslots, nullptr);
  s->ens.assign(s->nslots, nullptr);
  s->obs.assign(s->nslab, nullptr);
  for (int w = 0; w < s->nslots; ++w)
    if (int rc = mdc_ctx_create(device, &s->ctxs[w])) { mdc_stream_destroy(s); return rc; }
  *out = s;
  return MDC_OK;
}

int mdc_stream_destroy(mdc_stream* s) {
  if (!s) return MDC_OK;
  for (auto& kv : s->strips) { if (kv.second.obs) mdc_obs_destroy(kv.second.obs); if (kv.second.ens) mdc_ens_destroy(kv.second.ens); }
  for (mdc_obs* o : s->obs) if (o) mdc_obs_destroy(o);
  for (mdc_ens* e : s->ens) if (e) mdc_ens_destroy(e);
  if (!s->ctxs.empty() && s->ctxs[0]) {
    if (s->pool) mdc_dev_free(s->ctxs[0], s->pool);
    if (s->xbuf) mdc_dev_free(s->ctxs[0], s->xbuf);
  }
  if (s->comm && g_nccl.h) g_nccl.CommDestroy(s->comm);
  for (mdc_ctx* c : s->ctxs) if (c) mdc_ctx_destroy(c);
  delete s;
  return MDC_OK;
}

const char* mdc_stream_last_error(const mdc_stream* s) { return s ? s->err : "null stream handle"; }
int mdc_stream_slabs(const mdc_stream* s) { return s ? s->nslab : 0; }
int mdc_stream_slots(const mdc_stream* s) { return s ? s->nslots : 0; }

int mdc_stream_timings(const mdc_stream* s, double* edge_halo_ms, double* stream_ms, int64_t* halo_rows) {
  if (!s) return MDC_ERR_INVALID;
  if (edge_halo_ms) *edge_halo_ms = s->last_edge_ms;
  if (stream_ms) *stream_ms = s->last_stream_ms;
  if (halo_rows) *halo_rows = s->last_halo_rows;
  return MDC_OK;
}

int mdc_comm_get_unique_id(void* id, int bytes) {
  if (!id || bytes < (int)sizeof(NcclId)) return MDC_ERR_INVALID;
  if (!g_nccl.load()) return MDC_ERR_UNSUPPORTED;
  NcclId u;
  if (g_nccl.GetUniqueId(&u)) return MDC_ERR_CUDA;
  std::memcpy(id, &u, sizeof(u));
  return MDC_OK;
}

int mdc_comm_init(mdc_stream* s, const void* id, int rank, int nranks) {
  if (!s || !id || nranks < 1 || rank < 0 || rank >= nranks) return MDC_ERR_INVALID;
  int y0, y1;
  slab_bounds(s->cfg.gny, rank, nranks, y0, y1);
  if (y0 != s->cfg.row0 || y1 != s->cfg.row1)
    return fail(s, MDC_ERR_INVALID, "comm_init: rank %d of %d owns rows [%d, %d), the stream was created for [%d, %d)", rank, nranks, y0, y1,
                s->cfg.row0, s->cfg.row1);
  if (!g_nccl.load()) return fail(s, MDC_ERR_UNSUPPORTED, "%s", g_nccl.err.c_str());
  NcclId u;
  std::memcpy(&u, id, sizeof(u));
  MDC_RT(s, s->ctxs[0], mdc_ctx_sync(s->ctxs[0]));       // (makes the context's device current)
  const int rc = g_nccl.CommInitRank(&s->comm, nranks, u, rank);
  if (rc) return fail(s, MDC_ERR_CUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString(rc));
  s->rank = rank;
  s->world = nranks;
  return MDC_OK;
}

int mdc_comm_max(mdc_stream* s, double* value) {
  if (!s || !value) return MDC_ERR_INVALID;
  if (!s->comm) return MDC_OK;
  void* d = nullptr;
  MDC_RT(s, s->ctxs[0], mdc_dev_malloc(s->ctxs[0], 8, &d));
  int rc = mdc_dev_copy(s->ctxs[0], d, value, 8, 1);
  if (!rc) rc = g_nccl.AllReduce(d, d, 1, kNcclFloat64, kNcclMax, s->comm, mdc_ctx_stream(s->ctxs[0])) ? MDC_ERR_CUDA : 0;
  if (!rc) rc = mdc_dev_copy(s->ctxs[0], value, d, 8, 2);
  mdc_dev_free(s->ctxs[0], d);
  return rc ? fail(s, rc, "comm_max failed") : MDC_OK;
}

int mdc_comm_allgather_rows(mdc_stream* s, double* const* members) {
  // every rank holds full host members [nz][gny][gnx] of which it analysed its own rows: pass every rank's rows to
  // all the others (device staging, one group of ncclSend / ncclRecv per member)
  if (!s || !members) return MDC_ERR_INVALID;
  if (!s->comm || s->world < 2) return MDC_OK;
  const int gnx = s->cfg.gnx, gny = s->cfg.gny, nz = s->cfg.nz, W = s->world, me = s->rank;
  mdc_ctx* ctx = s->ctxs[0];
  const int64_t plane = (int64_t)gny * gnx;
  if (int rc = grow_dev(s, &s->xbuf, &s->xbuf_bytes, plane * nz * 8)) return rc;
  double* dev = static_cast<double*>(s->xbuf);
  void* stream = mdc_ctx_stream(ctx);
  for (int m = 0; m < s->cfg.k; ++m) {
    int y0, y1;
    slab_bounds(gny, me, W, y0, y1);
    for (int l = 0; l < nz; ++l)
      MDC_RT(s, ctx, mdc_dev_copy(ctx, dev + l * plane + (int64_t)y0 * gnx, members[m] + l * plane + (int64_t)y0 * gnx, (int64_t)(y1 - y0) * gnx * 8, 1));
    int rc = g_nccl.GroupStart();
    for (int r = 0; r < W && !rc; ++r) {
      if (r == me) continue;
      int a0, a1;
      slab_bounds(gny, r, W, a0, a1);
      for (int l = 0; l < nz && !rc; ++l) {
        rc = g_nccl.Send(dev + l * plane + (int64_t)y0 * gnx, (size_t)(y1 - y0) * gnx, kNcclFloat64, r, s->comm, stream);
        if (!rc) rc = g_nccl.Recv(dev + l * plane + (int64_t)a0 * gnx, (size_t)(a1 - a0) * gnx, kNcclFloat64, r, s->comm, stream);
      }
    }
    if (!rc) rc = g_nccl.GroupEnd();
    if (rc) return fail(s, MDC_ERR_CUDA, "NCCL row all-gather: %s", g_nccl.GetErrorString(rc));
    MDC_RT(s, ctx, mdc_ctx_sync(ctx));
    for (int r = 0; r < W; ++r) {
      if (r == me) continue;
      int a0, a1;
      slab_bounds(gny, r, W, a0, a1);
      for (int l = 0; l < nz; ++l)
        MDC_RT(s, ctx, mdc_dev_copy(ctx, members[m] + l * plane + (int64_t)a0 * gnx, dev + l * plane + (int64_t)a0 * gnx, (int64_t)(a1 - a0) * gnx * 8, 2));
    }
  }
  return MDC_OK;
}


// ---- sharded analysis of GEOGRAPHIC observations (one process per GPU; csrc/geo_api.inl for the pieces).  What
// metada_b200/parallel.py: GeoSlabLetkf does, from C++: every rank locates ALL observations on the global geography
// (ownership = the slab of the located row), applies H to its own, sends rank r the own observations inside the
// bounding box of r's columns widened by the reach of the radius (rows with latitude / longitude / level / variable,
// grouped ncclSend / ncclRecv; the row counts travel first -- the box test runs on the device, so they are not
// recomputed on the host), analyses its rows on its window of the geography (global frame: the lattice of the
// bucket index, hence the result, is the one-store run's) and writes them back to the host members.
int mdc_geo_sharded_analyse(mdc_stream* s, double* const* members, const double* glat, const double* glon, int nvc,
                            const double* vc, int nvar, const int32_t* var_nlev, int64_t P, const double* olat,
                            const double* olon, const double* olev, const double* oval, const double* oerr,
                            const uint8_t* ovalid, const int32_t* ovar, const mdc_letkf_params* prm, mdc_letkf_stats* st) {
  if (!s || !members || !glat || !glon || !prm || !st || P < 0) return MDC_ERR_INVALID;
  const int gnx = s->cfg.gnx, gny = s->cfg.gny, nz = s->cfg.nz, k = s->cfg.k, W = s->comm ? s->world : 1, me = s->comm ? s->rank : 0;
  mdc_ctx* ctx = s->ctxs[0];
  int y0, y1;
  slab_bounds(gny, me, W, y0, y1);
  if (y0 != s->cfg.row0 || y1 != s->cfg.row1) return fail(s, MDC_ERR_INVALID, "geo_sharded_analyse: the stream's row range is not this rank's slab");
  const int halo = y1 < gny ? 1 : 0;
  mdc_ens *whole = nullptr, *ens = nullptr;
  mdc_obs *all = nullptr, *own = nullptr;
  auto cleanup = [&]() {
    if (own) mdc_obs_destroy(own);
    if (all) mdc_obs_destroy(all);
    if (ens) mdc_ens_destroy(ens);
    if (whole) mdc_ens_destroy(whole);
  };
#define MDC_RTC(call)                                                                                         \
  do {                                                                                                        \
    int rc_ = (call);                                                                                         \
    if (rc_) { fail(s, rc_, "%s: %s", #call, mdc_last_error(ctx)); cleanup(); return rc_; }                   \
  } while (0)
  // the global geography (one level, one member: it only locates and hands out windows)
  MDC_RTC(mdc_ens_create(ctx, gnx, gny, 1, 1, &whole));
  MDC_RTC(mdc_ens_set_geography(whole, glat, glon, nvc, vc));
  std::vector<int32_t> oy((size_t)std::max<int64_t>(P, 1));
  MDC_RTC(mdc_obs_create_geographic(ctx, P, olat, olon, olev, oval, oerr, ovalid, nullptr, &all));
  MDC_RTC(mdc_obs_locate(all, whole));
  MDC_RTC(mdc_obs_download_grid_coords(all, nullptr, oy.data(), nullptr));
  mdc_obs_destroy(all);
  all = nullptr;
  std::vector<int64_t> idx;
  for (int64_t a = 0; a < P; ++a)
    if (W == 1 || owner_of_row(oy[(size_t)a], gny, W) == me) idx.push_back(a);
  const size_t n = idx.size();
  std::vector<double> la(n), lo(n), le(n), va(n), er(n);
  std::vector<uint8_t> ok(n);
  std::vector<int32_t> vr(n);
  for (size_t i = 0; i < n; ++i) {
    const int64_t a = idx[i];
    la[i] = olat[a]; lo[i] = olon[a]; le[i] = olev ? olev[a] : 0.0; va[i] = oval[a]; er[i] = oerr[a];
    ok[i] = ovalid ? ovalid[a] : 1; vr[i] = ovar ? ovar[a] : 0;
  }
  // this rank's slab (+ the halo row of the IDW stencil) on its window of the geography
  MDC_RTC(mdc_ens_create(ctx, gnx, (y1 - y0) + halo, nz, k, &ens));
  MDC_RTC(mdc_ens_set_domain(ens, 0, y0, gnx, gny, gnx, y1 - y0));
  MDC_RTC(mdc_ens_set_geography_from(ens, whole));
  if (nvar > 0) MDC_RTC(mdc_ens_set_variables(ens, nvar, var_nlev));
  MDC_RTC(mdc_ens_upload_members_rows(ens, 0, k, members, gny, y0));
  MDC_RTC(mdc_obs_create_geographic(ctx, (int64_t)n, la.data(), lo.data(), le.data(), va.data(), er.data(), ok.data(), idx.data(), &own));
  if (ovar) MDC_RTC(mdc_obs_set_variables(own, vr.data()));
  MDC_RTC(mdc_obs_locate(own, whole));
  MDC_RTC(mdc_hx_idw4(ens, own));
  s->last_halo_rows = 0;
  if (W > 1) {
    // every rank's box from the global arrays: bounding box of its columns in the geography's frame + the reach
    double lon_c, umin, umax, latmin, latmax;
    MDC_RTC(mdc_ens_geography_frame(whole, &lon_c, &umin, &umax, &latmin, &latmax));
    const double pi = 3.14159265358979323846, delta = std::max(prm->radius, 0.0) / 6371.0;
    const double phic = std::max(std::fabs(latmin), std::fabs(latmax)) * pi / 180.0;
    const double dlat = delta * 180.0 / pi * (1.0 + 1e-9) + 1e-12;
    const double dlon = std::asin(std::min(1.0, std::sin(delta) / std::cos(phic))) * 180.0 / pi * (1.0 + 1e-9) + 1e-12;
    std::vector<double> box((size_t)W * 4);
    for (int r = 0; r < W; ++r) {
      int a0, a1;
      slab_bounds(gny, r, W, a0, a1);
      double b0 = 1e300, b1 = -1e300, c0 = 1e300, c1 = -1e300;
      for (int64_t g = (int64_t)a0 * gnx; g < (int64_t)a1 * gnx; ++g) {
        const double u = (glon[g] - lon_c) - 360.0 * std::rint((glon[g] - lon_c) / 360.0);
        b0 = std::min(b0, glat[g]); b1 = std::max(b1, glat[g]); c0 = std::min(c0, u); c1 = std::max(c1, u);
      }
      box[4 * r] = b0 - dlat; box[4 * r + 1] = b1 + dlat; box[4 * r + 2] = c0 - dlon; box[4 * r + 3] = c1 + dlon;
    }
    const int rd = mdc_obs_row_doubles(own);
    std::vector<int64_t> nsend((size_t)W, 0), nrecv((size_t)W, 0);
    for (int r = 0; r < W; ++r)
      if (r != me) MDC_RTC(mdc_obs_pack_rows_geo(own, box[4 * r], box[4 * r + 1], box[4 * r + 2], box[4 * r + 3], lon_c, nullptr, 0, &nsend[(size_t)r]));
    // counts first (one double per peer), then the rows
    void* stream = mdc_ctx_stream(ctx);
    if (int rc = grow_dev(s, &s->xbuf, &s->xbuf_bytes, (int64_t)2 * W * 8)) { cleanup(); return rc; }
    {
      std::vector<double> c((size_t)2 * W, 0.0);
      for (int r = 0; r < W; ++r) c[(size_t)r] = (double)nsend[(size_t)r];
      MDC_RTC(mdc_dev_copy(ctx, s->xbuf, c.data(), (int64_t)2 * W * 8, 1));
      double* d = static_cast<double*>(s->xbuf);
      int rc = g_nccl.GroupStart();
      for (int r = 0; r < W && !rc; ++r) {
        if (r == me) continue;
        rc = g_nccl.Send(d + r, 1, kNcclFloat64, r, s->comm, stream);
        if (!rc) rc = g_nccl.Recv(d + W + r, 1, kNcclFloat64, r, s->comm, stream);
      }
      if (!rc) rc = g_nccl.GroupEnd();
      if (rc) { fail(s, MDC_ERR_CUDA, "NCCL count exchange: %s", g_nccl.GetErrorString(rc)); cleanup(); return MDC_ERR_CUDA; }
      MDC_RTC(mdc_ctx_sync(ctx));
      MDC_RTC(mdc_dev_copy(ctx, c.data(), s->xbuf, (int64_t)2 * W * 8, 2));
      for (int r = 0; r < W; ++r) nrecv[(size_t)r] = r == me ? 0 : (int64_t)c[(size_t)(W + r)];
    }
    int64_t tot = 0;
    for (int r = 0; r < W; ++r) tot += nsend[(size_t)r] + nrecv[(size_t)r];
    void* rows = nullptr;
    MDC_RTC(mdc_dev_malloc(ctx, std::max<int64_t>(tot, 1) * rd * 8, &rows));
    std::vector<double*> sp((size_t)W, nullptr), rp((size_t)W, nullptr);
    {
      double* p = static_cast<double*>(rows);
      for (int r = 0; r < W; ++r) { sp[(size_t)r] = p; p += nsend[(size_t)r] * rd; }
      for (int r = 0; r < W; ++r) { rp[(size_t)r] = p; p += nrecv[(size_t)r] * rd; }
    }
    int rc = 0;
    for (int r = 0; r < W && !rc; ++r) {
      if (r == me || nsend[(size_t)r] == 0) continue;
      int64_t got = 0;
      rc = mdc_obs_pack_rows_geo(own, box[4 * r], box[4 * r + 1], box[4 * r + 2], box[4 * r + 3], lon_c, sp[(size_t)r], nsend[(size_t)r], &got);
      if (!rc && got != nsend[(size_t)r]) rc = MDC_ERR_INVALID;
    }
    if (!rc) {
      rc = g_nccl.GroupStart() ? MDC_ERR_CUDA : 0;
      for (int r = 0; r < W && !rc; ++r) {
        if (r == me) continue;
        if (nsend[(size_t)r] > 0) rc = g_nccl.Send(sp[(size_t)r], (size_t)(nsend[(size_t)r] * rd), kNcclFloat64, r, s->comm, stream) ? MDC_ERR_CUDA : 0;
        if (!rc && nrecv[(size_t)r] > 0) rc = g_nccl.Recv(rp[(size_t)r], (size_t)(nrecv[(size_t)r] * rd), kNcclFloat64, r, s->comm, stream) ? MDC_ERR_CUDA : 0;
      }
      if (!rc) rc = g_nccl.GroupEnd() ? MDC_ERR_CUDA : 0;
      if (!rc) rc = mdc_ctx_sync(ctx);
    }
    for (int r = 0; r < W && !rc; ++r)
      if (r != me && nrecv[(size_t)r] > 0) { rc = mdc_obs_append_rows(own, rp[(size_t)r], nrecv[(size_t)r]); s->last_halo_rows += nrecv[(size_t)r]; }
    if (!rc) rc = mdc_ctx_sync(ctx);
    mdc_dev_free(ctx, rows);
    if (rc) { fail(s, rc, "geographic halo exchange: %s", mdc_last_error(ctx)); cleanup(); return rc; }
  }
  MDC_RTC(mdc_letkf_analyse(ens, own, prm, st));
  MDC_RTC(mdc_ens_download_members_rows(ens, 0, k, members, gny, y0, y1 - y0));
#undef MDC_RTC
  cleanup();
  return MDC_OK;
}

int mdc_stream_analyse(mdc_stream* s, double* const* members, int host_row0, int host_ny, int64_t P, const int32_t* ox,
                       const int32_t* oy, const int32_t* oz, const double* oval, const double* oerr, const uint8_t* ovalid,
                       const mdc_letkf_params* params, mdc_letkf_stats* out) {
  if (!s || !members || !params || !out || P < 0 || (P > 0 && (!ox || !oy || !oval || !oerr))) return MDC_ERR_INVALID;
  if ((int)std::floor(params->radius) > s->reach)
    return fail(s, MDC_ERR_INVALID, "stream was planned for radius < %d, analyse got %g: the observation halos would miss rows", s->reach + 1,
                params->radius);
  if (params->radius_v > 0.0 && params->mode != MDC_MODE_CANONICAL) return fail(s, MDC_ERR_UNSUPPORTED, "vertical localisation needs MDC_MODE_CANONICAL");
  if (host_row0 > s->cfg.row0 || host_row0 + host_ny < std::min(s->cfg.row1 + 1, s->cfg.gny))
    return fail(s, MDC_ERR_INVALID, "host members cover rows [%d, %d): rows [%d, %d] are needed (own rows + one halo row for H)", host_row0,
                host_row0 + host_ny, s->cfg.row0, std::min(s->cfg.row1, s->cfg.gny - 1));
  const ObsView o{P, ox, oy, oz, oval, oerr, ovalid};
  std::vector<std::pair<const double*, int64_t>> ext_top, ext_bottom;
  const auto t0 = std::chrono::steady_clock::now();
  s->last_halo_rows = 0;
  if (s->world > 1)
    if (int rc = exchange_halo(s, members, host_row0, host_ny, o, ext_top, ext_bottom)) return rc;
  const auto t1 = std::chrono::steady_clock::now();
  const int rc = run_pipeline(s, members, host_row0, host_ny, o, params, ext_top, ext_bottom, out);
  const auto t2 = std::chrono::steady_clock::now();
  s->last_edge_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  s->last_stream_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
  return rc;
}

}  // extern "C"
