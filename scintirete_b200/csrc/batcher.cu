// batcher.cu — micro-batching at the API edge.
//
// The reference's search API is one query per call: Collection.Search (collection.go:193-204) ->
// HNSW.Search (hnsw.go:292-350), one goroutine per request under an RLock
// (vector_operations_test.go:345-362 drives it with concurrent callers). A GPU answers one query
// in about the time it answers a thousand, so the drop-in for that call site coalesces the calls
// that are in flight at the same moment into one batched launch:
//
//   scn_batcher_search(b, q, k, ef, ...)    blocking, one query, callable from any thread
//
// Leader / follower, no background thread (nothing to start or stop from Go): the first caller to
// find no collector active becomes the leader of the next batch. It waits until the batch is full,
// or until `window_us` has passed AND fewer than two batches are executing (so under load batches
// grow to whatever arrived while the GPU was busy: natural batching), then takes every pending
// request, runs ONE batched search per distinct (k, ef) through the ordinary host-buffer entry
// points, scatters the results into the callers' buffers and wakes them. Results are exactly those
// of the batched entry points, i.e. identical to the reference's per-query Search.
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <map>

#include "store.h"

using namespace scn;

namespace {

struct Request {
  const float* q;
  uint32_t k, ef;
  uint64_t* out_ids;
  float* out_dist;
  uint32_t* out_count;
  int32_t rc = SCN_OK;
  std::string err;
  bool done = false;
};

}  // namespace

struct scn_batcher {
  scn_store* s = nullptr;
  int32_t kind = 0;  // 0 = exact flat scan, 1 = HNSW
  uint32_t max_batch = 1024;
  uint32_t window_us = 100;

  std::mutex mu;
  std::condition_variable cv_leader;  // batch full / a running batch finished
  std::condition_variable cv_done;    // results delivered
  std::vector<Request*> pending;
  bool collecting = false;
  int running = 0;

  // statistics
  uint64_t n_calls = 0, n_batches = 0, n_launches = 0, max_seen = 0;
};

static void run_batch(scn_batcher* b, std::vector<Request*>& batch) {
  scn_store* s = b->s;
  const uint32_t dim = s->dim;
  // one launch per distinct (k, ef); callers of one collection normally all use the same pair
  std::map<std::pair<uint32_t, uint32_t>, std::vector<Request*>> groups;
  for (Request* r : batch) groups[{r->k, r->ef}].push_back(r);
  std::vector<float> q;
  std::vector<uint64_t> ids;
  std::vector<float> dist;
  std::vector<uint32_t> counts;
  for (auto& g : groups) {
    const uint32_t k = g.first.first, ef = g.first.second;
    std::vector<Request*>& rs = g.second;
    const size_t n = rs.size();
    q.resize(n * dim);
    ids.resize(n * k);
    dist.resize(n * k);
    counts.resize(n);
    for (size_t i = 0; i < n; ++i) std::memcpy(q.data() + i * dim, rs[i]->q, dim * sizeof(float));
    int32_t rc = b->kind == 1 ? scn_search_hnsw(s, q.data(), n, k, ef, ids.data(), dist.data(), counts.data())
                              : scn_search_flat(s, q.data(), n, k, ids.data(), dist.data(), counts.data());
    const std::string err = rc == SCN_OK ? std::string() : std::string(scn_last_error());
    for (size_t i = 0; i < n; ++i) {
      Request* r = rs[i];
      r->rc = rc;
      if (rc == SCN_OK) {
        std::memcpy(r->out_ids, ids.data() + i * k, k * sizeof(uint64_t));
        std::memcpy(r->out_dist, dist.data() + i * k, k * sizeof(float));
        if (r->out_count) *r->out_count = counts[i];
      } else {
        r->err = err;
      }
    }
  }
  std::lock_guard<std::mutex> lk(b->mu);
  b->n_launches += groups.size();
}

extern "C" {

int32_t scn_batcher_create(scn_store* s, int32_t kind, uint32_t max_batch, uint32_t window_us, scn_batcher** out) {
  if (!out) return fail(SCN_ERR_INVALID_PARAMETERS, "out is NULL");
  *out = nullptr;
  if (!s) return fail(SCN_ERR_INVALID_PARAMETERS, "store is NULL");
  if (kind != 0 && kind != 1) return fail(SCN_ERR_INVALID_PARAMETERS, "batcher kind must be 0 (flat) or 1 (hnsw)");
  if (max_batch == 0 || max_batch > (1u << 20)) return fail(SCN_ERR_INVALID_PARAMETERS, "max_batch must be in [1, 2^20]");
  if (window_us > 1000000) return fail(SCN_ERR_INVALID_PARAMETERS, "window above one second");
  scn_batcher* b = new scn_batcher();
  b->s = s;
  b->kind = kind;
  b->max_batch = max_batch;
  b->window_us = window_us;
  *out = b;
  return SCN_OK;
}

int32_t scn_batcher_destroy(scn_batcher* b) {
  if (!b) return SCN_OK;
  {
    std::unique_lock<std::mutex> lk(b->mu);
    if (!b->pending.empty() || b->collecting || b->running)
      return fail(SCN_ERR_INVALID_PARAMETERS, "batcher destroyed while searches are in flight");
  }
  delete b;
  return SCN_OK;
}

int32_t scn_batcher_search(scn_batcher* b, const float* q, uint32_t k, uint32_t ef, uint64_t* out_ids, float* out_dist,
                           uint32_t* out_count) {
  if (!b) return fail(SCN_ERR_INVALID_PARAMETERS, "batcher is NULL");
  if (!q || !out_ids || !out_dist) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  if (k == 0 || k > 1024) return fail(SCN_ERR_INVALID_PARAMETERS, "top_k must be in [1, 1024]");
  if (b->kind == 1 && (ef == 0 || ef > 4096)) return fail(SCN_ERR_INVALID_PARAMETERS, "ef_search must be in [1, 4096]");
  Request r;
  r.q = q;
  r.k = k;
  r.ef = b->kind == 1 ? ef : 0;
  r.out_ids = out_ids;
  r.out_dist = out_dist;
  r.out_count = out_count;

  std::unique_lock<std::mutex> lk(b->mu);
  b->pending.push_back(&r);
  b->n_calls++;
  if (b->pending.size() >= b->max_batch) b->cv_leader.notify_all();
  if (!b->collecting) {
    // ---- leader of the next batch ----
    b->collecting = true;
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(b->window_us);
    for (;;) {
      if (b->pending.size() >= b->max_batch) break;
      const bool window_over = std::chrono::steady_clock::now() >= deadline;
      if (window_over && b->running < 2) break;
      if (window_over) b->cv_leader.wait(lk);  // wait for a running batch to finish (or a full batch)
      else b->cv_leader.wait_until(lk, deadline);
    }
    std::vector<Request*> batch;
    batch.swap(b->pending);
    b->collecting = false;  // the next arrival starts collecting the following batch
    b->running++;
    b->n_batches++;
    b->max_seen = std::max<uint64_t>(b->max_seen, batch.size());
    lk.unlock();
    run_batch(b, batch);
    lk.lock();
    b->running--;
    for (Request* x : batch) x->done = true;
    b->cv_done.notify_all();
    b->cv_leader.notify_all();
  } else {
    b->cv_done.wait(lk, [&] { return r.done; });
  }
  lk.unlock();
  if (r.rc != SCN_OK) return fail(r.rc, "%s", r.err.c_str());
  return SCN_OK;
}

int32_t scn_batcher_stats(scn_batcher* b, uint64_t* out, int32_t n) {
  if (!b || !out) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  std::lock_guard<std::mutex> lk(b->mu);
  const uint64_t v[4] = {b->n_calls, b->n_batches, b->n_launches, b->max_seen};
  for (int32_t i = 0; i < n && i < 4; ++i) out[i] = v[i];
  return SCN_OK;
}

}  // extern "C"
