// shards.cu — one collection row-sharded over the GPUs of one box, behind ONE host-buffer call.
//
// This is the multi-device form of the drop-in boundary ("scn_store_create(const
// int* devices, int ndev, ...)"; §8e): the reference server is a single process whose
// core.VectorIndex.Search (internal/core/interfaces.go:87-111) is one blocking call, so the
// replacement for it drives all G GPUs from inside that call:
//
//   scn_shards_search_flat(sh, q, nq, k, out_ids, out_dist, out_counts)
//
// One worker thread per device (bound to it for good, own stream) runs that device's rank of the
// fused exchange of exchange.cu: it copies ITS 1/G slice of the caller's query buffer to its GPU,
// the slices are gathered over NVLink, every GPU scans its rows for the whole batch, the top-k
// lists travel to the owner of each query slice as P2P stores, and each worker writes the merged
// results of its slice straight into the caller's output buffers. Results are bit-identical to
// the single-GPU scn_search_flat over the same rows.
//
// Row placement: shard g owns the global rows [g*per, (g+1)*per), per = ceil(capacity_rows / G),
// filled in insertion order; rows beyond the declared capacity go to the last shard. Contiguous
// blocks are what make `row_base + local row` a global row, i.e. what keeps the merge order equal
// to the flat oracle's (distance, insertion row) order. Auto-assigned ids are global row + 1
// (collection.go:57,115-116).
#include <condition_variable>
#include <cstring>
#include <functional>
#include <thread>

#include "store.h"

using namespace scn;

extern "C" int32_t scn_exchange_create(int32_t, uint32_t, uint32_t, uint64_t, uint32_t, uint32_t, scn_exchange**);

namespace {

// A worker thread pinned to one device; jobs are run in submission order.
struct Worker {
  int device = 0;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv, cv_done;
  std::function<int32_t()> job;
  bool has_job = false, stop = false, done = false;
  int32_t rc = SCN_OK;
  std::string err;

  void start(int dev) {
    device = dev;
    th = std::thread([this] {
      cudaSetDevice(device);
      std::unique_lock<std::mutex> lk(mu);
      for (;;) {
        cv.wait(lk, [this] { return has_job || stop; });
        if (stop) return;
        std::function<int32_t()> j = std::move(job);
        has_job = false;
        lk.unlock();
        const int32_t r = j();
        std::string e = (r != SCN_OK) ? std::string(scn_last_error()) : std::string();
        lk.lock();
        rc = r;
        err = std::move(e);
        done = true;
        cv_done.notify_all();
      }
    });
  }
  void submit(std::function<int32_t()> j) {
    std::lock_guard<std::mutex> lk(mu);
    job = std::move(j);
    has_job = true;
    done = false;
    cv.notify_all();
  }
  int32_t wait(std::string* e) {
    std::unique_lock<std::mutex> lk(mu);
    cv_done.wait(lk, [this] { return done; });
    if (rc != SCN_OK && e) *e = err;
    return rc;
  }
  void shutdown() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
      cv.notify_all();
    }
    if (th.joinable()) th.join();
  }
};

}  // namespace

struct scn_shards {
  uint32_t world = 0, dim = 0;
  int32_t metric = 0;
  uint64_t capacity = 0, per = 0;   // rows per shard
  uint64_t rows = 0;                // global rows appended so far
  bool auto_ids = true;
  std::vector<int32_t> devices;
  std::vector<scn_store*> stores;
  std::vector<Worker*> workers;
  // exchanges are sized for (max_nq, k); rebuilt when a call needs more
  std::vector<scn_exchange*> ex;
  uint64_t ex_max_nq = 0;
  uint32_t ex_k = 0;
  std::mutex mu;  // one sharded search at a time (a collective over all devices)
  bool shared_device = false;  // two shards on one GPU (tests): the exchange runs in host-synchronised steps

  uint64_t row_base(uint32_t g) const { return (uint64_t)g * per; }
};

namespace {

// run fn(g) on every worker, return the first failure (message re-raised on the calling thread)
int32_t run_all(scn_shards* sh, const std::function<int32_t(uint32_t)>& fn) {
  for (uint32_t g = 0; g < sh->world; ++g) sh->workers[g]->submit([g, &fn] { return fn(g); });
  int32_t rc = SCN_OK;
  std::string err;
  for (uint32_t g = 0; g < sh->world; ++g) {
    std::string e;
    const int32_t r = sh->workers[g]->wait(&e);
    if (r != SCN_OK && rc == SCN_OK) {
      rc = r;
      err = e;
    }
  }
  if (rc != SCN_OK) return fail(rc, "%s", err.c_str());
  return SCN_OK;
}

void drop_exchanges(scn_shards* sh) {
  for (scn_exchange* e : sh->ex) scn_exchange_destroy(e);
  sh->ex.clear();
  sh->ex_max_nq = 0;
  sh->ex_k = 0;
}

int32_t ensure_exchanges(scn_shards* sh, uint64_t nq, uint32_t k) {
  if (!sh->ex.empty() && sh->ex_k == k && sh->ex_max_nq >= nq) return SCN_OK;
  drop_exchanges(sh);
  const uint64_t max_nq = std::max<uint64_t>(1024, next_pow2((uint32_t)std::min<uint64_t>(nq, 1u << 30)));
  sh->ex.assign(sh->world, nullptr);
  for (uint32_t g = 0; g < sh->world; ++g) {
    const int32_t rc = scn_exchange_create(sh->devices[g], g, sh->world, max_nq, k, sh->dim, &sh->ex[g]);
    if (rc != SCN_OK) {
      drop_exchanges(sh);
      return rc;
    }
  }
  for (uint32_t g = 0; g < sh->world; ++g) {
    const int32_t rc = scn_exchange_connect_local(sh->ex[g], sh->ex.data());
    if (rc != SCN_OK) {
      drop_exchanges(sh);
      return rc;
    }
  }
  sh->ex_max_nq = max_nq;
  sh->ex_k = k;
  return SCN_OK;
}

}  // namespace

extern "C" {

int32_t scn_shards_create(const int32_t* devices, int32_t ndev, uint32_t dim, int32_t metric, uint64_t capacity_rows,
                          scn_shards** out) {
  if (!out) return fail(SCN_ERR_INVALID_PARAMETERS, "out is NULL");
  *out = nullptr;
  if (!devices || ndev < 1 || ndev > 16) return fail(SCN_ERR_INVALID_PARAMETERS, "between 1 and 16 devices are required");
  scn_shards* sh = new scn_shards();
  sh->world = (uint32_t)ndev;
  sh->dim = dim;
  sh->metric = metric;
  sh->capacity = std::max<uint64_t>(capacity_rows, 1);
  sh->per = (sh->capacity + sh->world - 1) / sh->world;
  sh->devices.assign(devices, devices + ndev);
  for (int a = 0; a < ndev; ++a)
    for (int c = a + 1; c < ndev; ++c) sh->shared_device |= devices[a] == devices[c];
  for (int g = 0; g < ndev; ++g) {
    scn_store* s = nullptr;
    const int32_t rc = scn_store_create(devices[g], dim, metric, &s);
    if (rc != SCN_OK) {
      for (scn_store* t : sh->stores) scn_store_destroy(t);
      delete sh;
      return rc;
    }
    s->auto_base = sh->row_base((uint32_t)g);
    sh->stores.push_back(s);
  }
  for (int g = 0; g < ndev; ++g) {
    Worker* w = new Worker();
    w->start(devices[g]);
    sh->workers.push_back(w);
  }
  *out = sh;
  return SCN_OK;
}

int32_t scn_shards_destroy(scn_shards* sh) {
  if (!sh) return SCN_OK;
  for (Worker* w : sh->workers) {
    w->shutdown();
    delete w;
  }
  drop_exchanges(sh);
  for (scn_store* s : sh->stores) scn_store_destroy(s);
  delete sh;
  return SCN_OK;
}

int32_t scn_shards_count(scn_shards* sh) { return sh ? (int32_t)sh->world : 0; }

scn_store* scn_shards_store(scn_shards* sh, int32_t i) {
  if (!sh || i < 0 || (uint32_t)i >= sh->world) return nullptr;
  return sh->stores[i];
}

// Appends n vectors in insertion order: global row = rows appended so far. ids == NULL assigns
// global row + 1.
int32_t scn_shards_append(scn_shards* sh, const float* vecs, const uint64_t* ids, uint64_t n) {
  if (!sh) return fail(SCN_ERR_INVALID_PARAMETERS, "shard set is NULL");
  if (n == 0) return SCN_OK;
  if (!vecs) return fail(SCN_ERR_INVALID_PARAMETERS, "vectors pointer is NULL");
  std::lock_guard<std::mutex> lk(sh->mu);
  // cut [rows, rows + n) at the shard boundaries
  struct Part {
    uint32_t g;
    uint64_t off, cnt;
  };
  std::vector<Part> parts;
  uint64_t r = sh->rows, left = n, off = 0;
  while (left) {
    const uint32_t g = (uint32_t)std::min<uint64_t>(r / sh->per, sh->world - 1);
    const uint64_t end = (g == sh->world - 1) ? ~0ull : sh->row_base(g + 1);
    const uint64_t cnt = std::min<uint64_t>(left, end - r);
    parts.push_back({g, off, cnt});
    r += cnt;
    off += cnt;
    left -= cnt;
  }
  // a shard's local row must equal global row - row_base: shards fill strictly in order
  for (const Part& p : parts)
    if (sh->stores[p.g]->rows != (sh->rows + p.off) - sh->row_base(p.g))
      return fail(SCN_ERR_INSERT_FAILED, "shard %u is out of step with the global row counter", p.g);
  std::vector<int32_t> rcs(parts.size(), SCN_OK);
  const int32_t rc = run_all(sh, [&](uint32_t g) -> int32_t {
    for (const Part& p : parts)
      if (p.g == g) {
        const int32_t r2 = scn_store_append(sh->stores[g], vecs + p.off * sh->dim, ids ? ids + p.off : nullptr, p.cnt);
        if (r2 != SCN_OK) return r2;
      }
    return SCN_OK;
  });
  // (a failed part leaves the earlier shards of this call appended: report, the caller rebuilds)
  SCN_TRY(rc);
  sh->rows += n;
  return SCN_OK;
}

int32_t scn_shards_mark_deleted(scn_shards* sh, const uint64_t* ids, uint64_t n) {
  if (!sh) return fail(SCN_ERR_INVALID_PARAMETERS, "shard set is NULL");
  if (n && !ids) return fail(SCN_ERR_INVALID_PARAMETERS, "ids pointer is NULL");
  std::lock_guard<std::mutex> lk(sh->mu);
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t row;
    uint32_t owner = sh->world;
    for (uint32_t g = 0; g < sh->world && owner == sh->world; ++g)
      if (sh->stores[g]->lookup(ids[i], &row)) owner = g;
    if (owner == sh->world) return fail(SCN_ERR_VECTOR_NOT_FOUND, "vector %llu not found", (unsigned long long)ids[i]);
    const uint64_t id = ids[i];
    SCN_TRY(run_all(sh, [&](uint32_t g) -> int32_t { return g == owner ? scn_store_mark_deleted(sh->stores[g], &id, 1) : SCN_OK; }));
  }
  return SCN_OK;
}

int32_t scn_shards_stats(scn_shards* sh, scn_stats* out) {
  if (!sh || !out) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  std::memset(out, 0, sizeof *out);
  for (scn_store* s : sh->stores) {
    scn_stats t;
    SCN_TRY(scn_store_stats(s, &t));
    out->rows += t.rows;
    out->live_rows += t.live_rows;
    out->capacity_rows += t.capacity_rows;
    out->device_bytes += t.device_bytes;
  }
  out->dim = sh->dim;
  out->metric = sh->metric;
  out->device = -1;
  out->max_layer = -1;
  return SCN_OK;
}

int32_t scn_shards_set_option(scn_shards* sh, const char* name, int64_t value) {
  if (!sh) return fail(SCN_ERR_INVALID_PARAMETERS, "shard set is NULL");
  for (scn_store* s : sh->stores) SCN_TRY(scn_set_option(s, name, value));
  return SCN_OK;
}

int32_t scn_shards_search_flat(scn_shards* sh, const float* q, uint64_t nq, uint32_t k, uint64_t* out_ids, float* out_dist,
                               uint32_t* out_counts) {
  if (!sh) return fail(SCN_ERR_INVALID_PARAMETERS, "shard set is NULL");
  if (nq && !q) return fail(SCN_ERR_INVALID_PARAMETERS, "query pointer is NULL");
  if (k == 0) return fail(SCN_ERR_INVALID_PARAMETERS, "top_k must be positive");
  if (k > 1024) return fail(SCN_ERR_INVALID_PARAMETERS, "top_k above 1024 is not supported");
  if (nq && (!out_ids || !out_dist)) return fail(SCN_ERR_INVALID_PARAMETERS, "output pointer is NULL");
  if (nq >= (1ull << 31)) return fail(SCN_ERR_INVALID_PARAMETERS, "too many queries in one call");
  if (nq == 0) return SCN_OK;
  std::lock_guard<std::mutex> lk(sh->mu);
  SCN_TRY(ensure_exchanges(sh, nq, k));
  if (!sh->shared_device)
    return run_all(sh, [&](uint32_t g) -> int32_t {
      uint64_t lo = 0, cnt = 0;
      SCN_TRY(scn_exchange_slice(sh->ex[g], nq, g, &lo, &cnt));
      return scn_search_flat_exchange(sh->stores[g], sh->ex[g], q + lo * sh->dim, nq, k, sh->row_base(g), out_ids + lo * k,
                                      out_dist + lo * k, out_counts ? out_counts + lo : nullptr);
    });
  // Shards that share a GPU: a kernel must never spin on a flag that another launch on the same GPU
  // is to raise (nothing guarantees that the two run at the same time). Each of the three steps ends
  // with a stream synchronisation and a host barrier, so every wait kernel finds its flags raised.
  std::vector<HostExchangeCall*> calls(sh->world, nullptr);
  int32_t rc = run_all(sh, [&](uint32_t g) -> int32_t {
    uint64_t lo = 0, cnt = 0;
    SCN_TRY(scn_exchange_slice(sh->ex[g], nq, g, &lo, &cnt));
    int32_t r = SCN_OK;
    calls[g] = host_exchange_begin(sh->stores[g], sh->ex[g], q + lo * sh->dim, nq, k, sh->row_base(g), &r);
    return r != SCN_OK ? r : host_exchange_sync(calls[g]);
  });
  if (rc == SCN_OK)
    rc = run_all(sh, [&](uint32_t g) -> int32_t {
      SCN_TRY(host_exchange_search(calls[g]));
      return host_exchange_sync(calls[g]);
    });
  if (rc != SCN_OK) {
    std::string msg = scn_last_error();
    run_all(sh, [&](uint32_t g) -> int32_t {
      host_exchange_abort(calls[g]);
      return SCN_OK;
    });
    drop_exchanges(sh);   // the ranks are out of step: start over with fresh buffers next time
    return fail(rc, "%s", msg.c_str());
  }
  return run_all(sh, [&](uint32_t g) -> int32_t {
    uint64_t lo = 0, cnt = 0;
    SCN_TRY(scn_exchange_slice(sh->ex[g], nq, g, &lo, &cnt));
    return host_exchange_finish(calls[g], out_ids + lo * k, out_dist + lo * k, out_counts ? out_counts + lo : nullptr);
  });
}

}  // extern "C"
