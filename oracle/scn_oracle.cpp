// scn_oracle.cpp — CPU ORACLE. TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A line-faithful CPU restatement of the reference's search hot path, used only as the parity
// checker (tests/, __graft_entry__.smoke()) and as the timed CPU baseline (bench.py cpu_baseline /
// --impl reference). Nothing under scintirete_b200/ may import, link or call this file.
//
// Parity status: PINNED against every known-answer row of the reference's own tests for this
// path (internal/core/algorithm/distance_test.go, hnsw_test.go, hnsw_graph_state_test.go) — see
// tests/test_oracle_golden.py. The reference itself (Go) cannot be executed in this image (no Go
// toolchain, tree needs protoc/flatc codegen), so there is no oracle/_ref build; search results
// at scale are pinned only by this restatement.
//
// Follows (all paths relative to /root/reference):
//   internal/core/algorithm/distance.go  (whole file)
//   internal/core/algorithm/hnsw.go      17-104 node, 128-145 ctor, 148-257 build/insert,
//                                        260-289 delete, 292-350 search, 458-634 helpers,
//                                        669-699 sorts, 703-804 export/import
// Arithmetic contract: float accumulators, source order, separate multiply and add roundings
// (compile with -ffp-contract=off, no -ffast-math, no -march=native), sqrtf == Go's
// float32(math.Sqrt(float64(x))) (double rounding of sqrt is innocuous for 24->53 bits).
//
// Known, documented deviations (none affects search arithmetic):
//   * selectLayer's uniform variate: Go's math/rand (ALFG seeded through a 607-entry table that
//     is not in /root/reference) cannot be reproduced here; a splitmix64 stream seeded with
//     HNSWParams.Seed is used. Level *distribution* is identical, per-seed sequences are not.
//     Tests that need a fixed structure pass explicit levels (scn_oracle_hnsw_insert_level).
//   * findNewEntrypoint iterates a Go map (random order); here nodes are visited in insertion
//     order and the first node with the maximal layer wins.
//   * `visited` is an epoch-stamped array instead of a Go map (same set semantics).
//   * sortCandidates is a full insertion sort in the reference; both lists are already sorted
//     except for the one element just appended/replaced at the tail, so a single tail insertion
//     yields the identical permutation (stable: shifts only while prev.Distance > key.Distance).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

enum : int32_t { METRIC_UNSPEC = 0, METRIC_L2 = 1, METRIC_COS = 2, METRIC_IP = 3 };  // types.go:14-19

// ---- distance.go -------------------------------------------------------------------------

// distance.go:21-32
inline float l2_distance(const float* a, size_t na, const float* b, size_t nb) {
  if (na != nb) return std::numeric_limits<float>::infinity();
  float sum = 0.0f;
  for (size_t i = 0; i < na; ++i) {
    float diff = a[i] - b[i];
    sum += diff * diff;
  }
  return sqrtf(sum);
}

// distance.go:53-82
inline float cosine_distance(const float* a, size_t na, const float* b, size_t nb) {
  if (na != nb) return std::numeric_limits<float>::infinity();
  float dot = 0.0f, norm_a = 0.0f, norm_b = 0.0f;
  for (size_t i = 0; i < na; ++i) {
    dot += a[i] * b[i];
    norm_a += a[i] * a[i];
    norm_b += b[i] * b[i];
  }
  norm_a = sqrtf(norm_a);
  norm_b = sqrtf(norm_b);
  if (norm_a == 0.0f || norm_b == 0.0f) return 1.0f;
  float cs = dot / (norm_a * norm_b);
  if (cs > 1.0f) cs = 1.0f;
  else if (cs < -1.0f) cs = -1.0f;
  return 1.0f - cs;
}

// distance.go:104-116
inline float ip_distance(const float* a, size_t na, const float* b, size_t nb) {
  if (na != nb) return std::numeric_limits<float>::infinity();
  float dot = 0.0f;
  for (size_t i = 0; i < na; ++i) dot += a[i] * b[i];
  return -dot;
}

typedef float (*dist_fn)(const float*, size_t, const float*, size_t);

// distance.go:129-140
inline dist_fn calculator_for(int32_t metric) {
  switch (metric) {
    case METRIC_L2: return l2_distance;
    case METRIC_COS: return cosine_distance;
    case METRIC_IP: return ip_distance;
    default: return nullptr;  // ErrInvalidParameters("unsupported distance metric")
  }
}

// ---- hnsw.go -----------------------------------------------------------------------------

struct Candidate {  // hnsw.go:669-672
  uint64_t id;
  float dist;
};

struct Node {  // hnsw.go:17-26
  uint64_t id;
  std::vector<float> vec;
  bool deleted = false;
  std::vector<std::vector<uint64_t>> conn;  // per layer
};

struct SearchStats {
  uint64_t evals = 0;  // distCalc.Distance calls
  uint64_t hops = 0;   // expansions (pops that were not the terminating one)
};

// tail insertion == the reference's insertion sort on an array sorted except at the tail
inline void tail_insert(std::vector<Candidate>& v, size_t first) {  // hnsw.go:675-686
  if (v.size() - first < 2) return;
  Candidate key = v.back();
  size_t j = v.size() - 1;
  while (j > first && v[j - 1].dist > key.dist) {
    v[j] = v[j - 1];
    --j;
  }
  v[j] = key;
}

struct Hnsw {
  int M, efC, efS, max_layers;
  int64_t seed;
  int32_t metric;
  dist_fn dist;
  std::vector<Node> nodes;  // insertion order
  std::unordered_map<uint64_t, uint32_t> index_of;
  uint64_t entrypoint = 0;  // 0 == none (hnsw.go:210, 296)
  int max_layer = -1;
  int size = 0;
  uint64_t rng_state;

  // scratch for the single-threaded mutators
  std::vector<uint32_t> visit_epoch;
  uint32_t epoch = 0;

  double next_float64() {  // stand-in for rng.Float64() (see header)
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= (z >> 31);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
  }

  int select_layer() {  // hnsw.go:458-469
    double mL = 1.0 / std::log(2.0);
    double u = next_float64();
    if (u <= 0.0) u = 1.0 / 9007199254740992.0;  // reference would produce +Inf -> UB; clamp
    int level = (int)std::floor(-std::log(u) * mL);
    if (level >= max_layers) level = max_layers - 1;
    return level;
  }

  Node* find(uint64_t id) {
    auto it = index_of.find(id);
    return it == index_of.end() ? nullptr : &nodes[it->second];
  }
  const Node* find(uint64_t id) const {
    auto it = index_of.find(id);
    return it == index_of.end() ? nullptr : &nodes[it->second];
  }

  int node_layer(uint64_t id) const {  // hnsw.go:472-484
    const Node* n = find(id);
    if (!n) return -1;
    for (int i = (int)n->conn.size() - 1; i >= 0; --i)
      if (!n->conn[i].empty()) return i;
    return 0;
  }

  // hnsw.go:487-557. `epochs`/`ep` provide the visited set for this call.
  void search_layer(const float* q, size_t dim, const std::vector<uint64_t>& entry, int ef, int layer,
                    std::vector<uint32_t>& epochs, uint32_t& ep, std::vector<uint64_t>& out,
                    SearchStats* st) const {
    out.clear();
    if (epochs.size() < nodes.size()) epochs.resize(nodes.size(), 0);
    if (++ep == 0) {  // wrapped: reset stamps
      std::fill(epochs.begin(), epochs.end(), 0);
      ep = 1;
    }
    std::vector<Candidate> W;  // `candidates`
    for (uint64_t e : entry) {  // 492-498
      auto it = index_of.find(e);
      if (it == index_of.end()) continue;
      const Node& n = nodes[it->second];
      if (n.deleted) continue;
      float d = dist(q, dim, n.vec.data(), n.vec.size());
      if (st) st->evals++;
      W.push_back({e, d});
      epochs[it->second] = ep;
    }
    if (W.empty()) return;  // 500-502
    std::stable_sort(W.begin(), W.end(),
                     [](const Candidate& a, const Candidate& b) { return a.dist < b.dist; });  // 505
    std::vector<Candidate> C(W);  // `dynamic`, 507-508
    size_t c_first = 0;
    while (c_first < C.size()) {  // 510
      Candidate cur = C[c_first++];  // 512-513
      if ((int)W.size() >= ef && cur.dist > W[ef - 1].dist) break;  // 516-518
      if (st) st->hops++;
      const Node& node = nodes[index_of.find(cur.id)->second];
      if (layer >= (int)node.conn.size()) continue;  // GetConnections 45-50
      const std::vector<uint64_t>& nbrs = node.conn[layer];
      for (uint64_t nb : nbrs) {  // 522
        uint32_t nbi = index_of.find(nb)->second;
        if (epochs[nbi] == ep) continue;  // 523-525
        const Node& nn = nodes[nbi];
        if (nn.deleted) continue;  // 527-530: not marked visited, not traversed
        epochs[nbi] = ep;          // 532
        float d = dist(q, dim, nn.vec.data(), nn.vec.size());  // 533
        if (st) st->evals++;
        if ((int)W.size() < ef) {  // 536-538
          W.push_back({nb, d});
          C.push_back({nb, d});
          tail_insert(W, 0);  // 545
          tail_insert(C, c_first);  // 546
        } else if (d < W[ef - 1].dist) {  // 539-542
          W[ef - 1] = {nb, d};
          C.push_back({nb, d});
          // W may be longer than ef only when more than ef entry points were given; the
          // reference sorts the whole slice, so re-sort the whole slice here as well.
          if ((int)W.size() == ef) tail_insert(W, 0);
          else std::stable_sort(W.begin(), W.end(), [](const Candidate& a, const Candidate& b) {
                 return a.dist < b.dist; });
          tail_insert(C, c_first);
        }
      }
    }
    size_t n = std::min<size_t>((size_t)ef, W.size());  // 551-556
    out.resize(n);
    for (size_t i = 0; i < n; ++i) out[i] = W[i].id;
  }

  // hnsw.go:560-583
  std::vector<uint64_t> select_neighbors(const float* q, size_t dim, const std::vector<uint64_t>& cand,
                                         int max_conn) const {
    if ((int)cand.size() <= max_conn) return cand;
    std::vector<Candidate> items(cand.size());
    for (size_t i = 0; i < cand.size(); ++i) {
      const Node* n = find(cand[i]);
      items[i] = {cand[i], dist(q, dim, n->vec.data(), n->vec.size())};
    }
    std::stable_sort(items.begin(), items.end(),
                     [](const Candidate& a, const Candidate& b) { return a.dist < b.dist; });
    std::vector<uint64_t> r(max_conn);
    for (int i = 0; i < max_conn; ++i) r[i] = items[i].id;
    return r;
  }

  static void add_connection(Node& n, int layer, uint64_t id) {  // hnsw.go:78-89
    if (layer < (int)n.conn.size()) {
      for (uint64_t c : n.conn[layer])
        if (c == id) return;
      n.conn[layer].push_back(id);
    }
  }

  void prune_connections(Node& node, int layer) {  // hnsw.go:586-614
    int max_conn = (layer == 0) ? M * 2 : M;
    if (layer >= (int)node.conn.size()) return;
    std::vector<uint64_t>& conns = node.conn[layer];
    if ((int)conns.size() <= max_conn) return;
    std::vector<Candidate> cand;
    cand.reserve(conns.size());
    for (uint64_t cid : conns) {
      const Node* cn = find(cid);
      if (cn && !cn->deleted)
        cand.push_back({cid, dist(node.vec.data(), node.vec.size(), cn->vec.data(), cn->vec.size())});
    }
    std::stable_sort(cand.begin(), cand.end(),
                     [](const Candidate& a, const Candidate& b) { return a.dist < b.dist; });
    std::vector<uint64_t> kept;
    int keep = std::min<int>(max_conn, (int)cand.size());
    for (int i = 0; i < keep; ++i) {
      bool dup = false;
      for (uint64_t k : kept) dup |= (k == cand[i].id);
      if (!dup) kept.push_back(cand[i].id);
    }
    conns.swap(kept);
  }

  // hnsw.go:190-257. level < 0 -> draw from select_layer().
  int insert_vector(uint64_t id, const float* vec, size_t dim, int level) {
    if (index_of.count(id)) return 3007;  // ErrInvalidParameters: already exists (192-194)
    int layer = level >= 0 ? std::min(level, max_layers - 1) : select_layer();  // 197
    Node n;
    n.id = id;
    n.vec.assign(vec, vec + dim);
    n.conn.resize(layer + 1);  // NewHNSWNode(..., layer+1) 200
    nodes.push_back(std::move(n));
    uint32_t self = (uint32_t)nodes.size() - 1;
    index_of[id] = self;
    size++;
    if (layer > max_layer) max_layer = layer;  // 205-207 (before the descent)
    if (entrypoint == 0) {                      // 210-213
      entrypoint = id;
      return 0;
    }
    std::vector<uint64_t> eps{entrypoint}, tmp;
    const float* v = nodes[self].vec.data();
    for (int lc = max_layer; lc > layer; --lc) {  // 219-221
      search_layer(v, dim, eps, 1, lc, visit_epoch, epoch, tmp, nullptr);
      eps = tmp;
    }
    for (int lc = std::min(layer, max_layer); lc >= 0; --lc) {  // 224
      search_layer(v, dim, eps, efC, lc, visit_epoch, epoch, tmp, nullptr);  // 225
      int max_conn = (lc == 0) ? M * 2 : M;                                   // 228-231
      std::vector<uint64_t> sel = select_neighbors(v, dim, tmp, max_conn);   // 233
      for (uint64_t nb : sel) {                                              // 236-246
        add_connection(nodes[self], lc, nb);
        auto it = index_of.find(nb);
        if (it != index_of.end()) {
          add_connection(nodes[it->second], lc, id);
          prune_connections(nodes[it->second], lc);
        }
      }
      eps = sel;  // 248
    }
    if (layer > node_layer(entrypoint)) entrypoint = id;  // 252-254
    return 0;
  }

  void find_new_entrypoint() {  // hnsw.go:617-634 (insertion order instead of Go map order)
    entrypoint = 0;
    int best = -1;
    for (const Node& n : nodes) {
      if (n.deleted) continue;
      int l = node_layer(n.id);
      if (l > best) {
        best = l;
        entrypoint = n.id;
      }
    }
    max_layer = best;
  }

  int remove(uint64_t id) {  // hnsw.go:260-289
    Node* n = find(id);
    if (!n) return 3004;  // ErrVectorNotFound
    if (n->deleted) return 0;
    n->deleted = true;
    size--;
    if (entrypoint == id) find_new_entrypoint();
    return 0;
  }

  // hnsw.go:292-350
  int search(const float* q, size_t dim, int topk, int ef_opt, uint64_t* out_ids, float* out_dist,
             std::vector<uint32_t>& epochs, uint32_t& ep, SearchStats* st) const {
    if (entrypoint == 0 || size == 0) return 0;  // 296-298
    int ef = efS;
    if (ef_opt > 0) ef = ef_opt;  // 300-303
    std::vector<uint64_t> eps{entrypoint}, tmp;
    for (int lc = max_layer; lc > 0; --lc) {  // 309-311
      search_layer(q, dim, eps, 1, lc, epochs, ep, tmp, st);
      eps = tmp;
    }
    search_layer(q, dim, eps, ef, 0, epochs, ep, tmp, st);  // 314
    std::vector<Candidate> res;
    for (uint64_t cid : tmp) {  // 319-339
      if ((int)res.size() >= topk) break;
      const Node* n = find(cid);
      if (n->deleted) continue;
      res.push_back({cid, dist(q, dim, n->vec.data(), n->vec.size())});
      if (st) st->evals++;
    }
    std::stable_sort(res.begin(), res.end(),
                     [](const Candidate& a, const Candidate& b) { return a.dist < b.dist; });  // 342
    if ((int)res.size() > topk) res.resize(topk);                                             // 345-347
    for (size_t i = 0; i < res.size(); ++i) {
      out_ids[i] = res[i].id;
      out_dist[i] = res[i].dist;
    }
    return (int)res.size();
  }
};

template <class F>
void parallel_for(uint64_t n, int nthreads, F f) {
  if (nthreads <= 1 || n <= 1) {
    for (uint64_t i = 0; i < n; ++i) f(i, 0);
    return;
  }
  std::atomic<uint64_t> next{0};
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&, t] {
      for (;;) {
        uint64_t i = next.fetch_add(1);
        if (i >= n) break;
        f(i, t);
      }
    });
  for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

// ---- scalar distance API (distance.go) ---------------------------------------------------

// Returns 0 and writes *out, or 3007 (ErrorCodeInvalidParameters) for an unsupported metric.
int32_t scn_oracle_distance(int32_t metric, const float* a, uint64_t na, const float* b, uint64_t nb,
                            float* out) {
  dist_fn f = calculator_for(metric);
  if (!f) return 3007;
  *out = f(a, na, b, nb);
  return 0;
}

// distance.go:144-150 — one query against n row-major targets of the same dimension.
int32_t scn_oracle_batch_distance(int32_t metric, const float* q, uint64_t dim, const float* targets,
                                  uint64_t n, float* out) {
  dist_fn f = calculator_for(metric);
  if (!f) return 3007;
  for (uint64_t i = 0; i < n; ++i) out[i] = f(q, dim, targets + i * dim, dim);
  return 0;
}

// distance.go:154-172. Returns 1 if `out` holds a normalised copy, 0 if the zero vector was
// returned unchanged (out == copy of in).
int32_t scn_oracle_normalize(const float* v, uint64_t n, float* out) {
  float norm = 0.0f;
  for (uint64_t i = 0; i < n; ++i) norm += v[i] * v[i];
  norm = sqrtf(norm);
  if (norm == 0.0f) {
    std::memcpy(out, v, n * sizeof(float));
    return 0;
  }
  for (uint64_t i = 0; i < n; ++i) out[i] = v[i] / norm;
  return 1;
}

// distance.go:175-181
float scn_oracle_magnitude(const float* v, uint64_t n) {
  float sum = 0.0f;
  for (uint64_t i = 0; i < n; ++i) sum += v[i] * v[i];
  return sqrtf(sum);
}

// distance.go:184-192
float scn_oracle_dot(const float* a, uint64_t na, const float* b, uint64_t nb) {
  if (na != nb) return 0.0f;
  float p = 0.0f;
  for (uint64_t i = 0; i < na; ++i) p += a[i] * b[i];
  return p;
}

// ---- flat (exact) scan: BatchDistance + stable ascending sort, ties -> lower row ------------
// The reference has no flat scan; SURVEY.md §8c defines it as BatchDistance (distance.go:144-150)
// over rows in ID order followed by a stable ascending sort and truncation to k. `ids` may be
// NULL (id = row + 1, collection.go:57,115-116); `deleted` may be NULL (byte per row).
// Missing results are padded with id 0 / +Inf. counts[q] receives the number of valid results.
int32_t scn_oracle_flat_search(int32_t metric, const float* db, uint64_t n, uint64_t dim,
                               const uint64_t* ids, const uint8_t* deleted, const float* q, uint64_t nq,
                               uint32_t k, uint64_t* out_ids, float* out_dist, uint32_t* counts,
                               int32_t nthreads) {
  dist_fn f = calculator_for(metric);
  if (!f) return 3007;
  parallel_for(nq, nthreads, [&](uint64_t qi, int) {
    const float* qv = q + qi * dim;
    std::vector<std::pair<float, uint64_t>> top;  // (dist,row), sorted ascending, size <= k
    top.reserve(k + 1);
    for (uint64_t r = 0; r < n; ++r) {
      if (deleted && deleted[r]) continue;
      float d = f(qv, dim, db + r * dim, dim);
      // stable ascending order over rows: a later row only enters if strictly smaller than the
      // current k-th (equal distance keeps the earlier row first). NaN never enters a full list.
      if (top.size() < k) {
        size_t j = top.size();
        top.emplace_back(d, r);
        while (j > 0 && top[j - 1].first > d) {
          top[j] = top[j - 1];
          --j;
        }
        top[j] = {d, r};
      } else if (k > 0 && d < top[k - 1].first) {
        size_t j = k - 1;
        while (j > 0 && top[j - 1].first > d) {
          top[j] = top[j - 1];
          --j;
        }
        top[j] = {d, r};
      }
    }
    for (uint32_t i = 0; i < k; ++i) {
      if (i < top.size()) {
        out_ids[qi * k + i] = ids ? ids[top[i].second] : top[i].second + 1;
        out_dist[qi * k + i] = top[i].first;
      } else {
        out_ids[qi * k + i] = 0;
        out_dist[qi * k + i] = std::numeric_limits<float>::infinity();
      }
    }
    if (counts) counts[qi] = (uint32_t)top.size();
  });
  return 0;
}

// ---- HNSW (hnsw.go) -----------------------------------------------------------------------

void* scn_oracle_hnsw_new(int32_t M, int32_t ef_construction, int32_t ef_search, int32_t max_layers,
                          int64_t seed, int32_t metric) {  // hnsw.go:128-145
  dist_fn f = calculator_for(metric);
  if (!f) return nullptr;
  Hnsw* h = new Hnsw();
  h->M = M;
  h->efC = ef_construction;
  h->efS = ef_search;
  h->max_layers = max_layers;
  h->seed = seed;
  h->metric = metric;
  h->dist = f;
  h->rng_state = (uint64_t)seed;
  return h;
}

void scn_oracle_hnsw_free(void* p) { delete (Hnsw*)p; }

int32_t scn_oracle_hnsw_insert(void* p, uint64_t id, const float* vec, uint64_t dim) {
  int r = ((Hnsw*)p)->insert_vector(id, vec, dim, -1);
  return r ? 5002 : 0;  // Insert wraps the cause in ErrInsertFailed (hnsw.go:181-183)
}

int32_t scn_oracle_hnsw_insert_level(void* p, uint64_t id, const float* vec, uint64_t dim, int32_t level) {
  int r = ((Hnsw*)p)->insert_vector(id, vec, dim, level);
  return r ? 5002 : 0;
}

// hnsw.go:148-174 — clear, then serial insert in slice order. ids may be NULL (1..n).
int32_t scn_oracle_hnsw_build(void* p, const uint64_t* ids, const float* vecs, uint64_t n, uint64_t dim) {
  Hnsw* h = (Hnsw*)p;
  h->nodes.clear();
  h->index_of.clear();
  h->entrypoint = 0;
  h->max_layer = -1;
  h->size = 0;
  h->nodes.reserve(n);
  h->index_of.reserve(n * 2);
  for (uint64_t i = 0; i < n; ++i)
    if (h->insert_vector(ids ? ids[i] : i + 1, vecs + i * dim, dim, -1)) return 5000;  // ErrIndexBuildFailed
  return 0;
}

int32_t scn_oracle_hnsw_delete(void* p, uint64_t id) { return ((Hnsw*)p)->remove(id); }

void scn_oracle_hnsw_set_ef_search(void* p, int32_t ef) { ((Hnsw*)p)->efS = ef; }  // hnsw.go:449-453
int32_t scn_oracle_hnsw_size(void* p) { return ((Hnsw*)p)->size; }
int32_t scn_oracle_hnsw_layers(void* p) {  // hnsw.go:394-401
  Hnsw* h = (Hnsw*)p;
  return h->max_layer < 0 ? 0 : h->max_layer + 1;
}
uint64_t scn_oracle_hnsw_entrypoint(void* p) { return ((Hnsw*)p)->entrypoint; }
int32_t scn_oracle_hnsw_max_layer(void* p) { return ((Hnsw*)p)->max_layer; }
uint64_t scn_oracle_hnsw_node_count(void* p) { return ((Hnsw*)p)->nodes.size(); }

// One query. ef <= 0 -> index default. Returns the number of results (<= topk).
int32_t scn_oracle_hnsw_search(void* p, const float* q, uint64_t dim, int32_t topk, int32_t ef,
                               uint64_t* out_ids, float* out_dist, uint64_t* stats /*[2] evals,hops or NULL*/) {
  Hnsw* h = (Hnsw*)p;
  std::vector<uint32_t> epochs;
  uint32_t ep = 0;
  SearchStats st;
  int n = h->search(q, dim, topk, ef, out_ids, out_dist, epochs, ep, &st);
  if (stats) {
    stats[0] = st.evals;
    stats[1] = st.hops;
  }
  return n;
}

// nq independent Search calls, one per worker thread at a time (the reference's
// goroutine-per-request model under RLock, hnsw.go:293). Outputs padded with id 0 / +Inf.
int32_t scn_oracle_hnsw_search_batch(void* p, const float* q, uint64_t nq, uint64_t dim, int32_t topk,
                                     int32_t ef, uint64_t* out_ids, float* out_dist, uint32_t* counts,
                                     int32_t nthreads, uint64_t* stats /*[2] totals or NULL*/) {
  Hnsw* h = (Hnsw*)p;
  int nt = std::max(1, nthreads);
  std::vector<std::vector<uint32_t>> epochs(nt);
  std::vector<uint32_t> eps(nt, 0);
  std::vector<SearchStats> sts(nt);
  parallel_for(nq, nt, [&](uint64_t qi, int t) {
    std::vector<uint64_t> ids(topk);
    std::vector<float> ds(topk);
    int n = h->search(q + qi * dim, dim, topk, ef, ids.data(), ds.data(), epochs[t], eps[t], &sts[t]);
    for (int i = 0; i < topk; ++i) {
      out_ids[qi * topk + i] = i < n ? ids[i] : 0;
      out_dist[qi * topk + i] = i < n ? ds[i] : std::numeric_limits<float>::infinity();
    }
    if (counts) counts[qi] = (uint32_t)n;
  });
  if (stats) {
    stats[0] = stats[1] = 0;
    for (auto& s : sts) {
      stats[0] += s.evals;
      stats[1] += s.hops;
    }
  }
  return 0;
}

// searchLayer exposed for unit tests (hnsw.go:487-557). Returns count written to out (<= ef).
int32_t scn_oracle_hnsw_search_layer(void* p, const float* q, uint64_t dim, const uint64_t* entry,
                                     uint32_t n_entry, int32_t ef, int32_t layer, uint64_t* out) {
  Hnsw* h = (Hnsw*)p;
  std::vector<uint32_t> epochs;
  uint32_t ep = 0;
  std::vector<uint64_t> e(entry, entry + n_entry), r;
  h->search_layer(q, dim, e, ef, layer, epochs, ep, r, nullptr);
  std::copy(r.begin(), r.end(), out);
  return (int32_t)r.size();
}

// ---- graph export / import (hnsw.go:703-804), flattened ------------------------------------
// Node order = insertion order. list_counts[i] = len(node.Connections) (== level+1).
// Edge layout: for node i, for layer l in [0, list_counts[i]): edge_counts[off_i + l] edges,
// concatenated in `edges` (neighbour IDs, in stored order).
void scn_oracle_hnsw_export_sizes(void* p, uint64_t* n_nodes, uint64_t* n_lists, uint64_t* n_edges) {
  Hnsw* h = (Hnsw*)p;
  uint64_t nl = 0, ne = 0;
  for (const Node& n : h->nodes) {
    nl += n.conn.size();
    for (auto& c : n.conn) ne += c.size();
  }
  *n_nodes = h->nodes.size();
  *n_lists = nl;
  *n_edges = ne;
}

void scn_oracle_hnsw_export(void* p, uint64_t* ids, uint8_t* deleted, int32_t* list_counts,
                            uint32_t* edge_counts, uint64_t* edges, float* vectors /*may be NULL*/) {
  Hnsw* h = (Hnsw*)p;
  uint64_t li = 0, ei = 0, i = 0;
  for (const Node& n : h->nodes) {
    ids[i] = n.id;
    deleted[i] = n.deleted ? 1 : 0;
    list_counts[i] = (int32_t)n.conn.size();
    if (vectors) std::memcpy(vectors + i * n.vec.size(), n.vec.data(), n.vec.size() * sizeof(float));
    for (auto& c : n.conn) {
      edge_counts[li++] = (uint32_t)c.size();
      for (uint64_t e : c) edges[ei++] = e;
    }
    ++i;
  }
}

// ImportGraphState: replaces all nodes, then sets entrypoint / maxLayer / size verbatim (791-793).
int32_t scn_oracle_hnsw_import(void* p, uint64_t n_nodes, uint64_t dim, const uint64_t* ids,
                               const uint8_t* deleted, const int32_t* list_counts,
                               const uint32_t* edge_counts, const uint64_t* edges, const float* vectors,
                               uint64_t entrypoint, int32_t max_layer, int32_t size) {
  Hnsw* h = (Hnsw*)p;
  h->nodes.clear();
  h->index_of.clear();
  h->nodes.reserve(n_nodes);
  uint64_t li = 0, ei = 0;
  for (uint64_t i = 0; i < n_nodes; ++i) {
    Node n;
    n.id = ids[i];
    n.deleted = deleted[i] != 0;
    n.vec.assign(vectors + i * dim, vectors + (i + 1) * dim);
    n.conn.resize(list_counts[i]);
    for (int l = 0; l < list_counts[i]; ++l) {
      uint32_t c = edge_counts[li++];
      n.conn[l].assign(edges + ei, edges + ei + c);
      ei += c;
    }
    h->index_of[n.id] = (uint32_t)h->nodes.size();
    h->nodes.push_back(std::move(n));
  }
  h->entrypoint = entrypoint;
  h->max_layer = max_layer;
  h->size = size;
  return 0;
}

}  // extern "C"
