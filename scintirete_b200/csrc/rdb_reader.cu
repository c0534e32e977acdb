// rdb_reader.cu — RDB snapshot -> device store, without the host detour.
//
// The reference restores a collection by parsing the FlatBuffers snapshot into Go structs
// (rdb.go:179-237, 661-741: one heap slice per vector, ids and edges as decimal strings, one float
// accessor call per element), converting them to core.HNSWGraphState (rdb.go:1027-1091) and
// importing that into the index (database.go:398-493 -> hnsw.go:749-804). scn_store_load_rdb reads
// the same file (schemas/flatbuffers/rdb.fbs, written by rdb.go:239-533) and goes straight to the
// flat device layout: vectors are copied out of the FlatBuffers `elements` arrays into row-major
// staging blocks, ids/edges are parsed once, and the graph is handed to scn_graph_upload.
//
// The FlatBuffers wire format is read directly (no flatc here): a table starts with an int32 offset
// back to its vtable {u16 vtable bytes, u16 table bytes, u16 field offsets...}; field i lives at
// table + vtable[2 + i] (0 = absent); strings/vectors/sub-tables are u32 offsets relative to the
// field's own position; vectors and strings start with a u32 length. Every access is bounds
// checked: a damaged file yields ErrorCodeCorruptedData (4002), never a wild read.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "store.h"

using namespace scn;

namespace {

constexpr int32_t ERR_RECOVERY = 4001;   // utils.ErrorCodeRecoveryFailed
constexpr int32_t ERR_CORRUPT = 4002;    // utils.ErrorCodeCorruptedData
constexpr int32_t ERR_DB_NOT_FOUND = 3000;
constexpr int32_t ERR_COLL_NOT_FOUND = 3002;

struct Buf {
  const uint8_t* p = nullptr;
  size_t n = 0;
  bool ok = true;  // cleared by any out-of-range access

  bool in(size_t off, size_t len) const { return off <= n && len <= n - off; }
  template <class T>
  T rd(size_t off) {
    T v{};
    if (!in(off, sizeof(T))) {
      ok = false;
      return v;
    }
    std::memcpy(&v, p + off, sizeof(T));
    return v;
  }
};

struct Table {
  Buf* b = nullptr;
  size_t pos = 0;      // table start
  size_t vt = 0;       // vtable start
  uint16_t vt_len = 0;
  bool present = false;

  // byte position of field `id`, 0 if absent
  size_t field(int id) const {
    if (!present) return 0;
    const size_t slot = 4 + 2 * (size_t)id;
    if (slot + 2 > vt_len) return 0;
    const uint16_t off = b->rd<uint16_t>(vt + slot);
    return off ? pos + off : 0;
  }
  template <class T>
  T scalar(int id, T dflt) const {
    const size_t f = field(id);
    return f ? b->rd<T>(f) : dflt;
  }
  // position a u32 offset field points at (0 if absent)
  size_t indirect(int id) const {
    const size_t f = field(id);
    if (!f) return 0;
    const uint32_t o = b->rd<uint32_t>(f);
    const size_t t = f + o;
    if (o == 0 || !b->in(t, 4)) {
      b->ok = false;
      return 0;
    }
    return t;
  }
};

Table table_at(Buf* b, size_t pos) {
  Table t;
  t.b = b;
  if (!pos || !b->in(pos, 4)) {
    b->ok = false;
    return t;
  }
  const int32_t so = b->rd<int32_t>(pos);
  const int64_t vt = (int64_t)pos - so;
  if (vt < 0 || !b->in((size_t)vt, 4)) {
    b->ok = false;
    return t;
  }
  t.pos = pos;
  t.vt = (size_t)vt;
  t.vt_len = b->rd<uint16_t>(t.vt);
  if (t.vt_len < 4 || !b->in(t.vt, t.vt_len)) {
    b->ok = false;
    return t;
  }
  t.present = true;
  return t;
}

Table sub_table(const Table& t, int id) {
  const size_t p = t.indirect(id);
  if (!p) return Table{t.b};
  return table_at(t.b, p);
}

struct Vec {
  size_t data = 0;   // first element
  uint32_t len = 0;
};

Vec vector_of(const Table& t, int id, size_t elem_size) {
  Vec v;
  const size_t p = t.indirect(id);
  if (!p) return v;
  v.len = t.b->rd<uint32_t>(p);
  v.data = p + 4;
  if (!t.b->in(v.data, (size_t)v.len * elem_size)) {
    t.b->ok = false;
    v.len = 0;
  }
  return v;
}

// element i of a vector of offsets (tables or strings) -> target position
size_t vec_target(Buf* b, const Vec& v, uint32_t i) {
  const size_t f = v.data + (size_t)i * 4;
  const uint32_t o = b->rd<uint32_t>(f);
  const size_t t = f + o;
  if (o == 0 || !b->in(t, 4)) {
    b->ok = false;
    return 0;
  }
  return t;
}

struct Str {
  const char* s = "";
  uint32_t len = 0;
};

Str string_at(Buf* b, size_t pos) {
  Str r;
  if (!pos) return r;
  const uint32_t len = b->rd<uint32_t>(pos);
  if (!b->in(pos + 4, len)) {
    b->ok = false;
    return r;
  }
  r.s = reinterpret_cast<const char*>(b->p + pos + 4);
  r.len = len;
  return r;
}

bool str_eq(const Str& a, const char* z) { return std::strlen(z) == a.len && std::memcmp(a.s, z, a.len) == 0; }

// strconv.ParseUint(s, 10, 64): digits only, no sign, no blanks, no overflow
bool parse_u64(const Str& a, uint64_t* out) {
  if (a.len == 0 || a.len > 20) return false;
  uint64_t v = 0;
  for (uint32_t i = 0; i < a.len; ++i) {
    const char c = a.s[i];
    if (c < '0' || c > '9') return false;
    const uint64_t d = (uint64_t)(c - '0');
    if (v > (UINT64_MAX - d) / 10) return false;
    v = v * 10 + d;
  }
  *out = v;
  return true;
}

// field ids = declaration order in schemas/flatbuffers/rdb.fbs
enum { RDB_VERSION = 0, RDB_TIMESTAMP = 1, RDB_DATABASES = 2 };
enum { DB_NAME = 0, DB_COLLECTIONS = 1 };
enum { COLL_NAME = 0, COLL_CONFIG = 1, COLL_VECTORS = 2, COLL_GRAPH = 3, COLL_VECTOR_COUNT = 4, COLL_DELETED_COUNT = 5 };
enum { CFG_NAME = 0, CFG_METRIC = 1, CFG_HNSW = 2 };
enum { HP_M = 0, HP_EFC = 1, HP_EFS = 2, HP_MAX_LAYERS = 3, HP_SEED = 4 };
enum { G_NODES = 0, G_ENTRY = 1, G_MAX_LAYER = 2, G_SIZE = 3 };
enum { N_ID = 0, N_ELEMENTS = 1, N_METADATA = 2, N_DELETED = 3, N_LAYERS = 4, N_MAX_LAYER = 5 };
enum { LC_LAYER = 0, LC_IDS = 1 };

}  // namespace

extern "C" {

int32_t scn_store_load_rdb(const char* path, const char* database, const char* collection, int32_t device,
                           scn_store** out, scn_rdb_info* info) {
  if (!out) return fail(SCN_ERR_INVALID_PARAMETERS, "out is NULL");
  *out = nullptr;
  if (!path || !database || !collection) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  // ---- read the whole file (rdb.go:191-195 does the same) ----
  std::FILE* f = std::fopen(path, "rb");
  if (!f) return fail(ERR_RECOVERY, "failed to read RDB file %s: %s", path, std::strerror(errno));
  std::vector<uint8_t> data;
  {
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (sz < 8) {
      std::fclose(f);
      return fail(ERR_CORRUPT, "RDB file %s is too short to be a snapshot", path);
    }
    data.resize((size_t)sz);
    const size_t got = std::fread(data.data(), 1, data.size(), f);
    std::fclose(f);
    if (got != data.size()) return fail(ERR_RECOVERY, "failed to read RDB file %s", path);
  }
  Buf b;
  b.p = data.data();
  b.n = data.size();
  const Table root = table_at(&b, b.rd<uint32_t>(0));
  if (!b.ok) return fail(ERR_CORRUPT, "RDB file %s: bad root table", path);

  // ---- locate database / collection by name ----
  Table coll{&b};
  {
    const Vec dbs = vector_of(root, RDB_DATABASES, 4);
    bool db_found = false;
    for (uint32_t i = 0; i < dbs.len && b.ok && !coll.present; ++i) {
      const Table db = table_at(&b, vec_target(&b, dbs, i));
      if (!b.ok || !str_eq(string_at(&b, db.indirect(DB_NAME)), database)) continue;
      db_found = true;
      const Vec cs = vector_of(db, DB_COLLECTIONS, 4);
      for (uint32_t j = 0; j < cs.len && b.ok; ++j) {
        const Table c = table_at(&b, vec_target(&b, cs, j));
        if (b.ok && str_eq(string_at(&b, c.indirect(COLL_NAME)), collection)) {
          coll = c;
          break;
        }
      }
    }
    if (!b.ok) return fail(ERR_CORRUPT, "RDB file %s: damaged database / collection tables", path);
    if (!db_found) return fail(ERR_DB_NOT_FOUND, "database %s not found in %s", database, path);
    if (!coll.present) return fail(ERR_COLL_NOT_FOUND, "collection %s not found in database %s", collection, database);
  }

  // ---- configuration (rdb.go:635-659) ----
  scn_rdb_info inf;
  std::memset(&inf, 0, sizeof inf);
  const Table cfg = sub_table(coll, COLL_CONFIG);
  inf.metric = cfg.scalar<int8_t>(CFG_METRIC, 0);
  const Table hp = sub_table(cfg, CFG_HNSW);
  inf.m = hp.scalar<int32_t>(HP_M, 0);
  inf.ef_construction = hp.scalar<int32_t>(HP_EFC, 0);
  inf.ef_search = hp.scalar<int32_t>(HP_EFS, 0);
  inf.max_layers = hp.scalar<int32_t>(HP_MAX_LAYERS, 0);
  inf.seed = hp.scalar<int64_t>(HP_SEED, 0);
  inf.vector_count = coll.scalar<int64_t>(COLL_VECTOR_COUNT, 0);
  inf.deleted_count = coll.scalar<int64_t>(COLL_DELETED_COUNT, 0);
  if (!b.ok) return fail(ERR_CORRUPT, "RDB file %s: damaged collection config", path);

  // ---- graph: restore refuses a snapshot without it (database.go:461-464) ----
  const Table graph = sub_table(coll, COLL_GRAPH);
  if (!b.ok) return fail(ERR_CORRUPT, "RDB file %s: damaged HNSW graph", path);
  if (!graph.present)
    return fail(ERR_RECOVERY, "HNSW graph state missing in RDB for collection %s - cannot restore without graph data", collection);
  uint64_t entry = 0;
  if (!parse_u64(string_at(&b, graph.indirect(G_ENTRY)), &entry))
    return fail(ERR_CORRUPT, "failed to parse entry point ID");  // rdb.go:1078-1081
  inf.entry_id = entry;
  inf.max_layer = graph.scalar<int32_t>(G_MAX_LAYER, 0);
  inf.graph_size = graph.scalar<int32_t>(G_SIZE, 0);
  const Vec nodes = vector_of(graph, G_NODES, 4);
  inf.nodes = nodes.len;
  inf.has_graph = 1;

  // ---- first pass: ids, dimension, list shapes ----
  const uint64_t n = nodes.len;
  std::vector<uint64_t> ids(n);
  std::vector<int32_t> list_counts(n);
  std::vector<size_t> node_pos(n);
  std::vector<uint64_t> dead;
  uint32_t dim = 0;
  for (uint64_t i = 0; i < n; ++i) {
    const Table nd = table_at(&b, vec_target(&b, nodes, (uint32_t)i));
    if (!b.ok) return fail(ERR_CORRUPT, "failed to parse HNSW node");  // rdb.go:670-673
    node_pos[i] = nd.pos;
    if (!parse_u64(string_at(&b, nd.indirect(N_ID)), &ids[i])) return fail(ERR_CORRUPT, "failed to parse node ID");  // rdb.go:1038-1041
    const Vec el = vector_of(nd, N_ELEMENTS, 4);
    if (i == 0) dim = el.len;
    if (el.len != dim || dim == 0)
      return fail(SCN_ERR_DIMENSION_MISMATCH, "node %llu has dimension %u, expected %u", (unsigned long long)ids[i], el.len, dim);
    const int32_t ml = nd.scalar<int32_t>(N_MAX_LAYER, 0);
    if (ml < 0 || ml > 255) return fail(ERR_CORRUPT, "node %llu: max_layer %d out of range", (unsigned long long)ids[i], ml);
    list_counts[i] = ml + 1;  // rdb.go:1044: len(connections) = MaxLayer + 1
    if (nd.scalar<uint8_t>(N_DELETED, 0)) dead.push_back(ids[i]);
    if (!b.ok) return fail(ERR_CORRUPT, "failed to parse HNSW node");
  }
  inf.dim = dim;
  inf.deleted = dead.size();
  if (info) *info = inf;

  // ---- store + vectors, in file order, through row-major staging blocks ----
  scn_store* s = nullptr;
  if (n == 0) {
    return fail(ERR_RECOVERY, "collection %s holds no nodes: its dimension is unknown, nothing to restore", collection);
  }
  SCN_TRY(scn_store_create(device, dim, inf.metric, &s));
  auto bail = [&](int32_t rc) {
    const std::string msg = scn_last_error();
    scn_store_destroy(s);
    return fail(rc, "%s", msg.c_str());
  };
  int32_t rc = scn_store_reserve(s, n);
  if (rc != SCN_OK) return bail(rc);
  {
    const uint64_t blk = std::max<uint64_t>(1, std::min<uint64_t>(n, (64ull << 20) / ((uint64_t)dim * 4)));
    std::vector<float> stage(blk * dim);
    for (uint64_t r0 = 0; r0 < n; r0 += blk) {
      const uint64_t m = std::min(blk, n - r0);
      for (uint64_t i = 0; i < m; ++i) {
        const Table nd = table_at(&b, node_pos[r0 + i]);
        const Vec el = vector_of(nd, N_ELEMENTS, 4);
        std::memcpy(stage.data() + i * dim, b.p + el.data, (size_t)dim * 4);  // little-endian float32, as stored
      }
      rc = scn_store_append(s, stage.data(), ids.data() + r0, m);
      if (rc != SCN_OK) return bail(rc);
    }
  }

  // ---- edges (rdb.go:1050-1061): lists beyond max_layer are dropped, unparsable ids skipped ----
  std::vector<uint32_t> edge_counts;
  std::vector<uint64_t> edges;
  {
    std::vector<std::vector<uint64_t>> lists;
    for (uint64_t i = 0; i < n; ++i) {
      const Table nd = table_at(&b, node_pos[i]);
      const int32_t nl = list_counts[i];
      lists.assign((size_t)nl, {});
      const Vec lcs = vector_of(nd, N_LAYERS, 4);
      for (uint32_t j = 0; j < lcs.len && b.ok; ++j) {
        const Table lc = table_at(&b, vec_target(&b, lcs, j));
        if (!b.ok) break;
        const int32_t layer = lc.scalar<int32_t>(LC_LAYER, 0);
        if (layer < 0 || layer >= nl) continue;
        const Vec cid = vector_of(lc, LC_IDS, 4);
        for (uint32_t e = 0; e < cid.len && b.ok; ++e) {
          uint64_t v;
          if (parse_u64(string_at(&b, vec_target(&b, cid, e)), &v)) lists[(size_t)layer].push_back(v);
        }
      }
      if (!b.ok) {
        scn_store_destroy(s);
        return fail(ERR_CORRUPT, "failed to parse layer connections");  // rdb.go:713-716
      }
      for (auto& l : lists) {
        edge_counts.push_back((uint32_t)l.size());
        edges.insert(edges.end(), l.begin(), l.end());
      }
    }
  }
  const int32_t m_param = inf.m > 0 ? inf.m : 16;
  rc = scn_graph_upload(s, m_param, inf.max_layer, entry, n, ids.data(), list_counts.data(), edge_counts.data(), edges.data());
  if (rc != SCN_OK) return bail(rc);
  if (!dead.empty()) {
    rc = scn_store_restore_deleted(s, dead.data(), dead.size());  // entry point and maxLayer stay verbatim (hnsw.go:791-793)
    if (rc != SCN_OK) return bail(rc);
  }
  *out = s;
  return SCN_OK;
}

}  // extern "C"
