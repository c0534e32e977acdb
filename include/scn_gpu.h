/*
 * scn_gpu.h — C ABI of libscn_gpu.so, the B200 (sm_100a) device-side replacement for the search
 * hot path of Scintirete's internal/core/algorithm package.
 *
 * This is the drop-in boundary: the entry points below are exactly what a cgo binding of the
 * reference's index/distance interfaces would call (see INTEGRATION.md for the Go side). Plain
 * pointers and sizes only; the caller owns every buffer; the library never keeps a caller pointer
 * past return. All citations are relative to the reference tree.
 *
 *   reference interface                                   replaced / mirrored by
 *   ----------------------------------------------------  ---------------------------------------
 *   algorithm.NewHNSW / NewDistanceCalculator              scn_store_create           (hnsw.go:128-145, distance.go:129-140)
 *   HNSW.Insert / Build (vector storage half)              scn_store_append[_dev]     (hnsw.go:148-187, collection.go:71-149)
 *   HNSW.Delete (soft delete flag)                         scn_store_mark_deleted     (hnsw.go:260-289)
 *   HNSW.ImportGraphState / ExportGraphState hand-off      scn_graph_upload           (hnsw.go:703-804, interfaces.go:137-151)
 *   HNSW.Search + searchLayer + rerank                     scn_search_hnsw[_dev]      (hnsw.go:292-350, 487-557)
 *   BatchDistance + stable sort (flat / exact scan)        scn_search_flat[_dev]      (distance.go:144-150)
 *   HNSW.Search rerank step on caller-chosen candidates    scn_rerank                 (hnsw.go:317-347)
 *   DistanceCalculator.Distance / BatchDistance            scn_distance_batch         (distance.go:21-32, 53-82, 104-116, 144-150)
 *   HNSW.Insert / Build (graph construction half)          scn_hnsw_insert            (hnsw.go:190-257, 560-614)
 *   HNSW.Size / MemoryUsage / GetStatistics                scn_store_stats            (hnsw.go:375-443)
 *   per-shard top-k merge (new: row-sharded multi-GPU)     scn_merge_topk_dev, scn_search_flat_exchange[_dev]
 *   VectorIndex.Search over a collection sharded over G GPUs scn_shards_search_flat     (interfaces.go:87-111)
 *   RDBManager.Load + RestoreFromSnapshot + ImportGraph    scn_store_load_rdb         (rdb.go:179-237, database.go:398-493)
 *   Collection.Compact (drop deleted, rebuild)             scn_store_compact          (collection.go:283-313)
 *   Collection.Search called by many goroutines            scn_batcher_search         (collection.go:193-204)
 *
 * Status codes are the reference's utils.ErrorCode numbers (internal/utils/errors.go:11-49) so the
 * Go shim can wrap them with utils.NewError(code, scn_last_error()).
 *
 * Threading: search entry points are re-entrant and may be called concurrently from any OS thread
 * (the reference allows concurrent Search under RLock, hnsw.go:293). Mutators (append, mark_deleted,
 * graph_upload, clear) require the caller's write lock (hnsw.go:178, 261, 750), exactly as in the
 * reference. All calls block until outputs are valid, except the *_dev variants, which enqueue on
 * the given CUDA stream and return.
 */
#ifndef SCN_GPU_H_
#define SCN_GPU_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define SCN_API __attribute__((visibility("default")))
#else
#define SCN_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* types.DistanceMetric (pkg/types/types.go:12-19); values cross the ABI unchanged. */
enum {
  SCN_METRIC_UNSPECIFIED = 0,
  SCN_METRIC_L2 = 1,
  SCN_METRIC_COSINE = 2,
  SCN_METRIC_INNER_PRODUCT = 3
};

/* utils.ErrorCode subset (internal/utils/errors.go:11-49). 0 = success. */
enum {
  SCN_OK = 0,
  SCN_ERR_INTERNAL = 1000,           /* CUDA / driver failure */
  SCN_ERR_RESOURCE = 1003,           /* out of device memory */
  SCN_ERR_VECTOR_NOT_FOUND = 3004,
  SCN_ERR_DIMENSION_MISMATCH = 3005,
  SCN_ERR_INVALID_PARAMETERS = 3007,
  SCN_ERR_INDEX_BUILD_FAILED = 5000,
  SCN_ERR_SEARCH_FAILED = 5001,
  SCN_ERR_INSERT_FAILED = 5002
};

typedef struct scn_store scn_store;

typedef struct scn_stats {
  uint64_t rows;          /* rows appended (including soft-deleted) */
  uint64_t live_rows;     /* rows not soft-deleted == HNSW.Size() */
  uint64_t capacity_rows;
  uint64_t device_bytes;  /* bytes of HBM held by the store */
  uint32_t dim;
  int32_t metric;
  int32_t device;
  int32_t has_graph;
  int32_t max_layer;      /* HNSW maxLayer, -1 if no graph */
  int32_t m;              /* HNSW M of the uploaded graph */
  uint64_t entry_id;      /* HNSW entrypoint, 0 = none */
  uint64_t graph_edges;
} scn_stats;

/* Thread-local message for the last non-zero status returned on this thread. */
SCN_API const char* scn_last_error(void);

/* Number of kernels launched by this library in this process so far (all threads). */
SCN_API uint64_t scn_launch_count(void);

/* Pinned (page-locked) host memory for query / result buffers. The host-buffer entry points accept
 * any host pointer: pinned buffers are DMA-ed directly, pageable ones (a Go slice) go through the
 * library's own pinned staging chunks, which costs one extra memcpy. */
SCN_API int32_t scn_host_alloc(uint64_t bytes, void** out);
SCN_API int32_t scn_host_free(void* p);

/* ---- device vector store ------------------------------------------------------------------ */

/* metric must be 1, 2 or 3 (else SCN_ERR_INVALID_PARAMETERS, like NewDistanceCalculator). */
SCN_API int32_t scn_store_create(int32_t device, uint32_t dim, int32_t metric, scn_store** out);
SCN_API int32_t scn_store_destroy(scn_store* s);
SCN_API int32_t scn_store_reserve(scn_store* s, uint64_t rows);
SCN_API int32_t scn_store_clear(scn_store* s);

/* Appends n row-major fp32 vectors (host memory). ids == NULL assigns id = row + 1, the
 * reference's auto-increment (collection.go:57,115-116). id 0 and duplicate ids are rejected
 * (SCN_ERR_INVALID_PARAMETERS; hnsw.go:192-194, 210). Norms and the bf16 mirror are computed on
 * the device. */
SCN_API int32_t scn_store_append(scn_store* s, const float* vecs, const uint64_t* ids, uint64_t n);
/* Same, with vectors already in device memory (ids, if given, in host memory). */
SCN_API int32_t scn_store_append_dev(scn_store* s, const float* d_vecs, const uint64_t* ids, uint64_t n);

/* Soft delete (HNSW.Delete). Unknown id -> SCN_ERR_VECTOR_NOT_FOUND; already deleted is a no-op. */
SCN_API int32_t scn_store_mark_deleted(scn_store* s, const uint64_t* ids, uint64_t n);
/* Restore form (ImportGraphState, hnsw.go:749-804): sets the soft-delete flags of a snapshot and
 * leaves entrypoint / maxLayer exactly as uploaded (hnsw.go:791-793) — a snapshot whose entry point
 * is flagged deleted then answers every search with no result, like the reference. */
SCN_API int32_t scn_store_restore_deleted(scn_store* s, const uint64_t* ids, uint64_t n);

/* Collection.Compact (collection.go:283-313): physically removes the soft-deleted rows. Surviving
 * rows keep their insertion order and ids; the graph is dropped (the reference rebuilds the index
 * from the survivors and the host hands the new graph over with scn_graph_upload). out_removed
 * (may be NULL) = rows dropped. */
SCN_API int32_t scn_store_compact(scn_store* s, uint64_t* out_removed);

SCN_API int32_t scn_store_stats(scn_store* s, scn_stats* out);

/* Copies row vectors back to the host by id (HNSW.Get); out is [n][dim]. */
SCN_API int32_t scn_store_get(scn_store* s, const uint64_t* ids, uint64_t n, float* out);

/* ---- HNSW graph hand-off -------------------------------------------------------------------
 * Flattened core.HNSWGraphState: for node i (any order), node_ids[i], list_counts[i] =
 * len(Connections) (level + 1); then for each of its lists, in layer order, edge_counts[...]
 * neighbour ids concatenated in `edges`. Every node id and neighbour id must already be in the
 * store. m is HNSWParams.M (layer-0 lists hold at most 2*m, upper lists m; longer lists are
 * rejected). entry_id / max_layer are HNSW.entrypoint / maxLayer verbatim. Replaces any previous
 * graph. */
SCN_API int32_t scn_graph_upload(scn_store* s, int32_t m, int32_t max_layer, uint64_t entry_id, uint64_t n_nodes,
                         const uint64_t* node_ids, const int32_t* list_counts, const uint32_t* edge_counts,
                         const uint64_t* edges);

/* ---- GPU-assisted graph construction -------------------------------------------
 * HNSW.Insert / Build with the reference's SERIAL semantics (hnsw.go:148-257: searchLayer with
 * efConstruction, selectNeighbors 560-583, pruneConnections 586-614, entry-point rule 252-254): the
 * next n rows of the store that are not in its graph yet (rows graph_nodes .. graph_nodes + n - 1,
 * in append order) are inserted one after the other. The searches of a window of upcoming inserts
 * run speculatively on the device against one snapshot; inserts are committed strictly in order and
 * only with a search that is provably the one the serial algorithm would have run (expansion logs
 * checked against every adjacency change since the snapshot), otherwise the insert is searched again.
 * The resulting graph is edge for edge the one the reference's insertVector builds for the same
 * level draws. levels[i] = selectLayer()'s draw for the i-th new node (hnsw.go:458-469), made by the
 * caller from its own rand stream (the Go shim uses the CPU index's). m = HNSWParams.M (<= 32),
 * 2*m <= ef_construction <= 1024. Works on an empty graph, after scn_graph_upload, and repeatedly. */
typedef struct scn_build_stats {
  uint64_t inserted, rounds, searches, conflicts, table_overflows;
  uint64_t distance_evals, expansions;   /* device counters over all (speculative) searches */
  /* what ended the rounds: [0] a neighbour was added to a list expanded while W was not full, [1] an added
   * neighbour would have been admitted, [2] an admitted neighbour left the list, [3] entry point / maxLayer
   * moved, [4] expansion log incomplete, [5] beyond the window's pair-distance matrix */
  uint64_t conflict_kind[6];
  double seconds, device_seconds, commit_seconds;   /* total; waiting for the device; host commit + list upload */
} scn_build_stats;
SCN_API int32_t scn_hnsw_insert(scn_store* s, uint64_t n, const int32_t* levels, int32_t m, int32_t ef_construction,
                        scn_build_stats* stats /* may be NULL */);
/* The store's graph as flattened core.HNSWGraphState (the layout scn_graph_upload takes), e.g. to
 * hand a device-built graph to the host index for persistence (ExportGraphState, hnsw.go:703-746). */
SCN_API int32_t scn_graph_export_sizes(scn_store* s, uint64_t* n_nodes, uint64_t* n_lists, uint64_t* n_edges);
SCN_API int32_t scn_graph_export(scn_store* s, uint64_t* node_ids, int32_t* list_counts, uint32_t* edge_counts, uint64_t* edges,
                         uint64_t* entry_id, int32_t* max_layer);

/* ---- restore from an RDB snapshot ---------------------------------------------
 * Reads one collection of a snapshot written by the reference (schemas/flatbuffers/rdb.fbs,
 * rdb.go:239-533) straight into a new device store: vectors, ids, soft-delete flags and the HNSW
 * graph, with the semantics of RDBManager.Load -> ConvertHNSWGraphSnapshot -> ImportGraphState
 * (rdb.go:179-237, 1027-1091; database.go:398-493; hnsw.go:749-804): lists above a node's max_layer
 * are dropped, unparsable neighbour ids are skipped, a snapshot without graph state is refused.
 * Errors: 4001 (recovery failed), 4002 (corrupted data), 3000 / 3002 (database / collection not in
 * the file), 3005 (nodes of different dimension). info (may be NULL) receives the collection's
 * configuration and counts, also when the load itself fails after parsing. */
typedef struct scn_rdb_info {
  int32_t metric;          /* CollectionConfig.metric */
  uint32_t dim;            /* length of the nodes' element arrays */
  int32_t m, ef_construction, ef_search, max_layers;   /* HNSWParams */
  int64_t seed;
  uint64_t nodes;          /* nodes in the graph (including soft-deleted) */
  uint64_t deleted;        /* of those, soft-deleted */
  uint64_t entry_id;       /* HNSWGraph.entrypoint_id */
  int32_t max_layer;       /* HNSWGraph.max_layer */
  int32_t graph_size;      /* HNSWGraph.size */
  int64_t vector_count, deleted_count;   /* CollectionSnapshot counters */
  int32_t has_graph;
} scn_rdb_info;
SCN_API int32_t scn_store_load_rdb(const char* path, const char* database, const char* collection, int32_t device,
                           scn_store** out, scn_rdb_info* info);

/* ---- search (host buffers; blocking) -------------------------------------------------------
 * q is [nq][dim] fp32. Outputs are [nq][k]: ids (0 = none) and distances (+Inf = none), sorted by
 * (distance, insertion row) ascending; out_counts[nq] (may be NULL) = number of valid results. */

/* Exact scan over all live rows. Identical to BatchDistance + stable sort + truncate. */
SCN_API int32_t scn_search_flat(scn_store* s, const float* q, uint64_t nq, uint32_t k, uint64_t* out_ids,
                        float* out_dist, uint32_t* out_counts);

/* HNSW.Search for nq queries. ef <= 0 is invalid here: the Go shim resolves
 * SearchParams.EfSearch / HNSWParams.EfSearch before the call (hnsw.go:300-303). Result count per
 * query is min(k, ef, reachable), as in the reference. */
SCN_API int32_t scn_search_hnsw(scn_store* s, const float* q, uint64_t nq, uint32_t k, uint32_t ef, uint64_t* out_ids,
                        float* out_dist, uint32_t* out_counts);

/* Exact distances for caller-chosen candidates, then top-k: cand_ids is [nq][ncand] (0 = empty
 * slot; unknown or deleted ids are skipped). */
SCN_API int32_t scn_rerank(scn_store* s, const float* q, uint64_t nq, const uint64_t* cand_ids, uint32_t ncand, uint32_t k,
                   uint64_t* out_ids, float* out_dist, uint32_t* out_counts);

/* out[i*nx + j] = Distance(q_i, x_j) with the reference's exact fp32 arithmetic. */
SCN_API int32_t scn_distance_batch(int32_t device, int32_t metric, const float* q, uint64_t nq, const float* x, uint64_t nx,
                           uint32_t dim, float* out);

/* The vector helpers of distance.go:152-192, batched over n vectors of `dim` floats, in the
 * reference's sequential fp32 order (results are bit-identical to the Go functions):
 *   SCN_VEC_MAGNITUDE  out[i]       = VectorMagnitude(a_i)            (distance.go:175-181)   out: [n]
 *   SCN_VEC_NORMALIZE  out[i][...]  = NormalizeVector(a_i); a zero vector comes back unchanged
 *                                     (distance.go:154-172)                                   out: [n][dim]
 *   SCN_VEC_DOT        out[i]       = DotProduct(a_i, b_i)            (distance.go:184-192)   out: [n]
 * b is only read by SCN_VEC_DOT. (DotProduct's length-mismatch case, which returns 0, cannot arise
 * here: both operands have `dim` elements; the host mirror handles it before calling.) */
#define SCN_VEC_MAGNITUDE 1
#define SCN_VEC_NORMALIZE 2
#define SCN_VEC_DOT 3
SCN_API int32_t scn_vector_ops(int32_t device, int32_t op, const float* a, const float* b, uint64_t n, uint32_t dim, float* out);

/* ---- search (device buffers; asynchronous on `stream`, a cudaStream_t) ---------------------- */
SCN_API int32_t scn_search_flat_dev(scn_store* s, const float* d_q, uint64_t nq, uint32_t k, uint64_t* d_out_ids,
                            float* d_out_dist, uint32_t* d_out_counts, void* stream);
SCN_API int32_t scn_search_hnsw_dev(scn_store* s, const float* d_q, uint64_t nq, uint32_t k, uint32_t ef, uint64_t* d_out_ids,
                            float* d_out_dist, uint32_t* d_out_counts, void* stream);

/* Row-sharded search, shard-local half: like scn_search_flat_dev but additionally writes
 * d_out_keys[nq][k], the 64-bit merge keys (order-preserving distance bits << 32 | global row,
 * global row = row_base + local row; ~0 = none), so shards can be merged exactly. */
SCN_API int32_t scn_search_flat_shard_dev(scn_store* s, const float* d_q, uint64_t nq, uint32_t k, uint64_t row_base,
                                  uint64_t* d_out_keys, uint64_t* d_out_ids, void* stream);
/* Merge of G per-shard results laid out [G][nq][k] (e.g. the output of an NCCL all-gather) into
 * the global top-k with the flat scan's (distance, row) order. */
SCN_API int32_t scn_merge_topk_dev(int32_t device, const uint64_t* d_keys, const uint64_t* d_ids, uint32_t n_shards, uint64_t nq,
                           uint32_t k, uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_counts, void* stream);

/* ---- row-sharded search with both exchanges fused over NVLink peer memory ------------------------
 * One scn_exchange per rank (= per GPU / row shard). A batch of nq queries is cut into `world`
 * contiguous slices (scn_exchange_slice); rank r answers slice r. Per call every rank (1) gathers the
 * batch over NVLink from the slices each rank holds (optional), (2) scans its rows for the whole
 * batch, (3) stores each top-k (key, id) list into the buffer of the rank that owns the query
 * (16-byte P2P stores) and raises a flag there, (4) merges its own slice from local memory. Replaces
 * H2D of the whole batch on every rank + all_gather + merge of the NCCL formulation; results are
 * bit-identical to the single-GPU search. Every rank must call for the same batch (same nq), like a
 * collective; ranks driven from one process must use different streams or devices.
 *   between processes : exchange the 64-byte handles (any transport), then scn_exchange_connect
 *   inside one process: scn_exchange_connect_local with the peers' objects (or use scn_shards_*)
 * dim = query dimension for the NVLink query gather (0: not used, q_is_slice must be 0). */
#define SCN_IPC_HANDLE_BYTES 64
typedef struct scn_exchange scn_exchange;
SCN_API int32_t scn_exchange_create(int32_t device, uint32_t rank, uint32_t world, uint64_t max_nq, uint32_t k, uint32_t dim,
                            scn_exchange** out);
SCN_API int32_t scn_exchange_local_handle(scn_exchange* ex, void* out_handle);
SCN_API int32_t scn_exchange_connect(scn_exchange* ex, const void* handles /* [world][64] */);
SCN_API int32_t scn_exchange_connect_local(scn_exchange* ex, scn_exchange* const* peers /* [world] */);
SCN_API int32_t scn_exchange_destroy(scn_exchange* ex);
/* Slice of a batch of nq queries that `rank` answers: queries [first, first + count). */
SCN_API int32_t scn_exchange_slice(const scn_exchange* ex, uint64_t nq, uint32_t rank, uint64_t* out_first, uint64_t* out_count);
/* d_q: the whole batch [nq][dim] (q_is_slice = 0) or this rank's slice only (q_is_slice = 1). The
 * outputs hold the results of this rank's slice: [count][k]. */
SCN_API int32_t scn_search_flat_exchange_dev(scn_store* s, scn_exchange* ex, const float* d_q, int32_t q_is_slice, uint64_t nq,
                                     uint32_t k, uint64_t row_base, uint64_t* d_out_ids, float* d_out_dist,
                                     uint32_t* d_out_counts, void* stream);
/* Host-buffer, blocking form of one rank's call: q_slice = this rank's slice of the batch, outputs =
 * the results of that slice. SCN_ERR_SEARCH_FAILED if a peer missed the 5 s arrival time-out. */
SCN_API int32_t scn_search_flat_exchange(scn_store* s, scn_exchange* ex, const float* q_slice, uint64_t nq, uint32_t k,
                                 uint64_t row_base, uint64_t* out_ids, float* out_dist, uint32_t* out_counts);
/* Synchronises `stream`; SCN_ERR_SEARCH_FAILED if a peer missed the 5 s arrival time-out since the
 * previous status call (the results of such a call are empty). Clears the condition. */
SCN_API int32_t scn_exchange_status(scn_exchange* ex, void* stream);

/* ---- one collection row-sharded over the GPUs of one box (single process) -----------------------
 * The multi-device form of the boundary: core.VectorIndex.Search (interfaces.go:87-111) stays ONE
 * blocking host-buffer call and drives all G GPUs (one worker thread per device inside the library,
 * the fused exchange above between them). Shard g owns the global rows [g*per, (g+1)*per),
 * per = ceil(capacity_rows / G), filled in insertion order; rows beyond the declared capacity go to
 * the last shard. ids == NULL assigns global row + 1 (collection.go:57,115-116). Results are
 * bit-identical to scn_search_flat of one store holding all rows. */
typedef struct scn_shards scn_shards;
SCN_API int32_t scn_shards_create(const int32_t* devices, int32_t ndev, uint32_t dim, int32_t metric, uint64_t capacity_rows,
                          scn_shards** out);
SCN_API int32_t scn_shards_destroy(scn_shards* sh);
SCN_API int32_t scn_shards_count(scn_shards* sh);
SCN_API scn_store* scn_shards_store(scn_shards* sh, int32_t i);   /* shard i (options, statistics); owned by sh */
SCN_API int32_t scn_shards_append(scn_shards* sh, const float* vecs, const uint64_t* ids, uint64_t n);
SCN_API int32_t scn_shards_mark_deleted(scn_shards* sh, const uint64_t* ids, uint64_t n);
SCN_API int32_t scn_shards_stats(scn_shards* sh, scn_stats* out);   /* summed over the shards; device = -1 */
SCN_API int32_t scn_shards_set_option(scn_shards* sh, const char* name, int64_t value);
SCN_API int32_t scn_shards_search_flat(scn_shards* sh, const float* q, uint64_t nq, uint32_t k, uint64_t* out_ids, float* out_dist,
                               uint32_t* out_counts);

/* ---- micro-batching of single-query calls ------------------------------------
 * The reference's API is one query per call (Collection.Search, collection.go:193-204; HNSW.Search,
 * hnsw.go:292-350) issued by many goroutines at once. scn_batcher_search is the drop-in for that
 * call: blocking, one query, callable from any number of threads; calls in flight at the same time
 * are coalesced into one batched launch (leader/follower, no background thread). kind: 0 = exact
 * flat scan, 1 = HNSW (ef as in scn_search_hnsw; ignored for kind 0). A batch is dispatched when it
 * holds max_batch queries, or window_us after its first query arrived provided fewer than two
 * batches are executing (under load batches grow to whatever arrived meanwhile). Results are those
 * of scn_search_flat / scn_search_hnsw for the same query. */
typedef struct scn_batcher scn_batcher;
SCN_API int32_t scn_batcher_create(scn_store* s, int32_t kind, uint32_t max_batch, uint32_t window_us, scn_batcher** out);
SCN_API int32_t scn_batcher_destroy(scn_batcher* b);
SCN_API int32_t scn_batcher_search(scn_batcher* b, const float* q, uint32_t k, uint32_t ef, uint64_t* out_ids, float* out_dist,
                           uint32_t* out_count);
/* out[0] calls, [1] batches, [2] batched launches, [3] largest batch so far. */
SCN_API int32_t scn_batcher_stats(scn_batcher* b, uint64_t* out, int32_t n);

/* ---- tuning / introspection ---------------------------------------------------------------- */
/* Options (defaults are the measured best; the alternatives are kept as switches for comparison and
 * are covered by the parity tests — none of them changes a result):
 *   "flat_path"         0 = auto, 1 = exact CUDA-core scan only, 2 = tensor-core filter + exact rerank
 *   "tensor_min_batch"  smallest batch that takes the tensor-core filter (default 1)
 *   "overfetch"         candidates kept per query and column block by the tensor filter (0 = auto)
 *   "tensor_hint"       1 = lists of a query start from the threshold of the finished ones (default);
 *                       "tensor_hint_target" rows of the shard that should beat a published threshold (0 = 3 k'')
 *   "tensor_bn"         128 forces 128-row tiles (0 = auto); "tensor_chunks" row chunks per query block (0 = auto)
 *   "tensor_pair"       1 = CTA-pair filter kernel (tcgen05 cta_group::2) for 448 < dim <= 512 and 576 < dim <= 768 at
 *                       batches of >= 256 queries (default), 0 = the single-CTA kernel everywhere
 *   "tensor_fused"      candidate merge, exact rerank and certificate behind the filter in one launch per batch
 *                       (one block per query): 1 always, 0 never (three grids), -1 auto (batches of <= 512 queries)
 *   "tensor_pair_ew"    pair kernel: epilogue warps per TMEM lane quarter (each gates 128 / n columns of a tile into its own
 *                       list): 0 = auto (2), 1, 2
 *   "tensor_share"      1 = the candidate lists of a query exchange bounds while the filter builds them where that pays (rows of
 *                       <= 256 elements, >= 16 lists per query, at most two waves of work items), 2 = whenever a query has
 *                       >= 16 lists, 0 = every list on its own
 *   "pdl"               1 = the short kernels behind the tensor filter are launched chained (programmatic dependent
 *                       launch: the start-up of a kernel overlaps the kernel before it), 0 = ordinary launches
 *   "hnsw_gather"       row gather of hnsw_search: -1 auto, 0 registers (LDG.256), 1 / 2 / 3 shared-memory
 *                       stages of 512 B / 2 x 512 B / 256 B; "hnsw_gather_long" the auto choice for rows > 512 B
 *   "hnsw_global"       1 = visited tables in global memory (default), 0 = in shared memory
 *   "hnsw_hash"         entries of the visited table (0 = auto); "hnsw_per_sm" cap on resident queries per SM
 *   "hnsw_early"        1 = rows requested before the visited test (default)
 *   "hnsw_exact_ties"   1 = a walk that ends while an element pushed out of the candidate list still has exactly the
 *                       distance of its last entry (the reference would go on expanding it, hnsw.go:516-518) is redone
 *                       by the exact walk kernel: results equal the reference's even through exact float ties. 0
 *                       (default): such walks — 1 of 10 000 at 1M x 128 — are returned as they are (one re-walk is a
 *                       single-warp latency chain, +0.9 ms on a 3.9 ms batch). The build (scn_hnsw_insert) is always exact.
 *   "build_window"      scn_hnsw_insert: inserts searched speculatively per round (0 = adaptive, 1 = none)
 *   "auto_id_base"      (empty store only) auto-assigned ids become value + row + 1: a row shard of a larger
 *                       collection numbers its rows globally
 *   "profile"           1 = record per-kernel CUDA-event timings and the device counters */
SCN_API int32_t scn_set_option(scn_store* s, const char* name, int64_t value);
/* Per-kernel CUDA-event timings accumulated since the previous call (option "profile" = 1):
 * names[i] -> total ms[i] over counts[i] launches. Synchronises on the recorded events, then
 * clears them. Returns the number of entries written. */
SCN_API int32_t scn_last_timings(scn_store* s, const char** names, float* ms, uint32_t* counts, int32_t max_entries);
/* Test hook: raw filter scores of the tensor-core path, out_scores[nq][rows] (host), where
 * score = aux[row] + c * (bf16(q) . mirror[row]) with c = -2 (L2) or -1 (inner product, cosine). */
SCN_API int32_t scn_debug_tensor_scores(scn_store* s, const float* q, uint64_t nq, float* out_scores);
/* Counters of the last search. Flat: [0] queries served by the tensor path, [1] of those, queries
 * whose first certificate failed (all their chunk candidates were then reranked), [2] queries that
 * failed the second certificate too and were re-scanned exactly. HNSW (option "profile" = 1):
 * [0] distance evaluations, [1] expansions, [2] walks redone by the exact walk kernel (visited table
 * overflow, or a distance tie at the edge of the candidate list that held to the end of the walk). */
SCN_API int32_t scn_last_counters(scn_store* s, uint64_t* out, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* SCN_GPU_H_ */
