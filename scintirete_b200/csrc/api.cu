// api.cu — search entry points of the C ABI (include/scn_gpu.h): argument checks, staging of
// host buffers, path dispatch. No arithmetic lives here.
#include <cmath>

#include "store.h"

using namespace scn;

static int32_t check_search_args(scn_store* s, const void* q, uint64_t nq, uint32_t k, const void* out_ids,
                                 const void* out_dist) {
  if (!s) return fail(SCN_ERR_INVALID_PARAMETERS, "store is NULL");
  if (nq && !q) return fail(SCN_ERR_INVALID_PARAMETERS, "query pointer is NULL");
  if (k == 0) return fail(SCN_ERR_INVALID_PARAMETERS, "top_k must be positive");  // vector_ops.go:186-188
  if (k > 1024) return fail(SCN_ERR_INVALID_PARAMETERS, "top_k above 1024 is not supported");
  if (nq && (!out_ids || !out_dist)) return fail(SCN_ERR_INVALID_PARAMETERS, "output pointer is NULL");
  if (nq >= (1ull << 31)) return fail(SCN_ERR_INVALID_PARAMETERS, "too many queries in one call");
  return SCN_OK;
}

// shard-local flat search producing sorted keys [nq][k]
int32_t scn::flat_keys(scn_store* s, const float* d_q, uint64_t nq, uint32_t k, uint64_t row_base, uint64_t* d_keys,
                         cudaStream_t stream, Profiler* prof) {
  if (row_base + s->rows >= (uint64_t)ROW_NONE)
    return fail(SCN_ERR_INVALID_PARAMETERS, "global row index exceeds 32 bits");
  bool tensor = false;
  if (s->opt_flat_path == 2) {
    if (!tensor_path_supported(s, k))
      return fail(SCN_ERR_INVALID_PARAMETERS, "tensor-core flat path unsupported for dim=%u k=%u", s->dim, k);
    tensor = true;
  } else if (s->opt_flat_path == 0) {
    tensor = tensor_path_supported(s, k) && (int64_t)nq >= s->opt_tensor_min_batch && s->rows >= 4096;
  }
  if (tensor) return flat_search_tensor(s, d_q, nq, k, row_base, d_keys, stream, prof);
  SCN_CUDA(cudaMemsetAsync(s->d_counters, 0, 4 * sizeof(unsigned long long), stream));
  return flat_search_exact(s, d_q, nullptr, nullptr, nq, k, row_base, d_keys, stream, prof);
}

extern "C" {

int32_t scn_search_flat_dev(scn_store* s, const float* d_q, uint64_t nq, uint32_t k, uint64_t* d_out_ids,
                            float* d_out_dist, uint32_t* d_out_counts, void* stream) {
  SCN_TRY(check_search_args(s, d_q, nq, k, d_out_ids, d_out_dist));
  if (nq == 0) return SCN_OK;
  DeviceGuard g(s->device);
  cudaStream_t st = (cudaStream_t)stream;
  Profiler prof(s, st);
  Scratch scratch(st);
  uint64_t* d_keys = nullptr;
  SCN_TRY(scratch.alloc(&d_keys, nq * k));
  SCN_TRY(flat_keys(s, d_q, nq, k, 0, d_keys, st, &prof));
  prof.begin("keys_to_results");
  SCN_TRY(keys_to_results(s, d_keys, nq, 0, d_out_ids, d_out_dist, d_out_counts, k, st, s->rows > 0));   // (an empty store ends in a memset, not a kernel)
  prof.end();
  prof.collect();
  return SCN_OK;
}

int32_t scn_search_flat_shard_dev(scn_store* s, const float* d_q, uint64_t nq, uint32_t k, uint64_t row_base,
                                  uint64_t* d_out_keys, uint64_t* d_out_ids, void* stream) {
  SCN_TRY(check_search_args(s, d_q, nq, k, d_out_keys, d_out_keys));
  if (nq == 0) return SCN_OK;
  DeviceGuard g(s->device);
  cudaStream_t st = (cudaStream_t)stream;
  Profiler prof(s, st);
  SCN_TRY(flat_keys(s, d_q, nq, k, row_base, d_out_keys, st, &prof));
  if (d_out_ids) SCN_TRY(keys_to_results(s, d_out_keys, nq, row_base, d_out_ids, nullptr, nullptr, k, st));
  prof.collect();
  return SCN_OK;
}

int32_t scn_search_flat(scn_store* s, const float* q, uint64_t nq, uint32_t k, uint64_t* out_ids, float* out_dist,
                        uint32_t* out_counts) {
  SCN_TRY(check_search_args(s, q, nq, k, out_ids, out_dist));
  if (nq == 0) return SCN_OK;
  DeviceGuard g(s->device);
  cudaStream_t st = thread_stream(s->device);
  Scratch scratch(st);
  float* d_q = nullptr;
  uint64_t* d_ids = nullptr;
  float* d_dist = nullptr;
  uint32_t* d_counts = nullptr;
  SCN_TRY(scratch.alloc(&d_q, nq * s->dim));
  SCN_TRY(scratch.alloc(&d_ids, nq * k));
  SCN_TRY(scratch.alloc(&d_dist, nq * k));
  SCN_TRY(scratch.alloc(&d_counts, nq));
  SCN_TRY(copy_to_device(d_q, q, nq * s->dim * sizeof(float), st));
  SCN_TRY(scn_search_flat_dev(s, d_q, nq, k, d_ids, d_dist, d_counts, st));
  SCN_CUDA(cudaMemcpyAsync(out_ids, d_ids, nq * k * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  SCN_CUDA(cudaMemcpyAsync(out_dist, d_dist, nq * k * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (out_counts) SCN_CUDA(cudaMemcpyAsync(out_counts, d_counts, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  SCN_CUDA(cudaStreamSynchronize(st));
  return SCN_OK;
}

int32_t scn_search_hnsw_dev(scn_store* s, const float* d_q, uint64_t nq, uint32_t k, uint32_t ef, uint64_t* d_out_ids,
                            float* d_out_dist, uint32_t* d_out_counts, void* stream) {
  SCN_TRY(check_search_args(s, d_q, nq, k, d_out_ids, d_out_dist));
  if (ef == 0 || ef > 4096) return fail(SCN_ERR_INVALID_PARAMETERS, "ef_search must be in [1, 4096]");
  if (nq == 0) return SCN_OK;
  DeviceGuard g(s->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (!s->has_graph || s->entry_id == 0 || s->live == 0) {
    // hnsw.go:296-298: empty index -> no results
    Scratch scratch(st);
    uint64_t* d_keys = nullptr;
    SCN_TRY(scratch.alloc(&d_keys, nq * k));
    SCN_CUDA(cudaMemsetAsync(d_keys, 0xFF, nq * k * sizeof(uint64_t), st));
    SCN_TRY(keys_to_results(s, d_keys, nq, 0, d_out_ids, d_out_dist, d_out_counts, k, st));
    return SCN_OK;
  }
  Profiler prof(s, st);
  SCN_TRY(hnsw_search(s, d_q, nq, k, ef, d_out_ids, d_out_dist, d_out_counts, st, &prof));
  prof.collect();
  return SCN_OK;
}

int32_t scn_search_hnsw(scn_store* s, const float* q, uint64_t nq, uint32_t k, uint32_t ef, uint64_t* out_ids,
                        float* out_dist, uint32_t* out_counts) {
  SCN_TRY(check_search_args(s, q, nq, k, out_ids, out_dist));
  if (nq == 0) return SCN_OK;
  DeviceGuard g(s->device);
  cudaStream_t st = thread_stream(s->device);
  Scratch scratch(st);
  float* d_q = nullptr;
  uint64_t* d_ids = nullptr;
  float* d_dist = nullptr;
  uint32_t* d_counts = nullptr;
  SCN_TRY(scratch.alloc(&d_q, nq * s->dim));
  SCN_TRY(scratch.alloc(&d_ids, nq * k));
  SCN_TRY(scratch.alloc(&d_dist, nq * k));
  SCN_TRY(scratch.alloc(&d_counts, nq));
  SCN_TRY(copy_to_device(d_q, q, nq * s->dim * sizeof(float), st));
  SCN_TRY(scn_search_hnsw_dev(s, d_q, nq, k, ef, d_ids, d_dist, d_counts, st));
  SCN_CUDA(cudaMemcpyAsync(out_ids, d_ids, nq * k * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  SCN_CUDA(cudaMemcpyAsync(out_dist, d_dist, nq * k * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (out_counts) SCN_CUDA(cudaMemcpyAsync(out_counts, d_counts, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  unsigned long long failed = 0;
  SCN_CUDA(cudaMemcpyAsync(&failed, s->d_counters + 3, sizeof failed, cudaMemcpyDeviceToHost, st));
  SCN_CUDA(cudaStreamSynchronize(st));
  if (failed)
    return fail(SCN_ERR_SEARCH_FAILED, "%llu walks visited more rows than the largest visited table holds (ef=%u)", failed, ef);
  return SCN_OK;
}

int32_t scn_rerank(scn_store* s, const float* q, uint64_t nq, const uint64_t* cand_ids, uint32_t ncand, uint32_t k,
                   uint64_t* out_ids, float* out_dist, uint32_t* out_counts) {
  SCN_TRY(check_search_args(s, q, nq, k, out_ids, out_dist));
  if (nq == 0) return SCN_OK;
  if (ncand == 0 || !cand_ids) return fail(SCN_ERR_INVALID_PARAMETERS, "no candidates given");
  DeviceGuard g(s->device);
  cudaStream_t st = thread_stream(s->device);
  // ids -> rows on the host (the id map lives there); unknown ids and 0 become empty slots
  std::vector<uint32_t> rows((size_t)nq * ncand);
  for (size_t i = 0; i < rows.size(); ++i) {
    uint32_t r;
    rows[i] = (cand_ids[i] != 0 && s->lookup(cand_ids[i], &r)) ? r : ROW_NONE;
  }
  Scratch scratch(st);
  float* d_q = nullptr;
  uint32_t* d_rows = nullptr;
  uint64_t* d_keys = nullptr;
  uint64_t* d_ids = nullptr;
  float* d_dist = nullptr;
  uint32_t* d_counts = nullptr;
  SCN_TRY(scratch.alloc(&d_q, nq * s->dim));
  SCN_TRY(scratch.alloc(&d_rows, rows.size()));
  SCN_TRY(scratch.alloc(&d_keys, nq * k));
  SCN_TRY(scratch.alloc(&d_ids, nq * k));
  SCN_TRY(scratch.alloc(&d_dist, nq * k));
  SCN_TRY(scratch.alloc(&d_counts, nq));
  SCN_TRY(copy_to_device(d_q, q, nq * s->dim * sizeof(float), st));
  SCN_CUDA(cudaMemcpyAsync(d_rows, rows.data(), rows.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  SCN_TRY(rerank_rows(s, d_q, nq, d_rows, ncand, k, 0, d_keys, st));
  SCN_TRY(keys_to_results(s, d_keys, nq, 0, d_ids, d_dist, d_counts, k, st));
  SCN_CUDA(cudaMemcpyAsync(out_ids, d_ids, nq * k * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  SCN_CUDA(cudaMemcpyAsync(out_dist, d_dist, nq * k * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (out_counts) SCN_CUDA(cudaMemcpyAsync(out_counts, d_counts, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  SCN_CUDA(cudaStreamSynchronize(st));
  return SCN_OK;
}

int32_t scn_distance_batch(int32_t device, int32_t metric, const float* q, uint64_t nq, const float* x, uint64_t nx,
                           uint32_t dim, float* out) {
  if (metric != M_L2 && metric != M_COS && metric != M_IP)
    return fail(SCN_ERR_INVALID_PARAMETERS, "unsupported distance metric");
  if (nq == 0 || nx == 0) return SCN_OK;
  if (!q || !x || !out) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  if (dim == 0) return fail(SCN_ERR_INVALID_PARAMETERS, "dimension must be positive");
  DeviceGuard g(device);
  cudaStream_t st = thread_stream(device);
  Scratch scratch(st);
  float *d_q = nullptr, *d_x = nullptr, *d_o = nullptr;
  SCN_TRY(scratch.alloc(&d_q, nq * dim));
  SCN_TRY(scratch.alloc(&d_x, nx * dim));
  SCN_TRY(scratch.alloc(&d_o, nq * nx));
  SCN_CUDA(cudaMemcpyAsync(d_q, q, nq * dim * sizeof(float), cudaMemcpyHostToDevice, st));
  SCN_CUDA(cudaMemcpyAsync(d_x, x, nx * dim * sizeof(float), cudaMemcpyHostToDevice, st));
  SCN_TRY(distance_batch(metric, d_q, nq, d_x, nx, dim, d_o, st));
  SCN_CUDA(cudaMemcpyAsync(out, d_o, nq * nx * sizeof(float), cudaMemcpyDeviceToHost, st));
  SCN_CUDA(cudaStreamSynchronize(st));
  return SCN_OK;
}

int32_t scn_vector_ops(int32_t device, int32_t op, const float* a, const float* b, uint64_t n, uint32_t dim, float* out) {
  if (op != SCN_VEC_MAGNITUDE && op != SCN_VEC_NORMALIZE && op != SCN_VEC_DOT)
    return fail(SCN_ERR_INVALID_PARAMETERS, "unknown vector operation %d", op);
  if (n == 0) return SCN_OK;
  if (!a || !out || (op == SCN_VEC_DOT && !b)) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  DeviceGuard g(device);
  cudaStream_t st = thread_stream(device);
  Scratch scratch(st);
  // dim == 0: magnitude and dot of empty vectors are 0 and an empty vector normalises to itself
  const uint64_t n_in = n * dim, n_out = (op == SCN_VEC_NORMALIZE) ? n * dim : n;
  if (n_out == 0) return SCN_OK;
  float *d_a = nullptr, *d_b = nullptr, *d_o = nullptr;
  SCN_TRY(scratch.alloc(&d_a, std::max<uint64_t>(n_in, 1)));
  SCN_TRY(scratch.alloc(&d_o, n_out));
  if (n_in) SCN_CUDA(cudaMemcpyAsync(d_a, a, n_in * sizeof(float), cudaMemcpyHostToDevice, st));
  if (op == SCN_VEC_DOT) {
    SCN_TRY(scratch.alloc(&d_b, std::max<uint64_t>(n_in, 1)));
    if (n_in) SCN_CUDA(cudaMemcpyAsync(d_b, b, n_in * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  SCN_TRY(vector_ops(op, d_a, d_b, n, dim, d_o, st));
  SCN_CUDA(cudaMemcpyAsync(out, d_o, n_out * sizeof(float), cudaMemcpyDeviceToHost, st));
  SCN_CUDA(cudaStreamSynchronize(st));
  return SCN_OK;
}

int32_t scn_debug_tensor_scores(scn_store* s, const float* q, uint64_t nq, float* out_scores) {
  if (!s || !q || !out_scores) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  if (!tensor_path_supported(s, 10)) return fail(SCN_ERR_INVALID_PARAMETERS, "tensor-core path unsupported for dim=%u", s->dim);
  DeviceGuard g(s->device);
  cudaStream_t st = thread_stream(s->device);
  Scratch scratch(st);
  float *d_q = nullptr, *d_o = nullptr;
  SCN_TRY(scratch.alloc(&d_q, nq * s->dim));
  SCN_TRY(scratch.alloc(&d_o, nq * s->rows));
  SCN_TRY(copy_to_device(d_q, q, nq * s->dim * sizeof(float), st));
  SCN_TRY(tensor_debug_scores(s, d_q, nq, d_o, st));
  SCN_CUDA(cudaMemcpyAsync(out_scores, d_o, nq * s->rows * sizeof(float), cudaMemcpyDeviceToHost, st));
  SCN_CUDA(cudaStreamSynchronize(st));
  return SCN_OK;
}

int32_t scn_merge_topk_dev(int32_t device, const uint64_t* d_keys, const uint64_t* d_ids, uint32_t n_shards, uint64_t nq,
                           uint32_t k, uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_counts, void* stream) {
  if (!d_keys || !d_ids || !d_out_ids || !d_out_dist) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  if (k == 0 || k > 1024) return fail(SCN_ERR_INVALID_PARAMETERS, "top_k must be in [1, 1024]");
  DeviceGuard g(device);
  return merge_topk(d_keys, d_ids, n_shards, nq, k, d_out_ids, d_out_dist, d_out_counts, (cudaStream_t)stream);
}

}  // extern "C"
