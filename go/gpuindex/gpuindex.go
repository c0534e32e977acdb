//go:build gpu && cgo && linux && amd64

// Package gpuindex is the cgo shim that puts libscn_gpu.so (include/scn_gpu.h) behind the
// reference's own core.HNSWIndex / core.VectorIndex interfaces
// (internal/core/interfaces.go:87-134). It is the Go half of the drop-in boundary.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain. Every call made here
// has an identical, tested twin in scintirete_b200/index.py (ctypes over the same C ABI); keep the
// two in lock-step. Build inside the reference tree with
//
//	CGO_ENABLED=1 go build -tags gpu ./...
//
// and the environment described in INTEGRATION.md.
//
// Design: GPUIndex embeds the reference's CPU *algorithm.HNSW. Graph construction
// (Insert/Build: searchLayer with efConstruction, selectNeighbors, pruneConnections —
// hnsw.go:190-257, 560-614), persistence hand-off (ExportGraphState/ImportGraphState) and Get stay
// on the embedded index, unchanged. Search, SearchBatch and SearchExact run on the GPU over a
// device-memory mirror that is brought up to date lazily (dirty flag) before the next search.
package gpuindex

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../scintirete_b200 -lscn_gpu -Wl,-rpath,${SRCDIR}/../../scintirete_b200
#include <stdlib.h>
#include "scn_gpu.h"
*/
import "C"

import (
	"context"
	"fmt"
	"math"
	"runtime"
	"sync"
	"unsafe"

	"github.com/scintirete/scintirete/internal/core"
	"github.com/scintirete/scintirete/internal/core/algorithm"
	"github.com/scintirete/scintirete/internal/utils"
	"github.com/scintirete/scintirete/pkg/types"
)

// GPUIndex implements core.HNSWIndex.
type GPUIndex struct {
	core.HNSWIndex // embedded CPU index: build, persistence, Get, statistics

	mu     sync.RWMutex // guards store + dirty; the embedded index has its own lock
	store  *C.scn_store
	dim    int
	metric types.DistanceMetric
	params types.HNSWParams
	dirty  bool // CPU graph changed since the last upload
	flat   bool // "flat-gpu": Search is the exact scan, no graph needed

	// batcher coalesces the one-query Search calls of concurrent goroutines (collection.go:193-204
	// is called once per request) into batched launches; nil = every Search is its own launch.
	batcher *C.scn_batcher
}

// EnableMicroBatching routes Search through scn_batcher_search: a batch is dispatched when it holds
// maxBatch queries or windowMicros after its first query arrived.
func (g *GPUIndex) EnableMicroBatching(maxBatch, windowMicros int) error {
	g.mu.Lock()
	defer g.mu.Unlock()
	kind := C.int32_t(1)
	if g.flat {
		kind = 0
	}
	return scnErr(C.scn_batcher_create(g.store, kind, C.uint32_t(maxBatch), C.uint32_t(windowMicros), &g.batcher))
}

func scnErr(rc C.int32_t) error {
	if rc == 0 {
		return nil
	}
	// status codes are utils.ErrorCode numbers (internal/utils/errors.go:11-49)
	return utils.NewError(utils.ErrorCode(int(rc)), C.GoString(C.scn_last_error()))
}

// New creates the CPU index and its device mirror. dim must be known up front (the reference
// learns it from the first insert, collection.go:80-109; the factory passes CollectionConfig's).
func New(params types.HNSWParams, metric types.DistanceMetric, dim int, device int, flat bool) (*GPUIndex, error) {
	cpu, err := algorithm.NewHNSW(params, metric) // also validates the metric (distance.go:129-140)
	if err != nil {
		return nil, err
	}
	g := &GPUIndex{HNSWIndex: cpu, dim: dim, metric: metric, params: params, flat: flat}
	if rc := C.scn_store_create(C.int32_t(device), C.uint32_t(dim), C.int32_t(metric), &g.store); rc != 0 {
		return nil, scnErr(rc)
	}
	runtime.SetFinalizer(g, func(g *GPUIndex) { g.Close() })
	return g, nil
}

func (g *GPUIndex) Close() {
	g.mu.Lock()
	defer g.mu.Unlock()
	if g.batcher != nil {
		C.scn_batcher_destroy(g.batcher)
		g.batcher = nil
	}
	if g.store != nil {
		C.scn_store_destroy(g.store)
		g.store = nil
	}
}

// ---- mutators: CPU index first, then mark the mirror dirty (vectors are appended eagerly) ----

func (g *GPUIndex) Insert(ctx context.Context, v types.Vector) error {
	if len(v.Elements) != g.dim {
		return utils.ErrInvalidVectorDimension(fmt.Sprintf("vector has dimension %d, expected %d", len(v.Elements), g.dim))
	}
	if err := g.HNSWIndex.Insert(ctx, v); err != nil {
		return err
	}
	g.mu.Lock()
	defer g.mu.Unlock()
	id := C.uint64_t(v.ID)
	// cgo rule: the library copies out of Go memory before returning, no pointer is retained
	rc := C.scn_store_append(g.store, (*C.float)(unsafe.Pointer(&v.Elements[0])), &id, 1)
	g.dirty = true
	return scnErr(rc)
}

func (g *GPUIndex) Build(ctx context.Context, vs []types.Vector) error {
	if err := g.HNSWIndex.Build(ctx, vs); err != nil {
		return err
	}
	g.mu.Lock()
	defer g.mu.Unlock()
	if rc := C.scn_store_clear(g.store); rc != 0 {
		return scnErr(rc)
	}
	if len(vs) == 0 {
		return nil
	}
	flat := make([]float32, 0, len(vs)*g.dim)
	ids := make([]C.uint64_t, len(vs))
	for i, v := range vs {
		flat = append(flat, v.Elements...)
		ids[i] = C.uint64_t(v.ID)
	}
	rc := C.scn_store_append(g.store, (*C.float)(unsafe.Pointer(&flat[0])), &ids[0], C.uint64_t(len(vs)))
	g.dirty = true
	return scnErr(rc)
}

func (g *GPUIndex) Delete(ctx context.Context, id string) error {
	if err := g.HNSWIndex.Delete(ctx, id); err != nil {
		return err
	}
	var vid uint64
	fmt.Sscanf(id, "%d", &vid) // same parse as hnsw.go:265-268; already validated above
	g.mu.Lock()
	defer g.mu.Unlock()
	cid := C.uint64_t(vid)
	rc := C.scn_store_mark_deleted(g.store, &cid, 1)
	g.dirty = true // Delete may move the entrypoint (hnsw.go:283-285)
	return scnErr(rc)
}

func (g *GPUIndex) ImportGraphState(st core.HNSWGraphState) error {
	if err := g.HNSWIndex.ImportGraphState(st); err != nil {
		return err
	}
	g.mu.Lock()
	defer g.mu.Unlock()
	if rc := C.scn_store_clear(g.store); rc != 0 {
		return scnErr(rc)
	}
	flat := make([]float32, 0, len(st.Nodes)*g.dim)
	ids := make([]C.uint64_t, 0, len(st.Nodes))
	var dead []C.uint64_t
	for id, n := range st.Nodes {
		flat = append(flat, n.Vector...)
		ids = append(ids, C.uint64_t(id))
		if n.Deleted {
			dead = append(dead, C.uint64_t(id))
		}
	}
	if len(ids) > 0 {
		if rc := C.scn_store_append(g.store, (*C.float)(unsafe.Pointer(&flat[0])), &ids[0], C.uint64_t(len(ids))); rc != 0 {
			return scnErr(rc)
		}
	}
	if len(dead) > 0 {
		if rc := C.scn_store_mark_deleted(g.store, &dead[0], C.uint64_t(len(dead))); rc != 0 {
			return scnErr(rc)
		}
	}
	g.dirty = true
	return nil
}

// syncGraph flattens ExportGraphState (hnsw.go:703-745) into the arrays scn_graph_upload takes.
// Caller holds g.mu for writing.
func (g *GPUIndex) syncGraph() error {
	if !g.dirty || g.flat {
		g.dirty = false
		return nil
	}
	st := g.HNSWIndex.ExportGraphState()
	n := len(st.Nodes)
	ids := make([]C.uint64_t, 0, n)
	lists := make([]C.int32_t, 0, n)
	var counts []C.uint32_t
	var edges []C.uint64_t
	for id, node := range st.Nodes {
		ids = append(ids, C.uint64_t(id))
		lists = append(lists, C.int32_t(len(node.Connections)))
		for _, l := range node.Connections {
			counts = append(counts, C.uint32_t(len(l)))
			for _, e := range l {
				edges = append(edges, C.uint64_t(e))
			}
		}
	}
	var pi *C.uint64_t
	var pl *C.int32_t
	var pc *C.uint32_t
	var pe *C.uint64_t
	if n > 0 {
		pi, pl = &ids[0], &lists[0]
	}
	if len(counts) > 0 {
		pc = &counts[0]
	}
	if len(edges) > 0 {
		pe = &edges[0]
	}
	rc := C.scn_graph_upload(g.store, C.int32_t(g.params.M), C.int32_t(st.MaxLayer), C.uint64_t(st.EntryPoint),
		C.uint64_t(n), pi, pl, pc, pe)
	if rc == 0 {
		g.dirty = false
	}
	return scnErr(rc)
}

// ---- search ----

func (g *GPUIndex) ef(p types.SearchParams) int { // hnsw.go:300-303
	if p.EfSearch != nil && *p.EfSearch > 0 {
		return *p.EfSearch
	}
	return g.HNSWIndex.GetParameters().EfSearch
}

// SearchBatch answers len(queries)/dim queries in one device pass. queries is row-major.
func (g *GPUIndex) SearchBatch(ctx context.Context, queries []float32, p types.SearchParams) ([][]types.SearchResult, error) {
	if p.TopK <= 0 {
		return nil, utils.ErrInvalidParameters("top_k must be positive")
	}
	if len(queries) == 0 || len(queries)%g.dim != 0 {
		return nil, utils.ErrInvalidVectorDimension(fmt.Sprintf("query length %d is not a multiple of %d", len(queries), g.dim))
	}
	nq := len(queries) / g.dim
	g.mu.RLock()
	if g.dirty {
		g.mu.RUnlock()
		g.mu.Lock()
		err := g.syncGraph()
		g.mu.Unlock()
		if err != nil {
			return nil, err
		}
		g.mu.RLock()
	}
	defer g.mu.RUnlock()
	k := p.TopK
	ids := make([]C.uint64_t, nq*k)
	dist := make([]C.float, nq*k)
	cnt := make([]C.uint32_t, nq)
	var rc C.int32_t
	if g.flat {
		rc = C.scn_search_flat(g.store, (*C.float)(unsafe.Pointer(&queries[0])), C.uint64_t(nq), C.uint32_t(k), &ids[0], &dist[0], &cnt[0])
	} else {
		rc = C.scn_search_hnsw(g.store, (*C.float)(unsafe.Pointer(&queries[0])), C.uint64_t(nq), C.uint32_t(k),
			C.uint32_t(g.ef(p)), &ids[0], &dist[0], &cnt[0])
	}
	if rc != 0 {
		return nil, scnErr(rc)
	}
	out := make([][]types.SearchResult, nq)
	for q := 0; q < nq; q++ {
		res := make([]types.SearchResult, 0, int(cnt[q]))
		for j := 0; j < int(cnt[q]); j++ {
			id := uint64(ids[q*k+j])
			r := types.SearchResult{Vector: types.Vector{ID: id}, Distance: float32(dist[q*k+j])}
			// re-attach Elements / Metadata from the host node map, as hnsw.go:331-335 shares them
			if v, err := g.HNSWIndex.Get(ctx, fmt.Sprintf("%d", id)); err == nil && v != nil {
				r.Vector = *v
			}
			res = append(res, r)
		}
		out[q] = res
	}
	return out, nil
}

// Search keeps the reference's one-query signature (interfaces.go:98).
func (g *GPUIndex) Search(ctx context.Context, query []float32, p types.SearchParams) ([]types.SearchResult, error) {
	if len(query) != g.dim {
		return nil, utils.ErrInvalidVectorDimension(fmt.Sprintf("query has dimension %d, expected %d", len(query), g.dim))
	}
	if g.batcher != nil && p.TopK > 0 {
		return g.searchCoalesced(ctx, query, p)
	}
	r, err := g.SearchBatch(ctx, query, p)
	if err != nil {
		return nil, err
	}
	return r[0], nil
}

// searchCoalesced is Search through the micro-batcher: this goroutine blocks in cgo until the
// batch its query joined has been answered.
func (g *GPUIndex) searchCoalesced(ctx context.Context, query []float32, p types.SearchParams) ([]types.SearchResult, error) {
	g.mu.RLock()
	if g.dirty {
		g.mu.RUnlock()
		g.mu.Lock()
		err := g.syncGraph()
		g.mu.Unlock()
		if err != nil {
			return nil, err
		}
		g.mu.RLock()
	}
	defer g.mu.RUnlock()
	k := p.TopK
	ids := make([]C.uint64_t, k)
	dist := make([]C.float, k)
	var cnt C.uint32_t
	rc := C.scn_batcher_search(g.batcher, (*C.float)(unsafe.Pointer(&query[0])), C.uint32_t(k), C.uint32_t(g.ef(p)),
		&ids[0], &dist[0], &cnt)
	if rc != 0 {
		return nil, scnErr(rc)
	}
	res := make([]types.SearchResult, 0, int(cnt))
	for j := 0; j < int(cnt); j++ {
		id := uint64(ids[j])
		r := types.SearchResult{Vector: types.Vector{ID: id}, Distance: float32(dist[j])}
		if v, err := g.HNSWIndex.Get(ctx, fmt.Sprintf("%d", id)); err == nil && v != nil {
			r.Vector = *v
		}
		res = append(res, r)
	}
	return res, nil
}

// Compact is the device half of Collection.Compact (collection.go:283-313): the caller rebuilds the
// CPU index from the surviving vectors (index.Build); here the deleted rows leave HBM and the stale
// graph is dropped, to be uploaded again before the next search.
func (g *GPUIndex) Compact() (int, error) {
	g.mu.Lock()
	defer g.mu.Unlock()
	var removed C.uint64_t
	rc := C.scn_store_compact(g.store, &removed)
	g.dirty = true
	return int(removed), scnErr(rc)
}

// SearchExact is the flat ground truth over the same rows.
func (g *GPUIndex) SearchExact(ctx context.Context, queries []float32, topK int) ([]uint64, []float32, error) {
	nq := len(queries) / g.dim
	ids := make([]C.uint64_t, nq*topK)
	dist := make([]C.float, nq*topK)
	g.mu.RLock()
	defer g.mu.RUnlock()
	rc := C.scn_search_flat(g.store, (*C.float)(unsafe.Pointer(&queries[0])), C.uint64_t(nq), C.uint32_t(topK), &ids[0], &dist[0], nil)
	if rc != 0 {
		return nil, nil, scnErr(rc)
	}
	oi := make([]uint64, len(ids))
	od := make([]float32, len(dist))
	for i := range ids {
		oi[i], od[i] = uint64(ids[i]), float32(dist[i])
	}
	return oi, od, nil
}

func (g *GPUIndex) MemoryUsage() int64 {
	var st C.scn_stats
	g.mu.RLock()
	defer g.mu.RUnlock()
	C.scn_store_stats(g.store, &st)
	return g.HNSWIndex.MemoryUsage() + int64(st.device_bytes)
}

// GPUDistance is a core.DistanceCalculator whose batched form runs on the device.
type GPUDistance struct {
	Metric types.DistanceMetric
	Device int
}

func (d GPUDistance) Distance(a, b []float32) float32 {
	if len(a) != len(b) {
		return float32(math.Inf(1)) // distance.go:22-24
	}
	var out C.float
	C.scn_distance_batch(C.int32_t(d.Device), C.int32_t(d.Metric), (*C.float)(unsafe.Pointer(&a[0])), 1,
		(*C.float)(unsafe.Pointer(&b[0])), 1, C.uint32_t(len(a)), &out)
	return float32(out)
}
func (d GPUDistance) DistanceType() types.DistanceMetric { return d.Metric }
func (d GPUDistance) IsSimilarity() bool                 { return false }

// The batched forms of the helpers in distance.go:152-192 (NormalizeVector, VectorMagnitude,
// DotProduct): n vectors of dim floats each, flat; results carry the bits of the Go loops.
func vectorOps(device int, op C.int32_t, a, b []float32, n, dim int, out []float32) error {
	if n == 0 {
		return nil
	}
	var pb *C.float
	if b != nil {
		pb = (*C.float)(unsafe.Pointer(&b[0]))
	}
	return scnErr(C.scn_vector_ops(C.int32_t(device), op, (*C.float)(unsafe.Pointer(&a[0])), pb, C.uint64_t(n),
		C.uint32_t(dim), (*C.float)(unsafe.Pointer(&out[0]))))
}

// NormalizeVectors: out[i] = algorithm.NormalizeVector(a[i]); zero vectors come back unchanged.
func NormalizeVectors(device int, a []float32, n, dim int) ([]float32, error) {
	out := make([]float32, n*dim)
	return out, vectorOps(device, C.SCN_VEC_NORMALIZE, a, nil, n, dim, out)
}

// VectorMagnitudes: out[i] = algorithm.VectorMagnitude(a[i]).
func VectorMagnitudes(device int, a []float32, n, dim int) ([]float32, error) {
	out := make([]float32, n)
	return out, vectorOps(device, C.SCN_VEC_MAGNITUDE, a, nil, n, dim, out)
}

// DotProducts: out[i] = algorithm.DotProduct(a[i], b[i]); a and b must hold n*dim floats each
// (the reference's length-mismatch case returns 0 and needs no device call).
func DotProducts(device int, a, b []float32, n, dim int) ([]float32, error) {
	out := make([]float32, n)
	if len(a) != len(b) {
		return out, nil
	}
	return out, vectorOps(device, C.SCN_VEC_DOT, a, b, n, dim, out)
}

// Factory implements core.IndexFactory (interfaces.go:187-196) for "hnsw-gpu" and "flat-gpu".
type Factory struct{ Device int }

func (f Factory) CreateIndex(cfg types.IndexConfig) (core.VectorIndex, error) {
	params := types.DefaultHNSWParams()
	dim, _ := cfg.Parameters["dim"].(int)
	if v, ok := cfg.Parameters["m"].(int); ok {
		params.M = v
	}
	if v, ok := cfg.Parameters["ef_construction"].(int); ok {
		params.EfConstruction = v
	}
	if v, ok := cfg.Parameters["ef_search"].(int); ok {
		params.EfSearch = v
	}
	return New(params, cfg.Metric, dim, f.Device, cfg.Type == "flat-gpu")
}
func (f Factory) SupportedMetrics() []types.DistanceMetric {
	return []types.DistanceMetric{types.DistanceMetricL2, types.DistanceMetricCosine, types.DistanceMetricInnerProduct}
}
func (f Factory) DefaultParameters() map[string]interface{} {
	return map[string]interface{}{"m": 16, "ef_construction": 200, "ef_search": 50}
}
