//go:build gpu && cgo && linux && amd64

// Package gpuindex is the cgo shim that puts libscn_gpu.so (include/scn_gpu.h) behind the
// reference's own core.HNSWIndex / core.VectorIndex interfaces
// (internal/core/interfaces.go:87-134). It is the Go half of the drop-in boundary.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain. Every call made here
// has an identical, tested twin in scintirete_b200/index.py and sharding.py (ctypes over the same C
// ABI; tests/test_abi.py checks that every C.scn_* call below names a declared entry point with the
// declared number of arguments); keep the two in lock-step. Build inside the reference tree with
//
//	CGO_ENABLED=1 go build -tags gpu ./...
//
// and the environment described in INTEGRATION.md.
//
// Design: GPUIndex IS the index — it does not wrap algorithm.HNSW. The host side keeps what the
// reference's node map keeps besides the graph (hnsw.go:17-26): id -> {Elements, Metadata, Deleted},
// in insertion order (= device row order). Vectors, norms, the bf16 mirror and the whole graph live
// in HBM. Insert/Build link nodes with scn_hnsw_insert (reference-serial semantics, searches on the
// GPU); the level of every node is drawn here, from rand.New(rand.NewSource(params.Seed)) with the
// formula of selectLayer (hnsw.go:458-469), i.e. from the very stream the CPU index would have used,
// so the graph is the one algorithm.HNSW builds. ExportGraphState / ImportGraphState hand the graph
// to and from persistence unchanged (hnsw.go:703-804).
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
	"math/rand"
	"runtime"
	"sort"
	"sync"
	"unsafe"

	"github.com/scintirete/scintirete/internal/core"
	"github.com/scintirete/scintirete/internal/core/algorithm"
	"github.com/scintirete/scintirete/internal/utils"
	"github.com/scintirete/scintirete/pkg/types"
)

// call runs one C entry point and, on failure, reads its message. scn_last_error is thread-local:
// the goroutine is pinned to its OS thread for the pair, so the message read is the failing call's.
func call(f func() C.int32_t) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	rc := f()
	if rc == 0 {
		return nil
	}
	// status codes are utils.ErrorCode numbers (internal/utils/errors.go:11-49)
	return utils.NewError(utils.ErrorCode(int(rc)), C.GoString(C.scn_last_error()))
}

type hostNode struct {
	elements []float32
	metadata map[string]interface{}
	deleted  bool
}

// GPUIndex implements core.HNSWIndex.
type GPUIndex struct {
	mu     sync.RWMutex // many Search (RLock), exclusive mutators (Lock) — as hnsw.go:178, 261, 293, 750
	store  *C.scn_store
	dim    int
	metric types.DistanceMetric
	params types.HNSWParams
	rng    *rand.Rand // selectLayer's stream (hnsw.go:143, 461)
	flat   bool       // "flat-gpu": Search is the exact scan, no graph is built
	failed error      // set when the device mirror may have diverged; every later call returns it

	nodes map[uint64]*hostNode
	order []uint64 // insertion order == device row order
	size  int      // live nodes (hnsw.go:375-379)

	// batcher coalesces the one-query Search calls of concurrent goroutines (collection.go:193-204
	// is called once per request) into batched launches; nil = every Search is its own launch.
	batcher *C.scn_batcher
}

// New creates the index. dim must be known up front (the reference learns it from the first insert,
// collection.go:80-109; the factory passes CollectionConfig's).
func New(params types.HNSWParams, metric types.DistanceMetric, dim int, device int, flat bool) (*GPUIndex, error) {
	if _, err := algorithm.NewDistanceCalculator(metric); err != nil { // distance.go:129-140
		return nil, err
	}
	g := &GPUIndex{dim: dim, metric: metric, params: params, flat: flat, nodes: map[uint64]*hostNode{},
		rng: rand.New(rand.NewSource(params.Seed))}
	if err := call(func() C.int32_t {
		return C.scn_store_create(C.int32_t(device), C.uint32_t(dim), C.int32_t(metric), &g.store)
	}); err != nil {
		return nil, err
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

// EnableMicroBatching routes Search through scn_batcher_search: a batch is dispatched when it holds
// maxBatch queries or windowMicros after its first query arrived.
func (g *GPUIndex) EnableMicroBatching(maxBatch, windowMicros int) error {
	g.mu.Lock()
	defer g.mu.Unlock()
	kind := C.int32_t(1)
	if g.flat {
		kind = 0
	}
	return call(func() C.int32_t {
		return C.scn_batcher_create(g.store, kind, C.uint32_t(maxBatch), C.uint32_t(windowMicros), &g.batcher)
	})
}

// selectLayer is hnsw.go:458-469 verbatim (one Float64 per insert).
func (g *GPUIndex) selectLayer() int32 {
	mL := 1.0 / math.Log(2.0)
	level := int(math.Floor(-math.Log(g.rng.Float64()) * mL))
	if level >= g.params.MaxLayers {
		level = g.params.MaxLayers - 1
	}
	if level < 0 {
		level = 0
	}
	return int32(level)
}

// link inserts the last len(levels) appended rows into the device graph. Caller holds g.mu.
func (g *GPUIndex) link(levels []int32) error {
	if g.flat || len(levels) == 0 {
		return nil
	}
	return call(func() C.int32_t {
		return C.scn_hnsw_insert(g.store, C.uint64_t(len(levels)), (*C.int32_t)(unsafe.Pointer(&levels[0])),
			C.int32_t(g.params.M), C.int32_t(g.params.EfConstruction), nil)
	})
}

// ---- mutators ----------------------------------------------------------------------------------

func (g *GPUIndex) Insert(ctx context.Context, v types.Vector) error {
	g.mu.Lock()
	defer g.mu.Unlock()
	if g.failed != nil {
		return g.failed
	}
	if len(v.Elements) != g.dim {
		return utils.ErrInsertFailed(fmt.Sprintf("failed to insert vector %d: dimension %d, expected %d", v.ID, len(v.Elements), g.dim))
	}
	if _, dup := g.nodes[v.ID]; dup || v.ID == 0 { // hnsw.go:192-194 (wrapped by Insert, 181-183); id 0 = "no entrypoint"
		return utils.ErrInsertFailed(fmt.Sprintf("failed to insert vector %d", v.ID))
	}
	id := C.uint64_t(v.ID)
	// cgo rule: the library copies out of Go memory before returning, no pointer is retained
	if err := call(func() C.int32_t {
		return C.scn_store_append(g.store, (*C.float)(unsafe.Pointer(&v.Elements[0])), &id, 1)
	}); err != nil {
		return err // nothing was changed on either side
	}
	level := []int32{g.selectLayer()}
	if err := g.link(level); err != nil {
		// the row is stored but not linked: host and device would diverge from here on
		g.failed = utils.ErrIndexOperationFailed("device index unusable after a failed insert: " + err.Error())
		return err
	}
	g.nodes[v.ID] = &hostNode{elements: v.Elements, metadata: v.Metadata}
	g.order = append(g.order, v.ID)
	g.size++
	return nil
}

func (g *GPUIndex) Build(ctx context.Context, vs []types.Vector) error { // hnsw.go:148-174
	g.mu.Lock()
	defer g.mu.Unlock()
	if err := call(func() C.int32_t { return C.scn_store_clear(g.store) }); err != nil {
		return err
	}
	g.nodes, g.order, g.size, g.failed = map[uint64]*hostNode{}, nil, 0, nil
	if len(vs) == 0 {
		return nil
	}
	flat := make([]float32, 0, len(vs)*g.dim)
	ids := make([]C.uint64_t, len(vs))
	levels := make([]int32, len(vs))
	for i, v := range vs {
		if len(v.Elements) != g.dim {
			return utils.ErrIndexBuildFailed(fmt.Sprintf("vector %d has dimension %d, expected %d", v.ID, len(v.Elements), g.dim))
		}
		flat = append(flat, v.Elements...)
		ids[i] = C.uint64_t(v.ID)
		levels[i] = g.selectLayer() // slice order, as the serial Build draws them
	}
	if err := call(func() C.int32_t {
		return C.scn_store_append(g.store, (*C.float)(unsafe.Pointer(&flat[0])), &ids[0], C.uint64_t(len(vs)))
	}); err != nil {
		return err
	}
	if err := g.link(levels); err != nil {
		g.failed = utils.ErrIndexBuildFailed(err.Error())
		return g.failed
	}
	for _, v := range vs {
		g.nodes[v.ID] = &hostNode{elements: v.Elements, metadata: v.Metadata}
		g.order = append(g.order, v.ID)
	}
	g.size = len(vs)
	return nil
}

func parseID(id string) (uint64, error) { // hnsw.go:265-268
	var vid uint64
	if _, err := fmt.Sscanf(id, "%d", &vid); err != nil {
		return 0, utils.ErrInvalidParameters(fmt.Sprintf("invalid ID format: %s", id))
	}
	return vid, nil
}

func (g *GPUIndex) Delete(ctx context.Context, id string) error { // hnsw.go:260-289
	vid, err := parseID(id)
	if err != nil {
		return err
	}
	g.mu.Lock()
	defer g.mu.Unlock()
	n, ok := g.nodes[vid]
	if !ok {
		return utils.ErrVectorNotFound(id)
	}
	if n.deleted {
		return nil
	}
	cid := C.uint64_t(vid)
	// the device picks the new entry point when the old one is deleted (findNewEntrypoint, 617-634)
	if err := call(func() C.int32_t { return C.scn_store_mark_deleted(g.store, &cid, 1) }); err != nil {
		return err
	}
	n.deleted = true
	g.size--
	return nil
}

// ImportGraphState replaces everything (hnsw.go:749-804). Nodes are appended in ascending id order:
// a Go map has no order, and the exact scan breaks distance ties by row, so the order must be a
// function of the data (auto-assigned ids ascend in insertion order, collection.go:115-116).
func (g *GPUIndex) ImportGraphState(st core.HNSWGraphState) error {
	g.mu.Lock()
	defer g.mu.Unlock()
	if err := call(func() C.int32_t { return C.scn_store_clear(g.store) }); err != nil {
		return err
	}
	g.nodes, g.order, g.size, g.failed = map[uint64]*hostNode{}, nil, 0, nil
	ids := make([]uint64, 0, len(st.Nodes))
	for id := range st.Nodes {
		ids = append(ids, id)
	}
	sort.Slice(ids, func(a, b int) bool { return ids[a] < ids[b] })
	flat := make([]float32, 0, len(ids)*g.dim)
	cids := make([]C.uint64_t, 0, len(ids))
	lists := make([]C.int32_t, 0, len(ids))
	var counts []C.uint32_t
	var edges, dead []C.uint64_t
	for _, id := range ids {
		n := st.Nodes[id]
		flat = append(flat, n.Vector...)
		cids = append(cids, C.uint64_t(id))
		lists = append(lists, C.int32_t(len(n.Connections)))
		for _, l := range n.Connections {
			counts = append(counts, C.uint32_t(len(l)))
			for _, e := range l {
				edges = append(edges, C.uint64_t(e))
			}
		}
		if n.Deleted {
			dead = append(dead, C.uint64_t(id))
		}
		g.nodes[id] = &hostNode{elements: n.Vector, metadata: n.Metadata, deleted: n.Deleted}
	}
	g.order = ids
	g.size = st.Size // verbatim (hnsw.go:793)
	if len(ids) == 0 {
		return nil
	}
	if err := call(func() C.int32_t {
		return C.scn_store_append(g.store, (*C.float)(unsafe.Pointer(&flat[0])), &cids[0], C.uint64_t(len(cids)))
	}); err != nil {
		return err
	}
	if !g.flat {
		var pc *C.uint32_t
		var pe *C.uint64_t
		if len(counts) > 0 {
			pc = &counts[0]
		}
		if len(edges) > 0 {
			pe = &edges[0]
		}
		if err := call(func() C.int32_t {
			return C.scn_graph_upload(g.store, C.int32_t(g.params.M), C.int32_t(st.MaxLayer), C.uint64_t(st.EntryPoint),
				C.uint64_t(len(cids)), &cids[0], &lists[0], pc, pe)
		}); err != nil {
			return err
		}
	}
	if len(dead) > 0 {
		// restore form: entry point and maxLayer stay exactly as imported (hnsw.go:791-793)
		return call(func() C.int32_t { return C.scn_store_restore_deleted(g.store, &dead[0], C.uint64_t(len(dead))) })
	}
	return nil
}

// ExportGraphState rebuilds core.HNSWGraphState from the device graph (hnsw.go:703-746): vectors and
// metadata are shared with the host nodes, edges are fresh copies.
func (g *GPUIndex) ExportGraphState() core.HNSWGraphState {
	g.mu.RLock()
	defer g.mu.RUnlock()
	st := core.HNSWGraphState{Nodes: make(map[uint64]*core.HNSWNodeState, len(g.nodes)), MaxLayer: -1, Size: g.size}
	var nn, nl, ne C.uint64_t
	if g.flat || call(func() C.int32_t { return C.scn_graph_export_sizes(g.store, &nn, &nl, &ne) }) != nil || nn == 0 {
		for id, n := range g.nodes {
			st.Nodes[id] = &core.HNSWNodeState{ID: id, Vector: n.elements, Metadata: n.metadata, Deleted: n.deleted,
				Connections: [][]uint64{{}}}
		}
		return st
	}
	ids := make([]C.uint64_t, nn)
	lists := make([]C.int32_t, nn)
	counts := make([]C.uint32_t, nl+1)
	edges := make([]C.uint64_t, ne+1)
	var entry C.uint64_t
	var maxLayer C.int32_t
	if call(func() C.int32_t {
		return C.scn_graph_export(g.store, &ids[0], &lists[0], &counts[0], &edges[0], &entry, &maxLayer)
	}) != nil {
		return st
	}
	li, ei := 0, 0
	for i := range ids {
		id := uint64(ids[i])
		n := g.nodes[id]
		conn := make([][]uint64, int(lists[i]))
		for l := range conn {
			c := int(counts[li])
			li++
			conn[l] = make([]uint64, c)
			for t := 0; t < c; t++ {
				conn[l][t] = uint64(edges[ei])
				ei++
			}
		}
		st.Nodes[id] = &core.HNSWNodeState{ID: id, Vector: n.elements, Metadata: n.metadata, Deleted: n.deleted, Connections: conn}
	}
	st.EntryPoint, st.MaxLayer = uint64(entry), int(maxLayer)
	return st
}

// ---- search ------------------------------------------------------------------------------------

func (g *GPUIndex) ef(p types.SearchParams) int { // hnsw.go:300-303
	if p.EfSearch != nil && *p.EfSearch > 0 {
		return *p.EfSearch
	}
	return g.params.EfSearch
}

// decorate re-attaches Elements / Metadata from the host nodes, shared like hnsw.go:331-335 shares them.
func (g *GPUIndex) decorate(ids []C.uint64_t, dist []C.float, n int) []types.SearchResult {
	res := make([]types.SearchResult, 0, n)
	for j := 0; j < n; j++ {
		id := uint64(ids[j])
		r := types.SearchResult{Vector: types.Vector{ID: id}, Distance: float32(dist[j])}
		if hn, ok := g.nodes[id]; ok {
			r.Vector.Elements, r.Vector.Metadata = hn.elements, hn.metadata
		}
		res = append(res, r)
	}
	return res
}

// SearchBatch answers len(queries)/dim queries in one device pass. queries is row-major (any Go
// slice: pageable memory is staged through the library's pinned chunks; see AllocPinned).
func (g *GPUIndex) SearchBatch(ctx context.Context, queries []float32, p types.SearchParams) ([][]types.SearchResult, error) {
	if p.TopK <= 0 {
		return nil, utils.ErrInvalidParameters("top_k must be positive")
	}
	if len(queries) == 0 || len(queries)%g.dim != 0 {
		return nil, utils.ErrInvalidVectorDimension(fmt.Sprintf("query length %d is not a multiple of %d", len(queries), g.dim))
	}
	nq := len(queries) / g.dim
	g.mu.RLock()
	defer g.mu.RUnlock()
	if g.failed != nil {
		return nil, g.failed
	}
	k := p.TopK
	ids := make([]C.uint64_t, nq*k)
	dist := make([]C.float, nq*k)
	cnt := make([]C.uint32_t, nq)
	err := call(func() C.int32_t {
		if g.flat {
			return C.scn_search_flat(g.store, (*C.float)(unsafe.Pointer(&queries[0])), C.uint64_t(nq), C.uint32_t(k), &ids[0], &dist[0], &cnt[0])
		}
		return C.scn_search_hnsw(g.store, (*C.float)(unsafe.Pointer(&queries[0])), C.uint64_t(nq), C.uint32_t(k),
			C.uint32_t(g.ef(p)), &ids[0], &dist[0], &cnt[0])
	})
	if err != nil {
		return nil, err
	}
	out := make([][]types.SearchResult, nq)
	for q := 0; q < nq; q++ {
		out[q] = g.decorate(ids[q*k:], dist[q*k:], int(cnt[q]))
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
	defer g.mu.RUnlock()
	if g.failed != nil {
		return nil, g.failed
	}
	k := p.TopK
	ids := make([]C.uint64_t, k)
	dist := make([]C.float, k)
	var cnt C.uint32_t
	if err := call(func() C.int32_t {
		return C.scn_batcher_search(g.batcher, (*C.float)(unsafe.Pointer(&query[0])), C.uint32_t(k), C.uint32_t(g.ef(p)),
			&ids[0], &dist[0], &cnt)
	}); err != nil {
		return nil, err
	}
	return g.decorate(ids, dist, int(cnt)), nil
}

// SearchExact is the flat ground truth over the same rows.
func (g *GPUIndex) SearchExact(ctx context.Context, queries []float32, topK int) ([]uint64, []float32, error) {
	nq := len(queries) / g.dim
	ids := make([]C.uint64_t, nq*topK)
	dist := make([]C.float, nq*topK)
	g.mu.RLock()
	defer g.mu.RUnlock()
	if err := call(func() C.int32_t {
		return C.scn_search_flat(g.store, (*C.float)(unsafe.Pointer(&queries[0])), C.uint64_t(nq), C.uint32_t(topK), &ids[0], &dist[0], nil)
	}); err != nil {
		return nil, nil, err
	}
	oi := make([]uint64, len(ids))
	od := make([]float32, len(dist))
	for i := range ids {
		oi[i], od[i] = uint64(ids[i]), float32(dist[i])
	}
	return oi, od, nil
}

// ---- read side of core.VectorIndex / core.HNSWIndex ------------------------------------------------

func (g *GPUIndex) Get(ctx context.Context, id string) (*types.Vector, error) { // hnsw.go:353-373
	vid, err := parseID(id)
	if err != nil {
		return nil, err
	}
	g.mu.RLock()
	defer g.mu.RUnlock()
	n, ok := g.nodes[vid]
	if !ok || n.deleted {
		return nil, utils.ErrVectorNotFound(id)
	}
	return &types.Vector{ID: vid, Elements: n.elements, Metadata: n.metadata}, nil
}

func (g *GPUIndex) Size() int {
	g.mu.RLock()
	defer g.mu.RUnlock()
	return g.size
}

func (g *GPUIndex) stats() C.scn_stats {
	var st C.scn_stats
	C.scn_store_stats(g.store, &st)
	return st
}

func (g *GPUIndex) MemoryUsage() int64 {
	g.mu.RLock()
	defer g.mu.RUnlock()
	return int64(g.stats().device_bytes)
}

func (g *GPUIndex) GetParameters() types.HNSWParams { return g.params }

func (g *GPUIndex) SetEfSearch(ef int) { // hnsw.go:449-453
	g.mu.Lock()
	defer g.mu.Unlock()
	g.params.EfSearch = ef
}

func (g *GPUIndex) GetLayers() int { // hnsw.go:394-401
	g.mu.RLock()
	defer g.mu.RUnlock()
	if ml := int(g.stats().max_layer); ml >= 0 {
		return ml + 1
	}
	return 0
}

func (g *GPUIndex) GetGraphStatistics() types.GraphStats { // hnsw.go:404-443 (MaxDegree is not tracked on the device)
	g.mu.RLock()
	defer g.mu.RUnlock()
	st := g.stats()
	avg := 0.0
	if g.size > 0 {
		avg = float64(st.graph_edges) / float64(g.size)
	}
	return types.GraphStats{Layers: int(st.max_layer) + 1, Nodes: g.size, Connections: int(st.graph_edges), AvgDegree: avg,
		MemoryUsage: int64(st.device_bytes)}
}

func (g *GPUIndex) GetStatistics() interface{} { return g.GetGraphStatistics() }

// Compact is the device half of Collection.Compact (collection.go:283-313): the deleted rows leave
// HBM for good; the caller then calls Build with the surviving vectors, exactly as it does with the
// CPU index (collection.go:299-311).
func (g *GPUIndex) Compact() (int, error) {
	g.mu.Lock()
	defer g.mu.Unlock()
	var removed C.uint64_t
	err := call(func() C.int32_t { return C.scn_store_compact(g.store, &removed) })
	return int(removed), err
}

// ---- one flat collection over several GPUs (single process) -------------------------------------------

// ShardedFlat is the exact-scan index row-sharded over the GPUs of the box: Search stays ONE blocking
// call (interfaces.go:98) and drives all of them (scn_shards_*; SURVEY.md §8b/§8e).
type ShardedFlat struct {
	mu     sync.RWMutex
	shards *C.scn_shards
	dim    int
}

func NewShardedFlat(devices []int, dim int, metric types.DistanceMetric, capacityRows uint64) (*ShardedFlat, error) {
	devs := make([]C.int32_t, len(devices))
	for i, d := range devices {
		devs[i] = C.int32_t(d)
	}
	s := &ShardedFlat{dim: dim}
	if err := call(func() C.int32_t {
		return C.scn_shards_create(&devs[0], C.int32_t(len(devs)), C.uint32_t(dim), C.int32_t(metric), C.uint64_t(capacityRows), &s.shards)
	}); err != nil {
		return nil, err
	}
	runtime.SetFinalizer(s, func(s *ShardedFlat) { s.Close() })
	return s, nil
}

func (s *ShardedFlat) Close() {
	s.mu.Lock()
	defer s.mu.Unlock()
	if s.shards != nil {
		C.scn_shards_destroy(s.shards)
		s.shards = nil
	}
}

// Append adds len(ids) row-major vectors in insertion order (ids nil: global row + 1).
func (s *ShardedFlat) Append(vecs []float32, ids []uint64) error {
	s.mu.Lock()
	defer s.mu.Unlock()
	n := len(vecs) / s.dim
	var pi *C.uint64_t
	if ids != nil {
		pi = (*C.uint64_t)(unsafe.Pointer(&ids[0]))
	}
	return call(func() C.int32_t {
		return C.scn_shards_append(s.shards, (*C.float)(unsafe.Pointer(&vecs[0])), pi, C.uint64_t(n))
	})
}

func (s *ShardedFlat) Delete(id uint64) error {
	s.mu.Lock()
	defer s.mu.Unlock()
	cid := C.uint64_t(id)
	return call(func() C.int32_t { return C.scn_shards_mark_deleted(s.shards, &cid, 1) })
}

// SearchBatch: ids / distances are [nq][k] (0 / +Inf padded), counts the valid entries per query.
func (s *ShardedFlat) SearchBatch(queries []float32, k int) ([]uint64, []float32, []uint32, error) {
	nq := len(queries) / s.dim
	ids := make([]uint64, nq*k)
	dist := make([]float32, nq*k)
	cnt := make([]uint32, nq)
	s.mu.RLock()
	defer s.mu.RUnlock()
	err := call(func() C.int32_t {
		return C.scn_shards_search_flat(s.shards, (*C.float)(unsafe.Pointer(&queries[0])), C.uint64_t(nq), C.uint32_t(k),
			(*C.uint64_t)(unsafe.Pointer(&ids[0])), (*C.float)(unsafe.Pointer(&dist[0])), (*C.uint32_t)(unsafe.Pointer(&cnt[0])))
	})
	return ids, dist, cnt, err
}

// AllocPinned returns a page-locked []float32 of n elements (query batches that are reused should live
// in one: the library DMAs pinned buffers directly) and the function that frees it.
func AllocPinned(n int) ([]float32, func(), error) {
	var p unsafe.Pointer
	if err := call(func() C.int32_t { return C.scn_host_alloc(C.uint64_t(n*4), &p) }); err != nil {
		return nil, nil, err
	}
	return unsafe.Slice((*float32)(p), n), func() { C.scn_host_free(p) }, nil
}

// ---- DistanceCalculator and the vector helpers ---------------------------------------------------------

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
	return call(func() C.int32_t {
		return C.scn_vector_ops(C.int32_t(device), op, (*C.float)(unsafe.Pointer(&a[0])), pb, C.uint64_t(n),
			C.uint32_t(dim), (*C.float)(unsafe.Pointer(&out[0])))
	})
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
	if v, ok := cfg.Parameters["max_layers"].(int); ok {
		params.MaxLayers = v
	}
	if v, ok := cfg.Parameters["seed"].(int64); ok {
		params.Seed = v
	}
	return New(params, cfg.Metric, dim, f.Device, cfg.Type == "flat-gpu")
}
func (f Factory) SupportedMetrics() []types.DistanceMetric {
	return []types.DistanceMetric{types.DistanceMetricL2, types.DistanceMetricCosine, types.DistanceMetricInnerProduct}
}
func (f Factory) DefaultParameters() map[string]interface{} {
	return map[string]interface{}{"m": 16, "ef_construction": 200, "ef_search": 50, "max_layers": 16}
}
