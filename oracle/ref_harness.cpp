// TEST INFRASTRUCTURE (oracle/) — never linked into the product library.
//
// Thin extern "C" harness around the UNMODIFIED reference traversal, compiled
// from the sources where they lie under /root/reference (see oracle/Makefile):
//   third_party/faiss/faiss/impl/HNSW.cpp         (HNSW::search :1081-1155,
//        greedy_update_nearest :1012-1063, MinimaxHeap :1263-1509, add_with_locks)
//   third_party/faiss/faiss/impl/HNSW_search.cpp  (search_from_candidates :369-850,
//        neighbor_range :299-362, fetch_neighbors :141-251)
// The harness replaces only what the reference itself treats as pluggable:
//   * the DistanceComputer (reference: ZmqDistanceComputer, impl/HNSW_zmq.h:26-153,
//     which ships ids to a Python process) becomes an in-process computer —
//     either over stored fp32 vectors with the canonical summation order of
//     oracle/canon_dist.h, or a C callback (the recompute oracle: a Python
//     fp32 BertModel forward), exactly the seam `distances_batch` exposes;
//   * the per-query driver loop of IndexHNSW.cpp:318-439 (hnsw_search /
//     IndexHNSW::search), restated here because that file pulls in the whole
//     faiss library + libzmq; the loop is: heapify results, set_query,
//     HNSW::search, reorder, negate distances for IP.
// Nothing here is copied from the reference; it only calls its public API.
#include <faiss/impl/AuxIndexStructures.h>
#include <faiss/impl/DistanceComputer.h>
#include <faiss/impl/HNSW.h>
#include <faiss/impl/ResultHandler.h>

#include <omp.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

#include "canon_dist.h"

using faiss::HNSW;
using faiss::idx_t;

extern "C" {
// Never called on this path (faiss::rand_smooth_vectors only); keeps the
// link free of a BLAS dependency.
int sgemm_(const char*, const char*, int*, int*, int*, const float*, const float*,
           int*, const float*, int*, float*, float*, int*) {
    fprintf(stderr, "ref_harness: unexpected sgemm_ call\n");
    abort();
}
typedef void (*ref_dist_cb)(void* ctx, const float* q, int64_t n, const int64_t* ids, float* out);
}

namespace {

struct FlatDC : faiss::DistanceComputer {
    const float* base;
    const float* q = nullptr;
    int d;
    bool ip;
    FlatDC(const float* base, int d, bool ip) : base(base), d(d), ip(ip) {}
    void set_query(const float* x) override { q = x; }
    const float* get_query() override { return q; }
    float operator()(idx_t i) override {
        const float* v = base + (size_t)i * d;
        return ip ? -lb2_canon_ip(q, v, d) : lb2_canon_l2(q, v, d);
    }
    float symmetric_dis(idx_t i, idx_t j) override {
        const float* a = base + (size_t)i * d;
        const float* b = base + (size_t)j * d;
        return ip ? -lb2_canon_ip(a, b, d) : lb2_canon_l2(a, b, d);
    }
};

struct CallbackDC : faiss::DistanceComputer {
    ref_dist_cb cb;
    void* ctx;
    const float* q = nullptr;
    CallbackDC(ref_dist_cb cb, void* ctx) : cb(cb), ctx(ctx) {}
    void set_query(const float* x) override { q = x; }
    const float* get_query() override { return q; }
    float operator()(idx_t i) override {
        float out;
        int64_t id = i;
        cb(ctx, q, 1, &id, &out);
        return out;
    }
    void distances_batch(const std::vector<idx_t>& ids, std::vector<float>& out) override {
        if (ids.empty()) return;
        static_assert(sizeof(idx_t) == sizeof(int64_t), "idx_t");
        cb(ctx, q, (int64_t)ids.size(), (const int64_t*)ids.data(), out.data());
    }
    float symmetric_dis(idx_t, idx_t) override { return 0.f; }
};

struct RefIndex {
    int d;
    bool ip;
    HNSW hnsw;
    const float* vectors = nullptr;  // borrowed
    RefIndex(int d, int M, bool ip) : d(d), ip(ip), hnsw(M) {}
};

}  // namespace

extern "C" {

void* ref_new(int d, int M, int metric_is_ip) {
    return new RefIndex(d, M, metric_is_ip != 0);
}

void ref_free(void* h) { delete (RefIndex*)h; }

void ref_set_vectors(void* h, const float* x) { ((RefIndex*)h)->vectors = x; }

// Graph construction through the reference's HNSW::add_with_locks, highest
// level first (the order hnsw_add_vertices uses, IndexHNSW.cpp:59-280), ids
// ascending inside a level.  nthreads==1 gives a deterministic graph.
int ref_build(void* h, int64_t n, const float* x, int efConstruction, int nthreads) {
    RefIndex* R = (RefIndex*)h;
    R->vectors = x;
    HNSW& hnsw = R->hnsw;
    hnsw.efConstruction = efConstruction;
    int max_level = hnsw.prepare_level_tab((size_t)n, false);
    std::vector<omp_lock_t> locks(n);
    for (auto& l : locks) omp_init_lock(&l);
    std::vector<std::vector<int>> by_level(max_level + 1);
    for (int64_t i = 0; i < n; i++) by_level[hnsw.levels[i] - 1].push_back((int)i);
    if (nthreads < 1) nthreads = 1;
    for (int lvl = max_level; lvl >= 0; lvl--) {
        const std::vector<int>& pts = by_level[lvl];
        // per-node degree cap used by add_links_starting_from (shrink_neighbor_list):
        // set as hnsw_add_vertices does, IndexHNSW.cpp:130-140 (M0 on level 0, unbounded above)
        hnsw.ems = std::vector<int>((size_t)n, lvl == 0 ? hnsw.nb_neighbors(0)
                                                        : std::numeric_limits<int>::max());
#pragma omp parallel num_threads(nthreads) if (nthreads > 1)
        {
            faiss::VisitedTable vt((int)n);
            FlatDC dc(x, R->d, R->ip);
#pragma omp for schedule(static, 1)
            for (int64_t j = 0; j < (int64_t)pts.size(); j++) {
                int pt = pts[j];
                dc.set_query(x + (size_t)pt * R->d);
                hnsw.add_with_locks(dc, lvl, pt, locks, vt, false);
            }
        }
    }
    for (auto& l : locks) omp_destroy_lock(&l);
    return 0;
}

int64_t ref_ntotal(void* h) { return (int64_t)((RefIndex*)h)->hnsw.levels.size(); }
int ref_max_level(void* h) { return ((RefIndex*)h)->hnsw.max_level; }
int ref_entry_point(void* h) { return ((RefIndex*)h)->hnsw.entry_point; }
int ref_ef_construction(void* h) { return ((RefIndex*)h)->hnsw.efConstruction; }
int ref_ef_search(void* h) { return ((RefIndex*)h)->hnsw.efSearch; }
int64_t ref_neighbors_size(void* h) { return (int64_t)((RefIndex*)h)->hnsw.neighbors.size(); }
const int* ref_neighbors(void* h) { return ((RefIndex*)h)->hnsw.neighbors.data(); }
const size_t* ref_offsets(void* h) { return ((RefIndex*)h)->hnsw.offsets.data(); }
const int* ref_levels(void* h) { return ((RefIndex*)h)->hnsw.levels.data(); }
int ref_n_cum(void* h) { return (int)((RefIndex*)h)->hnsw.cum_nneighbor_per_level.size(); }
const int* ref_cum(void* h) { return ((RefIndex*)h)->hnsw.cum_nneighbor_per_level.data(); }
int ref_n_probas(void* h) { return (int)((RefIndex*)h)->hnsw.assign_probas.size(); }
const double* ref_probas(void* h) { return ((RefIndex*)h)->hnsw.assign_probas.data(); }

// Install a compact-CSR graph (the arrays of the reference's .index file,
// index_read.cpp:523-813) so the reference's own neighbor_range /
// fetch_neighbors CSR branches are the ones exercised.
void* ref_from_csr(int d, int M, int metric_is_ip, int64_t ntotal, const int* levels,
                   int64_t n_level_ptr, const uint64_t* level_ptr, const uint64_t* node_offsets,
                   int64_t n_nbrs, const int* nbrs, int entry_point, int max_level) {
    RefIndex* R = new RefIndex(d, M, metric_is_ip != 0);
    HNSW& g = R->hnsw;
    g.levels.assign(levels, levels + ntotal);
    g.storage_is_compact = true;
    g.compact_level_ptr = faiss::MaybeOwnedVector<size_t>(
            std::vector<size_t>(level_ptr, level_ptr + n_level_ptr));
    g.compact_node_offsets = faiss::MaybeOwnedVector<size_t>(
            std::vector<size_t>(node_offsets, node_offsets + ntotal + 1));
    g.compact_neighbors_data = faiss::MaybeOwnedVector<HNSW::storage_idx_t>(
            std::vector<HNSW::storage_idx_t>(nbrs, nbrs + n_nbrs));
    g.entry_point = entry_point;
    g.max_level = max_level;
    return R;
}

// Per-query restatement of hnsw_search()/IndexHNSW::search (IndexHNSW.cpp:318-439).
// ndis/nhops are the reference's own HNSWStats counters, reported per query.
int ref_search(void* h, int64_t nq, const float* q, int64_t k, int ef, int beam_size,
               int batch_size, int check_relative_distance, float* D, int64_t* I,
               int64_t* ndis, int64_t* nhops, ref_dist_cb cb, void* ctx, int nthreads) {
    RefIndex* R = (RefIndex*)h;
    const HNSW& hnsw = R->hnsw;
    if (!cb && !R->vectors) return -1;
    faiss::SearchParametersHNSW params;
    params.efSearch = ef;
    params.beam_size = beam_size;
    params.batch_size = batch_size;
    params.check_relative_distance = check_relative_distance != 0;
    using RH = faiss::HeapBlockResultHandler<HNSW::C>;
    RH bres((size_t)nq, D, I, (size_t)k);
    if (nthreads < 1) nthreads = 1;
    int64_t ntotal = (int64_t)hnsw.levels.size();
#pragma omp parallel num_threads(nthreads) if (nthreads > 1 && nq > 1)
    {
        faiss::VisitedTable vt((int)ntotal);
        RH::SingleResultHandler res(bres);
        std::unique_ptr<faiss::DistanceComputer> dc;
        if (cb)
            dc.reset(new CallbackDC(cb, ctx));
        else
            dc.reset(new FlatDC(R->vectors, R->d, R->ip));
#pragma omp for schedule(dynamic, 1)
        for (int64_t i = 0; i < nq; i++) {
            res.begin((size_t)i);
            dc->set_query(q + (size_t)i * R->d);
            faiss::HNSWStats st = hnsw.search(*dc, res, vt, &params, nullptr);
            if (ndis) ndis[i] = (int64_t)st.ndis;
            if (nhops) nhops[i] = (int64_t)st.nhops;
            res.end();
        }
    }
    if (R->ip) {
        for (int64_t i = 0; i < nq * k; i++) D[i] = -D[i];
    }
    return 0;
}

// PQ-guided pruning (impl/HNSW_search.cpp:442-465, 645-750): the reference only prunes after
// HNSW::load_pq_pruning_data (:253-297) has read DiskANN-style pivots + compressed codes.
int ref_load_pq(void* h, const char* pivots_path, const char* compressed_path) {
    return ((RefIndex*)h)->hnsw.load_pq_pruning_data(pivots_path, compressed_path) ? 0 : -1;
}

// ref_search with the three pruning knobs of SearchParametersHNSW (impl/HNSW.h:54-73)
int ref_search_pq(void* h, int64_t nq, const float* q, int64_t k, int ef, int beam_size, int batch_size,
                  int check_relative_distance, float pq_pruning_ratio, int local_prune, float send_neigh_times_ratio,
                  float* D, int64_t* I, int64_t* ndis, int64_t* nhops, ref_dist_cb cb, void* ctx, int nthreads) {
    RefIndex* R = (RefIndex*)h;
    const HNSW& hnsw = R->hnsw;
    if (!cb && !R->vectors) return -1;
    faiss::SearchParametersHNSW params;
    params.efSearch = ef;
    params.beam_size = beam_size;
    params.batch_size = batch_size;
    params.check_relative_distance = check_relative_distance != 0;
    params.pq_pruning_ratio = pq_pruning_ratio;
    params.local_prune = local_prune != 0;
    params.send_neigh_times_ratio = send_neigh_times_ratio;
    using RH = faiss::HeapBlockResultHandler<HNSW::C>;
    RH bres((size_t)nq, D, I, (size_t)k);
    if (nthreads < 1) nthreads = 1;
    int64_t ntotal = (int64_t)hnsw.levels.size();
    int rc = 0;
#pragma omp parallel num_threads(nthreads) if (nthreads > 1 && nq > 1)
    {
        faiss::VisitedTable vt((int)ntotal);
        RH::SingleResultHandler res(bres);
        std::unique_ptr<faiss::DistanceComputer> dc;
        if (cb)
            dc.reset(new CallbackDC(cb, ctx));
        else
            dc.reset(new FlatDC(R->vectors, R->d, R->ip));
#pragma omp for schedule(dynamic, 1)
        for (int64_t i = 0; i < nq; i++) {
            res.begin((size_t)i);
            dc->set_query(q + (size_t)i * R->d);
            try {
                faiss::HNSWStats st = hnsw.search(*dc, res, vt, &params, nullptr);
                if (ndis) ndis[i] = (int64_t)st.ndis;
                if (nhops) nhops[i] = (int64_t)st.nhops;
            } catch (const std::exception& e) {
                fprintf(stderr, "ref_search_pq: %s\n", e.what());
                rc = -2;
            }
            res.end();
        }
    }
    if (R->ip) {
        for (int64_t i = 0; i < nq * k; i++) D[i] = -D[i];
    }
    return rc;
}

// MinimaxHeap exports (impl/HNSW.cpp:1263-1509) for unit-level pinning of the oracle's heap.
void* ref_mmh_new(int n) { return new HNSW::MinimaxHeap(n); }
void ref_mmh_free(void* p) { delete (HNSW::MinimaxHeap*)p; }
void ref_mmh_push(void* p, int i, float v) { ((HNSW::MinimaxHeap*)p)->push(i, v); }
int ref_mmh_size(void* p) { return ((HNSW::MinimaxHeap*)p)->size(); }
int ref_mmh_pop_min(void* p, float* v) { return ((HNSW::MinimaxHeap*)p)->pop_min(v); }
int ref_mmh_count_below(void* p, float t) { return ((HNSW::MinimaxHeap*)p)->count_below(t); }

}  // extern "C"
