// TEST INFRASTRUCTURE (oracle/) — never linked into the product library.
//
// extern "C" window onto the reference's own DiskANN search primitives, compiled from the sources where they lie
// (third_party/DiskANN/src/pq.cpp, logger.cpp, ann_exception.cpp; include/neighbor.h is header-only):
//   * diskann::NeighborPriorityQueue / diskann::Neighbor          (include/neighbor.h:14-152)
//   * diskann::FixedChunkPQTable::{load_pq_centroid_bin, preprocess_query, populate_chunk_distances}
//     and diskann::aggregate_coords / pq_dist_lookup               (src/pq.cpp:49-219, 302-340)
// The search loop itself (src/pq_flash_index.cpp) needs libaio / protobuf / Boost and is NOT built; see the header
// of vamana_oracle.c.  tests/test_vamana_oracle.py drives these against the C restatement.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mkl.h"
#include "neighbor.h"
#include "pq.h"

// training-only dependencies of pq.cpp, never reached from the query functions
#define LB2_UNREACHABLE(name) { std::fprintf(stderr, "diskann_ref_harness: %s reached\n", name); std::abort(); }
extern "C" {
void cblas_sgemm(CBLAS_LAYOUT, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, MKL_INT, MKL_INT, MKL_INT, float, const float*, MKL_INT,
                 const float*, MKL_INT, float, float*, MKL_INT) LB2_UNREACHABLE("cblas_sgemm")
int LAPACKE_sgesdd(int, char, MKL_INT, MKL_INT, float*, MKL_INT, float*, float*, MKL_INT, float*, MKL_INT) LB2_UNREACHABLE("LAPACKE_sgesdd")
}
namespace math_utils {
void compute_closest_centers(float*, size_t, size_t, float*, size_t, size_t, uint32_t*, std::vector<size_t>*, float*)
    LB2_UNREACHABLE("compute_closest_centers")
}
namespace kmeans {
float run_lloyds(float*, size_t, size_t, float*, const size_t, const size_t, std::vector<size_t>*, uint32_t*) LB2_UNREACHABLE("run_lloyds")
void kmeanspp_selecting_pivots(float*, size_t, size_t, float*, size_t) LB2_UNREACHABLE("kmeanspp_selecting_pivots")
}

template <typename T>
void gen_random_slice(const std::string, double, float*&, size_t&, size_t&) LB2_UNREACHABLE("gen_random_slice")
template void gen_random_slice<uint8_t>(const std::string, double, float*&, size_t&, size_t&);
template void gen_random_slice<int8_t>(const std::string, double, float*&, size_t&, size_t&);
template void gen_random_slice<float>(const std::string, double, float*&, size_t&, size_t&);

extern "C" {

// ---- NeighborPriorityQueue
void* dref_npq_new(size_t capacity) { return new diskann::NeighborPriorityQueue(capacity); }
void dref_npq_free(void* q) { delete static_cast<diskann::NeighborPriorityQueue*>(q); }
void dref_npq_insert(void* q, uint32_t id, float dist) {
    static_cast<diskann::NeighborPriorityQueue*>(q)->insert(diskann::Neighbor(id, dist));
}
void dref_npq_closest_unexpanded(void* q, uint32_t* id, float* dist) {
    diskann::Neighbor n = static_cast<diskann::NeighborPriorityQueue*>(q)->closest_unexpanded();
    *id = n.id; *dist = n.distance;
}
int dref_npq_has_unexpanded(void* q) { return static_cast<diskann::NeighborPriorityQueue*>(q)->has_unexpanded_node() ? 1 : 0; }
size_t dref_npq_size(void* q) { return static_cast<diskann::NeighborPriorityQueue*>(q)->size(); }
void dref_npq_get(void* q, size_t i, uint32_t* id, float* dist, int* expanded) {
    diskann::Neighbor n = (*static_cast<diskann::NeighborPriorityQueue*>(q))[i];
    *id = n.id; *dist = n.distance; *expanded = n.expanded ? 1 : 0;
}

// ---- FixedChunkPQTable
void* dref_pq_load(const char* pivots_path, size_t n_chunks) {
    auto* t = new diskann::FixedChunkPQTable();
    try { t->load_pq_centroid_bin(pivots_path, n_chunks); } catch (...) { delete t; return nullptr; }
    return t;
}
void dref_pq_free(void* t) { delete static_cast<diskann::FixedChunkPQTable*>(t); }
uint32_t dref_pq_num_chunks(void* t) { return static_cast<diskann::FixedChunkPQTable*>(t)->get_num_chunks(); }
void dref_pq_preprocess_query(void* t, float* q) { static_cast<diskann::FixedChunkPQTable*>(t)->preprocess_query(q); }
void dref_pq_populate_chunk_distances(void* t, const float* q, float* lut) {
    static_cast<diskann::FixedChunkPQTable*>(t)->populate_chunk_distances(q, lut);
}
void dref_pq_lookup(const uint32_t* ids, uint64_t n_ids, const uint8_t* all_codes, uint64_t n_chunks, const float* lut,
                    uint8_t* scratch, float* out) {
    diskann::aggregate_coords(ids, n_ids, all_codes, n_chunks, scratch);
    diskann::pq_dist_lookup(scratch, n_ids, n_chunks, lut, out);
}

}  // extern "C"
