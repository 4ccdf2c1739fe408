// Device-side data model of the DiskANN/Vamana search path (PQ beam search + one deferred re-rank).
#pragma once
#include "common.cuh"

namespace lb2 {

enum { VAM_L2 = 0, VAM_MIPS = 1, VAM_COSINE = 2 };
enum { VAM_DEFERRED_FETCH = 1, VAM_SKIP_SEARCH_REORDER = 2 };

// Index resident in HBM.  Same content as PQFlashIndex keeps in RAM / reads from disk
// (DiskANN/include/pq_flash_index.h): PQ codes + transposed pivot tables, adjacency as a padded [n, R]
// matrix (a row = one coalesced 4R-byte read instead of a 4 KB sector), optional full-precision coordinates.
struct DevVamana {
    int64_t n = 0;
    int data_dim = 0;  // PQ ndims / coordinate count (raw dim + 1 for MIPS)
    int raw_dim = 0;   // embedding dimension the caller's queries and the encoder have
    int metric = 0;
    int R = 0;
    int n_chunks = 0;
    int n_medoids = 0;
    const int32_t* nbrs = nullptr;          // [n, R] valid first, -1 padded
    const uint8_t* codes = nullptr;         // [n, n_chunks]
    const float* tables_tr = nullptr;       // [data_dim, 256]
    const float* centroid = nullptr;        // [data_dim]
    const uint32_t* chunk_offsets = nullptr;
    const uint32_t* medoids = nullptr;
    const float* centroid_data = nullptr;   // [n_medoids, data_dim] or null
    const float* coords = nullptr;          // [n, data_dim] or null
    float max_base_norm = 0.f;
};

// Buffers of one search call (a wave of queries).
struct VamanaWork {
    int64_t nq = 0;
    int L = 64, beam = 1, k = 10, cap_full = 0, flags = 0;
    uint32_t io_limit = 0xffffffffu;
    const float* queries = nullptr;  // [nq, raw_dim]
    float* aq = nullptr;             // [nq, data_dim] normalised / extended query (aligned_query_T)
    float* qnorm = nullptr;          // [nq]
    float* qrot = nullptr;           // [nq, data_dim] centred query
    float* lut = nullptr;            // [nq, n_chunks, 256]
    uint32_t* visited = nullptr;     // [slots, vis_words]
    int64_t vis_words = 0;
    int slots = 0;
    int warps = 8;                   // warps per block of the traversal kernel
    uint32_t* full_ids = nullptr;    // [nq, cap_full] expanded nodes in expansion order (full_retset)
    float* full_dist = nullptr;      // [nq, cap_full]
    int* n_full = nullptr;           // [nq]
    long long* cmps = nullptr;       // [nq]
    long long* hops = nullptr;       // [nq]
    int* next_query = nullptr;
    int* error_flag = nullptr;
    // deferred fetch: de-duplicated work list for the encoder
    uint32_t* stamp = nullptr;       // [n]
    int* slot_of = nullptr;          // [n] row of the node in E
    unsigned long long* claim = nullptr;  // packed (n_unique << 40 | n_tokens)
    int* uniq_node = nullptr;
    int* seq_start = nullptr;
    const uint64_t* tok_off = nullptr;
    int max_pos = 0;
    uint32_t call_epoch = 0;
    const float* E = nullptr;        // [n_unique, raw_dim]
    float* outD = nullptr;
    int64_t* outI = nullptr;
};

void vamana_plan(const DevVamana& v, VamanaWork& w, int num_sms);  // sets w.slots (queries in flight, one warp each) and w.warps
bool vamana_launch_prepare(cudaStream_t s, const DevVamana& v, const VamanaWork& w);
bool vamana_launch_search(cudaStream_t s, const DevVamana& v, const VamanaWork& w, int num_sms);
bool vamana_launch_collect(cudaStream_t s, const DevVamana& v, const VamanaWork& w);
bool vamana_launch_rerank(cudaStream_t s, const DevVamana& v, const VamanaWork& w);

}  // namespace lb2
