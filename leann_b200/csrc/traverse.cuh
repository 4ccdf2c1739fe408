// Device-side data model of the traversal stage (K1).
#pragma once
#include "common.cuh"

namespace lb2 {

// Compact-CSR HNSW graph resident in HBM.  Same three arrays as the reference's
// HNSW::compact_node_offsets / compact_level_ptr / compact_neighbors_data
// (faiss/impl/HNSW.h:194-197), offsets narrowed to what the sizes allow.
struct DevGraph {
    int64_t ntotal = 0;
    int d = 0;
    int metric_ip = 1;
    int entry_point = -1;
    int max_level = -1;
    int maxdeg0 = 0, maxdeg_up = 0;
    const uint64_t* node_offsets = nullptr;  // [ntotal+1]
    const uint64_t* level_ptr = nullptr;     // [sum(levels)+ntotal]
    const int32_t* nbrs = nullptr;           // [nnz]
};

enum Phase : int { PH_IDLE = 0, PH_FETCH = 1, PH_ENTRY = 2, PH_GREEDY = 3, PH_BASE_INIT = 4, PH_BASE = 5 };

// Search parameters, mirroring faiss::SearchParametersHNSW (faiss/impl/HNSW.h:54-73)
struct TravParams {
    int ef;         // efSearch
    int hcap;       // candidate heap capacity = max(efSearch, k)   (HNSW.cpp:1117)
    int k;
    int beam;       // beam_size
    int batch_size; // batch_size (0 = disabled)
    int check_rel;  // check_relative_distance
    int cap_req;    // max ids one hop can request
    int p2;         // next power of two >= cap_new (sort buffer)
    int cap_new;    // max unvisited neighbours one hop can gather (= cap_req without PQ pruning)
    // PQ-guided pruning (faiss/impl/HNSW_search.cpp:442-465, 645-750): 0 off, 1 global queue, 2 local, 3 proportional
    int pq_mode;
    float pq_ratio; // pq_select_ratio = 1 - pq_pruning_ratio
};

// PQ pruning data: PQPrunerDataLoader (faiss/impl/pq.h:12-43) + HNSW::pq_codes (HNSW_search.cpp:253-297), device copies
struct PqDev {
    const float* tables_tr = nullptr;        // [ndims, 256]
    const float* centroid = nullptr;         // [ndims]
    const uint32_t* chunk_offsets = nullptr; // [n_chunks + 1]
    const uint8_t* codes = nullptr;          // [ntotal, n_chunks]
    int ndims = 0, n_chunks = 0;
};

// Per-slot traversal state (one slot = one in-flight query, owned by one warp) + batch plumbing.
struct TravState {
    int S = 0;
    // slot scalars
    int* phase; int* qid; int* level; int* nearest; int* prev_nearest; float* d_nearest;
    int* hk; int* hnvalid; int* nstep; int* pend_beam; int* n_req;
    long long* ndis; long long* nhops;
    // slot arrays
    int* heap_ids; float* heap_dis;   // [S, hcap]
    int* res_ids; float* res_dis;     // [S, k]
    int* req_ids;                     // [S, cap_req]
    uint32_t* visited;                // [S, vis_words]
    int64_t vis_words;
    // batch
    const float* queries;  // [nq, d] device
    int64_t nq;
    float* outD; int64_t* outI; long long* out_ndis; long long* out_nhops;
    int* next_query; int* n_done;
    // scoring source
    int recompute;              // 1: fresh embeddings E (recompute stage); 0: stored vectors
    const float* vectors;       // [ntotal, d] (stored mode)
    const float* E;             // [cap_unique, d] embeddings of the previous hop's unique nodes
    // recompute-stage work list produced by each hop (deduplicated across queries)
    uint32_t* stamp;            // [ntotal] last hop (epoch) that requested the node
    int* slot_of[2];            // [ntotal] x2 (hop parity): row of the node in E / uniq_node
    unsigned long long* claim;  // packed (n_unique << 40 | n_tokens)
    int* uniq_node;             // [cap_unique]
    int* seq_start;             // [cap_unique] first packed token row of the passage
    const uint64_t* tok_off;    // [ntotal+1] passage token offsets
    int max_pos;
    uint32_t epoch;             // hop counter (>= 1)
    // de-duplication scope of the recompute work list:
    //   hop  (call_scope = 0): a node requested by several queries in the SAME hop is encoded once;
    //   call (call_scope = 1): ... once per search call — later hops of any query reuse the row (E keeps
    //        one row per distinct node for the duration of the call; nothing persists across calls).
    int call_scope;
    uint32_t stamp_value;       // value written to stamp[]: hop epoch (hop scope) or call epoch (call scope)
    int row_base_hop;           // first E row of this hop's new nodes (0 in hop scope)
    // PQ pruning state (allocated only when a search asks for pruning)
    PqDev pq;
    float* pq_lut;              // [S, n_chunks * 256] per-query distance tables
    float* pq_qprep;            // [S, ndims] preprocessed query (scratch of the table build)
    float* pq_qd; int* pq_qid;  // [S, pq_cap] the query's PQ candidates, ascending (distance, id): the content of the
                                //   reference's pq_candidate_queue (global mode) / pq_candidate_set (proportional mode)
    int* pq_qn; int* pq_qhead;  // [S] valid range [head, n)
    int pq_cap;
    int* error;                 // != 0: a slot overflowed its PQ candidate array
};

size_t step_smem_bytes(const TravParams& p, int d, int warps);
bool launch_step(cudaStream_t s, const DevGraph& g, const TravParams& p, const TravState& st, int max_iters,
                 int num_sms);
bool launch_init_slots(cudaStream_t s, const TravState& st);

}  // namespace lb2
