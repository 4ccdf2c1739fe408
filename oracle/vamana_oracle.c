/* TEST INFRASTRUCTURE (oracle/) — never linked into the product library.
 *
 * Plain-C restatement of the reference's DiskANN/Vamana query path as LEANN drives it:
 *   DiskannSearcher.search (leann-backend-diskann/leann_backend_diskann/diskann_backend.py:383-472) always calls
 *   StaticDiskIndex::batch_search (third_party/DiskANN/python/src/static_disk_index.cpp:88-118) with
 *   recompute_neighbors = false, so traversal runs on PQ distances only and, when recompute_embeddings is on
 *   (USE_DEFERRED_FETCH), the expanded nodes are re-scored once at the end from freshly computed embeddings.
 * All file:line citations are under /root/reference/packages/leann-backend-diskann/third_party/DiskANN/.
 *
 * PARITY STATUS: PINNED.  The search loop below restates src/pq_flash_index.cpp:1779-2906 and is checked against the
 * reference's own PQFlashIndex::load + cached_beam_search compiled from that file (oracle/_ref/
 * libleann_ref_diskann_flash.so, diskann_flash_harness.cpp: declaration-only stand-ins for libaio / Boost / the protoc
 * header, in-process file reader and embedding fetch): identical ids, hop / I/O / comparison counts and distances within
 * 2e-5 on the committed fixtures (outputs of the compiled reference are stored in tests/golden/vamana_small_expected.npz)
 * and on fresh indexes of both file layouts, l2 / mips / cosine, beam widths 1-8, io_limit (tests/test_vamana_oracle.py).
 * Its primitives are additionally pinned operation by operation: the queue (include/neighbor.h:39-152) and the PQ
 * arithmetic (src/pq.cpp:180-340) against oracle/_ref/libleann_ref_diskann.so.
 *
 * Two deliberate choices, both inside the 1e-3 distance tolerance and both shared with the CUDA path so ids match
 * bit for bit: (1) full-precision distances use the canonical lane-strided fp32 order of canon_dist.h instead of the
 * reference's AVX2 order (include/distance.h DistanceL2Float); (2) the squared norm in
 * preprocess_fetched_embeddings uses that same order.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "canon_dist.h"

enum { VO_L2 = 0, VO_MIPS = 1, VO_COSINE = 2 };
enum { VO_DEFERRED_FETCH = 1, VO_SKIP_SEARCH_REORDER = 2 };

typedef struct {
    int64_t n;
    int data_dim;   /* _data_dim = PQ ndims = stored coordinate count (raw dim + 1 for MIPS)          */
    int metric;
    int R;
    const int32_t* nbrs;   /* [n, R], valid ids first, padded with -1                                  */
    const uint32_t* deg;   /* [n]                                                                      */
    const uint32_t* medoids; int n_medoids;
    const float* centroid_data; /* [n_medoids, data_dim] or NULL (single medoid)                       */
    int n_chunks;
    const float* pivots;         /* [256, data_dim]  "tables" (src/pq.cpp:93)                          */
    const float* centroid;       /* [data_dim]                                                          */
    const uint32_t* chunk_offsets; /* [n_chunks + 1]                                                    */
    const uint8_t* codes;        /* [n, n_chunks]                                                       */
    float max_base_norm;
} vo_index;

/* ---------------------------------------------------------------- PQ: src/pq.cpp:158-166 (tables_tr), 180-219, 302-340 */
float* vo_pq_tables_tr(const vo_index* x) {
    float* t = (float*)malloc(sizeof(float) * 256 * (size_t)x->data_dim);
    for (int i = 0; i < 256; i++)
        for (int j = 0; j < x->data_dim; j++) t[(size_t)j * 256 + i] = x->pivots[(size_t)i * x->data_dim + j];
    return t;
}

void vo_pq_preprocess_query(const vo_index* x, float* query_vec) { /* no rotation matrix (OPQ) */
    for (int d = 0; d < x->data_dim; d++) query_vec[d] -= x->centroid[d];
}

void vo_pq_populate_chunk_distances(const vo_index* x, const float* tables_tr, const float* query_vec, float* dist_vec) {
    memset(dist_vec, 0, sizeof(float) * 256 * (size_t)x->n_chunks);
    for (int chunk = 0; chunk < x->n_chunks; chunk++) {
        float* chunk_dists = dist_vec + 256 * (size_t)chunk;
        for (uint32_t j = x->chunk_offsets[chunk]; j < x->chunk_offsets[chunk + 1]; j++) {
            const float* centers_dim_vec = tables_tr + 256 * (size_t)j;
            for (int idx = 0; idx < 256; idx++) {
                /* source: double diff = c - q (a float subtraction); chunk += (float)(diff * diff).  The product of two
                 * floats is exact in double, so the cast rounds once — it IS the float product — and the reference's own
                 * build flags (-O3 -mavx2 -mfma, GCC's default -ffp-contract=fast; DiskANN/CMakeLists.txt:429-441) fuse it
                 * with the accumulation.  Verified against the compiled pq.cpp (tests/test_vamana_oracle.py). */
                float diff = centers_dim_vec[idx] - query_vec[j];
                chunk_dists[idx] = fmaf(diff, diff, chunk_dists[idx]);
            }
        }
    }
}

void vo_pq_dist_lookup(const uint8_t* pq_ids, size_t n_pts, size_t n_chunks, const float* pq_dists, float* dists_out) {
    memset(dists_out, 0, n_pts * sizeof(float));
    for (size_t chunk = 0; chunk < n_chunks; chunk++) {
        const float* chunk_dists = pq_dists + 256 * chunk;
        for (size_t idx = 0; idx < n_pts; idx++) dists_out[idx] += chunk_dists[pq_ids[n_chunks * idx + chunk]];
    }
}

static void compute_dists(const vo_index* x, const float* lut, const uint32_t* ids, size_t n_ids, float* out, uint8_t* scratch) {
    for (size_t i = 0; i < n_ids; i++) /* aggregate_coords, src/pq.cpp:324-331 */
        memcpy(scratch + i * x->n_chunks, x->codes + (size_t)ids[i] * x->n_chunks, (size_t)x->n_chunks);
    vo_pq_dist_lookup(scratch, n_ids, (size_t)x->n_chunks, lut, out);
}

/* ---------------------------------------------------------------- NeighborPriorityQueue: include/neighbor.h:14-152 */
typedef struct {
    size_t size, capacity, cur;
    uint32_t* id; float* dist; uint8_t* expanded;
} vo_npq;

static int nbr_less(uint32_t ia, float da, uint32_t ib, float db) { return da < db || (da == db && ia < ib); }

vo_npq* vo_npq_new(size_t capacity) {
    vo_npq* q = (vo_npq*)calloc(1, sizeof(vo_npq));
    q->capacity = capacity;
    q->id = (uint32_t*)calloc(capacity + 1, sizeof(uint32_t));
    q->dist = (float*)calloc(capacity + 1, sizeof(float));
    q->expanded = (uint8_t*)calloc(capacity + 1, 1);
    return q;
}
void vo_npq_free(vo_npq* q) { free(q->id); free(q->dist); free(q->expanded); free(q); }

void vo_npq_insert(vo_npq* q, uint32_t id, float dist) {
    if (q->size == q->capacity && nbr_less(q->id[q->size - 1], q->dist[q->size - 1], id, dist)) return;
    size_t lo = 0, hi = q->size;
    while (lo < hi) {
        size_t mid = (lo + hi) >> 1;
        if (nbr_less(id, dist, q->id[mid], q->dist[mid])) hi = mid;
        else if (q->id[mid] == id) return;
        else lo = mid + 1;
    }
    if (lo < q->capacity) {
        memmove(&q->id[lo + 1], &q->id[lo], (q->size - lo) * sizeof(uint32_t));
        memmove(&q->dist[lo + 1], &q->dist[lo], (q->size - lo) * sizeof(float));
        memmove(&q->expanded[lo + 1], &q->expanded[lo], (q->size - lo));
    }
    q->id[lo] = id; q->dist[lo] = dist; q->expanded[lo] = 0;
    if (q->size < q->capacity) q->size++;
    if (lo < q->cur) q->cur = lo;
}

void vo_npq_closest_unexpanded(vo_npq* q, uint32_t* id, float* dist) {
    q->expanded[q->cur] = 1;
    size_t pre = q->cur;
    while (q->cur < q->size && q->expanded[q->cur]) q->cur++;
    *id = q->id[pre]; *dist = q->dist[pre];
}
int vo_npq_has_unexpanded(const vo_npq* q) { return q->cur < q->size; }
size_t vo_npq_size(const vo_npq* q) { return q->size; }
void vo_npq_get(const vo_npq* q, size_t i, uint32_t* id, float* dist, int* expanded) {
    *id = q->id[i]; *dist = q->dist[i]; *expanded = q->expanded[i];
}

/* ---------------------------------------------------------------- preprocess_fetched_embeddings: src/pq_flash_index.cpp:1723-1777 */
void vo_preprocess_embedding(const vo_index* x, const float* emb, int dim, float* out /* [data_dim] */) {
    if (x->metric == VO_MIPS) {
        const int m = x->data_dim - 1;
        float norm_sq = lb2_canon_ip(emb, emb, m < dim ? m : dim);
        for (int i = 0; i < m; i++) out[i] = (i < dim ? emb[i] : 0.0f) / x->max_base_norm;
        float res = 1 - (norm_sq / (x->max_base_norm * x->max_base_norm));
        res = res <= 0 ? 0 : sqrtf(res);
        out[m] = res;
    } else if (x->metric == VO_COSINE) {
        float norm = sqrtf(lb2_canon_ip(emb, emb, dim));
        for (int i = 0; i < x->data_dim; i++) {
            float v = i < dim ? emb[i] : 0.0f;
            out[i] = norm > 0 ? v / norm : v;
        }
    } else {
        for (int i = 0; i < x->data_dim; i++) out[i] = i < dim ? emb[i] : 0.0f;
    }
}

/* ---------------------------------------------------------------- query normalisation: src/pq_flash_index.cpp:1819-1848 */
void vo_prepare_query(const vo_index* x, const float* query, float* aq /* [data_dim], zero-initialised */, float* query_norm_out) {
    const int D = x->data_dim;
    float query_norm = 0;
    if (x->metric == VO_MIPS || x->metric == VO_COSINE) {
        const int inherent = x->metric == VO_COSINE ? D : D - 1;
        /* query_norm += q*q, contracted to an fma by the reference's build flags (see populate_chunk_distances) */
        for (int i = 0; i < inherent; i++) { aq[i] = query[i]; query_norm = fmaf(query[i], query[i], query_norm); }
        if (x->metric == VO_MIPS) aq[D - 1] = 0;
        query_norm = sqrtf(query_norm);
        for (int i = 0; i < inherent; i++) aq[i] = aq[i] / query_norm;
    } else {
        for (int i = 0; i < D; i++) aq[i] = query[i];
    }
    *query_norm_out = query_norm;
}

/* ---------------------------------------------------------------- cached_beam_search: src/pq_flash_index.cpp:1779-2906 */
typedef struct { uint32_t id; float dist; } vo_nb;
static int cmp_nb(const void* a, const void* b) {
    const vo_nb* x = (const vo_nb*)a; const vo_nb* y = (const vo_nb*)b;
    if (nbr_less(x->id, x->dist, y->id, y->dist)) return -1;
    if (nbr_less(y->id, y->dist, x->id, x->dist)) return 1;
    return 0;
}

/* query: raw [dim] (dim = data_dim - 1 for MIPS, data_dim otherwise).
 * coords: stored node coordinates [n, data_dim] (recompute off) or NULL; emb: fresh embeddings [n, dim] (deferred fetch) or NULL.
 * full_ids / full_dists: the expanded nodes in expansion order and their final distances (cap_full entries).
 * stats[0..2] = cmps, n_ios (= expansions), n_hops.  Returns 0, or -1 if cap_full overflowed. */
int vo_search(const vo_index* x, const float* tables_tr, const float* query, int dim, int k, int L, int beam_width, uint32_t io_limit,
              int flags, const float* coords, const float* emb, int64_t* out_ids, float* out_dists, uint32_t* full_ids,
              float* full_dists, int cap_full, int* n_full_out, int64_t* stats) {
    const int D = x->data_dim;
    const int deferred = (flags & VO_DEFERRED_FETCH) != 0, skip_reorder = (flags & VO_SKIP_SEARCH_REORDER) != 0;
    float* aq = (float*)calloc((size_t)D + 8, sizeof(float));      /* aligned_query_T            :1812 */
    float* qrot = (float*)calloc((size_t)D + 8, sizeof(float));    /* rotated_query              :1814 */
    float* lut = (float*)malloc(sizeof(float) * 256 * (size_t)x->n_chunks);
    float* dist_scratch = (float*)malloc(sizeof(float) * (size_t)(x->R + 1));
    uint8_t* code_scratch = (uint8_t*)malloc((size_t)(x->R + 1) * x->n_chunks);
    uint8_t* visited = (uint8_t*)calloc((size_t)x->n, 1);
    float* tmp = (float*)malloc(sizeof(float) * ((size_t)D + 8));
    uint32_t* frontier = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(beam_width + 1));
    float* frontier_d = (float*)malloc(sizeof(float) * (size_t)(beam_width + 1));
    float query_norm = 0;
    vo_prepare_query(x, query, aq, &query_norm);
    (void)dim;
    memcpy(qrot, aq, sizeof(float) * (size_t)D);
    vo_pq_preprocess_query(x, qrot);                           /* :1859 */
    vo_pq_populate_chunk_distances(x, tables_tr, qrot, lut);   /* :1862 */

    vo_npq* retset = vo_npq_new((size_t)L);                    /* :2110-2111 */
    int n_full = 0, overflow = 0;
    /* medoid seed :2119-2158 */
    uint32_t best_medoid = 0; float best_dist = FLT_MAX;
    for (int m = 0; m < x->n_medoids; m++) {
        float cur = x->centroid_data ? lb2_canon_l2(aq, x->centroid_data + (size_t)D * m, D) : 0.0f;
        if (cur < best_dist) { best_medoid = x->medoids[m]; best_dist = cur; }
    }
    compute_dists(x, lut, &best_medoid, 1, dist_scratch, code_scratch);
    vo_npq_insert(retset, best_medoid, dist_scratch[0]);
    visited[best_medoid] = 1;
    int64_t cmps = 0, hops = 0; uint32_t num_ios = 0;

    while (vo_npq_has_unexpanded(retset) && num_ios < io_limit) {       /* :2180 */
        int nf = 0; uint32_t num_seen = 0;
        while (vo_npq_has_unexpanded(retset) && nf < beam_width && num_seen < (uint32_t)beam_width) {  /* :2190-2211 */
            vo_npq_closest_unexpanded(retset, &frontier[nf], &frontier_d[nf]);
            num_seen++; nf++;
        }
        if (nf > 0) hops++;
        num_ios += (uint32_t)nf;                                        /* :2236-2243 */
        for (int f = 0; f < nf; f++) {                                  /* :2419-2612, batch_recompute = false */
            const uint32_t node = frontier[f];
            float cur_expanded_dist;
            if (skip_reorder) cur_expanded_dist = frontier_d[f];        /* compute_dists(&node_id, 1): same PQ sum as at insertion */
            else if (deferred) cur_expanded_dist = 0.0f;
            else cur_expanded_dist = lb2_canon_l2(aq, coords + (size_t)node * D, D);
            if (n_full < cap_full) { full_ids[n_full] = node; full_dists[n_full] = cur_expanded_dist; n_full++; }
            else overflow = 1;
            const uint32_t nnbrs = x->deg[node];
            uint32_t* ids = (uint32_t*)(x->nbrs + (size_t)node * x->R);
            compute_dists(x, lut, ids, nnbrs, dist_scratch, code_scratch);   /* prune_node_nbrs is a no-op: :2018-2021 */
            for (uint32_t m = 0; m < nnbrs; m++) {
                const uint32_t id = ids[m];
                if (!visited[id]) {
                    visited[id] = 1;
                    cmps++;
                    vo_npq_insert(retset, id, dist_scratch[m]);
                }
            }
        }
    }

    if (deferred) {                                                     /* :2661-2759 */
        for (int i = 0; i < n_full; i++) {
            vo_preprocess_embedding(x, emb + (size_t)full_ids[i] * dim, dim, tmp);
            full_dists[i] = lb2_canon_l2(aq, tmp, D);
        }
    }
    vo_nb* fr = (vo_nb*)malloc(sizeof(vo_nb) * (size_t)(n_full + 1));
    for (int i = 0; i < n_full; i++) { fr[i].id = full_ids[i]; fr[i].dist = full_dists[i]; }
    qsort(fr, (size_t)n_full, sizeof(vo_nb), cmp_nb);                   /* :2761 (ties ordered by id: total order) */
    for (int i = 0; i < k; i++) {                                       /* :2861-2883 */
        if (i >= n_full) { out_ids[i] = -1; out_dists[i] = FLT_MAX; continue; } /* the reference reads past the end here */
        out_ids[i] = fr[i].id;
        float d = fr[i].dist;
        if (x->metric == VO_MIPS) {
            d = -d;
            if (x->max_base_norm != 0) d *= (x->max_base_norm * query_norm);
        }
        out_dists[i] = d;
    }
    *n_full_out = n_full;
    if (stats) { stats[0] = cmps; stats[1] = num_ios; stats[2] = hops; }
    free(fr); vo_npq_free(retset);
    free(aq); free(qrot); free(lut); free(dist_scratch); free(code_scratch); free(visited); free(tmp); free(frontier); free(frontier_d);
    return overflow ? -1 : 0;
}

int vo_search_batch(const vo_index* x, int64_t nq, const float* queries, int dim, int k, int L, int beam_width, uint32_t io_limit,
                    int flags, const float* coords, const float* emb, int64_t* out_ids, float* out_dists, uint32_t* full_ids,
                    float* full_dists, int cap_full, int* n_full, int64_t* stats /* [nq, 3] */, int nthreads) {
    float* tr = vo_pq_tables_tr(x);
    int rc = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t q = 0; q < nq; q++) {
        int r = vo_search(x, tr, queries + q * dim, dim, k, L, beam_width, io_limit, flags, coords, emb, out_ids + q * k,
                          out_dists + q * k, full_ids + q * (int64_t)cap_full, full_dists + q * (int64_t)cap_full, cap_full,
                          n_full + q, stats ? stats + q * 3 : NULL);
        if (r) {
#pragma omp atomic write
            rc = r;
        }
    }
    free(tr);
    return rc;
}
