/* libleann_b200 — C ABI of the B200-native selective-recompute search path.
 *
 * Drop-in boundary for LEANN's HNSW backend search path.  Every entry point names the
 * reference interface it replaces (paths under /root/reference/packages/):
 *
 *   lb2_open            faiss.read_index(file, IO_FLAG_MMAP, HNSWIndexConfig)
 *                         leann-backend-hnsw/leann_backend_hnsw/hnsw_backend.py:141-151
 *                         -> third_party/faiss/faiss/impl/index_read.cpp:1402-1490, 523-813
 *   lb2_search          IndexHNSW::search(n, x, k, D, I, SearchParametersHNSW*)
 *                         hnsw_backend.py:241-248 -> faiss/IndexHNSW.cpp:419-439
 *   lb2_search_params   faiss::SearchParametersHNSW            faiss/impl/HNSW.h:54-73
 *   lb2_set_passages /  the embedding server's start-up inputs (passages file + model):
 *   lb2_set_encoder       leann-core/src/leann/embedding_server_manager.py:151-174,
 *                         hnsw_embedding_server.py:36-96 ; they feed the recompute stage that
 *                         replaces ZmqDistanceComputer::distances_batch
 *                         (faiss/impl/HNSW_zmq.cpp:579-654 -> hnsw_embedding_server.py:147-211)
 *   lb2_set_vectors     the IndexFlat storage of a non-pruned index (recompute_embeddings=False)
 *   lb2_encode_ids      embedding-server "ids -> embeddings" branch   hnsw_embedding_server.py:213-284
 *   lb2_encode_tokens   embedding-server "texts -> embeddings" branch hnsw_embedding_server.py:134-145
 *                         (compute_query_embedding, leann-core/src/leann/searcher_base.py:86-128)
 *   lb2_load_pq_pruning HNSW::load_pq_pruning_data(pq_pivots_path, pq_compressed_path)
 *                         faiss/impl/HNSW_search.cpp:253-297 (DiskANN-format PQ files; enables the PQ-guided pruning
 *                         branch of search_from_candidates, :442-465, 645-750)
 *   lb2_close           Index destructor / EmbeddingServerManager.stop_server
 *   lb2_last_error      FaissException text surfaced as Python RuntimeError by SWIG
 *
 * Conventions: plain pointers and sizes, no C++ or torch types; functions return 0 on success
 * and a negative code on failure (message via lb2_last_error(), thread-local); nothing throws
 * across the boundary.  Pointers are HOST pointers unless the name says device.  The library
 * needs a CUDA device of compute capability 10.x; there is no CPU path.
 */
#ifndef LEANN_B200_H
#define LEANN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LB2_OK 0
#define LB2_ERR_ARG (-1)
#define LB2_ERR_IO (-2)
#define LB2_ERR_CUDA (-3)
#define LB2_ERR_STATE (-4)
#define LB2_ERR_UNSUPPORTED (-5)

typedef struct lb2_index lb2_index; /* opaque */

/* faiss::SearchParametersHNSW (faiss/impl/HNSW.h:54-73) + the backend's recompute switch */
typedef struct {
    int32_t efSearch;                /* "complexity" of LeannSearcher.search; default 64          */
    int32_t beam_size;               /* "beam_width"; candidates expanded per hop; default 1       */
    int32_t batch_size;              /* expand until >= batch_size neighbours gathered; 0 = off    */
    int32_t check_relative_distance; /* stop rule of HNSW_search.cpp:583-592; default 1            */
    float pq_pruning_ratio;          /* "prune_ratio": share of PQ-ranked candidates NOT scored exactly; like the     */
    int32_t local_prune;             /* reference, the three pruning knobs act only after lb2_load_pq_pruning()       */
    float send_neigh_times_ratio;    /* ("local" / "proportional" strategies; default strategy = global queue)        */
    int32_t recompute;               /* 1: recompute embeddings on the GPU (default); 0: stored    */
} lb2_search_params;

typedef struct {
    int64_t ndis;            /* sum over queries of HNSWStats.ndis  (faiss/impl/HNSW.h:353-361)   */
    int64_t nhops;           /* sum over queries of HNSWStats.nhops                                 */
    int64_t n_recomputed;    /* passages encoded (after cross-query de-duplication per hop)        */
    int64_t n_requested;     /* (query, node) distance requests = ndis + entry points              */
    int64_t n_tokens;        /* tokens pushed through the encoder                                   */
    int64_t n_steps;         /* traversal kernel launches                                           */
    int64_t n_kernel_launches; /* all kernels launched by this call                                 */
    double gpu_ms;           /* device time of the call (CUDA events on the launching stream)      */
    double encoder_ms;       /* of which recompute stage                                           */
    double gemm_ms;          /* of which tcgen05 GEMMs (only when LB2_PROFILE_GEMM=1)               */
    double gemm_flops;       /* algorithmic flops of those GEMMs                                    */
    double attention_ms;     /* attention kernels (profiling on)                                    */
    double norm_ms;          /* LayerNorm kernels (profiling on)                                    */
    int64_t n_encoder_passes;/* encoder passes (each <= passages_per_pass passages) of the call           */
} lb2_search_stats;

typedef struct {
    int64_t ntotal;
    int32_t d;
    int32_t metric_type; /* faiss MetricType: 0 inner product, 1 L2 */
    int32_t entry_point;
    int32_t max_level;
    int64_t n_edges;
    int32_t max_degree_level0;
    int32_t max_degree_upper;
    int32_t has_vectors;
    int32_t has_passages;
    int32_t has_encoder;
    int32_t device;
} lb2_index_info;

/* BERT-family encoder description (architecture constants of the model card) */
typedef struct {
    int32_t vocab_size, hidden, layers, heads, ffn, max_pos, type_vocab;
    float ln_eps;
    int32_t pooling;   /* 0 masked mean, 1 CLS */
    int32_t normalize; /* 1: L2-normalise (Normalize module) */
} lb2_encoder_config;

const char* lb2_last_error(void);
int lb2_version(void);

/* Reads the reference's compact-CSR `.index` file and uploads the graph to `device`. */
lb2_index* lb2_open(const char* index_path, int device);
void lb2_close(lb2_index* idx);
int lb2_info(const lb2_index* idx, lb2_index_info* out);

/* Stored-vector mode (non-pruned index): x is [ntotal, d] fp32. */
int lb2_set_vectors(lb2_index* idx, const float* x);
int lb2_set_vectors_device(lb2_index* idx, const float* d_x); /* same, source already in device memory */

/* Pre-tokenised passage store: passage i = tokens[offsets[i] .. offsets[i+1]) (WordPiece ids incl.
 * [CLS]/[SEP], as the reference's tokenizer would produce them), ntotal+1 offsets. */
int lb2_set_passages(lb2_index* idx, const uint16_t* tokens, const uint64_t* offsets);

/* Weights: one fp32 blob, tensors in this order (nn.Linear weights are [out, in]):
 *   word_emb[V,H] pos_emb[P,H] type_emb[T,H] emb_ln_g[H] emb_ln_b[H]
 *   per layer: w_qkv[3H,H] (q,k,v rows stacked) b_qkv[3H] w_o[H,H] b_o[H] ln1_g[H] ln1_b[H]
 *              w_1[F,H] b_1[F] w_2[H,F] b_2[H] ln2_g[H] ln2_b[H]
 * Stored on the device as fp16 matrices / fp32 vectors (the reference loads fp16 too). */
/* PQ-guided pruning tables: DiskANN-format <prefix>_pq_pivots.bin / <prefix>_pq_compressed.bin over the index's
 * vectors (MIPS-extended: ndims = d + 1).  Without this call prune_ratio / pruning_strategy are ignored, as in the reference. */
int lb2_load_pq_pruning(lb2_index* idx, const char* pq_pivots_path, const char* pq_compressed_path);

size_t lb2_encoder_weight_count(const lb2_encoder_config* cfg);
int lb2_set_encoder(lb2_index* idx, const lb2_encoder_config* cfg, const float* weights, size_t n_floats);

void lb2_default_params(lb2_search_params* p);

/* q: [nq, d] fp32 (already normalised for cosine); D: [nq, k]; I: [nq, k].  Labels are the
 * internal int64 ids; distances are +inner product (descending) for IP metrics, squared L2
 * (ascending) otherwise; unfilled slots are (-1, -/+FLT_MAX) like faiss. */
int lb2_search(lb2_index* idx, int64_t nq, const float* q, int64_t k, float* D, int64_t* I,
               const lb2_search_params* params, lb2_search_stats* stats);
/* Same, all three buffers already resident in device memory. */
int lb2_search_device(lb2_index* idx, int64_t nq, const float* d_q, int64_t k, float* d_D, int64_t* d_I,
                      const lb2_search_params* params, lb2_search_stats* stats);
/* Per-query HNSWStats of the last search call ([nq] each, host). */
int lb2_last_query_stats(lb2_index* idx, int64_t nq, int64_t* ndis, int64_t* nhops);

/* Recompute stage alone. */
int lb2_encode_ids(lb2_index* idx, int64_t n, const int64_t* ids, float* out /* [n, d] */);
int lb2_encode_tokens(lb2_index* idx, int64_t n, const uint16_t* tokens, const uint64_t* offsets /* n+1 */,
                      float* out /* [n, d] */);
/* passages [first, first+n) of the attached store -> DEVICE buffer d_out [n, d] (index building) */
int lb2_encode_range_device(lb2_index* idx, int64_t first, int64_t n, float* d_out);

/* Tunables (0 keeps the current value): traversal slots in flight, passages per encoder pass. */
int lb2_configure(lb2_index* idx, int32_t slots, int32_t passages_per_pass);

/* Generic knobs: "slots", "passages_per_pass", "profile" (0/1: per-kernel-class CUDA-event timing in the stats),
 * "dedup_scope": 0 (default) = a passage requested by several queries in the same hop is encoded once;
 * 1 = once per search call (later hops of any query in the call reuse the embedding; one fp32 row per distinct
 * passage lives in HBM until the call returns — identical results, fewer recomputes, nothing persists). */
int lb2_set_option(lb2_index* idx, const char* key, int64_t value);

/* ------------------------------------------------------------------------------------------------
 * DiskANN / Vamana backend (leann-backend-diskann).  Same handle type; lb2_set_passages / lb2_set_encoder /
 * lb2_encode_* / lb2_set_option / lb2_close work on it unchanged.
 *
 *   lb2_diskann_open     StaticDiskIndex<float>(metric, index_path_prefix, num_threads, num_nodes_to_cache,
 *                          cache_mechanism, zmq_port, pq_prefix, partition_prefix)
 *                          leann-backend-diskann/leann_backend_diskann/diskann_backend.py:363-372
 *                          -> third_party/DiskANN/python/src/static_disk_index.cpp:15-48
 *                          -> PQFlashIndex::load, third_party/DiskANN/src/pq_flash_index.cpp:887-911, 1017-1467
 *                        Opens the same files: <p>_pq_pivots.bin, <p>_pq_compressed.bin, <p>_disk.index (+_medoids.bin,
 *                        _centroids.bin, _max_base_norm.bin) or, with a partition prefix, <pp>_partition.bin and
 *                        <pp>_disk_graph.index.
 *   lb2_diskann_search   StaticDiskIndex::batch_search(queries, num_queries, knn, complexity, beam_width, num_threads,
 *                          USE_DEFERRED_FETCH, skip_search_reorder, recompute_beighbor_embeddings, dedup_node_dis,
 *                          prune_ratio, batch_recompute, global_pruning)
 *                          diskann_backend.py:452-467 -> static_disk_index.cpp:88-118
 *                          -> PQFlashIndex::cached_beam_search, pq_flash_index.cpp:1779-2906,
 *                        including the deferred fetch_embeddings round trip (:2661-2759 -> diskann_embedding_server.py)
 *                        which becomes one encoder pass over the de-duplicated expanded nodes of the whole batch.
 *   lb2_diskann_params   the trailing arguments of batch_search; `metric` is given at open time as in the reference.
 */
typedef struct {
    int32_t complexity;          /* L: candidate list size                                   (l_search)        */
    int32_t beam_width;          /* nodes expanded per iteration                                               */
    int32_t deferred_fetch;      /* USE_DEFERRED_FETCH = recompute_embeddings: re-rank from fresh embeddings   */
    int32_t skip_search_reorder; /* keep PQ distances of the expanded nodes                                    */
    int32_t recompute_neighbors; /* must be 0: LEANN always passes false (diskann_backend.py:449)              */
    int32_t dedup_node_dis;      /* only meaningful with recompute_neighbors; accepted and ignored             */
    int32_t batch_recompute;     /* with recompute_neighbors == 0 this only regroups identical work; ignored   */
    int32_t global_pruning;      /* PQ pruning acts only inside recompute_neighbors (prune_node_nbrs,          */
    float prune_ratio;           /*   pq_flash_index.cpp:2018-2021 returns at once otherwise); ignored         */
    uint32_t io_limit;           /* 0 = unlimited (the python binding passes UINT32_MAX)                       */
} lb2_diskann_params;

typedef struct {
    int64_t npts;
    int32_t dim;            /* embedding dimension queries must have (data_dim - 1 for mips)   */
    int32_t data_dim;       /* stored coordinate count = PQ dims                               */
    int32_t n_chunks;       /* PQ bytes per vector                                             */
    int32_t max_degree;
    int32_t metric;         /* 0 l2, 1 mips, 2 cosine                                          */
    int32_t n_medoids;
    int32_t partitioned;    /* opened from <pp>_partition.bin + <pp>_disk_graph.index          */
    int32_t has_coords;     /* full-precision coordinates loaded (recompute_embeddings=False)  */
    float max_base_norm;
    int32_t pad;
    int64_t n_edges;
} lb2_diskann_info_t;

#define LB2_METRIC_L2 0
#define LB2_METRIC_MIPS 1
#define LB2_METRIC_COSINE 2

lb2_index* lb2_diskann_open(const char* index_prefix, const char* partition_prefix /* NULL or "" = none */, int metric,
                            int device);
int lb2_diskann_info(const lb2_index* idx, lb2_diskann_info_t* out);
void lb2_diskann_default_params(lb2_diskann_params* p);
/* q: [nq, dim] host; D: [nq, k] host; I: [nq, k] host (ids widened to int64; -1 / FLT_MAX where fewer than k nodes
 * were expanded).  Distances follow the reference: squared L2 for l2 and cosine (on unit vectors), and for mips
 * -(L2 in the extended space) * max_base_norm * |q| (pq_flash_index.cpp:2873-2881). */
int lb2_diskann_search(lb2_index* idx, int64_t nq, const float* q, int64_t k, float* D, int64_t* I,
                       const lb2_diskann_params* params, lb2_search_stats* stats);
int lb2_diskann_search_device(lb2_index* idx, int64_t nq, const float* d_q, int64_t k, float* d_D, int64_t* d_I,
                              const lb2_diskann_params* params, lb2_search_stats* stats);
/* test / diagnostics hook: the expanded nodes (full_retset) of the queries of the last search call, in expansion
 * order; ids is [nq, cap], n_full [nq].  Valid for calls of at most one wave (4096 queries). */
int lb2_diskann_last_expansions(lb2_index* idx, int64_t nq, int32_t cap, uint32_t* ids, int32_t* n_full);

/* ------------------------------------------------------------------------------------------------
 * Graph construction (tooling around the path): the two device stages of a batch-parallel
 * HNSW build; all pointers are DEVICE pointers, launches go to the default stream, nothing synchronises.
 *
 *   lb2_build_insert_search   per point: greedy descent over the upper levels + efConstruction search on level 0
 *                               HNSW::add_with_locks -> search_neighbors_to_add, faiss/impl/HNSW.cpp:839-894, 609-720
 *   lb2_build_select          neighbour-selection heuristic over sorted candidates
 *                               HNSW::shrink_neighbor_list, faiss/impl/HNSW.cpp:426-468
 * The batch schedule, link merging and the CSR writer are leann_b200/graph_build.py.
 */
/* x: [n, d] fp16 (d % 8 == 0); adj0: [n, cap0] level-0 lists, -1 padded (empty rows = not inserted yet);
 * up_row[n]: first row of a node's level-1 list in up_adj [rows, capU] (level l list = row up_row + l - 1), -1 for
 * level-0-only nodes; entry / max_level: entry point and its (0-based) level; points[npts]: ids whose vectors are the
 * queries; out_ids / out_dist: [npts, ef] nearest found, ascending, (-1, FLT_MAX) padded; the point itself is never
 * returned.  workspace: lb2_build_workspace_bytes(ef, cap0) bytes. */
int lb2_build_insert_search(const void* d_x_f16, int64_t n, int32_t d, int32_t metric_ip, const int32_t* d_adj0,
                            int32_t cap0, const int32_t* d_up_row, const int32_t* d_up_adj, int32_t capU, int32_t entry,
                            int32_t max_level, const int32_t* d_points, int64_t npts, int32_t ef, int32_t* d_out_ids,
                            float* d_out_dist, void* d_workspace, size_t workspace_bytes);
size_t lb2_build_workspace_bytes(int32_t ef, int32_t cap0);
/* pd: [b, K, K] pairwise candidate distances (fp16, or fp32 when pd_is_f32); dn: [b, K] node-to-candidate distances,
 * ascending; cand: [b, K] ids, -1 padded at the end; out: [b, keep] kept ids / distances, (-1, FLT_MAX) padded.
 * fill (0..keep): lists shorter than that are topped up with the nearest rejected candidates (the reference's
 * keep_max_size_level0 branch with a settable floor; 0 = the default behaviour). */
int lb2_build_select(const void* d_pd, int32_t pd_is_f32, const float* d_dn, const int32_t* d_cand, int64_t b, int32_t K,
                     int32_t keep, int32_t fill, int32_t* d_out_ids, float* d_out_dist);

/* ---- kernel-level hooks for the unit tests (device pointers, default stream, synchronous) ---- */
int lb2_test_gemm_f16(const void* dA, const void* dW, const float* dbias, const void* dres, void* dC, int M, int N,
                      int K, int epilogue /* 0 bias, 1 bias+gelu, 2 bias+residual */);
/* C = LayerNorm(A . W^T + bias + residual) * gamma + beta, N = 384 (the fused residual projections of a BERT block) */
int lb2_test_gemm_res_ln_f16(const void* dA, const void* dW, const float* dbias, const void* dres, const float* dgamma,
                             const float* dbeta, float eps, void* dC, int M, int N, int K);
int lb2_test_layernorm_f16(const void* din, const float* dg, const float* db, void* dout, int rows, int hidden,
                           float eps);
/* qkv: HEAD-MAJOR packed [heads][n_tokens][3*head_dim] fp16 (q|k|v per token), n_tokens = sum(len);
 * d_seq_start / d_seq_len: device int32[n_seq] (first row, length); ctx: [n_tokens, hidden] */
int lb2_test_attention_f16(const void* dqkv, const int32_t* d_seq_start, const int32_t* d_seq_len, int n_seq,
                           int n_tokens, int hidden, int heads, int max_len, void* dctx);
/* grouped-output GEMM (the QKV projection): C is [N / c_group][M][c_group] */
int lb2_test_gemm_grouped_f16(const void* dA, const void* dW, const float* dbias, void* dC, int M, int N, int K,
                              int c_group);

#ifdef __cplusplus
}
#endif
#endif /* LEANN_B200_H */
