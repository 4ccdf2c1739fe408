// K1 — best-first HNSW traversal, one warp per in-flight query ("slot").
//
// Restates, per slot, the reference's HNSW::search (faiss/impl/HNSW.cpp:1081-1155):
// entry-point distance, greedy descent over levels max_level..1
// (greedy_update_nearest, HNSW.cpp:1012-1063) and the level-0 loop
// search_from_candidates (faiss/impl/HNSW_search.cpp:369-850, default non-PQ branch) with
// its MinimaxHeap (HNSW.cpp:1263-1509) and top-k result heap (ResultHandler.h:214-250).
// The distances_batch() call of the reference (a ZMQ round trip to the embedding server,
// HNSW_zmq.cpp:579-654) becomes: emit the hop's node ids -> [recompute stage on the same
// GPU] -> score (-q.e or |q-e|^2) and fold on the next invocation.  In stored-vector mode
// (recompute_embeddings=False) no stage sits in between and the kernel is persistent: each
// warp drives its queries to completion and pulls new ones from a global counter.
//
// Parity notes (all replicated on purpose):
//  * candidates heap = faiss binary max-heap keyed (dis, id); pop_min takes the RIGHTMOST
//    minimum of the array; count_below also counts popped entries      (HNSW.cpp:1466-1509)
//  * stop rule "count_below(d0) >= efSearch" uses efSearch, heap capacity is max(ef, k)
//  * a hop's new nodes are deduplicated and folded in ASCENDING id order (std::set,
//    HNSW_search.cpp:623-640); results admitted iff dis < threshold  (ResultHandler.h:234-238)
//  * greedy levels score ALL neighbours (no visited filter), level 0 re-scores them.
// Distances use the canonical summation order of oracle/canon_dist.h (lane-strided fmaf,
// xor-butterfly), so ids, distances, ndis and nhops are bit-identical to the oracle.
#include <float.h>

#include "traverse.cuh"

namespace lb2 {

namespace {

constexpr int STEP_WARPS = 4;

__device__ __forceinline__ bool gt2(float a1, float b1, int a2, int b2) {
    return (a1 > b1) || ((a1 == b1) && (a2 > b2));  // CMax::cmp2, utils/ordered_key_value.h:74-76
}

// --- faiss heap primitives on (float, int) arrays, 1-based inside (utils/Heap.h:47-150) ---
__device__ void heap_pop(int k, float* val, int* ids) {
    val--; ids--;
    const float v = val[k];
    const int id = ids[k];
    int i = 1;
    for (;;) {
        const int i1 = i << 1, i2 = i1 + 1;
        if (i1 > k) break;
        const int c = (i2 == k + 1 || gt2(val[i1], val[i2], ids[i1], ids[i2])) ? i1 : i2;
        if (gt2(v, val[c], id, ids[c])) break;
        val[i] = val[c]; ids[i] = ids[c]; i = c;
    }
    val[i] = val[k]; ids[i] = ids[k];
}
__device__ void heap_push(int k, float* val, int* ids, float v, int id) {
    val--; ids--;
    int i = k;
    while (i > 1) {
        const int f = i >> 1;
        if (!gt2(v, val[f], id, ids[f])) break;
        val[i] = val[f]; ids[i] = ids[f]; i = f;
    }
    val[i] = v; ids[i] = id;
}
__device__ void heap_replace_top(int k, float* val, int* ids, float v, int id) {
    val--; ids--;
    int i = 1;
    for (;;) {
        const int i1 = i << 1, i2 = i1 + 1;
        if (i1 > k) break;
        const int c = (i2 == k + 1 || gt2(val[i1], val[i2], ids[i1], ids[i2])) ? i1 : i2;
        if (gt2(v, val[c], id, ids[c])) break;
        val[i] = val[c]; ids[i] = ids[c]; i = c;
    }
    val[i] = v; ids[i] = id;
}

struct Warp {
    // shared-memory views of this warp's slot
    float* q;
    int* h_ids; float* h_dis;
    float* r_dis; int* r_ids;
    int* rq; float* rq_dis;  // rq: the hop's gathered neighbours (sort buffer, p2); rq_dis: distances of the requests
    int* req;                // the hop's distance requests, in the order they are folded (== rq without PQ pruning)
    float* pqd;              // PQ distances of the gathered neighbours (p2; PQ pruning only)
};

// MinimaxHeap::push (HNSW.cpp:1263-1274); lane 0 only
__device__ void mm_push(Warp& w, int& hk, int& hnvalid, int hcap, int id, float v) {
    if (hk == hcap) {
        if (v >= w.h_dis[0]) return;
        if (w.h_ids[0] != -1) --hnvalid;
        heap_pop(hk--, w.h_dis, w.h_ids);
    }
    heap_push(++hk, w.h_dis, w.h_ids, v, id);
    ++hnvalid;
}

// MinimaxHeap::pop_min (HNSW.cpp:1466-1497): rightmost minimum among valid entries; whole warp
__device__ int mm_pop_min(Warp& w, int hk, int& hnvalid, float* vmin, int lane) {
    float bd = FLT_MAX;
    int bi = -1;
    for (int i = lane; i < hk; i += 32) {
        if (w.h_ids[i] != -1) {
            const float dd = w.h_dis[i];
            if (bi == -1 || dd <= bd) { bd = dd; bi = i; }
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi != -1 && (bi == -1 || od < bd || (od == bd && oi > bi))) { bd = od; bi = oi; }
    }
    if (bi == -1) return -1;
    const int ret = w.h_ids[bi];
    __syncwarp();
    if (lane == 0) w.h_ids[bi] = -1;
    __syncwarp();
    --hnvalid;
    *vmin = bd;
    return ret;
}

// MinimaxHeap::count_below (HNSW.cpp:1500-1509): over all k slots, popped ones included
__device__ int mm_count_below(const Warp& w, int hk, float thresh, int lane) {
    int n = 0;
    for (int i = lane; i < hk; i += 32) n += (w.h_dis[i] < thresh) ? 1 : 0;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    return n;
}

// canonical distance: see oracle/canon_dist.h
__device__ __forceinline__ float canon_dist(const float* q_s, const float* __restrict__ e, int d, int metric_ip,
                                            int lane) {
    float acc = 0.f;
    if (metric_ip) {
        for (int j = lane; j < d; j += 32) acc = fmaf(q_s[j], __ldg(e + j), acc);
    } else {
        for (int j = lane; j < d; j += 32) {
            const float t = q_s[j] - __ldg(e + j);
            acc = fmaf(t, t, acc);
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return metric_ip ? -acc : acc;
}

// CSR addressing (HNSW_search.cpp:299-344)
__device__ __forceinline__ void neighbor_range(const DevGraph& g, int node, int level, uint64_t* b, uint64_t* e) {
    const uint64_t ps = g.node_offsets[node], pe = g.node_offsets[node + 1];
    const int nlev = static_cast<int>(pe - ps) - 1;
    if (level < 0 || level >= nlev) { *b = 0; *e = 0; return; }
    *b = g.level_ptr[ps + level];
    *e = g.level_ptr[ps + level + 1];
}

// ascending bitonic sort of a[0..P) (P power of two) by one warp
__device__ void warp_sort(int* a, int P, int lane) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < P; i += 32) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const int x = a[i], y = a[ixj];
                    const bool up = ((i & k) == 0);
                    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncwarp();
        }
    }
}

// in-place removal of adjacent duplicates from sorted a[0..n); returns the new length
__device__ int warp_unique(int* a, int n, int lane) {
    int out = 0;
    int prev_last = -2147483647 - 1;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const int v = (i < n) ? a[i] : 0;
        int pv = __shfl_up_sync(0xffffffffu, v, 1);
        if (lane == 0) pv = prev_last;
        const bool keep = (i < n) && (base + lane == 0 || v != pv);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int last_valid = min(31, n - base - 1);
        prev_last = __shfl_sync(0xffffffffu, v, last_valid);
        __syncwarp();
        if (keep) a[out + __popc(m & ((1u << lane) - 1))] = v;
        out += __popc(m);
        __syncwarp();
    }
    return out;
}


// ------------------------------------------------------------------------------------------------------------------
// PQ-guided pruning (faiss/impl/HNSW_search.cpp:442-465, 645-750; arithmetic of faiss/impl/pq.cpp in its own operation
// order: double-precision differences, float accumulation — bit-identical tables and PQ distances).

// query -> (append 0, normalise, centre) -> per-chunk distance table.  HNSW_search.cpp:447-460, pq.cpp:140-190
__device__ void pq_build_lut(const PqDev& pq, const float* q_s, int d, float* qprep, float* lut, int lane) {
    const int dim = pq.ndims;
    for (int i = lane; i < dim; i += 32) qprep[i] = (i < dim - 1 && i < d) ? q_s[i] : 0.f;
    __syncwarp();
    float fn = 1.f;
    if (lane == 0) {
        double norm_sq = 0.0;
        for (int i = 0; i < dim; i++) norm_sq = __dadd_rn(norm_sq, __dmul_rn((double)qprep[i], (double)qprep[i]));
        double norm = __dsqrt_rn(norm_sq);
        if (norm <= 0.0) norm = 1.0;
        fn = __double2float_rn(norm);
    }
    fn = __shfl_sync(0xffffffffu, fn, 0);
    for (int i = lane; i < dim; i += 32) qprep[i] = __fsub_rn(__fdiv_rn(qprep[i], fn), pq.centroid[i]);
    __syncwarp();
    for (int c = 0; c < pq.n_chunks; c++) {
        const uint32_t j0 = pq.chunk_offsets[c], j1 = pq.chunk_offsets[c + 1];
        for (int idx = lane; idx < 256; idx += 32) {
            float acc = 0.f;
            for (uint32_t j = j0; j < j1; j++) {
                const double diff = __dsub_rn((double)pq.tables_tr[256 * (size_t)j + idx], (double)qprep[j]);
                acc = __fadd_rn(acc, __double2float_rn(__dmul_rn(diff, diff)));
            }
            lut[256 * c + idx] = acc;
        }
    }
    __syncwarp();
}

// pq_distance_lookup (pq.cpp:207-224): chunk-major float accumulation from 0
__device__ __forceinline__ float pq_lookup(const PqDev& pq, const float* lut, int id) {
    const uint8_t* code = pq.codes + (size_t)id * pq.n_chunks;
    float acc = 0.f;
    for (int c = 0; c < pq.n_chunks; c++) acc = __fadd_rn(acc, __ldcg(lut + 256 * c + code[c]));
    return acc;
}

__device__ __forceinline__ bool pair_lt(float d1, int i1, float d2, int i2) { return d1 < d2 || (d1 == d2 && i1 < i2); }

// ascending bitonic sort of the pairs (kd[i], ki[i]), i < P (power of two), by (distance, id): std::pair ordering
__device__ void warp_sort_pairs(float* kd, int* ki, int P, int lane) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < P; i += 32) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const float d1 = kd[i], d2 = kd[ixj];
                    const int i1 = ki[i], i2 = ki[ixj];
                    const bool up = ((i & k) == 0);
                    if (pair_lt(d2, i2, d1, i1) == up) { kd[i] = d2; ki[i] = i2; kd[ixj] = d1; ki[ixj] = i1; }
                }
            }
            __syncwarp();
        }
    }
}

// number of entries of the ascending array (ad, ai)[lo, hi) that are <  (d, id)   [strict]   /   <= (d, id)   [!strict]
__device__ __forceinline__ int pairs_rank(const float* ad, const int* ai, int lo, int hi, float d, int id, bool strict) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const float md = ad[mid];
        const int mi = ai[mid];
        const bool go_right = strict ? pair_lt(md, mi, d, id) : !pair_lt(d, id, md, mi);
        if (go_right) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Merge the hop's sorted batch (bd, bi)[0, nb) (shared memory) into the slot's ascending candidate array (qd, qi)[head, qn)
// (global memory).  unique: drop batch entries already present (std::set semantics of the proportional mode).
// `pos` is nb ints of scratch in shared memory.  Returns the new qn, or -1 on overflow.
__device__ int pq_merge(float* qd, int* qi, int head, int qn, int cap, float* bd, int* bi, int nb, bool unique, int* pos, int lane) {
    // rank of every batch entry among the old entries, and (set mode) whether it is already there
    int kept = 0;
    for (int b0 = 0; b0 < nb; b0 += 32) {
        const int j = b0 + lane;
        int r = 0;
        bool keep = false;
        if (j < nb) {
            r = pairs_rank(qd, qi, head, qn, bd[j], bi[j], true);          // old entries strictly smaller
            keep = !(unique && r < qn && qd[r] == bd[j] && qi[r] == bi[j]);
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int o = kept + __popc(m & ((1u << lane) - 1));
            pos[o] = r;                                                     // compacted: rank in the OLD array
            const float dd = bd[j]; const int ii = bi[j];
            __syncwarp(m);
            bd[o] = dd; bi[o] = ii;                                         // o <= j: in-place compaction is safe chunk by chunk
        }
        kept += __popc(m);
        __syncwarp();
    }
    nb = kept;
    if (qn + nb > cap) return -1;
    // move the old entries up, highest chunk first: entry i goes to i + (number of kept batch entries ranked <= i)
    for (int hi = qn; hi > head; hi -= 32) {
        const int i = hi - 1 - lane;
        float dd = 0.f; int ii = 0; int shift = 0;
        if (i >= head) {
            dd = qd[i]; ii = qi[i];
            int lo2 = 0, hi2 = nb;   // batch entries whose rank (insertion point in the old array) is <= i
            while (lo2 < hi2) { const int mid = (lo2 + hi2) >> 1; if (pos[mid] <= i) lo2 = mid + 1; else hi2 = mid; }
            shift = lo2;
        }
        __syncwarp();
        if (i >= head && shift) { qd[i + shift] = dd; qi[i + shift] = ii; }
        __syncwarp();
    }
    for (int j = lane; j < nb; j += 32) { qd[pos[j] + j] = bd[j]; qi[pos[j] + j] = bi[j]; }
    __syncwarp();
    return qn + nb;
}

__global__ void __launch_bounds__(STEP_WARPS * 32)
hnsw_step_kernel(const DevGraph g, const TravParams p, const TravState st, int max_iters) {
    extern __shared__ __align__(16) uint8_t step_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * STEP_WARPS + warp;
    if (slot >= st.S) return;
    int phase = st.phase[slot];
    if (phase == PH_IDLE) return;

    const int d = g.d;
    const int dpad = (d + 3) & ~3;
    const size_t per_warp = (size_t)dpad * 4 + (size_t)p.hcap * 8 + (size_t)p.k * 8 + (size_t)p.p2 * 4 + (size_t)p.cap_req * 4 +
                            (p.pq_mode ? (size_t)p.cap_req * 4 + (size_t)p.p2 * 4 : 0);
    uint8_t* base = step_smem + ((per_warp + 15) & ~size_t(15)) * warp;
    Warp w;
    w.q = reinterpret_cast<float*>(base);
    w.h_ids = reinterpret_cast<int*>(w.q + dpad);
    w.h_dis = reinterpret_cast<float*>(w.h_ids + p.hcap);
    w.r_dis = w.h_dis + p.hcap;
    w.r_ids = reinterpret_cast<int*>(w.r_dis + p.k);
    w.rq = w.r_ids + p.k;
    w.rq_dis = reinterpret_cast<float*>(w.rq + p.p2);
    w.req = p.pq_mode ? reinterpret_cast<int*>(w.rq_dis + p.cap_req) : w.rq;
    w.pqd = p.pq_mode ? reinterpret_cast<float*>(w.req + p.cap_req) : nullptr;

    // ---- load slot state
    int qid = st.qid[slot], level = st.level[slot], nearest = st.nearest[slot], prev_nearest = st.prev_nearest[slot];
    float d_nearest = st.d_nearest[slot];
    int hk = st.hk[slot], hnvalid = st.hnvalid[slot], nstep = st.nstep[slot], pend_beam = st.pend_beam[slot];
    int n_req = st.n_req[slot];
    long long ndis = st.ndis[slot], nhops = st.nhops[slot];
    int* g_hids = st.heap_ids + (size_t)slot * p.hcap;
    float* g_hdis = st.heap_dis + (size_t)slot * p.hcap;
    int* g_rids = st.res_ids + (size_t)slot * p.k;
    float* g_rdis = st.res_dis + (size_t)slot * p.k;
    int* g_req = st.req_ids + (size_t)slot * p.cap_req;
    uint32_t* vis = st.visited + (size_t)slot * st.vis_words;
    if (phase != PH_FETCH) {
        for (int i = lane; i < hk; i += 32) { w.h_ids[i] = g_hids[i]; w.h_dis[i] = g_hdis[i]; }
        for (int i = lane; i < p.k; i += 32) { w.r_ids[i] = g_rids[i]; w.r_dis[i] = g_rdis[i]; }
        for (int i = lane; i < n_req; i += 32) w.req[i] = g_req[i];
        const float* qg = st.queries + (size_t)qid * d;
        for (int j = lane; j < d; j += 32) w.q[j] = qg[j];
    }
    __syncwarp();
    // hop scope: rows are per hop, so the table is double-buffered by hop parity (this launch reads what the
    // previous launch wrote while claiming new rows); call scope: a node's row is assigned once per call and
    // only read by later launches, so one table serves both.
    const int* slot_rd = st.call_scope ? st.slot_of[0] : st.slot_of[(st.epoch + 1) & 1];
    int* slot_wr = st.call_scope ? st.slot_of[0] : st.slot_of[st.epoch & 1];
    bool have_pending = (phase != PH_FETCH) && (n_req > 0);

    for (int iter = 0; iter < max_iters; iter++) {
        // =========================== score + fold the pending requests ===========================
        if (have_pending) {
            for (int i = 0; i < n_req; i++) {
                const int node = w.req[i];
                const float* e = st.recompute ? st.E + (size_t)slot_rd[node] * d : st.vectors + (size_t)node * d;
                float dist = canon_dist(w.q, e, d, g.metric_ip, lane);
                if (dist != dist) {  // NaN embedding: the reference raises (HNSW_zmq.cpp:383-400); flag it and keep the heaps sane
                    if (lane == 0) atomicExch(st.error, 3);
                    dist = FLT_MAX;
                }
                if (lane == 0) w.rq_dis[i] = dist;
            }
            __syncwarp();
            if (phase == PH_ENTRY) {  // HNSW.cpp:1104-1105
                d_nearest = w.rq_dis[0];
                nearest = g.entry_point;
                level = g.max_level;
                phase = (level >= 1) ? PH_GREEDY : PH_BASE_INIT;
            } else if (phase == PH_GREEDY) {  // HNSW.cpp:1043-1061
                for (int i = 0; i < n_req; i++) {
                    const float dd = w.rq_dis[i];
                    if (dd < d_nearest) { d_nearest = dd; nearest = w.req[i]; }
                }
                ndis += n_req;
                nhops++;
                if (nearest == prev_nearest) {
                    level--;
                    if (level < 1) phase = PH_BASE_INIT;
                }
            } else {  // PH_BASE: HNSW_search.cpp:765-787
                if (lane == 0) {
                    float thr = w.r_dis[0];
                    for (int i = 0; i < n_req; i++) {
                        const float dd = w.rq_dis[i];
                        const int id = w.req[i];
                        if (dd < thr) { heap_replace_top(p.k, w.r_dis, w.r_ids, dd, id); thr = w.r_dis[0]; }
                        mm_push(w, hk, hnvalid, p.hcap, id, dd);
                    }
                }
                hk = __shfl_sync(0xffffffffu, hk, 0);
                hnvalid = __shfl_sync(0xffffffffu, hnvalid, 0);
                __syncwarp();
                ndis += n_req;
                nstep += pend_beam;
                if (!p.check_rel && nstep > p.ef) hnvalid = -1;  // forces the finish branch below
            }
            n_req = 0;
            have_pending = false;
        }

        // =========================== advance until new requests are emitted ===========================
        for (;;) {
            if (phase == PH_FETCH) {
                int qi = 0;
                if (lane == 0) qi = atomicAdd(st.next_query, 1);
                qi = __shfl_sync(0xffffffffu, qi, 0);
                if (qi >= st.nq) { phase = PH_IDLE; break; }
                qid = qi;
                const float* qg = st.queries + (size_t)qid * d;
                for (int j = lane; j < d; j += 32) w.q[j] = qg[j];
                for (int i = lane; i < p.k; i += 32) { w.r_dis[i] = FLT_MAX; w.r_ids[i] = -1; }  // ResultHandler.h:226-231
                hk = 0; hnvalid = 0; nstep = 0; ndis = 0; nhops = 0; pend_beam = 0;
                __syncwarp();
                phase = PH_ENTRY;
                if (lane == 0) w.req[0] = g.entry_point;
                n_req = 1;
                __syncwarp();
                break;
            }
            if (phase == PH_GREEDY) {  // HNSW.cpp:1025-1041
                uint64_t b, e;
                neighbor_range(g, nearest, level, &b, &e);
                const int n = static_cast<int>(e - b);
                prev_nearest = nearest;
                if (n == 0) {
                    nhops++;
                    level--;
                    if (level < 1) phase = PH_BASE_INIT;
                    continue;
                }
                for (int i = lane; i < n; i += 32) w.req[i] = g.nbrs[b + i];
                n_req = n;
                __syncwarp();
                break;
            }
            if (phase == PH_BASE_INIT) {  // HNSW.cpp:1117-1121, HNSW_search.cpp:478-490
                if (lane == 0) {
                    mm_push(w, hk, hnvalid, p.hcap, nearest, d_nearest);
                    if (d_nearest < w.r_dis[0]) heap_replace_top(p.k, w.r_dis, w.r_ids, d_nearest, nearest);
                    atomicOr(&vis[nearest >> 5], 1u << (nearest & 31));
                }
                hk = __shfl_sync(0xffffffffu, hk, 0);
                hnvalid = __shfl_sync(0xffffffffu, hnvalid, 0);
                nstep = 0;
                __syncwarp();
                if (p.pq_mode) {  // HNSW_search.cpp:447-460: the query's PQ table; fresh candidate queue / set
                    pq_build_lut(st.pq, w.q, d, st.pq_qprep + (size_t)slot * st.pq.ndims,
                                 st.pq_lut + (size_t)slot * st.pq.n_chunks * 256, lane);
                    if (lane == 0) { st.pq_qn[slot] = 0; st.pq_qhead[slot] = 0; }
                    __syncwarp();
                }
                phase = PH_BASE;
                continue;
            }
            // ---- PH_BASE
            if (hnvalid <= 0) {  // while (candidates.size() > 0) failed, or the no-dis-check step cap hit
                // ---- finish: HNSW_search.cpp:793-801 stats, ResultHandler end(), IndexHNSW.cpp:433-438 sign flip
                nhops += nstep;
                if (lane == 0) {
                    float* oD = st.outD + (size_t)qid * p.k;
                    int64_t* oI = st.outI + (size_t)qid * p.k;
                    // heap_reorder (utils/Heap.h:421-450): repeatedly pop the max to the back
                    int ii = 0;
                    for (int i = 0; i < p.k; i++) {
                        const float v = w.r_dis[0];
                        const int id = w.r_ids[0];
                        heap_pop(p.k - i, w.r_dis, w.r_ids);
                        w.r_dis[p.k - ii - 1] = v;
                        w.r_ids[p.k - ii - 1] = id;
                        if (id != -1) ii++;
                    }
                    for (int i = 0; i < p.k; i++) {
                        if (i < ii) {
                            const float v = w.r_dis[p.k - ii + i];
                            oD[i] = g.metric_ip ? -v : v;
                            oI[i] = w.r_ids[p.k - ii + i];
                        } else {
                            oD[i] = g.metric_ip ? -FLT_MAX : FLT_MAX;
                            oI[i] = -1;
                        }
                    }
                    if (st.out_ndis) st.out_ndis[qid] = ndis;
                    if (st.out_nhops) st.out_nhops[qid] = nhops;
                    atomicAdd(st.n_done, 1);
                }
                {   // fresh visited set for the next query (reference: epoch table, AuxIndexStructures.h:173-198)
                    uint4* v4 = reinterpret_cast<uint4*>(vis);
                    const int64_t n4 = st.vis_words >> 2;
                    for (int64_t i = lane; i < n4; i += 32) v4[i] = make_uint4(0, 0, 0, 0);
                }
                __syncwarp();
                phase = PH_FETCH;
                continue;
            }
            // pop the beam (HNSW_search.cpp:530-615)
            int nbeam = 0, total_neighbors = 0, nnew = 0;
            for (;;) {
                if (hnvalid <= 0) break;
                if (p.batch_size > 0) { if (!(nbeam == 0 || total_neighbors < p.batch_size)) break; }
                else if (nbeam >= p.beam) break;
                float d0 = 0.f;
                const int v0 = mm_pop_min(w, hk, hnvalid, &d0, lane);
                if (v0 < 0) break;
                if (p.check_rel) {
                    if (mm_count_below(w, hk, d0, lane) >= p.ef) break;
                }
                uint64_t b, e;
                neighbor_range(g, v0, 0, &b, &e);
                const int n = static_cast<int>(e - b);
                int cnt = 0;
                for (int i0 = 0; i0 < n; i0 += 32) {
                    const int i = i0 + lane;
                    int v1 = -1;
                    bool fresh = false;
                    if (i < n) {
                        v1 = g.nbrs[b + i];
                        fresh = ((__ldcg(&vis[v1 >> 5]) >> (v1 & 31)) & 1u) == 0;  // L2 read: bits are set with atomics
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, fresh);
                    if (nnew + __popc(m) > p.cap_new) {  // cannot happen for sizes derived in search_impl; never write past the buffer
                        if (lane == 0) atomicExch(st.error, 4);
                        break;
                    }
                    if (fresh) w.rq[nnew + __popc(m & ((1u << lane) - 1))] = v1;
                    nnew += __popc(m);
                    cnt += __popc(m);
                }
                nbeam++;
                total_neighbors = static_cast<int>(static_cast<float>(total_neighbors) +
                                                   static_cast<float>(cnt) * (p.pq_mode ? p.pq_ratio : 1.0f));  // :567-568, :610-611
                __syncwarp();
            }
            if (nbeam == 0) continue;  // HNSW_search.cpp:618-620
            // dedup across the beam in ascending id order (std::set, :623-640), mark visited (:753-759)
            int P = 1;
            while (P < nnew) P <<= 1;
            for (int i = nnew + lane; i < P; i += 32) w.rq[i] = 2147483647;
            __syncwarp();
            if (nnew > 1) {
                warp_sort(w.rq, P, lane);
                nnew = warp_unique(w.rq, nnew, lane);
            }
            if (p.pq_mode) {
                // ---- :645-750: PQ distances of the hop's unvisited neighbours, then one of three selections; only the
                // selected nodes are marked visited and scored exactly, in the order they are selected
                const float* lut = st.pq_lut + (size_t)slot * st.pq.n_chunks * 256;
                for (int i = lane; i < P; i += 32) w.pqd[i] = (i < nnew) ? pq_lookup(st.pq, lut, w.rq[i]) : FLT_MAX;
                __syncwarp();
                if (nnew > 1) warp_sort_pairs(w.pqd, w.rq, P, lane);
                int ns = 0;
                if (p.pq_mode == 2) {  // local: the closest pq_select_ratio share of this hop's neighbours (:683-707)
                    const size_t num = static_cast<size_t>(p.pq_ratio * static_cast<float>(static_cast<size_t>(nnew)));
                    ns = num < static_cast<size_t>(nnew) ? static_cast<int>(num) : nnew;
                    for (int i = lane; i < ns; i += 32) w.req[i] = w.rq[i];
                } else {
                    float* qd = st.pq_qd + (size_t)slot * st.pq_cap;
                    int* qi = st.pq_qid + (size_t)slot * st.pq_cap;
                    int head = st.pq_qhead[slot], qn = st.pq_qn[slot];
                    const int merged = pq_merge(qd, qi, head, qn, st.pq_cap, w.pqd, w.rq, nnew, p.pq_mode == 3, w.req, lane);
                    if (merged < 0) { if (lane == 0) atomicExch(st.error, 1); } else qn = merged;
                    if (p.pq_mode == 3) {  // proportional: std::set, the max(1, n_new * ratio) smallest, erased (:714-731)
                        int num = static_cast<int>(static_cast<float>(static_cast<size_t>(nnew)) * p.pq_ratio);
                        if (num < 1) num = 1;
                        ns = (qn - head) < num ? (qn - head) : num;
                        for (int i = lane; i < ns; i += 32) w.req[i] = qi[head + i];
                        head += ns;
                    } else {  // global: look at the max(1, size * ratio) smallest of everything seen, skip the visited (:732-750)
                        int num = static_cast<int>(static_cast<float>(static_cast<size_t>(qn)) * p.pq_ratio);
                        if (num < 1) num = 1;
                        if (num > qn) num = qn;
                        for (int b0 = 0; b0 < num; b0 += 32) {
                            const int i = b0 + lane;
                            bool take = false;
                            int id = -1;
                            if (i < num) {
                                id = qi[i];
                                const bool dup = i > 0 && qi[i - 1] == id;  // a re-pushed node sits next to its first copy
                                take = !dup && (((__ldcg(&vis[id >> 5]) >> (id & 31)) & 1u) == 0);
                            }
                            const unsigned m = __ballot_sync(0xffffffffu, take);
                            if (take) {
                                const int o = ns + __popc(m & ((1u << lane) - 1));
                                if (o < p.cap_req) w.req[o] = id;
                            }
                            ns += __popc(m);
                        }
                        if (ns > p.cap_req) { if (lane == 0) atomicExch(st.error, 2); ns = p.cap_req; }
                    }
                    __syncwarp();
                    if (lane == 0) { st.pq_qn[slot] = qn; st.pq_qhead[slot] = head; }
                }
                __syncwarp();
                for (int i = lane; i < ns; i += 32) {
                    const int v1 = w.req[i];
                    atomicOr(&vis[v1 >> 5], 1u << (v1 & 31));
                }
                nnew = ns;
            } else {
                for (int i = lane; i < nnew; i += 32) {
                    const int v1 = w.rq[i];
                    atomicOr(&vis[v1 >> 5], 1u << (v1 & 31));
                }
            }
            __syncwarp();
            if (nnew == 0) {  // distances_batch on an empty set; :781-787
                nstep += nbeam;
                if (!p.check_rel && nstep > p.ef) hnvalid = -1;
                continue;
            }
            pend_beam = nbeam;
            n_req = nnew;
            break;
        }
        if (phase == PH_IDLE) break;
        // =========================== emit ===========================
        if (st.recompute) {
            // claim a row of the hop's unique work list for every node nobody requested yet this hop
            for (int i = lane; i < n_req; i += 32) {
                const int node = w.req[i];
                const uint32_t old = atomicMax(&st.stamp[node], st.stamp_value);
                if (old < st.stamp_value) {
                    int len = static_cast<int>(st.tok_off[node + 1] - st.tok_off[node]);
                    len = len < st.max_pos ? len : st.max_pos;
                    const unsigned long long packed =
                        atomicAdd(st.claim, (1ull << 40) | static_cast<unsigned long long>(len));
                    const int us = static_cast<int>(packed >> 40);
                    st.uniq_node[us] = node;
                    st.seq_start[us] = static_cast<int>(packed & ((1ull << 40) - 1));
                    slot_wr[node] = st.row_base_hop + us;
                }
            }
            have_pending = false;  // distances arrive with the next launch
            break;
        }
        have_pending = true;
    }

    // ---- store slot state
    if (lane == 0) {
        st.phase[slot] = phase; st.qid[slot] = qid; st.level[slot] = level; st.nearest[slot] = nearest;
        st.prev_nearest[slot] = prev_nearest; st.d_nearest[slot] = d_nearest; st.hk[slot] = hk;
        st.hnvalid[slot] = hnvalid; st.nstep[slot] = nstep; st.pend_beam[slot] = pend_beam; st.n_req[slot] = n_req;
        st.ndis[slot] = ndis; st.nhops[slot] = nhops;
    }
    if (phase != PH_IDLE) {
        for (int i = lane; i < hk; i += 32) { g_hids[i] = w.h_ids[i]; g_hdis[i] = w.h_dis[i]; }
        for (int i = lane; i < p.k; i += 32) { g_rids[i] = w.r_ids[i]; g_rdis[i] = w.r_dis[i]; }
        for (int i = lane; i < n_req; i += 32) g_req[i] = w.req[i];
    }
}

__global__ void init_slots_kernel(TravState st) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < st.S) {
        st.phase[i] = PH_FETCH;
        st.qid[i] = -1; st.level[i] = 0; st.nearest[i] = -1; st.prev_nearest[i] = -1; st.d_nearest[i] = 0.f;
        st.hk[i] = 0; st.hnvalid[i] = 0; st.nstep[i] = 0; st.pend_beam[i] = 0; st.n_req[i] = 0;
        st.ndis[i] = 0; st.nhops[i] = 0;
    }
    if (i == 0) { *st.next_query = 0; *st.n_done = 0; }
}

}  // namespace

size_t step_smem_bytes(const TravParams& p, int d, int warps) {
    const int dpad = (d + 3) & ~3;
    const size_t per_warp = (size_t)dpad * 4 + (size_t)p.hcap * 8 + (size_t)p.k * 8 + (size_t)p.p2 * 4 + (size_t)p.cap_req * 4 +
                            (p.pq_mode ? (size_t)p.cap_req * 4 + (size_t)p.p2 * 4 : 0);
    return ((per_warp + 15) & ~size_t(15)) * warps;
}

bool launch_init_slots(cudaStream_t s, const TravState& st) {
    init_slots_kernel<<<(st.S + 127) / 128, 128, 0, s>>>(st);
    LB2_CUDA_OK(cudaGetLastError());
    return true;
}

bool launch_step(cudaStream_t s, const DevGraph& g, const TravParams& p, const TravState& st, int max_iters,
                 int num_sms) {
    (void)num_sms;
    const size_t smem = step_smem_bytes(p, g.d, STEP_WARPS);
    if (smem > 200 * 1024) {
        set_error("search parameters need %zu bytes of shared memory per CTA (efSearch/beam too large)", smem);
        return false;
    }
    if (smem > 48 * 1024)  // per device, so not cached in a process-wide static (a handle may live on any GPU)
        LB2_CUDA_OK(cudaFuncSetAttribute(hnsw_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (st.S + STEP_WARPS - 1) / STEP_WARPS;
    hnsw_step_kernel<<<grid, STEP_WARPS * 32, smem, s>>>(g, p, st, max_iters);
    LB2_CUDA_OK(cudaGetLastError());
    return true;
}

}  // namespace lb2
