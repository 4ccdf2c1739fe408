// Graph construction on the GPU — tooling around the search path, not the path itself.
//
// The reference builds its HNSW graph by inserting one point at a time: greedy descent through the
// upper levels, an efConstruction-wide best-first search on every level the point lives on
// (search_neighbors_to_add, faiss/impl/HNSW.cpp:609-720), neighbour selection with the
// "keep a candidate only if no kept neighbour is closer to it than the new point is" heuristic
// (HNSW::shrink_neighbor_list, impl/HNSW.cpp:426-468) and reverse links that are shrunk with the
// same heuristic when a list overflows (add_link, impl/HNSW.cpp:510-552; driver
// add_with_locks / hnsw_add_vertices, impl/HNSW.cpp:839-894, IndexHNSW.cpp:59-280).
// Hours at 10 M points on CPU.  Here the same construction runs batch-parallel:
//
//   build_insert_search_kernel   one warp per point of the batch: greedy descent over the (finished)
//                                upper levels, then the efConstruction search over the level-0 graph
//                                built so far.  fp16 vectors (768 B per 384-d row, gathered as 16-byte
//                                words by 8-lane groups, 4 neighbours per warp pass), fp32 accumulate;
//                                candidate list = one sorted array with an "expanded" bit per entry
//                                in shared memory (equivalent to the reference's results +
//                                candidates queues: expand the closest unexpanded entry of the best
//                                ef; stop when there is none); visited set = per-warp open-addressing
//                                hash table in HBM/L2 (no O(N) state per in-flight point).
//   build_select_kernel          the neighbour-selection heuristic over a sorted candidate row, given
//                                the pairwise candidate distances (one batched GEMM on the host side,
//                                leann_b200/graph_build.py); one warp per row, kept set in registers
//                                and shared memory, ballot over the kept entries.
//
// Both kernels are HBM/L2-latency bound gathers (integer/byte work, no tensor cores); the batch
// schedule, reverse-link merging and the conversion to the reference's CSR file live in
// leann_b200/graph_build.py.  Nothing here is on the query path.
#include <float.h>

#include <algorithm>

#include "../../include/leann_b200.h"
#include "common.cuh"

namespace lb2 {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int BS_WARPS = 8;
constexpr int EXPANDED = static_cast<int>(0x80000000u);
constexpr unsigned HASH_EMPTY = 0xffffffffu;
constexpr int MAX_EF = 256;
constexpr int MAX_NEW = 128;  // >= widest adjacency row

struct BuildArgs {
    const __half* x;  // [n, d] fp16
    long long n;
    int d;
    int metric_ip;
    const int* adj0;  // [n, cap0], -1 padded
    int cap0;
    const int* up_row;  // [n] first row of the node's level-1 list in up_adj, -1 when the node lives on level 0 only
    const int* up_adj;  // [rows, capU]
    int capU;
    int entry;
    int max_level;      // level of the entry point (0-based)
    const int* points;  // [npts] query points (their vectors are read from x)
    long long npts;
    int ef;
    int* out_ids;     // [npts, ef] ascending distance, -1 padded
    float* out_dist;  // [npts, ef]
    unsigned* hash;   // [warps in the grid][1 << hash_bits]
    int hash_bits;
    unsigned long long* counter;
};

// distance of q (fp32 in shared memory) to row `node`, computed by the 8 lanes of one group;
// every lane of the group returns the full sum.  d % 8 == 0.
__device__ __forceinline__ float group_dist(const float* q_s, const __half* __restrict__ x, long long node, int d,
                                            int metric_ip, int gl) {
    const uint4* row = reinterpret_cast<const uint4*>(x + node * d);
    const int chunks = d >> 3;
    float acc = 0.f;
    for (int c = gl; c < chunks; c += 8) {
        const uint4 v = __ldg(row + c);
        const float4 q0 = *reinterpret_cast<const float4*>(q_s + c * 8);
        const float4 q1 = *reinterpret_cast<const float4*>(q_s + c * 8 + 4);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
        const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&v.z));
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
        if (metric_ip) {
            acc = fmaf(q0.x, a.x, acc); acc = fmaf(q0.y, a.y, acc); acc = fmaf(q0.z, b.x, acc); acc = fmaf(q0.w, b.y, acc);
            acc = fmaf(q1.x, e.x, acc); acc = fmaf(q1.y, e.y, acc); acc = fmaf(q1.z, f.x, acc); acc = fmaf(q1.w, f.y, acc);
        } else {
            float t;
            t = q0.x - a.x; acc = fmaf(t, t, acc); t = q0.y - a.y; acc = fmaf(t, t, acc);
            t = q0.z - b.x; acc = fmaf(t, t, acc); t = q0.w - b.y; acc = fmaf(t, t, acc);
            t = q1.x - e.x; acc = fmaf(t, t, acc); t = q1.y - e.y; acc = fmaf(t, t, acc);
            t = q1.z - f.x; acc = fmaf(t, t, acc); t = q1.w - f.y; acc = fmaf(t, t, acc);
        }
    }
    acc += __shfl_xor_sync(FULL, acc, 1);
    acc += __shfl_xor_sync(FULL, acc, 2);
    acc += __shfl_xor_sync(FULL, acc, 4);
    return metric_ip ? -acc : acc;
}

// distances of ids[0..n) -> dist[0..n), 4 per warp pass
__device__ __forceinline__ void warp_dists(const float* q_s, const BuildArgs& a, const int* ids, float* dist, int n, int lane) {
    const int grp = lane >> 3, gl = lane & 7;
    for (int i0 = 0; i0 < n; i0 += 4) {
        const int i = i0 + grp;
        const int node = (i < n) ? ids[i] : ids[n - 1];
        const float dd = group_dist(q_s, a.x, node, a.d, a.metric_ip, gl);
        if (i < n && gl == 0) dist[i] = dd;
    }
    __syncwarp();
}

// visited.insert(id): true when the id was not in the table
__device__ __forceinline__ bool hash_insert(unsigned* tab, unsigned mask, int bits, unsigned id) {
    unsigned h = (id * 2654435761u) >> (32 - bits);
    for (unsigned probe = 0; probe <= mask; probe++) {
        const unsigned old = atomicCAS(&tab[h], HASH_EMPTY, id);
        if (old == HASH_EMPTY) return true;
        if (old == id) return false;
        h = (h + 1) & mask;
    }
    return false;  // table full: treat as visited
}

// insert (id, dist) into the ascending list l_*[0..sz) of capacity ef; whole warp
__device__ __forceinline__ void list_insert(int* l_id, float* l_d, int& sz, int ef, int id, float dist, int lane) {
    if (sz == ef && !(dist < l_d[ef - 1])) return;
    int cnt = 0;
    for (int i = lane; i < sz; i += 32) cnt += (l_d[i] <= dist) ? 1 : 0;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
    const int pos = cnt;
    const int last = (sz < ef) ? sz : ef - 1;  // entries [pos, last) move one place right, top chunk first
    for (int hi = last; hi > pos; hi -= 32) {
        const int i = hi - 1 - lane;
        int ti = 0;
        float td = 0.f;
        if (i >= pos) { ti = l_id[i]; td = l_d[i]; }
        __syncwarp();
        if (i >= pos) { l_id[i + 1] = ti; l_d[i + 1] = td; }
        __syncwarp();
    }
    if (lane == 0) { l_id[pos] = id; l_d[pos] = dist; }
    sz = last + 1;
    __syncwarp();
}

__global__ void __launch_bounds__(BS_WARPS * 32)
build_insert_search_kernel(const BuildArgs a) {
    extern __shared__ __align__(16) uint8_t bs_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = ((size_t)a.d * 4 + (size_t)a.ef * 8 + MAX_NEW * 8 + 15) & ~size_t(15);
    uint8_t* base = bs_smem + per_warp * warp;
    float* q_s = reinterpret_cast<float*>(base);
    int* l_id = reinterpret_cast<int*>(q_s + a.d);
    float* l_d = reinterpret_cast<float*>(l_id + a.ef);
    int* nw_id = reinterpret_cast<int*>(l_d + a.ef);
    float* nw_d = reinterpret_cast<float*>(nw_id + MAX_NEW);
    const unsigned hmask = (1u << a.hash_bits) - 1;
    unsigned* tab = a.hash + ((size_t)(blockIdx.x * BS_WARPS + warp) << a.hash_bits);
    const unsigned hash_limit = (hmask + 1) - ((hmask + 1) >> 2);  // stop expanding at 75 % load

    for (;;) {
        long long pi = 0;
        if (lane == 0) pi = static_cast<long long>(atomicAdd(a.counter, 1ull));
        pi = __shfl_sync(FULL, pi, 0);
        if (pi >= a.npts) break;
        const int me = a.points[pi];
        {   // query vector -> fp32 in shared memory; fresh visited table
            const __half* xr = a.x + (long long)me * a.d;
            for (int j = lane; j < a.d; j += 32) q_s[j] = __half2float(xr[j]);
            uint4* t4 = reinterpret_cast<uint4*>(tab);
            const int n4 = (hmask + 1) >> 2;
            for (int i = lane; i < n4; i += 32) t4[i] = make_uint4(HASH_EMPTY, HASH_EMPTY, HASH_EMPTY, HASH_EMPTY);
        }
        __syncwarp();
        // ---- greedy descent over the upper levels (impl/HNSW.cpp:868-871 -> greedy_update_nearest)
        int nearest = a.entry;
        float d_nearest;
        if (lane == 0) nw_id[0] = nearest;
        __syncwarp();
        warp_dists(q_s, a, nw_id, nw_d, 1, lane);
        d_nearest = nw_d[0];
        for (int level = a.max_level; level >= 1; level--) {
            for (;;) {
                const int r0 = a.up_row[nearest];
                int n = 0;
                if (r0 >= 0) {
                    const int* row = a.up_adj + (long long)(r0 + level - 1) * a.capU;
                    for (int i0 = 0; i0 < a.capU; i0 += 32) {
                        const int i = i0 + lane;
                        const int v = (i < a.capU) ? row[i] : -1;
                        const unsigned m = __ballot_sync(FULL, v >= 0);
                        if (v >= 0) nw_id[n + __popc(m & ((1u << lane) - 1))] = v;
                        n += __popc(m);
                    }
                }
                __syncwarp();
                if (n == 0) break;
                warp_dists(q_s, a, nw_id, nw_d, n, lane);
                int best = -1;
                float bd = d_nearest;
                for (int i = 0; i < n; i++) {
                    const float dd = nw_d[i];
                    if (dd < bd) { bd = dd; best = nw_id[i]; }
                }
                __syncwarp();
                if (best < 0) break;
                nearest = best;
                d_nearest = bd;
            }
        }
        // ---- level 0: efConstruction search over the graph built so far
        int sz = 0;
        unsigned n_vis = 1;
        if (lane == 0) hash_insert(tab, hmask, a.hash_bits, (unsigned)nearest);
        list_insert(l_id, l_d, sz, a.ef, nearest, d_nearest, lane);
        for (;;) {
            int pos = -1;
            for (int b0 = 0; b0 < sz && pos < 0; b0 += 32) {
                const int i = b0 + lane;
                const bool un = (i < sz) && ((l_id[i] & EXPANDED) == 0);
                const unsigned m = __ballot_sync(FULL, un);
                if (m) pos = b0 + __ffs(m) - 1;
            }
            if (pos < 0 || n_vis > hash_limit) break;
            const int cur = l_id[pos];
            __syncwarp();
            if (lane == 0) l_id[pos] = cur | EXPANDED;
            const int* row = a.adj0 + (long long)cur * a.cap0;
            int nnew = 0;
            for (int i0 = 0; i0 < a.cap0; i0 += 32) {
                const int i = i0 + lane;
                const int v = (i < a.cap0) ? __ldcg(row + i) : -1;
                const bool fresh = (v >= 0) && (v != me) && hash_insert(tab, hmask, a.hash_bits, (unsigned)v);
                const unsigned m = __ballot_sync(FULL, fresh);
                if (fresh) nw_id[nnew + __popc(m & ((1u << lane) - 1))] = v;
                nnew += __popc(m);
            }
            n_vis += nnew;
            __syncwarp();
            if (nnew == 0) continue;
            warp_dists(q_s, a, nw_id, nw_d, nnew, lane);
            for (int i = 0; i < nnew; i++) list_insert(l_id, l_d, sz, a.ef, nw_id[i], nw_d[i], lane);
        }
        // results: ascending, without the point itself (it is in the list when it already lives in the graph: repair sweeps)
        int* oi = a.out_ids + pi * a.ef;
        float* od = a.out_dist + pi * a.ef;
        int wr = 0;
        for (int b0 = 0; b0 < sz; b0 += 32) {
            const int i = b0 + lane;
            const int id = (i < sz) ? (l_id[i] & ~EXPANDED) : me;
            const bool keep = id != me;
            const unsigned m = __ballot_sync(FULL, keep);
            if (keep) {
                const int o = wr + __popc(m & ((1u << lane) - 1));
                oi[o] = id;
                od[o] = l_d[i];
            }
            wr += __popc(m);
        }
        for (int i = wr + lane; i < a.ef; i += 32) { oi[i] = -1; od[i] = FLT_MAX; }
        __syncwarp();
    }
}

// HNSW::shrink_neighbor_list over a sorted candidate row (closest first; invalid entries = -1 at the end).
//   pd [b, K, K]  pairwise candidate distances (fp16 or fp32),  dn [b, K] distance of the node to each candidate.
// Keeps candidate j iff no already kept i has pd[j][i] < dn[j]; stops at `keep`.  fill > 0: a list that ends with fewer than
// `fill` links is topped up with the nearest rejected candidates ("outsiders", the keep_max_size_level0 branch of the
// reference, :459-467, with a settable floor instead of max_size).  Output ids/dist [b, keep], -1 / FLT_MAX padded.
template <typename PT>
__global__ void __launch_bounds__(128)
build_select_kernel(const PT* __restrict__ pd, const float* __restrict__ dn, const int* __restrict__ cand, long long b, int K,
                    int keep, int fill, int* __restrict__ out_ids, float* __restrict__ out_dist) {
    __shared__ int s_kept[4][MAX_NEW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 4 + warp;
    if (row >= b) return;
    int* kept = s_kept[warp];
    const PT* p = pd + row * K * K;
    const float* dr = dn + row * K;
    const int* cr = cand + row * K;
    int nk = 0;
    for (int j = 0; j < K && nk < keep; j++) {
        const int id = cr[j];
        if (id < 0) break;
        const float dj = dr[j];
        bool dom = false;
        for (int i = lane; i < nk; i += 32) dom |= static_cast<float>(p[(long long)j * K + kept[i]]) < dj;
        if (!__any_sync(FULL, dom)) {
            if (lane == 0) {
                kept[nk] = j;
                out_ids[row * keep + nk] = id;
                out_dist[row * keep + nk] = dj;
            }
            nk++;
        }
        __syncwarp();
    }
    if (nk < fill) {  // outsiders, nearest first; kept[] is ascending, so one merge-like scan finds the rejected ones
        const int nk0 = nk;
        int ki = 0;
        for (int j = 0; j < K && nk < fill; j++) {
            const int id = cr[j];
            if (id < 0) break;
            while (ki < nk0 && kept[ki] < j) ki++;
            if (ki < nk0 && kept[ki] == j) continue;
            if (lane == 0) {
                out_ids[row * keep + nk] = id;
                out_dist[row * keep + nk] = dr[j];
            }
            nk++;
        }
        __syncwarp();
    }
    for (int i = nk + lane; i < keep; i += 32) {
        out_ids[row * keep + i] = -1;
        out_dist[row * keep + i] = FLT_MAX;
    }
}

// visited-table entries per warp: at most ef expansions x cap0 neighbours can be marked; the search stops at 75 % load
int visited_bits(int ef, int cap0) {
    int bits = 10;
    while ((1ll << bits) < (long long)ef * cap0 && bits < 20) bits++;
    return bits;
}

}  // namespace
}  // namespace lb2

using namespace lb2;

extern "C" {

int lb2_build_insert_search(const void* d_x_f16, int64_t n, int32_t d, int32_t metric_ip, const int32_t* d_adj0,
                            int32_t cap0, const int32_t* d_up_row, const int32_t* d_up_adj, int32_t capU, int32_t entry,
                            int32_t max_level, const int32_t* d_points, int64_t npts, int32_t ef, int32_t* d_out_ids,
                            float* d_out_dist, void* d_workspace, size_t workspace_bytes) {
    if (!d_x_f16 || !d_adj0 || !d_points || !d_out_ids || !d_out_dist || n <= 0 || npts < 0 || d <= 0 || (d & 7) ||
        ef < 1 || ef > MAX_EF || cap0 < 1 || cap0 > MAX_NEW || capU > MAX_NEW || entry < 0 || entry >= n ||
        (max_level > 0 && (!d_up_row || !d_up_adj))) {
        set_error("lb2_build_insert_search: bad arguments (d must be a multiple of 8, ef <= %d, row widths <= %d)", MAX_EF, MAX_NEW);
        return LB2_ERR_ARG;
    }
    if (npts == 0) return LB2_OK;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        set_error("lb2_build_insert_search: no CUDA device");
        return LB2_ERR_CUDA;
    }
    const int bits = visited_bits(ef, cap0);
    const size_t smem = (((size_t)d * 4 + (size_t)ef * 8 + MAX_NEW * 8 + 15) & ~size_t(15)) * BS_WARPS;
    if (smem > 200 * 1024) {
        set_error("lb2_build_insert_search: d / ef need %zu bytes of shared memory per CTA", smem);
        return LB2_ERR_ARG;
    }
    int per_sm = static_cast<int>(std::min<size_t>(4, (220 * 1024) / smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sms * per_sm;
    grid = std::min<long long>(grid, (npts + BS_WARPS - 1) / BS_WARPS);
    const size_t need = 16 + (size_t)grid * BS_WARPS * (sizeof(unsigned) << bits);
    if (!d_workspace || workspace_bytes < need) {
        set_error("lb2_build_insert_search: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
        return LB2_ERR_ARG;
    }
    BuildArgs a;
    a.x = static_cast<const __half*>(d_x_f16); a.n = n; a.d = d; a.metric_ip = metric_ip; a.adj0 = d_adj0; a.cap0 = cap0;
    a.up_row = d_up_row; a.up_adj = d_up_adj; a.capU = capU; a.entry = entry; a.max_level = max_level; a.points = d_points;
    a.npts = npts; a.ef = ef; a.out_ids = d_out_ids; a.out_dist = d_out_dist;
    a.counter = static_cast<unsigned long long*>(d_workspace);
    a.hash = reinterpret_cast<unsigned*>(static_cast<uint8_t*>(d_workspace) + 16);
    a.hash_bits = bits;
    if (cudaMemsetAsync(d_workspace, 0, 16, 0) != cudaSuccess) { set_error("cudaMemsetAsync failed"); return LB2_ERR_CUDA; }
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(build_insert_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    build_insert_search_kernel<<<(unsigned)grid, BS_WARPS * 32, smem, 0>>>(a);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("build_insert_search_kernel launch failed: %s", cudaGetErrorString(e)); return LB2_ERR_CUDA; }
    return LB2_OK;
}

size_t lb2_build_workspace_bytes(int32_t ef, int32_t cap0) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return 16 + (size_t)sms * 4 * BS_WARPS * (sizeof(unsigned) << visited_bits(ef, cap0));  // at most 4 CTAs per SM
}

int lb2_build_select(const void* d_pd, int32_t pd_is_f32, const float* d_dn, const int32_t* d_cand, int64_t b, int32_t K,
                     int32_t keep, int32_t fill, int32_t* d_out_ids, float* d_out_dist) {
    if (!d_pd || !d_dn || !d_cand || !d_out_ids || !d_out_dist || b < 0 || K < 1 || keep < 1 || keep > MAX_NEW || fill < 0 ||
        fill > keep) {
        set_error("lb2_build_select: bad arguments (keep <= %d)", MAX_NEW);
        return LB2_ERR_ARG;
    }
    if (b == 0) return LB2_OK;
    const unsigned grid = static_cast<unsigned>((b + 3) / 4);
    if (pd_is_f32)
        build_select_kernel<float><<<grid, 128>>>(static_cast<const float*>(d_pd), d_dn, d_cand, b, K, keep, fill, d_out_ids, d_out_dist);
    else
        build_select_kernel<__half><<<grid, 128>>>(static_cast<const __half*>(d_pd), d_dn, d_cand, b, K, keep, fill, d_out_ids, d_out_dist);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("build_select_kernel launch failed: %s", cudaGetErrorString(e)); return LB2_ERR_CUDA; }
    return LB2_OK;
}

}  // extern "C"
