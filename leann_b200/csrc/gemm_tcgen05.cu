// Recompute stage, dense part: C[M,N] = epilogue(A[M,K] . W[N,K]^T + bias) on the
// 5th-generation tensor cores (tcgen05.mma, fp16 operands, fp32 accumulation in TMEM).
//
// This is the arithmetic the reference runs inside SentenceTransformer.encode
// (packages/leann-core/src/leann/embedding_compute.py:231-239; fp16 weights and
// activations, :157-158) for every hop's batch of neighbour chunks: the four
// nn.Linear layers of each BERT block (QKV, attention output, FFN up, FFN down).
//
// Kernel shape (one persistent CTA per SM, 384 threads, warp-specialised):
//   warp 0 lane 0 : TMA producer   (cp.async.bulk.tensor 2D, 128B swizzle, 4-stage ring)
//   warp 1 lane 0 : MMA issuer     (tcgen05.mma cta_group::1 kind::f16, M=128, N=192, K=16)
//   warp 2        : TMEM allocator (2 accumulator stages x 192 columns -> 512 columns)
//   warps 4..11   : epilogue       (tcgen05.ld 32x32b, register double-buffered -> bias / GELU /
//                                   residual -> fp16 -> global); two warps per TMEM lane quarter
// Pipelines: smem full/empty mbarriers between TMA and MMA, TMEM full/empty mbarriers
// between MMA and epilogue, so the epilogue of tile i overlaps the MMAs of tile i+1.
// M is ragged (varlen-packed tokens): TMA zero-fills rows past M, stores are row-masked.
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace lb2 {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 fp16 = 128 B = one swizzle span
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 384;  // 4 control warps + 8 epilogue warps
constexpr int EPI_WARPS = 8;

template <int BLOCK_N, int STAGES>
struct GemmSmem {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024 /*alignment slack*/;
};

// HF "gelu" (erf form: x * Phi(x)), the activation of all-MiniLM-L6-v2 / bge-base BERT blocks.
// The K=384 GEMMs leave ~12 issue slots per output element before the epilogue, not the tensor
// pipe, becomes the limiter, and erff() costs ~30.  Phi(x) = 0.5 (1 + tanh(x (c0 + c1 x^2 + c2 x^4)))
// with minimax-fitted coefficients reproduces erf-GELU to 2.5e-5 absolute / 4.5e-4 relative
// (|gelu| > 0.05) — below the fp16 rounding of the stored activation — in 7 FP ops + 1 MUFU.
__device__ __forceinline__ float gelu_fast(float x) {
    const float x2 = fminf(x * x, 64.0f);  // the fit is monotone up to |x| = 8, tanh saturated long before
    float p = fmaf(x2, -3.51516789e-04f, 3.70056460e-02f);
    p = fmaf(x2, p, 7.97507884e-01f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * p));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}

template <int BLOCK_N, int STAGES, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_f16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   __half* __restrict__ C, const float* __restrict__ bias, const __half* __restrict__ residual,
                   int M, int N, int K) {
    using L = GemmSmem<BLOCK_N, STAGES>;
    constexpr int TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                            : (2 * BLOCK_N <= 256) ? 256 : 512;
    static_assert(2 * BLOCK_N <= 512, "two accumulator stages must fit TMEM");
    static_assert(BLOCK_N % 64 == 0 && BLOCK_N % 16 == 0 && BLOCK_N <= 256, "UMMA N constraint / two column halves");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_m = (M + BLOCK_M - 1) / BLOCK_M;
    const int num_n = N / BLOCK_N;
    const int num_k = K / BLOCK_K;
    const int num_tiles = num_m * num_n;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; i++) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            ptx::mbar_init(&tmem_full[i], 1);
            ptx::mbar_init(&tmem_empty[i], EPI_WARPS);  // one arrive per epilogue warp
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_ptr, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / num_n, n_blk = tile % num_n;
                for (int kb = 0; kb < num_k; kb++) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * L::STAGE_BYTES;
                    uint8_t* sb = sa + L::A_BYTES;
                    ptx::mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                    ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
                    ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = ptx::make_idesc_f16(BLOCK_M, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                ptx::mbar_wait(&tmem_empty[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BLOCK_N;
                for (int kb = 0; kb < num_k; kb++) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * L::STAGE_BYTES);
                    const uint64_t a_desc = ptx::make_sw128_kmajor_desc(sa);
                    const uint64_t b_desc = ptx::make_sw128_kmajor_desc(sa + L::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; k++) {
                        // advance 16 fp16 = 32 B along K inside the 128B swizzle span: +2 in 16B units
                        ptx::umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    }
                    ptx::umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tmem_full[as]);  // accumulator complete
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: 8 warps, two per TMEM lane quarter, each owning half of the tile's columns =====
        const int quarter = warp & 3;          // TMEM lanes [32*quarter, +32) are the ones this warp may read
        const int half = (warp - 4) >> 2;      // column half of the accumulator
        constexpr int NCH = BLOCK_N / 32 / 2;  // 32-column chunks per warp
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
            const int m_blk = tile / num_n, n_blk = tile % num_n;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            ptx::mbar_wait(&tmem_full[as], aphase);
            ptx::tc_fence_after();
            const int row = m_blk * BLOCK_M + quarter * 32 + lane;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BLOCK_N + half * NCH * 32;
            const int colbase = n_blk * BLOCK_N + half * NCH * 32;
            uint32_t r[2][32];
            ptx::tmem_ld_32x32(taddr, r[0]);
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                ptx::tmem_ld_wait();
                if (c + 1 < NCH) ptx::tmem_ld_32x32(taddr + (c + 1) * 32, r[(c + 1) & 1]);  // overlaps the math below
                const uint32_t(&acc)[32] = r[c & 1];
                const int col0 = colbase + c * 32;
                if (row < M) {
                    const size_t off = static_cast<size_t>(row) * N + col0;
                    uint4 resv[4];
                    if (EPI == EPI_BIAS_RES) {
                        const uint4* rp = reinterpret_cast<const uint4*>(residual + off);
#pragma unroll
                        for (int j = 0; j < 4; j++) resv[j] = __ldg(rp + j);
                    }
                    uint4 outv[4];
                    __half2* oh = reinterpret_cast<__half2*>(outv);
                    const __half2* rh = reinterpret_cast<const __half2*>(resv);
                    const float4* bp = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float4 b4 = __ldg(bp + j);  // warp-uniform address: one broadcast transaction
                        float v0 = __uint_as_float(acc[4 * j]) + b4.x;
                        float v1 = __uint_as_float(acc[4 * j + 1]) + b4.y;
                        float v2 = __uint_as_float(acc[4 * j + 2]) + b4.z;
                        float v3 = __uint_as_float(acc[4 * j + 3]) + b4.w;
                        if (EPI == EPI_BIAS_GELU) {
                            v0 = gelu_fast(v0); v1 = gelu_fast(v1); v2 = gelu_fast(v2); v3 = gelu_fast(v3);
                        }
                        if (EPI == EPI_BIAS_RES) {
                            const float2 ra = __half22float2(rh[2 * j]), rb = __half22float2(rh[2 * j + 1]);
                            v0 += ra.x; v1 += ra.y; v2 += rb.x; v3 += rb.y;
                        }
                        oh[2 * j] = __floats2half2_rn(v0, v1);
                        oh[2 * j + 1] = __floats2half2_rn(v2, v3);
                    }
                    uint4* cp = reinterpret_cast<uint4*>(C + off);
#pragma unroll
                    for (int j = 0; j < 4; j++) cp[j] = outv[j];
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty[as]);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace

// Row-major [rows, cols] fp16 matrix, box = [box_rows, 64 cols], 128B swizzle.
bool make_tmap_f16_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return false;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {BLOCK_K, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu", (int)r, (unsigned long long)rows,
                  (unsigned long long)cols);
        return false;
    }
    return true;
}

constexpr int GEMM_BLOCK_N = 192;
constexpr int GEMM_STAGES = 4;

template <int EPI>
static cudaError_t launch_gemm(cudaStream_t stream, const CUtensorMap& ta, const CUtensorMap& tb, __half* C,
                               const float* bias, const __half* residual, int M, int N, int K, int num_sms) {
    using L = GemmSmem<GEMM_BLOCK_N, GEMM_STAGES>;
    auto kern = gemm_f16_tn_kernel<GEMM_BLOCK_N, GEMM_STAGES, EPI>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int tiles = ((M + BLOCK_M - 1) / BLOCK_M) * (N / GEMM_BLOCK_N);
    const int grid = tiles < num_sms ? tiles : num_sms;
    kern<<<grid, GEMM_THREADS, L::TOTAL, stream>>>(ta, tb, C, bias, residual, M, N, K);
    return cudaGetLastError();
}

// A [M,K] fp16 row-major, W [N,K] fp16 row-major (nn.Linear layout), C [M,N] fp16.
bool gemm_f16(cudaStream_t stream, const __half* A, const CUtensorMap* tmap_w, const __half* W, const float* bias,
              const __half* residual, __half* C, int M, int N, int K, int epi, int num_sms) {
    if (M <= 0) return true;
    if (N % GEMM_BLOCK_N != 0 || K % BLOCK_K != 0) {
        set_error("gemm_f16: unsupported shape N=%d K=%d (need N%%192==0, K%%64==0)", N, K);
        return false;
    }
    CUtensorMap ta, tb_local;
    if (!make_tmap_f16_2d(&ta, A, (uint64_t)M, (uint64_t)K, BLOCK_M)) return false;
    if (!tmap_w) {
        if (!make_tmap_f16_2d(&tb_local, W, (uint64_t)N, (uint64_t)K, GEMM_BLOCK_N)) return false;
        tmap_w = &tb_local;
    }
    cudaError_t e;
    switch (epi) {
        case EPI_BIAS: e = launch_gemm<EPI_BIAS>(stream, ta, *tmap_w, C, bias, residual, M, N, K, num_sms); break;
        case EPI_BIAS_GELU: e = launch_gemm<EPI_BIAS_GELU>(stream, ta, *tmap_w, C, bias, residual, M, N, K, num_sms); break;
        case EPI_BIAS_RES: e = launch_gemm<EPI_BIAS_RES>(stream, ta, *tmap_w, C, bias, residual, M, N, K, num_sms); break;
        default: set_error("gemm_f16: bad epilogue %d", epi); return false;
    }
    if (e != cudaSuccess) {
        set_error("gemm_f16 launch: %s", cudaGetErrorString(e));
        return false;
    }
    return true;
}

int gemm_block_n() { return GEMM_BLOCK_N; }

}  // namespace lb2
