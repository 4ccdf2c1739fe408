// Microbenchmark: issue rate of tcgen05.mma kind::f16 from shared memory, no TMA traffic, no epilogue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I leann_b200/csrc scripts/umma_rate.cu -o /tmp/umma_rate
// Prints tensor clocks per MMA for cta_group::1 (M = 128) and cta_group::2 (M = 256) at several N, all SMs busy.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace lb2;

constexpr int BG_BYTES = 32768;
constexpr int SMEM = 128 * 128 + 256 * 128 + 1024 + 64 + BG_BYTES;  // A, B, barrier, background-traffic region  // A: 128 rows x 128 B, B: up to 256 rows x 128 B

template <bool PAIR>
__device__ void body(int N, int iters, int bg, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 128 * 128 + 256 * 128);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
    volatile uint32_t* done = tptr + 1;
    uint8_t* bgbuf = smem + 128 * 128 + 256 * 128 + 1024;
    const int warp = threadIdx.x >> 5;
    const int rank = PAIR ? ptx::cluster_ctarank() : 0;
    for (int i = threadIdx.x; i < (128 * 128 + 256 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); *done = 0; }
    if (warp == 0) {
        if (PAIR) { ptx::tmem_alloc_pair(tptr, 512); ptx::tmem_relinquish_pair(); }
        else { ptx::tmem_alloc(tptr, 512); ptx::tmem_relinquish(); }
    }
    ptx::fence_async_smem();
    ptx::tc_fence_before();
    if (PAIR) ptx::cluster_sync(); else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tm = *tptr;
    long long t0 = 0, t1 = 0;
    if (warp == 1) {
        const uint32_t idesc = ptx::make_idesc_f16(PAIR ? 256 : 128, N);
        const uint64_t a = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem));
        const uint64_t b = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + 128 * 128));
        t0 = clock64();
        if (rank == 0 && ptx::elect_one()) {
            for (int it = 0; it < iters; it++) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (PAIR) ptx::umma_f16_pair(tm + (it & 1) * 256, a + 2 * k, b + 2 * k, idesc, 1);
                    else ptx::umma_f16(tm + (it & 1) * 256, a + 2 * k, b + 2 * k, idesc, 1);
                }
            }
            if (PAIR) ptx::umma_commit_pair(bar); else ptx::umma_commit(bar);
        }
        __syncwarp();
        ptx::mbar_wait(bar, 0);
        t1 = clock64();
        if ((threadIdx.x & 31) == 0) { out[blockIdx.x] = t1 - t0; *done = 1; }
    } else if (warp >= 4 && bg >= 3) {
        // background tensor-memory reads (what a GEMM epilogue does while the next tile's MMAs run): 4 or 8 warps, each
        // reading 32 lanes x 32 columns of fp32 per tcgen05.ld from the accumulator buffer the MMAs are NOT writing... both
        // buffers alternate, so in effect from both
        if (bg == 3 && warp >= 8) {}
        else {
            const uint32_t taddr = tm + (static_cast<uint32_t>((warp & 3) * 32) << 16);
            uint32_t r[32];
            long long n = 0;
            uint32_t acc = 0;
            while (!*done) {
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    ptx::tmem_ld_32x32(taddr + c * 32, r);
                    ptx::tmem_ld_wait();
                    acc += r[0] ^ r[31];
                }
                n += 8;
            }
            if ((threadIdx.x & 31) == 0 && warp == 4) out[512 + blockIdx.x] = n + (acc == 12345);
        }
    } else if (warp >= 2 && warp < 4 && bg && bg < 3) {
        // background shared-memory traffic (what TMA fills and the epilogue's staging do in the GEMM): 16 B per lane per access
        const uint32_t base = ptx::smem_u32(bgbuf) + (threadIdx.x - 64) * 16;  // warps 2 and 3
        uint32_t x0 = 1, x1 = 2, x2 = 3, x3 = 4;
        long long n = 0;
        while (!*done) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const uint32_t addr = base + ((j * 2048) & (BG_BYTES - 1));
                if (bg == 1) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(addr));
                else asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(x0), "r"(x1), "r"(x2), "r"(x3) : "memory");
            }
            n += 16;
        }
        if ((threadIdx.x & 31) == 0 && warp == 2) out[512 + blockIdx.x] = n + (x0 == 12345);
    }
    ptx::tc_fence_before();
    if (PAIR) ptx::cluster_sync(); else __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        if (PAIR) ptx::tmem_dealloc_pair(tm, 512); else ptx::tmem_dealloc(tm, 512);
    }
}
__global__ void __launch_bounds__(384, 1) k1(int N, int iters, int bg, long long* out) { body<false>(N, iters, bg, out); }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1) k2(int N, int iters, int bg, long long* out) { body<true>(N, iters, bg, out); }

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* d; cudaMalloc(&d, sizeof(long long) * 1024); cudaMemset(d, 0, sizeof(long long) * 1024);
    cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    const int iters = 20000;
    for (int bg = 0; bg < 5; bg++)
    for (int pair = 0; pair < 2; pair++)
        for (int N : {64, 128, 192, 256}) {
            if (bg && N != 192 && N != 256) continue;
            const int grid = pair ? (sms & ~1) : sms;
            float ms = 0; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (pair) k2<<<grid, 384, SMEM>>>(N, iters, bg, d); else k1<<<grid, 384, SMEM>>>(N, iters, bg, d);
                cudaEventRecord(e1);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                cudaEventElapsedTime(&ms, e0, e1);
            }
            long long h[1024]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            double cyc = 0; int n = 0;
            for (int i = 0; i < grid; i += pair ? 2 : 1) { cyc += (double)h[i]; n++; }
            cyc /= n;
            const double mmas = 4.0 * iters;
            const double flops = 2.0 * (pair ? 256.0 : 128.0) * N * 16 * mmas * (pair ? grid / 2 : grid);
            // smem modes: 2 warps x 32 lanes x 16 B per access; TMEM modes: 4 / 8 warps x 4 KB per tcgen05.ld
            const double bgb = bg == 0 ? 0.0 : bg < 3 ? (double)h[512] * 64 * 16 / cyc : (double)h[512] * (bg == 3 ? 4 : 8) * 4096 / cyc;
            printf("cta_group::%d M=%d N=%3d bg=%s: %7.1f clk per MMA (%.0f MAC/clk/SM), %.3f ms -> %.0f TFLOP/s, %.2f GHz, background %.0f B/clk/SM\n",
                   pair + 1, pair ? 256 : 128, N, bg == 0 ? "none" : bg == 1 ? "ld.shared" : bg == 2 ? "st.shared" : bg == 3 ? "tcgen05.ld x4 warps" : "tcgen05.ld x8 warps", cyc / mmas,
                   128.0 * N * 16 / (cyc / mmas), ms, flops / ms / 1e9, cyc / ms / 1e6, bgb);
        }
    return 0;
}
