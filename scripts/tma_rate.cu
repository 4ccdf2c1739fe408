// Microbenchmark: TMA load throughput L2 -> shared memory per SM with all SMs loading, as a function of the bytes in flight.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I leann_b200/csrc scripts/tma_rate.cu -o /tmp/tma_rate -lcuda
// Every CTA streams [rows x 64] fp16 boxes (SWIZZLE_128B, 128 B rows) of a matrix that fits L2 (or not: `mb` argument).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace lb2;

__global__ void __launch_bounds__(64, 1) k(const __grid_constant__ CUtensorMap map, int stages, int box_rows, int iters, int row_blocks,
                                           long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
    const int box_bytes = box_rows * 128;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; i++) ptx::mbar_init(&bars[i], 1);
        ptx::fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        // whole warp, one elected issuer; cheap index arithmetic (the loop must not be the limiter)
        const long long t0 = clock64();
        int s = 0, par = 0, blk = (blockIdx.x * 37) % row_blocks, kc = 0;
        for (int i = 0; i < iters + stages; i++) {
            if (i >= stages) ptx::mbar_wait(&bars[s], par ^ 1);
            if (i < iters && ptx::elect_one()) {
                ptx::mbar_expect_tx(&bars[s], box_bytes);
                ptx::tma_load_2d(smem + s * box_bytes, &map, &bars[s], kc * 64, blk * box_rows);
            }
            __syncwarp();
            if (++kc == 6) { kc = 0; blk += 149; if (blk >= row_blocks) blk -= row_blocks; }
            if (++s == stages) { s = 0; par ^= 1; }
        }
        if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
    }
}

int main(int argc, char** argv) {
    const int mb = argc > 1 ? atoi(argv[1]) : 64;   // matrix size in MB (64: lives in L2; 2048: streams from HBM)
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t rows = (size_t)mb * 1024 * 1024 / 768;
    void* d; cudaMalloc(&d, rows * 768); cudaMemset(d, 0, rows * 768);
    long long* out; cudaMalloc(&out, 8 * 1024);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
    printf("matrix %d MB, %d SMs\n", mb, sms);
    for (int box_rows : {64, 128, 256})
        for (int stages : {1, 2, 4, 8, 12}) {
            if (stages * box_rows * 128 > 200 * 1024) continue;
            CUtensorMap map;
            cuuint64_t dims[2] = {384, rows}; cuuint64_t strides[1] = {768}; cuuint32_t box[2] = {64, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
            if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
            const int iters = 4000;
            for (int rep = 0; rep < 2; rep++) {
                k<<<sms, 64, 202 * 1024>>>(map, stages, box_rows, iters, (int)(rows / box_rows), out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            long long h[256]; cudaMemcpy(h, out, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
            double cyc = 0; for (int i = 0; i < sms; i++) cyc += (double)h[i]; cyc /= sms;
            const double bytes = (double)iters * box_rows * 128;
            printf("box %3d rows (%2d KB) x %2d in flight: %6.1f B/clk/SM, %7.0f clk per box, latency ~%5.0f clk\n", box_rows, box_rows / 8, stages,
                   bytes / cyc, cyc / iters, cyc / iters * stages);
        }
    return 0;
}
