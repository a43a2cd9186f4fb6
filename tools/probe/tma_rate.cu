// How many small TMA boxes per second does an SM take?  Each warp keeps DEPTH boxes (32 x ROWS bytes, u8) in flight from a
// `span`-byte region of a 2-D tensor with 512-byte rows (the ring's geometry at 400x240); lane 0 issues, all lanes wait.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate tma_rate.cu && ./tma_rate
// Prints, per (warps per SM, depth, region size): boxes per microsecond per SM and the implied cycles per box.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ROWS, int DEPTH>
__global__ void __launch_bounds__(128) k_rate(const __grid_constant__ CUtensorMap tm, int n_boxes, uint32_t rows_total, uint32_t seed, unsigned long long* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int SLOT = (32 * ROWS + 127) / 128 * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* box = smem + (size_t)warp * (DEPTH * SLOT + 128);
    uint64_t* bar = reinterpret_cast<uint64_t*>(box + DEPTH * SLOT);
    if (lane < DEPTH) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[lane])) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    uint32_t rng = seed * 2654435761u + (blockIdx.x * 4 + warp) * 40503u + 12345u;
    auto issue = [&](int slot) {
        rng = rng * 1664525u + 1013904223u;
        const int x = (int)((rng >> 8) % 30u) * 16, y = (int)((rng >> 13) % (rows_total - ROWS));
        if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar[slot])), "r"(32 * ROWS) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         :: "r"(smem_u32(box + slot * SLOT)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(x), "r"(y), "r"(smem_u32(&bar[slot])) : "memory");
        }
    };
    for (int s = 0; s < DEPTH; s++) issue(s);
    uint32_t phase = 0, acc = 0;
    for (int i = 0; i < n_boxes; i++) {
        const int slot = i % DEPTH;
        if (slot == 0 && i) phase ^= 1u;
        uint32_t done;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar[slot])), "r"(phase) : "memory");
        } while (!done);
        acc += box[slot * SLOT + lane];
        __syncwarp();
        if (i + DEPTH < n_boxes) issue(slot);
    }
    if (acc == 0xFFFFFFFFu) *sink = acc;
}

template <int ROWS, int DEPTH>
static void run(PFN_cuTensorMapEncodeTiled_v12000 enc, uint8_t* d, size_t bytes, int ctas_per_sm, int n_boxes, unsigned long long* sink) {
    const cuuint64_t rows = bytes / 512;
    const cuuint64_t dims[2] = {512, rows}; const cuuint64_t str[1] = {512}; const cuuint32_t bx[2] = {32, ROWS}; const cuuint32_t es[2] = {1, 1};
    CUtensorMap tm;
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, str, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return; }
    constexpr int SLOT = (32 * ROWS + 127) / 128 * 128;
    const size_t smem = 4 * (DEPTH * SLOT + 128);
    cudaFuncSetAttribute(k_rate<ROWS, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm;
    k_rate<ROWS, DEPTH><<<grid, 128, smem>>>(tm, n_boxes / 4, (uint32_t)rows, 1u, sink);
    cudaEventRecord(e0);
    k_rate<ROWS, DEPTH><<<grid, 128, smem>>>(tm, n_boxes, (uint32_t)rows, 2u, sink);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return; }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double boxes = (double)grid * 4 * n_boxes, per_sm_us = boxes / 148 / (ms * 1e3);
    printf("rows %2d depth %d warps/SM %2d region %6.0f MB: %.3f ms, %6.1f boxes/us/SM = %5.1f cycles/box at 1.965 GHz, %.2f TB/s of box bytes\n",
           ROWS, DEPTH, ctas_per_sm * 4, bytes / 1e6, ms, per_sm_us, 1965.0 / per_sm_us, boxes * 32 * ROWS / (ms * 1e-3) / 1e12);
}

int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    const size_t big = (size_t)1200 << 20;
    uint8_t* d; cudaMalloc(&d, big); cudaMemset(d, 7, big);
    unsigned long long* sink; cudaMalloc(&sink, 8);
    for (size_t bytes : {(size_t)32 << 20, big}) {          // L2-resident vs mostly DRAM
        run<17, 2>(enc, d, bytes, 4, 2000, sink);
        run<17, 2>(enc, d, bytes, 9, 2000, sink);
        run<17, 4>(enc, d, bytes, 9, 2000, sink);
        run<9, 4>(enc, d, bytes, 9, 2000, sink);
        run<17, 8>(enc, d, bytes, 6, 2000, sink);
    }
    return 0;
}
