// Canonical programming-guide TMA sample (2-D, int tensor, cuda::barrier) plus variants, one per process run:
//   ./tma_probe2 0   canonical 2D int32 64x64 box
//   ./tma_probe2 1   2D u8 box 32x17
//   ./tma_probe2 2   3D u8 box 32x17x1
//   ./tma_probe2 3   3D u8 box 32x16x1
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

template <int RANK, int BYTES>
__global__ void k(const __grid_constant__ CUtensorMap tm, int x, int y, int z, uint8_t* out) {
    __shared__ alignas(128) uint8_t buf[BYTES];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        if constexpr (RANK == 2) cde::cp_async_bulk_tensor_2d_global_to_shared(buf, &tm, x, y, bar);
        else cde::cp_async_bulk_tensor_3d_global_to_shared(buf, &tm, x, y, z, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, BYTES);
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < BYTES; i += blockDim.x) out[i] = buf[i];
}

int main(int argc, char** argv) {
    const int v = argc > 1 ? atoi(argv[1]) : 0;
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    printf("variant %d entry %p q %d\n", v, fn, (int)q);
    uint8_t* d; cudaMalloc(&d, 64 << 20); cudaMemset(d, 7, 64 << 20);
    uint8_t* out; cudaMalloc(&out, 1 << 16);
    CUtensorMap tm; CUresult r; cudaError_t e;
    if (v == 0) {
        const cuuint64_t dims[2] = {1024, 1024}; const cuuint64_t str[1] = {1024 * 4}; const cuuint32_t box[2] = {64, 64}; const cuuint32_t es[2] = {1, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode %d\n", (int)r);
        k<2, 64 * 64 * 4><<<1, 128>>>(tm, 0, 0, 0, out);
    } else if (v == 1) {
        const cuuint64_t dims[2] = {512, 360}; const cuuint64_t str[1] = {512}; const cuuint32_t box[2] = {32, 17}; const cuuint32_t es[2] = {1, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode %d\n", (int)r);
        k<2, 32 * 17><<<1, 128>>>(tm, argc > 2 ? atoi(argv[2]) : 3, argc > 3 ? atoi(argv[3]) : 5, 0, out);
    } else {
        const int bh = v == 2 ? 17 : 16;
        const cuuint64_t dims[3] = {512, 360, 6}; const cuuint64_t str[2] = {512, 512 * 360 + 256}; const cuuint32_t box[3] = {32, (cuuint32_t)bh, 1}; const cuuint32_t es[3] = {1, 1, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode %d\n", (int)r);
        if (v == 2) k<3, 32 * 17><<<1, 128>>>(tm, 3, 5, 1, out); else k<3, 32 * 16><<<1, 128>>>(tm, 3, 5, 1, out);
    }
    e = cudaDeviceSynchronize();
    printf("kernel -> %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
