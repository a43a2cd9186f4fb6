// Stand-alone probe of the TMA forms k_inter uses: rank-3 u8 tensor (S, R, P), boxes 32x17x1 and 16x9x1, arbitrary
// (also negative / out-of-range) start coordinates.  Prints mismatches against the expected zero-filled window.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tm, int bw, int bh, const int* coords, int n, uint8_t* out) {
    __shared__ __align__(128) uint8_t buf[1024];
    __shared__ uint64_t bar;
    const uint32_t b = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    for (int k = 0; k < n; k++) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bw * bh) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         :: "r"(smem_u32(buf)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(coords[3 * k]), "r"(coords[3 * k + 1]), "r"(coords[3 * k + 2]), "r"(b) : "memory");
        }
        uint32_t done;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(b), "r"(k & 1) : "memory");
        } while (!done);
        for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[k * 1024 + i] = buf[i];
        __syncthreads();
    }
}

int main() {
    const int S = 512, R = 360, P = 6;
    const size_t pic = (size_t)S * R + 256;
    std::vector<uint8_t> h(pic * P);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 2654435761u >> 13);
    uint8_t* d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    const cuuint64_t dims[3] = {S, R, P}; const cuuint64_t str[2] = {S, pic}; const cuuint32_t es[3] = {1, 1, 1};
    const int boxes[2][2] = {{32, 17}, {16, 9}};
    int bad_total = 0;
    for (int bk = 0; bk < 2; bk++) {
        CUtensorMap tm; const cuuint32_t box[3] = {(cuuint32_t)boxes[bk][0], (cuuint32_t)boxes[bk][1], 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box %dx%d encode -> %d\n", boxes[bk][0], boxes[bk][1], (int)r);
        std::vector<int> c = {0, 0, 0,  5, 7, 1,  -3, 10, 2,  500, 20, 3,  100, -2, 4,  100, 350, 5,  17, 240, 0,  273, 245, 5, 1, 1, 5, 255, 300, 2, 3, 343, 1};
        const int n = (int)c.size() / 3;
        int* dc; cudaMalloc(&dc, c.size() * 4); cudaMemcpy(dc, c.data(), c.size() * 4, cudaMemcpyHostToDevice);
        uint8_t* dout; cudaMalloc(&dout, n * 1024); cudaMemset(dout, 0xEE, n * 1024);
        probe<<<1, 64>>>(tm, boxes[bk][0], boxes[bk][1], dc, n, dout);
        cudaError_t e = cudaDeviceSynchronize();
        printf("  kernel -> %s\n", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<uint8_t> o(n * 1024); cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost);
        for (int k = 0; k < n; k++) {
            int bad = 0;
            for (int y = 0; y < boxes[bk][1]; y++) for (int x = 0; x < boxes[bk][0]; x++) {
                const int gx = c[3 * k] + x, gy = c[3 * k + 1] + y, gp = c[3 * k + 2];
                const uint8_t want = (gx < 0 || gx >= S || gy < 0 || gy >= R) ? 0 : h[gp * pic + (size_t)gy * S + gx];
                if (o[k * 1024 + y * boxes[bk][0] + x] != want) bad++;
            }
            printf("  coords (%d,%d,%d): %d mismatches\n", c[3 * k], c[3 * k + 1], c[3 * k + 2], bad);
            bad_total += bad;
        }
    }
    printf("TOTAL mismatches %d\n", bad_total);
    return bad_total != 0;
}
