// Which cheaper formulations of the reference's Moflex colour arithmetic (MD:300-305) give the SAME BYTE for every possible input?
// The inputs are discrete -- Y in 0..255, U and V multiples of 1/4 in [-128, 127] (one, two or four chroma samples averaged) -- so
// the question is decided by enumeration: 256 x 1021 pairs for R and B, 256 x 1021 x 1021 triples for G.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o bgra_formula bgra_formula.cu && ./bgra_formula
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <cmath>

__device__ __forceinline__ int to_byte(float v) { v = fminf(fmaxf(v, 0.0f), 255.0f); return (int)v; }   // clamp, truncate (MD:313-320)
__device__ __forceinline__ float ref_tail(float c) { return __fdiv_rn(__fmul_rn(__fsub_rn(c, 16.0f), 255.0f), 239.0f); }

struct Consts { float r, r_up, k, k_up, k_dn, m16k, m16k_up, m16k_dn; };

// variant: 0 exact division (sanity), 1 (c-16)*255*r, 2 same with r_up, 3 (c-16)*k, 4 k_up, 5 k_dn, 6 fma(c, k, -16k), 7 fma(c, k_up, -16 k_up), 8 fma(c, k_dn, -16 k_dn)
__device__ __forceinline__ float tail(int variant, float c, const Consts& K) {
    const float d = __fsub_rn(c, 16.0f);
    switch (variant) {
    case 0: { const float x = __fmul_rn(d, 255.0f), q = __fmul_rn(x, K.r); return __fmaf_rn(__fmaf_rn(-239.0f, q, x), K.r, q); }
    case 1: return __fmul_rn(__fmul_rn(d, 255.0f), K.r);
    case 2: return __fmul_rn(__fmul_rn(d, 255.0f), K.r_up);
    case 3: return __fmul_rn(d, K.k);
    case 4: return __fmul_rn(d, K.k_up);
    case 5: return __fmul_rn(d, K.k_dn);
    case 6: return __fmaf_rn(c, K.k, K.m16k);
    case 7: return __fmaf_rn(c, K.k_up, K.m16k_up);
    default: return __fmaf_rn(c, K.k_dn, K.m16k_dn);
    }
}

constexpr int NV = 9, NF = 6;   // NF: variants with the colour matrix contracted too
// fused variants: 0 fma matrix + exact division, 1 fma matrix + (c-16)*255*r, 2 fma matrix + fma(c,k_up,-16k_up),
// 3 two FFMAs with premultiplied constants (k_up), 4 the same with k, 5 separate matrix ops + fma(c,k_up,-16k_up) on R/B and (c-16)*255*r on G (the candidate)
__device__ __forceinline__ int fusedR(int f, float Y, float V, const Consts& K) {
    const float c = __fmaf_rn(1.420f, V, Y), cs = __fadd_rn(Y, __fmul_rn(1.420f, V));
    switch (f) {
    case 0: return to_byte(tail(0, c, K));
    case 1: return to_byte(tail(1, c, K));
    case 2: return to_byte(tail(7, c, K));
    case 3: return to_byte(__fmaf_rn(V, 1.420f * K.k_up, __fmaf_rn(Y, K.k_up, K.m16k_up)));
    case 4: return to_byte(__fmaf_rn(V, 1.420f * K.k, __fmaf_rn(Y, K.k, K.m16k)));
    default: return to_byte(tail(7, cs, K));
    }
}
__device__ __forceinline__ int fusedB(int f, float Y, float U, const Consts& K) {
    const float c = __fmaf_rn(1.772f, U, Y), cs = __fadd_rn(Y, __fmul_rn(1.772f, U));
    switch (f) {
    case 0: return to_byte(tail(0, c, K));
    case 1: return to_byte(tail(1, c, K));
    case 2: return to_byte(tail(7, c, K));
    case 3: return to_byte(__fmaf_rn(U, 1.772f * K.k_up, __fmaf_rn(Y, K.k_up, K.m16k_up)));
    case 4: return to_byte(__fmaf_rn(U, 1.772f * K.k, __fmaf_rn(Y, K.k, K.m16k)));
    default: return to_byte(tail(7, cs, K));
    }
}
__device__ __forceinline__ int fusedG(int f, float Y, float U, float V, const Consts& K) {
    const float c = __fmaf_rn(-0.714f, V, __fmaf_rn(-0.344f, U, Y)), cs = __fsub_rn(__fsub_rn(Y, __fmul_rn(0.344f, U)), __fmul_rn(0.714f, V));
    switch (f) {
    case 0: return to_byte(tail(0, c, K));
    case 1: return to_byte(tail(1, c, K));
    case 2: return to_byte(tail(7, c, K));
    case 3: return to_byte(__fmaf_rn(V, -0.714f * K.k_up, __fmaf_rn(U, -0.344f * K.k_up, __fmaf_rn(Y, K.k_up, K.m16k_up))));
    case 4: return to_byte(__fmaf_rn(V, -0.714f * K.k, __fmaf_rn(U, -0.344f * K.k, __fmaf_rn(Y, K.k, K.m16k))));
    default: return to_byte(tail(1, cs, K));
    }
}
__global__ void k_search(Consts K, unsigned long long* bad /* [3][NV] */, unsigned long long* badf /* [3][NF] */) {
    // grid: x over (Y, U) pairs, loop over V inside
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 256 * 1021) return;
    const int yi = idx / 1021, ui = idx % 1021;
    const float Y = (float)yi, U = (float)(ui - 512) * 0.25f;
    unsigned long long loc[3][NV] = {}, locf[3][NF] = {};
    {   // B depends on (Y, U); R on (Y, V): enumerate it with U standing in for V
        const float cB = __fadd_rn(Y, __fmul_rn(1.772f, U)), cR = __fadd_rn(Y, __fmul_rn(1.420f, U));
        const int wB = to_byte(ref_tail(cB)), wR = to_byte(ref_tail(cR));
        for (int v = 0; v < NV; v++) { loc[2][v] += to_byte(tail(v, cB, K)) != wB; loc[0][v] += to_byte(tail(v, cR, K)) != wR; }
        for (int f = 0; f < NF; f++) { locf[2][f] += fusedB(f, Y, U, K) != wB; locf[0][f] += fusedR(f, Y, U, K) != wR; }
    }
    const float gu = __fsub_rn(Y, __fmul_rn(0.344f, U));
    for (int vi = 0; vi < 1021; vi++) {
        const float V = (float)(vi - 512) * 0.25f;
        const float cG = __fsub_rn(gu, __fmul_rn(0.714f, V));
        const int wG = to_byte(ref_tail(cG));
        for (int v = 0; v < NV; v++) loc[1][v] += to_byte(tail(v, cG, K)) != wG;
        for (int f = 0; f < NF; f++) locf[1][f] += fusedG(f, Y, U, V, K) != wG;
    }
    for (int c = 0; c < 3; c++) for (int v = 0; v < NV; v++) if (loc[c][v]) atomicAdd(&bad[c * NV + v], loc[c][v]);
    for (int c = 0; c < 3; c++) for (int f = 0; f < NF; f++) if (locf[c][f]) atomicAdd(&badf[c * NF + f], locf[c][f]);
}

int main() {
    Consts K;
    K.r = 1.0f / 239.0f; K.r_up = nextafterf(K.r, 1.0f);
    K.k = 255.0f / 239.0f; K.k_up = nextafterf(K.k, 2.0f); K.k_dn = nextafterf(K.k, 0.0f);
    K.m16k = -16.0f * K.k; K.m16k_up = -16.0f * K.k_up; K.m16k_dn = -16.0f * K.k_dn;
    unsigned long long* bad; cudaMalloc(&bad, sizeof(unsigned long long) * 3 * NV); cudaMemset(bad, 0, sizeof(unsigned long long) * 3 * NV);
    unsigned long long* badf; cudaMalloc(&badf, sizeof(unsigned long long) * 3 * NF); cudaMemset(badf, 0, sizeof(unsigned long long) * 3 * NF);
    k_search<<<(256 * 1021 + 255) / 256, 256>>>(K, bad, badf);
    unsigned long long h[3 * NV];
    if (cudaMemcpy(h, bad, sizeof h, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    const char* names[NV] = {"exact 3-instruction division (sanity)", "(c-16)*255*r", "(c-16)*255*r_up", "(c-16)*k", "(c-16)*k_up", "(c-16)*k_dn", "fma(c,k,-16k)", "fma(c,k_up,-16k_up)", "fma(c,k_dn,-16k_dn)"};
    printf("bytes that differ from the reference arithmetic over the whole input domain (R: 261k pairs, G: 267M triples, B: 261k pairs)\n");
    for (int v = 0; v < NV; v++) printf("  %-40s R %8llu  G %10llu  B %8llu\n", names[v], h[v], h[NV + v], h[2 * NV + v]);
    unsigned long long hf[3 * NF]; cudaMemcpy(hf, badf, sizeof hf, cudaMemcpyDeviceToHost);
    const char* fn[NF] = {"fma matrix + exact division", "fma matrix + (c-16)*255*r", "fma matrix + fma(c,k_up,-16k_up)", "FFMA chain, constants premultiplied by k_up", "FFMA chain, premultiplied by k", "CANDIDATE: R,B fma(c,k_up,-16k_up); G (c-16)*255*r"};
    for (int f = 0; f < NF; f++) printf("  %-52s R %8llu  G %10llu  B %8llu\n", fn[f], hf[f], hf[NF + f], hf[2 * NF + f]);
    return 0;
}
