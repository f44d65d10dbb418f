// Developer probe: shared-memory integer atomic (RED.ADD.U32) throughput vs a plain LDS+STS read-modify-write, per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atoms_probe atoms_probe.cu && ./atoms_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) probe(uint32_t *out, const uint32_t *idx, int iters, int warps_active) {
    __shared__ uint32_t acc[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) acc[i] = 0;
    __syncthreads();
    const int w = threadIdx.x >> 5;
    if (w >= warps_active) return;
    uint32_t a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = (unsigned)__cvta_generic_to_shared(&acc[idx[(threadIdx.x + 256 * k) & 2047] & 8191]);
    uint32_t v = threadIdx.x + 1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(a[k]), "r"(v) : "memory");
            else if (MODE == 1) { uint32_t o; asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(a[k]), "r"(v) : "memory"); v += o & 1; }
            else { uint32_t o; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(o) : "r"(a[k]) : "memory"); o += v; asm volatile("st.shared.u32 [%0], %1;" :: "r"(a[k]), "r"(o) : "memory"); }
        }
    }
    __syncwarp();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[threadIdx.x] + v;
}
int main() {
    uint32_t *d_out, *d_idx; uint32_t h[2048];
    cudaMalloc(&d_out, 148 * 8 * 256 * 4); cudaMalloc(&d_idx, 2048 * 4);
    const char *pat[3] = {"distinct banks (lane-linear)", "random rows (bank conflicts ~3.5x)", "same bank (32-way)"};
    const char *mode[3] = {"red.shared.add.u32", "atom.shared.add.u32 (+return)", "ld.shared + add + st.shared"};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int p = 0; p < 3; ++p) {
        uint32_t s = 12345;
        for (int i = 0; i < 2048; ++i) { s = s * 1664525u + 1013904223u; h[i] = p == 0 ? (uint32_t)(i * 1 + (i / 32) * 32 * 3) : p == 1 ? (s >> 8) : (uint32_t)(i * 32); }
        cudaMemcpy(d_idx, h, sizeof(h), cudaMemcpyHostToDevice);
        for (int m = 0; m < 3; ++m) for (int wa : {1, 2, 4, 8}) {
            const int iters = 20000;
            float ms = 0;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (m == 0) probe<0><<<148, 256>>>(d_out, d_idx, iters, wa);
                if (m == 1) probe<1><<<148, 256>>>(d_out, d_idx, iters, wa);
                if (m == 2) probe<2><<<148, 256>>>(d_out, d_idx, iters, wa);
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            }
            const double ops = (double)iters * 8 * wa;                 // warp-level updates per SM
            printf("%-34s %-38s warps/SM %d: %.2f SM-cycles per warp update (at max clock)  %.3f ms\n", mode[m], pat[p], wa,
                   ms * 1e-3 * clk * 1e3 / ops, ms);
        }
    }
    return 0;
}
