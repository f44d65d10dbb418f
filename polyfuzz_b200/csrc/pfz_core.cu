// pfz_core.cu -- error plumbing, device info and the multi-level exclusive scan used by K1/K2.
#include <stdarg.h>
#include "pfz_common.cuh"

namespace pfz {
static thread_local char g_err[1024] = "";
unsigned long long g_launches = 0;
void set_error(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- scan -------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;                      // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// block-level exclusive scan of one tile; writes tile total to sums[blockIdx.x] when sums != NULL
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_kernel(const int32_t *__restrict__ in, int32_t *__restrict__ out,
                                                               int64_t n, int32_t *__restrict__ sums) {
    __shared__ int32_t warp_tot[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS];
    int32_t tsum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        tsum += v[i];
    }
    int32_t incl = warp_incl_scan(tsum);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    if (w == 0) {
        int32_t t = (lane < SCAN_THREADS / 32) ? warp_tot[lane] : 0;
        int32_t ti = warp_incl_scan(t);
        if (lane < SCAN_THREADS / 32) warp_tot[lane] = ti - t;
        if (lane == 31 && sums) sums[blockIdx.x] = ti;
    }
    __syncthreads();
    int32_t run = warp_tot[w] + incl - tsum;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(int32_t *__restrict__ out, int64_t n,
                                                              const int32_t *__restrict__ offs) {
    const int32_t add = offs[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) out[base + i] += add;
}

static int64_t n_tiles_of(int64_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

int scan_exclusive_i32(const int32_t *in, int32_t *out, int64_t n, void *ws, cudaStream_t st) {
    if (n <= 0) return 0;
    const int64_t nt = n_tiles_of(n);
    if (nt == 1) {
        scan_tile_kernel<<<1, SCAN_THREADS, 0, st>>>(in, out, n, nullptr);
        PFZ_LAUNCH_OK();
        return 0;
    }
    int32_t *sums = reinterpret_cast<int32_t *>(ws);
    // level-1 sums followed by deeper workspace (aligned to 256 B)
    char *next_ws = reinterpret_cast<char *>(ws) + ((nt * 4 + 255) / 256) * 256;
    scan_tile_kernel<<<(unsigned)nt, SCAN_THREADS, 0, st>>>(in, out, n, sums);
    PFZ_LAUNCH_OK();
    if (scan_exclusive_i32(sums, sums, nt, next_ws, st)) return 1;
    scan_add_kernel<<<(unsigned)nt, SCAN_THREADS, 0, st>>>(out, n, sums);
    PFZ_LAUNCH_OK();
    return 0;
}

// ---- integer-ALU throughput probe (the roofline denominator of K3, which is integer-issue bound) ------------
// 8 independent chains per thread, each alternating LOP3 (xor3) and IADD -- the op mix of the bit-parallel recurrences.
__global__ void __launch_bounds__(256) int_alu_probe_kernel(uint32_t *__restrict__ out, int iters, uint32_t seed) {
    uint32_t a[8];
    const uint32_t b = seed ^ (threadIdx.x * 2654435761u), c = seed * 3u + blockIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = b + i * 977u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) x ^= a[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x;
}
}  // namespace pfz

extern "C" {

int pfz_int_alu_probe(int32_t iters, uint32_t *scratch, int64_t *lane_ops_host, void *stream) {
    int dev = 0, sms = 0;
    PFZ_CUDA_OK(cudaGetDevice(&dev));
    PFZ_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8;
    pfz::int_alu_probe_kernel<<<blocks, 256, 0, pfz::as_stream(stream)>>>(scratch, iters, 12345u);
    PFZ_LAUNCH_OK();
    if (lane_ops_host) *lane_ops_host = (int64_t)blocks * 256 * (int64_t)iters * 64;
    return 0;
}

int pfz_abi_version(void) { return PFZ_ABI_VERSION; }
const char *pfz_last_error(void) { return pfz::g_err; }
int64_t pfz_launch_count(void) { return (int64_t)__atomic_load_n(&pfz::g_launches, __ATOMIC_RELAXED); }

int64_t pfz_scan_ws_bytes(int64_t n) {
    int64_t total = 256;
    while (n > pfz::SCAN_TILE) {
        n = pfz::n_tiles_of(n);
        total += ((n * 4 + 255) / 256) * 256;
    }
    return total;
}

int pfz_device_info(int32_t *sm_count, int32_t *smem_optin, int32_t *cc_major, int32_t *cc_minor) {
    int dev = 0;
    PFZ_CUDA_OK(cudaGetDevice(&dev));
    int v = 0;
    PFZ_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm_count = v;
    PFZ_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)); *smem_optin = v;
    PFZ_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v;
    PFZ_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v;
    return 0;
}
}
