// pfz_dense.cu -- K4: dense cosine top-k for pre-computed embeddings: C = X * Y^T on the 5th-generation
// tensor cores (tcgen05.mma, bf16 in, fp32 accumulate in TMEM) fed by TMA, with the per-row top-k fused
// into the epilogue so the n_from x n_to score matrix never exists in memory.
//
// Replaces the dense branch of polyfuzz/models/_utils.py:94-102 (sklearn cosine_similarity + argsort) as
// reached from polyfuzz/models/_embeddings.py:127-131 when the caller supplies embeddings.
//
// CTA = 128 from-rows (one TMEM lane per row).  Per 256-wide to-tile: K loop of 64-element (128-byte,
// SWIZZLE_128B) TMA boxes through a 4-stage shared-memory ring, tcgen05.mma M=128,N=256,K=16 issued by one
// thread, accumulator double-buffered in TMEM (2 x 256 columns) so the epilogue of tile t overlaps the MMAs of
// tile t+1.  Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = epilogue (thread =
// row, sorted top-k in registers, key (score desc, index asc)).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include "pfz_common.cuh"

namespace pfz {

constexpr int DM = 128, DN = 256, DK = 64, DSTAGES = 4, UMMA_K = 16;
constexpr int A_BYTES = DM * DK * 2, B_BYTES = DN * DK * 2;          // 16 KB, 32 KB
constexpr int DENSE_THREADS = 256;

__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile("{\n\t.reg .pred p;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}"
                 :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *map, int c0, int c1, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void tc_mma(unsigned d_tmem, uint64_t adesc, uint64_t bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(unsigned smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);               // start address  (bits 0-13)
    d |= (uint64_t)1 << 16;                                    // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                          // stride byte offset (bits 32-45)
    d |= (uint64_t)1 << 46;                                    // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                                    // layout type: SWIZZLE_128B
    return d;
}
// instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major, M=128, N=256
__device__ __forceinline__ unsigned umma_idesc() {
    unsigned d = 0;
    d |= 1u << 4;                       // c_format = F32
    d |= 1u << 7;                       // a_format = BF16
    d |= 1u << 10;                      // b_format = BF16
    d |= (unsigned)(DN >> 3) << 17;     // n_dim
    d |= (unsigned)(DM >> 4) << 24;     // m_dim
    return d;
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct DenseParams {
    int n_from, n_to, d; int k; float min_sim; int self_match; long long from_base, to_base;
    int n_splits; int n_mblocks; int n_ntiles;
    int32_t *top_idx; double *top_val;          // [n_splits][n_from][k]
};

template <int KMAX>
__global__ void __launch_bounds__(DENSE_THREADS, 1) dense_cos_topk_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                          const __grid_constant__ CUtensorMap map_y, const DenseParams P) {
    extern __shared__ __align__(1024) unsigned char dsm_raw[];
    // 128-byte-swizzled TMA/UMMA tiles need 1024-byte alignment: align by hand (the launch adds 1 KB of slack)
    unsigned char *dsm = dsm_raw + ((1024u - (s32(dsm_raw) & 1023u)) & 1023u);
    // [A stages][B stages] then barriers
    unsigned char *sa = dsm;
    unsigned char *sb = dsm + DSTAGES * A_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(dsm + DSTAGES * (A_BYTES + B_BYTES));
    // bars: full[0..S), empty[S..2S), tmem_full[2S..2S+2), tmem_empty[2S+2..2S+4)
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(bars + 2 * DSTAGES + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned bar0 = s32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (DSTAGES + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * DSTAGES + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * DSTAGES + 2 + a); };

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < DSTAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(s32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;

    const int n_units = P.n_mblocks * P.n_splits;
    const int tiles_per = (P.n_ntiles + P.n_splits - 1) / P.n_splits;
    const int n_kblk = (P.d + DK - 1) / DK;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; unsigned phase = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int mb = u % P.n_mblocks, sp = u / P.n_mblocks;
                const int t_lo = sp * tiles_per, t_hi = min(P.n_ntiles, t_lo + tiles_per);
                for (int t = t_lo; t < t_hi; ++t) {
                    for (int kb = 0; kb < n_kblk; ++kb) {
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        mbar_expect_tx(full_bar(stage), A_BYTES + B_BYTES);
                        tma_load_2d(s32(sa + stage * A_BYTES), &map_x, kb * DK, mb * DM, full_bar(stage));
                        tma_load_2d(s32(sb + stage * B_BYTES), &map_y, kb * DK, t * DN, full_bar(stage));
                        if (++stage == DSTAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const unsigned idesc = umma_idesc();
            int stage = 0; unsigned phase = 0; int acc = 0; unsigned acc_phase = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int sp = u / P.n_mblocks;
                const int t_lo = sp * tiles_per, t_hi = min(P.n_ntiles, t_lo + tiles_per);
                for (int t = t_lo; t < t_hi; ++t) {
                    mbar_wait(tempty_bar(acc), acc_phase ^ 1);            // epilogue drained this accumulator
                    tc_fence_after();
                    const unsigned d_tmem = tmem_base + (unsigned)(acc * DN);
                    for (int kb = 0; kb < n_kblk; ++kb) {
                        mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        const uint64_t adesc = umma_desc(s32(sa + stage * A_BYTES));
                        const uint64_t bdesc = umma_desc(s32(sb + stage * B_BYTES));
#pragma unroll
                        for (int kk = 0; kk < DK / UMMA_K; ++kk) {
                            // advance 16 elements = 32 bytes along K inside the 128-byte swizzled row: +2 in the address field
                            tc_mma(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (kb | kk) ? 1u : 0u);
                        }
                        tc_commit(empty_bar(stage));                      // smem slot free once these MMAs retire
                        if (++stage == DSTAGES) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(tfull_bar(acc));                            // accumulator complete
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;                                          // == warp % 4: TMEM lane quarter of this warp
        const int row_in_blk = ew * 32 + lane;
        int acc = 0; unsigned acc_phase = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int mb = u % P.n_mblocks, sp = u / P.n_mblocks;
            const int t_lo = sp * tiles_per, t_hi = min(P.n_ntiles, t_lo + tiles_per);
            const int row = mb * DM + row_in_blk;
            const long long self_col = P.from_base + row - P.to_base;    // local to-column of the diagonal
            float tv[KMAX]; int ti[KMAX];
#pragma unroll
            for (int q = 0; q < KMAX; ++q) { tv[q] = P.min_sim; ti[q] = -1; }
            float kv = P.min_sim; int ki = -1;
            for (int t = t_lo; t < t_hi; ++t) {
                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                const unsigned taddr = tmem_base + ((unsigned)(ew * 32) << 16) + (unsigned)(acc * DN);
                for (int c0 = 0; c0 < DN; c0 += 32) {
                    unsigned r[32];
                    tmem_ld32(taddr + (unsigned)c0, r);
                    // pass 1 (branch-free, 2-3 instructions per value): which of the 32 scores rank before the k-th key?
                    // (score desc, index asc); the sentinel (min_sim, -1) makes the test strict while the list is not full
                    const int colb = t * DN + c0;
                    unsigned mask = 0u;
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const float sc = __uint_as_float(r[q]);
                        if (sc > kv || (sc == kv && colb + q < ki)) mask |= 1u << q;
                    }
                    // pass 2 (rare after the first tiles; the insertion code exists once, not 32 times -- the unrolled
                    // version overflowed the instruction cache and ran ~30x slower)
                    while (mask) {
                        const int q = __ffs(mask) - 1; mask &= mask - 1;
                        float sc = 0.f;
#pragma unroll
                        for (int z = 0; z < 32; ++z) if (z == q) sc = __uint_as_float(r[z]);
                        const int col = colb + q;
                        if (!(sc > kv || (sc == kv && col < ki))) continue;          // the key may have risen meanwhile
                        if (col >= P.n_to || (P.self_match && (long long)col == self_col)) continue;
                        float cv = sc; int ci = (int)(P.to_base + col);
#pragma unroll
                        for (int z = 0; z < KMAX; ++z) {
                            if (z < P.k) {
                                const bool before = cv > tv[z] || (cv == tv[z] && (ti[z] < 0 || ci < ti[z]));
                                if (before) { const float fv = tv[z]; const int fi = ti[z]; tv[z] = cv; ti[z] = ci; cv = fv; ci = fi; }
                                if (z == P.k - 1) { kv = tv[z]; ki = ti[z] < 0 ? -1 : ti[z] - (int)P.to_base; }
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(tempty_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (row < P.n_from) {
                const size_t o = ((size_t)sp * P.n_from + row) * P.k;
#pragma unroll
                for (int z = 0; z < KMAX; ++z)
                    if (z < P.k) { P.top_idx[o + z] = ti[z]; P.top_val[o + z] = ti[z] >= 0 ? (double)tv[z] : 0.0; }
            }
        }
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
    }
}

// ---- 2-CTA variant (cta_group::2) -----------------------------------------------------------------------------------------
// A CTA pair (cluster of 2, same TPC) computes M = 256 from-rows x N = 256 to-rows per MMA: each CTA stages ITS 128 rows of X
// and ITS half (128 rows) of the Y tile, the leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 reading both CTAs'
// shared memory, and each CTA's TMEM receives the accumulators of its own 128 rows.  Per CTA and K step the L2 -> SM traffic is
// 16 + 16 KB instead of 16 + 32 KB (the 1-CTA kernel streams 96 B/cycle/SM at full tensor rate and is L2->SM bound), which
// also frees shared memory for 6 stages.  Barriers: `full` lives in the leader only (both CTAs' TMA complete_tx land there,
// peer bit of the address cleared); `empty` / `tmem_full` exist per CTA and are signalled by multicast commits; `tmem_empty`
// lives in the leader and collects the 2 x 128 epilogue threads (remote arrive from the peer).
constexpr int D2STAGES = 6;
constexpr int BH_BYTES = (DN / 2) * DK * 2;                           // this CTA's half of the Y tile: 16 KB
constexpr unsigned PEER_MASK = 0xFEFFFFFFu;                          // clears the CTA-rank bit of a shared::cluster address

__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(unsigned dst, const CUtensorMap *map, int c0, int c1, unsigned leader_bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(leader_bar) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(unsigned bar) {
    asm volatile("{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
                 "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_2sm(unsigned d_tmem, uint64_t adesc, uint64_t bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(unsigned local_bar) {      // arrive on the barrier at the same offset in CTA 0
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(local_bar & PEER_MASK) : "memory");
}
__device__ __forceinline__ unsigned umma_idesc_2sm() {                        // M = 256 (pair), N = 256
    unsigned d = 0;
    d |= 1u << 4; d |= 1u << 7; d |= 1u << 10;
    d |= (unsigned)(DN >> 3) << 17;
    d |= (unsigned)((2 * DM) >> 4) << 24;
    return d;
}

template <int KMAX>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DENSE_THREADS, 1)
dense_cos_topk2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y, const DenseParams P) {
    extern __shared__ __align__(1024) unsigned char dsm_raw[];
    unsigned char *dsm = dsm_raw + ((1024u - (s32(dsm_raw) & 1023u)) & 1023u);
    unsigned char *sa = dsm;
    unsigned char *sb = dsm + D2STAGES * A_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(dsm + D2STAGES * (A_BYTES + BH_BYTES));
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(bars + 2 * D2STAGES + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned rank = cluster_ctarank();
    const unsigned bar0 = s32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (D2STAGES + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * D2STAGES + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * D2STAGES + 2 + a); };

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < D2STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(s32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                                  // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;

    const int n_mpairs = (P.n_from + 2 * DM - 1) / (2 * DM);
    const int n_units = n_mpairs * P.n_splits;
    const int tiles_per = (P.n_ntiles + P.n_splits - 1) / P.n_splits;
    const int n_kblk = (P.d + DK - 1) / DK;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; unsigned phase = 0;
            for (int u = pair; u < n_units; u += n_pairs) {
                const int mp = u % n_mpairs, sp = u / n_mpairs;
                const int t_lo = sp * tiles_per, t_hi = min(P.n_ntiles, t_lo + tiles_per);
                for (int t = t_lo; t < t_hi; ++t) {
                    for (int kb = 0; kb < n_kblk; ++kb) {
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * (A_BYTES + BH_BYTES));
                        tma_load_2d_2sm(s32(sa + stage * A_BYTES), &map_x, kb * DK, mp * 2 * DM + (int)rank * DM, full_bar(stage) & PEER_MASK);
                        tma_load_2d_2sm(s32(sb + stage * BH_BYTES), &map_y, kb * DK, t * DN + (int)rank * (DN / 2), full_bar(stage) & PEER_MASK);
                        if (++stage == D2STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            const unsigned idesc = umma_idesc_2sm();
            int stage = 0; unsigned phase = 0; int acc = 0; unsigned acc_phase = 0;
            for (int u = pair; u < n_units; u += n_pairs) {
                const int sp = u / n_mpairs;
                const int t_lo = sp * tiles_per, t_hi = min(P.n_ntiles, t_lo + tiles_per);
                for (int t = t_lo; t < t_hi; ++t) {
                    mbar_wait(tempty_bar(acc), acc_phase ^ 1);            // both CTAs' epilogues drained this accumulator
                    tc_fence_after();
                    const unsigned d_tmem = tmem_base + (unsigned)(acc * DN);
                    for (int kb = 0; kb < n_kblk; ++kb) {
                        mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        const uint64_t adesc = umma_desc(s32(sa + stage * A_BYTES));
                        const uint64_t bdesc = umma_desc(s32(sb + stage * BH_BYTES));
#pragma unroll
                        for (int kk = 0; kk < DK / UMMA_K; ++kk)
                            tc_mma_2sm(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (kb | kk) ? 1u : 0u);
                        tc_commit_2sm(empty_bar(stage));                  // frees the stage in BOTH CTAs
                        if (++stage == D2STAGES) { stage = 0; phase ^= 1; }
                    }
                    tc_commit_2sm(tfull_bar(acc));                        // accumulators complete in both CTAs
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;
        const int row_in_blk = ew * 32 + lane;
        int acc = 0; unsigned acc_phase = 0;
        for (int u = pair; u < n_units; u += n_pairs) {
            const int mp = u % n_mpairs, sp = u / n_mpairs;
            const int t_lo = sp * tiles_per, t_hi = min(P.n_ntiles, t_lo + tiles_per);
            const int row = mp * 2 * DM + (int)rank * DM + row_in_blk;
            const long long self_col = P.from_base + row - P.to_base;
            float tv[KMAX]; int ti[KMAX];
#pragma unroll
            for (int q = 0; q < KMAX; ++q) { tv[q] = P.min_sim; ti[q] = -1; }
            float kv = P.min_sim; int ki = -1;
            for (int t = t_lo; t < t_hi; ++t) {
                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                const unsigned taddr = tmem_base + ((unsigned)(ew * 32) << 16) + (unsigned)(acc * DN);
                for (int c0 = 0; c0 < DN; c0 += 32) {
                    unsigned r[32];
                    tmem_ld32(taddr + (unsigned)c0, r);
                    const int colb = t * DN + c0;
                    unsigned mask = 0u;
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const float sc = __uint_as_float(r[q]);
                        if (sc > kv || (sc == kv && colb + q < ki)) mask |= 1u << q;
                    }
                    while (mask) {
                        const int q = __ffs(mask) - 1; mask &= mask - 1;
                        float sc = 0.f;
#pragma unroll
                        for (int z = 0; z < 32; ++z) if (z == q) sc = __uint_as_float(r[z]);
                        const int col = colb + q;
                        if (!(sc > kv || (sc == kv && col < ki))) continue;
                        if (col >= P.n_to || (P.self_match && (long long)col == self_col)) continue;
                        float cv = sc; int ci = (int)(P.to_base + col);
#pragma unroll
                        for (int z = 0; z < KMAX; ++z) {
                            if (z < P.k) {
                                const bool before = cv > tv[z] || (cv == tv[z] && (ti[z] < 0 || ci < ti[z]));
                                if (before) { const float fv = tv[z]; const int fi = ti[z]; tv[z] = cv; ti[z] = ci; cv = fv; ci = fi; }
                                if (z == P.k - 1) { kv = tv[z]; ki = ti[z] < 0 ? -1 : ti[z] - (int)P.to_base; }
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive_leader(tempty_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (row < P.n_from) {
                const size_t o = ((size_t)sp * P.n_from + row) * P.k;
#pragma unroll
                for (int z = 0; z < KMAX; ++z)
                    if (z < P.k) { P.top_idx[o + z] = ti[z]; P.top_val[o + z] = ti[z] >= 0 ? (double)tv[z] : 0.0; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                                  // neither CTA frees TMEM while the pair still computes
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
    }
}

// rows -> l2-normalised bf16 (fp64 or fp32 in); zero rows stay zero.  One warp per row.
template <typename T>
__global__ void __launch_bounds__(256) rows_normalize_bf16_kernel(const T *__restrict__ x, int64_t ld, int n_rows, int d, int d_pad, int normalize,
                                                                  __nv_bfloat16 *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int r = gw; r < n_rows; r += nw) {
        const T *row = x + (int64_t)r * ld;
        float ss = 0.f;
        if (normalize) {
            for (int c = lane; c < d; c += 32) { const float v = (float)row[c]; ss += v * v; }
#pragma unroll
            for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(FULL, ss, o);
        }
        const float inv = (normalize && ss > 0.f) ? rsqrtf(ss) : 1.f;
        for (int c = lane; c < d_pad; c += 32) out[(int64_t)r * d_pad + c] = __float2bfloat16(c < d ? (float)row[c] * inv : 0.f);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static int make_map(EncodeTiledFn enc, CUtensorMap *m, const void *base, int n_rows, int d, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)n_rows};
    cuuint64_t strides[1] = {(cuuint64_t)d * 2};
    cuuint32_t box[2] = {(cuuint32_t)DK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return 1; }
    return 0;
}

}  // namespace pfz

using namespace pfz;

extern "C" {

int pfz_rows_to_bf16(const void *x, int32_t is_f64, int64_t ld, int32_t n_rows, int32_t d, int32_t d_pad, int32_t normalize, void *out_bf16,
                     void *stream) {
    if (n_rows <= 0) return 0;
    PFZ_REQUIRE(d_pad >= d && d_pad % 8 == 0, "pfz_rows_to_bf16: d_pad %d must be >= d and a multiple of 8", d_pad);
    int grid = (n_rows + 7) / 8; if (grid > 148 * 16) grid = 148 * 16;
    if (is_f64) rows_normalize_bf16_kernel<double><<<grid, 256, 0, as_stream(stream)>>>((const double *)x, ld, n_rows, d, d_pad, normalize, (__nv_bfloat16 *)out_bf16);
    else        rows_normalize_bf16_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float *)x, ld, n_rows, d, d_pad, normalize, (__nv_bfloat16 *)out_bf16);
    PFZ_LAUNCH_OK();
    return 0;
}

int pfz_dense_cos_topk(const void *x_bf16, const void *y_bf16, int32_t n_from, int32_t n_to, int32_t d, int32_t k, double min_similarity,
                       int32_t self_match, int64_t from_index_base, int64_t to_index_base, int32_t n_splits, int32_t *top_idx, double *top_val,
                       void *stream) {
    PFZ_REQUIRE(k >= 1 && k <= 32, "pfz_dense_cos_topk: k=%d unsupported (1..32)", k);
    PFZ_REQUIRE(d >= 8 && d % 8 == 0, "pfz_dense_cos_topk: d=%d must be a multiple of 8 (16-byte row pitch for TMA)", d);
    PFZ_REQUIRE(((uintptr_t)x_bf16 % 16) == 0 && ((uintptr_t)y_bf16 % 16) == 0, "pfz_dense_cos_topk: operands must be 16-byte aligned");
    if (n_from <= 0) return 0;
    PFZ_REQUIRE(n_to > 0, "pfz_dense_cos_topk: empty to-matrix");
    cudaStream_t st = as_stream(stream);
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
        void *fn = nullptr; cudaDriverEntryPointQueryResult qres;
        PFZ_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        PFZ_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
        enc = (EncodeTiledFn)fn;
    }
    const char *env2 = getenv("PFZ_K4_2CTA");                   // cta_group::2 variant (CTA pairs, half the to-operand traffic per SM): the
    const bool two_cta = env2 ? atoi(env2) != 0 : true;         // default since it was validated on hardware (13.66 vs 14.04 ms at 100k x 100k x 768); 0 = single-CTA kernel
    CUtensorMap mx, my;
    if (make_map(enc, &mx, x_bf16, n_from, d, DM)) return 1;
    if (make_map(enc, &my, y_bf16, n_to, d, two_cta ? DN / 2 : DN)) return 1;
    DenseParams P;
    P.n_from = n_from; P.n_to = n_to; P.d = d; P.k = k; P.min_sim = (float)min_similarity; P.self_match = self_match;
    P.from_base = from_index_base; P.to_base = to_index_base;
    P.n_mblocks = (n_from + DM - 1) / DM; P.n_ntiles = (n_to + DN - 1) / DN;
    PFZ_REQUIRE(n_splits >= 1 && n_splits <= P.n_ntiles, "pfz_dense_cos_topk: n_splits %d out of range (1..%d)", n_splits, P.n_ntiles);
    P.n_splits = n_splits; P.top_idx = top_idx; P.top_val = top_val;
    int dev = 0, sms = 0;
    PFZ_CUDA_OK(cudaGetDevice(&dev));
    PFZ_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (two_cta) {
        const size_t smem2 = (size_t)D2STAGES * (A_BYTES + BH_BYTES) + (2 * D2STAGES + 4) * 8 + 16 + 1024;
        const int n_mpairs = (n_from + 2 * DM - 1) / (2 * DM);
        int pairs = n_mpairs * n_splits; if (pairs > sms / 2) pairs = sms / 2;
#define PFZ_DENSE2_LAUNCH(KM)                                                                                                 \
    do {                                                                                                                      \
        PFZ_CUDA_OK(cudaFuncSetAttribute(dense_cos_topk2_kernel<KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)); \
        dense_cos_topk2_kernel<KM><<<2 * pairs, DENSE_THREADS, smem2, st>>>(mx, my, P);                                        \
    } while (0)
        if (k <= 4) PFZ_DENSE2_LAUNCH(4); else if (k <= 10) PFZ_DENSE2_LAUNCH(10); else if (k <= 16) PFZ_DENSE2_LAUNCH(16); else PFZ_DENSE2_LAUNCH(32);
#undef PFZ_DENSE2_LAUNCH
        PFZ_LAUNCH_OK();
        return 0;
    }
    const size_t smem = (size_t)DSTAGES * (A_BYTES + B_BYTES) + (2 * DSTAGES + 4) * 8 + 16 + 1024;
    int grid = P.n_mblocks * n_splits; if (grid > sms) grid = sms;
#define PFZ_DENSE_LAUNCH(KM)                                                                                                \
    do {                                                                                                                    \
        PFZ_CUDA_OK(cudaFuncSetAttribute(dense_cos_topk_kernel<KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        dense_cos_topk_kernel<KM><<<grid, DENSE_THREADS, smem, st>>>(mx, my, P);                                             \
    } while (0)
    if (k <= 4) PFZ_DENSE_LAUNCH(4); else if (k <= 10) PFZ_DENSE_LAUNCH(10); else if (k <= 16) PFZ_DENSE_LAUNCH(16); else PFZ_DENSE_LAUNCH(32);
#undef PFZ_DENSE_LAUNCH
    PFZ_LAUNCH_OK();
    return 0;
}
}
