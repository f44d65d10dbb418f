// pfz_spcos_block.cu -- K2, from-row-BLOCK variant (PFZ_K2_BLOCK): sparse cosine + fused per-row top-k where one
// warp scores a block of up to 8 from-rows against a to-tile at a time, so that ONE load of a posting chunk serves
// every from-row of the block that contains the term.
//
// Replaces sparse_dot_topn.awesome_cossim_topn (call site polyfuzz/models/_utils.py:82) and the reference's Python
// post-processing (polyfuzz/models/_utils.py:84-91, 128-146), like the other K2 variants (pfz_spcos.cu); results are
// bit-identical to them (same exact fp64 re-scoring, same ranking key).
//
// Why: ncu on the per-row kernels (profiles/k2_r01_bench_metrics.txt) showed them bound by instruction issue -- ~93
// warp instructions per 32 postings, most of it fetching postings and per-(row, tile) bookkeeping.  On name-like
// data a few hundred n-grams carry > 90 % of the postings ("inc", "llc", "cor", ...), so from-rows share their heavy
// terms.  Here
//   * from-rows are CLUSTERED by their three heaviest terms (sort of a 64-bit key), so a block of 8 consecutive rows
//     shares most heavy terms (company names: one posting load serves ~4 rows on average);
//   * a block table (per block: distinct terms ascending, per term the list of (row-in-block, fp32 weight)) is built
//     once per call; per (block, tile) the warp tables one work item per <= 32 postings of one term and consumes the
//     items with a software-pipelined loop: posting chunk -> registers once, then one shared-memory
//     read-modify-write (ld, fma, st) per from-row that has the term, unrolled per row count (jump table);
//   * accumulators acc[8][tile] are fp32 and only FILTER (as in PFZ_K2_DENSE32): at the end of the (block, tile)
//     unit each row's accumulators are scanned (and cleared) against the row's threshold (k-th key - MARGIN, or a
//     lane-maxima bound while the list is still filling); cells above it are re-scored exactly -- fp64, ascending
//     terms, products rounded before the add -- by merging the two CSR rows.
// Idle lanes of a partial chunk add w = 0 to a padding cell acc[f][tile] (rows are tile + 4 floats apart), so the inner
// loop carries no predicates.
#include <stdlib.h>
#include "pfz_common.cuh"

namespace pfz {

constexpr double K2B_MARGIN = 3e-5;     // see K2_MARGIN in pfz_spcos.cu: |fp32 sum - exact| < 7.7e-6 for rows <= 128 terms
constexpr int BF = 8;                   // from-rows per block
constexpr int FV_CAP = 512;             // (row, weight) entries per block table: a block with more is split in two halves
constexpr int ITEM_CAP = 96;            // work items per batch (one term contributes <= tile/32 <= 64)
constexpr int DEPTH = 4;                // posting chunks in flight per warp
constexpr int RANK_CAP = 16383;         // term ranks are capped to 14 bits in the clustering key

struct __align__(16) BlockDesc { int pos0; int nrows; int base; int nterms; };
struct __align__(16) BItem { int off; int cnt; unsigned fvs; int nf; };

// ---- preparation kernels -------------------------------------------------------------------------------------
// terms sorted by document frequency in the to-shard, heaviest first: key = (~df) << 32 | term
__global__ void blk_term_key_kernel(const int32_t *__restrict__ seg, int n_vocab, int n_tiles, uint64_t *__restrict__ keys, int64_t n_pad) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_pad; t += (int64_t)gridDim.x * blockDim.x) {
        uint64_t k = ~0ull;
        if (t < n_vocab) {
            const uint32_t df = (uint32_t)(seg[(t + 1) * n_tiles] - seg[t * n_tiles]);
            k = ((uint64_t)(0xffffffffu - df) << 32) | (uint32_t)t;
        }
        keys[t] = k;
    }
}
__global__ void blk_term_rank_kernel(const uint64_t *__restrict__ keys, int n_vocab, int32_t *__restrict__ rank) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_vocab; i += gridDim.x * blockDim.x)
        rank[(uint32_t)(keys[i] & 0xffffffffu)] = min(i, RANK_CAP);
}
// from-row clustering key: the ranks of the row's three heaviest terms, then the row id (22 bits)
__global__ void blk_row_key_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, const int32_t *__restrict__ rank,
                                   int n_from, uint64_t *__restrict__ keys, int64_t n_pad) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_pad; r += (int64_t)gridDim.x * blockDim.x) {
        uint64_t k = ~0ull;
        if (r < n_from) {
            int r1 = RANK_CAP, r2 = RANK_CAP, r3 = RANK_CAP;
            for (int p = indptr[r]; p < indptr[r + 1]; ++p) {
                const int x = rank[indices[p]];
                if (x < r1) { r3 = r2; r2 = r1; r1 = x; }
                else if (x < r2) { r3 = r2; r2 = x; }
                else if (x < r3) r3 = x;
            }
            k = ((uint64_t)r1 << 50) | ((uint64_t)r2 << 36) | ((uint64_t)r3 << 22) | (uint64_t)r;
        }
        keys[r] = k;
    }
}
__global__ void blk_perm_kernel(const uint64_t *__restrict__ keys, const int32_t *__restrict__ indptr, int n_from, int32_t *__restrict__ perm,
                                int32_t *__restrict__ nnz_perm) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p <= n_from; p += gridDim.x * blockDim.x) {
        if (p == n_from) { nnz_perm[p] = 0; continue; }
        const int r = (int)(keys[p] & 0x3fffffull);
        perm[p] = r;
        nnz_perm[p] = indptr[r + 1] - indptr[r];
    }
}

// Block tables: one warp per group of BF consecutive clustered positions.  A group whose rows hold more than FV_CAP
// entries is emitted as two descriptors of BF/2 rows (<= 4 x 128 entries); descriptor slots 2g, 2g+1 (nrows = 0: unused).
// Per descriptor, at offset base = pos_ptr[first position]:
//   blk_terms[base + u]  = u-th distinct term (ascending)
//   blk_fvdesc[base + u] = (start << 4) | count of its (row, weight) entries in blk_fv[base + start ...], rows ascending
//   blk_fv[base + e]     = { row-in-block * row_stride_bytes, fp32 weight bits }
__global__ void __launch_bounds__(128) blk_table_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                                        const double *__restrict__ data, int n_from, const int32_t *__restrict__ perm,
                                                        const int32_t *__restrict__ pos_ptr, int row_stride_bytes, int32_t *__restrict__ blk_terms,
                                                        int32_t *__restrict__ blk_fvdesc, uint2 *__restrict__ blk_fv, BlockDesc *__restrict__ descs,
                                                        int n_groups, int32_t *__restrict__ err_flag) {
    __shared__ uint32_t s_key[4][FV_CAP];
    __shared__ float s_val[4][FV_CAP];
    const int lane = lane_id(), w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t *keys = s_key[w]; float *vals = s_val[w];
    for (int g = blockIdx.x * 4 + w; g < n_groups; g += gridDim.x * 4) {
        const int p0 = g * BF, p1 = min(n_from, p0 + BF);
        const int total = pos_ptr[p1] - pos_ptr[p0];
        const int nsub = total <= FV_CAP ? 1 : 2;
        for (int sub = 0; sub < 2; ++sub) {
            BlockDesc d; d.pos0 = 0; d.nrows = 0; d.base = 0; d.nterms = 0;
            if (sub < nsub) {
                const int q0 = nsub == 1 ? p0 : min(p1, p0 + sub * (BF / 2));
                const int q1 = nsub == 1 ? p1 : min(p1, q0 + BF / 2);
                const int base = pos_ptr[q0];
                const int cnt = pos_ptr[q1] - base;
                if (cnt > FV_CAP) { if (lane == 0) atomicExch(err_flag, 1); }          // a row with > 128 terms: caller's contract broken
                else if (q1 > q0) {
                    for (int f = 0; f < q1 - q0; ++f) {
                        const int row = perm[q0 + f];
                        const int a0 = indptr[row], m = indptr[row + 1] - a0, o = pos_ptr[q0 + f] - base;
                        for (int e = lane; e < m; e += 32) { keys[o + e] = ((uint32_t)indices[a0 + e] << 3) | (uint32_t)f; vals[o + e] = (float)data[a0 + e]; }
                    }
                    int P = 1;
                    while (P < cnt) P <<= 1;
                    for (int e = cnt + lane; e < P; e += 32) { keys[e] = 0xffffffffu; vals[e] = 0.f; }
                    __syncwarp();
                    for (int k = 2; k <= P; k <<= 1) {
                        for (int j = k >> 1; j > 0; j >>= 1) {
                            for (int t = lane; t < (P >> 1); t += 32) {
                                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                                const int ix = i | j;
                                const uint32_t a = keys[i], b = keys[ix];
                                const bool up = (i & k) == 0;
                                if ((a > b) == up) { keys[i] = b; keys[ix] = a; const float va = vals[i]; vals[i] = vals[ix]; vals[ix] = va; }
                            }
                            __syncwarp();
                        }
                    }
                    int nd = 0;
                    for (int e0 = 0; e0 < cnt; e0 += 32) {
                        const int e = e0 + lane;
                        bool head = false; uint32_t term = 0;
                        if (e < cnt) {
                            const uint32_t key = keys[e];
                            term = key >> 3;
                            head = (e == 0) || (keys[e - 1] >> 3) != term;
                            blk_fv[base + e] = make_uint2((key & 7u) * (uint32_t)row_stride_bytes, __float_as_uint(vals[e]));
                        }
                        const unsigned hm = __ballot_sync(FULL, head);
                        if (head) {
                            int len = 1;
                            while (e + len < cnt && (keys[e + len] >> 3) == term) ++len;
                            const int u = nd + __popc(hm & lt);
                            blk_terms[base + u] = (int32_t)term;
                            blk_fvdesc[base + u] = (e << 4) | len;
                        }
                        nd += __popc(hm);
                    }
                    d.pos0 = q0; d.nrows = q1 - q0; d.base = base; d.nterms = nd;
                    __syncwarp();
                }
            }
            if (lane == 0) descs[2 * g + sub] = d;
        }
    }
}

__global__ void blk_pack_kernel(const uint16_t *__restrict__ post_idx, const float *__restrict__ post_val32, const int32_t *__restrict__ nnz_ptr,
                                uint2 *__restrict__ post_pk) {
    const int64_t n = *nnz_ptr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        post_pk[i] = make_uint2((uint32_t)post_idx[i], __float_as_uint(post_val32[i]));
}

// ---- main kernel -----------------------------------------------------------------------------------------------
struct BlockParams {
    const int32_t *a_indptr; const int32_t *a_indices; const double *a_data; int n_from;
    const int32_t *perm; const BlockDesc *descs; int n_desc;
    const int32_t *blk_terms; const int32_t *blk_fvdesc; const uint2 *blk_fv;
    const int32_t *seg; const uint2 *post_pk;
    const int32_t *b_indptr; const int32_t *b_indices; const double *b_data;
    int tile, n_tiles, n_to, k; double min_sim; int self_match; int64_t from_base, to_base; int n_splits;
    int32_t *top_idx; double *top_val; int32_t *counter;
};

__device__ __forceinline__ unsigned sm_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds_f32(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_f32(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void lds_fv(unsigned a, unsigned &off, float &v) {
    unsigned b;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(off), "=r"(b) : "r"(a) : "memory");
    v = __uint_as_float(b);
}

// one posting chunk (already in registers: cell = address of acc[0][row], w = weight) applied to the N from-rows that
// contain the term: all loads, then all fmas, then all stores (distinct rows => distinct addresses)
template <int N>
__device__ __forceinline__ void rmw_group(unsigned cell, unsigned fvs, float w) {
    unsigned a[N]; float v[N], o[N];
#pragma unroll
    for (int q = 0; q < N; ++q) { unsigned off; lds_fv(fvs + 8u * q, off, v[q]); a[q] = cell + off; }
#pragma unroll
    for (int q = 0; q < N; ++q) o[q] = lds_f32(a[q]);
#pragma unroll
    for (int q = 0; q < N; ++q) o[q] = fmaf(v[q], w, o[q]);
#pragma unroll
    for (int q = 0; q < N; ++q) sts_f32(a[q], o[q]);
}
__device__ __forceinline__ void consume_item(int nf, unsigned cell, unsigned fvs, float w) {
    switch (nf) {
        case 1: rmw_group<1>(cell, fvs, w); break;
        case 2: rmw_group<2>(cell, fvs, w); break;
        case 3: rmw_group<3>(cell, fvs, w); break;
        case 4: rmw_group<4>(cell, fvs, w); break;
        case 5: rmw_group<5>(cell, fvs, w); break;
        case 6: rmw_group<6>(cell, fvs, w); break;
        case 7: rmw_group<7>(cell, fvs, w); break;
        case 8: rmw_group<8>(cell, fvs, w); break;
        default: break;                                     // padding item
    }
}
// fetch work item `item_s` and its posting chunk: lanes below cnt load (tile-local row, weight); the others address the
// padding cell (row index == tile) with weight 0
__device__ __forceinline__ void load_slot(unsigned item_s, const uint2 *pk_lane, int lane, unsigned acc_s, unsigned pad_row, unsigned &cell, float &w,
                                          unsigned &fvs, int &nf) {
    unsigned off; int cnt;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(off), "=r"(cnt), "=r"(fvs), "=r"(nf) : "r"(item_s) : "memory");
    unsigned jl = pad_row, wb = 0u;
    asm volatile("{ .reg .pred p; setp.lt.s32 p, %2, %3; @p ld.global.nc.v2.u32 {%0, %1}, [%4]; }"
                 : "+r"(jl), "+r"(wb) : "r"(lane), "r"(cnt), "l"(pk_lane + off) : "memory");
    cell = acc_s + (jl << 2);
    w = __uint_as_float(wb);
}

__device__ __forceinline__ bool blk_key_before(double sa, int ia, double sb, int ib) { return (sa > sb) || (sa == sb && ia < ib); }
__device__ __forceinline__ double blk_exact_dot(const int32_t *__restrict__ ai, const double *__restrict__ av, int an,
                                                const int32_t *__restrict__ bi, const double *__restrict__ bv, int bn) {
    double s = 0.0;
    int p = 0, q = 0;
    while (p < an && q < bn) {
        const int ca = ai[p], cb = bi[q];
        if (ca == cb) { s = __dadd_rn(s, __dmul_rn(av[p], bv[q])); ++p; ++q; }
        else if (ca < cb) ++p; else ++q;
    }
    return s;
}
__device__ __forceinline__ float blk_float_floor(double x) {        // largest float not above x (x >= 0)
    float f = (float)x;
    if ((double)f > x) f = __uint_as_float(__float_as_uint(f) - 1u);
    return f;
}
// descending bitonic sort of one value per lane: lane r ends up with the r-th largest
__device__ __forceinline__ float warp_sort_desc(float x, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const float o = __shfl_xor_sync(FULL, x, j);
            const bool keep_max = ((lane & j) == 0) == ((lane & k) == 0);
            x = keep_max ? fmaxf(x, o) : fminf(x, o);
        }
    }
    return x;
}

// per-warp shared-memory arena (bytes), T = tile
__host__ __device__ inline size_t blk_arena_bytes(int T) {
    return (size_t)BF * (T + 4) * 4            // acc
           + (size_t)(ITEM_CAP + 2 * DEPTH) * 16  // items
           + (size_t)FV_CAP * 8                // terms + fvdesc (uint2)
           + (size_t)FV_CAP * 8                // fv
           + (size_t)BF * 32 * 8 + (size_t)BF * 32 * 4   // top-k lists
           + 64 * 4                            // candidates
           + 512;                              // per-row scalars
}

__global__ void __launch_bounds__(32) spcos_block_kernel(const BlockParams P) {
    extern __shared__ __align__(16) unsigned char dyn[];
    const int lane = lane_id();
    const unsigned lt = (1u << lane) - 1u;
    const int T = P.tile, TS = T + 4, K = P.k;
    unsigned char *base = dyn;
    float *acc = reinterpret_cast<float *>(base);                       base += (size_t)BF * TS * 4;
    BItem *items = reinterpret_cast<BItem *>(base);                     base += (size_t)(ITEM_CAP + 2 * DEPTH) * 16;
    uint2 *termtab = reinterpret_cast<uint2 *>(base);                   base += (size_t)FV_CAP * 8;      // {term, fvdesc}
    uint2 *fvtab = reinterpret_cast<uint2 *>(base);                     base += (size_t)FV_CAP * 8;
    double *list_v = reinterpret_cast<double *>(base);                  base += (size_t)BF * 32 * 8;
    int *list_i = reinterpret_cast<int *>(base);                        base += (size_t)BF * 32 * 4;
    int *cand = reinterpret_cast<int *>(base);                          base += 64 * 4;
    double *kv_s = reinterpret_cast<double *>(base);                    // [BF]
    int *ki_s = reinterpret_cast<int *>(base + 64);                     // [BF]
    int *row_s = ki_s + BF, *a0_s = row_s + BF, *m_s = a0_s + BF, *stau_s = m_s + BF, *sjl_s = stau_s + BF;
    float *thr_s = reinterpret_cast<float *>(sjl_s + BF);               // [BF]

    for (int q = lane; q < BF * TS; q += 32) acc[q] = 0.f;
    __syncwarp();
    const unsigned acc_s = sm_u32(acc), items_s = sm_u32(items), fv_s = sm_u32(fvtab);
    const uint2 *pk_lane = P.post_pk + lane;
    const int32_t *__restrict__ seg = P.seg;
    const int n_tiles = P.n_tiles;

    const int split = blockIdx.y;
    const int tiles_per = (P.n_tiles + P.n_splits - 1) / P.n_splits;
    const int tau_lo = split * tiles_per;
    const int tau_hi = min(P.n_tiles, tau_lo + tiles_per);
    int32_t *counter = P.counter + split;
    const float thr0 = blk_float_floor(fmax(P.min_sim - K2B_MARGIN, 0.0));

    for (;;) {
        int di = 0;
        if (lane == 0) di = atomicAdd(counter, 1);
        di = __shfl_sync(FULL, di, 0);
        if (di >= P.n_desc) break;
        const BlockDesc D = P.descs[di];
        const int nrows = D.nrows, nterms = D.nterms;
        if (nrows == 0) continue;
        // per-row state
        int m_lane = 0;
        if (lane < BF) {
            int row = 0, a0 = 0, m = 0, st = -1, sj = 0;
            if (lane < nrows) {
                row = P.perm[D.pos0 + lane];
                a0 = P.a_indptr[row]; m = P.a_indptr[row + 1] - a0;
                const int64_t self_j = P.from_base + row - P.to_base;       // local to-row of the diagonal
                if (P.self_match && self_j >= 0 && self_j < (int64_t)P.n_to) { st = (int)(self_j / T); sj = (int)(self_j - (int64_t)st * T); }
            }
            m_lane = m;
            row_s[lane] = row; a0_s[lane] = a0; m_s[lane] = m; stau_s[lane] = st; sjl_s[lane] = sj;
            kv_s[lane] = P.min_sim; ki_s[lane] = -1; thr_s[lane] = thr0;
        }
        for (int q = lane; q < BF * 32; q += 32) { list_v[q] = P.min_sim; list_i[q] = -1; }
        for (int u = lane; u < nterms; u += 32) termtab[u] = make_uint2((unsigned)P.blk_terms[D.base + u], (unsigned)P.blk_fvdesc[D.base + u]);
        // (row, weight) entries: the table holds exactly the sum of the rows' nnz
        int nfv_total = m_lane;
#pragma unroll
        for (int dlt = 16; dlt; dlt >>= 1) nfv_total += __shfl_xor_sync(FULL, nfv_total, dlt);
        for (int e = lane; e < nfv_total; e += 32) fvtab[e] = P.blk_fv[D.base + e];
        __syncwarp();

        int ncand = 0;
        // exact scoring + insertion of up to 32 queued candidates (one per lane)
        auto score_round = [&](int n_round) {
            double sc = 0.0; int j = -1, f = 0; bool cnd = false;
            if (lane < n_round) {
                const int c = cand[lane];
                f = (int)((unsigned)c >> 28);
                const int jloc = c & 0x0fffffff;
                const int b0 = P.b_indptr[jloc];
                const int a0 = a0_s[f];
                sc = blk_exact_dot(P.a_indices + a0, P.a_data + a0, m_s[f], P.b_indices + b0, P.b_data + b0, P.b_indptr[jloc + 1] - b0);
                j = (int)(P.to_base + jloc);
                cnd = blk_key_before(sc, j, kv_s[f], ki_s[f]);
                if (stau_s[f] >= 0 && jloc == stau_s[f] * T + sjl_s[f]) cnd = false;
            }
            unsigned cm = __ballot_sync(FULL, cnd);
            while (cm) {
                const int src = __ffs(cm) - 1;
                const double cs = shfl_d(sc, src);
                const int cj = __shfl_sync(FULL, j, src);
                const int cf = __shfl_sync(FULL, f, src);
                double lv = list_v[cf * 32 + lane]; int li = list_i[cf * 32 + lane];
                const bool stays = (lane < K) && blk_key_before(lv, li, cs, cj);
                const int pos = __popc(__ballot_sync(FULL, stays));
                const double uv = __shfl_up_sync(FULL, lv, 1);
                const int ui = __shfl_up_sync(FULL, li, 1);
                if (lane > pos) { lv = uv; li = ui; }
                else if (lane == pos) { lv = cs; li = cj; }
                if (lane < K) { list_v[cf * 32 + lane] = lv; list_i[cf * 32 + lane] = li; }
                const double nkv = shfl_d(lv, K - 1);
                const int nki = __shfl_sync(FULL, li, K - 1);
                if (lane == 0) {
                    kv_s[cf] = nkv; ki_s[cf] = nki;
                    if (nki >= 0) thr_s[cf] = blk_float_floor(fmax(nkv - K2B_MARGIN, 0.0));
                }
                __syncwarp();
                cnd = cnd && lane != src && (f != cf || blk_key_before(sc, j, nkv, nki));
                cm = __ballot_sync(FULL, cnd);
            }
        };
        auto flush_full_round = [&]() {
            score_round(32);
            __syncwarp();
            const int rest = ncand - 32;
            int mv = 0;
            if (lane < rest) mv = cand[32 + lane];
            __syncwarp();
            if (lane < rest) cand[lane] = mv;
            ncand = rest;
            __syncwarp();
        };

        for (int tau = tau_lo; tau < tau_hi; ++tau) {
            bool any_post = false;
            for (int tb = 0; tb < nterms; tb += 32) {
                const int u = tb + lane;
                int s = 0, len = 0; unsigned fvd = 0u;
                if (u < nterms) {
                    const uint2 tt = termtab[u];
                    fvd = tt.y;
                    const int64_t c = (int64_t)tt.x * n_tiles + tau;
                    s = seg[c];
                    len = seg[c + 1] - s;
                }
                const int nch = (len + 31) >> 5;
                const int incl = warp_incl_scan(nch);
                const int n_total = __shfl_sync(FULL, incl, 31);
                if (n_total == 0) continue;
                any_post = true;
                for (int start = 0; start < n_total;) {
                    const bool inb = nch > 0 && incl - nch >= start && incl <= start + ITEM_CAP;
                    const unsigned bm = __ballot_sync(FULL, inb);
                    const int endv = __shfl_sync(FULL, incl, 31 - __clz(bm));
                    const int N = endv - start;
                    if (inb) {
                        int o = incl - nch - start, so = s, rem = len;
                        const unsigned fvs = fv_s + (fvd >> 4) * 8u;
                        const int nf = (int)(fvd & 15u);
                        while (rem > 0) { BItem it; it.off = so; it.cnt = rem; it.fvs = fvs; it.nf = nf; items[o] = it; ++o; so += 32; rem -= 32; }
                    }
                    if (lane < 2 * DEPTH) { BItem it; it.off = 0; it.cnt = 0; it.fvs = fv_s; it.nf = 0; items[N + lane] = it; }
                    __syncwarp();
                    {
                        unsigned cell[DEPTH], fvs[DEPTH]; float w[DEPTH]; int nf[DEPTH];
#pragma unroll
                        for (int d = 0; d < DEPTH; ++d) load_slot(items_s + d * 16, pk_lane, lane, acc_s, (unsigned)T, cell[d], w[d], fvs[d], nf[d]);
                        unsigned it_s = items_s + DEPTH * 16;
                        for (int b = 0; b < N; b += DEPTH) {
#pragma unroll
                            for (int d = 0; d < DEPTH; ++d) {
                                consume_item(nf[d], cell[d], fvs[d], w[d]);
                                load_slot(it_s + d * 16, pk_lane, lane, acc_s, (unsigned)T, cell[d], w[d], fvs[d], nf[d]);
                            }
                            it_s += DEPTH * 16;
                        }
                    }
                    __syncwarp();
                    start = endv;
                }
            }
            if (!any_post) continue;
            // scan + clear: every row's accumulators against the row's threshold
            for (int f = 0; f < nrows; ++f) {
                float4 *rowp = reinterpret_cast<float4 *>(acc + (size_t)f * TS);
                if (stau_s[f] == tau) { if (lane == 0) acc[(size_t)f * TS + sjl_s[f]] = 0.f; __syncwarp(); }   // the diagonal never competes
                float gate = thr_s[f];
                if (ki_s[f] < 0) {
                    // list not full yet: the K-th largest of the 32 lane maxima bounds the unit's K-th best score from below
                    float mx = 0.f;
                    for (int c = lane; c < (T >> 2); c += 32) { const float4 v = rowp[c]; mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w)); }
                    const float srt = warp_sort_desc(mx, lane);
                    const float kth = __shfl_sync(FULL, srt, K - 1);
                    if (kth > 0.f) gate = fmaxf(gate, blk_float_floor(fmax((double)kth - K2B_MARGIN, 0.0)));
                }
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int c0 = 0; c0 < (T >> 2); c0 += 32) {
                    const int c = c0 + lane;
                    float4 v = z;
                    if (c < (T >> 2)) { v = rowp[c]; rowp[c] = z; }
                    const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
                    if (!__any_sync(FULL, mx > gate)) continue;
#pragma unroll
                    for (int comp = 0; comp < 4; ++comp) {
                        const float x = comp == 0 ? v.x : comp == 1 ? v.y : comp == 2 ? v.z : v.w;
                        const bool take = x > gate;
                        const unsigned tm = __ballot_sync(FULL, take);
                        if (tm == 0u) continue;
                        if (take) cand[ncand + __popc(tm & lt)] = (f << 28) | (tau * T + c * 4 + comp);
                        ncand += __popc(tm);
                        __syncwarp();
                        if (ncand >= 32) flush_full_round();
                    }
                }
            }
            __syncwarp();
        }
        if (ncand > 0) { __syncwarp(); score_round(ncand); __syncwarp(); ncand = 0; }
        for (int f = 0; f < nrows; ++f) {
            if (lane < K) {
                const size_t o = ((size_t)split * P.n_from + row_s[f]) * K + lane;
                const int li = list_i[f * 32 + lane];
                P.top_idx[o] = li;
                P.top_val[o] = (li >= 0) ? list_v[f * 32 + lane] : 0.0;
            }
        }
        __syncwarp();
    }
}

static int blk_grid(int64_t work, int threads, int cap) {
    int64_t g = (work + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}
static int64_t pow2_at_least(int64_t n) { int64_t p = 2; while (p < n) p <<= 1; return p; }
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct BlockWs {
    size_t term_keys, term_rank, row_keys, perm, pos_ptr, scan_ws, blk_terms, blk_fvdesc, blk_fv, descs, counters, err, total;
};
static BlockWs block_ws_layout(int64_t n_from, int64_t nnz_cap, int64_t n_vocab, int n_splits) {
    BlockWs L; size_t o = 0;
    const int64_t vp = pow2_at_least(n_vocab), np = pow2_at_least(n_from);
    const int64_t n_groups = (n_from + BF - 1) / BF;
    L.term_keys = o; o += align256((size_t)vp * 8);
    L.term_rank = o; o += align256((size_t)(n_vocab + 1) * 4);
    L.row_keys = o; o += align256((size_t)np * 8);
    L.perm = o; o += align256((size_t)(n_from + 1) * 4);
    L.pos_ptr = o; o += align256((size_t)(n_from + 2) * 4);
    L.scan_ws = o; o += align256((size_t)pfz_scan_ws_bytes(n_from + 2));
    L.blk_terms = o; o += align256((size_t)(nnz_cap + 1) * 4);
    L.blk_fvdesc = o; o += align256((size_t)(nnz_cap + 1) * 4);
    L.blk_fv = o; o += align256((size_t)(nnz_cap + 1) * 8);
    L.descs = o; o += align256((size_t)(2 * n_groups + 1) * sizeof(BlockDesc));
    L.counters = o; o += align256((size_t)(n_splits + 1) * 4);
    L.err = o; o += 256;
    L.total = o;
    return L;
}

}  // namespace pfz

using namespace pfz;

extern "C" {

int64_t pfz_spcos_block_ws_bytes(int32_t n_from, int64_t nnz_cap_from, int32_t n_vocab, int32_t n_splits) {
    return (int64_t)block_ws_layout(n_from, nnz_cap_from, n_vocab, n_splits).total;
}

int pfz_index_pack32(const uint16_t *post_idx, const float *post_val32, const int32_t *nnz_dev, void *post_pk, void *stream) {
    blk_pack_kernel<<<148 * 8, 256, 0, as_stream(stream)>>>(post_idx, post_val32, nnz_dev, reinterpret_cast<uint2 *>(post_pk));
    PFZ_LAUNCH_OK();
    return 0;
}

int pfz_spcos_topk_block(const int32_t *a_indptr, const int32_t *a_indices, const double *a_data, int32_t n_from, int64_t nnz_cap_from,
                         const int32_t *seg, const void *post_pk, const int32_t *b_indptr, const int32_t *b_indices, const double *b_data,
                         int32_t n_vocab, int32_t tile, int32_t n_tiles, int32_t n_to, int32_t k, double min_similarity, int32_t self_match,
                         int64_t from_index_base, int64_t to_index_base, int32_t n_splits, int32_t *top_idx, double *top_val,
                         int32_t *err_flag_dev, void *ws, void *stream) {
    PFZ_REQUIRE(k >= 1 && k <= 32, "pfz_spcos_topk_block: k=%d unsupported (1..32)", k);
    PFZ_REQUIRE(tile >= 128 && tile <= 2048 && (tile % 128) == 0, "pfz_spcos_topk_block: tile %d must be a multiple of 128 in 128..2048", tile);
    PFZ_REQUIRE(n_splits >= 1 && n_splits <= n_tiles, "pfz_spcos_topk_block: n_splits %d out of range", n_splits);
    PFZ_REQUIRE(n_from < (1 << 22), "pfz_spcos_topk_block: n_from %d exceeds the 22-bit row id of the clustering key", n_from);
    PFZ_REQUIRE((int64_t)n_to < (1ll << 28), "pfz_spcos_topk_block: n_to %d exceeds 2^28", n_to);
    PFZ_REQUIRE(n_vocab < (1 << 29), "pfz_spcos_topk_block: n_vocab too large");
    if (n_from <= 0) return 0;
    cudaStream_t st = as_stream(stream);
    int dev = 0, sms = 0, smem_max = 0;
    PFZ_CUDA_OK(cudaGetDevice(&dev));
    PFZ_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    PFZ_CUDA_OK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t arena = (blk_arena_bytes(tile) + 15) & ~(size_t)15;
    PFZ_REQUIRE(arena <= (size_t)smem_max, "pfz_spcos_topk_block: tile %d needs %zu B shared memory > %d available", tile, arena, smem_max);
    const BlockWs L = block_ws_layout(n_from, nnz_cap_from, n_vocab, n_splits);
    char *w = reinterpret_cast<char *>(ws);
    uint64_t *term_keys = reinterpret_cast<uint64_t *>(w + L.term_keys);
    int32_t *term_rank = reinterpret_cast<int32_t *>(w + L.term_rank);
    uint64_t *row_keys = reinterpret_cast<uint64_t *>(w + L.row_keys);
    int32_t *perm = reinterpret_cast<int32_t *>(w + L.perm);
    int32_t *pos_ptr = reinterpret_cast<int32_t *>(w + L.pos_ptr);
    int32_t *blk_terms = reinterpret_cast<int32_t *>(w + L.blk_terms);
    int32_t *blk_fvdesc = reinterpret_cast<int32_t *>(w + L.blk_fvdesc);
    uint2 *blk_fv = reinterpret_cast<uint2 *>(w + L.blk_fv);
    BlockDesc *descs = reinterpret_cast<BlockDesc *>(w + L.descs);
    int32_t *counters = reinterpret_cast<int32_t *>(w + L.counters);
    const int64_t vp = pow2_at_least(n_vocab), np = pow2_at_least(n_from);
    const int n_groups = (n_from + BF - 1) / BF;

    blk_term_key_kernel<<<blk_grid(vp, 256, 148 * 8), 256, 0, st>>>(seg, n_vocab, n_tiles, term_keys, vp);
    PFZ_LAUNCH_OK();
    if (pfz_sort_u64(term_keys, vp, stream)) return 1;
    blk_term_rank_kernel<<<blk_grid(n_vocab, 256, 148 * 8), 256, 0, st>>>(term_keys, n_vocab, term_rank);
    PFZ_LAUNCH_OK();
    blk_row_key_kernel<<<blk_grid(np, 256, 148 * 16), 256, 0, st>>>(a_indptr, a_indices, term_rank, n_from, row_keys, np);
    PFZ_LAUNCH_OK();
    if (pfz_sort_u64(row_keys, np, stream)) return 1;
    blk_perm_kernel<<<blk_grid(n_from + 1, 256, 148 * 16), 256, 0, st>>>(row_keys, a_indptr, n_from, perm, pos_ptr);
    PFZ_LAUNCH_OK();
    if (scan_exclusive_i32(pos_ptr, pos_ptr, (int64_t)n_from + 1, w + L.scan_ws, st)) return 1;
    blk_table_kernel<<<blk_grid((int64_t)n_groups * 32, 128, 148 * 16), 128, 0, st>>>(a_indptr, a_indices, a_data, n_from, perm, pos_ptr, (tile + 4) * 4,
                                                                                      blk_terms, blk_fvdesc, blk_fv, descs, n_groups, err_flag_dev);
    PFZ_LAUNCH_OK();
    PFZ_CUDA_OK(cudaMemsetAsync(counters, 0, sizeof(int32_t) * (size_t)n_splits, st));
    BlockParams P{a_indptr, a_indices, a_data, n_from, perm, descs, 2 * n_groups, blk_terms, blk_fvdesc, blk_fv, seg,
                  reinterpret_cast<const uint2 *>(post_pk), b_indptr, b_indices, b_data, tile, n_tiles, n_to, k, min_similarity, self_match,
                  from_index_base, to_index_base, n_splits, top_idx, top_val, counters};
    PFZ_CUDA_OK(cudaFuncSetAttribute(spcos_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)arena));
    int occ = 0;
    PFZ_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spcos_block_kernel, 32, arena));
    if (occ < 1) occ = 1;
    int gx = sms * occ;
    if (gx > 2 * n_groups) gx = 2 * n_groups;
    if (n_splits > 1) { gx = (gx + n_splits - 1) / n_splits; if (gx < 1) gx = 1; }
    spcos_block_kernel<<<dim3(gx, n_splits), 32, arena, st>>>(P);
    PFZ_LAUNCH_OK();
    return 0;
}
}
