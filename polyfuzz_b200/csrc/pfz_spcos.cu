// pfz_spcos.cu -- K2: sparse cosine (from CSR x to inverted index) with a fused per-row top-k, plus
// the inverted-index build and the top-k list merge.
//
// Replaces sparse_dot_topn.awesome_cossim_topn (call site polyfuzz/models/_utils.py:82) and the
// reference's Python post-processing (polyfuzz/models/_utils.py:84-91, 128-146).
//
// Layout in HBM
//   from matrix : CSR, int32 indptr/indices (ascending per row), float64 data
//   to index    : postings grouped by (term, to-tile): seg[t*n_tiles + tau] .. seg[t*n_tiles + tau + 1]
//                 post_idx int32 (to-row local to the shard), post_val float64
//   A to-tile is `tile` consecutive to-rows; one warp owns one (from-row, tile) unit at a time with a
//   private fp64 accumulator array acc[tile] in shared memory, so that
//     - the per (from-row, to-row) additions happen in ascending term order (the warp walks the
//       from-row's terms in order; inside one term every to-row occurs at most once => no conflicts),
//     - no atomics and no block barriers are needed (only __syncwarp between terms).
//   First touches are detected by acc == 0 (all weights are > 0) and recorded in a per-warp list, so
//   the selection phase visits touched to-rows only (work ~ postings, not ~ n_from * n_to).
#include "pfz_common.cuh"

namespace pfz {

// ---- inverted index build ---------------------------------------------------------------------
__global__ void index_count_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, int n_rows, int tile, int n_tiles,
                                   int32_t *__restrict__ cnt) {
    const int lane = lane_id();
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int r = gw; r < n_rows; r += nw) {
        const int tau = r / tile;
        for (int p = indptr[r] + lane; p < indptr[r + 1]; p += 32)
            atomicAdd(&cnt[(int64_t)indices[p] * n_tiles + tau], 1);
    }
}

__global__ void index_fill_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, const double *__restrict__ data,
                                  int n_rows, int tile, int n_tiles, const int32_t *__restrict__ seg, int32_t *__restrict__ cur,
                                  uint16_t *__restrict__ post_idx, double *__restrict__ post_val) {
    const int lane = lane_id();
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int r = gw; r < n_rows; r += nw) {
        const int tau = r / tile;
        for (int p = indptr[r] + lane; p < indptr[r + 1]; p += 32) {
            const int64_t c = (int64_t)indices[p] * n_tiles + tau;
            const int pos = seg[c] + atomicAdd(&cur[c], 1);
            post_idx[pos] = (uint16_t)(r - tau * tile);
            post_val[pos] = data[p];
        }
    }
}

// ---- K2 --------------------------------------------------------------------------------------
// ranking key: (score desc, idx asc); "a before b"
__device__ __forceinline__ bool key_before(double sa, int ia, double sb, int ib) {
    return (sa > sb) || (sa == sb && ia < ib);
}

struct SpcosParams {
    const int32_t *a_indptr; const int32_t *a_indices; const double *a_data; int n_from;
    const int32_t *seg; const uint16_t *post_idx; const double *post_val;
    int n_vocab, tile, n_tiles, n_to;
    int k; double min_sim; int self_match; int64_t from_base, to_base;
    int n_splits; const double *excl_val; const int32_t *excl_idx;
    int32_t *top_idx; double *top_val; int32_t *row_counter;
};

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) spcos_list_kernel(const SpcosParams P) {
    extern __shared__ __align__(16) unsigned char dyn[];
    const int lane = lane_id();
    const int w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    // per-warp arenas: acc double[tile] | touched uint16[tile]
    double *acc = reinterpret_cast<double *>(dyn) + (size_t)w * P.tile;
    uint16_t *touched = reinterpret_cast<uint16_t *>(dyn + (size_t)WARPS * P.tile * 8) + (size_t)w * P.tile;
    for (int q = lane; q < P.tile; q += 32) acc[q] = 0.0;
    __syncwarp();

    const int split = blockIdx.y;
    const int tiles_per = (P.n_tiles + P.n_splits - 1) / P.n_splits;
    const int tau_lo = split * tiles_per;
    const int tau_hi = min(P.n_tiles, tau_lo + tiles_per);
    int32_t *counter = P.row_counter + split;

    for (;;) {
        int i = 0;
        if (lane == 0) i = atomicAdd(counter, 1);
        i = __shfl_sync(FULL, i, 0);
        if (i >= P.n_from) break;

        const int a0 = P.a_indptr[i];
        const int m = P.a_indptr[i + 1] - a0;
        // top-k list: lane r holds rank r (r < k); initial entries (min_sim, -1) reject score <= min_sim
        double tv = P.min_sim; int ti = -1;
        double kv = P.min_sim; int ki = -1;             // current k-th entry (threshold)
        // exclusive lower key for paging (candidates must rank strictly after it)
        double xv = 0.0; int xi = -1; bool has_x = false;
        if (P.excl_val) { xv = P.excl_val[i]; xi = P.excl_idx[i]; has_x = xi >= 0; }
        const int64_t self_j = P.from_base + i - P.to_base;     // local to-row of the diagonal

        for (int tau = tau_lo; tau < tau_hi; ++tau) {
            int ntouched = 0;
            for (int tb = 0; tb < m; tb += 32) {
                const int kk = tb + lane;
                int s = 0, len = 0; double v = 0.0;
                if (kk < m) {
                    const int t = P.a_indices[a0 + kk];
                    v = P.a_data[a0 + kk];
                    const int64_t c = (int64_t)t * P.n_tiles + tau;
                    s = P.seg[c];
                    len = P.seg[c + 1] - s;
                }
                unsigned live = __ballot_sync(FULL, len > 0);
                while (live) {                                   // ascending lane == ascending term
                    const int src = __ffs(live) - 1; live &= live - 1;
                    const int ss = __shfl_sync(FULL, s, src);
                    const int sl = __shfl_sync(FULL, len, src);
                    const double sv = shfl_d(v, src);
                    for (int c0 = 0; c0 < sl; c0 += 32) {
                        const int q = c0 + lane;
                        bool first = false; int jl = 0;
                        if (q < sl) {
                            jl = P.post_idx[ss + q];
                            const double prod = __dmul_rn(sv, P.post_val[ss + q]);
                            const double old = acc[jl];
                            acc[jl] = __dadd_rn(old, prod);
                            first = (old == 0.0);
                        }
                        const unsigned fm = __ballot_sync(FULL, first);
                        if (first) touched[ntouched + __popc(fm & lt)] = (uint16_t)jl;
                        ntouched += __popc(fm);
                    }
                    __syncwarp();
                }
            }
            // selection over touched to-rows
            for (int c0 = 0; c0 < ntouched; c0 += 32) {
                const int q = c0 + lane;
                double sc = 0.0; int j = -1; bool cand = false;
                if (q < ntouched) {
                    const int jl = touched[q];
                    sc = acc[jl];
                    acc[jl] = 0.0;
                    const int jloc = tau * P.tile + jl;
                    j = (int)(P.to_base + jloc);
                    cand = key_before(sc, j, kv, ki);
                    if (P.self_match && (int64_t)jloc == self_j) cand = false;
                    if (has_x && !key_before(xv, xi, sc, j)) cand = false;
                }
                unsigned cm = __ballot_sync(FULL, cand);
                while (cm) {
                    const int src = __ffs(cm) - 1; cm &= cm - 1;
                    const double cs = shfl_d(sc, src);
                    const int cj = __shfl_sync(FULL, j, src);
                    if (!key_before(cs, cj, kv, ki)) continue;   // threshold may have risen meanwhile
                    // insert: position = number of list entries that stay before the candidate
                    const bool stays = (lane < P.k) && key_before(tv, ti, cs, cj);
                    const int pos = __popc(__ballot_sync(FULL, stays));
                    const double uv = __shfl_up_sync(FULL, tv, 1);
                    const int ui = __shfl_up_sync(FULL, ti, 1);
                    if (lane > pos) { tv = uv; ti = ui; }
                    else if (lane == pos) { tv = cs; ti = cj; }
                    kv = shfl_d(tv, P.k - 1);
                    ki = __shfl_sync(FULL, ti, P.k - 1);
                }
            }
            __syncwarp();
        }
        if (lane < P.k) {
            const size_t o = ((size_t)split * P.n_from + i) * P.k + lane;
            P.top_idx[o] = ti;
            P.top_val[o] = (ti >= 0) ? tv : 0.0;
        }
    }
}


// ---- K2, dense-regime variant ----------------------------------------------------------------
// Same ownership (one warp = one (from-row, to-tile) unit, private acc[tile]) but tuned for inputs
// where a from-row touches a sizeable fraction of every tile (company names: ~24 %):
//   * no touched list: a to-row becomes a candidate the moment its running sum first passes the
//     current k-th key (partial sums only grow: all weights > 0); candidates are flagged in a tile
//     bitmap (rare after the first tiles) and examined with their FINAL sum at the end of the unit;
//   * acc is cleared densely with 16-byte stores (tile/64 instructions per lane);
//   * per term: one 16-byte broadcast read of {start,len,weight}; up to 4 posting chunks are loaded
//     before any is consumed (memory-level parallelism without relying on occupancy alone).
struct __align__(16) TermSeg { int s; int len; double v; };

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) spcos_dense_kernel(const SpcosParams P) {
    extern __shared__ __align__(16) unsigned char dyn[];
    const int lane = lane_id();
    const int w = threadIdx.x >> 5;
    const int T = P.tile;
    const int nwords = T >> 5;
    // per-warp arena: acc double[T] | terms TermSeg[32] | bitmap uint32[T/32]
    const size_t arena = (size_t)T * 8 + 32 * sizeof(TermSeg) + (size_t)nwords * 4;
    unsigned char *base = dyn + (size_t)w * ((arena + 15) & ~(size_t)15);
    double *acc = reinterpret_cast<double *>(base);
    TermSeg *terms = reinterpret_cast<TermSeg *>(base + (size_t)T * 8);
    unsigned *bitmap = reinterpret_cast<unsigned *>(base + (size_t)T * 8 + 32 * sizeof(TermSeg));
    for (int q = lane; q < T; q += 32) acc[q] = 0.0;
    for (int q = lane; q < nwords; q += 32) bitmap[q] = 0u;
    __syncwarp();

    const int split = blockIdx.y;
    const int tiles_per = (P.n_tiles + P.n_splits - 1) / P.n_splits;
    const int tau_lo = split * tiles_per;
    const int tau_hi = min(P.n_tiles, tau_lo + tiles_per);
    const int ntau = tau_hi - tau_lo;
    int32_t *counter = P.row_counter + split;

    for (;;) {
        int i = 0;
        if (lane == 0) i = atomicAdd(counter, 1);
        i = __shfl_sync(FULL, i, 0);
        if (i >= P.n_from) break;

        const int a0 = P.a_indptr[i];
        const int m = P.a_indptr[i + 1] - a0;
        double tv = P.min_sim; int ti = -1;
        double kv = P.min_sim; int ki = -1;
        double xv = 0.0; int xi = -1; bool has_x = false;
        if (P.excl_val) { xv = P.excl_val[i]; xi = P.excl_idx[i]; has_x = xi >= 0; }
        const int64_t self_j = P.from_base + i - P.to_base;
        // start with the tile that holds the diagonal: in sorted real-world lists the best matches sit
        // near the row itself, so the k-th key is high from the first unit on (order does not affect results)
        int first_tau = tau_lo;
        if (P.self_match && self_j >= (int64_t)tau_lo * T && self_j < (int64_t)tau_hi * T) first_tau = (int)(self_j / T);

        // rows with <= 32 terms (the common case) keep their terms in registers across tiles
        int t_reg = 0; double v_reg = 0.0; int prev_end = 0;
        if (m <= 32 && lane < m) { t_reg = P.a_indices[a0 + lane]; v_reg = P.a_data[a0 + lane]; }

        for (int it = 0; it < ntau; ++it) {
            int tau = first_tau + it; if (tau >= tau_hi) tau -= ntau;
            bool any_post = false;
            for (int tb = 0; tb < m; tb += 32) {
                const int kk = tb + lane;
                int s = 0, len = 0; double v = 0.0;
                if (kk < m) {
                    int t;
                    if (m <= 32) { t = t_reg; v = v_reg; } else { t = P.a_indices[a0 + kk]; v = P.a_data[a0 + kk]; }
                    const int64_t c = (int64_t)t * P.n_tiles + tau;
                    const int e = P.seg[c + 1];
                    s = (m <= 32 && it > 0 && tau != tau_lo) ? prev_end : P.seg[c];
                    prev_end = e;
                    len = e - s;
                }
                TermSeg me; me.s = s; me.len = len; me.v = v;
                terms[lane] = me;
                unsigned live = __ballot_sync(FULL, len > 0);
                __syncwarp();
                any_post |= live != 0u;
                while (live) {                                   // ascending lane == ascending term
                    const int src = __ffs(live) - 1; live &= live - 1;
                    const TermSeg ts = terms[src];               // 16-byte broadcast
                    const uint16_t *pi = P.post_idx + ts.s;
                    const double *pv = P.post_val + ts.s;
                    for (int c0 = 0; c0 < ts.len; c0 += 128) {
                        int jl[4]; double wv[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int q = c0 + u * 32 + lane;
                            jl[u] = -1; wv[u] = 0.0;
                            if (q < ts.len) { jl[u] = pi[q]; wv[u] = pv[q]; }
                        }
                        bool cross = false; int cj = 0;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (jl[u] >= 0) {
                                const double old = acc[jl[u]];
                                const double nw = __dadd_rn(old, __dmul_rn(ts.v, wv[u]));
                                acc[jl[u]] = nw;
                                // first time the running sum passes the k-th key (ties: inclusive once the list is full)
                                const bool pn = (nw > kv) || (nw == kv && ki >= 0);
                                const bool po = (old > kv) || (old == kv && ki >= 0);
                                if (pn && !po) { atomicOr(&bitmap[jl[u] >> 5], 1u << (jl[u] & 31)); }
                            }
                        }
                        (void)cross; (void)cj;
                    }
                    __syncwarp();
                }
            }
            if (!any_post) continue;
            // candidates: bitmap words -> final sums -> exact key test -> insertion
            for (int w0 = 0; w0 < nwords; w0 += 32) {
                unsigned bits = 0u;
                if (w0 + lane < nwords) { bits = bitmap[w0 + lane]; if (bits) bitmap[w0 + lane] = 0u; }
                while (__ballot_sync(FULL, bits != 0u)) {
                    double sc = 0.0; int j = -1; bool cand = false;
                    if (bits) {
                        const int b = __ffs(bits) - 1; bits &= bits - 1;
                        const int jl = ((w0 + lane) << 5) + b;
                        sc = acc[jl];
                        const int jloc = tau * T + jl;
                        j = (int)(P.to_base + jloc);
                        cand = key_before(sc, j, kv, ki);
                        if (P.self_match && (int64_t)jloc == self_j) cand = false;
                        if (has_x && !key_before(xv, xi, sc, j)) cand = false;
                    }
                    unsigned cm = __ballot_sync(FULL, cand);
                    while (cm) {
                        const int src = __ffs(cm) - 1;
                        const double cs = shfl_d(sc, src);
                        const int cjx = __shfl_sync(FULL, j, src);
                        const bool stays = (lane < P.k) && key_before(tv, ti, cs, cjx);
                        const int pos = __popc(__ballot_sync(FULL, stays));
                        const double uv = __shfl_up_sync(FULL, tv, 1);
                        const int ui = __shfl_up_sync(FULL, ti, 1);
                        if (lane > pos) { tv = uv; ti = ui; }
                        else if (lane == pos) { tv = cs; ti = cjx; }
                        kv = shfl_d(tv, P.k - 1);
                        ki = __shfl_sync(FULL, ti, P.k - 1);
                        // prune: drop the inserted lane and everything the new k-th key now rejects
                        cand = cand && lane != src && key_before(sc, j, kv, ki);
                        cm = __ballot_sync(FULL, cand);
                    }
                }
            }
            __syncwarp();
            // dense clear, 16 B per lane per store
            {
                double2 *a2 = reinterpret_cast<double2 *>(acc);
                const double2 z = make_double2(0.0, 0.0);
                for (int q = lane; q < (T >> 1); q += 32) a2[q] = z;
            }
            __syncwarp();
        }
        if (lane < P.k) {
            const size_t o = ((size_t)split * P.n_from + i) * P.k + lane;
            P.top_idx[o] = ti;
            P.top_val[o] = (ti >= 0) ? tv : 0.0;
        }
    }
}

// merge: one warp per row; lists are individually sorted but that is not relied upon
__global__ void __launch_bounds__(256) topk_merge_kernel(const int32_t *__restrict__ idx, const double *__restrict__ val, int n_lists, int n_from,
                                                         int k_in, int k_out, int32_t *__restrict__ out_idx, double *__restrict__ out_val) {
    const int lane = lane_id();
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int i = gw; i < n_from; i += nw) {
        double tv = 0.0; int ti = -1;       // sentinel entries: idx -1 ranks after everything valid
        double kv = 0.0; int ki = -1; bool full = false;
        int filled = 0;
        for (int l = 0; l < n_lists; ++l) {
            const size_t b = ((size_t)l * n_from + i) * k_in;
            for (int c0 = 0; c0 < k_in; c0 += 32) {
                const int q = c0 + lane;
                double sc = 0.0; int j = -1;
                if (q < k_in) { j = idx[b + q]; sc = val[b + q]; }
                unsigned cm = __ballot_sync(FULL, j >= 0);
                while (cm) {
                    const int src = __ffs(cm) - 1; cm &= cm - 1;
                    const double cs = shfl_d(sc, src);
                    const int cj = __shfl_sync(FULL, j, src);
                    if (full && !key_before(cs, cj, kv, ki)) continue;
                    const bool stays = (lane < filled) && key_before(tv, ti, cs, cj);
                    const int pos = __popc(__ballot_sync(FULL, stays));
                    const double uv = __shfl_up_sync(FULL, tv, 1);
                    const int ui = __shfl_up_sync(FULL, ti, 1);
                    if (lane > pos) { tv = uv; ti = ui; }
                    else if (lane == pos) { tv = cs; ti = cj; }
                    if (filled < k_out) ++filled;
                    full = filled == k_out;
                    kv = shfl_d(tv, k_out - 1);
                    ki = __shfl_sync(FULL, ti, k_out - 1);
                }
            }
        }
        if (lane < k_out) {
            const bool ok = lane < filled;
            out_idx[(size_t)i * k_out + lane] = ok ? ti : -1;
            out_val[(size_t)i * k_out + lane] = ok ? tv : 0.0;
        }
    }
}

static int grid_for2(int64_t work_items, int threads, int cap) {
    int64_t g = (work_items + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

}  // namespace pfz

using namespace pfz;

extern "C" {

int pfz_index_build(const int32_t *indptr, const int32_t *indices, const double *data, int32_t n_rows, int32_t n_vocab, int32_t tile,
                    int32_t n_tiles, int32_t *seg, uint16_t *post_idx, double *post_val, void *ws, void *stream) {
    PFZ_REQUIRE(tile > 0 && tile <= 65536, "pfz_index_build: tile %d out of range (1..65536)", tile);
    PFZ_REQUIRE((int64_t)n_tiles * tile >= n_rows, "pfz_index_build: n_tiles*tile < n_rows");
    const int64_t ncell = (int64_t)n_vocab * n_tiles;
    PFZ_REQUIRE(ncell + 1 < (1ll << 31), "pfz_index_build: n_vocab*n_tiles = %lld too large", (long long)ncell);
    cudaStream_t st = as_stream(stream);
    // ws layout: cur int32[ncell+1] | scan ws
    int32_t *cur = reinterpret_cast<int32_t *>(ws);
    void *sws = reinterpret_cast<char *>(ws) + ((((size_t)ncell + 1) * 4 + 255) / 256) * 256;
    PFZ_CUDA_OK(cudaMemsetAsync(seg, 0, ((size_t)ncell + 1) * 4, st));
    if (n_rows > 0) {
        index_count_kernel<<<grid_for2((int64_t)n_rows * 32, 256, 148 * 16), 256, 0, st>>>(indptr, indices, n_rows, tile, n_tiles, seg);
        PFZ_LAUNCH_OK();
    }
    if (scan_exclusive_i32(seg, seg, ncell + 1, sws, st)) return 1;
    if (n_rows > 0) {
        PFZ_CUDA_OK(cudaMemsetAsync(cur, 0, ((size_t)ncell + 1) * 4, st));
        index_fill_kernel<<<grid_for2((int64_t)n_rows * 32, 256, 148 * 16), 256, 0, st>>>(indptr, indices, data, n_rows, tile, n_tiles, seg, cur,
                                                                                            post_idx, post_val);
        PFZ_LAUNCH_OK();
    }
    return 0;
}

int pfz_spcos_topk(const int32_t *a_indptr, const int32_t *a_indices, const double *a_data, int32_t n_from, const int32_t *seg,
                   const uint16_t *post_idx, const double *post_val, int32_t n_vocab, int32_t tile, int32_t n_tiles, int32_t n_to, int32_t k,
                   double min_similarity, int32_t self_match, int64_t from_index_base, int64_t to_index_base, int32_t n_splits,
                   const double *excl_val, const int32_t *excl_idx, int32_t *top_idx, double *top_val, int32_t *row_counter,
                   int32_t variant, void *stream) {
    PFZ_REQUIRE(k >= 1 && k <= 32, "pfz_spcos_topk: k=%d unsupported (1..32 per call; page with excl_* for more)", k);
    PFZ_REQUIRE(tile > 0 && tile <= 65536 && (tile % 64) == 0, "pfz_spcos_topk: tile %d must be a multiple of 64 in 64..65536", tile);
    PFZ_REQUIRE(n_splits >= 1 && n_splits <= n_tiles, "pfz_spcos_topk: n_splits %d out of range", n_splits);
    PFZ_REQUIRE(variant == PFZ_K2_LIST || variant == PFZ_K2_DENSE, "pfz_spcos_topk: unknown variant %d", variant);
    if (n_from <= 0) return 0;
    cudaStream_t st = as_stream(stream);
    int dev = 0, sms = 0, smem_max = 0;
    PFZ_CUDA_OK(cudaGetDevice(&dev));
    PFZ_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    PFZ_CUDA_OK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    PFZ_CUDA_OK(cudaMemsetAsync(row_counter, 0, sizeof(int32_t) * (size_t)n_splits, st));
    SpcosParams P{a_indptr, a_indices, a_data, n_from, seg, post_idx, post_val, n_vocab, tile, n_tiles, n_to, k, min_similarity, self_match,
                  from_index_base, to_index_base, n_splits, excl_val, excl_idx, top_idx, top_val, row_counter};
    auto launch = [&](auto kernel, int warps, size_t smem) -> int {
        PFZ_REQUIRE(smem <= (size_t)smem_max, "pfz_spcos_topk: tile %d needs %zu B shared memory > %d available", tile, smem, smem_max);
        PFZ_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        PFZ_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, warps * 32, smem));
        if (occ < 1) occ = 1;
        int gx = sms * occ;
        const int need = (n_from + warps - 1) / warps;
        if (gx > need) gx = need;
        if (n_splits > 1) { gx = (gx + n_splits - 1) / n_splits; if (gx < 1) gx = 1; }
        kernel<<<dim3(gx, n_splits), warps * 32, smem, st>>>(P);
        PFZ_LAUNCH_OK();
        return 0;
    };
    if (variant == PFZ_K2_LIST) {
        constexpr int WARPS = 8;
        return launch(spcos_list_kernel<WARPS>, WARPS, (size_t)WARPS * tile * 10);
    }
    constexpr int WARPS = 4;
    const size_t arena = (((size_t)tile * 8 + 32 * sizeof(TermSeg) + (size_t)(tile >> 5) * 4) + 15) & ~(size_t)15;
    return launch(spcos_dense_kernel<WARPS>, WARPS, (size_t)WARPS * arena);
}

int pfz_topk_merge(const int32_t *idx, const double *val, int32_t n_lists, int32_t n_from, int32_t k_in, int32_t k_out, int32_t *out_idx,
                   double *out_val, void *stream) {
    PFZ_REQUIRE(k_out >= 1 && k_out <= 32, "pfz_topk_merge: k_out=%d unsupported (1..32)", k_out);
    PFZ_REQUIRE(n_lists >= 1 && k_in >= 1, "pfz_topk_merge: bad n_lists/k_in");
    if (n_from <= 0) return 0;
    topk_merge_kernel<<<grid_for2((int64_t)n_from * 32, 256, 148 * 16), 256, 0, as_stream(stream)>>>(idx, val, n_lists, n_from, k_in, k_out, out_idx, out_val);
    PFZ_LAUNCH_OK();
    return 0;
}
}
