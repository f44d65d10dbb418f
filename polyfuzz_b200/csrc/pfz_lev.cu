// pfz_lev.cu -- K3: all-pairs edit distance (bit-parallel Myers/Hyyro Levenshtein and Hyyro LCS/Indel)
// over the |from| x |to| grid with a fused per-row arg-best.
//
// Replaces rapidfuzz's scorer loop as the reference calls it:
//     polyfuzz/models/_rapidfuzz.py:99-113  process.extractOne(q, to_list, score_cutoff, scorer=fuzz.ratio)
//     polyfuzz/models/_distance.py:89-102   [scorer(q, t) for t in to_list]; np.argmax
//
// Mapping: one warp = one from-string (the bit-vector "pattern", its match masks Peq in shared memory),
// one lane = one to-string (the "text") at a time.  To-strings are pre-sorted by length and stored in
// groups of 32, transposed and packed 4 symbols per 32-bit word, so the 32 lanes of a warp read one
// coalesced 128-byte line per 4 dynamic-programming columns and run nearly the same trip count.
// Symbols are bytes: the host maps the code points that occur in the from-strings to 1..255 and every
// other code point to 0 ("matches nothing") -- equality among text-only symbols never matters.
//
// Patterns of <= 32 symbols use 32-bit words (half the integer work), <= 64 one 64-bit word, longer ones
// NW 64-bit blocks with horizontal carries (Hyyro 2003).  Work is integer-ALU bound.
#include "pfz_common.cuh"

namespace pfz {

// ---- to-list layout ---------------------------------------------------------------------------
// sorted position p (by length) belongs to group p/32, lane p%32.
//   grp_word_off[g] : first 32-bit word of group g in `packed`; group g holds ceil(maxlen_g/4) x 32 words
//   slen[p], sorig[p]: length and original index of sorted string p
__global__ void __launch_bounds__(256) lev_pack_kernel(const uint32_t *__restrict__ blob, const int64_t *__restrict__ offsets,
                                                       const int32_t *__restrict__ order, int n_to, const uint8_t *__restrict__ sym_table,
                                                       const int64_t *__restrict__ grp_word_off, uint32_t *__restrict__ packed,
                                                       int32_t *__restrict__ slen) {
    const int lane = lane_id();
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int n_grp = (n_to + 31) >> 5;
    for (int g = gw; g < n_grp; g += nw) {
        const int p = g * 32 + lane;
        int64_t beg = 0; int len = 0;
        if (p < n_to) { const int o = order[p]; beg = offsets[o]; len = (int)(offsets[o + 1] - beg); slen[p] = len; }
        int mx = len;
#pragma unroll
        for (int d = 16; d; d >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, d));
        const int nwords = (mx + 3) >> 2;
        uint32_t *dst = packed + grp_word_off[g];
        for (int wi = 0; wi < nwords; ++wi) {
            uint32_t word = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int q = wi * 4 + b;
                if (q < len) {
                    const uint32_t c = blob[beg + q];
                    word |= (uint32_t)(c < 0x110000u ? sym_table[c] : 0) << (8 * b);
                }
            }
            dst[(size_t)wi * 32 + lane] = word;
        }
    }
}

struct LevParams {
    const uint32_t *from_blob; const int64_t *from_off; const int32_t *from_ids; int n_ids;     // patterns of this class
    const uint8_t *sym_table;
    const uint32_t *packed; const int64_t *grp_word_off; const int32_t *slen; const int32_t *sorig; int n_to;
    int metric; double cutoff; int exclude_self; int64_t self_shift;
    int n_splits;                       // to-groups are split over blockIdx.y
    int32_t *part_idx; double *part_score; int32_t *part_dist;     // [n_splits][n_from]
    int32_t *matrix; int64_t matrix_ld;                           // optional full matrix [n_from][n_to]
    int n_from; int32_t *counter;
};

__device__ __forceinline__ double score_of(int metric, int d, int la, int lb) {
    if (metric == PFZ_METRIC_NORM_LEV) { const int m = max(la, lb); return m ? 1.0 - (double)d / (double)m : 1.0; }
    if (metric == PFZ_METRIC_RATIO) { const int m = la + lb; return m ? (1.0 - (double)d / (double)m) * 100.0 : 100.0; }
    return -(double)d;                  // raw distances: best = smallest
}

template <typename W> struct WordOps;
template <> struct WordOps<uint32_t> { static constexpr int BITS = 32; static __device__ __forceinline__ int pop(uint32_t x) { return __popc(x); } };
template <> struct WordOps<uint64_t> { static constexpr int BITS = 64; static __device__ __forceinline__ int pop(uint64_t x) { return __popcll(x); } };

// One warp scores pattern `pat` against every to-string of its split.  LCS = false: Levenshtein (Myers 1999,
// blocks: Hyyro 2003); LCS = true: longest common subsequence (Hyyro 2004) -> Indel = la + lb - 2*LCS.
template <typename W, int NW, bool LCS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) lev_kernel(const LevParams P) {
    constexpr int B = WordOps<W>::BITS;
    extern __shared__ __align__(16) unsigned char dyn[];
    const int lane = lane_id();
    const int w = threadIdx.x >> 5;
    W *peq = reinterpret_cast<W *>(dyn) + (size_t)w * 256 * NW;              // peq[sym * NW + block]
    const int split = blockIdx.y;
    const int n_grp = (P.n_to + 31) >> 5;
    const int per = (n_grp + P.n_splits - 1) / P.n_splits;
    const int g_lo = split * per, g_hi = min(n_grp, g_lo + per);
    int32_t *counter = P.counter + split;

    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(counter, 1);
        q = __shfl_sync(FULL, q, 0);
        if (q >= P.n_ids) break;
        const int i = P.from_ids[q];
        const int64_t fb = P.from_off[i];
        const int m = (int)(P.from_off[i + 1] - fb);
        // match masks
        for (int e = lane; e < 256 * NW; e += 32) peq[e] = 0;
        __syncwarp();
        for (int p = lane; p < m; p += 32) {
            const uint32_t c = P.from_blob[fb + p];
            const int s = c < 0x110000u ? P.sym_table[c] : 0;
            if (s) {
                if (sizeof(W) == 8) atomicOr(reinterpret_cast<unsigned long long *>(&peq[s * NW + p / B]), 1ull << (p % B));
                else atomicOr(reinterpret_cast<unsigned *>(&peq[s * NW + p / B]), 1u << (p % B));
            }
        }
        __syncwarp();
        const int last_bit = (m - 1) & (B - 1);          // bit of row m inside the last block (m > 0)
        const int last_blk = m > 0 ? (m - 1) / B : 0;

        double best_s = 0.0; int best_j = -1, best_d = -1;
        for (int g = g_lo; g < g_hi; ++g) {
            const int p = g * 32 + lane;
            const bool have = p < P.n_to;
            const int n = have ? P.slen[p] : 0;
            const int orig = have ? P.sorig[p] : -1;
            int nmax = n;
#pragma unroll
            for (int d = 16; d; d >>= 1) nmax = max(nmax, __shfl_xor_sync(FULL, nmax, d));
            const uint32_t *src = P.packed + P.grp_word_off[g] + lane;
            int dist;
            if (!LCS) {
                W Pv[NW], Mv[NW];
#pragma unroll
                for (int b = 0; b < NW; ++b) { Pv[b] = ~(W)0; Mv[b] = 0; }
                int score = m;
                uint32_t nextw = nmax > 0 ? src[0] : 0u;
                for (int j0 = 0; j0 < nmax; j0 += 4) {
                    const uint32_t word = nextw;
                    if (j0 + 4 < nmax) nextw = src[(size_t)((j0 >> 2) + 1) * 32];      // prefetch: the recurrence below is a long dependent chain
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb) {
                        if (j0 + bb < n) {
                            const int s = (word >> (8 * bb)) & 0xff;
                            int hin = 1;                              // D[0][j] - D[0][j-1] = +1
#pragma unroll
                            for (int b = 0; b < NW; ++b) {
                                if (b <= last_blk) {
                                    W Eq = peq[s * NW + b];
                                    const W pv = Pv[b], mv = Mv[b];
                                    const W Xv = Eq | mv;
                                    if (hin < 0) Eq |= 1;
                                    const W Xh = (((Eq & pv) + pv) ^ pv) | Eq;
                                    W Ph = mv | ~(Xh | pv);
                                    W Mh = pv & Xh;
                                    const int top = (b == last_blk) ? last_bit : B - 1;
                                    const int hout = (int)((Ph >> top) & 1) - (int)((Mh >> top) & 1);
                                    Ph <<= 1; Mh <<= 1;
                                    if (hin < 0) Mh |= 1; else if (hin > 0) Ph |= 1;
                                    Pv[b] = Mh | ~(Xv | Ph);
                                    Mv[b] = Ph & Xv;
                                    hin = hout;
                                }
                            }
                            score += hin;                            // horizontal delta of row m
                        }
                    }
                }
                dist = m > 0 ? score : n;
            } else {
                W S[NW];
#pragma unroll
                for (int b = 0; b < NW; ++b) S[b] = ~(W)0;
                uint32_t nextw = nmax > 0 ? src[0] : 0u;
                for (int j0 = 0; j0 < nmax; j0 += 4) {
                    const uint32_t word = nextw;
                    if (j0 + 4 < nmax) nextw = src[(size_t)((j0 >> 2) + 1) * 32];
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb) {
                        if (j0 + bb < n) {
                            const int s = (word >> (8 * bb)) & 0xff;
                            unsigned carry = 0;
#pragma unroll
                            for (int b = 0; b < NW; ++b) {
                                if (b <= last_blk) {
                                    const W Eq = peq[s * NW + b];
                                    const W x = S[b], u = x & Eq;
                                    // S' = (S + (S & Eq)) | (S - (S & Eq)); u is a subset of x, so x - u = x & ~Eq (no borrow);
                                    // the addition carries across blocks
                                    const W sum = x + u; unsigned c1 = sum < x; const W sum2 = sum + carry; c1 |= (sum2 < sum); carry = c1;
                                    S[b] = sum2 | (x & ~Eq);
                                }
                            }
                        }
                    }
                }
                int lcs = 0;
#pragma unroll
                for (int b = 0; b < NW; ++b) {
                    if (m > 0 && b <= last_blk) {
                        W z = ~S[b];
                        if (b == last_blk && last_bit != B - 1) z &= (((W)1 << (last_bit + 1)) - 1);
                        lcs += WordOps<W>::pop(z);
                    }
                }
                dist = m + n - 2 * lcs;
            }
            if (have) {
                if (P.matrix) P.matrix[(int64_t)i * P.matrix_ld + orig] = dist;
                const double sc = score_of(P.metric, dist, m, n);
                bool ok = !(P.exclude_self && (int64_t)orig == (int64_t)i + P.self_shift);
                if ((P.metric == PFZ_METRIC_NORM_LEV || P.metric == PFZ_METRIC_RATIO) && !(sc >= P.cutoff)) ok = false;
                if (ok && (best_j < 0 || sc > best_s || (sc == best_s && orig < best_j))) { best_s = sc; best_j = orig; best_d = dist; }
            }
        }
        // first maximal score = lowest original index among the maxima
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            const double os = shfl_d(best_s, lane ^ d);
            const int oj = __shfl_xor_sync(FULL, best_j, d), od = __shfl_xor_sync(FULL, best_d, d);
            if (oj >= 0 && (best_j < 0 || os > best_s || (os == best_s && oj < best_j))) { best_s = os; best_j = oj; best_d = od; }
        }
        if (lane == 0) {
            const size_t o = (size_t)split * P.n_from + i;
            P.part_idx[o] = best_j; P.part_score[o] = best_j >= 0 ? best_s : 0.0; P.part_dist[o] = best_d;
        }
        __syncwarp();
    }
}

__global__ void lev_merge_kernel(const int32_t *__restrict__ part_idx, const double *__restrict__ part_score, const int32_t *__restrict__ part_dist,
                                 int n_splits, int n_from, int32_t *__restrict__ best_idx, double *__restrict__ best_score,
                                 int32_t *__restrict__ best_dist) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_from; i += gridDim.x * blockDim.x) {
        double bs = 0.0; int bj = -1, bd = -1;
        for (int s = 0; s < n_splits; ++s) {
            const size_t o = (size_t)s * n_from + i;
            const int j = part_idx[o];
            if (j < 0) continue;
            const double sc = part_score[o];
            if (bj < 0 || sc > bs || (sc == bs && j < bj)) { bs = sc; bj = j; bd = part_dist[o]; }
        }
        best_idx[i] = bj; best_score[i] = bj >= 0 ? bs : 0.0; best_dist[i] = bd;
    }
}

template <typename W, int NW, bool LCS>
static int launch_lev(const LevParams &P, int sms, cudaStream_t st) {
    constexpr int WARPS = 4;
    const size_t smem = (size_t)WARPS * 256 * NW * sizeof(W);
    auto kernel = lev_kernel<W, NW, LCS, WARPS>;
    PFZ_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PFZ_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, WARPS * 32, smem));
    if (occ < 1) occ = 1;
    int gx = sms * occ;
    const int need = (P.n_ids + WARPS - 1) / WARPS;
    if (gx > need) gx = need;
    if (gx < 1) gx = 1;
    kernel<<<dim3(gx, P.n_splits), WARPS * 32, smem, st>>>(P);
    PFZ_LAUNCH_OK();
    return 0;
}

}  // namespace pfz

using namespace pfz;

extern "C" {

int pfz_lev_pack(const uint32_t *to_blob, const int64_t *to_offsets, const int32_t *order, int32_t n_to, const uint8_t *sym_table,
                 const int64_t *grp_word_off, uint32_t *packed, int32_t *slen, void *stream) {
    if (n_to <= 0) return 0;
    const int n_grp = (n_to + 31) / 32;
    int grid = (n_grp + 7) / 8; if (grid > 148 * 8) grid = 148 * 8;
    lev_pack_kernel<<<grid, 256, 0, as_stream(stream)>>>(to_blob, to_offsets, order, n_to, sym_table, grp_word_off, packed, slen);
    PFZ_LAUNCH_OK();
    return 0;
}

int pfz_lev_argbest(const uint32_t *from_blob, const int64_t *from_offsets, int32_t n_from, const int32_t *from_ids, int32_t n_ids,
                    int32_t n_words, const uint8_t *sym_table, const uint32_t *packed, const int64_t *grp_word_off, const int32_t *slen,
                    const int32_t *sorig, int32_t n_to, int32_t metric, double score_cutoff, int32_t exclude_self, int64_t self_shift,
                    int32_t n_splits, int32_t *part_idx, double *part_score, int32_t *part_dist, int32_t *matrix, int64_t matrix_ld,
                    int32_t *counter, void *stream) {
    PFZ_REQUIRE(metric >= PFZ_METRIC_LEV && metric <= PFZ_METRIC_RATIO, "pfz_lev_argbest: unknown metric %d", metric);
    PFZ_REQUIRE(n_words == 0 || n_words == 1 || n_words == 2 || n_words == 4 || n_words == 8 || n_words == 16,
                "pfz_lev_argbest: n_words %d unsupported (0 = 32-bit word, 1, 2, 4, 8, 16 64-bit words)", n_words);
    PFZ_REQUIRE(n_splits >= 1, "pfz_lev_argbest: n_splits < 1");
    if (n_ids <= 0 || n_to < 0) return 0;
    cudaStream_t st = as_stream(stream);
    int dev = 0, sms = 0;
    PFZ_CUDA_OK(cudaGetDevice(&dev));
    PFZ_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    PFZ_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(int32_t) * (size_t)n_splits, st));
    LevParams P{from_blob, from_offsets, from_ids, n_ids, sym_table, packed, grp_word_off, slen, sorig, n_to, metric, score_cutoff,
                exclude_self, self_shift, n_splits, part_idx, part_score, part_dist, matrix, matrix_ld, n_from, counter};
    const bool lcs = (metric == PFZ_METRIC_INDEL || metric == PFZ_METRIC_RATIO);
#define PFZ_LEV_CASE(NWv, Wt)                                                            \
    return lcs ? launch_lev<Wt, NWv, true>(P, sms, st) : launch_lev<Wt, NWv, false>(P, sms, st)
    switch (n_words) {
        case 0: PFZ_LEV_CASE(1, uint32_t);
        case 1: PFZ_LEV_CASE(1, uint64_t);
        case 2: PFZ_LEV_CASE(2, uint64_t);
        case 4: PFZ_LEV_CASE(4, uint64_t);
        case 8: PFZ_LEV_CASE(8, uint64_t);
        default: PFZ_LEV_CASE(16, uint64_t);
    }
#undef PFZ_LEV_CASE
}

int pfz_lev_merge(const int32_t *part_idx, const double *part_score, const int32_t *part_dist, int32_t n_splits, int32_t n_from,
                  int32_t *best_idx, double *best_score, int32_t *best_dist, void *stream) {
    if (n_from <= 0) return 0;
    int grid = (n_from + 255) / 256; if (grid > 148 * 8) grid = 148 * 8;
    lev_merge_kernel<<<grid, 256, 0, as_stream(stream)>>>(part_idx, part_score, part_dist, n_splits, n_from, best_idx, best_score, best_dist);
    PFZ_LAUNCH_OK();
    return 0;
}
}
