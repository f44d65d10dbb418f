// pfz_fuzz.cu -- K3b: rapidfuzz's token / partial / weighted scorers over the |from| x |to| grid with a fused per-row
// arg-best (process.extractOne semantics: the first to-string with the maximal score >= score_cutoff).
//
// Replaces the scorer loop of the reference's edit-distance matchers for the scorers beyond fuzz.ratio:
//     polyfuzz/models/_rapidfuzz.py:48       scorer=fuzz.WRatio, the default of RapidFuzz and of PolyFuzz("EditDistance")
//     polyfuzz/models/_rapidfuzz.py:24-37    partial_ratio, token_sort_ratio, token_set_ratio, token_ratio, partial_token_*, QRatio
//     polyfuzz/models/_rapidfuzz.py:106-108  process.extractOne(query, to_list, scorer=..., score_cutoff=...)
// The scorer definitions (incl. the order of the double-precision operations and the way score_cutoff is threaded through
// WRatio) are those of rapidfuzz 3.x as restated and pinned on rapidfuzz's published known answers in oracle/fuzz.py.
//
// Mapping (as K3, pfz_lev.cu): one warp = one from-string, whose bit-vector match masks Peq[symbol] live in shared memory
// for THREE derived patterns -- a itself, S(a) = its whitespace tokens sorted and joined, U(a) = its distinct tokens sorted and
// joined; one lane = one to-string at a time (to-strings pre-sorted by length, groups of 32, transposed, 4 byte-symbols per
// word), with the same three variants b, S(b), U(b).  Every Indel distance is a bit-parallel LCS (Hyyro 2004):
//   ratio                LCS(a, b)
//   token_sort_ratio     LCS(S(a), S(b))
//   token_set_ratio      token-id sets (sorted ids per string, 64-bit Bloom signature as a pre-filter): no common token ->
//                        LCS(U(a), U(b)); a common token and one side a subset -> 100; else the differences are joined per
//                        pair and scored by a small dynamic programme in local memory
//   partial_ratio        the shorter string against every window of the longer one (prefixes shorter than it, all windows of
//                        its length, suffixes): windows of the to-string restart the recurrence at the window start; windows of
//                        the from-string mask Peq to the window's bits; prefix windows fall out of ONE pass (the number of zero
//                        bits among the first i rows is LCS(a[:i], b))
//   WRatio               the weighted combination (0.95 / 0.9 / 0.6, length-ratio switches 1.5 and 8) of the above.
// Symbols are bytes (the host maps the code points of the from-list to 1..255, everything else to 0 = matches nothing).
#include "pfz_common.cuh"

namespace pfz {

constexpr int FZ_MAXLEN = 255;          // code points per string for these scorers (the host checks)
enum { FZ_RATIO = 0, FZ_QRATIO = 1, FZ_PARTIAL = 2, FZ_TSORT = 3, FZ_TSET = 4, FZ_TRATIO = 5, FZ_PTSORT = 6, FZ_PTSET = 7, FZ_PTRATIO = 8, FZ_WRATIO = 9 };

struct FuzzSide {                       // one string list with its derived variants (device pointers)
    const uint32_t *blob[3]; const int64_t *off[3];         // UTF-32 code points + offsets of s, S(s), U(s)
    const int32_t *tok_ptr; const int32_t *tok_ids;         // distinct token ids per string, ascending
    const uint64_t *sig;                                    // Bloom signature of the token ids
    const int32_t *n_tok_all;                               // number of tokens incl. duplicates
};
struct FuzzParams {
    FuzzSide F;                                             // from-strings (patterns)
    FuzzSide T;                                             // to-strings: blobs used by the per-pair slow path only
    const int32_t *from_ids; int n_ids;                     // from-rows of this word class
    const uint8_t *sym_table;
    const uint32_t *packed[3]; const int64_t *grp_off[3]; const int32_t *slen[3];   // to-side transposed layouts of b, S(b), U(b)
    const int32_t *sorig; int n_to;
    const uint32_t *tok_blob; const int64_t *tok_off;       // token texts (code points) by token id
    int scorer; double cutoff; int exclude_self; int64_t self_shift;
    int n_splits; int32_t *part_idx; double *part_score; int n_from; int32_t *counter;
};

// ---- scalar scorer algebra (double, no contraction) -- mirrors oracle/fuzz.py line by line -------------------------------------
__device__ __forceinline__ double fz_norm_sim(int dist, int lensum) {          // Indel.normalized_similarity
    const double nd = lensum ? __ddiv_rn((double)dist, (double)lensum) : 0.0;
    return __dsub_rn(1.0, nd);
}
__device__ __forceinline__ double fz_ratio(int lcs, int la, int lb, double cutoff) {
    const double ns = fz_norm_sim(la + lb - 2 * lcs, la + lb);
    return ns >= __ddiv_rn(cutoff, 100.0) ? __dmul_rn(ns, 100.0) : 0.0;
}
__device__ __forceinline__ double fz_norm_distance(int dist, int lensum, double cutoff) {
    const double sc = lensum ? __dsub_rn(100.0, __ddiv_rn(__dmul_rn(100.0, (double)dist), (double)lensum)) : 100.0;
    return sc >= cutoff ? sc : 0.0;
}
__device__ __forceinline__ double fz_partial_from_best(double best_ns, bool both_empty, double cutoff) {
    if (both_empty) return 100.0;
    const double res = __dmul_rn(best_ns, 100.0);
    return res >= cutoff ? res : 0.0;
}

// ---- bit-parallel LCS over NW 64-bit blocks ------------------------------------------------------------------------------------
template <int NW>
struct Bp {
    uint64_t S[NW];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int b = 0; b < NW; ++b) S[b] = ~0ull;
    }
    // one text symbol; mask (may be null) restricts the pattern to a window of its rows
    __device__ __forceinline__ void step(const uint64_t *__restrict__ peq, int s, const uint64_t *mask) {
        unsigned carry = 0;
#pragma unroll
        for (int b = 0; b < NW; ++b) {
            uint64_t Eq = peq[s * NW + b];
            if (mask) Eq &= mask[b];
            const uint64_t x = S[b], u = x & Eq;
            const uint64_t sum = x + u; unsigned c1 = sum < x; const uint64_t sum2 = sum + carry; c1 |= (sum2 < sum); carry = c1;
            S[b] = sum2 | (x & ~Eq);
        }
    }
    // LCS(pattern[lo:hi), text so far): zero bits among rows lo..hi-1
    __device__ __forceinline__ int zeros(int lo, int hi) const {
        int z = 0;
#pragma unroll
        for (int b = 0; b < NW; ++b) {
            const int l = max(lo - 64 * b, 0), h = min(hi - 64 * b, 64);
            if (h > l) {
                uint64_t m = (h == 64 ? ~0ull : ((1ull << h) - 1ull)) & ~((1ull << l) - 1ull);
                z += __popcll(~S[b] & m);
            }
        }
        return z;
    }
};

struct TextRef { const uint32_t *src; int n; };           // lane-strided packed words of one to-string variant
__device__ __forceinline__ int text_sym(const TextRef &t, int j) { return (t.src[(size_t)(j >> 2) * 32] >> (8 * (j & 3))) & 0xff; }

template <int NW>
__device__ int fz_lcs(const uint64_t *__restrict__ peq, int m, const TextRef &t) {
    if (m == 0 || t.n == 0) return 0;
    Bp<NW> bp; bp.init();
    for (int j = 0; j < t.n; ++j) bp.step(peq, text_sym(t, j), nullptr);
    return bp.zeros(0, m);
}

// best normalised Indel similarity of partial_ratio(pattern, text): rapidfuzz's window set (see the file header)
template <int NW>
__device__ double fz_partial_best(const uint64_t *__restrict__ peq, int la, const TextRef &t) {
    const int lb = t.n;
    double best = 0.0;
    if (la == 0 || lb == 0) return 0.0;                    // (both empty is handled by the caller)
    if (la <= lb) {                                        // shorter = pattern; windows of the text
        const int ls = la, ll = lb;
        Bp<NW> bp; bp.init();
        for (int j = 0; j + 1 < ls; ++j) {                 // prefixes text[:j+1], one pass
            bp.step(peq, text_sym(t, j), nullptr);
            best = fmax(best, fz_norm_sim(ls + (j + 1) - 2 * bp.zeros(0, ls), ls + j + 1));
        }
        for (int i = 0; i < ll; ++i) {                     // windows text[i : i+ls] (i < ll-ls) and suffixes text[i:] (i >= ll-ls)
            const int e = min(ll, i + ls);
            bp.init();
            for (int j = i; j < e; ++j) bp.step(peq, text_sym(t, j), nullptr);
            best = fmax(best, fz_norm_sim(ls + (e - i) - 2 * bp.zeros(0, ls), ls + (e - i)));
        }
    }
    if (la > lb || (la == lb && best != 1.0)) {            // shorter = text; windows of the pattern
        const int ls = lb, ll = la;
        Bp<NW> bp; bp.init();
        for (int j = 0; j < ls; ++j) bp.step(peq, text_sym(t, j), nullptr);
        for (int i = 1; i < ls; ++i)                       // prefixes pattern[:i]: rows 0..i-1 of the final state
            best = fmax(best, fz_norm_sim(ls + i - 2 * bp.zeros(0, i), ls + i));
        for (int i = 0; i < ll; ++i) {                     // windows pattern[i : i+ls] and suffixes pattern[i:]
            const int e = min(ll, i + ls);
            uint64_t mask[NW];
#pragma unroll
            for (int b = 0; b < NW; ++b) {
                const int l = max(i - 64 * b, 0), h = min(e - 64 * b, 64);
                mask[b] = (h > l) ? ((h == 64 ? ~0ull : ((1ull << h) - 1ull)) & ~((1ull << l) - 1ull)) : 0ull;
            }
            bp.init();
            for (int j = 0; j < ls; ++j) bp.step(peq, text_sym(t, j), mask);
            best = fmax(best, fz_norm_sim(ls + (e - i) - 2 * bp.zeros(i, e), ls + (e - i)));
        }
    }
    return best;
}

// ---- token sets ----------------------------------------------------------------------------------------------------------------
struct TokInfo { int n_common; int sect_len; int ab_len; int ba_len; int n_ab; int n_ba; };
// merge of the two ascending id lists: |intersection| and the joined lengths of intersection / differences (tokens + single spaces)
__device__ TokInfo fz_tok_info(const int32_t *a, int na, const int32_t *b, int nb, const int64_t *__restrict__ tok_off) {
    TokInfo r{0, 0, 0, 0, 0, 0};
    int p = 0, q = 0, sect = 0, ab = 0, ba = 0;
    while (p < na || q < nb) {
        const int x = p < na ? a[p] : 0x7fffffff, y = q < nb ? b[q] : 0x7fffffff;
        if (x == y) { sect += (int)(tok_off[x + 1] - tok_off[x]); ++r.n_common; ++p; ++q; }
        else if (x < y) { ab += (int)(tok_off[x + 1] - tok_off[x]); ++r.n_ab; ++p; }
        else { ba += (int)(tok_off[y + 1] - tok_off[y]); ++r.n_ba; ++q; }
    }
    r.sect_len = sect + max(r.n_common - 1, 0);
    r.ab_len = ab + max(r.n_ab - 1, 0);
    r.ba_len = ba + max(r.n_ba - 1, 0);
    return r;
}
// Indel distance of the joined differences (tokens of a not in b, sorted | tokens of b not in a, sorted): textbook LCS rows in
// local memory; token ids ascend in the lists but the JOIN order is the lexicographic order of the token texts, which is the id
// order by construction (the host numbers the tokens in sorted order)
__device__ int fz_diff_indel(const int32_t *a, int na, const int32_t *b, int nb, const uint32_t *__restrict__ tok_blob,
                             const int64_t *__restrict__ tok_off) {
    uint32_t sa[FZ_MAXLEN + 1], sb[FZ_MAXLEN + 1];
    uint8_t row[FZ_MAXLEN + 2];
    int la = 0, lb = 0;
    {
        int p = 0, q = 0;
        while (p < na || q < nb) {
            const int x = p < na ? a[p] : 0x7fffffff, y = q < nb ? b[q] : 0x7fffffff;
            if (x == y) { ++p; ++q; }
            else if (x < y) {
                if (la) sa[la++] = 0x20u;
                for (int64_t c = tok_off[x]; c < tok_off[x + 1] && la <= FZ_MAXLEN; ++c) sa[la++] = tok_blob[c];
                ++p;
            } else {
                if (lb) sb[lb++] = 0x20u;
                for (int64_t c = tok_off[y]; c < tok_off[y + 1] && lb <= FZ_MAXLEN; ++c) sb[lb++] = tok_blob[c];
                ++q;
            }
        }
    }
    for (int j = 0; j <= lb; ++j) row[j] = 0;
    for (int i = 1; i <= la; ++i) {
        int diag = 0;
        const uint32_t ca = sa[i - 1];
        for (int j = 1; j <= lb; ++j) {
            const int up = row[j];
            const int bestv = (ca == sb[j - 1]) ? diag + 1 : max(up, (int)row[j - 1]);
            diag = up;
            row[j] = (uint8_t)bestv;
        }
    }
    return la + lb - 2 * (int)row[lb];
}

template <int NW, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) fuzz_kernel(const FuzzParams P) {
    extern __shared__ __align__(16) unsigned char dyn[];
    const int lane = lane_id();
    const int w = threadIdx.x >> 5;
    uint64_t *peq_all = reinterpret_cast<uint64_t *>(dyn) + (size_t)w * 3 * 256 * NW;      // peq[variant][sym * NW + block]
    const int split = blockIdx.y;
    const int n_grp = (P.n_to + 31) >> 5;
    const int per = (n_grp + P.n_splits - 1) / P.n_splits;
    const int g_lo = split * per, g_hi = min(n_grp, g_lo + per);
    int32_t *counter = P.counter + split;
    const int sc_id = P.scorer;
    const bool need_sorted = sc_id == FZ_TSORT || sc_id == FZ_TRATIO || sc_id == FZ_PTSORT || sc_id == FZ_PTRATIO || sc_id == FZ_WRATIO;
    const bool need_uniq = sc_id == FZ_TSET || sc_id == FZ_TRATIO || sc_id == FZ_PTSET || sc_id == FZ_PTRATIO || sc_id == FZ_WRATIO;

    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(counter, 1);
        q = __shfl_sync(FULL, q, 0);
        if (q >= P.n_ids) break;
        const int i = P.from_ids[q];
        int lens[3];
        for (int v = 0; v < 3; ++v) {
            const int64_t fb = P.F.off[v][i];
            const int m = (int)(P.F.off[v][i + 1] - fb);
            lens[v] = m;
            uint64_t *peq = peq_all + (size_t)v * 256 * NW;
            if ((v == 1 && !need_sorted) || (v == 2 && !need_uniq)) continue;
            for (int e = lane; e < 256 * NW; e += 32) peq[e] = 0;
            __syncwarp();
            for (int p = lane; p < m; p += 32) {
                const uint32_t c = P.F.blob[v][fb + p];
                const int s = c < 0x110000u ? P.sym_table[c] : 0;
                if (s) atomicOr(reinterpret_cast<unsigned long long *>(&peq[s * NW + p / 64]), 1ull << (p % 64));
            }
            __syncwarp();
        }
        const int la = lens[0], las = lens[1], lau = lens[2];
        const int32_t *atok = P.F.tok_ids + P.F.tok_ptr[i];
        const int na = P.F.tok_ptr[i + 1] - P.F.tok_ptr[i];
        const int na_all = P.F.n_tok_all[i];
        const uint64_t asig = P.F.sig[i];
        const uint64_t *peq0 = peq_all, *peq1 = peq_all + 256 * NW, *peq2 = peq_all + 2 * 256 * NW;

        double best_s = 0.0; int best_j = -1;
        for (int g = g_lo; g < g_hi; ++g) {
            const int p = g * 32 + lane;
            if (p >= P.n_to) continue;
            const int orig = P.sorig[p];
            if (P.exclude_self && (int64_t)orig == (int64_t)i + P.self_shift) continue;
            TextRef t0{P.packed[0] + P.grp_off[0][g] + lane, P.slen[0][p]};
            TextRef t1{P.packed[1] + P.grp_off[1][g] + lane, P.slen[1][p]};
            TextRef t2{P.packed[2] + P.grp_off[2][g] + lane, P.slen[2][p]};
            const int lb = t0.n;
            const double cut = P.cutoff;
            // token-set facts of the pair, computed on demand
            bool tok_done = false; TokInfo ti{0, 0, 0, 0, 0, 0};
            const int32_t *btok = P.T.tok_ids + P.T.tok_ptr[orig];
            const int nb = P.T.tok_ptr[orig + 1] - P.T.tok_ptr[orig];
            const int nb_all = P.T.n_tok_all[orig];
            auto tok = [&]() {
                if (!tok_done) {
                    if ((asig & P.T.sig[orig]) == 0ull) {   // no common token possible: differences are the whole distinct-token strings
                        ti.n_common = 0; ti.sect_len = 0; ti.ab_len = lau; ti.ba_len = t2.n; ti.n_ab = na; ti.n_ba = nb;
                    } else ti = fz_tok_info(atok, na, btok, nb, P.tok_off);
                    tok_done = true;
                }
            };
            auto token_sort = [&](double c) { return fz_ratio(fz_lcs<NW>(peq1, las, t1), las, t1.n, c); };
            auto token_set = [&](double c) -> double {
                if (c > 100.0) return 0.0;
                if (na == 0 || nb == 0) return 0.0;
                tok();
                if (ti.n_common && (ti.n_ab == 0 || ti.n_ba == 0)) return 100.0;
                const int sect_len = ti.sect_len;
                const int sect_ab_len = sect_len + (sect_len != 0) + ti.ab_len;
                const int sect_ba_len = sect_len + (sect_len != 0) + ti.ba_len;
                double result = 0.0;
                const double cd = ceil(__dmul_rn((double)(sect_ab_len + sect_ba_len), __dsub_rn(1.0, __ddiv_rn(c, 100.0))));
                const int dist = ti.n_common == 0 ? (lau + t2.n - 2 * fz_lcs<NW>(peq2, lau, t2))
                                                  : fz_diff_indel(atok, na, btok, nb, P.tok_blob, P.tok_off);
                if ((double)dist <= cd) result = fz_norm_distance(dist, sect_ab_len + sect_ba_len, c);
                if (!sect_len) return result;
                const double r_ab = fz_norm_distance((sect_len != 0) + ti.ab_len, sect_len + sect_ab_len, c);
                const double r_ba = fz_norm_distance((sect_len != 0) + ti.ba_len, sect_len + sect_ba_len, c);
                return fmax(result, fmax(r_ab, r_ba));
            };
            auto partial = [&](const uint64_t *peq, int m, const TextRef &t, double c) {
                return fz_partial_from_best(fz_partial_best<NW>(peq, m, t), m == 0 && t.n == 0, c);
            };
            auto partial_token_ratio = [&](double c) -> double {
                tok();
                if (ti.n_common) return 100.0;
                const double result = partial(peq1, las, t1, c);
                if (na_all == na && nb_all == nb) return result;
                c = fmax(c, result);
                return fmax(result, partial(peq2, lau, t2, c));
            };

            double sc = 0.0;
            switch (sc_id) {
                case FZ_RATIO: sc = fz_ratio(fz_lcs<NW>(peq0, la, t0), la, lb, cut); break;
                case FZ_QRATIO: sc = (la == 0 || lb == 0) ? 0.0 : fz_ratio(fz_lcs<NW>(peq0, la, t0), la, lb, cut); break;
                case FZ_PARTIAL: sc = partial(peq0, la, t0, cut); break;
                case FZ_TSORT: sc = token_sort(cut); break;
                case FZ_TSET: sc = token_set(cut); break;
                case FZ_TRATIO: sc = fmax(token_set(cut), token_sort(cut)); break;
                case FZ_PTSORT: sc = partial(peq1, las, t1, cut); break;
                case FZ_PTSET: {
                    if (na == 0 || nb == 0) { sc = 0.0; break; }
                    tok();
                    sc = ti.n_common ? 100.0 : partial(peq2, lau, t2, cut);
                    break;
                }
                case FZ_PTRATIO: sc = partial_token_ratio(cut); break;
                default: {                                  // WRatio
                    if (la == 0 || lb == 0) { sc = 0.0; break; }
                    const double len_ratio = la > lb ? __ddiv_rn((double)la, (double)lb) : __ddiv_rn((double)lb, (double)la);
                    double end_ratio = fz_ratio(fz_lcs<NW>(peq0, la, t0), la, lb, cut);
                    double c = cut;
                    if (len_ratio < 1.5) {
                        c = __ddiv_rn(fmax(c, end_ratio), 0.95);
                        sc = fmax(end_ratio, __dmul_rn(fmax(token_set(c), token_sort(c)), 0.95));
                        break;
                    }
                    const double ps = len_ratio <= 8.0 ? 0.9 : 0.6;
                    c = __ddiv_rn(fmax(c, end_ratio), ps);
                    end_ratio = fmax(end_ratio, __dmul_rn(partial(peq0, la, t0, c), ps));
                    c = __ddiv_rn(fmax(c, end_ratio), 0.95);
                    sc = fmax(end_ratio, __dmul_rn(__dmul_rn(partial_token_ratio(c), 0.95), ps));
                }
            }
            if (sc >= P.cutoff && (best_j < 0 || sc > best_s || (sc == best_s && orig < best_j))) { best_s = sc; best_j = orig; }
        }
        // first maximal score = lowest original index among the maxima
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            const double os = shfl_d(best_s, lane ^ d);
            const int oj = __shfl_xor_sync(FULL, best_j, d);
            if (oj >= 0 && (best_j < 0 || os > best_s || (os == best_s && oj < best_j))) { best_s = os; best_j = oj; }
        }
        if (lane == 0) {
            const size_t o = (size_t)split * P.n_from + i;
            P.part_idx[o] = best_j; P.part_score[o] = best_j >= 0 ? best_s : 0.0;
        }
        __syncwarp();
    }
}

template <int NW>
static int launch_fuzz(const FuzzParams &P, int sms, cudaStream_t st) {
    constexpr int WARPS = NW == 1 ? 4 : NW == 2 ? 2 : 1;
    const size_t smem = (size_t)WARPS * 3 * 256 * NW * 8;
    auto kernel = fuzz_kernel<NW, WARPS>;
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

/* ptrs: 38 device pointers in the order of PfzFuzzArgs below (one flat array keeps the C ABI free of structs) */
int pfz_fuzz_argbest(const void *const *ptrs, int32_t n_ptrs, int32_t n_from, int32_t n_ids, int32_t n_words, int32_t n_to, int32_t scorer,
                     double score_cutoff, int32_t exclude_self, int64_t self_shift, int32_t n_splits, void *stream) {
    PFZ_REQUIRE(n_ptrs == 38, "pfz_fuzz_argbest: expected 38 pointers, got %d", n_ptrs);
    PFZ_REQUIRE(scorer >= FZ_RATIO && scorer <= FZ_WRATIO, "pfz_fuzz_argbest: unknown scorer %d", scorer);
    PFZ_REQUIRE(n_words == 1 || n_words == 2 || n_words == 4, "pfz_fuzz_argbest: n_words %d unsupported (1, 2, 4: strings up to 255 code points)", n_words);
    PFZ_REQUIRE(n_splits >= 1, "pfz_fuzz_argbest: n_splits < 1");
    if (n_ids <= 0 || n_to <= 0) return 0;
    cudaStream_t st = as_stream(stream);
    int dev = 0, sms = 0;
    PFZ_CUDA_OK(cudaGetDevice(&dev));
    PFZ_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FuzzParams P;
    int k = 0;
    auto side = [&](FuzzSide &S) {
        for (int v = 0; v < 3; ++v) { S.blob[v] = (const uint32_t *)ptrs[k++]; S.off[v] = (const int64_t *)ptrs[k++]; }
        S.tok_ptr = (const int32_t *)ptrs[k++]; S.tok_ids = (const int32_t *)ptrs[k++]; S.sig = (const uint64_t *)ptrs[k++];
        S.n_tok_all = (const int32_t *)ptrs[k++];
    };
    side(P.F); side(P.T);                                                           // 2 x 10
    P.from_ids = (const int32_t *)ptrs[k++]; P.sym_table = (const uint8_t *)ptrs[k++];   // 22
    for (int v = 0; v < 3; ++v) { P.packed[v] = (const uint32_t *)ptrs[k++]; P.grp_off[v] = (const int64_t *)ptrs[k++]; P.slen[v] = (const int32_t *)ptrs[k++]; }   // 31
    P.sorig = (const int32_t *)ptrs[k++]; P.tok_blob = (const uint32_t *)ptrs[k++]; P.tok_off = (const int64_t *)ptrs[k++];     // 34
    P.part_idx = (int32_t *)ptrs[k++]; P.part_score = (double *)ptrs[k++]; P.counter = (int32_t *)ptrs[k++];                  // 37
    k++;                                                                            // 38: reserved
    P.n_ids = n_ids; P.n_to = n_to; P.scorer = scorer; P.cutoff = score_cutoff; P.exclude_self = exclude_self; P.self_shift = self_shift;
    P.n_splits = n_splits; P.n_from = n_from;
    PFZ_CUDA_OK(cudaMemsetAsync(P.counter, 0, sizeof(int32_t) * (size_t)n_splits, st));
    if (n_words == 1) return launch_fuzz<1>(P, sms, st);
    if (n_words == 2) return launch_fuzz<2>(P, sms, st);
    return launch_fuzz<4>(P, sms, st);
}
}
