// Word error counts of the evaluation step right after the hot path: `calculate_wer` of pretrain.py:216-227 scores,
// per clip, jiwer.wer(target, pred) of the strings  " \n = \n ".join(idx2string(unpad(bar)) for bar in bars).  jiwer's default
// transform collapses whitespace runs and splits on blanks, so the scored WORDS of a clip are: for each bar the tokens before the
// first <eos> whose label is not pure whitespace ("\t", "\n"), with one "=" word between consecutive bars; and
// wer = Levenshtein(reference words, hypothesis words) / #reference words.
//
// One CTA per clip: (1) both word sequences are built in shared memory (first-<eos> search per bar, order-preserving compaction
// with warp ballots), (2) the edit-distance table is swept by anti-diagonals -- cell (i, j) needs (i-1, j), (i, j-1), (i-1, j-1),
// i.e. the two previous diagonals, kept in three rotating shared-memory rows -- with one __syncthreads per diagonal.
// Integer work, bit-exact against the row-by-row DP of oracle/metrics_oracle.py.
#include "common.cuh"

namespace {

constexpr int WER_THREADS = 256;

// words of one clip -> out[0..n); returns n (every thread gets the value).  tok: (bars, L) int64 rows.
__device__ int build_words(const long long* __restrict__ tok, int bars, int L, int eos, int skip_a, int skip_b, int sep, int* out,
                           int* s_len, int* s_cnt) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    for (int b = tid; b < bars; b += blockDim.x) s_len[b] = L;
    __syncthreads();
    for (int i = tid; i < bars * L; i += blockDim.x)
        if (tok[i] == eos) atomicMin(&s_len[i / L], i % L);
    __syncthreads();
    // kept words per bar (one warp per bar), then the bars' start offsets (bar b also contributes its leading "=" when b > 0)
    for (int b = warp; b < bars; b += nw) {
        int c = 0;
        for (int p = lane; p < s_len[b]; p += 32) {
            const long long t = tok[(size_t)b * L + p];
            c += (t != skip_a && t != skip_b);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) s_cnt[b] = c;
    }
    __syncthreads();
    if (tid == 0) {
        int off = 0;
        for (int b = 0; b < bars; ++b) {
            const int c = s_cnt[b];
            s_cnt[b] = off + (b > 0);              // where bar b's first word goes
            if (b > 0) out[off] = sep;
            off += c + (b > 0);
        }
        s_cnt[bars] = off;
    }
    __syncthreads();
    for (int b = warp; b < bars; b += nw) {
        int base = s_cnt[b];
        for (int p0 = 0; p0 < s_len[b]; p0 += 32) {
            const int p = p0 + lane;
            long long t = 0;
            bool keep = false;
            if (p < s_len[b]) {
                t = tok[(size_t)b * L + p];
                keep = (t != skip_a && t != skip_b);
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) out[base + __popc(m & ((1u << lane) - 1u))] = (int)t;
            base += __popc(m);
        }
    }
    __syncthreads();
    return s_cnt[bars];
}

__global__ void __launch_bounds__(WER_THREADS) wer_counts_kernel(const long long* __restrict__ hyp, const long long* __restrict__ ref, int bars,
                                                                 int Lh, int Lr, int eos, int skip_a, int skip_b, int sep,
                                                                 int* __restrict__ dist, int* __restrict__ nref, int* __restrict__ nhyp) {
    extern __shared__ int sm[];
    const int clip = blockIdx.x, tid = threadIdx.x;
    const int maxh = bars * Lh + bars, maxr = bars * Lr + bars;
    int* hw = sm;                       // hypothesis words
    int* rw = hw + maxh;                // reference words
    int* d0 = rw + maxr;                // three diagonals, indexed by i (reference position 0..m)
    int* d1 = d0 + maxr + 1;
    int* d2 = d1 + maxr + 1;
    int* s_len = d2 + maxr + 1;         // [bars]
    int* s_cnt = s_len + bars;          // [bars + 1]
    const int n = build_words(hyp + (size_t)clip * bars * Lh, bars, Lh, eos, skip_a, skip_b, sep, hw, s_len, s_cnt);
    const int m = build_words(ref + (size_t)clip * bars * Lr, bars, Lr, eos, skip_a, skip_b, sep, rw, s_len, s_cnt);
    // D[i][j]: i reference words vs j hypothesis words.  diagonal d = i + j; `cur` receives diagonal d from p1 (d-1) and p2 (d-2)
    int *p2 = d0, *p1 = d1, *cur = d2;
    for (int d = 0; d <= m + n; ++d) {
        const int ilo = max(0, d - n), ihi = min(m, d);
        for (int i = ilo + tid; i <= ihi; i += blockDim.x) {
            const int j = d - i;
            int v;
            if (i == 0) v = j;
            else if (j == 0) v = i;
            else {
                const int sub = p2[i - 1] + (rw[i - 1] != hw[j - 1]);
                v = min(sub, min(p1[i - 1] + 1, p1[i] + 1));
            }
            cur[i] = v;
        }
        __syncthreads();
        int* t = p2; p2 = p1; p1 = cur; cur = t;
    }
    if (tid == 0) {
        dist[clip] = p1[m];             // diagonal m + n holds the single cell (m, n)
        nref[clip] = m;
        nhyp[clip] = n;
    }
}

}  // namespace

PA2S_API int pa2s_wer_counts(void* stream, const long long* hyp, const long long* ref, int nclips, int bars, int Lh, int Lr, int eos,
                             int skip_a, int skip_b, int sep, int* dist, int* nref, int* nhyp) {
    if (nclips < 0 || bars <= 0 || Lh <= 0 || Lr <= 0) return -1;
    if (nclips == 0) return 0;
    const size_t smem = sizeof(int) * ((size_t)(bars * Lh + bars) + (size_t)(bars * Lr + bars) + 3 * (size_t)(bars * Lr + bars + 1) + 2 * bars + 1);
    if (smem > 227 * 1024) return -2;
    PA2S_TRY(cudaFuncSetAttribute(wer_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wer_counts_kernel<<<nclips, WER_THREADS, smem, (cudaStream_t)stream>>>(hyp, ref, bars, Lh, Lr, eos, skip_a, skip_b, sep, dist, nref, nhyp);
    PA2S_CHECK_LAST();
    return 0;
}
