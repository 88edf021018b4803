// VQT epilogue (utilities.py:246-253), NLL losses (pretrain.py:56-93) and the clip + Adadelta update
// (pretrain.py:125-128, pretrain.yaml:44-47).
#include "common.cuh"
#include <math.h>

namespace {

// C: (rows, 2*nb) filterbank responses, (re, im) interleaved per bin.  mag: (rows, nb).  One clip = rows_per_clip rows.
// valid_rows (or NULL): frames of each clip that exist (1 + n_samples/hop of the un-padded clip); later rows are the zero padding of
// pad_spectrogram (asap.py:345-349): they do not enter the clip maximum and come out as exact zeros.
__global__ void vqt_mag_kernel(const float2* __restrict__ C, float* __restrict__ mag, unsigned int* __restrict__ clip_max,
                               long long n, int nb, int rows_per_clip, const int* __restrict__ valid_rows) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.f;
    int clip = 0;
    if (i < n) {
        float2 v = __ldg(C + i);
        m = sqrtf(v.x * v.x + v.y * v.y);
        const long long row = i / nb;
        clip = (int)(row / rows_per_clip);
        if (valid_rows != nullptr && (int)(row - (long long)clip * rows_per_clip) >= valid_rows[clip]) m = 0.f;
        mag[i] = m;
    }
    // a warp may straddle two clips only at a clip boundary; reduce when uniform, else fall back to per-lane atomics
    int clip0 = __shfl_sync(0xffffffffu, clip, 0);
    bool uniform = __all_sync(0xffffffffu, (i >= n) || clip == clip0);
    if (uniform) {
        float wm = warp_max(m);
        if ((threadIdx.x & 31) == 0 && wm > 0.f) atomicMax(clip_max + clip0, __float_as_uint(wm));
    } else if (i < n && m > 0.f) {
        atomicMax(clip_max + clip, __float_as_uint(m));
    }
}

// librosa.amplitude_to_db(|V|, ref=max, amin=1e-5, top_db=80)/80 + 1
__global__ void vqt_logscale_kernel(float* __restrict__ mag, const unsigned int* __restrict__ clip_max, long long n, int nb,
                                    int rows_per_clip, const int* __restrict__ valid_rows) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long row = i / nb;
    int clip = (int)(row / rows_per_clip);
    if (valid_rows != nullptr && (int)(row - (long long)clip * rows_per_clip) >= valid_rows[clip]) { mag[i] = 0.f; return; }
    const float amin = 1e-5f;
    float ref = fmaxf(amin, __uint_as_float(clip_max[clip]));
    float db = 20.f * log10f(fmaxf(amin, mag[i])) - 20.f * log10f(ref);
    float top = 20.f * log10f(fmaxf(amin, __uint_as_float(clip_max[clip]))) - 20.f * log10f(ref);   // log_spec.max() (= 0)
    db = fmaxf(db, top - 80.f);
    mag[i] = db * (1.f / 80.f) + 1.f;
}

// acc[0] += sum over rows (target != ignore) of logp[row, target]; acc[1] += count
__global__ void nll_fwd_kernel(const float* __restrict__ logp, const long long* __restrict__ tgt, long long rows, int V,
                               long long ignore, float* __restrict__ acc) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float s = 0.f, c = 0.f;
    if (r < rows) {
        long long t = tgt[r];
        if (t != ignore) { s = __ldg(logp + r * V + t); c = 1.f; }
    }
    s = warp_sum(s); c = warp_sum(c);
    if ((threadIdx.x & 31) == 0 && c > 0.f) { atomicAdd(acc, s); atomicAdd(acc + 1, c); }
}
// grad[row, target] = -gout / count   (grad pre-zeroed)
__global__ void nll_bwd_kernel(float* __restrict__ grad, const long long* __restrict__ tgt, long long rows, int V, long long ignore,
                               const float* __restrict__ acc, const float* __restrict__ gout) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    long long t = tgt[r];
    if (t != ignore) grad[r * V + t] = -gout[0] / acc[1];
}

__global__ void sumsq_kernel(const float4* __restrict__ g, long long n4, const float* __restrict__ tail, int ntail, double* __restrict__ out) {
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = __ldg(g + i);
        s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < ntail) s += (double)tail[threadIdx.x] * tail[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double ws[32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += ws[i];
        atomicAdd(out, t);
    }
}

// clip_grad_norm_(max_norm) + Adadelta.  sumsq[0] = sum g^2 (after any all-reduce averaging).  Skips the update when the
// norm is not finite (speechbrain check_gradients semantics).
__global__ void adadelta_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ sq, float* __restrict__ acc,
                                long long n, const double* __restrict__ sumsq, float max_norm, float lr, float rho, float eps,
                                float* __restrict__ norm_out) {
    double nrm = sqrt(sumsq[0]);
    if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out != nullptr) norm_out[0] = (float)nrm;
    if (!isfinite(nrm)) return;
    float coef = fminf(1.f, max_norm / ((float)nrm + 1e-6f));
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i] * coef;
        float s = rho * sq[i] + (1.f - rho) * gi * gi;
        float d = sqrtf(acc[i] + eps) / sqrtf(s + eps) * gi;
        acc[i] = rho * acc[i] + (1.f - rho) * d * d;
        sq[i] = s;
        p[i] -= lr * d;
    }
}


// argmax over the vocabulary + cut at the first <eos> (pretrain.py:97-117 `pred = outs.argmax(-1)` and `unpad`,
// pretrain.py:245-249): one CTA per sequence (clip, bar), one warp per row; ties resolve to the LOWEST index like
// torch.argmax; NaN rows resolve to the first NaN (torch treats NaN as the maximum).
__global__ void greedy_tokens_kernel(const float* __restrict__ logp, int L, int V, int eos, long long* __restrict__ tokens,
                                     int* __restrict__ lengths) {
    const long long seq = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __shared__ int first_eos;
    if (threadIdx.x == 0) first_eos = L;
    __syncthreads();
    for (int r = warp; r < L; r += nw) {
        const float* row = logp + (seq * L + r) * V;
        float best = -INFINITY;
        int bi = V;
        bool bnan = false;
        for (int c = lane; c < V; c += 32) {
            const float x = __ldg(row + c);
            const bool xn = x != x;
            if (bnan) continue;
            if (xn || x > best || bi == V) { best = x; bi = c; bnan = xn; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const bool on = ob != ob;
            bool take;
            if (oi == V) take = false;
            else if (bi == V) take = true;
            else if (bnan || on) take = on && (!bnan || oi < bi);
            else take = ob > best || (ob == best && oi < bi);
            if (take) { best = ob; bi = oi; bnan = on; }
        }
        if (lane == 0) {
            tokens[seq * L + r] = bi;
            if (bi == eos) atomicMin(&first_eos, r);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) lengths[seq] = first_eos;
}

// Rational resampling up/down with a polyphase FIR (audio ingest, utilities.py:242 `librosa.load(sr=16000)`):
//   y[c][m] = sum_i x[c][i] * h[half + m*down - i*up],   0 <= half + m*down - i*up < L
// h (L = 2*half+1 taps, already scaled by `up`) is designed on the host.  One thread per output sample; x is read through L1 (the
// windows of neighbouring outputs overlap almost entirely), h through the read-only cache.
__global__ void resample_poly_kernel(const float* __restrict__ x, int C, long long n_in, int up, int down, const float* __restrict__ h,
                                     int L, int half, float* __restrict__ y, long long n_out) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (m >= n_out) return;
    const long long t = m * down + half;
    long long i_lo = (t - (L - 1) + up - 1) / up;            // ceil((t - (L-1)) / up), t - (L-1) may be negative
    if (t - (L - 1) < 0) i_lo = 0;
    long long i_hi = t / up;
    if (i_hi > n_in - 1) i_hi = n_in - 1;
    const float* xc = x + (long long)c * n_in;
    double acc = 0.0;                                        // fp64 accumulation: ~20..60 taps per output, keeps 1e-7 parity with the fp64 oracle
    for (long long i = i_lo; i <= i_hi; ++i) acc += (double)__ldg(xc + i) * (double)__ldg(h + (t - i * up));
    y[(long long)c * n_out + m] = (float)acc;
}

}  // namespace

PA2S_API int pa2s_vqt_post(void* stream, const float* C, float* out, unsigned int* clip_max, int nclips, int rows_per_clip, int nb,
                          const int* valid_rows) {
    cudaStream_t st = (cudaStream_t)stream;
    long long n = (long long)nclips * rows_per_clip * nb;
    PA2S_TRY(cudaMemsetAsync(clip_max, 0, sizeof(unsigned int) * nclips, st));
    vqt_mag_kernel<<<ceil_div(n, 256), 256, 0, st>>>((const float2*)C, out, clip_max, n, nb, rows_per_clip, valid_rows);
    PA2S_CHECK_LAST();
    vqt_logscale_kernel<<<ceil_div(n, 256), 256, 0, st>>>(out, clip_max, n, nb, rows_per_clip, valid_rows);
    PA2S_CHECK_LAST();
    return 0;
}

// second half of the VQT epilogue on magnitudes the filterbank GEMM has already written (pa2s_gemm_bf16_tma_mag), in place
PA2S_API int pa2s_vqt_logscale(void* stream, float* mag, const unsigned int* clip_max, int nclips, int rows_per_clip, int nb,
                              const int* valid_rows) {
    long long n = (long long)nclips * rows_per_clip * nb;
    vqt_logscale_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(mag, clip_max, n, nb, rows_per_clip, valid_rows);
    PA2S_CHECK_LAST();
    return 0;
}

PA2S_API int pa2s_nll_fwd(void* stream, const float* logp, const long long* tgt, long long rows, int V, long long ignore, float* acc2) {
    cudaStream_t st = (cudaStream_t)stream;
    PA2S_TRY(cudaMemsetAsync(acc2, 0, 2 * sizeof(float), st));
    nll_fwd_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(logp, tgt, rows, V, ignore, acc2);
    PA2S_CHECK_LAST();
    return 0;
}

PA2S_API int pa2s_nll_bwd(void* stream, float* grad, const long long* tgt, long long rows, int V, long long ignore,
                          const float* acc2, const float* gout) {
    cudaStream_t st = (cudaStream_t)stream;
    PA2S_TRY(cudaMemsetAsync(grad, 0, sizeof(float) * rows * V, st));
    nll_bwd_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(grad, tgt, rows, V, ignore, acc2, gout);
    PA2S_CHECK_LAST();
    return 0;
}

PA2S_API int pa2s_sumsq(void* stream, const float* g, long long n, double* out, int zero_first) {
    cudaStream_t st = (cudaStream_t)stream;
    if (zero_first) PA2S_TRY(cudaMemsetAsync(out, 0, sizeof(double), st));
    long long n4 = ((uintptr_t)g % 16 == 0) ? n / 4 : 0;
    int ntail = (int)(n - n4 * 4);
    if (ntail > 256) { n4 = 0; ntail = 0; return -1; }
    sumsq_kernel<<<592, 256, 0, st>>>((const float4*)g, n4, g + n4 * 4, ntail, out);
    PA2S_CHECK_LAST();
    return 0;
}

PA2S_API int pa2s_adadelta(void* stream, float* p, const float* g, float* sq, float* acc, long long n, const double* sumsq,
                           float max_norm, float lr, float rho, float eps, float* norm_out) {
    adadelta_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(p, g, sq, acc, n, sumsq, max_norm, lr, rho, eps, norm_out);
    PA2S_CHECK_LAST();
    return 0;
}

// pad_spectrogram (datasets/syn.py:46-58, asap.py:338-350) for a whole batch: clip b's frames [row_off[b], row_off[b+1]) of the
// packed ragged buffer go to out[b][0][0 .. min(n_b, Tmax)) and the remaining frames of the clip are zero.  Pure HBM stream:
// one float4 per thread (F % 4 == 0) or one float; grid-stride over B*Tmax*F so that 148 x 8 CTAs cover any size.
template <typename V>
__global__ void pad_spectrograms_kernel(const V* __restrict__ src, const long long* __restrict__ row_off, V* __restrict__ out,
                                        long long total, int Tmax, int Fv) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int f = (int)(i % Fv);
        const long long bt = i / Fv;
        const int t = (int)(bt % Tmax);
        const int b = (int)(bt / Tmax);
        const long long r0 = __ldg(row_off + b), n = __ldg(row_off + b + 1) - r0;
        V v;
        if (t < n) v = __ldcs(src + (r0 + t) * Fv + f);
        else memset(&v, 0, sizeof(V));
        __stcs(out + i, v);
    }
}

// Audio ingest in front of the VQT (datasets/asap.py:83-86): mono = mean over channels, then audio / max|audio|.
// Pass 1 keeps the per-clip maximum of |mono| as the bit pattern of a non-negative float (atomicMax on uint is order-preserving
// for those; a NaN sample has a larger pattern than every finite value and so survives, like torch.max); pass 2 divides with an
// IEEE division (__fdiv_rn), so the result is bit-identical to torch's `audio / torch.max(torch.abs(audio))`.
__global__ void mono_absmax_kernel(const float* __restrict__ in, int C, long long n, float* __restrict__ mono, unsigned int* __restrict__ amax) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float inv = 1.0f / (float)C;
    unsigned int m = 0u;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float s = __ldcs(in + i);
        for (int c = 1; c < C; ++c) s += __ldcs(in + (long long)c * n + i);
        if (C > 1) s *= inv;
        mono[i] = s;
        m = max(m, __float_as_uint(fabsf(s)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m != 0u) atomicMax(amax, m);
}
__global__ void peak_divide_kernel(float* __restrict__ mono, long long n, const unsigned int* __restrict__ amax) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float peak = __uint_as_float(*amax);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) mono[i] = __fdiv_rn(mono[i], peak);
}

PA2S_API int pa2s_mono_peak_normalize(void* stream, const float* audio, int channels, long long n, float* out, unsigned int* scratch) {
    if (channels <= 0 || n < 0) return -1;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    PA2S_TRY(cudaMemsetAsync(scratch, 0, sizeof(unsigned int), st));
    const int grid = (int)min((long long)148 * 8, (n + 255) / 256);
    mono_absmax_kernel<<<grid, 256, 0, st>>>(audio, channels, n, out, scratch);
    PA2S_CHECK_LAST();
    peak_divide_kernel<<<grid, 256, 0, st>>>(out, n, scratch);
    PA2S_CHECK_LAST();
    return 0;
}

PA2S_API int pa2s_pad_spectrograms(void* stream, const float* packed, const long long* row_off, int B, int Tmax, int F, float* out) {
    if (B < 0 || Tmax <= 0 || F <= 0) return -1;
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (F % 4 == 0) && ((uintptr_t)packed % 16 == 0) && ((uintptr_t)out % 16 == 0);
    const int Fv = vec ? F / 4 : F;
    const long long total = (long long)B * Tmax * Fv;
    const int grid = (int)min((long long)148 * 8, (total + 255) / 256);
    if (vec) pad_spectrograms_kernel<float4><<<grid, 256, 0, st>>>((const float4*)packed, row_off, (float4*)out, total, Tmax, Fv);
    else pad_spectrograms_kernel<float><<<grid, 256, 0, st>>>(packed, row_off, out, total, Tmax, Fv);
    PA2S_CHECK_LAST();
    return 0;
}

PA2S_API int pa2s_greedy_tokens(void* stream, const float* logp, long long nseq, int L, int V, int eos, long long* tokens, int* lengths) {
    if (nseq <= 0 || L <= 0 || V <= 0) return nseq == 0 ? 0 : -1;
    greedy_tokens_kernel<<<(unsigned)nseq, 256, 0, (cudaStream_t)stream>>>(logp, L, V, eos, tokens, lengths);
    PA2S_CHECK_LAST();
    return 0;
}

PA2S_API int pa2s_resample_poly(void* stream, const float* x, int channels, long long n_in, int up, int down, const float* h, int taps,
                                float* y, long long n_out) {
    if (channels <= 0 || n_out <= 0) return 0;
    if (up < 1 || down < 1 || taps < 1 || (taps & 1) == 0) return -1;
    resample_poly_kernel<<<dim3(ceil_div(n_out, 256), channels), 256, 0, (cudaStream_t)stream>>>(x, channels, n_in, up, down, h, taps,
                                                                                               taps / 2, y, n_out);
    PA2S_CHECK_LAST();
    return 0;
}
