// conv1 of the ConvStack (models.py:526: Conv2d(1, 20, 3x3, padding 1, no bias) on the (B,1,T,F) spectrogram) and its weight
// gradient, exact fp32.  With ONE input channel there is nothing to contract on the tensor cores (K = 9): both kernels are pure
// HBM streams over the 20-channel tensors (738 MB at B=16), so they are laid out for coalescing instead of for reuse:
// two threads per pixel, 10 output channels each, a warp touches 1 280 contiguous bytes of y / G per access, the nine input
// taps come from L1/L2 (the spectrogram is 37 MB), and nothing goes through shared memory in the pixel loop.
//   forward : y[p][co] = sum_tap W[co][tap] x[p+tap] (same fmaf order as conv3x3_kernel<1,20>: bit-identical), plus per-thread
//             BatchNorm sums reduced once per CTA;
//   wgrad   : dW[co][tap] = sum_p dy[p][co] x[p+tap] with dy = BatchNorm/ReLU backward of (G, y) formed on the fly (conv.cu
//             load_bwd), 90 accumulators per thread, reduced once per CTA.
#include "common.cuh"

namespace {

constexpr int C1 = 20, HC = 10, PXB = 128, NT1 = 2 * PXB;

// sum over the 16 lanes of equal parity (lanes 2k hold channel half 0, lanes 2k+1 half 1)
__device__ __forceinline__ float parity_sum(float v) {
#pragma unroll
    for (int o = 16; o > 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void load_taps(const float* __restrict__ X, int b, int t, int f, int T, int F, float (&x)[9]) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int tt = t + ky - 1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int ff = f + kx - 1;
            x[ky * 3 + kx] = (tt >= 0 && tt < T && ff >= 0 && ff < F) ? __ldg(X + ((size_t)b * T + tt) * F + ff) : 0.f;
        }
    }
}

__global__ void __launch_bounds__(NT1, 2) conv1_fwd_kernel(const float* __restrict__ X, const float* __restrict__ W, float* __restrict__ Y,
                                                           float* __restrict__ partial, int B, int T, int F) {
    __shared__ float sred[2 * C1];
    const int tid = threadIdx.x, px = tid >> 1, half = tid & 1;
    float w[HC][9];
#pragma unroll
    for (int c = 0; c < HC; ++c)
#pragma unroll
        for (int k = 0; k < 9; ++k) w[c][k] = __ldg(W + (half * HC + c) * 9 + k);
    float ps[HC], pq[HC];
#pragma unroll
    for (int c = 0; c < HC; ++c) { ps[c] = 0.f; pq[c] = 0.f; }
    const int nfb = (F + PXB - 1) / PXB;
    const long long ntiles = (long long)B * T * nfb;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int fb = (int)(tile % nfb);
        const long long bt = tile / nfb;
        const int t = (int)(bt % T), b = (int)(bt / T);
        const int f = fb * PXB + px;
        if (f >= F) continue;
        float x[9];
        load_taps(X, b, t, f, T, F, x);
        float y[HC];
#pragma unroll
        for (int c = 0; c < HC; ++c) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) a = fmaf(x[k], w[c][k], a);
            y[c] = a;
            ps[c] += a;
            pq[c] = fmaf(a, a, pq[c]);
        }
        float2* dst = reinterpret_cast<float2*>(Y + (((size_t)b * T + t) * F + f) * C1 + half * HC);
#pragma unroll
        for (int q = 0; q < HC / 2; ++q) dst[q] = make_float2(y[2 * q], y[2 * q + 1]);
    }
    if (partial == nullptr) return;
    if (tid < 2 * C1) sred[tid] = 0.f;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < HC; ++c) {
        const float s = parity_sum(ps[c]), q = parity_sum(pq[c]);
        if ((tid & 31) < 2) { atomicAdd(&sred[half * HC + c], s); atomicAdd(&sred[C1 + half * HC + c], q); }
    }
    __syncthreads();
    if (tid < 2 * C1) partial[(size_t)blockIdx.x * 2 * C1 + tid] = sred[tid];
}

__global__ void __launch_bounds__(NT1, 2) conv1_wgrad_kernel(const float* __restrict__ X, const float* __restrict__ G, const float* __restrict__ Yraw,
                                                             const float* __restrict__ zs, const float* __restrict__ zb,
                                                             const float* __restrict__ mean, const float* __restrict__ invstd,
                                                             const float* __restrict__ k1, const float* __restrict__ k2,
                                                             const float* __restrict__ k3, float* __restrict__ partial, int B, int T, int F) {
    __shared__ float4 cst[C1][2];                 // per channel: {zs, zb, mean, invstd}, {k1, k2, k3, -}
    __shared__ float sred[C1 * 9];
    const int tid = threadIdx.x, px = tid >> 1, half = tid & 1;
    if (tid < C1) {
        cst[tid][0] = make_float4(__ldg(zs + tid), __ldg(zb + tid), __ldg(mean + tid), __ldg(invstd + tid));
        cst[tid][1] = make_float4(__ldg(k1 + tid), __ldg(k2 + tid), __ldg(k3 + tid), 0.f);
    }
    if (tid < C1 * 9) sred[tid] = 0.f;
    __syncthreads();
    float acc[HC][9];
#pragma unroll
    for (int c = 0; c < HC; ++c)
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[c][k] = 0.f;
    const int nfb = (F + PXB - 1) / PXB;
    const long long ntiles = (long long)B * T * nfb;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int fb = (int)(tile % nfb);
        const long long bt = tile / nfb;
        const int t = (int)(bt % T), b = (int)(bt / T);
        const int f = fb * PXB + px;
        if (f >= F) continue;
        const size_t base = (((size_t)b * T + t) * F + f) * C1 + half * HC;
        float2 g2[HC / 2], y2[HC / 2];
#pragma unroll
        for (int q = 0; q < HC / 2; ++q) {
            g2[q] = __ldg(reinterpret_cast<const float2*>(G + base) + q);
            y2[q] = __ldg(reinterpret_cast<const float2*>(Yraw + base) + q);
        }
        float x[9];
        load_taps(X, b, t, f, T, F, x);
#pragma unroll
        for (int c = 0; c < HC; ++c) {
            const float4 a0 = cst[half * HC + c][0], a1 = cst[half * HC + c][1];
            const float y = (c & 1) ? y2[c >> 1].y : y2[c >> 1].x;
            const float gin = (c & 1) ? g2[c >> 1].y : g2[c >> 1].x;
            const float z = fmaf(y, a0.x, a0.y);
            const float g = z > 0.f ? gin : 0.f;
            const float xh = (y - a0.z) * a0.w;
            const float dy = a1.x * (g - a1.y - xh * a1.z);
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[c][k] = fmaf(dy, x[k], acc[c][k]);
        }
    }
#pragma unroll
    for (int c = 0; c < HC; ++c)
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float s = parity_sum(acc[c][k]);
            if ((tid & 31) < 2) atomicAdd(&sred[(half * HC + c) * 9 + k], s);
        }
    __syncthreads();
    if (tid < C1 * 9) partial[(size_t)blockIdx.x * C1 * 9 + tid] = sred[tid];
}

}  // namespace

// Y (B,T,F,20) = conv3x3(X (B,T,F), W (20,1,3,3) torch layout), zero padding; partial (or NULL): nctas rows of [sum y (20), sum y^2 (20)].
PA2S_API int pa2s_conv1_fwd(void* stream, int B, int T, int F, const float* X, const float* W, float* Y, float* partial, int nctas) {
    if (B <= 0 || T <= 0 || F <= 0 || nctas <= 0) return -1;
    conv1_fwd_kernel<<<nctas, NT1, 0, (cudaStream_t)stream>>>(X, W, Y, partial, B, T, F);
    PA2S_CHECK_LAST();
    return 0;
}

// partial: nctas rows of dW (20*1*3*3, torch order) for dy = k1*(G*(Yraw*zs+zb > 0) - k2 - (Yraw-mean)*invstd*k3); sum with pa2s_reduce_rows.
PA2S_API int pa2s_conv1_wgrad(void* stream, int B, int T, int F, const float* X, const float* G, const float* Yraw, const float* zs,
                              const float* zb, const float* mean, const float* invstd, const float* k1, const float* k2, const float* k3,
                              float* partial, int nctas) {
    if (B <= 0 || T <= 0 || F <= 0 || nctas <= 0) return -1;
    conv1_wgrad_kernel<<<nctas, NT1, 0, (cudaStream_t)stream>>>(X, G, Yraw, zs, zb, mean, invstd, k1, k2, k3, partial, B, T, F);
    PA2S_CHECK_LAST();
    return 0;
}
