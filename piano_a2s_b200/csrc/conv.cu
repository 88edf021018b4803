// ConvStack kernels (models.py:463-543): 3x3 convolution over NHWC activations with the previous layer's
// BatchNorm-apply + ReLU fused into the operand load and this layer's BatchNorm batch statistics fused into the
// epilogue; matching data-gradient / weight-gradient kernels with the BatchNorm backward transform fused into
// their operand loads; BatchNorm reductions and finalisation (train-mode batch statistics, running buffers).
//
// Layout: activations are (B, T, F, C) fp32, channels innermost ("NHWC"); the reference's NCHW feature index
// c*F+f of the `out` Linear (models.py:537) is handled by permuting that weight once per step on the host side.
// Only raw (pre-BN) conv outputs are ever stored: relu(bn(y)) is recomputed from y and the per-channel
// (scale, shift) wherever it is consumed.
#include "common.cuh"

namespace {

constexpr int TH = 4, TW = 32, CT = 128;      // output tile (rows of t, cols of f), threads per CTA
constexpr int PH = TH + 2, PW = TW + 2;

struct XformFwd {          // a = relu?(x*scale + shift)
    const float* scale; const float* shift; int relu;
};
struct XformBwd {          // dy = k1*(g - k2 - xhat*k3), g = G*(z>0), z = y*zs+zb, xhat=(y-mean)*invstd
    const float* Y; const float* zs; const float* zb; const float* mean; const float* invstd;
    const float* k1; const float* k2; const float* k3;
};

struct ConvArgs {
    const float* X;        // MODE 0: raw input (pre-BN of previous layer or the spectrogram); MODE 1: G = dL/d(relu out)
    const float* W;        // packed [9][CIN][COUT]
    float* Y;              // (B,T,F,COUT)
    float* partial;        // [gridDim.x*gridDim.y*gridDim.z][2*COUT] or null
    int B, T, F, ntile;
    XformFwd xf; XformBwd xb;
};

__device__ __forceinline__ float load_fwd(const float* X, size_t idx, int ci, const XformFwd& f) {
    float v = __ldg(X + idx);
    if (f.scale != nullptr) {
        v = fmaf(v, __ldg(f.scale + ci), __ldg(f.shift + ci));
        if (f.relu) v = fmaxf(v, 0.f);
    }
    return v;
}
__device__ __forceinline__ float load_bwd(const float* G, size_t idx, int c, const XformBwd& b) {
    float y = __ldg(b.Y + idx);
    float z = fmaf(y, __ldg(b.zs + c), __ldg(b.zb + c));
    float g = z > 0.f ? __ldg(G + idx) : 0.f;
    float xh = (y - __ldg(b.mean + c)) * __ldg(b.invstd + c);
    return __ldg(b.k1 + c) * (g - __ldg(b.k2 + c) - xh * __ldg(b.k3 + c));
}

template <int CIN, int COUT, int MODE>
__global__ void __launch_bounds__(CT) conv3x3_kernel(ConvArgs a) {
    constexpr int CINP = (CIN % 2 == 0) ? CIN + 1 : CIN;
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;                                 // 9*CIN*COUT
    float* Ps = smem + 9 * CIN * COUT;                // PH*PW*CINP
    __shared__ float s_stat[CT / 32][2 * COUT];

    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int b = blockIdx.z, f0 = blockIdx.x * TW;
    for (int i = tid; i < 9 * CIN * COUT; i += CT) Ws[i] = __ldg(a.W + i);

    float st_sum = 0.f, st_sq = 0.f;                  // thread c < 2*COUT accumulates this CTA's statistics
    for (int it = 0; it < a.ntile; ++it) {
        const int t0 = (blockIdx.y * a.ntile + it) * TH;
        if (t0 >= a.T) break;
        __syncthreads();                              // previous tile's patch fully consumed (and Ws visible)
        for (int i = tid; i < PH * PW * CIN; i += CT) {
            int ci = i % CIN, pix = i / CIN;
            int c = pix % PW, r = pix / PW;
            int t = t0 - 1 + r, f = f0 - 1 + c;
            float v = 0.f;
            if (t >= 0 && t < a.T && f >= 0 && f < a.F) {
                size_t idx = (((size_t)b * a.T + t) * a.F + f) * CIN + ci;
                v = (MODE == 0) ? load_fwd(a.X, idx, ci, a.xf) : load_bwd(a.X, idx, ci, a.xb);
            }
            Ps[pix * CINP + ci] = v;
        }
        __syncthreads();

        float acc[COUT];
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[co] = 0.f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap % 3;
            const float* prow = Ps + ((ty + dy) * PW + tx + dx) * CINP;
            const float* wt = Ws + tap * CIN * COUT;
#pragma unroll 4
            for (int ci = 0; ci < CIN; ++ci) {
                const float x = prow[ci];
                const float4* w4 = reinterpret_cast<const float4*>(wt + ci * COUT);
#pragma unroll
                for (int q = 0; q < COUT / 4; ++q) {
                    float4 w = w4[q];
                    acc[q * 4 + 0] = fmaf(x, w.x, acc[q * 4 + 0]);
                    acc[q * 4 + 1] = fmaf(x, w.y, acc[q * 4 + 1]);
                    acc[q * 4 + 2] = fmaf(x, w.z, acc[q * 4 + 2]);
                    acc[q * 4 + 3] = fmaf(x, w.w, acc[q * 4 + 3]);
                }
            }
        }
        const int t = t0 + ty, f = f0 + tx;
        const bool valid = t < a.T && f < a.F;
        if (valid) {
            float4* out = reinterpret_cast<float4*>(a.Y + (((size_t)b * a.T + t) * a.F + f) * COUT);
#pragma unroll
            for (int q = 0; q < COUT / 4; ++q)
                out[q] = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
        }
        if (a.partial != nullptr) {
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
                float v = valid ? acc[co] : 0.f;
                float s = warp_sum(v), q = warp_sum(v * v);
                if (tx == 0) { s_stat[ty][co] = s; s_stat[ty][COUT + co] = q; }
            }
            __syncthreads();
            if (tid < 2 * COUT) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < CT / 32; ++w) s += s_stat[w][tid];
                st_sum += s;
            }
        }
    }
    if (a.partial != nullptr && tid < 2 * COUT) {
        size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        a.partial[cta * 2 * COUT + tid] = st_sum;
    }
    (void)st_sq;
}

// ---------------------------------------------------------------------------------------------------------
// Weight gradient: dW[co][ci][tap] = sum_p dy[p,co] * a_in[p+tap,ci]   (persistent CTAs, register accumulation)
// ---------------------------------------------------------------------------------------------------------
struct WgradArgs {
    const float* Xin;      // raw input of this layer (pre-BN output of the previous layer, or the spectrogram)
    const float* G;        // dL/d(relu out) of this layer
    float* partial;        // [gridDim.x][COUT*CIN*9]
    int B, T, F;
    int tiles_f, tiles_t;
    XformFwd xf; XformBwd xb;
};

template <int CIN, int COUT, int COB, int NTHR, int PS = 1>
__global__ void __launch_bounds__(NTHR) conv3x3_wgrad_kernel(WgradArgs a) {
    constexpr int CINP = (CIN % 2 == 0) ? CIN + 1 : CIN;
    constexpr int COUTP = COUT + 4;
    constexpr int NCOMP1 = (COUT / COB) * CIN;        // compute threads per pixel slice
    constexpr int NCOMP = NCOMP1 * PS;                // PS pixel slices (for tiny channel counts) are reduced at the end
    static_assert(NCOMP <= NTHR, "thread budget");
    __shared__ float red[PS > 1 ? COUT * CIN * 9 : 1];
    extern __shared__ __align__(16) float smem[];
    float* Ps = smem;                                 // PH*PW*CINP
    float* Ds = smem + ((PH * PW * CINP + 3) / 4) * 4; // TH*TW*COUTP (16B aligned rows)
    const int tid = threadIdx.x;
    const int ps = tid / NCOMP1, t1 = tid % NCOMP1;
    const int ci = t1 % CIN, cog = t1 / CIN;          // lanes vary in ci -> conflict-free patch reads
    float acc[COB][9];
#pragma unroll
    for (int i = 0; i < COB; ++i)
#pragma unroll
        for (int j = 0; j < 9; ++j) acc[i][j] = 0.f;

    const long long ntiles = (long long)a.B * a.tiles_t * a.tiles_f;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tf = (int)(tile % a.tiles_f);
        const int tt = (int)((tile / a.tiles_f) % a.tiles_t);
        const int b = (int)(tile / ((long long)a.tiles_f * a.tiles_t));
        const int t0 = tt * TH, f0 = tf * TW;
        __syncthreads();
        for (int i = tid; i < PH * PW * CIN; i += NTHR) {
            int c_ = i % CIN, pix = i / CIN;
            int c = pix % PW, r = pix / PW;
            int t = t0 - 1 + r, f = f0 - 1 + c;
            float v = 0.f;
            if (t >= 0 && t < a.T && f >= 0 && f < a.F)
                v = load_fwd(a.Xin, (((size_t)b * a.T + t) * a.F + f) * CIN + c_, c_, a.xf);
            Ps[pix * CINP + c_] = v;
        }
        for (int i = tid; i < TH * TW * COUT; i += NTHR) {
            int co = i % COUT, pix = i / COUT;
            int c = pix % TW, r = pix / TW;
            int t = t0 + r, f = f0 + c;
            float v = 0.f;
            if (t < a.T && f < a.F)
                v = load_bwd(a.G, (((size_t)b * a.T + t) * a.F + f) * COUT + co, co, a.xb);
            Ds[pix * COUTP + co] = v;
        }
        __syncthreads();
        if (tid < NCOMP) {
            for (int r = 0; r < TH; ++r) {
#pragma unroll 4
                for (int c = ps; c < TW; c += PS) {
                    float d[COB];
                    const float* dp = Ds + (r * TW + c) * COUTP + cog * COB;
#pragma unroll
                    for (int i = 0; i < COB; ++i) d[i] = dp[i];
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        float x = Ps[((r + tap / 3) * PW + c + tap % 3) * CINP + ci];
#pragma unroll
                        for (int i = 0; i < COB; ++i) acc[i][tap] = fmaf(d[i], x, acc[i][tap]);
                    }
                }
            }
        }
    }
    float* out = a.partial + (size_t)blockIdx.x * COUT * CIN * 9;
    if (PS == 1) {
        if (tid < NCOMP) {
#pragma unroll
            for (int i = 0; i < COB; ++i)
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) out[((cog * COB + i) * CIN + ci) * 9 + tap] = acc[i][tap];
        }
    } else {
        __syncthreads();
        for (int i = tid; i < COUT * CIN * 9; i += NTHR) red[i] = 0.f;
        __syncthreads();
        if (tid < NCOMP) {
#pragma unroll
            for (int i = 0; i < COB; ++i)
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) atomicAdd(&red[((cog * COB + i) * CIN + ci) * 9 + tap], acc[i][tap]);
        }
        __syncthreads();
        for (int i = tid; i < COUT * CIN * 9; i += NTHR) out[i] = red[i];
    }
}

// ---------------------------------------------------------------------------------------------------------
// Reductions over rows of partials
// ---------------------------------------------------------------------------------------------------------
template <typename TIN>
__global__ void reduce_rows_kernel(const TIN* __restrict__ partial, int R, int N, double* __restrict__ out64,
                                   float* __restrict__ out32, int accumulate) {
    // one CTA per column block of 32; 8 row-lanes x 32 columns, double accumulation
    __shared__ double s[8][33];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
    double acc = 0.0;
    if (col < N)
        for (int r = rl; r < R; r += 8) acc += (double)__ldg(partial + (size_t)r * N + col);
    s[rl][threadIdx.x & 31] = acc;
    __syncthreads();
    if (rl == 0 && col < N) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s[i][threadIdx.x & 31];
        if (out64 != nullptr) out64[col] = accumulate ? out64[col] + t : t;
        if (out32 != nullptr) out32[col] = accumulate ? out32[col] + (float)t : (float)t;
    }
}

// Column sums of a TALL (R, N) matrix, stage 1: CTA (x, y) sums rows [y*rpc, (y+1)*rpc) of columns [32x, 32x+32) in double
// (8 row-lanes, 4 rows in flight per lane) and writes partial[y][col]; stage 2 is reduce_rows_kernel<double> over the chunks.
// Deterministic (no atomics): bias gradients of the GRU / attention projections sum 19 216 rows (models.py:63-67,444).
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ X, int R, int N, int rpc,
                                                             double* __restrict__ partial) {
    __shared__ double s[8][33];
    const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + lane;
    const int r0 = blockIdx.y * rpc, r1 = min(R, r0 + rpc);
    double acc = 0.0;
    if (col < N) {
        const float* p = X + col;
        int r = r0 + rl;
        for (; r + 24 < r1; r += 32) {
            const float a0 = __ldg(p + (size_t)r * N), a1 = __ldg(p + (size_t)(r + 8) * N);
            const float a2 = __ldg(p + (size_t)(r + 16) * N), a3 = __ldg(p + (size_t)(r + 24) * N);
            acc += ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
        }
        for (; r < r1; r += 8) acc += (double)__ldg(p + (size_t)r * N);
    }
    s[rl][lane] = acc;
    __syncthreads();
    if (rl == 0 && col < N) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s[i][lane];
        partial[(size_t)blockIdx.y * N + col] = t;
    }
}

// Per-channel sums over an (npix, C) tensor, C % 4 == 0:  MODE 0: [sum x, sum x^2];
// MODE 1 (BatchNorm backward): [sum g, sum g*xhat] with g = G*(mask)*(z>0).
struct StatArgs {
    const float* X;        // MODE 0: values; MODE 1: raw y (pre-BN)
    const float* G;        // MODE 1: upstream gradient wrt relu output
    const float* mask;     // MODE 1: optional dropout mask (already scaled), same shape
    const float* zs; const float* zb; const float* mean; const float* invstd;
    float* partial;        // [gridDim.x][2*C]
    long long n4;          // number of float4 elements
    int C;
};

template <int MODE>
__global__ void __launch_bounds__(256) colstats_kernel(StatArgs a) {
    extern __shared__ float s[];                      // 2*C
    const int C = a.C, c4n = C / 4;
    const int nthr = (blockDim.x / c4n) * c4n;        // threads used: multiple of C/4 so each thread owns fixed channels
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s[i] = 0.f;
    __syncthreads();
    if ((int)threadIdx.x < nthr) {
        const long long gthreads = (long long)gridDim.x * nthr;
        const long long g = (long long)blockIdx.x * nthr + threadIdx.x;
        const int c0 = (int)(g % c4n) * 4;
        float s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
        float zs[4], zb[4], mu[4], is[4];
        if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                zs[j] = a.zs[c0 + j]; zb[j] = a.zb[c0 + j]; mu[j] = a.mean[c0 + j]; is[j] = a.invstd[c0 + j];
            }
        }
        for (long long i = g; i < a.n4; i += gthreads) {
            float4 x = __ldg(reinterpret_cast<const float4*>(a.X) + i);
            float xv[4] = {x.x, x.y, x.z, x.w};
            if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { s0[j] += xv[j]; s1[j] = fmaf(xv[j], xv[j], s1[j]); }
            } else {
                float4 gg = __ldg(reinterpret_cast<const float4*>(a.G) + i);
                float gv[4] = {gg.x, gg.y, gg.z, gg.w};
                if (a.mask != nullptr) {
                    float4 m = __ldg(reinterpret_cast<const float4*>(a.mask) + i);
                    gv[0] *= m.x; gv[1] *= m.y; gv[2] *= m.z; gv[3] *= m.w;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float z = fmaf(xv[j], zs[j], zb[j]);
                    float gj = z > 0.f ? gv[j] : 0.f;
                    s0[j] += gj;
                    s1[j] = fmaf(gj, (xv[j] - mu[j]) * is[j], s1[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { atomicAdd(&s[c0 + j], s0[j]); atomicAdd(&s[C + c0 + j], s1[j]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) a.partial[(size_t)blockIdx.x * 2 * C + i] = s[i];
}

// BatchNorm finalisation from fp64 sums [sum x (C), sum x^2 (C)] over `count` elements per channel.
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* running_mean,
                                   float* running_var, float* scale, float* shift, float* mean_out, float* invstd_out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (count <= 0.0) count = sums[2 * C];          // SyncBatchNorm: the element count was all-reduced with the sums (uneven rank batches)
    double mean = sums[c] / count;
    double var = sums[C + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    double invstd = 1.0 / sqrt(var + (double)eps);
    float sc = (float)((double)gamma[c] * invstd);
    scale[c] = sc;
    shift[c] = (float)((double)beta[c] - mean * (double)gamma[c] * invstd);
    mean_out[c] = (float)mean;
    invstd_out[c] = (float)invstd;
    if (running_mean != nullptr) {
        double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
        running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
    }
}

// Eval-mode affine from running statistics.
__global__ void bn_eval_affine_kernel(int C, const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                                      float* scale, float* shift, float* mean_out, float* invstd_out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float invstd = 1.0f / sqrtf(rv[c] + eps);
    scale[c] = gamma[c] * invstd;
    shift[c] = beta[c] - rm[c] * gamma[c] * invstd;
    mean_out[c] = rm[c];
    invstd_out[c] = invstd;
}

// BatchNorm backward finalisation from fp64 sums [sum g (C), sum g*xhat (C)].
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, double count, int C, const float* __restrict__ gamma,
                                       const float* __restrict__ invstd, float* dgamma, float* dbeta, float* k1, float* k2, float* k3) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (count <= 0.0) count = sums[2 * C];
    double sg = sums[c], sgx = sums[C + c];
    dbeta[c] += (float)sg;
    dgamma[c] += (float)sgx;
    k1[c] = gamma[c] * invstd[c];
    k2[c] = (float)(sg / count);
    k3[c] = (float)(sgx / count);
}

// z (npix,C) raw -> out = relu(z*scale+shift) * mask   (out_bn + ReLU + dropout of models.py:539-541)
__global__ void bn_relu_mask_kernel(const float4* __restrict__ Z, const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float4* __restrict__ mask, float4* __restrict__ out, long long n4, int C) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    int c0 = (int)((i * 4) % C);
    float4 z = __ldg(Z + i);
    float4 o;
    o.x = fmaxf(fmaf(z.x, scale[c0 + 0], shift[c0 + 0]), 0.f);
    o.y = fmaxf(fmaf(z.y, scale[c0 + 1], shift[c0 + 1]), 0.f);
    o.z = fmaxf(fmaf(z.z, scale[c0 + 2], shift[c0 + 2]), 0.f);
    o.w = fmaxf(fmaf(z.w, scale[c0 + 3], shift[c0 + 3]), 0.f);
    if (mask != nullptr) {
        float4 m = __ldg(mask + i);
        o.x *= m.x; o.y *= m.y; o.z *= m.z; o.w *= m.w;
    }
    out[i] = o;
}

// dy = k1*(g - k2 - xhat*k3) materialised (used for the small (B*T,256) `out_bn` tensor).
__global__ void bn_bwd_apply_kernel(const float4* __restrict__ G, const float4* __restrict__ Y, const float4* __restrict__ mask,
                                    XformBwd b, float4* __restrict__ out, long long n4, int C) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    int c0 = (int)((i * 4) % C);
    float4 g = __ldg(G + i), y = __ldg(Y + i);
    float gv[4] = {g.x, g.y, g.z, g.w}, yv[4] = {y.x, y.y, y.z, y.w}, o[4];
    if (mask != nullptr) {
        float4 m = __ldg(mask + i);
        gv[0] *= m.x; gv[1] *= m.y; gv[2] *= m.z; gv[3] *= m.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int c = c0 + j;
        float z = fmaf(yv[j], b.zs[c], b.zb[c]);
        float gj = z > 0.f ? gv[j] : 0.f;
        float xh = (yv[j] - b.mean[c]) * b.invstd[c];
        o[j] = b.k1[c] * (gj - b.k2[c] - xh * b.k3[c]);
    }
    out[i] = make_float4(o[0], o[1], o[2], o[3]);
}

template <int CIN, int COUT, int MODE>
int launch_conv(cudaStream_t st, ConvArgs a) {
    constexpr int CINP = (CIN % 2 == 0) ? CIN + 1 : CIN;
    size_t smem = (size_t)(9 * CIN * COUT + PH * PW * CINP) * sizeof(float);
    PA2S_TRY(cudaFuncSetAttribute(conv3x3_kernel<CIN, COUT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(a.F, TW), ceil_div(ceil_div(a.T, TH), a.ntile), a.B);
    conv3x3_kernel<CIN, COUT, MODE><<<grid, CT, smem, st>>>(a);
    PA2S_CHECK_LAST();
    return 0;
}

template <int CIN, int COUT, int COB, int NTHR, int PS = 1>
int launch_wgrad(cudaStream_t st, WgradArgs a, int nctas) {
    constexpr int CINP = (CIN % 2 == 0) ? CIN + 1 : CIN;
    size_t smem = (size_t)(((PH * PW * CINP + 3) / 4) * 4 + TH * TW * (COUT + 4)) * sizeof(float);
    PA2S_TRY(cudaFuncSetAttribute(conv3x3_wgrad_kernel<CIN, COUT, COB, NTHR, PS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3x3_wgrad_kernel<CIN, COUT, COB, NTHR, PS><<<nctas, NTHR, smem, st>>>(a);
    PA2S_CHECK_LAST();
    return 0;
}

}  // namespace

// Number of CTAs (= rows of `partial`) pa2s_conv3x3_fwd will write for this shape.
PA2S_API int pa2s_conv3x3_num_partials(int B, int T, int F, int ntile) {
    return ceil_div(F, TW) * ceil_div(ceil_div(T, TH), ntile) * B;
}

// mode 0: forward conv of relu?(X*in_scale+in_shift) (in_scale may be null = identity); writes raw Y and, if
//         `partial` != null, per-CTA [sum y, sum y^2] rows.
// mode 1: data gradient: input is dy reconstructed from (G, Yraw, bn constants); W must be the flipped/transposed pack.
PA2S_API int pa2s_conv3x3(void* stream, int mode, int B, int T, int F, int Cin, int Cout, const float* X, const float* Wpacked,
                          float* Y, float* partial, int ntile,
                          const float* in_scale, const float* in_shift, int in_relu,
                          const float* Yraw, const float* zs, const float* zb, const float* mean, const float* invstd,
                          const float* k1, const float* k2, const float* k3) {
    ConvArgs a;
    a.X = X; a.W = Wpacked; a.Y = Y; a.partial = partial; a.B = B; a.T = T; a.F = F; a.ntile = ntile > 0 ? ntile : 1;
    a.xf = XformFwd{in_scale, in_shift, in_relu};
    a.xb = XformBwd{Yraw, zs, zb, mean, invstd, k1, k2, k3};
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) {
        if (Cin == 1 && Cout == 20) return launch_conv<1, 20, 0>(st, a);
        if (Cin == 20 && Cout == 20) return launch_conv<20, 20, 0>(st, a);
        if (Cin == 20 && Cout == 40) return launch_conv<20, 40, 0>(st, a);
        if (Cin == 40 && Cout == 40) return launch_conv<40, 40, 0>(st, a);
    } else {
        if (Cin == 20 && Cout == 20) return launch_conv<20, 20, 1>(st, a);
        if (Cin == 40 && Cout == 20) return launch_conv<40, 20, 1>(st, a);
        if (Cin == 40 && Cout == 40) return launch_conv<40, 40, 1>(st, a);
    }
    return -1;   // unsupported channel configuration
}

// dW partials: `partial` is [nctas][Cout*Cin*9] (torch (Cout,Cin,3,3) order); reduce with pa2s_reduce_rows.
PA2S_API int pa2s_conv3x3_wgrad(void* stream, int B, int T, int F, int Cin, int Cout, const float* Xin, const float* G,
                                float* partial, int nctas, const float* in_scale, const float* in_shift, int in_relu,
                                const float* Yraw, const float* zs, const float* zb, const float* mean, const float* invstd,
                                const float* k1, const float* k2, const float* k3) {
    WgradArgs a;
    a.Xin = Xin; a.G = G; a.partial = partial; a.B = B; a.T = T; a.F = F;
    a.tiles_f = ceil_div(F, TW); a.tiles_t = ceil_div(T, TH);
    a.xf = XformFwd{in_scale, in_shift, in_relu};
    a.xb = XformBwd{Yraw, zs, zb, mean, invstd, k1, k2, k3};
    cudaStream_t st = (cudaStream_t)stream;
    if (Cin == 1 && Cout == 20) return launch_wgrad<1, 20, 1, 160, 8>(st, a, nctas);
    if (Cin == 20 && Cout == 20) return launch_wgrad<20, 20, 4, 128>(st, a, nctas);
    if (Cin == 20 && Cout == 40) return launch_wgrad<20, 40, 4, 224>(st, a, nctas);
    if (Cin == 40 && Cout == 40) return launch_wgrad<40, 40, 4, 416>(st, a, nctas);
    return -1;
}

PA2S_API int pa2s_reduce_rows(void* stream, const float* partial, int R, int N, double* out64, float* out32, int accumulate) {
    reduce_rows_kernel<float><<<ceil_div(N, 32), 256, 0, (cudaStream_t)stream>>>(partial, R, N, out64, out32, accumulate);
    PA2S_CHECK_LAST();
    return 0;
}

PA2S_API int pa2s_colsum(void* stream, const float* X, int R, int N, double* scratch, int nchunks, float* out32, int accumulate) {
    if (R <= 0 || N <= 0 || nchunks <= 0) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    const int rpc = ceil_div(R, nchunks);
    nchunks = ceil_div(R, rpc);
    colsum_partial_kernel<<<dim3(ceil_div(N, 32), nchunks), 256, 0, st>>>(X, R, N, rpc, scratch);
    PA2S_CHECK_LAST();
    reduce_rows_kernel<double><<<ceil_div(N, 32), 256, 0, st>>>(scratch, nchunks, N, nullptr, out32, accumulate);
    PA2S_CHECK_LAST();
    return 0;
}

// mode 0: partial[cta] = [sum x, sum x^2]; mode 1: [sum g, sum g*xhat].  Returns rows written via *nrows (fixed grid).
PA2S_API int pa2s_colstats(void* stream, int mode, const float* X, const float* G, const float* mask, long long npix, int C,
                           const float* zs, const float* zb, const float* mean, const float* invstd, float* partial, int nctas) {
    if (C % 4 != 0 || C / 4 > 256) return -1;
    StatArgs a;
    a.X = X; a.G = G; a.mask = mask; a.zs = zs; a.zb = zb; a.mean = mean; a.invstd = invstd;
    a.partial = partial; a.n4 = npix * C / 4; a.C = C;
    size_t smem = 2 * C * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) colstats_kernel<0><<<nctas, 256, smem, st>>>(a);
    else colstats_kernel<1><<<nctas, 256, smem, st>>>(a);
    PA2S_CHECK_LAST();
    return 0;
}

PA2S_API int pa2s_bn_finalize(void* stream, const double* sums, double count, int C, const float* gamma, const float* beta,
                              float eps, float momentum, float* running_mean, float* running_var,
                              float* scale, float* shift, float* mean, float* invstd) {
    bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, count, C, gamma, beta, eps, momentum,
                                                                        running_mean, running_var, scale, shift, mean, invstd);
    PA2S_CHECK_LAST();
    return 0;
}

PA2S_API int pa2s_bn_eval_affine(void* stream, int C, const float* gamma, const float* beta, const float* rm, const float* rv,
                                 float eps, float* scale, float* shift, float* mean, float* invstd) {
    bn_eval_affine_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(C, gamma, beta, rm, rv, eps, scale, shift, mean, invstd);
    PA2S_CHECK_LAST();
    return 0;
}

PA2S_API int pa2s_bn_bwd_finalize(void* stream, const double* sums, double count, int C, const float* gamma, const float* invstd,
                                  float* dgamma, float* dbeta, float* k1, float* k2, float* k3) {
    bn_bwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, count, C, gamma, invstd, dgamma, dbeta, k1, k2, k3);
    PA2S_CHECK_LAST();
    return 0;
}

PA2S_API int pa2s_bn_relu_mask(void* stream, const float* Z, const float* scale, const float* shift, const float* mask,
                               float* out, long long npix, int C) {
    if (C % 4 != 0) return -1;
    long long n4 = npix * C / 4;
    bn_relu_mask_kernel<<<ceil_div(n4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)Z, scale, shift, (const float4*)mask,
                                                                          (float4*)out, n4, C);
    PA2S_CHECK_LAST();
    return 0;
}

PA2S_API int pa2s_bn_bwd_apply(void* stream, const float* G, const float* Yraw, const float* mask, long long npix, int C,
                               const float* zs, const float* zb, const float* mean, const float* invstd,
                               const float* k1, const float* k2, const float* k3, float* out) {
    if (C % 4 != 0) return -1;
    long long n4 = npix * C / 4;
    XformBwd b{Yraw, zs, zb, mean, invstd, k1, k2, k3};
    bn_bwd_apply_kernel<<<ceil_div(n4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)G, (const float4*)Yraw, (const float4*)mask,
                                                                          b, (float4*)out, n4, C);
    PA2S_CHECK_LAST();
    return 0;
}
