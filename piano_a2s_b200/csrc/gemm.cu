// Exact-fp32 GEMM on the FFMA pipe: C = op(A) * op(B) (+bias) (+C), batched / split-K, with an optional fused
// per-column affine(+ReLU) transform on one operand (BatchNorm-apply fused into the consumer's operand load).
//
// This is the fp32 reference-precision contraction of the library (the `out` Linear of ConvStack
// (models.py:504,539), GRU input projections (models.py:63-67), the attention encoder projection
// (models.py:444,458) and every weight-gradient contraction).  The tcgen05 path in tc_gemm.cu replaces it
// where a split-bf16 tensor-core contraction meets the parity tolerance.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256, PAD = 4;

struct GemmArgs {
    const float* A; const float* B; float* C; const float* bias;
    int M, N, K;
    long long lda, ldb, ldc;
    long long sA, sB, sC;     // batch strides (elements)
    int batch, splitk, kchunk;
    int accumulate, atomic;
    // operand transform: x' = relu?(x*scale[col % period] + shift[col % period]) on the operand's contiguous index
    const float* t_scale; const float* t_shift; int t_period; int t_relu; int t_on_b;
    int vecA, vecB;
};

__device__ __forceinline__ float xform(float x, long long col, const GemmArgs& g) {
    int c = (int)(col % g.t_period);
    float y = fmaf(x, __ldg(g.t_scale + c), __ldg(g.t_shift + c));
    return g.t_relu ? fmaxf(y, 0.f) : y;
}

// Load a (BMN x BK) operand tile into registers.
//  KC=true : memory is [mn][k] (k contiguous, leading dim ld);  KC=false: memory is [k][mn] (mn contiguous).
template <bool KC>
__device__ __forceinline__ void load_tile(const float* __restrict__ P, long long ld, int mn0, int k0, int MN, int kend,
                                          bool vec, bool tf, const GemmArgs& g, float4 (&r)[2]) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int f = tid + i * NT;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (KC) {
            int row = f >> 2, kq = (f & 3) * 4;
            int m = mn0 + row, k = k0 + kq;
            if (m < MN) {
                const float* p = P + (long long)m * ld + k;
                if (vec && k + 3 < kend) {
                    v = __ldg(reinterpret_cast<const float4*>(p));
                } else {
                    if (k + 0 < kend) v.x = __ldg(p + 0);
                    if (k + 1 < kend) v.y = __ldg(p + 1);
                    if (k + 2 < kend) v.z = __ldg(p + 2);
                    if (k + 3 < kend) v.w = __ldg(p + 3);
                }
                if (tf) {   // transform keyed on the contiguous (k) index
                    if (k + 0 < kend) v.x = xform(v.x, k + 0, g);
                    if (k + 1 < kend) v.y = xform(v.y, k + 1, g);
                    if (k + 2 < kend) v.z = xform(v.z, k + 2, g);
                    if (k + 3 < kend) v.w = xform(v.w, k + 3, g);
                }
            }
        } else {
            int kk = f >> 5, mq = (f & 31) * 4;
            int k = k0 + kk, m = mn0 + mq;
            if (k < kend) {
                const float* p = P + (long long)k * ld + m;
                if (vec && m + 3 < MN) {
                    v = __ldg(reinterpret_cast<const float4*>(p));
                } else {
                    if (m + 0 < MN) v.x = __ldg(p + 0);
                    if (m + 1 < MN) v.y = __ldg(p + 1);
                    if (m + 2 < MN) v.z = __ldg(p + 2);
                    if (m + 3 < MN) v.w = __ldg(p + 3);
                }
                if (tf) {   // transform keyed on the contiguous (mn) index
                    if (m + 0 < MN) v.x = xform(v.x, m + 0, g);
                    if (m + 1 < MN) v.y = xform(v.y, m + 1, g);
                    if (m + 2 < MN) v.z = xform(v.z, m + 2, g);
                    if (m + 3 < MN) v.w = xform(v.w, m + 3, g);
                }
            }
        }
        r[i] = v;
    }
}

template <bool KC>
__device__ __forceinline__ void store_tile(float (*S)[BM + PAD], const float4 (&r)[2]) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int f = tid + i * NT;
        if (KC) {
            int row = f >> 2, kq = (f & 3) * 4;
            S[kq + 0][row] = r[i].x; S[kq + 1][row] = r[i].y; S[kq + 2][row] = r[i].z; S[kq + 3][row] = r[i].w;
        } else {
            int kk = f >> 5, mq = (f & 31) * 4;
            *reinterpret_cast<float4*>(&S[kk][mq]) = r[i];
        }
    }
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(NT) gemm_f32_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];
    const int z = blockIdx.z;
    const int bz = z / g.splitk, sk = z % g.splitk;
    const float* A = g.A + (long long)bz * g.sA;
    const float* B = g.B + (long long)bz * g.sB;
    float* C = g.C + (long long)bz * g.sC;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = sk * g.kchunk;
    const int kend = min(g.K, kbeg + g.kchunk);
    const bool tfA = g.t_scale != nullptr && !g.t_on_b;
    const bool tfB = g.t_scale != nullptr && g.t_on_b;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    float4 ra[2], rb[2];
    // A is [m][k] when !TA (k contiguous); B is [n][k] when TB (k contiguous).
    load_tile<!TA>(A, g.lda, m0, kbeg, g.M, kend, g.vecA, tfA, g, ra);
    load_tile<TB>(B, g.ldb, n0, kbeg, g.N, kend, g.vecB, tfB, g, rb);
    store_tile<!TA>(As[0], ra);
    store_tile<TB>(Bs[0], rb);
    __syncthreads();
    int buf = 0;
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        const bool more = k0 + BK < kend;
        if (more) {
            load_tile<!TA>(A, g.lda, m0, k0 + BK, g.M, kend, g.vecA, tfA, g, ra);
            load_tile<TB>(B, g.ldb, n0, k0 + BK, g.N, kend, g.vecB, tfB, g, rb);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) {
            store_tile<!TA>(As[buf ^ 1], ra);
            store_tile<TB>(Bs[buf ^ 1], rb);
        }
        __syncthreads();
        buf ^= 1;
    }

    const bool use_atomic = g.atomic || g.splitk > 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.bias != nullptr && sk == 0) v += __ldg(g.bias + n);
            float* c = C + (long long)m * g.ldc + n;
            if (use_atomic) atomicAdd(c, v);
            else if (g.accumulate) *c += v;
            else *c = v;
        }
    }
}


// ---- skinny contractions (the bar-level decoder: batch-of-clips x weight matrix) ----------------------------------------------
// The bar GRU cell, the attention query and the time-signature / key heads (models.py:117-132, 242-286) contract a (B x K)
// activation block, B <= 32 clips, with a weight matrix: a 128 x 128 tile would be 90 % padding and leave 140 SMs idle.
constexpr int SK_MAX = 32;      // max rows of the skinny operand

__device__ __forceinline__ void skinny_store(float* c, float v, int atomic, int accumulate) {
    if (atomic) atomicAdd(c, v);
    else if (accumulate) *c += v;
    else *c = v;
}

// C[m][n] = sum_k A[m][k] * B[n][k] (+bias[n]);  M <= 32.  One warp per pair of output columns: lanes stride over k with
// 128-bit loads of the two weight rows (read once overall), the M activation rows come from L1, butterfly reduction at the end.
template <int MR>
__global__ void __launch_bounds__(256) gemm_skinny_nt_kernel(GemmArgs g) {
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int n0 = warp * 2;
    if (n0 >= g.N) return;
    const bool two = n0 + 1 < g.N;
    const float* b0 = g.B + (long long)n0 * g.ldb;
    const float* b1 = g.B + (long long)(two ? n0 + 1 : n0) * g.ldb;
    float acc0[MR], acc1[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) { acc0[m] = 0.f; acc1[m] = 0.f; }
    const bool vec = g.vecA && g.vecB;
    const int K4 = vec ? (g.K & ~3) : 0;
    for (int k = lane * 4; k < K4; k += 128) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(b0 + k)), w1 = __ldg(reinterpret_cast<const float4*>(b1 + k));
#pragma unroll
        for (int m = 0; m < MR; ++m) {
            if (m < g.M) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(g.A + (long long)m * g.lda + k));
                acc0[m] = fmaf(a.x, w0.x, fmaf(a.y, w0.y, fmaf(a.z, w0.z, fmaf(a.w, w0.w, acc0[m]))));
                acc1[m] = fmaf(a.x, w1.x, fmaf(a.y, w1.y, fmaf(a.z, w1.z, fmaf(a.w, w1.w, acc1[m]))));
            }
        }
    }
    for (int k = K4 + lane; k < g.K; k += 32) {
        const float w0 = __ldg(b0 + k), w1 = __ldg(b1 + k);
#pragma unroll
        for (int m = 0; m < MR; ++m) {
            if (m < g.M) {
                const float a = __ldg(g.A + (long long)m * g.lda + k);
                acc0[m] = fmaf(a, w0, acc0[m]);
                acc1[m] = fmaf(a, w1, acc1[m]);
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MR; ++m) {
        const float s0 = warp_sum(acc0[m]), s1 = warp_sum(acc1[m]);
        if (lane == (m & 31) && m < g.M) {
            skinny_store(g.C + (long long)m * g.ldc + n0, s0 + (g.bias ? __ldg(g.bias + n0) : 0.f), g.atomic, g.accumulate);
            if (two) skinny_store(g.C + (long long)m * g.ldc + n0 + 1, s1 + (g.bias ? __ldg(g.bias + n0 + 1) : 0.f), g.atomic, g.accumulate);
        }
    }
}

// C[m][n] = sum_k A[m][k] * B[k][n] (+bias[n]);  M <= 32.  CTA = 32 output columns x 8 k-slices; a warp reads 128-byte rows of
// B, the A values are warp-uniform; the 8 slices are summed through shared memory.
template <int MR>
__global__ void __launch_bounds__(256) gemm_skinny_nn_kernel(GemmArgs g) {
    __shared__ float red[8][MR][33];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + lane;
    // gridDim.y CTAs split K (their partial sums meet in C through atomics: the launcher zero-fills C and sets g.atomic), 8 warps split a CTA's range
    const int kcta = (g.K + gridDim.y - 1) / gridDim.y;
    const int k_lo = blockIdx.y * kcta, k_hi = min(g.K, k_lo + kcta);
    const int kper = (k_hi - k_lo + 7) / 8;
    const int kb = k_lo + slice * kper, ke = min(k_hi, kb + kper);
    float acc[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) acc[m] = 0.f;
    if (n < g.N) {
        // 8 rows of B in flight per thread (the loop is a chain of dependent L2 round trips otherwise: 75 us for 16 x 1024 x 1024)
        for (int k0 = kb; k0 < ke; k0 += 8) {
            float b[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = (k0 + j < ke) ? __ldg(g.B + (long long)(k0 + j) * g.ldb + n) : 0.f;
#pragma unroll
            for (int m = 0; m < MR; ++m) {
                if (m < g.M) {
                    const float* arow = g.A + (long long)m * g.lda + k0;
                    if (k0 + 8 <= ke && (g.lda & 3) == 0 && (k0 & 3) == 0 && ((size_t)g.A & 15) == 0) {
                        const float4 a0 = __ldg(reinterpret_cast<const float4*>(arow)), a1 = __ldg(reinterpret_cast<const float4*>(arow) + 1);
                        acc[m] = fmaf(a0.x, b[0], fmaf(a0.y, b[1], fmaf(a0.z, b[2], fmaf(a0.w, b[3], acc[m]))));
                        acc[m] = fmaf(a1.x, b[4], fmaf(a1.y, b[5], fmaf(a1.z, b[6], fmaf(a1.w, b[7], acc[m]))));
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (k0 + j < ke) acc[m] = fmaf(__ldg(arow + j), b[j], acc[m]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MR; ++m) red[slice][m][lane] = acc[m];
    __syncthreads();
    for (int m = slice; m < g.M; m += 8) {
        float v = 0.f;
#pragma unroll
        for (int s = 0; s < 8; ++s) v += red[s][m][lane];
        if (n < g.N) skinny_store(g.C + (long long)m * g.ldc + n, v + ((g.bias && blockIdx.y == 0) ? __ldg(g.bias + n) : 0.f), g.atomic, g.accumulate);
    }
}

// C[i][j] = sum_{m < K} A[m][i] * B[m][j];  K <= 32 (weight gradient of a skinny Linear: dW = dy^T x).  CTA tile 32 x 128,
// both operand slabs staged in shared memory, 16 outputs per thread, coalesced 128-bit stores.
__global__ void __launch_bounds__(256) gemm_skinny_tn_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[SK_MAX][32], Bs[SK_MAX][128];
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 128;
    for (int e = threadIdx.x; e < g.K * 32; e += 256) {
        const int m = e >> 5, i = e & 31;
        As[m][i] = (i0 + i < g.M) ? __ldg(g.A + (long long)m * g.lda + i0 + i) : 0.f;
    }
    for (int e = threadIdx.x; e < g.K * 128; e += 256) {
        const int m = e >> 7, j = e & 127;
        Bs[m][j] = (j0 + j < g.N) ? __ldg(g.B + (long long)m * g.ldb + j0 + j) : 0.f;
    }
    __syncthreads();
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int m = 0; m < g.K; ++m) {
        const float4 av = *reinterpret_cast<const float4*>(&As[m][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[m][tx * 4]);
        const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        const int i = i0 + ty * 4 + x;
        if (i >= g.M) continue;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int j = j0 + tx * 4 + y;
            if (j < g.N) skinny_store(g.C + (long long)i * g.ldc + j, acc[x][y], g.atomic, g.accumulate);
        }
    }
}

// Returns true if the contraction was handled by a skinny kernel.
bool launch_skinny(const GemmArgs& g0, int transA, int transB, cudaStream_t st) {
    if (g0.batch != 1 || g0.t_scale != nullptr || g0.K <= 0) return false;
    GemmArgs g = g0;
    if (g.splitk > 1) g.atomic = 1;
    g.splitk = 1;
    if (!transA && g.M <= SK_MAX && g.N >= 64) {
        if (transB) {
            const int warps = (g.N + 1) / 2, grid = ceil_div(warps, 8);
            if (g.M <= 16) gemm_skinny_nt_kernel<16><<<grid, 256, 0, st>>>(g);
            else gemm_skinny_nt_kernel<32><<<grid, 256, 0, st>>>(g);
        } else {
            // few column blocks (N / 32) and a long K: split K over gridDim.y so that ~150 CTAs share the weight matrix
            const int nblk = ceil_div(g.N, 32);
            int ks = 1;
            if (!g.atomic && g.K >= 512) ks = max(1, min(g.K / 128, 160 / nblk));
            if (ks > 1) {
                // the K slices add into C: zero it first unless the caller accumulates into what is there
                if (!g.accumulate && cudaMemset2DAsync(g.C, (size_t)g.ldc * sizeof(float), 0, (size_t)g.N * sizeof(float), (size_t)g.M, st) != cudaSuccess) ks = 1;
                else g.atomic = 1;
            }
            const dim3 grid(nblk, ks);
            if (g.M <= 16) gemm_skinny_nn_kernel<16><<<grid, 256, 0, st>>>(g);
            else gemm_skinny_nn_kernel<32><<<grid, 256, 0, st>>>(g);
        }
        return true;
    }
    if (transA && !transB && g.K <= SK_MAX && g.bias == nullptr && (long long)g.M * g.N >= 4096) {
        gemm_skinny_tn_kernel<<<dim3(ceil_div(g.N, 128), ceil_div(g.M, 32)), 256, 0, st>>>(g);
        return true;
    }
    return false;
}

}  // namespace

unsigned long long g_pa2s_launches = 0;

PA2S_API unsigned long long pa2s_launch_count(void) { return g_pa2s_launches; }

// See include/pa2s.h for the argument contract.
PA2S_API int pa2s_gemm_f32(void* stream, int transA, int transB, int M, int N, int K,
                           const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc,
                           const float* bias, int accumulate, int atomic,
                           int batch, long long strideA, long long strideB, long long strideC,
                           const float* t_scale, const float* t_shift, int t_period, int t_relu, int t_on_b,
                           int splitk) {
    if (M <= 0 || N <= 0 || batch <= 0) return 0;
    GemmArgs g;
    g.A = A; g.B = B; g.C = C; g.bias = bias;
    g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
    g.sA = strideA; g.sB = strideB; g.sC = strideC;
    g.batch = batch;
    if (splitk < 1) splitk = 1;
    if (splitk > 1) atomic = 1;          // a split-K request always means "add into C", even if K is too small to split
    int kchunk = ceil_div(ceil_div(K, splitk), BK) * BK;
    if (kchunk < BK) kchunk = BK;
    splitk = ceil_div(K > 0 ? K : 1, kchunk);
    g.splitk = splitk; g.kchunk = kchunk;
    g.accumulate = accumulate; g.atomic = atomic;
    g.t_scale = t_scale; g.t_shift = t_shift; g.t_period = t_period > 0 ? t_period : 1; g.t_relu = t_relu; g.t_on_b = t_on_b;
    g.vecA = (lda % 4 == 0) && (strideA % 4 == 0) && ((uintptr_t)A % 16 == 0);
    g.vecB = (ldb % 4 == 0) && (strideB % 4 == 0) && ((uintptr_t)B % 16 == 0);
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), batch * splitk);
    cudaStream_t st = (cudaStream_t)stream;
    if (launch_skinny(g, transA, transB, st)) {
        PA2S_CHECK_LAST();
        return 0;
    }
    if (!transA && !transB) gemm_f32_kernel<false, false><<<grid, NT, 0, st>>>(g);
    else if (!transA && transB) gemm_f32_kernel<false, true><<<grid, NT, 0, st>>>(g);
    else if (transA && !transB) gemm_f32_kernel<true, false><<<grid, NT, 0, st>>>(g);
    else gemm_f32_kernel<true, true><<<grid, NT, 0, st>>>(g);
    PA2S_CHECK_LAST();
    return 0;
}
