// 3x3 convolution (ConvStack conv2..conv4, models.py:481-498, :526-534) as an implicit GEMM on the tcgen05 tensor cores,
// forward and data-gradient, with the neighbouring BatchNorm/ReLU work fused into the operand load / epilogue.
//
// Formulation.  Activations are (B,T,F,C) fp32, channels innermost.  A frame row is padded to fp = f+1 in [0, F+2) and cut
// into blocks of 128 padded positions; one tile = (clip b, row t, block fb) = the 128 rows (TMEM lanes) of a UMMA
// accumulator.  For tap (ky,kx) the A operand row i is the input pixel (t+ky-1, fp+kx-1).  The loader warps stage
// "windows": 130 consecutive padded positions (fb*128-1 ...) of ONE input row as bf16 planes
// [8-channel group][position][8 ch] at a 16-byte pitch.  In that no-swizzle K-major UMMA layout the row index advances
// by exactly 16 bytes, so the three kx taps of a window are the SAME shared-memory bytes addressed with a descriptor
// start address shifted by kx*16 bytes.  A CTA walks DOWN a strip (b, fb): tile t uses the windows of rows t-1, t, t+1,
// so consecutive tiles share two of their three windows and every input pixel is loaded, transformed and split ONCE
// per strip: 9 taps are fed from one new window per tile, no im2col copy.
// Halo positions (fp = 0, F+1 and t = -1, T) are staged as zeros and their output rows are simply not stored.
//
// Precision: fp32 activations/weights are split into bf16 hi + lo and accumulated in TMEM (fp32) as
// hi*hi + hi*lo + lo*hi (nsplit = 3), or hi*hi only (nsplit = 1).
//
// Roles (416 threads, 1 CTA/SM, persistent over tiles): warps 0-3 epilogue (tcgen05.ld -> fp32 NHWC store + BatchNorm
// batch-statistics partial sums), warp 4 TMEM alloc + single-thread MMA issue, warps 5-12 window loaders (LDG.128 ->
// BatchNorm-apply+ReLU of the previous layer, or the BatchNorm/ReLU backward transform for the data gradient -> split -> STS.128).
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int BM = 128;
constexpr int WENT = 130;          // window entries used (128 rows + kx in {0,1,2})
constexpr int WPIX = 137;          // plane pitch in 16-byte units (odd mod 8: conflict-free plane-strided stores)
constexpr int NWIN = 5;            // window ring slots: three rows in use by the MMAs + two being filled
constexpr int N_EPI_WARPS = 4, N_LOAD_WARPS = 8;
constexpr int NTHREADS = (N_EPI_WARPS + 1 + N_LOAD_WARPS) * 32;
constexpr int N_LOAD_THREADS = N_LOAD_WARPS * 32;

// A CTA owns the contiguous range [begin, end) of the linearised (strip, row) space, strip = b * nfb + fb; it is
// walked as chunks of consecutive rows of one strip.
struct Chunk { int b, fb, t0, n; };            // rows [t0, t0+n) of strip (b, fb)
struct Walk {
    long long pos, end; int T, nfb;
    __device__ __forceinline__ void init(int B, int T_, int F) {
        T = T_; nfb = (F + 2 + BM - 1) / BM;
        const long long total = (long long)B * nfb * T;
        const long long per = (total + gridDim.x - 1) / gridDim.x;
        pos = (long long)blockIdx.x * per;
        end = pos + per < total ? pos + per : total;
    }
    __device__ __forceinline__ bool next(Chunk& c) {
        if (pos >= end) return false;
        const int strip = (int)(pos / T);
        c.t0 = (int)(pos - (long long)strip * T);
        c.b = strip / nfb; c.fb = strip - c.b * nfb;
        const long long left = end - pos;
        c.n = (int)(left < (long long)(T - c.t0) ? left : (long long)(T - c.t0));
        pos += c.n;
        return true;
    }
};
static int conv_grid(int B, int T, int F) {
    const long long total = (long long)B * ((F + 2 + BM - 1) / BM) * T;
    return (int)(total < 148 ? total : 148);
}

struct TcConvArgs {
    const float* X;        // mode 0: raw input (B,T,F,CIN);  mode 1: G = dL/d(relu out) of this layer (B,T,F,CIN=channels of dy)
    const uint4* Wpack;    // packed bf16 weights, see pa2s_tc_conv_pack
    float* Y;              // (B,T,F,COUT)
    float* partial;        // [gridDim.x*4][2*COUT] per-warp [sum y, sum y^2] or null
    int B, T, F, nsplit;
    // mode 0 operand transform (scale may be null = identity)
    const float* scale; const float* shift; int relu;
    // mode 1 operand transform: dy = k1*(g - k2 - xhat*k3), g = G*(z>0), z = y*zs+zb, xhat = (y-mean)*invstd
    const float* Yraw; const float* zs; const float* zb; const float* mean; const float* invstd;
    const float* k1; const float* k2; const float* k3;
};

__device__ __forceinline__ void split8v(const float (&x)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * i]), h1 = __float2bfloat16_rn(x[2 * i + 1]);
        __nv_bfloat16 l0 = __float2bfloat16_rn(x[2 * i] - __bfloat162float(h0));
        __nv_bfloat16 l1 = __float2bfloat16_rn(x[2 * i + 1] - __bfloat162float(h1));
        h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int CIN, int COUT, int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) tc_conv_kernel(TcConvArgs a) {
    constexpr int CINP = (CIN + 15) / 16 * 16, COUTP = (COUT + 15) / 16 * 16;
    constexpr int NG = CINP / 8, NGR = (CIN + 7) / 8, KS = CINP / 16;
    constexpr int PLANE_BYTES = WPIX * 16;
    constexpr int SLOT_BYTES = 2 * NG * PLANE_BYTES;                 // hi planes then lo planes
    constexpr int WBLK_BYTES = 2 * COUTP * 16;                        // one (tap, ks, split) weight block: 2 k-groups x COUTP rows
    constexpr int W_BYTES = 9 * KS * 2 * WBLK_BYTES;
    constexpr int TM_COLS = 64;                                       // TMEM columns per accumulator buffer
    static_assert(COUTP <= TM_COLS, "accumulator width");

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* wsm = smem;                                              // W_BYTES
    uint8_t* win = smem + ((W_BYTES + 127) / 128) * 128;              // NWIN * SLOT_BYTES
    __shared__ uint64_t full_bar[NWIN], empty_bar[NWIN], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float cst[7][64];        // per-channel transform constants (mode 0: scale, shift; mode 1: zs,zb,mean,invstd,k1,k2,k3)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = a.T, F = a.F;
    if (tid < 64) {
        const bool in = tid < CIN;
        if (MODE == 0) {
            cst[0][tid] = (in && a.scale != nullptr) ? a.scale[tid] : 1.f;
            cst[1][tid] = (in && a.scale != nullptr) ? a.shift[tid] : 0.f;
        } else {
            cst[0][tid] = in ? a.zs[tid] : 0.f; cst[1][tid] = in ? a.zb[tid] : 0.f; cst[2][tid] = in ? a.mean[tid] : 0.f;
            cst[3][tid] = in ? a.invstd[tid] : 0.f; cst[4][tid] = in ? a.k1[tid] : 0.f; cst[5][tid] = in ? a.k2[tid] : 0.f;
            cst[6][tid] = in ? a.k3[tid] : 0.f;
        }
    }

    // weights -> smem (already bf16, already in UMMA layout); zero the window ring once (pad planes / pad entries stay zero)
    for (int i = tid; i < W_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4*>(wsm)[i] = __ldg(a.Wpack + i);
    for (int i = tid; i < NWIN * SLOT_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4*>(win)[i] = make_uint4(0, 0, 0, 0);
    if (warp == N_EPI_WARPS) {
        if (lane == 0) {
            for (int s = 0; s < NWIN; ++s) { mbar_init(&full_bar[s], N_LOAD_THREADS); mbar_init(&empty_bar[s], 1); }
            for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], N_EPI_WARPS * 32); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        tmem_alloc(&tmem_base_s, 2 * TM_COLS);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    Walk walk;
    walk.init(a.B, T, F);
    Chunk ch;

    if (warp < N_EPI_WARPS) {
        // ================================================================================= epilogue
        float ssum[2] = {0.f, 0.f}, ssq[2] = {0.f, 0.f};              // lane c keeps channels c and c+32
        uint32_t it = 0;
        while (walk.next(ch)) {
            const int fp = ch.fb * BM + warp * 32 + lane;
            const bool valid = (fp >= 1) && (fp <= F);
            for (int k = 0; k < ch.n; ++k, ++it) {
                const int acc = it & 1;
                mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
                tc_fence_after();
                float* yrow = a.Y + (((size_t)ch.b * T + ch.t0 + k) * F + (fp - 1)) * COUT;
#pragma unroll
                for (int c0 = 0; c0 < COUTP; c0 += 16) {
                    float v[16];
                    tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * TM_COLS + c0), v);
                    if (valid) {
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd)
                            if (c0 + 4 * qd < COUT)
                                reinterpret_cast<float4*>(yrow + c0)[qd] = make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
                    }
                    if (a.partial != nullptr) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int c = c0 + i;
                            if (c < COUT) {
                                const float x = valid ? v[i] : 0.f;
                                const float s1 = warp_sum(x), s2 = warp_sum(x * x);
                                if (lane == (c & 31)) { ssum[c >> 5] += s1; ssq[c >> 5] += s2; }
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&tempty_bar[acc]);
            }
        }
        if (a.partial != nullptr) {
            float* pr = a.partial + ((size_t)blockIdx.x * N_EPI_WARPS + warp) * 2 * COUT;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = lane + 32 * h;
                if (c < COUT) { pr[c] = ssum[h]; pr[COUT + c] = ssq[h]; }
            }
        }
    } else if (warp == N_EPI_WARPS) {
        // ================================================================================= MMA issuer (warp-uniform)
        const uint32_t idesc = make_idesc(BM, COUTP, 0, 0);
        const uint32_t w_base = smem_u32(wsm), win_base = smem_u32(win);
        const uint64_t wdesc0 = make_desc(w_base, COUTP * 16, 128);
        uint32_t it = 0, wbase = 0;                                   // wbase = ring index of the chunk's first window (row t0-1)
        while (walk.next(ch)) {
            for (int k = 0; k < ch.n; ++k, ++it) {
                const int acc = it & 1;
                mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TM_COLS);
                const bool last = (k == ch.n - 1);
                for (int ky = 0; ky < 3; ++ky) {
                    const uint32_t wi = wbase + k + ky;               // window of input row t0 + k + ky - 1
                    const int slot = wi % NWIN;
                    if (k == 0 || ky == 2) {                          // the two older windows were awaited by the previous tile
                        mbar_wait(&full_bar[slot], (wi / NWIN) & 1);
                        tc_fence_after();
                    }
                    const uint64_t dah0 = make_desc(win_base + slot * SLOT_BYTES, PLANE_BYTES, 128);
                    const uint64_t dal0 = desc_advance(dah0, NG * PLANE_BYTES);
                    const uint64_t dbw = desc_advance(wdesc0, (uint32_t)(ky * 3 * KS * 2 * WBLK_BYTES));
                    if (elect_one()) {
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                            for (int ks = 0; ks < KS; ++ks) {
                                const uint32_t aoff = (uint32_t)(2 * ks * PLANE_BYTES + kx * 16);
                                const uint64_t dah = desc_advance(dah0, aoff), dal = desc_advance(dal0, aoff);
                                const uint64_t dbh = desc_advance(dbw, (uint32_t)(((kx * KS + ks) * 2) * WBLK_BYTES));
                                const uint64_t dbl = desc_advance(dbh, WBLK_BYTES);
                                tc_mma(d_tmem, dah, dbh, idesc, (ky | kx | ks) != 0);
                                if (a.nsplit > 1) {
                                    tc_mma(d_tmem, dah, dbl, idesc, 1);
                                    tc_mma(d_tmem, dal, dbh, idesc, 1);
                                }
                            }
                        }
                        // a window is free once the last tile that reads it has been issued: row t0+k-1 after tile k,
                        // all three after the chunk's last tile
                        if (ky == 0 || last) tc_commit(&empty_bar[slot]);
                        if (ky == 2) tc_commit(&tfull_bar[acc]);
                    }
                    __syncwarp();
                }
            }
            wbase += ch.n + 2;
        }
    } else {
        // ================================================================================= window loaders
        // Each thread owns NU units (position j, 8-channel group g) of every window.  The raw loads of window w+1 are
        // issued before window w is converted, so the L2 latency overlaps the BatchNorm transform / bf16 split work.
        const int ltid = tid - (N_EPI_WARPS + 1) * 32;
        const bool want_lo = a.nsplit > 1;
        constexpr int NU = (WENT * NGR + N_LOAD_THREADS - 1) / N_LOAD_THREADS;
        constexpr int NV = MODE == 0 ? 2 : 4;          // float4 per unit: x (8 ch)  |  g (8 ch) + y (8 ch)
        constexpr bool HALF_LAST = (CIN % 8) != 0;      // CIN = 20: the last real plane holds 4 channels only
        struct Raw { float4 v[NU][NV]; bool ok[NU]; };
        Raw cur, nxt;
        // window = the 130 padded positions fb*128-1 .. fb*128+128 of input row trow of clip b
        auto issue = [&](Raw& r, int b, int fb, int trow) {
            const bool row_ok = trow >= 0 && trow < T;
            const size_t rowbase = ((size_t)b * T + (row_ok ? trow : 0)) * F;
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                const int u = ltid + i * N_LOAD_THREADS;
                r.ok[i] = false;
#pragma unroll
                for (int k = 0; k < NV; ++k) r.v[i][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (u >= WENT * NGR) continue;
                const int j = u / NGR, g = u - j * NGR;
                const int fp = fb * BM - 1 + j;
                if (!row_ok || fp < 1 || fp > F) continue;
                r.ok[i] = true;
                const size_t base = (rowbase + (fp - 1)) * CIN + 8 * g;
                const bool half = HALF_LAST && (g == NGR - 1);
                r.v[i][0] = __ldg(reinterpret_cast<const float4*>(a.X + base));
                if (!half) r.v[i][1] = __ldg(reinterpret_cast<const float4*>(a.X + base) + 1);
                if (MODE == 1) {
                    r.v[i][NV - 2] = __ldg(reinterpret_cast<const float4*>(a.Yraw + base));
                    if (!half) r.v[i][NV - 1] = __ldg(reinterpret_cast<const float4*>(a.Yraw + base) + 1);
                }
            }
        };
        auto convert_store = [&](const Raw& r, uint8_t* hi_base, uint8_t* lo_base) {
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                const int u = ltid + i * N_LOAD_THREADS;
                if (u >= WENT * NGR) continue;
                const int j = u / NGR, g = u - j * NGR;
                float x[8] = {r.v[i][0].x, r.v[i][0].y, r.v[i][0].z, r.v[i][0].w, r.v[i][1].x, r.v[i][1].y, r.v[i][1].z, r.v[i][1].w};
                if (r.ok[i]) {
                    if (MODE == 0) {
                        if (a.scale != nullptr) {
                            const float4 s0 = *reinterpret_cast<const float4*>(&cst[0][8 * g]), s1 = *reinterpret_cast<const float4*>(&cst[0][8 * g + 4]);
                            const float4 h0 = *reinterpret_cast<const float4*>(&cst[1][8 * g]), h1 = *reinterpret_cast<const float4*>(&cst[1][8 * g + 4]);
                            const float sc8[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                            const float sh8[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                if (8 * g + e < CIN) {            // (pad channels keep their staged zeros)
                                    float y = fmaf(x[e], sc8[e], sh8[e]);
                                    x[e] = a.relu ? fmaxf(y, 0.f) : y;
                                }
                            }
                        }
                    } else {
                        const float yy[8] = {r.v[i][NV - 2].x, r.v[i][NV - 2].y, r.v[i][NV - 2].z, r.v[i][NV - 2].w,
                                             r.v[i][NV - 1].x, r.v[i][NV - 1].y, r.v[i][NV - 1].z, r.v[i][NV - 1].w};
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int c = 8 * g + e;
                            if (c < CIN) {
                                const float z = fmaf(yy[e], cst[0][c], cst[1][c]);
                                const float gi = z > 0.f ? x[e] : 0.f;
                                const float xh = (yy[e] - cst[2][c]) * cst[3][c];
                                x[e] = cst[4][c] * (gi - cst[5][c] - xh * cst[6][c]);
                            } else {
                                x[e] = 0.f;
                            }
                        }
                    }
                }
                uint4 hi, lo;
                split8_packed(x, hi, lo);
                const int off = (g * WPIX + j) * 16;
                *reinterpret_cast<uint4*>(hi_base + off) = hi;
                if (want_lo) *reinterpret_cast<uint4*>(lo_base + off) = lo;
            }
        };
        // flattened window sequence of this CTA: for every chunk the rows t0-1 .. t0+n
        Chunk nch;
        bool have = walk.next(ch);
        int j = 0;                                                    // window j of chunk ch = input row t0 - 1 + j
        uint32_t wi = 0;
        if (have) issue(cur, ch.b, ch.fb, ch.t0 - 1);
        while (have) {
            // successor window
            bool nhave = true; int nj = j + 1; nch = ch;
            if (nj == ch.n + 2) { nj = 0; nhave = walk.next(nch); }
            if (nhave) issue(nxt, nch.b, nch.fb, nch.t0 - 1 + nj);
            const int slot = wi % NWIN;
            mbar_wait(&empty_bar[slot], ((wi / NWIN) & 1) ^ 1);
            uint8_t* hi_base = win + (size_t)slot * SLOT_BYTES;
            convert_store(cur, hi_base, hi_base + NG * PLANE_BYTES);
            fence_proxy_async();
            mbar_arrive(&full_bar[slot]);
            cur = nxt; ch = nch; j = nj; have = nhave; ++wi;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == N_EPI_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * TM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight gradient on the tensor cores:  dW[co][ci][ky][kx] = sum_q dy[q][co] * a_in[q + (ky-1)(F+2) + (kx-1)][ci].
// The reduction runs over pixels, so BOTH operands are MN-major UMMA operands (K = pixel index at a 16-byte pitch):
//   A = dy  (M = 64 >= COUT channels)  staged per tile as planes [8-channel group][128 pixels][8 ch]
//   B = a_in(N = CINP channels)        the forward kernel's three row-windows; tap kx = descriptor start + kx*16 bytes
// Nine accumulators (one per tap, CINP columns each, 64 lanes) stay in TMEM for ALL tiles of the persistent CTA and are read
// out once at the end into a per-CTA partial (summed by pa2s_reduce_rows).
constexpr int APIX = 129;          // dy plane pitch in 16-byte units (odd: conflict-free plane-strided stores)
constexpr int NASLOT = 2;

struct TcWgradArgs {
    const float* Xin;      // raw input of this layer (B,T,F,CIN)
    const float* G;        // dL/d(relu out) of this layer (B,T,F,COUT)
    float* partial;        // [gridDim.x][COUT*CIN*9]
    int B, T, F, nsplit;
    const float* scale; const float* shift; int relu;                                  // a_in = relu?(Xin*scale+shift)
    const float* Yraw; const float* zs; const float* zb; const float* mean; const float* invstd;
    const float* k1; const float* k2; const float* k3;                                 // dy transform
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(NTHREADS, 1) tc_conv_wgrad_kernel(TcWgradArgs a) {
    constexpr int CINP = (CIN + 15) / 16 * 16;
    constexpr int NGB = CINP / 8, NGRB = (CIN + 7) / 8;          // a_in planes (padded / real)
    constexpr int NGA = 8, NGRA = (COUT + 7) / 8;                 // dy planes: M = 64
    constexpr int PLANE_BYTES = WPIX * 16, SLOT_BYTES = 2 * NGB * PLANE_BYTES;
    constexpr int APLANE_BYTES = APIX * 16, ASLOT_BYTES = 2 * NGA * APLANE_BYTES;
    constexpr int TM_COLS = 512;
    static_assert(9 * CINP <= TM_COLS && COUT <= 64, "accumulators must fit TMEM");

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* win = smem;                                           // NWIN * SLOT_BYTES
    uint8_t* asm_ = smem + NWIN * SLOT_BYTES;                      // NASLOT * ASLOT_BYTES
    __shared__ uint64_t full_bar[NWIN], empty_bar[NWIN], afull_bar[NASLOT], aempty_bar[NASLOT], done_bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = a.T, F = a.F, PWD = F + 2;
    const int tiles_per_clip = (T * PWD + BM - 1) / BM;
    const long long ntiles = (long long)a.B * tiles_per_clip;

    for (int i = tid; i < (NWIN * SLOT_BYTES + NASLOT * ASLOT_BYTES) / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (warp == N_EPI_WARPS) {
        if (lane == 0) {
            for (int s = 0; s < NWIN; ++s) { mbar_init(&full_bar[s], N_LOAD_THREADS); mbar_init(&empty_bar[s], 1); }
            for (int s = 0; s < NASLOT; ++s) { mbar_init(&afull_bar[s], N_LOAD_THREADS); mbar_init(&aempty_bar[s], 1); }
            mbar_init(&done_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        tmem_alloc(&tmem_base_s, TM_COLS);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < N_EPI_WARPS) {
        // ================================================================================= final read-out
        mbar_wait(&done_bar, 0);
        tc_fence_after();
        const int co = 16 * warp + lane;                           // M = 64: row r lives in lane (r%16) + 32*(r/16)
        float* out = a.partial + (size_t)blockIdx.x * COUT * CIN * 9;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
            for (int c0 = 0; c0 < CINP; c0 += 16) {
                float v[16];
                tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(tap * CINP + c0), v);
                if (lane < 16 && co < COUT) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < CIN) out[((size_t)co * CIN + c0 + i) * 9 + tap] = v[i];
                }
            }
        }
        tc_fence_before();
    } else if (warp == N_EPI_WARPS) {
        // ================================================================================= MMA issuer (warp-uniform)
        {
            const uint32_t idesc = make_idesc(64, CINP, 1, 1);
            const uint32_t win_base = smem_u32(win), a_base0 = smem_u32(asm_);
            uint32_t it = 0, wi = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int aslot = it % NASLOT;
                mbar_wait(&afull_bar[aslot], (it / NASLOT) & 1);
                tc_fence_after();
                const uint64_t dah0 = make_desc(a_base0 + aslot * ASLOT_BYTES, 128, APLANE_BYTES);
                const uint64_t dal0 = desc_advance(dah0, NGA * APLANE_BYTES);
                for (int ky = 0; ky < 3; ++ky, ++wi) {
                    const int slot = wi % NWIN;
                    mbar_wait(&full_bar[slot], (wi / NWIN) & 1);
                    tc_fence_after();
                    const uint64_t dbh0 = make_desc(win_base + slot * SLOT_BYTES, 128, PLANE_BYTES);
                    const uint64_t dbl0 = desc_advance(dbh0, NGB * PLANE_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const uint32_t d_tmem = tmem_base + (uint32_t)((ky * 3 + kx) * CINP);
#pragma unroll
                            for (int ks = 0; ks < BM / 16; ++ks) {
                                const uint64_t dah = desc_advance(dah0, ks * 256), dal = desc_advance(dal0, ks * 256);
                                const uint64_t dbh = desc_advance(dbh0, (kx + 16 * ks) * 16), dbl = desc_advance(dbl0, (kx + 16 * ks) * 16);
                                tc_mma(d_tmem, dah, dbh, idesc, (it | (uint32_t)ks) != 0);
                                if (a.nsplit > 1) {
                                    tc_mma(d_tmem, dah, dbl, idesc, 1);
                                    tc_mma(d_tmem, dal, dbh, idesc, 1);
                                }
                            }
                        }
                        tc_commit(&empty_bar[slot]);
                        if (ky == 2) tc_commit(&aempty_bar[aslot]);
                    }
                    __syncwarp();
                }
            }
            if (elect_one()) tc_commit(&done_bar);
            __syncwarp();
        }
    } else {
        // ================================================================================= loaders (4 stages per tile)
        const int ltid = tid - (N_EPI_WARPS + 1) * 32;
        const bool want_lo = a.nsplit > 1;
        constexpr int NUW = (WENT * NGRB + N_LOAD_THREADS - 1) / N_LOAD_THREADS;
        constexpr int NUA = (BM * NGRA + N_LOAD_THREADS - 1) / N_LOAD_THREADS;
        constexpr int NU = NUW > NUA ? NUW : NUA;
        constexpr bool HALF_B = (CIN % 8) != 0, HALF_A = (COUT % 8) != 0;
        struct Raw { float4 v[NU][4]; bool ok[NU]; };
        Raw cur, nxt;
        // stage st: 0 = dy tile (A operand), 1..3 = a_in window ky = st-1 (B operand)
        auto issue = [&](Raw& r, long long tile, int st) {
            const int b = (int)(tile / tiles_per_clip);
            const int q0 = (int)(tile % tiles_per_clip) * BM;
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                r.ok[i] = false;
#pragma unroll
                for (int k = 0; k < 4; ++k) r.v[i][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int u = ltid + i * N_LOAD_THREADS;
                if (st == 0) {
                    if (i >= NUA || u >= BM * NGRA) continue;
                    const int j = u / NGRA, g = u - j * NGRA;
                    const int q = q0 + j;
                    const int t = q / PWD, fp = q - t * PWD;
                    if (t >= T || fp < 1 || fp > F) continue;
                    r.ok[i] = true;
                    const size_t base = (((size_t)b * T + t) * F + (fp - 1)) * COUT + 8 * g;
                    const bool half = HALF_A && (g == NGRA - 1);
                    r.v[i][0] = __ldg(reinterpret_cast<const float4*>(a.G + base));
                    r.v[i][2] = __ldg(reinterpret_cast<const float4*>(a.Yraw + base));
                    if (!half) {
                        r.v[i][1] = __ldg(reinterpret_cast<const float4*>(a.G + base) + 1);
                        r.v[i][3] = __ldg(reinterpret_cast<const float4*>(a.Yraw + base) + 1);
                    }
                } else {
                    if (i >= NUW || u >= WENT * NGRB) continue;
                    const int j = u / NGRB, g = u - j * NGRB;
                    const int w = q0 + (st - 2) * PWD - 1 + j;
                    if (w < 0) continue;
                    const int t = w / PWD, fp = w - t * PWD;
                    if (t >= T || fp < 1 || fp > F) continue;
                    r.ok[i] = true;
                    const size_t base = (((size_t)b * T + t) * F + (fp - 1)) * CIN + 8 * g;
                    const bool half = HALF_B && (g == NGRB - 1);
                    r.v[i][0] = __ldg(reinterpret_cast<const float4*>(a.Xin + base));
                    if (!half) r.v[i][1] = __ldg(reinterpret_cast<const float4*>(a.Xin + base) + 1);
                }
            }
        };
        auto convert_store = [&](const Raw& r, int st, uint8_t* hi_base, uint8_t* lo_base) {
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                const int u = ltid + i * N_LOAD_THREADS;
                float x[8] = {r.v[i][0].x, r.v[i][0].y, r.v[i][0].z, r.v[i][0].w, r.v[i][1].x, r.v[i][1].y, r.v[i][1].z, r.v[i][1].w};
                int off;
                if (st == 0) {
                    if (i >= NUA || u >= BM * NGRA) continue;
                    const int j = u / NGRA, g = u - j * NGRA;
                    off = (g * APIX + j) * 16;
                    if (r.ok[i]) {
                        const float yy[8] = {r.v[i][2].x, r.v[i][2].y, r.v[i][2].z, r.v[i][2].w, r.v[i][3].x, r.v[i][3].y, r.v[i][3].z, r.v[i][3].w};
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int c = 8 * g + e;
                            if (c < COUT) {
                                const float z = fmaf(yy[e], __ldg(a.zs + c), __ldg(a.zb + c));
                                const float gi = z > 0.f ? x[e] : 0.f;
                                const float xh = (yy[e] - __ldg(a.mean + c)) * __ldg(a.invstd + c);
                                x[e] = __ldg(a.k1 + c) * (gi - __ldg(a.k2 + c) - xh * __ldg(a.k3 + c));
                            } else {
                                x[e] = 0.f;
                            }
                        }
                    }
                } else {
                    if (i >= NUW || u >= WENT * NGRB) continue;
                    const int j = u / NGRB, g = u - j * NGRB;
                    off = (g * WPIX + j) * 16;
                    if (r.ok[i] && a.scale != nullptr) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int c = 8 * g + e;
                            if (c < CIN) {
                                float y = fmaf(x[e], __ldg(a.scale + c), __ldg(a.shift + c));
                                x[e] = a.relu ? fmaxf(y, 0.f) : y;
                            }
                        }
                    }
                }
                uint4 hi, lo;
                split8_packed(x, hi, lo);
                *reinterpret_cast<uint4*>(hi_base + off) = hi;
                if (want_lo) *reinterpret_cast<uint4*>(lo_base + off) = lo;
            }
        };
        long long tile = blockIdx.x;
        int st = 0;
        uint32_t wi = 0, it = 0;
        if (tile < ntiles) issue(cur, tile, 0);
        while (tile < ntiles) {
            long long ntile = tile; int nst = st + 1;
            if (nst == 4) { nst = 0; ntile += gridDim.x; }
            if (ntile < ntiles) issue(nxt, ntile, nst);
            if (st == 0) {
                const int aslot = it % NASLOT;
                mbar_wait(&aempty_bar[aslot], ((it / NASLOT) & 1) ^ 1);
                uint8_t* hb = asm_ + (size_t)aslot * ASLOT_BYTES;
                convert_store(cur, 0, hb, hb + NGA * APLANE_BYTES);
                fence_proxy_async();
                mbar_arrive(&afull_bar[aslot]);
                ++it;
            } else {
                const int slot = wi % NWIN;
                mbar_wait(&empty_bar[slot], ((wi / NWIN) & 1) ^ 1);
                uint8_t* hb = win + (size_t)slot * SLOT_BYTES;
                convert_store(cur, st, hb, hb + NGB * PLANE_BYTES);
                fence_proxy_async();
                mbar_arrive(&full_bar[slot]);
                ++wi;
            }
            cur = nxt; tile = ntile; st = nst;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == N_EPI_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TM_COLS);
    }
}

template <int CIN, int COUT>
int launch_tc_wgrad(cudaStream_t st, const TcWgradArgs& a, int* grid_out) {
    constexpr int CINP = (CIN + 15) / 16 * 16;
    constexpr int SMEM = NWIN * (2 * (CINP / 8) * WPIX * 16) + NASLOT * (2 * 8 * APIX * 16) + 1024;
    const long long ntiles = (long long)a.B * ((a.T * (a.F + 2) + BM - 1) / BM);
    const int grid = (int)(ntiles < 148 ? ntiles : 148);
    if (grid_out) *grid_out = grid;
    PA2S_TRY(cudaFuncSetAttribute(tc_conv_wgrad_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    tc_conv_wgrad_kernel<CIN, COUT><<<grid, NTHREADS, SMEM, st>>>(a);
    PA2S_CHECK_LAST();
    return 0;
}

// Pack fp32 conv weights (Cout,Cin,3,3) into the bf16 UMMA blocks the kernel expects:
//   [tap][ks][split(hi,lo)][kgroup(2)][n (NOUTP)][8]   element k = ks*16 + kgroup*8 + e  (input channel), n = output channel.
// transpose_flip = 0: forward  (n = co, k = ci, tap = ky*3+kx);
// transpose_flip = 1: data gradient (n = ci, k = co, tap = (2-ky)*3 + (2-kx)):  conv of dy with the flipped, transposed filter.
// sel (three-way split W = W1 + W2 + W3 into bf16 pieces): 0 the blocks hold (W1, W2) -- the usual hi/lo --, 1 (W3, 0), 2 (W2, W1)
__global__ void tc_conv_pack_kernel(const float* __restrict__ W, int Cout, int Cin, int transpose_flip, __nv_bfloat16* __restrict__ out,
                                    int KIN, int NOUT, int KS, int NOUTP, int sel) {
    const int total = 9 * KS * 2 * 2 * NOUTP * 8;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int e = idx % 8, r = idx / 8;
        int n = r % NOUTP; r /= NOUTP;
        int kg = r % 2; r /= 2;
        int split = r % 2; r /= 2;
        int ks = r % KS; int tap = r / KS;
        int k = ks * 16 + kg * 8 + e;
        float w = 0.f;
        if (n < NOUT && k < KIN) {
            int ky = tap / 3, kx = tap % 3;
            if (!transpose_flip) w = W[((n * Cin + k) * 3 + ky) * 3 + kx];
            else w = W[((k * Cin + n) * 3 + (2 - ky)) * 3 + (2 - kx)];
        }
        const __nv_bfloat16 w1 = __float2bfloat16_rn(w);
        const float r1 = w - __bfloat162float(w1);
        const __nv_bfloat16 w2 = __float2bfloat16_rn(r1);
        const __nv_bfloat16 w3 = __float2bfloat16_rn(r1 - __bfloat162float(w2));
        __nv_bfloat16 v;
        if (sel == 0) v = split == 0 ? w1 : w2;
        else if (sel == 1) v = split == 0 ? w3 : __float2bfloat16_rn(0.f);
        else v = split == 0 ? w2 : w1;
        out[idx] = v;
    }
}

template <int CIN, int COUT, int MODE>
int launch_tc_conv(cudaStream_t st, const TcConvArgs& a) {
    constexpr int CINP = (CIN + 15) / 16 * 16, COUTP = (COUT + 15) / 16 * 16;
    constexpr int NG = CINP / 8, KS = CINP / 16;
    constexpr int W_BYTES = 9 * KS * 2 * (2 * COUTP * 16);
    constexpr int SMEM = ((W_BYTES + 127) / 128) * 128 + NWIN * (2 * NG * WPIX * 16) + 1024;
    PA2S_TRY(cudaFuncSetAttribute(tc_conv_kernel<CIN, COUT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const long long ntiles = (long long)a.B * ((a.T * (a.F + 2) + BM - 1) / BM);
    const int grid = (int)(ntiles < 148 ? ntiles : 148);
    tc_conv_kernel<CIN, COUT, MODE><<<grid, NTHREADS, SMEM, st>>>(a);
    PA2S_CHECK_LAST();
    return 0;
}

}  // namespace

// bytes of the packed weight buffer for a (Kin -> Nout) 3x3 convolution
PA2S_API int pa2s_tc_conv_pack_bytes(int Kin, int Nout) {
    int KS = (Kin + 15) / 16, NOUTP = (Nout + 15) / 16 * 16;
    return 9 * KS * 2 * 2 * NOUTP * 8 * 2;
}
// W: (Cout,Cin,3,3) fp32.  dgrad = 0 packs the forward filter (Kin = Cin, Nout = Cout); dgrad = 1 packs the flipped,
// transposed filter of the data gradient (Kin = Cout, Nout = Cin).
PA2S_API int pa2s_tc_conv_pack(void* stream, const float* W, int Cout, int Cin, int dgrad, void* out) {
    int KIN = dgrad ? Cout : Cin, NOUT = dgrad ? Cin : Cout;
    int KS = (KIN + 15) / 16, NOUTP = (NOUT + 15) / 16 * 16;
    tc_conv_pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(W, Cout, Cin, dgrad, (__nv_bfloat16*)out, KIN, NOUT, KS, NOUTP, 0);
    PA2S_CHECK_LAST();
    return 0;
}
// the same layout with other pieces of the three-way split W = W1 + W2 + W3 in the (hi, lo) blocks: sel 0 (W1, W2), 1 (W3, 0), 2 (W2, W1)
PA2S_API int pa2s_tc_conv_pack3(void* stream, const float* W, int Cout, int Cin, int dgrad, int sel, void* out) {
    if (sel < 0 || sel > 2) return -1;
    int KIN = dgrad ? Cout : Cin, NOUT = dgrad ? Cin : Cout;
    int KS = (KIN + 15) / 16, NOUTP = (NOUT + 15) / 16 * 16;
    tc_conv_pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(W, Cout, Cin, dgrad, (__nv_bfloat16*)out, KIN, NOUT, KS, NOUTP, sel);
    PA2S_CHECK_LAST();
    return 0;
}
// Rows of `partial` written by the forward kernel: 4 per CTA, 148 CTAs at most.
PA2S_API int pa2s_tc_conv_num_partials(int B, int T, int F) {
    long long ntiles = (long long)B * ((T * (F + 2) + BM - 1) / BM);
    return (int)(ntiles < 148 ? ntiles : 148) * N_EPI_WARPS;
}
// conv2d backward wrt weight on the tensor cores; `partial` has pa2s_tc_conv_wgrad_num_partials rows of Cout*Cin*9.
PA2S_API int pa2s_tc_conv_wgrad_num_partials(int B, int T, int F) {
    long long ntiles = (long long)B * ((T * (F + 2) + BM - 1) / BM);
    return (int)(ntiles < 148 ? ntiles : 148);
}
PA2S_API int pa2s_tc_conv3x3_wgrad(void* stream, int B, int T, int F, int Cin, int Cout, const float* Xin, const float* G,
                                   float* partial, int nsplit, const float* in_scale, const float* in_shift, int in_relu,
                                   const float* Yraw, const float* zs, const float* zb, const float* mean, const float* invstd,
                                   const float* k1, const float* k2, const float* k3) {
    TcWgradArgs a;
    a.Xin = Xin; a.G = G; a.partial = partial; a.B = B; a.T = T; a.F = F; a.nsplit = nsplit >= 3 ? 3 : 1;
    a.scale = in_scale; a.shift = in_shift; a.relu = in_relu;
    a.Yraw = Yraw; a.zs = zs; a.zb = zb; a.mean = mean; a.invstd = invstd; a.k1 = k1; a.k2 = k2; a.k3 = k3;
    cudaStream_t st = (cudaStream_t)stream;
    if (Cin == 20 && Cout == 20) return launch_tc_wgrad<20, 20>(st, a, nullptr);
    if (Cin == 20 && Cout == 40) return launch_tc_wgrad<20, 40>(st, a, nullptr);
    if (Cin == 40 && Cout == 40) return launch_tc_wgrad<40, 40>(st, a, nullptr);
    return -1;
}
// mode 0 / 1 as pa2s_conv3x3 (Cin = channels of the tensor being convolved, Cout = channels produced).
PA2S_API int pa2s_tc_conv3x3(void* stream, int mode, int B, int T, int F, int Cin, int Cout, const float* X, const void* Wpack,
                             float* Y, float* partial, int nsplit,
                             const float* in_scale, const float* in_shift, int in_relu,
                             const float* Yraw, const float* zs, const float* zb, const float* mean, const float* invstd,
                             const float* k1, const float* k2, const float* k3) {
    TcConvArgs a;
    a.X = X; a.Wpack = (const uint4*)Wpack; a.Y = Y; a.partial = partial; a.B = B; a.T = T; a.F = F; a.nsplit = nsplit >= 3 ? 3 : 1;
    a.scale = in_scale; a.shift = in_shift; a.relu = in_relu;
    a.Yraw = Yraw; a.zs = zs; a.zb = zb; a.mean = mean; a.invstd = invstd; a.k1 = k1; a.k2 = k2; a.k3 = k3;
    if ((long long)T * (F + 2) + 2 * (F + 2) + 256 > 0x7fffffffLL) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) {
        if (Cin == 20 && Cout == 20) return launch_tc_conv<20, 20, 0>(st, a);
        if (Cin == 20 && Cout == 40) return launch_tc_conv<20, 40, 0>(st, a);
        if (Cin == 40 && Cout == 40) return launch_tc_conv<40, 40, 0>(st, a);
    } else {
        if (Cin == 20 && Cout == 20) return launch_tc_conv<20, 20, 1>(st, a);
        if (Cin == 40 && Cout == 20) return launch_tc_conv<40, 20, 1>(st, a);
        if (Cin == 40 && Cout == 40) return launch_tc_conv<40, 40, 1>(st, a);
    }
    return -1;
}
