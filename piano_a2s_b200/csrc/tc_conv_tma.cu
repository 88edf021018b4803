// 3x3 convolution of ConvStack (conv2..conv4, models.py:481-498, :526-534) as implicit GEMMs on the tcgen05 tensor cores,
// forward, data gradient and weight gradient, fed ONLY by bulk async copies (TMA engine, cp.async.bulk) -- no thread touches
// operand data.
//
// Activation operands live in global memory as bf16 "planes", produced once per tensor by pa2s_planes_fwd / pa2s_planes_bwd
// (BatchNorm-apply + ReLU, resp. the BatchNorm/ReLU backward transform, fused with the fp32 -> bf16 hi/lo split) and consumed
// by two kernels each (forward conv + next layer's weight gradient; data gradient + weight gradient):
//
//     P[b][t' = t+1 in 0..T+1][piece (hi, lo)][g = 8-channel group][q = f+2 in 0..FP-1][8 channels]      (16-byte units)
//
// Rows t' = 0, T+1 and columns q < 2, q >= F+2 are zero (written by the producers), so the 3x3 halo needs no predication.
// A "window" = the 130 positions q0 .. q0+129, q0 = fb*128, of one row of one channel group: ONE contiguous 2080-byte chunk,
// and in shared memory exactly the UMMA no-swizzle canonical layout (K-major for the convolution: row = position at a 16-byte
// pitch; MN-major for the weight gradient: K = position).  The three kx taps of a window are the same bytes addressed with the
// descriptor start shifted by kx*16 B; a CTA walks DOWN a strip (b, fb) so consecutive tiles share two of their three windows:
// every activation byte is copied into shared memory once per strip.
//
// Warp roles, one persistent CTA per SM.  conv_tma3_kernel (forward / data gradient, the default): 320 threads = warps 0-7 two
// epilogue groups (one per TMEM accumulator: TMEM -> registers -> tap shift -> global, BatchNorm statistics), warp 8 TMEM
// allocation + elected-lane tcgen05.mma issue, warp 9 single-lane copy producer.  conv_tma_kernel (the round-1 formulation, kept as
// comparator) and conv_wgrad_tma_kernel: 192 threads = warps 0-3 epilogue / read-out, warp 4 MMA issue, warp 5 copy producer.
#include "tc_common.cuh"

#ifndef PA2S_CONV_NCAT
#define PA2S_CONV_NCAT 1          // 1: hi/lo weight blocks concatenated along N (2 MMAs per tap and k-step), 0: 3 MMAs of N = COUTP
#endif

namespace {
using namespace tc;

constexpr int BM = 128;
constexpr int WENT = 130;                    // window entries: 128 output positions + 2 halo
constexpr int PLANE_BYTES = WENT * 16;       // 2080
constexpr int NWIN = 5;                      // window ring: three rows in use by the MMAs + two in flight
constexpr int NTHREADS = 192;
constexpr int NT3 = 320;                     // conv_tma3_kernel: 2 x 4 epilogue warps, MMA warp, copy warp
__host__ __device__ constexpr int conv3_nwin(int w_bytes, int slot_bytes) {      // window ring of conv_tma3_kernel: what fits, 5..8
    const int n = (227 * 1024 - 12 * 1024 - ((w_bytes + 127) / 128) * 128) / slot_bytes;      // 12 KB: static shared memory + alignment
    return n > 8 ? 8 : n;
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct Geom {
    int B, T, F, nfb, FP, NP;                // FP = plane row pitch in positions, NP = pieces (1 or 2)
    int nfbc;                                // strips per row of conv_tma3_kernel
};
__host__ __device__ inline int geom_nfb(int F) { return (F + 2 + BM - 1) / BM; }
// conv_tma3_kernel: a tile = the 128 window entries q0 .. q0+127, q0 = fb*OT, and produces OT = 126 outputs f = q0-1 .. q0+124
constexpr int OT = 126;
__host__ __device__ inline int geom_nfbc(int F) { return (F + 1 + OT - 1) / OT; }
__host__ __device__ inline int geom_fp(int F) {
    const int a = geom_nfb(F) * BM + 8, b = ((geom_nfbc(F) - 1) * OT + WENT + 7) / 8 * 8;      // every window read stays inside its plane row
    return a > b ? a : b;
}
// 16-byte unit index of (b, t', piece, g, q) in a plane tensor with NGR stored channel groups
__device__ __forceinline__ size_t plane_unit(const Geom& g, int NGR, int b, int tp, int piece, int grp, int q) {
    return ((((size_t)b * (g.T + 2) + tp) * g.NP + piece) * NGR + grp) * g.FP + q;
}

// A CTA owns the contiguous range [begin, end) of the linearised (strip, row) space, strip = b * nfb + fb, walked as chunks
// of consecutive rows of one strip.
struct Chunk { int b, fb, t0, n; };
struct Walk {
    long long pos, end; int T, nfb;
    __device__ __forceinline__ void init(const Geom& g, int nfb_ = 0) {
        T = g.T; nfb = nfb_ > 0 ? nfb_ : g.nfb;
        const long long total = (long long)g.B * nfb * T;
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
// number of CTAs such that every CTA owns at least one tile under Walk's ceil-division
static int conv_grid(int B, int T, int F, int nfb = 0) {
    const long long total = (long long)B * (nfb > 0 ? nfb : geom_nfb(F)) * T;
    const long long g0 = total < 148 ? total : 148;
    const long long per = (total + g0 - 1) / g0;
    return (int)((total + per - 1) / per);
}

// ====================================================================================================================
// plane producers
// ====================================================================================================================
struct PlaneArgs {
    const float* X;            // fwd: raw conv output y (B,T,F,C);  bwd: G = dL/d relu(bn(y))
    const float* Yraw;         // bwd only
    uint4* P;
    Geom g;
    const float* c0; const float* c1; const float* c2; const float* c3; const float* c4; const float* c5; const float* c6;
    int relu;
    int low;                   // MODE 0: 1 = write the two LOWER pieces (p3, p2) of the three-way split instead of (p1, p2)
};
// MODE 0: a = relu?(x*c0[c] + c1[c])  (c0 null: identity)
// MODE 1: dy = c4*(gm - c5 - xhat*c6), gm = G*(y*c0+c1 > 0), xhat = (y-c2)*c3      (BatchNorm + ReLU backward, see conv.cu)
template <int C, int MODE>
__global__ void __launch_bounds__(128 * ((C + 7) / 8)) planes_kernel(PlaneArgs a) {
    constexpr int NGR = (C + 7) / 8;
    constexpr bool HALF = (C % 8) != 0;
    const int q = blockIdx.x * 128 + threadIdx.x, grp = threadIdx.y;
    const int tp = blockIdx.y, b = blockIdx.z;
    // per-channel constants of the backward transform: staged once per CTA (7 x C floats) and read as warp-wide broadcasts instead
    // of 7 global loads per element (48 -> 32 registers; measured: no change of the 1.03 ms of the 40-channel launch, which is
    // bound by its three strided 32-byte-per-lane streams, not by these loads)
    __shared__ float cs[MODE == 1 ? 7 : 1][NGR * 8];
    if (MODE == 1) {
        const int i = threadIdx.y * 128 + threadIdx.x;
        if (i < 7 * NGR * 8) {
            const int k = i / (NGR * 8), c = i % (NGR * 8);
            const float* src = k == 0 ? a.c0 : k == 1 ? a.c1 : k == 2 ? a.c2 : k == 3 ? a.c3 : k == 4 ? a.c4 : k == 5 ? a.c5 : a.c6;
            cs[k][c] = c < C ? __ldg(src + c) : 0.f;
        }
        __syncthreads();
    }
    if (q >= a.g.FP) return;
    const int f = q - 2, t = tp - 1;
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = 0.f;
    if (t >= 0 && t < a.g.T && f >= 0 && f < a.g.F) {
        const size_t base = (((size_t)b * a.g.T + t) * a.g.F + f) * C + 8 * grp;
        const bool half = HALF && grp == NGR - 1;
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(a.X + base));
        float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!half) v1 = __ldg(reinterpret_cast<const float4*>(a.X + base) + 1);
        x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
        if (MODE == 0) {
            if (a.c0 != nullptr) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int c = 8 * grp + e;
                    if (c < C) {
                        const float y = fmaf(x[e], __ldg(a.c0 + c), __ldg(a.c1 + c));
                        x[e] = a.relu ? fmaxf(y, 0.f) : y;
                    }
                }
            }
        } else {
            const float4 y0 = __ldg(reinterpret_cast<const float4*>(a.Yraw + base));
            float4 y1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!half) y1 = __ldg(reinterpret_cast<const float4*>(a.Yraw + base) + 1);
            const float yy[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = 8 * grp + e;
                if (c < C) {
                    const float z = fmaf(yy[e], cs[0][c], cs[1][c]);
                    const float gm = z > 0.f ? x[e] : 0.f;
                    const float xh = (yy[e] - cs[2][c]) * cs[3][c];
                    x[e] = cs[4][c] * (gm - cs[5][c] - xh * cs[6][c]);
                } else {
                    x[e] = 0.f;
                }
            }
        }
    }
    uint4 hi, lo;
    if (MODE == 0 && a.low) split8_packed_low(x, hi, lo);
    else split8_packed(x, hi, lo);
    a.P[plane_unit(a.g, NGR, b, tp, 0, grp, q)] = hi;
    if (a.g.NP > 1) a.P[plane_unit(a.g, NGR, b, tp, 1, grp, q)] = lo;
}

// ====================================================================================================================
// forward / data-gradient convolution
// ====================================================================================================================
struct ConvArgs2 {
    const uint4* P;        // input planes (CIN channels)
    const uint4* Wpack;    // packed bf16 weights (pa2s_tc_conv_pack)
    float* Y;              // (B,T,F,COUT) fp32
    float* partial;        // [gridDim.x*4][2*COUT] per-warp [sum y, sum y^2] or null
    Geom g;
    // conv_tma3_kernel as the data gradient of layer i: with sY = the raw output y of layer i-1 (same shape as Y) the statistics
    // become those of the BatchNorm/ReLU BACKWARD of layer i-1: [sum g, sum g * xhat], g = Y * (y*zs + zb > 0), xhat = (y - mu) * is
    // (what pa2s_colstats mode 1 computes in a separate pass over Y and y)
    const float* sY; const float* szs; const float* szb; const float* smu; const float* sis;
    int accumulate;        // conv_tma3_kernel: Y += result (passes 2 and 3 of the three-piece convolution)
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(NTHREADS, 1) conv_tma_kernel(ConvArgs2 a) {
    constexpr int CINP = (CIN + 15) / 16 * 16, COUTP = (COUT + 15) / 16 * 16;
    constexpr int NG = CINP / 8, NGR = (CIN + 7) / 8, KS = CINP / 16;
    constexpr int SLOT_BYTES = 2 * NG * PLANE_BYTES;                 // hi planes then lo planes
    constexpr int WBLK_BYTES = 2 * COUTP * 16;                        // one (tap, ks, split) weight block: 2 k-groups x COUTP rows
    constexpr int W_BYTES = 9 * KS * 2 * WBLK_BYTES;
    // bf16x3 (NP = 2): the hi and lo weight blocks sit side by side along N, so ONE MMA of N = 2*COUTP forms A_hi*W_hi (columns
    // [0,COUTP)) and A_hi*W_lo (columns [COUTP,2*COUTP)) from a single read of the A window; A_lo*W_hi follows with N = COUTP.
    // These small-N MMAs are bound by the shared-memory read of A (4 KB per MMA), so 2 reads per (tap, k-step) instead of 3.
    constexpr bool NCAT = PA2S_CONV_NCAT != 0;
    constexpr int TM_COLS = 2 * COUTP <= 64 ? 64 : 128;
    static_assert(2 * COUTP <= TM_COLS, "accumulator width");

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* wsm = smem;
    uint8_t* win = smem + ((W_BYTES + 127) / 128) * 128;
    __shared__ uint64_t full_bar[NWIN], empty_bar[NWIN], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = a.g.T, F = a.g.F;
    const int NP = a.g.NP;

    // pa2s_tc_conv_pack order is [tap][ks][split][kgroup][n]; staged as [tap][ks][kgroup][split][n] (16-byte units) so that
    // rows n = 0..2*COUTP-1 of a k-group are [W_hi | W_lo]
    for (int i = tid; i < W_BYTES / 16; i += NTHREADS) {
        const int n = i % COUTP, r = i / COUTP, kg = r & 1, split = (r >> 1) & 1, blk = r >> 2;
        reinterpret_cast<uint4*>(wsm)[((blk * 2 + kg) * 2 + split) * COUTP + n] = __ldg(a.Wpack + i);
    }
    for (int i = tid; i < NWIN * SLOT_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4*>(win)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 4) {
        if (lane == 0) {
            for (int s = 0; s < NWIN; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
            for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 128); }
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
    walk.init(a.g);
    Chunk ch;

    if (warp < 4) {
        // ================================================================================= epilogue
        // BatchNorm batch statistics: every thread (= one output pixel column of the tile) keeps its own running sum y, sum y^2
        // per channel over all its tiles (~500 pixels per thread at B=16); the 32 rows of a warp are combined ONCE after the
        // walk.  (Reducing per tile cost 10 shuffles per channel per tile and made the epilogue the slowest stage.)
        float ps[COUT], pq[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) { ps[c] = 0.f; pq[c] = 0.f; }
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
                    if (NCAT && NP > 1) {                             // + A_hi * W_lo
                        float v2[16];
                        tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * TM_COLS + COUTP + c0), v2);
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] += v2[i];
                    }
                    if (valid) {
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd)
                            if (c0 + 4 * qd < COUT)
                                reinterpret_cast<float4*>(yrow + c0)[qd] = make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
                    }
                    if (a.partial != nullptr && valid) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int c = c0 + i;
                            if (c < COUT) { ps[c] += v[i]; pq[c] = fmaf(v[i], v[i], pq[c]); }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&tempty_bar[acc]);
            }
        }
        if (a.partial != nullptr) {
            float* pr = a.partial + ((size_t)blockIdx.x * 4 + warp) * 2 * COUT;
#pragma unroll
            for (int c = 0; c < COUT; ++c) {
                const float s1 = warp_sum(ps[c]), s2 = warp_sum(pq[c]);
                if (lane == 0) { pr[c] = s1; pr[COUT + c] = s2; }
            }
        }
    } else if (warp == 4) {
        // ================================================================================= MMA issuer (warp-uniform)
        // All waits of a tile come first, then ONE elected block issues its 9 * KS * (2 or 3) instructions.  (An elected block per ky
        // cost a divergence + reconvergence and ~120 setup instructions three times per tile, during which the tensor pipe ran dry:
        // ~1000 of the 2670 cycles of a tile.  Everything stays warp-uniform so that the descriptors live in uniform registers --
        // a single-lane role makes the compiler wrap every tcgen05.mma in a waterfall loop.)
        const uint32_t idesc = make_idesc(BM, COUTP, 0, 0), idesc2 = make_idesc(BM, 2 * COUTP, 0, 0);
        const uint32_t w_base = smem_u32(wsm), win_base = smem_u32(win);
        const uint64_t wdesc0 = make_desc(w_base, 2 * COUTP * 16, 128);      // k-group pitch: hi rows + lo rows
        uint32_t it = 0, wbase = 0;                                   // wbase = ring index of the chunk's first window (row t0-1)
        while (walk.next(ch)) {
            for (int k = 0; k < ch.n; ++k, ++it) {
                const int acc = it & 1;
                mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
                const uint32_t wi0 = wbase + k;                       // windows wi0 + ky = input rows t0 + k + ky - 1
                if (k == 0) {                                         // the two older windows were awaited by the previous tile
                    mbar_wait(&full_bar[wi0 % NWIN], (wi0 / NWIN) & 1);
                    mbar_wait(&full_bar[(wi0 + 1) % NWIN], ((wi0 + 1) / NWIN) & 1);
                }
                mbar_wait(&full_bar[(wi0 + 2) % NWIN], ((wi0 + 2) / NWIN) & 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TM_COLS);
                const bool last = (k == ch.n - 1);
                // everything that varies at run time is formed OUTSIDE the elected block (warp-uniform -> uniform registers)
                uint32_t sl[3];
                uint64_t dA[3];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    sl[ky] = (wi0 + ky) % NWIN;
                    dA[ky] = make_desc(win_base + sl[ky] * SLOT_BYTES, PLANE_BYTES, 128);
                }
                if (elect_one()) {
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const uint32_t slot = sl[ky];
                        const uint64_t dah0 = dA[ky];
                        const uint64_t dal0 = desc_advance(dah0, NG * PLANE_BYTES);
                        const uint64_t dbw = desc_advance(wdesc0, (uint32_t)(ky * 3 * KS * 2 * WBLK_BYTES));
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                            for (int ks = 0; ks < KS; ++ks) {
                                const uint32_t aoff = (uint32_t)(2 * ks * PLANE_BYTES + kx * 16);
                                const uint64_t dah = desc_advance(dah0, aoff), dal = desc_advance(dal0, aoff);
                                const uint64_t dbh = desc_advance(dbw, (uint32_t)(((kx * KS + ks) * 2) * WBLK_BYTES));
                                if (NCAT && NP > 1) {
                                    tc_mma(d_tmem, dah, dbh, idesc2, (ky | kx | ks) != 0);     // A_hi * [W_hi | W_lo]
                                    tc_mma(d_tmem, dal, dbh, idesc, 1);                        // A_lo * W_hi
                                } else if (NP > 1) {                                           // three N = COUTP products
                                    tc_mma(d_tmem, dah, dbh, idesc, (ky | kx | ks) != 0);
                                    tc_mma(d_tmem, dah, desc_advance(dbh, COUTP * 16), idesc, 1);
                                    tc_mma(d_tmem, dal, dbh, idesc, 1);
                                } else {
                                    tc_mma(d_tmem, dah, dbh, idesc, (ky | kx | ks) != 0);
                                }
                            }
                        }
                        // a window is free once the last tile that reads it has been issued
                        if (ky == 0 || last) tc_commit(&empty_bar[slot]);
                        if (ky == 2) tc_commit(&tfull_bar[acc]);
                    }
                }
                __syncwarp();
            }
            wbase += ch.n + 2;
        }
    } else if (lane == 0) {
        // ================================================================================= copy producer (one thread)
        const uint32_t win_base = smem_u32(win);
        uint32_t wi = 0;
        while (walk.next(ch)) {
            for (int j = 0; j < ch.n + 2; ++j, ++wi) {                // input rows t0-1 .. t0+n  ->  t' = t0 + j
                const int slot = wi % NWIN;
                mbar_wait(&empty_bar[slot], ((wi / NWIN) & 1) ^ 1);
                mbar_expect_tx(&full_bar[slot], (uint32_t)(NP * NGR * PLANE_BYTES));
                for (int p = 0; p < NP; ++p)
                    for (int grp = 0; grp < NGR; ++grp)
                        bulk_g2s(win_base + slot * SLOT_BYTES + (p * NG + grp) * PLANE_BYTES,
                                 a.P + plane_unit(a.g, NGR, ch.b, ch.t0 + j, p, grp, ch.fb * BM), PLANE_BYTES, &full_bar[slot]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * TM_COLS);
    }
}

// ====================================================================================================================
// forward / data-gradient convolution, second formulation: the three kx taps share ONE read of the activation window.
//
// conv_tma_kernel issues one instruction group per tap with the A descriptor shifted by kx entries: 9 * KS * 2 instructions of
// N = 32..96 per tile, each reading its 4 KB A operand from shared memory again -- measured (tools/conv_prof.py, experiments in
// DESIGN.md): a tile is bound by the operand reads and by the issue overhead of its many small instructions, the tensor pipe is
// ~35 % busy.  Here A is the UNSHIFTED window (entries e = 0..127) and the B operand of a (ky, k-step) is the three kx filters
// side by side, N3 = 3 * COUT columns:   E_kx[e][co] = sum_ci a[e][ci] * W[ky][kx][ci][co].   The tap shift moves into the
// epilogue:   y[o][co] = E_0[o] + E_1[o+1] + E_2[o+2],  o = 0..125   (lane shuffles; the two rows a warp needs from its neighbour
// travel through shared memory).  3 * KS * 3 instructions per tile of N = 80 / 128, a third of the A reads.  A tile yields
// 126 outputs, strips advance by 126 positions (geom_nfbc).
// ====================================================================================================================
#ifdef PA2S_CONV_PROF
// development build only (PA2S_NVCC_DEFS=-DPA2S_CONV_PROF): clock64 stamps of the first 64 tiles of CTA 0 of the conv2 forward
// launch, read back by tools/conv_prof.py.  Slots: MMA warp 0 before / 1 after the accumulator wait, 2 after the window waits,
// 3 after issuing the tile; epilogue group of the tile 4 before / 5 after the wait for the accumulator, 6 after its stores.
__device__ unsigned long long g_conv_prof[64 * 8];
#define CONV_STAMP(slot_) do { if (PROF_ON && blockIdx.x == 0 && it < 64) g_conv_prof[it * 8 + (slot_)] = clock64(); } while (0)
#else
#define CONV_STAMP(slot_) do { } while (0)
#endif
template <int CIN, int COUT>
struct Conv3Cfg {
    static constexpr int CINP = (CIN + 15) / 16 * 16, NG = CINP / 8, NGR = (CIN + 7) / 8, KS = CINP / 16;
    static constexpr int CQ = (COUT + 7) / 8 * 8;                  // filter rows per kx block
    static constexpr int N3 = (3 * CQ + 15) / 16 * 16;             // instruction N
    static constexpr int SLOT_BYTES = 2 * NG * PLANE_BYTES;        // hi planes then lo planes
    static constexpr int WBLK_BYTES = 2 * 2 * N3 * 16;             // one (ky, ks) block: 2 k-groups x [hi rows | lo rows]
    static constexpr int W_BYTES = 3 * KS * WBLK_BYTES;
    static constexpr int NW = conv3_nwin(W_BYTES, SLOT_BYTES);
    static constexpr int SMEM = ((W_BYTES + 127) / 128) * 128 + NW * SLOT_BYTES + 1024;
    static constexpr int TM_COLS = N3 <= 64 ? 64 : 128;            // per accumulator; three accumulators, 4 * TM_COLS allocated (power of two)
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(NT3, 1) conv_tma3_kernel(ConvArgs2 a) {
    using Cfg = Conv3Cfg<CIN, COUT>;
    constexpr int NG = Cfg::NG, NGR = Cfg::NGR, KS = Cfg::KS, CQ = Cfg::CQ, N3 = Cfg::N3;
    constexpr int SLOT_BYTES = Cfg::SLOT_BYTES, WBLK_BYTES = Cfg::WBLK_BYTES, W_BYTES = Cfg::W_BYTES, NW = Cfg::NW, TM_COLS = Cfg::TM_COLS;
    constexpr int COUTP = (COUT + 15) / 16 * 16;                      // rows per block of the packed filter (pa2s_tc_conv_pack)
    constexpr int NCH = CQ / 8;                                       // 8-channel chunks of the epilogue
    static_assert(N3 <= TM_COLS && N3 <= 256, "accumulator width");

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* wsm = smem;
    uint8_t* win = smem + ((W_BYTES + 127) / 128) * 128;
    constexpr int NACCB = 3;                                          // TMEM accumulators: the MMA warp may run two tiles ahead of a read-out
    __shared__ uint64_t full_bar[NW], empty_bar[NW], tfull_bar[NACCB], tempty_bar[NACCB];
    __shared__ uint32_t tmem_base_s;
    __shared__ float cst[4][CQ];                                      // zs, zb, mu, is of the backward statistics
    __shared__ __align__(16) float xch[2][2][4][3][CQ];               // [epilogue group][tile parity][warp][E1 of lane 0, E2 of lane 0, E2 of lane 1][channel]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = a.g.T, F = a.g.F;
    const int NP = a.g.NP;

    // filter blocks: [ky][ks][kgroup][split][kx * CQ + n] (16-byte units), rows 3*CQ .. N3-1 zero.  Source order (pa2s_tc_conv_pack):
    // [tap = ky*3+kx][ks][split][kgroup][n < COUTP]
    for (int i = tid; i < W_BYTES / 16; i += NT3) reinterpret_cast<uint4*>(wsm)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < NW * SLOT_BYTES / 16; i += NT3) reinterpret_cast<uint4*>(win)[i] = make_uint4(0, 0, 0, 0);
    const bool dstat = a.sY != nullptr;
#ifdef PA2S_CONV_PROF
    const bool PROF_ON = (CIN == 20 && COUT == 20 && a.partial != nullptr && !dstat);
#endif
    if (dstat && tid < 4 * CQ) {
        const int kk = tid / CQ, c = tid % CQ;
        const float* src = kk == 0 ? a.szs : kk == 1 ? a.szb : kk == 2 ? a.smu : a.sis;
        cst[kk][c] = c < COUT ? __ldg(src + c) : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < 9 * KS * 2 * 2 * COUTP; i += NT3) {
        const int n = i % COUTP;
        int r = i / COUTP;
        const int kg = r & 1, split = (r >> 1) & 1;
        r >>= 2;
        const int ks = r % KS, tap = r / KS, ky = tap / 3, kx = tap % 3;
        if (n < CQ) reinterpret_cast<uint4*>(wsm)[((((ky * KS + ks) * 2 + kg) * 2 + split) * N3) + kx * CQ + n] = __ldg(a.Wpack + i);
    }
    if (warp == 8) {
        if (lane == 0) {
            for (int s = 0; s < NW; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
            for (int s = 0; s < NACCB; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 128); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        tmem_alloc(&tmem_base_s, 4 * TM_COLS);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    Walk walk;
    walk.init(a.g, a.g.nfbc);
    Chunk ch;

    if (warp < 8) {
        // ================================================================================= epilogue: two groups of four warps, group g
        // takes the tiles of accumulator g (it & 1 == g), so that a tile's read-out may last two tile periods
        const int grp = warp >> 2, wq = warp & 3;                     // wq = TMEM lane quarter
        float ps[COUT], pq[COUT];                                     // BatchNorm batch statistics, per thread over all its tiles
#pragma unroll
        for (int c = 0; c < COUT; ++c) { ps[c] = 0.f; pq[c] = 0.f; }
        const int o = wq * 32 + lane;                                 // output index of the tile = window entry of E_0
        uint32_t it = 0;
        while (walk.next(ch)) {
            const int f = ch.fb * OT + o - 1;
            const bool valid = (o < OT) && (f >= 0) && (f < F);
            for (int k = 0; k < ch.n; ++k, ++it) {
                if ((int)(it & 1) != grp) continue;
                const int acc = it % NACCB;
                const size_t yoff = (((size_t)ch.b * T + ch.t0 + k) * F + f) * COUT;
                float* yrow = a.Y + yoff;
                // backward statistics: this pixel's row of the layer below, requested before the wait for the accumulator
                float4 yp[COUT / 4];
                if (dstat && valid) {
#pragma unroll
                    for (int j = 0; j < COUT / 4; ++j) yp[j] = __ldg(reinterpret_cast<const float4*>(a.sY + yoff) + j);
                } else if (a.accumulate && valid) {               // Y += : the row written by the previous pass
#pragma unroll
                    for (int j = 0; j < COUT / 4; ++j) yp[j] = *(reinterpret_cast<const float4*>(yrow) + j);
                }
                if (lane == 0 && wq == 0) CONV_STAMP(4);
                mbar_wait(&tfull_bar[acc], (it / NACCB) & 1);
                tc_fence_after();
                if (lane == 0 && wq == 0) CONV_STAMP(5);
                const uint32_t tbase = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * TM_COLS);
                // pass 1: the rows the previous warp's last two outputs need (E_1 of lane 0, E_2 of lanes 0 and 1), all channels,
                // into shared memory; ONE barrier per tile (buffers alternate with the tile parity)
                float (*xw)[3][CQ] = xch[grp][(it >> 1) & 1];
#pragma unroll
                for (int c8 = 0; c8 < NCH; ++c8) {
                    uint32_t r1[8], r2[8];
                    tc_ld8_nowait(tbase + 1 * CQ + c8 * 8, r1);
                    tc_ld8_nowait(tbase + 2 * CQ + c8 * 8, r2);
                    tc_ld_wait();
                    if (lane < 2) {
                        float4* d1 = reinterpret_cast<float4*>(&xw[wq][lane == 0 ? 0 : 2][c8 * 8]);      // lane 0: E_1, lane 1: E_2
                        const uint32_t* s1 = lane == 0 ? r1 : r2;
                        d1[0] = make_float4(__uint_as_float(s1[0]), __uint_as_float(s1[1]), __uint_as_float(s1[2]), __uint_as_float(s1[3]));
                        d1[1] = make_float4(__uint_as_float(s1[4]), __uint_as_float(s1[5]), __uint_as_float(s1[6]), __uint_as_float(s1[7]));
                        if (lane == 0) {
                            float4* d2 = reinterpret_cast<float4*>(&xw[wq][1][c8 * 8]);
                            d2[0] = make_float4(__uint_as_float(r2[0]), __uint_as_float(r2[1]), __uint_as_float(r2[2]), __uint_as_float(r2[3]));
                            d2[1] = make_float4(__uint_as_float(r2[4]), __uint_as_float(r2[5]), __uint_as_float(r2[6]), __uint_as_float(r2[7]));
                        }
                    }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                const int wn = (wq + 1) & 3;                        // (warp 3's last two lanes are not outputs)
                // pass 2: y[o] = E_0[o] + E_1[o+1] + E_2[o+2], branch-free
#pragma unroll
                for (int c8 = 0; c8 < NCH; ++c8) {
                    uint32_t r0[8], r1[8], r2[8];
                    tc_ld8_nowait(tbase + 0 * CQ + c8 * 8, r0);
                    tc_ld8_nowait(tbase + 1 * CQ + c8 * 8, r1);
                    tc_ld8_nowait(tbase + 2 * CQ + c8 * 8, r2);
                    float x0[8], x1[8], x2[8];                        // neighbour rows (broadcast reads)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 q0 = reinterpret_cast<const float4*>(&xw[wn][0][c8 * 8])[h];
                        const float4 q1 = reinterpret_cast<const float4*>(&xw[wn][1][c8 * 8])[h];
                        const float4 q2 = reinterpret_cast<const float4*>(&xw[wn][2][c8 * 8])[h];
                        x0[4 * h] = q0.x; x0[4 * h + 1] = q0.y; x0[4 * h + 2] = q0.z; x0[4 * h + 3] = q0.w;
                        x1[4 * h] = q1.x; x1[4 * h + 1] = q1.y; x1[4 * h + 2] = q1.z; x1[4 * h + 3] = q1.w;
                        x2[4 * h] = q2.x; x2[4 * h + 1] = q2.y; x2[4 * h + 2] = q2.z; x2[4 * h + 3] = q2.w;
                    }
                    tc_ld_wait();
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float e1 = __shfl_down_sync(0xffffffffu, __uint_as_float(r1[i]), 1);
                        float e2 = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[i]), 2);
                        e1 = lane == 31 ? x0[i] : e1;
                        e2 = lane == 31 ? x2[i] : (lane == 30 ? x1[i] : e2);
                        v[i] = __uint_as_float(r0[i]) + e1 + e2;
                    }
                    if (valid) {
                        if (a.accumulate) {
                            const float4 o0 = yp[2 * c8], o1 = yp[(2 * c8 + 1) < COUT / 4 ? 2 * c8 + 1 : 0];
                            v[0] += o0.x; v[1] += o0.y; v[2] += o0.z; v[3] += o0.w;
                            v[4] += o1.x; v[5] += o1.y; v[6] += o1.z; v[7] += o1.w;
                        }
                        if (c8 * 8 < COUT) reinterpret_cast<float4*>(yrow + c8 * 8)[0] = make_float4(v[0], v[1], v[2], v[3]);
                        if (c8 * 8 + 4 < COUT) reinterpret_cast<float4*>(yrow + c8 * 8)[1] = make_float4(v[4], v[5], v[6], v[7]);
                        if (a.partial != nullptr && !dstat) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int c = c8 * 8 + i;
                                if (c < COUT) { ps[c] += v[i]; pq[c] = fmaf(v[i], v[i], pq[c]); }
                            }
                        } else if (a.partial != nullptr) {
                            float yv[8];
                            const float4 y0 = yp[2 * c8];
                            yv[0] = y0.x; yv[1] = y0.y; yv[2] = y0.z; yv[3] = y0.w;
                            if (c8 * 8 + 4 < COUT) {
                                const float4 y1 = yp[(2 * c8 + 1) < COUT / 4 ? 2 * c8 + 1 : 0];
                                yv[4] = y1.x; yv[5] = y1.y; yv[6] = y1.z; yv[7] = y1.w;
                            } else {
                                yv[4] = yv[5] = yv[6] = yv[7] = 0.f;
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int c = c8 * 8 + i;
                                if (c < COUT) {
                                    const float z = fmaf(yv[i], cst[0][c], cst[1][c]);
                                    const float gj = z > 0.f ? v[i] : 0.f;
                                    ps[c] += gj;
                                    pq[c] = fmaf(gj, (yv[i] - cst[2][c]) * cst[3][c], pq[c]);
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&tempty_bar[acc]);
                if (lane == 0 && wq == 0) CONV_STAMP(6);
            }
        }
        if (a.partial != nullptr) {
            float* pr = a.partial + ((size_t)blockIdx.x * 8 + warp) * 2 * COUT;
#pragma unroll
            for (int c = 0; c < COUT; ++c) {
                const float s1 = warp_sum(ps[c]), s2 = warp_sum(pq[c]);
                if (lane == 0) { pr[c] = s1; pr[COUT + c] = s2; }
            }
        }
    } else if (warp == 8) {
        // ================================================================================= MMA issuer (warp-uniform)
        const uint32_t idesc = make_idesc(BM, N3, 0, 0);
        const uint32_t w_base = smem_u32(wsm), win_base = smem_u32(win);
        const uint64_t wdesc0 = make_desc(w_base, 2 * N3 * 16, 128);         // k-group pitch: hi rows + lo rows
        uint32_t it = 0, wbase = 0;                                   // wbase = ring index of the chunk's first window (row t0-1)
        while (walk.next(ch)) {
            for (int k = 0; k < ch.n; ++k, ++it) {
                const int acc = it % NACCB;
                if (lane == 0) CONV_STAMP(0);
                mbar_wait(&tempty_bar[acc], ((it / NACCB) & 1) ^ 1);
                if (lane == 0) CONV_STAMP(1);
                const uint32_t wi0 = wbase + k;                       // windows wi0 + ky = input rows t0 + k + ky - 1
                if (k == 0) {                                         // the two older windows were awaited by the previous tile
                    mbar_wait(&full_bar[wi0 % NW], (wi0 / NW) & 1);
                    mbar_wait(&full_bar[(wi0 + 1) % NW], ((wi0 + 1) / NW) & 1);
                }
                mbar_wait(&full_bar[(wi0 + 2) % NW], ((wi0 + 2) / NW) & 1);
                tc_fence_after();
                if (lane == 0) CONV_STAMP(2);
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TM_COLS);
                const bool last = (k == ch.n - 1);
                uint32_t sl[3];
                uint64_t dA[3];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    sl[ky] = (wi0 + ky) % NW;
                    dA[ky] = make_desc(win_base + sl[ky] * SLOT_BYTES, PLANE_BYTES, 128);
                }
                if (elect_one()) {
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            const uint64_t dah = desc_advance(dA[ky], (uint32_t)(2 * ks * PLANE_BYTES));
                            const uint64_t dal = desc_advance(dah, NG * PLANE_BYTES);
                            const uint64_t dbh = desc_advance(wdesc0, (uint32_t)((ky * KS + ks) * WBLK_BYTES));
                            const uint64_t dbl = desc_advance(dbh, N3 * 16);
                            tc_mma(d_tmem, dah, dbh, idesc, (ky | ks) != 0);             // A_hi * W_hi
                            if (NP > 1) {
                                tc_mma(d_tmem, dah, dbl, idesc, 1);                       // A_hi * W_lo
                                tc_mma(d_tmem, dal, dbh, idesc, 1);                       // A_lo * W_hi
                            }
                        }
                        // a window is free once the last tile that reads it has been issued
                        if (ky == 0 || last) tc_commit(&empty_bar[sl[ky]]);
                        if (ky == 2) tc_commit(&tfull_bar[acc]);
                    }
                }
                __syncwarp();
                if (lane == 0) CONV_STAMP(3);
            }
            wbase += ch.n + 2;
        }
    } else if (lane == 0) {
        // ================================================================================= copy producer (one thread)
        const uint32_t win_base = smem_u32(win);
        uint32_t wi = 0;
        while (walk.next(ch)) {
            for (int j = 0; j < ch.n + 2; ++j, ++wi) {                // input rows t0-1 .. t0+n  ->  t' = t0 + j
                const int slot = wi % NW;
                mbar_wait(&empty_bar[slot], ((wi / NW) & 1) ^ 1);
                mbar_expect_tx(&full_bar[slot], (uint32_t)(NP * NGR * PLANE_BYTES));
                for (int p = 0; p < NP; ++p)
                    for (int grp = 0; grp < NGR; ++grp)
                        bulk_g2s(win_base + slot * SLOT_BYTES + (p * NG + grp) * PLANE_BYTES,
                                 a.P + plane_unit(a.g, NGR, ch.b, ch.t0 + j, p, grp, ch.fb * OT), PLANE_BYTES, &full_bar[slot]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 4 * TM_COLS);
    }
}

// ====================================================================================================================
// weight gradient:  dW[co][ci][ky][kx] = sum_pixels dy[pixel][co] * a_in[pixel + (ky-1, kx-1)][ci]
// The reduction runs over pixels, so BOTH operands are MN-major UMMA operands (K = position at a 16-byte pitch):
//   A = dy   (M = 64 >= COUT)  window of row t,  entries 1..128
//   B = a_in (N = CINP)        windows of rows t-1, t, t+1 (shared between consecutive tiles); tap kx = start + kx*16 bytes
// Nine accumulators (one per tap, CINP columns, 64 or 128 lanes) stay in TMEM for ALL tiles of the persistent CTA and are read out
// once at the end into a per-CTA partial (summed by pa2s_reduce_rows).
// ====================================================================================================================
constexpr int NASLOT = 3;
struct WgradArgs2 {
    const uint4* Pin;      // a_in planes (CIN channels)
    const uint4* Pdy;      // dy planes (COUT channels)
    float* partial;        // [2 * gridDim.x][COUT*CIN*9]
    Geom g;
};

template <int COUT> struct WgradMComp { static constexpr bool value = 2 * ((COUT + 7) / 8) <= 8 - 2; };
template <int CIN> struct WgradCat { static constexpr bool value = 2 * ((CIN + 7) / 8) * 8 * 9 <= 512; };

template <int CIN, int COUT>
__global__ void __launch_bounds__(NTHREADS, 1) conv_wgrad_tma_kernel(WgradArgs2 a) {
    constexpr int CINP = (CIN + 15) / 16 * 16;
    constexpr int NGRB = (CIN + 7) / 8;
    // NCATB (20 input channels): the lo planes of a_in follow its 3 hi planes directly, so that ONE N = 48 operand [x_hi | x_lo]
    // meets the stacked A = [dy_hi ; dy_lo]: all four piece products in a single M = 128 instruction per tap and k-step.  The
    // kernel is bound by the shared-memory operand reads (A = 4 KB per instruction), so instructions are what counts:
    // 5.5 KB per (tap, k-step) instead of 3 x 3 KB.  Nine taps x 2 x CINP columns do not fit TMEM for 40 channels.
    constexpr bool NCATB = WgradCat<CIN>::value;
    constexpr int NGB = NCATB ? NGRB : CINP / 8;
    constexpr int NACC = NCATB ? 2 * NGRB * 8 : CINP;             // accumulator columns per tap
    constexpr int NGRA = (COUT + 7) / 8;
    // MCOMP (20 output channels): both dy pieces fit ONE M = 64 operand -- rows 0..23 dy_hi, 24..47 dy_lo, 48..63 zero -- half the
    // A bytes of the M = 128 stacking (2 KB instead of 4 KB per instruction)
    constexpr bool MCOMP = WgradMComp<COUT>::value;
    constexpr int NGA = MCOMP ? NGRA : 8;                          // plane pitch between the dy pieces in a slot
    constexpr int SLOT_BYTES = 2 * NGB * PLANE_BYTES, ASLOT_BYTES = (MCOMP ? 8 : 16) * PLANE_BYTES;
    constexpr int TM_COLS = 512;
    static_assert(9 * NACC <= TM_COLS && COUT <= 64 && NACC % 16 == 0, "accumulators must fit TMEM");

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* win = smem;
    uint8_t* asm_ = smem + NWIN * SLOT_BYTES;
    __shared__ uint64_t full_bar[NWIN], empty_bar[NWIN], afull_bar[NASLOT], aempty_bar[NASLOT], done_bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NP = a.g.NP;

    for (int i = tid; i < (NWIN * SLOT_BYTES + NASLOT * ASLOT_BYTES) / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 4) {
        if (lane == 0) {
            for (int s = 0; s < NWIN; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
            for (int s = 0; s < NASLOT; ++s) { mbar_init(&afull_bar[s], 1); mbar_init(&aempty_bar[s], 1); }
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
    Walk walk;
    walk.init(a.g);
    Chunk ch;

    if (warp < 4) {
        // ================================================================================= final read-out
        mbar_wait(&done_bar, 0);
        tc_fence_after();
        const bool has_tiles = walk.pos < walk.end;                // a CTA without tiles never initialised its accumulators
        // one piece: M = 64, row r lives in lane (r%16) + 32*(r/16), the second partial row of this CTA is zero.
        // two pieces: M = 128, lane = row; rows 0..63 hold dy_hi * a_in, rows 64..127 dy_lo * a_in -> the two partial rows of this CTA
        // MCOMP: M = 64 with both pieces: row r = 16*warp + lane (lane < 16); rows 0..23 dy_hi channels, 24..47 dy_lo channels
        const bool stacked = MCOMP || NP > 1;
        const int row = MCOMP ? 16 * warp + lane : 32 * warp + lane;
        const int half = MCOMP ? (row >= 8 * NGRA ? 1 : 0) : (stacked ? (row >> 6) : 0);
        const int co = MCOMP ? (lane < 16 && row < 16 * NGRA ? row - half * 8 * NGRA : COUT) : (stacked ? (row & 63) : 16 * warp + lane);
        const bool writer = stacked ? true : lane < 16;
        float* out = a.partial + ((size_t)blockIdx.x * 2 + half) * COUT * CIN * 9;
        if (!stacked) {
            float* z = a.partial + ((size_t)blockIdx.x * 2 + 1) * COUT * CIN * 9;
            for (int i = tid; i < COUT * CIN * 9; i += 128) z[i] = 0.f;
        }
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
            if (NCATB && NP > 1) {                                 // columns [0, 24) = x_hi products, [24, 48) = x_lo products
                float v[NACC];
#pragma unroll
                for (int c0 = 0; c0 < NACC; c0 += 16)
                    tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(tap * NACC + c0), *reinterpret_cast<float(*)[16]>(&v[c0]));
                if (co < COUT) {
#pragma unroll
                    for (int i = 0; i < CIN; ++i) out[((size_t)co * CIN + i) * 9 + tap] = has_tiles ? v[i] + v[NACC / 2 + i] : 0.f;
                }
            } else {
#pragma unroll
                for (int c0 = 0; c0 < CINP; c0 += 16) {
                    float v[16];
                    tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(tap * NACC + c0), v);
                    if (writer && co < COUT) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (c0 + i < CIN) out[((size_t)co * CIN + c0 + i) * 9 + tap] = has_tiles ? v[i] : 0.f;
                    }
                }
            }
        }
        tc_fence_before();
    } else if (warp == 4) {
        // ================================================================================= MMA issuer (warp-uniform)
        // two pieces: A = [dy_hi ; dy_lo] stacked along M (the lo planes follow the hi planes in the slot), so that ONE M = 128
        // instruction per a_in piece yields dy_hi * x (rows 0..63) and dy_lo * x (rows 64..127): 2 full-rate MMAs per k-step
        // instead of 3 half-rate M = 64 ones (the lo * lo term comes for free).
        const uint32_t idesc = make_idesc((NP > 1 && !MCOMP) ? 128 : 64, (NCATB && NP > 1) ? NACC : CINP, 1, 1);
        const uint32_t win_base = smem_u32(win), a_base0 = smem_u32(asm_);
        uint32_t it = 0, wbase = 0;                                   // all waits of a tile first, then one elected block (see conv_tma_kernel)
        while (walk.next(ch)) {
            for (int k = 0; k < ch.n; ++k, ++it) {
                const int aslot = it % NASLOT;
                mbar_wait(&afull_bar[aslot], (it / NASLOT) & 1);
                const uint32_t wi0 = wbase + k;
                if (k == 0) {
                    mbar_wait(&full_bar[wi0 % NWIN], (wi0 / NWIN) & 1);
                    mbar_wait(&full_bar[(wi0 + 1) % NWIN], ((wi0 + 1) / NWIN) & 1);
                }
                mbar_wait(&full_bar[(wi0 + 2) % NWIN], ((wi0 + 2) / NWIN) & 1);
                tc_fence_after();
                const uint64_t dah0 = make_desc(a_base0 + aslot * ASLOT_BYTES + 16, 128, PLANE_BYTES);     // entries 1..128
                const bool last = (k == ch.n - 1);
                uint32_t sl[3];
                uint64_t dB[3];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    sl[ky] = (wi0 + ky) % NWIN;
                    dB[ky] = make_desc(win_base + sl[ky] * SLOT_BYTES, 128, PLANE_BYTES);
                }
                if (elect_one()) {
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const uint32_t slot = sl[ky];
                        const uint64_t dbh0 = dB[ky];
                        const uint64_t dbl0 = desc_advance(dbh0, NGB * PLANE_BYTES);
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const uint32_t d_tmem = tmem_base + (uint32_t)((ky * 3 + kx) * NACC);
#pragma unroll
                            for (int ks = 0; ks < BM / 16; ++ks) {
                                const uint64_t dah = desc_advance(dah0, ks * 256);
                                const uint64_t dbh = desc_advance(dbh0, (kx + 16 * ks) * 16), dbl = desc_advance(dbl0, (kx + 16 * ks) * 16);
                                tc_mma(d_tmem, dah, dbh, idesc, (it | (uint32_t)ks) != 0);
                                if (NP > 1 && !NCATB) tc_mma(d_tmem, dah, dbl, idesc, 1);
                            }
                        }
                        if (ky == 0 || last) tc_commit(&empty_bar[slot]);
                        if (ky == 2) tc_commit(&aempty_bar[aslot]);
                    }
                }
                __syncwarp();
            }
            wbase += ch.n + 2;
        }
        if (elect_one()) tc_commit(&done_bar);
        __syncwarp();
    } else if (lane == 0) {
        // ================================================================================= copy producer
        // order per chunk: [dy(0), in(-1), in(0), in(1)], [dy(1), in(2)], [dy(2), in(3)], ... = the MMA warp's wait order
        const uint32_t win_base = smem_u32(win), a_base0 = smem_u32(asm_);
        uint32_t wi = 0, it = 0;
        auto put_in = [&](int tp) {
            const int slot = wi % NWIN;
            mbar_wait(&empty_bar[slot], ((wi / NWIN) & 1) ^ 1);
            mbar_expect_tx(&full_bar[slot], (uint32_t)(NP * NGRB * PLANE_BYTES));
            for (int p = 0; p < NP; ++p)
                for (int grp = 0; grp < NGRB; ++grp)
                    bulk_g2s(win_base + slot * SLOT_BYTES + (p * NGB + grp) * PLANE_BYTES,
                             a.Pin + plane_unit(a.g, NGRB, ch.b, tp, p, grp, ch.fb * BM), PLANE_BYTES, &full_bar[slot]);
            ++wi;
        };
        while (walk.next(ch)) {
            for (int k = 0; k < ch.n; ++k, ++it) {
                const int aslot = it % NASLOT;
                mbar_wait(&aempty_bar[aslot], ((it / NASLOT) & 1) ^ 1);
                mbar_expect_tx(&afull_bar[aslot], (uint32_t)(NP * NGRA * PLANE_BYTES));
                for (int p = 0; p < NP; ++p)
                    for (int grp = 0; grp < NGRA; ++grp)
                        bulk_g2s(a_base0 + aslot * ASLOT_BYTES + (p * NGA + grp) * PLANE_BYTES,
                                 a.Pdy + plane_unit(a.g, NGRA, ch.b, ch.t0 + k + 1, p, grp, ch.fb * BM), PLANE_BYTES, &afull_bar[aslot]);
                if (k == 0) { put_in(ch.t0); put_in(ch.t0 + 1); }      // t' of input rows t0-1, t0
                put_in(ch.t0 + k + 2);                                  // t' of input row t0+k+1
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TM_COLS);
    }
}

template <int C, int MODE>
int launch_planes(cudaStream_t st, const PlaneArgs& a) {
    constexpr int NGR = (C + 7) / 8;
    dim3 grid((a.g.FP + 127) / 128, a.g.T + 2, a.g.B), block(128, NGR);
    planes_kernel<C, MODE><<<grid, block, 0, st>>>(a);
    PA2S_CHECK_LAST();
    return 0;
}
template <int CIN, int COUT>
int launch_conv(cudaStream_t st, const ConvArgs2& a) {
    constexpr int CINP = (CIN + 15) / 16 * 16, COUTP = (COUT + 15) / 16 * 16;
    constexpr int NG = CINP / 8, KS = CINP / 16;
    constexpr int W_BYTES = 9 * KS * 2 * (2 * COUTP * 16);
    constexpr int SMEM = ((W_BYTES + 127) / 128) * 128 + NWIN * (2 * NG * PLANE_BYTES) + 1024;
    static_assert(SMEM <= 227 * 1024, "shared memory");
    PA2S_TRY(cudaFuncSetAttribute(conv_tma_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    conv_tma_kernel<CIN, COUT><<<conv_grid(a.g.B, a.g.T, a.g.F), NTHREADS, SMEM, st>>>(a);
    PA2S_CHECK_LAST();
    return 0;
}
template <int CIN, int COUT>
int launch_conv3(cudaStream_t st, const ConvArgs2& a) {
    using Cfg = Conv3Cfg<CIN, COUT>;
    static_assert(Cfg::SMEM <= 227 * 1024 && Cfg::NW >= 5, "shared memory");
    PA2S_TRY(cudaFuncSetAttribute(conv_tma3_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    conv_tma3_kernel<CIN, COUT><<<conv_grid(a.g.B, a.g.T, a.g.F, a.g.nfbc), NT3, Cfg::SMEM, st>>>(a);
    PA2S_CHECK_LAST();
    return 0;
}
template <int CIN, int COUT>
int launch_wgrad(cudaStream_t st, const WgradArgs2& a) {
    constexpr int CINP = (CIN + 15) / 16 * 16;
    constexpr int NGB = WgradCat<CIN>::value ? (CIN + 7) / 8 : CINP / 8;
    constexpr int SMEM = NWIN * (2 * NGB * PLANE_BYTES) + NASLOT * ((WgradMComp<COUT>::value ? 8 : 16) * PLANE_BYTES) + 1024;
    static_assert(SMEM <= 227 * 1024, "shared memory");
    PA2S_TRY(cudaFuncSetAttribute(conv_wgrad_tma_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    conv_wgrad_tma_kernel<CIN, COUT><<<conv_grid(a.g.B, a.g.T, a.g.F), NTHREADS, SMEM, st>>>(a);
    PA2S_CHECK_LAST();
    return 0;
}
Geom make_geom(int B, int T, int F, int npieces) {
    Geom g;
    g.B = B; g.T = T; g.F = F; g.nfb = geom_nfb(F); g.FP = geom_fp(F); g.NP = npieces >= 2 ? 2 : 1; g.nfbc = geom_nfbc(F);
    return g;
}

}  // namespace

// bytes of the plane tensor of a (B,T,F,C) activation with `npieces` (1 or 2) bf16 pieces
PA2S_API long long pa2s_planes_bytes(int B, int T, int F, int C, int npieces) {
    return (long long)B * (T + 2) * (npieces >= 2 ? 2 : 1) * ((C + 7) / 8) * geom_fp(F) * 16;
}
// planes of relu?(X*scale + shift) (scale NULL: identity): the BatchNorm-apply + ReLU between two convolutions (models.py:525-534)
PA2S_API int pa2s_planes_fwd(void* stream, int B, int T, int F, int C, const float* X, const float* scale, const float* shift, int relu,
                             void* planes, int npieces) {
    PlaneArgs a;
    a.X = X; a.Yraw = nullptr; a.P = (uint4*)planes; a.g = make_geom(B, T, F, npieces);
    a.c0 = scale; a.c1 = shift; a.c2 = a.c3 = a.c4 = a.c5 = a.c6 = nullptr; a.relu = relu; a.low = 0;
    if (C == 20) return launch_planes<20, 0>((cudaStream_t)stream, a);
    if (C == 40) return launch_planes<40, 0>((cudaStream_t)stream, a);
    return -1;
}
// the same activation as its two LOWER bf16 pieces (p3, p2) of the three-way split x = p1 + p2 + p3 (pa2s_planes_fwd writes (p1, p2)):
// the operand of the third pass of the three-piece convolution of the exact-fp32 (eval) mode, see pa2s_conv_tma_acc
PA2S_API int pa2s_planes_fwd_low(void* stream, int B, int T, int F, int C, const float* X, const float* scale, const float* shift, int relu,
                                 void* planes) {
    PlaneArgs a;
    a.X = X; a.Yraw = nullptr; a.P = (uint4*)planes; a.g = make_geom(B, T, F, 2);
    a.c0 = scale; a.c1 = shift; a.c2 = a.c3 = a.c4 = a.c5 = a.c6 = nullptr; a.relu = relu; a.low = 1;
    if (C == 20) return launch_planes<20, 0>((cudaStream_t)stream, a);
    if (C == 40) return launch_planes<40, 0>((cudaStream_t)stream, a);
    return -1;
}
// planes of dy = k1*(G*(Yraw*zs+zb > 0) - k2 - (Yraw-mean)*invstd*k3): BatchNorm + ReLU backward (autograd of models.py:525-534)
PA2S_API int pa2s_planes_bwd(void* stream, int B, int T, int F, int C, const float* G, const float* Yraw, const float* zs, const float* zb,
                             const float* mean, const float* invstd, const float* k1, const float* k2, const float* k3,
                             void* planes, int npieces) {
    PlaneArgs a;
    a.low = 0;
    a.X = G; a.Yraw = Yraw; a.P = (uint4*)planes; a.g = make_geom(B, T, F, npieces);
    a.c0 = zs; a.c1 = zb; a.c2 = mean; a.c3 = invstd; a.c4 = k1; a.c5 = k2; a.c6 = k3; a.relu = 1;
    if (C == 20) return launch_planes<20, 1>((cudaStream_t)stream, a);
    if (C == 40) return launch_planes<40, 1>((cudaStream_t)stream, a);
    return -1;
}
static int g_conv_impl = 1;       // 1: conv_tma3_kernel (kx taps share the A read), 0: conv_tma_kernel (one instruction group per tap)
PA2S_API int pa2s_conv_tma_set_impl(int impl) { g_conv_impl = impl ? 1 : 0; return 0; }
PA2S_API int pa2s_conv_tma_get_impl(void) { return g_conv_impl; }
PA2S_API int pa2s_conv_tma_num_partials(int B, int T, int F) {
    return g_conv_impl ? conv_grid(B, T, F, geom_nfbc(F)) * 8 : conv_grid(B, T, F) * 4;      // every row is written by the kernel
}
PA2S_API int pa2s_conv_tma_wgrad_num_partials(int B, int T, int F) { return 2 * conv_grid(B, T, F); }   // two rows per CTA
// Y (B,T,F,Cout) = conv3x3 of the planes tensor (Cin channels) with Wpack (pa2s_tc_conv_pack: dgrad = 0 the forward filter,
// dgrad = 1 the flipped / transposed filter, which makes this the data gradient: planes = dy, Cin = channels of dy).
// partial (or NULL): pa2s_conv_tma_num_partials rows of [sum y, sum y^2].
static int conv_tma_launch(void* stream, int B, int T, int F, int Cin, int Cout, const ConvArgs2& a);
PA2S_API int pa2s_conv_tma(void* stream, int B, int T, int F, int Cin, int Cout, const void* planes, int npieces, const void* Wpack,
                           float* Y, float* partial) {
    ConvArgs2 a = {};
    a.P = (const uint4*)planes; a.Wpack = (const uint4*)Wpack; a.Y = Y; a.partial = partial; a.g = make_geom(B, T, F, npieces);
    return conv_tma_launch(stream, B, T, F, Cin, Cout, a);
}
// The data gradient of a layer fused with the statistics pass of the BatchNorm/ReLU backward of the layer below (what
// pa2s_colstats(mode 1, X = Yraw, G = Y) computes): partial = pa2s_conv_tma_num_partials rows of [sum g, sum g * xhat] over Cout
// channels.  Yraw: raw convolution output of the layer below, (B,T,F,Cout) like Y; zs, zb: its BatchNorm scale / shift (ReLU mask),
// mean, invstd: its batch statistics.  Only with the conv_tma3 kernel (pa2s_conv_tma_set_impl(1), the default): -2 otherwise.
PA2S_API int pa2s_conv_tma_dgrad_stats(void* stream, int B, int T, int F, int Cin, int Cout, const void* planes, int npieces,
                                       const void* Wpack, float* Y, const float* Yraw, const float* zs, const float* zb,
                                       const float* mean, const float* invstd, float* partial) {
    if (!g_conv_impl) return -2;
    ConvArgs2 a = {};
    a.P = (const uint4*)planes; a.Wpack = (const uint4*)Wpack; a.Y = Y; a.partial = partial; a.g = make_geom(B, T, F, npieces);
    a.sY = Yraw; a.szs = zs; a.szb = zb; a.smu = mean; a.sis = invstd;
    return conv_tma_launch(stream, B, T, F, Cin, Cout, a);
}
// Y += conv3x3(planes, Wpack): the second and third pass of the three-piece (fp32-level) convolution of the eval mode:
//   pass 1  pa2s_conv_tma      planes (p1, p2), pack sel 0 (W1, W2):  a1 W1 + a1 W2 + a2 W1
//   pass 2  pa2s_conv_tma_acc  planes (p1, p2), pack sel 1 (W3, 0):   a1 W3 + a2 W3
//   pass 3  pa2s_conv_tma_acc  planes (p3, p2) (pa2s_planes_fwd_low), pack sel 2 (W2, W1):  a3 W2 + a3 W1 + a2 W2
// = every product of the pieces of a = a1 + a2 + a3 and W = W1 + W2 + W3 except a3 W3 (2^-32): the accuracy of an fp32 FMA chain.
PA2S_API int pa2s_conv_tma_acc(void* stream, int B, int T, int F, int Cin, int Cout, const void* planes, const void* Wpack, float* Y) {
    if (!g_conv_impl) return -2;
    ConvArgs2 a = {};
    a.P = (const uint4*)planes; a.Wpack = (const uint4*)Wpack; a.Y = Y; a.partial = nullptr; a.g = make_geom(B, T, F, 2);
    a.accumulate = 1;
    return conv_tma_launch(stream, B, T, F, Cin, Cout, a);
}
static int conv_tma_launch(void* stream, int B, int T, int F, int Cin, int Cout, const ConvArgs2& a) {
    cudaStream_t st = (cudaStream_t)stream;
    if (g_conv_impl) {
        if (Cin == 20 && Cout == 20) return launch_conv3<20, 20>(st, a);
        if (Cin == 20 && Cout == 40) return launch_conv3<20, 40>(st, a);
        if (Cin == 40 && Cout == 40) return launch_conv3<40, 40>(st, a);
        if (Cin == 40 && Cout == 20) return launch_conv3<40, 20>(st, a);
        return -1;
    }
    if (Cin == 20 && Cout == 20) return launch_conv<20, 20>(st, a);
    if (Cin == 20 && Cout == 40) return launch_conv<20, 40>(st, a);
    if (Cin == 40 && Cout == 40) return launch_conv<40, 40>(st, a);
    if (Cin == 40 && Cout == 20) return launch_conv<40, 20>(st, a);
    return -1;
}
#ifdef PA2S_CONV_PROF
PA2S_API int pa2s_conv_tma_prof_read(unsigned long long* host_out) {
    PA2S_TRY(cudaDeviceSynchronize());
    PA2S_TRY(cudaMemcpyFromSymbol(host_out, g_conv_prof, sizeof(g_conv_prof)));
    return 0;
}
#endif
// partial: pa2s_conv_tma_wgrad_num_partials rows of Cout*Cin*9 in torch (Cout,Cin,3,3) order.
PA2S_API int pa2s_conv_tma_wgrad(void* stream, int B, int T, int F, int Cin, int Cout, const void* planes_in, const void* planes_dy,
                                 int npieces, float* partial) {
    WgradArgs2 a;
    a.Pin = (const uint4*)planes_in; a.Pdy = (const uint4*)planes_dy; a.partial = partial; a.g = make_geom(B, T, F, npieces);
    cudaStream_t st = (cudaStream_t)stream;
    if (Cin == 20 && Cout == 20) return launch_wgrad<20, 20>(st, a);
    if (Cin == 20 && Cout == 40) return launch_wgrad<20, 40>(st, a);
    if (Cin == 40 && Cout == 40) return launch_wgrad<40, 40>(st, a);
    return -1;
}
