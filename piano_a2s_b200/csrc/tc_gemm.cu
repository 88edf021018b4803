// tcgen05 (5th-gen tensor core) GEMM with fp32 in / fp32 out at ~fp32 accuracy: every fp32 operand element x is
// split on the fly into bf16 hi = rn(x), lo = rn(x - hi) and the product is accumulated in TMEM (fp32) as
//   A_hi*B_hi + A_hi*B_lo + A_lo*B_hi            ("bf16x3", relative error ~2^-17 per product),
// or as A_hi*B_hi only (nsplit = 1, the bf16 configuration of BASELINE configs[2..3]).
//
// Same contract as pa2s_gemm_f32 (include/pa2s.h): C = op(A) op(B) (+bias)(+C), batched / split-K, optional fused
// per-column affine+ReLU on one operand.  Warp-specialised persistent kernel, one CTA per SM:
//   warps 0-3  epilogue: tcgen05.ld the 128 x BN accumulator (one TMEM lane = one output row per thread) -> C
//   warp  4    allocates TMEM, initialises the mbarriers, one lane issues every tcgen05.mma / tcgen05.commit
//   warps 5-12 operand loaders: LDG.128 fp32 -> (BatchNorm affine+ReLU) -> bf16 hi/lo split -> 16-byte units stored
//              in the UMMA no-swizzle canonical layout ("planes" of 8 contiguous elements at a 16-byte pitch), so
//              a K-major source needs no transposition and an MN-major source is consumed as an MN-major operand.
// Shared-memory ring of 4 stages (BK = 32), two TMEM accumulator buffers (epilogue of tile i overlaps MMAs of i+1).
#include "tc_common.cuh"

namespace {

constexpr int BM = 128;            // UMMA M
constexpr int BNMAX = 256;         // max UMMA N / TMEM columns per accumulator buffer
constexpr int BK = 32;             // K per pipeline stage (4 groups of 8)
constexpr int NSTAGE = 4;
constexpr int N_EPI_WARPS = 4, N_LOAD_WARPS = 8;
constexpr int NTHREADS = (N_EPI_WARPS + 1 + N_LOAD_WARPS) * 32;      // 416
constexpr int N_LOAD_THREADS = N_LOAD_WARPS * 32;
// per stage: A hi, A lo (4 groups x 128 rows x 16 B), B hi, B lo (4 groups x BNMAX rows x 16 B)
constexpr int A_PLANE_BYTES = 4 * BM * 16;         // 8 KB  (one of hi / lo)
constexpr int B_PLANE_BYTES = 4 * BNMAX * 16;      // 16 KB
constexpr int STAGE_BYTES = 2 * A_PLANE_BYTES + 2 * B_PLANE_BYTES;   // 48 KB
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024;
constexpr int TPMAX = 256;        // max period of the fused operand affine (kept in shared memory)

struct TcArgs {
    const float* A; const float* B; float* C; const float* bias;
    int M, N, K, BN;               // BN = UMMA N of this launch (multiple of 16, <= 256)
    long long lda, ldb, ldc, sA, sB, sC;
    int batch, splitk, kchunk, tiles_m, tiles_n;
    int accumulate, atomic, nsplit;
    const float* t_scale; const float* t_shift; int t_period; int t_relu; int t_on_b;
    int vecA, vecB, vecC;
};

using namespace tc;

// ---- operand staging ------------------------------------------------------------------------------------------
// A loader thread owns, per pipeline stage, NIT "units" of one operand: 8 consecutive fp32 along the source's contiguous
// index.  Loads of a whole stage (both operands) are issued before any of them is consumed, and the loads of stage s+1
// are issued before stage s is converted, so the global/L2 latency overlaps the bf16 split work.
//  KMAJ source P[mn][k]: unit (row r, k-group kg)      -> smem (kg*rows + r)*16      (UMMA K-major, LBO = rows*16, SBO = 128)
//  MN   source P[k][mn]: unit (k row kk, mn-group mg)  -> smem (mg*BK + kk)*16       (UMMA MN-major, SBO = BK*16, LBO = 128)
template <int NIT>
struct Units {
    float4 v[NIT][2];
    int c0[NIT];            // contiguous-index of element 0 (transform key / guard), -1 = unit not loaded (zeros)
    int off[NIT];           // smem byte offset, -1 = unit not owned in this stage
};

template <bool KMAJ, int NIT>
__device__ __forceinline__ void units_issue(Units<NIT>& u, const float* __restrict__ P, long long ld, int mn0, int MN, int rows, int k0,
                                            int kend, bool vec, int ltid) {
    const int lane = ltid & 31, lw = ltid >> 5;
    const int i8 = lane & 7, g4 = lane >> 3;
    const int nblk = KMAJ ? rows / 8 : (BK / 8) * (rows / 32);
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int blk = lw + it * N_LOAD_WARPS;
        u.off[it] = -1; u.c0[it] = -1;
        u.v[it][0] = make_float4(0.f, 0.f, 0.f, 0.f); u.v[it][1] = u.v[it][0];
        if (blk >= nblk) continue;
        int slow, fast0, limit;     // slow index (row of the source), first contiguous index, limit of the contiguous index
        bool slow_ok;
        if (KMAJ) {
            const int r = blk * 8 + i8, kg = g4;
            slow = mn0 + r; slow_ok = slow < MN; fast0 = k0 + kg * 8; limit = kend;
            u.off[it] = (kg * rows + r) * 16;
        } else {
            const int kb = blk % (BK / 8), mb = blk / (BK / 8);
            const int kk = kb * 8 + i8, mg = mb * 4 + g4;
            slow = k0 + kk; slow_ok = slow < kend; fast0 = mn0 + mg * 8; limit = MN;
            u.off[it] = (mg * BK + kk) * 16;
        }
        if (!slow_ok || fast0 >= limit) continue;
        u.c0[it] = fast0;
        const float* p = P + (long long)slow * ld + fast0;
        if (vec && fast0 + 7 < limit) {
            u.v[it][0] = __ldg(reinterpret_cast<const float4*>(p));
            u.v[it][1] = __ldg(reinterpret_cast<const float4*>(p) + 1);
        } else {
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = (fast0 + i < limit) ? __ldg(p + i) : 0.f;
            u.v[it][0] = make_float4(x[0], x[1], x[2], x[3]);
            u.v[it][1] = make_float4(x[4], x[5], x[6], x[7]);
        }
    }
}

template <int NIT>
__device__ __forceinline__ void units_store(const Units<NIT>& u, int limit, bool tf, const TcArgs& g, const float* s_scale,
                                            const float* s_shift, uint8_t* hi_plane, uint8_t* lo_plane, bool want_lo) {
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        if (u.off[it] < 0) continue;
        float x[8] = {u.v[it][0].x, u.v[it][0].y, u.v[it][0].z, u.v[it][0].w, u.v[it][1].x, u.v[it][1].y, u.v[it][1].z, u.v[it][1].w};
        if (tf && u.c0[it] >= 0) {
            int c = u.c0[it] % g.t_period;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (u.c0[it] + i < limit) {
                    float y = fmaf(x[i], s_scale[c], s_shift[c]);
                    x[i] = g.t_relu ? fmaxf(y, 0.f) : y;
                }
                c = (c + 1 == g.t_period) ? 0 : c + 1;
            }
        }
        uint4 hi, lo;
        split8_packed(x, hi, lo);
        *reinterpret_cast<uint4*>(hi_plane + u.off[it]) = hi;
        if (want_lo) *reinterpret_cast<uint4*>(lo_plane + u.off[it]) = lo;
    }
}

// coordinates of one pipeline stage of this CTA's work list
struct StageCoord {
    long long tile; int kb, nkb, tm, tn, bz, kbeg, kend;
    bool valid;
};
__device__ __forceinline__ void coord_set_tile(StageCoord& c, long long tile, long long ntiles, const TcArgs& g) {
    c.tile = tile; c.valid = tile < ntiles; c.kb = 0;
    if (!c.valid) return;
    c.tn = (int)(tile % g.tiles_n);
    c.tm = (int)((tile / g.tiles_n) % g.tiles_m);
    const int z = (int)(tile / ((long long)g.tiles_n * g.tiles_m));
    c.bz = z / g.splitk;
    const int sk = z % g.splitk;
    c.kbeg = sk * g.kchunk; c.kend = min(g.K, c.kbeg + g.kchunk);
    c.nkb = (c.kend - c.kbeg + BK - 1) / BK;
}
__device__ __forceinline__ void coord_next(StageCoord& c, long long ntiles, const TcArgs& g) {
    if (++c.kb >= c.nkb) coord_set_tile(c, c.tile + gridDim.x, ntiles, g);
}

template <bool A_KMAJ, bool B_KMAJ>
__global__ void __launch_bounds__(NTHREADS, 1) tc_gemm_kernel(TcArgs g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full_bar[NSTAGE], empty_bar[NSTAGE], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_scale[TPMAX], s_shift[TPMAX];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int BN = g.BN;
    if (g.t_scale != nullptr)
        for (int i = tid; i < g.t_period; i += NTHREADS) { s_scale[i] = g.t_scale[i]; s_shift[i] = g.t_shift[i]; }

    if (warp == N_EPI_WARPS) {
        if (lane == 0) {
            for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], N_LOAD_THREADS); mbar_init(&empty_bar[s], 1); }
            for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], N_EPI_WARPS * 32); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    const int ktiles_total = g.splitk;                         // K chunks
    const long long ntiles = (long long)g.batch * g.splitk * g.tiles_m * g.tiles_n;
    (void)ktiles_total;

    if (warp < N_EPI_WARPS) {
        // ===================================================================== epilogue
        uint32_t it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int tn = (int)(tile % g.tiles_n);
            const int tm = (int)((tile / g.tiles_n) % g.tiles_m);
            const int z = (int)(tile / ((long long)g.tiles_n * g.tiles_m));
            const int bz = z / g.splitk, sk = z % g.splitk;
            const int acc = it & 1;
            mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
            tc_fence_after();
            const int m = tm * BM + warp * 32 + lane;
            float* crow = g.C + (long long)bz * g.sC + (long long)m * g.ldc;
            const bool use_atomic = g.atomic || g.splitk > 1;
            for (int c0 = 0; c0 < BN; c0 += 16) {
                float v[16];
                tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * BNMAX + c0), v);
                if (m < g.M && g.vecC && !use_atomic && !g.accumulate && tn * BN + c0 + 15 < g.N) {
                    const int n = tn * BN + c0;
                    if (g.bias != nullptr && sk == 0) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] += __ldg(g.bias + n + i);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        reinterpret_cast<float4*>(crow + n)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                } else if (m < g.M) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int n = tn * BN + c0 + i;
                        if (n < g.N) {
                            float o = v[i];
                            if (g.bias != nullptr && sk == 0) o += __ldg(g.bias + n);
                            if (use_atomic) atomicAdd(crow + n, o);
                            else if (g.accumulate) crow[n] += o;
                            else crow[n] = o;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
        }
    } else if (warp == N_EPI_WARPS) {
        // ===================================================================== MMA issuer (warp-uniform)
        {
            const uint32_t idesc = make_idesc(BM, BN, A_KMAJ ? 0 : 1, B_KMAJ ? 0 : 1);
            const uint32_t smem_base = smem_u32(smem);
            const uint32_t a_lbo = A_KMAJ ? BM * 16 : 128, a_sbo = A_KMAJ ? 128 : BK * 16;
            const uint32_t b_lbo = B_KMAJ ? (uint32_t)BN * 16 : 128, b_sbo = B_KMAJ ? 128 : BK * 16;
            const uint32_t a_step = A_KMAJ ? (uint32_t)(2 * BM * 16) : 256u;      // bytes per k-step of 16
            const uint32_t b_step = B_KMAJ ? (uint32_t)(2 * BN * 16) : 256u;
            uint32_t it = 0, kit = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int z = (int)(tile / ((long long)g.tiles_n * g.tiles_m));
                const int sk = z % g.splitk;
                const int kbeg = sk * g.kchunk, kend = min(g.K, kbeg + g.kchunk);
                const int nkb = (kend - kbeg + BK - 1) / BK;
                const int acc = it & 1;
                mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BNMAX);
                for (int kb = 0; kb < nkb; ++kb, ++kit) {
                    const int s = kit % NSTAGE;
                    mbar_wait(&full_bar[s], (kit / NSTAGE) & 1);
                    tc_fence_after();
                    const uint32_t st = smem_base + s * STAGE_BYTES;
                    const uint64_t dah0 = make_desc(st, a_lbo, a_sbo), dal0 = desc_advance(dah0, A_PLANE_BYTES);
                    const uint64_t dbh0 = make_desc(st + 2 * A_PLANE_BYTES, b_lbo, b_sbo), dbl0 = desc_advance(dbh0, B_PLANE_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < BK / 16; ++ks) {
                            const uint64_t dah = desc_advance(dah0, ks * a_step), dal = desc_advance(dal0, ks * a_step);
                            const uint64_t dbh = desc_advance(dbh0, ks * b_step), dbl = desc_advance(dbl0, ks * b_step);
                            tc_mma(d_tmem, dah, dbh, idesc, (kb | ks) != 0);
                            if (g.nsplit > 1) {
                                tc_mma(d_tmem, dah, dbl, idesc, 1);
                                tc_mma(d_tmem, dal, dbh, idesc, 1);
                            }
                        }
                        tc_commit(&empty_bar[s]);
                        if (kb == nkb - 1) tc_commit(&tfull_bar[acc]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===================================================================== operand loaders
        const int ltid = tid - (N_EPI_WARPS + 1) * 32;
        const bool tfA = g.t_scale != nullptr && !g.t_on_b, tfB = g.t_scale != nullptr && g.t_on_b;
        const bool want_lo = g.nsplit > 1;
        constexpr int NA = 2, NB = 4;                 // units per thread per stage (A: 128 rows, B: up to 256 rows)
        Units<NA> ua, ua_n;
        Units<NB> ub, ub_n;
        StageCoord cur, nxt;
        coord_set_tile(cur, blockIdx.x, ntiles, g);
        auto issue = [&](const StageCoord& c, Units<NA>& xa, Units<NB>& xb) {
            const float* A = g.A + (long long)c.bz * g.sA;
            const float* B = g.B + (long long)c.bz * g.sB;
            const int k0 = c.kbeg + c.kb * BK;
            units_issue<A_KMAJ, NA>(xa, A, g.lda, c.tm * BM, g.M, BM, k0, c.kend, g.vecA, ltid);
            units_issue<B_KMAJ, NB>(xb, B, g.ldb, c.tn * BN, g.N, BN, k0, c.kend, g.vecB, ltid);
        };
        if (cur.valid) issue(cur, ua, ub);
        uint32_t kit = 0;
        while (cur.valid) {
            nxt = cur;
            coord_next(nxt, ntiles, g);
            if (nxt.valid) issue(nxt, ua_n, ub_n);          // next stage's loads are in flight while this one is converted
            const int s = kit % NSTAGE;
            mbar_wait(&empty_bar[s], ((kit / NSTAGE) & 1) ^ 1);
            uint8_t* st = smem + (size_t)s * STAGE_BYTES;
            units_store<NA>(ua, A_KMAJ ? cur.kend : g.M, tfA, g, s_scale, s_shift, st, st + A_PLANE_BYTES, want_lo);
            units_store<NB>(ub, B_KMAJ ? cur.kend : g.N, tfB, g, s_scale, s_shift, st + 2 * A_PLANE_BYTES, st + 2 * A_PLANE_BYTES + B_PLANE_BYTES, want_lo);
            fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor-core (async) proxy
            mbar_arrive(&full_bar[s]);
            ua = ua_n; ub = ub_n; cur = nxt; ++kit;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == N_EPI_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

}  // namespace

// Returns 1 if (shape, layout) is handled by the tensor-core kernel, else 0 (caller uses pa2s_gemm_f32).
PA2S_API int pa2s_gemm_tc_supported(int M, int N, int K, int batch) {
    return (M >= 1 && N >= 1 && K >= 1 && batch >= 1) ? 1 : 0;
}

// Same argument contract as pa2s_gemm_f32 plus `nsplit` (3 = bf16x3 ~fp32 accuracy, 1 = plain bf16 operands).
PA2S_API int pa2s_gemm_tc(void* stream, int transA, int transB, int M, int N, int K,
                          const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc,
                          const float* bias, int accumulate, int atomic,
                          int batch, long long strideA, long long strideB, long long strideC,
                          const float* t_scale, const float* t_shift, int t_period, int t_relu, int t_on_b,
                          int splitk, int nsplit) {
    if (M <= 0 || N <= 0 || batch <= 0) return 0;
    if (K <= 0) return -1;
    if (t_scale != nullptr && t_period > TPMAX) return -1;
    TcArgs g;
    g.A = A; g.B = B; g.C = C; g.bias = bias; g.M = M; g.N = N; g.K = K;
    int BN = N >= BNMAX ? BNMAX : ((N + 31) / 32) * 32;      // multiple of 32 keeps both loader flavours whole
    g.BN = BN;
    g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.sA = strideA; g.sB = strideB; g.sC = strideC;
    g.batch = batch;
    if (splitk < 1) splitk = 1;
    if (splitk > 1) atomic = 1;
    int kchunk = ceil_div(ceil_div(K, splitk), BK) * BK;
    splitk = ceil_div(K, kchunk);
    g.splitk = splitk; g.kchunk = kchunk;
    g.tiles_m = ceil_div(M, BM); g.tiles_n = ceil_div(N, BN);
    g.accumulate = accumulate; g.atomic = atomic; g.nsplit = nsplit >= 3 ? 3 : 1;
    g.t_scale = t_scale; g.t_shift = t_shift; g.t_period = t_period > 0 ? t_period : 1; g.t_relu = t_relu; g.t_on_b = t_on_b;
    g.vecA = (lda % 4 == 0) && (strideA % 4 == 0) && ((uintptr_t)A % 16 == 0);
    g.vecB = (ldb % 4 == 0) && (strideB % 4 == 0) && ((uintptr_t)B % 16 == 0);
    g.vecC = (ldc % 4 == 0) && (strideC % 4 == 0) && ((uintptr_t)C % 16 == 0);
    long long ntiles = (long long)batch * splitk * g.tiles_m * g.tiles_n;
    int grid = (int)(ntiles < 148 ? ntiles : 148);
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(AK, BK_)                                                                                                      \
    do {                                                                                                                     \
        PA2S_TRY(cudaFuncSetAttribute(tc_gemm_kernel<AK, BK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));    \
        tc_gemm_kernel<AK, BK_><<<grid, NTHREADS, SMEM_BYTES, st>>>(g);                                                      \
    } while (0)
    if (!transA && transB) LAUNCH(true, true);
    else if (!transA && !transB) LAUNCH(true, false);
    else if (transA && transB) LAUNCH(false, true);
    else LAUNCH(false, false);
#undef LAUNCH
    PA2S_CHECK_LAST();
    return 0;
}
