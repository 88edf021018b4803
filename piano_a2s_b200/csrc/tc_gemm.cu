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
#include "common.cuh"
#include <cuda_bf16.h>

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

struct TcArgs {
    const float* A; const float* B; float* C; const float* bias;
    int M, N, K, BN;               // BN = UMMA N of this launch (multiple of 16, <= 256)
    long long lda, ldb, ldc, sA, sB, sC;
    int batch, splitk, kchunk, tiles_m, tiles_n;
    int accumulate, atomic, nsplit;
    const float* t_scale; const float* t_shift; int t_period; int t_relu; int t_on_b;
    int vecA, vecB, vecC;
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor, no swizzle (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=0 [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16, BF16 x BF16 -> F32.
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                       // c_format = F32
    d |= 1u << 7;                       // a_format = BF16
    d |= 1u << 10;                      // b_format = BF16
    d |= (uint32_t)a_mn_major << 15;
    d |= (uint32_t)b_mn_major << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
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

// Loads 8 consecutive fp32 at P[idx0 .. idx0+8) (guarded by `limit` on the contiguous index c0+i), applies the
// optional affine/ReLU keyed on the contiguous index.
__device__ __forceinline__ void load8(const float* __restrict__ P, bool row_ok, long long off, int c0, int limit, bool vec, bool tf,
                                      const TcArgs& g, float (&x)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 0.f;
    if (!row_ok) return;
    if (vec && c0 + 7 < limit) {
        float4 a = __ldg(reinterpret_cast<const float4*>(P + off));
        float4 b = __ldg(reinterpret_cast<const float4*>(P + off) + 1);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (c0 + i < limit) x[i] = __ldg(P + off + i);
    }
    if (tf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (c0 + i < limit) {
                int c = (c0 + i) % g.t_period;
                float y = fmaf(x[i], __ldg(g.t_scale + c), __ldg(g.t_shift + c));
                x[i] = g.t_relu ? fmaxf(y, 0.f) : y;
            }
        }
    }
}

// Fill one operand's hi/lo planes for one stage.
//  KMAJ source: P[mn][k]   -> unit (mn row r, k-group kg): smem (kg*ROWS + r)*16
//  MN   source: P[k][mn]   -> unit (k row kk, mn-group mg): smem (mg*BK + kk)*16
template <bool KMAJ>
__device__ __forceinline__ void load_operand(const float* __restrict__ P, long long ld, int mn0, int MN, int rows, int k0, int kend,
                                             bool vec, bool tf, const TcArgs& g, uint8_t* hi_plane, uint8_t* lo_plane, int ltid,
                                             bool want_lo) {
    const int lane = ltid & 31, lw = ltid >> 5;
    const int i8 = lane & 7, g4 = lane >> 3;
    if (KMAJ) {
        // blocks of 8 rows x 4 k-groups per warp
        for (int rb = lw; rb < rows / 8; rb += N_LOAD_WARPS) {
            const int r = rb * 8 + i8, kg = g4;
            const int m = mn0 + r, k = k0 + kg * 8;
            float x[8];
            load8(P, m < MN, (long long)m * ld + k, k, kend, vec, tf, g, x);
            uint4 hi, lo;
            split8(x, hi, lo);
            const int off = (kg * rows + r) * 16;
            *reinterpret_cast<uint4*>(hi_plane + off) = hi;
            if (want_lo) *reinterpret_cast<uint4*>(lo_plane + off) = lo;
        }
    } else {
        // blocks of 8 k-rows x 4 mn-groups per warp; (BK/8) * (rows/32) blocks
        const int nblk = (BK / 8) * (rows / 32);
        for (int blk = lw; blk < nblk; blk += N_LOAD_WARPS) {
            const int kb = blk % (BK / 8), mb = blk / (BK / 8);
            const int kk = kb * 8 + i8, mg = mb * 4 + g4;
            const int k = k0 + kk, m = mn0 + mg * 8;
            float x[8];
            load8(P, k < kend, (long long)k * ld + m, m, MN, vec, tf, g, x);
            uint4 hi, lo;
            split8(x, hi, lo);
            const int off = (mg * BK + kk) * 16;
            *reinterpret_cast<uint4*>(hi_plane + off) = hi;
            if (want_lo) *reinterpret_cast<uint4*>(lo_plane + off) = lo;
        }
    }
}

template <bool A_KMAJ, bool B_KMAJ>
__global__ void __launch_bounds__(NTHREADS, 1) tc_gemm_kernel(TcArgs g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full_bar[NSTAGE], empty_bar[NSTAGE], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int BN = g.BN;

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
        // ===================================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc(BM, BN, A_KMAJ ? 0 : 1, B_KMAJ ? 0 : 1);
            const uint32_t smem_base = smem_u32(smem);
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
                    const uint32_t a_hi = smem_base + s * STAGE_BYTES, a_lo = a_hi + A_PLANE_BYTES;
                    const uint32_t b_hi = a_hi + 2 * A_PLANE_BYTES, b_lo = b_hi + B_PLANE_BYTES;
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
                        // K-major operand: two k-groups (planes) per MMA, plane pitch rows*16; MN-major: 16 k rows at 16 B
                        const uint32_t a_off = A_KMAJ ? (uint32_t)(2 * ks * BM * 16) : (uint32_t)(ks * 256);
                        const uint32_t b_off = B_KMAJ ? (uint32_t)(2 * ks * BN * 16) : (uint32_t)(ks * 256);
                        const uint32_t a_lbo = A_KMAJ ? BM * 16 : 128, a_sbo = A_KMAJ ? 128 : BK * 16;
                        const uint32_t b_lbo = B_KMAJ ? (uint32_t)BN * 16 : 128, b_sbo = B_KMAJ ? 128 : BK * 16;
                        const uint64_t dah = make_desc(a_hi + a_off, a_lbo, a_sbo), dal = make_desc(a_lo + a_off, a_lbo, a_sbo);
                        const uint64_t dbh = make_desc(b_hi + b_off, b_lbo, b_sbo), dbl = make_desc(b_lo + b_off, b_lbo, b_sbo);
                        tc_mma(d_tmem, dah, dbh, idesc, (kb | ks) != 0);
                        if (g.nsplit > 1) {
                            tc_mma(d_tmem, dah, dbl, idesc, 1);
                            tc_mma(d_tmem, dal, dbh, idesc, 1);
                        }
                    }
                    tc_commit(&empty_bar[s]);
                }
                tc_commit(&tfull_bar[acc]);
            }
        }
    } else {
        // ===================================================================== operand loaders
        const int ltid = tid - (N_EPI_WARPS + 1) * 32;
        const bool tfA = g.t_scale != nullptr && !g.t_on_b, tfB = g.t_scale != nullptr && g.t_on_b;
        const bool want_lo = g.nsplit > 1;
        uint32_t kit = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int tn = (int)(tile % g.tiles_n);
            const int tm = (int)((tile / g.tiles_n) % g.tiles_m);
            const int z = (int)(tile / ((long long)g.tiles_n * g.tiles_m));
            const int bz = z / g.splitk, sk = z % g.splitk;
            const int kbeg = sk * g.kchunk, kend = min(g.K, kbeg + g.kchunk);
            const int nkb = (kend - kbeg + BK - 1) / BK;
            const float* A = g.A + (long long)bz * g.sA;
            const float* B = g.B + (long long)bz * g.sB;
            for (int kb = 0; kb < nkb; ++kb, ++kit) {
                const int s = kit % NSTAGE;
                mbar_wait(&empty_bar[s], ((kit / NSTAGE) & 1) ^ 1);
                uint8_t* st = smem + (size_t)s * STAGE_BYTES;
                const int k0 = kbeg + kb * BK;
                load_operand<A_KMAJ>(A, g.lda, tm * BM, g.M, BM, k0, kend, g.vecA, tfA, g, st, st + A_PLANE_BYTES, ltid, want_lo);
                load_operand<B_KMAJ>(B, g.ldb, tn * BN, g.N, BN, k0, kend, g.vecB, tfB, g, st + 2 * A_PLANE_BYTES,
                                     st + 2 * A_PLANE_BYTES + B_PLANE_BYTES, ltid, want_lo);
                fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor-core (async) proxy
                mbar_arrive(&full_bar[s]);
            }
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
