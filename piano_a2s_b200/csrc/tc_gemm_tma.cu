// TMA-fed tcgen05 GEMM on pre-split bf16 operands, fp32 accumulation in TMEM, fp32 output.
//
//   C[b] (M x N, fp32) (+)= sum over terms (i, j), i + j <= max(nA, nB) - 1, of  A_i[b] (M x K) * B_j[b]^T (N x K)
//
// A fp32 operand x is represented by 1..3 bf16 "pieces" x ~ p0 + p1 (+ p2) (pa2s_split_bf16): 2 pieces + the 3 terms
// p0*q0 + p0*q1 + p1*q0 reproduce an fp32 product to ~2^-17 ("bf16x3"), 3 pieces + 6 terms to ~2^-24, 1 piece is plain
// bf16.  Pieces live in global memory as bf16 matrices and are moved by TMA only (cp.async.bulk.tensor, 128-byte
// swizzle) -- no thread touches operand data:
//   K-major  operand: rows = M (or N), inner (contiguous) index = k;  box 64 k x 128 (or BN) rows
//   MN-major operand: rows = k, inner index = m (or n);               boxes 64 mn x 64 k, one per 64 columns of the tile
// so all four transposition cases of BLAS are fed without a transposing copy (UMMA descriptors: SWIZZLE_128B, K-major
// SBO = 1024 B; MN-major LBO = 8192 B (next 64-wide column block), SBO = 1024 B (next 8 k)).
// Rows may overlap in memory (row pitch < row length): with pitch = hop this is the VQT filterbank contraction over
// framed audio (utilities.py:246) with the frame matrix never materialised.
//
// One persistent CTA per SM, 192 threads: warps 0-3 epilogue (tcgen05.ld, one TMEM lane = one output row per thread),
// warp 4 TMEM allocation + single-lane tcgen05.mma issue, warp 5 single-lane TMA producer.  Pipelines: smem ring
// full/empty (TMA <-> MMA), two TMEM accumulators tfull/tempty (MMA <-> epilogue).
#include "tc_common.cuh"
#include <cuda.h>

namespace {
using namespace tc;

constexpr int BM = 128, BK = 64, BNMAX = 256;
constexpr int NSTAGE_MAX = 8;
constexpr int NTHREADS = 192;
constexpr int A_PIECE_BYTES = BM * BK * 2;          // 16 KB
constexpr int SMEM_BUDGET = 227 * 1024 - 2048;      // dynamic shared memory available to the stage ring

struct TmaGemmArgs {
    float* C; const float* bias;
    int M, N, K, BN;
    long long ldc, sC;
    int batch, splitk, kchunk, tiles_m, tiles_n;
    int atomic;
    int nA, nB, a_mn, b_mn, nstage;
    int a_bz, b_bz;                 // 0: the operand is shared by all batches
    uint32_t stage_bytes, b_piece_bytes;
    // complex-magnitude epilogue (the VQT filterbank, utilities.py:246-253): columns are (re, im) pairs; instead of C the kernel writes
    // mag[bz][m][n/2] = |re + i im| and folds the per-batch maximum into clip_max[bz] (bit pattern of a non-negative float)
    float* mag; unsigned int* clip_max; const int* valid_rows; long long sMag;
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* tm, uint32_t dst, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// SWIZZLE_128B shared-memory matrix descriptor
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return make_desc(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)2 << 61);
}

struct TileCoord { int tm, tn, bz, sk, kbeg, nkb; };
__device__ __forceinline__ TileCoord decode_tile(long long tile, const TmaGemmArgs& g) {
    TileCoord c;
    c.tn = (int)(tile % g.tiles_n);
    c.tm = (int)((tile / g.tiles_n) % g.tiles_m);
    const int z = (int)(tile / ((long long)g.tiles_n * g.tiles_m));
    c.bz = z / g.splitk; c.sk = z % g.splitk;
    c.kbeg = c.sk * g.kchunk;
    const int kend = min(g.K, c.kbeg + g.kchunk);
    c.nkb = (kend - c.kbeg + BK - 1) / BK;
    return c;
}

__global__ void __launch_bounds__(NTHREADS, 1) tc_gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmB, TmaGemmArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t full_bar[NSTAGE_MAX], empty_bar[NSTAGE_MAX], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int BN = g.BN;

    if (warp == 4) {
        if (lane == 0) {
            for (int s = 0; s < g.nstage; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
            for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 128); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        tmem_alloc(&tmem_base_s, 512);
    } else if (warp == 5 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const long long ntiles = (long long)g.batch * g.splitk * g.tiles_m * g.tiles_n;

    if (warp < 4) {
        // ===================================================================== epilogue
        uint32_t it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const TileCoord c = decode_tile(tile, g);
            const int acc = it & 1;
            mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
            tc_fence_after();
            const int m = c.tm * BM + warp * 32 + lane;
            float* crow = g.C + (long long)c.bz * g.sC + (long long)m * g.ldc;
            const bool use_atomic = g.atomic || g.splitk > 1;
            const bool vec = !use_atomic && (g.ldc % 4 == 0) && (g.sC % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
            if (g.mag != nullptr) {
                // fused |.| + per-clip maximum: rows past the clip's own frames (valid_rows) are written as zeros and stay out of the max
                const bool row_ok = m < g.M && (g.valid_rows == nullptr || m < __ldg(g.valid_rows + c.bz));
                float* mrow = g.mag + (long long)c.bz * g.sMag + (long long)m * (g.N / 2);
                float rmax = 0.f;
                for (int c0 = 0; c0 < BN; c0 += 16) {
                    const int n0 = c.tn * BN + c0;
                    if (n0 >= g.N) break;                               // warp-uniform
                    float v[16];
                    tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * BNMAX + c0), v);
                    if (m >= g.M) continue;
                    float mg[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        mg[i] = row_ok ? sqrtf(v[2 * i] * v[2 * i] + v[2 * i + 1] * v[2 * i + 1]) : 0.f;
                        if (n0 + 2 * i < g.N) rmax = fmaxf(rmax, mg[i]);
                    }
                    if (n0 + 15 < g.N) {
                        reinterpret_cast<float4*>(mrow + n0 / 2)[0] = make_float4(mg[0], mg[1], mg[2], mg[3]);
                        reinterpret_cast<float4*>(mrow + n0 / 2)[1] = make_float4(mg[4], mg[5], mg[6], mg[7]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) if (n0 + 2 * i < g.N) mrow[n0 / 2 + i] = mg[i];
                    }
                }
                rmax = warp_max(rmax);                                  // (a tile never straddles two batches: bz is tile-uniform)
                if (lane == 0 && rmax > 0.f) atomicMax(g.clip_max + c.bz, __float_as_uint(rmax));
                tc_fence_before();
                mbar_arrive(&tempty_bar[acc]);
                continue;
            }
            for (int c0 = 0; c0 < BN; c0 += 16) {
                const int n0 = c.tn * BN + c0;
                if (n0 >= g.N) break;                                   // warp-uniform
                float v[16];
                tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * BNMAX + c0), v);
                if (m >= g.M) continue;
                if (g.bias != nullptr && c.sk == 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) if (n0 + i < g.N) v[i] += __ldg(g.bias + n0 + i);
                }
                if (vec && n0 + 15 < g.N) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        reinterpret_cast<float4*>(crow + n0)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (n0 + i < g.N) {
                            if (use_atomic) atomicAdd(crow + n0 + i, v[i]);
                            else crow[n0 + i] = v[i];
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
        }
    } else if (warp == 4) {
        // ===================================================================== MMA issuer (warp-uniform control flow)
        const uint32_t idesc = make_idesc(BM, BN, g.a_mn, g.b_mn);
        const uint32_t a_lbo = g.a_mn ? 8192u : 16u, b_lbo = g.b_mn ? 8192u : 16u;
        const uint32_t a_step = g.a_mn ? 2048u : 32u, b_step = g.b_mn ? 2048u : 32u;     // bytes per UMMA_K = 16
        const int order = (g.nA > g.nB ? g.nA : g.nB) - 1;
        uint32_t it = 0, kit = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const TileCoord c = decode_tile(tile, g);
            const int acc = it & 1;
            mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BNMAX);
            for (int kb = 0; kb < c.nkb; ++kb, ++kit) {
                const int s = kit % g.nstage;
                mbar_wait(&full_bar[s], (kit / g.nstage) & 1);
                tc_fence_after();
                const uint32_t st = smem_base + s * g.stage_bytes;
                const uint64_t da0 = make_desc_sw128(st, a_lbo, 1024);
                const uint64_t db0 = make_desc_sw128(st + g.nA * A_PIECE_BYTES, b_lbo, 1024);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
                        for (int i = 0; i < g.nA; ++i) {
                            const uint64_t da = desc_advance(da0, i * A_PIECE_BYTES + ks * a_step);
                            for (int j = 0; j < g.nB && i + j <= order; ++j) {
                                const uint64_t db = desc_advance(db0, j * g.b_piece_bytes + ks * b_step);
                                tc_mma(d_tmem, da, db, idesc, (kb | ks | i | j) != 0);
                            }
                        }
                    }
                    tc_commit(&empty_bar[s]);
                    if (kb == c.nkb - 1) tc_commit(&tfull_bar[acc]);
                }
                __syncwarp();
            }
        }
    } else if (lane == 0) {
        // ===================================================================== TMA producer (one thread)
        uint32_t kit = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const TileCoord c = decode_tile(tile, g);
            for (int kb = 0; kb < c.nkb; ++kb, ++kit) {
                const int s = kit % g.nstage;
                mbar_wait(&empty_bar[s], ((kit / g.nstage) & 1) ^ 1);
                mbar_expect_tx(&full_bar[s], g.stage_bytes);
                const uint32_t st = smem_base + s * g.stage_bytes;
                const int k0 = c.kbeg + kb * BK;
                for (int i = 0; i < g.nA; ++i) {
                    const uint32_t dst = st + i * A_PIECE_BYTES;
                    if (!g.a_mn) tma_load_4d(&tmA, dst, &full_bar[s], k0, c.tm * BM, i, c.bz * g.a_bz);
                    else
                        for (int j = 0; j < BM / 64; ++j) tma_load_4d(&tmA, dst + j * 8192, &full_bar[s], c.tm * BM + j * 64, k0, i, c.bz * g.a_bz);
                }
                for (int i = 0; i < g.nB; ++i) {
                    const uint32_t dst = st + g.nA * A_PIECE_BYTES + i * g.b_piece_bytes;
                    if (!g.b_mn) tma_load_4d(&tmB, dst, &full_bar[s], k0, c.tn * BN, i, c.bz * g.b_bz);
                    else
                        for (int j = 0; j < BN / 64; ++j) tma_load_4d(&tmB, dst + j * 8192, &full_bar[s], c.tn * BN + j * 64, k0, i, c.bz * g.b_bz);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---- fp32 -> bf16 pieces ---------------------------------------------------------------------------------------------
struct SplitArgs {
    const float* src; __nv_bfloat16* dst;
    long long rows, cols, ld_src, bs_src, ld_dst, piece_stride, bs_dst;
    int npieces, batch;
    const float* t_scale; const float* t_shift; int t_period, t_relu;
    int vec;
};
__global__ void __launch_bounds__(256) split_bf16_kernel(SplitArgs a) {
    const long long groups = (a.cols + 7) / 8;
    const long long total = a.rows * groups;
    const int b = blockIdx.y;
    const float* src = a.src + (long long)b * a.bs_src;
    __nv_bfloat16* dst = a.dst + (long long)b * a.bs_dst;
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (long long)gridDim.x * blockDim.x) {
        const long long r = u / groups;
        const long long c0 = (u - r * groups) * 8;
        const float* p = src + r * a.ld_src + c0;
        float x[8];
        if (a.vec && c0 + 7 < a.cols) {
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(p)), v1 = __ldg(reinterpret_cast<const float4*>(p) + 1);
            x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = (c0 + i < a.cols) ? __ldg(p + i) : 0.f;
        }
        if (a.t_scale != nullptr) {
            int c = (int)(c0 % a.t_period);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (c0 + i < a.cols) {
                    const float y = fmaf(x[i], __ldg(a.t_scale + c), __ldg(a.t_shift + c));
                    x[i] = a.t_relu ? fmaxf(y, 0.f) : y;
                }
                c = (c + 1 == a.t_period) ? 0 : c + 1;
            }
        }
        __nv_bfloat16* q = dst + r * a.ld_dst + c0;
        uint4 hi, lo;
        split8_packed(x, hi, lo);
        *reinterpret_cast<uint4*>(q) = hi;
        if (a.npieces == 2) {
            *reinterpret_cast<uint4*>(q + a.piece_stride) = lo;
        } else if (a.npieces == 3) {
            // x = p0 + p1 + p2: p1 = bf16(x - p0), p2 = bf16(x - p0 - p1)
            float r1[8];
            const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                r1[2 * i] = x[2 * i] - __uint_as_float(h[i] << 16);
                r1[2 * i + 1] = x[2 * i + 1] - __uint_as_float(h[i] & 0xffff0000u);
            }
            uint4 mid, low;
            split8_packed(r1, mid, low);
            *reinterpret_cast<uint4*>(q + a.piece_stride) = mid;
            *reinterpret_cast<uint4*>(q + 2 * a.piece_stride) = low;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 4-D map (inner, rows, piece, batch) over a bf16 operand; box = (64, box_rows, 1, 1), 128-byte swizzle, zero fill.
int encode_operand(CUtensorMap* tm, const void* ptr, long long inner, long long rows, long long ld, int pieces, long long piece_stride,
                   int batch, long long batch_stride, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (enc == nullptr) return -2;
    if (ld % 8 != 0 || piece_stride % 8 != 0 || batch_stride % 8 != 0 || (reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return -3;
    cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)pieces, (cuuint64_t)batch};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)(pieces > 1 ? piece_stride : ld * rows) * 2,
                             (cuuint64_t)(batch > 1 ? batch_stride : ld * rows * pieces) * 2};
    if (strides[1] == 0) strides[1] = 16;
    if (strides[2] == 0) strides[2] = 16;
    cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -4;
}

}  // namespace

PA2S_API int pa2s_split_bf16(void* stream, const float* src, long long rows, long long cols, long long ld_src, long long batch_stride_src,
                             void* dst, long long ld_dst, long long piece_stride, long long batch_stride_dst, int npieces, int batch,
                             const float* t_scale, const float* t_shift, int t_period, int t_relu) {
    if (rows <= 0 || cols <= 0 || batch <= 0) return 0;
    if (npieces < 1 || npieces > 3 || ld_dst % 8 != 0 || piece_stride % 8 != 0 || batch_stride_dst % 8 != 0) return -1;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) != 0) return -1;
    SplitArgs a;
    a.src = src; a.dst = (__nv_bfloat16*)dst; a.rows = rows; a.cols = cols; a.ld_src = ld_src; a.bs_src = batch_stride_src;
    a.ld_dst = ld_dst; a.piece_stride = piece_stride; a.bs_dst = batch_stride_dst; a.npieces = npieces; a.batch = batch;
    a.t_scale = t_scale; a.t_shift = t_shift; a.t_period = t_period > 0 ? t_period : 1; a.t_relu = t_relu;
    a.vec = (ld_src % 4 == 0) && (batch_stride_src % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const long long total = rows * ((cols + 7) / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    split_bf16_kernel<<<dim3((unsigned)blocks, (unsigned)batch), 256, 0, (cudaStream_t)stream>>>(a);
    PA2S_CHECK_LAST();
    return 0;
}

// Operand X (A: rows of C, B: columns of C): bf16 pieces at x + piece * x_piece_stride + batch * x_batch_stride (elements);
// x_mn = 0: stored [mn][k] with row pitch x_ld (K-major); x_mn = 1: stored [k][mn] (MN-major).  Pitches are multiples of 8
// elements, base pointers 16-byte aligned; x_batch_stride = 0 shares the operand between batches.  `atomic` (or splitk > 1) accumulates into C with atomicAdd.
static int launch_tma_gemm(void* stream, int M, int N, int K,
                           const void* A, long long a_ld, long long a_piece_stride, long long a_batch_stride, int a_pieces, int a_mn,
                           const void* B, long long b_ld, long long b_piece_stride, long long b_batch_stride, int b_pieces, int b_mn,
                           float* C, long long ldc, long long strideC, const float* bias, int atomic, int batch, int splitk,
                           float* mag, unsigned int* clip_max, const int* valid_rows, long long strideMag) {
    if (M <= 0 || N <= 0 || batch <= 0) return 0;
    if (K <= 0 || a_pieces < 1 || a_pieces > 3 || b_pieces < 1 || b_pieces > 3) return -1;
    TmaGemmArgs g;
    g.mag = mag; g.clip_max = clip_max; g.valid_rows = valid_rows; g.sMag = strideMag;
    if (mag != nullptr && (splitk > 1 || atomic || (N & 15) != 0 || ((uintptr_t)mag & 15) != 0 || (strideMag & 3) != 0)) return -1;
    g.C = C; g.bias = bias; g.M = M; g.N = N; g.K = K; g.ldc = ldc; g.sC = strideC; g.batch = batch;
    g.nA = a_pieces; g.nB = b_pieces; g.a_mn = a_mn ? 1 : 0; g.b_mn = b_mn ? 1 : 0;
    // UMMA N: multiple of 16 (64 for an MN-major B, whose boxes are 64 columns wide), shrunk until two stages fit
    const int gran = g.b_mn ? 64 : 16;
    int BN = N >= BNMAX ? BNMAX : ((N + gran - 1) / gran) * gran;
    while (2 * (a_pieces * A_PIECE_BYTES + b_pieces * BN * 128) > SMEM_BUDGET && BN > gran) BN = ((BN / 2 + gran - 1) / gran) * gran;
    g.BN = BN;
    g.b_piece_bytes = (uint32_t)BN * 128u;
    g.stage_bytes = (uint32_t)(a_pieces * A_PIECE_BYTES) + (uint32_t)b_pieces * g.b_piece_bytes;
    int nstage = SMEM_BUDGET / (int)g.stage_bytes;
    if (nstage < 1) return -1;
    g.nstage = nstage > NSTAGE_MAX ? NSTAGE_MAX : nstage;
    if (splitk < 1) splitk = 1;
    int kchunk = ceil_div(ceil_div(K, splitk), BK) * BK;
    splitk = ceil_div(K, kchunk);
    g.splitk = splitk; g.kchunk = kchunk;
    g.tiles_m = ceil_div(M, BM); g.tiles_n = ceil_div(N, BN);
    g.atomic = (atomic || splitk > 1) ? 1 : 0;
    g.a_bz = (batch > 1 && a_batch_stride != 0) ? 1 : 0;
    g.b_bz = (batch > 1 && b_batch_stride != 0) ? 1 : 0;
    const int a_nb = g.a_bz ? batch : 1, b_nb = g.b_bz ? batch : 1;
    CUtensorMap tmA, tmB;
    int rc;
    if (!g.a_mn) rc = encode_operand(&tmA, A, K, M, a_ld, a_pieces, a_piece_stride, a_nb, a_batch_stride, BM);
    else rc = encode_operand(&tmA, A, M, K, a_ld, a_pieces, a_piece_stride, a_nb, a_batch_stride, 64);
    if (rc != 0) return rc;
    if (!g.b_mn) rc = encode_operand(&tmB, B, K, N, b_ld, b_pieces, b_piece_stride, b_nb, b_batch_stride, BN);
    else rc = encode_operand(&tmB, B, N, K, b_ld, b_pieces, b_piece_stride, b_nb, b_batch_stride, 64);
    if (rc != 0) return rc;
    const long long ntiles = (long long)batch * splitk * g.tiles_m * g.tiles_n;
    const int grid = (int)(ntiles < 148 ? ntiles : 148);
    const int smem = g.nstage * (int)g.stage_bytes + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        PA2S_TRY(cudaFuncSetAttribute(tc_gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET + 1024));
        attr_set = true;
    }
    tc_gemm_tma_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(tmA, tmB, g);
    PA2S_CHECK_LAST();
    return 0;
}

PA2S_API int pa2s_gemm_bf16_tma(void* stream, int M, int N, int K,
                                const void* A, long long a_ld, long long a_piece_stride, long long a_batch_stride, int a_pieces, int a_mn,
                                const void* B, long long b_ld, long long b_piece_stride, long long b_batch_stride, int b_pieces, int b_mn,
                                float* C, long long ldc, long long strideC, const float* bias, int atomic, int batch, int splitk) {
    return launch_tma_gemm(stream, M, N, K, A, a_ld, a_piece_stride, a_batch_stride, a_pieces, a_mn, B, b_ld, b_piece_stride, b_batch_stride,
                           b_pieces, b_mn, C, ldc, strideC, bias, atomic, batch, splitk, nullptr, nullptr, nullptr, 0);
}

// The same contraction with the VQT epilogue fused (utilities.py:246-253): the N columns are (re, im) pairs of N/2 bins; writes
// mag[batch][m][N/2] = |.| and atomically folds each batch's maximum into clip_max[batch] (zero it first).  valid_rows (int32 per batch or
// NULL): rows >= valid_rows[batch] are written as zeros and excluded from the maximum (frames the un-padded clip does not have).
PA2S_API int pa2s_gemm_bf16_tma_mag(void* stream, int M, int N, int K,
                                    const void* A, long long a_ld, long long a_piece_stride, long long a_batch_stride, int a_pieces, int a_mn,
                                    const void* B, long long b_ld, long long b_piece_stride, long long b_batch_stride, int b_pieces, int b_mn,
                                    float* mag, long long strideMag, unsigned int* clip_max, const int* valid_rows, int batch) {
    if (mag == nullptr || clip_max == nullptr) return -1;
    return launch_tma_gemm(stream, M, N, K, A, a_ld, a_piece_stride, a_batch_stride, a_pieces, a_mn, B, b_ld, b_piece_stride, b_batch_stride,
                           b_pieces, b_mn, nullptr, 0, 0, nullptr, 0, batch, 1, mag, clip_max, valid_rows, strideMag);
}
