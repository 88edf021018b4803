// Multi-sequence persistent note decoder (NoteDecoder.decode_notes, models.py:366-420, for several bars of a staff at once).
//
// dec_persist.cu runs ONE (bar, staff) per launch; its step is bound by streaming every clip's encoder memory (3.69 MB per
// clip and step) from L2 and by three grid barriers.  But bars whose input token comes from the ground truth (teacher-forced
// bars, models.py:289-311) do not depend on the previous bar's predictions, and in the reverse pass NO bar depends on another:
// here one cooperative launch decodes NQ bars x B clips = R rows.  All rows of a clip attend over the same memory, so ONE pass
// over a clip's frames serves NQ queries (bytes per bar-step / NQ), the grid barriers and per-step latencies are shared, and
// the weight-stationary GEMV phases see NQ x more rows per weight read.
//
//   A(s)  attention, item = (clip, frame range): pass 1 streams exp(2 Ep) rows (scores of all NQ queries, 1 MUFU per element:
//         tanh(q+e) = 1 - 2/(1 + exp(2q) exp(2e))), pass 2 streams enc rows in 128-column blocks (contexts); partials per item,
//         last-arriving CTA of a clip combines.  + D(s-1): log-softmax / argmax / teacher forcing / next token, one WARP per row.
//   B(s)  GRU cell 528 -> 512: each CTA owns 8 hidden units (24 gate rows x 1040 weights in shared memory), rows in chunks of 16
//   C(s)  logits and next query: 7 of the 429 rows per CTA
//
// The reverse kernel has the same shape (P1 gate gradients | P2 W^T products | P3 attention backward, two passes).
#include "decm_args.cuh"
#include <cooperative_groups.h>
#include <cuda_bf16.h>

namespace {

constexpr int PG = 64;                 // CTAs of a persistent decoder grid (two staves run concurrently: 128 of 148 SMs)
constexpr int NT = 384;
constexpr int NW = NT / 32;            // 12 warps
constexpr int UPC = DD / PG;           // hidden units per CTA (8)
constexpr int GR = 3 * UPC;            // gate rows per CTA (24)
constexpr int CR = 7;                  // phase-C rows per CTA: PG*CR = 448 >= V + DA
constexpr int XP = 2 * DD + DE + 8;    // per-row state in shared memory: [h (512) | ctx (512) | tok (16)] (+8: pitch = 24 mod 32 banks, see mma_a_frag)
constexpr int XP4 = XP / 4;
constexpr int KM4 = 2 * DD / 4;        // 256 float4 columns of [h | ctx]
constexpr int TILE_MAX = 320;          // frames per attention item (host: NS >= ceil(T / TILE_MAX))
constexpr int NCB = 4, NFS = NW / NCB; // pass 2: 4 column blocks of 128 x 3 frame subsets
static_assert(NW == NCB * NFS, "warp layout of the context pass");

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// exp(2x) with the exponent clamped to +-21 (|x| <= 10.5 is exact; tanh is saturated to 1e-9 there): 1 + exp(2q) exp(2e) <= 1.8e18, so
// the product of TWO such terms stays finite and one reciprocal serves two elements: 1/a = b * rcp(a b), 1/b = a * rcp(a b).
__device__ __forceinline__ float exp2x(float x) { return expf(fminf(fmaxf(2.f * x, -21.f), 21.f)); }
// sum over two elements of w_k / (1 + q_k e_k) with one SFU reciprocal
__device__ __forceinline__ float pair_term(float q0, float e0, float w0, float q1, float e1, float w1) {
    const float a0 = fmaf(q0, e0, 1.f), a1 = fmaf(q1, e1, 1.f);
    return rcp_fast(a0 * a1) * fmaf(w0, a1, w1 * a0);
}
// Packed fp32 pairs (Blackwell FFMA2 / FMUL2: two fp32 operations per instruction on a register pair).  The attention passes are
// bound by instruction issue + latency at 12 warps per SM, so halving the FP instruction count of their inner loops is a direct gain.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
// acc += sum over the four elements of w_k / (1 + q_k e_k), two SFU reciprocals: elements (x, z) and (y, w) share one each
// (1/a = b rcp(ab)), everything else on register pairs (xy, zw)
__device__ __forceinline__ float2 quad_term(float4 q, float4 e, float4 w, float2 acc) {
    const float2 one = make_float2(1.f, 1.f);
    const float2 axy = ffma2(make_float2(q.x, q.y), make_float2(e.x, e.y), one);
    const float2 azw = ffma2(make_float2(q.z, q.w), make_float2(e.z, e.w), one);
    const float2 pr = fmul2(axy, azw);                                           // (a0 a2, a1 a3)
    const float2 r = make_float2(rcp_fast(pr.x), rcp_fast(pr.y));
    const float2 t = ffma2(make_float2(w.x, w.y), azw, fmul2(make_float2(w.z, w.w), axy));   // (w0 a2 + w2 a0, w1 a3 + w3 a1)
    return ffma2(r, t, acc);
}
// the reverse pass: (dxy, dzw) += ds * r (1 - r) for the four elements, r = 1 / (1 + q e)
__device__ __forceinline__ void quad_grad(float4 q, float4 e, float ds, float2& dxy, float2& dzw) {
    const float2 one = make_float2(1.f, 1.f), mone = make_float2(-1.f, -1.f), ds2 = make_float2(ds, ds);
    const float2 axy = ffma2(make_float2(q.x, q.y), make_float2(e.x, e.y), one);
    const float2 azw = ffma2(make_float2(q.z, q.w), make_float2(e.z, e.w), one);
    const float2 pr = fmul2(axy, azw);
    const float2 rp = make_float2(rcp_fast(pr.x), rcp_fast(pr.y));
    const float2 rxy = fmul2(azw, rp), rzw = fmul2(axy, rp);                     // 1 / a_xy, 1 / a_zw
    dxy = ffma2(fmul2(ds2, rxy), ffma2(rxy, mone, one), dxy);
    dzw = ffma2(fmul2(ds2, rzw), ffma2(rzw, mone, one), dzw);
}
// per-lane asynchronous copies global -> shared (a lane only ever reads back what it copied itself: no barrier needed, and the
// depth of the prefetch costs neither registers nor scoreboard slots)
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ int mslot(const DecMArgs& a, int s) { return a.save ? s : 0; }
__device__ __forceinline__ int mhslot(const DecMArgs& a, int s) { return a.save ? s : (s & 1); }

// Grid barrier for a co-resident grid (monotonic arrival counter); a lost CTA becomes a trap, not a hang (see dec_persist.cu).
__device__ __forceinline__ void grid_sync(unsigned int* sync, unsigned int& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(sync, 1u);
        unsigned int spins = 0;
        while (true) {
            unsigned int v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(sync) : "memory");
            if (v >= target) break;
            if (++spins > (1u << 24)) { atomicExch(sync + 1, 1u); __threadfence_system(); __trap(); }
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define SUB_BEGIN() unsigned long long sub_t = (a.prof != nullptr && threadIdx.x == 0 && blockIdx.x == 0) ? gtimer() : 0ull
#define SUB_MARK(i)                                                                  \
    do {                                                                             \
        if (a.prof != nullptr && threadIdx.x == 0 && blockIdx.x == 0) {              \
            const unsigned long long now_ = gtimer();                                \
            a.prof[i] += now_ - sub_t;                                               \
            sub_t = now_;                                                            \
        }                                                                            \
    } while (0)
#define PROF_MARK(i)                                                                 \
    do {                                                                             \
        if (a.prof != nullptr && threadIdx.x == 0 && blockIdx.x == 0) {              \
            const unsigned long long now_ = gtimer();                                \
            a.prof[i] += now_ - prof_t;                                              \
            prof_t = now_;                                                           \
        }                                                                            \
    } while (0)

// Sum N values held by every lane across the warp so that each lane ends up with N/32 complete sums:
// after the call v[i] (i < N/32) is the warp total of original element rs_base<N>(lane) + i.
template <int N>
__device__ __forceinline__ void reduce_scatter(float (&v)[N], int lane) {
    static_assert(N % 32 == 0, "N must be a multiple of the warp size");
#pragma unroll
    for (int off = 16, n = N / 2; off >= 1; off >>= 1, n >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float send = up ? v[i] : v[i + n];
            const float recv = __shfl_xor_sync(0xffffffffu, send, off);
            v[i] = (up ? v[i + n] : v[i]) + recv;
        }
    }
}
template <int N>
__device__ __forceinline__ int rs_base(int lane) {
    int base = 0;
#pragma unroll
    for (int off = 16, n = N / 2; off >= 1; off >>= 1, n >>= 1) base += (lane & off) ? n : 0;
    return base;
}

// ---- tensor-core path of the weight-stationary products (precision modes bf16x3 / bf16) ---------------------------------------
// A chunk of 16 staged rows is ONE m16 tile: C[16 x 8n] += A[16 x K] W^T, mma.sync.m16n8k16 bf16 with fp32 accumulation, the K range
// split over the 12 warps.  fp32 operands are split into bf16 hi + lo on the fly (activations) / at kernel start (weights) and
// hi*hi + lo*hi + hi*lo is accumulated (the bf16x3 scheme of the tcgen05 GEMMs: ~5e-6 relative).  Replaces the FFMA products and
// their warp-shuffle reductions (10x fewer instructions per chunk); the exact-fp32 mode (eval / greedy decode) keeps the FFMA path.
__device__ __forceinline__ void split2(float x0, float x1, unsigned int& hi, unsigned int& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const unsigned int*>(&h);
    lo = *reinterpret_cast<const unsigned int*>(&l);
}
__device__ __forceinline__ void mma16816(float (&c)[4], const unsigned int (&a)[4], unsigned int b0, unsigned int b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// A fragment of k-step `col0` (16 columns) from fp32 rows of pitch `pitch` floats: rows g / g+8, columns 2tg,2tg+1 / +8.  With
// pitch = 8 or 24 (mod 32) the 64-bit loads of a half warp hit 32 distinct banks.
__device__ __forceinline__ void mma_a_frag(const float* x, int pitch, int col0, int g, int tg, unsigned int (&hi)[4], unsigned int (&lo)[4]) {
    const float2 v0 = *reinterpret_cast<const float2*>(x + g * pitch + col0 + 2 * tg);
    const float2 v1 = *reinterpret_cast<const float2*>(x + (g + 8) * pitch + col0 + 2 * tg);
    const float2 v2 = *reinterpret_cast<const float2*>(x + g * pitch + col0 + 8 + 2 * tg);
    const float2 v3 = *reinterpret_cast<const float2*>(x + (g + 8) * pitch + col0 + 8 + 2 * tg);
    split2(v0.x, v0.y, hi[0], lo[0]);
    split2(v1.x, v1.y, hi[1], lo[1]);
    split2(v2.x, v2.y, hi[2], lo[2]);
    split2(v3.x, v3.y, hi[3], lo[3]);
}
// acc[j] += A (hi, lo) x rows j*8 .. j*8+7 of the packed weights (bf16x2 words, row pitch wp words), k-step ks
template <int NTILES>
__device__ __forceinline__ void mma_kstep(float (&acc)[NTILES][4], const unsigned int (&ahi)[4], const unsigned int (&alo)[4],
                                          const unsigned int* whi, const unsigned int* wlo, int wp, int ks, int g, int tg) {
#pragma unroll
    for (int j = 0; j < NTILES; ++j) {
        const int o = (j * 8 + g) * wp + ks * 8 + tg;
        const unsigned int bh0 = whi[o], bh1 = whi[o + 4], bl0 = wlo[o], bl1 = wlo[o + 4];
        mma16816(acc[j], ahi, bh0, bh1);
        mma16816(acc[j], alo, bh0, bh1);
        mma16816(acc[j], ahi, bl0, bl1);
    }
}
constexpr int WGP = (2 * DD + DE) / 2 + 4;   // packed row pitch of the gate weights (524 words = 12 mod 32: conflict-free B fragments)
constexpr int WCP = DD + 4;                  // ... of the phase-C rows (K = 1024 -> 512 words + 4)
constexpr int WTP = 3 * DD / 2 + 4;          // ... of the reverse kernel's W^T rows (K = 1536 -> 768 words + 4)

// per-query step counts, copied to shared memory at kernel start (dynamic indexing of the kernel-parameter array would force a
// local-memory copy of the whole argument block)
__shared__ int g_Sq[8];
__shared__ unsigned int g_tf[64];      // teacher-forcing coins of the launch (same reason)

constexpr int RED_FLOATS = NW * BT * GR;               // FFMA path: (4 clip groups x 8 warps) x (24 rows x 4 clips) = 3072; MMA path: 12 warps x 16 x 24
constexpr int WG_WORDS = 2 * GR * WGP;                 // >= GR * 2 * DD + GR * DE (fp32 layout of the FFMA path)
constexpr int WC_WORDS = 2 * 8 * WCP;                  // >= CR * 2 * DD
static_assert(WG_WORDS >= GR * 2 * DD + GR * DE && WC_WORDS >= CR * 2 * DD, "the packed layouts must cover the fp32 ones");
struct FwdSmem {
    float* Wg;      // FFMA: [GR][1024] rows g*8+u: [W_hh row | W_ih row, context columns], then Wtok [GR][16]
                    // MMA:  hi [GR][WGP] | lo [GR][WGP] bf16x2 words over k = [h | ctx | tok]
    float* Wc;      // FFMA: [CR][1024] W_out rows / W_h rows (zero beyond 512) / zero rows;  MMA: hi [8][WCP] | lo [8][WCP]
    float* Wtok;    // [GR][16] (FFMA path)
    float* bias;    // [4][8]       b_r (ih+hh), b_z (ih+hh), b_in, b_hn
    float* bc;      // [8]
    float* xs;      // [BT][XP]     staged rows of phases B / C; attention scratch of phase A
    float* red;     // RED_FLOATS
    float* v2;      // [DA]         -2 v
};
constexpr int FWD_SMEM_FLOATS = WG_WORDS + WC_WORDS + 32 + 8 + BT * XP + RED_FLOATS + DA;
__device__ __forceinline__ FwdSmem carve(float* sm) {
    FwdSmem s;
    s.Wg = sm; s.Wtok = sm + GR * 2 * DD; sm += WG_WORDS;
    s.Wc = sm; sm += WC_WORDS;
    s.bias = sm; sm += 32;
    s.bc = sm; sm += 8;
    s.xs = sm; sm += BT * XP;
    s.red = sm; sm += RED_FLOATS;
    s.v2 = sm; sm += DA;
    return s;
}
// attention scratch inside xs: Eq [NQ][DA] | sc [NQ][TILE_MAX] | ring [NW][RING] (per-lane cp.async rings of the two streaming passes),
// aliased after the loops by cred [NFS][NQ][DD]
constexpr int D1 = 4, D2 = 8;                 // frames in flight per warp: pass 1 (1 KB per frame), pass 2 (512 B per frame)
constexpr int RING = D2 * 32 * 4;             // floats per warp
static_assert(D1 * 32 * 8 == RING, "both passes use the same ring bytes");
static_assert(NQMAX * (DA + TILE_MAX) + NW * RING <= BT * XP && NFS * NQMAX * DD <= NW * RING, "attention scratch must fit the staged-row region");

// ------------------------------------------------------------------------------------------------ phase A (forward)
template <int NQ>
__device__ void attn_item(const DecMArgs& a, const FwdSmem& S, float sumv, int s, int b, int js) {
    __shared__ float wred[NW][NQMAX];
    __shared__ float Mq[NQMAX], Lq[NQMAX];
    __shared__ int is_last;
    constexpr int TS = TILE_MAX * (NQMAX / NQ);    // frames per item the score table holds: 320 (NQ >= 3), 800 (NQ = 2), 1600 (NQ = 1)
    float* Eq = S.xs;                              // [NQ][DA]
    float* sc = Eq + NQ * DA;                      // [NQ][TS]
    float* ringb = sc + NQ * TS;                   // [NW][RING]
    float* cred = ringb;                           // [NFS][NQ][DD] (after the streaming loops)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = a.T, B = a.B;
    const int t0 = js * a.tile, t1 = min(T, t0 + a.tile), nt = t1 - t0;
    float* myring = ringb + warp * RING;
    bool actq[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) actq[q] = s < g_Sq[q];
    SUB_BEGIN();
    __syncthreads();
    // ---- exp(2 q_s) of every active query of this clip
    for (int i = tid; i < NQ * DA; i += NT) {
        const int q = i / DA, k = i - q * DA;
        float e = 0.f;
        if (s < g_Sq[q]) {
            const size_t g = (size_t)a.r0 + q * B + b;
            e = exp2x(__ldcg(a.qs + ((size_t)mhslot(a, s) * a.Rtot + g) * DA + k));
            if (js == 0 && a.save) a.eqs[((size_t)s * a.Rtot + g) * DA + k] = e;
        }
        Eq[i] = e;
    }
    __syncthreads();
    SUB_MARK(8);
    // ---- pass 1: scores.  Warp w takes the frames t0 + w, t0 + w + 12, ...; a lane holds 8 of the 256 exp(2 Ep) values of a frame
    // (D1 frames in flight through its cp.async ring); G frames x NQ queries = up to 32 per-lane partial sums are reduced at once.
    {
        const float4 va = *reinterpret_cast<const float4*>(S.v2 + lane * 4);
        const float4 vb = *reinterpret_cast<const float4*>(S.v2 + 128 + lane * 4);
        float mq[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) mq[q] = -INFINITY;
        constexpr int G = (32 / NQ) < 8 ? (32 / NQ) : 8;
        const float4* ee = reinterpret_cast<const float4*>(a.Ee + (size_t)b * T * DA) + lane;
        float4* slot = reinterpret_cast<float4*>(myring) + lane;            // slot i: [i*64 + lane] and [i*64 + 32 + lane]
        const int base = rs_base<32>(lane);
        const int nf = (t1 - t0 - warp + NW - 1) / NW;                      // frames of this warp (<= 0: none)
#pragma unroll
        for (int i = 0; i < D1; ++i) {
            if (i < nf) {
                const size_t t = t0 + warp + i * NW;
                cp_async16(slot + i * 64, ee + t * (DA / 4));
                cp_async16(slot + i * 64 + 32, ee + t * (DA / 4) + 32);
            }
            cp_async_commit();
        }
        for (int f0 = 0; f0 < nf; f0 += G) {
            float val[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) val[i] = 0.f;
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const int f = f0 + i;
                if (f < nf) {                                                // warp-uniform
                    cp_async_wait<D1 - 1>();
                    const int sl = f % D1;
                    const float4 c0 = slot[sl * 64], c1 = slot[sl * 64 + 32];
                    if (f + D1 < nf) {
                        const size_t t = t0 + warp + (f + D1) * NW;
                        cp_async16(slot + sl * 64, ee + t * (DA / 4));
                        cp_async16(slot + sl * 64 + 32, ee + t * (DA / 4) + 32);
                    }
                    cp_async_commit();
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        if (!actq[q]) continue;                              // block-uniform
                        const float4 qa = *reinterpret_cast<const float4*>(Eq + q * DA + lane * 4);
                        const float4 qb = *reinterpret_cast<const float4*>(Eq + q * DA + 128 + lane * 4);
                        const float2 acc2 = quad_term(qb, c1, vb, quad_term(qa, c0, va, make_float2(0.f, 0.f)));
                        val[i * NQ + q] = acc2.x + acc2.y;
                    }
                }
            }
            reduce_scatter<32>(val, lane);
            // the lane now holds the warp total of element `base` = (frame base / NQ of the group, query base % NQ)
            if (base < G * NQ) {
                const int i = base / NQ, q = base - i * NQ;
                const int f = f0 + i;
                if (f < nf && s < g_Sq[q]) {
                    const float e = sumv + val[0];                           // sum_k v_k tanh(q_k + Ep_k)
                    sc[q * TS + warp + f * NW] = e;
#pragma unroll
                    for (int qq = 0; qq < NQ; ++qq)
                        if (qq == q) mq[qq] = fmaxf(mq[qq], e);
                }
            }
        }
        cp_async_wait<0>();
#pragma unroll
        for (int q = 0; q < NQ; ++q) mq[q] = warp_max(mq[q]);
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) wred[warp][q] = mq[q];
        }
    }
    __syncthreads();
    SUB_MARK(9);
    if (tid < NQ) {
        float m = -INFINITY;
#pragma unroll
        for (int w = 0; w < NW; ++w) m = fmaxf(m, wred[w][tid]);
        Mq[tid] = m;
    }
    __syncthreads();
    // ---- p = exp(score - max) in place (raw scores go to the saved attention rows), partial softmax sums
    {
        float ls[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            ls[q] = 0.f;
            if (!actq[q]) continue;
            const float m = Mq[q];
            float* araw = a.save ? a.attn + ((size_t)s * a.Rtot + a.r0 + q * B + b) * T + t0 : nullptr;
            for (int j = tid; j < nt; j += NT) {
                const float e = sc[q * TS + j];
                if (araw != nullptr) araw[j] = e;
                const float p = expf(e - m);
                sc[q * TS + j] = p;
                ls[q] += p;
            }
            ls[q] = warp_sum(ls[q]);
        }
        __syncthreads();                                             // everyone has read Mq / wred users are done
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) wred[warp][q] = ls[q];
        }
    }
    __syncthreads();
    if (tid < NQ) {
        float l = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) l += wred[w][tid];
        Lq[tid] = l;
    }
    SUB_MARK(10);
    // ---- pass 2: contexts.  Warp (cb, fs): 128-column block cb of the frames t0 + fs, t0 + fs + 3, ...; D2 frames in flight.
    {
        const int cb = warp & (NCB - 1), fs = warp / NCB;
        float4 acc[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* en = reinterpret_cast<const float4*>(a.enc + (size_t)b * T * DD) + cb * 32 + lane;
        float4* slot = reinterpret_cast<float4*>(myring) + lane;            // slot i: [i*32 + lane]
        const int nf = (t1 - t0 - fs + NFS - 1) / NFS;
#pragma unroll
        for (int i = 0; i < D2; ++i) {
            if (i < nf) cp_async16(slot + i * 32, en + (size_t)(t0 + fs + i * NFS) * (DD / 4));
            cp_async_commit();
        }
        for (int f0 = 0; f0 < nf; f0 += D2) {
#pragma unroll
            for (int i = 0; i < D2; ++i) {
                const int f = f0 + i;
                if (f >= nf) break;                                          // warp-uniform
                cp_async_wait<D2 - 1>();
                const float4 e = slot[i * 32];
                if (f + D2 < nf) cp_async16(slot + i * 32, en + (size_t)(t0 + fs + (f + D2) * NFS) * (DD / 4));
                cp_async_commit();
                const int j = fs + f * NFS;
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    if (!actq[q]) continue;
                    const float p = sc[q * TS + j];
                    acc[q].x = fmaf(p, e.x, acc[q].x);
                    acc[q].y = fmaf(p, e.y, acc[q].y);
                    acc[q].z = fmaf(p, e.z, acc[q].z);
                    acc[q].w = fmaf(p, e.w, acc[q].w);
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();                                                     // every warp is done with its ring: cred aliases it
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            *reinterpret_cast<float4*>(cred + ((size_t)(fs * NQ + q)) * DD + cb * 128 + lane * 4) = acc[q];
    }
    __syncthreads();
    SUB_MARK(11);
    // ---- item partials -> global
    for (int i = tid; i < NQ * (DD / 2); i += NT) {
        const int q = i / (DD / 2), c2 = i - q * (DD / 2);
        if (s >= g_Sq[q]) continue;
        float2 v = make_float2(0.f, 0.f);
#pragma unroll
        for (int f = 0; f < NFS; ++f) {
            const float2 w = *reinterpret_cast<const float2*>(cred + ((size_t)(f * NQ + q)) * DD + 2 * c2);
            v.x += w.x; v.y += w.y;
        }
        const int r = q * B + b;
        reinterpret_cast<float2*>(a.pc + ((size_t)r * a.NS + js) * DD)[c2] = v;
        if (c2 == 0) { a.pm[r * a.NS + js] = Mq[q]; a.pl[r * a.NS + js] = Lq[q]; }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int tk = atomicAdd(a.tickets + b, 1);
        is_last = (tk == a.NS - 1);
        if (is_last) a.tickets[b] = 0;
    }
    __syncthreads();
    SUB_MARK(12);
    if (!is_last) return;
    __threadfence();
    // ---- last CTA of the clip: combine the NS partials of every active query
    for (int i = tid; i < NQ * (DD / 2); i += NT) {
        const int q = i / (DD / 2), c2 = i - q * (DD / 2);
        if (s >= g_Sq[q]) continue;
        const int r = q * B + b;
        float M = -INFINITY;
        for (int j = 0; j < a.NS; ++j) M = fmaxf(M, __ldcg(a.pm + r * a.NS + j));
        float L = 0.f, c0 = 0.f, c1 = 0.f;
        for (int j = 0; j < a.NS; ++j) {
            const float mj = __ldcg(a.pm + r * a.NS + j);
            const float wgt = (mj == -INFINITY) ? 0.f : expf(mj - M);
            L = fmaf(__ldcg(a.pl + r * a.NS + j), wgt, L);
            const float2 pj = __ldcg(reinterpret_cast<const float2*>(a.pc + ((size_t)r * a.NS + j) * DD) + c2);
            c0 = fmaf(pj.x, wgt, c0);
            c1 = fmaf(pj.y, wgt, c1);
        }
        const float invL = 1.f / L;
        const size_t g = (size_t)a.r0 + r;
        reinterpret_cast<float2*>(a.ctxs + ((size_t)mslot(a, s) * a.Rtot + g) * DD)[c2] = make_float2(c0 * invL, c1 * invL);
        if (c2 == 0 && a.save) { a.ml[((size_t)s * a.Rtot + g) * 2] = M; a.ml[((size_t)s * a.Rtot + g) * 2 + 1] = invL; }
    }
}

// ------------------------------------------------------------------------------------------------ phase D
// Finalise step s of row (q, b) by ONE warp: log-softmax row, greedy token, teacher forcing, EOS bookkeeping, next input embedding.
__device__ void finalize_row(const DecMArgs& a, int s, int q, int b, int eos_id) {
    const int lane = threadIdx.x & 31;
    const int V = a.V, r = q * a.B + b;
    const size_t g = (size_t)a.r0 + r;
    constexpr int PER = 8;                                           // V <= 256
    float x[PER];
    float m = -INFINITY;
    int idx = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int vi = lane + 32 * i;
        x[i] = vi < V ? __ldcg(a.logits + (size_t)r * a.VP + vi) : -INFINITY;
        if (vi < V && (idx == 0x7fffffff || x[i] > m)) { m = x[i]; idx = vi; }      // first index of the maximum, like torch.argmax
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (om > m || (om == m && oi < idx)) { m = om; idx = oi; }
    }
    float e = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) e += (lane + 32 * i < V) ? expf(x[i] - m) : 0.f;
    e = warp_sum(e);
    const float lse = m + logf(e);
    const size_t ro = ((size_t)b * a.bars + a.k0 + q) * a.max_steps + s;
#pragma unroll
    for (int i = 0; i < PER; ++i)
        if (lane + 32 * i < V) a.logp[ro * V + lane + 32 * i] = x[i] - lse;
    int tok = 0;
    if (lane == 0) {
        const long long gtv = a.gt != nullptr ? a.gt[ro] : -1;
        const int tb = q * a.Spitch + s;
        const bool tf = (!a.inference) && a.has_tf && a.gt != nullptr && ((g_tf[tb >> 5] >> (tb & 31)) & 1u);
        tok = tf ? (int)gtv : idx;
        const bool hit = a.gt != nullptr ? (gtv == eos_id) : (idx == eos_id);
        if (hit) {
            a.lengths[g] = s + 1;
            if (a.eos[r] == 0) { a.eos[r] = 1; atomicAdd(a.counters, 1); }
        }
        if (b == 0) atomicAdd(a.counters + 1, 1);                     // executed (bar, step) pairs of this launch
        if (a.save && s + 1 <= a.S) a.toks[(size_t)(s + 1) * a.Rtot + g] = tok;
    }
    tok = __shfl_sync(0xffffffffu, tok, 0);
    if (lane < DE && s + 1 < g_Sq[q]) {
        const float mk = a.mask != nullptr ? a.mask[((size_t)(s + 1) * a.Rtot + g) * DE + lane] : 1.f;
        const float xv = a.emb[(size_t)tok * DE + lane] * mk;
        a.xbuf[(size_t)r * DX + lane] = xv;
        if (a.save) a.xtok[((size_t)(s + 1) * a.Rtot + g) * DE + lane] = xv;
    }
}

// ------------------------------------------------------------------------------------------------ staging of a row chunk
// parts bit 0 = h (hs[hs_slot]), bit 1 = ctx (ctxs[ctx_slot]), bit 2 = token embedding; rows [rb0, rb0+nb) of this launch
__device__ void stage_xs(const DecMArgs& a, const FwdSmem& S, int rb0, int nb, int parts, int hs_slot, int ctx_slot) {
    float4* xs4 = reinterpret_cast<float4*>(S.xs);
    const int tid = threadIdx.x;
    const int n4 = nb * (DD / 4);
    // every part with asynchronous 16-byte copies (L2 -> shared memory, no registers): ONE L2 round trip for the whole chunk
    if ((parts & 4) && tid < nb * (DE / 4)) cp_async16(xs4 + (tid >> 2) * XP4 + 256 + (tid & 3), a.xbuf + (size_t)(rb0 + (tid >> 2)) * DX + (tid & 3) * 4);
#pragma unroll
    for (int part = 0; part < 2; ++part) {
        if (!(parts & (1 << part))) continue;
        const float* src = part == 0 ? a.hs + ((size_t)hs_slot * a.Rtot + a.r0 + rb0) * DD : a.ctxs + ((size_t)ctx_slot * a.Rtot + a.r0 + rb0) * DD;
        for (int i = tid; i < n4; i += NT) cp_async16(xs4 + (i >> 7) * XP4 + part * 128 + (i & 127), src + (size_t)i * 4);
    }
    cp_async_commit();
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------ phase B
__device__ void gru_phase(const DecMArgs& a, const FwdSmem& S, int s, int rb0, int nb) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int col = tid & (KM4 - 1), warp = col >> 5;
    const float4* xs4 = reinterpret_cast<const float4*>(S.xs);
    const float4* wg4 = reinterpret_cast<const float4*>(S.Wg);
    const int base = rs_base<UPC * 4>(lane);
    if (tid < KM4) {
#pragma unroll 1
        for (int cg = 0; cg < BT / 4; ++cg) {
            if (cg * 4 >= nb) break;
            float4 xv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) xv[c] = xs4[(cg * 4 + c) * XP4 + col];
            float* dst = S.red + (cg * 8 + warp) * (GR * 4) + base;
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                float acc[UPC * 4];
#pragma unroll
                for (int u = 0; u < UPC; ++u) {
                    const float4 w = wg4[(g * UPC + u) * KM4 + col];
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[u * 4 + c] = dot4(w, xv[c]);
                }
                reduce_scatter<UPC * 4>(acc, lane);
                dst[g * UPC * 4] = acc[0];
            }
        }
    }
    __syncthreads();
    if (tid < UPC * BT) {
        const int br = tid >> 3, u = tid & 7;
        const int r = rb0 + br;
        if (br < nb && s < g_Sq[r / a.B]) {
            const int cg = br >> 2, c = br & 3;
            float g3[3], nh = 0.f;
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                const int idx = (g * UPC + u) * 4 + c;
                float lo = 0.f, hi = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) lo += S.red[(cg * 8 + w) * (GR * 4) + idx];       // hidden-state columns
#pragma unroll
                for (int w = 4; w < 8; ++w) hi += S.red[(cg * 8 + w) * (GR * 4) + idx];       // context columns
                const float* wt = S.Wtok + (g * UPC + u) * DE;
                const float* xt = S.xs + br * XP + 2 * DD;
#pragma unroll
                for (int k = 0; k < DE; ++k) hi = fmaf(wt[k], xt[k], hi);
                if (g < 2) g3[g] = lo + hi;
                else { g3[2] = hi; nh = lo; }
            }
            const int j = blockIdx.x * UPC + u;
            const size_t gr = (size_t)a.r0 + r;
            const float rr = sigmoidf_(g3[0] + S.bias[u]);
            const float z = sigmoidf_(g3[1] + S.bias[8 + u]);
            const float hnl = nh + S.bias[24 + u];
            const float n = tanhf(g3[2] + S.bias[16 + u] + rr * hnl);
            const float hp = S.xs[br * XP + j];
            const float hn = (1.f - z) * n + z * hp;
            a.hs[((size_t)mhslot(a, s + 1) * a.Rtot + gr) * DD + j] = hn;
            if (a.save) {
                float* gs = a.gates + ((size_t)s * a.Rtot + gr) * 4 * DD + j;
                gs[0] = rr; gs[DD] = z; gs[2 * DD] = n; gs[3 * DD] = hnl;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ phase C
// logits and next query for the CTA's 7 rows and the rows staged in xs ([h' | ctx]); s < 0: prologue (query only, every row)
__device__ void out_phase(const DecMArgs& a, const FwdSmem& S, int s, int qslot, int rb0, int nb) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int col = tid & (KM4 - 1), warp = col >> 5;
    const float4* xs4 = reinterpret_cast<const float4*>(S.xs);
    const float4* wc4 = reinterpret_cast<const float4*>(S.Wc);
    if (tid < KM4) {
        float4 wr[CR];
#pragma unroll
        for (int r = 0; r < CR; ++r) wr[r] = wc4[r * KM4 + col];
        const int base = rs_base<32>(lane);
#pragma unroll 1
        for (int cg = 0; cg < BT / 4; ++cg) {
            if (cg * 4 >= nb) break;
            float acc[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 x = xs4[(cg * 4 + c) * XP4 + col];
#pragma unroll
                for (int r = 0; r < CR; ++r) acc[r * 4 + c] = dot4(wr[r], x);
            }
            reduce_scatter<32>(acc, lane);
            S.red[(cg * 8 + warp) * 32 + base] = acc[0];
        }
    }
    __syncthreads();
    if (tid < CR * BT) {
        const int rr = tid >> 4, br = tid & 15;
        const int rg = blockIdx.x * CR + rr;
        const int r = rb0 + br;
        if (br < nb && rg < a.V + DA && (s < 0 || s < g_Sq[r / a.B])) {
            const int cg = br >> 2, c = br & 3;
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += S.red[(cg * 8 + w) * 32 + rr * 4 + c];
            if (rg < a.V) {
                if (s >= 0) a.logits[(size_t)r * a.VP + rg] = v + S.bc[rr];
            } else {
                a.qs[((size_t)qslot * a.Rtot + a.r0 + r) * DA + (rg - a.V)] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ phases B / C on the tensor cores
// B: the staged chunk [h | ctx | tok] (16 rows x 1040) times this CTA's 24 gate rows.  Warps 0-5 take the hidden-state k-steps
// (W_hh h), warps 6-11 the input k-steps (W_ih [ctx, tok]): the n gate needs the two sums apart.
__device__ void gru_phase_tc(const DecMArgs& a, const FwdSmem& S, int s, int rb0, int nb) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
    const unsigned int* whi = reinterpret_cast<const unsigned int*>(S.Wg);
    const unsigned int* wlo = whi + GR * WGP;
    float acc[3][4];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
    const int half = warp / 6, wi = warp - half * 6;
    const int ks1 = half ? (2 * DD + DE) / 16 : DD / 16;
    for (int ks = half * (DD / 16) + wi; ks < ks1; ks += 6) {
        unsigned int ahi[4], alo[4];
        mma_a_frag(S.xs, XP, ks * 16, g, tg, ahi, alo);
        mma_kstep<3>(acc, ahi, alo, whi, wlo, WGP, ks, g, tg);
    }
    float* dst = S.red + warp * (BT * GR);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        *reinterpret_cast<float2*>(dst + g * GR + j * 8 + 2 * tg) = make_float2(acc[j][0], acc[j][1]);
        *reinterpret_cast<float2*>(dst + (g + 8) * GR + j * 8 + 2 * tg) = make_float2(acc[j][2], acc[j][3]);
    }
    __syncthreads();
    if (tid < UPC * BT) {
        const int br = tid >> 3, u = tid & 7;
        const int r = rb0 + br;
        if (br < nb && s < g_Sq[r / a.B]) {
            float g3[3], nh = 0.f;
#pragma unroll
            for (int gi = 0; gi < 3; ++gi) {
                const int idx = br * GR + gi * 8 + u;
                float lo = 0.f, hi = 0.f;
#pragma unroll
                for (int w = 0; w < 6; ++w) lo += S.red[w * (BT * GR) + idx];             // W_hh h
#pragma unroll
                for (int w = 6; w < 12; ++w) hi += S.red[w * (BT * GR) + idx];            // W_ih [ctx, tok]
                if (gi < 2) g3[gi] = lo + hi;
                else { g3[2] = hi; nh = lo; }
            }
            const int j = blockIdx.x * UPC + u;
            const size_t gr = (size_t)a.r0 + r;
            const float rr = sigmoidf_(g3[0] + S.bias[u]);
            const float z = sigmoidf_(g3[1] + S.bias[8 + u]);
            const float hnl = nh + S.bias[24 + u];
            const float n = tanhf(g3[2] + S.bias[16 + u] + rr * hnl);
            const float hp = S.xs[br * XP + j];
            const float hn = (1.f - z) * n + z * hp;
            a.hs[((size_t)mhslot(a, s + 1) * a.Rtot + gr) * DD + j] = hn;
            if (a.save) {
                float* gs = a.gates + ((size_t)s * a.Rtot + gr) * 4 * DD + j;
                gs[0] = rr; gs[DD] = z; gs[2 * DD] = n; gs[3 * DD] = hnl;
            }
        }
    }
}
// C: [h' | ctx] (16 rows x 1024) times this CTA's 7 (+1 zero) logit / query rows
__device__ void out_phase_tc(const DecMArgs& a, const FwdSmem& S, int s, int qslot, int rb0, int nb) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
    const unsigned int* whi = reinterpret_cast<const unsigned int*>(S.Wc);
    const unsigned int* wlo = whi + 8 * WCP;
    float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
    for (int ks = warp; ks < 2 * DD / 16; ks += NW) {
        unsigned int ahi[4], alo[4];
        mma_a_frag(S.xs, XP, ks * 16, g, tg, ahi, alo);
        mma_kstep<1>(acc, ahi, alo, whi, wlo, WCP, ks, g, tg);
    }
    float* dst = S.red + warp * (BT * 8);
    *reinterpret_cast<float2*>(dst + g * 8 + 2 * tg) = make_float2(acc[0][0], acc[0][1]);
    *reinterpret_cast<float2*>(dst + (g + 8) * 8 + 2 * tg) = make_float2(acc[0][2], acc[0][3]);
    __syncthreads();
    if (tid < CR * BT) {
        const int rr = tid >> 4, br = tid & 15;
        const int rg = blockIdx.x * CR + rr;
        const int r = rb0 + br;
        if (br < nb && rg < a.V + DA && (s < 0 || s < g_Sq[r / a.B])) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) v += S.red[w * (BT * 8) + br * 8 + rr];
            if (rg < a.V) {
                if (s >= 0) a.logits[(size_t)r * a.VP + rg] = v + S.bc[rr];
            } else {
                a.qs[((size_t)qslot * a.Rtot + a.r0 + r) * DA + (rg - a.V)] = v;
            }
        }
    }
}

__device__ __forceinline__ bool chunk_active(const DecMArgs& a, int s, int rb0, int nb) {
    // rows are q-major: the chunk covers queries rb0 / B .. (rb0 + nb - 1) / B
    for (int q = rb0 / a.B; q <= (rb0 + nb - 1) / a.B; ++q)
        if (s < g_Sq[q]) return true;
    return false;
}

template <int NQ, bool TC>
__global__ void __launch_bounds__(NT, 1) decm_fwd_kernel(DecMArgs a, int eos_id) {
    extern __shared__ __align__(16) float smem_f[];
    __shared__ float s_sumv;
    const FwdSmem S = carve(smem_f);
    const int tid = threadIdx.x, cta = blockIdx.x, warp = tid >> 5;
    unsigned int target = 0;
    if (tid < 8) g_Sq[tid] = tid == 0 ? a.Sq[0] : tid == 1 ? a.Sq[1] : tid == 2 ? a.Sq[2] : tid == 3 ? a.Sq[3] : tid == 4 ? a.Sq[4] : tid == 5 ? a.Sq[5] : tid == 6 ? a.Sq[6] : a.Sq[7];
    if (tid == 32) {
#pragma unroll
        for (int i = 0; i < 64; ++i) g_tf[i] = a.tf_bits[i];        // (compile-time indices: read straight from the parameter bank)
    }

    // ---- one-time: this CTA's weight slices -> shared memory
    if (TC) {
        unsigned int* ghi = reinterpret_cast<unsigned int*>(S.Wg);
        unsigned int* glo = ghi + GR * WGP;
        constexpr int KW = (2 * DD + DE) / 2;                       // 520 bf16x2 words per gate row: k = [h | ctx | tok]
        for (int i = tid; i < GR * KW; i += NT) {
            const int r = i / KW, pw = i - r * KW, k = 2 * pw;
            const int row = (r / UPC) * DD + cta * UPC + (r % UPC);
            const float* src = k < DD ? a.W_hh + (size_t)row * DD + k
                             : k < 2 * DD ? a.W_ih + (size_t)row * DX + DE + (k - DD) : a.W_ih + (size_t)row * DX + (k - 2 * DD);
            split2(__ldg(src), __ldg(src + 1), ghi[r * WGP + pw], glo[r * WGP + pw]);
        }
        unsigned int* chi = reinterpret_cast<unsigned int*>(S.Wc);
        unsigned int* clo = chi + 8 * WCP;
        for (int i = tid; i < 8 * DD; i += NT) {
            const int rr = i / DD, pw = i - rr * DD, k = 2 * pw;
            const int rg = cta * CR + rr;
            float w0 = 0.f, w1 = 0.f;
            if (rr < CR && rg < a.V) { w0 = __ldg(a.W_out + (size_t)rg * 2 * DD + k); w1 = __ldg(a.W_out + (size_t)rg * 2 * DD + k + 1); }
            else if (rr < CR && rg < a.V + DA && k < DD) { w0 = __ldg(a.Wattn + (size_t)(rg - a.V) * 2 * DD + k); w1 = __ldg(a.Wattn + (size_t)(rg - a.V) * 2 * DD + k + 1); }
            split2(w0, w1, chi[rr * WCP + pw], clo[rr * WCP + pw]);
        }
    } else {
    for (int i = tid; i < GR * KM4; i += NT) {
        const int r = i / KM4, k4 = i % KM4;
        const int row = (r / UPC) * DD + cta * UPC + (r % UPC);
        const float4 w = k4 < 128 ? __ldg(reinterpret_cast<const float4*>(a.W_hh + (size_t)row * DD) + k4)
                                  : __ldg(reinterpret_cast<const float4*>(a.W_ih + (size_t)row * DX + DE) + (k4 - 128));
        reinterpret_cast<float4*>(S.Wg)[i] = w;
    }
    for (int i = tid; i < GR * DE; i += NT) {
        const int r = i / DE, k = i % DE;
        const int row = (r / UPC) * DD + cta * UPC + (r % UPC);
        S.Wtok[i] = a.W_ih[(size_t)row * DX + k];
    }
    for (int i = tid; i < CR * KM4; i += NT) {
        const int rr = i / KM4, k4 = i % KM4;
        const int rg = cta * CR + rr;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rg < a.V) w = __ldg(reinterpret_cast<const float4*>(a.W_out + (size_t)rg * 2 * DD) + k4);
        else if (rg < a.V + DA && k4 < 128) w = __ldg(reinterpret_cast<const float4*>(a.Wattn + (size_t)(rg - a.V) * 2 * DD) + k4);
        reinterpret_cast<float4*>(S.Wc)[i] = w;
    }
    }
    for (int i = tid; i < BT * XP; i += NT) S.xs[i] = 0.f;           // rows beyond a partial chunk must hold finite values
    if (tid < UPC) {
        const int j = cta * UPC + tid;
        S.bias[tid] = a.b_ih[j] + a.b_hh[j];
        S.bias[8 + tid] = a.b_ih[DD + j] + a.b_hh[DD + j];
        S.bias[16 + tid] = a.b_ih[2 * DD + j];
        S.bias[24 + tid] = a.b_hh[2 * DD + j];
        const int rg = cta * CR + tid;
        S.bc[tid] = (tid < CR && rg < a.V) ? a.b_out[rg] : 0.f;
    }
    if (tid < DA) S.v2[tid] = -2.f * a.v[tid];
    if (warp == 0) {                                               // sum_k v_k (fixed order: every CTA gets the same value)
        float sv = 0.f;
        for (int k = tid; k < DA; k += 32) sv += a.v[k];
        sv = warp_sum(sv);
        if (tid == 0) s_sumv = sv;
    }
    __syncthreads();
    const float sumv = s_sumv;

    const int B = a.B, R = NQ * B;
    const int nchunks = (R + BT - 1) / BT;
    const int nitems = B * a.NS;
    unsigned long long prof_t = gtimer();
    // ---- prologue: q_0 = W_h h_0 of every row
    for (int ch = 0; ch < nchunks; ++ch) {
        const int rb0 = ch * BT, nb = min(BT, R - rb0);
        __syncthreads();
        stage_xs(a, S, rb0, nb, 1, mhslot(a, 0), 0);
        __syncthreads();
        if (TC) out_phase_tc(a, S, -1, mhslot(a, 0), rb0, nb);
        else out_phase(a, S, -1, mhslot(a, 0), rb0, nb);
    }
    grid_sync(a.sync, target);
    PROF_MARK(6);

    int s = 0;
    for (; s < a.S; ++s) {
        // ---- D(s-1): one warp per row, rows spread over all CTAs
        if (s > 0) {
            for (int r = cta + PG * warp; r < R; r += PG * NW) {
                const int q = r / B;
                if (s - 1 < g_Sq[q]) finalize_row(a, s - 1, q, r - q * B, eos_id);
            }
        }
        // ---- A(s)
        for (int item = cta; item < nitems; item += PG) attn_item<NQ>(a, S, sumv, s, item / a.NS, item % a.NS);
        PROF_MARK(0);
        grid_sync(a.sync, target);
        PROF_MARK(1);
        if (a.inference && __ldcg(a.counters) >= R) break;          // every row has emitted <eos> (models.py:418-419)
        // ---- B(s)
        // (prefetching the next chunk's rows into registers during the MMAs was measured and dropped: 13.8 -> 17.4 us per step at NQ = 5;
        //  the cp.async staging below is one L2 round trip per chunk and costs no registers)
        for (int ch = 0; ch < nchunks; ++ch) {
            const int rb0 = ch * BT, nb = min(BT, R - rb0);
            if (!chunk_active(a, s, rb0, nb)) continue;
            __syncthreads();
            stage_xs(a, S, rb0, nb, 7, mhslot(a, s), mslot(a, s));
            __syncthreads();
            if (TC) gru_phase_tc(a, S, s, rb0, nb);
            else gru_phase(a, S, s, rb0, nb);
        }
        PROF_MARK(2);
        grid_sync(a.sync, target);
        PROF_MARK(3);
        // ---- C(s)
        for (int ch = 0; ch < nchunks; ++ch) {
            const int rb0 = ch * BT, nb = min(BT, R - rb0);
            if (!chunk_active(a, s, rb0, nb)) continue;
            __syncthreads();
            stage_xs(a, S, rb0, nb, 3, mhslot(a, s + 1), mslot(a, s));
            __syncthreads();
            if (TC) out_phase_tc(a, S, s, mhslot(a, s + 1), rb0, nb);
            else out_phase(a, S, s, mhslot(a, s + 1), rb0, nb);
        }
        PROF_MARK(4);
        grid_sync(a.sync, target);
        PROF_MARK(5);
    }
    if (s == a.S) {                                                 // loop ran to completion: finalise the last step of the longest rows
        for (int r = cta + PG * warp; r < R; r += PG * NW) {
            const int q = r / B;
            if (g_Sq[q] == a.S) finalize_row(a, a.S - 1, q, r - q * B, eos_id);
        }
    }
}

__global__ void decm_init_kernel(DecMArgs a, int sos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int R = a.NQ * a.B;
    if (i < R * DE) {
        const int r = i / DE, e = i % DE;
        const float m = a.mask != nullptr ? a.mask[((size_t)a.r0 + r) * DE + e] : 1.f;
        const float x = a.emb[(size_t)sos * DE + e] * m;
        a.xbuf[(size_t)r * DX + e] = x;
        if (a.save) a.xtok[((size_t)a.r0 + r) * DE + e] = x;
    }
    if (i < R && a.save) a.toks[a.r0 + i] = sos;
}

// ================================================================================================= backward
constexpr int NTB = 384;
constexpr int K3 = 3 * DD;
constexpr int RPB = 16;                // W^T rows per CTA (+1 token row on the first 16 CTAs)
static_assert(K3 / 4 == NTB, "one float4 column of the gate gradients per thread");
static_assert(PG == 64 && RPB * (PG / 2) == DD, "row slicing of the reverse kernel");
constexpr int UP = K3 + 8;             // pitch of the staged gate-gradient rows (8 mod 32 banks: conflict-free A fragments)
constexpr int UP4 = UP / 4;
struct BwdSmem {
    float* WT;      // FFMA: [RPB+1][K3] fp32;  MMA: hi [RPB][WTP] | lo [RPB][WTP] bf16x2 words, then the token row [K3] fp32
    float* U;       // [BT][UP]     gate gradients of the staged rows (P2) | scratch of P1 and P3
    float* red;     // FFMA: [4][12][64] + [12][BT];  MMA: [12][16][16] + [12][BT]
    float* Wq;      // [UPC*2][QP]
    float* v4;      // [DA]         4 v
};
constexpr int QP = 132;
constexpr int BWD_RED_FLOATS = 4 * 12 * 64 + 12 * BT;
constexpr int WT_WORDS = 2 * RPB * WTP + K3;           // >= (RPB + 1) * K3
static_assert(WT_WORDS >= (RPB + 1) * K3 && NW * BT * RPB + 12 * BT <= BWD_RED_FLOATS, "packed W^T layout / MMA partials");
constexpr int BWD_SMEM_FLOATS = WT_WORDS + BT * UP + BWD_RED_FLOATS + UPC * 2 * QP + DA;
// P3 scratch inside U: dc [NQ][DD] | Eq [NQ][DA] | dap [NCB][NQ][TILE_MAX] (ds in dap[0]) | ring [NW][RING], aliased afterwards by dqr [NW][DA]
static_assert(NQMAX * (DD + DA + NCB * TILE_MAX) + NW * RING <= BT * UP && DA <= RING, "attention-backward scratch must fit U");

// ---- P1: dh of this CTA's 8 hidden units, GRU gate gradients -> dgi_all / dgh_all / dh*z; ALL rows in one pass (blocks of P1R rows)
constexpr int P1R = 80;
static_assert(P1R * 2 * QP <= BT * UP, "dq rows of a P1 block must fit U");
__device__ void bwd_gates_all(const DecMArgs& a, const BwdSmem& S, int s, int R) {
    const int tid = threadIdx.x, cta = blockIdx.x;
    float* dqs = S.U;                                 // [P1R][2][QP]
    for (int rblk = 0; rblk < R; rblk += P1R) {
        const int nr = min(P1R, R - rblk);
        __syncthreads();
        for (int i = tid; i < nr * (DA / 4); i += NTB) {
            const int br = i >> 6;
            const int r = rblk + br;
            float* dst = dqs + (i >> 5) * QP + (i & 31) * 4;
            if (s + 1 < g_Sq[r / a.B]) cp_async16(dst, a.dq_all + ((size_t)(s + 1) * a.Rtot + a.r0 + rblk) * DA + (size_t)i * 4);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_commit();
        const int nitems = nr * 2 * UPC;
        for (int base = 0; base < nitems; base += NTB) {
            const int it = base + tid;
            const bool valid = it < nitems;
            const int br = it >> 4, u = (it >> 1) & 7, half = it & 1;
            const int r = rblk + br;
            const int Sr = valid ? g_Sq[r / a.B] : 0;
            const bool on = valid && s < Sr;
            const bool last_step = (s == Sr - 1);
            const int j = cta * UPC + u;
            const size_t sb = (size_t)s * a.Rtot + a.r0 + r;
            // the step's saved values first (their L2 latency overlaps the wait for the dq rows and the dot product)
            float dh = 0.f, carry = 0.f, rr = 0.f, z = 0.f, n = 0.f, hnl = 0.f, hp = 0.f;
            if (on && half == 0) {
                dh = __ldg(a.dhc_all + sb * 2 * DD + j);
                if (!last_step) carry = __ldcg(a.dh_carry + (size_t)r * DD + j);
                const float* gs = a.gates + sb * 4 * DD + j;
                rr = gs[0]; z = gs[DD]; n = gs[2 * DD]; hnl = gs[3 * DD];
                hp = a.hs[sb * DD + j];
            }
            if (base == 0) {
                cp_async_wait<0>();
                __syncthreads();
            }
            float dhq = 0.f;
            if (on && !last_step) {
                const float4* w4 = reinterpret_cast<const float4*>(S.Wq + (u * 2 + half) * QP);
                const float4* d4 = reinterpret_cast<const float4*>(dqs + (br * 2 + half) * QP);
#pragma unroll 8
                for (int i = 0; i < 32; ++i) dhq += dot4(w4[i], d4[i]);
            }
            dhq += __shfl_xor_sync(0xffffffffu, dhq, 1);
            if (half == 0 && on) {
                if (!last_step) dh += dhq + carry;
                const float dn_pre = dh * (1.f - z) * (1.f - n * n);
                const float dr_pre = dn_pre * hnl * rr * (1.f - rr);
                const float dz_pre = dh * (hp - n) * z * (1.f - z);
                float* gi = a.dgi_all + sb * K3 + j;
                float* gh = a.dgh_all + sb * K3 + j;
                gi[0] = dr_pre; gi[DD] = dz_pre; gi[2 * DD] = dn_pre;
                gh[0] = dr_pre; gh[DD] = dz_pre; gh[2 * DD] = dn_pre * rr;
                a.d_hc[(size_t)r * 2 * DD + j] = dh * z;
            }
        }
    }
}

// ---- P2: dx = dgi W_ih (CTAs < 32; 16 context columns + 1 token column), dh_prev = dgh W_hh + dh*z (CTAs >= 32)
// the chunk's gate-gradient rows -> registers (thread tid = float4 column tid of every row); issued one chunk ahead of their use
__device__ __forceinline__ void bwd_chunk_load(const DecMArgs& a, int s, int rb0, int nb, float4 (&r)[BT]) {
    const float* src = ((blockIdx.x < PG / 2) ? a.dgi_all : a.dgh_all) + ((size_t)s * a.Rtot + a.r0 + rb0) * K3;
#pragma unroll
    for (int br = 0; br < BT; ++br)
        if (br < nb) r[br] = ldcg4(src + ((size_t)br * (K3 / 4) + threadIdx.x) * 4);
}
template <bool TC>
__device__ void bwd_gemv_phase(const DecMArgs& a, const BwdSmem& S, int s, int rb0, int nb, float4 (&pre)[BT], bool staged, int next_rb0, int next_nb) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cta = blockIdx.x;
    const bool is_dx = cta < PG / 2;
    const bool has_tok = cta < DE;
    const float* src = (is_dx ? a.dgi_all : a.dgh_all) + ((size_t)s * a.Rtot + a.r0 + rb0) * K3;
    __syncthreads();
    {
        float4* u4 = reinterpret_cast<float4*>(S.U);
        if (staged) {
#pragma unroll
            for (int br = 0; br < BT; ++br)
                if (br < nb) u4[br * UP4 + tid] = pre[br];
            if (next_nb > 0) bwd_chunk_load(a, s, next_rb0, next_nb, pre);      // in flight during this chunk's MMAs
        } else {
            for (int br = 0; br < nb; ++br) cp_async16(u4 + br * UP4 + tid, src + ((size_t)br * (K3 / 4) + tid) * 4);
            cp_async_commit();
            cp_async_wait<0>();
        }
    }
    __syncthreads();
    const float4* x4 = reinterpret_cast<const float4*>(S.U);
    float* tokred = S.red + 4 * 12 * 64;
    if (TC) {
        const int g = lane >> 2, tg = lane & 3;
        const unsigned int* whi = reinterpret_cast<const unsigned int*>(S.WT);
        const unsigned int* wlo = whi + RPB * WTP;
        float acc[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
        for (int ks = warp; ks < K3 / 16; ks += NW) {
            unsigned int ahi[4], alo[4];
            mma_a_frag(S.U, UP, ks * 16, g, tg, ahi, alo);
            mma_kstep<2>(acc, ahi, alo, whi, wlo, WTP, ks, g, tg);
        }
        float* dst = S.red + warp * (BT * RPB);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            *reinterpret_cast<float2*>(dst + g * RPB + j * 8 + 2 * tg) = make_float2(acc[j][0], acc[j][1]);
            *reinterpret_cast<float2*>(dst + (g + 8) * RPB + j * 8 + 2 * tg) = make_float2(acc[j][2], acc[j][3]);
        }
        if (has_tok) {                                         // the one token row stays an FFMA row product
            const float4 w = reinterpret_cast<const float4*>(S.WT + 2 * RPB * WTP)[tid];
#pragma unroll 1
            for (int cg = 0; cg < BT / 4; ++cg) {
                if (cg * 4 >= nb) break;
                float tk[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) tk[c] = warp_sum(dot4(w, x4[(cg * 4 + c) * UP4 + tid]));
                if (lane < 4) tokred[warp * BT + cg * 4 + lane] = lane == 0 ? tk[0] : lane == 1 ? tk[1] : lane == 2 ? tk[2] : tk[3];
            }
        }
    } else {
    const float4* w4 = reinterpret_cast<const float4*>(S.WT);
    const int base = rs_base<32>(lane);
#pragma unroll 1
    for (int cg = 0; cg < BT / 4; ++cg) {
        if (cg * 4 >= nb) break;
        float4 xv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) xv[c] = x4[(cg * 4 + c) * UP4 + tid];
        float* dst = S.red + (cg * 12 + warp) * 64 + base;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float acc[32];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float4 w = w4[(h * 8 + r) * (K3 / 4) + tid];
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r * 4 + c] = dot4(w, xv[c]);
            }
            reduce_scatter<32>(acc, lane);
            dst[h * 32] = acc[0];
        }
        float tk[4] = {0.f, 0.f, 0.f, 0.f};
        if (has_tok) {
            const float4 w = w4[RPB * (K3 / 4) + tid];
#pragma unroll
            for (int c = 0; c < 4; ++c) tk[c] = warp_sum(dot4(w, xv[c]));
        }
        if (has_tok && lane < 4) tokred[warp * BT + cg * 4 + lane] = lane == 0 ? tk[0] : lane == 1 ? tk[1] : lane == 2 ? tk[2] : tk[3];
    }
    }
    __syncthreads();
    if (tid < RPB * BT) {
        const int rr = tid >> 4, br = tid & 15;
        const int r = rb0 + br;
        if (br < nb && s < g_Sq[r / a.B]) {
            float v = 0.f;
            if (TC) {
#pragma unroll
                for (int w = 0; w < NW; ++w) v += S.red[w * (BT * RPB) + br * RPB + rr];
            } else {
                const int cg = br >> 2, c = br & 3;
#pragma unroll
                for (int w = 0; w < 12; ++w) v += S.red[(cg * 12 + w) * 64 + rr * 4 + c];
            }
            if (is_dx) {
                a.dx[(size_t)r * DX + DE + cta * RPB + rr] = v;
            } else {
                const int k = (cta - PG / 2) * RPB + rr;
                a.dh_carry[(size_t)r * DD + k] = v + __ldcg(a.d_hc + (size_t)r * 2 * DD + k);
            }
        }
    } else if (has_tok && tid < RPB * BT + BT) {
        const int br = tid - RPB * BT;
        const int r = rb0 + br;
        if (br < nb && s < g_Sq[r / a.B]) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 12; ++w) v += tokred[w * BT + br];
            a.dxtok_all[((size_t)s * a.Rtot + a.r0 + r) * DE + cta] = v;
        }
    }
}

// ---- P3: attention backward for clip b, frames [t0,t1), all active queries
//   da[q][t] = dc_q . enc_t;  ds = attn (da - c0_q)  -> ds_all (and the normalised attention weight in place of the raw score);
//   dq_q[k]  = 4 v_k sum_t ds[q][t] r (1 - r),  r = 1 / (1 + Eq_q[k] Ee_t[k])                               -> dq_part; last arriver sums
template <int NQ>
__device__ void bwd_attn_item(const DecMArgs& a, const BwdSmem& S, int s, int b, int js) {
    __shared__ float c0s[NQMAX], Ms[NQMAX], iLs[NQMAX];
    __shared__ float wred[NW][NQMAX];
    __shared__ int is_last;
    float* dc = S.U;                               // [NQ][DD]
    float* Eq = dc + NQ * DD;                      // [NQ][DA]
    float* dap = Eq + NQ * DA;                     // [NCB][NQ][TILE_MAX]
    float* ringb = dap + NCB * NQ * TILE_MAX;      // [NW][RING]
    float* dqr = ringb;                            // [NW][DA] (after the streaming loops)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* myring = ringb + warp * RING;
    const int T = a.T, B = a.B;
    const int t0 = js * a.tile, t1 = min(T, t0 + a.tile), nt = t1 - t0;
    bool actq[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) actq[q] = s < g_Sq[q];
    SUB_BEGIN();
    __syncthreads();
    for (int i = tid; i < NQ * DD; i += NTB) {
        const int q = i / DD, d = i - q * DD;
        float v = 0.f;
        if (s < g_Sq[q]) {
            const int r = q * B + b;
            const size_t sb = (size_t)s * a.Rtot + a.r0 + r;
            v = __ldg(a.dhc_all + sb * 2 * DD + DD + d) + __ldcg(a.dx + (size_t)r * DX + DE + d);
            if (js == 0) a.dctx_all[sb * DD + d] = v;
        }
        dc[i] = v;
    }
    for (int i = tid; i < NQ * DA; i += NTB) {
        const int q = i / DA, k = i - q * DA;
        Eq[i] = (s < g_Sq[q]) ? __ldg(a.eqs + ((size_t)s * a.Rtot + a.r0 + q * B + b) * DA + k) : 0.f;
    }
    if (tid < NQ && s < g_Sq[tid]) {
        const size_t sb = (size_t)s * a.Rtot + a.r0 + tid * B + b;
        Ms[tid] = a.ml[sb * 2]; iLs[tid] = a.ml[sb * 2 + 1];
    }
    __syncthreads();
    // c0_q = dc_q . ctx_q
    {
        float c[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            c[q] = 0.f;
            if (!actq[q]) continue;
            const float* ctx = a.ctxs + ((size_t)s * a.Rtot + a.r0 + q * B + b) * DD;
            for (int d = tid; d < DD; d += NTB) c[q] += dc[q * DD + d] * __ldg(ctx + d);
            c[q] = warp_sum(c[q]);
        }
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) wred[warp][q] = c[q];
        }
    }
    __syncthreads();
    if (tid < NQ) {
        float c = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) c += wred[w][tid];
        c0s[tid] = c;
    }
    SUB_MARK(8);
    // ---- pass 1: da partials per 128-column block.  Warp (cb, fs), D2 frames in flight; G frames x NQ dot products reduced at once.
    {
        const int cb = warp & (NCB - 1), fs = warp / NCB;
        float4 dcq[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) dcq[q] = *reinterpret_cast<const float4*>(dc + q * DD + cb * 128 + lane * 4);
        constexpr int G = (32 / NQ) < 8 ? (32 / NQ) : 8;
        const float4* en = reinterpret_cast<const float4*>(a.enc + (size_t)b * T * DD) + cb * 32 + lane;
        float4* slot = reinterpret_cast<float4*>(myring) + lane;
        const int base = rs_base<32>(lane);
        const int nf = (t1 - t0 - fs + NFS - 1) / NFS;
#pragma unroll
        for (int i = 0; i < D2; ++i) {
            if (i < nf) cp_async16(slot + i * 32, en + (size_t)(t0 + fs + i * NFS) * (DD / 4));
            cp_async_commit();
        }
        for (int f0 = 0; f0 < nf; f0 += G) {
            float val[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) val[i] = 0.f;
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const int f = f0 + i;
                if (f < nf) {
                    cp_async_wait<D2 - 1>();
                    const int sl = f % D2;
                    const float4 e = slot[sl * 32];
                    if (f + D2 < nf) cp_async16(slot + sl * 32, en + (size_t)(t0 + fs + (f + D2) * NFS) * (DD / 4));
                    cp_async_commit();
#pragma unroll
                    for (int q = 0; q < NQ; ++q) val[i * NQ + q] = dot4(dcq[q], e);
                }
            }
            reduce_scatter<32>(val, lane);
            if (base < G * NQ) {
                const int i = base / NQ, q = base - i * NQ;
                const int f = f0 + i;
                if (f < nf) dap[(cb * NQ + q) * TILE_MAX + fs + f * NFS] = val[0];
            }
        }
        cp_async_wait<0>();
    }
    __syncthreads();
    SUB_MARK(9);
    // ---- ds = attn (da - c0); the saved raw score becomes the normalised weight (read by the encoder-gradient GEMM)
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        if (!actq[q]) continue;
        const size_t sb = (size_t)s * a.Rtot + a.r0 + q * B + b;
        float* arow = a.attn + sb * T + t0;
        float* dsrow = a.ds_all + sb * T + t0;
        const float M = Ms[q], iL = iLs[q], c0 = c0s[q];
        for (int j = tid; j < nt; j += NTB) {
            float da = 0.f;
#pragma unroll
            for (int cb = 0; cb < NCB; ++cb) da += dap[(cb * NQ + q) * TILE_MAX + j];
            const float aw = expf(arow[j] - M) * iL;
            arow[j] = aw;
            const float ds = aw * (da - c0);
            dsrow[j] = ds;
            dap[q * TILE_MAX + j] = ds;                              // (cb = 0 slot of query q is only read by this thread)
        }
    }
    __syncthreads();
    SUB_MARK(10);
    // ---- pass 2: dq.  Warp w takes the frames t0 + w, t0 + w + 12, ...; D1 frames in flight; a lane holds 8 of the 256 exp(2 Ep) values.
    float2 dq[NQ][4];                                            // per query: elements (0,1) (2,3) (128,129) (130,131) of the lane's columns
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) dq[q][k] = make_float2(0.f, 0.f);
    {
        const float4* ee = reinterpret_cast<const float4*>(a.Ee + (size_t)b * T * DA) + lane;
        float4* slot = reinterpret_cast<float4*>(myring) + lane;
        const int nf = (t1 - t0 - warp + NW - 1) / NW;
#pragma unroll
        for (int i = 0; i < D1; ++i) {
            if (i < nf) {
                const size_t t = t0 + warp + i * NW;
                cp_async16(slot + i * 64, ee + t * (DA / 4));
                cp_async16(slot + i * 64 + 32, ee + t * (DA / 4) + 32);
            }
            cp_async_commit();
        }
        for (int f0 = 0; f0 < nf; f0 += D1) {
#pragma unroll
            for (int i = 0; i < D1; ++i) {
                const int f = f0 + i;
                if (f >= nf) break;
                cp_async_wait<D1 - 1>();
                const float4 c0 = slot[i * 64], c1 = slot[i * 64 + 32];
                if (f + D1 < nf) {
                    const size_t t = t0 + warp + (f + D1) * NW;
                    cp_async16(slot + i * 64, ee + t * (DA / 4));
                    cp_async16(slot + i * 64 + 32, ee + t * (DA / 4) + 32);
                }
                cp_async_commit();
                const int j = warp + f * NW;
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    if (!actq[q]) continue;
                    const float ds = dap[q * TILE_MAX + j];
                    const float4 qa = *reinterpret_cast<const float4*>(Eq + q * DA + lane * 4);
                    const float4 qb = *reinterpret_cast<const float4*>(Eq + q * DA + 128 + lane * 4);
                    quad_grad(qa, c0, ds, dq[q][0], dq[q][1]);           // r (1 - r) = (1 - tanh^2) / 4
                    quad_grad(qb, c1, ds, dq[q][2], dq[q][3]);
                }
            }
        }
        cp_async_wait<0>();
    }
    SUB_MARK(11);
    // ---- cross-warp sums, one query at a time
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        if (!actq[q]) continue;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            *reinterpret_cast<float2*>(dqr + warp * DA + lane * 4 + 2 * i) = dq[q][i];
            *reinterpret_cast<float2*>(dqr + warp * DA + 128 + lane * 4 + 2 * i) = dq[q][2 + i];
        }
        __syncthreads();
        if (tid < DA) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) t += dqr[w * DA + tid];
            a.dq_part[((size_t)(q * B + b) * a.NS + js) * DA + tid] = t * S.v4[tid];
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int tk = atomicAdd(a.tickets + b, 1);
        is_last = (tk == a.NS - 1);
        if (is_last) a.tickets[b] = 0;
    }
    __syncthreads();
    SUB_MARK(12);
    if (!is_last) return;
    __threadfence();
    for (int i = tid; i < NQ * DA; i += NTB) {
        const int q = i / DA, k = i - q * DA;
        if (s >= g_Sq[q]) continue;
        const int r = q * B + b;
        float t = 0.f;
        for (int j = 0; j < a.NS; ++j) t += __ldcg(a.dq_part + ((size_t)r * a.NS + j) * DA + k);
        a.dq_all[((size_t)s * a.Rtot + a.r0 + r) * DA + k] = t;
    }
}

template <int NQ, bool TC>
__global__ void __launch_bounds__(NTB, 1) decm_bwd_kernel(DecMArgs a) {
    extern __shared__ __align__(16) float smem_f[];
    BwdSmem S;
    S.WT = smem_f;
    S.U = S.WT + WT_WORDS;
    S.red = S.U + BT * UP;
    S.Wq = S.red + BWD_RED_FLOATS;
    S.v4 = S.Wq + UPC * 2 * QP;
    const int tid = threadIdx.x, cta = blockIdx.x;
    unsigned int target = 0;
    if (tid < 8) g_Sq[tid] = tid == 0 ? a.Sq[0] : tid == 1 ? a.Sq[1] : tid == 2 ? a.Sq[2] : tid == 3 ? a.Sq[3] : tid == 4 ? a.Sq[4] : tid == 5 ? a.Sq[5] : tid == 6 ? a.Sq[6] : a.Sq[7];
    {
        const bool is_dx = cta < PG / 2;
        if (TC) {
            unsigned int* whi = reinterpret_cast<unsigned int*>(S.WT);
            unsigned int* wlo = whi + RPB * WTP;
            for (int i = tid; i < RPB * (K3 / 2); i += NTB) {
                const int r = i / (K3 / 2), pw = i - r * (K3 / 2);
                const float* row = is_dx ? a.W_ihT + (size_t)(DE + cta * RPB + r) * K3 : a.W_hhT + (size_t)((cta - PG / 2) * RPB + r) * K3;
                split2(__ldg(row + 2 * pw), __ldg(row + 2 * pw + 1), whi[r * WTP + pw], wlo[r * WTP + pw]);
            }
            for (int i = tid; i < K3 / 4; i += NTB)
                reinterpret_cast<float4*>(S.WT + 2 * RPB * WTP)[i] =
                    cta < DE ? __ldg(reinterpret_cast<const float4*>(a.W_ihT + (size_t)cta * K3) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
        for (int i = tid; i < (RPB + 1) * (K3 / 4); i += NTB) {
            const int r = i / (K3 / 4), k4 = i % (K3 / 4);
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < RPB) {
                const float* row = is_dx ? a.W_ihT + (size_t)(DE + cta * RPB + r) * K3 : a.W_hhT + (size_t)((cta - PG / 2) * RPB + r) * K3;
                w = __ldg(reinterpret_cast<const float4*>(row) + k4);
            } else if (cta < DE) {
                w = __ldg(reinterpret_cast<const float4*>(a.W_ihT + (size_t)cta * K3) + k4);
            }
            reinterpret_cast<float4*>(S.WT)[i] = w;
        }
        }
        for (int i = tid; i < BT * UP; i += NTB) S.U[i] = 0.f;
        for (int i = tid; i < UPC * DA; i += NTB)
            S.Wq[(i >> 7) * QP + (i & 127)] = a.W_hT[(size_t)cta * UPC * DA + i];
        if (tid < DA) S.v4[tid] = 4.f * a.v[tid];
    }
    __syncthreads();
    const int B = a.B, R = NQ * B;
    const int nchunks = (R + BT - 1) / BT;
    const int nitems = B * a.NS;
    unsigned long long prof_t = gtimer();
    for (int s = a.S - 1; s >= 0; --s) {
        bwd_gates_all(a, S, s, R);
        PROF_MARK(0);
        grid_sync(a.sync, target);
        PROF_MARK(1);
        if (TC) {
            // the next chunk's rows travel L2 -> registers while the current chunk's MMAs run
            float4 pre[BT];
            int ch = 0;
            while (ch < nchunks && !chunk_active(a, s, ch * BT, min(BT, R - ch * BT))) ++ch;
            if (ch < nchunks) bwd_chunk_load(a, s, ch * BT, min(BT, R - ch * BT), pre);
            while (ch < nchunks) {
                const int rb0 = ch * BT, nb = min(BT, R - rb0);
                int nx = ch + 1;
                while (nx < nchunks && !chunk_active(a, s, nx * BT, min(BT, R - nx * BT))) ++nx;
                bwd_gemv_phase<TC>(a, S, s, rb0, nb, pre, true, nx * BT, nx < nchunks ? min(BT, R - nx * BT) : 0);
                ch = nx;
            }
        } else {
            float4 none[BT];
            for (int ch = 0; ch < nchunks; ++ch) {
                const int rb0 = ch * BT, nb = min(BT, R - rb0);
                if (chunk_active(a, s, rb0, nb)) bwd_gemv_phase<TC>(a, S, s, rb0, nb, none, false, 0, 0);
            }
        }
        PROF_MARK(2);
        grid_sync(a.sync, target);
        PROF_MARK(3);
        for (int item = cta; item < nitems; item += PG) bwd_attn_item<NQ>(a, S, s, item / a.NS, item % a.NS);
        PROF_MARK(4);
        grid_sync(a.sync, target);
        PROF_MARK(5);
    }
    // ---- tail: dh_0 = dh_prev of step 0 + dq_0 W_h  -> a.dhq
    for (int ch = 0; ch < nchunks; ++ch) {
        const int rb0 = ch * BT, nb = min(BT, R - rb0);
        float* dqs = S.U;
        __syncthreads();
        for (int i = tid; i < nb * (DA / 4); i += NTB)
            *reinterpret_cast<float4*>(dqs + (i >> 5) * QP + (i & 31) * 4) = ldcg4(a.dq_all + ((size_t)a.r0 + rb0) * DA + (size_t)i * 4);
        __syncthreads();
        if (tid < 2 * UPC * BT) {
            const int br = tid >> 4, u = (tid >> 1) & 7, half = tid & 1;
            float dhq = 0.f;
            if (br < nb) {
                const float4* w4 = reinterpret_cast<const float4*>(S.Wq + (u * 2 + half) * QP);
                const float4* d4 = reinterpret_cast<const float4*>(dqs + (br * 2 + half) * QP);
#pragma unroll 8
                for (int i = 0; i < 32; ++i) dhq += dot4(w4[i], d4[i]);
            }
            dhq += __shfl_xor_sync(0xffffffffu, dhq, 1);
            if (half == 0 && br < nb) {
                const int j = cta * UPC + u;
                a.dhq[(size_t)(rb0 + br) * DD + j] = dhq + __ldcg(a.dh_carry + (size_t)(rb0 + br) * DD + j);
            }
        }
    }
}

// dlogits[s,row,:] = dlogp[b,bar,s,:] - exp(logp[b,bar,s,:]) * sum_v dlogp[b,bar,s,v]   (log_softmax backward), zero padded to VP;
// rows past their sequence (s >= Sq[q]) get zeros
__global__ void decm_dlogits_kernel(DecMArgs a) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int R = a.NQ * a.B;
    if (threadIdx.x < 8) g_Sq[threadIdx.x] = threadIdx.x == 0 ? a.Sq[0] : threadIdx.x == 1 ? a.Sq[1] : threadIdx.x == 2 ? a.Sq[2] : threadIdx.x == 3 ? a.Sq[3] : threadIdx.x == 4 ? a.Sq[4] : threadIdx.x == 5 ? a.Sq[5] : threadIdx.x == 6 ? a.Sq[6] : a.Sq[7];
    __syncthreads();
    if (row >= a.S * R) return;
    const int s = row / R, r = row - s * R;
    const int q = r / a.B, b = r - q * a.B;
    float* out = a.dlogits_all + ((size_t)s * a.Rtot + a.r0 + r) * a.VP;
    if (s >= g_Sq[q]) {
        for (int vi = lane; vi < a.VP; vi += 32) out[vi] = 0.f;
        return;
    }
    const size_t ro = (((size_t)b * a.bars + a.k0 + q) * a.max_steps + s) * a.V;
    float sg = 0.f;
    for (int vi = lane; vi < a.V; vi += 32) sg += __ldg(a.dlogp + ro + vi);
    sg = warp_sum(sg);
    for (int vi = lane; vi < a.VP; vi += 32) {
        float d = 0.f;
        if (vi < a.V) d = __ldg(a.dlogp + ro + vi) - expf(__ldg(a.logp + ro + vi)) * sg;
        out[vi] = d;
    }
}

// Deferred accumulations over steps and queries (off the sequential chain):
//   dEp[b,t,k] = sum_{s,q} ds v_k (1 - u^2),  dv_k = sum_{s,q,b,t} ds u,   u = 1 - 2 r,  1 - u^2 = 4 r (1 - r),  r = 1 / (1 + Eq Ee)
constexpr int DEF_FPW = 2, DEF_WARPS = 8, DEF_FPB = DEF_FPW * DEF_WARPS;
__global__ void __launch_bounds__(DEF_WARPS * 32) decm_attn_deferred_kernel(DecMArgs a) {
    __shared__ float acc[DEF_WARPS][DA];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = a.T, B = a.B;
    if (tid < 8) g_Sq[tid] = tid == 0 ? a.Sq[0] : tid == 1 ? a.Sq[1] : tid == 2 ? a.Sq[2] : tid == 3 ? a.Sq[3] : tid == 4 ? a.Sq[4] : tid == 5 ? a.Sq[5] : tid == 6 ? a.Sq[6] : a.Sq[7];
    __syncthreads();
    const int tA = blockIdx.x * DEF_FPB + warp * DEF_FPW;
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(a.v) + lane), v1 = __ldg(reinterpret_cast<const float4*>(a.v) + 32 + lane);
    const float vk[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    float ev[DEF_FPW][8], dE[DEF_FPW][8], dv[8];
    bool ok[DEF_FPW];
#pragma unroll
    for (int f = 0; f < DEF_FPW; ++f) {
        ok[f] = tA + f < T;
        const float4* ep = reinterpret_cast<const float4*>(a.Ee + ((size_t)b * T + min(tA + f, T - 1)) * DA);
        const float4 e0 = __ldg(ep + lane), e1 = __ldg(ep + 32 + lane);
        ev[f][0] = e0.x; ev[f][1] = e0.y; ev[f][2] = e0.z; ev[f][3] = e0.w; ev[f][4] = e1.x; ev[f][5] = e1.y; ev[f][6] = e1.z; ev[f][7] = e1.w;
#pragma unroll
        for (int i = 0; i < 8; ++i) dE[f][i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) dv[i] = 0.f;
    if (tA < T) {
        for (int q = 0; q < a.NQ; ++q) {
            const int Sq = g_Sq[q];
            for (int s = 0; s < Sq; ++s) {
                const size_t sb = (size_t)s * a.Rtot + a.r0 + q * B + b;
                const float4 q0 = __ldg(reinterpret_cast<const float4*>(a.eqs + sb * DA) + lane);
                const float4 q1 = __ldg(reinterpret_cast<const float4*>(a.eqs + sb * DA) + 32 + lane);
                const float qk[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
                for (int f = 0; f < DEF_FPW; ++f) {
                    const float ds = ok[f] ? __ldg(a.ds_all + sb * T + tA + f) : 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float r = rcp_fast(fmaf(qk[i], ev[f][i], 1.f));
                        dE[f][i] = fmaf(ds, fmaf(-r, r, r), dE[f][i]);
                        dv[i] = fmaf(ds, fmaf(-2.f, r, 1.f), dv[i]);
                    }
                }
            }
        }
#pragma unroll
        for (int f = 0; f < DEF_FPW; ++f) {
            if (!ok[f]) continue;
            float4* dep = reinterpret_cast<float4*>(a.dEp + ((size_t)b * T + tA + f) * DA);
            dep[lane] = make_float4(4.f * vk[0] * dE[f][0], 4.f * vk[1] * dE[f][1], 4.f * vk[2] * dE[f][2], 4.f * vk[3] * dE[f][3]);
            dep[32 + lane] = make_float4(4.f * vk[4] * dE[f][4], 4.f * vk[5] * dE[f][5], 4.f * vk[6] * dE[f][6], 4.f * vk[7] * dE[f][7]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[warp][lane * 4 + i] = dv[i]; acc[warp][128 + lane * 4 + i] = dv[4 + i]; }
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < DEF_WARPS; ++w) t += acc[w][tid];
    a.dv_part[((size_t)b * gridDim.x + blockIdx.x) * DA + tid] = t;
}

// Ee = exp(2 Ep), elementwise
__global__ void decm_exp2x_kernel(const float4* __restrict__ x, float4* __restrict__ y, long long n4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = __ldg(x + i);
    y[i] = make_float4(exp2x(v.x), exp2x(v.y), exp2x(v.z), exp2x(v.w));
}

int check_args(const DecMArgs& a, bool bwd) {
    if (a.NQ < 1 || a.NQ > NQMAX || a.NQ > 8) return -2;
    if (a.V + DA > PG * CR || a.V > 256 || a.sync == nullptr) return -2;
    // frames per attention item: the forward's score table holds TILE_MAX * (NQMAX / NQ) per query, the reverse pass TILE_MAX
    if (a.tile > TILE_MAX * (bwd ? 1 : NQMAX / a.NQ) || a.NS * a.tile < a.T || a.NS < 1) return -3;
    for (int q = 0; q < a.NQ; ++q)
        if (a.Sq[q] < 1 || a.Sq[q] > a.S) return -4;
    return 0;
}

template <int NQ, bool TC>
int launch_fwd2(DecMArgs& a, int eos_id, cudaStream_t st) {
    const size_t smem = (size_t)FWD_SMEM_FLOATS * sizeof(float);
    PA2S_TRY(cudaFuncSetAttribute(decm_fwd_kernel<NQ, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* kargs[] = {(void*)&a, (void*)&eos_id};
    PA2S_TRY(cudaLaunchCooperativeKernel((const void*)decm_fwd_kernel<NQ, TC>, dim3(PG), dim3(NT), kargs, smem, st));
    PA2S_COUNT_LAUNCH();
    return 0;
}
template <int NQ>
int launch_fwd(DecMArgs& a, int eos_id, cudaStream_t st) {
    return a.tc ? launch_fwd2<NQ, true>(a, eos_id, st) : launch_fwd2<NQ, false>(a, eos_id, st);
}
template <int NQ, bool TC>
int launch_bwd2(DecMArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)BWD_SMEM_FLOATS * sizeof(float);
    PA2S_TRY(cudaFuncSetAttribute(decm_bwd_kernel<NQ, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* kargs[] = {(void*)&a};
    PA2S_TRY(cudaLaunchCooperativeKernel((const void*)decm_bwd_kernel<NQ, TC>, dim3(PG), dim3(NTB), kargs, smem, st));
    PA2S_COUNT_LAUNCH();
    return 0;
}
template <int NQ>
int launch_bwd(DecMArgs& a, cudaStream_t st) {
    return a.tc ? launch_bwd2<NQ, true>(a, st) : launch_bwd2<NQ, false>(a, st);
}

}  // namespace

PA2S_API int pa2s_decm_args_size(void) { return (int)sizeof(DecMArgs); }
PA2S_API int pa2s_decm_max_queries(void) { return NQMAX; }
PA2S_API int pa2s_decm_tile_max(void) { return TILE_MAX; }
PA2S_API int pa2s_decm_grid(void) { return PG; }
PA2S_API int pa2s_decm_deferred_blocks(int T) { return (T + DEF_FPB - 1) / DEF_FPB; }

PA2S_API int pa2s_exp2x(void* stream, const float* x, float* y, long long n) {
    if (n % 4 != 0) return -1;
    decm_exp2x_kernel<<<ceil_div(n / 4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)x, (float4*)y, n / 4);
    PA2S_CHECK_LAST();
    return 0;
}

// All steps of NQ sequences x B clips in one cooperative launch.
PA2S_API int pa2s_decm_fwd(void* stream, const void* args, int sos_id, int eos_id) {
    DecMArgs a = *reinterpret_cast<const DecMArgs*>(args);
    cudaStream_t st = (cudaStream_t)stream;
    if (a.B <= 0 || a.S <= 0) return 0;
    const int rc = check_args(a, false);
    if (rc != 0) return rc;
    decm_init_kernel<<<ceil_div(a.NQ * a.B * DE, 128), 128, 0, st>>>(a, sos_id);
    PA2S_CHECK_LAST();
    switch (a.NQ) {
        case 1: return launch_fwd<1>(a, eos_id, st);
        case 2: return launch_fwd<2>(a, eos_id, st);
        case 3: return launch_fwd<3>(a, eos_id, st);
        case 4: return launch_fwd<4>(a, eos_id, st);
        case 5: return launch_fwd<5>(a, eos_id, st);
    }
    return -2;
}

// log-softmax backward of every (step, row) -> dlogits_all; the caller then forms dhc_all = dlogits_all @ W_out with one GEMM
PA2S_API int pa2s_decm_dlogits(void* stream, const void* args) {
    DecMArgs a = *reinterpret_cast<const DecMArgs*>(args);
    if (a.B <= 0 || a.S <= 0) return 0;
    decm_dlogits_kernel<<<ceil_div((long long)a.S * a.NQ * a.B, 8), 256, 0, (cudaStream_t)stream>>>(a);
    PA2S_CHECK_LAST();
    return 0;
}

// Reverse pass over all saved steps of all rows: the sequential chain (cooperative kernel) ...
PA2S_API int pa2s_decm_bwd_chain(void* stream, const void* args) {
    DecMArgs a = *reinterpret_cast<const DecMArgs*>(args);
    cudaStream_t st = (cudaStream_t)stream;
    if (a.B <= 0 || a.S <= 0) return 0;
    const int rc = check_args(a, true);
    if (rc != 0) return rc;
    if (a.dhc_all == nullptr || a.ds_all == nullptr || a.eqs == nullptr || a.ml == nullptr) return -2;
    switch (a.NQ) {
        case 1: return launch_bwd<1>(a, st);
        case 2: return launch_bwd<2>(a, st);
        case 3: return launch_bwd<3>(a, st);
        case 4: return launch_bwd<4>(a, st);
        case 5: return launch_bwd<5>(a, st);
    }
    return -2;
}
// ... and the parallel dEp / dv accumulation over (step, query) (reads ds_all, eqs, Ee, v)
PA2S_API int pa2s_decm_bwd_deferred(void* stream, const void* args) {
    DecMArgs a = *reinterpret_cast<const DecMArgs*>(args);
    if (a.B <= 0 || a.S <= 0) return 0;
    if (a.ds_all == nullptr || a.dEp == nullptr || a.dv_part == nullptr) return -2;
    decm_attn_deferred_kernel<<<dim3(ceil_div(a.T, DEF_FPB), a.B), DEF_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    PA2S_CHECK_LAST();
    return 0;
}
