// Persistent note-level decoder (NoteDecoder.decode_notes, models.py:366-420): ALL steps of one (bar, staff) in ONE
// cooperative kernel of 64 CTAs.  The GRU / output / query weights are sliced across the CTAs and stay in shared
// memory for the whole sequence, the recurrent state of every clip stays in shared memory between the phases of a
// step, and the CTAs meet at three grid barriers per step:
//
//   A(s)  attention over the encoder memory (clip x frame-range items, last-arriver combine)   [needs q_s]
//         + D(s-1): log-softmax / argmax / teacher forcing / next-token embedding / EOS        [needs logits_{s-1}]
//   B(s)  GRU cell 528 -> 512: each CTA owns 8 hidden units = 24 gate rows x 1040 weights      [needs ctx_s, tok_s]
//   C(s)  logits_s = W_out [h'; ctx] + b and q_{s+1} = W_h h': 7 of the 429 rows per CTA       [needs h'_s]
//
// The reverse kernel has the same shape: P1 dhq = dq_{s+1} W_h | P2 GRU backward (W^T row-sliced) | P3 attention
// backward.  Everything that does not sit on the sequential chain is hoisted out of the loop: the softmax/out-projection
// gradient is one GEMM before it, dEp / dv (accumulations over steps) one kernel after it, weight gradients GEMMs.
#include "dec_args.cuh"
#include <cooperative_groups.h>

namespace {

constexpr int PG = 64;                 // CTAs of a persistent decoder grid (two staves run concurrently: 128 of 148 SMs)
#ifndef PA2S_DEC_NT
#define PA2S_DEC_NT 384
#endif
#ifndef PA2S_DEC_PD
#define PA2S_DEC_PD 1
#endif
// frames per warp in flight in the attention phases (register ring depth).  Measured (tools/prof_decoder.py, B=16, S=80, forward /
// backward us per step): PD=1 32.4 / 32.5, PD=2 36.5 / 33.3, PD=3 38.3 / 49.3, PD=4 51.4 / 67.4 -- a second frame costs 24 registers
// per thread, which at 384 threads x 168 registers spills inside the frame loop (ptxas: 0 / 156 / 496 / 816 bytes of spill stores).
constexpr int PD = PA2S_DEC_PD;
constexpr int NT = PA2S_DEC_NT;        // threads per CTA: 12 warps stream the attention memory, the first 8 own the GEMV columns
constexpr int NW = NT / 32;
constexpr int UPC = DD / PG;           // hidden units per CTA (8)
constexpr int GR = 3 * UPC;            // gate rows per CTA (24): r[8], z[8], n[8]
constexpr int CR = 7;                  // phase-C rows per CTA: PG*CR = 448 >= V + DA
constexpr int XP = 2 * DD + DE;        // per-clip state row in shared memory: [h (512) | ctx (512) | tok (16)]
constexpr int XP4 = XP / 4;
constexpr int KM4 = 2 * DD / 4;        // float4 columns of the main part (256 = one per thread)
static_assert(KM4 <= NT, "one float4 column of [h|ctx] per thread of the first 8 warps");
constexpr int TEAMS = NT / KM4;        // groups of 256 threads that split the clip groups of the GEMV phases (1 at 384 threads, 2 at 512)

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// streaming 16-byte load of the encoder memory (read once per step by exactly one warp)
#ifndef PA2S_DEC_LD
#define PA2S_DEC_LD 0
#endif
__device__ __forceinline__ float4 lds4(const float4* p) {
#if PA2S_DEC_LD == 0
    return __ldg(p);
#else
    float4 v;
#if PA2S_DEC_LD == 1
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
#elif PA2S_DEC_LD == 2
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
#else
    asm volatile("ld.global.nc.L1::evict_first.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
#endif
    return v;
#endif
}

// Grid barrier for a co-resident grid: monotonically increasing arrival counter.  A watchdog turns a lost CTA into a
// LOUD failure instead of a hang: it records the error flag (sync[1], for a post-mortem read) and traps, so the launch
// fails with a CUDA error at the next synchronisation and no later kernel (Adadelta in particular) consumes the
// inconsistent state the barrier would otherwise have let through.
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

// optional phase timing (CTA 0, thread 0): prof[i] += ns since the previous mark
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define PROF_MARK(i)                                                                 \
    do {                                                                             \
        if (a.prof != nullptr && threadIdx.x == 0 && blockIdx.x == 0) {              \
            const unsigned long long now_ = gtimer();                                \
            a.prof[i] += now_ - prof_t;                                              \
            prof_t = now_;                                                           \
        }                                                                            \
    } while (0)

// sub-phase timing inside a device function (CTA 0, thread 0): prof[i] += ns since sub_t0, without moving the phase clock
#define SUB_BEGIN() unsigned long long sub_t = (a.prof != nullptr && threadIdx.x == 0 && blockIdx.x == 0) ? gtimer() : 0ull
#define SUB_MARK(i)                                                                  \
    do {                                                                             \
        if (a.prof != nullptr && threadIdx.x == 0 && blockIdx.x == 0) {              \
            const unsigned long long now_ = gtimer();                                \
            a.prof[i] += now_ - sub_t;                                               \
            sub_t = now_;                                                            \
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

// phase B: (4 clip groups x 8 warps) x (24 rows x 4 clips); phase C: 32 x 32.  The per-warp partial contexts of phase A (NW x 512
// floats) live in the CONTEXT columns of xs instead, which are dead between phase C of one step and the staging of phase B of the
// next (row w of xs = warp w: needs NW <= BT); that keeps 16 warps within the 227 KB of shared memory.
constexpr int RED_FLOATS = 4 * 8 * GR * 4;
static_assert(NW <= BT && RED_FLOATS >= 32 * 32, "partial contexts alias xs rows; phase C scratch");
struct FwdSmem {
    float* Wg;      // [GR][1024]   rows g*8+u: [W_hh row | W_ih row, context columns]
    float* Wc;      // [CR][1024]   W_out rows / W_h rows (zero beyond 512) / zero rows
    float* Wtok;    // [GR][16]     W_ih row, token columns
    float* bias;    // [4][8]       b_r (ih+hh), b_z (ih+hh), b_in, b_hn
    float* bc;      // [8]          b_out of the phase-C rows (0 for query rows)
    float* xs;      // [BT][XP]
    float* red;     // RED_FLOATS: cross-warp reduction scratch of phases B and C
    float* qv;      // [DA]
    float* vv;      // [DA]
};

__device__ __forceinline__ FwdSmem carve(float* sm) {
    FwdSmem s;
    s.Wg = sm; sm += GR * 2 * DD;
    s.Wc = sm; sm += CR * 2 * DD;
    s.Wtok = sm; sm += GR * DE;
    s.bias = sm; sm += 32;
    s.bc = sm; sm += 8;
    s.xs = sm; sm += BT * XP;
    s.red = sm; sm += RED_FLOATS;
    s.qv = sm; sm += DA;
    s.vv = sm; sm += DA;
    return s;
}
constexpr int FWD_SMEM_FLOATS = GR * 2 * DD + CR * 2 * DD + GR * DE + 32 + 8 + BT * XP + RED_FLOATS + 2 * DA;

// ------------------------------------------------------------------------------------------------ phase A
// Attention for item (clip b, frame range js) of step s in ONE pass over the frames: every warp keeps an online-softmax
// partial (running max, sum, 512-wide context) while the Ep / enc rows of its next two frames are already in flight;
// warps are combined in shared memory, CTAs of a clip by the last one to arrive.
struct AttnRows { float4 e0[1], e1[1], en[1][4]; };

__device__ void attn_item(const DecArgs& a, const FwdSmem& S, int s, int b, int js) {
    __shared__ float wm[NW], wl[NW];
    __shared__ int is_last;
    constexpr int FU = 1;                                       // frames per warp iteration (one more is in flight)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = a.T;
    const int t0 = js * a.tile, t1 = min(T, t0 + a.tile);
    const float* q = a.qs + ((size_t)hslot(a, s) * a.B + b) * DA;
    SUB_BEGIN();
    __syncthreads();
    if (tid < DA) S.qv[tid] = __ldcg(q + tid);
    __syncthreads();
    SUB_MARK(8);
    float* araw = a.attn + ((size_t)slot(a, s) * a.B + b) * T;
    float m = -INFINITY, l = 0.f;
    float4 cacc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) cacc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    {
        const float4 q0 = *reinterpret_cast<const float4*>(S.qv + lane * 4);
        const float4 q1 = *reinterpret_cast<const float4*>(S.qv + 128 + lane * 4);
        const float4 v0 = *reinterpret_cast<const float4*>(S.vv + lane * 4);
        const float4 v1 = *reinterpret_cast<const float4*>(S.vv + 128 + lane * 4);
        auto load = [&](AttnRows& r, int tb) {
#pragma unroll
            for (int u = 0; u < FU; ++u) {
                const int t = min(tb + u, t1 - 1);
                const float4* ep = reinterpret_cast<const float4*>(a.Ep + ((size_t)b * T + t) * DA);
                r.e0[u] = lds4(ep + lane);
                r.e1[u] = lds4(ep + 32 + lane);
                const float4* en = reinterpret_cast<const float4*>(a.enc + ((size_t)b * T + t) * DD);
#pragma unroll
                for (int j = 0; j < 4; ++j) r.en[u][j] = lds4(en + j * 32 + lane);
            }
        };
        // PD frames per warp in flight: slot i of the register ring is refilled as soon as its frame has been consumed
        AttnRows ring[PD];
        const int tb0 = t0 + warp * FU;
#pragma unroll
        for (int i = 0; i < PD; ++i)
            if (tb0 + i * NW * FU < t1) load(ring[i], tb0 + i * NW * FU);
        for (int tb = tb0; tb < t1; tb += PD * NW * FU) {
#pragma unroll
            for (int i = 0; i < PD; ++i) {
                const int tc = tb + i * NW * FU;
                if (tc >= t1) break;                            // warp-uniform
                const AttnRows cur = ring[i];
                if (tc + PD * NW * FU < t1) load(ring[i], tc + PD * NW * FU);
#pragma unroll
                for (int u = 0; u < FU; ++u) {
                    float e = v0.x * tanh_fast(q0.x + cur.e0[u].x) + v0.y * tanh_fast(q0.y + cur.e0[u].y) + v0.z * tanh_fast(q0.z + cur.e0[u].z) +
                              v0.w * tanh_fast(q0.w + cur.e0[u].w) + v1.x * tanh_fast(q1.x + cur.e1[u].x) + v1.y * tanh_fast(q1.y + cur.e1[u].y) +
                              v1.z * tanh_fast(q1.z + cur.e1[u].z) + v1.w * tanh_fast(q1.w + cur.e1[u].w);
                    e = warp_sum(e);
                    if (tc + u < t1) {                          // warp-uniform
                        if (lane == 0) araw[tc + u] = e;        // raw score; normalised by the combining CTA
                        if (e > m) {
                            const float sc = expf(m - e);       // 0 on the first frame (m = -inf)
                            l *= sc;
#pragma unroll
                            for (int j = 0; j < 4; ++j) { cacc[j].x *= sc; cacc[j].y *= sc; cacc[j].z *= sc; cacc[j].w *= sc; }
                            m = e;
                        }
                        const float p = expf(e - m);
                        l += p;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            cacc[j].x = fmaf(p, cur.en[u][j].x, cacc[j].x);
                            cacc[j].y = fmaf(p, cur.en[u][j].y, cacc[j].y);
                            cacc[j].z = fmaf(p, cur.en[u][j].z, cacc[j].z);
                            cacc[j].w = fmaf(p, cur.en[u][j].w, cacc[j].w);
                        }
                    }
                }
            }
        }
    }
    SUB_MARK(9);
    float* part = S.xs + DD;                                    // per-warp partial contexts: context columns of xs row `warp`
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(part + warp * XP + j * 128 + lane * 4) = cacc[j];
    if (lane == 0) { wm[warp] = m; wl[warp] = l; }
    __syncthreads();
    float M = -INFINITY, L = 0.f, c0 = 0.f, c1 = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) M = fmaxf(M, wm[w]);
    if (tid < DD / 2) {
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float wgt = (wm[w] == -INFINITY) ? 0.f : expf(wm[w] - M);
            L = fmaf(wl[w], wgt, L);
            const float2 pw = *reinterpret_cast<const float2*>(part + w * XP + 2 * tid);
            c0 = fmaf(pw.x, wgt, c0);
            c1 = fmaf(pw.y, wgt, c1);
        }
        reinterpret_cast<float2*>(a.pc + ((size_t)b * a.NS + js) * DD)[tid] = make_float2(c0, c1);
        if (tid == 0) { a.pm[b * a.NS + js] = M; a.pl[b * a.NS + js] = L; }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int tk = atomicAdd(a.tickets + b, 1);
        is_last = (tk == a.NS - 1);
        if (is_last) a.tickets[b] = 0;
    }
    __syncthreads();
    SUB_MARK(10);
    if (!is_last) return;
    __threadfence();
    M = -INFINITY;
    for (int j = 0; j < a.NS; ++j) M = fmaxf(M, __ldcg(a.pm + b * a.NS + j));
    L = 0.f;
    for (int j = 0; j < a.NS; ++j) {
        const float mj = __ldcg(a.pm + b * a.NS + j);
        L = fmaf(__ldcg(a.pl + b * a.NS + j), (mj == -INFINITY) ? 0.f : expf(mj - M), L);
    }
    const float invL = 1.f / L;
    if (tid < DD / 2) {
        c0 = 0.f; c1 = 0.f;
        for (int j = 0; j < a.NS; ++j) {
            const float mj = __ldcg(a.pm + b * a.NS + j);
            const float wgt = (mj == -INFINITY) ? 0.f : expf(mj - M);
            const float2 pj = __ldcg(reinterpret_cast<const float2*>(a.pc + ((size_t)b * a.NS + j) * DD) + tid);
            c0 = fmaf(pj.x, wgt, c0);
            c1 = fmaf(pj.y, wgt, c1);
        }
        float* cs = a.ctxs + ((size_t)slot(a, s) * a.B + b) * DD;
        reinterpret_cast<float2*>(cs)[tid] = make_float2(c0 * invL, c1 * invL);
    }
    for (int t = tid; t < T; t += NT) araw[t] = expf(__ldcg(araw + t) - M) * invL;
    SUB_MARK(11);
}

// ------------------------------------------------------------------------------------------------ phase D
// Finalise step s for clip b: log-softmax row, greedy token, teacher forcing, EOS bookkeeping, next input embedding.
__device__ void finalize_step(const DecArgs& a, int s, int b, int eos_id) {
    __shared__ float redf[NW];
    __shared__ int redi[NW];
    __shared__ int s_tok;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int V = a.V;
    __syncthreads();
    const float x = tid < V ? __ldcg(a.logits + (size_t)b * a.VP + tid) : -INFINITY;
    float m = x; int idx = tid < V ? tid : 0x7fffffff;          // first index of the maximum, like torch.argmax
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (om > m || (om == m && oi < idx)) { m = om; idx = oi; }
    }
    if (lane == 0) { redf[warp] = m; redi[warp] = idx; }
    __syncthreads();
    m = redf[0]; idx = redi[0];
#pragma unroll
    for (int i = 1; i < NW; ++i)
        if (redf[i] > m || (redf[i] == m && redi[i] < idx)) { m = redf[i]; idx = redi[i]; }
    __syncthreads();
    float e = tid < V ? expf(x - m) : 0.f;
    e = warp_sum(e);
    if (lane == 0) redf[warp] = e;
    __syncthreads();
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) sum += redf[i];
    const float lse = m + logf(sum);
    if (tid < V) a.logp[((size_t)b * a.max_steps + s) * V + tid] = x - lse;
    if (tid == 0) {
        const long long g = a.gt != nullptr ? a.gt[(size_t)b * a.max_steps + s] : -1;
        const bool tf = (!a.inference) && a.use_gt != nullptr && a.use_gt[s] != 0 && a.gt != nullptr;
        const int tok = tf ? (int)g : idx;
        s_tok = tok;
        const bool hit = a.gt != nullptr ? (g == eos_id) : (idx == eos_id);
        if (hit) {
            a.lengths[b] = s + 1;
            if (a.eos[b] == 0) { a.eos[b] = 1; atomicAdd(a.counters, 1); }
        }
        if (b == 0) atomicAdd(a.counters + 1, 1);
        if (a.save && s + 1 <= a.S) a.toks[(size_t)(s + 1) * a.B + b] = tok;
    }
    __syncthreads();
    if (tid < DE && s + 1 < a.S) {
        const float mk = a.mask != nullptr ? a.mask[((size_t)(s + 1) * a.B + b) * DE + tid] : 1.f;
        const float xv = a.emb[(size_t)s_tok * DE + tid] * mk;
        a.xbuf[(size_t)b * DX + tid] = xv;
        if (a.save) a.xtok[((size_t)(s + 1) * a.B + b) * DE + tid] = xv;
    }
}

// ------------------------------------------------------------------------------------------------ phase B
// GRU cell for the CTA's 8 hidden units and the clips [bb0, bb0+nb) staged in xs.
__device__ void gru_phase(const DecArgs& a, const FwdSmem& S, int s, int bb0, int nb) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int col = tid & (KM4 - 1), warp = col >> 5, team = tid / KM4;      // TEAMS groups of 256 threads share the clip groups
    const float4* xs4 = reinterpret_cast<const float4*>(S.xs);
    const float4* wg4 = reinterpret_cast<const float4*>(S.Wg);
    const int base = rs_base<UPC * 4>(lane);
#pragma unroll 1
    for (int cg = team; cg < BT / 4; cg += TEAMS) {
        if (cg * 4 >= nb || team >= TEAMS) break;
        float4 xv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) xv[c] = xs4[(cg * 4 + c) * XP4 + col];
        float* dst = S.red + (cg * 8 + warp) * (GR * 4) + base;
#pragma unroll
        for (int g = 0; g < 3; ++g) {                  // one gate (8 rows x 4 clips) at a time keeps the register tile small
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
    __syncthreads();
    if (tid < UPC * BT) {
        const int b = tid >> 3, u = tid & 7;
        if (b < nb) {
            const int cg = b >> 2, c = b & 3;
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
                const float* xt = S.xs + b * XP + 2 * DD;
#pragma unroll
                for (int k = 0; k < DE; ++k) hi = fmaf(wt[k], xt[k], hi);
                if (g < 2) g3[g] = lo + hi;
                else { g3[2] = hi; nh = lo; }
            }
            const int j = blockIdx.x * UPC + u, bg = bb0 + b;
            const float r = sigmoidf_(g3[0] + S.bias[u]);
            const float z = sigmoidf_(g3[1] + S.bias[8 + u]);
            const float hnl = nh + S.bias[24 + u];
            const float n = tanhf(g3[2] + S.bias[16 + u] + r * hnl);
            const float hp = S.xs[b * XP + j];
            const float hn = (1.f - z) * n + z * hp;
            a.hs[((size_t)hslot(a, s + 1) * a.B + bg) * DD + j] = hn;
            if (a.save) {
                float* gs = a.gates + ((size_t)s * a.B + bg) * 4 * DD + j;
                gs[0] = r; gs[DD] = z; gs[2 * DD] = n; gs[3 * DD] = hnl;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ phase C
// logits and next query for the CTA's 7 rows and the clips staged in xs ([h' | ctx]).
__device__ void out_phase(const DecArgs& a, const FwdSmem& S, int qslot, int bb0, int nb, bool only_q) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int col = tid & (KM4 - 1), warp = col >> 5, team = tid / KM4;
    const float4* xs4 = reinterpret_cast<const float4*>(S.xs);
    const float4* wc4 = reinterpret_cast<const float4*>(S.Wc);
    float4 wr[CR];
#pragma unroll
    for (int r = 0; r < CR; ++r) wr[r] = wc4[r * KM4 + col];
    const int base = rs_base<32>(lane);
#pragma unroll 1
    for (int cg = team; cg < BT / 4; cg += TEAMS) {
        if (cg * 4 >= nb || team >= TEAMS) break;
        float acc[32];                                 // 8 rows (7 real) x 4 clips
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
    __syncthreads();
    if (tid < CR * BT) {
        const int rr = tid >> 4, b = tid & 15;
        const int rg = blockIdx.x * CR + rr;
        if (b < nb && rg < a.V + DA) {
            const int cg = b >> 2, c = b & 3;
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += S.red[(cg * 8 + w) * 32 + rr * 4 + c];
            if (rg < a.V) {
                if (!only_q) a.logits[(size_t)(bb0 + b) * a.VP + rg] = v + S.bc[rr];
            } else {
                a.qs[((size_t)qslot * a.B + bb0 + b) * DA + (rg - a.V)] = v;
            }
        }
    }
}

// stage clip rows into xs: parts bit 0 = h (from hs[hs_slot]), bit 1 = ctx (from ctxs[ctx_slot]), bit 2 = token embedding
__device__ void stage_xs(const DecArgs& a, const FwdSmem& S, int bb0, int nb, int parts, int hs_slot, int ctx_slot) {
    float4* xs4 = reinterpret_cast<float4*>(S.xs);
    const int tid = threadIdx.x;
    constexpr int NLD = (BT * (DD / 4) + NT - 1) / NT;          // float4 loads per thread and part
    const int n4 = nb * (DD / 4);
    // all global loads of the call are issued before the first shared-memory store: one L2 round trip per part
    float4 tk = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool has_tk = (parts & 4) && tid < nb * (DE / 4);
    if (has_tk) tk = ldcg4(a.xbuf + (size_t)(bb0 + (tid >> 2)) * DX + (tid & 3) * 4);
#pragma unroll
    for (int part = 0; part < 2; ++part) {
        if (!(parts & (1 << part))) continue;
        const float* src = part == 0 ? a.hs + ((size_t)hs_slot * a.B + bb0) * DD : a.ctxs + ((size_t)ctx_slot * a.B + bb0) * DD;
        float4 v[NLD];
#pragma unroll
        for (int j = 0; j < NLD; ++j) {
            const int i = tid + j * NT;
            if (i < n4) v[j] = ldcg4(src + (size_t)i * 4);
        }
#pragma unroll
        for (int j = 0; j < NLD; ++j) {
            const int i = tid + j * NT;
            if (i < n4) xs4[(i >> 7) * XP4 + part * 128 + (i & 127)] = v[j];
        }
    }
    if (has_tk) xs4[(tid >> 2) * XP4 + 256 + (tid & 3)] = tk;
}

__global__ void __launch_bounds__(NT, 1) dec_persist_fwd_kernel(DecArgs a, int eos_id) {
    extern __shared__ __align__(16) float smem_f[];
    const FwdSmem S = carve(smem_f);
    const int tid = threadIdx.x, cta = blockIdx.x;
    unsigned int target = 0;

    // ---- one-time: this CTA's weight slices -> shared memory
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
    if (tid < UPC) {
        const int j = cta * UPC + tid;
        S.bias[tid] = a.b_ih[j] + a.b_hh[j];
        S.bias[8 + tid] = a.b_ih[DD + j] + a.b_hh[DD + j];
        S.bias[16 + tid] = a.b_ih[2 * DD + j];
        S.bias[24 + tid] = a.b_hh[2 * DD + j];
        const int rg = cta * CR + tid;
        S.bc[tid] = (tid < CR && rg < a.V) ? a.b_out[rg] : 0.f;
    }
    if (tid < DA) S.vv[tid] = a.v[tid];
    for (int i = tid; i < BT * XP; i += NT) S.xs[i] = 0.f;
    __syncthreads();

    const int B = a.B;
    const int nchunks = (B + BT - 1) / BT;
    const bool resident = nchunks == 1;            // the recurrent state of all clips stays in shared memory across phases
    const int nitems = B * a.NS;

    unsigned long long prof_t = gtimer();
    // ---- prologue: q_0 = W_h h_0
    for (int ch = 0; ch < nchunks; ++ch) {
        const int bb0 = ch * BT, nb = min(BT, B - bb0);
        __syncthreads();
        stage_xs(a, S, bb0, nb, 1, hslot(a, 0), 0);
        __syncthreads();
        out_phase(a, S, hslot(a, 0), bb0, nb, true);
    }
    grid_sync(a.sync, target);
    PROF_MARK(6);

    int s = 0;
    for (; s < a.S; ++s) {
        // ---- A(s) + D(s-1)
        for (int item = cta; item < nitems; item += PG) {
            const int b = item / a.NS, js = item - b * a.NS;
            if (js == 0 && s > 0) finalize_step(a, s - 1, b, eos_id);
            attn_item(a, S, s, b, js);
        }
        PROF_MARK(0);
        grid_sync(a.sync, target);
        PROF_MARK(1);
        if (a.inference && __ldcg(a.counters) >= B) break;        // every clip has emitted <eos> (models.py:418-419)
        // ---- B(s)
        for (int ch = 0; ch < nchunks; ++ch) {
            const int bb0 = ch * BT, nb = min(BT, B - bb0);
            __syncthreads();
            stage_xs(a, S, bb0, nb, resident ? 6 : 7, hslot(a, s), slot(a, s));
            __syncthreads();
            gru_phase(a, S, s, bb0, nb);
        }
        PROF_MARK(2);
        grid_sync(a.sync, target);
        PROF_MARK(3);
        // ---- C(s)
        for (int ch = 0; ch < nchunks; ++ch) {
            const int bb0 = ch * BT, nb = min(BT, B - bb0);
            __syncthreads();
            stage_xs(a, S, bb0, nb, resident ? 1 : 3, hslot(a, s + 1), slot(a, s));
            __syncthreads();
            out_phase(a, S, hslot(a, s + 1), bb0, nb, false);
        }
        PROF_MARK(4);
        grid_sync(a.sync, target);
        PROF_MARK(5);
    }
    if (s == a.S)                                                  // loop ran to completion: finalise the last step
        for (int b = cta; b < B; b += PG) finalize_step(a, a.S - 1, b, eos_id);
}

__global__ void dec_persist_init_kernel(DecArgs a, int sos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.B * DE) {
        const int b = i / DE, e = i % DE;
        const float m = a.mask != nullptr ? a.mask[(size_t)b * DE + e] : 1.f;
        const float x = a.emb[(size_t)sos * DE + e] * m;
        a.xbuf[(size_t)b * DX + e] = x;
        if (a.save) a.xtok[(size_t)b * DE + e] = x;
    }
    if (i < a.B && a.save) a.toks[i] = sos;
}


// ================================================================================================= backward
constexpr int NTB = 384;               // threads of the reverse kernel: one float4 column of the 1536 gate gradients each
constexpr int K3 = 3 * DD;             // GRU gate rows (reduction length of the transposed products)
constexpr int RPB = 16;                // W^T rows per CTA (+1 token row on the first 16 CTAs)
static_assert(K3 / 4 == NTB, "one float4 column of the gate gradients per thread");
static_assert(PG == 64 && RPB * (PG / 2) == DD, "row slicing of the reverse kernel");

struct BwdSmem {
    float* WT;      // [RPB+1][K3]  rows of W_ih^T (context / token inputs) on CTAs < 32, of W_hh^T on CTAs >= 32
    float* U;       // [BT][K3]     gate gradients of the staged clips (P2) | scratch of P1 and P3
    float* red;     // [4][12][64] cross-warp reduction scratch + [12][BT] for the token row
    float* Wq;      // [UPC*2][QP]  W_h^T rows of this CTA's hidden units, split in two 128-column halves (padded pitch)
};
constexpr int QP = 132;                // pitch of the half-rows: 16 half-rows hit 8 distinct 16-byte bank groups
constexpr int BWD_RED_FLOATS = 4 * 12 * 64 + 12 * BT;
constexpr int BWD_SMEM_FLOATS = (RPB + 1) * K3 + BT * K3 + BWD_RED_FLOATS + UPC * 2 * 132;

// ---- P1: dh of this CTA's 8 hidden units (all clips), GRU gate gradients -> dgi_all / dgh_all / dh*z
__device__ void bwd_gates_phase(const DecArgs& a, const BwdSmem& S, int s, int bb0, int nb) {
    const int tid = threadIdx.x, cta = blockIdx.x;
    float* dqs = S.U;                                 // [BT][DA]
    const bool last_step = (s == a.S - 1);
    __syncthreads();
    if (!last_step)
        for (int i = tid; i < nb * (DA / 4); i += NTB)
            *reinterpret_cast<float4*>(dqs + (i >> 5) * QP + (i & 31) * 4) = ldcg4(a.dq_all + ((size_t)(s + 1) * a.B + bb0) * DA + (size_t)i * 4);
    __syncthreads();
    if (tid < 2 * UPC * BT) {
        const int b = tid >> 4, u = (tid >> 1) & 7, half = tid & 1;
        float dhq = 0.f;
        if (!last_step && b < nb) {
            const float4* w4 = reinterpret_cast<const float4*>(S.Wq + (u * 2 + half) * QP);
            const float4* d4 = reinterpret_cast<const float4*>(dqs + (b * 2 + half) * QP);
#pragma unroll 8
            for (int i = 0; i < 32; ++i) dhq += dot4(w4[i], d4[i]);
        }
        dhq += __shfl_xor_sync(0xffffffffu, dhq, 1);
        if (half == 0 && b < nb) {
            const int j = cta * UPC + u, bg = bb0 + b;
            const size_t sb = (size_t)s * a.B + bg;
            float dh = __ldg(a.dhc_all + sb * 2 * DD + j);
            if (!last_step) dh += dhq + __ldcg(a.dh_carry + (size_t)bg * DD + j);
            else if (a.dh_last != nullptr) dh += a.dh_last[(size_t)bg * DD + j];
            const float* gs = a.gates + sb * 4 * DD + j;
            const float rr = gs[0], z = gs[DD], n = gs[2 * DD], hnl = gs[3 * DD];
            const float hp = a.hs[sb * DD + j];
            const float dn_pre = dh * (1.f - z) * (1.f - n * n);
            const float dr_pre = dn_pre * hnl * rr * (1.f - rr);
            const float dz_pre = dh * (hp - n) * z * (1.f - z);
            float* gi = a.dgi_all + sb * K3 + j;
            float* gh = a.dgh_all + sb * K3 + j;
            gi[0] = dr_pre; gi[DD] = dz_pre; gi[2 * DD] = dn_pre;
            gh[0] = dr_pre; gh[DD] = dz_pre; gh[2 * DD] = dn_pre * rr;
            a.d_hc[(size_t)bg * 2 * DD + j] = dh * z;                      // direct path of dh_prev
        }
    }
}

// ---- P2: dx = dgi W_ih (CTAs < 32; 16 context columns + 1 token column), dh_prev = dgh W_hh + dh*z (CTAs >= 32)
__device__ void bwd_gemv_phase(const DecArgs& a, const BwdSmem& S, int s, int bb0, int nb) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cta = blockIdx.x;
    const bool is_dx = cta < PG / 2;
    const bool has_tok = cta < DE;
    const float* src = (is_dx ? a.dgi_all : a.dgh_all) + ((size_t)s * a.B + bb0) * K3;
    __syncthreads();
    {
        // NTB == K3/4: thread tid moves float4 column tid of every staged clip; 8 loads in flight per thread
        float4* u4 = reinterpret_cast<float4*>(S.U);
#pragma unroll
        for (int b0 = 0; b0 < BT; b0 += 8) {
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (b0 + j < nb) v[j] = ldcg4(src + ((size_t)(b0 + j) * (K3 / 4) + tid) * 4);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (b0 + j < nb) u4[(b0 + j) * (K3 / 4) + tid] = v[j];
        }
    }
    __syncthreads();
    const float4* x4 = reinterpret_cast<const float4*>(S.U);
    const float4* w4 = reinterpret_cast<const float4*>(S.WT);
    const int base = rs_base<32>(lane);
    float* tokred = S.red + 4 * 12 * 64;                // [12][BT]
#pragma unroll 1
    for (int cg = 0; cg < BT / 4; ++cg) {
        if (cg * 4 >= nb) break;
        float4 xv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) xv[c] = x4[(cg * 4 + c) * (K3 / 4) + tid];
        float* dst = S.red + (cg * 12 + warp) * 64 + base;
#pragma unroll
        for (int h = 0; h < 2; ++h) {                  // 8 rows x 4 clips at a time keeps the register tile small
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
    __syncthreads();
    if (tid < RPB * BT) {
        const int r = tid >> 4, b = tid & 15;
        if (b < nb) {
            const int cg = b >> 2, c = b & 3, bg = bb0 + b;
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 12; ++w) v += S.red[(cg * 12 + w) * 64 + r * 4 + c];
            if (is_dx) {
                a.dx[(size_t)bg * DX + DE + cta * RPB + r] = v;
            } else {
                const int k = (cta - PG / 2) * RPB + r;
                a.dh_carry[(size_t)bg * DD + k] = v + __ldcg(a.d_hc + (size_t)bg * 2 * DD + k);
            }
        }
    } else if (has_tok && tid < RPB * BT + BT) {
        const int b = tid - RPB * BT;
        if (b < nb) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 12; ++w) v += tokred[w * BT + b];
            a.dxtok_all[((size_t)s * a.B + bb0 + b) * DE + cta] = v;
        }
    }
}

// ---- P3: attention backward for clip b, frames [t0,t1): d(score) -> ds_all, dq partials; last arriver sums dq -> dq_all[s]
__device__ void bwd_attn_item(const DecArgs& a, const BwdSmem& S, int s, int b, int js) {
    __shared__ float red12[12];
    __shared__ int is_last;
    float* dc = S.U;                 // [DD]
    float* qv = S.U + DD;            // [DA]
    float* vv = S.U + DD + DA;       // [DA]
    float* accq = S.U + DD + 2 * DA; // [12][DA]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = a.T;
    const int t0 = js * a.tile, t1 = min(T, t0 + a.tile);
    const size_t sb = (size_t)s * a.B + b;
    SUB_BEGIN();
    __syncthreads();
    for (int d = tid; d < DD; d += NTB) {
        const float v = __ldg(a.dhc_all + sb * 2 * DD + DD + d) + __ldcg(a.dx + (size_t)b * DX + DE + d);
        dc[d] = v;
        if (js == 0) a.dctx_all[sb * DD + d] = v;
    }
    if (tid < DA) { qv[tid] = a.qs[sb * DA + tid]; vv[tid] = a.v[tid]; }
    __syncthreads();
    const float* ctx = a.ctxs + sb * DD;
    float c0 = 0.f;
    for (int d = tid; d < DD; d += NTB) c0 += dc[d] * ctx[d];
    c0 = warp_sum(c0);
    if (lane == 0) red12[warp] = c0;
    __syncthreads();
    c0 = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) c0 += red12[i];

    float4 dcr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) dcr[i] = *reinterpret_cast<const float4*>(dc + i * 128 + lane * 4);
    const float4 q0 = *reinterpret_cast<const float4*>(qv + lane * 4);
    const float4 q1 = *reinterpret_cast<const float4*>(qv + 128 + lane * 4);
    const float4 v0 = *reinterpret_cast<const float4*>(vv + lane * 4);
    const float4 v1 = *reinterpret_cast<const float4*>(vv + 128 + lane * 4);
    float dq[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    SUB_MARK(8);
    const float* at = a.attn + sb * T;
    float* dsrow = a.ds_all + sb * T;
    const float vk[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    const float qk[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    struct Rows { float4 en[4], e0, e1; float aw; };
    auto load = [&](Rows& r, int t) {
        const float4* e4 = reinterpret_cast<const float4*>(a.enc + ((size_t)b * T + t) * DD);
#pragma unroll
        for (int i = 0; i < 4; ++i) r.en[i] = lds4(e4 + i * 32 + lane);
        const float4* ep = reinterpret_cast<const float4*>(a.Ep + ((size_t)b * T + t) * DA);
        r.e0 = lds4(ep + lane); r.e1 = lds4(ep + 32 + lane);
        r.aw = at[t];
    };
    Rows ring[PD];                                    // PD frames per warp in flight (register ring)
    constexpr int NWB = NTB / 32;
#pragma unroll
    for (int i = 0; i < PD; ++i)
        if (t0 + warp + i * NWB < t1) load(ring[i], t0 + warp + i * NWB);
    for (int tb = t0 + warp; tb < t1; tb += PD * NWB) {
#pragma unroll
        for (int i = 0; i < PD; ++i) {
            const int t = tb + i * NWB;
            if (t >= t1) break;                       // warp-uniform
            const Rows cur = ring[i];
            if (t + PD * NWB < t1) load(ring[i], t + PD * NWB);
            float da = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) da += dot4(dcr[k], cur.en[k]);
            da = warp_sum(da);
            const float ds = cur.aw * (da - c0);
            if (lane == 0) dsrow[t] = ds;
            const float ev[8] = {cur.e0.x, cur.e0.y, cur.e0.z, cur.e0.w, cur.e1.x, cur.e1.y, cur.e1.z, cur.e1.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float uu = tanh_fast(qk[k] + ev[k]);
                dq[k] = fmaf(ds * vk[k], 1.f - uu * uu, dq[k]);
            }
        }
    }
    SUB_MARK(9);
#pragma unroll
    for (int i = 0; i < 4; ++i) { accq[warp * DA + lane * 4 + i] = dq[i]; accq[warp * DA + 128 + lane * 4 + i] = dq[4 + i]; }
    __syncthreads();
    if (tid < DA) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 12; ++w) t += accq[w * DA + tid];
        a.dq_part[((size_t)b * a.NS + js) * DA + tid] = t;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int tk = atomicAdd(a.tickets + b, 1);
        is_last = (tk == a.NS - 1);
        if (is_last) a.tickets[b] = 0;
    }
    __syncthreads();
    SUB_MARK(10);
    if (!is_last) return;
    __threadfence();
    if (tid < DA) {
        float t = 0.f;
        for (int j = 0; j < a.NS; ++j) t += __ldcg(a.dq_part + ((size_t)b * a.NS + j) * DA + tid);
        a.dq_all[sb * DA + tid] = t;
    }
}

__global__ void __launch_bounds__(NTB, 1) dec_persist_bwd_kernel(DecArgs a) {
    extern __shared__ __align__(16) float smem_f[];
    BwdSmem S;
    S.WT = smem_f;
    S.U = S.WT + (RPB + 1) * K3;
    S.red = S.U + BT * K3;
    S.Wq = S.red + BWD_RED_FLOATS;
    const int tid = threadIdx.x, cta = blockIdx.x;
    unsigned int target = 0;
    // ---- one-time: W^T rows of this CTA (from the host-transposed copies) and its W_h^T rows
    {
        const bool is_dx = cta < PG / 2;
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
        for (int i = tid; i < UPC * DA; i += NTB)
            S.Wq[(i >> 7) * QP + (i & 127)] = a.W_hT[(size_t)cta * UPC * DA + i];      // half-row (u*2+half) = i / 128
    }
    __syncthreads();
    const int B = a.B;
    const int nchunks = (B + BT - 1) / BT;
    const int nitems = B * a.NS;
    unsigned long long prof_t = gtimer();
    for (int s = a.S - 1; s >= 0; --s) {
        for (int ch = 0; ch < nchunks; ++ch) bwd_gates_phase(a, S, s, ch * BT, min(BT, B - ch * BT));
        PROF_MARK(0);
        grid_sync(a.sync, target);
        PROF_MARK(1);
        for (int ch = 0; ch < nchunks; ++ch) bwd_gemv_phase(a, S, s, ch * BT, min(BT, B - ch * BT));
        PROF_MARK(2);
        grid_sync(a.sync, target);
        PROF_MARK(3);
        for (int item = cta; item < nitems; item += PG) bwd_attn_item(a, S, s, item / a.NS, item % a.NS);
        PROF_MARK(4);
        grid_sync(a.sync, target);
        PROF_MARK(5);
    }
    // ---- tail: dh_0 = dh_prev of step 0 + dq_0 W_h  -> a.dhq
    for (int ch = 0; ch < nchunks; ++ch) {
        const int bb0 = ch * BT, nb = min(BT, B - bb0);
        float* dqs = S.U;
        __syncthreads();
        for (int i = tid; i < nb * (DA / 4); i += NTB)
            *reinterpret_cast<float4*>(dqs + (i >> 5) * QP + (i & 31) * 4) = ldcg4(a.dq_all + (size_t)bb0 * DA + (size_t)i * 4);
        __syncthreads();
        if (tid < 2 * UPC * BT) {
            const int b = tid >> 4, u = (tid >> 1) & 7, half = tid & 1;
            float dhq = 0.f;
            if (b < nb) {
                const float4* w4 = reinterpret_cast<const float4*>(S.Wq + (u * 2 + half) * QP);
                const float4* d4 = reinterpret_cast<const float4*>(dqs + (b * 2 + half) * QP);
#pragma unroll 8
                for (int i = 0; i < 32; ++i) dhq += dot4(w4[i], d4[i]);
            }
            dhq += __shfl_xor_sync(0xffffffffu, dhq, 1);
            if (half == 0 && b < nb) {
                const int j = cta * UPC + u;
                a.dhq[(size_t)(bb0 + b) * DD + j] = dhq + __ldcg(a.dh_carry + (size_t)(bb0 + b) * DD + j);
            }
        }
    }
}

// dlogits[s,b,:] = dlogp[b,s,:] - exp(logp[b,s,:]) * sum_v dlogp[b,s,v]   (log_softmax backward), zero padded to VP
__global__ void dec_dlogits_kernel(DecArgs a) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= a.S * a.B) return;
    const int s = row / a.B, b = row % a.B;
    const size_t ro = ((size_t)b * a.max_steps + s) * a.V;
    float sg = 0.f;
    for (int vi = lane; vi < a.V; vi += 32) sg += __ldg(a.dlogp + ro + vi);
    sg = warp_sum(sg);
    for (int vi = lane; vi < a.VP; vi += 32) {
        float d = 0.f;
        if (vi < a.V) d = __ldg(a.dlogp + ro + vi) - expf(__ldg(a.logp + ro + vi)) * sg;
        a.dlogits_all[(size_t)row * a.VP + vi] = d;
    }
}

// Deferred accumulations over the steps (off the sequential chain):
//   dEp[b,t,k] = sum_s ds[s,b,t] v_k (1 - u^2),  dv_k = sum_{s,b,t} ds[s,b,t] u,   u = tanh(q[s,b,k] + Ep[b,t,k])
constexpr int DEF_FPW = 2, DEF_WARPS = 8, DEF_FPB = DEF_FPW * DEF_WARPS;
__global__ void __launch_bounds__(DEF_WARPS * 32) dec_attn_deferred_kernel(DecArgs a) {
    __shared__ float acc[DEF_WARPS][DA];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = a.T;
    const int tA = blockIdx.x * DEF_FPB + warp * DEF_FPW;
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(a.v) + lane), v1 = __ldg(reinterpret_cast<const float4*>(a.v) + 32 + lane);
    const float vk[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    float ev[DEF_FPW][8], dE[DEF_FPW][8], dv[8];
    bool ok[DEF_FPW];
#pragma unroll
    for (int f = 0; f < DEF_FPW; ++f) {
        ok[f] = tA + f < T;
        const float4* ep = reinterpret_cast<const float4*>(a.Ep + ((size_t)b * T + min(tA + f, T - 1)) * DA);
        const float4 e0 = __ldg(ep + lane), e1 = __ldg(ep + 32 + lane);
        ev[f][0] = e0.x; ev[f][1] = e0.y; ev[f][2] = e0.z; ev[f][3] = e0.w; ev[f][4] = e1.x; ev[f][5] = e1.y; ev[f][6] = e1.z; ev[f][7] = e1.w;
#pragma unroll
        for (int i = 0; i < 8; ++i) dE[f][i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) dv[i] = 0.f;
    if (tA < T) {
        for (int s = 0; s < a.S; ++s) {
            const size_t sb = (size_t)s * a.B + b;
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(a.qs + sb * DA) + lane);
            const float4 q1 = __ldg(reinterpret_cast<const float4*>(a.qs + sb * DA) + 32 + lane);
            const float qk[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int f = 0; f < DEF_FPW; ++f) {
                const float ds = ok[f] ? __ldg(a.ds_all + sb * T + tA + f) : 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float uu = tanh_fast(qk[i] + ev[f][i]);
                    dE[f][i] = fmaf(ds * vk[i], 1.f - uu * uu, dE[f][i]);
                    dv[i] = fmaf(ds, uu, dv[i]);
                }
            }
        }
#pragma unroll
        for (int f = 0; f < DEF_FPW; ++f) {
            if (!ok[f]) continue;
            float4* dep = reinterpret_cast<float4*>(a.dEp + ((size_t)b * T + tA + f) * DA);
            dep[lane] = make_float4(dE[f][0], dE[f][1], dE[f][2], dE[f][3]);
            dep[32 + lane] = make_float4(dE[f][4], dE[f][5], dE[f][6], dE[f][7]);
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

}  // namespace

PA2S_API int pa2s_dec_persist_grid(void) { return PG; }

// All S steps of one (bar, staff) in one cooperative launch.  a.NS * a.tile must cover T with NS <= 16;
// a.sync must point to >= 2 zeroed uint32.
PA2S_API int pa2s_note_decoder_fwd_persist(void* stream, const void* args, int sos_id, int eos_id) {
    DecArgs a = *reinterpret_cast<const DecArgs*>(args);
    cudaStream_t st = (cudaStream_t)stream;
    if (a.B <= 0 || a.S <= 0) return 0;
    if (a.V + DA > PG * CR || a.sync == nullptr) return -2;
    const size_t smem = (size_t)FWD_SMEM_FLOATS * sizeof(float);
    PA2S_TRY(cudaFuncSetAttribute(dec_persist_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dec_persist_init_kernel<<<ceil_div(a.B * DE, 128), 128, 0, st>>>(a, sos_id);
    PA2S_CHECK_LAST();
    void* kargs[] = {(void*)&a, (void*)&eos_id};
    PA2S_TRY(cudaLaunchCooperativeKernel((const void*)dec_persist_fwd_kernel, dim3(PG), dim3(NT), kargs, smem, st));
    PA2S_COUNT_LAUNCH();
    return 0;
}

// rows per clip of the dv partials written by pa2s_note_decoder_bwd_persist (dv_part must hold B * this rows of DA floats)
PA2S_API int pa2s_dec_deferred_blocks(int T) { return (T + DEF_FPB - 1) / DEF_FPB; }

// Reverse pass over the S saved steps in one cooperative launch, bracketed by the two parallel kernels that carry the
// work taken off the sequential chain.  Needs (besides the forward's saved state): dlogp, dlogits_all, dhc_all scratch is
// formed by the CALLER between pa2s_dec_dlogits and this call (dhc_all = dlogits_all @ W_out); sync = 2 zeroed uint32.
PA2S_API int pa2s_dec_dlogits(void* stream, const void* args) {
    DecArgs a = *reinterpret_cast<const DecArgs*>(args);
    if (a.B <= 0 || a.S <= 0) return 0;
    dec_dlogits_kernel<<<ceil_div(a.S * a.B, 8), 256, 0, (cudaStream_t)stream>>>(a);
    PA2S_CHECK_LAST();
    return 0;
}
// The two halves of pa2s_note_decoder_bwd_persist as separate entry points, so that a caller can put the parallel dEp / dv
// kernel (which only reads ds_all, qs, Ep, v) on another stream than the sequential chain.
PA2S_API int pa2s_note_decoder_bwd_chain(void* stream, const void* args) {
    DecArgs a = *reinterpret_cast<const DecArgs*>(args);
    cudaStream_t st = (cudaStream_t)stream;
    if (a.B <= 0 || a.S <= 0) return 0;
    if (a.sync == nullptr || a.dhc_all == nullptr || a.ds_all == nullptr) return -2;
    const size_t smem = (size_t)BWD_SMEM_FLOATS * sizeof(float);
    PA2S_TRY(cudaFuncSetAttribute(dec_persist_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* kargs[] = {(void*)&a};
    PA2S_TRY(cudaLaunchCooperativeKernel((const void*)dec_persist_bwd_kernel, dim3(PG), dim3(NTB), kargs, smem, st));
    PA2S_COUNT_LAUNCH();
    return 0;
}
PA2S_API int pa2s_note_decoder_bwd_deferred(void* stream, const void* args) {
    DecArgs a = *reinterpret_cast<const DecArgs*>(args);
    if (a.B <= 0 || a.S <= 0) return 0;
    if (a.ds_all == nullptr || a.dEp == nullptr || a.dv_part == nullptr) return -2;
    dec_attn_deferred_kernel<<<dim3(ceil_div(a.T, DEF_FPB), a.B), DEF_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    PA2S_CHECK_LAST();
    return 0;
}
PA2S_API int pa2s_note_decoder_bwd_persist(void* stream, const void* args) {
    const int rc = pa2s_note_decoder_bwd_chain(stream, args);
    return rc != 0 ? rc : pa2s_note_decoder_bwd_deferred(stream, args);
}
