// Argument block and small device helpers shared by the note-decoder kernels (decoder.cu: per-step kernels used for the
// single bar-level attention step; dec_persist.cu: persistent whole-sequence kernels).
#pragma once
#include "common.cuh"

namespace {

constexpr int DD = 512;    // decoder hidden (2*hidden_size)
constexpr int DA = 256;    // attention width (hidden_size)
constexpr int DE = 16;     // note embedding
constexpr int DX = DE + DD;
constexpr int BT = 16;     // batch chunk held in registers by the row-dot kernels

struct DecArgs {
    // problem
    int B, T, V, VP, S, max_steps, NS, tile, inference, save;
    int tile_pad, reserved_;   // host-filled: tile rounded up to 4 floats
    // encoder memory and attention module
    const float* enc;      // (B,T,DD)
    const float* Ep;       // (B,T,DA)
    const float* Wattn;    // (DA, 2*DD) row-major; W_h = [:, :DD]
    const float* v;        // (DA)
    // decoder weights
    const float* emb;      // (V, DE)
    const float* W_ih; const float* W_hh; const float* b_ih; const float* b_hh;   // (3DD,DX) (3DD,DD)
    const float* W_out; const float* b_out;                                         // (V, 2DD)
    // transposed copies for backward
    const float* W_outT;   // (2DD, VP) zero padded
    const float* W_hT;     // (DD, DA)
    const float* W_ihT;    // (DX, 3DD)
    const float* W_hhT;    // (DD, 3DD)
    // teacher forcing / dropout
    const long long* gt;   // (B, max_steps) or null
    const int* use_gt;     // (S) or null
    const float* mask;     // (S, B, DE) or null
    // outputs
    float* logp;           // (B, max_steps, V), pre-zeroed
    long long* lengths;    // (B) pre-set to max_steps
    int* eos;              // (B) zero
    int* counters;         // [0] eos_count, [1] steps executed
    // saved state (S-indexed when save, else slot 0 / ping-pong)
    float* hs;             // (S+1,B,DD)
    float* ctxs;           // (S,B,DD)
    float* attn;           // (S,B,T)
    float* gates;          // (S,B,4DD)
    float* qs;             // (S+1,B,DA)
    float* xtok;           // (S+1,B,DE)  dropped-out input embedding of each step
    int* toks;             // (S+1,B)
    // scratch
    float* xbuf;           // (B,DX)   [tok | ctx]
    float* hc;             // (B,2DD)  [h' | ctx]
    float* logits;         // (B,VP)
    float* pm; float* pl; float* pc;   // (B,NS) (B,NS) (B,NS,DD)
    int* tickets;          // (B) zero
    // backward
    const float* dlogp;    // (B,max_steps,V)
    float* dlogits_all;    // (S,B,VP)
    float* dgi_all;        // (S,B,3DD)
    float* dgh_all;        // (S,B,3DD)
    float* dq_all;         // (S+1,B,DA)
    float* dctx_all;       // (S,B,DD)
    float* dxtok_all;      // (S,B,DE)
    float* dEp;            // (B,T,DA) accumulated
    float* dv_part;        // (B*NS, DA) accumulated
    float* d_hc;           // (B,2DD)
    float* dhq;            // (B,DD)
    float* dx;             // (B,DX)
    float* dq_part;        // (B,NS,DA)
    float* dh_carry;       // (2,B,DD)
    const float* dh_last;  // (B,DD) or null: upstream gradient wrt the final hidden state (unused by the reference)
    // persistent kernels (dec_persist.cu)
    unsigned int* sync;    // [0] grid-barrier arrival counter, [1] error flag (barrier watchdog); zeroed by the host
    const float* dhc_all;  // (S,B,2DD) = dlogits_all @ W_out, formed by one GEMM before the reverse loop
    float* ds_all;         // (S,B,T) d(loss)/d(score), consumed by the deferred dEp / dv kernel
    float* dv;             // (DA) accumulated by the deferred kernel
    unsigned long long* prof;  // optional [8] ns accumulated per phase by CTA 0 (tools/prof_decoder.py), or null
};

__device__ __forceinline__ int slot(const DecArgs& a, int s) { return a.save ? s : 0; }
__device__ __forceinline__ int hslot(const DecArgs& a, int s) { return a.save ? s : (s & 1); }
__device__ __forceinline__ bool all_done(const DecArgs& a) { return *((volatile int*)a.counters) >= a.B; }

// tanh(x) = 1 - 2/(exp(2x)+1) on the SFU (ex2.approx + rcp.approx): absolute error ~1e-7, saturates correctly at +-1.
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = __expf(2.f * x);
    return 1.f - __fdividef(2.f, e + 1.f);
}

__device__ __forceinline__ float dot4(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }

// acc[b] += w[0..4*K4) . xs[b][0..4*K4) for b < nb; lanes stride over float4 columns.
__device__ __forceinline__ void warp_row_dot(const float* __restrict__ wrow, int K4, const float* xs, int pitch4, int nb,
                                             float (&acc)[BT]) {
    const int lane = threadIdx.x & 31;
    const float4* w4 = reinterpret_cast<const float4*>(wrow);
    const float4* x4 = reinterpret_cast<const float4*>(xs);
    constexpr int WB = 6;                       // weight float4s in flight per lane (covers K <= 768 per pass)
    for (int k0 = lane; k0 < K4; k0 += 32 * WB) {
        float4 w[WB];
#pragma unroll
        for (int j = 0; j < WB; ++j) {
            const int k = k0 + 32 * j;
            w[j] = k < K4 ? __ldg(w4 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < WB; ++j) {
            const int k = k0 + 32 * j;
            if (k < K4) {
#pragma unroll
                for (int b = 0; b < BT; ++b)
                    if (b < nb) acc[b] += dot4(w[j], x4[b * pitch4 + k]);
            }
        }
    }
}
__device__ __forceinline__ void warp_reduce_all(float (&acc)[BT]) {
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[b] = warp_sum(acc[b]);
}


}  // namespace
