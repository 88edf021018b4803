// Note-level attention decoder (NoteDecoder.decode_notes, models.py:366-420; AttentionLayer, models.py:452-461),
// forward and backward, driven step by step from a host loop inside ONE C call per (bar, staff).
//
// Algebra (exactly the reference's, re-associated so the step-invariant half is hoisted):
//   energy_t = v . tanh( W_h h + (W_e enc_t + b) )          W = [W_h | W_e] = attn.weight (256 x 1024)
//   Ep = enc W_e^T + b is computed ONCE per forward per attention module (a GEMM), q = W_h h once per step.
// Per step:  A  attention scores/softmax/context, split over NS CTAs per clip with a last-CTA combine
//            B  GRU cell (528 -> 512): rows of [W_ih | W_hh] sliced across CTAs, all clips per CTA
//            C  logits = W_out [h'; ctx] + b  and  q' = W_h h'   (row-sliced)
//            D  log_softmax, argmax, teacher forcing / next-token embedding (+dropout mask), EOS bookkeeping
// All control flow that the reference does on the host with B device->host syncs per step (models.py:411-419)
// is done on the device: an `eos_count` counter makes every later kernel of the call a no-op once all clips hit EOS.
#include "dec_args.cuh"

namespace {

// --------------------------------------------------------------------------------------------- init
__global__ void dec_init_kernel(DecArgs a, int sos) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.B * DE) {
        int b = i / DE, e = i % DE;
        float m = a.mask != nullptr ? a.mask[(size_t)b * DE + e] : 1.f;
        float x = a.emb[(size_t)sos * DE + e] * m;
        a.xbuf[(size_t)b * DX + e] = x;
        if (a.save) a.xtok[(size_t)b * DE + e] = x;
    }
    if (i < a.B && a.save) a.toks[i] = sos;
}

// --------------------------------------------------------------------------------------------- A: attention
// scores for frames [t0,t1) of clip b, local softmax statistics and partial context; last CTA of the clip combines.
__global__ void __launch_bounds__(256) dec_attn_kernel(DecArgs a, int s) {
    if (all_done(a)) return;
    extern __shared__ __align__(16) float sm[];
    float* qv = sm;                 // DA
    float* vv = sm + DA;            // DA
    float* sc = sm + 2 * DA;        // tile
    __shared__ float red[8];
    __shared__ int is_last;
    const int js = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = a.T;
    const int t0 = js * a.tile, t1 = min(T, t0 + a.tile);
    const float* q = a.qs + ((size_t)hslot(a, s) * a.B + b) * DA;
    qv[tid] = q[tid];
    vv[tid] = a.v[tid];
    __syncthreads();
    // scores: 4 frames per warp iteration, all 8 128-bit loads issued before the tanh work
    {
        const float4 q0 = *reinterpret_cast<const float4*>(qv + lane * 4);
        const float4 q1 = *reinterpret_cast<const float4*>(qv + 128 + lane * 4);
        const float4 v0 = *reinterpret_cast<const float4*>(vv + lane * 4);
        const float4 v1 = *reinterpret_cast<const float4*>(vv + 128 + lane * 4);
        constexpr int FU = 4;
        for (int tb = t0 + warp * FU; tb < t1; tb += 8 * FU) {
            float4 e0[FU], e1[FU];
#pragma unroll
            for (int u = 0; u < FU; ++u) {
                const int t = min(tb + u, t1 - 1);
                const float4* ep = reinterpret_cast<const float4*>(a.Ep + ((size_t)b * T + t) * DA);
                e0[u] = __ldg(ep + lane);
                e1[u] = __ldg(ep + 32 + lane);
            }
#pragma unroll
            for (int u = 0; u < FU; ++u) {
                float e = v0.x * tanh_fast(q0.x + e0[u].x) + v0.y * tanh_fast(q0.y + e0[u].y) + v0.z * tanh_fast(q0.z + e0[u].z) +
                          v0.w * tanh_fast(q0.w + e0[u].w) + v1.x * tanh_fast(q1.x + e1[u].x) + v1.y * tanh_fast(q1.y + e1[u].y) +
                          v1.z * tanh_fast(q1.z + e1[u].z) + v1.w * tanh_fast(q1.w + e1[u].w);
                e = warp_sum(e);
                if (lane == 0 && tb + u < t1) sc[tb + u - t0] = e;
            }
        }
    }
    __syncthreads();
    const int n = t1 - t0;
    float m = -INFINITY;
    for (int i = tid; i < n; i += 256) m = fmaxf(m, sc[i]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    __syncthreads();
    float l = 0.f;
    float* araw = a.attn + ((size_t)slot(a, s) * a.B + b) * T;
    for (int i = tid; i < n; i += 256) {
        float e = sc[i];
        araw[t0 + i] = e;                       // raw score; normalised by the combining CTA
        float p = expf(e - m);
        sc[i] = p;
        l += p;
    }
    l = warp_sum(l);
    if (lane == 0) red[warp] = l;
    __syncthreads();
    l = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) l += red[i];
    // partial context: warp w takes frames w, w+8, ...; lane owns d = 128*j + 4*lane .. +3 (j < 4); 2 frames in flight
    float c0 = 0.f, c1 = 0.f;
    {
        float4 cacc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) cacc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* ebase = reinterpret_cast<const float4*>(a.enc + ((size_t)b * T + t0) * DD);
        for (int i = warp; i < n; i += 16) {
            const int i2 = i + 8;
            const bool two = i2 < n;
            float4 ea[4], eb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) ea[j] = __ldg(ebase + (size_t)i * (DD / 4) + j * 32 + lane);
#pragma unroll
            for (int j = 0; j < 4; ++j) eb[j] = two ? __ldg(ebase + (size_t)i2 * (DD / 4) + j * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float pa = sc[i], pb = two ? sc[i2] : 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                cacc[j].x = fmaf(pa, ea[j].x, fmaf(pb, eb[j].x, cacc[j].x));
                cacc[j].y = fmaf(pa, ea[j].y, fmaf(pb, eb[j].y, cacc[j].y));
                cacc[j].z = fmaf(pa, ea[j].z, fmaf(pb, eb[j].z, cacc[j].z));
                cacc[j].w = fmaf(pa, ea[j].w, fmaf(pb, eb[j].w, cacc[j].w));
            }
        }
        float* part = sm + 2 * DA + a.tile_pad;          // 8 x DD per-warp partial rows (shared memory)
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(part + warp * DD + j * 128 + lane * 4) = cacc[j];
        __syncthreads();
#pragma unroll
        for (int w = 0; w < 8; ++w) { c0 += part[w * DD + 2 * tid]; c1 += part[w * DD + 2 * tid + 1]; }
    }
    if (n <= 0) { m = -INFINITY; l = 0.f; }
    float* pc = a.pc + ((size_t)b * a.NS + js) * DD;
    pc[2 * tid] = c0; pc[2 * tid + 1] = c1;
    if (tid == 0) { a.pm[b * a.NS + js] = m; a.pl[b * a.NS + js] = l; }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        int tk = atomicAdd(a.tickets + b, 1);
        is_last = (tk == a.NS - 1);
        if (is_last) a.tickets[b] = 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // combine
    float M = -INFINITY;
    for (int j = 0; j < a.NS; ++j) M = fmaxf(M, ((volatile float*)a.pm)[b * a.NS + j]);
    float L = 0.f;
    c0 = 0.f; c1 = 0.f;
    for (int j = 0; j < a.NS; ++j) {
        float mj = ((volatile float*)a.pm)[b * a.NS + j];
        float wgt = (mj == -INFINITY) ? 0.f : expf(mj - M);
        L = fmaf(((volatile float*)a.pl)[b * a.NS + j], wgt, L);
        const volatile float* pj = a.pc + ((size_t)b * a.NS + j) * DD;
        c0 = fmaf(pj[2 * tid], wgt, c0);
        c1 = fmaf(pj[2 * tid + 1], wgt, c1);
    }
    const float invL = 1.f / L;
    c0 *= invL; c1 *= invL;
    a.xbuf[(size_t)b * DX + DE + 2 * tid] = c0; a.xbuf[(size_t)b * DX + DE + 2 * tid + 1] = c1;
    a.hc[(size_t)b * 2 * DD + DD + 2 * tid] = c0; a.hc[(size_t)b * 2 * DD + DD + 2 * tid + 1] = c1;
    float* cs = a.ctxs + ((size_t)slot(a, s) * a.B + b) * DD;
    cs[2 * tid] = c0; cs[2 * tid + 1] = c1;
    for (int t = tid; t < T; t += 256) araw[t] = expf(((volatile float*)araw)[t] - M) * invL;
}

// --------------------------------------------------------------------------------------------- B: GRU cell
constexpr int GU = 4;                       // hidden units per CTA
constexpr int GW = 3 * GU;                  // warps per CTA (one gate row each)
__global__ void __launch_bounds__(GW * 32) dec_gru_kernel(DecArgs a, int s) {
    if (all_done(a)) return;
    extern __shared__ __align__(16) float sm[];
    float* xs = sm;                          // BT * DX
    float* hsm = sm + BT * DX;               // BT * DD
    __shared__ float gsum[2][GW][BT];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = warp / GU, jj = warp % GU;
    const int j0 = blockIdx.x * GU;
    const int row = g * DD + j0 + jj;
    const float* hin = a.hs + (size_t)hslot(a, s) * a.B * DD;
    float* hout = a.hs + (size_t)hslot(a, s + 1) * a.B * DD;
    for (int bb0 = 0; bb0 < a.B; bb0 += BT) {
        const int nb = min(BT, a.B - bb0);
        __syncthreads();
        for (int i = tid; i < nb * DX / 4; i += GW * 32)
            reinterpret_cast<float4*>(xs)[i] = reinterpret_cast<const float4*>(a.xbuf + (size_t)bb0 * DX)[i];
        for (int i = tid; i < nb * DD / 4; i += GW * 32)
            reinterpret_cast<float4*>(hsm)[i] = reinterpret_cast<const float4*>(hin + (size_t)bb0 * DD)[i];
        __syncthreads();
        float ai[BT], ah[BT];
#pragma unroll
        for (int b = 0; b < BT; ++b) { ai[b] = 0.f; ah[b] = 0.f; }
        warp_row_dot(a.W_ih + (size_t)row * DX, DX / 4, xs, DX / 4, nb, ai);
        warp_row_dot(a.W_hh + (size_t)row * DD, DD / 4, hsm, DD / 4, nb, ah);
        warp_reduce_all(ai);
        warp_reduce_all(ah);
        if (lane == 0) {
            const float bi = a.b_ih[row], bh = a.b_hh[row];
#pragma unroll
            for (int b = 0; b < BT; ++b) { gsum[0][warp][b] = ai[b] + bi; gsum[1][warp][b] = ah[b] + bh; }
        }
        __syncthreads();
        if (tid < GU * BT) {
            const int u = tid / BT, b = tid % BT;
            if (b < nb) {
                const int j = j0 + u, bg = bb0 + b;
                float r = sigmoidf_(gsum[0][u][b] + gsum[1][u][b]);
                float z = sigmoidf_(gsum[0][GU + u][b] + gsum[1][GU + u][b]);
                float hnl = gsum[1][2 * GU + u][b];
                float n = tanhf(gsum[0][2 * GU + u][b] + r * hnl);
                float hn = (1.f - z) * n + z * hsm[b * DD + j];
                hout[(size_t)bg * DD + j] = hn;
                a.hc[(size_t)bg * 2 * DD + j] = hn;
                if (a.save) {
                    float* gs = a.gates + ((size_t)s * a.B + bg) * 4 * DD + j;
                    gs[0] = r; gs[DD] = z; gs[2 * DD] = n; gs[3 * DD] = hnl;
                }
            }
        }
    }
}

// --------------------------------------------------------------------------------------------- C: logits + next q
// rows [0, V): logits[b,r] = W_out[r,:] . hc[b,:] + b_out[r];  rows [V, V+DA): q'[b,r-V] = W_h[r-V,:] . h'[b,:]
// With only_q != 0 (call prologue) h' is taken from hs[slot 0] and only q rows are produced.
constexpr int PW = 8;                       // warps (rows) per CTA
__global__ void __launch_bounds__(PW * 32) dec_post_kernel(DecArgs a, int s, int only_q) {
    if (all_done(a)) return;
    extern __shared__ __align__(16) float sm[];   // BT * 2DD
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = blockIdx.x * PW + warp;
    const int nrows = only_q ? DA : a.V + DA;
    const int qslot = only_q ? hslot(a, 0) : hslot(a, s + 1);
    const float* hsrc = only_q ? a.hs + (size_t)hslot(a, 0) * a.B * DD : nullptr;
    for (int bb0 = 0; bb0 < a.B; bb0 += BT) {
        const int nb = min(BT, a.B - bb0);
        __syncthreads();
        if (only_q) {
            for (int i = tid; i < nb * DD / 4; i += PW * 32) {
                int b = i / (DD / 4), k = i % (DD / 4);
                reinterpret_cast<float4*>(sm)[b * (2 * DD / 4) + k] = reinterpret_cast<const float4*>(hsrc + (size_t)(bb0 + b) * DD)[k];
            }
        } else {
            for (int i = tid; i < nb * 2 * DD / 4; i += PW * 32)
                reinterpret_cast<float4*>(sm)[i] = reinterpret_cast<const float4*>(a.hc + (size_t)bb0 * 2 * DD)[i];
        }
        __syncthreads();
        if (r < nrows) {
            float acc[BT];
#pragma unroll
            for (int b = 0; b < BT; ++b) acc[b] = 0.f;
            const bool is_logit = !only_q && r < a.V;
            const int qr = only_q ? r : r - a.V;
            if (is_logit) warp_row_dot(a.W_out + (size_t)r * 2 * DD, 2 * DD / 4, sm, 2 * DD / 4, nb, acc);
            else warp_row_dot(a.Wattn + (size_t)qr * 2 * DD, DD / 4, sm, 2 * DD / 4, nb, acc);
            warp_reduce_all(acc);
            if (lane == 0) {
#pragma unroll
                for (int b = 0; b < BT; ++b) {
                    if (b < nb) {
                        if (is_logit) a.logits[(size_t)(bb0 + b) * a.VP + r] = acc[b] + a.b_out[r];
                        else a.qs[((size_t)qslot * a.B + bb0 + b) * DA + qr] = acc[b];
                    }
                }
            }
        }
    }
}

// --------------------------------------------------------------------------------------------- D: finalise step
__global__ void __launch_bounds__(256) dec_fin_kernel(DecArgs a, int s, int eos_id) {
    if (all_done(a)) return;
    __shared__ float redf[8];
    __shared__ int redi[8];
    __shared__ int s_tok;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int V = a.V;
    float x = tid < V ? a.logits[(size_t)b * a.VP + tid] : -INFINITY;
    // max + argmax (first index of the maximum, like torch.argmax)
    float m = x; int idx = tid < V ? tid : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float om = __shfl_xor_sync(0xffffffffu, m, o);
        int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (om > m || (om == m && oi < idx)) { m = om; idx = oi; }
    }
    if (lane == 0) { redf[warp] = m; redi[warp] = idx; }
    __syncthreads();
    m = redf[0]; idx = redi[0];
#pragma unroll
    for (int i = 1; i < 8; ++i)
        if (redf[i] > m || (redf[i] == m && redi[i] < idx)) { m = redf[i]; idx = redi[i]; }
    __syncthreads();
    float e = tid < V ? expf(x - m) : 0.f;
    e = warp_sum(e);
    if (lane == 0) redf[warp] = e;
    __syncthreads();
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += redf[i];
    const float lse = m + logf(sum);
    if (tid < V) a.logp[((size_t)b * a.max_steps + s) * V + tid] = x - lse;
    if (tid == 0) {
        long long g = a.gt != nullptr ? a.gt[(size_t)b * a.max_steps + s] : -1;
        const bool tf = (!a.inference) && a.use_gt != nullptr && a.use_gt[s] != 0 && a.gt != nullptr;
        int tok = tf ? (int)g : idx;
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
        float mk = a.mask != nullptr ? a.mask[((size_t)(s + 1) * a.B + b) * DE + tid] : 1.f;
        float xv = a.emb[(size_t)s_tok * DE + tid] * mk;
        a.xbuf[(size_t)b * DX + tid] = xv;
        if (a.save) a.xtok[((size_t)(s + 1) * a.B + b) * DE + tid] = xv;
    }
}

// ============================================================================================= backward
// X1: d_hc = dlogit W_out (via W_out^T rows), dhq = dq W_h (via W_h^T rows).  tail != 0: only the dq part (after step 0).
__global__ void __launch_bounds__(PW * 32) dec_bwd_out_kernel(DecArgs a, int s, int tail) {
    extern __shared__ __align__(16) float sm[];
    float* dl = sm;                           // BT * VP
    float* dqs = sm + BT * a.VP;              // BT * DA
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = blockIdx.x * PW + warp;     // rows [0,2DD): d_hc; rows [2DD, 3DD): dhq
    const int V = a.V, VP = a.VP;
    const bool have_dq = tail || (s + 1 < a.S);
    for (int bb0 = 0; bb0 < a.B; bb0 += BT) {
        const int nb = min(BT, a.B - bb0);
        __syncthreads();
        if (!tail) {
            for (int b = warp; b < nb; b += PW) {     // dlogit = g - exp(logp) * sum(g)
                const size_t ro = ((size_t)(bb0 + b) * a.max_steps + s) * V;
                float sg = 0.f;
                for (int vi = lane; vi < V; vi += 32) sg += __ldg(a.dlogp + ro + vi);
                sg = warp_sum(sg);
                for (int vi = lane; vi < VP; vi += 32) {
                    float d = 0.f;
                    if (vi < V) d = __ldg(a.dlogp + ro + vi) - expf(__ldg(a.logp + ro + vi)) * sg;
                    dl[b * VP + vi] = d;
                    if (blockIdx.x == 0) a.dlogits_all[((size_t)s * a.B + bb0 + b) * VP + vi] = d;
                }
            }
        }
        for (int i = tid; i < nb * DA; i += PW * 32) {   // dq of the step after this one = sum of its split partials
            int b = i / DA, k = i % DA;
            float d = 0.f;
            if (have_dq)
                for (int j = 0; j < a.NS; ++j) d += a.dq_part[((size_t)(bb0 + b) * a.NS + j) * DA + k];
            dqs[i] = d;
            if (blockIdx.x == 0 && have_dq) a.dq_all[((size_t)(tail ? 0 : s + 1) * a.B + bb0 + b) * DA + k] = d;
        }
        __syncthreads();
        float acc[BT];
#pragma unroll
        for (int b = 0; b < BT; ++b) acc[b] = 0.f;
        if (r < 2 * DD) {
            if (!tail) {
                warp_row_dot(a.W_outT + (size_t)r * VP, VP / 4, dl, VP / 4, nb, acc);
                warp_reduce_all(acc);
                if (lane == 0)
#pragma unroll
                    for (int b = 0; b < BT; ++b)
                        if (b < nb) a.d_hc[(size_t)(bb0 + b) * 2 * DD + r] = acc[b];
            }
        } else if (r < 3 * DD) {
            const int k = r - 2 * DD;
            warp_row_dot(a.W_hT + (size_t)k * DA, DA / 4, dqs, DA / 4, nb, acc);
            warp_reduce_all(acc);
            if (lane == 0)
#pragma unroll
                for (int b = 0; b < BT; ++b)
                    if (b < nb) a.dhq[(size_t)(bb0 + b) * DD + k] = acc[b];
        }
    }
}

// X3: GRU backward.  Rows [0,DX): dx = dgi W_ih;  rows [DX, DX+DD): dh_prev = dgh W_hh + dh*z.
__global__ void __launch_bounds__(PW * 32) dec_bwd_gru_kernel(DecArgs a, int s) {
    extern __shared__ __align__(16) float sm[];       // BT * 3DD
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = blockIdx.x * PW + warp;
    const int first_h_block = (DX + PW - 1) / PW;      // blocks >= this one only hold dh rows
    const bool dx_block = (int)blockIdx.x * PW < DX;   // block contains dx rows -> stage dgi; else dgh
    // NOTE: DX = 528 = 66*8, so no block mixes dx and dh rows.
    const int pin = (s + 1) & 1, pout = s & 1;
    const float* dhc_in = a.dh_carry + (size_t)pin * a.B * DD;
    float* dhc_out = a.dh_carry + (size_t)pout * a.B * DD;
    const bool last_step = (s == a.S - 1);
    for (int bb0 = 0; bb0 < a.B; bb0 += BT) {
        const int nb = min(BT, a.B - bb0);
        __syncthreads();
        for (int i = tid; i < nb * DD; i += PW * 32) {
            const int b = i / DD, j = i % DD, bg = bb0 + b;
            const float* gs = a.gates + ((size_t)s * a.B + bg) * 4 * DD + j;
            const float rr = gs[0], z = gs[DD], n = gs[2 * DD], hnl = gs[3 * DD];
            const float hp = a.hs[((size_t)s * a.B + bg) * DD + j];
            float dh = a.d_hc[(size_t)bg * 2 * DD + j];
            if (!last_step) dh += a.dhq[(size_t)bg * DD + j] + dhc_in[(size_t)bg * DD + j];
            else if (a.dh_last != nullptr) dh += a.dh_last[(size_t)bg * DD + j];
            const float dn_pre = dh * (1.f - z) * (1.f - n * n);
            const float dr_pre = dn_pre * hnl * rr * (1.f - rr);
            const float dz_pre = dh * (hp - n) * z * (1.f - z);
            const float dn_h = dn_pre * rr;
            sm[b * 3 * DD + j] = dr_pre;
            sm[b * 3 * DD + DD + j] = dz_pre;
            sm[b * 3 * DD + 2 * DD + j] = dx_block ? dn_pre : dn_h;
            if (blockIdx.x == 0) {
                float* o = a.dgi_all + ((size_t)s * a.B + bg) * 3 * DD + j;
                o[0] = dr_pre; o[DD] = dz_pre; o[2 * DD] = dn_pre;
            } else if ((int)blockIdx.x == first_h_block) {
                float* o = a.dgh_all + ((size_t)s * a.B + bg) * 3 * DD + j;
                o[0] = dr_pre; o[DD] = dz_pre; o[2 * DD] = dn_h;
            }
        }
        __syncthreads();
        if (r < DX + DD) {
            float acc[BT];
#pragma unroll
            for (int b = 0; b < BT; ++b) acc[b] = 0.f;
            if (r < DX) {
                warp_row_dot(a.W_ihT + (size_t)r * 3 * DD, 3 * DD / 4, sm, 3 * DD / 4, nb, acc);
                warp_reduce_all(acc);
                if (lane == 0)
#pragma unroll
                    for (int b = 0; b < BT; ++b)
                        if (b < nb) {
                            a.dx[(size_t)(bb0 + b) * DX + r] = acc[b];
                            if (r < DE) a.dxtok_all[((size_t)s * a.B + bb0 + b) * DE + r] = acc[b];
                        }
            } else {
                const int k = r - DX;
                warp_row_dot(a.W_hhT + (size_t)k * 3 * DD, 3 * DD / 4, sm, 3 * DD / 4, nb, acc);
                warp_reduce_all(acc);
                if (lane == 0)
#pragma unroll
                    for (int b = 0; b < BT; ++b)
                        if (b < nb) {
                            const int bg = bb0 + b;
                            // + dh * z  (direct path), dh recomputed for this element
                            float dh = a.d_hc[(size_t)bg * 2 * DD + k];
                            if (!last_step) dh += a.dhq[(size_t)bg * DD + k] + dhc_in[(size_t)bg * DD + k];
                            else if (a.dh_last != nullptr) dh += a.dh_last[(size_t)bg * DD + k];
                            const float z = a.gates[((size_t)s * a.B + bg) * 4 * DD + DD + k];
                            dhc_out[(size_t)bg * DD + k] = acc[b] + dh * z;
                        }
            }
        }
    }
}

// X4: attention backward for clip b, frames [t0,t1).
__global__ void __launch_bounds__(256) dec_bwd_attn_kernel(DecArgs a, int s) {
    __shared__ __align__(16) float dc[DD];
    __shared__ __align__(16) float qv[DA];
    __shared__ __align__(16) float vv[DA];
    __shared__ float red[8];
    __shared__ float accq[8][DA];
    const int js = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = a.T;
    const int t0 = js * a.tile, t1 = min(T, t0 + a.tile);
    // d_ctx = (out-projection part) + (GRU input part)
    for (int d = tid; d < DD; d += 256) {
        float v = a.d_hc[(size_t)b * 2 * DD + DD + d] + a.dx[(size_t)b * DX + DE + d];
        dc[d] = v;
        if (js == 0) a.dctx_all[((size_t)s * a.B + b) * DD + d] = v;
    }
    qv[tid] = a.qs[((size_t)s * a.B + b) * DA + tid];
    vv[tid] = a.v[tid];
    __syncthreads();
    // c0 = d_ctx . ctx   ( = sum_t a_t * da_t )
    const float* ctx = a.ctxs + ((size_t)s * a.B + b) * DD;
    float c0 = dc[tid] * ctx[tid] + dc[tid + 256] * ctx[tid + 256];
    c0 = warp_sum(c0);
    if (lane == 0) red[warp] = c0;
    __syncthreads();
    c0 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) c0 += red[i];

    float4 dcr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) dcr[i] = *reinterpret_cast<const float4*>(dc + i * 128 + lane * 4);
    const float4 q0 = *reinterpret_cast<const float4*>(qv + lane * 4);
    const float4 q1 = *reinterpret_cast<const float4*>(qv + 128 + lane * 4);
    const float4 v0 = *reinterpret_cast<const float4*>(vv + lane * 4);
    const float4 v1 = *reinterpret_cast<const float4*>(vv + 128 + lane * 4);
    float dq[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const float* at = a.attn + ((size_t)s * a.B + b) * T;
    const float vk[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    const float qk[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    constexpr int FU = 2;                     // frames in flight per warp: 2 x (4 enc + 2 Ep + 2 dEp) 128-bit loads
    for (int tb = t0 + warp * FU; tb < t1; tb += 8 * FU) {
        float4 en[FU][4], e0[FU], e1[FU], d0[FU], d1[FU];
        float aw[FU];
#pragma unroll
        for (int u = 0; u < FU; ++u) {
            const int t = min(tb + u, t1 - 1);
            const float4* e4 = reinterpret_cast<const float4*>(a.enc + ((size_t)b * T + t) * DD);
#pragma unroll
            for (int i = 0; i < 4; ++i) en[u][i] = __ldg(e4 + i * 32 + lane);
            const float4* ep = reinterpret_cast<const float4*>(a.Ep + ((size_t)b * T + t) * DA);
            e0[u] = __ldg(ep + lane); e1[u] = __ldg(ep + 32 + lane);
            const float4* dep = reinterpret_cast<const float4*>(a.dEp + ((size_t)b * T + t) * DA);
            d0[u] = dep[lane]; d1[u] = dep[32 + lane];
            aw[u] = at[t];
        }
#pragma unroll
        for (int u = 0; u < FU; ++u) {
            const int t = tb + u;
            float da = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) da += dot4(dcr[i], en[u][i]);
            da = warp_sum(da);
            if (t >= t1) continue;            // warp-uniform
            const float ds = aw[u] * (da - c0);
            const float ev[8] = {e0[u].x, e0[u].y, e0[u].z, e0[u].w, e1[u].x, e1[u].y, e1[u].z, e1[u].w};
            float dp[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float uu = tanh_fast(qk[i] + ev[i]);
                dp[i] = ds * vk[i] * (1.f - uu * uu);
                dq[i] += dp[i];
                dv[i] = fmaf(ds, uu, dv[i]);
            }
            float4* dep = reinterpret_cast<float4*>(a.dEp + ((size_t)b * T + t) * DA);
            dep[lane] = make_float4(d0[u].x + dp[0], d0[u].y + dp[1], d0[u].z + dp[2], d0[u].w + dp[3]);
            dep[32 + lane] = make_float4(d1[u].x + dp[4], d1[u].y + dp[5], d1[u].z + dp[6], d1[u].w + dp[7]);
        }
    }
    // reduce dq, dv over the 8 warps
#pragma unroll
    for (int i = 0; i < 4; ++i) { accq[warp][lane * 4 + i] = dq[i]; accq[warp][128 + lane * 4 + i] = dq[4 + i]; }
    __syncthreads();
    {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += accq[w][tid];
        a.dq_part[((size_t)b * a.NS + js) * DA + tid] = t;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) { accq[warp][lane * 4 + i] = dv[i]; accq[warp][128 + lane * 4 + i] = dv[4 + i]; }
    __syncthreads();
    {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += accq[w][tid];
        a.dv_part[((size_t)b * a.NS + js) * DA + tid] += t;
    }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace

// The argument block is a plain C struct mirrored field-for-field by ctypes (see piano_a2s_b200/_lib.py).
PA2S_API int pa2s_dec_args_size(void) { return (int)sizeof(DecArgs); }

// Runs S decoding steps.  Returns 0 or a CUDA error code.
PA2S_API int pa2s_note_decoder_fwd(void* stream, const void* args, int sos_id, int eos_id) {
    DecArgs a = *reinterpret_cast<const DecArgs*>(args);
    cudaStream_t st = (cudaStream_t)stream;
    if (a.B <= 0 || a.S <= 0) return 0;
    a.tile_pad = (a.tile + 3) / 4 * 4;
    const size_t sm_attn = (size_t)(2 * DA + a.tile_pad + 8 * DD) * sizeof(float);
    const size_t sm_gru = (size_t)BT * (DX + DD) * sizeof(float);
    const size_t sm_post = (size_t)BT * 2 * DD * sizeof(float);
    PA2S_TRY((cudaError_t)set_smem(dec_attn_kernel, sm_attn));
    PA2S_TRY((cudaError_t)set_smem(dec_gru_kernel, sm_gru));
    PA2S_TRY((cudaError_t)set_smem(dec_post_kernel, sm_post));
    dec_init_kernel<<<ceil_div(a.B * DE, 128), 128, 0, st>>>(a, sos_id);
    PA2S_CHECK_LAST();
    dec_post_kernel<<<ceil_div(DA, PW), PW * 32, sm_post, st>>>(a, 0, 1);
    PA2S_CHECK_LAST();
    for (int s = 0; s < a.S; ++s) {
        dec_attn_kernel<<<dim3(a.NS, a.B), 256, sm_attn, st>>>(a, s);
        PA2S_CHECK_LAST();
        dec_gru_kernel<<<DD / GU, GW * 32, sm_gru, st>>>(a, s);
        PA2S_CHECK_LAST();
        dec_post_kernel<<<ceil_div(a.V + DA, PW), PW * 32, sm_post, st>>>(a, s, 0);
        PA2S_CHECK_LAST();
        dec_fin_kernel<<<a.B, 256, 0, st>>>(a, s, eos_id);
        PA2S_CHECK_LAST();
    }
    return 0;
}

// Reverse pass over the S saved steps; weight gradients are formed afterwards by GEMMs over the *_all buffers.
PA2S_API int pa2s_note_decoder_bwd(void* stream, const void* args) {
    DecArgs a = *reinterpret_cast<const DecArgs*>(args);
    cudaStream_t st = (cudaStream_t)stream;
    if (a.B <= 0 || a.S <= 0) return 0;
    const size_t sm_out = (size_t)BT * (a.VP + DA) * sizeof(float);
    const size_t sm_gru = (size_t)BT * 3 * DD * sizeof(float);
    PA2S_TRY((cudaError_t)set_smem(dec_bwd_out_kernel, sm_out));
    PA2S_TRY((cudaError_t)set_smem(dec_bwd_gru_kernel, sm_gru));
    for (int s = a.S - 1; s >= 0; --s) {
        dec_bwd_out_kernel<<<ceil_div(3 * DD, PW), PW * 32, sm_out, st>>>(a, s, 0);
        PA2S_CHECK_LAST();
        dec_bwd_gru_kernel<<<ceil_div(DX + DD, PW), PW * 32, sm_gru, st>>>(a, s);
        PA2S_CHECK_LAST();
        dec_bwd_attn_kernel<<<dim3(a.NS, a.B), 256, 0, st>>>(a, s);
        PA2S_CHECK_LAST();
    }
    dec_bwd_out_kernel<<<ceil_div(3 * DD, PW), PW * 32, sm_out, st>>>(a, 0, 1);
    PA2S_CHECK_LAST();
    return 0;
}

// Stand-alone single attention step (bar-level attention, models.py:241-242): q must already be in a.qs slot 0.
PA2S_API int pa2s_attn_step_fwd(void* stream, const void* args) {
    DecArgs a = *reinterpret_cast<const DecArgs*>(args);
    a.tile_pad = (a.tile + 3) / 4 * 4;
    const size_t sm_attn = (size_t)(2 * DA + a.tile_pad + 8 * DD) * sizeof(float);
    PA2S_TRY((cudaError_t)set_smem(dec_attn_kernel, sm_attn));
    dec_attn_kernel<<<dim3(a.NS, a.B), 256, sm_attn, (cudaStream_t)stream>>>(a, 0);
    PA2S_CHECK_LAST();
    return 0;
}
// Backward of the single step: expects d_ctx in a.d_hc[:, DD:] (+ a.dx[:, DE:]), writes dq_part / dv_part / dEp.
PA2S_API int pa2s_attn_step_bwd(void* stream, const void* args) {
    DecArgs a = *reinterpret_cast<const DecArgs*>(args);
    dec_bwd_attn_kernel<<<dim3(a.NS, a.B), 256, 0, (cudaStream_t)stream>>>(a, 0);
    PA2S_CHECK_LAST();
    return 0;
}
