// GRU recurrences.
//  * encoder BiGRU (models.py:63-67,77): persistent thread-block-cluster kernels.  One cluster of 8 CTAs per
//    (direction, batch group); every CTA keeps its 96x256 slice of W_hh in REGISTERS for the whole sequence, the
//    hidden state lives in shared memory and is exchanged through distributed shared memory with one cluster
//    barrier per time step.  The input projections (x W_ih^T + b_ih) are one GEMM outside.
//  * staff summariser BiGRU(16->32) over packed token sequences (models.py:107-111,164-189): one CTA per
//    (sample, direction), weights in registers, embedding gather / scatter-add fused.
//  * the gate non-linearity of a single GRU cell (bar-level GRU, models.py:117-120,247), forward and backward.
// Gate order is torch's (r, z, n):  r = s(gi_r+gh_r), z = s(gi_z+gh_z), n = tanh(gi_n + r*gh_n), h' = (1-z) n + z h.
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

constexpr int EH = 256;          // encoder hidden size per direction
constexpr int ECL = 8;           // CTAs per cluster
constexpr int EU = EH / ECL;     // units per CTA (32)
constexpr int ENT = 256;         // threads per CTA = EU * 8 k-slices

struct GruSeqArgs {
    const float* gi;      // (B,T,ND*3H)  x W_ih^T + b_ih, both directions
    const float* Whh;     // (ND,3H,H)
    const float* bhh;     // (ND,3H)
    float* out;           // (B,T,ND*H)
    float* gates;         // (B,T,ND,4H): r, z, n, hn_lin (= W_hn h + b_hn)
    float* hN;            // (ND,B,H)
    // backward
    const float* dOut;    // (B,T,ND*H)
    const float* dhN;     // (ND,B,H) or null
    float* dgi;           // (B,T,ND*3H)
    float* dgh;           // (B,T,ND*3H)
    int B, T, ND, G;
};

// 16-byte asynchronous global -> shared copy (LDGSTS).  Unlike a register load, it is not waited for by the release
// fence of the per-step cluster barrier, so a stage issued several steps ahead stays in flight across barriers.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
constexpr int EPF = 4;           // stages of per-step operands in flight (encoder recurrences)

// ---- state exchange of the recurrences without a cluster barrier --------------------------------------------------------------
// Every CTA of a cluster needs the 256 (fwd) / 768 (bwd) new values of all 8 CTAs before its next step.  Plain DSMEM stores +
// barrier.cluster cost ~215 cycles (store) + ~490 cycles (UCGABAR wait) per step; `st.async` delivers the value AND signals the
// destination CTA's mbarrier (complete_tx), so the consumer wakes as soon as the last byte has landed (try_wait: ~60-90 cycles).
// Double-buffered state + one mbarrier per buffer; a CTA can only start step t+1 once every peer has sent step t, and a peer sends
// only after it has finished reading the buffer of step t, so a buffer is never overwritten while it is being read.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 20)) __trap();           // a lost peer becomes a launch failure, not a hang
}
__device__ __forceinline__ void st_async_v4(unsigned raddr, float4 v, unsigned rmbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(raddr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)), "r"(rmbar)
                 : "memory");
}


constexpr int CPT = 4;           // hidden units per warp (8 warps x 4 = the CTA's 32 units); a warp's 32 lanes are 32 k-slices

// Sum NV (power of two <= 32) per-lane values across the warp: lane L returns the warp total of value index L >> (5 - log2 NV).
// Halving exchange (NV/2 + NV/4 + ... + 1 shuffles) followed by a plain butterfly over the remaining lane bits.
template <int NV>
__device__ __forceinline__ float reduce_scatter_nv(float (&v)[NV], int lane) {
    static_assert(NV == 1 || NV == 2 || NV == 4 || NV == 8 || NV == 16 || NV == 32, "NV must be a power of two <= 32");
    int off = 16;
#pragma unroll
    for (int n = NV / 2; n >= 1; n >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float send = up ? v[i] : v[i + n];
            const float recv = __shfl_xor_sync(0xffffffffu, send, off);
            v[i] = (up ? v[i + n] : v[i]) + recv;
        }
        off >>= 1;
    }
    float r = v[0];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        if (o <= off) r += __shfl_xor_sync(0xffffffffu, r, o);
    return r;
}
template <int NV> struct LaneShift { static constexpr int value = NV == 16 ? 1 : (NV == 8 ? 2 : (NV == 4 ? 3 : (NV == 2 ? 4 : 5))); };

// Thread layout of both recurrence kernels: warp w owns units 4w..4w+3 of the CTA's 32, lane l the k-slice {(q*32+l)*4..+3}.
// Each shared-memory read of the state vector is then a conflict-free 512-byte row shared by 12 (fwd) / 4 (bwd) weight rows,
// instead of the same 128 bytes re-read by every quarter warp.  After the reduce-scatter the total for (unit c, sample bb)
// sits in lanes (c*BG+bb) << SH, which do the gate math for that (unit, sample).
template <int BG, bool ASYNC>
__global__ void __launch_bounds__(ENT, 1) gru_seq_fwd_kernel(GruSeqArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / ECL;
    const int dir = cid / a.G, grp = cid % a.G;
    const int b0 = grp * BG;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NV = CPT * BG, SH = LaneShift<NV>::value;
    const int H = EH, T = a.T, ND = a.ND;

    __shared__ __align__(16) float hbuf[2][BG][EH];
    float* remote[ECL];
#pragma unroll
    for (int c = 0; c < ECL; ++c) remote[c] = cluster.map_shared_rank(&hbuf[0][0][0], c);

    // W_hh slice in registers: rows g*H + (rank*32 + warp*4 + c), columns (q*32 + lane)*4 + i
    float w[3][CPT][8];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int c = 0; c < CPT; ++c)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(
                    a.Whh + ((size_t)dir * 3 * H + g * H + rank * EU + warp * CPT + c) * H + (q * 32 + lane) * 4));
                w[g][c][q * 4 + 0] = v.x; w[g][c][q * 4 + 1] = v.y; w[g][c][q * 4 + 2] = v.z; w[g][c][q * 4 + 3] = v.w;
            }
    const int idx = lane >> SH, gb = idx % BG;
    const int u = warp * CPT + idx / BG, j = rank * EU + u, b = b0 + gb;
    const bool gate_thread = ((lane & ((1 << SH) - 1)) == 0) && (b < a.B);
    const float bhr = a.bhh[dir * 3 * H + j], bhz = a.bhh[dir * 3 * H + H + j], bhn = a.bhh[dir * 3 * H + 2 * H + j];
    for (int i = tid; i < 2 * BG * EH; i += ENT) (&hbuf[0][0][0])[i] = 0.f;
    __shared__ __align__(8) unsigned long long xbar[2];
    if (ASYNC && tid == 0) {
        mbar_init(&xbar[0], 1);
        mbar_init(&xbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();
    // async exchange: a warp's 4 units of a clip travel as one 16-byte st.async per destination CTA (lane of unit 0 sends)
    const int nbv = min(BG, a.B - b0);                               // clips of this cluster that exist
    const unsigned xbytes = (unsigned)(ECL * EU * nbv * sizeof(float));
    const bool send_thread = gate_thread && (idx / BG == 0);
    unsigned rbuf[ECL], rbar[ECL];
    if (ASYNC) {
#pragma unroll
        for (int c = 0; c < ECL; ++c) { rbuf[c] = mapa_u32(smem_u32(&hbuf[0][0][0]), c); rbar[c] = mapa_u32(smem_u32(&xbar[0]), c); }
    }

    int p = 0;
    // Input pre-activations (gi: a 118 MB stream) are staged EPF-1 steps ahead with asynchronous copies, so their DRAM latency
    // never sits on the sequential chain: threads 0..BG*3*8-1 each move 16 bytes of the [sample][gate][32 units] slice per step.
    __shared__ __align__(16) float gis[EPF][BG][3][EU];
    auto issue = [&](int step_) {
        if (step_ < T && tid < BG * 3 * 8) {
            const int seg = tid >> 3, q = tid & 7, bb = seg / 3, g = seg % 3;
            if (b0 + bb < a.B) {
                const int t_ = dir == 0 ? step_ : T - 1 - step_;
                cp_async16(&gis[step_ % EPF][bb][g][q * 4],
                           a.gi + ((size_t)(b0 + bb) * T + t_) * ND * 3 * H + dir * 3 * H + g * H + rank * EU + q * 4);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < EPF - 1; ++i) issue(i);
    cp_async_wait<EPF - 2>();
    __syncthreads();
    for (int step = 0; step < T; ++step) {
        const int t = dir == 0 ? step : T - 1 - step;
        issue(step + EPF - 1);
        if (ASYNC) {
            // arm the mbarrier of the buffer this step's results go to, then wait for the previous step's values of all 8 CTAs
            if (tid == 0 && step + 1 < T) mbar_arrive_expect_tx(&xbar[p ^ 1], xbytes);
            if (step > 0) mbar_wait(&xbar[p], (unsigned)(((step - 1) >> 1) & 1));
        }
        float acc[3][NV];
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int i = 0; i < NV; ++i) acc[g][i] = 0.f;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
#pragma unroll
            for (int bb = 0; bb < BG; ++bb) {
                const float4 hv = *reinterpret_cast<const float4*>(&hbuf[p][bb][(q * 32 + lane) * 4]);
#pragma unroll
                for (int g = 0; g < 3; ++g)
#pragma unroll
                    for (int c = 0; c < CPT; ++c) {
                        float v = acc[g][c * BG + bb];
                        v = fmaf(w[g][c][q * 4 + 0], hv.x, v);
                        v = fmaf(w[g][c][q * 4 + 1], hv.y, v);
                        v = fmaf(w[g][c][q * 4 + 2], hv.z, v);
                        v = fmaf(w[g][c][q * 4 + 3], hv.w, v);
                        acc[g][c * BG + bb] = v;
                    }
            }
        }
        const float sr = reduce_scatter_nv<NV>(acc[0], lane);
        const float sz = reduce_scatter_nv<NV>(acc[1], lane);
        const float sn = reduce_scatter_nv<NV>(acc[2], lane);
        float hn = 0.f, sv_r = 0.f, sv_z = 0.f, sv_n = 0.f, sv_hn = 0.f;
        if (gate_thread) {
            const float gir = gis[step % EPF][gb][0][u], giz = gis[step % EPF][gb][1][u], gin = gis[step % EPF][gb][2][u];
            float ghr = sr + bhr, ghz = sz + bhz, ghn = sn + bhn;
            float r = sigmoidf_(gir + ghr), z = sigmoidf_(giz + ghz);
            float n = tanhf(gin + r * ghn);
            float hp = hbuf[p][gb][j];
            hn = (1.f - z) * n + z * hp;
            if (!ASYNC) {
                const int off = ((p ^ 1) * BG + gb) * EH + j;
#pragma unroll
                for (int c = 0; c < ECL; ++c) remote[c][off] = hn;
            }
            sv_r = r; sv_z = z; sv_n = n; sv_hn = ghn;
        }
        if (ASYNC) {
            // gather the warp's 4 units of each clip into the lane of unit 0 (lanes (c*BG+bb) << SH) and send 16 bytes per peer
            const int src0 = (gb << SH);
            float4 v4;
            v4.x = hn;
            v4.y = __shfl_sync(0xffffffffu, hn, ((1 * BG) << SH) + src0);
            v4.z = __shfl_sync(0xffffffffu, hn, ((2 * BG) << SH) + src0);
            v4.w = __shfl_sync(0xffffffffu, hn, ((3 * BG) << SH) + src0);
            if (send_thread && step + 1 < T) {
                const unsigned off = (unsigned)((((p ^ 1) * BG + gb) * EH + rank * EU + warp * CPT) * sizeof(float));
                const unsigned boff = (unsigned)((p ^ 1) * sizeof(unsigned long long));
#pragma unroll
                for (int c = 0; c < ECL; ++c) st_async_v4(rbuf[c] + off, v4, rbar[c] + boff);
            }
        } else {
            // split barrier: the new state is on its way to the 8 CTAs; the global stores of this step (off the sequential chain)
            // are issued while the arrivals propagate
            cluster.barrier_arrive();
        }
        if (gate_thread) {
            a.out[((size_t)b * T + t) * ND * H + dir * H + j] = hn;
            if (a.gates != nullptr) {
                float* gs = a.gates + (((size_t)b * T + t) * ND + dir) * 4 * H + j;
                gs[0] = sv_r; gs[H] = sv_z; gs[2 * H] = sv_n; gs[3 * H] = sv_hn;
            }
            if (step == T - 1) a.hN[((size_t)dir * a.B + b) * H + j] = hn;
        }
        cp_async_wait<EPF - 2>();          // the stage of step+1 has landed (made visible CTA-wide by the barrier)
        if (ASYNC) __syncthreads();
        else cluster.barrier_wait();
        p ^= 1;
    }
    if (ASYNC) cluster.sync();             // no CTA leaves while a peer could still address its shared memory
}

template <int BG, bool ASYNC>
__global__ void __launch_bounds__(ENT, 1) gru_seq_bwd_kernel(GruSeqArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / ECL;
    const int dir = cid / a.G, grp = cid % a.G;
    const int b0 = grp * BG;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NV = CPT * BG, SH = LaneShift<NV>::value;
    const int H = EH, T = a.T, ND = a.ND;

    __shared__ __align__(16) float dbuf[2][BG][3 * EH];
    float* remote[ECL];
#pragma unroll
    for (int c = 0; c < ECL; ++c) remote[c] = cluster.map_shared_rank(&dbuf[0][0][0], c);

    // W_hh^T slice in registers: columns rank*32 + warp*4 + c, rows (q*32 + lane)*4 + i  (q < 6)
    float w[CPT][24];
#pragma unroll
    for (int c = 0; c < CPT; ++c)
#pragma unroll
        for (int q = 0; q < 6; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                w[c][q * 4 + i] = __ldg(a.Whh + ((size_t)dir * 3 * H + (q * 32 + lane) * 4 + i) * H + rank * EU + warp * CPT + c);

    const int idx = lane >> SH, gb = idx % BG;
    const int u = warp * CPT + idx / BG, j = rank * EU + u, b = b0 + gb;
    const bool gate_thread = ((lane & ((1 << SH) - 1)) == 0) && (b < a.B);
    float dhc = 0.f;
    if (gate_thread && a.dhN != nullptr) dhc = a.dhN[((size_t)dir * a.B + b) * H + j];
    for (int i = tid; i < 2 * BG * 3 * EH; i += ENT) (&dbuf[0][0][0])[i] = 0.f;
    __shared__ __align__(8) unsigned long long xbar[2];
    if (ASYNC && tid == 0) {
        mbar_init(&xbar[0], 1);
        mbar_init(&xbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();
    const int nbv = min(BG, a.B - b0);
    const unsigned xbytes = (unsigned)(ECL * EU * nbv * 3 * sizeof(float));
    const bool send_thread = gate_thread && (idx / BG == 0);
    unsigned rbuf[ECL], rbar[ECL];
    if (ASYNC) {
#pragma unroll
        for (int c = 0; c < ECL; ++c) { rbuf[c] = mapa_u32(smem_u32(&dbuf[0][0][0]), c); rbar[c] = mapa_u32(smem_u32(&xbar[0]), c); }
    }

    int p = 0;
    // Saved gates (157 MB per layer), previous state and upstream gradient of each step are staged EPF-1 steps ahead with
    // asynchronous copies: [r, z, n, hn_lin, h_prev, dOut][sample][32 units], 16 bytes per thread of the first BG*6*8.
    __shared__ __align__(16) float stg[EPF][6][BG][EU];
    auto issue = [&](int step_) {          // step_ counts DOWN from T-1; slot = step_ % EPF
        if (step_ >= 0 && tid < BG * 6 * 8) {
            const int seg = tid >> 3, q = tid & 7, bb = seg / 6, c = seg % 6;
            if (b0 + bb < a.B) {
                const int t_ = dir == 0 ? step_ : T - 1 - step_;
                const size_t bt_ = (size_t)(b0 + bb) * T + t_;
                float* dst = &stg[step_ % EPF][c][bb][q * 4];
                const int col = rank * EU + q * 4;
                if (c < 4) cp_async16(dst, a.gates + (bt_ * ND + dir) * 4 * H + c * H + col);
                else if (c == 5) cp_async16(dst, a.dOut + bt_ * ND * H + dir * H + col);
                else if (step_ > 0) cp_async16(dst, a.out + ((size_t)(b0 + bb) * T + (dir == 0 ? t_ - 1 : t_ + 1)) * ND * H + dir * H + col);
                else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < EPF - 1; ++i) issue(T - 1 - i);
    cp_async_wait<EPF - 2>();
    __syncthreads();
    for (int step = T - 1; step >= 0; --step) {
        const int t = dir == 0 ? step : T - 1 - step;
        float dh_direct = 0.f, g_r = 0.f, g_z = 0.f, g_n = 0.f, g_hn = 0.f;
        issue(step - (EPF - 1));
        const int it = T - 1 - step;                                     // iteration count: buffer / mbarrier it & 1, use it >> 1
        if (ASYNC && tid == 0) mbar_arrive_expect_tx(&xbar[p], xbytes);
        if (gate_thread) {
            const int sl = step % EPF;
            const float r = stg[sl][0][gb][u], z = stg[sl][1][gb][u], n = stg[sl][2][gb][u], hnl = stg[sl][3][gb][u];
            const float hp = stg[sl][4][gb][u], dout_t = stg[sl][5][gb][u];
            float dh = dout_t + dhc;
            float dn = dh * (1.f - z), dzv = dh * (hp - n);
            float dn_pre = dn * (1.f - n * n);
            float dr_pre = dn_pre * hnl * r * (1.f - r);
            float dhn_lin = dn_pre * r;
            float dz_pre = dzv * z * (1.f - z);
            dh_direct = dh * z;
            if (!ASYNC) {
                const int off = (p * BG + gb) * 3 * EH + j;
#pragma unroll
                for (int c = 0; c < ECL; ++c) {
                    remote[c][off] = dr_pre; remote[c][off + H] = dz_pre; remote[c][off + 2 * H] = dhn_lin;
                }
            }
            g_r = dr_pre; g_z = dz_pre; g_n = dn_pre; g_hn = dhn_lin;
        }
        if (ASYNC) {
            // the warp's 4 units of a clip as three 16-byte st.async per destination CTA (lane of unit 0 sends)
            const int src0 = (gb << SH);
            float4 vr, vz, vn;
            vr.x = g_r; vz.x = g_z; vn.x = g_hn;
            vr.y = __shfl_sync(0xffffffffu, g_r, ((1 * BG) << SH) + src0); vz.y = __shfl_sync(0xffffffffu, g_z, ((1 * BG) << SH) + src0);
            vn.y = __shfl_sync(0xffffffffu, g_hn, ((1 * BG) << SH) + src0);
            vr.z = __shfl_sync(0xffffffffu, g_r, ((2 * BG) << SH) + src0); vz.z = __shfl_sync(0xffffffffu, g_z, ((2 * BG) << SH) + src0);
            vn.z = __shfl_sync(0xffffffffu, g_hn, ((2 * BG) << SH) + src0);
            vr.w = __shfl_sync(0xffffffffu, g_r, ((3 * BG) << SH) + src0); vz.w = __shfl_sync(0xffffffffu, g_z, ((3 * BG) << SH) + src0);
            vn.w = __shfl_sync(0xffffffffu, g_hn, ((3 * BG) << SH) + src0);
            if (send_thread) {
                const unsigned off = (unsigned)(((p * BG + gb) * 3 * EH + rank * EU + warp * CPT) * sizeof(float));
                const unsigned boff = (unsigned)(p * sizeof(unsigned long long));
#pragma unroll
                for (int c = 0; c < ECL; ++c) {
                    st_async_v4(rbuf[c] + off, vr, rbar[c] + boff);
                    st_async_v4(rbuf[c] + off + H * (unsigned)sizeof(float), vz, rbar[c] + boff);
                    st_async_v4(rbuf[c] + off + 2 * H * (unsigned)sizeof(float), vn, rbar[c] + boff);
                }
            }
        } else {
            cluster.barrier_arrive();          // split barrier: the global stores below overlap the arrival latency
        }
        if (gate_thread) {
            const size_t bt = (size_t)b * T + t;
            float* gi = a.dgi + bt * ND * 3 * H + dir * 3 * H + j;
            gi[0] = g_r; gi[H] = g_z; gi[2 * H] = g_n;
            float* gh = a.dgh + bt * ND * 3 * H + dir * 3 * H + j;
            gh[0] = g_r; gh[H] = g_z; gh[2 * H] = g_hn;
        }
        cp_async_wait<EPF - 2>();          // the stage of step-1 has landed; the barrier makes it visible CTA-wide
        if (ASYNC) mbar_wait(&xbar[p], (unsigned)((it >> 1) & 1));
        else cluster.barrier_wait();
        float acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = 0.f;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
#pragma unroll
            for (int bb = 0; bb < BG; ++bb) {
                const float4 dv = *reinterpret_cast<const float4*>(&dbuf[p][bb][(q * 32 + lane) * 4]);
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    float v = acc[c * BG + bb];
                    v = fmaf(w[c][q * 4 + 0], dv.x, v);
                    v = fmaf(w[c][q * 4 + 1], dv.y, v);
                    v = fmaf(w[c][q * 4 + 2], dv.z, v);
                    v = fmaf(w[c][q * 4 + 3], dv.w, v);
                    acc[c * BG + bb] = v;
                }
            }
        }
        const float tot = reduce_scatter_nv<NV>(acc, lane);
        if (gate_thread) dhc = dh_direct + tot;
        if (ASYNC) __syncthreads();        // the staged operands of the next step (cp.async by other threads) are visible CTA-wide
        p ^= 1;
    }
    if (ASYNC) cluster.sync();             // no CTA leaves while a peer could still address its shared memory
}

// Reverse recurrence, second formulation.  dh_prev[k] = sum over the 768 gate rows j of W_hh[j][k] * dgate[j].  The kernel above gives
// each CTA the columns k of its 32 units and therefore needs ALL 768 gate gradients of a step from the 8 CTAs (12 KB per CTA and
// step through DSMEM, whose ~21 B/cycle makes that 585 cycles) plus a 16-way shuffle reduction.  Here a CTA keeps the ROWS of its
// own 32 units (96 x 256 weights, thread k holds column k in registers): it multiplies its own 96 gate gradients (broadcast reads
// from shared memory, no cross-lane reduction) into partial sums for all 256 k and sends each partial to the CTA that owns unit k --
// one 16-byte st.async per thread and step, 4 KB per CTA -- where the 8 partials are added.
template <int BG>
__global__ void __launch_bounds__(ENT, 1) gru_seq_bwd2_kernel(GruSeqArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / ECL;
    const int dir = cid / a.G, grp = cid % a.G;
    const int b0 = grp * BG;
    const int tid = threadIdx.x;
    const int H = EH, T = a.T, ND = a.ND;
    static_assert(ENT == EH, "one thread per state column");

    __shared__ __align__(16) float gsm[3 * EU][BG];            // gate gradients of this CTA's units: row g*32 + u
    __shared__ __align__(16) float pbuf[2][ECL][EU][BG];       // partial sums received: [buffer][source CTA][unit][clip]
    __shared__ __align__(16) float stg[EPF][6][BG][EU];
    __shared__ __align__(8) unsigned long long xbar[2];

    // rows g*H + rank*32 + u of W_hh, column tid
    float w[3 * EU];
#pragma unroll
    for (int jr = 0; jr < 3 * EU; ++jr)
        w[jr] = __ldg(a.Whh + ((size_t)dir * 3 * H + (jr / EU) * H + rank * EU + (jr % EU)) * H + tid);

    const int u = tid / BG, gb = tid % BG, j = rank * EU + u, b = b0 + gb;
    const bool gate_thread = tid < EU * BG && b < a.B;
    float dhc = 0.f;
    if (gate_thread && a.dhN != nullptr) dhc = a.dhN[((size_t)dir * a.B + b) * H + j];
    for (int i = tid; i < 3 * EU * BG; i += ENT) (&gsm[0][0])[i] = 0.f;
    if (tid == 0) {
        mbar_init(&xbar[0], 1);
        mbar_init(&xbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();
    const unsigned xbytes = (unsigned)(ECL * EU * BG * sizeof(float));
    const unsigned dst_cta = (unsigned)(tid / EU);                      // owner of unit k = tid
    const unsigned rdst = mapa_u32(smem_u32(&pbuf[0][rank][tid % EU][0]), dst_cta);
    const unsigned rbar = mapa_u32(smem_u32(&xbar[0]), dst_cta);

    auto issue = [&](int step_) {          // step_ counts DOWN from T-1; slot = step_ % EPF
        if (step_ >= 0 && tid < BG * 6 * 8) {
            const int seg = tid >> 3, q = tid & 7, bb = seg / 6, c = seg % 6;
            if (b0 + bb < a.B) {
                const int t_ = dir == 0 ? step_ : T - 1 - step_;
                const size_t bt_ = (size_t)(b0 + bb) * T + t_;
                float* dst = &stg[step_ % EPF][c][bb][q * 4];
                const int col = rank * EU + q * 4;
                if (c < 4) cp_async16(dst, a.gates + (bt_ * ND + dir) * 4 * H + c * H + col);
                else if (c == 5) cp_async16(dst, a.dOut + bt_ * ND * H + dir * H + col);
                else if (step_ > 0) cp_async16(dst, a.out + ((size_t)(b0 + bb) * T + (dir == 0 ? t_ - 1 : t_ + 1)) * ND * H + dir * H + col);
                else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < EPF - 1; ++i) issue(T - 1 - i);
    cp_async_wait<EPF - 2>();
    __syncthreads();
    int p = 0;
    for (int step = T - 1; step >= 0; --step) {
        const int t = dir == 0 ? step : T - 1 - step;
        const int it = T - 1 - step;
        issue(step - (EPF - 1));
        if (tid == 0) mbar_arrive_expect_tx(&xbar[p], xbytes);
        float dh_direct = 0.f;
        if (gate_thread) {
            const int sl = step % EPF;
            const float r = stg[sl][0][gb][u], z = stg[sl][1][gb][u], n = stg[sl][2][gb][u], hnl = stg[sl][3][gb][u];
            const float hp = stg[sl][4][gb][u], dout_t = stg[sl][5][gb][u];
            const float dh = dout_t + dhc;
            const float dn = dh * (1.f - z), dzv = dh * (hp - n);
            const float dn_pre = dn * (1.f - n * n);
            const float dr_pre = dn_pre * hnl * r * (1.f - r);
            const float dhn_lin = dn_pre * r;
            const float dz_pre = dzv * z * (1.f - z);
            dh_direct = dh * z;
            gsm[u][gb] = dr_pre; gsm[EU + u][gb] = dz_pre; gsm[2 * EU + u][gb] = dhn_lin;
            const size_t bt = (size_t)b * T + t;
            float* gi = a.dgi + bt * ND * 3 * H + dir * 3 * H + j;
            gi[0] = dr_pre; gi[H] = dz_pre; gi[2 * H] = dn_pre;
            float* gh = a.dgh + bt * ND * 3 * H + dir * 3 * H + j;
            gh[0] = dr_pre; gh[H] = dz_pre; gh[2 * H] = dhn_lin;
        }
        cp_async_wait<EPF - 2>();          // the stage of step-1 has landed; the barrier below makes it (and gsm) visible CTA-wide
        __syncthreads();
        // partial dh_prev[tid] of every clip from this CTA's 96 gate rows
        float acc[BG];
#pragma unroll
        for (int bb = 0; bb < BG; ++bb) acc[bb] = 0.f;
#pragma unroll
        for (int jr = 0; jr < 3 * EU; ++jr) {
#pragma unroll
            for (int bb = 0; bb < BG; ++bb) acc[bb] = fmaf(w[jr], gsm[jr][bb], acc[bb]);
        }
        {
            const unsigned off = (unsigned)(p * ECL * EU * BG * sizeof(float));
            const unsigned boff = (unsigned)(p * sizeof(unsigned long long));
            if (BG == 4) {
                st_async_v4(rdst + off, make_float4(acc[0], acc[BG > 1 ? 1 : 0], acc[BG > 2 ? 2 : 0], acc[BG > 3 ? 3 : 0]), rbar + boff);
            } else {
#pragma unroll
                for (int bb = 0; bb < BG; ++bb)
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                                 ::"r"(rdst + off + (unsigned)(bb * sizeof(float))), "r"(__float_as_uint(acc[bb])), "r"(rbar + boff) : "memory");
            }
        }
        mbar_wait(&xbar[p], (unsigned)((it >> 1) & 1));
        if (gate_thread) {
            float tot = 0.f;
#pragma unroll
            for (int c = 0; c < ECL; ++c) tot += pbuf[p][c][u][gb];
            dhc = dh_direct + tot;
        }
        p ^= 1;
    }
    cluster.sync();                        // no CTA leaves while a peer could still address its shared memory
}

template <typename K>
int launch_cluster(K kernel, cudaStream_t st, int nblocks, GruSeqArgs a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nblocks);
    cfg.blockDim = dim3(ENT);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ECL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    PA2S_TRY(cudaLaunchKernelEx(&cfg, kernel, a));
    PA2S_COUNT_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Staff summariser BiGRU (input 16, hidden 32) over packed sequences
// ---------------------------------------------------------------------------------------------------------
constexpr int SH = 32, SI = 16, SR = 3 * SH;

struct StaffArgs {
    const long long* tokens;   // (B, L)
    const long long* lengths;  // (B)
    const float* emb;          // (V, SI)
    const float* w_ih;         // (2, 96, SI)
    const float* w_hh;         // (2, 96, SH)
    const float* b_ih;         // (2, 96)
    const float* b_hh;         // (2, 96)
    float* hN;                 // (B, 2*SH)
    float* hs;                 // (B, 2, L, SH)   h after each executed step
    float* gates;              // (B, 2, L, 4*SH) r, z, n, hn_lin
    // backward
    const float* dhN;          // (B, 2*SH)
    float* d_emb;              // (V, SI)           atomically accumulated
    float* d_w_ih; float* d_w_hh; float* d_b_ih; float* d_b_hh;
    int B, L;
};

__global__ void __launch_bounds__(SR) staff_gru_fwd_kernel(StaffArgs a) {
    const int b = blockIdx.x, dir = blockIdx.y, row = threadIdx.x;
    __shared__ float xs[SI], hsm[SH], gis[SR], ghs[SR];
    float wi[SI], wh[SH];
#pragma unroll
    for (int i = 0; i < SI; ++i) wi[i] = a.w_ih[((size_t)dir * SR + row) * SI + i];
#pragma unroll
    for (int k = 0; k < SH; ++k) wh[k] = a.w_hh[((size_t)dir * SR + row) * SH + k];
    const float bi = a.b_ih[dir * SR + row], bh = a.b_hh[dir * SR + row];
    long long len = a.lengths[b];
    if (len > a.L) len = a.L;
    if (len < 0) len = 0;
    if (row < SH) hsm[row] = 0.f;
    __syncthreads();
    for (int step = 0; step < (int)len; ++step) {
        const int t = dir == 0 ? step : (int)len - 1 - step;
        if (row < SI) xs[row] = a.emb[a.tokens[(size_t)b * a.L + t] * SI + row];
        __syncthreads();
        float gi = bi, gh = bh;
#pragma unroll
        for (int i = 0; i < SI; ++i) gi = fmaf(wi[i], xs[i], gi);
#pragma unroll
        for (int k = 0; k < SH; ++k) gh = fmaf(wh[k], hsm[k], gh);
        gis[row] = gi; ghs[row] = gh;
        __syncthreads();
        if (row < SH) {
            float r = sigmoidf_(gis[row] + ghs[row]);
            float z = sigmoidf_(gis[SH + row] + ghs[SH + row]);
            float hnl = ghs[2 * SH + row];
            float n = tanhf(gis[2 * SH + row] + r * hnl);
            float hn = (1.f - z) * n + z * hsm[row];
            hsm[row] = hn;
            if (a.hs != nullptr) {
                size_t o = (((size_t)b * 2 + dir) * a.L + t);
                a.hs[o * SH + row] = hn;
                float* gs = a.gates + o * 4 * SH + row;
                gs[0] = r; gs[SH] = z; gs[2 * SH] = n; gs[3 * SH] = hnl;
            }
        }
        __syncthreads();
    }
    if (row < SH) a.hN[(size_t)b * 2 * SH + dir * SH + row] = hsm[row];
}

__global__ void __launch_bounds__(SR) staff_gru_bwd_kernel(StaffArgs a) {
    const int b = blockIdx.x, dir = blockIdx.y, row = threadIdx.x;
    __shared__ float xs[SI], hps[SH], dgi[SR], dgh[SR];
    __shared__ float wis[SR][SI + 1], whs[SR][SH + 1];
#pragma unroll
    for (int i = 0; i < SI; ++i) wis[row][i] = a.w_ih[((size_t)dir * SR + row) * SI + i];
#pragma unroll
    for (int k = 0; k < SH; ++k) whs[row][k] = a.w_hh[((size_t)dir * SR + row) * SH + k];
    float dwi[SI], dwh[SH], dbi = 0.f, dbh = 0.f;
#pragma unroll
    for (int i = 0; i < SI; ++i) dwi[i] = 0.f;
#pragma unroll
    for (int k = 0; k < SH; ++k) dwh[k] = 0.f;
    long long len = a.lengths[b];
    if (len > a.L) len = a.L;
    if (len < 0) len = 0;
    float dh = (row < SH) ? a.dhN[(size_t)b * 2 * SH + dir * SH + row] : 0.f;
    __syncthreads();
    for (int step = (int)len - 1; step >= 0; --step) {
        const int t = dir == 0 ? step : (int)len - 1 - step;
        const long long tok = a.tokens[(size_t)b * a.L + t];
        if (row < SI) xs[row] = a.emb[tok * SI + row];
        float dh_direct = 0.f;
        if (row < SH) {
            size_t o = (((size_t)b * 2 + dir) * a.L + t);
            const float* gs = a.gates + o * 4 * SH + row;
            float r = gs[0], z = gs[SH], n = gs[2 * SH], hnl = gs[3 * SH];
            float hp = 0.f;
            if (step > 0) {
                const int tp = dir == 0 ? t - 1 : t + 1;
                hp = a.hs[(((size_t)b * 2 + dir) * a.L + tp) * SH + row];
            }
            hps[row] = hp;
            float dn = dh * (1.f - z), dzv = dh * (hp - n);
            float dn_pre = dn * (1.f - n * n);
            float dr_pre = dn_pre * hnl * r * (1.f - r);
            float dz_pre = dzv * z * (1.f - z);
            dgi[row] = dr_pre; dgi[SH + row] = dz_pre; dgi[2 * SH + row] = dn_pre;
            dgh[row] = dr_pre; dgh[SH + row] = dz_pre; dgh[2 * SH + row] = dn_pre * r;
            dh_direct = dh * z;
        }
        __syncthreads();
        {
            const float gi = dgi[row], gh = dgh[row];
            dbi += gi; dbh += gh;
#pragma unroll
            for (int i = 0; i < SI; ++i) dwi[i] = fmaf(gi, xs[i], dwi[i]);
#pragma unroll
            for (int k = 0; k < SH; ++k) dwh[k] = fmaf(gh, hps[k], dwh[k]);
        }
        if (row < SH) {
            float acc = dh_direct;
            for (int r2 = 0; r2 < SR; ++r2) acc = fmaf(whs[r2][row], dgh[r2], acc);
            dh = acc;
        } else if (row < SH + SI) {
            const int i = row - SH;
            float acc = 0.f;
            for (int r2 = 0; r2 < SR; ++r2) acc = fmaf(wis[r2][i], dgi[r2], acc);
            atomicAdd(a.d_emb + tok * SI + i, acc);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < SI; ++i) atomicAdd(a.d_w_ih + ((size_t)dir * SR + row) * SI + i, dwi[i]);
#pragma unroll
    for (int k = 0; k < SH; ++k) atomicAdd(a.d_w_hh + ((size_t)dir * SR + row) * SH + k, dwh[k]);
    atomicAdd(a.d_b_ih + dir * SR + row, dbi);
    atomicAdd(a.d_b_hh + dir * SR + row, dbh);
}

// ---------------------------------------------------------------------------------------------------------
// Single GRU cell gate math (bar-level GRU): gi, gh are full pre-activations (biases included).
// ---------------------------------------------------------------------------------------------------------
__global__ void gru_gates_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ hprev,
                                     float* __restrict__ hnew, float* __restrict__ save, int B, int H) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H) return;
    int b = i / H, j = i % H;
    const float* pi = gi + (size_t)b * 3 * H + j;
    const float* ph = gh + (size_t)b * 3 * H + j;
    float r = sigmoidf_(pi[0] + ph[0]), z = sigmoidf_(pi[H] + ph[H]);
    float hnl = ph[2 * H];
    float n = tanhf(pi[2 * H] + r * hnl);
    hnew[i] = (1.f - z) * n + z * hprev[i];
    float* s = save + (size_t)b * 4 * H + j;
    s[0] = r; s[H] = z; s[2 * H] = n; s[3 * H] = hnl;
}

__global__ void gru_gates_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ save, const float* __restrict__ hprev,
                                     float* __restrict__ dgi, float* __restrict__ dgh, float* __restrict__ dhprev, int B, int H) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H) return;
    int b = i / H, j = i % H;
    const float* s = save + (size_t)b * 4 * H + j;
    float r = s[0], z = s[H], n = s[2 * H], hnl = s[3 * H];
    float d = dh[i], hp = hprev[i];
    float dn_pre = d * (1.f - z) * (1.f - n * n);
    float dr_pre = dn_pre * hnl * r * (1.f - r);
    float dz_pre = d * (hp - n) * z * (1.f - z);
    float* pi = dgi + (size_t)b * 3 * H + j;
    float* ph = dgh + (size_t)b * 3 * H + j;
    pi[0] = dr_pre; pi[H] = dz_pre; pi[2 * H] = dn_pre;
    ph[0] = dr_pre; ph[H] = dz_pre; ph[2 * H] = dn_pre * r;
    dhprev[i] = d * z;
}

}  // namespace

PA2S_API int pa2s_gru_seq_max_bg(void) { return 4; }

// 1 (default): the recurrences exchange their state with st.async + mbarrier; 0: DSMEM stores + cluster barrier (round 1)
static int g_gru_async = 1, g_gru_bwd2 = 1;
// 0: DSMEM stores + cluster barrier; 1 (default): st.async + mbarrier, reverse pass in the row-owner formulation (gru_seq_bwd2_kernel);
// 3: st.async + mbarrier with the column-owner reverse kernel
PA2S_API int pa2s_gru_seq_set_exchange(int mode) { g_gru_async = mode ? 1 : 0; g_gru_bwd2 = (mode == 1); return 0; }

// Encoder recurrence forward.  H must be 256.  `bg` in {1,2,4} = samples per cluster.
PA2S_API int pa2s_gru_seq_fwd(void* stream, int B, int T, int ND, int H, int bg, const float* gi, const float* Whh, const float* bhh,
                              float* out, float* gates, float* hN) {
    if (H != EH) return -1;
    if (((uintptr_t)gi & 15) != 0) return -1;          // staged with 16-byte asynchronous copies
    GruSeqArgs a = {};
    a.gi = gi; a.Whh = Whh; a.bhh = bhh; a.out = out; a.gates = gates; a.hN = hN; a.B = B; a.T = T; a.ND = ND;
    a.G = ceil_div(B, bg);
    int nblocks = ND * a.G * ECL;
    cudaStream_t st = (cudaStream_t)stream;
    if (g_gru_async) {
        if (bg == 1) return launch_cluster(gru_seq_fwd_kernel<1, true>, st, nblocks, a);
        if (bg == 2) return launch_cluster(gru_seq_fwd_kernel<2, true>, st, nblocks, a);
        if (bg == 4) return launch_cluster(gru_seq_fwd_kernel<4, true>, st, nblocks, a);
        return -1;
    }
    if (bg == 1) return launch_cluster(gru_seq_fwd_kernel<1, false>, st, nblocks, a);
    if (bg == 2) return launch_cluster(gru_seq_fwd_kernel<2, false>, st, nblocks, a);
    if (bg == 4) return launch_cluster(gru_seq_fwd_kernel<4, false>, st, nblocks, a);
    return -1;
}

PA2S_API int pa2s_gru_seq_bwd(void* stream, int B, int T, int ND, int H, int bg, const float* Whh, const float* out, const float* gates,
                              const float* dOut, const float* dhN, float* dgi, float* dgh) {
    if (H != EH) return -1;
    if ((((uintptr_t)out | (uintptr_t)gates | (uintptr_t)dOut) & 15) != 0) return -1;          // staged with 16-byte asynchronous copies
    GruSeqArgs a = {};
    a.Whh = Whh; a.out = const_cast<float*>(out); a.gates = const_cast<float*>(gates); a.dOut = dOut; a.dhN = dhN; a.dgi = dgi; a.dgh = dgh;
    a.B = B; a.T = T; a.ND = ND;
    a.G = ceil_div(B, bg);
    int nblocks = ND * a.G * ECL;
    cudaStream_t st = (cudaStream_t)stream;
    if (g_gru_async == 2 || (g_gru_async == 1 && g_gru_bwd2)) {
        if (bg == 1) return launch_cluster(gru_seq_bwd2_kernel<1>, st, nblocks, a);
        if (bg == 2) return launch_cluster(gru_seq_bwd2_kernel<2>, st, nblocks, a);
        if (bg == 4) return launch_cluster(gru_seq_bwd2_kernel<4>, st, nblocks, a);
        return -1;
    }
    if (g_gru_async) {
        if (bg == 1) return launch_cluster(gru_seq_bwd_kernel<1, true>, st, nblocks, a);
        if (bg == 2) return launch_cluster(gru_seq_bwd_kernel<2, true>, st, nblocks, a);
        if (bg == 4) return launch_cluster(gru_seq_bwd_kernel<4, true>, st, nblocks, a);
        return -1;
    }
    if (bg == 1) return launch_cluster(gru_seq_bwd_kernel<1, false>, st, nblocks, a);
    if (bg == 2) return launch_cluster(gru_seq_bwd_kernel<2, false>, st, nblocks, a);
    if (bg == 4) return launch_cluster(gru_seq_bwd_kernel<4, false>, st, nblocks, a);
    return -1;
}

PA2S_API int pa2s_staff_gru_fwd(void* stream, int B, int L, int I, int H, const long long* tokens, const long long* lengths,
                                const float* emb, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                                float* hN, float* hs, float* gates) {
    if (I != SI || H != SH) return -1;
    StaffArgs a = {};
    a.tokens = tokens; a.lengths = lengths; a.emb = emb; a.w_ih = w_ih; a.w_hh = w_hh; a.b_ih = b_ih; a.b_hh = b_hh;
    a.hN = hN; a.hs = hs; a.gates = gates; a.B = B; a.L = L;
    staff_gru_fwd_kernel<<<dim3(B, 2), SR, 0, (cudaStream_t)stream>>>(a);
    PA2S_CHECK_LAST();
    return 0;
}

PA2S_API int pa2s_staff_gru_bwd(void* stream, int B, int L, int I, int H, const long long* tokens, const long long* lengths,
                                const float* emb, const float* w_ih, const float* w_hh, const float* hs, const float* gates,
                                const float* dhN, float* d_emb, float* d_w_ih, float* d_w_hh, float* d_b_ih, float* d_b_hh) {
    if (I != SI || H != SH) return -1;
    StaffArgs a = {};
    a.tokens = tokens; a.lengths = lengths; a.emb = emb; a.w_ih = w_ih; a.w_hh = w_hh;
    a.hs = const_cast<float*>(hs); a.gates = const_cast<float*>(gates); a.dhN = dhN;
    a.d_emb = d_emb; a.d_w_ih = d_w_ih; a.d_w_hh = d_w_hh; a.d_b_ih = d_b_ih; a.d_b_hh = d_b_hh; a.B = B; a.L = L;
    staff_gru_bwd_kernel<<<dim3(B, 2), SR, 0, (cudaStream_t)stream>>>(a);
    PA2S_CHECK_LAST();
    return 0;
}

PA2S_API int pa2s_gru_gates_fwd(void* stream, int B, int H, const float* gi, const float* gh, const float* hprev, float* hnew, float* save) {
    gru_gates_fwd_kernel<<<ceil_div(B * H, 256), 256, 0, (cudaStream_t)stream>>>(gi, gh, hprev, hnew, save, B, H);
    PA2S_CHECK_LAST();
    return 0;
}

PA2S_API int pa2s_gru_gates_bwd(void* stream, int B, int H, const float* dh, const float* save, const float* hprev,
                                float* dgi, float* dgh, float* dhprev) {
    gru_gates_bwd_kernel<<<ceil_div(B * H, 256), 256, 0, (cudaStream_t)stream>>>(dh, save, hprev, dgi, dgh, dhprev, B, H);
    PA2S_CHECK_LAST();
    return 0;
}
