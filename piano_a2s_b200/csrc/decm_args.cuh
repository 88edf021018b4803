// Argument block of the multi-sequence persistent note decoder (dec_multi.cu).  Mirrored by ctypes in piano_a2s_b200/_lib.py
// (DecMArgs; pa2s_decm_args_size() lets the binding verify the layout).
//
// One launch decodes NQ sequences ("queries": bars of one staff) for each of B clips: R = NQ*B rows, row r = q*B + b.  All
// rows of a clip attend over the SAME encoder memory, so one pass over a clip's frames serves NQ queries.  Step-major
// saved-state buffers are shared by all launches of a forward pass (row pitch Rtot, this launch's first row r0), so that
// the reverse pass can run over ALL rows (all bars of the staff) in one launch.
#pragma once
#include "dec_args.cuh"

namespace {

constexpr int NQMAX = 5;

struct DecMArgs {
    int B, NQ, T, V, VP, S, max_steps, NS, tile, inference, save, Rtot, r0, bars, k0, Spitch;
    int tc;                  // 1: weight-stationary products on the tensor cores (bf16 hi/lo split, fp32 accumulate); 0: exact fp32 FFMA
    int Sq[8];               // executed steps of query q (1 <= Sq[q] <= S = max_q Sq[q])
    unsigned int tf_bits[64];// note-level teacher-forcing coins (models.py:404) of this launch, bit q * Spitch + s (NQ * Spitch <= 2048)
    int has_tf, pad_;        // 1: tf_bits is valid (training with targets); 0: every next token is the arg-max
    // encoder memory and attention module
    const float* enc;        // (B,T,DD)
    const float* Ee;         // (B,T,DA)  exp(2 * (enc W_e^T + b)): tanh(q + Ep) = 1 - 2 / (1 + exp(2q) * Ee)
    const float* Wattn;      // (DA, 2*DD) row-major; W_h = [:, :DD]
    const float* v;          // (DA)
    // decoder weights
    const float* emb;
    const float* W_ih; const float* W_hh; const float* b_ih; const float* b_hh;
    const float* W_out; const float* b_out;
    const float* W_hT;       // (DD, DA)
    const float* W_ihT;      // (DX, 3DD)
    const float* W_hhT;      // (DD, 3DD)
    // teacher forcing / dropout
    const long long* gt;     // (B, bars, max_steps) or null; row (q,b) reads bar k0+q
    const float* mask;       // (S, Rtot, DE) or null: dropout mask of the input embedding of (step, global row), like xtok
    // outputs
    float* logp;             // (B, bars, max_steps, V), pre-zeroed; row (q,b) writes bar k0+q
    long long* lengths;      // (Rtot) pre-set to max_steps, indexed by global row
    int* eos;                // (R) zero
    int* counters;           // [0] rows that have emitted <eos>, [1] steps executed
    // saved state, (step, global row) major when save, else slot 0 / ping-pong with pitch Rtot = R
    float* hs;               // (S+1,Rtot,DD)
    float* ctxs;             // (S,Rtot,DD)
    float* attn;             // (S,Rtot,T)   raw scores after the forward; normalised in place by the reverse pass
    float* gates;            // (S,Rtot,4DD)
    float* qs;               // (S+1,Rtot,DA)
    float* eqs;              // (S,Rtot,DA)  exp(2q)
    float* xtok;             // (S+1,Rtot,DE)
    int* toks;               // (S+1,Rtot)
    float* ml;               // (S,Rtot,2)   softmax maximum and 1/sum of every (step, row)
    // scratch (local rows)
    float* xbuf;             // (R,DX)
    float* logits;           // (R,VP)
    float* pm; float* pl; float* pc;   // (R,NS) (R,NS) (R,NS,DD)
    int* tickets;            // (B) zero
    unsigned int* sync;      // [0] grid-barrier arrival counter, [1] watchdog flag; zeroed by the host
    // backward (one launch over all rows: r0 = 0, R = Rtot)
    const float* dhc_all;    // (S,Rtot,2DD) = dlogits_all @ W_out
    float* dgi_all;          // (S,Rtot,3DD) zero-initialised (inactive rows stay zero)
    float* dgh_all;          // (S,Rtot,3DD) "
    float* dq_all;           // (S+1,Rtot,DA) "
    float* dctx_all;         // (S,Rtot,DD) "
    float* dxtok_all;        // (S,Rtot,DE) "
    float* ds_all;           // (S,Rtot,T) "
    float* dEp;              // (B,T,DA)
    float* dv_part;          // (B*nblk, DA)
    float* d_hc;             // (R,2DD)
    float* dhq;              // (R,DD)   gradient wrt h_0 of every row
    float* dx;               // (R,DX)
    float* dq_part;          // (R,NS,DA)
    float* dh_carry;         // (R,DD)
    const float* dlogp;      // (B,bars,max_steps,V)
    float* dlogits_all;      // (S,Rtot,VP)
    unsigned long long* prof;
};

}  // namespace
