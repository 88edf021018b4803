/* pa2s.h -- C ABI of libpa2s.so, the sm_100a kernel library behind piano-a2s's data-parallel hot path.
 *
 * The reference (wei-zeng98/piano-a2s @ ca8bc59) has NO native/FFI interface: its hot path is the Python class
 * `models.ScoreTranscription` (models.py:14-51) executed through stock PyTorch library calls, and its front end
 * is `utilities.get_VQT` (utilities.py:240-254) executed through librosa on the CPU.  The entry points below are
 * what a binding for that path attaches to; each cites the reference lines whose computation it replaces.  The
 * reference-side binding (a ctypes stub inside models.py) is shown in INTEGRATION.md.
 *
 * Conventions: every pointer is a DEVICE pointer to densely packed fp32 data unless stated otherwise; `stream` is a
 * cudaStream_t; every function only enqueues work on `stream` and returns 0 or a cudaError_t value (or -1 for an
 * unsupported shape).  No torch types cross this boundary.  Activations are channels-last: (B, T, F, C).
 */
#ifndef PA2S_H
#define PA2S_H
#ifdef __cplusplus
#define PA2S_API extern "C"
#else
#define PA2S_API
#endif

/* Number of kernels this library has launched in this process (bench.py `gpu_launches`). */
PA2S_API unsigned long long pa2s_launch_count(void);

/* ---- dense contractions ------------------------------------------------------------------------------------
 * C[b] = op(A[b]) op(B[b]) (+bias) (+C[b]);  op(A)(m,k) = transA ? A[k*lda+m] : A[m*lda+k];
 * op(B)(k,n) = transB ? B[n*ldb+k] : B[k*ldb+n].  `atomic`/`splitk>1` accumulate into C with atomicAdd (C must hold
 * the value to add to).  If t_scale != NULL one operand (B if t_on_b else A; it must be the non-transposed-A /
 * non-transposed-B form whose contiguous index is k resp. n) is read as relu?(x*t_scale[i%t_period]+t_shift[i%t_period]).
 * Replaces: F.linear of ConvStack.out (models.py:504,539), the nn.GRU input projections (models.py:63-67,117,353),
 * the encoder half of AttentionLayer.attn (models.py:444,458), the MLP heads (models.py:123-132), torch.bmm context
 * gradients, and every weight-gradient contraction autograd derives from them.  With lda = hop it is also the VQT
 * filterbank contraction over overlapping audio frames (utilities.py:246). */
PA2S_API int pa2s_gemm_f32(void* stream, int transA, int transB, int M, int N, int K,
                           const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc,
                           const float* bias, int accumulate, int atomic,
                           int batch, long long strideA, long long strideB, long long strideC,
                           const float* t_scale, const float* t_shift, int t_period, int t_relu, int t_on_b,
                           int splitk);

/* Same contract on the tcgen05 tensor cores (tc_gemm.cu): fp32 operands are split on the fly into bf16 hi/lo and
 * accumulated in TMEM as A_hi*B_hi + A_hi*B_lo + A_lo*B_hi (nsplit = 3, ~fp32 accuracy) or A_hi*B_hi (nsplit = 1). */
PA2S_API int pa2s_gemm_tc_supported(int M, int N, int K, int batch);
PA2S_API int pa2s_gemm_tc(void* stream, int transA, int transB, int M, int N, int K,
                          const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc,
                          const float* bias, int accumulate, int atomic,
                          int batch, long long strideA, long long strideB, long long strideC,
                          const float* t_scale, const float* t_shift, int t_period, int t_relu, int t_on_b,
                          int splitk, int nsplit);

/* TMA-fed tcgen05 GEMM on pre-split bf16 operands (tc_gemm_tma.cu).  pa2s_split_bf16 writes an fp32 matrix (optionally through
 * relu?(x*t_scale[c%t_period]+t_shift[c%t_period]), c = column) as 1..3 bf16 pieces x ~ p0+p1(+p2): piece p, batch b, row r at
 * dst + p*piece_stride + b*batch_stride_dst + r*ld_dst (elements; all multiples of 8).  pa2s_gemm_bf16_tma contracts
 * C[b] (+)= sum_{i+j<=max(nA,nB)-1} A_i[b] B_j[b]^T with fp32 TMEM accumulation; an operand is stored [mn][k] (x_mn = 0) or
 * [k][mn] (x_mn = 1) with row pitch x_ld, which may be smaller than the row length (overlapping frames: the VQT filterbank
 * contraction, utilities.py:246).  x_batch_stride = 0 shares the operand between batches.  Replaces the same reference
 * lines as pa2s_gemm_f32. */
PA2S_API int pa2s_split_bf16(void* stream, const float* src, long long rows, long long cols, long long ld_src, long long batch_stride_src,
                             void* dst, long long ld_dst, long long piece_stride, long long batch_stride_dst, int npieces, int batch,
                             const float* t_scale, const float* t_shift, int t_period, int t_relu);
PA2S_API int pa2s_gemm_bf16_tma(void* stream, int M, int N, int K,
                                const void* A, long long a_ld, long long a_piece_stride, long long a_batch_stride, int a_pieces, int a_mn,
                                const void* B, long long b_ld, long long b_piece_stride, long long b_batch_stride, int b_pieces, int b_mn,
                                float* C, long long ldc, long long strideC, const float* bias, int atomic, int batch, int splitk);

/* the same contraction with the VQT epilogue fused: N columns = (re, im) pairs; writes mag (batch, M, N/2) = |.| and folds each batch's
 * maximum into clip_max[batch] (uint32 view of a non-negative float, zeroed by the caller); rows >= valid_rows[batch] (or all rows when
 * NULL) ... are zeros and stay out of the maximum.  Replaces librosa.vqt + np.abs + the `ref=np.max` reduction of utilities.py:246-253. */
PA2S_API int pa2s_gemm_bf16_tma_mag(void* stream, int M, int N, int K,
                                    const void* A, long long a_ld, long long a_piece_stride, long long a_batch_stride, int a_pieces, int a_mn,
                                    const void* B, long long b_ld, long long b_piece_stride, long long b_batch_stride, int b_pieces, int b_mn,
                                    float* mag, long long strideMag, unsigned int* clip_max, const int* valid_rows, int batch);

/* ---- audio ingest (utilities.py:241-242 `librosa.load(sr=16000)`, datasets/asap.py:80-86) ---------------------------------
 * y[c][m] = sum_i x[c][i] * h[taps/2 + m*down - i*up]: rational resampling with a host-designed odd-length FIR (scaled by `up`);
 * n_out = ceil(n_in * up / down).  x: (channels, n_in), y: (channels, n_out). */
PA2S_API int pa2s_resample_poly(void* stream, const float* x, int channels, long long n_in, int up, int down, const float* h, int taps,
                                float* y, long long n_out);

/* ---- VQT front end (utilities.py:246-253) -------------------------------------------------------------------
 * C: (nclips*rows_per_clip, 2*nb) filterbank responses (re,im interleaved).  Writes
 * out = amplitude_to_db(|V|, ref=max over the clip, amin=1e-5, top_db=80)/80 + 1 as (nclips, rows_per_clip, nb).
 * valid_rows (int32 per clip, or NULL): frames the un-padded clip has; later rows are excluded from the maximum and written as zeros
 * (the zero padding of pad_spectrogram, datasets/asap.py:345-349). */
PA2S_API int pa2s_vqt_post(void* stream, const float* C, float* out, unsigned int* clip_max, int nclips, int rows_per_clip, int nb,
                          const int* valid_rows);
/* the dB / scale half alone, in place on magnitudes written by pa2s_gemm_bf16_tma_mag */
PA2S_API int pa2s_vqt_logscale(void* stream, float* mag, const unsigned int* clip_max, int nclips, int rows_per_clip, int nb,
                              const int* valid_rows);

/* ---- ConvStack (models.py:463-543) ---------------------------------------------------------------------------
 * mode 0: Y = conv3x3(relu?(X*in_scale+in_shift)) (in_scale NULL = identity), Wpacked = W.permute(2,3,1,0);
 *         partial (if not NULL) gets one [sum y, sum y^2] row per CTA (pa2s_conv3x3_num_partials rows).
 * mode 1: data gradient (conv2d backward wrt input): X = dL/d(relu out) of the layer, the BatchNorm+ReLU backward
 *         transform k1*(g-k2-xhat*k3) is applied on load; Wpacked = W.flip(2,3).permute(2,3,0,1). */
PA2S_API int pa2s_conv3x3_num_partials(int B, int T, int F, int ntile);
PA2S_API int pa2s_conv3x3(void* stream, int mode, int B, int T, int F, int Cin, int Cout, const float* X, const float* Wpacked,
                          float* Y, float* partial, int ntile,
                          const float* in_scale, const float* in_shift, int in_relu,
                          const float* Yraw, const float* zs, const float* zb, const float* mean, const float* invstd,
                          const float* k1, const float* k2, const float* k3);
/* The same two convolutions as an implicit GEMM on the tcgen05 tensor cores (tc_conv.cu; conv2..conv4 shapes).  Wpack is
 * produced by pa2s_tc_conv_pack (dgrad = 0 forward filter, 1 = flipped/transposed filter of the data gradient) into a buffer
 * of pa2s_tc_conv_pack_bytes(Kin, Nout) bytes; partial has pa2s_tc_conv_num_partials rows of [sum y, sum y^2]. */
PA2S_API int pa2s_tc_conv_pack_bytes(int Kin, int Nout);
PA2S_API int pa2s_tc_conv_pack(void* stream, const float* W, int Cout, int Cin, int dgrad, void* out);
PA2S_API int pa2s_tc_conv_num_partials(int B, int T, int F);
PA2S_API int pa2s_tc_conv3x3(void* stream, int mode, int B, int T, int F, int Cin, int Cout, const float* X, const void* Wpack,
                             float* Y, float* partial, int nsplit,
                             const float* in_scale, const float* in_shift, int in_relu,
                             const float* Yraw, const float* zs, const float* zb, const float* mean, const float* invstd,
                             const float* k1, const float* k2, const float* k3);
PA2S_API int pa2s_tc_conv_wgrad_num_partials(int B, int T, int F);
PA2S_API int pa2s_tc_conv3x3_wgrad(void* stream, int B, int T, int F, int Cin, int Cout, const float* Xin, const float* G,
                                   float* partial, int nsplit, const float* in_scale, const float* in_shift, int in_relu,
                                   const float* Yraw, const float* zs, const float* zb, const float* mean, const float* invstd,
                                   const float* k1, const float* k2, const float* k3);
/* The same three convolutions fed by bulk async copies only (tc_conv_tma.cu).  Activation operands are bf16 "plane" tensors
 * P[b][t+1][piece][8-channel group][f+2][8] with zero halos (pa2s_planes_bytes bytes), written once per tensor by
 * pa2s_planes_fwd (BatchNorm-apply + ReLU + hi/lo split; replaces bn_i + relu of models.py:525-534 on the operand side) or
 * pa2s_planes_bwd (BatchNorm/ReLU backward + split) and read by two kernels each.  pa2s_conv_tma is the forward convolution
 * (Wpack from pa2s_tc_conv_pack with dgrad = 0) or the data gradient (planes = dy, dgrad = 1); pa2s_conv_tma_wgrad the weight
 * gradient (partial: pa2s_conv_tma_wgrad_num_partials rows of Cout*Cin*9, torch order). */
PA2S_API long long pa2s_planes_bytes(int B, int T, int F, int C, int npieces);
PA2S_API int pa2s_planes_fwd(void* stream, int B, int T, int F, int C, const float* X, const float* scale, const float* shift, int relu,
                             void* planes, int npieces);
PA2S_API int pa2s_planes_bwd(void* stream, int B, int T, int F, int C, const float* G, const float* Yraw, const float* zs, const float* zb,
                             const float* mean, const float* invstd, const float* k1, const float* k2, const float* k3,
                             void* planes, int npieces);
/* 1 (default): conv_tma3_kernel, the three kx taps share one read of the activation window (126 outputs per tile);
 * 0: conv_tma_kernel, one instruction group per tap.  Same results up to fp32 summation order. */
PA2S_API int pa2s_conv_tma_set_impl(int impl);
PA2S_API int pa2s_conv_tma_get_impl(void);      /* returns the selection, not a status */
/* Three-piece (fp32-level) convolution of the eval mode (models.py:526-534 in exact fp32): a = a1 + a2 + a3, W = W1 + W2 + W3 (bf16 pieces);
 *   pass 1  pa2s_conv_tma      planes (a1, a2) [pa2s_planes_fwd],     pack3 sel 0 (W1, W2)
 *   pass 2  pa2s_conv_tma_acc  the same planes,                        pack3 sel 1 (W3, 0)
 *   pass 3  pa2s_conv_tma_acc  planes (a3, a2) [pa2s_planes_fwd_low], pack3 sel 2 (W2, W1)
 * sums every piece product except a3*W3 (2^-32 relative). */
PA2S_API int pa2s_planes_fwd_low(void* stream, int B, int T, int F, int C, const float* X, const float* scale, const float* shift, int relu,
                                 void* planes);
PA2S_API int pa2s_tc_conv_pack3(void* stream, const float* W, int Cout, int Cin, int dgrad, int sel, void* out);
PA2S_API int pa2s_conv_tma_acc(void* stream, int B, int T, int F, int Cin, int Cout, const void* planes, const void* Wpack, float* Y);
/* data gradient (planes = dy of layer i, Wpack packed with dgrad = 1) fused with the statistics pass of the BatchNorm/ReLU backward
 * of layer i-1 (reference: autograd of models.py:525-534): partial = pa2s_conv_tma_num_partials rows of [sum g, sum g*xhat],
 * g = Y * (Yraw*zs + zb > 0), xhat = (Yraw - mean) * invstd -- the sums pa2s_colstats(mode 1) forms in a separate pass. */
PA2S_API int pa2s_conv_tma_dgrad_stats(void* stream, int B, int T, int F, int Cin, int Cout, const void* planes, int npieces,
                                       const void* Wpack, float* Y, const float* Yraw, const float* zs, const float* zb,
                                       const float* mean, const float* invstd, float* partial);
PA2S_API int pa2s_conv_tma_num_partials(int B, int T, int F);
PA2S_API int pa2s_conv_tma_wgrad_num_partials(int B, int T, int F);
PA2S_API int pa2s_conv_tma(void* stream, int B, int T, int F, int Cin, int Cout, const void* planes, int npieces, const void* Wpack,
                           float* Y, float* partial);
PA2S_API int pa2s_conv_tma_wgrad(void* stream, int B, int T, int F, int Cin, int Cout, const void* planes_in, const void* planes_dy,
                                 int npieces, float* partial);
/* conv2d backward wrt weight; partial is [nctas][Cout*Cin*9] in torch (Cout,Cin,3,3) order. */
PA2S_API int pa2s_conv3x3_wgrad(void* stream, int B, int T, int F, int Cin, int Cout, const float* Xin, const float* G,
                                float* partial, int nctas, const float* in_scale, const float* in_shift, int in_relu,
                                const float* Yraw, const float* zs, const float* zb, const float* mean, const float* invstd,
                                const float* k1, const float* k2, const float* k3);
/* conv1 (models.py:526: Conv2d(1, 20, 3x3, padding 1, bias=False) on the spectrogram) as a coalesced fp32 stream: X (B,T,F),
 * W (20,1,3,3) in torch layout, Y (B,T,F,20); partial (or NULL) = nctas rows of [sum y (20), sum y^2 (20)] for the BatchNorm
 * batch statistics.  pa2s_conv1_wgrad: conv1.weight gradient for dy = k1*(G*(Yraw*zs+zb > 0) - k2 - (Yraw-mean)*invstd*k3)
 * (BatchNorm + ReLU backward formed on the fly); partial = nctas rows of 180 (torch order), summed by pa2s_reduce_rows. */
PA2S_API int pa2s_conv1_fwd(void* stream, int B, int T, int F, const float* X, const float* W, float* Y, float* partial, int nctas);
PA2S_API int pa2s_conv1_wgrad(void* stream, int B, int T, int F, const float* X, const float* G, const float* Yraw, const float* zs,
                              const float* zb, const float* mean, const float* invstd, const float* k1, const float* k2, const float* k3,
                              float* partial, int nctas);
/* out[n] (=|+=) sum_r partial[r][n], accumulated in fp64. */
PA2S_API int pa2s_reduce_rows(void* stream, const float* partial, int R, int N, double* out64, float* out32, int accumulate);
/* Same sum for a TALL matrix (bias gradients: R = B*T rows), two deterministic stages over `nchunks` row chunks;
 * scratch holds nchunks*N doubles.  Replaces the bias-gradient reductions autograd derives for nn.GRU / nn.Linear
 * (models.py:63-67,444). */
PA2S_API int pa2s_colsum(void* stream, const float* X, int R, int N, double* scratch, int nchunks, float* out32, int accumulate);
/* per-channel sums over (npix, C): mode 0 [sum x, sum x^2]; mode 1 [sum g, sum g*xhat] (BatchNorm backward). */
PA2S_API int pa2s_colstats(void* stream, int mode, const float* X, const float* G, const float* mask, long long npix, int C,
                           const float* zs, const float* zb, const float* mean, const float* invstd, float* partial, int nctas);
/* nn.BatchNorm{1,2}d (models.py:499-505) train-mode statistics -> affine, running buffers updated in place.
 * count <= 0: the element count is read from sums[2*C] (SyncBatchNorm: it was all-reduced together with the sums, so ranks may hold
 * different batch sizes); the same convention holds for pa2s_bn_bwd_finalize. */
PA2S_API int pa2s_bn_finalize(void* stream, const double* sums, double count, int C, const float* gamma, const float* beta,
                              float eps, float momentum, float* running_mean, float* running_var,
                              float* scale, float* shift, float* mean, float* invstd);
PA2S_API int pa2s_bn_eval_affine(void* stream, int C, const float* gamma, const float* beta, const float* rm, const float* rv,
                                 float eps, float* scale, float* shift, float* mean, float* invstd);
PA2S_API int pa2s_bn_bwd_finalize(void* stream, const double* sums, double count, int C, const float* gamma, const float* invstd,
                                  float* dgamma, float* dbeta, float* k1, float* k2, float* k3);
/* out = relu(Z*scale+shift)*mask : out_bn + ReLU + dropout(0.2) (models.py:539-541). */
PA2S_API int pa2s_bn_relu_mask(void* stream, const float* Z, const float* scale, const float* shift, const float* mask,
                               float* out, long long npix, int C);
PA2S_API int pa2s_bn_bwd_apply(void* stream, const float* G, const float* Yraw, const float* mask, long long npix, int C,
                               const float* zs, const float* zb, const float* mean, const float* invstd,
                               const float* k1, const float* k2, const float* k3, float* out);

/* ---- Encoder BiGRU recurrence (models.py:63-67,77): gi = x W_ih^T + b_ih for all directions, H = 256 ----------- */
PA2S_API int pa2s_gru_seq_max_bg(void);
/* state exchange of the encoder recurrences: 0 DSMEM stores + cluster barrier (round 1); 1 (default) st.async + mbarrier with the
 * row-owner reverse kernel (4 KB instead of 12 KB exchanged per CTA and step); 3 st.async + mbarrier with the column-owner reverse kernel */
PA2S_API int pa2s_gru_seq_set_exchange(int mode);
PA2S_API int pa2s_gru_seq_fwd(void* stream, int B, int T, int ND, int H, int bg, const float* gi, const float* Whh, const float* bhh,
                              float* out, float* gates, float* hN);
PA2S_API int pa2s_gru_seq_bwd(void* stream, int B, int T, int ND, int H, int bg, const float* Whh, const float* out, const float* gates,
                              const float* dOut, const float* dhN, float* dgi, float* dgh);
/* ---- staff summariser: note_emb -> packed BiGRU(16->32) -> h_n (models.py:107-111,164-189) -------------------- */
PA2S_API int pa2s_staff_gru_fwd(void* stream, int B, int L, int I, int H, const long long* tokens, const long long* lengths,
                                const float* emb, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                                float* hN, float* hs, float* gates);
PA2S_API int pa2s_staff_gru_bwd(void* stream, int B, int L, int I, int H, const long long* tokens, const long long* lengths,
                                const float* emb, const float* w_ih, const float* w_hh, const float* hs, const float* gates,
                                const float* dhN, float* d_emb, float* d_w_ih, float* d_w_hh, float* d_b_ih, float* d_b_hh);
/* ---- gate non-linearity of one GRU cell (bar-level GRU, models.py:117-120,247) --------------------------------- */
PA2S_API int pa2s_gru_gates_fwd(void* stream, int B, int H, const float* gi, const float* gh, const float* hprev, float* hnew, float* save);
PA2S_API int pa2s_gru_gates_bwd(void* stream, int B, int H, const float* dh, const float* save, const float* hprev,
                                float* dgi, float* dgh, float* dhprev);

/* ---- note-level attention decoder (models.py:366-420, 452-461) ------------------------------------------------
 * `args` points to a HOST struct DecArgs (layout in piano_a2s_b200/csrc/dec_args.cuh, mirrored by ctypes in
 * piano_a2s_b200/_lib.py; pa2s_dec_args_size() lets the binding verify the layout). */
PA2S_API int pa2s_dec_args_size(void);
PA2S_API int pa2s_note_decoder_fwd(void* stream, const void* args, int sos_id, int eos_id);
PA2S_API int pa2s_note_decoder_bwd(void* stream, const void* args);
/* persistent variants: all steps of the call in ONE cooperative kernel of pa2s_dec_persist_grid() CTAs (weights and the
 * recurrent state resident in shared memory, 3 grid barriers per step); args->sync = 2 zeroed uint32. */
PA2S_API int pa2s_dec_persist_grid(void);
PA2S_API int pa2s_note_decoder_fwd_persist(void* stream, const void* args, int sos_id, int eos_id);
/* reverse pass: (1) pa2s_dec_dlogits fills args->dlogits_all (log-softmax backward of every step), (2) the caller forms
 * args->dhc_all = dlogits_all @ W_out with pa2s_gemm_*, (3) pa2s_note_decoder_bwd_persist runs the sequential chain in one
 * cooperative kernel and then the deferred dEp / dv accumulation (dv_part: B * pa2s_dec_deferred_blocks(T) rows). */
PA2S_API int pa2s_dec_dlogits(void* stream, const void* args);
PA2S_API int pa2s_dec_deferred_blocks(int T);
PA2S_API int pa2s_note_decoder_bwd_persist(void* stream, const void* args);
/* the two halves of pa2s_note_decoder_bwd_persist on their own: the sequential chain (cooperative kernel), and the parallel
 * dEp / dv accumulation, which only reads ds_all / qs / Ep / v and may therefore run on another stream behind the chain. */
PA2S_API int pa2s_note_decoder_bwd_chain(void* stream, const void* args);
PA2S_API int pa2s_note_decoder_bwd_deferred(void* stream, const void* args);

/* ---- multi-sequence persistent note decoder (dec_multi.cu): NQ bars of one staff x B clips per launch ---------------------
 * Replaces NoteDecoder.decode_notes (models.py:366-420) for every run of bars whose input tokens do not depend on the previous
 * bar's predictions (teacher-forced bars, models.py:289-311), and the autograd of all bars of a staff in ONE reverse launch.
 * `args` points to a HOST struct DecMArgs (csrc/decm_args.cuh, mirrored by ctypes in piano_a2s_b200/_lib.py).  The attention
 * memory is passed as Ee = exp(2 * (enc W_e^T + b)) (pa2s_exp2x): tanh(q + Ep) = 1 - 2 / (1 + exp(2q) * Ee), one SFU op per
 * element.  Backward: pa2s_decm_dlogits, caller GEMM dhc_all = dlogits_all @ W_out, pa2s_decm_bwd_chain, pa2s_decm_bwd_deferred. */
PA2S_API int pa2s_decm_args_size(void);
PA2S_API int pa2s_decm_max_queries(void);
PA2S_API int pa2s_decm_tile_max(void);
PA2S_API int pa2s_decm_grid(void);
PA2S_API int pa2s_decm_deferred_blocks(int T);
PA2S_API int pa2s_exp2x(void* stream, const float* x, float* y, long long n);
PA2S_API int pa2s_decm_fwd(void* stream, const void* args, int sos_id, int eos_id);
PA2S_API int pa2s_decm_dlogits(void* stream, const void* args);
PA2S_API int pa2s_decm_bwd_chain(void* stream, const void* args);
PA2S_API int pa2s_decm_bwd_deferred(void* stream, const void* args);
PA2S_API int pa2s_attn_step_fwd(void* stream, const void* args);
PA2S_API int pa2s_attn_step_bwd(void* stream, const void* args);

/* ---- loss and optimiser (pretrain.py:56-93, 125-128; pretrain.yaml:44-55) -------------------------------------- */
PA2S_API int pa2s_nll_fwd(void* stream, const float* logp, const long long* tgt, long long rows, int V, long long ignore, float* acc2);
PA2S_API int pa2s_nll_bwd(void* stream, float* grad, const long long* tgt, long long rows, int V, long long ignore,
                          const float* acc2, const float* gout);
PA2S_API int pa2s_sumsq(void* stream, const float* g, long long n, double* out, int zero_first);
PA2S_API int pa2s_adadelta(void* stream, float* p, const float* g, float* sq, float* acc, long long n, const double* sumsq,
                           float max_norm, float lr, float rho, float eps, float* norm_out);
/* ---- token post-processing -----------------------------------------------------------------------
 * tokens[seq][r] = argmax_v logp[seq][r][v] (lowest index on ties, like torch.argmax) and lengths[seq] = index of the first
 * <eos> token (L if none): `pred = outs.argmax(-1)` + `unpad` of pretrain.py:97-117,245-249 / finetune.py:86-108 for all
 * (clip, bar) sequences of one staff in one launch.  tokens is int64 (nseq, L), lengths int32 (nseq). */
PA2S_API int pa2s_greedy_tokens(void* stream, const float* logp, long long nseq, int L, int V, int eos, long long* tokens, int* lengths);
/* ---- batch collation ---------------------------------------------------------------------------------
 * `pad_spectrogram` of datasets/syn.py:46-58 and datasets/asap.py:338-350 for B clips in one launch: `packed` holds the clips'
 * (n_b, F) fp32 spectrograms back to back, row_off (B+1 int64, device) their first rows; out (B, 1, Tmax, F) receives the first
 * min(n_b, Tmax) frames of each clip and zeros after them. */
/* ---- evaluation metric right after the path ------------------------------------------------------------
 * Per clip, the counts jiwer.wer(target, pred) of `calculate_wer` (pretrain.py:216-227) is made of: hyp / ref are (nclips, bars, L)
 * int64 token rows (greedy tokens / targets); a row ends at its first `eos`; tokens equal to skip_a / skip_b (the pure-whitespace
 * labels "\t" and "\n", which jiwer's whitespace collapsing removes) are dropped; `sep` (any id outside the vocabulary) stands for
 * the "=" word between bars.  dist = Levenshtein distance of the two word sequences, nref / nhyp their lengths (int32 per clip);
 * wer = dist / nref. */
PA2S_API int pa2s_wer_counts(void* stream, const long long* hyp, const long long* ref, int nclips, int bars, int Lh, int Lr, int eos,
                             int skip_a, int skip_b, int sep, int* dist, int* nref, int* nhyp);
/* ---- audio ingest (the arithmetic of datasets/asap.py:83-86) ---------------------------------------------
 * audio (channels, n) fp32 -> out (n): mean over channels, then divided by max|mean| (IEEE division; an all-zero clip gives NaN
 * like the reference's 0/0).  scratch: one uint32 on the device. */
PA2S_API int pa2s_mono_peak_normalize(void* stream, const float* audio, int channels, long long n, float* out, unsigned int* scratch);
PA2S_API int pa2s_pad_spectrograms(void* stream, const float* packed, const long long* row_off, int B, int Tmax, int F, float* out);
#endif
