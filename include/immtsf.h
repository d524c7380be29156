/*
 * immtsf.h -- C ABI of the B200-native (sm_100a) IMM-TSF text->time-series
 * fusion kernels.
 *
 * This is the drop-in boundary for the hot path named in BASELINE.json: the
 * reference's `fusions/FusionModel.py:98-113` (ttf -> mmf) and the four modules
 * behind it.  The reference has no FFI of its own (it is pure PyTorch); these
 * entry points are what a binding for that path calls instead of the ATen op
 * sequences listed in SURVEY.md section 2.2.  Each entry cites the reference
 * lines it replaces.
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer owned by the caller (fp32 unless the
 *     type says otherwise); nothing is allocated, freed or synchronised here;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value: 0 on success, negative on error (IMMTSF_ERR_*); the text
 *     of the last error on the calling thread is immtsf_last_error_string();
 *   - there is no CPU fallback and no other backend: on a device that is not
 *     compute capability 10.x every launch fails with IMMTSF_ERR_ARCH/LAUNCH;
 *   - ragged layout: notes are stored compacted, sample-major, in
 *     `[M_alloc, d]` row buffers with `offsets[B+1]` (int32).  `sumN =
 *     offsets[B]` lives on the device; kernels take it as `const int32_t*
 *     m_dev` so no host sync is needed.  Producers zero rows
 *     [sumN, roundup(sumN,128)) so that reductions over rows stay clean.
 *   - dropout masks are Philox4x32-10 functions of (seed, site, element index)
 *     (csrc/common.cuh: 16 random bits per element, one call per 8 elements),
 *     regenerated in backward; `drop_thr = floor(p*2^16)`, 0 disables dropout.
 */
#ifndef IMMTSF_H_
#define IMMTSF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMMTSF_ABI_VERSION 1

/* ---- library ---------------------------------------------------------- */
int immtsf_version(void);
const char* immtsf_last_error_string(void);
/* 1 if `device` is compute capability 10.x (B200), 0 otherwise, <0 on error */
int immtsf_device_supported(int device);
/* number of kernel launches issued by this library in this process (monotonic) */
unsigned long long immtsf_launch_count(void);

/* Dropout seeds are passed by value; when a device pointer is registered here (process-wide, NULL clears it)
 * every kernel adds *dev_ptr to its seed argument at run time.  A CUDA graph that captured one training step
 * (immtsf/runtime.py GraphedStep) bumps that word with immtsf_seed_advance inside the graph, so each replay
 * draws fresh Philox masks although the by-value seeds are frozen in the graph. */
int immtsf_set_seed_offset_ptr(const uint64_t* dev_ptr);
int immtsf_seed_advance(uint64_t* dev_ptr, uint64_t inc, void* stream);

/* Per-launch device timing of the tcgen05 GEMM kernel (bench.py's roofline): between begin and end every
 * gemm_tc_kernel launch is bracketed by a CUDA-event pair on its launch stream.  end() waits for them and returns
 * the number of records copied (shape, ragged_dim and milliseconds each). */
int immtsf_profile_begin(int max_records);
int immtsf_profile_end(int* M, int* N, int* K, int* ragged_dim, float* ms, int cap);
/* Diagnostics of the CTA-pair tcgen05 kernel: while buf != NULL every CTA writes clock64 stamps of its phases to
 * buf[cta * 8 + slot]: 0 entry, 1 prologue done, 2 first operands landed (leader), 3 MMAs issued (leader),
 * 4 accumulator complete, 5 stores done, 6 exit; slot 7 = globaltimer (ns) at entry.  NULL switches it off. */
int immtsf_gemm_trace(long long* buf);

/* ---- K1: padded -> ragged CSR (replaces the content mask of
 * fusions/TTF_RecAvg.py:69 / fusions/TTF_T2V_XAttn.py:107, the NaN guard of
 * :75 / :116 and M_txt of :110 / :124) ----------------------------------
 * notes [B,N,d_m], tau [B,N]  ->  note_mask [B*N] u8, offsets [B+1],
 * rows [B*N] (flat index b*N+n of each valid note), seg [B*N] (owning
 * sample), emb_flat [M_alloc,d_m], tau_flat [M_alloc], m_txt [B] u8.
 * flags[0] is set to 1 if any NaN is present in notes. */
int immtsf_csr_build(const float* notes, const float* tau, int B, int N, int d_m,
                     uint8_t* note_mask, int32_t* offsets, int32_t* rows, int32_t* seg,
                     float* emb_flat, float* tau_flat, uint8_t* m_txt, int32_t* flags,
                     int M_alloc, void* stream);
/* same, with the compacted rows written at a leading dimension ld_emb >= d_model (the rows then sit in the left columns of a
 * wider buffer, e.g. the [emb ; phi] operand of the collapsed T2V schedule) and, when emb_lo != NULL, their tcgen05 lo
 * operand x - trunc_tf32(x) written beside them (leading dimension ld_lo). */
int immtsf_csr_build_ex(const float* notes, const float* tau, int B, int N, int d_m, uint8_t* note_mask,
                        int32_t* offsets, int32_t* rows, int32_t* seg, float* emb_flat, int ld_emb, float* emb_lo,
                        int ld_lo, float* tau_flat, uint8_t* m_txt, int32_t* flags, int M_alloc, void* stream);
/* flags[slot] = 1 if x[0..n) holds a NaN (FusionModel.py:103,107,111) */
int immtsf_nan_check(const float* x, size_t n, int32_t* flags, int slot, void* stream);
/* zero rows [sumN, min(roundup(sumN,128), M_alloc)) of X[M_alloc, ncols] (ld) */
int immtsf_zero_pad_rows(float* X, int ld, int ncols, const int32_t* m_dev, int M_alloc, void* stream);

/* ---- dense projections (nn.Linear call sites: TTF_RecAvg.py:80,109;
 * TTF_T2V_XAttn.py:121,140,182; the in/out projections inside
 * nn.MultiheadAttention; MMF_GR_Add.py:46-47,54; MMF_XAttn_Add.py:68-70,83)
 * C[M,N] = alpha * op(A)[M,K] * op(B)[K,N] + beta * C + bias[N]
 *   transA=0: A is [M,K] row-major (lda)   transA=1: A is [K,M] row-major
 *   transB=0: B is [K,N] row-major (ldb)   transB=1: B is [N,K] row-major
 * ragged_dim: 0 none; 1: rows M bounded by *ragged (rows >= it are written
 * as 0 inside touched tiles); 2: contraction K bounded by *ragged (wgrad).
 * backend: 0 auto, 1 FFMA (CUDA cores, exact fp32), 2 tcgen05 3xTF32.
 * With backend 0, shapes with a dimension <= 32 (the channel count C) go to
 * streaming kernels (csrc/gemm_skinny.cu: exact fp32, one coalesced pass over
 * the wide operand, deterministic row-split reduction).
 * workspace: caller-owned device scratch for the tcgen05 backend's hi/lo
 * operand split and the row-split partial sums (>=
 * immtsf_gemm_workspace_bytes); with backend 0 and a NULL workspace the FFMA
 * kernel is used. */
int immtsf_gemm(int transA, int transB, int M, int N, int K, float alpha,
                const float* A, int lda, const float* B, int ldb, float beta,
                float* C, int ldc, const float* bias, const int32_t* ragged,
                int ragged_dim, int backend, void* workspace, size_t workspace_bytes,
                void* stream);
size_t immtsf_gemm_workspace_bytes(int transA, int transB, int M, int N, int K);
/* Same, with optional pre-split operands for the tcgen05 backend.  kind::tf32 reads fp32 containers and ignores
 * the low 13 mantissa bits, so the "hi" operand is the tensor itself; A_lo / B_lo = x - trunc_tf32(x) (from
 * immtsf_split_lo, same shape as the operand, leading dimension lda_lo / ldb_lo) let a caller that uses a tensor
 * in several products (forward, dgrad, wgrad) split it once.  NULL => split internally into the workspace.
 * C_lo (nullable, leading dimension ldc_lo >= roundup(N,4), multiple of 4): second output, C - trunc_tf32(C),
 * written by the epilogue that writes C, so that a product consuming C needs no split pass over it. */
int immtsf_gemm_ex(int transA, int transB, int M, int N, int K, float alpha,
                   const float* A, int lda, const float* A_lo, int lda_lo,
                   const float* B, int ldb, const float* B_lo, int ldb_lo, float beta,
                   float* C, int ldc, float* C_lo, int ldc_lo, const float* bias,
                   const int32_t* ragged, int ragged_dim, int backend, void* workspace,
                   size_t workspace_bytes, void* stream);
/* lo[rows][ld_lo] = src - trunc_tf32(src); rows bounded by roundup(*ragged,128) when ragged != NULL */
int immtsf_split_lo(const float* src, int ld, int rows, int cols, float* lo, int ld_lo,
                    const int32_t* ragged, void* stream);
/* Up to 4 independent products C_i = alpha_i * op(A_i) op(B_i) + beta_i * C_i in ONE launch (the d x d x d
 * weight-space folds / un-folds of MMF_XAttn_Add and TTF_T2V_XAttn, each too small to fill the GPU).  tcgen05 3xTF32,
 * 128 x 128 tiles, no bias / ragged bounds / split-K.  Every operand comes with its lo part; C_lo[i] (nullable) is the
 * second output of immtsf_gemm_ex.  All arrays are host arrays of length n. */
int immtsf_gemm_group(int n, const int* transA, const int* transB, const int* M, const int* N, const int* K,
                      const float* alpha, const float* const* A, const float* const* A_lo, const int* lda,
                      const int* lda_lo, const float* const* B, const float* const* B_lo, const int* ldb,
                      const int* ldb_lo, const float* beta, float* const* C, const int* ldc, float* const* C_lo,
                      const int* ldc_lo, void* stream);
/* dst [cols, rows] = src^T and (lo != NULL) lo = dst - trunc_tf32(dst): a weight transposed ONCE so that the data-gradient
 * product dx = dy W reads a K-major operand (the tcgen05 kernel's fast path) instead of an MN-major one. */
int immtsf_transpose_split(const float* src, int ld, int rows, int cols, float* dst, int ldd, float* lo, int ldl,
                           void* stream);

/* One launch for up to 16 small tensors (a module's weight matrices): task i reads src[i] (rows[i] x cols[i],
 * leading dimension ld_src[i]) and writes a plain copy to hi[i] (nullable: e.g. a slice of a packed operand) and
 * src - trunc_tf32(src) to lo[i] (nullable).  The arrays are host arrays of length n. */
int immtsf_multi_split(int n, const float* const* src, const int* ld_src, const int* rows, const int* cols,
                       float* const* hi, const int* ld_hi, float* const* lo, const int* ld_lo, void* stream);
/* kernel family immtsf_gemm would pick for this call: 1 FFMA, 2 tcgen05, 3 skinny streaming kernels */
int immtsf_gemm_plan(int transA, int transB, int M, int N, int K, const float* A, int lda,
                     const float* B, int ldb, const float* C, int ldc, int backend);
/* out[N] = beta*out + sum_m X[m, :]  (bias gradients: the `.sum(0)` autograd
 * emits for every nn.Linear bias on the path).  Rows bounded by *ragged when
 * given.  workspace (>= immtsf_gemm_workspace_bytes(0,0,1,N,M)) enables the
 * row-split two-pass kernel; without it a single-pass kernel is used. */
int immtsf_colsum(const float* X, int M, int N, int ldx, float* out, float beta,
                  const int32_t* ragged, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K2: TTF_RecAvg pooling (TTF_RecAvg.py:94-106): recency weights,
 * weighted mean over each ragged segment, LayerNorm, dropout ------------- */
int immtsf_recavg_pool_fwd(const float* Vp, int ldv, const float* tau_flat, const int32_t* offsets,
                           const float* t_hat, int t_hat_bstride, const float* log_sigma,
                           const float* gamma, const float* beta, int B, int T, int d, int N_max, float eps,
                           uint32_t drop_thr, uint64_t seed, float* E_drop, float* E_raw,
                           float* mean, float* rstd, float* wsum, void* stream);
/* dS: caller-owned scratch of B*T*(d+1) floats (the LayerNorm-backward rows [B*T, d] followed by one scalar per
 * row, written by the first of the two kernels and read by the second); N_max bounds the notes per sample
 * (grid size only).  dlog_sigma is a DOUBLE accumulator (zeroed by the caller): its terms cancel almost completely,
 * so they are summed in double from the thread level up. */
int immtsf_recavg_pool_bwd(const float* dE_drop, const float* E_raw, const float* mean, const float* rstd,
                           const float* wsum, const float* Vp, int ldv, const float* tau_flat,
                           const int32_t* offsets, const float* t_hat, int t_hat_bstride,
                           const float* log_sigma, const float* gamma, int B, int T, int d, int N_max,
                           uint32_t drop_thr, uint64_t seed, float* dS, float* dVp, int lddv, float* dgamma,
                           float* dbeta, double* dlog_sigma, void* stream);

/* ---- K2 at long segments x long windows (N_max, T >= 64): the pooling of TTF_RecAvg.py:100 is a dense [T x N_i] . [N_i x d]
 * product per sample and runs on immtsf_gemm_batched (tcgen05, 3xTF32); these are the kernels around it
 * (csrc/recavg_tc.cu).  Np = N_max rounded up to a multiple of 4 (the contraction / leading dimension).
 *   recavg_weights: Wn [B, T, Np] = w_nt / max(sum_n w_nt, 1e-6) (zero for n >= N_b); wsum [B*T] (nullable) = sum_n w_nt;
 *                   Cn [B, T, Np] (nullable, with csum [B*T]) = Wn * 2 (delta/sigma)^2, csum = sum_n Cn.
 *   Wn_lo / Cn_lo / dst_lo (nullable, same shapes): the operands' 3xTF32 lo parts x - trunc_tf32(x), written by the producer.
 *   csr_to_padded / padded_to_csr: ragged rows <-> [B, Np, d] with zero rows beyond N_b.
 *   recavg_dls: *dlog_sigma (double, accumulated) += sum_{r,j} dE_raw[r,j] (R[r,j] - csum[r] E_raw[r,j]), R = Cn V'. */
int immtsf_recavg_weights(const float* tau_flat, const int32_t* offsets, const float* t_hat, int t_hat_bstride,
                          const float* log_sigma, int B, int T, int Np, float* Wn, float* Cn, float* Wn_lo,
                          float* Cn_lo, float* wsum, float* csum, void* stream);
int immtsf_csr_to_padded(const float* src, int lds, const int32_t* offsets, int B, int Np, int d, float* dst,
                         float* dst_lo, void* stream);
int immtsf_padded_to_csr(const float* src, const int32_t* offsets, int B, int Np, int d, float* dst, int ldd,
                         void* stream);
int immtsf_recavg_dls(const float* dE_raw, const float* R, const float* E_raw, const float* csum, long rows, int d,
                      double* dlog_sigma, void* stream);

/* ---- Time2Vec (TTF_T2V_XAttn.py:20-24,136) written straight into the
 * [V' ; phi] concat buffer (:139); out_lo (nullable): phi - trunc_tf32(phi) into the
 * matching columns of that buffer's tcgen05 lo operand ------------------- */
int immtsf_time2vec_fwd(const float* tau_flat, const float* w_lin, const float* b_lin, const float* w_per,
                        const float* b_per, int d_tau, float* out, int ld, float* out_lo, int ld_lo,
                        const int32_t* m_dev, int M_alloc, void* stream);
int immtsf_time2vec_bwd(const float* dphi, int ld, const float* tau_flat, const float* w_per,
                        const float* b_per, int d_tau, float* dw_lin, float* db_lin, float* dw_per,
                        float* db_per, const int32_t* m_dev, int M_alloc, void* stream);

/* ---- K3: single-learned-query attention over each ragged segment
 * (TTF_T2V_XAttn.py:143-166 + the softmax/dropout/bmm inside
 * nn.MultiheadAttention), computed ONCE per note instead of once per
 * (note, query).  q [d] is already scaled by hd^-1/2.  KVp [M_alloc, 2d]:
 * keys in columns [0,d), values in [d,2d).  R = B*T rows when per_query
 * (training with attention dropout) else B rows. ------------------------- */
int immtsf_segattn_fwd(const float* q, const float* KVp, const int32_t* offsets, int B, int T, int H,
                       int d, int N_max, int per_query, uint32_t drop_thr, uint64_t seed,
                       float* attn_cat, float* probs, void* stream);
int immtsf_segattn_bwd(const float* d_attn_cat, const float* q, const float* KVp, const float* probs,
                       const int32_t* offsets, int B, int T, int H, int d, int N_max, int per_query,
                       uint32_t drop_thr, uint64_t seed, float* dKVp, float* dq_partial, void* stream);

/* ---- one head, train mode: the attention above + out-projection bias + learned query residual + LayerNorm + dropout
 * (TTF_T2V_XAttn.py:143-179) in ONE launch, eight query rows per CTA; attn_cat [B*T, d] (the pooled rows, saved for backward),
 * probs [sum N] and mean / rstd [B*T] are what immtsf_ln_bwd / immtsf_segattn_bwd need.  Same results as immtsf_segattn_fwd
 * (H = 1, per_query = 1) followed by immtsf_ln_fwd(x = attn_cat, xbias, res, valid = "sample has notes", rows_per_sample = T). */
int immtsf_segattn_ln_ok(int d, int N_max);
int immtsf_segattn_ln_fwd(const float* q, const float* KVp, const int32_t* offsets, int B, int T, int d, int N_max,
                          uint32_t drop_thr, uint64_t seed, const float* xbias, const float* res, const float* gamma,
                          const float* beta, float eps, float* attn_cat, float* probs, float* y, float* mean,
                          float* rstd, void* stream);

/* ---- K3b: per-(note, query) Time2Vec attention (fusions/TTF_T2V_XAttn_old.py:119-143 -- Time2Vec
 * of the clamped lag max(t_hat - tau, 0) of every (note, query) pair enters keys and values).  With
 * X_nt = A_n + W_phi phi_nt (A = V' W_a^T + b_kv once per note, KV_proj.weight = [W_a | W_phi]) the caller supplies
 * A [M_alloc, d], the per-note score base a_sc = A U^T [M_alloc, H] (u_h = W_k[h]^T q_h, q scaled by hd^-1/2) and
 * g = U W_phi [H, d_tau]; the kernel evaluates sin() on the fly, takes softmax_n(a_sc + g_h . phi_nt) over each ragged
 * segment, applies attention dropout and returns, for output row r = (b*T + t)*H + h,
 *   Z [B*T*H, d] = sum_n P~ A_n,  Phi [B*T*H, d_tau] = sum_n P~ phi_nt,  sp [B*T*H] = sum_n P~
 * (head output = W_v[h] (Z + W_phi Phi) + sp b_v[h]: GEMMs of the caller).  probs (nullable) [H*T*M_alloc]:
 * softmax before dropout at (h*T + t)*M_alloc + row.  H <= 8.  t_hat_bstride = 0 for a shared 1-D t_hat. */
int immtsf_t2vq_attn_fwd(const float* A, int lda, const float* a_sc, const float* g, const float* tau_flat,
                         const int32_t* offsets, const float* t_hat, int t_hat_bstride, const float* w_lin,
                         const float* b_lin, const float* w_per, const float* b_per, int B, int T, int H, int d,
                         int d_tau, int N_max, int M_alloc, uint32_t drop_thr, uint64_t seed, float* Z, float* Phi,
                         float* sp, float* probs, void* stream);
/* Backward, two launches: (1) per (query tile, sample) CTA: dP~, dropout + softmax backward, the Time2Vec parameter
 * partials of the tile; (2) per (column block, sample) CTA: dA [M_alloc, d] = sum over the sample's (t, h) rows of P~ dZ
 * (the pooling part only; the caller adds da U) and da [M_alloc, H] = sum_t dS.  dpart [B * tiles, 2+H, d_tau]
 * (tiles = immtsf_t2vq_bwd_tiles): row 0 = d w (unit 0: time2vec.linear, units k >= 1: periodic k-1), row 1 = d b,
 * rows 2.. = dg[h]; the caller sums the rows (immtsf_colsum).  workspace: >= immtsf_t2vq_bwd_workspace_bytes (P~ and dS of
 * every (h, t, note) between the two launches).  Every output element is owned by one thread: no atomics. */
int immtsf_t2vq_bwd_tiles(int T, int H, int N_max);
size_t immtsf_t2vq_bwd_workspace_bytes(int T, int H, int M_alloc);
int immtsf_t2vq_attn_bwd(const float* dZ, const float* dPhi, const float* dsp, const float* A, int lda, const float* g,
                         const float* probs, const float* tau_flat, const int32_t* offsets, const float* t_hat,
                         int t_hat_bstride, const float* w_lin, const float* b_lin, const float* w_per,
                         const float* b_per, int B, int T, int H, int d, int d_tau, int N_max, int M_alloc,
                         uint32_t drop_thr, uint64_t seed, float* dA, int lddA, float* da, float* dpart,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- residual + LayerNorm + dropout over d (TTF_T2V_XAttn.py:171-179) ---
 * z = (valid[row / rows_per_sample] ? x + xbias : 0) + res;  y = dropout(LN(z))
 * xbias [d] (nullable): a bias that belongs to x (the MHA out_proj bias when out_proj is folded into V). */
int immtsf_ln_fwd(const float* x, int ldx, const float* xbias, const float* res, const uint8_t* valid, int rows_per_sample,
                  const float* gamma, const float* beta, int R, int d, float eps, uint32_t drop_thr,
                  uint64_t seed, uint32_t site, float* y, float* mean, float* rstd, void* stream);
int immtsf_ln_bwd(const float* dy, const float* x, int ldx, const float* xbias, const float* res, const uint8_t* valid,
                  int rows_per_sample, const float* gamma, const float* mean, const float* rstd, int R,
                  int d, uint32_t drop_thr, uint64_t seed, uint32_t site, float* dx, float* dres,
                  float* dgamma, float* dbeta, void* stream);

/* ---- K5: MMF_GR_Add (MMF_GR_Add.py:43-60).  G4 [B*T, 4C] = [Y;E] [W_ih;W_g]^T
 * + [b_ih;b_g] comes from immtsf_gemm; the scan consumes columns [0,3C) and
 * the tail columns [3C,4C). ----------------------------------------------- */
/* gates (nullable): [B*T, 4C] scratch for (r, z, n, W_hn h + b_hn).  When given and C > 32 the wide recurrence is
 * used (one CTA per sample, W_hh rows in registers) and backward must be handed the same buffer. */
int immtsf_gru_scan_fwd(const float* G4, const float* w_hh, const float* b_hh, int B, int T, int C,
                        float* h_all, float* h_prev, float* gates, void* stream);
int immtsf_gr_tail_fwd(const float* Y, const float* G4, const float* h_all, const float* w_r,
                       const float* b_r, const float* gamma, const float* beta, const uint8_t* m_txt,
                       int B, int T, int C, float eps, uint32_t drop_thr, uint64_t seed, float* Y_out,
                       int32_t* flags, void* stream);
/* backward returns pre-activation gradients; every weight gradient is then an
 * immtsf_gemm / immtsf_colsum over rows:
 *   dG4[:, 3C:4C] (gate logits), d_delta [B*T,C] (-> dW_r = d_delta^T h_all),
 *   dh_out [B*T,C] (into the scan) */
int immtsf_gr_tail_bwd(const float* dY_out, const float* G4, const float* h_all, const float* w_r,
                       const float* b_r, const float* gamma, const float* beta, const uint8_t* m_txt,
                       int B, int T, int C, float eps, uint32_t drop_thr, uint64_t seed, float* dG4,
                       float* d_delta, float* dh_out, float* dgamma, float* dbeta, void* stream);
/*   dG4[:, 0:3C] = [da_r,da_z,da_n], dGh [B*T,3C] = [da_r,da_z,da_n*r]
 *   (-> dW_hh = dGh^T h_prev, db_hh = colsum dGh) */
int immtsf_gru_scan_bwd(const float* G4, const float* h_prev, const float* w_hh, const float* b_hh,
                        const float* dh_out, const float* gates, int B, int T, int C, float* dG4, float* dGh,
                        void* stream);

/* ---- K6: MMF_XAttn_Add core (MMF_XAttn_Add.py:73-80 + MHA internals):
 * per (sample, head) T x T attention; q,k,v are [B*T, d] with leading dims. */
int immtsf_xattn_core_fwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                          const uint8_t* m_txt, int B, int T, int H, int d, uint32_t drop_thr,
                          uint64_t seed, float* o, int ldo, float* probs, void* stream);
int immtsf_xattn_core_bwd(const float* d_o, int lddo, const float* q, int ldq, const float* k, int ldk,
                          const float* v, int ldv, const float* probs, const uint8_t* m_txt, int B, int T,
                          int H, int d, uint32_t drop_thr, uint64_t seed, float* dq, int lddq, float* dk,
                          int lddk, float* dv, int lddv, void* stream);
/* Rank-(C+1) query path for T <= 32 (csrc/xattn_small.cu).  In MMF_XAttn_Add the queries are a projection of the
 * C-channel series (q_i = W y_i + b, MMF_XAttn_Add.py:68 + in_proj_q), so q_i . k_j = [y_i ; 1] . kq_j with
 * kq_j = [W^T k_j ; b . k_j]: the caller makes kq [B*T, H*(C+1)] with one skinny product per head and q is never
 * formed.  Backward returns dv, Z [B*T, H*(C+1)] (Z_j = sum_i dS_ij [y_i ; 1]: dk = Z [W | b]^T and
 * d[W | b] = k^T Z are two skinny products) and dyh [H][B*T][C] (per-head query-side gradient into Y_ts). */
int immtsf_xattn_lowrank_ok(int T, int H, int d, int C);
int immtsf_xattn_lowrank_fwd(const float* y, int ldy, const float* kq, int ldkq, const float* v, int ldv,
                             const uint8_t* m_txt, int B, int T, int H, int d, int C, uint32_t drop_thr,
                             uint64_t seed, float* o, int ldo, float* probs, void* stream);
int immtsf_xattn_lowrank_bwd(const float* d_o, int lddo, const float* y, int ldy, const float* kq, int ldkq,
                             const float* v, int ldv, const float* probs, const uint8_t* m_txt, int B, int T,
                             int H, int d, int C, uint32_t drop_thr, uint64_t seed, float* dv, int lddv,
                             float* z, int ldz, float* dyh, void* stream);
/* MMF_XAttn_Add as a rank-(2C+1) form in E_txt (csrc/xattn_rank.cu; T <= 32, C <= 31, H*(C+1) <= 64).  Both ends of
 * the attention are C-dimensional (queries = projections of the C-channel series, output = residual_head of the
 * MHA output), so with R = [kq | vo] = E_txt [A ; G]^T + [a0 ; g0]  ([B*T, H*(C+1) + H*C], one skinny product;
 * A = Wq_aug^T in_k W_K, G = (W_r W_o) in_v W_V per head, formed by skinny weight-space products)
 *   score_ij = [y_i ; 1] . kq_j / sqrt(hd),   delta_i = sum_h sum_j P~_ij vo_j + bo
 * and no tensor of width d (q, k, v, o or their gradients) exists.  Backward returns dR = [Z | U] and dY. */
int immtsf_xattn_rank_ok(int T, int H, int d, int C);
int immtsf_xattn_rank_fwd(const float* y, int ldy, const float* r, int ldr, const float* bo, const uint8_t* m_txt,
                          int B, int T, int H, int d, int C, uint32_t drop_thr, uint64_t seed, float* delta_y,
                          float* probs, void* stream);
int immtsf_xattn_rank_bwd(const float* d_delta, const float* y, int ldy, const float* r, int ldr, const float* probs,
                          const uint8_t* m_txt, int B, int T, int H, int d, int C, uint32_t drop_thr, uint64_t seed,
                          float* dr, int lddr, float* dy, void* stream);

/* ---- the data half of the rank form in ONE launch per direction (csrc/xattn_rank_fused.cu): R = E Wr^T + br from the
 * sample's own rows of E, the T x (2C+1) attention, and the LayerNorm_C / dropout / kappa-blend tail (MMF_XAttn_Add.py:83-102)
 * -> Y_out; backward: tail, attention, dE = dR Wr and dWr = dR^T E in one pass over E, per-CTA partials reduced in CTA
 * order by a second launch.  small = [dbr (nr) | d bo (C) | dgamma (C) | dbeta (C)].  Sets FLAG_E when E holds a NaN and
 * FLAG_OUT when delta / Y_out do.  Needs immtsf_xattn_rank_fused_ok; operands 16-byte aligned, ld % 4 == 0. */
int immtsf_xattn_rank_fused_ok(int T, int H, int d, int C, int de);
size_t immtsf_xattn_rank_fused_bwd_workspace_bytes(int B, int H, int C, int de);
int immtsf_xattn_rank_fused_fwd(const float* e, int lde, int de, const float* wr, int ldwr, const float* br,
                                const float* y, int ldy, const float* bo, const float* gamma, const float* beta,
                                const uint8_t* m_txt, int B, int T, int H, int d, int C, float eps, float kappa,
                                uint32_t drop_thr, uint64_t seed, float* r, int ldr, float* delta_y, float* probs,
                                float* y_out, int32_t* flags, void* stream);
int immtsf_xattn_rank_fused_bwd(const float* dy_out, const float* delta_y, const float* gamma, const float* y, int ldy,
                                const float* r, int ldr, const float* probs, const uint8_t* m_txt, const float* e,
                                int lde, int de, const float* wr, int ldwr, int B, int T, int H, int d, int C,
                                float eps, float kappa, uint32_t drop_thr, uint64_t seed, float* de_out, int ldde,
                                float* dy, float* dwr, float* small, void* workspace, size_t workspace_bytes,
                                void* stream);
/* Large-T form of the same core (T > 32: the T x T contractions are dense products and run on tcgen05):
 *   S = Q K^T (immtsf_gemm_batched) -> immtsf_softmax_rows_fwd -> O = P~ V (immtsf_gemm_batched), and in backward
 *   dP~ = dO V^T -> immtsf_softmax_rows_bwd -> dQ = dS K, dK = dS^T Q, dV = P~^T dO.
 * Batched product: C(b1,b2)[M,N] = alpha * op(A(b1,b2)) op(B(b1,b2)) + beta * C(b1,b2) with X(b1,b2) = X + b1*x_s1 +
 * b2*x_s2 (element strides, multiples of 4).  3xTF32 like immtsf_gemm; tiles that overhang a batch are zero-filled.
 * A_lo / B_lo (nullable): lo = x - trunc_tf32(x) over the operand's flat extent, same layout as the operand
 * (immtsf_split_lo with rows = 1), for operands used by several products. */
int immtsf_gemm_batched(int transA, int transB, int M, int N, int K, float alpha, const float* A, const float* A_lo,
                        int lda, long a_s1, long a_s2, const float* B, const float* B_lo, int ldb, long b_s1,
                        long b_s2, float beta, float* C, int ldc, long c_s1, long c_s2, int batch1, int batch2,
                        void* workspace, size_t workspace_bytes, void* stream);
size_t immtsf_gemm_batched_workspace_bytes(int transA, int transB, int M, int N, int K, int lda, long a_s1, long a_s2,
                                           int ldb, long b_s1, long b_s2, int batch1, int batch2);
/* S, Pt, dS: [B, H, T, Tp] score buffers (Tp >= T, a multiple of 4).  fwd: S <- P = softmax(scale*S) in place,
 * Pt <- P * dropout keep-scale.  bwd: dS holds dP~ on entry and scale * P * (dP~ ks - D) on exit, Pt <- P ks. */
int immtsf_softmax_rows_fwd(float* S, float* Pt, const uint8_t* m_txt, int B, int H, int T, int Tp, float scale,
                            uint32_t drop_thr, uint64_t seed, void* stream);
int immtsf_softmax_rows_bwd(float* dS, const float* P, float* Pt, const uint8_t* m_txt, int B, int H, int T, int Tp,
                            float scale, uint32_t drop_thr, uint64_t seed, void* stream);
/* tail (MMF_XAttn_Add.py:84-102): Y_out = (Y + kappa * m * dropout(LN_C(delta_y))) / (1+kappa) */
int immtsf_xattn_tail_fwd(const float* Y, const float* delta_y, const float* gamma, const float* beta,
                          const uint8_t* m_txt, int B, int T, int C, float eps, float kappa,
                          uint32_t drop_thr, uint64_t seed, float* Y_out, int32_t* flags, void* stream);
int immtsf_xattn_tail_bwd(const float* dY_out, const float* delta_y, const float* gamma,
                          const uint8_t* m_txt, int B, int T, int C, float eps, float kappa,
                          uint32_t drop_thr, uint64_t seed, float* d_delta_y, float* dgamma, float* dbeta,
                          void* stream);

/* ---- small helpers ----------------------------------------------------- */
/* y[n] = alpha * x[n] (+ y[n] if accumulate) */
int immtsf_axpby(const float* x, float alpha, float* y, int accumulate, size_t n, void* stream);
/* out[r, c] = sum_{t<T} x[r*T + t, c]   ([R*T, d] -> [R, d]) */
int immtsf_group_sum_rows(const float* x, int ldx, int R, int T, int d, float* out, void* stream);

/* ---- next to the path: the masked-MSE training loss, lib/evaluation.py:17-69 compute_error(...,
 * "MSE", "mean") as called at :107-113, and the all-zero-mask check of :128-132 --------------------------------------
 * partial: err_cnt[0..C) = sum (truth-pred)^2 mask, err_cnt[C..2C) = sum mask over the rows (= samples x T) of this
 *   shard, deterministic (per-CTA partials added in CTA order by the last CTA; ticket: a zero-initialised device word
 *   the kernel leaves at zero); *empty_flag (nullable) is set when a sample's mask is all zero.
 * finalize: loss = sum_c err_c / (cnt_c + 1e-8) / count_nonzero(cnt); scale_c = 1 / ((cnt_c + 1e-8) count_nonzero(cnt)).
 *   Under batch sharding pass the all-reduced counts: the loss is then this shard's share of the global loss.
 * bwd: dpred = *gloss * 2 (pred - truth) mask * scale_c   (gloss NULL = 1). */
size_t immtsf_masked_mse_workspace_bytes(int C);
int immtsf_masked_mse_partial(const float* pred, const float* truth, const float* mask, long rows, int T, int C,
                              float* err_cnt, int32_t* empty_flag, unsigned int* ticket, void* workspace,
                              size_t workspace_bytes, void* stream);
int immtsf_masked_mse_finalize(const float* err, const float* cnt, int C, float* loss, float* scale, void* stream);
int immtsf_masked_mse_bwd(const float* pred, const float* truth, const float* mask, long rows, int C,
                          const float* scale, const float* gloss, float* dpred, void* stream);

/* ---- next to the path: the text-embedding store.  The reference keeps a Python list of (rel_time,
 * embedding row) per record (lib/parse_datasets.py:132-147), filters it per chunk window with a list comprehension
 * (:204-209) and pads / stacks the selected rows per batch (multimodal_collate :786-819).  Here every embedding row of every
 * record is resident once: emb_all [sumN, d_m] (ld), rel_all [sumN], entity_offsets [E+1].
 * window_count: counts[i] = #{j in record ent[i] : st[i] <= (double)rel_all[j] < hist_end[i]}   (one warp per chunk)
 * exclusive_scan_i32: offsets[0] = 0, offsets[i+1] = counts[0] + ... + counts[i]; *overflow_flag = 1 if a sum exceeds int32
 * window_fill: chunk_rows[chunk_offsets[i] + k] = row of the k-th selected note of chunk i IN FILE ORDER (the order of the
 *   reference's comprehension), chunk_tau[...] = (float)((double)rel - st[i])
 * batch_gather: sample b of a batch is chunk chunk_ids[b]; its selected rows are copied to emb_flat[offsets[b] + k] and
 *   their tau to tau_flat (offsets [B+1]: exclusive scan of the chunk sizes, made by the caller); N_max >= every size. */
int immtsf_window_count(const float* rel_all, const int32_t* entity_offsets, const int32_t* ent, const double* st,
                        const double* hist_end, int n, int32_t* counts, void* stream);
int immtsf_exclusive_scan_i32(const int32_t* counts, int n, int32_t* offsets, int32_t* overflow_flag, void* stream);
int immtsf_window_fill(const float* rel_all, const int32_t* entity_offsets, const int32_t* ent, const double* st,
                       const double* hist_end, int n, const int32_t* chunk_offsets, int32_t* chunk_rows,
                       float* chunk_tau, void* stream);
int immtsf_batch_gather(const float* emb_all, int ld, int d_m, const int32_t* chunk_offsets, const int32_t* chunk_rows,
                        const float* chunk_tau, const int32_t* chunk_ids, const int32_t* offsets, int B, int N_max,
                        float* emb_flat, int ld_out, float* tau_flat, void* stream);

/* ---- data-parallel all-reduce of the gradient statistics, in the NVSwitch (csrc/nvls.cu).  The reference has no
 * distributed code; this is the one collective of the path (8e).  mc_base = multicast address of a symmetric
 * arena, peer_bases_dev = DEVICE array of world pointers (this process's mapping of every rank's replica); the range
 * [off_floats, off_floats + n_floats) is summed over the ranks in place on every rank (multimem.ld_reduce / multimem.st of
 * this rank's slice between two rank barriers).  The first immtsf_nvls_flag_bytes(world) bytes at flag_off_floats of every
 * replica are barrier flags, zero before the first call.  err_flag_dev is set to 1 if a rank never arrives (bounded spin). */
size_t immtsf_nvls_flag_bytes(int world);
int immtsf_nvls_allreduce_f32(void* mc_base, void* const* peer_bases_dev, size_t off_floats, size_t n_floats,
                              size_t flag_off_floats, int rank, int world, int* err_flag_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IMMTSF_H_ */
