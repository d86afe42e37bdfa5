/* sessrec_b200 — C ABI of the B200-native session-recommendation training path.
 *
 * Drop-in boundary for the per-batch hot path of SpaceLearner/SessionRec-pytorch
 * (`src/models/{srgnn,niser,msgifsr}.py` forward / `nll_loss` / autograd backward).  The reference has no
 * FFI of its own: its boundary is the `nn.Module.forward(batched_graph) -> (B, V) log-probs` contract
 * consumed by `src/utils/train.py:94-101`.  Each entry point below replaces the library kernels the
 * reference reaches for one stage of that forward/backward (file:line given per function).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`; the caller owns all memory;
 *   - floating tensors are fp32 row-major, indices are int32 (the Python boundary converts from int64);
 *   - `stream` is a `cudaStream_t` (passed as void*); calls only enqueue work, they never synchronise;
 *   - return value: SRK_OK or a negative error code, message via srk_last_error() (thread-local);
 *   - d (embedding dim) must be a multiple of 4 and <= 1024; GAT heads are fixed to 8 like the reference.
 */
#ifndef SESSREC_B200_H
#define SESSREC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRK_OK 0
#define SRK_ERR_INVALID (-1)
#define SRK_ERR_CUDA (-2)
#define SRK_ERR_UNSUPPORTED (-3)

#define SRK_HEADS 8
#define SRK_MAX_GAT_INST 8

/* row-normalisation modes */
#define SRK_NORM_NONE 0
#define SRK_NORM_NISER 1   /* x / (||x|| + 1e-12) then x / ||x||   (niser.py:134-135,141-142) */
#define SRK_NORM_L2 2      /* F.normalize: x / max(||x||, 1e-12)    (msgifsr.py:252-253)        */
#define SRK_NORM_EPS 3     /* x / (||x|| + 1e-12)                   (niser.py:147-151)          */

/* dropout sites (shared with oracle/models.py) */
#define SRK_SITE_EMBED 0x100
#define SRK_SITE_READOUT 0x200
#define SRK_SITE_GGNN 0x300
#define SRK_SITE_GAT_SRC 0x1000
#define SRK_SITE_GAT_DST 0x1001
#define SRK_SITE_GAT_ATTN 0x1002

typedef struct srk_dropout {
  float p;        /* 0 disables */
  uint32_t site;  /* SRK_SITE_* (+ offset) */
  uint64_t seed;
} srk_dropout;

const char* srk_last_error(void);
int srk_version(void);
long long srk_launch_count(void); /* kernels launched by this library so far (process-wide) */

/* ---- dense building block -------------------------------------------------------------------------
 * C[rc(m), n] (op)= alpha * sum_k A[ra(m)*sa_m + k*sa_k] * B[rb(k)*sb_k + n*sb_n] (+ bias[n]).
 * Replaces the cuBLAS sgemm behind every nn.Linear / `@` on the path (srgnn.py:42-45,80-81,143;
 * gatconv.py:273-274; msgifsr.py:139-140,271).  split_k <= 0 picks a split automatically. */
int srk_gemm(int M, int N, int K, const float* A, long long sa_m, long long sa_k, const float* B, long long sb_k,
             long long sb_n, float* C, long long ldc, const int* a_idx, const int* b_idx, const int* c_idx,
             const float* bias, float alpha, int accumulate, int split_k, void* stream);

/* tcgen05 tensor-core GEMM, fp32-faithful via a 3xTF32 split (A = Ahi + Alo, B = Bhi + Blo, halves produced by
 * srk_split_tf32 or by the producing kernels).  form 0: A[M,K] B[N,K] (Z = s E^T, srgnn.py:146); form 1: A[M,K] B[K,N]
 * (dS = dZ E); form 2: A[K,M] B[K,N] (dE = dZ^T s).  TMA (SWIZZLE_128B) + tcgen05.mma kind::tf32 + TMEM accumulators.
 * accumulate != 0 uses an atomicAdd epilogue (required for split_k > 1).  N <= 256 for forms 1 and 2. */
int srk_umma_gemm(int form, int M, int N, int K, const float* Ahi, const float* Alo, long long lda, const float* Bhi,
                  const float* Blo, long long ldb, float* C, long long ldc, float alpha, int accumulate, int split_k,
                  void* stream);
/* The same tensor-core GEMM from PLAIN fp32 operands (every nn.Linear / `@` of the session encoder once the node count
 * makes it worth a tensor-core launch: srgnn.py:42-45,80-81,143; msgifsr.py:139-140,271; gatconv.py:273-274): one launch
 * splits both operands into dense TF32 hi / lo copies inside `scratch` (srk_tc_gemm_scratch_floats floats, 16-byte aligned),
 * then srk_umma_gemm's kernel runs.  bias[N] (optional) is added once.  Any N (forms 1 / 2 are tiled over 256 columns);
 * split_k <= 0 picks a split that fills the SMs when accumulate != 0. */
long long srk_tc_gemm_scratch_floats(int form, int M, int N, int K);
int srk_tc_gemm(int form, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb, float* C,
                long long ldc, const float* bias, float alpha, int accumulate, int split_k, float* scratch, void* stream);
/* Persistent forward scoring kernel with a fused log-sum-exp epilogue: Z[M, ldz] = alpha * A B^T (A[M,K], B[N,K] as TF32
 * hi/lo pairs), lse[M] = row log-sum-exp, nll[M] = lse - Z[m, labels[m]] (labels / nll may be NULL).  One CTA per SM,
 * 128 x 256 tiles, double-buffered TMEM.  part = scratch of 4 * ceil(N/256) * M + M floats.  Replaces srk_umma_gemm(form 0)
 * + srk_ce_rows_fwd (srgnn.py:146-147 / msgifsr.py:308-309 + train.py:99). */
int srk_umma_score_fwd(int M, int N, int K, const float* Ahi, const float* Alo, long long lda, const float* Bhi,
                       const float* Blo, long long ldb, float* Z, long long ldz, float alpha, const int* labels, float* lse,
                       float* nll, float* part, void* stream);
/* hi = x with the 13 low mantissa bits cleared (TF32-exact), lo = x - hi (exact); [rows, cols] with pitches ldx / ldo. */
int srk_split_tf32(const float* X, long long ldx, int rows, int cols, float* hi, float* lo, long long ldo, void* stream);

/* ---- fused scoring + cross-entropy head ("flash CE", csrc/flash_ce.cu) ---------------------------------
 * logits = scale * shat Ehat^T, loss = mean(logsumexp - label logit) (srgnn.py:145-147, niser.py:149-156,
 * msgifsr.py:276-309 + utils/train.py:99) WITHOUT writing the (B, V) logits: tcgen05.mma kind::f16 on bf16 hi/lo
 * pairs (x = hi + lo, products as hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM; ~1e-5 relative).  16 <= d <= 128,
 * d % 16 == 0.  Operands are bf16 bit patterns (uint16_t) with row pitches lds / lde (multiples of 8). */
/* hi = bf16(x) (round to nearest even), lo = bf16(x - hi); [rows, cols] with pitches ldx / ldo. */
int srk_split_bf16(const float* X, long long ldx, int rows, int cols, uint16_t* hi, uint16_t* lo, long long ldo, void* stream);
/* 1 when the fused head is built for this embedding dim, else 0 */
int srk_flash_ce_supported(int d);
/* floats of scratch (`part`) srk_flash_ce_fwd needs */
long long srk_flash_ce_part_floats(int B, int V);
/* lse[B] = row log-sum-exp of the logits; nll[B] (optional) = lse - logit of labels[b] (0 where the label lies outside
 * [0, V): catalog sharding). */
int srk_flash_ce_fwd(int B, int V, int d, const uint16_t* Shi, const uint16_t* Slo, long long lds, const uint16_t* Ehi,
                     const uint16_t* Elo, long long lde, float scale, const int* labels, float* lse, float* nll, float* part,
                     void* stream);
/* Fused evaluation head (evaluate(), utils/train.py:36-55): ids (best first; ties: smaller id first) and optionally values of
 * the K <= 32 largest logits of every row, kept by the soft-max threads of the scoring kernel while the tiles go by and merged
 * per row by a second small launch - `logits.topk(k=cutoff)` (train.py:49) without the (B, V) matrix.  scratch:
 * srk_flash_ce_topk_scratch_floats floats. */
long long srk_flash_ce_topk_scratch_floats(int B, int V, int K);
int srk_flash_ce_topk(int B, int V, int d, const uint16_t* Shi, const uint16_t* Slo, long long lds, const uint16_t* Ehi,
                      const uint16_t* Elo, long long lde, float scale, int K, int* out_idx, float* out_val, float* scratch,
                      void* stream);
/* number of [V, d] partial buffers srk_flash_ce_bwd writes (= ceil(B / 128)) */
int srk_flash_ce_bwd_parts(int B);
/* Backward of the mean loss: recomputes every 128 x 128 logit tile, dZ = gout[0] * scale * (softmax - onehot) / B
 * (gout NULL = 1), and returns dS[B, d] = dZ Ehat (overwritten) and dEpart[parts][V, d]: the sum over `parts` is
 * dEhat = dZ^T shat (srk_catalog_prep_bwd / srk_sum_parts add them up; every element of dEpart is written). */
int srk_flash_ce_bwd(int B, int V, int d, const uint16_t* Shi, const uint16_t* Slo, long long lds, const uint16_t* Ehi,
                     const uint16_t* Elo, long long lde, float scale, const int* labels, const float* lse, const float* gout,
                     float* dS, float* dEpart, void* stream);
/* Debug aid: device buffer of 11 * 64 * 8 int64 receiving clock64() stamps of CTA 0's roles per tile (NULL = off). */
int srk_flash_ce_set_trace(long long* trace_dev);
/* out[n] (+)= sum_p parts[p * stride + i] */
int srk_sum_parts(const float* parts, long long stride, int nparts, long long n, float* out, int accumulate, void* stream);

/* ---- item-embedding gather / scatter-add (K1 / K8) -------------------------------------------------
 * Forward: X[i] = norm_mode(dropout(E[iid[i]])), i < P.  Replaces `self.embedding(iid)` + feat_drop +
 * normalisation (srgnn.py:133; niser.py:133-135,141-142; msgifsr.py:247-253).  rnorm[i] receives the L2
 * norm of the dropped row (needed by backward).  x_first (optional, NISER only) receives the
 * once-normalised rows that feed the GGNN layers (niser.py:136). */
int srk_embed_gather_fwd(const float* E, const int* iid, int P, int d, int norm_mode, const srk_dropout* drop,
                         float* X, float* rnorm, float* x_first, void* stream);
/* Backward: dE[u] += sum over the occurrences i of item u of d(dropped row i): `perm` lists the P gather positions
 * sorted by item id, uoff[U+1] delimits each distinct item, uid[U] is its id.  Load-balanced over chunks of 8
 * occurrences; only items cut by a chunk boundary use atomics.  Replaces
 * `embedding_dense_backward` (autograd of srgnn.py:133). dX_first may be NULL. */
int srk_embed_scatter_bwd(const float* E, const int* iid, const int* perm, const int* uoff, const int* uid, int U,
                          int P, int d, int norm_mode, const srk_dropout* drop, const float* rnorm, const float* dX,
                          const float* dX_first, float* dE, void* stream);

/* The same with a workspace of srk_embed_scatter_ws_floats(P, d) floats: runs of one item that a chunk boundary cuts are
 * written as partial sums and added up in chunk order by a second small launch - no atomics, bit-reproducible from run to
 * run (ws == NULL: atomicAdd for the cut runs, as srk_embed_scatter_bwd). */
long long srk_embed_scatter_ws_floats(int P, int d);
int srk_embed_scatter_bwd_ws(const float* E, const int* iid, const int* perm, const int* uoff, const int* uid, int U, int P,
                             int d, int norm_mode, const srk_dropout* drop, const float* rnorm, const float* dX,
                             const float* dX_first, float* dE, float* ws, void* stream);

/* ---- catalog pre-pass (K6a) -------------------------------------------------------------------------
 * mode SRK_NORM_L2 (MSGIFSR): renormalise IN PLACE every row of E whose norm exceeds max_norm (> 0) by
 * max_norm / (norm + 1e-7) (`nn.Embedding(max_norm=1)`, msgifsr.py:162,276), then Ehat = F.normalize(E)
 * (msgifsr.py:278-279).  mode SRK_NORM_EPS (NISER): Ehat = E / (||E|| + 1e-12) (niser.py:149-151).
 * enorm[V] receives the (post-renorm) row norms.  Ehat_hi / Ehat_lo (optional, both or neither): TF32 split of
 * Ehat for srk_umma_gemm.  Ebf_hi / Ebf_lo (optional, both or neither): bf16 hi/lo split for srk_flash_ce_*. */
int srk_catalog_prep_fwd(float* E, int V, int d, int norm_mode, float max_norm, float* Ehat, float* enorm,
                         float* Ehat_hi, float* Ehat_lo, uint16_t* Ebf_hi, uint16_t* Ebf_lo, void* stream);
/* dE += backward of the row normalisation; dEhat is given as `nparts` partial sums, [V, d] each, V * d floats apart
 * (nparts = 1: a plain gradient). */
int srk_catalog_prep_bwd(const float* E, const float* Ehat, const float* enorm, const float* dEhat, int nparts, int V, int d,
                         int norm_mode, float* dE, void* stream);
/* In-place max_norm renorm of the rows touched by a gather (msgifsr.py:247); duplicates are safe. */
int srk_renorm_rows(float* E, const int* uid, int U, int d, float max_norm, void* stream);

/* ---- row normalisation -------------------------------------------------------------------------------
 * Y = norm_mode(X) row-wise over [R, d] (ldx/ldy row strides); rnorm[R] keeps ||x||. */
int srk_rownorm_fwd(const float* X, long long ldx, int R, int d, int norm_mode, float* Y, long long ldy, float* rnorm,
                    void* stream);
/* srk_rownorm_fwd + srk_split_bf16 of the result in one launch: Ybf_hi / Ybf_lo [R, d] (dense) = bf16 hi / lo of Y. */
int srk_rownorm_split_fwd(const float* X, long long ldx, int R, int d, int norm_mode, float* Y, long long ldy, float* rnorm,
                          uint16_t* Ybf_hi, uint16_t* Ybf_lo, void* stream);
int srk_rownorm_bwd(const float* X, long long ldx, const float* Y, long long ldy, const float* rnorm, const float* dY,
                    long long lddy, int R, int d, int norm_mode, float* dX, long long lddx, int accumulate,
                    void* stream);

/* ---- SemanticExpander tail (msgifsr.py:32-45 with reducer 'mean', then F.normalize :252-253) -----------------------
 * out[n] = normalize(0.5 * mean_t X[n, t, :] + 0.5 * h[n, :]); X = dropped gather [N, k, d], h = final GRU state.
 * Backward writes dh = 0.5 dpre and dX[n, t, :] = 0.5 / k dpre (the GRU BPTT then accumulates into dX). */
int srk_expander_combine_fwd(const float* X, const float* h, int N, int k, int d, float* out, float* rnorm, void* stream);
int srk_expander_combine_bwd(const float* out, const float* rnorm, const float* dout, int N, int k, int d, float* dh,
                             float* dX, void* stream);

/* ---- elementwise helpers -------------------------------------------------------------------------------
 * Y = dropout(X) over n elements (flat index = element index); accumulate: Y += dropout_mask * X. */
int srk_dropout_apply(const float* X, float* Y, long long n, const srk_dropout* drop, int accumulate, void* stream);
/* Y = dropout(X + A) (Y may alias X or A). */
int srk_dropout_apply_add(const float* X, const float* A, float* Y, long long n, const srk_dropout* drop, void* stream);
/* Y = dropout(X) plus its TF32 hi / lo split (Yhi + Ylo == Y exactly; operands of srk_umma_gemm) in one launch. */
int srk_dropout_apply_split(const float* X, float* Y, float* Yhi, float* Ylo, long long n, const srk_dropout* drop,
                            void* stream);
int srk_fill(float* X, long long n, float value, void* stream);
int srk_gather_rows(const float* X, const int* idx, int R, int d, float* Y, long long ldy, void* stream);
int srk_scatter_add_rows(const float* X, long long ldx, const int* idx, int R, int d, float* Y, void* stream);
int srk_colsum(const float* X, long long ldx, int R, int d, float* out, int accumulate, void* stream);

/* ---- attention readout (K5) ------------------------------------------------------------------------------
 * Rows F[R, d] are grouped in B contiguous segments seg[B+1]; u = F Wu^T (+bu) [R, d] and v [B, d] are
 * computed by the caller with srk_gemm.  e_i = <we, sigmoid(u_i + v_b)>, alpha = segment softmax,
 * g_b = sum alpha_i F_i.  Writes sr_in[b] = [F[last_b] | g_b] (row stride 2d), e[R] and the per-segment
 * softmax statistics ms[B, 2] = (max, sum); with_last == 0 leaves the first half of sr_in to the caller.  Replaces srgnn.py:82-86 / msgifsr.py:141-146 (+ the concat
 * at srgnn.py:141-142, msgifsr.py:269-270). last[B] = row index of each segment's "last" node. */
int srk_readout_fwd(const float* F, const float* u, const float* v, const float* we, const int* seg,
                    const int* last, int B, int d, int with_last, float* e, float* ms, float* sr_in, void* stream);
/* Fused per-session tail of the readout, forward (csrc/readout_fused.cu): given u = F Wu^T (+bu) [R, d] and v [B, d]
 * (srk_gemm), e / alpha / g as srk_readout_fwd, sr_in = [F[last] | g], s = sr_in Wsr^T, shat = norm_mode(s)
 * (SRK_NORM_NONE / L2 / EPS), rn_s = ||s|| (optional), sbh / sbl = bf16 hi/lo of shat (optional pair): one launch instead
 * of srk_readout_fwd + srk_gemm + srk_rownorm_fwd + srk_split_bf16 (srgnn.py:82-91,141-143; msgifsr.py:141-146,269-273).
 * WsrT[2d, d] = Wsr^T (srk_transpose): the lanes read consecutive output columns. */
int srk_readout_tail_fwd(const float* F, const float* u, const float* v, const float* we, const float* WsrT, const int* seg,
                         const int* last, int B, int d, int norm_mode, float* e, float* ms, float* sr_in, float* s,
                         float* shat, float* rn_s, uint16_t* sbh, uint16_t* sbl, void* stream);
/* Y[cols, rows] = X[rows, cols]^T */
int srk_transpose(const float* X, int rows, int cols, float* Y, void* stream);
/* Fused per-session head of the readout backward: d shat [B, d] -> ds = normalise-backward (written out) -> d sr_in =
 * ds Wsr -> attention backward: u / v are overwritten with d u / d v, dwe[d] accumulated, dF[R, d] = alpha d g (+ d l on
 * the last row) - one launch instead of srk_rownorm_bwd + srk_gemm + srk_readout_bwd.  The caller finishes with the
 * node-parallel GEMMs (dF += du Wu, dF[last] += dv Wv, and the weight gradients). */
int srk_readout_head_bwd(const float* F, const float* we, const float* Wsr, const int* seg, const int* last, int B, int d,
                         int norm_mode, const float* s, const float* shat, const float* rn_s, const float* sr_in,
                         const float* e, const float* ms, const float* dshat, float* u, float* v, float* ds, float* dF,
                         float* dwe, void* stream);
/* Backward: given d sr_in [B, 2d], overwrites u with du and v with dv IN PLACE, writes dF[R, d] = alpha_i *
 * dg_b, adds into dwe[d]. The caller finishes with GEMMs (dF += du Wu, dWu += du^T F, ...). */
int srk_readout_bwd(const float* F, float* u, float* v, const float* we, const int* seg, const int* last,
                    const float* e, const float* ms, const float* sr_in, const float* dsr_in, int B, int d, int with_last,
                    float* dF, float* dwe, void* stream);

/* ---- scoring head + cross-entropy (K6 / K7 elementwise parts) ----------------------------------------------
 * Z[B, V] (row stride ldz >= V) holds scale * sr Ehat^T (srk_gemm or the tcgen05 kernel).  Row kernel: lse[b] = logsumexp(Z[b]),
 * nll[b] = lse[b] - Z[b, label[b]] (skipped when labels == NULL); write_logp != 0 rewrites Z in place to log-probabilities
 * (srgnn.py:146-147, niser.py:152-156, msgifsr.py:308-309,321; nll_loss at train.py:99). */
int srk_ce_rows_fwd(float* Z, long long ldz, const int* labels, int B, int V, int write_logp, float* lse, float* nll,
                    void* stream);
/* loss = mean(nll) (deterministic single-block reduction). */
int srk_mean(const float* x, int n, float* out, void* stream);
/* Fused-loss backward: Z <- dZ = gscale[0] * scale * (exp(Z - lse) - onehot) / B in place (Z = logits).  With Zlo != NULL
 * the TF32 split is written instead: Z <- hi(dZ), Zlo <- dZ - hi(dZ) (operands of srk_umma_gemm). */
int srk_ce_rows_bwd(float* Z, long long ldz, const int* labels, const float* lse, const float* gscale, float scale,
                    int B, int V, int z_is_logp, float* Zlo, void* stream);
/* Same, restricted to the catalog columns [col0, col0 + ncols): lets the caller run the backward of the head chunk by
 * chunk so that each chunk's dZ hi/lo pair is still L2-resident when the two GEMMs consume it. */
int srk_ce_rows_bwd_cols(float* Z, long long ldz, const int* labels, const float* lse, const float* gscale, float scale,
                         int B, int col0, int ncols, float* Zlo, void* stream);
/* Compat backward from an arbitrary upstream gradient G[B, V] of the log-probs LP:
 * dZ = scale * (G - exp(LP) * rowsum(G)), written into DZ (may alias G). */
int srk_logp_bwd(const float* LP, long long ldlp, const float* G, long long ldg, float scale, int B, int V, float* DZ,
                 long long lddz, float* DZlo, void* stream);

/* ---- REnorm head of MSGIFSR (`--extra`, msgifsr.py:281-305) ----
 * score[b, v] = phi[b, 0] * softmax_{v in session b}(Z) + phi[b, 1] * softmax_{v not in session b}(Z): replaces the O(B)
 * Python mask loop and the two masked (B, V) soft-maxes.  The session's items are its order-1 nodes iid[seg[b] .. seg[b+1])
 * (unique within the session).  srk_renorm_head_fwd rewrites the scaled logits Z[B, ldz] in place into log score, given
 * lphi[B, 2] = log phi; zin[N] is scratch (N = seg[B]).  srk_renorm_head_bwd: from an upstream gradient G[B, ldg] of the
 * log-probs LP, or (G == NULL) from the labels of the mean-NLL loss times gscale[0], writes
 * DZ = scale * d loss / d Z (DZ may alias LP or G; optional TF32 lo part in DZlo) and dlphi[B, 2] = dl_scale * d loss / d log phi
 * (dl_scale = 1 / scale when G already carries the logit scale); tmp[2 N] scratch.
 * srk_gate_fwd / srk_gate_bwd: phi = softmax(W2 relu(h)) of `sc_sr[0]` (msgifsr.py:206,283): H[B, d] holds h = W1 s + b1 on
 * entry and relu(h) on exit, W2[2, d]; the backward returns da[B, 2] (gradient at the two gate logits) and dH[B, d]. */
int srk_renorm_head_fwd(float* Z, long long ldz, int B, int V, const int* iid, const int* seg, const float* lphi, float* zin,
                        void* stream);
int srk_renorm_head_bwd(const float* LP, long long ldlp, const float* G, long long ldg, const int* labels, const float* gscale,
                        float scale, float dl_scale, int B, int V, const int* iid, const int* seg, const float* lphi, float* tmp,
                        float* DZ, long long lddz, float* DZlo, float* dlphi, void* stream);
int srk_gate_fwd(float* H, const float* W2, int B, int d, float* lphi, void* stream);
int srk_gate_bwd(const float* Hr, const float* W2, const float* lphi, const float* dlphi, int B, int d, float* da, float* dH,
                 void* stream);

/* ---- order-fusion head of MSGIFSR (msgifsr.py:311-315,321): out = log sum_k a_k softmax(Z_k), a = softmax(alpha) ----
 * Zall[K][B][ldz] (head_stride floats apart) are the per-order scaled logits, lse[K][B] their row log-sum-exps, alpha[K]
 * the un-normalised mixture logits on the DEVICE (softmax taken in-kernel, no host sync).  srk_mix_bwd rewrites Zall in place with d loss / d Z_k (optionally as a TF32 hi/lo pair)
 * from an upstream gradient G[B, ldg] of the log-probs, or (G == NULL) from the labels of the fused mean-NLL loss; rsum[K][B]
 * receives sum_v G q_k, from which d alpha follows. */
int srk_mix_logp_fwd(const float* Zall, long long head_stride, long long ldz, const float* lse, const float* alpha, int K,
                     int B, int V, float* out, long long ldo, void* stream);
int srk_mix_loss_fwd(const float* nll, const float* alpha, int K, int B, float* loss_out, void* stream);
int srk_mix_bwd(float* Zall, float* Zlo_all, long long head_stride, long long ldz, const float* lse, const float* alpha,
                int K, int B, int V, const float* G, long long ldg, const int* labels, const float* gscale, float scale,
                float* rsum, void* stream);

int srk_mix_alpha_bwd(const float* rsum, const float* alpha, int K, int B, float* dalpha, void* stream);

/* Fused evaluation head (next-row: evaluate(), utils/train.py:36-55): per row the ids (and optionally values) of the k
 * largest logits, sorted descending, ties by ascending id.  Replaces materialising log-probs + `logits.topk(k)`. */
int srk_topk_rows(const float* Z, long long ldz, int B, int V, int k, int* out_idx, float* out_val, void* stream);

/* ---- GGNN layer (K2 / K3) ---------------------------------------------------------------------------------
 * Weighted-mean aggregation over in-edges and out-edges: NN[v] = [ sum_in w x[u] / sum_in w | sum_out w x[t] /
 * sum_out w ] (row stride 2d; 0 where a node has no such edge).  CSR by destination (in_*) and by source
 * (out_*), per-edge weights indexed by edge id.  Replaces the DGL UDF degree-bucketing update_all on the graph
 * and its reverse (srgnn.py:21-29,37-41). */
int srk_ggnn_aggregate_fwd(const float* X, int N, int d, const int* in_ptr, const int* in_src, const int* in_eid,
                           const int* out_ptr, const int* out_dst, const int* out_eid, const float* w, float* NN,
                           float* wsum, void* stream);
/* dX[u] (+)= sum_out (w/W_in[v]) dNN1[v] + sum_in (w/W_out[t]) dNN2[t]. */
int srk_ggnn_aggregate_bwd(const float* dNN, int N, int d, const int* in_ptr, const int* in_src, const int* in_eid,
                           const int* out_ptr, const int* out_dst, const int* out_eid, const float* w,
                           const float* wsum, float* dX, int accumulate, void* stream);
/* GRUCell pointwise part (torch gate order r, z, n): gi[N, 3d], gh[N, 3d] (biases included), h[N, d] ->
 * hnew[N, d]; saves nothing (backward recomputes the gates from gi/gh). srgnn.py:45. */
int srk_gru_pointwise_fwd(const float* gi, const float* gh, const float* h, int N, int d, float* hnew, void* stream);
/* dgi, dgh overwrite gi, gh IN PLACE; dh (+)= direct path z * dhnew. */
int srk_gru_pointwise_bwd(float* gi, float* gh, const float* h, const float* dhnew, int N, int d, float* dh,
                          int accumulate, void* stream);

/* ---- GAT / MSHGNN layer (K4) ----------------------------------------------------------------------------------
 * One "instance" = one relation of the heterograph in one direction (conv1 on the graph, conv2 on its reverse)
 * feeding destination nodes of one node type. Zel[N_src, 8d+8] = x'_src [W ; wl]^T (columns 8d..8d+7 are the
 * per-head source scores el), er[N_dst, 8] = x'_dst wr^T. */
typedef struct srk_gat_inst {
  const int* in_ptr;    /* [N_dst+1] CSR by destination (for this direction) */
  const int* in_src;    /* [M] source node of each CSR slot */
  const int* in_eid;    /* [M] edge id of each CSR slot (indexes att / dedge) */
  const int* out_ptr;   /* [N_src+1] CSR by source */
  const int* out_dst;   /* [M] */
  const int* out_eid;   /* [M] */
  const float* Zel;     /* [N_src, 8d+8] */
  const float* er;      /* [N_dst, 8] */
  const float* bias;    /* [8d] */
  const float* xdst;    /* [N_dst, d] residual input (dropped destination copy) */
  float* att;           /* [M, 8] attention after softmax, before attn_drop */
  float* dedge;         /* [M, 8] backward: d(pre-LeakyReLU score) */
  float* der;           /* [N_dst, 8] backward */
  float* dZel;          /* [N_src, 8d+8] backward */
  int n_src, n_dst, n_edges;
  uint32_t attn_site;   /* dropout site of attn_drop for this instance */
} srk_gat_inst;

/* wl[h, :] = sum_j attn_l[h, j] W[h*d + j, :], wr likewise: Waug[8d+8, d] = [W ; wl], wr[8, d]. (gatconv.py:285-286
 * reassociated so that el/er come out of the same GEMM as the projection.) */
int srk_gat_prep(const float* W, const float* attn_l, const float* attn_r, int d, float* Waug, float* wr, void* stream);
/* the same in one launch together with the TF32 hi / lo split of W_aug (dense [8d + 8, d] each) that srk_umma_gemm reads */
int srk_gat_prep_split(const float* W, const float* attn_l, const float* attn_r, int d, float* Waug, float* wr, float* Whi,
                       float* Wlo, void* stream);
/* dW += dWaug[:8d] + attn_l (x) dwl + attn_r (x) dwr ; dattn_l[h, j] += <W[h*d+j], dwl[h]> ; same for r. */
int srk_gat_prep_bwd(const float* W, const float* attn_l, const float* attn_r, const float* dWaug, const float* dwr,
                     int d, float* dW, float* dattn_l, float* dattn_r, void* stream);
/* Per destination node of one type: sum over instances of (edge-softmax attention aggregation + residual +
 * bias), max over the 8 heads, + mean of the session's input rows (segmean[B, d], node2seg[N]); optionally
 * L2-normalise.  Hpre = un-normalised result (NULL allowed when normalize != 0 is the only consumer), amax[N, d]
 * = winning head.  Replaces gatconv.py:294-311 + msgifsr.py:74-90 (+ 260-263 when normalize). n_inst == 0:
 * H = segmean broadcast (msgifsr.py:78-83 fallback). */
int srk_gat_aggregate_fwd(const srk_gat_inst* inst_host, int n_inst, int N, int d, const float* segmean,
                          const int* node2seg, const srk_dropout* attn_drop, int normalize, float* H, float* rnorm,
                          uint8_t* amax, void* stream);
/* Backward part 1 (per destination node): dH -> (normalise bwd) -> dHpre[N, d]; per instance writes dedge, der. */
int srk_gat_aggregate_bwd_dst(const srk_gat_inst* inst_host, int n_inst, int N, int d, const srk_dropout* attn_drop,
                              int normalize, const float* H, const float* rnorm, const uint8_t* amax,
                              const float* dH, float* dHpre, void* stream);
/* Backward part 2 (per source node of one instance): dZel[u] = [ sum_out att' dO[v] | sum_out dedge ]. */
int srk_gat_aggregate_bwd_src(const srk_gat_inst* inst_host, int d, const srk_dropout* attn_drop, const float* dHpre,
                              const uint8_t* amax, void* stream);
/* Same, and additionally writes the TF32 hi / lo split of dZel ([n_src, 8d + 8] each) for srk_umma_gemm. */
int srk_gat_aggregate_bwd_src_split(const srk_gat_inst* inst_host, int d, const srk_dropout* attn_drop, const float* dHpre,
                                    const uint8_t* amax, float* dZel_hi, float* dZel_lo, void* stream);
/* dbias[h, j] += sum_v [amax[v, j] == h] dHpre[v, j]. */
int srk_gat_bias_bwd(const float* dHpre, const uint8_t* amax, int N, int d, float* dbias, void* stream);
/* Segment mean over contiguous segments and its backward (msgifsr.py:86-87). */
int srk_segmean_fwd(const float* X, const int* seg, int B, int d, float* mean, void* stream);
int srk_segmean_bwd(const float* dHpre, const int* seg, int B, int d, float* dX, int accumulate, void* stream);

/* ---- optimizer (next-row: utils/train.py:70-74, torch.optim.Adam with L2-in-grad) ------------------------------
 * Flat buffers; decay[i] per element group is given by seg_off[S+1] / seg_decay[S]. step is 1-based; grad_scale
 * multiplies the raw gradient first (1/world_size after a data-parallel all-reduce).  A segment with seg_decay < 0 is
 * INACTIVE and left untouched (parameter, moments): torch.optim.Adam skips parameters whose .grad is None, which is what
 * the reference's forward leaves on every parameter it never reaches (SRGNN's dead GGNN layers, MSGIFSR's lint / linq /
 * link / beta / unused GRU ...). */
int srk_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                  const long long* seg_off, const float* seg_decay, int n_seg, float lr, float beta1, float beta2,
                  float eps, int step, float grad_scale, void* stream);
/* The same update in two launches around one [rows, d] table that starts `tab` floats into the flat buffer and spans
 * tab_span floats (rows * d + alignment padding).  part 0: the table rows NOT listed in rows_sorted[n_rows] (ascending);
 * part 1: every other element (outside the table + the listed rows).  Both parts together update every element exactly
 * once with the arithmetic of srk_adam_step; part 0 can run as soon as the gradient of the un-gathered rows is final. */
int srk_adam_step_split(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                        const long long* seg_off, const float* seg_decay, int n_seg, long long tab, int rows, int d,
                        long long tab_span, const int* rows_sorted, int n_rows, int part, float lr, float beta1,
                        float beta2, float eps, int step, float grad_scale, void* stream);

/* ---- native training step (MSGIFSR order 1, extra=False): utils/train.py:95-101 around msgifsr.py:241-323 ----
 * zero_grad + forward + nll_loss + backward (+ Adam) enqueued by ONE host call from a caller-provided device
 * workspace (srk_msgifsr_workspace_bytes).  slot_off_host: float offsets of the parameters inside the flat buffers,
 * order documented in csrc/step.cu.  phase 0 = all, 1 = up to the gradients (data parallel: all-reduce, then) 2 = Adam;
 * phase 3 = all, with the data-parallel gradient all-reduce (srk_comm_allreduce over the flat gradient buffer) enqueued by
 * the step itself between the backward pass and Adam (needs srk_comm_init). */
/* ---- collectives of the multi-GPU path (csrc/comm.cu; the reference is single-device) ----
 * One process per GPU; a process-wide NCCL communicator owned by this library (libnccl.so.2 is resolved with dlopen at run
 * time, nccl_path may be NULL) so that a training step's exchanges are enqueued from inside the native step, between its
 * kernels and inside its CUDA-graph replay.  Bootstrap: rank 0 calls srk_comm_unique_id, the caller hands the 128 bytes to
 * every rank (parallel.init_comm broadcasts them over torch.distributed), every rank calls srk_comm_init with its CUDA device
 * current.  All buffers are fp32 device buffers; calls only enqueue on `stream`. */
int srk_comm_unique_id(char* id128_host, const char* nccl_path);
int srk_comm_init(const char* id128_host, int rank, int world, const char* nccl_path);
int srk_comm_destroy(void);
int srk_comm_world(void);
int srk_comm_rank(void);
int srk_comm_nccl_version(void);
/* in-place all-reduce of buf[n]: op 0 = sum, 1 = max, 2 = average */
int srk_comm_allreduce(float* buf, long long n, int op, void* stream);
/* recv[world * n] <- every rank's send[n], in rank order */
int srk_comm_allgather(const float* send, float* recv, long long n, void* stream);
/* [rows, d] table whose rows are sharded over the ranks (balanced contiguous split, first rows % world ranks hold one more):
 * every owner broadcasts its rows in place - one grouped NCCL call - so that all replicas hold all updated rows again */
int srk_comm_share_rows(float* table, int rows, int d, void* stream);
/* Catalog-sharded scoring head (BASELINE config 5): rank r scores catalog rows [lo, hi) only.  srk_shard_labels:
 * labels_local[b] = labels[b] - lo where owned, else -1 (srk_flash_ce_fwd / srk_umma_score_fwd then return the LOCAL
 * log-sum-exp and, where owned, the NLL).  srk_shard_lse_pack builds the payload of the ONE sum all-reduce - pack[0:B] =
 * exp(lse_local - shift), pack[B:2B] = label logit where owned else 0 - and srk_shard_lse_unpack turns the reduced payload
 * into the global log-sum-exp and NLL.  shift = NULL: the constant `bound` (cosine heads, |logit| <= scale); else the
 * per-session max of lse_local over the ranks (one extra max all-reduce: SRGNN's unbounded logits). */
int srk_shard_labels(const int* labels, int B, int lo, int hi, int* labels_local, void* stream);
int srk_shard_lse_pack(const float* lse_local, const float* nll_local, const int* labels_local, const float* shift,
                       float bound, int B, float* pack, void* stream);
int srk_shard_lse_unpack(const float* pack, const float* shift, float bound, int B, float* lse, float* nll, void* stream);

/* CUDA-graph replay of the native step's backward half: after two warm-up steps per model configuration its ~35
 * launches (7 streams) are captured once (programmatic-dependent-launch edges included); later steps only rewrite the
 * kernel-node parameters that changed and issue one cudaGraphLaunch.  srk_set_graph_mode: 0 never, 1 always, 2 auto (default;
 * SESSREC_GRAPH=0/1 overrides): data-parallel steps always replay (the ranks share the host CPU); a single-rank step times
 * steps 2 and 3 on the host and on the device and replays from step 4 on when the enqueue time is at least 0.85 of the device
 * time, i.e. when the host is the bottleneck on this machine.  The counters tell how many steps were replayed, how many update passes had to
 * fall back to plain launches and how many nodes were rewritten. */
int srk_set_graph_mode(int on);
/* Whole-step graph: forward, backward and optimizer of a single-rank step (phase 0) as ONE graph.  An update pass leaves a
 * kernel node alone when its launch configuration and arguments are byte-identical to what the node holds, so with batches
 * padded to a fixed shape at a fixed address (srk_batch_build_padded) only the nodes that carry the dropout seed / the Adam
 * step count are rewritten and the host cost of a step is the body's bookkeeping + one cudaGraphLaunch.  mode: 0 never,
 * 1 always, 2 auto = when the batch header says it is padded (default; SESSREC_GRAPH_WHOLE=0/1 overrides). */
int srk_set_graph_whole(int mode);
/* Test hook: the nth_update-th next replay finds a "different kernel sequence" half-way through its update pass and must
 * fall back to plain launches for the backward half (0 = off). */
int srk_graph_inject_mismatch(int nth_update);
long long srk_graph_launches(void);
long long srk_graph_fallbacks(void);
long long srk_graph_node_updates(void); /* kernel nodes re-parameterised by update passes (unchanged nodes are skipped) */
long long srk_msgifsr_workspace_bytes(int B, int N, int M, int V, int d, int L);
int srk_msgifsr_train_step(const int* batch_dev, const int* batch_hdr_host, float* params, float* grads,
                           const long long* slot_off_host, int V, int d, int L, float dropout_p, uint64_t seed,
                           int use_umma, void* workspace, long long workspace_bytes, const float* one_dev, float* loss_out,
                           int do_adam, float* exp_avg, float* exp_avg_sq, long long n_flat, const long long* seg_off_dev,
                           const float* seg_decay_dev, int n_seg, float lr, float beta1, float beta2, float eps,
                           int adam_step, float grad_scale, int phase, int head_chunks, void* stream);

/* ---- native training step (SRGNN / NISER+): utils/train.py:95-101 around srgnn.py:131-148 / niser.py:130-157 ----
 * Same contract as srk_msgifsr_train_step for a session-graph batch (kind 'session').  slot_off_host: offsets (floats) into
 * the flat parameter / gradient buffers of embedding.weight; per layer gru.weight_ih, gru.weight_hh, gru.bias_ih,
 * gru.bias_hh, W1.weight, W2.weight; readout.fc_u.weight, readout.fc_v.weight, readout.fc_v.bias, readout.fc_e.weight,
 * fc_sr.weight.  niser: double-normalised gather, normalised session / catalog rows, logits x scale (niser.py:134-156).
 * dead_layers: evaluate the GGNN layers whose output the reference discards (srgnn.py:135-142) on a side stream.
 * gseed_dev: device scalar the backward is seeded with (1, or B_local / B_global under data parallelism).  flags / phase /
 * the optimizer arguments as in srk_msgifsr_train_step. */
long long srk_srgnn_workspace_bytes(int B, int N, int M, int V, int d, int L);
int srk_srgnn_train_step(const int* batch_dev, const int* batch_hdr_host, float* params, float* grads,
                         const long long* slot_off_host, int V, int d, int L, int niser, float scale, int dead_layers,
                         float dropout_p, uint64_t seed, int flags, void* workspace, long long workspace_bytes,
                         const float* gseed_dev, float* loss_out, int do_adam, float* exp_avg, float* exp_avg_sq,
                         long long n_flat, const long long* seg_off_dev, const float* seg_decay_dev, int n_seg, float lr,
                         float beta1, float beta2, float eps, int adam_step, float grad_scale, int phase, void* stream);


/* ---- native training step, MSGIFSR of any order K (k-gram node types): SemanticExpander (msgifsr.py:32-45), L heterogeneous
 * MSHGNN layers over intra_k / inter relations (msgifsr.py:47-91), multi-order read-out (msgifsr.py:124-155), fused head,
 * backward, Adam; without --extra / --fusion.  Same contract as srk_msgifsr_train_step.  slot_off_host (n_slots entries):
 * [0] embeddings.weight; 1 + ((l*2 + conv)*(K+1) + e)*4 + {0: attn_l, 1: attn_r, 2: bias, 3: fc.weight} of layers.l.conv{conv+1}
 * .mods.{intra1..intraK (e = k-1), inter (e = K)}; then per k = 2..K expander.GRUs.{k-2}.{weight_ih_l0, weight_hh_l0, bias_ih_l0,
 * bias_hh_l0}; then readout.fc_u.0.weight, readout.fc_u.0.bias, readout.fc_v.0.weight, readout.fc_e.0.weight, fc_sr.0.weight. */
long long srk_msgifsr_k_workspace_bytes(const int* batch_hdr_host, int V, int d, int L);
int srk_msgifsr_k_train_step(const int* batch_dev, const int* batch_hdr_host, float* params, float* grads,
                             const long long* slot_off_host, int n_slots, int V, int d, int L, float dropout_p, uint64_t seed,
                             int flags, void* workspace, long long workspace_bytes, const float* gseed_dev, float* loss_out,
                             int do_adam, float* exp_avg, float* exp_avg_sq, long long n_flat, const long long* seg_off_dev,
                             const float* seg_decay_dev, int n_seg, float lr, float beta1, float beta2, float eps, int adam_step,
                             float grad_scale, int phase, void* stream);

/* ---- native batch builder (host; next-row: utils/data/collate.py:61-85,87-217,219-256) ---------------------------
 * Sessions are given as a flat item array + offsets. kind 0 = session graph (weights, self-loop rule), kind 1 =
 * ccs heterograph of the given order. Two calls: srk_batch_size() returns the number of int32 words needed,
 * srk_batch_build() fills `out_host` (layout documented in DESIGN.md / batch.py) . */
long long srk_batch_size(const int* items_host, const int* offs_host, int B, int kind, int order);
long long srk_batch_build(const int* items_host, const int* offs_host, const int* labels_host, int B, int kind,
                          int order, int* out_host, long long out_words);

#ifdef __cplusplus
}
#endif
#endif /* SESSREC_B200_H */
