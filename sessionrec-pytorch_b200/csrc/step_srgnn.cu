// Native training step of SRGNN / NISER+: zero_grad + forward + nll_loss + backward + Adam for one batch in ONE host call.
//
// The body of the reference's training loop (src/utils/train.py:95-101) around the model forward of
// src/models/srgnn.py:131-148 / src/models/niser.py:130-157, composed in C++ from the same kernels the Python module
// drives stage by stage (srgnn.py::SRGNN._fwd/_bwd of this package is the readable twin and the parity reference of this
// file).  All temporaries come from a caller-provided device workspace; nothing is allocated, nothing synchronises.
//
// What differs from the MSGIFSR step (csrc/step.cu): the GGNN layers are dead code in the reference's output
// (srgnn.py:135-142: evaluated, then discarded) - they are still evaluated here so that the device does the work the
// reference does, but on a low-priority side stream beside the critical path, since nothing waits for them; and every
// dense product whose node count pays for a tensor-core launch (GGNN gate projections N x 2d x 3d, read-out projections
// N x d x d and their gradients) runs on the tcgen05 3xTF32 GEMM (srk_tc_gemm) instead of the fp32 CUDA-core kernel.
#include "step_common.cuh"

namespace {

constexpr int TYPE_TAB = 16, REL_TAB = 80;

struct SBatch {
  int B, N, M, U;
  const int *labels, *iid, *seg, *last, *perm, *uoff, *uid;
  const int *in_ptr, *in_src, *in_eid, *out_ptr, *out_dst, *out_eid;
  const float* w;
};

int parse_session_batch(const int* dev, const int* hdr, SBatch& b) {
  SRK_REQUIRE(hdr[0] == 0x53524B31, "srgnn step: not a SessionBatch buffer");
  SRK_REQUIRE(hdr[2] == 0 && hdr[3] == 1, "srgnn step: needs a session-graph batch (kind 'session', order 1)");
  b.B = hdr[1];
  b.labels = dev + hdr[7];
  const int* t = hdr + TYPE_TAB;
  b.N = t[0]; b.U = t[8];
  b.iid = dev + t[1]; b.seg = dev + t[2]; b.last = dev + t[3];
  b.perm = dev + t[5]; b.uoff = dev + t[6]; b.uid = dev + t[7];
  const int* r = hdr + REL_TAB;
  b.M = r[2];
  b.in_ptr = dev + r[5]; b.in_src = dev + r[6]; b.in_eid = dev + r[7];
  b.out_ptr = dev + r[8]; b.out_dst = dev + r[9]; b.out_eid = dev + r[10];
  b.w = reinterpret_cast<const float*>(dev + r[11]);
  return SRK_OK;
}

// one scratch region per stream for the operand splits of srk_tc_gemm (calls on one stream are ordered)
struct TcScratch {
  float* p;
  long long floats;
};

long long tc_min_macs() {
  static long long v = -1;
  if (v < 0) {
    const char* e = getenv("SESSREC_TC_MIN_MACS");
    v = e ? atoll(e) : (1LL << 24);
    const char* off = getenv("SESSREC_TC_GEMM");
    if (off && off[0] == '0') v = (1LL << 62);
  }
  return v;
}

// TF32 hi / lo pair of an operand that already exists (written by the operand's producer or by one split launch)
struct Pre {
  const float *hi = nullptr, *lo = nullptr;
  long long ld = 0;
};

bool tc_shape_ok(int M, int N, int K) { return (long long)M * N * K >= tc_min_macs() && K >= 32 && N >= 16; }

// C (op)= op(A) op(B) (+ bias): tensor cores when the product is big enough and nothing is row-indexed, else CUDA cores.
// accumulate: the split-K that fills the SMs is picked (C must hold the addend, or zeros).
int mmp(cudaStream_t st, const TcScratch& sc, bool tc, int form, int M, int N, int K, const float* A, long long lda, Pre pa,
        const float* Bm, long long ldb, Pre pb, float* C, long long ldc, const float* bias, int accumulate) {
  if (M <= 0 || N <= 0) return SRK_OK;
  const bool ok = tc && tc_shape_ok(M, N, K) && lda % 4 == 0 && ldb % 4 == 0 &&
                  ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Bm)) & 15u) == 0 &&
                  ((pa.hi && pb.hi) || srk_tc_gemm_scratch_floats(form, M, N, K) <= sc.floats) && (form != 2 || accumulate);
  if (ok)
    return srk_tc_gemm_pre(form, M, N, K, A, lda, pa.hi, pa.lo, pa.ld, Bm, ldb, pb.hi, pb.lo, pb.ld, C, ldc, bias, 1.0f, accumulate,
                           accumulate ? 0 : 1, sc.p, st);
  if (form == 0) return srk_gemm(M, N, K, A, lda, 1, Bm, 1, ldb, C, ldc, nullptr, nullptr, nullptr, bias, 1.f, accumulate, 0, st);
  if (form == 1) return srk_gemm(M, N, K, A, lda, 1, Bm, ldb, 1, C, ldc, nullptr, nullptr, nullptr, bias, 1.f, accumulate, 0, st);
  return srk_gemm(M, N, K, A, 1, lda, Bm, ldb, 1, C, ldc, nullptr, nullptr, nullptr, bias, 1.f, accumulate, 0, st);
}
int mm(cudaStream_t st, const TcScratch& sc, bool tc, int form, int M, int N, int K, const float* A, long long lda, const float* Bm,
       long long ldb, float* C, long long ldc, const float* bias, int accumulate) {
  return mmp(st, sc, tc, form, M, N, K, A, lda, Pre{}, Bm, ldb, Pre{}, C, ldc, bias, accumulate);
}

long long scratch_floats(int B, int N, int d) {
  // the largest operand pair of the step: gi = hn[N, 2d] W_ih[3d, 2d]^T (and d sr_in / fc_sr at B rows)
  long long a = srk_tc_gemm_scratch_floats(0, N, 3 * d, 2 * d);
  long long b = srk_tc_gemm_scratch_floats(1, B, 2 * d, d);
  long long c = srk_tc_gemm_scratch_floats(2, d, 2 * d, B > N ? B : N);
  long long m = a > b ? a : b;
  return (m > c ? m : c) + 64;
}

}  // namespace

extern "C" long long srk_srgnn_workspace_bytes(int B, int N, int M, int V, int d, int L) {
  const long long ldz = (V + 3) / 4 * 4;
  long long fl = 0;
  fl += 4LL * V * d + V;                                     // Ehat, Ehi / Elo (or bf16 pair), enorm
  fl += 3LL * N * d + N;                                     // X, x_first, F, rn
  fl += (long long)(L > 0) * (13LL * N * d + 2LL * N);       // dead GGNN layer temporaries (reused across layers): ft, NN, wsum, hn, gi, gh, hnew x 2
  fl += 2LL * N * d + 6LL * B * d + N + 4LL * B;             // u, e, v, ms, sr_in, s, shat, rn_s
  fl += 2LL * B * ldz + 4LL * B * d + 2LL * B + 4LL * ((V + 255) / 256) * B + B + 64;      // Z, Zlo, sh, sl, lse, nll, partials
  fl += (long long)V * d + B * d + srk_flash_ce_part_floats(B, V) + (long long)srk_flash_ce_bwd_parts(B) * V * d + 256;   // flash head
  fl += 4LL * B * d + 3LL * N * d + (long long)(N + 4) * d;  // dshat, ds, dsr_in, dF, dX, scatter partials
  fl += 3 * scratch_floats(B, N, d);                         // tensor-core operand splits: one region per stream
  fl += 4LL * N * d + 8LL * B * d + 14LL * d * d + 8192 + (long long)V * d;     // TF32 pairs made once: F, du, sr_in, ds, read-out weights; zero pool
  return fl * 4 + fl + (1 << 20);                            // floats -> bytes with 25% head-room + alignment slack
}

// Parameter slots (offsets in floats into the flat parameter / gradient buffers), in this order:
//   [0] embedding.weight
//   per layer l (6 slots each, starting at 1 + 6*l): gru.weight_ih, gru.weight_hh, gru.bias_ih, gru.bias_hh, W1.weight, W2.weight
//   then: readout.fc_u.weight, readout.fc_v.weight, readout.fc_v.bias, readout.fc_e.weight, fc_sr.weight
static int srgnn_body(const int* batch_dev, const int* batch_hdr_host, float* params, float* grads,
                      const long long* slot_off_host, int V, int d, int L, int niser, float scale, int dead_layers,
                      float dropout_p, uint64_t seed, int flags, void* workspace, long long workspace_bytes,
                      const float* gseed_dev, float* loss_out, int do_adam, float* exp_avg, float* exp_avg_sq, long long n_flat,
                      const long long* seg_off_dev, const float* seg_decay_dev, int n_seg, float lr, float beta1, float beta2,
                      float eps, int adam_step, float grad_scale, int phase, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SBatch b;
  SRK_TRY(parse_session_batch(batch_dev, batch_hdr_host, b));
  SRK_REQUIRE(L >= 0 && L <= 8, "srgnn step: 0..8 layers");
  SRK_REQUIRE(d % 4 == 0, "srgnn step: d must be a multiple of 4");
  const int B = b.B, N = b.N;
  const long long ldz = (V + 3) / 4 * 4;
  const bool umma = (flags & 1) && d <= 256;
  const bool fused_lse = (flags & 2) != 0;
  const bool flash = umma && (flags & 4) != 0 && srk_flash_ce_supported(d);
  const bool drop = dropout_p > 0.f;
  const int emb_mode = niser ? SRK_NORM_NISER : SRK_NORM_NONE;
  Arena ar{reinterpret_cast<uint8_t*>(workspace), (size_t)workspace_bytes, 0, true};
  auto P = [&](int slot) { return params + slot_off_host[slot]; };
  auto G = [&](int slot) { return grads + slot_off_host[slot]; };
  const int s_ro = 1 + 6 * L;       // readout.fc_u.weight, fc_v.weight, fc_v.bias, fc_e.weight, fc_sr.weight
  float* E = P(0);
  auto dcfg = [&](uint32_t site) { srk_dropout c; c.p = dropout_p; c.site = site; c.seed = seed; return c; };

  SideStreams* ss = srk_side_streams();
  StageTimer tm(st);
  // s2: weight gradients; sdead: the dead GGNN layers (lowest priority but one); s4: catalog-wide bulk passes
  cudaStream_t s1 = ss ? ss->s[1] : st, s2 = ss ? ss->s[4] : st, sdead = ss ? ss->s[5] : st, s4 = ss ? ss->s[6] : st;
  auto order = [&](cudaStream_t from, cudaStream_t to) { return ss ? ss->order(from, to) : (int)SRK_OK; };
  const bool graph_planned = srk_get_launch_ctx() != nullptr;     // the backward half will be captured / replayed
  SRK_TRY(srk_step_begin());
  // dS of the head accumulates (TMA reduce-add / split-K): it is zeroed by the launch that zeroes the gradient buffer
  // ... together with the two session-level products that run split-K (fc_sr and its data gradient: 16 row tiles only)
  const size_t bd = (size_t)b.B * d;
  // Wide fused head (d > 128): every session tile ADDS its table gradient into one [V, d] buffer (coalesced reductions in L2)
  // instead of writing B / 128 partial tables that a later pass sums (278 MB written + read at cfg2).  SRGNN scores the table
  // itself, so the buffer is the table's gradient rows (zeroed by zero_grad); NISER needs dEhat apart for the row-normalisation
  // backward: a zeroed buffer from the pool.  SESSREC_FCE_DE_ATOMIC=0 keeps the partial tables.
  static const bool de_atomic_on = [] { const char* e = getenv("SESSREC_FCE_DE_ATOMIC"); return !(e && e[0] == '0'); }();
  const bool de_atomic = flash && de_atomic_on && srk_flash_ce_bwd_parts(b.B) > 1;      // one session tile: nothing to add up
  const size_t zextra = (de_atomic && niser) ? (size_t)V * d : 0;
  float* zpool = ar.f(4 * bd + zextra);
  float *dshat = zpool, *s = zpool + bd, *dsr_in = zpool + 2 * bd;
  SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
  SRK_TRY(order(st, s4));
  SRK_TRY(srk_zero2_async(grads, sizeof(float) * (size_t)n_flat, zpool, sizeof(float) * (4 * bd + zextra), s4));
  // Every tensor-core product of the live path reads operands whose TF32 hi / lo pair is made ONCE: by the operand's producer
  // (dropout pass, read-out backward) or by one split launch per tensor - not once per product.  The read-out weights are
  // split here, beside the gather.
  const bool tc_ro = umma && d % 4 == 0 && tc_shape_ok(N, d, d);
  Pre Wu, Wsr;
  if (tc_ro) {
    const long long o_u = slot_off_host[s_ro], o_s = slot_off_host[s_ro + 4];
    const long long span = o_s + 2LL * d * d - o_u;
    SRK_TRY(order(st, s1));
    if (o_s > o_u && span % 4 == 0 && span <= 6LL * d * d + 4096) {          // fc_u .. fc_sr contiguous: one pass
      float *wh = ar.f((size_t)span), *wl = ar.f((size_t)span);
      SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
      SRK_TRY(srk_split_tf32(params + o_u, span, 1, (int)span, wh, wl, span, s1));
      Wu.hi = wh; Wu.lo = wl;
      Wsr.hi = wh + (o_s - o_u); Wsr.lo = wl + (o_s - o_u);
    } else {
      float *uh = ar.f((size_t)d * d), *ul = ar.f((size_t)d * d), *sh_ = ar.f(2 * (size_t)d * d), *sl_ = ar.f(2 * (size_t)d * d);
      SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
      SRK_TRY(srk_split_tf32(P(s_ro), d, d, d, uh, ul, d, s1));
      SRK_TRY(srk_split_tf32(P(s_ro + 4), 2 * d, d, 2 * d, sh_, sl_, 2 * d, s1));
      Wu.hi = uh; Wu.lo = ul; Wsr.hi = sh_; Wsr.lo = sl_;
    }
    Wu.ld = d; Wsr.ld = 2 * d;
  }
  tm.mark("zero_grad");

  const long long scf = scratch_floats(B, N, d);
  TcScratch sc_main{ar.f((size_t)scf), scf}, sc_w{ar.f((size_t)scf), scf}, sc_dead{ar.f((size_t)scf), scf};
  SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");

  // ---- forward -------------------------------------------------------------------------------------------
  // catalog pre-pass on s4 beside the encoder: NISER normalises every row (niser.py:149-151), SRGNN scores the table as it
  // is (srgnn.py:145-146); either way the head wants its bf16 (fused head) or TF32 (materialised scores) hi / lo pair
  float *Ehat = E, *enorm = nullptr, *Ehi = nullptr, *Elo = nullptr;
  uint16_t *Ebh = nullptr, *Ebl = nullptr;
  if (niser) {
    Ehat = ar.f((size_t)V * d);
    enorm = ar.f(V);
  }
  if (flash) {
    Ebh = reinterpret_cast<uint16_t*>(ar.raw((size_t)V * d * 2));
    Ebl = reinterpret_cast<uint16_t*>(ar.raw((size_t)V * d * 2));
  } else if (umma) {
    Ehi = ar.f((size_t)V * d);
    Elo = ar.f((size_t)V * d);
  }
  SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
  if (niser) {
    SRK_TRY(srk_catalog_prep_fwd(E, V, d, SRK_NORM_EPS, 0.0f, Ehat, enorm, Ehi, Elo, Ebh, Ebl, s4));
  } else if (flash) {
    SRK_TRY(srk_split_bf16(E, d, V, d, Ebh, Ebl, d, s4));
  } else if (umma) {
    SRK_TRY(srk_split_tf32(E, d, V, d, Ehi, Elo, d, s4));
  }
  tm.mark("catalog_prep");
  const bool dead = dead_layers && L > 0;
  float *X = ar.f((size_t)N * d), *rn = niser ? ar.f(N) : nullptr;
  float* x_first = (niser && dead) ? ar.f((size_t)N * d) : nullptr;
  SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
  srk_dropout dc_e = dcfg(SRK_SITE_EMBED + 1);
  SRK_TRY(srk_embed_gather_fwd(E, b.iid, N, d, emb_mode, drop ? &dc_e : nullptr, X, rn, x_first, st));
  tm.mark("gather");
  if (dead) {
    // srgnn.py:135-137: out = layer(mg, out) for every layer, result unused.  Forward only, beside everything else.
    SRK_TRY(order(st, sdead));
    const float* h = niser ? x_first : X;
    float *ft = drop ? ar.f((size_t)N * d) : nullptr, *NN = ar.f(2 * (size_t)N * d), *wsum = ar.f(2 * (size_t)N);
    float *hn = ar.f(2 * (size_t)N * d), *gi = ar.f(3 * (size_t)N * d), *gh = ar.f(3 * (size_t)N * d);
    float* hout[2] = {ar.f((size_t)N * d), ar.f((size_t)N * d)};
    SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
    for (int l = 0; l < L; ++l) {
      const int base = 1 + 6 * l;
      const float* x = h;
      if (drop) {
        srk_dropout dc = dcfg(SRK_SITE_GGNN + l);
        SRK_TRY(srk_dropout_apply(h, ft, (long long)N * d, &dc, 0, sdead));
        x = ft;
      }
      SRK_TRY(srk_ggnn_aggregate_fwd(x, N, d, b.in_ptr, b.in_src, b.in_eid, b.out_ptr, b.out_dst, b.out_eid, b.w, NN, wsum, sdead));
      SRK_TRY(mm(sdead, sc_dead, umma, 0, N, d, d, NN, 2 * d, P(base + 4), d, hn, 2 * d, nullptr, 0));
      SRK_TRY(mm(sdead, sc_dead, umma, 0, N, d, d, NN + d, 2 * d, P(base + 5), d, hn + d, 2 * d, nullptr, 0));
      SRK_TRY(mm(sdead, sc_dead, umma, 0, N, 3 * d, 2 * d, hn, 2 * d, P(base), 2 * d, gi, 3 * d, P(base + 2), 0));
      SRK_TRY(mm(sdead, sc_dead, umma, 0, N, 3 * d, d, h, d, P(base + 1), d, gh, 3 * d, P(base + 3), 0));
      SRK_TRY(srk_gru_pointwise_fwd(gi, gh, h, N, d, hout[l & 1], sdead));
      h = hout[l & 1];
    }
  }
  // read-out (with its own feat_drop on top of the embedding dropout, srgnn.py:79)
  const float* F = X;
  srk_dropout dc_r = dcfg(SRK_SITE_READOUT);
  Pre Fp, SRp, DSp, DUp;
  if (tc_ro) {
    float *fh = ar.f((size_t)N * d), *fl = ar.f((size_t)N * d), *rh = ar.f(2 * bd), *rl = ar.f(2 * bd);
    float *dh = ar.f(bd), *dl = ar.f(bd), *uh = ar.f((size_t)N * d), *ul = ar.f((size_t)N * d);
    SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
    Fp.hi = fh; Fp.lo = fl; Fp.ld = d;
    SRp.hi = rh; SRp.lo = rl; SRp.ld = 2 * d;
    DSp.hi = dh; DSp.lo = dl; DSp.ld = d;
    DUp.hi = uh; DUp.lo = ul; DUp.ld = d;
  }
  if (drop) {
    float* Fd = ar.f((size_t)N * d);
    SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
    if (tc_ro) SRK_TRY(srk_dropout_apply_split(X, Fd, const_cast<float*>(Fp.hi), const_cast<float*>(Fp.lo), (long long)N * d, &dc_r, st));
    else SRK_TRY(srk_dropout_apply(X, Fd, (long long)N * d, &dc_r, 0, st));
    F = Fd;
  } else if (tc_ro) {
    SRK_TRY(srk_split_tf32(X, d, N, d, const_cast<float*>(Fp.hi), const_cast<float*>(Fp.lo), d, st));
  }
  float *u = ar.f((size_t)N * d), *v = ar.f((size_t)B * d), *e = ar.f(N), *ms = ar.f(2 * (size_t)B);
  float* sr_in = ar.f(2 * (size_t)B * d);
  float *shat = niser ? ar.f((size_t)B * d) : s, *rn_s = niser ? ar.f(B) : nullptr;
  SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
  uint16_t *Sbh = nullptr, *Sbl = nullptr;
  if (flash) {
    Sbh = reinterpret_cast<uint16_t*>(ar.raw((size_t)B * d * 2));
    Sbl = reinterpret_cast<uint16_t*>(ar.raw((size_t)B * d * 2));
    SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
  }
  if (tc_ro) SRK_TRY(order(s1, st));            // the weight pairs
  SRK_TRY(order(st, s1));
  SRK_TRY(mmp(st, sc_main, umma, 0, N, d, d, F, d, Fp, P(s_ro), d, Wu, u, d, nullptr, 0));
  SRK_TRY(srk_gemm(B, d, d, F, d, 1, P(s_ro + 1), 1, d, v, d, b.last, nullptr, nullptr, P(s_ro + 2), 1.f, 0, 0, s1));
  SRK_TRY(order(s1, st));
  SRK_TRY(srk_readout_fwd(F, u, v, P(s_ro + 3), b.seg, b.last, B, d, drop ? 0 : 1, e, ms, sr_in, st));
  if (drop) SRK_TRY(srk_gather_rows(X, b.last, B, d, sr_in, 2 * d, st));      // sr_l uses the once-dropped rows
  if (tc_ro) SRK_TRY(srk_split_tf32(sr_in, 2 * d, B, 2 * d, const_cast<float*>(SRp.hi), const_cast<float*>(SRp.lo), 2 * d, st));
  // s comes zeroed from the pool: split-K over the 2d inputs (M = B gives 16 row tiles only)
  SRK_TRY(mmp(st, sc_main, umma, 0, B, d, 2 * d, sr_in, 2 * d, SRp, P(s_ro + 4), 2 * d, Wsr, s, d, nullptr, 1));
  if (niser) {
    if (flash) SRK_TRY(srk_rownorm_split_fwd(s, d, B, d, SRK_NORM_EPS, shat, d, rn_s, Sbh, Sbl, st));
    else SRK_TRY(srk_rownorm_fwd(s, d, B, d, SRK_NORM_EPS, shat, d, rn_s, st));
  } else if (flash) {
    SRK_TRY(srk_split_bf16(s, d, B, d, Sbh, Sbl, d, st));
  }
  tm.mark("readout_fwd");
  // scoring head + CE (needs the catalog pass)
  SRK_TRY(order(s4, st));
  float *Z = flash ? nullptr : ar.f((size_t)B * ldz), *lse = ar.f(B), *nll = ar.f(B);
  float *sh = nullptr, *sl = nullptr;
  SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
  if (flash) {
    float* part = ar.f((size_t)srk_flash_ce_part_floats(B, V));
    SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
    SRK_TRY(srk_flash_ce_fwd(B, V, d, Sbh, Sbl, d, Ebh, Ebl, d, scale, b.labels, lse, nll, part, st));
  } else if (umma) {
    sh = ar.f((size_t)B * d); sl = ar.f((size_t)B * d);
    SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
    SRK_TRY(srk_split_tf32(shat, d, B, d, sh, sl, d, st));
    if (fused_lse) {
      float* part = ar.f(4 * (size_t)((V + 255) / 256) * B + B);
      SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
      SRK_TRY(srk_umma_score_fwd(B, V, d, sh, sl, d, Ehi, Elo, d, Z, ldz, scale, b.labels, lse, nll, part, st));
    } else {
      SRK_TRY(srk_umma_gemm(0, B, V, d, sh, sl, d, Ehi, Elo, d, Z, ldz, scale, 0, 1, st));
      SRK_TRY(srk_ce_rows_fwd(Z, ldz, b.labels, B, V, 0, lse, nll, st));
    }
  } else {
    SRK_TRY(srk_gemm(B, V, d, shat, d, 1, Ehat, 1, d, Z, ldz, nullptr, nullptr, nullptr, nullptr, scale, 0, 0, st));
    SRK_TRY(srk_ce_rows_fwd(Z, ldz, b.labels, B, V, 0, lse, nll, st));
  }
  tm.mark("score_fwd+lse");
  // a captured backward half may not wait on work enqueued outside the capture: the dead layers are joined here then
  if (dead && graph_planned) SRK_TRY(order(sdead, st));
  // ---- forward / backward boundary: every side stream the backward touches has been joined into `st` ----------------
  SRK_TRY(srk_step_boundary());
  SRK_TRY(order(st, s2));
  SRK_TRY(srk_mean(nll, B, loss_out, s2));

  // ---- backward ------------------------------------------------------------------------------------------
  const int de_parts = flash ? srk_flash_ce_bwd_parts(B) : 1;
  float* Zlo = (umma && !flash) ? ar.f((size_t)B * ldz) : nullptr;
  // SRGNN scores the table itself: without row normalisation the head's table gradient IS d E, accumulated straight
  // into the (zeroed) gradient buffer by the materialised paths; the fused head leaves per-tile partials to be summed
  float* dEhat = de_atomic ? (niser ? zpool + 4 * bd : G(0)) : ((niser || flash) ? ar.f((size_t)de_parts * V * d) : G(0));
  SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
  if (flash) {
    SRK_TRY(srk_flash_ce_bwd_ex(B, V, d, Sbh, Sbl, d, Ebh, Ebl, d, scale, b.labels, lse, gseed_dev, dshat, dEhat, de_atomic ? 3 : 1, st));
  } else {
    SRK_TRY(srk_ce_rows_bwd(Z, ldz, b.labels, lse, gseed_dev, scale, B, V, 0, Zlo, st));
    if (umma) {
      int split = 148 / ((B + 127) / 128);
      const int nkb = (V + 31) / 32;
      if (split < 1) split = 1;
      if (split > nkb) split = nkb;
      SRK_TRY(srk_umma_gemm(1, B, d, V, Z, Zlo, ldz, Ehi, Elo, d, dshat, d, 1.0f, 1, split, st));
      SRK_TRY(srk_umma_gemm(2, V, d, B, Z, Zlo, ldz, sh, sl, d, dEhat, d, 1.0f, niser ? 0 : 1, 1, st));
    } else {
      if (niser) SRK_TRY(srk_zero_async(dEhat, sizeof(float) * (size_t)V * d, st));
      SRK_TRY(srk_gemm(B, d, V, Z, ldz, 1, Ehat, d, 1, dshat, d, nullptr, nullptr, nullptr, nullptr, 1.f, 1, 0, st));
      SRK_TRY(srk_gemm(V, d, B, Z, 1, ldz, shat, d, 1, dEhat, d, nullptr, nullptr, nullptr, nullptr, 1.f, 1, 0, st));
    }
  }
  tm.mark("ce_bwd+dS+dE");
  // the catalog-wide part of the table gradient stays on s4 beside the encoder backward; it only has to finish before
  // the scatter-add touches the same rows
  SRK_TRY(order(st, s4));
  if (niser) SRK_TRY(srk_catalog_prep_bwd(E, Ehat, enorm, dEhat, de_atomic ? 1 : de_parts, V, d, SRK_NORM_EPS, G(0), s4));
  else if (flash && !de_atomic) SRK_TRY(srk_sum_parts(dEhat, (long long)V * d, de_parts, (long long)V * d, G(0), 1, s4));
  const bool live = srk_launch_mode() == SRK_LAUNCH_DIRECT || srk_launch_mode() == SRK_LAUNCH_CAPTURE;
  if (ss && live) SRK_CUDA(cudaEventRecord(ss->ev_cat, s4));
  // Adam in two parts (see csrc/step.cu): the table rows this batch did not gather are final now
  const bool split_adam = ss != nullptr && phase == 0 && do_adam;
  const long long tab = slot_off_host[0];
  const long long tab_span = ((long long)V * d + 63) / 64 * 64;
  if (split_adam)
    SRK_TRY(srk_adam_step_split(params, grads, exp_avg, exp_avg_sq, n_flat, seg_off_dev, seg_decay_dev, n_seg, tab, V, d,
                                tab_span, b.uid, b.U, 0, lr, beta1, beta2, eps, adam_step, grad_scale, s4));
  tm.mark("catalog_bwd");
  float* ds = dshat;
  if (niser) {
    ds = ar.f((size_t)B * d);
    SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
    SRK_TRY(srk_rownorm_bwd(s, d, shat, d, rn_s, dshat, d, B, d, SRK_NORM_EPS, ds, d, 0, st));
  }
  float* dF = ar.f((size_t)N * d);
  SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
  if (tc_ro) SRK_TRY(srk_split_tf32(ds, d, B, d, const_cast<float*>(DSp.hi), const_cast<float*>(DSp.lo), d, st));
  // data gradients on the main stream, weight gradients (into the flat gradient buffer) on s2
  SRK_TRY(order(st, s2));
  SRK_TRY(mmp(st, sc_main, umma, 1, B, 2 * d, d, ds, d, DSp, P(s_ro + 4), 2 * d, Wsr, dsr_in, 2 * d, nullptr, 1));     // zeroed: split-K
  SRK_TRY(mmp(s2, sc_w, umma, 2, d, 2 * d, B, ds, d, DSp, sr_in, 2 * d, SRp, G(s_ro + 4), 2 * d, nullptr, 1));
  SRK_TRY(srk_readout_bwd_split(F, u, v, P(s_ro + 3), b.seg, b.last, e, ms, sr_in, dsr_in, B, d, drop ? 0 : 1, dF, G(s_ro + 3),
                                const_cast<float*>(DUp.hi), const_cast<float*>(DUp.lo), st));
  SRK_TRY(order(st, s2));
  SRK_TRY(mmp(st, sc_main, umma, 1, N, d, d, u, d, DUp, P(s_ro), d, Wu, dF, d, nullptr, 1));                     // u holds du
  SRK_TRY(mmp(s2, sc_w, umma, 2, d, d, N, u, d, DUp, F, d, Fp, G(s_ro), d, nullptr, 1));
  SRK_TRY(srk_gemm(B, d, d, v, d, 1, P(s_ro + 1), d, 1, dF, d, nullptr, nullptr, b.last, nullptr, 1.f, 1, 0, st));    // v holds dv
  SRK_TRY(srk_gemm(d, d, B, v, 1, d, F, d, 1, G(s_ro + 1), d, nullptr, b.last, nullptr, nullptr, 1.f, 1, 0, s2));
  SRK_TRY(srk_colsum(v, d, B, d, G(s_ro + 2), 1, s2));
  const float* dX = dF;
  if (drop) {
    float* dXd = ar.f((size_t)N * d);
    SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
    SRK_TRY(srk_dropout_apply(dF, dXd, (long long)N * d, &dc_r, 0, st));
    SRK_TRY(srk_scatter_add_rows(dsr_in, 2 * d, b.last, B, d, dXd, st));
    dX = dXd;
  }
  tm.mark("readout_bwd");
  if (ss && live) SRK_CUDA(cudaStreamWaitEvent(st, ss->ev_cat, 0));
  // The step is not bit-reproducible from run to run as a whole (split-K weight-gradient GEMMs, the d w_e reduction), so it
  // takes the one-launch scatter-add by default (atomicAdd on runs cut by a chunk boundary; measured 0.402 vs 0.418 ms per
  // step at cfg1); SESSREC_DETERMINISTIC_SCATTER=1 selects the two-pass variant without atomics.
  static const bool scatter_det = getenv("SESSREC_DETERMINISTIC_SCATTER") != nullptr;
  float* sws = !scatter_det ? nullptr : ar.f((size_t)srk_embed_scatter_ws_floats(N, d));  
  SRK_REQUIRE(ar.ok, "srgnn step: workspace too small");
  SRK_TRY(srk_embed_scatter_bwd_ws(E, b.iid, b.perm, b.uoff, b.uid, b.U, N, d, emb_mode, drop ? &dc_e : nullptr, rn, dX, nullptr, G(0),
                                   sws, st));
  SRK_TRY(order(s2, st));
  tm.mark("scatter");
  if (phase == 3) {
    // data parallel: the flat gradient buffer (every rank seeded its backward with B_local / B_global) is summed over the
    // ranks right here, behind the last gradient kernel (csrc/comm.cu), then Adam; behind the graph when one is replayed
    SRK_REQUIRE(srk_comm_world() > 1, "srgnn step: phase 3 needs a communicator (srk_comm_init)");
    SRK_TRY(order(s4, st));
    auto tail = [=]() -> int {
      SRK_TRY(srk_comm_allreduce(grads, n_flat, 0, st));
      if (do_adam)
        SRK_TRY(srk_adam_step(params, grads, exp_avg, exp_avg_sq, n_flat, seg_off_dev, seg_decay_dev, n_seg, lr, beta1, beta2, eps,
                              adam_step, grad_scale, st));
      return SRK_OK;
    };
    if (SrkLaunchCtx* lc = srk_get_launch_ctx()) lc->tail = tail;
    else SRK_TRY(tail());
  } else if (split_adam) {
    SRK_TRY(srk_adam_step_split(params, grads, exp_avg, exp_avg_sq, n_flat, seg_off_dev, seg_decay_dev, n_seg, tab, V, d,
                                tab_span, b.uid, b.U, 1, lr, beta1, beta2, eps, adam_step, grad_scale, st));
  } else if (phase == 0 && do_adam) {
    SRK_TRY(srk_adam_step(params, grads, exp_avg, exp_avg_sq, n_flat, seg_off_dev, seg_decay_dev, n_seg, lr, beta1, beta2, eps,
                          adam_step, grad_scale, st));
  }
  tm.mark("adam");
  SRK_TRY(order(s4, st));
  if (dead && !graph_planned) SRK_TRY(order(sdead, st));      // the workspace is reused by the next step
  tm.report();
  return SRK_OK;
}

// phase: 0 = everything; 1 = zero_grad + forward + backward only (the caller all-reduces the gradients); 2 = Adam only;
// 3 = everything with the data-parallel gradient all-reduce enqueued by the step itself (srk_comm_allreduce).
// flags: bit 0 tensor cores, bit 1 fused-LSE forward scoring kernel, bit 2 fused scoring + CE head (flash CE).
extern "C" int srk_srgnn_train_step(const int* batch_dev, const int* batch_hdr_host, float* params, float* grads,
                                    const long long* slot_off_host, int V, int d, int L, int niser, float scale,
                                    int dead_layers, float dropout_p, uint64_t seed, int flags, void* workspace,
                                    long long workspace_bytes, const float* gseed_dev, float* loss_out, int do_adam,
                                    float* exp_avg, float* exp_avg_sq, long long n_flat, const long long* seg_off_dev,
                                    const float* seg_decay_dev, int n_seg, float lr, float beta1, float beta2, float eps,
                                    int adam_step, float grad_scale, int phase, void* stream) {
  cudaStream_t caller = (cudaStream_t)stream;
  if (phase == 2) {
    return srk_adam_step(params, grads, exp_avg, exp_avg_sq, n_flat, seg_off_dev, seg_decay_dev, n_seg, lr, beta1, beta2,
                         eps, adam_step, grad_scale, caller);
  }
  int dev = 0;
  SRK_CUDA(cudaGetDevice(&dev));
  const unsigned long long key = (1ull << 63) | ((unsigned long long)(batch_hdr_host[1] & 0xFFFFF) << 43) |
                                 ((unsigned long long)(d & 0x3FF) << 33) | ((unsigned long long)(V & 0x3FFFF) << 15) |
                                 ((unsigned long long)(dev & 15) << 11) | ((unsigned long long)(L & 15) << 7) |
                                 ((unsigned long long)(flags & 7) << 4) | ((unsigned long long)(dropout_p > 0.f) << 3) |
                                 ((unsigned long long)(niser != 0) << 2) | ((unsigned long long)(dead_layers != 0) << 1) |
                                 (unsigned long long)(do_adam != 0 && phase == 0) | ((unsigned long long)(phase == 3) << 62);
  const bool whole = srk_step_want_whole(phase, batch_hdr_host[11]);
  return srk_step_driver(caller, key ^ ((unsigned long long)whole << 61), whole ? 1 : srk_step_want_graph(phase), [&](void* run) {
    return srgnn_body(batch_dev, batch_hdr_host, params, grads, slot_off_host, V, d, L, niser, scale, dead_layers, dropout_p, seed,
                      flags, workspace, workspace_bytes, gseed_dev, loss_out, do_adam, exp_avg, exp_avg_sq, n_flat, seg_off_dev,
                      seg_decay_dev, n_seg, lr, beta1, beta2, eps, adam_step, grad_scale, phase, run);
  }, whole);
}
