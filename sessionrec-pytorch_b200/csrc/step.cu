// Native training step: zero_grad + forward + nll_loss + backward + Adam for one batch in ONE host call.
//
// This is the body of the reference's training loop (src/utils/train.py:95-101) with the model forward of
// src/models/msgifsr.py:241-323 (order 1, extra=False) composed in C++ from the same kernels the Python modules
// drive stage by stage (msgifsr.py::MSGIFSR._fwd/_bwd is the readable twin and the parity reference of this file).
// All temporaries come from a caller-provided device workspace through a bump allocator; nothing is allocated,
// nothing synchronises; ~60 kernel launches are enqueued back to back on the given stream.
#include <chrono>
#include <map>
#include <mutex>

#include "step_common.cuh"

SideStreams* srk_side_streams() {
  static SideStreams per_dev[16];
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("SESSREC_STREAMS");
    enabled = !(e && e[0] == '0');
  }
  if (!enabled) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  SideStreams* ss = &per_dev[dev];
  if (!ss->ok && ss->init() != SRK_OK) return nullptr;
  return ss;
}


namespace {

constexpr int H = SRK_HEADS;

// batch buffer layout (csrc/batch_builder.cu)
constexpr int TYPE_TAB = 16, REL_TAB = 80;
struct BatchView {
  int B, N, M, U, P;
  const int *labels, *iid, *seg, *last, *node2seg, *perm, *uoff, *uid;
  const int *in_ptr, *in_src, *in_eid, *out_ptr, *out_dst, *out_eid;
};

int parse_batch(const int* dev, const int* hdr, BatchView& b) {
  SRK_REQUIRE(hdr[0] == 0x53524B31, "step: not a SessionBatch buffer");
  SRK_REQUIRE(hdr[2] == 1 && hdr[3] == 1, "step: MSGIFSR step needs a ccs batch of order 1");
  b.B = hdr[1];
  b.labels = dev + hdr[7];
  const int* t = hdr + TYPE_TAB;
  b.N = t[0]; b.U = t[8]; b.P = b.N;
  b.iid = dev + t[1]; b.seg = dev + t[2]; b.last = dev + t[3]; b.node2seg = dev + t[4];
  b.perm = dev + t[5]; b.uoff = dev + t[6]; b.uid = dev + t[7];
  const int* r = hdr + REL_TAB;
  b.M = r[2];
  b.in_ptr = dev + r[5]; b.in_src = dev + r[6]; b.in_eid = dev + r[7];
  b.out_ptr = dev + r[8]; b.out_dst = dev + r[9]; b.out_eid = dev + r[10];
  return SRK_OK;
}

int gemm(cudaStream_t st, int M, int N, int K, const float* A, long long sa_m, long long sa_k, const float* Bm, long long sb_k,
         long long sb_n, float* C, long long ldc, const int* a_idx = nullptr, const int* b_idx = nullptr,
         const int* c_idx = nullptr, const float* bias = nullptr, float alpha = 1.f, int accumulate = 0) {
  return srk_gemm(M, N, K, A, sa_m, sa_k, Bm, sb_k, sb_n, C, ldc, a_idx, b_idx, c_idx, bias, alpha, accumulate, 0, st);
}
// C[M,N] = X[M,K] W[N,K]^T
int linear_nt(cudaStream_t st, int M, int N, int K, const float* X, long long lda, const float* W, float* C, long long ldc,
              const int* a_idx = nullptr, const float* bias = nullptr) {
  return gemm(st, M, N, K, X, lda, 1, W, 1, K, C, ldc, a_idx, nullptr, nullptr, bias);
}
// C[M,N] (+)= A[M,K] Bm[K,N]
int mm_nn(cudaStream_t st, int M, int N, int K, const float* A, long long lda, const float* Bm, long long ldb, float* C,
          long long ldc, int accumulate, const int* c_idx = nullptr) {
  return gemm(st, M, N, K, A, lda, 1, Bm, ldb, 1, C, ldc, nullptr, nullptr, c_idx, nullptr, 1.f, accumulate);
}
// C[M,N] += A[K,M]^T Bm[K,N]
int mm_tn(cudaStream_t st, int M, int N, int K, const float* A, long long lda, const float* Bm, long long ldb, float* C,
          long long ldc, const int* b_idx = nullptr) {
  return gemm(st, M, N, K, A, 1, lda, Bm, ldb, 1, C, ldc, nullptr, b_idx, nullptr, nullptr, 1.f, 1);
}

struct InstRec {
  srk_gat_inst gi;
  const float *W, *al, *ar;
  float *gW, *gal, *gar, *gbias;
  float *Waug, *wr, *xs, *xd;
  float *Wh, *Wl, *xh, *xl;    // TF32 hi / lo of W_aug and of the source copy x_s (tensor-core projections)
  srk_dropout dcs, dcd;
  bool drop;
};

struct LayerRec {
  const float* in;          // layer input [N, d]
  float *Hout, *rn;
  uint8_t* amax;
  int n_inst;
  InstRec inst[2];
  int normalize;
};

}  // namespace

extern "C" long long srk_msgifsr_workspace_bytes(int B, int N, int M, int V, int d, int L) {
  const long long ldz = (V + 3) / 4 * 4;
  const long long ldzel = (long long)H * d + H;
  long long fl = 0;
  fl += 4LL * V * d + V;                                   // Ehat, Ehi, Elo, dEhat, enorm
  fl += (long long)N * d + N;                              // X, rnX
  fl += (long long)L * (2 * ((ldzel + H) * d + 2LL * N * d + N * ldzel + N * H + (long long)(M + 1) * H) + B * d + N * d + N + N * d / 4 + 64);
  fl += 2LL * N * d + 3LL * B * d + N + 2LL * B + 4LL * B * d + B;    // u, v, e, ms, sr_in, s, shat, rn_s
  fl += 2LL * B * ldz + 4LL * B * d + 2LL * B + 64 + 4LL * ((V + 255) / 256) * B + B;        // Z, Zlo, sh, sl, dshat, ds, lse, nll
  fl += 2LL * B * d + (long long)N * d + 2LL * d * d + (long long)(N + 4) * d;      // dsr_in, dF, W_sr^T, scatter partials
  // flash CE head: bf16 hi/lo of Ehat and shat, soft-max partials, one [V, d] dE partial per 128-session tile
  fl += (long long)V * d + B * d + srk_flash_ce_part_floats(B, V) + (long long)srk_flash_ce_bwd_parts(B) * V * d + 256;
  // backward per layer (reused across layers): dHpre, dfeat, per inst dedge, der, dZel, dWaug, dwr, tmp, tmp2
  fl += 8LL * N * d + 2 * ((long long)(M + 1) * H + N * H + N * ldzel + (ldzel + H) * d + 2LL * N * d);
  fl += (long long)L * 2 * (2 * ldzel * d + 2LL * N * d) + 2 * (2LL * N * ldzel);     // opt-in tensor-core projections: hi / lo copies
  fl += (long long)L * 2 * (ldzel * d + (long long)H * d + (long long)N * d + 192) + 2LL * B * d + 256 + (long long)V * d;      // zeroed pool
  return fl * 5 + (1 << 20);                               // floats -> bytes with 25% head-room + alignment slack
}

// Parameter slots (offsets in floats into the flat parameter / gradient buffers), in this order:
//   [0] embeddings.weight
//   per layer l (8 slots each, starting at 1 + 8*l): conv1.intra1.{attn_l, attn_r, bias, fc.weight}, conv2.intra1.{...}
//   then: readout.fc_u.0.weight, readout.fc_u.0.bias, readout.fc_v.0.weight, readout.fc_e.0.weight, fc_sr.0.weight
// The step as a sequence of srk_launch() calls.  Re-runnable: depending on the launch context of the calling thread it
// launches, is captured into a graph, or only re-parameterises the nodes of a captured graph (common.cuh).
static int step_body(const int* batch_dev, const int* batch_hdr_host, float* params, float* grads,
                     const long long* slot_off_host, int V, int d, int L, float dropout_p, uint64_t seed, int use_umma,
                     void* workspace, long long workspace_bytes, const float* one_dev, float* loss_out, int do_adam,
                     float* exp_avg, float* exp_avg_sq, long long n_flat, const long long* seg_off_dev,
                     const float* seg_decay_dev, int n_seg, float lr, float beta1, float beta2, float eps, int adam_step,
                     float grad_scale, int phase, int head_chunks, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BatchView b;
  SRK_TRY(parse_batch(batch_dev, batch_hdr_host, b));
  SRK_REQUIRE(L >= 1 && L <= 8, "step: 1..8 layers");
  const int B = b.B, N = b.N, M = b.M;
  const int ldzel = H * d + H;
  const bool umma = (use_umma & 1) && d <= 256;
  const bool fused_lse = (use_umma & 2) != 0;      // bit 1: persistent forward kernel with the fused LSE epilogue
  // bit 2: fused scoring + CE head (csrc/flash_ce.cu): no (B, V) logits in memory, bf16 x 3 tensor-core products
  const bool flash = umma && (use_umma & 4) != 0 && srk_flash_ce_supported(d);
  const bool drop = dropout_p > 0.f;
  // bit 3: catalog-sharded head (BASELINE config 5): this rank scores rows [lo, hi) of the table only; the session encoder
  // is replicated on the whole batch.  Exchanges, all enqueued here through csrc/comm.cu: ONE sum all-reduce of [2, B]
  // soft-max statistics (|logit| <= 12: constant shift), one of dS [B, d], and at the end of the step every owner
  // broadcasts its updated rows (the table gradient is never all-reduced: rank-local Adam on the owned rows).
  const int world = srk_comm_world(), rank = srk_comm_rank();
  const bool shard = (use_umma & 8) != 0 && world > 1;
  int lo = 0, hi = V;
  if (shard) {
    const int base = V / world, rem = V % world;
    lo = rank * base + (rank < rem ? rank : rem);
    hi = lo + base + (rank < rem ? 1 : 0);
  }
  const int Vl = hi - lo;              // catalog rows this rank scores
  const long long ldz = (Vl + 3) / 4 * 4;
  const bool dp_inside = phase == 3;   // data parallel: the gradient all-reduce is enqueued by this step
  SRK_REQUIRE(!dp_inside || world > 1, "step: phase 3 needs a communicator (srk_comm_init)");
  Arena ar{reinterpret_cast<uint8_t*>(workspace), (size_t)workspace_bytes, 0, true};
  auto P = [&](int slot) { return params + slot_off_host[slot]; };
  auto G = [&](int slot) { return grads + slot_off_host[slot]; };
  // buffers that cross the links sit at fixed workspace offsets (they depend on B and d only): the captured NCCL nodes
  // of a replayed step keep their pointers
  float *xpack = nullptr, *dshat_x = nullptr;
  int* labels_l = nullptr;
  if (shard) {
    xpack = ar.f(2 * (size_t)b.B);
    dshat_x = ar.f((size_t)b.B * d);
    labels_l = reinterpret_cast<int*>(ar.raw(sizeof(int) * (size_t)b.B));
  }
  const int s_ro = 1 + 8 * L;        // readout.fc_u.0.weight, .bias, fc_v, fc_e, fc_sr
  float* E = P(0);
  auto dcfg = [&](uint32_t site) { srk_dropout c; c.p = dropout_p; c.site = site; c.seed = seed; return c; };

  // with side streams the caller passes our own high-priority stream s[0] and orders it against the user's stream
  SideStreams* ss = srk_side_streams();
  StageTimer tm(st);
  // h1..h3: high-priority chains beside the critical path; s2 / s3: weight gradients; s4: catalog-wide bulk passes
  cudaStream_t h1 = ss ? ss->s[1] : st, h2 = ss ? ss->s[2] : st, h3 = ss ? ss->s[3] : st;
  cudaStream_t s2 = ss ? ss->s[4] : st, s3 = ss ? ss->s[5] : st, s4 = ss ? ss->s[6] : st;
  cudaStream_t s1 = h1;
  auto order = [&](cudaStream_t from, cudaStream_t to) { return ss ? ss->order(from, to) : (int)SRK_OK; };
  SRK_TRY(srk_step_begin());
  // zero_grad runs beside the forward pass; the first gradient is written after the head's backward.  Every scratch buffer of
  // the step that is accumulated into (split-K GEMM outputs, staged weight gradients, dS of the fused head) comes from ONE
  // pool that the same launch zeroes: 8 zero-fill launches less per step, three of them on the critical path.
  // wide fused head (d > 128): the session tiles add their table gradient into ONE zeroed [Vl, d] buffer (see step_srgnn.cu)
  static const bool de_atomic_on = [] { const char* e = getenv("SESSREC_FCE_DE_ATOMIC"); return !(e && e[0] == '0'); }();
  const bool de_atomic = flash && de_atomic_on && srk_flash_ce_bwd_parts(b.B) > 1;      // one session tile: nothing to add up
  const size_t zp_bytes = sizeof(float) * ((size_t)L * 2 * ((size_t)ldzel * d + (size_t)H * d + (size_t)N * d + 192) +
                                            2 * (size_t)b.B * d + 256 + (de_atomic ? (size_t)Vl * d : 0));
  Arena zp{ar.raw(zp_bytes), zp_bytes, 0, true};
  SRK_REQUIRE(ar.ok, "step: workspace too small");
  SRK_TRY(order(st, s4));
  SRK_TRY(srk_zero2_async(grads, sizeof(float) * (size_t)n_flat, zp.base, zp_bytes, s4));
  tm.mark("zero_grad");

  // ---- forward -------------------------------------------------------------------------------------------
  float *Ehat = ar.f((size_t)Vl * d), *enorm = ar.f(Vl);
  float *Ehi = nullptr, *Elo = nullptr;
  uint16_t *Ebh = nullptr, *Ebl = nullptr;
  if (flash) {
    Ebh = reinterpret_cast<uint16_t*>(ar.raw((size_t)Vl * d * 2));
    Ebl = reinterpret_cast<uint16_t*>(ar.raw((size_t)Vl * d * 2));
  } else if (umma) {
    Ehi = ar.f((size_t)Vl * d);
    Elo = ar.f((size_t)Vl * d);
  }
  float* El = E + (size_t)lo * d;      // the rows this rank scores (all of them without sharding)
  SRK_REQUIRE(ar.ok, "step: workspace too small");
  // nn.Embedding(max_norm=1) renorms the rows a lookup touches (msgifsr.py:247) and, at the scoring head, every row
  // (msgifsr.py:276).  The gather only needs the touched rows: they are renormed first on the critical path; the
  // catalog-wide pass (renorm of the remaining rows + normalisation + bf16 split) starts once the gather has read the
  // table and runs beside the encoder.
  // Sharded: the rows of other owners are renormed by their owners and come back with the end-of-step broadcast; the rows
  // the gather touches are renormed by every replica (same arithmetic on the same values).
  if (ss || shard) SRK_TRY(srk_renorm_rows(E, b.uid, b.U, d, 1.0f, st));
  else SRK_TRY(srk_catalog_prep_fwd(E, V, d, SRK_NORM_L2, 1.0f, Ehat, enorm, Ehi, Elo, Ebh, Ebl, st));
  tm.mark("catalog_prep");
  float *X = ar.f((size_t)N * d), *rnX = ar.f(N);
  srk_dropout dc_e = dcfg(SRK_SITE_EMBED + 1);
  SRK_TRY(srk_embed_gather_fwd(E, b.iid, N, d, SRK_NORM_L2, drop ? &dc_e : nullptr, X, rnX, nullptr, st));
  if (ss || shard) {
    SRK_TRY(order(st, s4));
    SRK_TRY(srk_catalog_prep_fwd(El, Vl, d, SRK_NORM_L2, 1.0f, Ehat, enorm, Ehi, Elo, Ebh, Ebl, s4));
  }
  srk_dropout dc_attn = dcfg(SRK_SITE_GAT_ATTN);

  tm.mark("gather");
  std::vector<LayerRec> layers(L);
  // The N x d x 8d projections of the GAT layers (Z | el = x_s W_aug^T, its data gradient dZel W_aug and its weight gradient
  // dZel^T x_s) run on the tcgen05 3xTF32 GEMM instead of the fp32 CUDA-core kernel (measured 0.441 -> 0.419 ms/step at
  // cfg1 for the two on the critical path); SESSREC_TC_ENCODER=0 switches back.
  const char* tce = getenv("SESSREC_TC_ENCODER");
  const bool tc_enc = umma && d <= 256 && d % 32 == 0 && !(tce && tce[0] == '0');
  // W_aug = [W ; a_l-contracted rows] and w_r depend on the parameters only: built beside the gather, the two convolutions
  // of a layer on two streams (the first projection of the step waits for this chain: gat_prep -> split, ~10 us per conv)
  SRK_TRY(order(st, s2));                        // after the previous step's optimizer update
  SRK_TRY(order(st, s3));
  for (int l = 0; l < L; ++l) {
    LayerRec& R = layers[l];
    R.n_inst = M > 0 ? 2 : 0;
    for (int c = 0; c < R.n_inst; ++c) {
      InstRec& I = R.inst[c];
      const int base = 1 + 8 * l + 4 * c;        // attn_l, attn_r, bias, fc.weight
      I.al = P(base); I.ar = P(base + 1); I.W = P(base + 3);
      I.gal = G(base); I.gar = G(base + 1); I.gbias = G(base + 2); I.gW = G(base + 3);
      I.Waug = ar.f((size_t)ldzel * d);
      I.wr = ar.f((size_t)H * d);
      SRK_REQUIRE(ar.ok, "step: workspace too small");
      cudaStream_t ps = c == 0 ? s2 : s3;
      I.Wh = I.Wl = nullptr;
      if (tc_enc) {
        I.Wh = ar.f((size_t)ldzel * d);
        I.Wl = ar.f((size_t)ldzel * d);
        SRK_REQUIRE(ar.ok, "step: workspace too small");
      }
      SRK_TRY(srk_gat_prep_split(I.W, I.al, I.ar, d, I.Waug, I.wr, I.Wh, I.Wl, ps));      // one launch incl. the TF32 split
    }
  }
  // W_sr^T for the fused read-out tail (the lanes read consecutive output columns)
  // opt-in (SESSREC_FUSED_READOUT=1): measured 2 % SLOWER than the unfused chain at cfg1 (0.437 vs 0.429 ms/step) although
  // it saves 4 launches - the per-session mat-vecs run through a cold L1 on 64 CTAs while the GEMM launches fill the machine
  const char* fro = getenv("SESSREC_FUSED_READOUT");
  const bool fused_ro = fro && fro[0] == '1';
  float* WsrT = nullptr;
  if (fused_ro) {
    WsrT = ar.f(2 * (size_t)d * d);
    SRK_REQUIRE(ar.ok, "step: workspace too small");
    SRK_TRY(srk_transpose(P(s_ro + 4), d, 2 * d, WsrT, s2));
  }
  SRK_TRY(order(s2, st));
  SRK_TRY(order(s3, st));
  const float* h = X;
  for (int l = 0; l < L; ++l) {
    LayerRec& R = layers[l];
    R.in = h;
    R.normalize = (l == L - 1);
    srk_gat_inst insts[2];
    // the layer input is ready on the main stream: four projection chains (conv x {source, destination} copy) and the
    // segment mean run side by side
    SRK_TRY(order(st, h1));
    SRK_TRY(order(st, h2));
    SRK_TRY(order(st, h3));
    SRK_TRY(order(st, s2));
    for (int c = 0; c < R.n_inst; ++c) {
      cudaStream_t zs = c == 0 ? st : h1;        // source copy: Z | el = x_s W_aug^T   (conv1 = graph, conv2 = reversed graph)
      cudaStream_t es = c == 0 ? h2 : h3;        // destination copy: er = x_d w_r^T
      InstRec& I = R.inst[c];
      const int base = 1 + 8 * l + 4 * c;
      const uint32_t slot = (uint32_t)((l * 2 + c) * 3);
      I.drop = drop;
      I.xs = const_cast<float*>(h);
      I.xd = const_cast<float*>(h);
      if (drop) {
        I.dcs = dcfg(SRK_SITE_GAT_SRC + 4 * slot);
        I.dcd = dcfg(SRK_SITE_GAT_DST + 4 * slot);
        I.xs = ar.f((size_t)N * d);
        I.xd = ar.f((size_t)N * d);
        SRK_REQUIRE(ar.ok, "step: workspace too small");
        if (tc_enc) {                             // the source copy together with its TF32 split
          I.xh = ar.f((size_t)N * d);
          I.xl = ar.f((size_t)N * d);
          SRK_REQUIRE(ar.ok, "step: workspace too small");
          SRK_TRY(srk_dropout_apply_split(h, I.xs, I.xh, I.xl, (long long)N * d, &I.dcs, zs));
        } else {
          SRK_TRY(srk_dropout_apply(h, I.xs, (long long)N * d, &I.dcs, 0, zs));
        }
        SRK_TRY(srk_dropout_apply(h, I.xd, (long long)N * d, &I.dcd, 0, es));
      }
      float* Zel = ar.f((size_t)N * ldzel);
      float* er = ar.f((size_t)N * H);
      float* att = ar.f((size_t)(M + 1) * H);
      SRK_REQUIRE(ar.ok, "step: workspace too small");
      if (tc_enc) {
        if (!drop) {
          I.xh = ar.f((size_t)N * d);
          I.xl = ar.f((size_t)N * d);
          SRK_REQUIRE(ar.ok, "step: workspace too small");
          SRK_TRY(srk_split_tf32(I.xs, d, N, d, I.xh, I.xl, d, zs));
        }
        SRK_TRY(srk_umma_gemm(0, N, ldzel, d, I.xh, I.xl, d, I.Wh, I.Wl, d, Zel, ldzel, 1.0f, 0, 1, zs));
      } else {
        SRK_TRY(linear_nt(zs, N, ldzel, d, I.xs, d, I.Waug, Zel, ldzel));
      }
      SRK_TRY(linear_nt(es, N, H, d, I.xd, d, I.wr, er, H));
      srk_gat_inst& g = I.gi;
      memset(&g, 0, sizeof(g));
      if (c == 0) {
        g.in_ptr = b.in_ptr; g.in_src = b.in_src; g.in_eid = b.in_eid;
        g.out_ptr = b.out_ptr; g.out_dst = b.out_dst; g.out_eid = b.out_eid;
      } else {            // conv2 runs on the reversed graph: the two CSRs swap roles
        g.in_ptr = b.out_ptr; g.in_src = b.out_dst; g.in_eid = b.out_eid;
        g.out_ptr = b.in_ptr; g.out_dst = b.in_src; g.out_eid = b.in_eid;
      }
      g.Zel = Zel; g.er = er; g.bias = P(base + 2); g.xdst = I.xd; g.att = att;
      g.n_src = N; g.n_dst = N; g.n_edges = M;
      g.attn_site = SRK_SITE_GAT_ATTN + 4 * slot;
      insts[c] = g;
    }
    float* segmean = ar.f((size_t)B * d);
    R.Hout = ar.f((size_t)N * d);
    R.rn = ar.f(N);
    R.amax = ar.raw((size_t)N * d);
    SRK_REQUIRE(ar.ok, "step: workspace too small");
    SRK_TRY(srk_segmean_fwd(h, b.seg, B, d, segmean, s2));
    SRK_TRY(order(h1, st));
    SRK_TRY(order(h2, st));
    SRK_TRY(order(h3, st));
    SRK_TRY(order(s2, st));
    SRK_TRY(srk_gat_aggregate_fwd(insts, R.n_inst, N, d, segmean, b.node2seg, drop ? &dc_attn : nullptr, R.normalize, R.Hout,
                                  R.rn, R.amax, st));
    h = R.Hout;
  }
  tm.mark("gat_fwd");
  const float* F = h;
  float *u = ar.f((size_t)N * d), *v = ar.f((size_t)B * d), *e = ar.f(N), *ms = ar.f(2 * (size_t)B);
  float *sr_in = ar.f(2 * (size_t)B * d), *s = fused_ro ? ar.f((size_t)B * d) : zp.f((size_t)B * d), *shat = ar.f((size_t)B * d), *rn_s = ar.f(B);
  SRK_REQUIRE(ar.ok, "step: workspace too small");
  uint16_t *Sbh = nullptr, *Sbl = nullptr;
  if (flash) {
    Sbh = reinterpret_cast<uint16_t*>(ar.raw((size_t)B * d * 2));
    Sbl = reinterpret_cast<uint16_t*>(ar.raw((size_t)B * d * 2));
    SRK_REQUIRE(ar.ok, "step: workspace too small");
  }
  SRK_TRY(order(st, s1));
  SRK_TRY(linear_nt(st, N, d, d, F, d, P(s_ro), u, d, nullptr, P(s_ro + 1)));
  SRK_TRY(linear_nt(s1, B, d, d, F, d, P(s_ro + 2), v, d, b.last, nullptr));
  SRK_TRY(order(s1, st));
  if (fused_ro) {
    // one launch: attention scores, segment soft-max, fc_sr, normalisation, bf16 split (csrc/readout_fused.cu)
    SRK_TRY(srk_readout_tail_fwd(F, u, v, P(s_ro + 3), WsrT, b.seg, b.last, B, d, SRK_NORM_L2, e, ms, sr_in, s, shat, rn_s, Sbh,
                                 Sbl, st));
  } else {
    SRK_TRY(srk_readout_fwd(F, u, v, P(s_ro + 3), b.seg, b.last, B, d, 1, e, ms, sr_in, st));
    // s comes zeroed from the pool: the split-K accumulate path without the zero launch srk_gemm would put in front
    SRK_TRY(gemm(st, B, d, 2 * d, sr_in, 2 * d, 1, P(s_ro + 4), 1, 2 * d, s, d, nullptr, nullptr, nullptr, nullptr, 1.f, 1));
    if (flash) SRK_TRY(srk_rownorm_split_fwd(s, d, B, d, SRK_NORM_L2, shat, d, rn_s, Sbh, Sbl, st));
    else SRK_TRY(srk_rownorm_fwd(s, d, B, d, SRK_NORM_L2, shat, d, rn_s, st));
  }
  tm.mark("readout_fwd");
  // scoring head + CE (needs the catalog pass)
  SRK_TRY(order(s4, st));
  float *Z = flash ? nullptr : ar.f((size_t)B * ldz), *lse = ar.f(B), *nll = ar.f(B);
  float *sh = nullptr, *sl = nullptr;
  SRK_REQUIRE(ar.ok, "step: workspace too small");
  // sharded: the head sees the label only where this rank owns its row; lse / nll are then LOCAL until the exchange below
  const int* hlabels = b.labels;
  if (shard) {
    SRK_TRY(srk_shard_labels(b.labels, B, lo, hi, labels_l, st));
    hlabels = labels_l;
  }
  if (flash) {
    float* part = ar.f((size_t)srk_flash_ce_part_floats(B, Vl));
    SRK_REQUIRE(ar.ok, "step: workspace too small");
    SRK_TRY(srk_flash_ce_fwd(B, Vl, d, Sbh, Sbl, d, Ebh, Ebl, d, 12.0f, hlabels, lse, nll, part, st));
  } else if (umma) {
    sh = ar.f((size_t)B * d); sl = ar.f((size_t)B * d);
    SRK_TRY(srk_split_tf32(shat, d, B, d, sh, sl, d, st));
    if (fused_lse) {
      float* part = ar.f(4 * (size_t)((Vl + 255) / 256) * B + B);
      SRK_REQUIRE(ar.ok, "step: workspace too small");
      SRK_TRY(srk_umma_score_fwd(B, Vl, d, sh, sl, d, Ehi, Elo, d, Z, ldz, 12.0f, hlabels, lse, nll, part, st));
    } else {
      SRK_TRY(srk_umma_gemm(0, B, Vl, d, sh, sl, d, Ehi, Elo, d, Z, ldz, 12.0f, 0, 1, st));
      SRK_TRY(srk_ce_rows_fwd(Z, ldz, hlabels, B, Vl, 0, lse, nll, st));
    }
  } else {
    SRK_TRY(gemm(st, B, Vl, d, shat, d, 1, Ehat, 1, d, Z, ldz, nullptr, nullptr, nullptr, nullptr, 12.0f));
    SRK_TRY(srk_ce_rows_fwd(Z, ldz, hlabels, B, Vl, 0, lse, nll, st));
  }
  if (shard) {
    // THE exchange of the sharded head: per session sum_v exp(logit - 12) over this rank's rows and the label logit where
    // owned - one [2, B] sum all-reduce over NVLink - then the global log-sum-exp / NLL (cosine logits: |logit| <= 12)
    SRK_TRY(srk_shard_lse_pack(lse, nll, labels_l, nullptr, 12.0f, B, xpack, st));
    SRK_TRY(srk_comm_allreduce(xpack, 2LL * B, 0, st));
    SRK_TRY(srk_shard_lse_unpack(xpack, nullptr, 12.0f, B, lse, nll, st));
  }

  tm.mark("score_fwd+lse");
  // ---- forward / backward boundary: every side stream has been joined into `st` -------------------------------------
  // The forward half is always launched kernel by kernel (the GPU starts working while the host is still enqueueing);
  // the backward half is what gets captured into / replayed from a CUDA graph (see srk_msgifsr_train_step).
  SRK_TRY(srk_step_boundary());
  // the backward needs the row log-sum-exps, not their mean: the loss reduction runs beside the head's backward (s2 is
  // joined before the optimizer step)
  SRK_TRY(order(st, s2));
  SRK_TRY(srk_mean(nll, B, loss_out, s2));
  // ---- backward ------------------------------------------------------------------------------------------
  const int de_parts = (flash && !de_atomic) ? srk_flash_ce_bwd_parts(B) : 1;
  float* Zlo = (umma && !flash) ? ar.f((size_t)B * ldz) : nullptr;
  const bool ds_pooled = flash && !shard;       // dS accumulates (TMA reduce-add): zeroed with the pool
  float *dshat = shard ? dshat_x : (ds_pooled ? zp.f((size_t)B * d) : ar.f((size_t)B * d));
  float* dEhat = de_atomic ? zp.f((size_t)Vl * d) : ar.f((size_t)de_parts * Vl * d);
  SRK_REQUIRE(ar.ok && zp.ok, "step: workspace too small");
  if (flash) {
    SRK_TRY(srk_flash_ce_bwd_ex(B, Vl, d, Sbh, Sbl, d, Ebh, Ebl, d, 12.0f, hlabels, lse, one_dev, dshat, dEhat,
                                (ds_pooled ? 1 : 0) | (de_atomic ? 2 : 0), st));
  } else if (umma) {
    SRK_TRY(srk_zero_async(dshat, sizeof(float) * (size_t)B * d, st));
    // Backward of the head, chunked over catalog columns so that each chunk's dZ hi/lo pair (2 x B x Vc x 4 bytes) is
    // still L2-resident when the two tensor-core GEMMs read it (the whole pair, 2 x 88 MB at cfg1, is not).
    int chunks = head_chunks < 1 ? 1 : head_chunks;
    int Vc = ((Vl + chunks - 1) / chunks + 255) / 256 * 256;
    for (int c0 = 0; c0 < Vl; c0 += Vc) {
      const int nc = Vl - c0 < Vc ? Vl - c0 : Vc;
      SRK_TRY(srk_ce_rows_bwd_cols(Z, ldz, hlabels, lse, one_dev, 12.0f, B, c0, nc, Zlo, st));
      const int nkb = (nc + 31) / 32;
      int split = 148 / ((B + 127) / 128);      // one wave of CTAs
      if (split < 1) split = 1;
      if (split > nkb) split = nkb;
      SRK_TRY(srk_umma_gemm(1, B, d, nc, Z + c0, Zlo + c0, ldz, Ehi + (size_t)c0 * d, Elo + (size_t)c0 * d, d, dshat, d, 1.0f, 1,
                            split, st));
      SRK_TRY(srk_umma_gemm(2, nc, d, B, Z + c0, Zlo + c0, ldz, sh, sl, d, dEhat + (size_t)c0 * d, d, 1.0f, 0, 1, st));
    }
  } else {
    SRK_TRY(srk_zero_async(dshat, sizeof(float) * (size_t)B * d, st));
    SRK_TRY(srk_ce_rows_bwd(Z, ldz, hlabels, lse, one_dev, 12.0f, B, Vl, 0, Zlo, st));
    SRK_TRY(srk_zero_async(dEhat, sizeof(float) * (size_t)Vl * d, st));
    SRK_TRY(gemm(st, B, d, Vl, Z, ldz, 1, Ehat, d, 1, dshat, d, nullptr, nullptr, nullptr, nullptr, 1.f, 1));
    SRK_TRY(gemm(st, Vl, d, B, Z, 1, ldz, shat, d, 1, dEhat, d, nullptr, nullptr, nullptr, nullptr, 1.f, 1));
  }
  tm.mark("ce_bwd+dS+dE");
  // gradients start here: zero_grad (s4) must be complete; the catalog backward (a [V, d] pass) stays on s4 and runs
  // beside the whole encoder backward, it only has to finish before the scatter-add touches the same rows
  // (zero_grad and the catalog pass on s4 were joined before the head's forward)
  SRK_TRY(order(st, s4));
  SRK_TRY(srk_catalog_prep_bwd(El, Ehat, enorm, dEhat, de_parts, Vl, d, SRK_NORM_L2, G(0) + (size_t)lo * d, s4));
  // sharded: every rank holds dS of its own catalog rows only - the second (and last) exchange of the head, while the
  // catalog backward above runs on s4
  if (shard) SRK_TRY(srk_comm_allreduce(dshat, (long long)B * d, 0, st));
  const bool live = srk_launch_mode() == SRK_LAUNCH_DIRECT || srk_launch_mode() == SRK_LAUNCH_CAPTURE;
  if (ss && live) SRK_CUDA(cudaEventRecord(ss->ev_cat, s4));
  // Adam in two parts: the table rows this batch did not gather have their final gradient now (the scatter-add only
  // touches gathered rows, and only those rows of E are read again by the backward), so their update - 95 % of the
  // optimizer's bytes - runs on s4 beside the encoder backward; the gathered rows and all other parameters follow at the end
  // (single device only: under data parallelism a row is final after the all-reduce, under sharding the owner updates it)
  const bool split_adam = ss != nullptr && phase == 0 && do_adam && !shard;
  const long long tab = slot_off_host[0];
  const long long tab_span = ((long long)V * d + 63) / 64 * 64;
  if (split_adam)
    SRK_TRY(srk_adam_step_split(params, grads, exp_avg, exp_avg_sq, n_flat, seg_off_dev, seg_decay_dev, n_seg, tab, V, d,
                                tab_span, b.uid, b.U, 0, lr, beta1, beta2, eps, adam_step, grad_scale, s4));
  tm.mark("catalog_bwd");
  float* ds = ar.f((size_t)B * d);
  float* dsr_in = ar.f(2 * (size_t)B * d);
  float* dF = ar.f((size_t)N * d);
  SRK_REQUIRE(ar.ok, "step: workspace too small");
  // data gradients on the main stream, weight gradients (mm_tn / colsum into the flat gradient buffer) on s2
  if (fused_ro) {
    // one launch: normalise backward, d sr_in = d s W_sr, attention backward (csrc/readout_fused.cu); then the projections
    SRK_TRY(srk_readout_head_bwd(F, P(s_ro + 3), P(s_ro + 4), b.seg, b.last, B, d, SRK_NORM_L2, s, shat, rn_s, sr_in, e, ms, dshat,
                                 u, v, ds, dF, G(s_ro + 3), st));
    SRK_TRY(order(st, s2));
    SRK_TRY(mm_nn(st, N, d, d, u, d, P(s_ro), d, dF, d, 1));                      // u holds du
    SRK_TRY(mm_nn(st, B, d, d, v, d, P(s_ro + 2), d, dF, d, 1, b.last));          // v holds dv
    SRK_TRY(mm_tn(s2, d, 2 * d, B, ds, d, sr_in, 2 * d, G(s_ro + 4), 2 * d));
    SRK_TRY(mm_tn(s2, d, d, N, u, d, F, d, G(s_ro), d));
    SRK_TRY(srk_colsum(u, d, N, d, G(s_ro + 1), 1, s2));
    SRK_TRY(mm_tn(s2, d, d, B, v, d, F, d, G(s_ro + 2), d, b.last));
  } else {
    SRK_TRY(srk_rownorm_bwd(s, d, shat, d, rn_s, dshat, d, B, d, SRK_NORM_L2, ds, d, 0, st));
    SRK_TRY(order(st, s2));
    SRK_TRY(mm_nn(st, B, 2 * d, d, ds, d, P(s_ro + 4), 2 * d, dsr_in, 2 * d, 0));
    SRK_TRY(mm_tn(s2, d, 2 * d, B, ds, d, sr_in, 2 * d, G(s_ro + 4), 2 * d));
    SRK_TRY(srk_readout_bwd(F, u, v, P(s_ro + 3), b.seg, b.last, e, ms, sr_in, dsr_in, B, d, 1, dF, G(s_ro + 3), st));
    SRK_TRY(order(st, s2));
    SRK_TRY(mm_nn(st, N, d, d, u, d, P(s_ro), d, dF, d, 1));                      // u holds du
    SRK_TRY(mm_tn(s2, d, d, N, u, d, F, d, G(s_ro), d));
    SRK_TRY(srk_colsum(u, d, N, d, G(s_ro + 1), 1, s2));
    SRK_TRY(mm_nn(st, B, d, d, v, d, P(s_ro + 2), d, dF, d, 1, b.last));          // v holds dv
    SRK_TRY(mm_tn(s2, d, d, B, v, d, F, d, G(s_ro + 2), d, b.last));
  }

  tm.mark("readout_bwd");
  // layers, last to first.  Scratch below is re-carved per layer from a fixed mark.
  const size_t mark = ar.off;
  const float* dH = dF;
  float* dfeat = nullptr;
  for (int l = L - 1; l >= 0; --l) {
    LayerRec& R = layers[l];
    // dfeat must outlive this iteration (it is the next layer's dH): alternate two halves of the scratch region
    ar.off = mark;
    float* bufA = ar.f((size_t)N * d);
    float* bufB = ar.f((size_t)N * d);
    dfeat = ((L - 1 - l) & 1) ? bufB : bufA;
    // d(layer input) = segment-mean term + per conv {source-copy term, destination-copy + residual term}: every term
    // is produced on its own stream into its own [N, d] slice and one pass adds them up
    float* parts = ar.f(5 * (size_t)N * d);
    const size_t nd = (size_t)N * d;
    float* dHpre = ar.f(nd);
    srk_gat_inst insts[2];
    float *dedge[2], *der[2], *dZel[2];
    for (int c = 0; c < R.n_inst; ++c) {
      dedge[c] = ar.f((size_t)(M + 1) * H);
      der[c] = ar.f((size_t)N * H);
      dZel[c] = ar.f((size_t)N * ldzel);
      R.inst[c].gi.dedge = dedge[c]; R.inst[c].gi.der = der[c]; R.inst[c].gi.dZel = dZel[c];
      insts[c] = R.inst[c].gi;
    }
    SRK_REQUIRE(ar.ok, "step: workspace too small");
    SRK_TRY(srk_gat_aggregate_bwd_dst(insts, R.n_inst, N, d, drop ? &dc_attn : nullptr, R.normalize, R.Hout, R.rn, R.amax, dH,
                                      dHpre, st));
    SRK_TRY(order(st, h1));
    SRK_TRY(order(st, h3));
    SRK_TRY(srk_segmean_bwd(dHpre, b.seg, B, d, parts, 0, h3));
    for (int c = 0; c < R.n_inst; ++c) {
      cudaStream_t zs = c == 0 ? st : h1, es = c == 0 ? h2 : h3, wsc = c == 0 ? s2 : s3;
      float* pz = parts + (1 + 2 * c) * nd;      // source-copy term
      float* pe = parts + (2 + 2 * c) * nd;      // destination-copy term + identity residual
      InstRec& I = R.inst[c];
      float *zh = nullptr, *zl = nullptr;
      if (tc_enc) {                               // dZel together with its TF32 hi / lo (data and weight gradient GEMMs)
        zh = ar.f((size_t)N * ldzel);
        zl = ar.f((size_t)N * ldzel);
        SRK_REQUIRE(ar.ok, "step: workspace too small");
        SRK_TRY(srk_gat_aggregate_bwd_src_split(&insts[c], d, drop ? &dc_attn : nullptr, dHpre, R.amax, zh, zl, zs));
      } else {
        SRK_TRY(srk_gat_aggregate_bwd_src(&insts[c], d, drop ? &dc_attn : nullptr, dHpre, R.amax, zs));
      }
      SRK_TRY(order(zs, es));
      SRK_TRY(order(zs, wsc));
      // weight gradients
      SRK_TRY(srk_gat_bias_bwd(dHpre, R.amax, N, d, I.gbias, wsc));
      float* dWaug = zp.f((size_t)ldzel * d);      // zeroed with the pool
      float* dwr = zp.f((size_t)H * d);
      SRK_REQUIRE(zp.ok, "step: zero pool too small");
      if (tc_enc) {                               // dW_aug[8d + 8, d] = dZel^T x_s, split over the N rows
        int split = 148 / ((ldzel + 127) / 128);
        SRK_TRY(srk_umma_gemm(2, ldzel, d, N, zh, zl, ldzel, I.xh, I.xl, d, dWaug, d, 1.0f, 1, split < 1 ? 1 : split, wsc));
      } else {
        SRK_TRY(mm_tn(wsc, ldzel, d, N, dZel[c], ldzel, I.xs, d, dWaug, d));
      }
      SRK_TRY(mm_tn(wsc, H, d, N, der[c], H, I.xd, d, dwr, d));
      SRK_TRY(srk_gat_prep_bwd(I.W, I.al, I.ar, dWaug, dwr, d, I.gW, I.gal, I.gar, wsc));
      // data gradients
      auto dz_times_waug = [&](float* out, bool out_zeroed) -> int {          // out[N, d] = dZel[N, 8d + 8] W_aug[8d + 8, d]
        if (!tc_enc) return mm_nn(zs, N, d, ldzel, dZel[c], ldzel, I.Waug, d, out, d, 0);
        if (!out_zeroed) SRK_TRY(srk_zero_async(out, sizeof(float) * (size_t)N * d, zs));
        int split = 148 / ((N + 127) / 128);
        if (split < 1) split = 1;
        return srk_umma_gemm(1, N, d, ldzel, zh, zl, ldzel, I.Wh, I.Wl, d, out, d, 1.0f, 1, split, zs);
      };
      if (!I.drop) {
        SRK_TRY(dz_times_waug(pz, false));
        SRK_TRY(mm_nn(es, N, d, H, der[c], H, I.wr, d, pe, d, 0));
        SRK_TRY(srk_dropout_apply(dHpre, pe, (long long)nd, nullptr, 1, es));            // residual
      } else {
        float* tmp = tc_enc ? zp.f(nd) : ar.f(nd);      // split-K accumulation target: zeroed with the pool
        float* tmp2 = ar.f(nd);
        SRK_REQUIRE(ar.ok && zp.ok, "step: workspace too small");
        SRK_TRY(dz_times_waug(tmp, tc_enc));
        SRK_TRY(srk_dropout_apply(tmp, pz, (long long)nd, &I.dcs, 0, zs));
        SRK_TRY(mm_nn(es, N, d, H, der[c], H, I.wr, d, tmp2, d, 0));
        SRK_TRY(srk_dropout_apply_add(dHpre, tmp2, pe, (long long)nd, &I.dcd, es));       // mask(residual + der w_r)
      }
    }
    SRK_TRY(order(h1, st));
    SRK_TRY(order(h2, st));
    SRK_TRY(order(h3, st));
    SRK_TRY(srk_sum_parts(parts, (long long)nd, 1 + 2 * R.n_inst, (long long)nd, dfeat, 0, st));
    if (l > 0) {                                 // the scratch region is re-carved by the next layer
      SRK_TRY(order(s2, st));
      SRK_TRY(order(s3, st));
    }
    dH = dfeat;
  }
  tm.mark("gat_bwd");
  // catalog backward done (not the early Adam part queued behind it): the scatter-add updates the same table rows
  if (ss && live) SRK_CUDA(cudaStreamWaitEvent(st, ss->ev_cat, 0));
  // The step is not bit-reproducible from run to run as a whole (split-K weight-gradient GEMMs, the d w_e reduction), so it
  // takes the one-launch scatter-add by default (atomicAdd on runs cut by a chunk boundary; measured 0.402 vs 0.418 ms per
  // step at cfg1); SESSREC_DETERMINISTIC_SCATTER=1 selects the two-pass variant without atomics.
  static const bool scatter_det = getenv("SESSREC_DETERMINISTIC_SCATTER") != nullptr;
  float* sws = !scatter_det ? nullptr : ar.f((size_t)srk_embed_scatter_ws_floats(b.P, d));
  SRK_REQUIRE(ar.ok, "step: workspace too small");
  SRK_TRY(srk_embed_scatter_bwd_ws(E, b.iid, b.perm, b.uoff, b.uid, b.U, b.P, d, SRK_NORM_L2, drop ? &dc_e : nullptr, rnX, dH,
                                   nullptr, G(0), sws, st));
  SRK_TRY(order(s2, st));
  SRK_TRY(order(s3, st));
  tm.mark("scatter");
  if (dp_inside) {
    // data parallel: the flat gradient buffer (every rank seeded its backward with B_local / B_global) is summed over the
    // ranks right here, behind the last gradient kernel, by this library's own communicator - no return to Python - and Adam
    // follows.  When the backward half is replayed from a CUDA graph the two are enqueued with plain calls behind the graph.
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
    if (shard) {
      SRK_TRY(order(s4, st));                    // the owned rows' head gradient (catalog backward on s4)
      // The encoder is replicated: every rank computed the gradients of the non-table parameters from the same inputs, up
      // to the summation order of atomics.  Averaging them (two small all-reduces around the table, ~MBs) makes the
      // replicas' updates bit-identical, so they cannot drift apart over a long run.
      SRK_TRY(srk_comm_allreduce(grads, tab, 2, st));
      SRK_TRY(srk_comm_allreduce(grads + tab + tab_span, n_flat - tab - tab_span, 2, st));
    }
    SRK_TRY(srk_adam_step(params, grads, exp_avg, exp_avg_sq, n_flat, seg_off_dev, seg_decay_dev, n_seg, lr, beta1, beta2, eps,
                          adam_step, grad_scale, st));
    // sharded: seg_decay marks the table rows of the other owners inactive (rank-local Adam); every owner now hands its
    // updated rows to the other replicas - one grouped broadcast instead of an all-reduce of the [V, d] table gradient
    if (shard) SRK_TRY(srk_comm_share_rows(E, V, d, st));
  }
  tm.mark("adam");
  SRK_TRY(order(s4, st));
  tm.report();
  return SRK_OK;
}

namespace {

int g_graphs_on = -1;                  // -1: read SESSREC_GRAPH on first use; 0 never, 1 always, 2 auto (data-parallel steps)
long long g_graph_launches = 0;        // steps issued as one cudaGraphLaunch
long long g_graph_node_updates = 0;
long long g_graph_fallbacks = 0;       // update passes that found a different kernel sequence
int g_inject_mismatch = 0;             // test hook: the n-th next update pass is made to fail half-way (0 = off)

struct GraphEntry {
  int seen = 0, fails = 0;
  bool bad = false;
  SrkStepGraph g;
  // auto policy of single-rank steps: host and device time of steps 2 and 3 (warm, plain launches) decide once, so that a
  // capture happens at step 4 - inside the warm-up of any training loop or benchmark
  static constexpr int NMEAS = 2;
  int decided = -1;                     // -1 not yet, 0 plain launches, 1 graph replay
  cudaEvent_t ev0[NMEAS] = {nullptr}, ev1[NMEAS] = {nullptr};
  double host_ms = 0.0;
};

}  // namespace

static int switch_launch_mode(SrkLaunchCtx* lc) {
  if (lc->mode == lc->mode_after_boundary) return SRK_OK;
  if (lc->mode_after_boundary == SRK_LAUNCH_CAPTURE) {
    SRK_CUDA(cudaStreamBeginCapture(lc->capture_stream, cudaStreamCaptureModeThreadLocal));
    lc->capturing = true;
  }
  lc->mode = lc->mode_after_boundary;
  return SRK_OK;
}

// first call of a step body: a whole-step graph starts here
int srk_step_begin() {
  SrkLaunchCtx* lc = srk_get_launch_ctx();
  return lc && lc->whole ? switch_launch_mode(lc) : (int)SRK_OK;
}

int srk_step_boundary() {
  if (SrkLaunchCtx* lc = srk_get_launch_ctx()) return switch_launch_mode(lc);
  return SRK_OK;
}

// Host cost: a step is ~60 kernels on 7 streams.  Launched one by one that is ~0.36 ms of CPU per step (2.7 us per
// launch + ~60 event record / wait calls + memsets), more than the GPU needs once several ranks share the host.  After two
// warm-up steps the sequence is therefore captured ONCE into a CUDA graph (per model configuration `key`); every later
// step only rewrites the kernel-node parameters (shapes and pointers change with the batch, the sequence does not) and
// issues one cudaGraphLaunch.  SESSREC_GRAPH=0 disables it; any mismatch falls back to plain launches for that step.
int srk_step_driver(cudaStream_t caller, unsigned long long key, int want_graph, const std::function<int(void*)>& body_on,
                    bool whole) {
  // The step runs on our own high-priority stream s[0] (the user's stream may be the legacy default stream, which can
  // neither be prioritised nor captured); it is ordered after / before the user's stream with events.
  SideStreams* ss0 = srk_side_streams();
  cudaStream_t run = ss0 ? ss0->s[0] : caller;
  auto body_on_run = [&]() { return body_on((void*)run); };
  auto body = [&]() {                     // plain launches
    if (ss0) SRK_TRY(ss0->order_always(caller, run));
    SRK_TRY(body_on_run());
    if (ss0) SRK_TRY(ss0->order_always(run, caller));
    return (int)SRK_OK;
  };
  const char* timing = getenv("SESSREC_STEP_TIMING");
  if (!want_graph || ss0 == nullptr || (timing && timing[0] != '0')) return body();

  static std::mutex mu;
  static std::map<unsigned long long, GraphEntry> cache;
  static const bool debug = getenv("SESSREC_GRAPH_DEBUG") != nullptr;
  std::lock_guard<std::mutex> lock(mu);
  GraphEntry& e = cache[key];
  ++e.seen;
  if (want_graph == 2 && !e.bad && e.decided < 0) {
    // Auto (single-rank steps): replay pays when the HOST is the bottleneck.  A cfg1 step needs ~0.29 ms of enqueue on a fast
    // host against 0.34 ms on the GPU, but 0.45 ms on a slower one and 0.7-0.8 ms when the process shares its core (measured:
    // replay then runs the step in 0.54 ms).  Steps 2 and 3 - warm, plain launches - are timed on both sides (wall clock inside
    // the call, CUDA events around the step) and the sums decide once per configuration at step 4.
    constexpr int NM = GraphEntry::NMEAS;
    if (e.seen == 1) return body();
    if (e.seen <= 1 + NM) {
      const int i = e.seen - 2;
      if (cudaEventCreate(&e.ev0[i]) != cudaSuccess || cudaEventCreate(&e.ev1[i]) != cudaSuccess) {
        cudaGetLastError();
        e.decided = 0;
        return body();
      }
      SRK_TRY(ss0->order_always(caller, run));
      SRK_CUDA(cudaEventRecord(e.ev0[i], run));
      const auto t0 = std::chrono::steady_clock::now();
      SRK_TRY(body_on_run());
      e.host_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      SRK_CUDA(cudaEventRecord(e.ev1[i], run));
      SRK_TRY(ss0->order_always(run, caller));
      return SRK_OK;
    }
    double dev_ms = 0.0;
    bool timed = cudaEventSynchronize(e.ev1[NM - 1]) == cudaSuccess;
    for (int i = 0; i < NM; ++i) {
      float ms = 0.f;
      timed = timed && cudaEventElapsedTime(&ms, e.ev0[i], e.ev1[i]) == cudaSuccess;
      dev_ms += ms;
      cudaEventDestroy(e.ev0[i]);
      cudaEventDestroy(e.ev1[i]);
      e.ev0[i] = e.ev1[i] = nullptr;
    }
    if (!timed) cudaGetLastError();
    e.decided = timed && e.host_ms >= 0.85 * dev_ms ? 1 : 0;
    if (debug)
      fprintf(stderr, "[sessrec graph] auto: steps 2..%d took %.3f ms of host enqueue, %.3f ms on the device -> %s\n", 1 + NM,
              e.host_ms, dev_ms, e.decided ? "graph replay of the backward half" : "plain launches");
  }
  if (want_graph == 2 && e.decided == 0) return body();
  if (e.bad || e.seen <= 2) return body();          // warm-up steps also run every one-time cudaFuncSetAttribute

  // forward: plain launches; backward: captured once, afterwards its kernel nodes are re-parameterised and replayed
  SrkLaunchCtx ctx;
  ctx.g = &e.g;
  ctx.capture_stream = run;
  ctx.whole = whole;
  const bool capture = e.g.exec == nullptr;
  ctx.mode_after_boundary = capture ? SRK_LAUNCH_CAPTURE : SRK_LAUNCH_UPDATE;
  if (!capture && g_inject_mismatch > 0 && --g_inject_mismatch == 0) ctx.fail_at = e.g.nodes.size() / 2;
  SRK_TRY(ss0->order_always(caller, run));
  srk_set_launch_ctx(&ctx);
  int rc = body_on_run();
  srk_set_launch_ctx(nullptr);
  bool ok = rc == SRK_OK && !ctx.failed;
  if (capture) {
    if (ctx.capturing) {
      const cudaError_t ce = cudaStreamEndCapture(run, &e.g.graph);
      cudaError_t ie = cudaSuccess;
      ok = ok && ce == cudaSuccess && e.g.graph != nullptr && (ie = cudaGraphInstantiate(&e.g.exec, e.g.graph, 0)) == cudaSuccess;
      if (debug)
        fprintf(stderr, "[sessrec graph] capture: rc %d failed %d reason %d end-capture %d instantiate %d nodes %zu -> %s\n", rc,
                (int)ctx.failed, ctx.fail_reason, (int)ce, (int)ie, e.g.nodes.size(), ok ? "ok" : "discarded");
    } else {
      ok = false;
      if (debug) fprintf(stderr, "[sessrec graph] capture: the step never reached its capture boundary (rc %d)\n", rc);
    }
    if (!ok) {
      cudaGetLastError();
      e.g.destroy();
      e.bad = true;
    }
  } else {
    ok = ok && ctx.cursor == e.g.nodes.size();
    g_graph_node_updates += (long long)ctx.updated;
    if (debug && ok) fprintf(stderr, "[sessrec graph] replay: %zu of %zu nodes rewritten\n", ctx.updated, e.g.nodes.size());
    if (!ok) {
      ++g_graph_fallbacks;
      if (debug)
        fprintf(stderr, "[sessrec graph] update pass failed: rc %d reason %d cuda %d at launch %zu of %zu\n", rc, ctx.fail_reason,
                ctx.fail_cuda, ctx.cursor, e.g.nodes.size());
      if (++e.fails > 3) {                            // this configuration keeps changing its kernel sequence
        e.g.destroy();
        e.bad = true;
      }
    }
  }
  std::function<int()> tail = ctx.tail;
  if (ok) {
    SRK_CUDA(cudaGraphLaunch(e.g.exec, run));
    ++g_graph_launches;
  } else {
    // The forward half is already enqueued.  Re-run the step's bookkeeping with the forward launches dropped and launch
    // the backward half kernel by kernel.
    SrkLaunchCtx redo;
    redo.mode = whole ? SRK_LAUNCH_DIRECT : SRK_LAUNCH_SKIP;     // whole-step graph: nothing has been enqueued yet
    redo.mode_after_boundary = SRK_LAUNCH_DIRECT;
    srk_set_launch_ctx(&redo);
    rc = body_on_run();
    srk_set_launch_ctx(nullptr);
    if (rc != SRK_OK) return rc;
    tail = redo.tail;
  }
  if (tail) SRK_TRY(tail());                 // plain calls behind the graph: collective + optimizer of a data-parallel step
  SRK_TRY(ss0->order_always(run, caller));
  return SRK_OK;
}

// 0 never, 1 always, 2 auto (data-parallel steps: phase 1)
// whole-step graph (forward + backward + optimizer as ONE graph): single-rank steps with the optimizer inside (phase 0).
// SESSREC_GRAPH_WHOLE=1 always, =0 never; default: when the batch says it was padded to a fixed shape (header word 11).
int srk_step_want_graph(int phase);
static int g_whole_mode = -1;
bool srk_step_want_whole(int phase, int padded) {
  int& mode = g_whole_mode;
  if (mode < 0) {
    const char* e = getenv("SESSREC_GRAPH_WHOLE");
    mode = !e ? 2 : (e[0] == '0' ? 0 : 1);
  }
  (void)srk_step_want_graph(phase);      // reads SESSREC_GRAPH on first use
  if (phase != 0 || g_graphs_on == 0) return false;
  return mode == 1 || (mode == 2 && padded != 0);
}

// 0 = plain launches, 1 = replay the backward half as a graph, 2 = let the driver decide from the timing of the second step
int srk_step_want_graph(int phase) {
  if (g_graphs_on < 0) {
    const char* e = getenv("SESSREC_GRAPH");
    g_graphs_on = !e ? 2 : (e[0] == '0' ? 0 : 1);
  }
  if (g_graphs_on != 2) return g_graphs_on;
  // auto: replay pays off when the host is the bottleneck.  Data-parallel steps: several ranks share the CPU (measured on
  // 8 x B200: 0.83 -> 0.32 ms of enqueue per step, 5.7 M -> 8.3 M sessions/s) - always.  A single rank: depends on the host
  // (see srk_step_driver) - measured.
  return (phase == 1 || phase == 3) ? 1 : 2;
}

// phase: 0 = everything; 1 = zero_grad + forward + backward only (no Adam): lets the caller all-reduce the gradients;
// 2 = Adam only.
extern "C" int srk_msgifsr_train_step(const int* batch_dev, const int* batch_hdr_host, float* params, float* grads,
                                      const long long* slot_off_host, int V, int d, int L, float dropout_p, uint64_t seed,
                                      int use_umma, void* workspace, long long workspace_bytes, const float* one_dev,
                                      float* loss_out, int do_adam, float* exp_avg, float* exp_avg_sq, long long n_flat,
                                      const long long* seg_off_dev, const float* seg_decay_dev, int n_seg, float lr,
                                      float beta1, float beta2, float eps, int adam_step, float grad_scale, int phase,
                                      int head_chunks, void* stream) {
  cudaStream_t caller = (cudaStream_t)stream;
  if (phase == 2) {
    return srk_adam_step(params, grads, exp_avg, exp_avg_sq, n_flat, seg_off_dev, seg_decay_dev, n_seg, lr, beta1, beta2,
                         eps, adam_step, grad_scale, caller);
  }
  int dev = 0;
  SRK_CUDA(cudaGetDevice(&dev));
  const int has_edges = batch_hdr_host[REL_TAB + 2] > 0;
  const unsigned long long key = ((unsigned long long)(batch_hdr_host[1] & 0xFFFFF) << 44) | ((unsigned long long)(d & 0x3FF) << 34) |
                                 ((unsigned long long)(V & 0x3FFFF) << 16) | ((unsigned long long)(dev & 15) << 12) |
                                 ((unsigned long long)(L & 15) << 8) | ((unsigned long long)(use_umma & 7) << 5) | ((unsigned long long)(phase == 3) << 62) |
                                 ((unsigned long long)(dropout_p > 0.f) << 4) | ((unsigned long long)(phase & 1) << 3) | ((unsigned long long)(use_umma & 8) << 60) |
                                 ((unsigned long long)(do_adam != 0) << 2) | ((unsigned long long)has_edges << 1) |
                                 (unsigned long long)(head_chunks > 1);
  const bool whole = srk_step_want_whole(phase, batch_hdr_host[11]);
  return srk_step_driver(caller, key ^ ((unsigned long long)whole << 63), whole ? 1 : srk_step_want_graph(phase), [&](void* run) {
    return step_body(batch_dev, batch_hdr_host, params, grads, slot_off_host, V, d, L, dropout_p, seed, use_umma, workspace,
                     workspace_bytes, one_dev, loss_out, do_adam, exp_avg, exp_avg_sq, n_flat, seg_off_dev, seg_decay_dev, n_seg,
                     lr, beta1, beta2, eps, adam_step, grad_scale, phase, head_chunks, run);
  }, whole);
}

extern "C" int srk_set_graph_mode(int on) {
  g_graphs_on = on < 0 ? 0 : (on > 2 ? 2 : on);      // 0 never, 1 always, 2 auto (srk_step_want_graph)
  return SRK_OK;
}

extern "C" int srk_set_graph_whole(int mode) {
  g_whole_mode = mode < 0 ? 0 : (mode > 2 ? 2 : mode);      // 0 never, 1 always, 2 auto (batches padded to a fixed shape)
  return SRK_OK;
}

extern "C" int srk_graph_inject_mismatch(int nth_update) {
  g_inject_mismatch = nth_update;
  return SRK_OK;
}

extern "C" long long srk_graph_launches(void) { return g_graph_launches; }
extern "C" long long srk_graph_node_updates(void) { return g_graph_node_updates; }
extern "C" long long srk_graph_fallbacks(void) { return g_graph_fallbacks; }
