// Native training step for MSGIFSR of order K >= 1 with k-gram node types (utils/train.py:95-101 around msgifsr.py:158-317
// without --extra / --fusion): zero_grad, SemanticExpander (msgifsr.py:32-45), L heterogeneous MSHGNN layers over the intra_k /
// inter relations (msgifsr.py:47-91, gatconv.py:254-319), the multi-order attention read-out (msgifsr.py:124-155), the fused
// scoring + cross-entropy head, the whole backward and Adam - ONE host call that enqueues ~350 kernels at order 3.
//
// csrc/step.cu is the tuned order-1 path (graph replay, tensor-core projections).  This file is the general one: the same kernels
// as the staged Python composition (msgifsr.py of this package), enqueued from C++ (4 us of host time per launch instead of the
// 16 us of a Python + ctypes launch) and forked over the step's streams: at order 3 a layer has 14 independent convolutions
// (7 relations, each on the graph and on the reversed graph).  Convolutions that share a parameter module (`inter`) stay on one
// stream because their weight gradients accumulate into the same rows; every data-gradient term of a node type is written to its
// own slice and one pass per type adds them up, so no two streams ever accumulate into the same buffer.
#include <vector>

#include "launch.cuh"
#include "step_common.cuh"

namespace {

constexpr int H = SRK_HEADS;
constexpr int TYPE_TAB = 16, REL_TAB = 80, TAB_W = 16, MAXK = 4;

struct TypeView {
  int N, U, P;
  const int *iid, *seg, *last, *node2seg, *perm, *uoff, *uid, *last_row, *row_of;
};
struct RelView {
  int st, dt, M, code;
  const int *in_ptr, *in_src, *in_eid, *out_ptr, *out_dst, *out_eid;
};
struct KBatch {
  int B, K, R, nrel;
  const int *labels, *row_seg;
  TypeView t[MAXK + 1];        // 1-based
  RelView rel[3 * MAXK];
};

int parse(const int* dev, const int* hdr, KBatch& b) {
  SRK_REQUIRE(hdr[0] == 0x53524B31, "msgifsr step: not a SessionBatch buffer");
  SRK_REQUIRE(hdr[2] == 1, "msgifsr step: needs a ccs batch");
  b.B = hdr[1]; b.K = hdr[3]; b.nrel = hdr[5]; b.R = hdr[6];
  SRK_REQUIRE(b.K >= 1 && b.K <= MAXK && b.nrel <= 3 * MAXK - 2, "msgifsr step: order %d unsupported", b.K);
  b.labels = dev + hdr[7];
  b.row_seg = dev + hdr[8];
  for (int k = 1; k <= b.K; ++k) {
    const int* t = hdr + TYPE_TAB + TAB_W * (k - 1);
    TypeView& v = b.t[k];
    v.N = t[0]; v.U = t[8]; v.P = t[0] * k;
    v.iid = dev + t[1]; v.seg = dev + t[2]; v.last = dev + t[3]; v.node2seg = dev + t[4];
    v.perm = dev + t[5]; v.uoff = dev + t[6]; v.uid = dev + t[7]; v.last_row = dev + t[9]; v.row_of = dev + t[10];
  }
  for (int r = 0; r < b.nrel; ++r) {
    const int* t = hdr + REL_TAB + TAB_W * r;
    RelView& v = b.rel[r];
    v.st = t[0]; v.dt = t[1]; v.M = t[2]; v.code = t[12];
    v.in_ptr = dev + t[5]; v.in_src = dev + t[6]; v.in_eid = dev + t[7];
    v.out_ptr = dev + t[8]; v.out_dst = dev + t[9]; v.out_eid = dev + t[10];
  }
  return SRK_OK;
}

// C[M,N] (+)= X[M,K] W[N,K]^T (+ bias), rows of X optionally through a_idx
int linear_nt(cudaStream_t st, int M, int N, int K, const float* X, long long lda, const float* W, float* C, long long ldc,
              const int* a_idx = nullptr, const float* bias = nullptr) {
  return srk_gemm(M, N, K, X, lda, 1, W, 1, K, C, ldc, a_idx, nullptr, nullptr, bias, 1.f, 0, 0, st);
}
// C[M,N] (+)= A[M,K] Bm[K,N], rows of C optionally through c_idx
int mm_nn(cudaStream_t st, int M, int N, int K, const float* A, long long lda, const float* Bm, long long ldb, float* C,
          long long ldc, int accumulate, const int* c_idx = nullptr) {
  return srk_gemm(M, N, K, A, lda, 1, Bm, ldb, 1, C, ldc, nullptr, nullptr, c_idx, nullptr, 1.f, accumulate, 0, st);
}
// C[M,N] += A[K,M]^T Bm[K,N], rows of Bm optionally through b_idx
int mm_tn(cudaStream_t st, int M, int N, int K, const float* A, long long lda, const float* Bm, long long ldb, float* C,
          long long ldc, const int* b_idx = nullptr) {
  return srk_gemm(M, N, K, A, 1, lda, Bm, ldb, 1, C, ldc, nullptr, b_idx, nullptr, nullptr, 1.f, 1, 0, st);
}

// dropout-site numbering of one (layer, conv, relation instance): msgifsr.py::gat_slot of this package / oracle/models.py
int gat_slot(int layer, int conv, bool inter, int st, int dt, int K) {
  const int r = inter ? K + (st == 1 ? dt - 2 : (K - 1) + st - 2) : st - 1;
  return (layer * 2 + conv) * (3 * K) + r;
}

struct Inst {
  srk_gat_inst gi;
  int st, dt, M, pslot;                  // pslot: first of the module's 4 parameter slots (attn_l, attn_r, bias, fc.weight)
  float *Waug, *wr, *xs, *xd;
  srk_dropout dcs, dcd;
  cudaStream_t sq, sp;                   // sq: the stream of this convolution's parameter module; sp: its own (round robin)
  float *src_term, *dst_term;            // slices of the per-type gradient-term buffers (backward)
  float *dWaug, *dwr;
};
struct TypeRec {
  std::vector<Inst> inst;
  float *Hout, *rn, *dHpre;
  uint8_t* amax;
};
struct LayerRec {
  const float* in[MAXK + 1];
  TypeRec t[MAXK + 1];
  int normalize;
};
struct ExpRec {
  float *Xk, *out, *rn;
  float *hs[MAXK + 1], *gi[MAXK], *gh[MAXK];
  srk_dropout dc;
};

long long ws_floats(const int* hdr, int V, int d, int L) {
  const int B = hdr[1], K = hdr[3], nrel = hdr[5], R = hdr[6];
  const long long ldzel = (long long)H * d + H;
  long long Nmax = 1, Mmax = 1, fl = 0;
  for (int k = 1; k <= K; ++k) {
    const long long N = hdr[TYPE_TAB + TAB_W * (k - 1)];
    Nmax = N > Nmax ? N : Nmax;
    fl += N * k * d * 2 + (k + 3) * N * d + 2LL * k * N * 3 * d + 4 * N + 512;      // expander fwd + bwd
  }
  for (int r = 0; r < nrel; ++r) {
    const long long M = hdr[REL_TAB + TAB_W * r + 2];
    Mmax = M > Mmax ? M : Mmax;
  }
  const long long per_inst = 2 * (ldzel + H) * d + 8 * Nmax * d + 2 * Nmax * ldzel + 2 * Nmax * H + 2 * (Mmax + 1) * H + 2048;
  const long long per_type = (long long)B * d + 5 * Nmax * d + 2 * Nmax + 1024;
  fl += (long long)L * (2LL * nrel * per_inst + K * per_type);
  fl += 4LL * V * d + V + (long long)V * d + srk_flash_ce_part_floats(B, V) + (long long)srk_flash_ce_bwd_parts(B) * V * d + 1024;
  fl += 4LL * R * d + 2LL * R + 12LL * B * d + 8LL * B + (long long)K * Nmax * d + 4096;      // read-out, head, dH
  fl += srk_embed_scatter_ws_floats((int)(Nmax * K), d);
  return fl;
}

int body(const int* batch_dev, const int* hdr, float* params, float* grads, const long long* slot, int V, int d, int L,
         float dropout_p, uint64_t seed, int flags, void* workspace, long long workspace_bytes, const float* gseed_dev,
         float* loss_out, int do_adam, float* exp_avg, float* exp_avg_sq, long long n_flat, const long long* seg_off_dev,
         const float* seg_decay_dev, int n_seg, float lr, float beta1, float beta2, float eps, int adam_step, float grad_scale,
         int phase, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  KBatch b;
  SRK_TRY(parse(batch_dev, hdr, b));
  const int B = b.B, K = b.K;
  SRK_REQUIRE(L >= 1 && L <= 8, "msgifsr step: 1..8 layers");
  SRK_REQUIRE(d % 4 == 0 && srk_flash_ce_supported(d), "msgifsr step: embedding dim %d not supported by the fused head", d);
  SRK_REQUIRE((flags & 5) == 5, "msgifsr step (order K): needs the tensor-core fused head");
  const bool drop = dropout_p > 0.f;
  const int ldzel = H * d + H;
  Arena ar{reinterpret_cast<uint8_t*>(workspace), (size_t)workspace_bytes, 0, true};
  auto P = [&](int s) { return params + slot[s]; };
  auto G = [&](int s) { return grads + slot[s]; };
  auto dcfg = [&](uint32_t site) { srk_dropout c; c.p = dropout_p; c.site = site; c.seed = seed; return c; };
  // slot table (built by msgifsr.py::_slot_offsets_k): [0] embeddings.weight; layers: 1 + ((l*2 + conv)*(K+1) + e)*4 with e = k-1
  // for intra_k and K for `inter`; then the expander GRUs of k = 2..K (weight_ih, weight_hh, bias_ih, bias_hh); then
  // readout.fc_u.0.{weight,bias}, readout.fc_v.0.weight, readout.fc_e.0.weight, fc_sr.0.weight
  const int s_gru = 1 + L * 2 * (K + 1) * 4;
  const int s_ro = s_gru + (K - 1) * 4;
  float* E = P(0);

  SideStreams* ss = srk_side_streams();
  cudaStream_t s4 = ss ? ss->s[6] : st;
  constexpr int NP = 6;
  cudaStream_t pool[NP];                    // [0] = the step's main stream
  for (int i = 0; i < NP; ++i) pool[i] = ss ? ss->s[i] : st;
  auto order = [&](cudaStream_t from, cudaStream_t to) { return ss ? ss->order(from, to) : (int)SRK_OK; };
  auto fork_all = [&]() -> int { for (int i = 1; i < NP; ++i) SRK_TRY(order(st, pool[i])); return SRK_OK; };
  auto join_all = [&]() -> int { for (int i = 1; i < NP; ++i) SRK_TRY(order(pool[i], st)); return SRK_OK; };
  auto type_stream = [&](int k) { return pool[(k - 1) % NP]; };
  SRK_TRY(srk_step_begin());
  SRK_TRY(order(st, s4));
  SRK_TRY(srk_zero_async(grads, sizeof(float) * (size_t)n_flat, s4));

  // ---- forward ---------------------------------------------------------------------------------------------------------
  // nn.Embedding(max_norm=1): the rows a lookup touches are renormed at the lookup, all rows at the scoring head (msgifsr.py:
  // 247,276); rows are independent, so ONE pass over the catalog does both before the gathers read the table
  float *Ehat = ar.f((size_t)V * d), *enorm = ar.f(V);
  uint16_t* Ebh = reinterpret_cast<uint16_t*>(ar.raw((size_t)V * d * 2));
  uint16_t* Ebl = reinterpret_cast<uint16_t*>(ar.raw((size_t)V * d * 2));
  SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
  SRK_TRY(srk_catalog_prep_fwd(E, V, d, SRK_NORM_L2, 1.0f, Ehat, enorm, nullptr, nullptr, Ebh, Ebl, st));

  const float* feat[MAXK + 1] = {nullptr};
  const int N1 = b.t[1].N;
  float *X1 = ar.f((size_t)N1 * d), *rnX = ar.f(N1);
  SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
  srk_dropout dc_e = dcfg(SRK_SITE_EMBED + 1);
  SRK_TRY(srk_embed_gather_fwd(E, b.t[1].iid, N1, d, SRK_NORM_L2, drop ? &dc_e : nullptr, X1, rnX, nullptr, st));
  feat[1] = X1;

  // SemanticExpander (msgifsr.py:32-45): a k-gram node = mean of its k item rows + GRU over them, normalised
  ExpRec ex[MAXK + 1];
  for (int k = 2; k <= K; ++k) {
    const TypeView& t = b.t[k];
    ExpRec& e = ex[k];
    const int N = t.N, g0 = s_gru + (k - 2) * 4;
    cudaStream_t st = type_stream(k);                 // shadows the main stream inside this loop body
    SRK_TRY(order(pool[0], st));                      // the catalog pass has renormed the table
    e.dc = dcfg(SRK_SITE_EMBED + k);
    e.Xk = ar.f((size_t)N * k * d);
    e.hs[0] = ar.f((size_t)N * d);
    SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
    SRK_TRY(srk_embed_gather_fwd(E, t.iid, N * k, d, SRK_NORM_NONE, drop ? &e.dc : nullptr, e.Xk, nullptr, nullptr, st));
    SRK_TRY(srk_zero_async(e.hs[0], sizeof(float) * (size_t)N * d, st));
    for (int step = 0; step < k; ++step) {
      e.gi[step] = ar.f((size_t)N * 3 * d);
      e.gh[step] = ar.f((size_t)N * 3 * d);
      e.hs[step + 1] = ar.f((size_t)N * d);
      SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
      SRK_TRY(linear_nt(st, N, 3 * d, d, e.Xk + (size_t)step * d, (long long)k * d, P(g0), e.gi[step], 3 * d, nullptr, P(g0 + 2)));
      SRK_TRY(linear_nt(st, N, 3 * d, d, e.hs[step], d, P(g0 + 1), e.gh[step], 3 * d, nullptr, P(g0 + 3)));
      SRK_TRY(srk_gru_pointwise_fwd(e.gi[step], e.gh[step], e.hs[step], N, d, e.hs[step + 1], st));
    }
    e.out = ar.f((size_t)N * d);
    e.rn = ar.f(N);
    SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
    SRK_TRY(srk_expander_combine_fwd(e.Xk, e.hs[k], N, k, d, e.out, e.rn, st));
    feat[k] = e.out;
  }
  for (int k = 2; k <= K; ++k) SRK_TRY(order(type_stream(k), st));

  // MSHGNN layers: conv1 over every relation, conv2 over every reversed relation (msgifsr.py:70-91); a relation without edges
  // is skipped like HeteroGraphConv skips it
  srk_dropout dc_attn = dcfg(SRK_SITE_GAT_ATTN);
  std::vector<LayerRec> layers(L);
  for (int l = 0; l < L; ++l) {
    LayerRec& R = layers[l];
    R.normalize = (l == L - 1);
    for (int k = 1; k <= K; ++k) R.in[k] = feat[k];
    SRK_TRY(fork_all());                           // the layer input is complete on the main stream
    int rr = 0;
    for (int conv = 0; conv < 2; ++conv)
      for (int r = 0; r < b.nrel; ++r) {
        const RelView& rv = b.rel[r];
        if (rv.M == 0) continue;
        const bool inter = rv.code >= 100;
        Inst I;
        memset(&I.gi, 0, sizeof(I.gi));
        I.st = conv ? rv.dt : rv.st;
        I.dt = conv ? rv.st : rv.dt;
        I.M = rv.M;
        const int e = inter ? K : rv.code - 1;
        I.pslot = 1 + ((l * 2 + conv) * (K + 1) + e) * 4;
        // sq: one stream per parameter module - convolutions that share a module (`inter`: 2 (K - 1) of them) accumulate their
        // weight gradients into the same rows, those kernels are serialised there; everything else of a convolution runs on sp
        I.sq = pool[(inter ? 1 + conv : 3 + conv * K + e) % NP];
        I.sp = pool[rr++ % NP];
        cudaStream_t st = I.sp;                     // shadows the main stream for this convolution's chain
        const int gs = gat_slot(l, conv, inter, I.st, I.dt, K);
        const int Ns = b.t[I.st].N, Nd = b.t[I.dt].N;
        I.Waug = ar.f((size_t)ldzel * d);
        I.wr = ar.f((size_t)H * d);
        SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
        SRK_TRY(srk_gat_prep(P(I.pslot + 3), P(I.pslot), P(I.pslot + 1), d, I.Waug, I.wr, st));
        I.xs = const_cast<float*>(feat[I.st]);
        I.xd = const_cast<float*>(feat[I.dt]);
        if (drop) {
          I.dcs = dcfg(SRK_SITE_GAT_SRC + 4 * gs);
          I.dcd = dcfg(SRK_SITE_GAT_DST + 4 * gs);
          I.xs = ar.f((size_t)Ns * d);
          I.xd = ar.f((size_t)Nd * d);
          SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
          SRK_TRY(srk_dropout_apply(feat[I.st], I.xs, (long long)Ns * d, &I.dcs, 0, st));
          SRK_TRY(srk_dropout_apply(feat[I.dt], I.xd, (long long)Nd * d, &I.dcd, 0, st));
        }
        float *Zel = ar.f((size_t)Ns * ldzel), *er = ar.f((size_t)Nd * H), *att = ar.f((size_t)(rv.M + 1) * H);
        SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
        SRK_TRY(linear_nt(st, Ns, ldzel, d, I.xs, d, I.Waug, Zel, ldzel));
        SRK_TRY(linear_nt(st, Nd, H, d, I.xd, d, I.wr, er, H));
        srk_gat_inst& g = I.gi;
        if (conv == 0) {
          g.in_ptr = rv.in_ptr; g.in_src = rv.in_src; g.in_eid = rv.in_eid;
          g.out_ptr = rv.out_ptr; g.out_dst = rv.out_dst; g.out_eid = rv.out_eid;
        } else {            // reversed graph: the two CSRs swap roles
          g.in_ptr = rv.out_ptr; g.in_src = rv.out_dst; g.in_eid = rv.out_eid;
          g.out_ptr = rv.in_ptr; g.out_dst = rv.in_src; g.out_eid = rv.in_eid;
        }
        g.Zel = Zel; g.er = er; g.bias = P(I.pslot + 2); g.xdst = I.xd; g.att = att;
        g.n_src = Ns; g.n_dst = Nd; g.n_edges = rv.M;
        g.attn_site = SRK_SITE_GAT_ATTN + 4 * gs;
        R.t[I.dt].inst.push_back(I);
      }
    SRK_TRY(join_all());
    SRK_TRY(fork_all());
    for (int k = 1; k <= K; ++k) {
      TypeRec& T = R.t[k];
      const int N = b.t[k].N;
      cudaStream_t st = type_stream(k);            // the K aggregations run side by side
      SRK_REQUIRE((int)T.inst.size() <= SRK_MAX_GAT_INST, "msgifsr step: %d convolutions into one node type", (int)T.inst.size());
      float* segmean = ar.f((size_t)B * d);
      T.Hout = ar.f((size_t)N * d);
      T.rn = ar.f(N);
      T.amax = ar.raw((size_t)N * d);
      SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
      SRK_TRY(srk_segmean_fwd(feat[k], b.t[k].seg, B, d, segmean, st));
      std::vector<srk_gat_inst> arr;
      for (const Inst& I : T.inst) arr.push_back(I.gi);
      SRK_TRY(srk_gat_aggregate_fwd(arr.data(), (int)arr.size(), N, d, segmean, b.t[k].node2seg, drop ? &dc_attn : nullptr, R.normalize,
                                    T.Hout, T.rn, T.amax, st));
    }
    SRK_TRY(join_all());
    for (int k = 1; k <= K; ++k) feat[k] = R.t[k].Hout;
  }

  // read-out over the per-session concatenation of all orders' rows (msgifsr.py:127-146); without --fusion only order 1's
  // score is returned (msgifsr.py:316-317)
  const float* rows = feat[1];
  const int* seg = b.t[1].seg;
  const int* last_row = b.t[1].last;
  int Rn = N1;
  if (K > 1) {
    Rn = b.R;
    float* rw = ar.f((size_t)Rn * d);
    SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
    SRK_TRY(srk_zero_async(rw, sizeof(float) * (size_t)Rn * d, st));
    for (int k = 1; k <= K; ++k) SRK_TRY(srk_scatter_add_rows(feat[k], d, b.t[k].row_of, b.t[k].N, d, rw, st));
    rows = rw;
    seg = b.row_seg;
    last_row = b.t[1].last_row;
  }
  float *u = ar.f((size_t)Rn * d), *v = ar.f((size_t)B * d), *e = ar.f(Rn), *ms = ar.f(2 * (size_t)B);
  float *sr_in = ar.f(2 * (size_t)B * d), *s = ar.f((size_t)B * d), *shat = ar.f((size_t)B * d), *rn_s = ar.f(B);
  uint16_t* Sbh = reinterpret_cast<uint16_t*>(ar.raw((size_t)B * d * 2));
  uint16_t* Sbl = reinterpret_cast<uint16_t*>(ar.raw((size_t)B * d * 2));
  float *lse = ar.f(B), *nll = ar.f(B), *part = ar.f((size_t)srk_flash_ce_part_floats(B, V));
  SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
  SRK_TRY(linear_nt(st, Rn, d, d, rows, d, P(s_ro), u, d, nullptr, P(s_ro + 1)));
  SRK_TRY(linear_nt(st, B, d, d, feat[1], d, P(s_ro + 2), v, d, b.t[1].last, nullptr));
  SRK_TRY(srk_readout_fwd(rows, u, v, P(s_ro + 3), seg, last_row, B, d, 1, e, ms, sr_in, st));
  SRK_TRY(linear_nt(st, B, d, 2 * d, sr_in, 2 * d, P(s_ro + 4), s, d));
  SRK_TRY(srk_rownorm_split_fwd(s, d, B, d, SRK_NORM_L2, shat, d, rn_s, Sbh, Sbl, st));
  SRK_TRY(srk_flash_ce_fwd(B, V, d, Sbh, Sbl, d, Ebh, Ebl, d, 12.0f, b.labels, lse, nll, part, st));
  SRK_TRY(srk_mean(nll, B, loss_out, st));

  // ---- backward --------------------------------------------------------------------------------------------------------
  const int de_parts = srk_flash_ce_bwd_parts(B);
  float *dshat = ar.f((size_t)B * d), *dEhat = ar.f((size_t)de_parts * V * d);
  SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
  SRK_TRY(order(s4, st));                        // zero_grad
  SRK_TRY(srk_flash_ce_bwd(B, V, d, Sbh, Sbl, d, Ebh, Ebl, d, 12.0f, b.labels, lse, gseed_dev, dshat, dEhat, st));
  // the catalog-wide part of the table gradient runs beside the encoder backward; the scatter-adds wait for it
  SRK_TRY(order(st, s4));
  SRK_TRY(srk_catalog_prep_bwd(E, Ehat, enorm, dEhat, de_parts, V, d, SRK_NORM_L2, G(0), s4));

  float *ds = ar.f((size_t)B * d), *dsr_in = ar.f(2 * (size_t)B * d), *drows = ar.f((size_t)Rn * d);
  SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
  SRK_TRY(srk_rownorm_bwd(s, d, shat, d, rn_s, dshat, d, B, d, SRK_NORM_L2, ds, d, 0, st));
  SRK_TRY(mm_nn(st, B, 2 * d, d, ds, d, P(s_ro + 4), 2 * d, dsr_in, 2 * d, 0));
  SRK_TRY(mm_tn(st, d, 2 * d, B, ds, d, sr_in, 2 * d, G(s_ro + 4), 2 * d));
  SRK_TRY(srk_readout_bwd(rows, u, v, P(s_ro + 3), seg, last_row, e, ms, sr_in, dsr_in, B, d, 1, drows, G(s_ro + 3), st));
  SRK_TRY(mm_nn(st, Rn, d, d, u, d, P(s_ro), d, drows, d, 1));                  // u holds du
  SRK_TRY(mm_tn(st, d, d, Rn, u, d, rows, d, G(s_ro), d));
  SRK_TRY(srk_colsum(u, d, Rn, d, G(s_ro + 1), 1, st));
  float* dH[MAXK + 1] = {nullptr};
  if (K == 1) {
    dH[1] = drows;
  } else {
    for (int k = 1; k <= K; ++k) {
      dH[k] = ar.f((size_t)b.t[k].N * d);
      SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
      SRK_TRY(srk_gather_rows(drows, b.t[k].row_of, b.t[k].N, d, dH[k], d, st));
    }
  }
  SRK_TRY(mm_nn(st, B, d, d, v, d, P(s_ro + 2), d, dH[1], d, 1, b.t[1].last));      // v holds dv
  SRK_TRY(mm_tn(st, d, d, B, v, d, feat[1], d, G(s_ro + 2), d, b.t[1].last));

  for (int l = L - 1; l >= 0; --l) {
    LayerRec& R = layers[l];
    // d(layer input of type k) = segment-mean term + one term per convolution out of / into the type: every term gets its own
    // [N_k, d] slice of parts[k] and one pass adds them up
    float *dfeat[MAXK + 1] = {nullptr}, *parts[MAXK + 1] = {nullptr};
    int nterm[MAXK + 1] = {0};
    for (int k = 1; k <= K; ++k) nterm[k] = 1;
    for (int k = 1; k <= K; ++k)
      for (Inst& I : R.t[k].inst) {
        ++nterm[I.st];
        ++nterm[k];
      }
    for (int k = 1; k <= K; ++k) {
      const size_t nd = (size_t)b.t[k].N * d;
      dfeat[k] = ar.f(nd);
      parts[k] = ar.f(nd * nterm[k]);
      nterm[k] = 1;                                  // slice 0 = segment-mean term; re-counted while the slices are handed out
    }
    SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
    for (int k = 1; k <= K; ++k)
      for (Inst& I : R.t[k].inst) {
        I.src_term = parts[I.st] + (size_t)b.t[I.st].N * d * nterm[I.st]++;
        I.dst_term = parts[k] + (size_t)b.t[k].N * d * nterm[k]++;
      }
    SRK_TRY(fork_all());
    for (int k = 1; k <= K; ++k) {
      TypeRec& T = R.t[k];
      const int N = b.t[k].N;
      cudaStream_t st = type_stream(k);
      T.dHpre = ar.f((size_t)N * d);
      std::vector<srk_gat_inst> arr;
      for (Inst& I : T.inst) {
        I.gi.dedge = ar.f((size_t)(I.M + 1) * H);
        I.gi.der = ar.f((size_t)N * H);
        I.gi.dZel = ar.f((size_t)I.gi.n_src * ldzel);
        arr.push_back(I.gi);
      }
      SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
      SRK_TRY(srk_gat_aggregate_bwd_dst(arr.data(), (int)arr.size(), N, d, drop ? &dc_attn : nullptr, R.normalize, T.Hout, T.rn, T.amax,
                                        dH[k], T.dHpre, st));
      SRK_TRY(srk_segmean_bwd(T.dHpre, b.t[k].seg, B, d, parts[k], 0, st));
    }
    SRK_TRY(join_all());
    SRK_TRY(fork_all());
    for (int k = 1; k <= K; ++k) {
      TypeRec& T = R.t[k];
      const int N = b.t[k].N;
      for (Inst& I : T.inst) {
        const int Ns = I.gi.n_src;
        cudaStream_t st = I.sp;
        SRK_TRY(srk_gat_aggregate_bwd_src(&I.gi, d, drop ? &dc_attn : nullptr, T.dHpre, T.amax, st));
        float *dWaug = I.dWaug = ar.f((size_t)ldzel * d), *dwr = I.dwr = ar.f((size_t)H * d);
        SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
        SRK_TRY(srk_zero2_async(dWaug, sizeof(float) * (size_t)ldzel * d, dwr, sizeof(float) * (size_t)H * d, st));
        SRK_TRY(mm_tn(st, ldzel, d, Ns, I.gi.dZel, ldzel, I.xs, d, dWaug, d));
        SRK_TRY(mm_tn(st, H, d, N, I.gi.der, H, I.xd, d, dwr, d));
        // source-copy term: mask_s(dZel W_aug); destination-copy term: mask_d(residual + der w_r)
        float* tmp2 = ar.f((size_t)N * d);
        if (!drop) {
          SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
          SRK_TRY(mm_nn(st, Ns, d, ldzel, I.gi.dZel, ldzel, I.Waug, d, I.src_term, d, 0));
        } else {
          float* tmp = ar.f((size_t)Ns * d);
          SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
          SRK_TRY(mm_nn(st, Ns, d, ldzel, I.gi.dZel, ldzel, I.Waug, d, tmp, d, 0));
          SRK_TRY(srk_dropout_apply(tmp, I.src_term, (long long)Ns * d, &I.dcs, 0, st));
        }
        SRK_TRY(mm_nn(st, N, d, H, I.gi.der, H, I.wr, d, tmp2, d, 0));
        SRK_TRY(srk_dropout_apply_add(T.dHpre, tmp2, I.dst_term, (long long)N * d, drop ? &I.dcd : nullptr, st));
      }
    }
    SRK_TRY(join_all());
    SRK_TRY(fork_all());
    for (int k = 1; k <= K; ++k) {
      const long long nd = (long long)b.t[k].N * d;
      SRK_TRY(srk_sum_parts(parts[k], nd, nterm[k], nd, dfeat[k], 0, type_stream(k)));
    }
    // the kernels that accumulate into a module's gradient rows, one module per stream
    for (int k = 1; k <= K; ++k)
      for (Inst& I : R.t[k].inst) {
        SRK_TRY(srk_gat_bias_bwd(R.t[k].dHpre, R.t[k].amax, b.t[k].N, d, G(I.pslot + 2), I.sq));
        SRK_TRY(srk_gat_prep_bwd(P(I.pslot + 3), P(I.pslot), P(I.pslot + 1), I.dWaug, I.dwr, d, G(I.pslot + 3), G(I.pslot), G(I.pslot + 1),
                                 I.sq));
      }
    SRK_TRY(join_all());
    for (int k = 1; k <= K; ++k) dH[k] = dfeat[k];
  }

  float* dXk_of[MAXK + 1] = {nullptr};
  SRK_TRY(fork_all());
  for (int k = 2; k <= K; ++k) {
    const TypeView& t = b.t[k];
    ExpRec& x = ex[k];
    const int N = t.N, g0 = s_gru + (k - 2) * 4;
    cudaStream_t st = type_stream(k);
    float *dXk = ar.f((size_t)N * k * d), *dh = ar.f((size_t)N * d), *dprev = ar.f((size_t)N * d);
    SRK_REQUIRE(ar.ok, "msgifsr step: workspace too small");
    dXk_of[k] = dXk;
    SRK_TRY(srk_expander_combine_bwd(x.out, x.rn, dH[k], N, k, d, dh, dXk, st));
    for (int step = k - 1; step >= 0; --step) {
      float *gi = x.gi[step], *gh = x.gh[step];
      SRK_TRY(srk_gru_pointwise_bwd(gi, gh, x.hs[step], dh, N, d, dprev, 0, st));      // gi, gh <- their gradients
      SRK_TRY(mm_tn(st, 3 * d, d, N, gi, 3 * d, x.Xk + (size_t)step * d, (long long)k * d, G(g0), d));
      SRK_TRY(srk_colsum(gi, 3 * d, N, 3 * d, G(g0 + 2), 1, st));
      SRK_TRY(mm_nn(st, N, d, 3 * d, gi, 3 * d, P(g0), d, dXk + (size_t)step * d, (long long)k * d, 1));
      SRK_TRY(mm_tn(st, 3 * d, d, N, gh, 3 * d, x.hs[step], d, G(g0 + 1), d));
      SRK_TRY(srk_colsum(gh, 3 * d, N, 3 * d, G(g0 + 3), 1, st));
      SRK_TRY(mm_nn(st, N, d, 3 * d, gh, 3 * d, P(g0 + 1), d, dprev, d, 1));
      float* t2 = dh; dh = dprev; dprev = t2;
    }
  }
  // the scatter-adds update table rows in place (and the catalog backward wrote the same rows): one after the other on the
  // main stream
  SRK_TRY(order(s4, st));
  SRK_TRY(srk_embed_scatter_bwd_ws(E, b.t[1].iid, b.t[1].perm, b.t[1].uoff, b.t[1].uid, b.t[1].U, b.t[1].P, d, SRK_NORM_L2,
                                   drop ? &dc_e : nullptr, rnX, dH[1], nullptr, G(0), nullptr, st));
  SRK_TRY(join_all());
  for (int k = 2; k <= K; ++k) {
    const TypeView& t = b.t[k];
    SRK_TRY(srk_embed_scatter_bwd_ws(E, t.iid, t.perm, t.uoff, t.uid, t.U, t.P, d, SRK_NORM_NONE, drop ? &ex[k].dc : nullptr, nullptr,
                                     dXk_of[k], nullptr, G(0), nullptr, st));
  }
  if (phase == 0 && do_adam)
    SRK_TRY(srk_adam_step(params, grads, exp_avg, exp_avg_sq, n_flat, seg_off_dev, seg_decay_dev, n_seg, lr, beta1, beta2, eps,
                          adam_step, grad_scale, st));
  return SRK_OK;
}

}  // namespace

extern "C" long long srk_msgifsr_k_workspace_bytes(const int* batch_hdr_host, int V, int d, int L) {
  return ws_floats(batch_hdr_host, V, d, L) * 5 + (1 << 20);      // floats -> bytes with 25 % head-room + alignment slack
}

// phase: 0 = everything; 1 = zero_grad + forward + backward only (the caller all-reduces the gradients); 2 = Adam only.
// flags: bit 0 tensor cores, bit 2 fused scoring + CE head (both required).
extern "C" int srk_msgifsr_k_train_step(const int* batch_dev, const int* batch_hdr_host, float* params, float* grads,
                                        const long long* slot_off_host, int n_slots, int V, int d, int L, float dropout_p,
                                        uint64_t seed, int flags, void* workspace, long long workspace_bytes,
                                        const float* gseed_dev, float* loss_out, int do_adam, float* exp_avg, float* exp_avg_sq,
                                        long long n_flat, const long long* seg_off_dev, const float* seg_decay_dev, int n_seg,
                                        float lr, float beta1, float beta2, float eps, int adam_step, float grad_scale, int phase,
                                        void* stream) {
  cudaStream_t caller = (cudaStream_t)stream;
  if (phase == 2)
    return srk_adam_step(params, grads, exp_avg, exp_avg_sq, n_flat, seg_off_dev, seg_decay_dev, n_seg, lr, beta1, beta2, eps,
                         adam_step, grad_scale, caller);
  const int K = batch_hdr_host[3];
  SRK_REQUIRE(K >= 1 && K <= MAXK, "msgifsr step: order %d unsupported", K);
  SRK_REQUIRE(n_slots == 1 + L * 2 * (K + 1) * 4 + (K - 1) * 4 + 5, "msgifsr step: slot table of %d entries does not fit order %d, %d layers",
              n_slots, K, L);
  return srk_step_driver(caller, 0, 0, [&](void* run) {
    return body(batch_dev, batch_hdr_host, params, grads, slot_off_host, V, d, L, dropout_p, seed, flags, workspace, workspace_bytes,
                gseed_dev, loss_out, do_adam, exp_avg, exp_avg_sq, n_flat, seg_off_dev, seg_decay_dev, n_seg, lr, beta1, beta2, eps,
                adam_step, grad_scale, phase, run);
  });
}
