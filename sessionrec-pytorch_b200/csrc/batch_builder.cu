// Native batch builder (host code): sessions -> batched session graphs in the flat int32 layout the kernels
// consume.  Re-designed equivalent of the reference's per-session Python collate
// (src/utils/data/collate.py:61-85 seq_to_session_graph, :87-217 seq_to_ccs_graph, :219-256 collate_fn +
// dgl.batch): sorted-unique item nodes, first-occurrence-ordered unique k-gram nodes, de-duplicated
// consecutive-pair edges (with multiplicities for the session graph), batch-global ids, plus what only the
// device path needs: CSR by destination and by source, node->session map, item-sorted scatter permutation.
//
// Buffer layout (int32 words), all offsets relative to the buffer start:
//   [0] magic 'SRK1'  [1] B  [2] kind  [3] K  [4] words used  [5] n_rel  [6] R (rows of all types)
//   [7] off_labels  [8] off_row_seg  [9] off_row_type  [10] off_row_node
//   type table  at word 16 + 16*(k-1):  N, off_iid, off_seg, off_last, off_node2seg, off_perm, off_uoff,
//                                        off_uid, U, off_last_row, off_row_of (mixed-row index of every node)
//   rel table   at word 80 + 16*r:      st, dt, M, off_src, off_dst, off_in_ptr, off_in_src, off_in_eid,
//                                        off_out_ptr, off_out_dst, off_out_eid, off_w (float bits, kind 0), code
//   rel order: intra1..intraK, then for k = 2..K: inter1_k, interk_1.  code = k for intra_k, 100+k for
//   inter1_k, 200+k for interk_1.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

constexpr int HDR = 16, TYPE_TAB = 16, REL_TAB = 80, TAB_W = 16, MAXK = 4, MAXREL = 3 * MAXK - 2;
constexpr int DATA0 = REL_TAB + TAB_W * MAXREL;
constexpr int SMALL_L = 32;   // sessions up to this many clicks take the branch-free path of build()

struct Rel {
  int st, dt, code;
  std::vector<int> src, dst, w;
};

struct Work {
  int B, kind, K;
  std::vector<int> iid[MAXK], seg[MAXK], last[MAXK];
  std::vector<Rel> rels;
};

inline void add_unique_pair(std::vector<int>& ps, std::vector<int>& pd, std::vector<int>* pw, size_t base, int a, int b) {
  for (size_t i = base; i < ps.size(); ++i)
    if (ps[i] == a && pd[i] == b) {
      if (pw) (*pw)[i] += 1;
      return;
    }
  ps.push_back(a);
  pd.push_back(b);
  if (pw) pw->push_back(1);
}

int build(const int* items, const int* offs, int B, int kind, int K, Work& wk) {
  wk.B = B; wk.kind = kind; wk.K = K;
  // the work area is thread-local and keeps its capacity from batch to batch: no allocation in steady state
  wk.rels.resize(3 * K - 2);
  auto reset = [](Rel& r, int st, int dt, int code) {
    r.st = st; r.dt = dt; r.code = code;
    r.src.clear(); r.dst.clear(); r.w.clear();
  };
  for (int k = 1; k <= K; ++k) reset(wk.rels[k - 1], k, k, k);
  for (int k = 2; k <= K; ++k) {
    reset(wk.rels[K + 2 * (k - 2)], 1, k, 100 + k);
    reset(wk.rels[K + 2 * (k - 2) + 1], k, 1, 200 + k);
  }
  for (int k = 0; k < K; ++k) {
    wk.iid[k].clear(); wk.last[k].clear();
    wk.seg[k].assign(1, 0);
  }
  static thread_local std::vector<int> uniq, s, gid[MAXK];
  for (int b = 0; b < B; ++b) {
    const int* seq = items + offs[b];
    const int L = offs[b + 1] - offs[b];
    SRK_REQUIRE(L >= 1, "batch: session %d is empty", b);
    s.resize(L);
    const bool small = L <= SMALL_L;
    if (small) {
      // Short session (the common case): rank every click among the session's distinct items by counting, with no
      // data-dependent branch - on real sessions the compare branches of sort / unique / lower_bound mispredict every
      // other time and cost more than the O(L^2) counts.
      int first[SMALL_L];
      int nu = 0;
      for (int i = 0; i < L; ++i) {
        const int x = seq[i];
        int seen = 0;
        for (int j = 0; j < i; ++j) seen += (seq[j] == x);
        first[i] = (seen == 0);
        nu += first[i];
      }
      uniq.resize(nu);
      for (int i = 0; i < L; ++i) {
        const int x = seq[i];
        int less = 0;
        for (int j = 0; j < L; ++j) less += (seq[j] < x) & first[j];
        s[i] = less;
        uniq[less] = x;
      }
    } else {
      uniq.assign(seq, seq + L);
      std::sort(uniq.begin(), uniq.end());
      uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
      for (int i = 0; i < L; ++i) s[i] = (int)(std::lower_bound(uniq.begin(), uniq.end(), seq[i]) - uniq.begin());
    }
    const int base1 = wk.seg[0].back();
    wk.iid[0].insert(wk.iid[0].end(), uniq.begin(), uniq.end());
    wk.seg[0].push_back(base1 + (int)uniq.size());
    wk.last[0].push_back(base1 + s[L - 1]);
    gid[0].assign(s.begin(), s.end());
    // k-gram node types
    for (int k = 2; k <= K; ++k) {
      std::vector<int>& g = gid[k - 1];
      const int ng = L - k + 1;
      g.assign(ng > 0 ? ng : 0, 0);
      const int basek = wk.seg[k - 1].back();
      std::vector<int>& rows = wk.iid[k - 1];
      const size_t row0 = rows.size();
      int cnt = 0;
      for (int j = 0; j < ng; ++j) {
        int found = -1;
        for (int c = 0; c < cnt && found < 0; ++c)
          if (std::equal(seq + j, seq + j + k, rows.begin() + row0 + (size_t)c * k)) found = c;
        if (found < 0) {
          rows.insert(rows.end(), seq + j, seq + j + k);
          found = cnt++;
        }
        g[j] = found;
      }
      if (cnt == 0) {                       // shorter than k: one dummy node made of the smallest item id
        rows.insert(rows.end(), k, uniq[0]);
        cnt = 1;
        wk.last[k - 1].push_back(basek);
      } else {
        wk.last[k - 1].push_back(basek + g[ng - 1]);
      }
      wk.seg[k - 1].push_back(basek + cnt);
    }
    // relations
    for (int k = 1; k <= K; ++k) {
      Rel& r = wk.rels[k - 1];
      const std::vector<int>& g = gid[k - 1];
      const int basek = wk.seg[k - 1][b];
      const size_t e0 = r.src.size();
      if (k == 1 && small) {
        // consecutive-click pairs of a short session: one small key per pair, duplicates found by counting
        int key[SMALL_L];
        const int ne = L - 1;
        for (int i = 0; i < ne; ++i) key[i] = g[i] * SMALL_L + g[i + 1];
        for (int i = 0; i < ne; ++i) {
          int before = 0, total = 0;
          for (int j = 0; j < i; ++j) before += (key[j] == key[i]);
          for (int j = 0; j < ne; ++j) total += (key[j] == key[i]);
          if (before == 0) {                  // first occurrence keeps the edge (and, for the session graph, the count)
            r.src.push_back(basek + g[i]);
            r.dst.push_back(basek + g[i + 1]);
            if (kind == 0) r.w.push_back(total);
          }
        }
      } else {
        for (int i = 0; i + 1 < (int)g.size(); ++i)
          add_unique_pair(r.src, r.dst, kind == 0 ? &r.w : nullptr, e0, basek + g[i], basek + g[i + 1]);
      }
      if (kind == 0 && r.src.size() == e0) {   // single click: self-loop, weight 1
        r.src.push_back(basek); r.dst.push_back(basek); r.w.push_back(1);
      }
    }
    for (int k = 2; k <= K; ++k) {
      Rel& f = wk.rels[K + 2 * (k - 2)];
      Rel& r = wk.rels[K + 2 * (k - 2) + 1];
      const std::vector<int>& g = gid[k - 1];
      const int basek = wk.seg[k - 1][b];
      const size_t f0 = f.src.size(), r0 = r.src.size();
      for (int i = 0; i < L - k; ++i) {
        add_unique_pair(f.src, f.dst, nullptr, f0, base1 + s[i], basek + g[i + 1]);
        add_unique_pair(r.src, r.dst, nullptr, r0, basek + g[i], base1 + s[i + k]);
      }
    }
  }
  return SRK_OK;
}

struct Writer {
  int* out;
  long long cap, pos;
  bool ok;
  int put(const int* p, size_t n) {
    long long at = pos;
    if (pos + (long long)n > cap) { ok = false; pos += (long long)n; return (int)at; }
    if (n) memcpy(out + pos, p, n * sizeof(int));
    pos += (long long)n;
    return (int)at;
  }
  int put(const std::vector<int>& v) { return put(v.data(), v.size()); }
};

void csr(const std::vector<int>& key, const std::vector<int>& other, int n, std::vector<int>& ptr,
         std::vector<int>& nbr, std::vector<int>& eid) {
  const int M = (int)key.size();
  ptr.assign(n + 1, 0);
  for (int e = 0; e < M; ++e) ptr[key[e] + 1]++;
  for (int i = 0; i < n; ++i) ptr[i + 1] += ptr[i];
  nbr.resize(M); eid.resize(M);
  std::vector<int> cur(ptr.begin(), ptr.end() - 1);
  for (int e = 0; e < M; ++e) {             // stable: edge-id order inside a node (DGL mailbox order)
    int p = cur[key[e]]++;
    nbr[p] = other[e];
    eid[p] = e;
  }
}

// perm = positions 0..P-1 ordered by ids[] (stable).  LSD radix on 11-bit digits: a batch holds a few thousand
// positions, where std::stable_sort with an indirect comparator costs ~100 ns per element.
void order_by_id(const std::vector<int>& ids, std::vector<int>& perm, std::vector<int>& tmp) {
  const int P = (int)ids.size();
  perm.resize(P);
  tmp.resize(P);
  int mx = 0, mn = 0;
  for (int i = 0; i < P; ++i) {
    perm[i] = i;
    mx = std::max(mx, ids[i]);
    mn = std::min(mn, ids[i]);
  }
  if (mn < 0) {                            // not an item id; keep the ordering defined anyway
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int c) { return ids[a] < ids[c]; });
    return;
  }
  int* a = perm.data();
  int* b = tmp.data();
  for (int shift = 0; shift < 31 && (mx >> shift) > 0; shift += 11) {
    int cnt[2049] = {0};
    for (int i = 0; i < P; ++i) cnt[((ids[a[i]] >> shift) & 2047) + 1]++;
    for (int i = 0; i < 2048; ++i) cnt[i + 1] += cnt[i];
    for (int i = 0; i < P; ++i) {
      const int p = a[i];
      b[cnt[(ids[p] >> shift) & 2047]++] = p;
    }
    std::swap(a, b);
  }
  if (a != perm.data()) memcpy(perm.data(), a, sizeof(int) * P);
}

long long emit(const Work& wk, const int* labels, int* out, long long cap) {
  Writer w{out, cap, DATA0, true};
  std::vector<int> hdr(DATA0, 0);
  const int B = wk.B, K = wk.K;
  hdr[0] = 0x53524B31; hdr[1] = B; hdr[2] = wk.kind; hdr[3] = K; hdr[5] = (int)wk.rels.size();
  std::vector<int> lab(labels, labels + B);
  hdr[7] = w.put(lab);
  // mixed readout rows: per session its s1 nodes, then s2 nodes, ... (msgifsr.py:127-138)
  std::vector<int> row_seg(1, 0), row_type, row_node, row_of[MAXK];
  for (int k = 0; k < K; ++k) row_of[k].assign(wk.iid[k].size() / (k + 1), 0);
  for (int b = 0; b < B; ++b) {
    for (int k = 0; k < K; ++k)
      for (int n = wk.seg[k][b]; n < wk.seg[k][b + 1]; ++n) {
        row_of[k][n] = (int)row_type.size();
        row_type.push_back(k + 1);
        row_node.push_back(n);
      }
    row_seg.push_back((int)row_type.size());
  }
  hdr[6] = (int)row_type.size();
  hdr[8] = w.put(row_seg); hdr[9] = w.put(row_type); hdr[10] = w.put(row_node);
  for (int k = 0; k < K; ++k) {
    int* t = hdr.data() + TYPE_TAB + TAB_W * k;
    const int kk = k + 1, N = (int)wk.iid[k].size() / kk;
    t[0] = N;
    t[1] = w.put(wk.iid[k]);
    t[2] = w.put(wk.seg[k]);
    t[3] = w.put(wk.last[k]);
    std::vector<int> n2s(N);
    for (int b = 0; b < B; ++b)
      for (int n = wk.seg[k][b]; n < wk.seg[k][b + 1]; ++n) n2s[n] = b;
    t[4] = w.put(n2s);
    // scatter permutation: gather positions sorted by item id (stable), distinct ids and their ranges
    const int P = N * kk;
    const std::vector<int>& ids = wk.iid[k];
    static thread_local std::vector<int> perm, perm_tmp;
    order_by_id(ids, perm, perm_tmp);
    std::vector<int> uoff, uid;
    for (int i = 0; i < P; ++i)
      if (i == 0 || ids[perm[i]] != ids[perm[i - 1]]) {
        uoff.push_back(i);
        uid.push_back(ids[perm[i]]);
      }
    uoff.push_back(P);
    t[5] = w.put(perm); t[6] = w.put(uoff); t[7] = w.put(uid);
    t[8] = (int)uid.size();
    std::vector<int> last_row(B);
    for (int b = 0; b < B; ++b) last_row[b] = row_of[k][wk.last[k][b]];
    t[9] = w.put(last_row);
    t[10] = w.put(row_of[k]);
  }
  std::vector<int> ptr, nbr, eid;
  for (size_t r = 0; r < wk.rels.size(); ++r) {
    const Rel& R = wk.rels[r];
    int* t = hdr.data() + REL_TAB + TAB_W * (int)r;
    const int Ns = (int)wk.iid[R.st - 1].size() / R.st, Nd = (int)wk.iid[R.dt - 1].size() / R.dt;
    t[0] = R.st; t[1] = R.dt; t[2] = (int)R.src.size();
    t[3] = w.put(R.src); t[4] = w.put(R.dst);
    csr(R.dst, R.src, Nd, ptr, nbr, eid);
    t[5] = w.put(ptr); t[6] = w.put(nbr); t[7] = w.put(eid);
    csr(R.src, R.dst, Ns, ptr, nbr, eid);
    t[8] = w.put(ptr); t[9] = w.put(nbr); t[10] = w.put(eid);
    if (wk.kind == 0) {
      std::vector<int> wf(R.w.size());
      for (size_t i = 0; i < R.w.size(); ++i) {
        float f = (float)R.w[i];
        memcpy(&wf[i], &f, 4);
      }
      t[11] = w.put(wf);
    }
    t[12] = R.code;
  }
  hdr[4] = (int)w.pos;
  if (!w.ok) return -w.pos;                // caller's buffer too small: -needed
  memcpy(out, hdr.data(), sizeof(int) * DATA0);
  return w.pos;
}

}  // namespace

extern "C" long long srk_batch_size(const int* items_host, const int* offs_host, int B, int kind, int order) {
  (void)items_host;
  const int K = kind == 0 ? 1 : order;
  long long T = offs_host[B] - offs_host[0];
  // nodes <= T per type (+B dummies), gram rows carry k items, edges <= T per relation (+B self-loops)
  long long per_type = 0;
  for (int k = 1; k <= K; ++k) per_type += (T + B) * (long long)(4 * k + 3) + 4LL * (B + 1);
  long long per_rel = 8LL * (T + B) + 2LL * (T + B + 1);
  return DATA0 + B + 3LL * (T + B) * K + (B + 1) + per_type + per_rel * (3LL * K - 2) + 64;
}

extern "C" long long srk_batch_build(const int* items_host, const int* offs_host, const int* labels_host, int B, int kind,
                                     int order, int* out_host, long long out_words) {
  if (B <= 0 || (kind != 0 && kind != 1)) {
    srk_set_error("batch_build: bad B=%d / kind=%d", B, kind);
    return SRK_ERR_INVALID;
  }
  const int K = kind == 0 ? 1 : order;
  if (K < 1 || K > MAXK) {
    srk_set_error("batch_build: order %d unsupported (1..%d)", K, MAXK);
    return SRK_ERR_UNSUPPORTED;
  }
  static thread_local Work wk;
  int rc = build(items_host, offs_host, B, kind, K, wk);
  if (rc != SRK_OK) return rc;
  long long used = emit(wk, labels_host, out_host, out_words);
  if (used < 0) {
    srk_set_error("batch_build: buffer of %lld words too small (need %lld)", out_words, -used);
    return SRK_ERR_INVALID;
  }
  return used;
}
