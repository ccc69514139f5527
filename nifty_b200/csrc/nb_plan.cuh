// nifty_b200 -- transform plan: grid geometry, |k| mode bins, twiddles, scratch, pass launchers.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>
#include "nb_backend.cuh"
#include "nb_passes.cuh"
#include "nb_passes2.cuh"
#include "nb_amp.cuh"

namespace nb {

struct PassCfg { int lg_n = 0, lg_R = 0, pitch = 1, grid = 1, block = 64, ahead = 0; size_t smem = 0; };

inline int ilog2(int64_t v) { int l = 0; while ((int64_t(1) << l) < v) ++l; return l; }
inline bool is_pow2(int64_t v) { return v >= 1 && (v & (v - 1)) == 0; }

// host-side grid description shared by Plan<float> and Plan<double>
struct GridInfo {
  int ndim = 0;
  int64_t shape[3] = {1, 1, 1};
  double dist[3] = {1, 1, 1};
  int n0 = 1, nm = 1, nl = 1;      // internal axes: first, middle (1 unless 3-D), last (contiguous)
  int h0 = 0, hm = 0, hl = 0;
  bool three = false;
  int64_t N = 1;
  double V = 1;
  int K = 0;
  std::vector<int> idxf;           // [h0+1][hm+1][hl+1] bin of the folded mode
  std::vector<double> um, rel, logvol;
  std::vector<int64_t> mult;
  std::vector<int> w_order, w_offs;   // CSR of W positions per bin
  int64_t nW = 0;

  // Mode binning; follows the arithmetic of correlated_field.py:134-176 (mode lengths) and :55-67
  // (unique with 1e-12 relative merge, mid-point binning) on the folded index range only: every
  // |k| value of the full grid occurs there, with multiplicity prod_i (1 or 2).
  int build(int ndim_, const int64_t* shp, const double* dst, bool want_csr = true) {
    ndim = ndim_;
    if (ndim < 1 || ndim > 3) return fail("nb200: only 1-, 2- and 3-dimensional grids are supported");
    N = 1; V = 1;
    for (int i = 0; i < ndim; ++i) {
      shape[i] = shp[i]; dist[i] = dst[i];
      if (!is_pow2(shp[i]) || shp[i] < 2)
        return fail("nb200: every grid extent must be a power of two >= 2 on the sm_100a path (got " + std::to_string(shp[i]) + ")");
      if (shp[i] > (1 << 14)) return fail("nb200: grid extent too large for one shared-memory line");
      if (!(dst[i] > 0)) return fail("nb200: distances must be positive");
      N *= shp[i]; V *= (double)shp[i] * dst[i];
    }
    if (ndim == 1) { n0 = 1; nm = 1; nl = (int)shp[0]; }
    else if (ndim == 2) { n0 = (int)shp[0]; nm = 1; nl = (int)shp[1]; }
    else { n0 = (int)shp[0]; nm = (int)shp[1]; nl = (int)shp[2]; three = true; }
    h0 = n0 / 2; hm = nm / 2; hl = nl / 2;
    // per-axis folded lengths, user axis order
    std::vector<double> ax[3];
    for (int i = 0; i < ndim; ++i) {
      double step = 1.0 / ((double)shp[i] * dst[i]);
      ax[i].resize(shp[i] / 2 + 1);
      for (int64_t k = 0; k <= shp[i] / 2; ++k) ax[i][k] = (double)k * step;
    }
    const size_t nf = (size_t)(h0 + 1) * (hm + 1) * (hl + 1);
    std::vector<double> len(nf);
    size_t q = 0;
    for (int f0 = 0; f0 <= h0; ++f0)
      for (int fm = 0; fm <= hm; ++fm)
        for (int fl = 0; fl <= hl; ++fl, ++q) {
          double v;
          if (ndim == 1) v = ax[0][fl];
          else if (ndim == 2) { double a = ax[0][f0], b = ax[1][fl]; v = std::sqrt(a * a + b * b); }
          else { double a = ax[0][f0], b = ax[1][fm], c = ax[2][fl]; v = a * a; v = v + b * b; v = v + c * c; v = std::sqrt(v); }
          len[q] = v;
        }
    std::vector<double> srt(len);
    std::sort(srt.begin(), srt.end());
    srt.erase(std::unique(srt.begin(), srt.end()), srt.end());
    const double tol = 1e-12 * srt.back();
    um.clear();
    for (size_t i = 0; i < srt.size(); ++i) {
      double nxt = (i + 1 < srt.size()) ? srt[i + 1] : 2.0 * srt.back();
      if (nxt - srt[i] > tol) um.push_back(srt[i]);
    }
    K = (int)um.size();
    if (K < 1) return fail("invalid harmonic mode(s) encountered");
    std::vector<double> bounds(K > 0 ? K - 1 : 0);
    for (int i = 0; i + 1 < K; ++i) bounds[i] = 0.5 * (um[i] + um[i + 1]);
    idxf.resize(nf);
    mult.assign(K, 0);
    q = 0;
    for (int f0 = 0; f0 <= h0; ++f0)
      for (int fm = 0; fm <= hm; ++fm)
        for (int fl = 0; fl <= hl; ++fl, ++q) {
          int b = (int)(std::lower_bound(bounds.begin(), bounds.end(), len[q]) - bounds.begin());
          idxf[q] = b;
          int64_t w = 1;
          if (f0 != 0 && 2 * f0 != n0) w *= 2;
          if (fm != 0 && 2 * fm != nm) w *= 2;
          if (fl != 0 && 2 * fl != nl) w *= 2;
          mult[b] += w;
        }
    for (int b = 0; b < K; ++b)
      if (mult[b] == 0) return fail("invalid harmonic mode(s) encountered");
    // make_grid / _log_modes (correlated_field.py:228-265)
    rel = um;
    for (int b = 1; b < K; ++b) rel[b] = std::log(rel[b]);
    if (K > 1) { double r1 = rel[1]; for (int b = 1; b < K; ++b) rel[b] -= r1; }
    logvol.clear();
    for (int b = 2; b < K; ++b) logvol.push_back(rel[b] - rel[b - 1]);
    // CSR of the folded partial-sum array W[l = a*nm + km][x], a in [0,h0], km in [0,nm), x in [0,hl]
    nW = (int64_t)(h0 + 1) * nm * (hl + 1);
    if (!want_csr) return 0;          // slab-decomposed plans build the CSR of their local planes only
    if (nW >= (int64_t(1) << 31)) return fail("nb200: grid too large for 32-bit mode-bin indices");
    w_offs.assign(K + 1, 0);
    auto wbin = [&](int64_t p) {
      int x = (int)(p % (hl + 1)); int64_t l = p / (hl + 1);
      int km = (int)(l % nm), a = (int)(l / nm);
      int fm = km <= nm - km ? km : nm - km;
      return idxf[((size_t)a * (hm + 1) + fm) * (hl + 1) + x];
    };
    for (int64_t p = 0; p < nW; ++p) w_offs[wbin(p) + 1]++;
    for (int b = 0; b < K; ++b) w_offs[b + 1] += w_offs[b];
    w_order.resize(nW);
    std::vector<int> cur(w_offs.begin(), w_offs.end() - 1);
    for (int64_t p = 0; p < nW; ++p) w_order[cur[wbin(p)]++] = (int)p;
    return 0;
  }
};

// Slab ownership of one axis of extent n over P ranks (3-D grids too large for one GPU).  The half range
// [0, n/2] is cut into P contiguous pieces of c = n/(2P) indices (the last piece also takes n/2); a rank
// owns its piece ("A" planes) and the mirrors n - a of those planes ("B" planes, a = 0 and n/2 have none),
// so the Hermitian partner of every line it transforms is local.  Local order: A planes ascending, then the
// B planes in the order of their partners; the count is padded to a multiple of 16 with zero planes.
struct AxisDist {
  int n = 0, P = 1, c = 0;
  std::vector<int> a0, cA, cB, skip, npad;   // per rank
  std::vector<int> owner, lidx;              // per global index
  void build(int n_, int P_) {
    n = n_; P = P_; c = n / (2 * P);
    a0.resize(P); cA.resize(P); cB.resize(P); skip.resize(P); npad.resize(P);
    owner.assign(n, -1); lidx.assign(n, -1);
    for (int r = 0; r < P; ++r) {
      a0[r] = r * c; cA[r] = c + (r == P - 1 ? 1 : 0); skip[r] = (r == 0) ? 1 : 0; cB[r] = c - skip[r];
      npad[r] = ((cA[r] + cB[r] + 15) / 16) * 16;
      for (int i = 0; i < cA[r]; ++i) {
        int a = a0[r] + i;
        owner[a] = r; lidx[a] = i;
        if (a != 0 && 2 * a != n) { owner[n - a] = r; lidx[n - a] = cA[r] + i - skip[r]; }
      }
    }
  }
  // local -> global index of rank r (-1 for padding)
  std::vector<int> loc2glob(int r) const {
    std::vector<int> m(npad[r], -1);
    for (int x = 0; x < n; ++x) if (owner[x] == r) m[lidx[x]] = x;
    return m;
  }
};

struct PlanBase {
  int dtype = 1, device = 0, hconv = 0;
  int reduce_chunks = 1;    // pieces of the last pass of a sample-averaged product handed to the reduction hook as they finish
  GridInfo g;
  // slab decomposition (3-D only): axis 0 of the latent array, last axis (= x2 planes) of position space
  bool dist = false;
  int rank = 0, world = 1;
  AxisDist d0, d2;
  int rows0 = 0, planes2 = 0;        // local (padded) extents; = n0 / nl when not distributed
  int64_t scratch_elems = 0;         // complex elements each exchange / scratch buffer must hold
  virtual ~PlanBase() {}
};

template <class T> std::vector<cplx<T>> make_twiddles(int n) {
  std::vector<cplx<T>> tw(n);
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int j = 0; j < n; ++j) {
    // exact octant symmetry: evaluate sin/cos on [0, pi/4] only
    int jj = j % n; long double c, s;
    int oct = (int)((8LL * jj) / n);           // 0..7
    long double base;
    switch (oct) {
      case 0: base = two_pi * jj / n; c = cosl(base); s = sinl(base); break;
      case 1: base = two_pi * (n / 4.0L - jj) / n; c = sinl(base); s = cosl(base); break;
      case 2: base = two_pi * (jj - n / 4.0L) / n; c = -sinl(base); s = cosl(base); break;
      case 3: base = two_pi * (n / 2.0L - jj) / n; c = -cosl(base); s = sinl(base); break;
      case 4: base = two_pi * (jj - n / 2.0L) / n; c = -cosl(base); s = -sinl(base); break;
      case 5: base = two_pi * (3 * n / 4.0L - jj) / n; c = -sinl(base); s = -cosl(base); break;
      case 6: base = two_pi * (jj - 3 * n / 4.0L) / n; c = sinl(base); s = -cosl(base); break;
      default: base = two_pi * (n - jj) / n; c = cosl(base); s = -sinl(base); break;
    }
    tw[j] = cmake<T>((T)c, (T)(-s));   // exp(-2 pi i j / n)
  }
  return tw;
}

// L2 prefetch of later-phase / next-wave rows is OFF by default: measured on B200 it added 30-50 % DRAM
// read traffic (lines evicted before use) for no gain (profiles/r1_notes.md); NB200_PREFETCH=1 re-enables it.
inline bool p1m_enabled() { const char* e = std::getenv("NB200_P1M"); return !(e && e[0] == '0'); }   // developer knob: 0 = old P1 for every prologue
// staged passes (nb_passes2.cuh) for line lengths 2^NB_FAST_LGMIN .. 4096; NB200_FAST=0 forces the generic bodies (A/B knob)
#ifndef NB_FAST_LGMIN
#define NB_FAST_LGMIN 5
#endif
inline bool fast_enabled() { const char* e = std::getenv("NB200_FAST"); return !(e && e[0] == '0'); }
#define NB_LG_SWITCH(lg, CALL)                                                                     \
  switch (lg) {                                                                                    \
    case 5: CALL(5); break; case 6: CALL(6); break; case 7: CALL(7); break; case 8: CALL(8); break; \
    case 9: CALL(9); break; case 10: CALL(10); break; case 11: CALL(11); break; case 12: CALL(12); break; \
    default: throw Error{"nb200: internal: no staged pass for this line length"};                 \
  }
inline bool prefetch_disabled() { const char* e = std::getenv("NB200_PREFETCH"); return !(e && e[0] == '1'); }

template <class T> struct Plan : PlanBase {
  int lg0 = 0, lgm = 0, lgl = 0;
  T hsign = 1;
  DevBuf<cplx<T>> tw0, twm, twl, S0, S1;
  DevBuf<fft_slot_t> pos0, posm, posl, poslh;
  DevBuf<cplx<T>> ctw0, ctwm, ctwl, ctwlh;     // compact per-stage twiddle tables of the four line FFTs
  DevBuf<int> idxf, w_order, w_offs, plane_loc0, src_mul3, src_mul5;
  DevBuf<long> src_off3, src_off5;
  cplx<T>* xS0 = nullptr; cplx<T>* xS1 = nullptr; cplx<T>* xS2 = nullptr;   // host-provided buffers (distributed plans)
  // chunked pipeline of the distributed plans: chunk c of every rank's half-range piece
  int nchunks = 1;
  DevBuf<int> omapA, omapB;                 // outer indices (k2 / k0) of PCa / PCb chunk launches, concatenated
  std::vector<int> omapA_off, omapB_off;    // [nchunks+1]
  std::vector<int> p3_line0, p3_nlines, p5_line0, p5_nlines;   // [nchunks] line ranges of this rank
  FftDev f0, fm, fl, flh;   // line FFTs of length n0, nm, nl and nl/2
  DevBuf<T> W, p3part, p5part;
  PassCfg c1, cA, c3, cB, c5;
  int n3part = 0, n5part = 0;   // number of per-CTA partial sums the last P3 / P5 launch wrote
  bool fast = true;             // staged passes allowed (read once at plan creation)
  bool staged_ok = false;       // float64, single GPU: the staged bodies apply
  bool chain_ok = false;        // ... and every pass of the metric chain is covered (line lengths 32 .. 4096)
  bool use_chain = false;       // set by the operator that runs the whole chain (ChainScope), read by run_*
  int p5f_forced = -1;          // NB200_P5F=1 / 0: register-resident last pass forced on / off (default: chosen per shape)
  bool p5_sidx = true;          // NB200_P5_SIDX=0: the last pass reads the bin indices of its epilogue from global memory
  bool p5f_staged_epi = false;    // NB200_P5F_SE=0: the register-resident last pass loads its epilogue inputs per thread
  bool p3_stage_jl = true;      // NB200_P3_SJL=0: per-thread global loads of the Jacobian weights in the staged axis-0 pass
  bool l2_prefetch = false;     // bulk L2 prefetch of the next tile's epilogue rows: measured slower (+50 % DRAM reads); NB200_L2PF=1 enables
  TmaDesc d_pca, d_p3, d_pcb, d_p5;
  int sms = 148;
  int seg_lg_lpb = 2;   // lanes per mode bin in the segment sum (power of two near the mean bin population)
  int seg_grid() const { int bpb = 256 >> seg_lg_lpb; return (g.K + bpb - 1) / bpb; }

  static PassCfg choose(int lg_n_line /*log2 complex elems per line*/, int64_t group_lines, int64_t n_groups_outer,
                        bool sequential_lines, int64_t total_lines, const char* env_R = "NB200_LGR_NONE", int cap_shift = 0) {
    PassCfg c;
    const size_t line_bytes = (size_t(1) << lg_n_line) * sizeof(cplx<T>);
    const size_t maxs = max_smem_per_block() - 2048;
    if (line_bytes > maxs) throw Error{"nb200: grid extent too large for one shared-memory line"};
    const size_t budget = 96 * 1024;
    const int sms = sm_count();
    auto ctas = [&](int lgR) {
      int64_t R = int64_t(1) << lgR;
      return sequential_lines ? (total_lines + R - 1) / R : n_groups_outer * ((group_lines + R - 1) / R);
    };
    // complex elements per CTA: 4096 for long lines, 1024 for short ones (measured optimum on B200:
    // 256^3 runs 9 % faster with R = 4 than with R = 16, 2048^2 / 4096^2 are best at 4096 elements)
    // P1 (cap_shift = 1) counts its REAL line length: 8192 reals for long lines, 1024 for short ones
    const int64_t elems = (lg_n_line + cap_shift >= 10) ? (int64_t(4096) << cap_shift) : 1024;
    int lgR = 0;
    while (lgR < 4 && (int64_t(2) << lgR) <= (sequential_lines ? total_lines : group_lines) &&
           (size_t(2) << lgR) * line_bytes <= budget && ((int64_t(2) << lgR) << (lg_n_line + cap_shift)) <= elems)
      ++lgR;
    while (lgR > 0 && ctas(lgR) < 2 * sms) --lgR;
    if (const char* e = std::getenv(env_R)) {   // tuning override (developer knob)
      int v = std::atoi(e);
      if (v >= 0 && v <= 4 && (size_t(1) << v) * line_bytes <= maxs && (int64_t(1) << v) <= (sequential_lines ? total_lines : group_lines)) lgR = v;
    }
    c.lg_R = lgR;
    c.pitch = 1 << lg_n_line;
    c.smem = (size_t(1) << lgR) * c.pitch * sizeof(cplx<T>);
    c.grid = (int)ctas(lgR);
    int64_t bf = (int64_t(1) << (lgR + lg_n_line)) / 8;
    int blk = 64;
    while (blk < 256 && blk < bf) blk *= 2;
    if (const char* e = std::getenv("NB200_BLOCK")) { int v = std::atoi(e); if (v == 64 || v == 128 || v == 256) blk = v; }
    c.block = blk;
    {   // CTAs resident on the whole device (shared-memory limited): distance of the input prefetch
      size_t per = c.smem + 2048;
      int per_sm = (int)std::min<size_t>(8, (228 * 1024) / per);
      c.ahead = std::max(1, per_sm) * sms;
      if (prefetch_disabled()) c.ahead = 0;
    }
    return c;
  }

  void init(int device_, int ndim, const int64_t* shp, const double* dst, int hconv_, int rank_ = 0, int world_ = 1) {
    device = device_; hconv = hconv_;
    fast = fast_enabled();
    dtype = sizeof(T) == 8 ? 1 : 0;
    hsign = hconv ? T(-1) : T(1);
    rank = rank_; world = world_; dist = world_ > 1;
    dev_set(device);
    if (g.build(ndim, shp, dst, !dist)) throw Error{last_error_ref()};
    lg0 = ilog2(g.n0); lgm = ilog2(g.nm); lgl = ilog2(g.nl);
    rows0 = g.n0; planes2 = g.nl;
    int cA0 = g.h0 + 1, cA2 = g.hl + 1;      // planes of the half ranges handled by P5 / P3
    std::vector<int> idxf_loc = g.idxf, w_order_loc = g.w_order, w_offs_loc = g.w_offs;
    int64_t nW_loc = g.nW;
    if (dist) {
      if (!g.three) throw Error{"nb200: slab decomposition needs a 3-D grid"};
      if (!is_pow2(world) || g.n0 < 2 * world || g.nl < 2 * world) throw Error{"nb200: world size must be a power of two <= n0/2 and <= n2/2"};
      d0.build(g.n0, world); d2.build(g.nl, world);
      rows0 = d0.npad[rank]; planes2 = d2.npad[rank];
      cA0 = d0.cA[rank]; cA2 = d2.cA[rank];
      const size_t plane = (size_t)(g.hm + 1) * (g.hl + 1);
      idxf_loc.assign(g.idxf.begin() + (size_t)d0.a0[rank] * plane, g.idxf.begin() + (size_t)(d0.a0[rank] + cA0) * plane);
      // CSR of the LOCAL partial sums W[(a_loc*nm + km)][x]
      nW_loc = (int64_t)cA0 * g.nm * (g.hl + 1);
      w_offs_loc.assign(g.K + 1, 0);
      auto wbin = [&](int64_t p) {
        int x = (int)(p % (g.hl + 1)); int64_t l = p / (g.hl + 1);
        int km = (int)(l % g.nm), a = (int)(l / g.nm);
        int fm = km <= g.nm - km ? km : g.nm - km;
        return idxf_loc[((size_t)a * (g.hm + 1) + fm) * (g.hl + 1) + x];
      };
      for (int64_t p = 0; p < nW_loc; ++p) w_offs_loc[wbin(p) + 1]++;
      for (int b = 0; b < g.K; ++b) w_offs_loc[b + 1] += w_offs_loc[b];
      w_order_loc.resize(nW_loc);
      std::vector<int> cur(w_offs_loc.begin(), w_offs_loc.end() - 1);
      for (int64_t p = 0; p < nW_loc; ++p) w_order_loc[cur[wbin(p)]++] = (int)p;
      // folded plane of the local bin table for every local latent row
      std::vector<int> pl(rows0, 0);
      for (int i = 0; i < d0.cA[rank]; ++i) pl[i] = i;
      for (int i = 0; i < d0.cB[rank]; ++i) pl[d0.cA[rank] + i] = i + d0.skip[rank];
      plane_loc0.upload(pl);
      // chunk tables of the two exchanges (element x of line l: recv[src_off[x] + l * src_mul[x]])
      auto tables = [&](const AxisDist& ax, int my_lines_planes, DevBuf<long>& off, DevBuf<int>& mul) {
        std::vector<long> roff(world, 0);
        for (int p = 1; p < world; ++p) roff[p] = roff[p - 1] + (long)my_lines_planes * g.nm * ax.npad[p - 1];
        std::vector<long> o(ax.n); std::vector<int> m(ax.n);
        for (int x = 0; x < ax.n; ++x) { o[x] = roff[ax.owner[x]] + ax.lidx[x]; m[x] = ax.npad[ax.owner[x]]; }
        off.upload(o); mul.upload(m);
      };
      tables(d0, cA2, src_off3, src_mul3);   // X1: lines (k2 in my half-range piece, k1) along global j0
      tables(d2, cA0, src_off5, src_mul5);   // X2: lines (k0 in my piece, k1) along global x2
    }
    tw0.upload(make_twiddles<T>(g.n0));
    twm.upload(make_twiddles<T>(g.nm));
    twl.upload(make_twiddles<T>(g.nl));
    idxf.upload(idxf_loc); w_order.upload(w_order_loc); w_offs.upload(w_offs_loc);
    auto mk = [&](int lg, DevBuf<fft_slot_t>& buf, DevBuf<cplx<T>>& ctw, int order) {
      if (lg > 14) throw Error{"nb200: line length above 2^14 is not supported"};
      std::vector<fft_slot_t> t(size_t(1) << lg);
      fill_pos_table(lg, t.data(), order);
      buf.upload(t);
      std::vector<cplx<T>> full = make_twiddles<T>(1 << lg), compact;
      fill_compact_twiddles(lg, full.data(), compact, order);
      ctw.upload(compact);
      return make_fft_dev(lg, buf.p, ctw.p, order);
    };
    int order1 = 0;
    if (const char* e = std::getenv("NB200_P1_R4")) order1 = (e[0] == '1');   // developer knob: radix-4 stage first in P1
    f0 = mk(lg0, pos0, ctw0, 0); fm = mk(lgm, posm, ctwm, 0); fl = mk(lgl, posl, ctwl, 0); flh = mk(lgl - 1, poslh, ctwlh, order1);
    W.alloc((size_t)nW_loc);
    const int64_t n0 = g.n0, nm = g.nm, nl = g.nl, h0 = g.h0, hl = g.hl;
    size_t sc = 0;
    if (dist) {
      int64_t sum0 = 0, sum2 = 0;
      for (int p = 0; p < world; ++p) { sum0 += d0.npad[p]; sum2 += d2.npad[p]; }
      sc = (size_t)std::max(std::max((int64_t)rows0 * (hl + 1) * nm, (int64_t)cA2 * nm * sum0),
                            std::max((h0 + 1) * (int64_t)planes2 * nm, (int64_t)cA0 * nm * sum2));
      scratch_elems = (int64_t)sc;           // buffers come from the host (torch tensors, for the all-to-all)
    } else {
      if (g.three) sc = (size_t)std::max(n0 * (hl + 1) * nm, (h0 + 1) * nl * nm);
      else sc = (size_t)std::max((hl + 1) * n0, (h0 + 1) * nl);
      S0.alloc(sc); S1.alloc(sc);
      scratch_elems = (int64_t)sc;
    }
    // P1: real lines of length nl -> complex FFT of nl/2
    if (g.three) c1 = choose(lgl - 1, nm, rows0, false, 0, "NB200_LGR1", 1);
    else c1 = choose(lgl - 1, n0, 1, false, 0, "NB200_LGR1", 1);
    c1.lg_n = lgl;
    if (g.three) {
      cA = choose(lgm, rows0, hl + 1, false, 0, "NB200_LGRC"); cA.lg_n = lgm;
      cB = choose(lgm, planes2, h0 + 1, false, 0, "NB200_LGRC"); cB.lg_n = lgm;
    }
    c3 = choose(lg0, 0, 0, true, (int64_t)cA2 * nm, "NB200_LGR3"); c3.lg_n = lg0;
    c5 = choose(lgl, 0, 0, true, (int64_t)cA0 * nm, "NB200_LGR5"); c5.lg_n = lgl;
    {
      double avg = (double)nW_loc / (double)g.K;
      seg_lg_lpb = 0;
      while (seg_lg_lpb < 6 && (1 << (seg_lg_lpb + 1)) <= avg / 3.0) ++seg_lg_lpb;   // ~3-6 entries per lane
    }
    sms = sm_count();
    p3part.alloc((size_t)2 * std::max(c3.grid, 2 * sms) + 64);
    p5part.alloc((size_t)std::max(c5.grid, 2 * sms) + 32);
    n3part = c3.grid; n5part = c5.grid;
    staged_ok = fast && sizeof(T) == 8 && !dist;
    {
      const int lgc = lgl - 1, n_r1 = g.three ? g.nm : g.n0;
      chain_ok = staged_ok && lgc >= NB_FAST_LGMIN && lgc <= 11 && n_r1 >= (F16_TILE >> lgc) && lg0 >= NB_FAST_LGMIN && lg0 <= 12 &&
                 lgl >= 6 && lgl <= 12 && (!g.three || (lgm >= NB_FAST_LGMIN && lgm <= 12));
      const bool chain_ok_shape = chain_ok;
      // measured (profiles/r3_notes.md): the chain wins where the lines are long (4096^2: 0.623 vs 0.700 ms, 2048^2: 0.216
      // vs 0.223); on short 3-D lines (256^3) the generic first / last passes are still faster (more resident warps for
      // their streaming phases: 0.632 vs 0.592 ms), so those shapes keep the generic bodies
      chain_ok = chain_ok && !g.three && lg0 >= 11 && lgl >= 11;
      if (const char* e = std::getenv("NB200_CHAIN")) { if (e[0] == '0') chain_ok = false; else if (e[0] == '1') chain_ok = chain_ok_shape; }
      if (const char* e = std::getenv("NB200_L2PF")) l2_prefetch = (e[0] == '1');
      if (const char* e = std::getenv("NB200_P5F")) p5f_forced = (e[0] == '1') ? 1 : 0;
      if (const char* e = std::getenv("NB200_P3_SJL")) p3_stage_jl = (e[0] == '1');
      if (const char* e = std::getenv("NB200_P5F_SE")) p5f_staged_epi = (e[0] == '1');
      if (const char* e = std::getenv("NB200_P5_SIDX")) p5_sidx = (e[0] == '1');
    }
    if (staged_ok) {
      // gather descriptors: P3 / P5 read column l of [n_line][(h+1) n_mid]; PCa / PCb read column k of the rows of one plane
      make_tma_desc(d_p3, g.three ? (const void*)s1() : (const void*)s0(), (uint64_t)(g.hl + 1) * g.nm, (uint64_t)g.n0, std::min(256, g.n0));
      make_tma_desc(d_p5, s1(), (uint64_t)(g.h0 + 1) * g.nm, (uint64_t)g.nl, std::min(256, g.nl));
      if (g.three) {
        make_tma_desc(d_pca, s0(), (uint64_t)(g.hl + 1), (uint64_t)g.n0 * g.nm, std::min(256, g.nm));
        make_tma_desc(d_pcb, s0(), (uint64_t)(g.h0 + 1), (uint64_t)g.nl * g.nm, std::min(256, g.nm));
      }
    }
    if (dist) set_chunks(1);
  }
  cplx<T>* s0() { return dist ? xS0 : S0.p; }
  cplx<T>* s1() { return dist ? xS1 : S1.p; }
  int64_t local_latent_grid() const { return (int64_t)rows0 * g.nm * g.nl; }     // xi entries held by this rank
  int64_t local_position_grid() const { return (int64_t)planes2 * g.nm * g.n0; }

  FoldGeom fold_geom() const {
    FoldGeom f;
    if (g.three) { f.n_o = g.n0; f.n_r = g.nm; } else { f.n_o = 1; f.n_r = g.n0; }
    f.n = g.nl; f.hr1 = f.n_r / 2 + 1; f.h1 = g.hl + 1;
    f.plane_loc = dist ? plane_loc0.p : nullptr;
    return f;
  }
  MirrorGeom mgeom(int n_a, int h_a, const AxisDist& ax) const {
    MirrorGeom m; m.n_a = n_a; m.h_a = h_a; m.lg_mid = lgm; m.dist = dist ? 1 : 0; m.a0 = 0; m.cA = 0; m.skip = 0;
    if (dist) { m.a0 = ax.a0[rank]; m.cA = ax.cA[rank]; m.skip = ax.skip[rank]; }
    return m;
  }
  MirrorGeom mg3() const { return mgeom(g.nl, g.hl, d2); }
  MirrorGeom mg5() const { return mgeom(g.n0, g.h0, d0); }
  // buffer roles.  single GPU: P1->S0, PCa S0->S1, P3 S1->S0 (2-D: S0->S1), PCb S0->S1, P5 reads S1.
  // distributed (three buffers so that an exchange can overlap the passes around it):
  //   P1->S2, PCa S2->S1, [all-to-all S1->S0], P3 S0->S2, PCb S2->S0, [all-to-all S0->S1], P5 reads S1.
  cplx<T>* p3_in() { return dist ? s0() : (g.three ? s1() : s0()); }
  cplx<T>* p3_out() { return dist ? xS2 : (g.three ? s0() : s1()); }
  // split [start, start+count) into nchunks contiguous pieces; piece c = [split(c), split(c+1))
  static int split_at(int start, int count, int c, int nch) { return start + (int)((int64_t)count * c / nch); }
  void set_chunks(int nch) {
    if (!dist) throw Error{"nb200_plan_set_chunks: not a slab-decomposed plan"};
    if (nch < 1 || nch > 16) throw Error{"nb200_plan_set_chunks: 1 <= nchunks <= 16"};
    nchunks = nch;
    auto build = [&](const AxisDist& ax, DevBuf<int>& dev, std::vector<int>& off, std::vector<int>& l0, std::vector<int>& nl) {
      std::vector<int> all; off.assign(nch + 1, 0); l0.assign(nch, 0); nl.assign(nch, 0);
      for (int c = 0; c < nch; ++c) {
        for (int q = 0; q < world; ++q)
          for (int a = split_at(ax.a0[q], ax.cA[q], c, nch); a < split_at(ax.a0[q], ax.cA[q], c + 1, nch); ++a) all.push_back(a);
        off[c + 1] = (int)all.size();
        int s = split_at(0, ax.cA[rank], c, nch), e = split_at(0, ax.cA[rank], c + 1, nch);
        l0[c] = s * g.nm; nl[c] = (e - s) * g.nm;
      }
      dev.upload(all);
    };
    build(d2, omapA, omapA_off, p3_line0, p3_nlines);   // exchange 1: k2 planes, consumed by P3
    build(d0, omapB, omapB_off, p5_line0, p5_nlines);   // exchange 2: k0 planes, consumed by P5
  }

  PointOp<T> make_op(int mode) const {
    PointOp<T> op;
    std::memset(&op, 0, sizeof(op));
    op.mode = mode; op.invV = T(1); op.sc = T(1); op.nl_exp = 1; op.nl_s = nullptr; op.nl_ds = nullptr;
    op.n_a_full = g.nl; op.n_mid = g.nm; op.n = g.n0;
    return op;
  }

  template <class Pro> void run_p1(stream_t st, const Pro& pro) {
    P1Params<T, Pro> p;
    p.lg_n = lgl; p.lg_R = c1.lg_R; p.pitch = c1.pitch; p.tw = twl.p; p.lg_tw = lgl; p.fft = flh; p.out = dist ? xS2 : s0(); p.ahead = c1.ahead; p.pro = pro;
    if (g.three) {
      p.n_o = rows0; p.n_r = g.nm; p.in_ostride = (long)g.nm * g.nl; p.in_rstride = g.nl;
      p.out_ostride = (long)(g.hl + 1) * g.nm; p.out_kstride = g.nm;
    } else {
      p.n_o = 1; p.n_r = g.n0; p.in_ostride = 0; p.in_rstride = g.nl; p.out_ostride = 0; p.out_kstride = g.n0;
    }
    // staged chain (nb_passes2.cuh): mirror pairs of rows, register-resident half-length transform, row-major output
    if constexpr (Pro::kStaged) {
      if (use_chain) {
        const int lgc = lgl - 1;
        P1FParams<T, Pro> q;
        q.lg_n = lgl; q.n_r = p.n_r; q.n_o = p.n_o; q.in_ostride = p.in_ostride; q.in_rstride = p.in_rstride;
        q.tw = twl.p; q.lg_tw = lgl; q.out = s0(); q.pro = pro;
        const int lpc = F16_TILE >> lgc;
        q.ntiles = (g.three ? rows0 : 1) * (p.n_r / lpc);
        const int grid = std::min(q.ntiles, 2 * sms);
#define NB_CALL(LG) launch<P1FBody<T, Pro, LG>>(grid, F16_NT, StageLayout<T, LG>::BYTES, st, q)
        switch (lgc) {
          case 5: NB_CALL(5); break; case 6: NB_CALL(6); break; case 7: NB_CALL(7); break; case 8: NB_CALL(8); break;
          case 9: NB_CALL(9); break; case 10: NB_CALL(10); break; default: NB_CALL(11); break;
        }
#undef NB_CALL
        return;
      }
    }
    // prologues that read the mode-bin table share one lookup per mirror quad (P1MBody); needs >= 1 pair of rows per CTA
    if constexpr (Pro::kUsesBins) {
      if (c1.lg_R >= 1 && p.n_r >= 2 && p1m_enabled()) {
        bool three = 3 * (c1.smem + 2048) <= size_t(100) * 1024;
        if (const char* e = std::getenv("NB200_P1M_MINB")) three = (e[0] == '3');     // developer knob
        if (three) launch<P1MBody<T, Pro, 3>>(c1.grid, c1.block, c1.smem, st, p);
        else launch<P1MBody<T, Pro, 2>>(c1.grid, c1.block, c1.smem, st, p);
        return;
      }
    }
    if (pro.aligned()) launch<P1Body<T, Pro, true>>(c1.grid, c1.block, c1.smem, st, p);
    else launch<P1Body<T, Pro, false>>(c1.grid, c1.block, c1.smem, st, p);
  }
  void run_pc(stream_t st, bool second, const int* omap = nullptr, int n_o_chunk = -1) {
    if (!g.three) return;
    PCParams<T> p;
    const PassCfg& c = second ? cB : cA;
    p.omap = omap;
    p.lg_n = lgm; p.lg_R = c.lg_R; p.pitch = c.pitch; p.tw = twm.p; p.lg_tw = lgm; p.fft = fm; p.in = dist ? xS2 : s0(); p.out = s1();
    if (!second) {   // [j0][k2][j1] -> [k2][k1][j0]   (j0: local rows when distributed)
      p.n_o = g.hl + 1; p.n_r = rows0; p.in_ostride = g.nm; p.in_rstride = (long)(g.hl + 1) * g.nm;
      p.out_ostride = (long)g.nm * rows0; p.out_kstride = rows0;
    } else {         // [k0][x2][x1] -> [k0][k1][x2]   (x2: local planes when distributed)
      if (dist) { p.in = xS2; p.out = s0(); }
      p.n_o = g.h0 + 1; p.n_r = planes2; p.in_ostride = (long)planes2 * g.nm; p.in_rstride = g.nm;
      p.out_ostride = (long)g.nm * planes2; p.out_kstride = planes2;
    }
    if (omap && n_o_chunk <= 0) return;
    if (use_chain) {
      PCFParams<T> q;
      q.desc = second ? d_pcb : d_pca;
      q.ncols = second ? g.h0 + 1 : g.hl + 1;
      q.nlines = (long)(second ? g.nl : g.n0) * q.ncols;
      q.omap = nullptr; q.tw = twm.p; q.out = s1();
      const int lpc = F16_TILE >> lgm;
      q.ntiles = (int)((q.nlines + lpc - 1) / lpc);
      const int gridf = std::min(q.ntiles, 2 * sms);
#define NB_CALL(LG) launch<PCFBody<T, LG>>(gridf, F16_NT, StageLayout<T, LG>::BYTES, st, q)
      NB_LG_SWITCH(lgm, NB_CALL)
#undef NB_CALL
      return;
    }
    int grid = c.grid;
    if (omap) grid = n_o_chunk * (p.n_r >> p.lg_R);
    if (3 * (c.smem + 2048) <= size_t(100) * 1024) launch<PCBody<T, 3>>(grid, c.block, c.smem, st, p);
    else launch<PCBody<T, 2>>(grid, c.block, c.smem, st, p);
  }
  // (the staged launches live in non-template members: nvcc turned the same switch inside the run_p3 / run_p5 member
  // templates into a host-side exit(1))
  void launch_p3f(stream_t st, const P3FParams<T>& q, int gridf) {
#define NB_CALL(LG) launch<P3FBody<T, LG>>(gridf, F16_NT, StageLayout<T, LG>::BYTES, st, q)
    NB_LG_SWITCH(lg0, NB_CALL)
#undef NB_CALL
  }
  void launch_p5f(stream_t st, const P5FParams<T>& q, int gridf) {
    if (q.staged_epi) {
#define NB_CALL(LG) launch<P5FBody<T, LG, true>>(gridf, F16_NT, P5FBody<T, LG, true>::BYTES_STAGED, st, q)
      NB_LG_SWITCH(lgl, NB_CALL)
#undef NB_CALL
      return;
    }
#define NB_CALL(LG) launch<P5FBody<T, LG, false>>(gridf, F16_NT, StageLayout<T, LG>::BYTES, st, q)
    NB_LG_SWITCH(lgl, NB_CALL)
#undef NB_CALL
  }
  template <bool FWD, bool ADJ> void run_p3(stream_t st, const PointOp<T>& op, int line0 = 0, int nlines = -1) {
    P3Params<T> p;
    p.line0 = line0;
    int grid3 = c3.grid;
    if (nlines >= 0) { if (nlines == 0) return; grid3 = (nlines + (1 << c3.lg_R) - 1) >> c3.lg_R; }
    p.lg_n = lg0; p.lg_R = c3.lg_R; p.mg = mg3(); p.pitch = c3.pitch; p.tw = tw0.p; p.lg_tw = lg0; p.fft = f0; p.hsign = hsign; p.ahead = c3.ahead;
    p.in = p3_in(); p.out = p3_out(); p.out_kstride = (long)planes2 * g.nm; p.op = op;
    p.src_off = dist ? src_off3.p : nullptr; p.src_mul = dist ? src_mul3.p : nullptr;
    const size_t sm = c3.smem + LINEINFO_BYTES;
    n3part = c3.grid;
    if constexpr (FWD && ADJ) {
      if (op.mode == PM_METRIC && use_chain) {
        P3FParams<T> q;
        q.desc = d_p3; q.mg = p.mg; q.tw = tw0.p; q.hsign = hsign; q.out = p.out;
        q.line0 = line0; q.nlines = nlines >= 0 ? nlines : p.mg.nlines(); q.op = op; q.prefetch = l2_prefetch ? 1 : 0;
        q.stage_jl = p3_stage_jl ? 1 : 0;
        const int lpc = F16_TILE >> lg0;
        q.ntiles = (q.nlines + lpc - 1) / lpc;
        const int gridf = std::min(q.ntiles, 2 * sms);
        n3part = gridf;
        launch_p3f(st, q, gridf);
        return;
      }
      if (op.mode == PM_METRIC) {
        // measured: 3 CTAs/SM win for short lines (256^3: 145 vs 166 us) and for 4096-point lines (204 vs 213 us),
        // 2 CTAs/SM for two 2048-point lines per CTA (56 vs 60 us)
        bool three = 3 * (sm + 2048) <= size_t(100) * 1024 || lg0 >= 12;
        if (const char* e = std::getenv("NB200_P3_MINB")) three = (e[0] == '3');     // developer knob
        if (three) launch<P3Body<T, true, true, PM_METRIC, 3>>(grid3, c3.block, sm, st, p);
        else launch<P3Body<T, true, true, PM_METRIC, 2>>(grid3, c3.block, sm, st, p);
      }
      else if (op.mode == PM_LINEARIZE) launch<P3Body<T, true, true, PM_LINEARIZE>>(grid3, c3.block, sm, st, p);
      else throw Error{"nb200: invalid pointwise mode for the fused pass"};
    } else if constexpr (FWD) {
      if (op.mode == PM_LINEARIZE) launch<P3Body<T, true, false, PM_LINEARIZE>>(grid3, c3.block, sm, st, p);
      else if (op.mode == PM_JVP_OUT) launch<P3Body<T, true, false, PM_JVP_OUT>>(grid3, c3.block, sm, st, p);
      else if (op.mode == PM_FIELD_OUT) launch<P3Body<T, true, false, PM_FIELD_OUT>>(grid3, c3.block, sm, st, p);
      else throw Error{"nb200: invalid pointwise mode for the forward pass"};
    } else {
      if (op.mode != PM_LOAD) throw Error{"nb200: invalid pointwise mode for the adjoint pass"};
      launch<P3Body<T, false, true, PM_LOAD>>(grid3, c3.block, sm, st, p);
    }
  }
  template <class Epi> void run_p5(stream_t st, const Epi& epi, int line0 = 0, int nlines = -1) {
    P5Params<T, Epi> p;
    p.line0 = line0;
    int grid5 = c5.grid;
    if (nlines >= 0) { if (nlines == 0) return; grid5 = (nlines + (1 << c5.lg_R) - 1) >> c5.lg_R; }
    p.lg_n = lgl; p.lg_R = c5.lg_R; p.mg = mg5(); p.hmid1 = g.hm + 1; p.pitch = c5.pitch; p.tw = twl.p; p.lg_tw = lgl; p.fft = fl; p.ahead = c5.ahead;
    p.hsign = hsign; p.in = s1(); p.epi = epi;
    p.src_off = dist ? src_off5.p : nullptr; p.src_mul = dist ? src_mul5.p : nullptr;
    size_t sm5 = c5.smem + LINEINFO_BYTES;
    n5part = c5.grid;
    p.gather = 0;
    p.sidx_off = 0;
    // measured (profiles/r3_notes.md): 180.4 -> 177.8 us on 4096-point lines, nothing on 2048, a loss on short 3-D lines
    if (Epi::BATCHED && p5_sidx && lgl >= 12) {       // bin-index rows behind the line buffers and the reduction scratch
      p.sidx_off = (int)(sm5 + 512);
      sm5 += (((size_t)1 << c5.lg_R) * ((1 << (lgl - 1)) + 1) * sizeof(int) + 15) & ~size_t(15);
    }
    if constexpr (Epi::BATCHED) {
      // register-resident P5 (P5FBody): its epilogue is latency bound at 16 warps per SM -- measured against the generic body
      // (which gathers its lines through the tensor map in a staged chain): 173 vs 178 us on 4096-point lines, 59 vs 58 us
      // on 2048, 232 vs 125 us at 256^3 -> chosen for 4096-point lines only (NB200_P5F=1 / 0 forces it on / off); its
      // variant with the epilogue rows staged in shared memory (NB200_P5F_SE=1) measured 210 us
      const bool p5f_here = p5f_forced >= 0 ? p5f_forced == 1 : (use_chain && lgl >= 12);
      if (p5f_here && staged_ok && lgl >= NB_FAST_LGMIN && lgl <= 12 && (p5f_staged_epi || epi.add != epi.out)) {    // (run(): read-only loads must not alias the output)
        P5FParams<T> q;
        q.desc = d_p5; q.contig = use_chain ? nullptr : p.in;
        q.mg = p.mg; q.hmid1 = p.hmid1; q.tw = twl.p; q.hsign = hsign; q.line0 = line0;
        q.nlines = nlines >= 0 ? nlines : p.mg.nlines(); q.epi = epi; q.prefetch = l2_prefetch ? 1 : 0; // (bulk copies of the epilogue rows need 16-byte aligned rows)
        q.staged_epi = (p5f_staged_epi && ((uintptr_t)epi.add % 16) == 0 && ((uintptr_t)epi.xi % 16) == 0) ? 1 : 0;
        if (!q.staged_epi && epi.add == epi.out) goto generic_p5;
        const int lpc = F16_TILE >> lgl;
        q.ntiles = (q.nlines + lpc - 1) / lpc;
        const int gridf = std::min(q.ntiles, 2 * sms);
        n5part = gridf;
        launch_p5f(st, q, gridf);
        return;
      }
    }
  generic_p5:
    if (use_chain) { p.gather = 1; p.desc = d_p5; }
    if (3 * (sm5 + 2048) <= size_t(100) * 1024) launch<P5Body<T, Epi, 3, false>>(grid5, c5.block, sm5, st, p);
    else launch<P5Body<T, Epi, 2, true>>(grid5, c5.block, sm5, st, p);
  }
  // natural (d0,d1,d2) -> reversed axes; for the T-layout <-> natural conversions
  void run_rev(stream_t st, const T* in, T* out, bool to_T) {
    RevParams<T> p; p.in = in; p.out = out;
    int a0 = g.n0, a1 = g.nm, a2 = g.nl;
    if (to_T) { p.d0 = a0; p.d1 = a1; p.d2 = a2; } else { p.d0 = a2; p.d1 = a1; p.d2 = a0; }
    int grid = p.d1 * ((p.d0 + 31) / 32) * ((p.d2 + 31) / 32);
    launch<RevBody<T>>(grid, 256, 32 * 33 * sizeof(T), st, p);
  }

  // ---- raw operators ----
  void hartley(stream_t st, const T* in, T* out) {
    ProPlain<T> pro; pro.x = in;
    run_p1(st, pro); run_pc(st, false);
    PointOp<T> op = make_op(PM_FIELD_OUT); op.natural = 1; op.pos_out = out;
    run_p3<true, false>(st, op);
  }
  void cf_apply(stream_t st, const T* amp, const T* xi, T offset, T* out) {
    ProAmp<T> pro; pro.xi = xi; pro.idxf = idxf.p; pro.amp = amp; pro.fg = fold_geom();
    run_p1(st, pro); run_pc(st, false);
    PointOp<T> op = make_op(PM_FIELD_OUT); op.natural = 1; op.pos_out = out; op.invV = T(1.0 / g.V); op.offset = offset;
    run_p3<true, false>(st, op);
  }
};

// marks an operator sequence whose passes all run staged (they share the row-major intermediate layouts)
template <class T> struct ChainScope {
  Plan<T>& P; bool prev;
  ChainScope(Plan<T>& p, bool on) : P(p), prev(p.use_chain) { P.use_chain = on; }
  ~ChainScope() { P.use_chain = prev; }
};

}  // namespace nb
