// nifty_b200 -- model (amplitude priors + likelihood) and linearisation objects; operator launch
// sequences for linearise / value+gradient / metric / sqrt-metric applications.
#pragma once
#include <functional>
#include "nb_plan.cuh"
#include "nb_chain.cuh"
#include "../../include/nifty_b200.h"

namespace nb {

// small helper kernels -------------------------------------------------------------------------
template <class T> struct ReduceColsParams { const T* partials; int n, ncol; T* out0; T* out1; };
template <class T> struct ReduceColsBody {
  typedef ReduceColsParams<T> Params;
  static NB_HD void run(Ctx& ctx, const Params& p, void* smem) {
    T a0 = 0, a1 = 0;
    NB_FOR(ctx, i, p.n) { a0 += p.partials[p.ncol * i]; if (p.ncol > 1) a1 += p.partials[p.ncol * i + 1]; }
    a0 = ctx.block_sum(a0, smem);
    a1 = ctx.block_sum(a1, smem);
    if (ctx.tid == 0) { if (p.out0) *p.out0 = a0; if (p.out1) *p.out1 = a1; }
  }
};

enum PosMapMode { PMAP_TRAFO = 0, PMAP_NRES = 1, PMAP_COPY = 2 };
template <class T> struct PosMapParams {
  int mode, lh_kind; long n; const T* s; const T* data; T w_scalar; const T* w_arr; T* out;
};
template <class T> struct PosMapBody {
  typedef PosMapParams<T> Params;
  static NB_HD void run(Ctx& ctx, const Params& p, void*) {
    for (long i = (long)ctx.bid * ctx.nthr + ctx.tid; i < p.n; i += (long)ctx.nblk * ctx.nthr) {
      T s = p.s[i], r;
      if (p.mode == PMAP_COPY) r = s;
      else if (p.lh_kind == LH_GAUSS) {
        T sw = nb_sqrt(p.w_arr ? p.w_arr[i] : p.w_scalar);
        r = (p.mode == PMAP_TRAFO) ? sw * s : sw * ((p.data ? p.data[i] : T(0)) - s);
      } else {
        r = (p.mode == PMAP_TRAFO) ? T(2) * nb_sqrt(s) : ((p.data ? p.data[i] : T(0)) - s) / nb_sqrt(s);
      }
      p.out[i] = r;
    }
  }
};

struct ModelBase { PlanBase* plan = nullptr; virtual ~ModelBase() {} };
struct LinBase { ModelBase* model = nullptr; virtual ~LinBase() {} };

template <class T> struct Lin;

template <class T> struct Model : ModelBase {
  Plan<T>* P = nullptr;
  AmpModel<T> am;
  T offset_mean = 0;
  int lh_kind = LH_GAUSS, nl_exp = 1;
  bool have_lh = false;
  T w_scalar = 1;
  DevBuf<T> ell, multT, dt, modes, data, w_arr;
  bool has_w_arr = false;
  // chain workspaces
  DevBuf<Aff<T>> agg, preaff;
  DevBuf<T> gbuf, partials, tmp_pos;
  DevBuf<cplx<T>> ad;
  DevBuf<unsigned> counters;
  int nchunksK = 1, nchunksJ = 1;
  Lin<T>* scratch_lin = nullptr;

  void init(Plan<T>* plan_, const nb200_model_desc& d);
  ~Model();

  int scan_e = 8;      // elements per thread of the scan kernels (scan_pick_e)
  size_t scan_smem() const { return scan_smem_bytes<T>(scan_e); }

  // fused chains (nb_chain.cuh): one cooperative launch per chain; launch geometry fixed at model creation
  struct ChainCfg {
    bool ok = false;
    int grid_t = 0; long per_t = 0; size_t smem_t = 0;                       // tangent
    int grid_c = 0; long per_c = 0; size_t smem_c = 0;                       // cotangent
    int dbg_twice = 0;
    int nph_t = 2, nph_c = 3;                                                // phases to run (developer timing aid)
  } chain;
  DevBuf<Aff<T>> cagg, crel;
  DevBuf<int> seg_b0, seg_lg;       // segment sum: first bin and log2(lanes per bin) of every CTA
  DevBuf<long long> cdbg;
  void init_chain();
};

template <class T> struct Lin : LinBase {
  Model<T>* M = nullptr;
  DevBuf<T> pos, amp, wS, Pb, s, jl, scal, ellv_buf, cv_buf;
  DevBuf<T> nl_s, nl_ds;       // tabulated pointwise map of the field at the position of the NEXT update (nb200_lin_set_pointwise)
  bool nl_ready = false;
  const T* ellv = nullptr; const T* cv = nullptr;   // effective du/dslope and du/dcutoff tables (Matern: per linearisation)
  bool valid = false;

  void init(Model<T>* m) {
    M = m; model = m;
    const GridInfo& g = m->P->g;
    pos.alloc((size_t)m->am.L); amp.alloc(g.K); wS.alloc(g.K); Pb.alloc(g.K);
    s.alloc((size_t)m->P->local_position_grid()); jl.alloc((size_t)m->P->local_position_grid()); scal.alloc(SC_COUNT);
    if (m->am.matern) { ellv_buf.alloc(g.K); cv_buf.alloc(g.K); ellv = ellv_buf.p; cv = cv_buf.p; }
    else { ellv = m->am.ell; cv = nullptr; }
  }

  // forward amplitude chain at this->pos
  void amp_forward(stream_t st) {
    Model<T>& m = *M; const int K = m.am.K;
    FwdElem<T> el; el.m = m.am; el.pos = pos.p;
    const int nch = m.am.has_dev ? m.nchunksK : 1;   // without deviations every element is the identity
    if (m.am.has_dev && m.nchunksK > 1) {
      ScanAggParams<T, FwdElem<T>> pa; pa.n = K; pa.elem = el; pa.agg = m.agg.p; pa.pre = m.preaff.p; pa.counter = m.counters.p + 5;
      if (m.scan_e == 8) launch<ScanAggBody<T, FwdElem<T>, 8>>(m.nchunksK, SCAN_NT, m.scan_smem(), st, pa); else launch<ScanAggBody<T, FwdElem<T>, 4>>(m.nchunksK, SCAN_NT, m.scan_smem(), st, pa);
    }
    FwdOut<T> fo; fo.m = m.am; fo.pos = pos.p; fo.P = Pb.p; fo.partials = m.partials.p;
    fo.counter = m.counters.p; fo.scal = scal.p; fo.ellv = ellv_buf.p; fo.cv = cv_buf.p;
    ScanApplyParams<T, FwdElem<T>, FwdOut<T>> pc; pc.n = K; pc.elem = el; pc.out = fo; pc.pre = m.preaff.p; pc.nchunks = nch;
    if (m.scan_e == 8) launch<ScanApplyBody<T, FwdElem<T>, FwdOut<T>, 8>>(m.nchunksK, SCAN_NT, m.scan_smem(), st, pc); else launch<ScanApplyBody<T, FwdElem<T>, FwdOut<T>, 4>>(m.nchunksK, SCAN_NT, m.scan_smem(), st, pc);
    AmpTabParams<T> pt2; pt2.ch = SCAN_NT * m.scan_e; pt2.m = m.am; pt2.P = Pb.p; pt2.amp = amp.p; pt2.wS = wS.p; pt2.partials = m.partials.p;
    pt2.counter = m.counters.p + 1; pt2.scal = scal.p; pt2.ellv = ellv; pt2.cv = cv;
    launch<AmpTabBody<T>>(m.nchunksK, SCAN_NT, 512, st, pt2);
  }

  // tangent chain: du table + (cj, da0) scalars for ProMetric
  void amp_tangent(stream_t st, const T* t) {
    Model<T>& m = *M; const int K = m.am.K;
    JvpElem<T> el; el.m = m.am; el.pos = pos.p; el.t = t; el.scal = scal.p;
    if (m.chain.ok) {
      TanChainParams<T> pp; pp.elem = el;
      JvpOut<T>& jo = pp.out; jo.m = m.am; jo.pos = pos.p; jo.t = t; jo.wS = wS.p; jo.amp = amp.p; jo.ellv = ellv; jo.cv = cv; jo.ad = m.ad.p;
      jo.partials = m.partials.p; jo.counter = m.counters.p + 2; jo.scal = scal.p;
      pp.n = K; pp.per = m.chain.per_t; pp.agg = m.cagg.p; pp.bar = m.counters.p + 8; pp.nph = m.chain.nph_t;
      launch_coop<TanChainBody<T, 4>>(m.chain.grid_t, SCAN_NT, m.chain.smem_t, st, pp);
      return;
    }
    const int nch = m.am.has_dev ? m.nchunksK : 1;
    if (m.am.has_dev && m.nchunksK > 1) {
      ScanAggParams<T, JvpElem<T>> pa; pa.n = K; pa.elem = el; pa.agg = m.agg.p; pa.pre = m.preaff.p; pa.counter = m.counters.p + 5;
      if (m.scan_e == 8) launch<ScanAggBody<T, JvpElem<T>, 8>>(m.nchunksK, SCAN_NT, m.scan_smem(), st, pa); else launch<ScanAggBody<T, JvpElem<T>, 4>>(m.nchunksK, SCAN_NT, m.scan_smem(), st, pa);
    }
    JvpOut<T> jo; jo.m = m.am; jo.pos = pos.p; jo.t = t; jo.wS = wS.p; jo.amp = amp.p; jo.ellv = ellv; jo.cv = cv; jo.ad = m.ad.p;
    jo.partials = m.partials.p; jo.counter = m.counters.p + 2; jo.scal = scal.p;
    ScanApplyParams<T, JvpElem<T>, JvpOut<T>> pc; pc.n = K; pc.elem = el; pc.out = jo; pc.pre = m.preaff.p; pc.nchunks = nch;
    if (m.scan_e == 8) launch<ScanApplyBody<T, JvpElem<T>, JvpOut<T>, 8>>(m.nchunksK, SCAN_NT, m.scan_smem(), st, pc); else launch<ScanApplyBody<T, JvpElem<T>, JvpOut<T>, 4>>(m.nchunksK, SCAN_NT, m.scan_smem(), st, pc);
  }
  ProMetric<T> pro_metric(const T* t) const {
    const Model<T>& m = *M;
    ProMetric<T> pro; pro.xi = pos.p + m.am.off_xi; pro.t = t + m.am.off_xi; pro.idxf = m.P->idxf.p;
    pro.ad = m.ad.p; pro.scal = scal.p + SC_CJ; pro.kappa = m.am.kind_power ? T(0.5) : T(1); pro.fg = m.P->fold_geom();
    return pro;
  }

  // cotangent chain after P5 filled W: segment sum, reverse scan, scalar leaves, dot product.
  // Distributed plans run it in two halves around an all-reduce of the bin sums (abar) and of the two
  // scalars xs = {sum of the position-space cotangent, xi-block of <add, out>}.
  void seg_sum(stream_t st, T* abar_out, const T* abar_in) {
    Model<T>& m = *M; Plan<T>& P = *m.P;
    SegSumParams<T> ps; ps.m = m.am; ps.W = P.W.p; ps.order = P.w_order.p; ps.offs = P.w_offs.p; ps.amp = amp.p;
    ps.g = abar_out ? nullptr : m.gbuf.p; ps.abar = abar_out; ps.abar_in = abar_in; ps.partials = m.partials.p;
    ps.counter = m.counters.p + 3; ps.scal = scal.p; ps.lg_lpb = P.seg_lg_lpb; ps.ellv = ellv; ps.cv = cv;
    launch<SegSumBody<T>>(P.seg_grid(), 256, (256 + 96) * sizeof(T), st, ps);
  }
  void vjp_chain(stream_t st, T* out, const T* add, const T* p3_src, int n_p3, const T* p5_src, int n_p5, T scl_factor) {
    Model<T>& m = *M; const int K = m.am.K;
    const long nj = m.am.has_dev ? (long)K - 2 : 0;
    VjpOut<T> vo; vo.m = m.am; vo.pos = pos.p; vo.g = m.gbuf.p; vo.wS = wS.p; vo.scal_in = scal.p; vo.out = out; vo.add = add;
    vo.partials = m.partials.p; vo.counter = m.counters.p + 4; vo.scal = scal.p;
    vo.p3_partials = p3_src; vo.n_p3 = n_p3; vo.p5_partials = p5_src; vo.n_p5 = n_p5; vo.scl_factor = scl_factor;
    if (nj > 0) {
      VjpElem<T> el; el.m = m.am; el.g = m.gbuf.p; el.wS = wS.p; el.scal = scal.p;
      if (m.nchunksJ > 1) {
        ScanAggParams<T, VjpElem<T>> pa; pa.n = nj; pa.elem = el; pa.agg = m.agg.p; pa.pre = m.preaff.p; pa.counter = m.counters.p + 5;
        if (m.scan_e == 8) launch<ScanAggBody<T, VjpElem<T>, 8>>(m.nchunksJ, SCAN_NT, m.scan_smem(), st, pa); else launch<ScanAggBody<T, VjpElem<T>, 4>>(m.nchunksJ, SCAN_NT, m.scan_smem(), st, pa);
      }
      ScanApplyParams<T, VjpElem<T>, VjpOut<T>> pc; pc.n = nj; pc.elem = el; pc.out = vo; pc.pre = m.preaff.p; pc.nchunks = m.nchunksJ;
      if (m.scan_e == 8) launch<ScanApplyBody<T, VjpElem<T>, VjpOut<T>, 8>>(m.nchunksJ, SCAN_NT, m.scan_smem(), st, pc); else launch<ScanApplyBody<T, VjpElem<T>, VjpOut<T>, 4>>(m.nchunksJ, SCAN_NT, m.scan_smem(), st, pc);
    } else {
      ScanApplyParams<T, NoElem<T>, VjpOut<T>> pc; pc.n = 0; pc.out = vo; pc.pre = nullptr; pc.nchunks = 1;
      if (m.scan_e == 8) launch<ScanApplyBody<T, NoElem<T>, VjpOut<T>, 8>>(1, SCAN_NT, m.scan_smem(), st, pc); else launch<ScanApplyBody<T, NoElem<T>, VjpOut<T>, 4>>(1, SCAN_NT, m.scan_smem(), st, pc);
    }
  }
  void amp_cotangent(stream_t st, T* out, const T* add, int p3_col, T scl_factor, bool use_p5_dot) {
    Plan<T>& P = *M->P;
    if (M->chain.ok) {
      Model<T>& m = *M; const int K = m.am.K;
      CotChainParams<T> pp;
      VjpOut<T>& vo = pp.out; vo.m = m.am; vo.pos = pos.p; vo.g = m.gbuf.p; vo.wS = wS.p; vo.scal_in = scal.p; vo.out = out; vo.add = add;
      vo.partials = m.partials.p; vo.counter = m.counters.p + 4; vo.scal = scal.p;
      vo.p3_partials = P.p3part.p + p3_col; vo.n_p3 = P.n3part; vo.p5_partials = P.p5part.p; vo.n_p5 = use_p5_dot ? P.n5part : 0; vo.scl_factor = scl_factor;
      pp.nj = m.am.has_dev ? (long)K - 2 : 0; pp.per = m.chain.per_c; pp.rel = m.crel.p; pp.agg = m.cagg.p; pp.bar = m.counters.p + 10; pp.nph = m.chain.nph_c;
      pp.W = P.W.p; pp.order = P.w_order.p; pp.offs = P.w_offs.p; pp.amp = amp.p; pp.ellv = ellv; pp.cv = cv; pp.g = m.gbuf.p;
      pp.segpart = m.partials.p + 4096; pp.scal = scal.p;
      pp.seg_b0 = m.seg_b0.p; pp.seg_lg = m.seg_lg.p;
      pp.dbg = m.cdbg.n ? m.cdbg.p : nullptr; pp.dbg_twice = m.chain.dbg_twice;
      launch_coop<CotChainBody<T, 4>>(m.chain.grid_c, SCAN_NT, m.chain.smem_c, st, pp);
#ifndef NB_EMU
      if (m.cdbg.n) {       // developer aid: where the CTAs spend their time (one line per launch on stderr)
        std::vector<long long> h(m.cdbg.n);
        stream_sync(st); d2h(h.data(), m.cdbg.p, h.size() * sizeof(long long), st); stream_sync(st);
        const int G = m.chain.grid_c; long long t0 = h[0];
        for (int c = 0; c < G; ++c) t0 = std::min(t0, h[(size_t)c * 12]);
        std::fprintf(stderr, "[nb200 cot-chain] grid %d, us since first CTA start (min/avg/max over CTAs):", G);
        const char* nm[11] = {"start", "summed", "seg-done", "ph1-start", "ph1-done", "ph2-start", "ph2-done", "-", "again-start", "again-summed", "again-done"};
        for (int s = 0; s < 11; ++s) {
          if (s == 7 || (s > 7 && !m.chain.dbg_twice)) continue;
          double mn = 1e30, mx = -1e30, av = 0;
          for (int c = 0; c < G; ++c) { double v = (h[(size_t)c * 12 + s] - t0) * 1e-3; mn = std::min(mn, v); mx = std::max(mx, v); av += v / G; }
          std::fprintf(stderr, "  %s %.1f/%.1f/%.1f", nm[s], mn, av, mx);
        }
        std::fprintf(stderr, "\n");
      }
#endif
      return;
    }
    seg_sum(st, nullptr, nullptr);
    vjp_chain(st, out, add, P.p3part.p + p3_col, P.n3part, P.p5part.p, use_p5_dot ? P.n5part : 0, scl_factor);
  }
  EpiAdjoint<T> epi_adjoint(T* out, const T* add, bool want_dot) const {
    const Model<T>& m = *M;
    EpiAdjoint<T> e; e.out = out + m.am.off_xi; e.add = add ? add + m.am.off_xi : nullptr; e.xi = pos.p + m.am.off_xi;
    e.idxf = m.P->idxf.p; e.amp = amp.p; e.W = m.P->W.p; e.invV = T(1.0 / m.P->g.V);
    e.partials = want_dot ? m.P->p5part.p : nullptr;
    return e;
  }

  void update(stream_t st, const T* pos_in, T* grad, bool add_prior) {
    Model<T>& m = *M; Plan<T>& P = *m.P;
    if (!m.have_lh) throw Error{"nb200: likelihood not set on this model"};
    d2d(pos.p, pos_in, (size_t)m.am.L * sizeof(T), st);
    amp_forward(st);
    ProAmp<T> pro; pro.xi = pos.p + m.am.off_xi; pro.idxf = P.idxf.p; pro.amp = amp.p; pro.fg = P.fold_geom();
    P.run_p1(st, pro); P.run_pc(st, false);
    PointOp<T> op = P.make_op(PM_LINEARIZE);
    op.invV = T(1.0 / P.g.V); op.offset = m.offset_mean; op.sc_ptr = scal.p + SC_SCALING;
    op.lh_kind = m.lh_kind; op.nl_exp = m.nl_exp; op.data = m.data.p; op.w_scalar = m.w_scalar;
    if (m.nl_exp == 2) {
      if (!nl_ready) throw Error{"nb200: this model has a tabulated non-linearity: call nb200_lin_set_pointwise before nb200_lin_update"};
      op.nl_s = nl_s.p; op.nl_ds = nl_ds.p; nl_ready = false;       // (valid for this position only)
    }
    op.w_arr = m.has_w_arr ? m.w_arr.p : nullptr; op.s_out = s.p; op.jl_out = jl.p; op.partials = P.p3part.p;
    if (grad) P.template run_p3<true, true>(st, op); else P.template run_p3<true, false>(st, op);
    ReduceColsParams<T> pr; pr.partials = P.p3part.p; pr.n = P.n3part; pr.ncol = 2;
    pr.out0 = scal.p + SC_ENERGY; pr.out1 = scal.p + SC_SUMCOT;
    launch<ReduceColsBody<T>>(1, 256, 512, st, pr);
    if (grad) {
      P.run_pc(st, true);
      P.run_p5(st, epi_adjoint(grad, add_prior ? pos.p : nullptr, false));
      amp_cotangent(st, grad, add_prior ? pos.p : nullptr, 1, m.am.scl_b, false);
    }
    valid = true;
  }

  // out = scale * J_a^T l_a l_b J_b t (+ add); this == lin_a (cotangent side), b == tangent side.
  // `add` may be t (the "+ 1" of the Hamiltonian metric), `out` itself (accumulate the products of several
  // linearisations in place: every entry is read and written by the same thread) or null; `want_dot` leaves
  // <add, out> in scal[SC_DOT] (meaningful for add == t, scale == 1: the CG curvature).
  // `piece_done(first, count)` (optional, needs !want_dot): the last pass runs in plan.reduce_chunks launches over ranges of
  // axis-0 planes and reports every finished range of `out` (in elements) right after the launch that completes it --
  // the caller's all-reduce of that range then overlaps the remaining launches; the hyper-parameter entries follow last.
  typedef std::function<void(long, long)> PieceFn;
  void metric_ex(stream_t st, Lin<T>* b, const T* t, T* out, const T* add, bool want_dot, T scale, const PieceFn* piece_done = nullptr) {
    Model<T>& m = *M; Plan<T>& P = *m.P;
    if (!valid || !b->valid) throw Error{"nb200: linearisation not initialised (call nb200_lin_update)"};
    ChainScope<T> chain(P, P.chain_ok);      // every pass of this sequence runs staged (row-major intermediates)
    b->amp_tangent(st, t);
    P.run_p1(st, b->pro_metric(t)); P.run_pc(st, false);
    PointOp<T> op = P.make_op(PM_METRIC);
    op.invV = T(1.0 / P.g.V) * scale; op.jl_a = jl.p; op.jl_b = b->jl.p; op.partials = P.p3part.p;
    if (m.am.has_scaling) { op.cshift_ptr = t + m.am.off_scl; op.cshift_scale = m.am.scl_b * scale; }
    P.template run_p3<true, true>(st, op);
    P.run_pc(st, true);
    const int nch = (piece_done && !want_dot && !P.dist) ? std::min(P.reduce_chunks, P.g.h0 + 1) : 1;
    if (nch <= 1) {
      P.run_p5(st, epi_adjoint(out, add, want_dot));
      amp_cotangent(st, out, add, 0, m.am.scl_b, want_dot);
      if (piece_done) (*piece_done)(0, m.am.L);
      return;
    }
    const long plane = (long)P.g.nm * P.g.nl;          // elements of one axis-0 plane of the excitations
    const int n_a = P.g.n0, h_a = P.g.h0, lgm = P.mg5().lg_mid;
    for (int c = 0; c < nch; ++c) {
      // lines of the planes a0 .. a1-1 of the half range; they also complete the mirror planes n_a - a (a != 0, h_a)
      const int a0 = Plan<T>::split_at(0, h_a + 1, c, nch), a1 = Plan<T>::split_at(0, h_a + 1, c + 1, nch);
      P.run_p5(st, epi_adjoint(out, add, false), a0 << lgm, (a1 - a0) << lgm);
      const int m0 = std::max(n_a - a1 + 1, h_a + 1), m1 = std::min(n_a - a0 + 1, n_a);      // mirror planes [m0, m1)
      if (a1 > a0 && m1 > m0 && m0 == a1) { (*piece_done)(m.am.off_xi + a0 * plane, (long)(m1 - a0) * plane); continue; }
      if (a1 > a0) (*piece_done)(m.am.off_xi + a0 * plane, (long)(a1 - a0) * plane);
      if (m1 > m0) (*piece_done)(m.am.off_xi + m0 * plane, (long)(m1 - m0) * plane);
    }
    amp_cotangent(st, out, add, 0, m.am.scl_b, false);
    if (m.am.off_xi > 0) (*piece_done)(0, m.am.off_xi);
    const long tail = m.am.off_xi + (long)n_a * plane;
    if (m.am.L > tail) (*piece_done)(tail, m.am.L - tail);
  }
  void metric(stream_t st, Lin<T>* b, const T* t, T* out, bool add_identity) {
    metric_ex(st, b, t, out, add_identity ? t : nullptr, add_identity, T(1));
  }

  void rsm(stream_t st, const T* t, T* out_nat, bool scaled) {
    Model<T>& m = *M; Plan<T>& P = *m.P;
    if (!valid) throw Error{"nb200: linearisation not initialised (call nb200_lin_update)"};
    amp_tangent(st, t);
    P.run_p1(st, pro_metric(t)); P.run_pc(st, false);
    PointOp<T> op = P.make_op(PM_JVP_OUT);
    op.invV = T(1.0 / P.g.V); op.natural = 1; op.pos_out = out_nat; op.jl_a = scaled ? jl.p : nullptr;
    if (scaled && m.am.has_scaling) { op.cshift_ptr = t + m.am.off_scl; op.cshift_scale = m.am.scl_b; }
    P.template run_p3<true, false>(st, op);
  }
  void lsm(stream_t st, const T* u_nat, T* out, bool scaled) {
    Model<T>& m = *M; Plan<T>& P = *m.P;
    if (!valid) throw Error{"nb200: linearisation not initialised (call nb200_lin_update)"};
    PointOp<T> op = P.make_op(PM_LOAD);
    op.natural = 1; op.in_pos = u_nat; op.jl_a = scaled ? jl.p : nullptr; op.partials = P.p3part.p;
    P.template run_p3<false, true>(st, op);
    P.run_pc(st, true);
    P.run_p5(st, epi_adjoint(out, nullptr, false));
    amp_cotangent(st, out, nullptr, 0, scaled ? m.am.scl_b : T(0), false);
  }
  // ---- slab-decomposed plans: the operator sequences cut at the two exchanges / the all-reduce -------------
  // code: 0 update.1 (amplitude, P1, PCa) | 1 update.2 (P3 linearise; local energy) | 2 metric.1 (tangent chain, P1, PCa)
  //       3 metric.2 (P3 fused, PCb) | 4 lsm.2 (P3 adjoint-only from local T-layout `in`, PCb) | 5 adjoint.3 (P5, local bin sums,
  //       xs = {p3 sum, xi dot}) | 6 adjoint.4 (finish with all-reduced abar / xs: hyper-parameter leaves, <add,out>)
  //       7 rsm.2 (P3 forward-only into local planes; after metric.1 + exchange 1); 1x / 2x: the chunked forms
  void dist_phase(stream_t st, int code, Lin<T>* b, const T* in, T* out, T* abar, T* xs, int flag, int chunk = -1) {
    Model<T>& m = *M; Plan<T>& P = *m.P;
    if (!P.dist) throw Error{"nb200_dist_phase: not a slab-decomposed plan"};
    if (!P.xS0 || !P.xS1 || !P.xS2) throw Error{"nb200_dist_phase: exchange buffers not set (nb200_plan_set_scratch)"};
    if (chunk >= P.nchunks) throw Error{"nb200_dist_phase: chunk index out of range"};
    // chunk < 0: the whole range in one launch
    auto pc = [&](bool second) {
      if (chunk < 0) { P.run_pc(st, second); return; }
      const std::vector<int>& off = second ? P.omapB_off : P.omapA_off;
      P.run_pc(st, second, (second ? P.omapB.p : P.omapA.p) + off[chunk], off[chunk + 1] - off[chunk]);
    };
    const int l3 = chunk < 0 ? 0 : P.p3_line0[chunk], n3 = chunk < 0 ? -1 : P.p3_nlines[chunk];
    const int l5 = chunk < 0 ? 0 : P.p5_line0[chunk], n5 = chunk < 0 ? -1 : P.p5_nlines[chunk];
    auto reduce_p3 = [&](T* o0, T* o1) {
      ReduceColsParams<T> pr; pr.partials = P.p3part.p; pr.n = P.n3part; pr.ncol = 2; pr.out0 = o0; pr.out1 = o1;
      launch<ReduceColsBody<T>>(1, 256, 512, st, pr);
    };
    switch (code) {
      case 0: case 10: {     // update.1: amplitude chain + P1 (+ PCa for code 0)
        if (!m.have_lh) throw Error{"nb200: likelihood not set on this model"};
        if (m.nl_exp == 2) throw Error{"nb200: tabulated non-linearities are not available on slab-decomposed plans"};
        d2d(pos.p, in, (size_t)m.am.L * sizeof(T), st);
        amp_forward(st);
        ProAmp<T> pro; pro.xi = pos.p + m.am.off_xi; pro.idxf = P.idxf.p; pro.amp = amp.p; pro.fg = P.fold_geom();
        P.run_p1(st, pro);
        if (code == 0) P.run_pc(st, false);
      } break;
      case 11: pc(false); break;                        // PCa (chunk)
      case 13: pc(true); break;                         // PCb (chunk)
      case 1: case 12: {     // update.2: P3 linearise (chunk); code 1 also reduces the local energy
        PointOp<T> op = P.make_op(PM_LINEARIZE);
        op.invV = T(1.0 / P.g.V); op.offset = m.offset_mean; op.sc_ptr = scal.p + SC_SCALING;
        op.lh_kind = m.lh_kind; op.nl_exp = m.nl_exp; op.data = m.data.p; op.w_scalar = m.w_scalar;
        op.w_arr = m.has_w_arr ? m.w_arr.p : nullptr; op.s_out = s.p; op.jl_out = jl.p; op.partials = P.p3part.p;
        if (flag & 1) P.template run_p3<true, true>(st, op, l3, n3);      // + first adjoint pass of dE/df (gradient)
        else P.template run_p3<true, false>(st, op, l3, n3);
        if (code == 1) { reduce_p3(scal.p + SC_ENERGY, scal.p + SC_SUMCOT); valid = true; }
      } break;
      case 14: reduce_p3(scal.p + SC_ENERGY, scal.p + SC_SUMCOT); valid = true; break;
      case 2: case 20: {     // metric.1: tangent chain + P1 (+ PCa for code 2)
        if (!valid) throw Error{"nb200: linearisation not initialised"};
        amp_tangent(st, in);
        P.run_p1(st, pro_metric(in));
        if (code == 2) P.run_pc(st, false);
      } break;
      case 3: case 21: {     // metric.2: fused P3 (chunk) (+ PCb for code 3)
        PointOp<T> op = P.make_op(PM_METRIC);
        op.invV = T(1.0 / P.g.V); op.jl_a = jl.p; op.jl_b = b->jl.p; op.partials = P.p3part.p;
        if (m.am.has_scaling) { op.cshift_ptr = in + m.am.off_scl; op.cshift_scale = m.am.scl_b; }
        P.template run_p3<true, true>(st, op, l3, n3);
        if (code == 3) P.run_pc(st, true);
      } break;
      case 4: case 22: {     // lsm.2: adjoint-only P3 from the local planes `in` (chunk) (+ PCb for code 4)
        PointOp<T> op = P.make_op(PM_LOAD);
        op.in_pos = in; op.jl_a = flag ? jl.p : nullptr; op.partials = P.p3part.p;
        P.template run_p3<false, true>(st, op, l3, n3);
        if (code == 4) P.run_pc(st, true);
      } break;
      case 7: case 25: {     // rsm.2: forward-only P3 into the local planes `out` (chunk); flag bit 0: scaled (l * d signal)
        PointOp<T> op = P.make_op(PM_JVP_OUT);
        op.invV = T(1.0 / P.g.V); op.natural = 0; op.pos_out = out; op.jl_a = (flag & 1) ? jl.p : nullptr;
        if ((flag & 1) && m.am.has_scaling) { op.cshift_ptr = in + m.am.off_scl; op.cshift_scale = m.am.scl_b; }
        P.template run_p3<true, false>(st, op, l3, n3);
      } break;
      case 23: P.run_p5(st, epi_adjoint(out, (flag & 1) ? in : nullptr, (flag & 1) != 0), l5, n5); break;   // P5 (chunk)
      case 5: case 24: {     // adjoint.3: (P5 for code 5) local bin sums, xs = {p3 sum, xi dot}
        if (code == 5) P.run_p5(st, epi_adjoint(out, (flag & 1) ? in : nullptr, (flag & 1) != 0));
        seg_sum(st, abar, nullptr);
        if (flag & 4) reduce_p3(nullptr, xs); else reduce_p3(xs, nullptr);   // bit 2: the gradient needs sum dE/df (column 1)
        if (flag & 1) {
          ReduceColsParams<T> pd; pd.partials = P.p5part.p; pd.n = P.n5part; pd.ncol = 1; pd.out0 = xs + 1; pd.out1 = nullptr;
          launch<ReduceColsBody<T>>(1, 256, 512, st, pd);
        } else {
          dev_zero(xs + 1, sizeof(T), st);
        }
      } break;
      case 6: {              // adjoint.4: finish with the all-reduced abar / xs
        seg_sum(st, nullptr, abar);
        T sf = (flag & 2) ? m.am.scl_b : T(0);
        vjp_chain(st, out, (flag & 1) ? in : nullptr, xs, 1, xs + 1, 1, sf);
      } break;
      default: throw Error{"nb200_dist_phase: unknown phase code"};
    }
  }

  // signal and its derivative with respect to the field, evaluated by the host at the position of the next update
  // (natural layout) -> T-layout tables
  void set_pointwise(stream_t st, const T* s_nat, const T* ds_nat) {
    Model<T>& m = *M; Plan<T>& P = *m.P;
    if (P.dist) throw Error{"nb200: tabulated non-linearities are not available on slab-decomposed plans"};
    if (m.nl_exp != 2) throw Error{"nb200_lin_set_pointwise: the model's non-linearity is not `tabulated` (nonlinearity = 2)"};
    if (!nl_s.n) { nl_s.alloc((size_t)P.local_position_grid()); nl_ds.alloc((size_t)P.local_position_grid()); }
    P.run_rev(st, s_nat, nl_s.p, true);
    P.run_rev(st, ds_nat, nl_ds.p, true);
    nl_ready = true;
  }

  void posmap(stream_t st, int mode, T* out_nat) {
    Model<T>& m = *M; Plan<T>& P = *m.P;
    if (!valid) throw Error{"nb200: linearisation not initialised (call nb200_lin_update)"};
    PosMapParams<T> pm; pm.mode = mode; pm.lh_kind = m.lh_kind; pm.s = s.p; pm.data = m.data.p;
    pm.n = P.dist ? (long)P.planes2 * P.g.nm * P.g.n0 : (long)P.g.N;       // slab-decomposed: the local planes only
    pm.w_scalar = m.w_scalar; pm.w_arr = m.has_w_arr ? m.w_arr.p : nullptr; pm.out = m.tmp_pos.p;
    int grid = (int)std::min<int64_t>((pm.n + 255) / 256, 148 * 8);
    if (P.dist) pm.out = out_nat;      // slab-decomposed plans hand out the local planes in the internal layout
    launch<PosMapBody<T>>(grid, 256, 0, st, pm);
    if (!P.dist) P.run_rev(st, m.tmp_pos.p, out_nat, false);
  }
};

}  // namespace nb
