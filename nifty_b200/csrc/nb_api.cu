// nifty_b200 -- C ABI (include/nifty_b200.h).  Compiled by nvcc for sm_100a into libniftyb200.so;
// tests/emu compiles the same file with -DNB_EMU as plain C++ (sequential host execution of the
// kernel bodies) to check host logic and index arithmetic where no GPU is available.
#include "nb_model.cuh"
#include "nb_vec.cuh"
#include "nb_bluestein.cuh"
#include <limits>
#include <memory>

using namespace nb;

struct nb200_plan { PlanBase* impl; };
struct nb200_model { ModelBase* impl; nb200_plan* plan; int dtype; };
struct nb200_lin { LinBase* impl; nb200_model* model; int dtype; };

namespace nb {

template <class T> void Model<T>::init(Plan<T>* plan_, const nb200_model_desc& d) {
  P = plan_; plan = plan_;
  const GridInfo& g = P->g;
  std::memset(&am, 0, sizeof(am));
  am.K = g.K; am.kind_power = d.kind_power; am.has_flu = d.has_fluctuations; am.has_asp = d.has_asperity;
  am.has_dev = d.has_deviations && g.K > 2; am.has_scaling = d.has_scaling;
  am.matern = d.amplitude_type == 1; am.renorm = d.renormalize_amplitude;
  am.ctf_a = (T)d.cutoff_a; am.ctf_b = (T)d.cutoff_b; am.off_ctf = d.off_cutoff;
  if (d.amplitude_type != 0 && d.amplitude_type != 1) throw Error{"nb200: amplitude_type must be 0 (non-parametric) or 1 (Matern)"};
  if (am.matern && am.has_dev) throw Error{"nb200: the Matern amplitude has no spectrum deviations"};
  am.V = (T)g.V;
  am.flu_a = (T)d.fluct_a; am.flu_b = (T)d.fluct_b; am.slp_a = (T)d.slope_a; am.slp_b = (T)d.slope_b;
  am.flx_a = (T)d.flex_a; am.flx_b = (T)d.flex_b; am.asp_a = (T)d.asp_a; am.asp_b = (T)d.asp_b;
  am.zm_a = (T)d.zeromode_a; am.zm_b = (T)d.zeromode_b; am.scl_a = (T)d.scaling_a; am.scl_b = (T)d.scaling_b;
  am.off_flu = d.off_fluct; am.off_slp = d.off_slope; am.off_flx = d.off_flex; am.off_asp = d.off_asp;
  am.off_spec = d.off_spectrum; am.off_zm = d.off_zeromode; am.off_xi = d.off_xi; am.off_scl = d.off_scaling;
  am.L = d.latent_size;
  if (d.has_deviations && g.K <= 2) throw Error{"nb200: has_deviations set but the grid has <= 2 unique modes"};
  auto chk = [&](int64_t off, int64_t len, const char* nm) {
    if (off < 0 || off + len > d.latent_size) throw Error{std::string("nb200: leaf offset out of range: ") + nm};
  };
  chk(d.off_xi, P->local_latent_grid(), "xi"); chk(d.off_zeromode, 1, "zeromode"); chk(d.off_slope, 1, "loglogavgslope");
  if (am.has_flu) chk(d.off_fluct, 1, "fluctuations");
  if (am.has_dev) { chk(d.off_flex, 1, "flexibility"); chk(d.off_spectrum, 2 * (int64_t)(g.K - 2), "spectrum"); }
  if (am.has_dev && am.has_asp) chk(d.off_asp, 1, "asperity");
  if (am.has_scaling) chk(d.off_scaling, 1, "scaling");
  if (am.matern) chk(d.off_cutoff, 1, "cutoff");
  offset_mean = (T)d.offset_mean;
  std::vector<T> e(g.K), mu(g.K), dtv(g.logvol.size());
  for (int b = 0; b < g.K; ++b) { e[b] = (T)g.rel[b]; mu[b] = (T)g.mult[b]; }
  for (size_t j = 0; j < dtv.size(); ++j) dtv[j] = (T)g.logvol[j];
  ell.upload(e); multT.upload(mu); dt.upload(dtv);
  am.ell = ell.p; am.mult = multT.p; am.dt = dt.p;
  { std::vector<T> kk(g.K); for (int b = 0; b < g.K; ++b) kk[b] = (T)g.um[b]; modes.upload(kk); am.modes = modes.p; }
  scan_e = scan_pick_e(g.K);
  const int scan_ch = SCAN_NT * scan_e;
  nchunksK = (g.K + scan_ch - 1) / scan_ch;
  nchunksJ = std::max(1, (g.K - 2 + scan_ch - 1) / scan_ch);
  agg.alloc(nchunksK + 1); preaff.alloc(nchunksK + 2);
  ad.alloc(g.K); gbuf.alloc(g.K);
  size_t np = std::max<size_t>(3 * (size_t)nchunksK + 3, 3 * (size_t)P->seg_grid() + 3);
  np = std::max<size_t>(np, 4 * 2048);
  partials.alloc(np);
  counters.alloc(16);
  tmp_pos.alloc((size_t)P->local_position_grid());
  init_chain();
  scratch_lin = new Lin<T>();
  scratch_lin->init(this);
}

// launch geometry of the fused chain kernels (nb_chain.cuh): all CTAs must be resident (cooperative launch)
template <class T> void Model<T>::init_chain() {
  const GridInfo& g = P->g;
  chain = ChainCfg();
  if (P->dist) return;                                // slab-decomposed plans all-reduce the bin sums between the halves
  if (const char* e = std::getenv("NB200_COOP")) { if (e[0] == '0') return; }
  constexpr int E = 4, CH = SCAN_NT * E;
  const int sms = P->sms > 0 ? P->sms : sm_count();
  const long K = g.K, nj = am.has_dev ? K - 2 : 0;
  long gcap = 1L << 30;                               // test knob: few CTAs -> several rounds per CTA
  if (const char* e = std::getenv("NB200_COOP_MAXGRID")) gcap = std::max(1L, std::atol(e));
  // tangent
  chain.smem_t = TanChainBody<T, E>::smem_bytes();
  const int rt = coop_blocks_per_sm<TanChainBody<T, E>>(chain.smem_t);
  if (rt < 1) return;
  {
    const long nch = (K + CH - 1) / CH, gmax = std::min(gcap, (long)rt * sms);
    const long rounds = (nch + gmax - 1) / gmax;
    chain.per_t = rounds * CH;
    chain.grid_t = (int)((K + chain.per_t - 1) / chain.per_t);
  }
  // cotangent: the scan range of a CTA and its range of bins in the segment sum
  chain.smem_c = CotChainBody<T, E>::smem_bytes();
  const int rc = coop_blocks_per_sm<CotChainBody<T, E>>(chain.smem_c);
  if (rc < 1) return;
  {
    const long nchj = std::max<long>(1, (nj + CH - 1) / CH), gmax = std::min(gcap, (long)rc * sms);
    const long rounds = (nchj + gmax - 1) / gmax;
    chain.per_c = rounds * CH;
    const long gscan = std::max<long>(1, (nj + chain.per_c - 1) / chain.per_c);
    // segment sum: spread the bins over all resident CTAs (at least 64 bins each)
    chain.grid_c = (int)std::max<long>(gscan, std::min<long>(gmax, (K + 63) / 64));
    // bin ranges of equal cost (one unit per W position + two per bin), lanes per bin from the mean population of the range
    std::vector<int> b0((size_t)chain.grid_c + 1, (int)K), lgv((size_t)chain.grid_c, 0);
    const double total = (double)g.w_offs[K] + 2.0 * (double)K;
    long b = 0;
    for (int c = 0; c < chain.grid_c; ++c) {
      b0[c] = (int)b;
      const double target = total * (c + 1) / chain.grid_c;
      while (b < K && (double)g.w_offs[b + 1] + 2.0 * (double)(b + 1) <= target) ++b;
      if (c == chain.grid_c - 1) b = K;
      const long nbin = b - b0[c];
      const double mean = nbin > 0 ? (double)(g.w_offs[b] - g.w_offs[b0[c]]) / (double)nbin : 0.0;
      int lg = 0;
      while (lg < 5 && (double)(4 << lg) < mean) ++lg;       // <= ~4 positions per lane
      lgv[c] = lg;
    }
    b0[chain.grid_c] = (int)K;
    seg_b0.upload(b0); seg_lg.upload(lgv);
  }
  cagg.alloc((size_t)std::max(chain.grid_t, chain.grid_c) + 8);
  crel.alloc((size_t)std::max<long>(nj, 1));
  if (const char* e = std::getenv("NB200_COOP_DBG")) { if (e[0] >= '1') cdbg.alloc((size_t)chain.grid_c * 12); chain.dbg_twice = e[0] == '2'; }
  if (const char* e = std::getenv("NB200_COOP_NPH_T")) chain.nph_t = std::atoi(e);
  if (const char* e = std::getenv("NB200_COOP_NPH_C")) chain.nph_c = std::atoi(e);
  chain.ok = true;
}
template <class T> Model<T>::~Model() { delete scratch_lin; }

// ---- conjugate gradient driver ------------------------------------------------------------------
template <class T> struct CgWork {
  DevBuf<T> r, d, q, tm, cgs, partials;
  DevBuf<int> cgi;
  DevBuf<unsigned> counter;
  long L = 0;
  void ensure(long n) {
    if (L == n) return;
    r.alloc(n); d.alloc(n); q.alloc(n); tm.alloc(n); cgs.alloc(CG_NSCAL); partials.alloc(4 * 2048); cgi.alloc(CGI_NINT); counter.alloc(4);
    L = n;
  }
};
template <class T> CgWork<T>& cg_work() { static thread_local CgWork<T> w; return w; }

// out = sum_i scale * M_i t (+ t on the caller that owns the identity), then the caller's reduction hook (an in-stream
// all-reduce over the ranks that hold the other sample points): the sample-averaged metric of `_kl_met`
// (nifty/re/optimize_kl.py:117-144) without leaving the device.  `ws` is unused scratch for future use.
template <class T>
void multi_metric(Lin<T>** lins, int n, T scale, bool identity_here, stream_t st, const T* t, T* out, nb200_reduce_hook hook, void* user) {
  if (n <= 0) throw Error{"nb200: multi_metric needs at least one linearisation (ranks without samples still pass one and scale 0)"};
  typename Lin<T>::PieceFn piece = [&](long first, long count) { hook(user, out + first, (int64_t)count, (void*)st); };
  for (int i = 0; i < n; ++i) {
    const T* add = (i == 0) ? (identity_here ? t : nullptr) : out;
    lins[i]->metric_ex(st, lins[i], t, out, add, false, scale, (hook && i == n - 1) ? &piece : nullptr);
  }
  if (hook) hook(user, nullptr, 0, (void*)st);       // every range has been handed over: results must be complete in stream order
}

template <class T>
int cg_solve(Lin<T>* lin, Lin<T>* lin_b, stream_t st, const T* j, T* x, const nb200_cg_opts& o, nb200_cg_result* res,
             Lin<T>** multi = nullptr, int n_multi = 0, T scale = T(1), bool identity_here = true, nb200_reduce_hook hook = nullptr,
             void* user = nullptr) {
  Model<T>& m = *lin->M;
  const long L = m.am.L;
  CgWork<T>& w = cg_work<T>();
  w.ensure(L);
  const int vgrid = (int)std::min<long>((L + 1023) / 1024, 148 * 8);
  int norm_ord = o.norm_ord;
  if (norm_ord != 0 && norm_ord != 1 && norm_ord != 2) return fail("nb200_cg_solve: norm_ord must be 1, 2 or 0 (inf)");
  const long maxiter_fallback = 20 * L;
  long miniter = o.miniter >= 0 ? o.miniter : std::min<long>(6, o.maxiter >= 0 ? o.maxiter : maxiter_fallback);
  long maxiter = o.maxiter >= 0 ? o.maxiter : std::max<long>(std::min<long>(200, maxiter_fallback), miniter);
  for (int k = 0; k < o.n_frozen; ++k) {
    if (!o.frozen || o.frozen[2 * k] < 0 || o.frozen[2 * k + 1] > L || o.frozen[2 * k] > o.frozen[2 * k + 1])
      return fail("nb200_cg_solve: invalid frozen range");
  }
  // frozen entries: the operator output is cleared there after every application, so that (with j and x0 zero
  // on them) every Krylov vector stays in the liquid subspace -- the solve of the restricted operator
  auto clear_frozen = [&](T* v) {
    for (int k = 0; k < o.n_frozen; ++k)
      if (o.frozen[2 * k + 1] > o.frozen[2 * k]) dev_zero(v + o.frozen[2 * k], (size_t)(o.frozen[2 * k + 1] - o.frozen[2 * k]) * sizeof(T), st);
  };
  const bool own_dot = lin_b != nullptr || multi != nullptr;     // curvature <d, q> from a separate reduction
  auto op = [&](const T* t, T* out) {
    if (multi) { multi_metric<T>(multi, n_multi, scale, identity_here, st, t, out, hook, user); clear_frozen(out); }
    else if (!lin_b) { lin->metric(st, lin, t, out, true); clear_frozen(out); }
    else {
      lin_b->metric(st, lin, t, w.tm.p, true); clear_frozen(w.tm.p);
      lin->metric(st, lin_b, w.tm.p, out, true); clear_frozen(out);
    }
  };
  CgParams<T> p;
  std::memset(&p, 0, sizeof(p));
  p.n = L; p.pos = x; p.r = w.r.p; p.d = w.d.p; p.q = w.q.p; p.j = j; p.cgs = w.cgs.p; p.cgi = w.cgi.p;
  p.partials = w.partials.p; p.counter = w.counter.p;
  p.absdelta = o.absdelta >= 0 ? (T)o.absdelta : T(-1);
  p.resnorm = o.resnorm >= 0 ? (T)o.resnorm : T(-1);
  p.eps = T(6) * std::numeric_limits<T>::epsilon(); p.tiny = T(6) * std::numeric_limits<T>::min();
  p.norm_ord = norm_ord; p.miniter = (int)miniter; p.raise_nonposdef = o.raise_nonposdef;
  p.curv_ptr = own_dot ? w.cgs.p + CG_CURV : lin->scal.p + SC_DOT;
  dev_zero(w.cgi.p, CGI_NINT * sizeof(int), st);
  dev_zero(w.cgs.p, CG_NSCAL * sizeof(T), st);
  if (o.absdelta < 0 && o.resnorm < 0) {
    // resnorm = max(tol * norm(j), atol)  (conjugate_gradient.py:103-105); one-off host read
    if (norm_ord != 2) return fail("nb200_cg_solve: default resnorm needs norm_ord=2 (pass resnorm explicitly)");
    DotParams<T> dp; dp.n = L; dp.x = j; dp.y = j; dp.partials = w.partials.p; dp.counter = w.counter.p + 1; dp.out = w.cgs.p + CG_NORM;
    launch<DotBody<T>>(vgrid, 256, 512, st, dp);
    T jj = 0; d2h(&jj, w.cgs.p + CG_NORM, sizeof(T), st); stream_sync(st);
    p.resnorm = (T)std::max((double)o.tol * std::sqrt((double)jj), (double)o.atol);
  }
  int nfev = 0;
  if (o.x0_is_zero) { p.mode = CGM_INIT_ZERO; p.iter = 0; launch<CgStepBody<T>>(vgrid, 256, 1024, st, p); }
  else { op(x, w.q.p); ++nfev; p.mode = CGM_INIT; p.iter = 0; launch<CgStepBody<T>>(vgrid, 256, 1024, st, p); }
  int host_cgi[CGI_NINT] = {0};
  const int every = std::max(1, o.check_every);
  long i = 0;
  d2h(host_cgi, w.cgi.p, sizeof(host_cgi), st); stream_sync(st);
  if (host_cgi[CGI_STATUS] == CGS_RUNNING) {
    for (i = 1; i <= maxiter; ++i) {
      op(w.d.p, w.q.p);
      if (own_dot) {
        DotParams<T> dp; dp.n = L; dp.x = w.d.p; dp.y = w.q.p; dp.partials = w.partials.p; dp.counter = w.counter.p + 1; dp.out = w.cgs.p + CG_CURV;
        launch<DotBody<T>>(vgrid, 256, 512, st, dp);
      }
      p.iter = (int)i;
      if (i % 20 == 0) {   // N_RESET (conjugate_gradient.py:17,168-172)
        p.mode = CGM_POS_ONLY; launch<CgStepBody<T>>(vgrid, 256, 1024, st, p);
        op(x, w.q.p);
        p.mode = CGM_RESID; launch<CgStepBody<T>>(vgrid, 256, 1024, st, p);
      } else {
        p.mode = CGM_NORMAL; launch<CgStepBody<T>>(vgrid, 256, 1024, st, p);
      }
      launch<CgDirBody<T>>(vgrid, 256, 0, st, p);
      if (i % every == 0 || i == maxiter) {
        d2h(host_cgi, w.cgi.p, sizeof(host_cgi), st); stream_sync(st);
        if (host_cgi[CGI_STATUS] != CGS_RUNNING) break;
      }
    }
  }
  T hs[CG_NSCAL];
  d2h(hs, w.cgs.p, sizeof(hs), st); stream_sync(st);
  int nit, info, err = 0;
  if (host_cgi[CGI_STATUS] != CGS_RUNNING) { nit = host_cgi[CGI_ITER]; info = host_cgi[CGI_INFO]; err = host_cgi[CGI_ERROR]; }
  else { nit = (int)maxiter; info = (int)maxiter; }
  // POS_ONLY + bad curvature on a reset iteration leaves the status to the following RESID launch; covered above
  res->info = info; res->nit = nit; res->nfev = nfev + nit + nit / 20; res->error = err;
  res->energy = (double)hs[CG_ENERGY]; res->gamma = (double)hs[CG_GAMMA];
  return 0;
}

}  // namespace nb

// ---- dispatch helpers ------------------------------------------------------------------------------
#define NB_TRY try {
#define NB_CATCH                                             \
  }                                                          \
  catch (const nb::Error& e) { return nb::fail(e.msg); }     \
  catch (const std::exception& e) { return nb::fail(e.what()); } \
  catch (...) { return nb::fail("nb200: unknown error"); }

template <class F64, class F32> static int by_dtype(int dtype, F64 f64, F32 f32) { return dtype == 1 ? f64() : f32(); }
#define NB_DISPATCH(dtype, TT, ...)                                 \
  if ((dtype) == 1) { typedef double TT; __VA_ARGS__ } else { typedef float TT; __VA_ARGS__ }

extern "C" {

const char* nb200_last_error(void) { return nb::last_error_ref().c_str(); }
int nb200_version(void) { return 100; }

int nb200_plan_create(nb200_plan** plan, int device, int ndim, const int64_t* shape_host, const double* distances_host,
                      int dtype, int hartley_convention) {
  NB_TRY
  if (!plan) return fail("nb200_plan_create: null output");
  if (dtype != 0 && dtype != 1) return fail("nb200_plan_create: dtype must be 0 (float32) or 1 (float64)");
  if (hartley_convention != 0 && hartley_convention != 1) return fail("nb200_plan_create: invalid hartley convention");
#ifndef NB_EMU
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device)
    return fail("nb200_plan_create: no CUDA device available (there is no CPU fallback)");
#endif
  std::unique_ptr<nb200_plan> h(new nb200_plan{nullptr});
  NB_DISPATCH(dtype, TT, { auto* p = new Plan<TT>(); h->impl = p; p->init(device, ndim, shape_host, distances_host, hartley_convention); })
  *plan = h.release();
  return 0;
  NB_CATCH
}
void nb200_plan_destroy(nb200_plan* plan) { if (plan) { delete plan->impl; delete plan; } }
int64_t nb200_plan_num_modes(const nb200_plan* plan) { return plan->impl->g.K; }
int64_t nb200_plan_size(const nb200_plan* plan) { return plan->impl->g.N; }
double nb200_plan_total_volume(const nb200_plan* plan) { return plan->impl->g.V; }
int nb200_plan_mode_lengths(const nb200_plan* plan, double* out) { const auto& g = plan->impl->g; std::copy(g.um.begin(), g.um.end(), out); return 0; }
int nb200_plan_mode_multiplicity(const nb200_plan* plan, int64_t* out) { const auto& g = plan->impl->g; std::copy(g.mult.begin(), g.mult.end(), out); return 0; }
int nb200_plan_relative_log_mode_lengths(const nb200_plan* plan, double* out) { const auto& g = plan->impl->g; std::copy(g.rel.begin(), g.rel.end(), out); return 0; }
int nb200_plan_log_volume(const nb200_plan* plan, double* out) { const auto& g = plan->impl->g; std::copy(g.logvol.begin(), g.logvol.end(), out); return 0; }
int nb200_plan_power_distributor(const nb200_plan* plan, int32_t* out) {
  const auto& g = plan->impl->g;
  int64_t q = 0;
  for (int i0 = 0; i0 < g.n0; ++i0)
    for (int im = 0; im < g.nm; ++im)
      for (int il = 0; il < g.nl; ++il, ++q) {
        int f0 = std::min(i0, g.n0 - i0), fm = std::min(im, g.nm - im), fl = std::min(il, g.nl - il);
        out[q] = g.idxf[((size_t)f0 * (g.hm + 1) + fm) * (g.hl + 1) + fl];
      }
  return 0;
}

int nb200_plan_create_dist(nb200_plan** plan, int device, int ndim, const int64_t* shape_host, const double* distances_host,
                           int dtype, int hartley_convention, int rank, int world) {
  NB_TRY
  if (!plan) return fail("nb200_plan_create_dist: null output");
  if (dtype != 0 && dtype != 1) return fail("nb200_plan_create_dist: dtype must be 0 (float32) or 1 (float64)");
  if (ndim != 3) return fail("nb200_plan_create_dist: slab decomposition needs a 3-D grid");
  if (world < 1 || rank < 0 || rank >= world) return fail("nb200_plan_create_dist: invalid rank / world");
#ifndef NB_EMU
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device)
    return fail("nb200_plan_create_dist: no CUDA device available (there is no CPU fallback)");
#endif
  std::unique_ptr<nb200_plan> h(new nb200_plan{nullptr});
  NB_DISPATCH(dtype, TT, { auto* p = new Plan<TT>(); h->impl = p; p->init(device, ndim, shape_host, distances_host, hartley_convention, rank, world); })
  *plan = h.release();
  return 0;
  NB_CATCH
}
int nb200_plan_dist_info(const nb200_plan* plan, int64_t* out, int64_t nout) {
  const PlanBase& p = *plan->impl;
  std::vector<int64_t> v = {p.rank, p.world, p.rows0, p.planes2, p.scratch_elems, p.g.n0, p.g.nm, p.g.nl, p.g.K};
  for (int r = 0; r < p.world; ++r) v.push_back(p.dist ? p.d0.npad[r] : p.g.n0);
  for (int r = 0; r < p.world; ++r) v.push_back(p.dist ? p.d2.npad[r] : p.g.nl);
  for (int r = 0; r < p.world; ++r) v.push_back(p.dist ? p.d0.cA[r] : p.g.h0 + 1);
  for (int r = 0; r < p.world; ++r) v.push_back(p.dist ? p.d2.cA[r] : p.g.hl + 1);
  if ((int64_t)v.size() > nout) return fail("nb200_plan_dist_info: output too small");
  std::copy(v.begin(), v.end(), out);
  return 0;
}
int nb200_plan_local_map(const nb200_plan* plan, int axis, int32_t* out_host) {
  const PlanBase& p = *plan->impl;
  if (!p.dist) return fail("nb200_plan_local_map: not a slab-decomposed plan");
  std::vector<int> m = (axis == 0) ? p.d0.loc2glob(p.rank) : p.d2.loc2glob(p.rank);
  std::copy(m.begin(), m.end(), out_host);
  return 0;
}
int nb200_plan_set_scratch(nb200_plan* plan, void* s0, void* s1, void* s2) {
  NB_TRY
  NB_DISPATCH(plan->impl->dtype, TT, { auto* P = static_cast<Plan<TT>*>(plan->impl); P->xS0 = (cplx<TT>*)s0; P->xS1 = (cplx<TT>*)s1; P->xS2 = (cplx<TT>*)s2; })
  return 0;
  NB_CATCH
}
int nb200_plan_set_reduce_chunks(nb200_plan* plan, int nchunks) {
  NB_TRY
  if (!plan || nchunks < 1 || nchunks > 64) return fail("nb200_plan_set_reduce_chunks: 1 <= nchunks <= 64");
  plan->impl->reduce_chunks = nchunks;
  return 0;
  NB_CATCH
}
int nb200_plan_set_chunks(nb200_plan* plan, int nchunks) {
  NB_TRY
  NB_DISPATCH(plan->impl->dtype, TT, { static_cast<Plan<TT>*>(plan->impl)->set_chunks(nchunks); })
  return 0;
  NB_CATCH
}
int nb200_dist_phase(nb200_lin* lin_a, nb200_lin* lin_b, void* stream, int code, int chunk, const void* in, void* out, void* abar,
                     void* xs, int flag) {
  NB_TRY
  if (!lin_a) return fail("nb200_dist_phase: null linearisation");
  if (lin_b && lin_b->model != lin_a->model) return fail("nb200_dist_phase: linearisations of different models");
  NB_DISPATCH(lin_a->dtype, TT, {
    Lin<TT>* a = static_cast<Lin<TT>*>(lin_a->impl);
    Lin<TT>* b = lin_b ? static_cast<Lin<TT>*>(lin_b->impl) : a;
    if (code == 2 || code == 20) b->dist_phase((stream_t)stream, code, b, (const TT*)in, (TT*)out, (TT*)abar, (TT*)xs, flag, chunk);   // tangent side
    else a->dist_phase((stream_t)stream, code, b, (const TT*)in, (TT*)out, (TT*)abar, (TT*)xs, flag, chunk);
  })
  return 0;
  NB_CATCH
}

int nb200_hartley(nb200_plan* plan, void* stream, const void* in, void* out) {
  NB_TRY
  NB_DISPATCH(plan->impl->dtype, TT, { static_cast<Plan<TT>*>(plan->impl)->hartley((stream_t)stream, (const TT*)in, (TT*)out); })
  return 0;
  NB_CATCH
}
int nb200_hartley_chirpz(nb200_plan* padded_plan, void* stream, const int64_t* n, const void* tab, const void* in, void* work, void* out) {
  NB_TRY
  if (!padded_plan || !n || !tab || !in || !work || !out) return fail("nb200_hartley_chirpz: null argument");
  NB_DISPATCH(padded_plan->impl->dtype, TT, {
    Plan<TT>& P = *static_cast<Plan<TT>*>(padded_plan->impl);
    if (P.dist) return fail("nb200_hartley_chirpz: not available on slab-decomposed plans");
    stream_t st = (stream_t)stream;
    ChirpParams<TT> p; std::memset(&p, 0, sizeof(p));
    const int nd = P.g.ndim, lead = 3 - nd;
    p.ntot = 1; p.mtot = 1;
    for (int a = 0; a < 3; ++a) {
      p.n[a] = a < lead ? 1 : n[a - lead];
      p.m[a] = a < lead ? 1 : P.g.shape[a - lead];
      if (p.n[a] < 1 || (p.n[a] > 1 && p.m[a] < 2 * p.n[a] - 1)) return fail("nb200_hartley_chirpz: the padded extent must be >= 2 n - 1 along every axis");
      p.lg_m[a] = ilog2(p.m[a]);
      p.ntot *= p.n[a]; p.mtot *= p.m[a];
    }
    const cplx<TT>* t = (const cplx<TT>*)tab;
    for (int a = 0; a < 3; ++a) { p.cc[a] = t; t += p.n[a]; }
    for (int a = 0; a < 3; ++a) { p.G[a] = t; t += p.m[a]; }
    p.x = (const TT*)in; p.w = (TT*)work; p.out = (TT*)out; p.s = P.hsign;
    auto grid = [](long cnt) { return (int)std::min<long>((cnt + 1023) / 1024 + 1, 148 * 8); };
    launch<ChirpPadBody<TT>>(grid(p.mtot), 256, 0, st, p);
    P.hartley(st, p.w, p.w); P.hartley(st, p.w + p.mtot, p.w + p.mtot);
    launch<ChirpFilterBody<TT>>(grid(p.mtot / 2), 256, 0, st, p);
    P.hartley(st, p.w, p.w); P.hartley(st, p.w + p.mtot, p.w + p.mtot);
    launch<ChirpCropBody<TT>>(grid(p.ntot), 256, 0, st, p);
  })
  return 0;
  NB_CATCH
}
int nb200_cf_apply(nb200_plan* plan, void* stream, const void* amp, const void* xi, double offset, void* out) {
  NB_TRY
  NB_DISPATCH(plan->impl->dtype, TT, { static_cast<Plan<TT>*>(plan->impl)->cf_apply((stream_t)stream, (const TT*)amp, (const TT*)xi, (TT)offset, (TT*)out); })
  return 0;
  NB_CATCH
}
int nb200_cf_apply_adjoint(nb200_plan* plan, void* stream, const void* amp, const void* xi, const void* cot, void* xi_bar, void* amp_bar) {
  NB_TRY
  if ((xi == nullptr) != (amp_bar == nullptr)) return fail("nb200_cf_apply_adjoint: xi and amp_bar must be given together");
  NB_DISPATCH(plan->impl->dtype, TT, {
    Plan<TT>& P = *static_cast<Plan<TT>*>(plan->impl);
    stream_t st = (stream_t)stream;
    PointOp<TT> op = P.make_op(PM_LOAD); op.natural = 1; op.in_pos = (const TT*)cot;
    P.template run_p3<false, true>(st, op);
    P.run_pc(st, true);
    EpiAdjoint<TT> e; e.out = (TT*)xi_bar; e.add = nullptr; e.xi = (const TT*)xi; e.idxf = P.idxf.p; e.amp = (const TT*)amp;
    e.W = xi ? P.W.p : nullptr; e.invV = TT(1.0 / P.g.V); e.partials = nullptr;
    P.run_p5(st, e);
    if (amp_bar) {
      SegSumParams<TT> ps; std::memset(&ps, 0, sizeof(ps));
      ps.m.K = P.g.K; ps.W = P.W.p; ps.order = P.w_order.p; ps.offs = P.w_offs.p; ps.abar = (TT*)amp_bar; ps.lg_lpb = P.seg_lg_lpb; ps.ellv = nullptr; ps.cv = nullptr;
      launch<SegSumBody<TT>>(P.seg_grid(), 256, (256 + 96) * sizeof(TT), st, ps);
    }
  })
  return 0;
  NB_CATCH
}

int nb200_cf_apply_batch(nb200_plan* plan, void* stream, const void* amp, int64_t amp_stride, const void* xi, double offset, void* out,
                         int64_t batch) {
  if (!plan || !amp || !xi || !out || batch < 0 || amp_stride < 0) return fail("nb200_cf_apply_batch: invalid argument");
  const int64_t N = plan->impl->g.N;
  const size_t w = plan->impl->dtype == 1 ? 8 : 4;
  for (int64_t i = 0; i < batch; ++i) {
    const int rc = nb200_cf_apply(plan, stream, (const char*)amp + (size_t)(i * amp_stride) * w, (const char*)xi + (size_t)(i * N) * w, offset,
                                  (char*)out + (size_t)(i * N) * w);
    if (rc) return rc;
  }
  return 0;
}
int nb200_cf_apply_adjoint_batch(nb200_plan* plan, void* stream, const void* amp, int64_t amp_stride, const void* xi, const void* cot,
                                 void* xi_bar, void* amp_bar, int64_t amp_bar_stride, int64_t batch) {
  if (!plan || !amp || !cot || !xi_bar || batch < 0 || amp_stride < 0) return fail("nb200_cf_apply_adjoint_batch: invalid argument");
  if ((xi == nullptr) != (amp_bar == nullptr)) return fail("nb200_cf_apply_adjoint_batch: xi and amp_bar must be given together");
  if (amp_bar && amp_bar_stride != plan->impl->g.K) return fail("nb200_cf_apply_adjoint_batch: amp_bar_stride must be K (one table of bin sums per item)");
  const int64_t N = plan->impl->g.N;
  const size_t w = plan->impl->dtype == 1 ? 8 : 4;
  for (int64_t i = 0; i < batch; ++i) {
    const int rc = nb200_cf_apply_adjoint(plan, stream, (const char*)amp + (size_t)(i * amp_stride) * w,
                                          xi ? (const char*)xi + (size_t)(i * N) * w : nullptr, (const char*)cot + (size_t)(i * N) * w,
                                          (char*)xi_bar + (size_t)(i * N) * w, amp_bar ? (char*)amp_bar + (size_t)(i * amp_bar_stride) * w : nullptr);
    if (rc) return rc;
  }
  return 0;
}

int nb200_model_create(nb200_model** model, nb200_plan* plan, const nb200_model_desc* d) {
  NB_TRY
  if (!model || !plan || !d) return fail("nb200_model_create: null argument");
  std::unique_ptr<nb200_model> h(new nb200_model{nullptr, plan, plan->impl->dtype});
  NB_DISPATCH(h->dtype, TT, { auto* m = new Model<TT>(); h->impl = m; m->init(static_cast<Plan<TT>*>(plan->impl), *d); })
  *model = h.release();
  return 0;
  NB_CATCH
}
void nb200_model_destroy(nb200_model* model) { if (model) { delete model->impl; delete model; } }

int nb200_model_set_likelihood(nb200_model* model, void* stream, int kind, int nonlinearity, const void* data,
                               double noise_cov_inv_scalar, const void* noise_cov_inv_array) {
  NB_TRY
  if (kind != 0 && kind != 1) return fail("nb200_model_set_likelihood: kind must be 0 (Gaussian) or 1 (Poissonian)");
  if (nonlinearity < 0 || nonlinearity > 2) return fail("nb200_model_set_likelihood: nonlinearity must be 0 (identity), 1 (exp) or 2 (tabulated)");
  NB_DISPATCH(model->dtype, TT, {
    Model<TT>& m = *static_cast<Model<TT>*>(model->impl);
    if (nonlinearity != 1 && m.am.has_scaling) return fail("nb200: scaling requires the exp non-linearity");
    stream_t st = (stream_t)stream;
    m.lh_kind = kind; m.nl_exp = nonlinearity; m.w_scalar = (TT)noise_cov_inv_scalar;
    const size_t npos = (size_t)m.P->local_position_grid();
    m.data.alloc(npos);
    m.has_w_arr = noise_cov_inv_array != nullptr;
    if (m.has_w_arr) m.w_arr.alloc(npos);
    if (m.P->dist) {   // slab-decomposed plans take the LOCAL planes in the internal reversed-axis layout
      if (data) d2d(m.data.p, data, npos * sizeof(TT), st);
      if (m.has_w_arr) d2d(m.w_arr.p, noise_cov_inv_array, npos * sizeof(TT), st);
    } else {
      if (data) m.P->run_rev(st, (const TT*)data, m.data.p, true);
      if (m.has_w_arr) m.P->run_rev(st, (const TT*)noise_cov_inv_array, m.w_arr.p, true);
    }
    stream_sync(st);
    m.have_lh = true;
  })
  return 0;
  NB_CATCH
}

int nb200_lin_create(nb200_lin** lin, nb200_model* model) {
  NB_TRY
  std::unique_ptr<nb200_lin> h(new nb200_lin{nullptr, model, model->dtype});
  NB_DISPATCH(h->dtype, TT, { auto* l = new Lin<TT>(); h->impl = l; l->init(static_cast<Model<TT>*>(model->impl)); })
  *lin = h.release();
  return 0;
  NB_CATCH
}
void nb200_lin_destroy(nb200_lin* lin) { if (lin) { delete lin->impl; delete lin; } }

int nb200_lin_update(nb200_lin* lin, void* stream, const void* pos, void* grad, int add_prior) {
  NB_TRY
  NB_DISPATCH(lin->dtype, TT, { static_cast<Lin<TT>*>(lin->impl)->update((stream_t)stream, (const TT*)pos, (TT*)grad, add_prior != 0); })
  return 0;
  NB_CATCH
}
int nb200_lin_set_pointwise(nb200_lin* lin, void* stream, const void* signal, const void* dsignal) {
  NB_TRY
  if (!lin || !signal || !dsignal) return fail("nb200_lin_set_pointwise: null argument");
  NB_DISPATCH(lin->dtype, TT, { static_cast<Lin<TT>*>(lin->impl)->set_pointwise((stream_t)stream, (const TT*)signal, (const TT*)dsignal); })
  return 0;
  NB_CATCH
}
int nb200_lin_energy(nb200_lin* lin, void* stream, double* energy_host) {
  NB_TRY
  NB_DISPATCH(lin->dtype, TT, {
    Lin<TT>* l = static_cast<Lin<TT>*>(lin->impl);
    TT e = 0; d2h(&e, l->scal.p + SC_ENERGY, sizeof(TT), (stream_t)stream); stream_sync((stream_t)stream);
    *energy_host = (double)e;
  })
  return 0;
  NB_CATCH
}
int nb200_lin_amplitude(nb200_lin* lin, void* stream, void* amp_out) {
  NB_TRY
  NB_DISPATCH(lin->dtype, TT, {
    Lin<TT>* l = static_cast<Lin<TT>*>(lin->impl);
    d2d(amp_out, l->amp.p, (size_t)l->M->am.K * sizeof(TT), (stream_t)stream);
  })
  return 0;
  NB_CATCH
}
int nb200_lin_signal(nb200_lin* lin, void* stream, void* out) {
  NB_TRY
  NB_DISPATCH(lin->dtype, TT, { static_cast<Lin<TT>*>(lin->impl)->posmap((stream_t)stream, PMAP_COPY, (TT*)out); })
  return 0;
  NB_CATCH
}
int nb200_cf_forward(nb200_model* model, void* stream, const void* pos, void* field_out) {
  NB_TRY
  NB_DISPATCH(model->dtype, TT, {
    Model<TT>& m = *static_cast<Model<TT>*>(model->impl);
    Lin<TT>& l = *m.scratch_lin;
    stream_t st = (stream_t)stream;
    d2d(l.pos.p, pos, (size_t)m.am.L * sizeof(TT), st);
    l.amp_forward(st);
    m.P->cf_apply(st, l.amp.p, l.pos.p + m.am.off_xi, m.offset_mean, (TT*)field_out);
  })
  return 0;
  NB_CATCH
}
int nb200_cf_amplitude(nb200_model* model, void* stream, const void* pos, void* amp_out) {
  NB_TRY
  NB_DISPATCH(model->dtype, TT, {
    Model<TT>& m = *static_cast<Model<TT>*>(model->impl);
    Lin<TT>& l = *m.scratch_lin;
    stream_t st = (stream_t)stream;
    d2d(l.pos.p, pos, (size_t)m.am.L * sizeof(TT), st);
    l.amp_forward(st);
    d2d(amp_out, l.amp.p, (size_t)m.am.K * sizeof(TT), st);
  })
  return 0;
  NB_CATCH
}
int nb200_metric(nb200_lin* lin, void* stream, const void* t, void* out, int add_identity) {
  NB_TRY
  NB_DISPATCH(lin->dtype, TT, { Lin<TT>* l = static_cast<Lin<TT>*>(lin->impl); l->metric((stream_t)stream, l, (const TT*)t, (TT*)out, add_identity != 0); })
  return 0;
  NB_CATCH
}
int nb200_metric_pair(nb200_lin* lin_a, nb200_lin* lin_b, void* stream, const void* t, void* out, int add_identity) {
  NB_TRY
  if (lin_a->model != lin_b->model) return fail("nb200_metric_pair: linearisations of different models");
  NB_DISPATCH(lin_a->dtype, TT, {
    static_cast<Lin<TT>*>(lin_a->impl)->metric((stream_t)stream, static_cast<Lin<TT>*>(lin_b->impl), (const TT*)t, (TT*)out, add_identity != 0);
  })
  return 0;
  NB_CATCH
}
int nb200_rsm(nb200_lin* lin, void* stream, const void* t, void* out_pos, int scaled) {
  NB_TRY
  NB_DISPATCH(lin->dtype, TT, { static_cast<Lin<TT>*>(lin->impl)->rsm((stream_t)stream, (const TT*)t, (TT*)out_pos, scaled != 0); })
  return 0;
  NB_CATCH
}
int nb200_lsm(nb200_lin* lin, void* stream, const void* u_pos, void* out, int scaled) {
  NB_TRY
  NB_DISPATCH(lin->dtype, TT, { static_cast<Lin<TT>*>(lin->impl)->lsm((stream_t)stream, (const TT*)u_pos, (TT*)out, scaled != 0); })
  return 0;
  NB_CATCH
}
int nb200_transformation(nb200_lin* lin, void* stream, void* out_pos) {
  NB_TRY
  NB_DISPATCH(lin->dtype, TT, { static_cast<Lin<TT>*>(lin->impl)->posmap((stream_t)stream, PMAP_TRAFO, (TT*)out_pos); })
  return 0;
  NB_CATCH
}
int nb200_normalized_residual(nb200_lin* lin, void* stream, void* out_pos) {
  NB_TRY
  NB_DISPATCH(lin->dtype, TT, { static_cast<Lin<TT>*>(lin->impl)->posmap((stream_t)stream, PMAP_NRES, (TT*)out_pos); })
  return 0;
  NB_CATCH
}

void nb200_cg_default_opts(nb200_cg_opts* o) {
  o->absdelta = -1; o->resnorm = -1; o->tol = 1e-5; o->atol = 0; o->norm_ord = 2; o->miniter = -1; o->maxiter = -1;
  o->raise_nonposdef = 1; o->check_every = 4; o->x0_is_zero = 0; o->n_frozen = 0; o->reserved = 0; o->frozen = nullptr;
}
int nb200_cg_solve(nb200_lin* lin, nb200_lin* lin_b, void* stream, const void* j, void* x, const nb200_cg_opts* opts,
                   nb200_cg_result* result_host) {
  NB_TRY
  if (!lin || !j || !x || !opts || !result_host) return fail("nb200_cg_solve: null argument");
  if (lin_b && lin_b->model != lin->model) return fail("nb200_cg_solve: linearisations of different models");
  NB_DISPATCH(lin->dtype, TT, {
    return cg_solve<TT>(static_cast<Lin<TT>*>(lin->impl), lin_b ? static_cast<Lin<TT>*>(lin_b->impl) : nullptr, (stream_t)stream,
                        (const TT*)j, (TT*)x, *opts, result_host);
  })
  NB_CATCH
}

int nb200_metric_multi(nb200_lin** lins, int n_lins, double scale, int identity_here, void* stream, const void* t, void* out,
                       nb200_reduce_hook hook, void* user) {
  NB_TRY
  if (!lins || n_lins < 1 || !t || !out) return fail("nb200_metric_multi: invalid argument");
  for (int i = 1; i < n_lins; ++i) if (lins[i]->model != lins[0]->model) return fail("nb200_metric_multi: linearisations of different models");
  NB_DISPATCH(lins[0]->dtype, TT, {
    std::vector<Lin<TT>*> v(n_lins);
    for (int i = 0; i < n_lins; ++i) v[i] = static_cast<Lin<TT>*>(lins[i]->impl);
    multi_metric<TT>(v.data(), n_lins, (TT)scale, identity_here != 0, (stream_t)stream, (const TT*)t, (TT*)out, hook, user);
  })
  return 0;
  NB_CATCH
}
int nb200_cg_solve_multi(nb200_lin** lins, int n_lins, double scale, int identity_here, void* stream, const void* j, void* x,
                         const nb200_cg_opts* opts, nb200_cg_result* result_host, nb200_reduce_hook hook, void* user) {
  NB_TRY
  if (!lins || n_lins < 1 || !j || !x || !opts || !result_host) return fail("nb200_cg_solve_multi: invalid argument");
  for (int i = 1; i < n_lins; ++i) if (lins[i]->model != lins[0]->model) return fail("nb200_cg_solve_multi: linearisations of different models");
  NB_DISPATCH(lins[0]->dtype, TT, {
    std::vector<Lin<TT>*> v(n_lins);
    for (int i = 0; i < n_lins; ++i) v[i] = static_cast<Lin<TT>*>(lins[i]->impl);
    return cg_solve<TT>(v[0], nullptr, (stream_t)stream, (const TT*)j, (TT*)x, *opts, result_host, v.data(), n_lins, (TT)scale,
                        identity_here != 0, hook, user);
  })
  NB_CATCH
}

int nb200_vec_axpby(nb200_plan* plan, void* stream, int64_t n, double a, const void* x, double b, const void* y, void* out) {
  NB_TRY
  NB_DISPATCH(plan->impl->dtype, TT, {
    AxpbyParams<TT> p; p.n = n; p.a = (TT)a; p.b = (TT)b; p.x = (const TT*)x; p.y = (const TT*)y; p.out = (TT*)out;
    launch<AxpbyBody<TT>>((int)std::min<int64_t>((n + 1023) / 1024 + 1, 148 * 8), 256, 0, (stream_t)stream, p);
  })
  return 0;
  NB_CATCH
}
int nb200_vec_dot(nb200_plan* plan, void* stream, int64_t n, const void* x, const void* y, double* out_host) {
  NB_TRY
  NB_DISPATCH(plan->impl->dtype, TT, {
    CgWork<TT>& w = cg_work<TT>();
    if (w.partials.n == 0) { w.partials.alloc(4 * 2048); w.counter.alloc(4); w.cgs.alloc(CG_NSCAL); w.cgi.alloc(CGI_NINT); }
    DotParams<TT> dp; dp.n = n; dp.x = (const TT*)x; dp.y = (const TT*)y; dp.partials = w.partials.p; dp.counter = w.counter.p + 1; dp.out = w.cgs.p + CG_NORM;
    launch<DotBody<TT>>((int)std::min<int64_t>((n + 1023) / 1024 + 1, 148 * 8), 256, 512, (stream_t)stream, dp);
    TT v = 0; d2h(&v, w.cgs.p + CG_NORM, sizeof(TT), (stream_t)stream); stream_sync((stream_t)stream);
    *out_host = (double)v;
  })
  return 0;
  NB_CATCH
}

int nb200_vec_stats(nb200_plan* plan, void* stream, int64_t n, const void* x, double* out_host) {
  NB_TRY
  NB_DISPATCH(plan->impl->dtype, TT, {
    CgWork<TT>& w = cg_work<TT>();
    if (w.partials.n == 0) { w.partials.alloc(4 * 2048); w.counter.alloc(4); w.cgs.alloc(CG_NSCAL); w.cgi.alloc(CGI_NINT); }
    StatsParams<TT> sp; sp.n = n; sp.x = (const TT*)x; sp.partials = w.partials.p; sp.counter = w.counter.p + 1; sp.out = w.cgs.p + CG_NORM;
    launch<StatsBody<TT>>((int)std::min<int64_t>((n + 1023) / 1024 + 1, 148 * 8), 256, 1024, (stream_t)stream, sp);
    TT v[2] = {0, 0}; d2h(v, w.cgs.p + CG_NORM, 2 * sizeof(TT), (stream_t)stream); stream_sync((stream_t)stream);
    out_host[0] = (double)v[0]; out_host[1] = (double)v[1];
  })
  return 0;
  NB_CATCH
}

// Per-kernel event timing: begin() arms it, end() synchronises and writes "name count total_ms" lines.
int nb200_timing_begin(void) {
#ifndef NB_EMU
  nb::kernel_timer().recs.clear(); nb::kernel_timer().on = true;
#endif
  return 0;
}
int nb200_timing_end(char* buf, int64_t buflen) {
  NB_TRY
  std::string out;
#ifndef NB_EMU
  nb::KernelTimer& kt = nb::kernel_timer();
  kt.on = false;
  NB_CUDA_CHECK(cudaDeviceSynchronize());
  std::vector<std::string> names; std::vector<double> tot; std::vector<long> cnt;
  for (auto& r : kt.recs) {
    float ms = 0; NB_CUDA_CHECK(cudaEventElapsedTime(&ms, r.a, r.b));
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    size_t i = 0; for (; i < names.size(); ++i) if (names[i] == r.name) break;
    if (i == names.size()) { names.push_back(r.name); tot.push_back(0); cnt.push_back(0); }
    tot[i] += ms; cnt[i] += 1;
  }
  kt.recs.clear();
  for (size_t i = 0; i < names.size(); ++i) out += names[i] + " " + std::to_string(cnt[i]) + " " + std::to_string(tot[i]) + "\n";
#endif
  if (buf && buflen > 0) { size_t n = std::min<size_t>(out.size(), (size_t)buflen - 1); std::memcpy(buf, out.data(), n); buf[n] = 0; }
  return 0;
  NB_CATCH
}

unsigned long long nb200_launch_count(void) {
#ifdef NB_EMU
  return 0;
#else
  return nb::launch_counter();
#endif
}

}  // extern "C"
