// nifty_b200 -- fused vector kernels of conjugate gradient on flat latent vectors.
//
// One iteration of `_cg` (nifty/re/conjugate_gradient.py:139-204) touches the vectors twice:
//   CgStep : alpha = gamma_prev / <d,q>;  pos -= alpha d;  r -= alpha q;  and in the same sweep the
//            partial sums of gamma = <r,r>, |r|_1, |r|_inf and energy = <(r-j)/2, pos>; the last
//            block finishes the sums in fixed order and evaluates every stopping rule ON THE DEVICE
//            (tiny-gamma, resnorm, energy increase, absdelta) -- no host round trip per scalar
//   CgDir  : d = max(0, gamma/gamma_prev) d + r
// <d,q> itself comes for free from the epilogue of the metric kernels (SC_DOT).  All kernels are
// no-ops once the status word leaves 0, so the host may enqueue iterations ahead of its polls.
#pragma once
#include "nb_common.cuh"
#include "nb_amp.cuh"

namespace nb {

enum CgScal { CG_GAMMA_PREV = 0, CG_GAMMA, CG_ENERGY, CG_NORM, CG_ALPHA, CG_BETA, CG_CURV, CG_NSCAL = 8 };
enum CgInt { CGI_STATUS = 0, CGI_ITER, CGI_INFO, CGI_ERROR, CGI_NINT = 8 };
enum CgStatus { CGS_RUNNING = 0, CGS_DONE = 1 };
enum CgStepMode { CGM_NORMAL = 0, CGM_POS_ONLY = 1, CGM_RESID = 2, CGM_INIT = 3, CGM_INIT_ZERO = 4 };

template <class T> struct CgParams {
  int mode, iter;
  long n;
  T* pos; T* r; T* d; const T* q; const T* j;
  T* cgs; int* cgi; const T* curv_ptr;
  T* partials;   // [nblk][4]
  unsigned* counter;
  // stopping rules
  T absdelta, resnorm, eps, tiny;   // < 0: disabled
  int norm_ord, miniter, raise_nonposdef;
};

template <class T> struct CgStepBody {
  typedef CgParams<T> Params;
  static NB_HD void run(Ctx& ctx, const Params& p, void* smem) {
    if (p.cgi[CGI_STATUS] != CGS_RUNNING) return;
    const bool init = (p.mode == CGM_INIT || p.mode == CGM_INIT_ZERO);
    T curv = init ? T(1) : *p.curv_ptr;
    T gprev = p.cgs[CG_GAMMA_PREV];
    bool bad = !init && p.mode != CGM_RESID && !(curv > T(0));
    T alpha = bad ? T(0) : gprev / curv;
    T a_g = 0, a_1 = 0, a_m = 0, a_e = 0;
    const long stride = (long)ctx.nblk * ctx.nthr;
    if (bad) {
      // zero / negative curvature (conjugate_gradient.py:149-165)
      if (curv < T(0) && !p.raise_nonposdef && p.iter == 1) {
        T f = gprev / (-curv);
        for (long i = (long)ctx.bid * ctx.nthr + ctx.tid; i < p.n; i += stride) p.pos[i] = -f * p.j[i];
      }
    } else {
      for (long i = (long)ctx.bid * ctx.nthr + ctx.tid; i < p.n; i += stride) {
        T x, rr, jj = p.j[i];
        if (p.mode == CGM_NORMAL) {
          T dd = p.d[i];
          x = p.pos[i] - alpha * dd; p.pos[i] = x;
          rr = p.r[i] - alpha * p.q[i]; p.r[i] = rr;
        } else if (p.mode == CGM_POS_ONLY) {
          p.pos[i] = p.pos[i] - alpha * p.d[i];
          continue;
        } else if (p.mode == CGM_RESID || p.mode == CGM_INIT) {
          x = p.pos[i]; rr = p.q[i] - jj; p.r[i] = rr;
          if (init) p.d[i] = rr;
        } else {   // CGM_INIT_ZERO: pos = 0, r = d = -j
          x = 0; p.pos[i] = 0; rr = -jj; p.r[i] = rr; p.d[i] = rr;
        }
        a_g += rr * rr;
        T ar = rr < 0 ? -rr : rr;
        a_1 += ar; a_m = ar > a_m ? ar : a_m;
        a_e += T(0.5) * (rr - jj) * x;
      }
    }
    if (p.mode == CGM_POS_ONLY && !bad) {
      if (ctx.bid == 0 && ctx.tid == 0) p.cgs[CG_ALPHA] = alpha;
      return;
    }
    { T v3[3] = {a_g, a_1, a_e}; ctx.template block_sum_n<3>(v3, smem); a_g = v3[0]; a_1 = v3[1]; a_e = v3[2]; }
    a_m = ctx.block_max(a_m, smem);
    if (ctx.tid == 0) {
      p.partials[4 * ctx.bid] = a_g; p.partials[4 * ctx.bid + 1] = a_1;
      p.partials[4 * ctx.bid + 2] = a_m; p.partials[4 * ctx.bid + 3] = a_e;
    }
    if (ctx.last_block(p.counter)) {
      T gamma = block_total(ctx, p.partials, ctx.nblk, 4, smem);
      T n1 = block_total(ctx, p.partials + 1, ctx.nblk, 4, smem);
      T en = block_total(ctx, p.partials + 3, ctx.nblk, 4, smem);
      T mx = 0;
      NB_FOR(ctx, i, ctx.nblk) { T v = p.partials[4 * i + 2]; mx = v > mx ? v : mx; }
      mx = ctx.block_max(mx, smem);
      if (ctx.tid == 0) {
        int status = CGS_RUNNING, info = -1, err = 0;
        if (bad) {
          if (p.raise_nonposdef) { err = (curv == T(0)) ? 1 : 2; info = -1; }
          else info = 0;
          status = CGS_DONE;
        } else if (init) {
          p.cgs[CG_GAMMA_PREV] = gamma; p.cgs[CG_GAMMA] = gamma; p.cgs[CG_ENERGY] = en;
          if (gamma == T(0)) { status = CGS_DONE; info = 0; }
        } else {
          T nrm = p.norm_ord == 2 ? nb_sqrt(gamma) : (p.norm_ord == 1 ? n1 : mx);
          T energy = p.cgs[CG_ENERGY];
          T ediff = energy - en;
          T aen = en < 0 ? -en : en;
          p.cgs[CG_GAMMA] = gamma; p.cgs[CG_NORM] = nrm; p.cgs[CG_ALPHA] = alpha;
          if (gamma >= T(0) && gamma <= p.tiny) { status = CGS_DONE; info = 0; }
          else if (p.resnorm >= T(0) && nrm < p.resnorm && p.iter >= p.miniter) { status = CGS_DONE; info = 0; }
          else if (ediff < -p.eps * aen) {
            status = CGS_DONE;
            if (p.raise_nonposdef) { err = 3; info = -1; } else info = p.iter;
          } else if (p.absdelta >= T(0) && ediff < p.absdelta && p.iter >= p.miniter) { status = CGS_DONE; info = 0; }
          else {
            T beta = gamma / gprev;
            p.cgs[CG_BETA] = beta > T(0) ? beta : T(0);
            p.cgs[CG_GAMMA_PREV] = gamma;
          }
          p.cgs[CG_ENERGY] = en;
        }
        if (status != CGS_RUNNING) { p.cgi[CGI_ITER] = p.iter; p.cgi[CGI_INFO] = info; p.cgi[CGI_ERROR] = err; p.cgi[CGI_STATUS] = status; }
      }
    }
  }
};

template <class T> struct CgDirBody {
  typedef CgParams<T> Params;
  static NB_HD void run(Ctx& ctx, const Params& p, void*) {
    if (p.cgi[CGI_STATUS] != CGS_RUNNING) return;
    T beta = p.cgs[CG_BETA];
    const long stride = (long)ctx.nblk * ctx.nthr;
    for (long i = (long)ctx.bid * ctx.nthr + ctx.tid; i < p.n; i += stride) p.d[i] = p.d[i] * beta + p.r[i];
  }
};

// out = a x + b y (either pointer may alias out); dot product into a device scalar
template <class T> struct AxpbyParams { long n; T a, b; const T* x; const T* y; T* out; };
template <class T> struct AxpbyBody {
  typedef AxpbyParams<T> Params;
  static NB_HD void run(Ctx& ctx, const Params& p, void*) {
    const long stride = (long)ctx.nblk * ctx.nthr;
    for (long i = (long)ctx.bid * ctx.nthr + ctx.tid; i < p.n; i += stride) {
      T v = p.a * p.x[i];
      if (p.y) v += p.b * p.y[i];
      p.out[i] = v;
    }
  }
};
template <class T> struct DotParams { long n; const T* x; const T* y; T* partials; unsigned* counter; T* out; };
template <class T> struct DotBody {
  typedef DotParams<T> Params;
  static NB_HD void run(Ctx& ctx, const Params& p, void* smem) {
    T a = 0;
    const long stride = (long)ctx.nblk * ctx.nthr;
    for (long i = (long)ctx.bid * ctx.nthr + ctx.tid; i < p.n; i += stride) a += p.x[i] * p.y[i];
    a = ctx.block_sum(a, smem);
    if (ctx.tid == 0) p.partials[ctx.bid] = a;
    if (ctx.last_block(p.counter)) {
      T t = block_total(ctx, p.partials, ctx.nblk, 1, smem);
      if (ctx.tid == 0) *p.out = t;
    }
  }
};

// (sum x, sum x^2): the two moments behind reduced_residual_stats (nifty/re/minisanity.py:17-21)
template <class T> struct StatsParams { long n; const T* x; T* partials; unsigned* counter; T* out; };
template <class T> struct StatsBody {
  typedef StatsParams<T> Params;
  static NB_HD void run(Ctx& ctx, const Params& p, void* smem) {
    T v[2] = {0, 0};
    const long stride = (long)ctx.nblk * ctx.nthr;
    for (long i = (long)ctx.bid * ctx.nthr + ctx.tid; i < p.n; i += stride) { T x = p.x[i]; v[0] += x; v[1] += x * x; }
    ctx.template block_sum_n<2>(v, smem);
    if (ctx.tid == 0) { p.partials[2 * ctx.bid] = v[0]; p.partials[2 * ctx.bid + 1] = v[1]; }
    if (ctx.last_block(p.counter)) {
      T t[2] = {0, 0};
      NB_FOR(ctx, i, ctx.nblk) { t[0] += p.partials[2 * (size_t)i]; t[1] += p.partials[2 * (size_t)i + 1]; }
      ctx.template block_sum_n<2>(t, smem);
      if (ctx.tid == 0) { p.out[0] = t[0]; p.out[1] = t[1]; }
    }
  }
};

}  // namespace nb
