/* nifty_b200 -- C ABI of the B200-native MGVI/geoVI inner loop (libniftyb200.so).
 *
 * This is the drop-in boundary of the hot path of NIFTy.re: every entry point names the reference
 * interface it replaces (paths relative to the reference tree, nifty/re/...).  The reference has
 * no FFI of its own (it is pure Python on JAX); these are the symbols a `jax.ffi` custom-call
 * shim, or any other host binding (ctypes, cgo, ...), binds.  See INTEGRATION.md.
 *
 * Conventions
 *  - All array arguments are DEVICE pointers owned by the caller (XLA / torch allocator) unless
 *    the parameter name ends in `_host`.  The callee never frees or retains them beyond the
 *    call, except `nb200_model_set_likelihood` which copies what it needs.
 *  - `dtype`: 0 = float32, 1 = float64; every array of one plan has the plan's dtype.
 *  - `stream` is a `cudaStream_t` passed as `void*`; all work is enqueued on it, nothing in a
 *    hot call allocates or synchronises (calls that return host scalars say so).
 *  - Latent vectors are FLAT buffers of `latent_size` elements; leaf offsets are given in
 *    `nb200_model_desc` (the host keeps the pytree <-> flat mapping, sorted-key order in NIFTy).
 *  - Harmonic-space / latent arrays (xi, excitations, gradients) are in natural C order.
 *    Position-space arrays passed in/out through this API are also in natural C order; inside,
 *    the library keeps position-space state in reversed-axis order.
 *  - Return value 0 = success; otherwise `nb200_last_error()` (thread local) describes the failure.
 *    Nothing aborts or throws across the ABI.  There is NO CPU fallback: without a CUDA device
 *    `nb200_plan_create` fails.
 *  - A plan owns transform scratch: use one plan per concurrently used stream.
 */
#ifndef NIFTY_B200_H
#define NIFTY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nb200_plan nb200_plan;   /* grid geometry + transform plan (twiddles, mode bins, scratch) */
typedef struct nb200_model nb200_model; /* amplitude model + likelihood bound to a plan               */
typedef struct nb200_lin nb200_lin;     /* linearisation state of a model at one latent position       */

const char* nb200_last_error(void);
int nb200_version(void);
/* number of kernel launches issued by this library in this process (instrumentation) */
unsigned long long nb200_launch_count(void);
/* per-kernel CUDA-event timing (instrumentation for bench.py): begin() arms it; end() synchronises the
 * device and writes one "mangled_kernel_body_name launches total_ms" line per kernel into buf */
int nb200_timing_begin(void);
int nb200_timing_end(char* buf, int64_t buflen);

/* ---- plan: grid, mode bins, transform ------------------------------------------------------ */

/* Replaces make_grid / get_fourier_mode_distributor / _unique_mode_distributor
 * (correlated_field.py:238-294, :134-176, :55-67): builds the |k| bins of a regular Fourier
 * grid (1e-12 relative merge tolerance, RuntimeError semantics -> error code) plus the transform
 * plan.  ndim in {1,2,3}; every extent a power of two >= 2 (other shapes: explicit error).
 * hartley_convention: 0 = "non_canonical_hartley" (Re+Im, nifty/config.py:44 default), 1 = "canonical_hartley". */
int nb200_plan_create(nb200_plan** plan, int device, int ndim, const int64_t* shape_host,
                      const double* distances_host, int dtype, int hartley_convention);
void nb200_plan_destroy(nb200_plan* plan);
int64_t nb200_plan_num_modes(const nb200_plan* plan);                       /* K */
int64_t nb200_plan_size(const nb200_plan* plan);                            /* N */
double nb200_plan_total_volume(const nb200_plan* plan);
/* host copies of the static tables of RegularFourierGrid (correlated_field.py:179-200, 228-235) */
int nb200_plan_mode_lengths(const nb200_plan* plan, double* out_host /*K*/);
int nb200_plan_mode_multiplicity(const nb200_plan* plan, int64_t* out_host /*K*/);
int nb200_plan_relative_log_mode_lengths(const nb200_plan* plan, double* out_host /*K*/);
int nb200_plan_log_volume(const nb200_plan* plan, double* out_host /*K-2*/);
/* full power_distributor (int32, natural order, N entries) for API compatibility */
int nb200_plan_power_distributor(const nb200_plan* plan, int32_t* out_host /*N*/);

/* ---- slab-decomposed 3-D grids (no reference counterpart: SURVEY.md 8e.2; fields too large for one GPU) ----
 * One process per GPU.  Axis 0 of the latent array and the last axis of position space are cut into slabs such
 * that the Hermitian partner of every line is local.  The library runs the local passes;
 * the HOST performs the two all-to-all exchanges and one all-reduce per product with its own communicator
 * (torch.distributed / NCCL) on the buffers it registered with nb200_plan_set_scratch. */
int nb200_plan_create_dist(nb200_plan** plan, int device, int ndim, const int64_t* shape_host, const double* distances_host,
                           int dtype, int hartley_convention, int rank, int world);
/* out = {rank, world, local latent rows (padded), local position planes (padded), scratch elements (complex),
 *        n0, n1, n2, K, rows per rank [world], planes per rank [world], half-range rows per rank [world], half-range planes per rank [world]} */
int nb200_plan_dist_info(const nb200_plan* plan, int64_t* out_host, int64_t nout);
/* local -> global index of the owned rows (axis = 0) or position planes (axis = 2); -1 marks zero padding */
int nb200_plan_local_map(const nb200_plan* plan, int axis, int32_t* out_host);
/* three device buffers of `scratch elements` complex numbers each: pass buffers and all-to-all send / receive space */
int nb200_plan_set_scratch(nb200_plan* plan, void* s0, void* s1, void* s2);
/* cut every rank's half-range piece into `nchunks` pieces so that the host can pipeline an exchange with the
 * passes before / after it (chunk c of pass X, then exchange chunk c while pass X works on chunk c+1, ...) */
int nb200_plan_set_chunks(nb200_plan* plan, int nchunks);
/* one segment of an operator between exchanges (chunk < 0: whole range); codes are listed in DESIGN.md section 7 */
int nb200_dist_phase(nb200_lin* lin_a, nb200_lin* lin_b, void* stream, int code, int chunk, const void* in, void* out, void* abar,
                     void* xs, int flag);

/* hartley(p, axes=all) (correlated_field.py:24-30): out = Re(fftn(in)) +/- Im(fftn(in)), unnormalised */
int nb200_hartley(nb200_plan* plan, void* stream, const void* in, void* out);

/* hartley(p) (correlated_field.py:24-30) on a grid whose extents are NOT powers of two -- the reference's own parity case (3, 3)
 * (test/test_re/test_correlated_field.py:123-124) and the 3618^2 point of its published benchmark (padded extents are limited to 8192 in float64, 16384 in float32) -- as a chirp convolution
 * (Bluestein) through the power-of-two passes of `padded_plan`, whose extents must be >= 2 n - 1 along every axis.
 *   n    : the logical extents (host, ndim of the padded plan entries)
 *   tab  : device table of complex numbers (interleaved re, im, the plan's dtype): for each of the three right-aligned axes (missing
 *          leading axes count one entry equal to 1) conj(c_j) = exp(-i pi j^2 / n), j < n; then for each axis the M-point DFT of the
 *          even chirp filter g_m = c_|m| (|m| < n, zero elsewhere), the last axis scaled by 1 / prod(M)
 *   in   : n-grid, out: n-grid, work: 2 prod(M) elements of scratch
 * Sequence on the stream: pad + chirp, 2 in-place hartley, spectrum product on mirror pairs, 2 in-place hartley, crop + chirp. */
int nb200_hartley_chirpz(nb200_plan* padded_plan, void* stream, const int64_t* n, const void* tab, const void* in, void* work,
                         void* out);

/* correlated_field(p) core (correlated_field.py:909-912, :882-887) as a bilinear operator:
 *   out = offset + (1/V) hartley(amp[power_distributor] * xi)
 * `amp` is the K-entry table azm * normalized_amplitude with amp[0] = zeromode * V. */
int nb200_cf_apply(nb200_plan* plan, void* stream, const void* amp, const void* xi, double offset, void* out);
/* its transpose (what jax.linear_transpose / jax.vjp generate, likelihood.py:304,619):
 *   xi_bar = amp[pd] * g,  amp_bar[b] = sum_{k in bin b} xi_k g_k,  g = (1/V) hartley(cot).
 * xi / amp_bar may be NULL together (linear-in-xi transpose only). */
int nb200_cf_apply_adjoint(nb200_plan* plan, void* stream, const void* amp, const void* xi, const void* cot,
                           void* xi_bar, void* amp_bar);

/* Batched forms: what `jax.vmap` of the two operators above asks of the custom call (optimize_kl.py:106,135 maps over the
 * sample axis with every leaf batched; test/test_re/test_empirical_power_spectrum.py:39 `VModel(cf, in_axes="xi")` batches
 * the excitations only).  `batch` independent applications on consecutive grids: xi / out / cot / xi_bar advance by one
 * grid (N elements) per item, amp / amp_bar by `amp_stride` / `amp_bar_stride` elements (K: one table per item; 0 for
 * amp: one table shared by the batch).  amp_bar_stride must be K or amp_bar NULL. */
int nb200_cf_apply_batch(nb200_plan* plan, void* stream, const void* amp, int64_t amp_stride, const void* xi, double offset, void* out,
                         int64_t batch);
int nb200_cf_apply_adjoint_batch(nb200_plan* plan, void* stream, const void* amp, int64_t amp_stride, const void* xi, const void* cot,
                                 void* xi_bar, void* amp_bar, int64_t amp_bar_stride, int64_t batch);

/* ---- model: CorrelatedFieldMaker (one sub-grid) + likelihood --------------------------------- */

typedef struct nb200_model_desc {
  /* add_fluctuations(..., non_parametric_kind=) correlated_field.py:661-755; NonParametricAmplitude :398-516 */
  int32_t kind_power;        /* 1 "power", 0 "amplitude" */
  int32_t has_fluctuations;  /* fluctuations is not None */
  int32_t has_deviations;    /* flexibility is not None and K > 2 (spectrum leaf present) */
  int32_t has_asperity;
  int32_t has_scaling;       /* signal = scaling * nl(cf), scaling ~ LogNormalPrior, shape (1,) (demos/re/0_intro.py:39,57) */
  int32_t reserved;
  double offset_mean;        /* set_amplitude_total_offset (:583-659) */
  /* prior parametrisation value = a + b*xi (normal) or exp(a + b*xi) (lognormal); a, b as produced by
   * lognormal_moments / normal_prior (num/stats_distributions.py:42-98) */
  double zeromode_a, zeromode_b;
  double fluct_a, fluct_b;
  double slope_a, slope_b;
  double flex_a, flex_b;
  double asp_a, asp_b;
  double scaling_a, scaling_b;
  /* offsets of the leaves in the flat latent vector (elements); -1 if the leaf is absent */
  int64_t off_xi, off_zeromode, off_fluct, off_slope, off_flex, off_asp, off_spectrum, off_scaling;
  int64_t latent_size;
  /* add_fluctuations_matern / MaternAmplitude (correlated_field.py:302-395, 661-755): amplitude_type 1.
   * Leaves: <prefix>scale -> fluct_a/b + off_fluct, <prefix>loglogslope -> slope_a/b + off_slope, <prefix>cutoff (lognormal) */
  int32_t amplitude_type;         /* 0 non-parametric, 1 Matern */
  int32_t renormalize_amplitude;  /* Matern only */
  double cutoff_a, cutoff_b;
  int64_t off_cutoff;
} nb200_model_desc;

int nb200_model_create(nb200_model** model, nb200_plan* plan, const nb200_model_desc* desc_host);
void nb200_model_destroy(nb200_model* model);

/* Gaussian(data, noise_cov_inv).amend(signal) / Poissonian(data).amend(signal)
 * (likelihood_impl.py:83-138, 203-251; likelihood.py:546-633).  kind: 0 Gaussian, 1 Poissonian.
 * nonlinearity: 0 identity, 1 exp, 2 tabulated (nb200_lin_set_pointwise).  `data` natural order, plan dtype (Poisson counts converted by
 * the host).  Diagonal noise: scalar, or array if noise_cov_inv_array != NULL. */
int nb200_model_set_likelihood(nb200_model* model, void* stream, int kind, int nonlinearity, const void* data,
                               double noise_cov_inv_scalar, const void* noise_cov_inv_array);

/* ---- linearisation ------------------------------------------------------------------------------ */

int nb200_lin_create(nb200_lin** lin, nb200_model* model);
void nb200_lin_destroy(nb200_lin* lin);

/* Linearise at `pos`: amplitude tables, signal and the cached Jacobian/metric weights
 * (what jax.linearize recomputes on EVERY LikelihoodWithModel.metric call, likelihood.py:618).
 * If grad != NULL also writes d(likelihood energy)/d pos (+ pos if add_prior), i.e. the gradient of
 * _StandardHamiltonian (optimize_kl.py:67-87).  The energy is left on the device (nb200_lin_energy). */
int nb200_lin_update(nb200_lin* lin, void* stream, const void* pos, void* grad, int add_prior);
/* Models with nonlinearity = 2 ("tabulated": any pointwise map signal = f(correlated field), the `Model(lambda x: f(cf(x)))`
 * of nifty/re/model.py:146-181): before every nb200_lin_update the host evaluates f and f' at the field of that position
 * (nb200_cf_forward gives the field) and hands both over, natural order, plan dtype.  Everything after the
 * linearisation (metric, sqrt-metrics, CG, ...) is unchanged: it only sees the cached signal and Jacobian weights. */
int nb200_lin_set_pointwise(nb200_lin* lin, void* stream, const void* signal, const void* dsignal);
/* likelihood energy of the last nb200_lin_update (synchronises the stream) */
int nb200_lin_energy(nb200_lin* lin, void* stream, double* energy_host);

/* amplitude table azm*normalized_amplitude (K entries, amp[0] = zeromode*V) of the linearisation point */
int nb200_lin_amplitude(nb200_lin* lin, void* stream, void* amp_out);
/* signal(pos) and correlated field in natural order (Model.__call__, correlated_field.py:909-912) */
int nb200_lin_signal(nb200_lin* lin, void* stream, void* out);
int nb200_cf_forward(nb200_model* model, void* stream, const void* pos, void* field_out);
/* CorrelatedFieldMaker.amplitude (correlated_field.py:824-838): amp_out[K] = [azm V, amp_1, ..., amp_{K-1}] at pos,
 * i.e. the O(K) amplitude chain alone (power_spectrum = amp^2, normalized amplitudes = amp / azm; :807-845) */
int nb200_cf_amplitude(nb200_model* model, void* stream, const void* pos, void* amp_out);

/* LikelihoodWithModel.metric (likelihood.py:613-621) fused: out = J^T M J t (+ t if add_identity,
 * i.e. _ham_metric evi.py:83-85).  Leaves <t, out> in a device scalar for conjugate gradient. */
int nb200_metric(nb200_lin* lin, void* stream, const void* t, void* out, int add_identity);
/* geoVI building block (evi.py:167-172): out = LSM_a( RSM_b(t) ) (+ t) = J_a^T l_a l_b J_b t (+ t) */
int nb200_metric_pair(nb200_lin* lin_a, nb200_lin* lin_b, void* stream, const void* t, void* out, int add_identity);

/* right_sqrt_metric / forward-mode derivative (likelihood.py:306-332, 625-626):
 *   scaled=1: out = l * d signal/d pos . t   (RSM);  scaled=0: out = d field/d pos . t  (field JVP) */
int nb200_rsm(nb200_lin* lin, void* stream, const void* t, void* out_pos, int scaled);
/* left_sqrt_metric / reverse-mode (likelihood.py:284-304, 622-623):
 *   scaled=1: out = J_signal^T (l * u) (LSM); scaled=0: out = J_field^T u (field VJP) */
int nb200_lsm(nb200_lin* lin, void* stream, const void* u_pos, void* out, int scaled);
/* transformation (likelihood.py:334-351, 630-633): Gaussian sqrt(N^-1) s, Poissonian 2 sqrt(s) */
int nb200_transformation(nb200_lin* lin, void* stream, void* out_pos);
/* normalized_residual (likelihood_impl.py:128-129, 238-239) */
int nb200_normalized_residual(nb200_lin* lin, void* stream, void* out_pos);

/* ---- conjugate gradient on the device (conjugate_gradient.py:77-214 `_cg`) -------------------- */

typedef struct nb200_cg_opts {
  double absdelta;     /* < 0: None */
  double resnorm;      /* < 0: None */
  double tol, atol;    /* used when both absdelta and resnorm are None (:103-105) */
  int32_t norm_ord;    /* 1, 2, or 0 for inf */
  int32_t miniter;     /* < 0: default (:95-98) */
  int32_t maxiter;     /* < 0: default */
  int32_t raise_nonposdef;
  int32_t check_every; /* host polls the device status every this many iterations (>=1) */
  int32_t x0_is_zero;  /* x0 = None in the reference */
  /* point estimates / constants (likelihood.py:399-499 LikelihoodPartial, evi.py:62-85): n_frozen half-open
   * ranges [lo, hi) of the flat latent vector on which the operator acts as the identity, i.e. the solve runs
   * in the subspace of the other ("liquid") entries.  j and x0 must be zero there; x stays zero there. */
  int32_t n_frozen;
  int32_t reserved;
  const int64_t* frozen;   /* host array of 2*n_frozen entries, or NULL */
} nb200_cg_opts;

typedef struct nb200_cg_result {
  int32_t info;        /* reference semantics: 0 converged, i>0 stopped at iteration i, <0 error */
  int32_t nit, nfev;
  int32_t error;       /* 0 ok, 1 zero curvature, 2 negative curvature, 3 energy increased */
  double energy;
  double gamma;        /* last <r, r> */
} nb200_cg_result;

void nb200_cg_default_opts(nb200_cg_opts* opts);
/* Solve (metric(lin) + 1) x = j, or metric_pair-composed geoVI operator if lin_b != NULL:
 *   op(t) = nb200_metric(lin, t) + t                                     (lin_b == NULL)
 *   op(t): tm = metric_pair(lin_b, lin, t) + t; out = metric_pair(lin, lin_b, tm) + tm   (evi.py:167-172, lin = x, lin_b = e)
 * x holds x0 on entry (ignored if x0_is_zero) and the solution on return. Synchronises the stream. */
int nb200_cg_solve(nb200_lin* lin, nb200_lin* lin_b, void* stream, const void* j, void* x,
                   const nb200_cg_opts* opts, nb200_cg_result* result_host);

/* ---- sample-averaged operator of the KL (optimize_kl.py:117-144 `_kl_met`, :90-114 `_kl_vg`) ----
 * out = sum_i scale * J_i^T M_i J_i t  (+ t if identity_here), accumulated in place by the last-pass epilogues of the
 * n_lins local linearisations (no extra vector passes), followed by `hook(user, out, n_elems, stream)` if given: the
 * caller's in-stream all-reduce (SUM) over the ranks that hold the other sample points (NCCL through torch.distributed
 * in the Python binding; XLA's all-reduce in a jax.ffi binding).  With scale = 1 / n_samples_total and identity_here
 * on exactly one rank the all-reduced result is mean_i(metric(x_i, t) + t).  A rank without samples passes one
 * linearisation with scale = 0. */
/* The hook is called once for every finished range [buf, buf + n_elems) of `out` -- the ranges partition the vector; with
 * nb200_plan_set_reduce_chunks(plan, n > 1) the last pass of the last linearisation runs in n launches and hands over the
 * excitation rows each launch completes right after it, so that the all-reduce of one piece travels over NVLink while the
 * next piece is computed (the hyper-parameter entries follow after the cotangent chain) -- and finally with buf == NULL,
 * n_elems == 0: every range has been handed over, the hook must make all its reductions complete in stream order
 * (e.g. wait on the communication stream).  A hook may simply all-reduce synchronously and ignore the final call. */
typedef void (*nb200_reduce_hook)(void* user, void* buf, int64_t n_elems, void* stream);
int nb200_plan_set_reduce_chunks(nb200_plan* plan, int nchunks);
int nb200_metric_multi(nb200_lin** lins, int n_lins, double scale, int identity_here, void* stream, const void* t, void* out,
                       nb200_reduce_hook hook, void* user);
/* conjugate gradient (same rules / result as nb200_cg_solve) on that operator: the Newton-CG inner solve of
 * kl_minimize (optimize_kl.py:540-591 -> optimize.py:335) without leaving the device; <d, q> is a fused reduction. */
int nb200_cg_solve_multi(nb200_lin** lins, int n_lins, double scale, int identity_here, void* stream, const void* j, void* x,
                         const nb200_cg_opts* opts, nb200_cg_result* result_host, nb200_reduce_hook hook, void* user);

/* ---- fused vector algebra on flat latent vectors (tree_math vdot/norm/axpy; evi.py:136) ---------- */
int nb200_vec_axpby(nb200_plan* plan, void* stream, int64_t n, double a, const void* x, double b, const void* y, void* out);
int nb200_vec_dot(nb200_plan* plan, void* stream, int64_t n, const void* x, const void* y, double* out_host);
/* out_host[0] = sum x_i, out_host[1] = sum x_i^2: the moments of reduced_residual_stats (minisanity.py:17-21, 30-82) */
int nb200_vec_stats(nb200_plan* plan, void* stream, int64_t n, const void* x, double* out_host);

#ifdef __cplusplus
}
#endif
#endif /* NIFTY_B200_H */
