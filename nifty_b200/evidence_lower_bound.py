"""``estimate_evidence_lower_bound`` with the interface of ``nifty/re/evidence_lower_bound.py`` (eigenvalue path).

ELBO = -<H(xi)>_samples + (N + Tr log Lambda) / 2 with Lambda the inverse of the metric ``M + 1`` at the mean
(:414-700).  ``Tr log Lambda = -sum_i log lambda_i`` over the eigenvalues of ``M + 1``; all but ``min(#data, #latent)``
of them equal one, the largest ones are found with ARPACK in batches, each batch on the operator projected onto the
complement of the eigenvectors found so far (:176-411), and the unresolved tail enters ``lower_error`` (:826-832).

Every operator application is one fused metric-vector product on the device (``nb200_metric`` with the identity added,
or ``nb200_lsm`` + ``nb200_rsm`` for the data-space operator ``R M^(1/2) ... `` of :154-173); SciPy only sees a
``LinearOperator`` on host vectors.  ``trace_log_method="slq"`` is the stochastic Lanczos quadrature of the same trace-log
(``lanczos.py``): plain (``n_eigenvalues = 0``), or the reference's hybrid -- the largest eigenvalues exactly, the remainder on
probes deflated by their eigenvectors, optionally bracketed by Gauss-Radau quadratures (``use_radau_as_bound=True``).  The
``analytic_prior_term`` evaluates the prior energy in closed form from ``tr (M + 1)^-1``; a stored eigensystem can be resumed
(``resume_eigenvectors`` / ``resume_eigenvalues``); ``slq_kwargs`` / ``slq_jit`` raise.
"""

from __future__ import annotations

import logging
import os
from typing import Optional

import numpy as np
import scipy.linalg as slg
import scipy.sparse.linalg as ssl
import torch

from .evi import Samples
from .likelihood import LikelihoodWithModel

logger = logging.getLogger("nifty_b200")


def _orthonormality_error(vecs: np.ndarray, n_probes: int) -> float:
    """max |V^T V p - p| over a few random p (:100-109)."""
    if vecs.size == 0:
        return 0.0
    k = vecs.shape[1]
    probes = np.random.default_rng(0).standard_normal((k, min(n_probes, k)))
    return float(np.max(np.abs(vecs.conj().T @ (vecs @ probes) - probes)))


def _deflated(op: ssl.LinearOperator, vecs: Optional[np.ndarray]) -> ssl.LinearOperator:
    """P op P with P = 1 - V V^T (:28-65): the spectrum of ``op`` without the directions already found."""
    if vecs is None:
        return op

    def proj(x):
        return x - vecs @ (vecs.conj().T @ x)

    return ssl.LinearOperator(shape=op.shape, dtype=op.dtype, matvec=lambda x: proj(op.matvec(proj(x))))


def _largest_eigenvalues(op, size, n_eigenvalues, tot_dofs, *, min_lh_eval, eigenvalue_shift, solver_shift, n_batches, tol,
                         early_stop, verbose, output_directory, prefix, orthonormalize, every, threshold, n_probes,
                         resume_eigenvectors=None, resume_eigenvalues=None):
    """The ``n_eigenvalues`` largest eigenvalues of ``op`` (:176-411): dense when all relevant ones are requested,
    otherwise ARPACK batch by batch with deflation, stopping early once the smallest one found is within
    ``min_lh_eval`` of the value every remaining eigenvalue has (``eigenvalue_shift``).  ``resume_*``: an eigensystem computed
    earlier (:199-268) -- checked against the operator, sorted, truncated, and used as the first deflation set."""
    if n_eigenvalues > tot_dofs:
        raise ValueError("Number of requested eigenvalues exceeds the number of relevant degrees of freedom!")
    if resume_eigenvalues is not None and resume_eigenvectors is None:
        raise ValueError("resume_eigenvalues requires resume_eigenvectors.")
    r_vals, r_vecs = None, None
    if resume_eigenvectors is not None:
        r_vecs = np.asarray(resume_eigenvectors, dtype=np.float64)
        if r_vecs.ndim != 2:
            raise ValueError("resume_eigenvectors must be a 2D array.")
        if r_vecs.shape[0] != size:
            raise ValueError("resume_eigenvectors does not match the operator size.")
        estimated = np.array([r_vecs[:, i] @ op.matvec(r_vecs[:, i]) for i in range(r_vecs.shape[1])])      # Rayleigh quotients (:86-90)
        r_vals = estimated if resume_eigenvalues is None else np.asarray(resume_eigenvalues, dtype=np.float64)
        if r_vals.ndim != 1:
            raise ValueError("resume_eigenvalues must be a 1D array.")
        if r_vals.size != r_vecs.shape[1]:
            raise ValueError("resume_eigenvalues and resume_eigenvectors have mismatched sizes.")
        if not np.allclose(r_vals, estimated, rtol=1e-5, atol=1e-8):
            raise ValueError("The resumed eigensystem does not match the selected operator. Check trace_log_space and the eigensystem source.")
        order = np.argsort(-r_vals)
        r_vals, r_vecs = r_vals[order][:max(n_eigenvalues, 0)], r_vecs[:, order][:, :max(n_eigenvalues, 0)]
        if orthonormalize and threshold is not None and r_vecs.shape[1] and _orthonormality_error(r_vecs, n_probes) > threshold:
            r_vecs, _ = np.linalg.qr(r_vecs)
        if r_vals.size > tot_dofs:
            raise ValueError("Number of provided eigenvectors exceeds relevant degrees of freedom.")
        if r_vals.size == 0:
            r_vals, r_vecs = None, None
    if n_eigenvalues <= 0:
        return np.asarray([], dtype=np.float64), None
    if r_vals is not None and (r_vals.size >= n_eigenvalues or (early_stop and abs(eigenvalue_shift - np.min(r_vals)) < min_lh_eval)):
        return r_vals, r_vecs

    def save(vals, vecs):
        if output_directory is None:
            return
        d = output_directory or "."
        os.makedirs(d, exist_ok=True)
        np.save(os.path.join(d, f"{prefix}_eigenvalues.npy"), vals)
        if vecs is not None:
            np.save(os.path.join(d, f"{prefix}_eigenvectors.npy"), vecs)

    if tot_dofs == n_eigenvalues and r_vals is None:
        if verbose:
            logger.info(f"Computing all {tot_dofs} relevant eigenvalues.")
        dense = np.column_stack([op.matvec(e) for e in np.identity(size)])
        vals, vecs = slg.eigh(dense, subset_by_index=[size - tot_dofs, size - 1])
        order = np.argsort(-vals)
        save(vals[order], vecs[:, order])
        return vals[order], vecs[:, order]
    n_pre = 0 if r_vals is None else r_vals.size
    base, rem = divmod(n_eigenvalues - n_pre, n_batches)
    batches = [b for b in [base + 1] * rem + [base] * (n_batches - rem) if b > 0]
    solver_op = op if solver_shift == 0.0 else ssl.LinearOperator(shape=op.shape, dtype=op.dtype,
                                                                   matvec=lambda x: op.matvec(x) + solver_shift * x)
    vals, vecs = r_vals, r_vecs
    for count, batch in enumerate(batches, start=1):
        if verbose:
            logger.info(f"\nNumber of eigenvalues being computed: {batch}")
        bv, bw = ssl.eigsh(_deflated(solver_op, vecs), k=batch, tol=tol, return_eigenvectors=True, which="LM")
        bv = np.real_if_close(bv - solver_shift)
        order = np.argsort(-bv)
        bv, bw = bv[order], bw[:, order]
        vals = bv if vals is None else np.concatenate((vals, bv))
        vecs = bw if vecs is None else np.hstack((vecs, bw))
        if orthonormalize:
            err = _orthonormality_error(vecs, n_probes) if threshold is not None else None
            if (err is not None and err > threshold) or count % every == 0:
                vecs, _ = np.linalg.qr(vecs)
        save(vals, vecs)
        if verbose:
            logger.info(f"Eigenvalue progress: {vals.size}/{n_eigenvalues}")
        if early_stop and abs(eigenvalue_shift - np.min(vals)) < min_lh_eval:
            break
    return vals, vecs


def _analytic_prior(likelihood, samples, eigenvalues, use_data, trace_inv_remainder, trace_inv_se, n_unit):
    """The prior energy of the ELBO in closed form (:809-824, 953-979): ``1/2 (tr (M + 1)^-1 + <mean, mean>)`` with the trace split
    into the exact eigenvalues, the SLQ remainder and the ``n_unit`` eigenvalues that equal one."""
    pos = samples.pos
    prior_mean_sq = float(likelihood.vdot(pos, pos)) if pos is not None else 0.0
    trace_inv_exact = float(np.sum(1.0 / (eigenvalues + float(use_data)))) if eigenvalues.size else 0.0
    total = trace_inv_exact + trace_inv_remainder + float(n_unit)
    prior_term = 0.5 * (total + prior_mean_sq)
    return prior_term, {"trace_inv_exact": trace_inv_exact, "trace_inv_slq": float(trace_inv_remainder), "trace_inv_const": float(n_unit),
                        "trace_inv_se": float(trace_inv_se), "trace_inv_total": total, "prior_mean_sq": prior_mean_sq,
                        "prior_term": prior_term}


def estimate_evidence_lower_bound(likelihood: LikelihoodWithModel, samples: Samples, n_eigenvalues, *, compute_all=False,
                                  min_lh_eval=1e-3, n_batches=10, tol=0.0, verbose=True, metric_jit=True,
                                  output_directory: Optional[str] = None, save_eigensystem_prefix="metric",
                                  resume_eigenvectors=None, resume_eigenvalues=None, orthonormalize_eigenvectors=True,
                                  orthonormalize_every_n_batches=8, orthonormalize_threshold=1e-6, orthonormalize_n_probes=2,
                                  trace_log_method="eigsh", trace_log_space="signal", analytic_prior_term=False, **slq_options):
    """Returns ``(elbo_samples, stats)`` (:414-1042, eigenvalue path).  ``stats`` holds ``elbo_mean``, ``elbo_std``,
    ``elbo_se``, ``elbo_up``, ``elbo_lw`` and ``lower_error`` with the reference's definitions (:959-1005).

    Differences from the reference's signature: ``output_directory`` defaults to ``None`` (the reference's default ``""``
    writes the eigensystem into the working directory); resuming from a stored eigensystem and the SLQ options raise."""
    if not isinstance(samples, Samples):
        raise TypeError("samples attribute should be of type `Samples`.")
    if not isinstance(likelihood, LikelihoodWithModel):
        raise TypeError("likelhood is not an instance of `Likelihood`.")
    trace_log_method, trace_log_space = trace_log_method.lower(), trace_log_space.lower()
    if trace_log_method not in ("eigsh", "slq"):
        raise ValueError("trace_log_method must be 'eigsh' or 'slq'.")
    if trace_log_space not in ("auto", "signal", "data"):
        raise ValueError("trace_log_space must be 'auto', 'signal', or 'data'.")
    slq_order = int(slq_options.pop("slq_order", 30))
    slq_num_samples = int(slq_options.pop("slq_num_samples", 16))
    slq_key = slq_options.pop("slq_key", None)
    use_radau_as_bound = bool(slq_options.pop("use_radau_as_bound", False))
    if any(v is not None and v is not False for v in slq_options.values()):
        raise NotImplementedError("slq_kwargs / slq_jit are not provided on the B200 path; trace_log_method='eigsh', the stochastic Lanczos "
                                  "quadrature trace_log_method='slq' (plain, or hybrid with n_eigenvalues exact eigenvalues deflated, optionally "
                                  "with use_radau_as_bound=True) and analytic_prior_term are")
    if likelihood.signal.cf.plan.dist:
        raise NotImplementedError("estimate_evidence_lower_bound on slab-decomposed fields is not supported")
    if orthonormalize_eigenvectors and resume_eigenvectors is not None and resume_eigenvalues is None:
        raise ValueError("resume_eigenvalues is required when orthonormalize_eigenvectors=True.")
    if orthonormalize_eigenvectors and (not isinstance(orthonormalize_every_n_batches, int) or orthonormalize_every_n_batches < 1):
        raise ValueError("orthonormalize_every_n_batches must be a positive integer.")

    rt, dtype = likelihood.rt, likelihood.dtype
    pos = samples.pos
    lin, _ = likelihood.lin_at(pos)
    metric_size = int(likelihood.layout.size)
    data_shape = tuple(likelihood.signal.target_shape)
    n_data = int(np.prod(data_shape))
    n_relevant = min(n_data, metric_size)
    use_data = trace_log_space == "data" or (trace_log_space == "auto" and n_data <= metric_size)

    def to_dev(x, shape):
        return rt.asarray(np.ascontiguousarray(x, dtype=np.float64).reshape(shape), dtype)

    if use_data:      # u -> RSM(LSM(u)) on data vectors; eigenvalues of M are those of this operator (:154-173)
        op_size, eig_shift, solver_shift, log_np = n_data, 0.0, 1.0, np.log1p

        def matvec(x):
            return lin.rsm(lin.lsm(to_dev(x, data_shape), scaled=True), scaled=True).reshape(-1).to(torch.float64).cpu().numpy()
    else:
        op_size, eig_shift, solver_shift, log_np = metric_size, 1.0, 0.0, np.log

        def matvec(x):
            return lin.metric(to_dev(x, (metric_size,)), add_identity=True).to(torch.float64).cpu().numpy()
    op = ssl.LinearOperator(shape=(op_size, op_size), dtype=np.float64, matvec=matvec)

    if trace_log_method == "slq":
        # tr log M by stochastic Lanczos quadrature on the same operator (nifty_b200/lanczos.py; every Lanczos step is one fused
        # device product): signal space log det(metric + 1), data space log det(1 + RSM LSM).  With n_eigenvalues > 0 the HYBRID of
        # the reference (:833-937): the largest eigenvalues exactly (ARPACK, deflation), the remainder by SLQ on probes projected
        # onto the complement of their eigenvectors; use_radau_as_bound=True adds the Gauss-Radau bracket of the remainder with
        # nodes at the two ends of the remaining spectrum, [1, smallest exact eigenvalue] of the shifted operator.
        from .lanczos import slq_gauss_radau
        rng = slq_key if isinstance(slq_key, np.random.Generator) else np.random.default_rng(0 if slq_key is None else slq_key)
        shift = 0.0 if not use_data else 1.0
        if not isinstance(n_eigenvalues, (int, np.integer)) or n_eigenvalues < 0:
            raise ValueError("n_eigenvalues must be a non-negative integer.")
        if n_eigenvalues > n_relevant:
            raise ValueError("Number of requested eigenvalues exceeds the number of relevant degrees of freedom!")
        exact_log, eigenvalues, eigenvectors = 0.0, np.asarray([]), None
        if n_eigenvalues > 0:
            eigenvalues, eigenvectors = _largest_eigenvalues(
                op, op_size, int(n_eigenvalues), n_relevant, min_lh_eval=min_lh_eval, eigenvalue_shift=eig_shift, solver_shift=solver_shift,
                n_batches=n_batches, tol=tol, early_stop=False, verbose=verbose, output_directory=output_directory,
                prefix=f"{save_eigensystem_prefix}_{'data' if use_data else 'signal'}", orthonormalize=orthonormalize_eigenvectors,
                every=orthonormalize_every_n_batches, threshold=orthonormalize_threshold, n_probes=orthonormalize_n_probes,
                resume_eigenvectors=resume_eigenvectors, resume_eigenvalues=resume_eigenvalues)
            exact_log = float(np.sum(log_np(eigenvalues)))
        if use_radau_as_bound and eigenvalues.size == 0:
            raise ValueError("use_radau_as_bound=True requires a valid upper spectral endpoint from at least one exact eigenvalue.")
        remainder, slq_se, tail_hi = 0.0, 0.0, None
        trace_inv_remainder, trace_inv_se = 0.0, 0.0
        if n_relevant - eigenvalues.size > 0:
            if slq_num_samples < 2 and eigenvalues.size > 0:
                raise ValueError("Estimating an SLQ remainder requires at least two probes to quantify stochastic uncertainty.")

            def dev_matvec(v):
                return torch.as_tensor(matvec(v.cpu().numpy()), dtype=torch.float64) + shift * v

            radau = {}
            if use_radau_as_bound:
                lam_min = 1.0                                         # eigenvalue_shift of the SHIFTED operator in both spaces
                radau = dict(lam_min=lam_min, lam_max=max(lam_min, float(np.min(eigenvalues)) + shift), compute_radau=True)
            extra = {"inv": lambda x: 1.0 / x - 1.0} if analytic_prior_term else None     # centred: unit eigenvalues contribute nothing
            out = slq_gauss_radau(dev_matvec, torch.log, min(slq_order, op_size), slq_num_samples, rng, shape0=op_size,
                                  deflate_eigvecs=eigenvectors, extra_fns=extra, **radau)
            remainder, slq_se = out["estimate"], out["stochastic_se"]
            if analytic_prior_term:
                trace_inv_remainder = (n_relevant - eigenvalues.size) + out["extra_inv_estimate"]
                trace_inv_se = out["extra_inv_se"]
            if use_radau_as_bound:
                lo, hi = out["radau_lo"], out["radau_hi"]
                if not np.isfinite(lo) or not np.isfinite(hi):
                    raise ValueError("Gauss-Radau quadrature failed because an endpoint is too close to the Lanczos spectrum.")
                tail_hi = max(lo, hi)
        logdet = exact_log + remainder
        lower = 0.5 * max(0.0, tail_hi - remainder) if tail_hi is not None else (0.5 * slq_se if eigenvalues.size else 0.0)
        posterior_contribution = -0.5 * logdet + 0.5 * metric_size
        pts = [samples[i] for i in range(len(samples))]
        stats = {}
        prior_term = 0.0
        if analytic_prior_term:
            prior_term, pstats = _analytic_prior(likelihood, samples, eigenvalues, use_data, trace_inv_remainder, trace_inv_se,
                                                 metric_size - n_relevant)
            stats.update(pstats)
            if trace_inv_se > 0.0:
                lower += 0.5 * trace_inv_se
            energies = [likelihood.energy(s) for s in pts]                     # the prior part is analytic (:959)
        else:
            energies = [likelihood.energy(s) + 0.5 * likelihood.vdot(s, s) for s in pts]
        elbo_samples = np.array([posterior_contribution - h - prior_term for h in energies])
        mean = float(np.mean(elbo_samples)) if len(pts) else float("nan")
        std = float(np.std(elbo_samples, ddof=1)) if len(pts) > 1 else float("nan")
        stats.update({"lower_error": lower, "slq_stochastic_se": 0.5 * slq_se, "slq_remainder": remainder, "exact_log": exact_log,
                      "trace_log_exact": exact_log, "trace_log_slq": remainder, "trace_log_se": slq_se,
                      "elbo_lw": mean - std - lower, "elbo_mean": mean, "elbo_up": mean + std,
                      "elbo_std": std, "elbo_se": std / np.sqrt(len(pts)) if len(pts) > 0 else 0.0})
        return elbo_samples, stats
    if compute_all:
        n_eigenvalues = n_relevant
    if not isinstance(n_eigenvalues, (int, np.integer)):
        raise TypeError("n_eigenvalues must be an integer.")
    if n_eigenvalues < 0:
        raise ValueError("n_eigenvalues must be non-negative.")
    if n_relevant > 0 and n_eigenvalues == 0:
        raise ValueError("trace_log_method='eigsh' requires at least one eigenvalue. "
                         "Use trace_log_method='slq' to estimate the full trace stochastically.")
    prefix = f"{save_eigensystem_prefix}_{'data' if use_data else 'signal'}"
    eigenvalues, _ = _largest_eigenvalues(
        op, op_size, int(n_eigenvalues), n_relevant, min_lh_eval=min_lh_eval, eigenvalue_shift=eig_shift, solver_shift=solver_shift,
        n_batches=n_batches, tol=tol, early_stop=not compute_all, verbose=verbose, output_directory=output_directory, prefix=prefix,
        orthonormalize=orthonormalize_eigenvectors, every=orthonormalize_every_n_batches, threshold=orthonormalize_threshold,
        n_probes=orthonormalize_n_probes, resume_eigenvectors=resume_eigenvectors, resume_eigenvalues=resume_eigenvalues)
    if verbose:
        logger.info(f"\nComputed {eigenvalues.size} largest eigenvalues (out of {n_relevant} relevant degrees of freedom).")

    log_eigs = log_np(eigenvalues)
    tr_log_lat_cov = -0.5 * float(np.sum(log_eigs))                                             # :826-827
    lower_error = 0.5 * (n_relevant - log_eigs.size) * float(np.min(log_eigs)) if log_eigs.size else 0.0   # :828-832
    posterior_contribution = tr_log_lat_cov + 0.5 * metric_size                                  # :956
    pts = [samples[i] for i in range(len(samples))]
    stats, prior_term = {}, 0.0
    if analytic_prior_term:
        if n_relevant > log_eigs.size:
            raise ValueError("analytic_prior_term requires trace_log_method='slq' or compute_all=True when not all eigenvalues are computed.")
        prior_term, stats = _analytic_prior(likelihood, samples, eigenvalues, use_data, 0.0, 0.0, metric_size - n_relevant)
        ham = [likelihood.energy(s) for s in pts]
    else:
        ham = [likelihood.energy(s) + 0.5 * likelihood.vdot(s, s) for s in pts]                  # StandardHamiltonian (:67-87)
    elbo_samples = np.array([posterior_contribution - h - prior_term for h in ham])
    stats["lower_error"] = lower_error
    mean = float(np.mean(elbo_samples)) if len(pts) else float("nan")
    std = float(np.std(elbo_samples, ddof=1)) if len(pts) > 1 else float("nan")
    stats["elbo_lw"], stats["elbo_mean"], stats["elbo_up"] = mean - std - lower_error, mean, mean + std
    stats["elbo_std"] = std
    stats["elbo_se"] = std / np.sqrt(len(pts)) if len(pts) > 0 else 0.0
    if verbose:
        logger.info(f"\nELBO decomposition (in log units)\nELBO mean : {mean:.4e} (lower: {stats['elbo_lw']:.4e}, "
                    f"upper: {stats['elbo_up']:.4e})\nELBO std  : {std:.4e}")
    return elbo_samples, stats
