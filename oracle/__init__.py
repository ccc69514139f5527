"""CPU oracle for the NIFTy.re MGVI/geoVI inner loop -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy float64 restatement of the algorithm that
``nifty.re`` (reference tree ``/root/reference/nifty/re``) runs for the
correlated-field forward model, its JVP/VJP, the Gaussian/Poissonian
metric-vector product, conjugate gradient, Newton-CG, the MGVI linear sample
draw, the geoVI non-linear sample update and the sample-averaged KL.

Rules (enforced by tests/test_layout_rules.py):

* Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
  ``cpu_baseline`` / ``--impl reference`` leg may import this package, and
  only as the *checker* or the *CPU baseline*, never as the product path.
* ``nifty_b200`` (the product) must never import it; the product fails
  loudly when the CUDA library is missing.

Pinning status: ``nifty.re`` itself cannot run in the build container or on
the GPU box (JAX is not installed, no network).  The oracle is pinned against
the reference's *second* implementation of the same mathematics,
``nifty.cl`` (imported unmodified from ``/root/reference`` through two
out-of-tree shims, see ``tests/golden/make_golden.py``): field values, JVP,
VJP, metric and energies agree to <= 1e-12 relative, and the committed
fixtures under ``tests/golden/`` hold those ``nifty.cl`` outputs.  The
reference's own test-suite pins ``nifty.re`` to ``nifty.cl`` for exactly
these quantities (``test/test_re/test_correlated_field.py:116-192``), there
are no stored golden vectors in the reference tree.  CG / Newton-CG /
sample-draw control flow follows ``nifty.re`` line by line and is pinned by
the reference's own known-answer tests restated in ``tests/test_oracle_*.py``
(``test/test_re/test_ncg.py:112-154``, ``test/test_re/test_evi.py:136-204``).
"""

from .correlated_field import (  # noqa: F401
    CorrelatedFieldOracle,
    FourierGrid,
    fourier_mode_distributor,
    hartley,
    lognormal_moments,
    make_fourier_grid,
)
from .likelihood import GaussianOracle, PoissonianOracle, SignalOracle  # noqa: F401
from .solvers import cg, newton_cg, CGResult, NewtonResult  # noqa: F401
from .vi import (  # noqa: F401
    draw_linear_residual,
    wiener_filter_posterior_mean,
    nonlinearly_update_residual,
    kl_value_and_grad,
    kl_metric,
    kl_minimize,
    tree_size,
    Layout,
)
