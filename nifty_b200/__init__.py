"""nifty_b200 -- B200-native MGVI/geoVI inner loop behind the ``nifty.re`` interface.

Host side (this package): the reference's names and call protocol.  Device side
(``nifty_b200/csrc`` -> ``lib/libniftyb200.so``): hand-written sm_100a CUDA behind the C ABI of
``include/nifty_b200.h``.  Importing the package does not need a GPU; using it does.
"""

from ._capi import NB200Error  # noqa: F401
from ._runtime import Lin, ModelHandle, Plan, Runtime, default_runtime  # noqa: F401
from .prior import (LogNormalPrior, NormalPrior, laplace_prior, lognormal_invprior, lognormal_moments, lognormal_prior,  # noqa: F401
                    normal_invprior, normal_prior, uniform_prior)
from .tree import Layout  # noqa: F401
from .correlated_field import CorrelatedField, CorrelatedFieldMaker, get_fourier_mode_distributor, hartley, make_grid  # noqa: F401
from .likelihood import (Gaussian, Likelihood, LikelihoodPartial, LikelihoodSum, LikelihoodWithModel, OperatorLikelihood,  # noqa: F401
                         Poissonian,
                         SignalModel)
from .conjugate_gradient import CGResults, HamiltonianMetric, cg, static_cg  # noqa: F401
from .optimize import OptimizeResults, minimize, newton_cg, static_newton_cg  # noqa: F401
from .evi import (Samples, concatenate_zip, draw_linear_residual, draw_residual, nonlinearly_update_residual,  # noqa: F401
                  random_like, random_split, sample_likelihood, wiener_filter_posterior)
from .optimize_kl import OptimizeVI, OptimizeVIState, get_status_message, optimize_kl  # noqa: F401
from .minisanity import ChiSqStats, minisanity, reduced_residual_stats  # noqa: F401
from .tree_math import (Vector, get_map, lmap, mean, mean_and_std, norm, ravel, size, smap, stack, unstack, vdot, where,  # noqa: F401
                        zeros_like)
from .model import Initializer, LazyModel, Model, VModel, WrappedCall  # noqa: F401
from .evidence_lower_bound import estimate_evidence_lower_bound  # noqa: F401
from .outer import OuterCorrelatedField, OuterLikelihood  # noqa: F401
from .bluestein import BluesteinCorrelatedField, BluesteinHartley, fourier_mode_tables  # noqa: F401
from . import lanczos  # noqa: F401
from .lanczos import lanczos_tridiag, slq_gauss_radau, stochastic_logdet_from_lanczos, stochastic_lq_logdet  # noqa: F401
