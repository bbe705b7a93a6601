"""bo_pr_b200 -- B200-native implementation of bo_pr's probabilistic-reparameterization acquisition hot path.

Public names mirror the reference (``discrete_mixed_bo``): the PR acquisition wrappers, the PR input transforms, the
mixed-kernel factory and the multi-start optimisation drivers.  All arithmetic on the path runs in ``libprb200.so``
(hand-written sm_100a CUDA behind the C ABI in ``include/prb200.h``); there is no CPU fallback.
"""
from .acquisition import (  # noqa: F401
    AcquisitionFunction,
    ConstrainedExpectedImprovement,
    ExpectedHypervolumeImprovement,
    ExpectedImprovement,
    ModelListAcquisitionFunction,
    PosteriorMean,
    UpperConfidenceBound,
    is_nonnegative,
)
from .input import (  # noqa: F401
    AnalyticProbabilisticReparameterizationInputTransform,
    CategoricalSpec,
    ChainedInputTransform,
    LatentCategoricalEmbedding,
    LatentCategoricalSpec,
    MCProbabilisticReparameterizationInputTransform,
    Normalize,
    OneHotToNumeric,
)
from .kernels import (  # noqa: F401
    AdditiveKernel,
    CategoricalKernel,
    MaternKernel,
    ProductKernel,
    ScaleKernel,
    get_kernel,
)
from .models import FixedNoiseGP, GPState, ModelListGP  # noqa: F401
from .multi_objective import FastNondominatedPartitioning, NondominatedPartitioning, is_non_dominated  # noqa: F401
from .probabilistic_reparameterization import (  # noqa: F401
    AnalyticProbabilisticReparameterization,
    MCProbabilisticReparameterization,
    get_probabilistic_reparameterization_input_transform,
)
from .wrapper import AcquisitionFunctionWrapper  # noqa: F401
from .gen import gen_candidates_scipy, gen_candidates_torch, get_best_candidates  # noqa: F401
from .initializers import (  # noqa: F401
    gen_batch_initial_conditions,
    initialize_q_batch,
    initialize_q_batch_nonneg,
)
from .optimize import optimize_acqf, optimize_acqf_mixed  # noqa: F401
from .finalize import finalize_candidates  # noqa: F401
from .rffs import (  # noqa: F401
    RandomFourierFeatures,
    RFFSampleModel,
    get_gp_sample_w_transforms,
    get_gp_samples,
)
from .trust_region import TurboState, trust_region_bounds, trust_region_center, update_state  # noqa: F401

__version__ = "0.1.0"
