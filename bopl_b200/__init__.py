"""bopl_b200 — B200-native (sm_100a) implementation of BOPL's data-parallel hot path: the multi-output GP posterior
and the uEI / uEI_affine acquisition over large candidate batches, behind the reference's own Python interfaces.

    from bopl_b200 import MultiOutputGP, Utility, UtilityDistribution, uEI, uEI_affine

The computation lives in libbopl_b200.so (hand-written CUDA, C-ABI in include/bopl_b200.h).  Importing this package
does not need a GPU; constructing a model does, and there is no CPU fallback.
"""
from . import _lib
from ._lib import BoplError
from .utility import Utility, UtilityDistribution
from .anchor_points import ObjectiveAnchorPointsGenerator, BoxSpace


def __getattr__(name):
    # engine / models / acquisition import torch lazily through these names
    if name in ("MultiOutputGP", "BasicModel", "state_from_hyperparameters", "state_from_reference_model", "fit_exact_gp"):
        from . import models
        return getattr(models, name)
    if name in ("uEI", "uEI_affine", "uEI_constrained", "AcquisitionBase"):
        from . import acquisition
        return getattr(acquisition, name)
    if name in ("AcquisitionOptimizer", "OptLbfgs", "apply_optimizer"):
        from . import optimization
        return getattr(optimization, name)
    if name in ("AcquisitionFunction", "Random", "uTS", "ParEGO", "SamplePath", "Sequential", "chebyshev_scalarization"):
        from . import sampling_policies
        return getattr(sampling_policies, name)
    if name in ("Engine", "GPState"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
