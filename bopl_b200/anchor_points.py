"""The batch entry of the hot path: candidate generation -> acquisition scores -> top-k anchors.

Drop-in for `GPyOpt/optimization/anchor_points_generator.py:9-98` (AnchorPointsGenerator /
ObjectiveAnchorPointsGenerator) as used by `bopl/optimization_services/acquisition_optimizer.py:79-84`.
When the objective is the `acquisition_function` of a GPU acquisition (bopl_b200.acquisition.uEI) the scores never
come back to the host: scoring and top-k selection run in one library call (`bopl_uei_host` with k > 0).

Tie order: the reference's `np.argsort` is an unstable introsort, so the order among exactly equal scores is
unspecified there; here it is "best score first, then lowest candidate index" (SURVEY.md §7.1-9).
"""
import numpy as np


def samples_multidimensional_uniform(bounds, points_count, seed=None):
    """Same draw as GPyOpt/experiment_design/random_design.py:57-71 (one `uniform(size=(N, dim))` then per-column affine)."""
    dim = len(bounds)
    rng = np.random.RandomState(seed) if seed is not None else np.random
    Z = rng.uniform(size=(points_count, dim))
    for k in range(dim):
        Z[:, k] = (bounds[k][1] - bounds[k][0]) * Z[:, k] + bounds[k][0]
    return Z


class BoxSpace(object):
    """Minimal continuous design space (the part of GPyOpt's Design_space the path touches)."""

    def __init__(self, bounds):
        self.bounds = [tuple(map(float, b)) for b in bounds]
        self.dimensionality = len(self.bounds)

    def get_bounds(self):
        return list(self.bounds)

    get_continuous_bounds = get_bounds

    def has_constraints(self):
        return False


def latin_design(bounds, points_count):
    """Centred latin hypercube as GPyOpt/experiment_design/latin_design.py:18-41 draws it through pyDOE
    (`lhs(dim, n, criterion='center')`: one unused `np.random.rand(n, dim)`, then one `np.random.permutation` of the bin
    centres per dimension), scaled to the bounds."""
    dim = len(bounds)
    cut = np.linspace(0, 1, points_count + 1)
    np.random.rand(points_count, dim)                     # pyDOE draws it before branching on the criterion
    centre = (cut[:points_count] + cut[1:points_count + 1]) / 2
    H = np.zeros((points_count, dim))
    for j in range(dim):
        H[:, j] = np.random.permutation(centre)
    b = np.asarray(bounds, dtype=np.float64)
    return b[:, 0][None, :] + H * (b[:, 1] - b[:, 0])[None, :]


def initial_design(design_type, space, init_points_count, seed=None):
    if design_type == "random":
        return samples_multidimensional_uniform(space.get_bounds(), init_points_count, seed)
    if design_type == "latin":
        return latin_design(space.get_bounds(), init_points_count)
    raise NotImplementedError("only the 'random' and 'latin' designs are on the accelerated path "
                              "(acquisition_optimizer.py:79,266)")


class AnchorPointsGenerator(object):
    def __init__(self, space, design_type, num_samples):
        self.space = space
        self.design_type = design_type
        self.num_samples = num_samples

    def get_anchor_point_scores(self, X):
        raise NotImplementedError("get_anchor_point_scores is not implemented in the parent class.")

    def get(self, num_anchor=20, duplicate_manager=None, unique=False, context_manager=None, get_scores=False):
        X = initial_design(self.design_type, self.space, self.num_samples)
        return self.select(X, num_anchor, get_scores)

    def select(self, X, num_anchor, get_scores=False):
        scores = self.get_anchor_point_scores(X)
        k = min(len(scores), num_anchor)
        order = np.argsort(scores, kind="stable")[:k]
        if get_scores:
            return X[order, :], scores[order]
        return X[order, :]


class ObjectiveAnchorPointsGenerator(AnchorPointsGenerator):
    """Locations with the minimum objective (= -acquisition) value.  If `objective` is the bound
    `acquisition_function` of a bopl_b200 acquisition exposing `top_candidates`, the fused GPU path is taken."""

    def __init__(self, space, design_type, objective, num_samples=18):
        super(ObjectiveAnchorPointsGenerator, self).__init__(space, design_type, num_samples)
        self.objective = objective

    def get_anchor_point_scores(self, X):
        return self.objective(X).flatten()

    def select(self, X, num_anchor, get_scores=False):
        acq = getattr(self.objective, "__self__", None)
        fused = acq is not None and getattr(self.objective, "__name__", "") == "acquisition_function" and \
            hasattr(acq, "top_candidates")
        if not fused:
            return super(ObjectiveAnchorPointsGenerator, self).select(X, num_anchor, get_scores)
        k = min(len(X), num_anchor)
        val, idx = acq.top_candidates(X, k)
        if get_scores:
            return X[idx, :], -val                               # scores are -acq (base.py:40), ascending
        return X[idx, :]
