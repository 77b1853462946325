"""
Composition wrappers that every Verde example puts around the spline:
``Vector`` (one scalar gridder per data component, verde/vector.py:28-141) and
``Chain`` (gridders applied in sequence to each other's residuals,
verde/chain.py:17-138).  Pure host glue: they only ever call ``fit`` /
``filter`` / ``predict`` on their members, so the B200 gridders slot in
unchanged.
"""
from sklearn.utils.validation import check_is_fitted

from .base import BaseGridder, check_data, check_fit_input
from .coords import get_region


class Vector(BaseGridder):
    """
    Fit one estimator per component of multi-component data.

    ``components`` is a list of gridders (e.g. ``[Spline(damping=...)] * 2`` must be
    two distinct instances, as in Verde); ``predict`` returns one array per component.
    """

    def __init__(self, components):
        super().__init__()
        self.components = components

    def fit(self, coordinates, data, weights=None):
        if not isinstance(data, tuple):
            raise ValueError("Data must be a tuple of arrays. {} given.".format(type(data)))
        if weights is not None and not isinstance(weights, tuple):
            raise ValueError("Weights must be a tuple of arrays. {} given.".format(type(weights)))
        coordinates, data, weights = check_fit_input(coordinates, data, weights)
        self.region_ = get_region(coordinates[:2])
        for estimator, data_comp, weight_comp in zip(self.components, data, weights):
            estimator.fit(coordinates, data_comp, weight_comp)
        return self

    def predict(self, coordinates):
        check_is_fitted(self, ["region_"])
        return tuple(comp.predict(coordinates) for comp in self.components)


class Chain(BaseGridder):
    """
    Apply ``steps`` = [(name, gridder), ...] in sequence: each step is fitted to the
    residuals of the previous ones (through ``filter``); ``predict`` sums the steps
    that can predict.
    """

    def __init__(self, steps):
        super().__init__()
        self.steps = steps

    @property
    def named_steps(self):
        "Access the steps by name"
        return dict(self.steps)

    def fit(self, coordinates, data, weights=None):
        self.region_ = get_region(coordinates[:2])
        args = coordinates, data, weights
        for _, step in self.steps:
            args = step.filter(*args)
        return self

    def predict(self, coordinates):
        check_is_fitted(self, ["region_"])
        result = None
        for _, step in self.steps:
            if hasattr(step, "predict"):
                predicted = check_data(step.predict(coordinates))
                if result is None:
                    result = [0 for _ in range(len(predicted))]
                for i, pred in enumerate(predicted):
                    result[i] += pred
        if len(result) == 1:
            return result[0]
        return tuple(result)
