"""Optimisation rules, operators, averaging and the ParamSpaceSGD step loop.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates src/optimization/rules.jl (DoG, DoWG), Optimisers.jl Descent / Adam
(third-party), src/optimization/clip_scale.jl,
src/optimization/proximal_location_scale_entropy.jl, src/optimization/averaging.jl and
the init / step / output of src/algorithms/common.jl:40-120.
"""

from __future__ import annotations

import numpy as np

from .family import MvLocationScale


# ---- rules: state = init(x); state, dx' = apply(state, x, dx); x <- x - dx' ---------------
class Descent:
    """Optimisers.Descent(eta): dx' = eta * dx."""

    def __init__(self, eta=0.1):
        self.eta = eta

    def init(self, x):
        return None

    def apply(self, state, x, dx):
        return state, x.dtype.type(self.eta) * dx


class Adam:
    """Optimisers.Adam(eta=1e-3, beta=(0.9, 0.999), epsilon=1e-8):
    mt = b1 mt + (1-b1) dx; vt = b2 vt + (1-b2) dx^2;
    dx' = mt / (1 - b1^t) / (sqrt(vt / (1 - b2^t)) + eps) * eta."""

    def __init__(self, eta=1e-3, beta=(0.9, 0.999), epsilon=1e-8):
        self.eta, self.beta, self.epsilon = eta, beta, epsilon

    def init(self, x):
        return (np.zeros_like(x), np.zeros_like(x), (self.beta[0], self.beta[1]))

    def apply(self, state, x, dx):
        T = x.dtype.type
        mt, vt, bt = state
        b1, b2 = self.beta
        mt = T(b1) * mt + T(1 - b1) * dx
        vt = T(b2) * vt + T(1 - b2) * dx * dx
        dxp = mt / T(1 - bt[0]) / (np.sqrt(vt / T(1 - bt[1])) + T(self.epsilon)) * T(self.eta)
        return (mt, vt, (bt[0] * b1, bt[1] * b2)), dxp


class DoG:
    """src/optimization/rules.jl:48-64."""

    def __init__(self, alpha=1e-6):
        self.alpha = alpha

    def init(self, x):                                          # :52-54
        T = x.dtype.type
        return (x.copy(), T(0), T(self.alpha) * (1 + np.linalg.norm(x)))

    def apply(self, state, x, dx):                              # :56-64
        x0, v, r = state
        r = max(np.sqrt(np.sum((x - x0) ** 2)), r)
        v = v + np.sum(dx * dx)
        eta = r / np.sqrt(v)
        return (x0, v, r), dx * eta

    @staticmethod
    def stepsize(state):                                        # proximal...jl:36-39
        _, v, r = state
        return r / np.sqrt(v)


class DoWG(DoG):
    """src/optimization/rules.jl:17-34."""

    def apply(self, state, x, dx):
        x0, v, r = state
        r = max(np.sqrt(np.sum((x - x0) ** 2)), r)
        r2 = r * r
        v = v + r2 * np.sum(dx * dx)
        eta = r2 / np.sqrt(v)
        return (x0, v, r), dx * eta

    @staticmethod
    def stepsize(state):                                        # proximal...jl:41-44
        _, v, r = state
        return r * r / np.sqrt(v)


# ---- operators ----------------------------------------------------------------------------
class IdentityOperator:
    def apply(self, q_template, opt_rule, opt_state, params):   # src/AdvancedVI.jl:199
        return params


class ClipScale:
    """src/optimization/clip_scale.jl:8-29: diag(scale) <- max(diag(scale), eps)."""

    def __init__(self, epsilon=1e-5):
        self.epsilon = epsilon

    def apply(self, q_template: MvLocationScale, opt_rule, opt_state, params):
        q = q_template.restructure(params)
        eps = params.dtype.type(self.epsilon)
        if hasattr(q, "scale_factors"):                 # MvLocationScaleLowRank, clip_scale.jl:31-41
            q.scale_diag = np.maximum(q.scale_diag, eps)
            return q.destructure()
        if q.is_meanfield:
            q.scale = np.maximum(q.scale, eps)
        else:
            i = np.diag_indices(len(q))
            q.scale[i] = np.maximum(q.scale[i], eps)
        return q.destructure()


class ProximalLocationScaleEntropy:
    """src/optimization/proximal_location_scale_entropy.jl:46-61:
    L_ii <- L_ii + (sqrt(L_ii^2 + 4 gamma) - L_ii) / 2 with gamma from the rule (:34-44)."""

    def apply(self, q_template: MvLocationScale, opt_rule, opt_state, params):
        if isinstance(opt_rule, Descent):
            gamma = opt_rule.eta
        elif isinstance(opt_rule, (DoG, DoWG)):
            gamma = opt_rule.stepsize(opt_state)
        else:
            raise TypeError("ProximalLocationScaleEntropy does not support this rule")  # :29-33
        q = q_template.restructure(params)
        gamma = params.dtype.type(gamma)
        if q.is_meanfield:
            s = q.scale
            q.scale = s + (np.sqrt(s * s + 4 * gamma) - s) / 2
        else:
            i = np.diag_indices(len(q))
            s = q.scale[i]
            q.scale[i] = s + (np.sqrt(s * s + 4 * gamma) - s) / 2
        return q.destructure()


# ---- averaging (src/optimization/averaging.jl) --------------------------------------------
class NoAveraging:
    def init(self, x):
        return x

    def apply(self, state, x):
        return x

    def value(self, state):
        return state


class PolynomialAveraging:
    def __init__(self, eta=8):
        self.eta = eta

    def init(self, x):                      # :42
        return (x.copy(), 1)

    def apply(self, state, x):              # :44-51
        T = x.dtype.type
        eta = T(self.eta)
        x_bar, t = state
        w = (eta + 1) / (T(t) + eta)
        x_bar = (1 - w) * x_bar + w * x
        return (x_bar, t + 1)

    def value(self, state):                 # :53
        return state[0]


# ---- ParamSpaceSGD init / step / output (src/algorithms/common.jl:40-120) ------------------
class SGDState:
    def __init__(self, params, opt_st, avg_st, iteration=0):
        self.params, self.opt_st, self.avg_st, self.iteration = params, opt_st, avg_st, iteration


def sgd_init(q_init: MvLocationScale, rule, averager):
    params = q_init.destructure()                      # :47
    return SGDState(params, rule.init(params), averager.init(params))


def sgd_step(state: SGDState, q_template, grad_fn, rule, operator, averager):
    """One iteration of `step` (common.jl:69-104).  grad_fn(params, iteration) ->
    (value, grad, info_dict).  Raises on a non-finite objective value (:83-89)."""
    state.iteration += 1
    value, grad, info = grad_fn(state.params, state.iteration)
    if not np.isfinite(value):
        raise RuntimeError(f"The objective value is {value}. This indicates that the "
                           "optimization run diverged.")
    state.opt_st, dxp = rule.apply(state.opt_st, state.params, grad)   # Optimisers.update! :92
    params = state.params - dxp
    params = operator.apply(q_template, rule, state.opt_st, params)     # :93
    state.avg_st = averager.apply(state.avg_st, params)                 # :94
    state.params = params
    return info


def sgd_output(state: SGDState, q_template, averager):                   # :63-67
    return q_template.restructure(averager.value(state.avg_st))
