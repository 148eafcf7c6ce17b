"""Beta-Bernoulli posterior of a motif and the scores derived from it (host side, float64).

Mirror of nanomotif/model.py:11-126 and of the two scoring formulas that consume it
(predictive_evaluation_score, nanomotif/find_motifs_bin.py:1360-1379; MotifSearcher.
_priority_function, :901-924).  The counts come from the GPU as exact integers; these O(#motifs)
scalar formulas stay in float64 on the host, vectorised over motif batches.
"""
from __future__ import annotations

import numpy as np
from scipy.special import psi

DEFAULT_PRIOR_ALPHA = 5  # nanomotif/model.py:8-9
DEFAULT_PRIOR_BETA = 5


class BetaBernoulliModel:
    """Same attributes (`_alpha`, `_beta`, `_alpha_prior`, `_beta_prior`) and methods as the reference."""

    def __init__(self, alpha=DEFAULT_PRIOR_ALPHA, beta=DEFAULT_PRIOR_BETA):
        self._alpha = alpha
        self._beta = beta
        self._alpha_prior = alpha
        self._beta_prior = beta

    def __getstate__(self):
        return dict(_alpha=self._alpha, _beta=self._beta, _alpha_prior=self._alpha_prior, _beta_prior=self._beta_prior)

    def __setstate__(self, state):
        for k in ("_alpha", "_beta", "_alpha_prior", "_beta_prior"):
            setattr(self, k, state[k])

    def get_raw_counts(self):
        return self._alpha - self._alpha_prior, self._beta - self._beta_prior

    def update(self, n_positives, n_negatives):
        self._alpha += n_positives
        self._beta += n_negatives

    def reset(self):
        self._alpha, self._beta = self._alpha_prior, self._beta_prior

    def mean(self):
        return self._alpha / (self._alpha + self._beta)

    def variance(self):
        a, b = self._alpha, self._beta
        return (a * b) / ((a + b) ** 2 * (a + b + 1))

    def standard_deviation(self):
        return np.sqrt(self.variance())

    def posterior_predictive(self, n_positives, n_negatives):
        if n_positives + n_negatives == 0:
            return 0.0
        total = psi(self._alpha + self._beta)
        return n_positives * (psi(self._alpha) - total) + n_negatives * (psi(self._beta) - total)

    def posterior_predictive_per_obs(self, n_positives, n_negatives):
        n_new = n_positives + n_negatives
        if n_new == 0:
            return 0.0
        return self.posterior_predictive(n_positives, n_negatives) / n_new

    def __repr__(self):
        return f"BetaBernoulliModel(alpha={self._alpha}, beta={self._beta})"

    __str__ = __repr__


def predictive_evaluation_score(next_model, current_model) -> float:
    """score(next | current), find_motifs_bin.py:1360-1379."""
    a_n, b_n = next_model._alpha, next_model._beta
    extra_pos, extra_neg = current_model._alpha - a_n, current_model._beta - b_n
    ppc_next = next_model.posterior_predictive_per_obs(a_n, b_n)
    ppc_extra = next_model.posterior_predictive_per_obs(extra_pos, extra_neg)
    return (next_model.mean() / current_model.mean()) * (ppc_next - ppc_extra)


def predictive_evaluation_scores(alpha_next, beta_next, alpha_cur, beta_cur) -> np.ndarray:
    """Vectorised predictive_evaluation_score over arrays of posterior parameters."""
    a_n = np.asarray(alpha_next, dtype=np.float64)
    b_n = np.asarray(beta_next, dtype=np.float64)
    a_c = np.asarray(alpha_cur, dtype=np.float64)
    b_c = np.asarray(beta_cur, dtype=np.float64)
    total = psi(a_n + b_n)
    e1, e0 = psi(a_n) - total, psi(b_n) - total

    def ppc(x, y):
        n = x + y
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(n == 0, 0.0, (x * e1 + y * e0) / np.where(n == 0, 1.0, n))

    ratio = (a_n / (a_n + b_n)) / (a_c / (a_c + b_c))
    return ratio * (ppc(a_n, b_n) - ppc(a_c - a_n, b_c - b_n))


def priority(next_model, root_model) -> float:
    """(1 - a_n/a_root) * (b_n/b_root), find_motifs_bin.py:915-923."""
    try:
        d_alpha = 1 - (next_model._alpha / root_model._alpha)
    except ZeroDivisionError:
        d_alpha = 1
    try:
        d_beta = next_model._beta / root_model._beta
    except ZeroDivisionError:
        d_beta = 1
    return d_alpha * d_beta
