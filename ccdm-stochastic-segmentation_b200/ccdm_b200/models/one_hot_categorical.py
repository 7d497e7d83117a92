"""``OneHotCategoricalBCHW`` with the reference's surface (ddpm/models/one_hot_categorical.py:10-54).

Callers use it to draw x_T (``OneHotCategoricalBCHW(logits=zeros).sample()``,
eval_cdm.py:162-163, evaluate_lidc_uncertainty.py:100).  The reference subclasses
``torch.distributions.OneHotCategorical``, whose ``sample`` is
``torch.multinomial(probs_2d, 1, True)`` = ``argmax(p / E)`` with
``E = empty_like(p).exponential_(1)``.  This class draws the
same ``E`` from the same global generator in the same (pixel-major, class-minor)
order, so samples are bit-identical to the reference for a given seed and device,
without the distribution's argument validation (five host synchronisations per
construction on CUDA).
"""
from typing import Optional

import torch

__all__ = ["OneHotCategoricalBCHW"]


class OneHotCategoricalBCHW:
    def __init__(self, probs: Optional[torch.Tensor] = None, logits: Optional[torch.Tensor] = None, validate_args=None):
        if (probs is None) == (logits is None):
            raise ValueError("Either `probs` or `logits` must be specified, but not both.")
        if probs is not None and probs.ndim < 2:
            raise ValueError("`probs.ndim` should be at least 2")
        if logits is not None and logits.ndim < 2:
            raise ValueError("`logits.ndim` should be at least 2")
        if probs is not None:
            p = self.channels_last(probs)
            self.probs = p / p.sum(-1, keepdim=True)  # Categorical.__init__
        else:
            lg = self.channels_last(logits)
            lg = lg - lg.logsumexp(dim=-1, keepdim=True)
            self.probs = torch.softmax(lg, dim=-1)
        self._num_events = self.probs.shape[-1]

    @staticmethod
    def channels_last(arr: torch.Tensor) -> torch.Tensor:
        return arr.permute((0,) + tuple(range(2, arr.ndim)) + (1,))

    @staticmethod
    def channels_second(arr: torch.Tensor) -> torch.Tensor:
        return arr.permute((0, arr.ndim - 1) + tuple(range(1, arr.ndim - 1)))

    def sample(self, sample_shape=torch.Size()):
        if len(sample_shape):
            raise NotImplementedError("sample_shape is not used on the hot path")
        p2 = self.probs.reshape(-1, self._num_events)
        e = torch.empty_like(p2).exponential_(1)
        idx = (p2 / e).argmax(dim=-1)
        res = torch.nn.functional.one_hot(idx.reshape(self.probs.shape[:-1]), self._num_events).to(self.probs)
        return self.channels_second(res)

    def max_prob_sample(self):
        res = torch.nn.functional.one_hot(self.probs.argmax(dim=-1), self._num_events)
        return self.channels_second(res)

    def prob_sample(self):
        return self.channels_second(self.probs)
