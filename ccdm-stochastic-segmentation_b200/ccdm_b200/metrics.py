"""LIDC sample-diversity metrics on the device.

Mirrors the three functions the reference's LIDC evaluators call on every batch
(/root/reference/ddpm/utils.py:129-174, duplicated in evaluation/evaluate_lidc_uncertainty.py:27-73 and called at :108-123):

    batched_distance(x, y)                                       -> [B, N, M]   1 - mean_{c >= 1} IoU_c
    calc_batched_generalised_energy_distance(s0, s1, K)          -> (GED [B], diversity_0 [B], diversity_1 [B])
    batched_hungarian_matching(s0, s1, K)                        -> list of B Hungarian-matched mean IoUs

The reference one-hot encodes both sets on the host and reduces a [B, N, M, H*W, K] boolean broadcast with numpy
(after a device-to-host copy of all samples).  Here the label maps stay on the GPU as uint8, one kernel counts
intersections and unions per (image, sample pair, class) in integers (`ccdm_pairwise_distance`), and only the [B, N, M]
double matrices travel; the assignment problem itself (N, M <= ~100) is scipy's `linear_sum_assignment` on the host, like
the reference.  Inputs: integer label maps [B, N, H, W] / [B, M, H, W] (any integer dtype, values < K), CUDA tensors.
"""
import ctypes

import numpy as np
import torch

from . import _lib

__all__ = ["batched_distance", "calc_batched_generalised_energy_distance", "batched_hungarian_matching"]


def _labels_u8(t: torch.Tensor) -> torch.Tensor:
    if t.device.type != "cuda":
        raise _lib.CcdmError("ccdm_b200.metrics works on CUDA tensors; there is no CPU path")
    if t.dim() < 3:
        raise ValueError("label maps must be [B, N, ...]")
    return t.reshape(t.shape[0], t.shape[1], -1).to(torch.uint8).contiguous()


def batched_distance(x: torch.Tensor, y: torch.Tensor, num_classes: int) -> torch.Tensor:
    """utils.py:136-142 on label maps: [B, N, ...] x [B, M, ...] -> float64 [B, N, M] (background class 0 excluded)."""
    L = _lib.lib()
    xs, ys = _labels_u8(x), _labels_u8(y)
    if xs.shape[0] != ys.shape[0] or xs.shape[2] != ys.shape[2]:
        raise ValueError(f"incompatible label sets {tuple(x.shape)} / {tuple(y.shape)}")
    B, N, n_pix = xs.shape
    M = ys.shape[1]
    out = torch.empty((B, N, M), dtype=torch.float64, device=xs.device)
    sp = _lib.stream_ptr(torch.cuda.current_stream(xs.device))
    _lib.check(L.ccdm_pairwise_distance(xs.data_ptr(), ys.data_ptr(), B, N, M, n_pix, int(num_classes), out.data_ptr(), sp),
               "pairwise_distance")
    return out


def calc_batched_generalised_energy_distance(samples_dist_0: torch.Tensor, samples_dist_1: torch.Tensor, num_classes: int):
    """utils.py:145-158 -> (2 * cross - diversity_0 - diversity_1, diversity_0, diversity_1), float64 [B] each (numpy arrays,
    like the reference returns)."""
    cross = batched_distance(samples_dist_0, samples_dist_1, num_classes).mean(dim=(1, 2))
    d0 = batched_distance(samples_dist_0, samples_dist_0, num_classes).mean(dim=(1, 2))
    d1 = batched_distance(samples_dist_1, samples_dist_1, num_classes).mean(dim=(1, 2))
    return (2 * cross - d0 - d1).cpu().numpy(), d0.cpu().numpy(), d1.cpu().numpy()


def batched_hungarian_matching(samples_dist_0: torch.Tensor, samples_dist_1: torch.Tensor, num_classes: int):
    """utils.py:161-174: per image, the mean IoU of the optimal one-to-one matching between the two sets."""
    from scipy.optimize import linear_sum_assignment
    cost = batched_distance(samples_dist_0, samples_dist_1, num_classes).cpu().numpy()
    return [float((1 - cost[i])[linear_sum_assignment(cost[i])].mean()) for i in range(cost.shape[0])]
