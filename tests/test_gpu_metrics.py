"""the LIDC diversity metrics on the device against a numpy restatement of the reference's functions
(/root/reference/ddpm/utils.py:129-174; restated because `np.bool` no longer exists in numpy 2 -- the arithmetic is the
reference's line for line).  Integer counting + double divisions: equal to numpy to 1e-12."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _iou(x, y, axis=-1):  # utils.py:129-132
    with np.errstate(invalid="ignore", divide="ignore"):
        iou_ = (x & y).sum(axis) / (x | y).sum(axis)
    iou_[np.isnan(iou_)] = 1.
    return iou_


def _batched_distance(x, y):  # utils.py:136-142
    per_class_iou = _iou(x[:, :, None], y[:, None, :], axis=-2)
    return 1 - per_class_iou[..., 1:].mean(-1)


def _onehot(s, K):
    return np.eye(K)[s.reshape(*s.shape[:2], -1)].astype(bool)


def _ged(s0, s1, K):  # utils.py:145-158
    a, b = _onehot(s0, K), _onehot(s1, K)
    cross = np.mean(_batched_distance(a, b), axis=(1, 2))
    d0 = np.mean(_batched_distance(a, a), axis=(1, 2))
    d1 = np.mean(_batched_distance(b, b), axis=(1, 2))
    return 2 * cross - d0 - d1, d0, d1


def _hungarian(s0, s1, K):  # utils.py:161-174
    from scipy.optimize import linear_sum_assignment
    cost = _batched_distance(_onehot(s0, K), _onehot(s1, K))
    return [(1 - cost[i])[linear_sum_assignment(cost[i])].mean() for i in range(cost.shape[0])]


@pytest.mark.parametrize("B,N,M,H,W,K", [(3, 4, 4, 32, 32, 2), (2, 6, 4, 17, 23, 2), (2, 3, 5, 16, 16, 20), (1, 16, 4, 128, 128, 2)])
def test_metrics_match_the_reference_functions(cuda_device, B, N, M, H, W, K):
    from ccdm_b200 import metrics
    rng = np.random.default_rng(5 + K + N)
    s0 = rng.integers(0, K, (B, N, H, W)).astype(np.int64)
    s1 = rng.integers(0, K, (B, M, H, W)).astype(np.int64)
    s0[0, 0] = 0          # an all-background sample: 0/0 -> IoU 1 against another empty one
    s1[0, 0] = 0
    if K > 2:
        s0[:, :, :4] = np.minimum(s0[:, :, :4], 3)  # some classes absent from parts of the maps
    t0, t1 = torch.as_tensor(s0).cuda(), torch.as_tensor(s1).cuda()
    d = metrics.batched_distance(t0, t1, K).cpu().numpy()
    np.testing.assert_allclose(d, _batched_distance(_onehot(s0, K), _onehot(s1, K)), rtol=0, atol=1e-12)
    ged, d0, d1 = metrics.calc_batched_generalised_energy_distance(t0, t1, K)
    rg, r0, r1 = _ged(s0, s1, K)
    np.testing.assert_allclose(ged, rg, rtol=0, atol=1e-12)
    np.testing.assert_allclose(d0, r0, rtol=0, atol=1e-12)
    np.testing.assert_allclose(d1, r1, rtol=0, atol=1e-12)
    np.testing.assert_allclose(metrics.batched_hungarian_matching(t0, t1, K), _hungarian(s0, s1, K), rtol=0, atol=1e-12)


def test_metrics_refuse_cpu_tensors(cuda_device):
    from ccdm_b200 import _lib, metrics
    with pytest.raises(_lib.CcdmError):
        metrics.batched_distance(torch.zeros(1, 2, 4, 4, dtype=torch.uint8), torch.zeros(1, 2, 4, 4, dtype=torch.uint8), 2)
