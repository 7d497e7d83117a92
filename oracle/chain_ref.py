"""CPU restatement of the T-step reverse chain (test infrastructure, see package docstring).

Restates /root/reference/ddpm/models/diffusion_denoising.py:164-215
(``DenoisingModel.forward_denoising``) on top of ``unet_ref.unet_forward`` and the
C posterior/draw routines, with the noise of every categorical draw made explicit:
the reference's ``OneHotCategoricalBCHW(probs).sample()`` is
``argmax(p / E)``, ``E = torch.empty(B*H*W, K).exponential_(1)``
(one_hot_categorical.py:30-32 -> torch.multinomial).
"""
import numpy as np
import torch

from . import cdm
from .unet_ref import unet_forward


def torch_noise(shape):
    """The draw the reference consumes from torch's global CPU generator."""
    return torch.empty(shape, dtype=torch.float32).exponential_(1).numpy()


def labels_to_onehot(labels, K):
    """uint8 [B,H,W] -> fp32 one-hot [B,K,H,W]."""
    lab = torch.as_tensor(np.asarray(labels)).long()
    return torch.nn.functional.one_hot(lab, K).permute(0, 3, 1, 2).float()


@torch.no_grad()
def reverse_chain(sd, labels_T, image, feature_condition, alphas, cumalphas, time_steps, init_t=None,
                  step_T_sample="majority", noise_fn=torch_noise, head_channels=32, num_heads=1,
                  feature_condition_idx=None, K=None, record=None, theta_fn=None):
    """Run the chain from integer labels ``labels_T`` (uint8 [B,H,W]).

    Returns (labels [B,H,W] uint8, probs [B,H,W,K] fp32 normalised posterior of the
    last step).  The reference returns one-hot int64 of ``labels`` for
    ``majority``/None and ``probs`` (as BCHW) for ``confidence`` (:208-212).
    ``record``: optional list receiving a dict per step.
    ``theta_fn(step, t, labels)``: optional override of the UNet call (teacher
    forcing with precomputed x0 predictions).
    """
    labels = np.ascontiguousarray(labels_T, np.uint8)
    B, H, W = labels.shape
    probs = None
    ts = cdm.t_values(time_steps, init_t)
    for i, t in enumerate(ts):
        if theta_fn is None:
            x = labels_to_onehot(labels, K)
            tt = torch.full((B,), float(t))
            theta = unet_forward(sd, x, image, feature_condition, tt, head_channels, num_heads, feature_condition_idx)
            theta = theta.permute(0, 2, 3, 1).contiguous().numpy()
        else:
            theta = np.ascontiguousarray(theta_fn(i, t, labels), np.float32)
        a_t, ca_tm1 = cdm.step_scalars(alphas, cumalphas, t)
        post = cdm.posterior_closed(labels, theta, a_t, ca_tm1)
        if t > 1:
            noise = noise_fn((B * H * W, theta.shape[-1]))
            new_labels, probs = cdm.draw(post, noise, 0)
        else:
            noise = None
            new_labels, probs = cdm.draw(post, None, 2 if step_T_sample == "confidence" else 1)
        if record is not None:
            record.append(dict(t=t, labels_in=labels.copy(), theta=theta, posterior=post, noise=noise,
                               labels_out=new_labels.copy(), probs=probs))
        labels = new_labels
    return labels, probs
