"""``DiffusionModel`` / ``DenoisingModel`` with the reference's surface, backed by the CUDA engine.

Reference: /root/reference/ddpm/models/diffusion_denoising.py.  ``DenoisingModel.forward``
keeps the reference's dispatch (:144-159) and return convention (:206-214); the T-step
loop of ``forward_denoising`` (:164-215) -- UNet, ``theta_post_prob``, clamp, categorical
draw -- runs as one captured CUDA graph per step in ``libccdm_b200.so``.

Knobs that do not exist in the reference (plain attributes; the defaults are the drop-in behaviour):
  ``precision``  'exact' (default): fp32-grade arithmetic ON THE TENSOR CORES -- every activation and weight is carried as
                 fp16 hi + lo (22 significand bits), every product is three tcgen05 MMAs with fp32 accumulation,
                 GroupNorm / SiLU / softmax / posterior in fp32.  Held to the same parity tolerances as 'fp32'.
                 'fp32': fp32 storage, FFMA kernels (the in-library yardstick; ~4x slower).
                 'bf16': bf16 storage, single bf16 MMAs (~2x faster than 'exact'; x0 differs by up to ~1e-2, so a
                 free-running chain decorrelates from the reference's after some tens of steps -- a valid sampler
                 of the same model, not a reproduction of the reference's samples).
  ``noise``      'torch' (default: consume the device's global generator exactly like the reference's
                 ``torch.multinomial``, one ``exponential_`` draw of [B*H*W, K] per step) | 'philox' (in-kernel counter RNG
                 keyed by (seed, global sample index, step, pixel): sharding-invariant, no noise traffic, the whole
                 chain replays as CUDA graphs -- ~1.0x (LIDC) .. 1.3x (Cityscapes) faster; what the benchmark uses)
  ``seed``, ``sample_offset``  Philox key / global index of local sample 0.
  ``tile_batch``  0 (default): every conv picks its tile height from the batch it is given (fastest).  N > 0: as if the batch
                 were N -- the tiling, and with it the summation order of the GroupNorm statistics, no longer depends on how
                 the samples are batched, so in 'exact' mode a sample's result is bit-identical on 1 GPU or 8, in a batch of
                 64 or of 2 (``sample_sharded`` sets 64; costs throughput only when the real batch is much smaller).
"""
import logging
import math
from typing import Optional, Tuple, Union

import numpy as np
import torch
from torch import Tensor, nn

from .. import _lib

LOGGER = logging.getLogger(__name__)

__all__ = ["DiffusionModel", "DenoisingModel", "linear_schedule", "cosine_schedule", "reverse_t_values"]


def linear_schedule(time_steps: int, start=1e-2, end=0.2) -> Tuple[Tensor, Tensor, Tensor]:
    """:18-22 -- betas linear in [start, end]; cumalphas = cumprod(1 - betas)."""
    betas = torch.linspace(start, end, time_steps)
    alphas = 1 - betas
    return betas, alphas, torch.cumprod(alphas, dim=0)


def cosine_schedule(time_steps: int, s: float = 8e-3) -> Tuple[Tensor, Tensor, Tensor]:
    """:25-39 -- including its quirks: the ``s`` argument is ignored (0.008 always);
    ``cumalphas`` comes from an fp32 tensor expression that is NOT normalised by f(0)
    while ``betas`` come from float64 ratios clipped at 0.999, so
    ``cumalphas != cumprod(alphas)``.  Same torch ops as the reference -> same bits."""
    s = 0.008
    steps = torch.arange(0, time_steps)
    cumalphas = torch.cos(((steps / time_steps + s) / (1 + s)) * (math.pi / 2)) ** 2

    def f(u):
        return math.cos((u + s) / (1.0 + s) * math.pi / 2) ** 2

    betas = torch.tensor([min(1 - f((i + 1) / time_steps) / f(i / time_steps), 0.999) for i in range(time_steps)])
    return betas, 1 - betas, cumalphas


def reverse_t_values(time_steps: int, init_t: Optional[int] = None):
    """The t grid of the reverse chain (:167-187): all of T..1, or, for ``init_t = 10000 + K'``,
    K' values ``round(linspace(T, 1, K'))`` (python round = half-to-even)."""
    if init_t is None:
        init_t = time_steps
    if init_t > 10000:
        k = init_t % 10000
        assert 0 < k <= time_steps
        if k == time_steps:
            return list(range(k, 0, -1))
        LOGGER.warning(f"Override default {time_steps} time steps with {k}.")
        return [round(v) for v in np.linspace(time_steps, 1, k)]
    return list(range(init_t, 0, -1))


class DiffusionModel(nn.Module):
    """Schedule buffers + class count (:42-66).  The training-only transition kernels
    (``q_xt_given_x0``, ``theta_post``; trainer.py:257,263) are outside the sampler's scope;
    ``theta_post_prob`` is provided on device tensors through the C ABI."""
    betas: Tensor
    alphas: Tensor
    cumalphas: Tensor

    def __init__(self, schedule: str, time_steps: int, num_classes: int, schedule_params=None):
        super().__init__()
        fn = {"linear": linear_schedule, "cosine": cosine_schedule}[schedule]
        betas, alphas, cumalphas = fn(time_steps, **schedule_params) if schedule_params is not None else fn(time_steps)
        self.register_buffer("betas", betas)
        self.register_buffer("alphas", alphas)
        self.register_buffer("cumalphas", cumalphas)
        self.num_classes = num_classes

    @property
    def time_steps(self):
        return len(self.betas)

    def step_scalars(self, t: int):
        """(alpha_t, cumalpha_{t-1}) as the posterior uses them (:107-113)."""
        t0 = t - 1
        if t0 == 0:
            return 0.0, 1.0
        return float(self.alphas[t0]), float(self.cumalphas[t0 - 1])

    @torch.no_grad()
    def theta_post_prob(self, xt: Tensor, theta_x0: Tensor, t: Tensor) -> Tensor:
        """:99-128 on CUDA tensors: xt one-hot [B,K,H,W], theta_x0 [B,K,H,W], t int [B] (all equal)."""
        _lib.require_device()
        L = _lib.lib()
        B, K, H, W = theta_x0.shape
        tv = t.reshape(-1).tolist()
        if any(v != tv[0] for v in tv):
            raise NotImplementedError("theta_post_prob with per-sample t is only needed by the training loss")
        a, c = self.step_scalars(int(tv[0]))
        theta = theta_x0.float().permute(0, 2, 3, 1).contiguous()
        labels = xt.argmax(dim=1).to(torch.uint8).contiguous()
        out = torch.empty_like(theta)
        sp = _lib.stream_ptr(torch.cuda.current_stream(theta.device))
        _lib.check(L.ccdm_posterior_draw(theta.data_ptr(), labels.data_ptr(), H * W, B, K, a, c, _lib.DRAW_POSTERIOR,
                                         _lib.NOISE_PHILOX, None, 0, 0, 0, None, out.data_ptr(), None, sp), "posterior_draw")
        return out.permute(0, 3, 1, 2)


class DenoisingModel(nn.Module):
    def __init__(self, diffusion: DiffusionModel, unet: nn.Module, dataset_file: str, step_T_sample: str = "majority"):
        super().__init__()
        self.diffusion = diffusion
        self.unet = unet
        self.dataset_file = dataset_file
        self.step_T_sample = step_T_sample
        self.precision = "exact"
        self.noise = "torch"
        self.seed = 0
        self.sample_offset = 0
        self.tile_batch = 0
        self._sched_host = None

    @property
    def time_steps(self):
        return self.diffusion.time_steps

    def forward(self, x: Tensor, condition: Tensor, feature_condition: Tensor = None, t: Optional[Tensor] = None,
                label_ref_logits: Optional[Tensor] = None, validation: bool = False) -> Union[Tensor, dict]:
        if self.training:  # :147-152
            if not isinstance(t, Tensor):
                raise ValueError("'t' needs to be a Tensor at training time")
            if not isinstance(x, Tensor):
                raise ValueError("'x' needs to be a Tensor at training time")
            return self.forward_step(x, condition, feature_condition, t)
        if validation:  # :154-155
            return self.forward_step(x, condition, feature_condition, t)
        if t is None:  # :156-157
            return self.forward_denoising(x, condition, feature_condition, label_ref_logits=label_ref_logits)
        return self.forward_denoising(x, condition, feature_condition, int(t.item()), label_ref_logits)  # :159

    # Philox draw index of the chain's initial state; the reverse steps use 0, 1, 2, ...
    X_T_DRAW = 0xFFFFFFFF

    @torch.no_grad()
    def draw_x_T(self, batch: int, height: int, width: int, device=None) -> Tensor:
        """x_T drawn ON THE DEVICE: a uniform label per pixel, uint8 ``[batch, height, width]``, accepted as ``x`` by
        ``forward``.  The reference builds it on the host as
        ``OneHotCategoricalBCHW(logits=zeros(B, K, H, W)).sample()`` -- an exponential race over equal probabilities
        (evaluate_lidc_uncertainty.py:100, eval_cdm.py:162-163) -- and copies B*K*H*W floats to the GPU.  Here the same
        race runs in one kernel on Philox noise keyed by (``seed``, ``sample_offset`` + b, ``X_T_DRAW``, pixel): one
        byte per pixel, no host copy, and independent of how the batch is sharded over ranks."""
        L = _lib.lib()
        dev = torch.device(device) if device is not None else next(self.unet.parameters()).device
        if dev.type != "cuda":
            raise _lib.CcdmError("draw_x_T needs the model on a CUDA device")
        K = self.diffusion.num_classes
        labels = torch.empty((batch, height, width), dtype=torch.uint8, device=dev)
        sp = _lib.stream_ptr(torch.cuda.current_stream(dev))
        _lib.check(L.ccdm_uniform_labels(int(self.seed), self.X_T_DRAW, int(self.sample_offset), batch, height * width, K,
                                         labels.data_ptr(), sp), "uniform_labels")
        return labels

    @torch.no_grad()
    def sample_many(self, condition: Tensor, n_samples: int, feature_condition: Optional[Tensor] = None,
                    init_t: Optional[int] = None) -> dict:
        """N stochastic segmentations per image in ONE batched chain, with everything the evaluators do around the sampler
        call moved onto the device:

        * x_T is drawn on the device (``draw_x_T``) instead of on the host + a copy of [B*N, K, H, W] floats
          (evaluate_lidc_uncertainty.py:100);
        * the image (and the feature map) are NOT replicated: the kernels read entry ``sample // N`` -- the reference does
          ``image.repeat_interleave(N, dim=0)`` (evaluate_lidc_uncertainty.py:96);
        * the vote over the N samples runs as one kernel on the uint8 label maps: ``mean_onehot`` [B, K, H, W] is the mean
          of the N one-hot maps (what ``predict_multiple`` accumulates, eval_cdm.py:176-193, and what
          ``prediction.reshape(B, N, ...)`` feeds the metrics, evaluate_lidc_uncertainty.py:103), ``majority`` its argmax.

        ``condition`` [B, C_img, H, W], ``feature_condition`` [B, F, H/8, W/8] or None.  Noise is the in-kernel Philox
        stream keyed by (``seed``, ``sample_offset`` + global sample index), so the result does not depend on how the
        B*N samples are batched or sharded.  Returns ``labels`` uint8 [B, N, H, W], ``mean_onehot`` fp32 [B, K, H, W] (in
        ``confidence`` mode: the mean of the N normalised probability maps) and ``majority`` uint8 [B, H, W]."""
        L = _lib.lib()
        if n_samples < 1:
            raise ValueError("n_samples must be >= 1")
        B_img, _, H, W = condition.shape
        B, K = B_img * n_samples, self.diffusion.num_classes
        dev = next(self.unet.parameters()).device
        t_values = reverse_t_values(self.time_steps, init_t)
        alphas, cumalphas = self._schedule_host()
        confidence = self.step_T_sample == "confidence"
        base = int(self.sample_offset)  # global index of this call's first SAMPLE (image-major, sample-minor)
        x = self.draw_x_T(B, H, W, dev)
        engine = self.unet.engine(self.precision)
        engine.tile_batch = int(self.tile_batch)
        labels, probs = engine.run_chain(
            x, condition, feature_condition, t_values, alphas, cumalphas, _lib.DRAW_CONFIDENCE if confidence else _lib.DRAW_MAJORITY,
            noise="philox", seed=self.seed, sample0=base, img_rep=n_samples)
        freq = torch.empty((B_img, K, H, W), dtype=torch.float32, device=dev)
        majority = torch.empty((B_img, H, W), dtype=torch.uint8, device=dev)
        sp = _lib.stream_ptr(torch.cuda.current_stream(dev))
        _lib.check(L.ccdm_vote(labels.data_ptr(), B_img, n_samples, H * W, K, freq.data_ptr(), majority.data_ptr(), sp), "vote")
        if confidence and probs is not None:
            freq = probs.view(B_img, n_samples, H, W, K).mean(dim=1).permute(0, 3, 1, 2)
            majority = freq.argmax(dim=1).to(torch.uint8)
        return {"labels": labels.view(B_img, n_samples, H, W), "mean_onehot": freq, "majority": majority}

    def forward_step(self, x: Tensor, condition: Tensor, feature_condition: Tensor, t: Tensor) -> dict:
        """One denoiser evaluation (:161-162).  Inference only: no autograd graph is built."""
        self.unet.precision = self.precision
        return self.unet(x, condition, feature_condition=feature_condition, timesteps=t)

    def _schedule_host(self):
        d = self.diffusion
        key = (d.alphas.data_ptr(), d.alphas._version, d.cumalphas.data_ptr(), d.cumalphas._version)
        if self._sched_host is None or self._sched_host[0] != key:
            self._sched_host = (key, d.alphas.detach().float().cpu().tolist(), d.cumalphas.detach().float().cpu().tolist())
        return self._sched_host[1], self._sched_host[2]

    @torch.no_grad()
    def forward_denoising(self, x: Optional[Tensor], condition: Tensor, feature_condition: Tensor,
                          init_t: Optional[int] = None, label_ref_logits: Optional[Tensor] = None) -> dict:
        if label_ref_logits is not None:
            # The reference's guidance branch (:172-174,199-202) reads attributes that are never
            # defined (guidance_scale_weights, guidance_fn, ...): it cannot run there either.
            raise NotImplementedError("label_ref_logits guidance is not implemented (undefined in the reference as well)")
        t_values = reverse_t_values(self.time_steps, init_t)
        alphas, cumalphas = self._schedule_host()
        confidence = self.step_T_sample == "confidence"
        if not (self.step_T_sample is None or self.step_T_sample in ("majority", "confidence")):
            raise ValueError(f"step_T_sample={self.step_T_sample!r}")
        engine = self.unet.engine(self.precision)
        engine.tile_batch = int(self.tile_batch)
        labels, probs = engine.run_chain(x, condition, feature_condition, t_values, alphas, cumalphas,
                                         _lib.DRAW_CONFIDENCE if confidence else _lib.DRAW_MAJORITY,
                                         noise=self.noise, seed=self.seed, sample0=self.sample_offset)
        K = self.diffusion.num_classes
        if t_values[-1] != 1:
            # chain stopped above t=1 (init_t given as a plain int < ... never on shipped paths):
            # the reference would return the last *sampled* one-hot x_t as fp32
            out = self._onehot(labels, K, torch.float32)
        elif confidence:
            out = probs.permute(0, 3, 1, 2)      # fp32 normalised probabilities (:211-212)
        else:
            out = self._onehot(labels, K, torch.int64)  # int64 one-hot (:208-210)
        return {"diffusion_out": out}

    @staticmethod
    def _onehot(labels: Tensor, K: int, dtype) -> Tensor:
        L = _lib.lib()
        B, H, W = labels.shape
        oh = torch.empty((B, H, W, K), dtype=torch.int64, device=labels.device)
        sp = _lib.stream_ptr(torch.cuda.current_stream(labels.device))
        _lib.check(L.ccdm_labels_to_onehot_i64(labels.data_ptr(), B * H * W, K, oh.data_ptr(), sp), "labels_to_onehot")
        if dtype != torch.int64:
            oh = oh.to(dtype)
        return oh.permute(0, 3, 1, 2)  # same NHWC-strided BCHW view the reference returns
