"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run here (the build container), never on the GPU box:

    python tests/golden/make_golden.py

Imports /root/reference/ddpm/models the way SURVEY.md 8c describes
(``sys.path.insert(0, '/root/reference/ddpm'); import models`` -- the top-level
``ddpm`` package needs ignite, the ``models`` sub-package only torch/numpy), fills
the weights with the repo's deterministic synthetic recipe
(``ccdm_b200.synthetic``; a fresh reference model is degenerate, SURVEY.md
section 7 item 9) and records what the reference computes at the hot-path
boundary.  Everything is CPU fp32, torch's global generator seeded explicitly.
The reference has no tests / golden vectors of its own;
these files are what pins the oracle and the CUDA path to it.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference/ddpm")
sys.path.insert(0, os.path.join(ROOT, "ccdm-stochastic-segmentation_b200"))

import models  # noqa: E402  (the reference)
from models.one_hot_categorical import OneHotCategoricalBCHW  # noqa: E402
from models.diffusion_denoising import DiffusionModel  # noqa: E402

from ccdm_b200.synthetic import fill_synthetic_, synthetic_inputs  # noqa: E402

UNET_PARAMS = dict(base_channels=32, channel_mult=None, attention_resolutions=[32, 16, 8], num_heads=1,
                   num_head_channels=32, softmax_output=True)
DINO = dict(type="dino", model="dino_vits8", channels=384, conditioning="concat_pixels_concat_features",
            output_stride=8, scale="single", train=False, source_layer=11, target_layer=10)


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def onehot(labels, K):
    return torch.nn.functional.one_hot(labels.long(), K).permute(0, 3, 1, 2).float()


def gen_schedules():
    out = {}
    for T in (100, 250, 1000):
        d = DiffusionModel("cosine", T, 2, {"s": 0.008})
        out[f"cosine{T}_betas"] = d.betas.numpy()
        out[f"cosine{T}_alphas"] = d.alphas.numpy()
        out[f"cosine{T}_cumalphas"] = d.cumalphas.numpy()
    d = DiffusionModel("linear", 250, 2, None)
    out["linear250_betas"], out["linear250_alphas"], out["linear250_cumalphas"] = (
        d.betas.numpy(), d.alphas.numpy(), d.cumalphas.numpy())
    save("schedules.npz", **out)


def gen_t_values():
    # diffusion_denoising.py:178-187 through the public forward(): record the t's the UNet sees.
    class Spy(torch.nn.Module):
        def __init__(self, K):
            super().__init__()
            self.K, self.seen = K, []

        def forward(self, x, cond, feat, t):
            self.seen.append(int(t[0].item()))
            return {"diffusion_out": torch.full_like(x, 1.0 / self.K)}

    out = {}
    for T, req in [(250, None), (250, 10007), (250, 10025), (250, 10250), (250, 10100), (1000, 10050), (100, 10010),
                   (250, 10002), (250, 10001), (250, 40)]:
        spy = Spy(2)
        dm = models.DenoisingModel(DiffusionModel("cosine", T, 2, {"s": 0.008}), spy, "datasets.lidc", "majority").eval()
        x = onehot(torch.zeros(1, 4, 4, dtype=torch.uint8), 2)
        torch.manual_seed(0)
        dm(x, torch.zeros(1, 1, 4, 4), None, None if req is None else torch.as_tensor(req))
        out[f"T{T}_req{req}"] = np.asarray(spy.seen, np.int32)
    save("t_values.npz", **out)


def gen_posterior():
    out = {}
    for K in (2, 20):
        for T in (250, 1000):
            d = DiffusionModel("cosine", T, K, {"s": 0.008})
            g = torch.Generator().manual_seed(100 + K + T)
            B, H, W = 6, 8, 8
            theta = torch.softmax(3.0 * torch.randn((B, K, H, W), generator=g), dim=1)
            # some hard (near one-hot) and some exactly uniform predictions
            theta[0] = torch.softmax(40.0 * torch.randn((K, H, W), generator=g), dim=0)
            theta[1] = 1.0 / K
            labels = torch.randint(0, K, (B, H, W), generator=g, dtype=torch.uint8)
            t = torch.tensor([1, 2, 3, T // 2, T - 1, T])
            post = d.theta_post_prob(onehot(labels, K), theta, t)
            tag = f"K{K}_T{T}"
            out[tag + "_theta"] = theta.permute(0, 2, 3, 1).contiguous().numpy()
            out[tag + "_labels"] = labels.numpy()
            out[tag + "_t"] = t.numpy().astype(np.int32)
            out[tag + "_post"] = post.permute(0, 2, 3, 1).contiguous().numpy()
    save("posterior.npz", **out)


def gen_draw():
    out = {}
    for K in (2, 20):
        g = torch.Generator().manual_seed(7 + K)
        B, H, W = 3, 16, 16
        probs = torch.softmax(2.0 * torch.randn((B, K, H, W), generator=g), dim=1)
        probs[0, 0] = 0.0  # exercises the 1e-12 clamp (diffusion_denoising.py:204)
        probs = torch.clamp(probs, min=1e-12)
        torch.manual_seed(1000 + K)
        sample = OneHotCategoricalBCHW(probs=probs).sample()
        torch.manual_seed(1000 + K)
        noise = torch.empty(B * H * W, K).exponential_(1)
        dist = OneHotCategoricalBCHW(probs=probs)
        out[f"K{K}_probs_in"] = probs.permute(0, 2, 3, 1).contiguous().numpy()
        out[f"K{K}_noise"] = noise.numpy()
        out[f"K{K}_sample_labels"] = sample.argmax(1).numpy().astype(np.uint8)
        out[f"K{K}_majority_labels"] = dist.max_prob_sample().argmax(1).numpy().astype(np.uint8)
        out[f"K{K}_confidence"] = dist.prob_sample().permute(0, 2, 3, 1).contiguous().numpy()
        # x_T: OneHotCategoricalBCHW(logits=zeros).sample()  (eval_cdm.py:162-163)
        torch.manual_seed(2000 + K)
        xT = OneHotCategoricalBCHW(logits=torch.zeros(B, K, H, W)).sample()
        torch.manual_seed(2000 + K)
        noiseT = torch.empty(B * H * W, K).exponential_(1)
        out[f"K{K}_xT_noise"] = noiseT.numpy()
        out[f"K{K}_xT_labels"] = xT.argmax(1).numpy().astype(np.uint8)
    save("draw.npz", **out)


def build(T, C_img, H, W, K, step_T_sample, fce=None, channel_mult=None):
    p = dict(UNET_PARAMS)
    p["channel_mult"] = channel_mult
    torch.manual_seed(0)
    m = models.build_model(T, "cosine", {"s": 0.008}, [(C_img, H, W), (K, H, W)], (C_img, H, W), "unet_openai", p,
                           "datasets.lidc" if K == 2 else "datasets.cityscapes", step_T_sample, fce).eval()
    fill_synthetic_(m.unet, 0)
    return m


def noise_digest(seed, shape, n):
    torch.manual_seed(seed)
    h = hashlib.sha256()
    for _ in range(n):
        h.update(torch.empty(shape).exponential_(1).numpy().tobytes())
    return np.frombuffer(h.digest(), np.uint8).copy()


def gen_unet_and_chain(tag, T, B, C_img, H, W, K, fce, channel_mult, t_probe, steps):
    m = build(T, C_img, H, W, K, "majority", fce, channel_mult)
    image, feat, labels = synthetic_inputs(B, C_img, H, W, K, 384 if fce else 0)
    x = onehot(labels, K)
    out = dict(cfg=np.asarray([T, B, C_img, H, W, K, 1 if fce else 0, steps], np.int32),
               channel_mult=np.asarray(m.unet.channel_mult, np.float32))
    with torch.no_grad():
        for t in t_probe:
            tt = torch.full((B,), float(t))
            x0 = m.unet(x, image, feat, tt)["diffusion_out"]
            out[f"x0pred_t{t}"] = x0.permute(0, 2, 3, 1).contiguous().numpy()
        # full strided chain, both last-step modes (diffusion_denoising.py:206-212)
        for mode in ("majority", "confidence"):
            m.step_T_sample = mode
            torch.manual_seed(42)
            res = m(x, image, feat, t=torch.as_tensor(10000 + steps))["diffusion_out"]
            if mode == "majority":
                assert res.dtype == torch.int64
                out["chain_majority_labels"] = res.argmax(1).numpy().astype(np.uint8)
            else:
                assert res.dtype == torch.float32
                out["chain_confidence_probs"] = res.permute(0, 2, 3, 1).contiguous().numpy()
    out["chain_noise_sha256"] = noise_digest(42, (B * H * W, K), steps - 1)
    save(f"{tag}.npz", **out)


MANIFEST_CASES = {
    # tag: (C_img, H, W, K, DINO concat?, channel_mult, base_channels)
    "lidc128": (1, 128, 128, 2, False, None, 32),
    "lidc64": (1, 64, 64, 2, False, None, 32),
    "cs256x512": (3, 256, 512, 20, True, None, 32),
    "cs64x128": (3, 64, 128, 20, True, (1, 1, 2, 2, 4, 4), 32),
    "base64_256": (3, 256, 256, 5, False, None, 64),
}


def gen_manifest():
    """state_dict_manifest.json: key names, order and shapes of the reference ``DenoisingModel.state_dict()`` for five
    configurations -- what a strict ``load_state_dict`` of a reference checkpoint requires of ``ccdm_b200.models``."""
    import json
    man = {}
    for tag, (C, H, W, K, fce, mult, base) in MANIFEST_CASES.items():
        p = dict(UNET_PARAMS, channel_mult=mult, base_channels=base)
        m = models.build_model(250, "cosine", {"s": 0.008}, [(C, H, W), (K, H, W)], (C, H, W), "unet_openai", p, "d", "majority",
                               DINO if fce else None)
        man[tag] = dict(args=[C, H, W, K, fce, list(mult) if mult else None, base],
                        keys=[[k, list(v.shape)] for k, v in m.state_dict().items()])
    path = os.path.join(HERE, "state_dict_manifest.json")
    with open(path, "w") as fh:
        json.dump(man, fh)
    print(f"state_dict_manifest.json: {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    torch.set_num_threads(8)
    if sys.argv[1:] == ["manifest"]:
        gen_manifest()
        sys.exit(0)
    gen_manifest()
    gen_schedules()
    gen_t_values()
    gen_posterior()
    gen_draw()
    gen_unet_and_chain("lidc64", 250, 2, 1, 64, 64, 2, None, None, (250, 37, 1), 7)
    gen_unet_and_chain("lidc128", 250, 1, 1, 128, 128, 2, None, None, (250, 1), 5)
    gen_unet_and_chain("cs64x128", 250, 1, 3, 64, 128, 20, DINO, (1, 1, 2, 2, 4, 4), (250, 100), 4)
