"""End-to-end parity on the B200: UNet evaluation and the reverse chain through the
reference-shaped Python surface (ccdm_b200.models), against (a) the fixtures generated from
the unmodified reference and (b) the CPU oracle on seeded inputs.

Two parity-grade precision modes are held to the SAME tolerances: 'fp32' (FFMA kernels) and 'exact' (tensor cores on
fp16 hi + lo operands, three MMAs per product -- the default of DenoisingModel).  Tolerances:
  * x0 prediction (softmax probabilities):   max |dp| <= 2e-4
  * posterior log-probabilities:              max |d log p| <= 1e-3 on entries with p >= 1e-6
  * sampled labels, teacher-forced per step:  identical wherever the top-2 race margin
    |log(p_i/E_i) - log(p_j/E_j)| > 1e-3 (near-ties may flip under a different fp32 summation order)
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, ROOT, build_ours, golden

pytestmark = pytest.mark.gpu

X0_TOL = 2e-4
RACE_MARGIN = 1e-3
REPORT = {}
PARITY_MODES = ["fp32", "exact"]


def _report(key, **vals):
    REPORT[key] = {k: (float(v) if not isinstance(v, (list, str)) else v) for k, v in vals.items()}
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_report.json"), "w") as fh:
        json.dump(REPORT, fh, indent=1, sort_keys=True)


def _case(tag, step_T_sample="majority"):
    from ccdm_b200.synthetic import synthetic_inputs
    T, B, C_img, H, W, K, fce, mult, t_probe, steps = GOLDEN_CASES[tag]
    m = build_ours(T, C_img, H, W, K, step_T_sample, fce, mult).cuda()
    image, feat, labels = synthetic_inputs(B, C_img, H, W, K, 384 if fce else 0)
    return m, image, feat, labels


def _onehot(labels, K):
    return torch.nn.functional.one_hot(labels.long(), K).permute(0, 3, 1, 2).float()


@pytest.mark.parametrize("prec", PARITY_MODES)
@pytest.mark.parametrize("tag", ["lidc64", "lidc128", "cs64x128"])
def test_unet_matches_reference_fixture(cuda_device, tag, prec):
    T, B, C_img, H, W, K, fce, mult, t_probe, steps = GOLDEN_CASES[tag]
    g = golden(tag + ".npz")
    m, image, feat, labels = _case(tag)
    m.unet.precision = prec
    x = _onehot(labels, K).cuda()
    for t in t_probe:
        out = m.unet(x, image.cuda(), feat.cuda() if feat is not None else None, torch.full((B,), float(t)).cuda())
        p = out["diffusion_out"]
        assert p.shape == (B, K, H, W) and out["logits"] is None
        err = np.abs(p.permute(0, 2, 3, 1).cpu().numpy() - g[f"x0pred_t{t}"]).max()
        _report(f"unet_{prec}_{tag}_t{t}", max_abs_err=err)
        assert err <= X0_TOL, f"{tag} t={t}: max|dp|={err}"


@pytest.mark.parametrize("prec", PARITY_MODES)
def test_unet_layerwise_vs_oracle(cuda_device, prec):
    """Every fused kernel's output against the torch fp32 restatement, layer by layer."""
    from oracle import unet_ref
    tag = "lidc64"
    T, B, C_img, H, W, K, fce, mult, t_probe, steps = GOLDEN_CASES[tag]
    m, image, feat, labels = _case(tag)
    taps = {}
    unet_ref.unet_forward({k: v.cpu() for k, v in m.unet.state_dict().items()}, _onehot(labels, K), image, feat,
                          torch.full((B,), 37.0), taps=taps)
    tr = m.unet.engine(prec).trace_step(labels.cuda(), image.cuda(), None, 37.0)
    worst = {}
    for name, ref in taps.items():
        key = name if name in tr else (name + ".op" if name + ".op" in tr else name + ".conv")
        if key not in tr:
            continue
        got = tr[key].permute(0, 3, 1, 2).cpu()
        scale = float(ref.abs().max()) + 1e-6
        worst[name] = float((got - ref).abs().max()) / scale
    assert len(worst) >= 30
    _report(f"layerwise_{prec}_lidc64", worst_rel=max(worst.values()), worst_layer=max(worst, key=worst.get))
    bad = {k: v for k, v in worst.items() if v > 1e-4}
    assert not bad, bad


def test_forward_step_validation_path_per_sample_t(cuda_device):
    """DenoisingModel.forward(validation=True) = one UNet evaluation with a per-sample t (:154-155)."""
    from oracle import unet_ref
    tag = "lidc64"
    T, B, C_img, H, W, K, fce, mult, t_probe, steps = GOLDEN_CASES[tag]
    m, image, feat, labels = _case(tag)
    t = torch.tensor([3.0, 201.0])
    ref = unet_ref.unet_forward({k: v.cpu() for k, v in m.unet.state_dict().items()}, _onehot(labels, K), image, None, t)
    out = m(_onehot(labels, K).cuda(), image.cuda(), None, t.cuda(), validation=True)["diffusion_out"]
    assert float((out.cpu() - ref).abs().max()) <= X0_TOL
    m.train()
    with pytest.raises(ValueError):
        m(_onehot(labels, K).cuda(), image.cuda(), None, None)
    out2 = m(_onehot(labels, K).cuda(), image.cuda(), None, t.cuda())["diffusion_out"]
    assert torch.equal(out, out2)


def _teacher_forced_check(m, image, feat, fce, K, T, record):
    """Replay every recorded GPU step on the CPU oracle with the same inputs and noise."""
    from oracle import cdm, unet_ref
    sd = {k: v.cpu() for k, v in m.unet.state_dict().items()}
    al, ca = m.diffusion.alphas.cpu().numpy(), m.diffusion.cumalphas.cpu().numpy()
    stats = dict(max_dlogp=0.0, mismatch_outside_margin=0, mismatch_total=0, pixels=0, max_dx0=0.0)
    for rec in record:
        t = rec["t"]
        lab_in = rec["labels_in"].cpu().numpy()
        B = lab_in.shape[0]
        theta = unet_ref.unet_forward(sd, _onehot(torch.as_tensor(lab_in), K), image, feat, torch.full((B,), float(t)),
                                      feature_condition_idx=10 if fce else None).permute(0, 2, 3, 1).contiguous().numpy()
        theta_gpu = torch.softmax(rec["logits"].float(), dim=-1).cpu().numpy()
        stats["max_dx0"] = max(stats["max_dx0"], float(np.abs(theta - theta_gpu).max()))
        a, c = cdm.step_scalars(al, ca, t)
        post = cdm.posterior_closed(lab_in, theta, a, c)
        post_gpu = cdm.posterior_closed(lab_in, theta_gpu, a, c)
        big = post >= 1e-6
        stats["max_dlogp"] = max(stats["max_dlogp"], float(np.abs(np.log(post[big]) - np.log(np.maximum(post_gpu[big], 1e-30))).max()))
        if t > 1:
            noise = rec["noise"].cpu().numpy().reshape(post.shape)
            lab, pn = cdm.draw(post, noise, 0)
            score = np.log(pn) - np.log(noise)
        else:
            lab, pn = cdm.draw(post, None, 1)
            score = np.log(pn)
        top2 = np.sort(score, axis=-1)[..., -2:]
        margin = top2[..., 1] - top2[..., 0]
        got = rec["labels_out"].cpu().numpy()
        diff = got != lab
        stats["mismatch_total"] += int(diff.sum())
        stats["mismatch_outside_margin"] += int((diff & (margin > RACE_MARGIN)).sum())
        stats["pixels"] += diff.size
    return stats


@pytest.mark.parametrize("prec", PARITY_MODES)
@pytest.mark.parametrize("tag,noise", [("lidc64", "torch"), ("lidc64", "philox"), ("cs64x128", "torch")])
def test_chain_teacher_forced_vs_oracle(cuda_device, tag, noise, prec):
    T, B, C_img, H, W, K, fce, mult, t_probe, steps = GOLDEN_CASES[tag]
    m, image, feat, labels = _case(tag)
    from ccdm_b200 import _lib
    from ccdm_b200.models.diffusion_denoising import reverse_t_values
    ts = reverse_t_values(T, 10000 + steps)
    record = []
    torch.manual_seed(7)
    al, ca = m._schedule_host()
    m.unet.engine(prec).run_chain(_onehot(labels, K).cuda(), image.cuda(), feat.cuda() if feat is not None else None, ts,
                                  al, ca, _lib.DRAW_MAJORITY, noise=noise, seed=1234, record=record)
    assert [r["t"] for r in record] == ts
    stats = _teacher_forced_check(m, image, feat, fce, K, T, record)
    _report(f"teacher_forced_{prec}_{tag}_{noise}", **stats)
    assert stats["max_dx0"] <= X0_TOL
    assert stats["max_dlogp"] <= 1e-3
    assert stats["mismatch_outside_margin"] == 0
    assert stats["mismatch_total"] <= 1e-4 * stats["pixels"] + 2


@pytest.mark.parametrize("prec", PARITY_MODES)
@pytest.mark.parametrize("tag", ["lidc64", "cs64x128"])
def test_chain_replays_reference_fixture_with_injected_noise(cuda_device, tag, prec):
    """The reference's own chain output (CPU generator, seed 42), reproduced on the GPU by injecting
    the same exponential draws.  Free-running: one near-tie flip would propagate, so agreement is
    reported and required to be >= 99.9 %; in practice it is exact."""
    T, B, C_img, H, W, K, fce, mult, t_probe, steps = GOLDEN_CASES[tag]
    g = golden(tag + ".npz")
    torch.manual_seed(42)
    noises = [torch.empty(B * H * W, K).exponential_(1) for _ in range(steps - 1)]
    for mode in ("majority", "confidence"):
        m, image, feat, labels = _case(tag, mode)
        m.precision = prec
        m.noise = [n.cuda() for n in noises]
        out = m(_onehot(labels, K).cuda(), image.cuda(), feat.cuda() if feat is not None else None,
                t=torch.as_tensor(10000 + steps))["diffusion_out"]
        assert out.shape == (B, K, H, W)
        if mode == "majority":
            assert out.dtype == torch.int64 and not out.is_contiguous()  # NHWC-strided view, like the reference
            agree = float((out.argmax(1).cpu().numpy() == g["chain_majority_labels"]).mean())
            _report(f"fixture_chain_{prec}_{tag}_majority", agreement=agree)
            assert agree >= 0.999
        else:
            assert out.dtype == torch.float32
            err = np.abs(out.permute(0, 2, 3, 1).cpu().numpy() - g["chain_confidence_probs"])
            frac_close = float((err.max(axis=-1) < 1e-3).mean())
            _report(f"fixture_chain_{prec}_{tag}_confidence", frac_close=frac_close, median_err=float(np.median(err)))
            assert frac_close >= 0.995


def test_torch_generator_parity_with_reference_sampler(cuda_device):
    """on the B200, torch.multinomial(p, 1, True) == argmax(p / exponential_)
    under the same seed -- the identity the 'torch' noise mode relies on -- and our OneHotCategoricalBCHW
    draws the same x_T as torch.distributions' (the class the reference wraps)."""
    from ccdm_b200.models import OneHotCategoricalBCHW
    p = torch.rand(4096, 20, device="cuda")
    p = p / p.sum(-1, keepdim=True)
    torch.manual_seed(5)
    a = torch.multinomial(p, 1, True).squeeze(1)
    torch.manual_seed(5)
    b = (p / torch.empty_like(p).exponential_(1)).argmax(-1)
    assert torch.equal(a, b)
    torch.manual_seed(11)
    ours = OneHotCategoricalBCHW(logits=torch.zeros(2, 20, 16, 16, device="cuda")).sample()
    torch.manual_seed(11)
    ref = torch.distributions.OneHotCategorical(logits=torch.zeros(2, 16, 16, 20, device="cuda")).sample().permute(0, 3, 1, 2)
    assert torch.equal(ours, ref) and ours.stride() == ref.stride()


def test_chain_is_deterministic_and_sharding_invariant(cuda_device):
    """Philox mode: same seed -> same bits; samples [2,4) run as their own shard == samples 2..3 of the batch."""
    tag = "lidc64"
    T, _, C_img, H, W, K, fce, mult, t_probe, steps = GOLDEN_CASES[tag]
    from ccdm_b200.synthetic import synthetic_inputs
    m = build_ours(T, C_img, H, W, K, "majority").cuda()
    image, _, labels = synthetic_inputs(4, C_img, H, W, K)
    m.noise, m.seed = "philox", 77
    tt = torch.as_tensor(10000 + 6)
    for prec in ("fp32", "exact"):
        m.precision, m.sample_offset = prec, 0
        full = m(_onehot(labels, K).cuda(), image.cuda(), None, t=tt)["diffusion_out"]
        again = m(_onehot(labels, K).cuda(), image.cuda(), None, t=tt)["diffusion_out"]
        assert torch.equal(full, again)  # deterministic in every mode
        m.sample_offset = 2
        shard = m(_onehot(labels[2:], K).cuda(), image[2:].cuda(), None, t=tt)["diffusion_out"]
        if prec == "fp32":
            assert torch.equal(full[2:], shard)
        else:
            # default tiling follows the batch: the fp32 per-item partial sums of the GroupNorm statistics differ at the 1e-7
            # level between a batch of 4 and a batch of 2, a near-tie may flip
            agree = float((full[2:] == shard).float().mean())
            _report("exact_shard_agreement_default_tiling", agreement=agree)
            assert agree >= 0.999, agree
            # batch-independent tiling (what sample_sharded sets): bit-identical
            m.tile_batch, m.sample_offset = 64, 0
            full = m(_onehot(labels, K).cuda(), image.cuda(), None, t=tt)["diffusion_out"]
            m.sample_offset = 2
            shard = m(_onehot(labels[2:], K).cuda(), image[2:].cuda(), None, t=tt)["diffusion_out"]
            m.tile_batch = 0
            assert torch.equal(full[2:], shard)


def test_sub_batch_lanes_do_not_change_the_result(cuda_device):
    """engine.lanes > 1 splits the batch into sub-batch programs on separate streams (experimental); Philox noise is keyed
    by the global sample index, so labels must be identical to the single-program run (exact in fp32 mode)."""
    tag = "lidc64"
    T, _, C_img, H, W, K, fce, mult, t_probe, steps = GOLDEN_CASES[tag]
    from ccdm_b200.synthetic import synthetic_inputs
    image, _, labels = synthetic_inputs(5, C_img, H, W, K)
    tt = torch.as_tensor(10000 + 5)
    for prec in ("fp32", "exact", "bf16"):
        m = build_ours(T, C_img, H, W, K, "majority").cuda()
        m.noise, m.seed, m.precision = "philox", 21, prec
        eng = m.unet.engine(prec)
        eng.lanes = 1
        one = m(_onehot(labels, K).cuda(), image.cuda(), None, t=tt)["diffusion_out"]
        eng.lanes = 3
        three = m(_onehot(labels, K).cuda(), image.cuda(), None, t=tt)["diffusion_out"]
        eng.lanes = 1
        if prec == "fp32":
            assert torch.equal(one, three)
        else:
            # bf16 / tensor-core mode: conv outputs do not depend on the batch split, but the fp32 partial sums of the
            # GroupNorm statistics are grouped per CTA, so a different split changes their summation order (1e-7
            # relative) and a near-tie can flip: agreement, not bit equality
            agree = float((one == three).float().mean())
            _report(f"{prec}_lanes_agreement", agreement=agree)
            assert agree >= (0.999 if prec == "exact" else 0.995), agree


def test_x_T_drawn_on_device_and_label_input(cuda_device):
    """DenoisingModel.draw_x_T: uniform uint8 labels from the device-side race, keyed by the global sample
    index; forward() takes them as they are and gives the result of the one-hot float input the reference passes."""
    tag = "lidc64"
    T, _, C_img, H, W, K, fce, mult, t_probe, steps = GOLDEN_CASES[tag]
    from ccdm_b200.synthetic import synthetic_inputs
    image, _, _ = synthetic_inputs(4, C_img, H, W, K)
    m = build_ours(T, C_img, H, W, K, "majority").cuda()
    m.noise, m.seed = "philox", 11
    xt = m.draw_x_T(4, H, W)
    assert xt.dtype == torch.uint8 and tuple(xt.shape) == (4, H, W) and int(xt.max()) < K
    frac = float((xt == 1).float().mean())
    assert 0.47 < frac < 0.53, frac                      # K = 2: a fair coin per pixel
    assert not torch.equal(xt[0], xt[1])                 # samples differ
    m.sample_offset = 2                                  # rank 1 of 2 draws ITS samples of the same batch
    assert torch.equal(m.draw_x_T(2, H, W), xt[2:4])
    m.sample_offset = 0
    m.seed = 12
    assert not torch.equal(m.draw_x_T(4, H, W), xt)      # and the seed matters
    m.seed = 11
    tt = torch.as_tensor(10000 + 3)
    a = m(xt, image.cuda(), None, t=tt)["diffusion_out"]
    b = m(_onehot(xt.cpu(), K).cuda(), image.cuda(), None, t=tt)["diffusion_out"]
    assert a.dtype == torch.int64 and torch.equal(a, b)
    with pytest.raises(Exception):
        build_ours(T, C_img, H, W, K, "majority").draw_x_T(1, H, W)  # model on the CPU: loud


def test_graph_replay_equals_eager_launches(cuda_device):
    tag = "lidc64"
    T, B, C_img, H, W, K, fce, mult, t_probe, steps = GOLDEN_CASES[tag]
    m, image, feat, labels = _case(tag)
    m.noise, m.seed = "philox", 5
    eng = m.unet.engine(m.precision)
    tt = torch.as_tensor(10000 + 5)
    eng.use_graph = True
    a = m(_onehot(labels, K).cuda(), image.cuda(), None, t=tt)["diffusion_out"]
    eng.use_graph = False
    b = m(_onehot(labels, K).cuda(), image.cuda(), None, t=tt)["diffusion_out"]
    eng.use_graph = True
    assert torch.equal(a, b)


def test_weight_cache_follows_parameter_updates(cuda_device):
    """SURVEY.md 8b staleness hazard: in-place writes / load_state_dict after the first forward must be seen."""
    tag = "lidc64"
    T, B, C_img, H, W, K, fce, mult, t_probe, steps = GOLDEN_CASES[tag]
    m, image, feat, labels = _case(tag)
    x, t = _onehot(labels, K).cuda(), torch.full((B,), 9.0).cuda()
    a = m.unet(x, image.cuda(), None, t)["diffusion_out"].clone()
    with torch.no_grad():
        m.unet.out._modules["2"].weight.mul_(0.0)
        m.unet.out._modules["2"].bias.mul_(0.0)
    b = m.unet(x, image.cuda(), None, t)["diffusion_out"]
    assert float((b - 1.0 / K).abs().max()) < 1e-6  # zero head -> uniform prediction
    from ccdm_b200.synthetic import fill_synthetic_
    fill_synthetic_(m.unet, 0)  # load_state_dict path
    c = m.unet(x, image.cuda(), None, t)["diffusion_out"]
    assert torch.equal(a, c)


def test_no_fallbacks(cuda_device):
    tag = "lidc64"
    T, B, C_img, H, W, K, fce, mult, t_probe, steps = GOLDEN_CASES[tag]
    from ccdm_b200 import _lib
    from ccdm_b200.synthetic import synthetic_inputs
    m = build_ours(T, C_img, H, W, K)  # parameters on the CPU
    image, _, labels = synthetic_inputs(B, C_img, H, W, K)
    with pytest.raises(_lib.CcdmError):
        m(_onehot(labels, K), image, None)
    with pytest.raises(NotImplementedError):
        m.cuda()(_onehot(labels, K).cuda(), image.cuda(), None, label_ref_logits=torch.zeros(1))


# ---------------------------------------------------------------------------------------------
# bf16 ("fast") precision: bf16 activation storage, tcgen05 tensor-core convs.
# Stated tolerance (precision ladder: bf16 autocast on the reference gives
# max |d x0| 2.1e-2 and flips 3e-4 (mean) .. 6.3e-3 (worst step) of the sampled labels per step):
#   max |d x0| <= 6e-2, mean |d x0| <= 4e-3, teacher-forced label mismatch <= 1e-2 per chain.
# ---------------------------------------------------------------------------------------------
BF16_X0_MAX, BF16_X0_MEAN, BF16_LABEL_FRAC = 6e-2, 4e-3, 1e-2


@pytest.mark.parametrize("tag", ["lidc64", "lidc128", "cs64x128"])
def test_bf16_unet_vs_reference_fixture(cuda_device, tag):
    T, B, C_img, H, W, K, fce, mult, t_probe, steps = GOLDEN_CASES[tag]
    g = golden(tag + ".npz")
    m, image, feat, labels = _case(tag)
    m.unet.precision = "bf16"
    x = _onehot(labels, K).cuda()
    for t in t_probe:
        p = m.unet(x, image.cuda(), feat.cuda() if feat is not None else None, torch.full((B,), float(t)).cuda())["diffusion_out"]
        err = np.abs(p.permute(0, 2, 3, 1).cpu().numpy() - g[f"x0pred_t{t}"])
        flips = float((p.argmax(1).cpu().numpy() != g[f"x0pred_t{t}"].argmax(-1)).mean())
        _report(f"bf16_unet_{tag}_t{t}", max_abs_err=err.max(), mean_abs_err=err.mean(), argmax_flips=flips)
        assert err.max() <= BF16_X0_MAX and err.mean() <= BF16_X0_MEAN, (err.max(), err.mean())
    eng = m.unet.engine("bf16")
    assert eng.program(B, H, W, 1).n_tc > 40  # the tensor-core kernels are what ran


def test_bf16_layerwise_vs_oracle(cuda_device):
    from oracle import unet_ref
    tag = "lidc64"
    T, B, C_img, H, W, K, fce, mult, t_probe, steps = GOLDEN_CASES[tag]
    m, image, feat, labels = _case(tag)
    taps = {}
    unet_ref.unet_forward({k: v.cpu() for k, v in m.unet.state_dict().items()}, _onehot(labels, K), image, feat,
                          torch.full((B,), 37.0), taps=taps)
    tr = m.unet.engine("bf16").trace_step(labels.cuda(), image.cuda(), None, 37.0)
    worst = {}
    for name, ref in taps.items():
        key = name if name in tr else (name + ".op" if name + ".op" in tr else name + ".conv")
        if key not in tr:
            continue
        got = tr[key].permute(0, 3, 1, 2).cpu()
        worst[name] = float((got - ref).abs().max()) / (float(ref.abs().max()) + 1e-6)
    _report("bf16_layerwise_lidc64", worst_rel=max(worst.values()), worst_layer=max(worst, key=worst.get),
            first_layers=[round(worst[k], 5) for k in list(worst)[:6]])
    bad = {k: v for k, v in worst.items() if v > 5e-2}
    assert not bad, bad


@pytest.mark.parametrize("tag", ["lidc64", "cs64x128"])
def test_bf16_chain_teacher_forced_label_agreement(cuda_device, tag):
    T, B, C_img, H, W, K, fce, mult, t_probe, steps = GOLDEN_CASES[tag]
    m, image, feat, labels = _case(tag)
    from ccdm_b200 import _lib
    from ccdm_b200.models.diffusion_denoising import reverse_t_values
    ts = reverse_t_values(T, 10000 + steps)
    record = []
    al, ca = m._schedule_host()
    m.unet.engine("bf16").run_chain(_onehot(labels, K).cuda(), image.cuda(), feat.cuda() if feat is not None else None, ts,
                                    al, ca, _lib.DRAW_MAJORITY, noise="philox", seed=99, record=record)
    stats = _teacher_forced_check(m, image, feat, fce, K, T, record)
    frac = stats["mismatch_total"] / stats["pixels"]
    _report(f"bf16_teacher_forced_{tag}", label_mismatch_frac=frac, **stats)
    assert stats["max_dx0"] <= BF16_X0_MAX
    assert frac <= BF16_LABEL_FRAC


def test_sample_many_indexes_the_image_and_votes_on_device(cuda_device):
    """N samples per image in one chain without image.repeat_interleave(N) (the kernels read entry
    sample // N), x_T drawn on the device, the vote over the N samples as one kernel.  Same labels as the evaluator-style
    call on the replicated batch (evaluate_lidc_uncertainty.py:96-103); the vote equals the mean of the one-hot maps."""
    from ccdm_b200.synthetic import synthetic_inputs
    for tag, n in (("lidc64", 3), ("cs64x128", 2)):
        T, _, C_img, H, W, K, fce, mult, t_probe, steps = GOLDEN_CASES[tag]
        m = build_ours(T, C_img, H, W, K, "majority", fce, mult).cuda()
        image, feat, _ = synthetic_inputs(2, C_img, H, W, K, 384 if fce else 0)
        m.noise, m.seed = "philox", 31
        for prec in ("exact", "bf16"):
            m.precision = prec
            out = m.sample_many(image.cuda(), n, feat.cuda() if feat is not None else None, init_t=10000 + 4)
            assert out["labels"].dtype == torch.uint8 and tuple(out["labels"].shape) == (2, n, H, W)
            # the evaluator's way: replicate, draw the same x_T, one call
            x = m.draw_x_T(2 * n, H, W)
            ref = m(x, image.repeat_interleave(n, dim=0).cuda(), feat.repeat_interleave(n, dim=0).cuda() if feat is not None else None,
                    t=torch.as_tensor(10000 + 4))["diffusion_out"]
            assert torch.equal(out["labels"].reshape(2 * n, H, W).long(), ref.argmax(1))
            onehot = torch.nn.functional.one_hot(out["labels"].long(), K).float()          # [2, n, H, W, K]
            assert torch.allclose(out["mean_onehot"], onehot.mean(dim=1).permute(0, 3, 1, 2), atol=1e-6)
            assert torch.equal(out["majority"].long(), out["mean_onehot"].argmax(dim=1))
    with pytest.raises(ValueError):
        m.sample_many(image.cuda(), 0)


@pytest.mark.parametrize("prec", PARITY_MODES + ["bf16"])
def test_feature_fold_equals_the_per_step_concatenation(cuda_device, prec):
    """input_blocks[10] with the constant DINO channels folded into per-chain maps (the default) against the same
    block run on the 448-channel concatenation every step (engine.fold_features = False): the UNet outputs agree to the
    storage precision of the two maps, and both sit inside the tolerance against the reference fixture."""
    T, B, C_img, H, W, K, fce, mult, t_probe, steps = GOLDEN_CASES["cs64x128"]
    g = golden("cs64x128.npz")
    m, image, feat, labels = _case("cs64x128")
    m.unet.precision = prec
    x = _onehot(labels, K).cuda()
    eng = m.unet.engine(prec)
    outs = {}
    for fold in (True, False):
        eng.fold_features = fold
        prog = eng.program(B, H, W, rows_per_sample=1)
        assert bool(prog.fmaps) == fold
        outs[fold] = m.unet(x, image.cuda(), feat.cuda(), torch.full((B,), float(t_probe[0])).cuda())["diffusion_out"].permute(0, 2, 3, 1).cpu().numpy()
    eng.fold_features = True
    d = float(np.abs(outs[True] - outs[False]).max())
    e = [float(np.abs(outs[f] - g[f"x0pred_t{t_probe[0]}"]).max()) for f in (True, False)]
    _report(f"feature_fold_{prec}", folded_vs_unfolded=d, folded_vs_fixture=e[0], unfolded_vs_fixture=e[1])
    tol = X0_TOL if prec != "bf16" else 6e-2
    assert d <= tol and e[0] <= tol and e[1] <= tol, (d, e)
