"""Drop-in tier (tier v): the bodies of the reference's evaluators, restated line for line against
``ccdm_b200.models`` -- what a maintainer gets after the one-line import switch of INTEGRATION.md.

* ``Tester.test_step``                       /root/reference/evaluation/evaluate_lidc_uncertainty.py:90-125
* ``eval_lidc_uncertainty`` model set-up      :185-200 (``_build_model`` x2, ``PolyakAverager``, checkpoint load :157-161)
* ``Evaluator.load_objects``                  /root/reference/evaluation/eval_cdm.py:131-144 (strict ``load_state_dict``)
* ``Evaluator.predict_single / predict / predict_multiple``   eval_cdm.py:160-193

The evaluation modules themselves cannot be imported (ignite, cityscapesscripts are absent -- SURVEY 8c); the reference
MODEL can, from the bytecode under oracle/_ref, and provides (a) a checkpoint written by ``torch.save`` with the
reference's own state_dict and (b) the numbers to compare with (CPU, same x_T, same exponential draws).
"""
import os

import numpy as np
import pytest
import torch

from conftest import DINO, UNET_PARAMS

pytestmark = pytest.mark.gpu


def _reference_models():
    from oracle.build_ref import load_reference_models
    ref = load_reference_models()
    if ref is None:
        pytest.skip("oracle/_ref (reference bytecode) has not been built")
    return ref


def _build(models, T, C_img, H, W, K, step_T_sample, fce=None, channel_mult=None):
    # trainer.py:589-603 `_build_model`
    p = dict(UNET_PARAMS)
    p["channel_mult"] = channel_mult
    return models.build_model(time_steps=T, schedule="cosine", schedule_params={"s": 0.008}, input_shapes=[(C_img, H, W), (K, H, W)],
                              cond_encoded_shape=(C_img, H, W), backbone="unet_openai", backbone_params=p,
                              dataset_file="datasets.lidc" if K == 2 else "datasets.cityscapes", step_T_sample=step_T_sample,
                              feature_cond_encoder=fce)


def _polyak_init(model, average_model):
    # polyak.py:18-26 PolyakAverager.init_average_model: in-place writes into the live parameter storage
    with torch.no_grad():
        dst_dict = average_model.state_dict()
        for key, value in model.state_dict().items():
            dst_dict[key][...] = value


def _checkpoint(tmp_path, ref_models, T, C_img, H, W, K, fce, mult):
    """{"model", "average_model"} as trainer.py:357-376 saves it: two UNet state_dicts of the REFERENCE classes."""
    from ccdm_b200.synthetic import fill_synthetic_
    ref = _build(ref_models, T, C_img, H, W, K, "majority", fce, mult)
    ref_avg = _build(ref_models, T, C_img, H, W, K, "majority", fce, mult)
    fill_synthetic_(ref.unet, 3)
    fill_synthetic_(ref_avg.unet, 0)
    path = os.path.join(tmp_path, "checkpoint.pt")
    torch.save({"model": ref.unet.state_dict(), "average_model": ref_avg.unet.state_dict()}, path)
    return path, ref_avg


def test_lidc_tester_test_step_body(cuda_device, tmp_path):
    from ccdm_b200 import models
    from ccdm_b200.models import OneHotCategoricalBCHW
    ref_models = _reference_models()
    T, C_img, H, W, K = 12, 1, 64, 64, 2
    device = cuda_device
    path, ref_avg = _checkpoint(str(tmp_path), ref_models, T, C_img, H, W, K, None, None)

    # evaluate_lidc_uncertainty.py:190-200
    model, average_model = [_build(models, T, C_img, H, W, K, "majority").to(device) for _ in range(2)]
    _polyak_init(model, average_model)
    checkpoint = torch.load(path, map_location=device)
    average_model.unet.load_state_dict(checkpoint["average_model"], strict=True)  # ModelCheckpoint.load_objects({"average_model": unet})

    # Tester.test_step :90-125
    num_samples = [2, 3]
    g = torch.Generator().manual_seed(1)
    image = torch.randn(2, C_img, H, W, generator=g)
    labels = torch.nn.functional.one_hot(torch.randint(0, K, (2, 4, H, W), generator=g), K).permute(0, 1, 4, 2, 3).float()  # [B, experts, K, H, W]
    max_num_samples = np.max(num_samples)
    image = image.to(device)
    image = image.repeat_interleave(max_num_samples, dim=0)
    average_model.eval()
    torch.manual_seed(0)
    x = OneHotCategoricalBCHW(logits=torch.zeros(labels[:, 0].repeat_interleave(max_num_samples, dim=0).shape, device=labels.device)).sample().to(device)
    assert x.device.type == "cuda" and not x.is_contiguous()  # drawn on the CPU, NHWC-strided view, copied over
    # the chain's noise: the reference consumes the global generator; inject the same CPU draws on both sides
    torch.manual_seed(42)
    noises = [torch.empty(x.shape[0] * H * W, K).exponential_(1) for _ in range(T - 1)]
    average_model.noise = [n.to(device) for n in noises]
    prediction = average_model(x, image)['diffusion_out']
    assert prediction.dtype == torch.int64 and tuple(prediction.shape) == (6, K, H, W)
    prediction = prediction.reshape(labels.shape[0], -1, *labels.shape[2:])
    labels = labels.to(device)
    labels = labels.argmax(dim=2)
    for idx, samples in enumerate(num_samples):
        pred_np = prediction[:, :samples].argmax(dim=2).cpu().numpy()
        assert pred_np.shape == (2, samples, H, W)
        lcm = np.lcm(samples, labels.shape[1])
        assert prediction[:, :samples].repeat_interleave(lcm // samples, dim=1).argmax(dim=2).shape[1] == lcm
    mean_prediction = torch.log(prediction).mean(dim=(1))  # int64 one-hot -> log gives 0 / -inf, as with the reference
    assert mean_prediction.dtype == torch.float32 and tuple(mean_prediction.shape) == (2, K, H, W)

    # the reference model itself, CPU, same x_T and the same draws (torch.multinomial == exponential race, one draw per step)
    torch.manual_seed(42)
    with torch.no_grad():
        want = ref_avg.eval()(x.cpu(), image.cpu())['diffusion_out']
    agree = float((want.argmax(1) == prediction.reshape(6, K, H, W).argmax(1).cpu()).float().mean())
    assert agree >= 0.999, agree


def test_cityscapes_evaluator_predict_bodies(cuda_device, tmp_path):
    from ccdm_b200 import models
    from ccdm_b200.models import OneHotCategoricalBCHW
    ref_models = _reference_models()
    T, C_img, H, W, K, mult = 10, 3, 64, 128, 20, (1, 1, 2, 2, 4, 4)
    device = cuda_device
    path, ref_avg = _checkpoint(str(tmp_path), ref_models, T, C_img, H, W, K, DINO, mult)
    ref_avg.step_T_sample = "confidence"

    class Evaluator:  # eval_cdm.py:79-193, the members the predict* bodies touch
        pass
    self = Evaluator()
    self.model = _build(models, T, C_img, H, W, K, "confidence", DINO, mult).to(device)
    self.average_model = _build(models, T, C_img, H, W, K, "confidence", DINO, mult).to(device)
    self.num_classes, self.num_evaluations, self.eval_voting_strategy = K, 2, "confidence"
    # load_objects :131-144
    checkpoint = torch.load(path, map_location=device)
    self.model.unet.load_state_dict(checkpoint["model"], True)
    self.average_model.unet.load_state_dict(checkpoint["average_model"], True)

    def predict(xt, condition, feature_condition, label_ref_logits=None):  # :168-174
        self.average_model.eval()
        ret = self.average_model(x=xt, condition=condition, feature_condition=feature_condition, label_ref_logits=label_ref_logits)
        assert ("diffusion_out" in ret)
        return ret["diffusion_out"]

    def predict_single(condition, image, feature_condition, label_ref_logits=None):  # :160-166
        label_shape = (image.shape[0], self.num_classes, *image.shape[2:])
        xt = OneHotCategoricalBCHW(logits=torch.zeros(label_shape, device=image.device)).sample()
        prediction = predict(xt, condition, feature_condition, label_ref_logits)
        return prediction

    def predict_multiple(image, condition, feature_condition):  # :176-193
        assert (self.num_evaluations > 1)
        for i in range(self.num_evaluations):
            prediction_onehot_i = predict_single(image, condition, feature_condition)
            if self.eval_voting_strategy == 'confidence':
                if i == 0:
                    prediction_onehot_total = torch.zeros_like(prediction_onehot_i)
                prediction_onehot_total += prediction_onehot_i * (1 / self.num_evaluations)
            elif self.eval_voting_strategy == 'majority':
                raise NotImplementedError()
            else:
                raise ValueError()
        return prediction_onehot_total

    g = torch.Generator().manual_seed(2)
    image = torch.randn(1, C_img, H, W, generator=g).to(device)
    feature_condition = torch.randn(1, 384, H // 8, W // 8, generator=g).to(device)
    torch.manual_seed(3)
    total = predict_multiple(image, image, feature_condition)  # infer_step :206-211 passes condition == image
    assert total.dtype == torch.float32 and tuple(total.shape) == (1, K, H, W)
    assert float((total.sum(dim=1) - 1).abs().max()) < 1e-5  # a mean of normalised probability maps
    # numbers: predict() on a given x_T with injected draws vs the reference model on the CPU
    torch.manual_seed(5)
    xt = OneHotCategoricalBCHW(logits=torch.zeros(1, K, H, W)).sample()
    torch.manual_seed(42)
    noises = [torch.empty(H * W, K).exponential_(1) for _ in range(T - 1)]
    self.average_model.noise = [n.to(device) for n in noises]
    got = predict(xt.to(device), image, feature_condition)
    torch.manual_seed(42)
    with torch.no_grad():
        want = ref_avg.eval()(x=xt, condition=image.cpu(), feature_condition=feature_condition.cpu(), label_ref_logits=None)["diffusion_out"]
    err = (got.cpu() - want).abs().amax(dim=1)
    assert float((err < 1e-3).float().mean()) >= 0.995, float((err < 1e-3).float().mean())
    # guidance is refused loudly, exactly where the reference would fail on undefined attributes
    with pytest.raises(NotImplementedError):
        predict(xt.to(device), image, feature_condition, label_ref_logits=torch.zeros(1))
