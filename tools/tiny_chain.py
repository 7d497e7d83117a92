"""A tiny bf16 reverse chain (LIDC-shaped 64x64, B=2, 3 strided steps) for compute-sanitizer / ncu smoke runs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ccdm-stochastic-segmentation_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from ccdm_b200 import models  # noqa: E402
from ccdm_b200.synthetic import fill_synthetic_, synthetic_inputs  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
H = W = int(sys.argv[2]) if len(sys.argv) > 2 else 64
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
p = dict(base_channels=32, channel_mult=None, attention_resolutions=[32, 16, 8], num_heads=1, num_head_channels=32, softmax_output=True)
m = models.build_model(250, "cosine", {"s": 0.008}, [(1, H, W), (2, H, W)], (1, H, W), "unet_openai", p, "datasets.lidc", "majority", None).eval()
fill_synthetic_(m.unet, 0)
m = m.cuda()
m.precision, m.noise, m.seed = prec, "philox", 1
m.unet.engine(prec).use_graph = os.environ.get("CCDM_NO_GRAPH", "0") != "1"
image, _, labels = synthetic_inputs(B, 1, H, W, 2)
x = torch.nn.functional.one_hot(labels.long(), 2).permute(0, 3, 1, 2).float().cuda()
out = m(x, image.cuda(), None, t=torch.as_tensor(10003))["diffusion_out"]
torch.cuda.synchronize()
print("tiny chain ok", tuple(out.shape), int(out.argmax(1).sum()))
