"""Informational: the UNMODIFIED reference (bytecode under oracle/_ref) run eagerly by PyTorch on the same B200,
fp32, the bench's workloads and batch sizes -- the only "Blackwell path" that existed before this repo.  Prints one JSON line
per workload: ms per reverse step and the samples/s the full T-step chain would reach.  Not part of bench.py's contract."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ccdm-stochastic-segmentation_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from ccdm_b200.synthetic import synthetic_inputs  # noqa: E402

dev = torch.device("cuda", 0)
for name, n_steps in (("lidc", 10), ("cityscapes", 4)):
    wl = bench.WORKLOADS[name]
    m, kind = bench.build_model(wl, dev, reference=True)
    if kind != "reference":
        print(json.dumps({"workload": name, "unavailable": "oracle/_ref bytecode missing"}))
        continue
    B = wl["B"]
    image, feat, labels = synthetic_inputs(B, wl["C_img"], wl["H"], wl["W"], wl["K"], 384 if wl["fce"] else 0)
    x = torch.nn.functional.one_hot(labels.long(), wl["K"]).permute(0, 3, 1, 2).float().to(dev)
    image = image.to(dev)
    feat = feat.to(dev) if feat is not None else None
    torch.manual_seed(0)
    with torch.no_grad():
        m(x, image, feat, t=torch.as_tensor(10000 + 2))  # warm-up: cuDNN autotune, allocator
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m(x, image, feat, t=torch.as_tensor(10000 + n_steps))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    per_step = dt / n_steps
    print(json.dumps({"workload": wl["name"] if "name" in wl else name, "impl": "reference, PyTorch eager fp32 on the B200", "batch": B,
                      "steps_timed": n_steps, "ms_per_reverse_step": round(per_step * 1e3, 3),
                      "samples_per_s_full_chain": round(B / (per_step * wl["T"]), 3), "torch": torch.__version__,
                      "tf32": bool(torch.backends.cudnn.allow_tf32), "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}))
    del m
    torch.cuda.empty_cache()
