"""Free-running label agreement between the precision modes over the FULL T-step chain (same inputs, same in-kernel
Philox noise): fp32 (FFMA) vs exact (fp16x2 tensor cores) vs bf16.  The chain is chaotic -- one flipped pixel changes every
later UNet input -- so this is the number that says what a precision mode does to the samples a user gets.

    python tools/chain_agreement.py [lidc|cityscapes] [batch] [T] > profiles/rNN_chain_agreement_<workload>.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ccdm-stochastic-segmentation_b200")]
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "lidc"
    wl = dict(bench.WORKLOADS[name])
    B = int(sys.argv[2]) if len(sys.argv) > 2 else (16 if name == "lidc" else 2)
    if len(sys.argv) > 3:
        wl["T"] = int(sys.argv[3])
    res = measure(wl, B, modes=("fp32", "exact", "bf16"))
    res["vs_reference_eager_gpu"] = reference_leg(wl, B)
    print(json.dumps(res))


def reference_leg(wl, B, seed=7):
    """The UNMODIFIED reference (bytecode under oracle/_ref) run eagerly by PyTorch on this GPU, against this library in
    the reference's noise mode (noise='torch': both consume the device generator identically -- one exponential_ draw of
    [B*H*W, K] per step), same x_T, same seed.  Also the reference against ITSELF with cuDNN/cuBLAS TF32 on (its default
    on this GPU) and off: the noise floor of "identical samples" for this chaotic chain."""
    from ccdm_b200.synthetic import synthetic_inputs
    dev = torch.device("cuda", torch.cuda.current_device())
    ref, kind = bench.build_model(wl, dev, reference=True)
    if kind != "reference":
        return {"unavailable": "oracle/_ref bytecode missing"}
    ours, _ = bench.build_model(wl, dev)
    image, feat, labels = synthetic_inputs(B, wl["C_img"], wl["H"], wl["W"], wl["K"], 384 if wl["fce"] else 0, seed=77)
    x = torch.nn.functional.one_hot(labels.long(), wl["K"]).permute(0, 3, 1, 2).float().to(dev)
    image, feat = image.to(dev), (feat.to(dev) if feat is not None else None)
    out = {}

    def run_ref(tf32):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.manual_seed(seed)
        with torch.no_grad():
            return ref(x, image, feat)["diffusion_out"].argmax(1)

    labs = {"reference_tf32_on": run_ref(True), "reference_tf32_off": run_ref(False)}
    torch.backends.cudnn.allow_tf32 = True
    ours.noise = "torch"
    for prec in ("exact", "fp32", "bf16"):
        ours.precision = prec
        torch.manual_seed(seed)
        labs["ours_" + prec] = ours(x, image, feat)["diffusion_out"].argmax(1)
    names = list(labs)
    for i in range(len(names)):
        for j in range(i + 1, len(names)):
            out[f"{names[i]}_vs_{names[j]}"] = float((labs[names[i]] == labs[names[j]]).float().mean())
    out["note"] = ("final labels after the full free-running chain, torch generator seed %d on every run; reference_tf32_on is how the "
                   "reference runs on this GPU out of the box (torch.backends.cudnn.allow_tf32 defaults to True)" % seed)
    return out


def measure(wl, B, modes=("exact", "bf16"), seed=2024, dev=None):
    """{(a, b): fraction of equal final labels} for every pair of `modes`, plus per-mode chain seconds."""
    from ccdm_b200.synthetic import synthetic_inputs
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    m, _ = bench.build_model(wl, dev)
    image, feat, labels = synthetic_inputs(B, wl["C_img"], wl["H"], wl["W"], wl["K"], 384 if wl["fce"] else 0, seed=77)
    m.noise, m.seed = "philox", seed
    out, secs = {}, {}
    for prec in modes:
        m.precision = prec
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out[prec] = m(labels.to(dev), image.to(dev), feat.to(dev) if feat is not None else None)["diffusion_out"].argmax(1)
        e1.record()
        torch.cuda.synchronize()
        secs[prec] = e0.elapsed_time(e1) / 1e3
    res = {"workload": wl["name"], "batch": B, "T": wl["T"], "noise": f"philox seed {seed}", "chain_seconds": secs, "agreement": {},
           "per_sample_min": {}}
    ms = list(modes)
    for i in range(len(ms)):
        for j in range(i + 1, len(ms)):
            eq = (out[ms[i]] == out[ms[j]]).float()
            res["agreement"][f"{ms[i]}_vs_{ms[j]}"] = float(eq.mean())
            res["per_sample_min"][f"{ms[i]}_vs_{ms[j]}"] = float(eq.flatten(1).mean(1).min())
    return res


if __name__ == "__main__":
    main()
