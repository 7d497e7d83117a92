"""Milestone timeline (ns, CTA 0) of conv_tma launches of the LIDC step: where the latency of a launch goes."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ccdm-stochastic-segmentation_b200"))
import torch
import bench
from ccdm_b200 import _lib
from ccdm_b200.models.diffusion_denoising import reverse_t_values
from ccdm_b200.synthetic import synthetic_inputs

wlname = sys.argv[1] if len(sys.argv) > 1 else "lidc"
wl = bench.WORKLOADS[wlname]
L = _lib.lib()
dev = torch.device("cuda", 0)
m, _ = bench.build_model(wl, dev)
eng = m.unet.engine(sys.argv[2] if len(sys.argv) > 2 else "bf16")
B = wl["B"]
image, feat, labels = synthetic_inputs(B, wl["C_img"], wl["H"], wl["W"], wl["K"], 384 if wl["fce"] else 0)
al, ca = m._schedule_host()
eng.run_chain(labels.to(dev), image.to(dev), feat.to(dev) if feat is not None else None, reverse_t_values(wl["T"], 10003), al, ca,
              _lib.DRAW_MAJORITY, noise="philox", seed=1)
prog = eng.program(B, wl["H"], wl["W"])
names = ["start", "setup", "affine", "raw0", "xf0", "mma0", "epi0", "flush", "end"]
seen = set()
sp = _lib.stream_ptr(eng.stream)
buf = (ctypes.c_uint64 * 400)()
with torch.cuda.stream(eng.stream):
    for i in range(prog.n_ops):
        o = prog._op_array[i]
        if o.kind != _lib.OP_CONV or not L.ccdm_conv_uses_tma(ctypes.byref(o)):
            continue
        c = bench.op_class(o)
        if c in seen:
            continue
        seen.add(c)
        for _ in range(3):
            _lib.check(L.ccdm_launch_op(ctypes.byref(o), sp))
        eng.stream.synchronize()
        L.ccdm_debug_conv_trace(buf, 400)
        t = [int(v) for v in buf[:9]]
        rel = [(v - t[0]) / 1e3 if v >= t[0] else float("nan") for v in t]
        print(f"{c:44s} " + " ".join(f"{n}={r:6.1f}" for n, r in zip(names[1:], rel[1:])))
        cfg = (ctypes.c_int32 * 16)()
        L.ccdm_conv_tc_config(ctypes.byref(o), cfg)
        grid = min(int(cfg[13]), 160)
        st = sorted(int(v) - t[0] for v in buf[80:80 + grid])
        en = sorted(int(v) - t[0] for v in buf[240:240 + grid])
        dur = sorted(int(buf[240 + k]) - int(buf[80 + k]) for k in range(grid))
        print(f"      CTAs {grid} items {cfg[12]}: start min/med/max {st[0] / 1e3:.1f}/{st[grid // 2] / 1e3:.1f}/{st[-1] / 1e3:.1f}  "
              f"end {en[0] / 1e3:.1f}/{en[grid // 2] / 1e3:.1f}/{en[-1] / 1e3:.1f}  duration {dur[0] / 1e3:.1f}/{dur[grid // 2] / 1e3:.1f}/{dur[-1] / 1e3:.1f} us")
        if "128x128" in c or "64x64" in c or "256x512" in c:
            tl = [int(v) for v in buf[16:80]]
            for role, rn in enumerate(("tma", "xform", "mma", "epi")):
                row = []
                for item in range(8):
                    b, e = tl[(role * 8 + item) * 2], tl[(role * 8 + item) * 2 + 1]
                    row.append(f"[{(b - t[0]) / 1e3:5.1f},{(e - t[0]) / 1e3:5.1f}]" if b >= t[0] and e >= t[0] else "[  -  ]")
                print(f"      {rn:6s} " + " ".join(row))
