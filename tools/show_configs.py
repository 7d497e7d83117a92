"""Prints the tile / pipeline configuration of every op of one reverse step (dry run, no GPU)."""
import ctypes, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ccdm-stochastic-segmentation_b200")]
import torch
import bench
from ccdm_b200 import _lib

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "lidc"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else wl["B"]
m, _ = bench.build_model(wl)
eng = m.unet.engine(sys.argv[3] if len(sys.argv) > 3 else "bf16", dry_run=True)
prog = eng.program(B, wl["H"], wl["W"])
eng.weights.refresh()
prog.bind(8)
L = _lib.lib()
names = "PL R Wt MB WN NT n_cc NS res acc2 tmem tiles items grid smem chunks".split()
for i in range(prog.n_ops):
    o = prog._op_array[i]
    out = (ctypes.c_int32 * 16)()
    tag = bench.op_class(o)
    if o.kind == _lib.OP_CONV and L.ccdm_conv_tc_config(ctypes.byref(o), out) == 0:
        print(f"{i:3d} {tag:42s} " + " ".join(f"{n}={v}" for n, v in zip(names, out)))
    else:
        print(f"{i:3d} {tag:42s} (not on the tcgen05 kernel)")
