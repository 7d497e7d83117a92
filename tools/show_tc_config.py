"""Print the tile / pipeline configuration the tensor-core conv kernels pick for given shapes (host only)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ccdm-stochastic-segmentation_b200"))
from ccdm_b200 import _lib
L = _lib.lib()
names = ["PL", "R", "Wt", "MB", "NQ/WN", "NT", "n_cc", "NS", "resident", "acc2", "tmem", "tiles", "items", "grid", "smem", "chunks"]
def show(B, H, W, C0, C1, Cout, k=3, S0=0, S1=0, up=0, gn=1):
    op = _lib.Op(kind=_lib.OP_CONV, dtype=_lib.DT_BF16, out_dtype=_lib.DT_BF16, B=B, Hin=H // (2 if up else 1), Win=W // (2 if up else 1), Hout=H, Wout=W, C0=C0, C1=C1,
                 Cout=Cout, ksize=k, stride=1, S0=S0, S1=S1, upsample=up, gn=gn, silu=gn)
    out = (ctypes.c_int32 * 16)()
    rc = L.ccdm_conv_tc_config(ctypes.byref(op), out)
    tma = L.ccdm_conv_uses_tma(ctypes.byref(op))
    print(f"B{B} {H}x{W} {C0}+{C1}->{Cout} k{k} skip{S0}+{S1} up{up}: rc={rc} tma={tma} " + " ".join(f"{n}={v}" for n, v in zip(names, out)))
if __name__ == "__main__":
    for B in (64, 8):
        show(B, 128, 128, 32, 0, 32); show(B, 128, 128, 32, 32, 32); show(B, 128, 128, 32, 0, 32, S0=32, S1=32)
        show(B, 64, 64, 32, 0, 32); show(B, 64, 64, 32, 32, 32); show(B, 32, 32, 64, 0, 64); show(B, 32, 32, 64, 0, 64, S0=64, S1=64)
        show(B, 16, 16, 96, 0, 96); show(B, 8, 8, 128, 0, 128); show(B, 8, 8, 128, 0, 384, k=1); show(B, 128, 128, 32, 0, 2)
    show(8, 256, 512, 32, 0, 32); show(8, 256, 512, 32, 32, 32); show(8, 32, 64, 64, 384, 64)
