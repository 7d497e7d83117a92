"""ctypes binding of ``libccdm_b200.so`` (C ABI: ``include/ccdm_b200.h``).

The library is the product: there is no Python/PyTorch fallback.  Importing this
module without the built library raises; calling into it without a B200 raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CCDM_B200_LIB: another build of the same library (A/B timing of two builds on one GPU box)
LIB_PATH = os.environ.get("CCDM_B200_LIB") or os.path.join(_HERE, "libccdm_b200.so")

# constants of include/ccdm_b200.h
DT_F32, DT_BF16, DT_F16X2 = 0, 1, 2
F16X2_SCALE_LOG2 = 4
DRAW_SAMPLE, DRAW_MAJORITY, DRAW_CONFIDENCE, DRAW_X0, DRAW_POSTERIOR = 0, 1, 2, 3, 4
NOISE_TENSOR, NOISE_PHILOX = 0, 1
OP_INPUT_CONV, OP_CONV, OP_ATTENTION, OP_HEAD, OP_ENCODE_INPUT = 1, 2, 3, 4, 5
ABI_VERSION = 4


class StepEntry(ctypes.Structure):
    _fields_ = [("t", ctypes.c_float), ("alpha_t", ctypes.c_float), ("cumalpha_tm1", ctypes.c_float),
                ("mode", ctypes.c_int32), ("draw", ctypes.c_uint32), ("emb_row", ctypes.c_int32),
                ("pad0", ctypes.c_int32), ("pad1", ctypes.c_int32)]


_I32 = ["kind", "dtype", "B", "Hin", "Win", "Hout", "Wout", "C0", "C1", "Cout", "ksize", "stride", "upsample", "gn", "silu",
        "S0", "S1", "heads", "head_dim", "K", "C_img", "emb_off", "emb_cols", "emb_bstride", "noise_mode", "sample0",
        "out_dtype", "src_kind", "exact", "acc_shift", "img_rep", "st_slots0", "st_slots1", "st_ips0", "st_ips1", "st_items0", "st_items1",
        "st_grid0", "st_grid1", "st_rows0", "st_rows1", "tile_batch", "gn_cpg", "gn_off"]
_U64 = ["seed", "src0", "src1", "stat0", "stat1", "gamma", "beta", "weight", "bias", "emb", "skip0", "skip1", "skip_w", "res",
        "out", "ostat", "part", "ticket", "labels_in", "labels_out", "image", "noise", "probs_out", "noise_out", "steps",
        "step_ptr"]


class Op(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in _I32] + [(n, ctypes.c_uint64) for n in _U64]

    def __init__(self, **kw):
        super().__init__()
        self.emb_off = -1
        for k, v in kw.items():
            if k not in _I32 and k not in _U64:
                raise AttributeError(f"ccdm_op has no field {k}")
            setattr(self, k, int(v))


class CcdmError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CcdmError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a).  ccdm_b200 has no CPU or PyTorch fallback.")
    L = ctypes.CDLL(LIB_PATH)
    L.ccdm_last_error.restype = ctypes.c_char_p
    L.ccdm_plan_create.restype = ctypes.c_void_p
    L.ccdm_plan_create.argtypes = [ctypes.POINTER(Op), ctypes.c_int]
    L.ccdm_plan_destroy.argtypes = [ctypes.c_void_p]
    L.ccdm_plan_destroy.restype = None
    L.ccdm_plan_num_launches.argtypes = [ctypes.c_void_p]
    L.ccdm_plan_set_noise.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
    L.ccdm_plan_step.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.ccdm_plan_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.ccdm_plan_profile.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_void_p]
    L.ccdm_debug_conv_trace.argtypes = [ctypes.POINTER(ctypes.c_uint64), ctypes.c_int]
    L.ccdm_launch_op.argtypes = [ctypes.POINTER(Op), ctypes.c_void_p]
    L.ccdm_sizeof_op.restype = ctypes.c_size_t
    L.ccdm_sizeof_step_entry.restype = ctypes.c_size_t
    L.ccdm_time_table.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 6 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    L.ccdm_onehot_to_labels.argtypes = [ctypes.c_void_p] + [ctypes.c_int64] * 4 + [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_void_p]
    L.ccdm_labels_to_onehot_i64.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    L.ccdm_nchw_to_nhwc_stats.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 5 + [ctypes.c_void_p] * 2 + [ctypes.c_int, ctypes.c_void_p]
    L.ccdm_pairwise_distance.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_size_t,
                                         ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    L.ccdm_vote.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                            ctypes.c_void_p]
    L.ccdm_posterior_draw.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                      ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32,
                                      ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.ccdm_philox_bits.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int,
                                   ctypes.c_void_p, ctypes.c_void_p]
    L.ccdm_uniform_labels.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_void_p]
    L.ccdm_op_part_floats.restype = ctypes.c_size_t
    L.ccdm_op_part_floats.argtypes = [ctypes.POINTER(Op)]
    L.ccdm_conv_uses_tc.argtypes = [ctypes.POINTER(Op)]
    L.ccdm_conv_uses_tma.argtypes = [ctypes.POINTER(Op)]
    L.ccdm_conv_stat_layout.argtypes = [ctypes.POINTER(Op), ctypes.POINTER(ctypes.c_int32)]
    L.ccdm_conv_tc_nt.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.ccdm_conv_tc_config.argtypes = [ctypes.POINTER(Op), ctypes.POINTER(ctypes.c_int32)]
    L.ccdm_conv_part_floats.restype = ctypes.c_size_t
    L.ccdm_conv_part_floats.argtypes = [ctypes.c_int] * 4
    VP, CI = ctypes.c_void_p, ctypes.c_int
    L.ccdm_vit_linear.argtypes = [VP, VP, VP, VP, CI, CI, CI, CI, CI, CI, VP, VP]
    L.ccdm_vit_layernorm.argtypes = [VP, VP, VP, CI, CI, CI, ctypes.c_float, VP, VP]
    L.ccdm_vit_patch_embed.argtypes = [VP, VP, VP, VP, VP, CI, CI, CI, CI, CI, CI, VP, VP]
    L.ccdm_vit_pos_embed.argtypes = [VP, CI, CI, CI, CI, ctypes.c_double, ctypes.c_double, VP, VP]
    L.ccdm_vit_descriptor.argtypes = [VP, CI, CI, CI, CI, CI, CI, CI, CI, VP, VP]
    if L.ccdm_abi_version() != ABI_VERSION:
        raise CcdmError(f"ABI mismatch: library {L.ccdm_abi_version()} vs binding {ABI_VERSION}; rebuild")
    if L.ccdm_sizeof_op() != ctypes.sizeof(Op) or L.ccdm_sizeof_step_entry() != ctypes.sizeof(StepEntry):
        raise CcdmError("ccdm_op / ccdm_step_entry layout differs between the header and _lib.py")
    _lib = L
    return L


def check(rc: int, what: str = ""):
    if rc != 0:
        raise CcdmError(f"{what or 'libccdm_b200'} failed ({rc}): {lib().ccdm_last_error().decode(errors='replace')}")


def require_device():
    """Raise unless a B200-class (sm_10x) device is current."""
    import torch
    if not torch.cuda.is_available():
        raise CcdmError("ccdm_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    check(lib().ccdm_check_device(), "ccdm_check_device")


def stream_ptr(stream) -> ctypes.c_void_p:
    return ctypes.c_void_p(stream.cuda_stream)
