"""ctypes/numpy front end of ``ccdm_oracle.c`` (test infrastructure, see package docstring).

Arrays are pixel-major / class-minor: ``theta[p, c]`` with ``p`` running over
``B*H*W`` pixels, the order ``OneHotCategoricalBCHW`` hands to
``torch.multinomial`` (/root/reference/ddpm/models/one_hot_categorical.py:25-38).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libccdm_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ccdm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libccdm_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def cosine_schedule(T):
    b, a, c = (np.empty(T, np.float32) for _ in range(3))
    lib().ccdm_oracle_cosine_schedule(ctypes.c_int(T), _p(b, ctypes.c_float), _p(a, ctypes.c_float), _p(c, ctypes.c_float))
    return b, a, c


def linear_schedule(T, start=1e-2, end=0.2):
    b, a, c = (np.empty(T, np.float32) for _ in range(3))
    lib().ccdm_oracle_linear_schedule(ctypes.c_int(T), ctypes.c_float(start), ctypes.c_float(end),
                                      _p(b, ctypes.c_float), _p(a, ctypes.c_float), _p(c, ctypes.c_float))
    return b, a, c


def t_values(T, init_t=None):
    out = np.empty(T, np.int32)
    n = lib().ccdm_oracle_t_values(ctypes.c_int(T), ctypes.c_int(0 if init_t is None else int(init_t)), _p(out, ctypes.c_int))
    if n < 0:
        raise AssertionError("0 < K <= time_steps violated (diffusion_denoising.py:180)")
    return [int(v) for v in out[:n]]


def step_scalars(alphas, cumalphas, t):
    alphas = np.ascontiguousarray(alphas, np.float32)
    cumalphas = np.ascontiguousarray(cumalphas, np.float32)
    a, c = ctypes.c_float(), ctypes.c_float()
    lib().ccdm_oracle_step_scalars(_p(alphas, ctypes.c_float), _p(cumalphas, ctypes.c_float), ctypes.c_int(int(t)),
                                   ctypes.byref(a), ctypes.byref(c))
    return np.float32(a.value), np.float32(c.value)


def _posterior(fn, xt_label, theta, alpha_t, cumalpha_tm1):
    theta = np.ascontiguousarray(theta, np.float32)
    K = theta.shape[-1]
    lab = np.ascontiguousarray(xt_label, np.uint8).reshape(-1)
    n = lab.size
    assert theta.size == n * K
    out = np.empty_like(theta)
    fn(_p(lab, ctypes.c_uint8), _p(theta, ctypes.c_float), ctypes.c_size_t(n), ctypes.c_int(K),
       ctypes.c_float(alpha_t), ctypes.c_float(cumalpha_tm1), _p(out, ctypes.c_float))
    return out


def posterior_literal(xt_label, theta, alpha_t, cumalpha_tm1):
    """O(K^2) form, diffusion_denoising.py:115-128."""
    return _posterior(lib().ccdm_oracle_posterior_literal, xt_label, theta, alpha_t, cumalpha_tm1)


def posterior_closed(xt_label, theta, alpha_t, cumalpha_tm1):
    """O(K) closed form; the operation order the CUDA kernel follows."""
    return _posterior(lib().ccdm_oracle_posterior_closed, xt_label, theta, alpha_t, cumalpha_tm1)


def softmax(logits):
    logits = np.ascontiguousarray(logits, np.float32)
    K = logits.shape[-1]
    out = np.empty_like(logits)
    lib().ccdm_oracle_softmax(_p(logits, ctypes.c_float), ctypes.c_size_t(logits.size // K), ctypes.c_int(K), _p(out, ctypes.c_float))
    return out


def draw(post, noise, mode):
    """mode 0 sample / 1 majority / 2 confidence -> (labels uint8, normalised probs)."""
    post = np.ascontiguousarray(post, np.float32)
    K = post.shape[-1]
    n = post.size // K
    labels = np.empty(post.shape[:-1], np.uint8)
    probs = np.empty_like(post)
    if mode == 0:
        noise = np.ascontiguousarray(noise, np.float32)
        assert noise.size == post.size
        npt = _p(noise, ctypes.c_float)
    else:
        npt = None
    lib().ccdm_oracle_draw(_p(post, ctypes.c_float), npt, ctypes.c_size_t(n), ctypes.c_int(K), ctypes.c_int(mode),
                           _p(labels, ctypes.c_uint8), _p(probs, ctypes.c_float))
    return labels, probs


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, np.uint32).copy()
    k = np.asarray(key, np.uint32).copy()
    o = np.empty(4, np.uint32)
    lib().ccdm_oracle_philox4x32_10(_p(c, ctypes.c_uint32), _p(k, ctypes.c_uint32), _p(o, ctypes.c_uint32))
    return o


def philox_bits(seed, draw_idx, sample0, n_samples, n_pix, K):
    bits = np.empty((n_samples, n_pix, K), np.uint32)
    lib().ccdm_oracle_philox_bits(ctypes.c_uint64(seed), ctypes.c_uint32(draw_idx), ctypes.c_uint32(sample0),
                                  ctypes.c_uint32(n_samples), ctypes.c_uint32(n_pix), ctypes.c_int(K), _p(bits, ctypes.c_uint32))
    return bits


def bits_to_exponential(bits):
    bits = np.ascontiguousarray(bits, np.uint32)
    e = np.empty(bits.shape, np.float32)
    lib().ccdm_oracle_bits_to_exponential(_p(bits, ctypes.c_uint32), ctypes.c_size_t(bits.size), _p(e, ctypes.c_float))
    return e


def uniform_labels(noise):
    noise = np.ascontiguousarray(noise, np.float32)
    K = noise.shape[-1]
    out = np.empty(noise.shape[:-1], np.uint8)
    lib().ccdm_oracle_uniform_labels(_p(noise, ctypes.c_float), ctypes.c_size_t(noise.size // K), ctypes.c_int(K), _p(out, ctypes.c_uint8))
    return out
