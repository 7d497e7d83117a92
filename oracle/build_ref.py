"""Compile the UNMODIFIED reference sampler into bytecode under ``oracle/_ref/`` (test infrastructure).

The reference is pure Python; it "cannot travel" to the GPU box as source (reference sources are
never copied into this repo).  What the bench's CPU arm needs is the reference's own
implementation, so -- exactly like a C reference would be compiled to ``oracle/_ref/*.so`` from
the sources where they lie -- this recipe byte-compiles the handful of modules on the hot path
from ``/root/reference`` into sourceless bytecode files (``.pycbin``: plain ``.pyc`` content under an extension that
snapshot / sync tools do not filter out):

    /root/reference/ddpm/models/{__init__,builder,diffusion_denoising,one_hot_categorical}.py
    /root/reference/ddpm/models/unet_openai/{__init__,unet,nn,fp16_util,attention}.py

``oracle/_ref/`` is git-ignored (never enters history) and not gpurun-ignored (travels with the
snapshot).  Import with ``load_reference_models()``: returns the reference's ``models`` package,
or None when ``oracle/_ref`` has not been built (the callers then fall back to the oracle port
and say so).  Runs only where ``/root/reference`` exists (the build container).
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/ddpm/models"
OUT = os.path.join(HERE, "_ref", "refmodels")
FILES = ["__init__.py", "builder.py", "diffusion_denoising.py", "one_hot_categorical.py", "unet_openai/__init__.py",
         "unet_openai/unet.py", "unet_openai/nn.py", "unet_openai/fp16_util.py", "unet_openai/attention.py"]


EXT = ".pycbin"


def build_ref() -> bool:
    if not os.path.isdir(REF_SRC):
        return False
    for rel in FILES:
        dst = os.path.join(OUT, rel[:-3] + EXT)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(os.path.join(REF_SRC, rel), cfile=dst, dfile=f"<reference>/ddpm/models/{rel}", doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    return True


class _RefFinder(importlib.abc.MetaPathFinder):
    """Resolves ``refmodels[.sub...]`` to the bytecode files under ``oracle/_ref/refmodels``."""

    def find_spec(self, fullname, path=None, target=None):
        parts = fullname.split(".")
        if parts[0] != "refmodels":
            return None
        base = os.path.join(OUT, *parts[1:])
        init = os.path.join(base, "__init__" + EXT)
        if os.path.isfile(init):
            loader = importlib.machinery.SourcelessFileLoader(fullname, init)
            return importlib.util.spec_from_file_location(fullname, init, loader=loader, submodule_search_locations=[base])
        f = base + EXT
        if os.path.isfile(f):
            return importlib.util.spec_from_file_location(fullname, f, loader=importlib.machinery.SourcelessFileLoader(fullname, f))
        return None


def load_reference_models():
    """The reference's ``ddpm.models`` package from bytecode, or None."""
    if not os.path.exists(os.path.join(OUT, "__init__" + EXT)):
        return None
    if not any(isinstance(f, _RefFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _RefFinder())
    return importlib.import_module("refmodels")


if __name__ == "__main__":
    print("built" if build_ref() else "reference sources not present; nothing built")
