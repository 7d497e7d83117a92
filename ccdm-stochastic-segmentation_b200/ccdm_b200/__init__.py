"""ccdm_b200 -- B200-native categorical-diffusion sampler (the CCDM reverse-process hot path).

Public surface mirrors the reference's ``ddpm.models`` package:
``ccdm_b200.models.build_model`` / ``DenoisingModel`` / ``UNetModel`` /
``OneHotCategoricalBCHW``.  All arithmetic runs in ``libccdm_b200.so``
(hand-written sm_100a CUDA behind the C ABI of ``include/ccdm_b200.h``).
"""
__version__ = "0.1.0"
