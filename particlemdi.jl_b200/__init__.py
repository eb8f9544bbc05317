"""pmdi-b200: B200-native conditional-SMC allocation sweep behind ParticleMDI's `pmdi()`.

Host-side mirror of the reference interface (src/ParticleMDI.jl:31-36 exports) for the hot
path only; the compute lives in ``csrc/`` behind the C-ABI of ``include/pmdi_cuda.h``.
"""
from . import dataprep, synth  # noqa: F401
from .dataprep import coerce_categorical, gaussian_normalise  # noqa: F401
