"""moonshine_b200 -- B200-native (sm_100a) implementation of Moonshine's
encoder-decoder transcription loop behind the reference's C ABI.

The product is ``moonshine_b200/lib/libmoonshine.so`` (C++/CUDA, no torch, no
ORT).  This Python package is the host-side mirror of the reference's Python
binding for that path (``language-bindings/python/src/moonshine_voice``):
``Transcriber`` / ``ModelArch`` / ``TranscriptLine`` over ctypes, plus the
weight-container tooling (``weights.py``).  There is no CPU fallback: loading
fails loudly if the CUDA library is missing.
"""
from .arch import ARCHS, ModelDims, dims_for_arch  # noqa: F401
