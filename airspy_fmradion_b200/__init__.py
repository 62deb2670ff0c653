"""B200-native streaming demodulation hot path of airspy-fmradion (FM broadcast / AM / narrow-band FM),
behind the reference's FmDecoder::process / AmDecoder::process / NbfmDecoder::process block API.

csrc/        hand-written sm_100a CUDA kernels + the C ABI (include/fmradion_b200.h)
decoder.py   host-side mirror of the reference's decoder classes over that ABI
"""
from .decoder import AmDecoder, FmDecoder, NbfmDecoder  # noqa: F401
from ._capi import FmrError  # noqa: F401
