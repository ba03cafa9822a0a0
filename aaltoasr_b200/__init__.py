"""aaltoasr_b200 -- B200-native acoustic front-end for AaltoASR.

Host-side mirror (Python flavour, for tests / bench / scripting) of the reference's
FeatureGenerator / HmmSet / phone_probs surface on top of the C ABI in
``include/akugpu.h`` (``libakugpu.so``).  The compute lives in the CUDA library; this
package never computes features or likelihoods itself and raises if the library or a
CUDA device is missing -- there is no CPU fallback.
"""
from ._lib import AkuGpuError, load_library, library_path
from .engine import AkuGpu, F32, F64
from .hostapi import FeatureGenerator, HmmSet, PhoneProbs, SpeakerConfig, StreamSession, parse_speaker_file

PPToolbox = PhoneProbs      # the name of the reference's SWIG class (aku/swig/PPToolbox.i:59-66)

__all__ = ["AkuGpu", "AkuGpuError", "F32", "F64", "FeatureGenerator", "HmmSet", "PhoneProbs", "PPToolbox", "SpeakerConfig", "StreamSession", "parse_speaker_file",
           "load_library", "library_path"]
