"""Small driver for ncu captures: a few launches of the scorer at the config-2 model size."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aaltoasr_b200 import AkuGpu, synth  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n_utts = int(sys.argv[2]) if len(sys.argv) > 2 else 27
eng = AkuGpu(0)
eng.frontend_load_config_text(synth.mfcc39_config())
base = [synth.synth_audio(2000 + i, 160000) for i in range(3)]
pcm = np.concatenate([base[i % 3] for i in range(n_utts)])
uo = np.arange(n_utts + 1, dtype=np.int64) * 160000
feats, fo = eng.features(pcm, uo, dtype=np.float32)
model = synth.synth_diag_model(2999, feats[:3000].astype(np.float64), 5000, 16)
eng.set_scorer_variant(variant)
eng.model_load_diag(model["mix_offsets"], model["mix_gauss"], model["mix_weight"], model["means"], model["covs"])
eng.set_chunk_frames(0)
for _ in range(3):
    eng.phone_probs(pcm, uo, lnabytes=2, discard=True)
print("done", fo[-1], "frames")
