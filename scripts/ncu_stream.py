"""A few single-frame calls of the streaming scorer at the config-2 model for an ncu capture of gmm_stream_kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aaltoasr_b200 import AkuGpu, F32, synth
eng = AkuGpu(0)
eng.frontend_load_config_text(synth.mfcc39_config(16000))
pcm = np.concatenate([synth.synth_audio(2000 + i, 160000) for i in range(2)])
feats, _ = eng.features(pcm, np.array([0, 160000, 320000]), dtype=np.float64)
S, M = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (5000, 16)
m = synth.synth_diag_model(2999, feats, S, M)
eng.model_load_diag(m["mix_offsets"], m["mix_gauss"], m["mix_weight"], m["means"], m["covs"])
x = feats[100:101].astype(np.float32)
for _ in range(6):
    eng.gmm_logprobs(x, precision=F32, tiny=1e-30)
eng.close()
