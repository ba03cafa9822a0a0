import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aaltoasr_b200 import AkuGpu, synth
import torch
n_utts = 120
eng = AkuGpu(0)
eng.frontend_load_config_text(synth.mfcc39_config())
base = [synth.synth_audio(2000 + i, 160000) for i in range(4)]
pcm = np.concatenate([base[i % 4] for i in range(n_utts)])
uo = np.arange(n_utts + 1, dtype=np.int64) * 160000
feats, fo = eng.features(pcm, uo, dtype=np.float32)
model = synth.synth_diag_model(2999, feats[:5000].astype(np.float64), 5000, 16)
eng.model_load_diag(model["mix_offsets"], model["mix_gauss"], model["mix_weight"], model["means"], model["covs"])
pcm_d = torch.from_numpy(pcm).cuda()
F = int(fo[-1])
for _ in range(2):
    eng.phone_probs(pcm_d, uo, lnabytes=2, discard=True)
eng.stage_times_reset(True)
eng.phone_probs(pcm_d, uo, lnabytes=2, discard=True)
st = eng.stage_times()
flops = F * 80000.0 * 40 * 4
print("AKUGPU_DBG=%s: gmm %.2f ms  %.2f TFLOP/s" % (os.environ.get("AKUGPU_DBG", "0"), st["gmm"][0], flops / st["gmm"][0] / 1e9))
