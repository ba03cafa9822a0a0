"""One feature-only pass (16 kHz, 256-sample window, 39-dim MFCC+d+dd) for an ncu capture of the front-end kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aaltoasr_b200 import AkuGpu, synth
eng = AkuGpu(0)
eng.frontend_load_config_text(synth.mfcc39_config(16000))
one = synth.synth_audio(3016, 60 * 16000, 16000)
pcm = torch.from_numpy(np.tile(one, 16)).cuda()
uo = np.arange(17, dtype=np.int64) * one.size
fo = eng.frame_offsets(uo)
out = torch.empty((int(fo[-1]), 39), dtype=torch.float32, device="cuda")
for _ in range(3):
    eng.features(pcm, uo, out=out)
torch.cuda.synchronize()
eng.close()
