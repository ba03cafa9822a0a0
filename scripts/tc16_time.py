"""Times the default scorer (gmm_lna on resident features) at the config-2 model; AKUGPU_DBG toggles kernel experiments."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aaltoasr_b200 import AkuGpu, F32, F64, synth
import torch
eng = AkuGpu(0)
eng.frontend_load_config_text(synth.mfcc39_config())
base = [synth.synth_audio(2000 + i, 160000) for i in range(8)]
n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 120
S, M = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (5000, 16)
pcm = np.concatenate([base[i % 8] for i in range(n_utts)])
uo = np.arange(n_utts + 1, dtype=np.int64) * 160000
feats, fo = eng.features(pcm, uo, dtype=np.float32)
F = int(fo[-1])
feats_d = torch.from_numpy(feats).cuda()
model = synth.synth_diag_model(2999, feats[:20000].astype(np.float64), S, M)
eng.model_load_diag(model["mix_offsets"], model["mix_gauss"], model["mix_weight"], model["means"], model["covs"])
out = torch.empty((F, S * 2), dtype=torch.uint8, device="cuda")
eng.gmm_lna(feats_d, lnabytes=2, out=out)
torch.cuda.synchronize()
eng.stage_times_reset(True)
t0 = time.time()
eng.gmm_lna(feats_d, lnabytes=2, out=out)
torch.cuda.synchronize(); dt = time.time() - t0
st = eng.stage_times(); eng.stage_times_reset(False)
print("DBG=%s %dx%d: %.1f ms for %d frames, %.2f M frames/s; gmm %.2f ms (%d launches) lna %.2f ms" % (
    os.environ.get("AKUGPU_DBG", "0"), S, M, dt * 1e3, F, F / dt / 1e6, st["gmm"][0], st["gmm"][1], st["lna"][0]), flush=True)
