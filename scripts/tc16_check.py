"""Checks the fp16x2 tensor-core scorer (default) against the oracle fixture, the F64 path and the other F32 kernels."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aaltoasr_b200 import AkuGpu, F32, F64, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def load(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["model"] = {k[6:]: d[k] for k in list(d) if k.startswith("model_")}
    return d

eng = AkuGpu(0)
g = load("ref_small")
m = g["model"]
want = np.log(g["lik"])
for variant in (0, 4, 2):
    eng.set_scorer_variant(variant)
    eng.model_load_diag(m["mix_offsets"], m["mix_gauss"], m["mix_weight"], m["means"], m["covs"])
    ll = eng.gmm_score(g["feats"].astype(np.float32), precision=F32).astype(np.float64)
    err = np.abs(ll - want)
    print("ref_small variant %d: max abs err %.3e  mean %.3e  (ll range %.1f..%.1f)" % (variant, err.max(), err.mean(), want.min(), want.max()), flush=True)

import torch
eng.frontend_load_config_text(synth.mfcc39_config())
base = [synth.synth_audio(2000 + i, 160000) for i in range(8)]
n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 120
pcm = np.concatenate([base[i % 8] for i in range(n_utts)])
uo = np.arange(n_utts + 1, dtype=np.int64) * 160000
feats, fo = eng.features(pcm, uo, dtype=np.float32)
F = int(fo[-1])
feats_d = torch.from_numpy(feats).cuda()
model = synth.synth_diag_model(2999, feats[:20000].astype(np.float64), 5000, 16)
eng.set_scorer_variant(0)
eng.model_load_diag(model["mix_offsets"], model["mix_gauss"], model["mix_weight"], model["means"], model["covs"])
nchk = 2048
want = np.log(eng.gmm_score(feats[:nchk].astype(np.float64), precision=F64))
outs = {}
for variant in (0, 4, 2):
    eng.set_scorer_variant(variant)
    eng.model_load_diag(model["mix_offsets"], model["mix_gauss"], model["mix_weight"], model["means"], model["covs"])
    ll = eng.gmm_score(feats[:nchk], precision=F32).astype(np.float64)
    ok = want > -80
    err = np.abs(ll - want)[ok]
    print("5000x16 variant %d: max abs err %.3e  mean %.3e  rms %.3e" % (variant, err.max(), err.mean(), np.sqrt((err**2).mean())), flush=True)
    out = torch.empty((F, 5000 * 2), dtype=torch.uint8, device="cuda")
    eng.gmm_lna(feats_d, lnabytes=2, out=out)
    torch.cuda.synchronize(); t0 = time.time()
    eng.stage_times_reset(True)
    eng.gmm_lna(feats_d, lnabytes=2, out=out)
    torch.cuda.synchronize(); dt = time.time() - t0
    st = eng.stage_times(); eng.stage_times_reset(False)
    print("5000x16 variant %d: %.1f ms for %d frames, %.2f M frames/s; stage ms gmm %.1f (%d launches) lna %.1f" % (
        variant, dt * 1e3, F, F / dt / 1e6, st["gmm"][0], st["gmm"][1], st["lna"][0]), flush=True)
    outs[variant] = out[:200000].cpu().numpy().view(">u2").astype(np.int64)
    del out
for v in (0, 4):
    a, b = outs[v], outs[2]
    print("codes variant %d vs FP32-pipe kernel: max |diff| %d, differing %.3f%%" % (v, np.abs(a - b).max(), 100 * (a != b).mean()))
# range overflow: a frame far outside the model must fall back (and still give finite, correct-order results)
bad = feats[:256].copy(); bad[7, 3] = 3.0e6
eng.set_scorer_variant(0)
eng.model_load_diag(model["mix_offsets"], model["mix_gauss"], model["mix_weight"], model["means"], model["covs"])
l0 = eng.launch_count()
ll = eng.gmm_score(bad, precision=F32)
print("overflow frame: launches %d, finite %s, row7 max %.3e; other rows max diff vs clean %.2e" % (
    eng.launch_count() - l0, np.isfinite(ll).all(), ll[7].max(), np.abs(np.delete(ll, 7, 0) - np.delete(eng.gmm_score(feats[:256], precision=F32), 7, 0)).max()))
