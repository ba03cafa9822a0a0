"""Checks the experimental tensor-core scorer (variant 3) against the oracle fixtures; run under `timeout`."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aaltoasr_b200 import AkuGpu, F32, F64, synth

def load(name):
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["model"] = {k[6:]: d[k] for k in list(d) if k.startswith("model_")}
    return d

eng = AkuGpu(0)
g = load("ref_small")
m = g["model"]
eng.set_scorer_variant(3)
eng.model_load_diag(m["mix_offsets"], m["mix_gauss"], m["mix_weight"], m["means"], m["covs"])
ll = eng.gmm_score(g["feats"].astype(np.float32), precision=F32).astype(np.float64)
want = np.log(g["lik"])
err = np.abs(ll - want)
print("diag ref_small: max abs err %.3e  mean %.3e  (ll range %.1f..%.1f)" % (err.max(), err.mean(), want.min(), want.max()))
g = load("ref_full")
m = g["model"]
idx = np.nonzero(m["full_mask"])[0]
sub_off = np.arange(0, len(idx) + 1, 3, dtype=np.int32)
eng.model_load_full(sub_off, np.arange(len(idx), dtype=np.int32), np.ones(len(idx)), m["means"][idx], m["full_covs"][idx])
ll = eng.gmm_score(g["feats"].astype(np.float32), precision=F32).astype(np.float64)
eng.set_scorer_variant(0)
eng.model_load_full(sub_off, np.arange(len(idx), dtype=np.int32), np.ones(len(idx)), m["means"][idx], m["full_covs"][idx])
want = np.log(eng.gmm_score(g["feats"], precision=F64))
err = np.abs(ll - want)
print("full ref_full: max abs err %.3e  mean %.3e  (ll range %.1f..%.1f)" % (err.max(), err.mean(), want.min(), want.max()))
if len(sys.argv) > 1:
    # throughput at the config-2 model (diag) and config-5 model (full)
    import torch
    eng.frontend_load_config_text(synth.mfcc39_config())
    base = [synth.synth_audio(2000 + i, 160000) for i in range(4)]
    n_utts = 60
    pcm = np.concatenate([base[i % 4] for i in range(n_utts)])
    uo = np.arange(n_utts + 1, dtype=np.int64) * 160000
    feats, fo = eng.features(pcm, uo, dtype=np.float32)
    F = int(fo[-1])
    feats_d = torch.from_numpy(feats).cuda()
    model = synth.synth_diag_model(2999, feats[:5000].astype(np.float64), 5000, 16)
    out = torch.empty((F, 5000 * 2), dtype=torch.uint8, device="cuda")
    for variant in (0, 3):
        eng.set_scorer_variant(variant)
        eng.model_load_diag(model["mix_offsets"], model["mix_gauss"], model["mix_weight"], model["means"], model["covs"])
        eng.gmm_lna(feats_d, lnabytes=2, out=out)
        torch.cuda.synchronize(); t0 = time.time()
        eng.gmm_lna(feats_d, lnabytes=2, out=out)
        torch.cuda.synchronize(); dt = time.time() - t0
        if variant == 0:
            ref_out = out.clone()
        else:
            d = (out.view(torch.int16).int() - ref_out.view(torch.int16).int())
            a = out.cpu().numpy().view(">u2").astype(np.int64); b = ref_out.cpu().numpy().view(">u2").astype(np.int64)
            print("   codes vs fp32 kernel: max |diff| %d, differing %.3f%%" % (np.abs(a - b).max(), 100 * (a != b).mean()))
        print("diag 5000x16 variant %d: %.1f ms, %.0f frames/s" % (variant, dt * 1e3, F / dt))
    # config 5: 2000 states x 16 full-covariance
    rng = np.random.default_rng(5999)
    S, M, D = 2000, 16, 39
    G = S * M
    sd = feats[:5000].astype(np.float64).std(axis=0)
    means = feats[rng.integers(0, 5000, G)].astype(np.float64) + 0.3 * sd * rng.standard_normal((G, D))
    A = rng.standard_normal((G, D, 4)) * sd[None, :, None]
    full = np.einsum("gik,gjk->gij", A, A) * 0.1
    full[:, np.arange(D), np.arange(D)] += rng.uniform(0.5, 2, (G, D)) * sd ** 2
    off = np.arange(0, G + 1, M, dtype=np.int32)
    w = rng.dirichlet(np.ones(M), S).reshape(-1)
    Fs = 18944
    out5 = torch.empty((Fs, S * 2), dtype=torch.uint8, device="cuda")
    for variant in (3, 0):
        eng.set_scorer_variant(variant)
        t0 = time.time()
        eng.model_load_full(off, np.arange(G, dtype=np.int32), w, means, full)
        print("   full model pack (variant %d): %.1f s" % (variant, time.time() - t0))
        nfr = Fs if variant == 3 else 2048
        eng.gmm_lna(feats_d[:nfr], lnabytes=2, out=out5[:nfr])
        torch.cuda.synchronize(); t0 = time.time()
        eng.gmm_lna(feats_d[:nfr], lnabytes=2, out=out5[:nfr])
        torch.cuda.synchronize(); dt = time.time() - t0
        print("full 2000x16 variant %d: %d frames in %.1f ms, %.0f frames/s" % (variant, nfr, dt * 1e3, nfr / dt))
