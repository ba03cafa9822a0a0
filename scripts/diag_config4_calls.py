"""Diagnostic: per-call wall / stage times of config-4 style sub-batch calls (10000 x 32, device PCM slices)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aaltoasr_b200 import AkuGpu, F32, synth, multigpu as mg

eng = AkuGpu(0)
stream = torch.cuda.Stream(); eng.set_stream(stream.cuda_stream)
eng.frontend_load_config_text(synth.mfcc39_config(16000))
clips = [synth.synth_audio(4000 + i, 240000, 16000) for i in range(8)]
feats, _ = eng.features(np.concatenate(clips), np.arange(9, dtype=np.int64) * 240000, dtype=np.float64)
S, M = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (10000, 32)
model = synth.synth_diag_model(4999, feats, S, M)
eng.model_load_diag(model["mix_offsets"], model["mix_gauss"], model["mix_weight"], model["means"], model["covs"])
rng = np.random.default_rng(4000)
lens = (rng.uniform(5, 15, 150) * 16000).astype(np.int64)
pcm = np.concatenate([clips[i % 8][:n] for i, n in enumerate(lens)])
uo = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
nfr = np.array([eng.num_frames(int(n)) for n in lens])
pcm_d = torch.from_numpy(pcm).cuda()
sched = mg.sub_batches(nfr, 37888)
slots = [torch.empty(37888 * S * 2, dtype=torch.uint8, device="cuda") for _ in range(2)]
for rep in range(2):
    for k, (u0, u1, f0, n) in enumerate(sched):
        a, b = int(uo[u0]), int(uo[u1])
        eng.stage_times_reset(True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.phone_probs(pcm_d[a:b], uo[u0:u1 + 1] - uo[u0], precision=F32, lnabytes=2, out=slots[k & 1])
        torch.cuda.synchronize(); wall = 1e3 * (time.perf_counter() - t0)
        st = eng.stage_times()
        print("rep %d call %d: %d utts %d frames  wall %.2f ms  stages fe %.2f gmm %.2f lna %.2f" % (
            rep, k, u1 - u0, n, wall, st["frontend"][0], st["gmm"][0], st["lna"][0]), flush=True)
eng.stage_times_reset(False)
t0 = time.perf_counter()
for k, (u0, u1, f0, n) in enumerate(sched):
    a, b = int(uo[u0]), int(uo[u1])
    eng.phone_probs(pcm_d[a:b], uo[u0:u1 + 1] - uo[u0], precision=F32, lnabytes=2, out=slots[k & 1])
torch.cuda.synchronize()
print("timers off: %.2f ms for %d frames" % (1e3 * (time.perf_counter() - t0), int(nfr.sum())))
eng.close()
