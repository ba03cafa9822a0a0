"""Quick on-GPU probe: pipe rates + per-stage timings of the config-2 workload (subset)."""
import json
import sys
import os
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aaltoasr_b200 import AkuGpu, F32, F64, synth  # noqa: E402

n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 100
eng = AkuGpu(0)
rates = eng.pipe_rates()
print("pipe rates (lane-ops/s):", json.dumps({k: "%.3e" % v for k, v in rates.items()}))
cfg = synth.mfcc39_config()
eng.frontend_load_config_text(cfg)
base = [synth.synth_audio(2000 + i, 160000) for i in range(10)]
pcm = np.concatenate([base[i % 10] for i in range(n_utts)])
uo = np.arange(n_utts + 1, dtype=np.int64) * 160000
t0 = time.time()
feats, fo = eng.features(pcm, uo, dtype=np.float32)
print("features: %d frames in %.3f s (first call)" % (fo[-1], time.time() - t0))
model = synth.synth_diag_model(2999, feats[:20000].astype(np.float64), 5000, 16)
import torch
pcm_d = torch.from_numpy(pcm).cuda()
F = int(fo[-1])
out_d = torch.empty((F, 5000 * 2), dtype=torch.uint8, device="cuda")
for variant in (1, 2):
    eng.set_scorer_variant(variant)
    eng.model_load_diag(model["mix_offsets"], model["mix_gauss"], model["mix_weight"], model["means"], model["covs"])
    for chunk in (16384, 0, 75776):
        eng.set_chunk_frames(chunk)
        eng.phone_probs(pcm_d, uo, lnabytes=2, out=out_d)
        eng.stage_times_reset(True)
        torch.cuda.synchronize()
        t0 = time.time()
        eng.phone_probs(pcm_d, uo, lnabytes=2, out=out_d)
        torch.cuda.synchronize()
        dt = time.time() - t0
        st = eng.stage_times()
        eng.stage_times_reset(False)
        gmm_ms = st["gmm"][0]
        flops = F * 80000.0 * 40 * 2 * 2   # 2 FMA per (frame, comp, dim incl. pad), 2 flop each
        print("variant %d chunk %6d: wall %.3f s = %.0f frames/s | fe %.1f ms gmm %.1f ms lna %.1f ms | gmm %.2f TFLOP/s, %.0f frames/s"
              % (variant, chunk, dt, F / dt, st["frontend"][0], gmm_ms, st["lna"][0], flops / gmm_ms / 1e9, F / gmm_ms * 1e3))
eng.set_scorer_variant(0)
eng.model_load_diag(model["mix_offsets"], model["mix_gauss"], model["mix_weight"], model["means"], model["covs"])
eng.set_chunk_frames(0)
# parity mode throughput on a slice
t0 = time.time()
eng.gmm_lna(feats[:4096], precision=F64, lnabytes=2)
print("F64 parity mode: 4096 frames in %.3f s = %.0f frames/s" % (time.time() - t0, 4096 / (time.time() - t0)))
# e2e with pinned host buffers
pcm_p = torch.from_numpy(pcm).pin_memory()
out_p = torch.empty((F, 5000 * 2), dtype=torch.uint8).pin_memory()
eng.phone_probs(pcm_p, uo, lnabytes=2, out=out_p)
t0 = time.time()
eng.phone_probs(pcm_p, uo, lnabytes=2, out=out_p)
dt = time.time() - t0
print("e2e pinned host: %.3f s = %.0f frames/s, D2H %.1f GB/s" % (dt, F / dt, out_p.numel() / dt / 1e9))
