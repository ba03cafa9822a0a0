"""Streaming regime: latency of scoring a small tile of frames -- what a decoder that calls
HmmSet::precompute_likelihoods per frame (aku/HmmSet.cc:485-501, decoder/decode-stream.cc:178-207) would see -- and the
parameter-sweep bandwidth it implies (every call reads the whole packed model: L2-resident after the first call)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aaltoasr_b200 import AkuGpu, F32, F64, synth
import torch

eng = AkuGpu(0)
stream = torch.cuda.Stream()
eng.set_stream(stream.cuda_stream)
eng.frontend_load_config_text(synth.mfcc39_config())
pcm = np.concatenate([synth.synth_audio(2000 + i, 160000) for i in range(4)])
uo = np.arange(5, dtype=np.int64) * 160000
feats, fo = eng.features(pcm, uo, dtype=np.float32)
rows = []
for S, M in ((5000, 16), (10000, 32)):
    model = synth.synth_diag_model(2999, feats.astype(np.float64), S, M)
    eng.model_load_diag(model["mix_offsets"], model["mix_gauss"], model["mix_weight"], model["means"], model["covs"])
    G = S * M
    param_bytes = G * 192 * 2          # packed fp16 image B' = [Bh | Bl], 3 k-blocks of 64
    for F in (1, 8, 32, 128, 512, 2048):
        fd = torch.from_numpy(feats[:F].copy()).cuda()
        out = torch.empty((F, S), dtype=torch.float32, device="cuda")
        for _ in range(5):
            eng.gmm_score(fd, precision=F32, out=out)
        torch.cuda.synchronize()
        reps = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(reps):
                eng.gmm_score(fd, precision=F32, out=out)
            e1.record(stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / reps
        dev = e0.elapsed_time(e1) / reps * 1e-3
        rows.append({"states": S, "mix": M, "frames_per_call": F, "device_us_per_call": dev * 1e6, "wall_us_per_call": wall * 1e6,
                     "frames_per_s": F / wall, "param_sweep_GBps": param_bytes / dev / 1e9})
        print("%5d x %2d  F=%4d: %8.1f us/call device, %8.1f us wall, %10.0f frames/s, parameter sweep %7.1f GB/s" % (
            S, M, F, dev * 1e6, wall * 1e6, F / wall, param_bytes / dev / 1e9), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/stream_bench.json", "w"), indent=1)
