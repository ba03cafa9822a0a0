"""Diagnostic: where do ~34 ms per call go in the front-end stage of config-4 style calls (30 jittered utterances)?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aaltoasr_b200 import AkuGpu, synth

eng = AkuGpu(0)
stream = torch.cuda.Stream(); eng.set_stream(stream.cuda_stream)
eng.frontend_load_config_text(synth.mfcc39_config(16000))
clip = synth.synth_audio(1, 240000, 16000)
rng = np.random.default_rng(0)

def run(tag, lens, offset, reps=6, out_dev=True):
    pcm = np.concatenate([clip[:n] for n in lens])
    uo = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    big = torch.zeros(offset + pcm.size + 8, dtype=torch.int16, device="cuda")
    big[offset:offset + pcm.size] = torch.from_numpy(pcm).cuda()
    view = big[offset:offset + pcm.size]
    fo = eng.frame_offsets(uo)
    out = torch.empty((int(fo[-1]), 39), dtype=torch.float32, device="cuda")
    ts = []
    for r in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.features(view, uo, out=out)
        torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    print("%-40s frames %6d  ms per call: %s" % (tag, int(fo[-1]), " ".join("%.2f" % t for t in ts)), flush=True)

run("30 x 10 s, offset 0", [160000] * 30, 0)
run("30 x 10 s, offset 1 (odd)", [160000] * 30, 1)
run("30 x 10 s, offset 12345", [160000] * 30, 12345)
lens = (rng.uniform(5, 15, 30) * 16000).astype(np.int64)
run("30 jittered, offset 0", lens, 0)
run("30 jittered, offset 777", lens, 777)
lens2 = (rng.uniform(5, 15, 30) * 16000).astype(np.int64)
run("30 other jittered", lens2, 0)
run("1000 x 10 s", [160000] * 1000, 0, reps=3)
run("120 jittered", (rng.uniform(5, 15, 120) * 16000).astype(np.int64), 0)
eng.close()
