"""gmm_tc16_kernel with B' multicast over clusters of frame tiles: the scores must not change by a bit.
usage: python scripts/tc16_cluster_check.py          (runs itself once per AKUGPU_TC16_CLUSTER setting, compares the dumps)"""
import os, subprocess, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

if len(sys.argv) > 1:
    from aaltoasr_b200 import AkuGpu, F32, synth
    eng = AkuGpu(0)
    eng.frontend_load_config_text(synth.mfcc39_config(16000))
    pcm = np.concatenate([synth.synth_audio(2000 + i, 160000) for i in range(4)])
    feats, _ = eng.features(pcm, np.arange(5, dtype=np.int64) * 160000, dtype=np.float32)
    S, K = (5000, 16) if sys.argv[2] == "c2" else (10000, 32)
    model = synth.synth_diag_model(2999, feats.astype(np.float64), S, K)
    eng.model_load_diag(model["mix_offsets"], model["mix_gauss"], model["mix_weight"], model["means"], model["covs"])
    x = feats[:int(sys.argv[3])]
    out = eng.gmm_score(x, precision=F32)
    rec = eng.gmm_lna(x, precision=F32, lnabytes=2)
    print(hashlib.sha256(out.tobytes()).hexdigest(), hashlib.sha256(rec.tobytes()).hexdigest(), float(out[5, 7]), flush=True)
    eng.close()
    sys.exit(0)

ok = True
for model, frames in (("c2", 1024), ("c2", 4096 + 256), ("c4", 512)):
    res = {}
    for cl in ("1", "2", "4"):
        env = dict(os.environ, AKUGPU_TC16_CLUSTER=cl)
        try:
            r = subprocess.run([sys.executable, __file__, "run", model, str(frames)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=150)
            res[cl] = r.stdout.decode().strip() if r.returncode == 0 else "FAILED rc=%d %s" % (r.returncode, r.stderr.decode()[-300:])
        except subprocess.TimeoutExpired:
            res[cl] = "TIMEOUT"
        print(model, frames, "cluster", cl, res[cl], flush=True)
    ok = ok and res["1"] == res["2"] == res["4"] and "FAILED" not in res["1"] and "TIMEOUT" not in res["1"]
print("bit-identical across cluster sizes:", ok)
sys.exit(0 if ok else 1)
