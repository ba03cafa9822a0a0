"""Accuracy of the throughput-mode scorers on an all-full-covariance pool (fixture ref_full) against the double path."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aaltoasr_b200 import AkuGpu, F32, F64
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_full.npz"))
m = {k[6:]: z[k] for k in z.files if k.startswith("model_")}
idx = np.nonzero(m["full_mask"])[0]
off = np.arange(0, len(idx) + 1, 3, dtype=np.int32)
eng = AkuGpu(0)
for variant in (0, 4):
    eng.set_scorer_variant(variant)
    eng.model_load_full(off, np.arange(len(idx), dtype=np.int32), np.ones(len(idx)), m["means"][idx], m["full_covs"][idx])
    ll = eng.gmm_score(z["feats"].astype(np.float32), precision=F32).astype(np.float64)
    want = np.log(eng.gmm_score(z["feats"].astype(np.float32).astype(np.float64), precision=F64))
    err = np.abs(ll - want)
    print("full pool variant %d: max abs err %.3e mean %.3e (ll range %.1f..%.1f)" % (variant, err.max(), err.mean(), want.min(), want.max()))
