"""Error of the throughput scorers (default tc16, bf16x3, FP32 pipe) against the double path on random models of other shapes."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aaltoasr_b200 import AkuGpu, F32, F64
eng = AkuGpu(0)
for D, max_mix in [(39, 64), (47, 9), (63, 20), (70, 12)]:
    rng = np.random.default_rng(100 * D + max_mix)
    S = 37
    sizes = rng.integers(1, max_mix + 1, S); sizes[0] = max_mix
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    G = int(off[-1]); F = 300
    feats = rng.standard_normal((F, D)) * rng.uniform(0.5, 3.0, D) + rng.uniform(-2, 2, D)
    means = feats[rng.integers(0, F, G)] + 0.5 * rng.standard_normal((G, D))
    covs = rng.uniform(0.3, 2.0, (G, D))
    w = rng.uniform(0.1, 1.0, G)
    f32 = feats.astype(np.float32)
    for variant in (0, 4, 2):
        eng.set_scorer_variant(variant)
        eng.model_load_diag(off, np.arange(G, dtype=np.int32), w, means, covs)
        want = np.log(eng.gmm_score(f32.astype(np.float64), precision=F64))
        got = eng.gmm_score(f32, precision=F32).astype(np.float64)
        live = want > -100
        err = np.abs(got - want)[live]
        rel = (np.abs(got - want) / (1 + np.abs(want) / 40))[live]
        print("D=%d mix<=%d variant %d: max abs err %.2e (scaled %.2e), mean %.2e, live %d, ll range %.0f..%.0f" % (
            D, max_mix, variant, err.max(), rel.max(), err.mean(), live.sum(), want[live].min(), want[live].max()))
