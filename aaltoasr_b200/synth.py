"""Synthetic audio and acoustic models with fixed seeds.

Pure data generation on the host (numpy): inputs for tests and bench.py, no arithmetic
of the accelerated path.
"""
import numpy as np

MFCC_39_CFG = """module
{
  name audiofile
  type audiofile
  sample_rate %(sample_rate)d
}

module
{
  name fft
  type fft
  sources audiofile
}

module
{
  name mel
  type mel
  sources fft
}

module
{
  name power
  type power
  sources fft
}

module
{
  name mfcc
  type dct
  sources mel
}

module
{
  name mfcc_power
  type merge
  sources mfcc power
}

module
{
  name delta1
  type delta
  sources mfcc_power
}

module
{
  name delta2
  type delta
  sources delta1
}

module
{
  name final
  type merge
  sources mfcc_power delta1 delta2
}
"""


def mfcc39_config(sample_rate=16000):
    """The 39-dim MFCC+power+delta+delta-delta chain of aku/tests/mfcc_p_dd.feaconf."""
    return MFCC_39_CFG % {"sample_rate": sample_rate}


def synth_audio(seed, n_samples, sample_rate=16000):
    """1-pole low-passed Gaussian noise under a slow AM envelope, sigma ~ 2500, int16."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n_samples)
    y = np.empty(n_samples)
    # y[t] = 0.9 y[t-1] + x[t], vectorised in blocks via the closed-form filter
    from scipy.signal import lfilter
    y = lfilter([1.0], [1.0, -0.9], x)
    t = np.arange(n_samples) / float(sample_rate)
    y *= 0.55 + 0.45 * np.sin(2 * np.pi * 1.3 * t)
    y *= 2500.0 / max(y.std(), 1e-9)
    return np.clip(np.rint(y), -32768, 32767).astype(np.int16)


def synth_diag_model(seed, feats, n_states, n_mix, var_lo=0.5, var_hi=2.0, mean_jitter=0.5):
    """Diagonal GMM whose means are feature frames + noise, so that state log-likelihoods land
    in the range real models produce.  Returns dict(mix_offsets, mix_gauss, mix_weight, means, covs)."""
    rng = np.random.default_rng(seed)
    feats = np.asarray(feats, dtype=np.float64)
    D = feats.shape[1]
    G = n_states * n_mix
    sd = feats.std(axis=0) + 1e-6
    idx = rng.integers(0, feats.shape[0], size=G)
    means = feats[idx] + rng.standard_normal((G, D)) * (mean_jitter * sd)
    covs = rng.uniform(var_lo, var_hi, size=(G, D)) * (sd ** 2)
    w = rng.dirichlet(np.ones(n_mix), size=n_states)
    return dict(mix_offsets=np.arange(0, G + 1, n_mix, dtype=np.int32),
                mix_gauss=np.arange(G, dtype=np.int32),
                mix_weight=w.reshape(-1), means=means, covs=covs)
