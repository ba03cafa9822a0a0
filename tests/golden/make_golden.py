"""Generates the committed golden fixtures from the REFERENCE itself.

Run in the dev container (needs /root/reference and oracle/_ref built by oracle/build_ref.sh):
    python tests/golden/make_golden.py

  aku_tests.npz   the reference's own test fixtures for the feature path, re-encoded:
                  aku/tests/short.wav samples, the three .feaconf texts and the three golden
                  matrices mfcc_p_dd.ref / mfcc_cms_norm.ref / pre_test.ref (2-decimal ASCII).
  ref_small.npz   outputs of the reference's code (oracle/_ref) on a seeded synthetic case:
                  1.5 s of audio, 39-dim MFCC config, 24-state x 4-mix diagonal model with
                  unequal mixture sizes and shared Gaussians: float64 features (incl. frames
                  outside the file), intermediate module outputs, linear state likelihoods,
                  and the LNA files written by the literal aku/phone_probs.cc (2 and 4 bytes,
                  with and without normalisation).
  ref_full.npz    the same for a mixed diagonal / full-covariance pool (FullCovarianceGaussian, exponential form)
  ref_clust.npz   the Gaussian-clustering approximation (phone_probs -C x.gcl --eval-minc/--eval-ming) on the ref_small
                  model: .gcl text (12 clusters, three Gaussians left unlisted), state likelihoods of the reference's
                  HmmSet for three (min_clusters, min_gaussians) settings, and the LNA files of the literal phone_probs.
  ref_spk.npz     speaker adaptation through feature-module parameters (phone_probs -S x.spkc): a 39x39 lin_transform
                  ("cmllr") after the MFCC chain, two speakers with their own matrix + bias and a default speaker, a
                  three-line recipe (speaker A, speaker B, unknown speaker -> default); LNA files of the literal phone_probs.
  ref_pre.npz     the `pre` base module (stored float32 features, int32 dim header: feacat -H --raw-output) followed by a
                  delta module: the reference's output for frames -4 .. n+4 (first / last row replicated outside the file).
  ref_vtln.npz    the vtln module (VtlnModule: bilinear / piecewise-linear / SLAPT warps, Lanczos-sinc or linear
                  interpolation) between fft and mel, warp factors set per speaker through a speaker file: the
                  reference's 39-dim features for four configurations x three speakers.
  ref_modx.npz    the two remaining module types: sr_norm (Lanczos resampling over the frames a concat module stacked,
                  speech rate per speaker) and quanteq (per-channel power-law equalisation, parameters per speaker).
  ref_edge.npz    the same for a handmade edge-case model (underflow / denormal / floor regimes,
                  zero-variance dimensions, tiny weights).
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from aaltoasr_b200 import formats, synth   # noqa: E402
from oracle import oracle_np, ref          # noqa: E402

REF = os.environ.get("AKU_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def parse_ref(path):
    return np.array([[float(v) for v in line.split()] for line in open(path) if line.strip()])


def aku_tests():
    T = os.path.join(REF, "aku", "tests")
    pcm, sr = formats.read_wav(os.path.join(T, "short.wav"))
    np.savez_compressed(
        os.path.join(HERE, "aku_tests.npz"), short_wav=pcm, sample_rate=sr,
        mfcc_p_dd_cfg=open(os.path.join(T, "mfcc_p_dd.feaconf")).read(),
        mfcc_cms_norm_cfg=open(os.path.join(T, "mfcc_cms_norm.feaconf")).read(),
        pre_cfg=open(os.path.join(T, "pre.feaconf")).read(),
        mfcc_p_dd_ref=parse_ref(os.path.join(T, "mfcc_p_dd.ref")),
        mfcc_cms_norm_ref=parse_ref(os.path.join(T, "mfcc_cms_norm.ref")),
        pre_test_ref=parse_ref(os.path.join(T, "pre_test.ref")))


def small_model(feats, seed):
    rng = np.random.default_rng(seed)
    m = synth.synth_diag_model(seed, feats, 24, 4)
    # unequal mixture sizes + shared Gaussians + un-normalised weights
    off, mg, mw = [0], [], []
    for s in range(24):
        k = int(rng.integers(1, 7))
        idx = rng.integers(0, m["means"].shape[0], size=k)
        mg += list(idx)
        mw += list(rng.uniform(0.1, 3.0, size=k))
        off.append(len(mg))
    m["mix_offsets"], m["mix_gauss"], m["mix_weight"] = np.array(off, np.int32), np.array(mg, np.int32), np.array(mw)
    return m


def edge_model(feats, seed):
    """States engineered to hit: double underflow (< -745), float zero (-115..-104), float denormal
    (-103.3..-87.3), lp straddling -36.008, near-certain posteriors, 1e-24 weights, cov <= 0 dims."""
    rng = np.random.default_rng(seed)
    D = feats.shape[1]
    sd = feats.std(axis=0)
    mu0 = feats.mean(axis=0)
    shifts = [0.0, 0.5, 1.0, 1.5, 2.0, 2.4, 2.7, 2.9, 3.0, 3.1, 3.2, 3.3, 3.5, 4.0, 6.0, 9.0]
    means, covs, off, mg, mw = [], [], [0], [], []
    for s, sh in enumerate(shifts):
        k = 1 + (s % 3)
        for j in range(k):
            means.append(mu0 + sh * sd * np.sign(rng.standard_normal(D)) + 0.1 * sd * rng.standard_normal(D))
            cv = rng.uniform(0.5, 2.0, D) * sd ** 2
            if s % 5 == 4:
                cv[rng.integers(0, D)] = 0.0        # zero variance: dimension disabled, constant stays 0
            if s % 7 == 6:
                cv[rng.integers(0, D)] = -1.0
            covs.append(cv)
            mg.append(len(means) - 1)
            mw.append(1e-24 if (j == 1 and s % 2) else rng.uniform(0.2, 1.0))
        off.append(len(mg))
    return dict(mix_offsets=np.array(off, np.int32), mix_gauss=np.array(mg, np.int32), mix_weight=np.array(mw),
                means=np.array(means), covs=np.array(covs))


def full_model(feats, seed):
    """Mixed pool for BASELINE config 5 semantics: every other Gaussian has a full covariance
    diag(U(0.5,2) var) + 0.1 A A^T (A: D x 4), SPD by construction.  (A non-SPD covariance is undefined
    behaviour in the reference: set_covariance leaves the exponential parameters unsized and the next
    precompute_likelihoods reads past them, so it is not part of the fixture.)"""
    rng = np.random.default_rng(seed)
    D, S, M = feats.shape[1], 8, 3
    G = S * M
    sd = feats.std(axis=0)
    means = feats[rng.integers(0, feats.shape[0], G)] + 0.3 * sd * rng.standard_normal((G, D))
    A = rng.standard_normal((G, D, 4)) * sd[None, :, None]
    full = np.array([np.diag(rng.uniform(0.5, 2, D) * sd ** 2) + 0.1 * A[g] @ A[g].T for g in range(G)])
    mask = np.zeros(G, dtype=bool)
    mask[::2] = True
    covs = rng.uniform(0.5, 2, (G, D)) * sd ** 2
    return dict(mix_offsets=np.arange(0, G + 1, M, dtype=np.int32), mix_gauss=np.arange(G, dtype=np.int32),
                mix_weight=rng.dirichlet(np.ones(M), S).reshape(-1), means=means, covs=covs, full_covs=full,
                full_mask=mask)


def run_case(name, pcm, model, tmp):
    wav = os.path.join(tmp, name + ".wav")
    cfg = os.path.join(tmp, name + ".cfg")
    base = os.path.join(tmp, name)
    formats.write_wav(wav, pcm, 16000)
    open(cfg, "w").write(synth.mfcc39_config())
    formats.write_model(base, **model)
    feats, last, _ = ref.features(cfg, wav)
    feats_ext, _, _ = ref.features(cfg, wav, -12, feats.shape[0] + 12)
    mods = {m: ref.module_output(cfg, wav, m, -3, 12) for m in ("fft", "mel", "power", "mfcc", "delta1", "delta2")}
    M = ref.Model(base)
    lik = M.state_likelihoods(feats)
    gll = M.gaussian_loglik(feats[:64])
    M.close()
    rec = os.path.join(tmp, name + ".recipe")
    open(rec, "w").write("audio=%s lna=%s.lna\n" % (wav, name))
    lna = {}
    for nb in (2, 4):
        for nn in (0, 1):
            od = os.path.join(tmp, "o%d%d" % (nb, nn))
            os.makedirs(od, exist_ok=True)
            ref.phone_probs(cfg, base, rec, od, nb, extra=["-N"] if nn else [])
            lna["lna%d%s" % (nb, "_nonorm" if nn else "")] = np.frombuffer(
                open(os.path.join(od, name + ".lna"), "rb").read(), dtype=np.uint8)
    out = dict(pcm=pcm, cfg=synth.mfcc39_config(), feats=feats, feats_ext=feats_ext, ext_start=-12, last_frame=last,
               lik=lik, gauss_loglik64=gll, **{"mod_" + k: v for k, v in mods.items()}, **lna,
               **{"model_" + k: v for k, v in model.items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "frames", feats.shape, "states", lik.shape[1], "loglik range", np.log(lik).min(), np.log(lik).max())


def clust_case(pcm, model, tmp):
    """Clustering fixture on the ref_small inputs."""
    name = "ref_clust"
    rng = np.random.default_rng(7004)
    wav = os.path.join(tmp, name + ".wav"); cfg = os.path.join(tmp, name + ".cfg"); base = os.path.join(tmp, name)
    formats.write_wav(wav, pcm, 16000)
    open(cfg, "w").write(synth.mfcc39_config())
    formats.write_model(base, **model)
    feats, _, _ = ref.features(cfg, wav)
    G, C = model["means"].shape[0], 12
    sd = feats.std(axis=0)
    seeds = model["means"][rng.choice(G, C, replace=False)]
    assign = np.argmin((((model["means"][:, None, :] - seeds[None]) / sd) ** 2).sum(axis=2), axis=1)
    listed = np.ones(G, dtype=bool)
    listed[rng.choice(G, 3, replace=False)] = False
    order = rng.permutation(np.nonzero(listed)[0])
    gcl = "%d\n" % C + "".join("%d %d\n" % (g, assign[g]) for g in order)
    gpath = os.path.join(tmp, name + ".gcl")
    open(gpath, "w").write(gcl)
    gpath2 = os.path.join(tmp, name + "_nonl.gcl")
    open(gpath2, "w").write(gcl.rstrip("\n"))               # no trailing newline: same pairs, same EOF behaviour
    settings = [(0.0, 0.25), (0.3, 0.1), (0.0, 0.1)]
    out = dict(gcl=gcl, settings=np.array(settings), feats=feats, **{"model_" + k: v for k, v in model.items()})
    for k, (mc, mg) in enumerate(settings):
        M = ref.Model(base)
        M.read_clustering(gpath if k != 1 else gpath2)
        M.set_clustering_min_evals(mc, mg)
        out["lik%d" % k] = M.state_likelihoods(feats)
        M.close()
    rec = os.path.join(tmp, name + ".recipe")
    open(rec, "w").write("audio=%s lna=%s.lna\n" % (wav, name))
    for nb in (2, 4):
        od = os.path.join(tmp, "c%d" % nb)
        os.makedirs(od, exist_ok=True)
        ref.phone_probs(cfg, base, rec, od, nb, extra=["-C", gpath, "--eval-ming=0.25"])
        out["lna%d" % nb] = np.frombuffer(open(os.path.join(od, name + ".lna"), "rb").read(), dtype=np.uint8)
    M = ref.Model(base)
    exact = M.state_likelihoods(feats)
    M.close()
    out["lik_exact"] = exact
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "frames", feats.shape[0], "G", G, "C", C,
          "states differing from exact evaluation: %.1f%% / %.1f%% / %.1f%%" % tuple(
              100 * (out["lik%d" % k] != exact).mean() for k in range(3)))


SPK_TAIL = """
module
{
  name cmllr
  type lin_transform
  sources final
}
"""


def spk_case(pcm, model, tmp):
    name = "ref_spk"
    rng = np.random.default_rng(7005)
    cfg_text = synth.mfcc39_config() + SPK_TAIL
    cfg = os.path.join(tmp, name + ".cfg"); base = os.path.join(tmp, name)
    open(cfg, "w").write(cfg_text)
    formats.write_model(base, **model)
    D = 39
    def block(spk):
        A = np.eye(D) + 0.05 * rng.standard_normal((D, D))
        b = 0.3 * rng.standard_normal(D)
        return ("speaker %s\n{\n  cmllr\n  {\n    matrix %s\n    bias %s\n  }\n}\n\n" %
                (spk, " ".join("%.6g" % v for v in A.reshape(-1)), " ".join("%.6g" % v for v in b)))
    spkc = "speaker default\n{\n  feature cmllr\n  {\n  }\n}\n\n" + block("alice") + block("bob")
    spath = os.path.join(tmp, name + ".spkc")
    open(spath, "w").write(spkc)
    cuts = [pcm, pcm[:16000], pcm[5000:22000]]
    lines = []
    for i, (c, spk) in enumerate(zip(cuts, ["alice", "bob", "carol"])):
        w = os.path.join(tmp, "spk%d.wav" % i)
        formats.write_wav(w, c, 16000)
        lines.append("audio=%s lna=spk%d.lna speaker=%s" % (w, i, spk))
    rec = os.path.join(tmp, name + ".recipe")
    open(rec, "w").write("\n".join(lines) + "\n")
    od = os.path.join(tmp, "spk_out")
    os.makedirs(od, exist_ok=True)
    ref.phone_probs(cfg, base, rec, od, 2, extra=["-S", spath])
    out = dict(cfg=cfg_text, spkc=spkc, pcm=pcm, cut_ranges=np.array([[0, pcm.size], [0, 16000], [5000, 22000]]),
               speakers=np.array(["alice", "bob", "carol"]), **{"model_" + k: v for k, v in model.items()})
    for i in range(3):
        out["lna2_%d" % i] = np.frombuffer(open(os.path.join(od, "spk%d.lna" % i), "rb").read(), dtype=np.uint8)
    # the same utterances without -S (identity transform) must differ for alice / bob and agree for carol
    od2 = os.path.join(tmp, "spk_out0")
    os.makedirs(od2, exist_ok=True)
    ref.phone_probs(cfg, base, rec, od2, 2)
    for i in range(3):
        out["lna2_plain_%d" % i] = np.frombuffer(open(os.path.join(od2, "spk%d.lna" % i), "rb").read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "differs from identity:", [float((out["lna2_%d" % i] != out["lna2_plain_%d" % i]).mean()) for i in range(3)])


PRE_CFG = """module
{
  name pre
  type pre
  dim 39
}

module
{
  name d
  type delta
  sources pre
  width 2
}
"""


def pre_case(feats, tmp):
    rows = feats[:60].astype(np.float32)
    raw = os.path.join(tmp, "ref_pre.raw")
    with open(raw, "wb") as f:
        f.write(np.int32(rows.shape[1]).tobytes())
        f.write(rows.tobytes())
    cfg = os.path.join(tmp, "ref_pre.cfg")
    open(cfg, "w").write(PRE_CFG)
    out, last, _ = ref.features(cfg, raw)                      # until the reference reports eof
    ext, _, _ = ref.features(cfg, raw, -4, rows.shape[0] + 4)
    base = ref.module_output(cfg, raw, "pre", -4, rows.shape[0] + 4)
    np.savez_compressed(os.path.join(HERE, "ref_pre.npz"), cfg=PRE_CFG, rows=rows, out=out, ext=ext, ext_start=-4, base_ext=base,
                        last_frame=last)
    print("ref_pre rows", rows.shape, "frames until eof", out.shape[0], "last_frame", last)


def vtln_cfg(extra):
    base = synth.mfcc39_config()
    vt = "module\n{\n  name vtln\n  type vtln\n  sources fft\n%s}\n\n" % "".join("  %s\n" % e for e in extra)
    return base.replace("module\n{\n  name mel\n  type mel\n  sources fft\n}", vt + "module\n{\n  name mel\n  type mel\n  sources vtln\n}")


def vtln_case(pcm, tmp):
    wav = os.path.join(tmp, "vtln.wav")
    formats.write_wav(wav, pcm[:12000], 16000)
    variants = {"blin": [], "pwlin": ["pwlin_vtln 1", "pwlin_turnpoint 0.75"], "linear": ["sinc_interpolation_rad 0"],
                "slapt": ["slapt 1", "sinc_interpolation_rad 4", "lanczos_window 0"]}
    out = dict(pcm=pcm[:12000])
    for name, extra in variants.items():
        cfg_text = vtln_cfg(extra)
        assert "sources vtln" in cfg_text
        cfg = os.path.join(tmp, "vtln_%s.cfg" % name)
        open(cfg, "w").write(cfg_text)
        if name == "slapt":
            params = {"s1": "slapt_coef 0.02 -0.01", "s2": "slapt_coef -0.03"}
        else:
            params = {"s1": "warp_factor 0.9", "s2": "warp_factor 1.12"}
        spkc = "speaker default\n{\n  vtln\n  {\n  }\n}\n\n" + "".join(
            "speaker %s\n{\n  feature vtln\n  {\n    %s\n  }\n}\n\n" % (k, v) for k, v in params.items())
        sp = os.path.join(tmp, "vtln_%s.spkc" % name)
        open(sp, "w").write(spkc)
        out["cfg_" + name] = cfg_text
        out["spkc_" + name] = spkc
        for spk in ("s1", "s2", "other"):
            out["feats_%s_%s" % (name, spk)] = ref.features_spk(cfg, wav, sp, spk)
        plain, _, _ = ref.features(cfg, wav)
        assert np.array_equal(plain, out["feats_%s_other" % name])          # default speaker = warp 1
        print("ref_vtln", name, out["feats_%s_s1" % name].shape, "max |s1 - unwarped| %.3f" % np.abs(out["feats_%s_s1" % name] - plain).max())
    np.savez_compressed(os.path.join(HERE, "ref_vtln.npz"), **out)


def vtln_allpass_case(pcm, tmp):
    """VtlnModule with `all-pass 1` (aku/FeatureModules.cc:1717-1904): bilinear and SLAPT warps as cepstral matrices."""
    wav = os.path.join(tmp, "vtlnap.wav")
    formats.write_wav(wav, pcm[:8000], 16000)
    variants = {"blin": ["all-pass 1"], "slapt": ["all-pass 1", "slapt 1"]}
    out = dict(pcm=pcm[:8000])
    for name, extra in variants.items():
        cfg_text = vtln_cfg(extra)
        cfg = os.path.join(tmp, "vtlnap_%s.cfg" % name)
        open(cfg, "w").write(cfg_text)
        params = {"s1": "slapt_coef 0.02 -0.01", "s2": "slapt_coef -0.03"} if name == "slapt" else {"s1": "warp_factor 0.93", "s2": "warp_factor 1.08"}
        spkc = "speaker default\n{\n  vtln\n  {\n  }\n}\n\n" + "".join(
            "speaker %s\n{\n  feature vtln\n  {\n    %s\n  }\n}\n\n" % (k, v) for k, v in params.items())
        sp = os.path.join(tmp, "vtlnap_%s.spkc" % name)
        open(sp, "w").write(spkc)
        out["cfg_" + name] = cfg_text
        out["spkc_" + name] = spkc
        for spk in ("s1", "s2", "other"):
            out["feats_%s_%s" % (name, spk)] = ref.features_spk(cfg, wav, sp, spk)
        plain, _, _ = ref.features(vtln_cfg_path(tmp), wav)
        print("ref_vtln_allpass", name, out["feats_%s_s1" % name].shape,
              "max |s1 - no vtln| %.3f, |default - no vtln| %.3g" % (np.abs(out["feats_%s_s1" % name] - plain).max(),
                                                                     np.abs(out["feats_%s_other" % name] - plain).max()))
        P = oracle_np.Pipeline(cfg_text)
        for spk, prm in list(params.items()) + [("other", None)]:
            if prm:
                P.set_parameters("vtln", prm)
            else:
                P.set_parameters("vtln", "")
            mine = P.run(pcm[:8000])
            print("   oracle vs reference, speaker %s: max abs diff %.3g" % (spk, np.abs(mine - out["feats_%s_%s" % (name, spk)]).max()))
    np.savez_compressed(os.path.join(HERE, "ref_vtln_allpass.npz"), **out)


def vtln_cfg_path(tmp):
    """The same chain with a pass-through vtln module (warp 1, linear interpolation): the unwarped features."""
    p = os.path.join(tmp, "vtln_plain.cfg")
    open(p, "w").write(vtln_cfg(["sinc_interpolation_rad 0"]))
    return p


MODX_HEAD = """module
{
  name audiofile
  type audiofile
  sample_rate 16000
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

"""
SRNORM_CFG = MODX_HEAD + """module
{
  name stack
  type concat
  sources mel
  left 3
  right 3
}

module
{
  name srn
  type sr_norm
  sources stack
  in_frames 7
  out_frames 5
  lanczos_order 2
}
"""
QUANTEQ_CFG = MODX_HEAD + """module
{
  name qe
  type quanteq
  sources mel
}

module
{
  name mfcc
  type dct
  sources qe
}
"""


def modx_case(pcm, tmp):
    rng = np.random.default_rng(7006)
    wav = os.path.join(tmp, "modx.wav")
    formats.write_wav(wav, pcm[:12000], 16000)
    out = dict(pcm=pcm[:12000], cfg_srnorm=SRNORM_CFG, cfg_quanteq=QUANTEQ_CFG)
    vec = lambda a: " ".join("%.6g" % v for v in a)
    spk = {"srnorm": {"s1": "speech_rate 0.8", "s2": "speech_rate 1.3"},
           "quanteq": {"s1": "alpha %s\n    gamma %s\n    quant_max %s" % (vec(rng.uniform(0.3, 0.9, 21)), vec(rng.uniform(0.5, 1.5, 21)), vec(rng.uniform(4, 9, 21))),
                       "s2": "alpha %s\n    gamma %s\n    quant_max %s" % (vec(rng.uniform(0.3, 0.9, 21)), vec(rng.uniform(0.5, 1.5, 21)), vec(rng.uniform(4, 9, 21)))}}
    for name, cfg_text, mod in (("srnorm", SRNORM_CFG, "srn"), ("quanteq", QUANTEQ_CFG, "qe")):
        cfg = os.path.join(tmp, "modx_%s.cfg" % name)
        open(cfg, "w").write(cfg_text)
        spkc = "speaker default\n{\n  %s\n  {\n  }\n}\n\n" % mod + "".join(
            "speaker %s\n{\n  %s\n  {\n    %s\n  }\n}\n\n" % (k, mod, v) for k, v in spk[name].items())
        sp = os.path.join(tmp, "modx_%s.spkc" % name)
        open(sp, "w").write(spkc)
        out["spkc_" + name] = spkc
        for who in ("s1", "s2", "other"):
            out["feats_%s_%s" % (name, who)] = ref.features_spk(cfg, wav, sp, who)
        print("ref_modx", name, out["feats_%s_s1" % name].shape,
              "max |s1 - default| %.3f" % np.abs(out["feats_%s_s1" % name] - out["feats_%s_other" % name]).max())
    np.savez_compressed(os.path.join(HERE, "ref_modx.npz"), **out)


def cmllr_case(pcm, model, tmp):
    """Model-level CMLLR, global transform (`model cmllr`, unitmode UNIT_NO): state likelihoods of aku::HmmSet after
    aku::SpeakerConfig::set_speaker (plain and with Gaussian clustering), and the LNA files of the literal phone_probs
    -S for three speakers (the third falls back to the default speaker = no transform)."""
    name = "ref_cmllr"
    rng = np.random.default_rng(7010)
    cfg_text = synth.mfcc39_config()
    wav = os.path.join(tmp, name + ".wav"); cfg = os.path.join(tmp, name + ".cfg"); base = os.path.join(tmp, name)
    formats.write_wav(wav, pcm, 16000)
    open(cfg, "w").write(cfg_text)
    formats.write_model(base, **model)
    feats, _, _ = ref.features(cfg, wav)
    D = 39
    Ws = {}
    sd = feats.std(axis=0)
    def block(spk, flip):
        A = np.eye(D) + 0.03 * rng.standard_normal((D, D)) * (sd[:, None] / sd[None, :])
        if flip:
            A[3, 3] = -A[3, 3]              # the likelihood factor is |prod diag(A)|
        b = 0.15 * sd * rng.standard_normal(D)
        W = np.concatenate([b[:, None], A], axis=1)
        text = " ".join("%g" % v for v in W.reshape(-1))      # what the reference's own writer emits (ModelModules.cc:143)
        Ws[spk] = np.array([float(t) for t in text.split()]).reshape(D, D + 1)
        return "speaker %s\n{\n  model cmllr\n  {\n    unitmode UNIT_NO\n    w1 %s\n  }\n}\n\n" % (spk, text)
    spkc = "speaker default\n{\n  model cmllr\n  {\n    unitmode UNIT_NO\n  }\n}\n\n" + block("alice", False) + block("bob", True)
    spath = os.path.join(tmp, name + ".spkc")
    open(spath, "w").write(spkc)
    out = dict(cfg=cfg_text, spkc=spkc, pcm=pcm, feats=feats, W_alice=Ws["alice"], W_bob=Ws["bob"],
               **{"model_" + k: v for k, v in model.items()})
    M = ref.Model(base)
    plain = M.state_likelihoods(feats)
    for spk in ("alice", "bob", "carol", "alice"):        # back to alice: the transform is reloaded
        M.set_speaker(spath, spk)
        lik = M.state_likelihoods(feats)
        if "lik_" + spk in out:
            assert np.array_equal(out["lik_" + spk], lik)
        out["lik_" + spk] = lik
    M.close()
    assert np.array_equal(out["lik_carol"], plain)
    # with the Gaussian-clustering approximation on (cluster centres are wrapped too)
    z = np.load(os.path.join(HERE, "ref_clust.npz"))
    gpath = os.path.join(tmp, name + ".gcl")
    open(gpath, "w").write(str(z["gcl"]))
    out["gcl"] = str(z["gcl"])
    M = ref.Model(base)
    M.read_clustering(gpath)
    M.set_clustering_min_evals(0.0, 0.25)
    M.set_speaker(spath, "bob")
    out["lik_clust_bob"] = M.state_likelihoods(feats)
    M.close()
    # the literal tool
    cuts = [pcm, pcm[:16000], pcm[5000:22000]]
    lines = []
    for i, (c, spk) in enumerate(zip(cuts, ["alice", "bob", "carol"])):
        w = os.path.join(tmp, "cm%d.wav" % i)
        formats.write_wav(w, c, 16000)
        lines.append("audio=%s lna=cm%d.lna speaker=%s" % (w, i, spk))
    rec = os.path.join(tmp, name + ".recipe")
    open(rec, "w").write("\n".join(lines) + "\n")
    for tag, extra in (("", []), ("raw", ["-N"])):
        for nb in (2, 4):
            od = os.path.join(tmp, "cm_out%s%d" % (tag, nb))
            os.makedirs(od, exist_ok=True)
            ref.phone_probs(cfg, base, rec, od, nb, extra=["-S", spath] + extra)
            for i in range(3):
                out["lna%d%s_%d" % (nb, tag, i)] = np.frombuffer(open(os.path.join(od, "cm%d.lna" % i), "rb").read(), dtype=np.uint8)
    out["cut_ranges"] = np.array([[0, pcm.size], [0, 16000], [5000, 22000]])
    out["speakers"] = np.array(["alice", "bob", "carol"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "log-lik shift alice/bob vs plain (median):",
          float(np.median(np.log(out["lik_alice"]) - np.log(plain))), float(np.median(np.log(out["lik_bob"]) - np.log(plain))),
          "clustered differs from exact on %.1f%%" % (100 * (out["lik_clust_bob"] != out["lik_bob"]).mean()))


def cmllr_units_case(pcm, model, tmp):
    """Model-level CMLLR with regression classes (`model cmllr`, unitmode UNIT_PHONE / UNIT_MIX / UNIT_GAUSSIAN,
    aku/ModelModules.cc:62-95,172-236): state likelihoods of aku::HmmSet after SpeakerConfig::set_speaker for one speaker
    per unit mode (two transforms each, claiming overlapping Gaussians through the model's shared components), and the
    LNA file of the literal phone_probs -S for each."""
    name = "ref_cmllr_units"
    rng = np.random.default_rng(7020)
    cfg_text = synth.mfcc39_config()
    wav = os.path.join(tmp, name + ".wav"); cfg = os.path.join(tmp, name + ".cfg"); base = os.path.join(tmp, name)
    formats.write_wav(wav, pcm, 16000)
    open(cfg, "w").write(cfg_text)
    formats.write_model(base, **model)
    # eight three-state phones over the 24 states; context-dependent labels exercise Hmm::get_center_phone
    labels = ["a", "x-b+y", "c+z", "q-d", "e", "a-f+a", "g", "x-a+y"]
    phones = [(labels[p], [3 * p, 3 * p + 1, 3 * p + 2]) for p in range(8)]
    ph_text = "PHONE\n8\n" + "".join("%d 5 %s\n-1 -2 %d %d %d\n0 1 2 1\n1 0\n2 2 2 0.8 3 0.2\n3 2 3 0.8 4 0.2\n4 2 4 0.8 1 0.2\n"
                                      % (p + 1, labels[p], 3 * p, 3 * p + 1, 3 * p + 2) for p in range(8))
    open(base + ".ph", "w").write(ph_text)
    feats, _, _ = ref.features(cfg, wav)
    D = 39
    sd = feats.std(axis=0)
    def matrix(flip=False):
        A = np.eye(D) + 0.03 * rng.standard_normal((D, D)) * (sd[:, None] / sd[None, :])
        if flip:
            A[5, 5] = -A[5, 5]
        b = 0.15 * sd * rng.standard_normal(D)
        W = np.concatenate([b[:, None], A], axis=1)
        text = " ".join("%g" % v for v in W.reshape(-1))
        # the reference parses the values with str::str2float (aku/str.cc:261-282): they pass through a float
        return text, np.array([np.float32(float(t)) for t in text.split()], dtype=np.float64).reshape(D, D + 1)
    cases = {"phone": ("UNIT_PHONE", [["a", "d"], ["b", "g", "nosuch"]]),          # "a" = phones 0 and 7 (x-a+y)
             "mix": ("UNIT_MIX", [["7", "2", "19"], ["0", "1", "2", "3", "23"]]),      # mixture 2 in both: map order decides
             "gauss": ("UNIT_GAUSSIAN", [["5", "40", "41", "90"], ["0", "1", "2", "40", "17"]])}
    spkc = "speaker default\n{\n}\n\n"
    out = dict(cfg=cfg_text, pcm=pcm, feats=feats, ph=ph_text, **{"model_" + k: v for k, v in model.items()})
    for spk, (um, unit_lists) in cases.items():
        lines = ["    unitmode " + um]
        for i, units in enumerate(unit_lists):
            text, W = matrix(flip=(i == 1))
            lines.append("    w%d %s %s" % (i + 1, " ".join(units), text))
            out["W_%s_%d" % (spk, i)] = W
            out["units_%s_%d" % (spk, i)] = np.array(units)
        spkc += "speaker %s\n{\n  model cmllr\n  {\n%s\n  }\n}\n\n" % (spk, "\n".join(lines))
        out["unitmode_" + spk] = um
    spath = os.path.join(tmp, name + ".spkc")
    open(spath, "w").write(spkc)
    out["spkc"] = spkc
    M = ref.Model(base)
    plain = M.state_likelihoods(feats)
    out["lik_plain"] = plain
    for spk, (um, unit_lists) in cases.items():
        M.set_speaker(spath, spk)
        lik = M.state_likelihoods(feats)
        out["lik_" + spk] = lik
        trs = [(unit_lists[i], out["W_%s_%d" % (spk, i)]) for i in range(len(unit_lists))]
        g2t, ordered = oracle_np.cmllr_unit_assignment(um, trs, model, phones)
        mine = oracle_np.state_likelihoods(model, feats, cmllr_units=(g2t, [w for _, w in ordered]))
        rel = np.abs(mine - lik) / lik
        print(name, spk, um, "adapted Gaussians per transform:", [int((g2t == t).sum()) for t in range(len(trs))],
              "states changed: %d / 24" % int((np.abs(lik - plain).max(axis=0) > 0).sum()), "oracle max rel diff %.3g" % rel.max(),
              "bit-identical" if np.array_equal(mine, lik) else "")
        out["g2t_" + spk] = g2t
    # a speaker without a `model cmllr` entry leaves the previous speaker's transforms loaded (set_modules only touches
    # listed modules, aku/SpeakerConfig.cc:365-380)
    M.set_speaker(spath, "nobody")
    assert np.array_equal(M.state_likelihoods(feats), out["lik_gauss"])
    M.close()
    lines = []
    for i, spk in enumerate(cases):
        lines.append("audio=%s lna=cu%d.lna speaker=%s" % (wav, i, spk))
    rec = os.path.join(tmp, name + ".recipe")
    open(rec, "w").write("\n".join(lines) + "\n")
    for nb in (2, 4):
        od = os.path.join(tmp, "cu_out%d" % nb)
        os.makedirs(od, exist_ok=True)
        ref.phone_probs(cfg, base, rec, od, nb, extra=["-S", spath])
        for i, spk in enumerate(cases):
            out["lna%d_%s" % (nb, spk)] = np.frombuffer(open(os.path.join(od, "cu%d.lna" % i), "rb").read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def main():
    if not ref.available():
        raise SystemExit("oracle/_ref is not built: run oracle/build_ref.sh first")
    if sys.argv[1:] == ["vtln_allpass"]:
        with tempfile.TemporaryDirectory() as tmp:
            vtln_allpass_case(synth.synth_audio(7001, 24000), tmp)
        return
    if sys.argv[1:] == ["cmllr_units"]:
        with tempfile.TemporaryDirectory() as tmp:
            pcm = synth.synth_audio(7001, 24000)
            wav = os.path.join(tmp, "probe.wav"); cfg = os.path.join(tmp, "probe.cfg")
            formats.write_wav(wav, pcm, 16000)
            open(cfg, "w").write(synth.mfcc39_config())
            feats, _, _ = ref.features(cfg, wav)
            cmllr_units_case(pcm, small_model(feats, 7002), tmp)
        return
    if sys.argv[1:] == ["cmllr"]:             # only the newest fixture (the others are unchanged)
        with tempfile.TemporaryDirectory() as tmp:
            pcm = synth.synth_audio(7001, 24000)
            wav = os.path.join(tmp, "probe.wav"); cfg = os.path.join(tmp, "probe.cfg")
            formats.write_wav(wav, pcm, 16000)
            open(cfg, "w").write(synth.mfcc39_config())
            feats, _, _ = ref.features(cfg, wav)
            cmllr_case(pcm, small_model(feats, 7002), tmp)
        return
    aku_tests()
    with tempfile.TemporaryDirectory() as tmp:
        pcm = synth.synth_audio(7001, 24000)
        wav = os.path.join(tmp, "probe.wav"); cfg = os.path.join(tmp, "probe.cfg")
        formats.write_wav(wav, pcm, 16000)
        open(cfg, "w").write(synth.mfcc39_config())
        feats, _, _ = ref.features(cfg, wav)
        run_case("ref_small", pcm, small_model(feats, 7002), tmp)
        clust_case(pcm, small_model(feats, 7002), tmp)
        spk_case(pcm, small_model(feats, 7002), tmp)
        pre_case(feats, tmp)
        vtln_case(pcm, tmp)
        vtln_allpass_case(pcm, tmp)
        modx_case(pcm, tmp)
        cmllr_case(pcm, small_model(feats, 7002), tmp)
        cmllr_units_case(pcm, small_model(feats, 7002), tmp)
        run_case("ref_edge", pcm, edge_model(feats, 7003), tmp)
        run_case("ref_full", pcm, full_model(feats, 5999), tmp)


if __name__ == "__main__":
    main()
