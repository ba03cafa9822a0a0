"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI,
against the oracle and the committed reference fixtures.

Bars (BASELINE.json north_star):
  * frame counts / frame indexing: identical
  * features: <= 1e-5 absolute (only the FFT's internal rounding order differs)
  * GMM + LNA, parity mode (F64) on the reference's float64 features: LNA bytes identical
  * throughput mode (F32): float log-probs within 1e-4 relative; 2-byte codes within +-1
"""
import os

import numpy as np
import pytest

from aaltoasr_b200 import AkuGpuError, F32, F64
from oracle import oracle_np

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4          # north_star tolerance on float log-probs
FEAT_ABS_TOL = 1e-5


def lna4(rec):
    return np.ascontiguousarray(rec).view("<f4")


def codes2(rec):
    return np.ascontiguousarray(rec).view(">u2").astype(np.int64)


def load_model(engine, model):
    engine.model_load_diag(model["mix_offsets"], model["mix_gauss"], model["mix_weight"], model["means"], model["covs"])


# ------------------------------------------------------------------ features
def test_features_aku_goldens(engine, aku_tests):
    pcm = aku_tests["short_wav"]
    engine.frontend_load_config_text(aku_tests["mfcc_p_dd_cfg"])
    assert engine.num_frames(pcm.size) == 73 and engine.feature_dim == 39
    assert engine.frame_rate == 125.0 and engine.sample_rate == 16000
    out = engine.features_range(pcm, -10, 81)
    gold = aku_tests["mfcc_p_dd_ref"]
    assert np.abs(out - gold[:91]).max() <= 0.0051 and np.abs(out - gold[91:]).max() <= 0.0051
    pre = engine.features_range(pcm, 10, 61).astype(np.float32).astype(np.float64)
    assert np.abs(pre - aku_tests["pre_test_ref"]).max() <= 0.0051
    engine.frontend_load_config_text(aku_tests["mfcc_cms_norm_cfg"])
    out = engine.features_range(pcm, -15, 91)
    assert out.shape == (106, 39)
    assert np.abs(out - aku_tests["mfcc_cms_norm_ref"]).max() <= 0.0051
    # the oracle at full precision, incl. normalization / lin_transform / mean_subtractor / after-EOF frames
    P = oracle_np.Pipeline(aku_tests["mfcc_cms_norm_cfg"])
    assert np.abs(out - P.run(pcm, -15, 91)).max() <= 5e-5


def test_features_random_access(engine, aku_tests):
    """aku/tests/random_feature_test.cc: every access order yields identical frames."""
    pcm = aku_tests["short_wav"]
    engine.frontend_load_config_text(aku_tests["mfcc_p_dd_cfg"])
    seq = engine.features_range(pcm, -10, 81)
    rng = np.random.default_rng(1)
    for f in rng.integers(-10, 81, size=40):
        assert np.array_equal(engine.features_range(pcm, int(f), int(f) + 1)[0], seq[f + 10])


@pytest.mark.parametrize("case", ["ref_small", "ref_edge"])
def test_features_vs_reference(engine, case, request):
    g = request.getfixturevalue(case)
    engine.frontend_load_config_text(g["cfg"])
    f64, fo = engine.features(g["pcm"], dtype=np.float64)
    assert list(fo) == [0, g["feats"].shape[0]]
    assert np.abs(f64 - g["feats"]).max() <= FEAT_ABS_TOL
    f32, _ = engine.features(g["pcm"], dtype=np.float32)
    assert np.array_equal(f32, f64.astype(np.float32))
    s = int(g["ext_start"])
    ext = engine.features_range(g["pcm"], s, s + g["feats_ext"].shape[0])
    assert np.abs(ext - g["feats_ext"]).max() <= FEAT_ABS_TOL
    for mod, tol in (("fft", 2e-2), ("mel", 5e-6), ("power", 5e-6), ("mfcc", 2e-5), ("delta1", 1e-5), ("delta2", 1e-5)):
        got = engine.features_range(g["pcm"], -3, 12, module=mod)
        assert np.abs(got - g["mod_" + mod]).max() <= tol, mod


def test_features_batch_equals_single(engine, ref_small):
    """Ragged batch: utterances of different lengths in one call == one call each (bit-exact)."""
    pcm = ref_small["pcm"]
    engine.frontend_load_config_text(ref_small["cfg"])
    cuts = [pcm[:9000], pcm[3000:24000], pcm[100:400], pcm]
    uo = np.concatenate([[0], np.cumsum([c.size for c in cuts])])
    batch, fo = engine.features(np.concatenate(cuts), uo, dtype=np.float64)
    for k, c in enumerate(cuts):
        single, _ = engine.features(c, dtype=np.float64)
        assert single.shape[0] == engine.num_frames(c.size) == fo[k + 1] - fo[k]
        assert np.array_equal(batch[fo[k]:fo[k + 1]], single)


@pytest.mark.parametrize("sr", [8000, 16000, 48000])
def test_fused_frontend_identical_to_module_by_module(engine, sr, monkeypatch):
    """The fused kernels (fft+mel+power+dct+merge in one, delta+delta+merge+gather in one) perform the same operations
    in the same order as the per-module kernels: features are bit-identical, float and double output."""
    from aaltoasr_b200 import synth
    engine.frontend_load_config_text(synth.mfcc39_config(sr))
    pcm = synth.synth_audio(3100 + sr // 1000, sr * 2, sr)
    cuts = [pcm, pcm[: sr // 2], pcm[sr // 3:]]
    uo = np.concatenate([[0], np.cumsum([c.size for c in cuts])])
    for dt in (np.float64, np.float32):
        l0 = engine.launch_count()
        fused, fo = engine.features(np.concatenate(cuts), uo, dtype=dt)
        n_fused = engine.launch_count() - l0
        monkeypatch.setenv("AKUGPU_FE_NOFUSE", "1")
        l0 = engine.launch_count()
        plain, _ = engine.features(np.concatenate(cuts), uo, dtype=dt)
        n_plain = engine.launch_count() - l0
        monkeypatch.delenv("AKUGPU_FE_NOFUSE")
        assert np.array_equal(fused, plain)
        assert n_fused == 3 and n_plain >= 10, (n_fused, n_plain)      # row table + fused spectrum kernel + fused delta kernel


@pytest.mark.parametrize("sr,ww", [(8000, None), (16000, None), (16000, 512), (16000, 1024), (16000, 2048), (32000, None)])
def test_warp_fft_kernel_against_the_shared_memory_kernel(engine, sr, ww, monkeypatch):
    """Power-of-two windows run one warp per frame with the FFT in registers (fe_spectrum_wfft); the butterflies, twiddles
    and roundings are those of the shared-memory kernel (fe_spectrum_fft, AKUGPU_FE_OLDFFT=1): spectra and features agree
    up to the compiler's choice of which multiply-adds it contracts (a few 1e-6 relative on a spectrum bin)."""
    from aaltoasr_b200 import synth
    cfg = synth.mfcc39_config(sr)
    if ww:
        cfg = cfg.replace("sample_rate %d" % sr, "sample_rate %d\n  window_width %d" % (sr, ww))
    engine.frontend_load_config_text(cfg)
    pcm = synth.synth_audio(3200 + sr // 1000, sr, sr)
    new = engine.features(pcm, dtype=np.float64)[0]
    new_fft = engine.features_range(pcm, -2, 9, module="fft")
    monkeypatch.setenv("AKUGPU_FE_OLDFFT", "1")
    old = engine.features(pcm, dtype=np.float64)[0]
    old_fft = engine.features_range(pcm, -2, 9, module="fft")
    monkeypatch.delenv("AKUGPU_FE_OLDFFT")
    print("sr %d window %s: max |feature diff| %.3g, max relative spectrum diff %.3g, identical: %s" % (
        sr, ww, np.abs(new - old).max(), (np.abs(new_fft - old_fft) / np.maximum(np.abs(old_fft), 1e-30)).max(), np.array_equal(new, old)))
    assert np.abs(new - old).max() <= 3e-6
    assert (np.abs(new_fft - old_fft) <= 5e-6 * np.abs(old_fft) + 1e-3).all()


@pytest.mark.parametrize("sr,ww", [(8000, None), (16000, 512), (32000, None), (16000, 2048), (48000, None), (16000, 400),
                                   (24000, None), (16000, 1536), (12000, None)])
def test_features_sweep_vs_oracle(engine, sr, ww):
    """Config 3: other sample rates / window widths: 2^k and 3 * 2^k (192, 384, 768, 1536) go through the shared-memory
    FFT, 400 through the direct DFT."""
    from aaltoasr_b200 import synth
    cfg = synth.mfcc39_config(sr)
    if ww:
        cfg = cfg.replace("sample_rate %d" % sr, "sample_rate %d\n  window_width %d" % (sr, ww))
    pcm = synth.synth_audio(3000 + sr // 1000, sr // 2, sr)
    engine.frontend_load_config_text(cfg)
    P = oracle_np.Pipeline(cfg)
    got, fo = engine.features(pcm, dtype=np.float64)
    assert got.shape[0] == P.num_frames(pcm.size)
    want = P.run(pcm)
    assert np.abs(got - want).max() <= 5e-5, np.abs(got - want).max()


@pytest.mark.parametrize("variant", ["blin", "pwlin", "linear", "slapt"])
def test_vtln_module(engine, ref_vtln, variant, tmp_path):
    """vtln module between fft and mel (VtlnModule, aku/FeatureModules.cc:1505-1934): bilinear, piecewise-linear and
    SLAPT warps, Lanczos-sinc and linear interpolation, the warp set per speaker through a speaker file
    (SpeakerConfig -> FeatureModule::set_parameters).  Features within the front-end's usual 1e-5 of the reference's."""
    from aaltoasr_b200 import SpeakerConfig
    g = ref_vtln
    engine.frontend_load_config_text(g["cfg_" + variant])
    spkc = str(tmp_path / "v.spkc")
    open(spkc, "w").write(g["spkc_" + variant])
    sc = SpeakerConfig(engine)
    sc.read_speaker_file(spkc)
    got = {}
    for spk in ("s1", "other", "s2"):
        sc.set_speaker(spk)
        feats, _ = engine.features(g["pcm"], dtype=np.float64)
        want = g["feats_%s_%s" % (variant, spk)]
        assert feats.shape == want.shape
        assert np.abs(feats - want).max() <= 2e-5, (variant, spk, np.abs(feats - want).max())
        got[spk] = feats
    assert np.abs(got["s1"] - got["other"]).max() > 0.1          # the warp does something
    with pytest.raises(AkuGpuError, match="lanczos_window and all-pass"):
        engine.frontend_load_config_text(g["cfg_blin"].replace("type vtln", "type vtln\n  all-pass 1\n  lanczos_window 1"))


@pytest.mark.parametrize("variant", ["blin", "slapt"])
def test_vtln_all_pass(engine, variant, tmp_path):
    """vtln with `all-pass 1` (VtlnModule::create_all_pass_blin_transform / _slapt_transform / set_all_pass_transform,
    aku/FeatureModules.cc:1717-1904): the cepstral warp matrix is built on the host in double and applied by fe_vtln as
    full coefficient rows; features within the front-end's usual bar of the reference's, warp per speaker."""
    from aaltoasr_b200 import SpeakerConfig
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_vtln_allpass.npz"))
    engine.frontend_load_config_text(str(z["cfg_" + variant]))
    spkc = str(tmp_path / "v.spkc")
    open(spkc, "w").write(str(z["spkc_" + variant]))
    sc = SpeakerConfig(engine)
    sc.read_speaker_file(spkc)
    got = {}
    for spk in ("s1", "other", "s2", "s1"):
        sc.set_speaker(spk)
        feats, _ = engine.features(z["pcm"], dtype=np.float64)
        want = z["feats_%s_%s" % (variant, spk)]
        assert feats.shape == want.shape
        assert np.abs(feats - want).max() <= 2e-5, (variant, spk, np.abs(feats - want).max())
        got[spk] = feats
    assert np.abs(got["s1"] - got["other"]).max() > 0.1


@pytest.mark.parametrize("which", ["srnorm", "quanteq"])
def test_sr_norm_and_quanteq_modules(engine, ref_modx, which, tmp_path):
    """The last two module types of aku/FeatureModules.cc: sr_norm (:1936-2069, Lanczos resampling over the frames a concat
    module stacked, speech rate per speaker) and quanteq (:2072-2148, per-channel power law, parameters per speaker,
    identity without them): features within the front-end's usual bar of the reference's for three speakers."""
    from aaltoasr_b200 import SpeakerConfig
    g = ref_modx
    engine.frontend_load_config_text(g["cfg_" + which])
    spkc = str(tmp_path / "m.spkc")
    open(spkc, "w").write(g["spkc_" + which])
    sc = SpeakerConfig(engine)
    sc.read_speaker_file(spkc)
    got = {}
    for spk in ("s1", "other", "s2"):
        sc.set_speaker(spk)
        feats, _ = engine.features(g["pcm"], dtype=np.float64)
        want = g["feats_%s_%s" % (which, spk)]
        assert feats.shape == want.shape
        assert np.abs(feats - want).max() <= 3e-5, (which, spk, np.abs(feats - want).max())
        got[spk] = feats
    assert np.abs(got["s1"] - got["other"]).max() > 0.1


def test_pre_base_module(engine, ref_pre, aku_tests):
    """`pre` base module (stored float32 features, aku/FeatureModules.cc:603-755) followed by a delta module: equal
    to the reference's doubles, also outside the file (first / last row replicated); and the third aku/tests golden:
    feacat --raw-output of frames 10..60, read back through pre.feaconf (aku/tests/pre_test.script)."""
    g = ref_pre
    engine.frontend_load_config_text(g["cfg"])
    assert engine.feature_dim == 39 and engine.num_frames(60) == 60
    out, fo = engine.features_pre(g["rows"])
    assert list(fo) == [0, 60] and np.array_equal(out, g["out"])
    n = g["rows"].shape[0]
    ext = engine.features_pre_range(g["rows"], int(g["ext_start"]), n + 4)
    assert np.array_equal(ext, g["ext"])
    base = engine.features_pre_range(g["rows"], -4, n + 4, module="pre")
    assert np.array_equal(base, g["base_ext"])
    two, fo = engine.features_pre(np.concatenate([g["rows"], g["rows"][:17]]), [0, n, n + 17])      # batch of two "files"
    assert list(fo) == [0, n, n + 17] and np.array_equal(two[:n], g["out"])
    with pytest.raises(AkuGpuError, match="base module is `pre`"):
        engine.features(np.zeros(4000, np.int16))
    # aku/tests/pre_test: mfcc_p_dd features of frames 10..60 as raw float32 rows, then the bare `pre` configuration
    engine.frontend_load_config_text(aku_tests["mfcc_p_dd_cfg"])
    rows = engine.features_range(aku_tests["short_wav"], 10, 61, dtype=np.float32)
    engine.frontend_load_config_text(aku_tests["pre_cfg"])
    back, _ = engine.features_pre(rows)
    assert np.array_equal(back, rows.astype(np.float64))
    assert back.shape == aku_tests["pre_test_ref"].shape and np.abs(back - aku_tests["pre_test_ref"]).max() <= 0.0051
    with pytest.raises(AkuGpuError, match="base module is `audiofile`"):
        engine.frontend_load_config_text(aku_tests["mfcc_p_dd_cfg"])
        engine.features_pre(rows)


# ------------------------------------------------------------------ GMM + LNA
@pytest.mark.parametrize("case", ["ref_small", "ref_edge"])
def test_gmm_lna_parity_mode_bit_exact(engine, case, request):
    g = request.getfixturevalue(case)
    load_model(engine, g["model"])
    assert engine.num_states == g["lik"].shape[1] and engine.model_dim == 39
    lik = engine.gmm_score(g["feats"], precision=F64)
    # CUDA's exp() is within 1 ulp of glibc's; everything else is the same sequence of operations
    rel = np.abs(lik - g["lik"]) / g["lik"]
    assert rel.max() <= 4.5e-16, rel.max()
    assert (lik != g["lik"]).mean() < 0.35
    for nb in (2, 4):
        for nonorm in (False, True):
            rec = engine.gmm_lna(g["feats"], precision=F64, lnabytes=nb, normalize=not nonorm)
            want = g["lna%d%s" % (nb, "_nonorm" if nonorm else "")]
            assert engine.lna_header(nb) == bytes(want[:5])
            assert np.array_equal(rec.reshape(-1), want[5:]), (case, nb, nonorm, (rec.reshape(-1) != want[5:]).sum())


@pytest.mark.parametrize("variant", [3, 4])
def test_tensor_core_variant_tolerance(engine, ref_small, ref_full, variant):
    """tcgen05 scorers: variant 3 (= default) is the fp16x2-split kernel for diagonal pools (gmm_tc16.cu), variant 4
    forces the bf16x3-split kernel (gmm_tc.cu), which also serves full-covariance pools; fp32 accumulation in TMEM
    with leading terms and corrections in separate accumulators.  Diagonal pools meet the same bars as the
    FP32-pipe kernel (test_gmm_lna_throughput_mode_tolerance runs them too); the full-covariance contraction
    (K' = 4928) is held to 2e-4 absolute on the log-likelihoods."""
    engine.set_scorer_variant(variant)
    try:
        g = ref_small
        load_model(engine, g["model"])
        feats32 = g["feats"].astype(np.float32)
        ll = engine.gmm_score(feats32, precision=F32).astype(np.float64)
        assert np.abs(ll - np.log(g["lik"])).max() <= 3e-5
        got4 = lna4(engine.gmm_lna(feats32, precision=F32, lnabytes=4))
        want4 = lna4(g["lna4"][5:]).reshape(got4.shape)
        rel = np.abs(got4 - want4) / np.abs(want4)
        assert (rel <= REL_TOL).all(), rel.max()
        d = np.abs(codes2(engine.gmm_lna(feats32, precision=F32, lnabytes=2)) - codes2(g["lna2"][5:]).reshape(got4.shape))
        assert d.max() <= 1 and (d != 0).mean() <= 0.02, (d.max(), (d != 0).mean())
        # all-full pool
        m = ref_full["model"]
        idx = np.nonzero(m["full_mask"])[0]
        off = np.arange(0, len(idx) + 1, 3, dtype=np.int32)
        engine.model_load_full(off, np.arange(len(idx), dtype=np.int32), np.ones(len(idx)), m["means"][idx], m["full_covs"][idx])
        ll = engine.gmm_score(ref_full["feats"].astype(np.float32), precision=F32).astype(np.float64)
        want = np.log(engine.gmm_score(ref_full["feats"], precision=F64))
        assert np.abs(ll - want).max() <= 2e-4
    finally:
        engine.set_scorer_variant(0)


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
@pytest.mark.parametrize("case", ["ref_small", "ref_edge"])
def test_gmm_lna_throughput_mode_tolerance(engine, case, variant, request):
    g = request.getfixturevalue(case)
    engine.set_scorer_variant(variant)
    try:
        load_model(engine, g["model"])
        feats32 = g["feats"].astype(np.float32)
        ll = engine.gmm_score(feats32, precision=F32).astype(np.float64)
        want_ll = np.log(oracle_np.state_likelihoods(g["model"], feats32.astype(np.float64)))
        live = want_ll > -100            # below that the reference itself floors at 1e-50 / flushes
        err = (np.abs(ll - want_ll) / (1 + np.abs(want_ll) / 40))[live]   # fp32 error grows with the magnitude
        # variants 3/4 force the expanded-form kernels also on the edge model (means up to 9 sigma from the centre),
        # which the default sends to the direct-form kernel
        assert err.max() <= (3e-5 if variant >= 3 else 2e-5), err.max()
        for nonorm in (False, True):
            key = "_nonorm" if nonorm else ""
            got4 = lna4(engine.gmm_lna(feats32, precision=F32, lnabytes=4, normalize=not nonorm))
            want4 = lna4(g["lna4" + key][5:]).reshape(got4.shape)
            rel = np.abs(got4 - want4) / np.maximum(np.abs(want4), 1e-30)
            # entries the reference rounds through fp32 denormals are only emulated to fp32 accuracy;
            # un-normalised log-likelihoods cross zero, where only an absolute bound is meaningful
            ok = (rel <= REL_TOL) | (np.abs(got4 - want4) <= 2e-5)
            assert ok.mean() >= 0.999, (case, nonorm, 1 - ok.mean(), rel.max())
            if case == "ref_small":
                assert ok.all(), rel.max()
            got2 = codes2(engine.gmm_lna(feats32, precision=F32, lnabytes=2, normalize=not nonorm))
            want2 = codes2(g["lna2" + key][5:]).reshape(got2.shape)
            d = np.abs(got2 - want2)
            assert (d <= 1).mean() >= 0.999 and (d != 0).mean() <= 0.02, (d.max(), (d != 0).mean())
    finally:
        engine.set_scorer_variant(0)


@pytest.mark.parametrize("case", ["ref_small", "ref_edge"])
def test_lna_row_kernel_identical_to_column_kernel(engine, case, request, monkeypatch):
    """The bandwidth-oriented epilogue (lna_f32_rows, fp32-only arithmetic) writes the same bytes as the original
    column kernel (double expressions) -- both lnabytes, with and without normalisation, incl. the edge model."""
    g = request.getfixturevalue(case)
    load_model(engine, g["model"])
    feats32 = g["feats"].astype(np.float32)
    for nb in (2, 4):
        for normalize in (True, False):
            new = engine.gmm_lna(feats32, precision=F32, lnabytes=nb, normalize=normalize)
            monkeypatch.setenv("AKUGPU_LNA_OLD", "1")
            old = engine.gmm_lna(feats32, precision=F32, lnabytes=nb, normalize=normalize)
            monkeypatch.delenv("AKUGPU_LNA_OLD")
            assert np.array_equal(new, old), (case, nb, normalize, (new != old).mean())


def test_gaussian_clustering_parity(engine, ref_clust, tmp_path):
    """phone_probs -C x.gcl --eval-minc/--eval-ming (section 8f-1): state likelihoods within CUDA-exp ulps of the
    reference's HmmSet for three settings, LNA bytes identical to the literal phone_probs output; the .gcl reader
    (file) and the in-memory loader agree; throughput-mode requests are served by the same double path."""
    g = ref_clust
    load_model(engine, g["model"])
    gpath = str(tmp_path / "c.gcl")
    open(gpath, "w").write(g["gcl"])
    n, gi, ci = oracle_np.parse_clustering(g["gcl"])
    for k, (mc, mg) in enumerate(g["settings"]):
        if k % 2 == 0:
            engine.read_clustering(gpath)
        else:
            engine.set_clustering(n, gi, ci)
        engine.set_clustering_min_evals(mc, mg)
        lik = engine.gmm_score(g["feats"], precision=F64)
        want = g["lik%d" % k]
        rel = np.abs(lik - want) / want
        assert rel.max() <= 4.5e-16, (k, rel.max())
    engine.read_clustering(gpath)
    engine.set_clustering_min_evals(0.0, 0.25)
    for nb in (2, 4):
        want = g["lna%d" % nb]
        for prec in (F64, F32):
            rec = engine.gmm_lna(g["feats"], precision=prec, lnabytes=nb)
            assert np.array_equal(rec.reshape(-1), want[5:]), (nb, prec, (rec.reshape(-1) != want[5:]).sum())
    # switched off: exact evaluation again
    engine.use_clustering(False)
    lik = engine.gmm_score(g["feats"], precision=F64)
    assert (np.abs(lik - g["lik_exact"]) / g["lik_exact"]).max() <= 4.5e-16
    with pytest.raises(AkuGpuError, match="seems insensible"):
        engine.set_clustering(60, gi, ci)
    load_model(engine, g["model"])          # loading a model clears the clustering
    assert (np.abs(engine.gmm_score(g["feats"], precision=F64) - g["lik_exact"]) / g["lik_exact"]).max() <= 4.5e-16


def test_model_cmllr_parity(engine, ref_cmllr, tmp_path):
    """Global model-level CMLLR (`model cmllr`, unitmode UNIT_NO; ConstrainedMllr / AdaptedGaussian, aku/ModelModules.cc):
    state likelihoods within CUDA-exp ulps of the reference's HmmSet after SpeakerConfig::set_speaker -- two speakers
    (one with a negative diagonal element: the factor is |prod diag(A)|), the default speaker (no transform), a
    revisit, and with Gaussian clustering on; every throughput scorer holds its usual bar on the adapted model."""
    from aaltoasr_b200 import SpeakerConfig
    g = ref_cmllr
    load_model(engine, g["model"])
    spkc = str(tmp_path / "m.spkc")
    open(spkc, "w").write(g["spkc"])
    sc = SpeakerConfig(engine)
    sc.read_speaker_file(spkc)
    for spk in ("alice", "bob", "carol", "alice"):
        sc.set_speaker(spk)
        lik = engine.gmm_score(g["feats"], precision=F64)
        want = g["lik_" + spk]
        rel = np.abs(lik - want) / want
        assert rel.max() <= 2e-15, (spk, rel.max())
    assert not np.array_equal(g["lik_alice"], g["lik_carol"])
    # direct call; every throughput scorer on the adapted model
    engine.model_set_cmllr(g["W_bob"].astype(np.float32).astype(np.float64))
    feats32 = g["feats"].astype(np.float32)
    Wb = g["W_bob"].astype(np.float32).astype(np.float64)
    # what the scorers are handed: float32 features, adapted in double, stored as float32 again
    adapted, factor = oracle_np.cmllr_adapt(Wb, feats32.astype(np.float64))
    want_ll = np.log(oracle_np.state_likelihoods(g["model"], adapted.astype(np.float32).astype(np.float64)) * factor)
    assert np.abs(want_ll - np.log(g["lik_bob"])).max() <= 1e-5
    for variant in (0, 1, 2, 3, 4):
        engine.set_scorer_variant(variant)
        try:
            ll = engine.gmm_score(feats32, precision=F32).astype(np.float64)
            err = np.abs(ll - want_ll) / (1 + np.abs(want_ll) / 40)
            assert err.max() <= 3e-5, (variant, err.max())
            raw4 = lna4(engine.gmm_lna(feats32, precision=F32, lnabytes=4, normalize=False)).reshape(ll.shape)
            assert np.abs(raw4 - want_ll).max() <= 2e-4 < 0.07 < abs(np.log(factor)), variant    # the factor reaches the un-normalised stream
        finally:
            engine.set_scorer_variant(0)
    lp = engine.gmm_logprobs(g["feats"], tiny=1e-30)
    assert np.abs(lp - np.log(np.maximum(g["lik_bob"], 1e-30))).max() <= 2e-4
    # Gaussian clustering on top (centres are wrapped as well)
    gpath = str(tmp_path / "c.gcl")
    open(gpath, "w").write(g["gcl"])
    engine.read_clustering(gpath)
    engine.set_clustering_min_evals(0.0, 0.25)
    lik = engine.gmm_score(g["feats"], precision=F64)
    assert (np.abs(lik - g["lik_clust_bob"]) / g["lik_clust_bob"]).max() <= 2e-15
    engine.use_clustering(False)
    # removal, errors, and a model load clears it
    engine.model_set_cmllr(None)
    assert (np.abs(engine.gmm_score(g["feats"], precision=F64) - g["lik_carol"]) / g["lik_carol"]).max() <= 4.5e-16
    W = g["W_alice"].copy()
    W[5, 6] = 0.0
    with pytest.raises(AkuGpuError, match="diagonal of A"):
        engine.model_set_cmllr(W)
    with pytest.raises(ValueError):
        engine.model_set_cmllr(np.zeros((39, 39)))
    engine.model_set_cmllr(g["W_alice"])
    load_model(engine, g["model"])
    assert (np.abs(engine.gmm_score(g["feats"], precision=F64) - g["lik_carol"]) / g["lik_carol"]).max() <= 4.5e-16


@pytest.mark.parametrize("D,max_mix", [(5, 3), (13, 16), (26, 40), (39, 64), (47, 9), (63, 20), (70, 12), (39, 90)])
def test_throughput_scorers_other_shapes(engine, D, max_mix):
    """Every instantiation of the default scorer (K16 chunk counts 1-8 = feature dims up to 63, the streaming variant
    beyond, mixtures of 1..64 components) and the FP32-pipe fallback (> 64 components per state) against the double
    path on random diagonal models; frame counts that are not multiples of the tile."""
    rng = np.random.default_rng(100 * D + max_mix)
    S = 37
    sizes = rng.integers(1, max_mix + 1, S)
    sizes[0] = max_mix
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    G = int(off[-1])
    F = 300
    feats = rng.standard_normal((F, D)) * rng.uniform(0.5, 3.0, D) + rng.uniform(-2, 2, D)
    means = feats[rng.integers(0, F, G)] + 0.5 * rng.standard_normal((G, D))
    covs = rng.uniform(0.3, 2.0, (G, D))
    wts = rng.uniform(0.1, 1.0, G)
    f32 = feats.astype(np.float32)
    # variant 3 = the tensor-core kernels whatever the model's conditioning (these random models are sharper than the
    # expanded form likes: its error grows with 1/2 sum (mu - c)^2 / var); variant 0 = default, which sends such models
    # to the direct-form kernel and holds the usual bar
    for variant, bar in ((3, 3e-4), (0, 1e-4)):
        engine.set_scorer_variant(variant)
        try:
            engine.model_load_diag(off, np.arange(G, dtype=np.int32), wts, means, covs)
            want = np.log(engine.gmm_score(f32.astype(np.float64), precision=F64))
            for n in (F, 129, 1):
                got = engine.gmm_score(f32[:n], precision=F32).astype(np.float64)
                live = want[:n] > -100
                err = (np.abs(got - want[:n]) / (1 + np.abs(want[:n]) / 40))[live]
                assert live.sum() == 0 or err.max() <= bar, (D, max_mix, variant, n, err.max())
                assert np.isfinite(got).all()
        finally:
            engine.set_scorer_variant(0)
    # LNA through the same kernels: float records within the contract of the parity mode's
    a = lna4(engine.gmm_lna(f32, precision=F32, lnabytes=4)).astype(np.float64)
    b = lna4(engine.gmm_lna(f32.astype(np.float64), precision=F64, lnabytes=4)).astype(np.float64)
    ok = (np.abs(a - b) <= REL_TOL * np.abs(b)) | (np.abs(a - b) <= 2e-5)
    assert ok.mean() >= 0.999, (D, max_mix, 1 - ok.mean())


@pytest.mark.parametrize("case", ["ref_small", "ref_edge"])
def test_decoder_logprob_feed(engine, case, request):
    """akugpu_gmm_logprobs = what decoder/decode-stream.cc:191-207 hands to OneFrameAcoustics per frame:
    (float) log(max(state_likelihood, 1e-30)), un-normalised.  Parity mode equals the oracle up to the float cast of a
    1-ulp exp difference; throughput mode within the usual bar; one frame at a time gives the same rows."""
    g = request.getfixturevalue(case)
    load_model(engine, g["model"])
    want = np.log(np.maximum(g["lik"], 1e-30)).astype(np.float32)
    got64 = engine.gmm_logprobs(g["feats"], precision=F64, tiny=1e-30)
    assert got64.dtype == np.float32 and got64.shape == want.shape
    assert np.abs(got64.astype(np.float64) - want).max() <= 8e-6 and (got64 != want).mean() < 0.01
    assert got64.min() >= np.float32(np.log(1e-30))
    got32 = engine.gmm_logprobs(g["feats"].astype(np.float32), precision=F32, tiny=1e-30)
    err = np.abs(got32.astype(np.float64) - want) / (1 + np.abs(want) / 40)
    assert err.max() <= 3e-5, err.max()
    for f in (0, 17, want.shape[0] - 1):       # the per-frame loop of the decoder
        row = engine.gmm_logprobs(g["feats"][f:f + 1], precision=F64, tiny=1e-30)
        assert np.array_equal(row[0], got64[f])
    with pytest.raises(AkuGpuError, match="tiny must be > 0"):
        engine.gmm_logprobs(g["feats"], precision=F64, tiny=0.0)


def test_hybrid_scoring_of_ill_conditioned_states(engine, ref_edge):
    """States with a component too sharp / too far from the centre for the expanded (GEMM) form are scored by the
    direct-form FP32-pipe kernel in the same pass, the rest by the tensor-core kernel: the edge-case model (means up to 9
    sigma out) in the default mode holds the direct-form bar on every state; a model where every second state is sharp
    goes to the FP32-pipe kernel entirely."""
    g = ref_edge
    load_model(engine, g["model"])
    assert engine.scorer_in_use() == 5 and engine.expanded_form_q() > 200
    feats32 = g["feats"].astype(np.float32)
    l0 = engine.launch_count()
    ll = engine.gmm_score(feats32, precision=F32).astype(np.float64)
    assert engine.launch_count() - l0 == 3                    # tensor-core kernel + FP32-pipe kernel + transpose
    want = np.log(oracle_np.state_likelihoods(g["model"], feats32.astype(np.float64)))
    live = want > -100
    err = (np.abs(ll - want) / (1 + np.abs(want) / 40))[live]
    assert err.max() <= 2e-5, err.max()
    for nb in (2, 4):                                         # and through the LNA epilogue (normaliser pass + row kernel)
        got = engine.gmm_lna(feats32, precision=F32, lnabytes=nb)
        ref = engine.gmm_lna(feats32.astype(np.float64), precision=F64, lnabytes=nb)
        if nb == 4:
            a, b = lna4(got).astype(np.float64), lna4(ref).astype(np.float64)
            ok = (np.abs(a - b) <= REL_TOL * np.abs(b)) | (np.abs(a - b) <= 2e-5)
            assert ok.mean() >= 0.999
        else:
            d = np.abs(codes2(got) - codes2(ref))
            assert (d <= 1).mean() >= 0.999
    # half of the states sharp: no tensor-core image at all
    m = dict(g["model"])
    covs = m["covs"].copy()
    off = m["mix_offsets"]
    for s in range(0, len(off) - 1, 2):
        covs[m["mix_gauss"][off[s]]] = np.abs(covs[m["mix_gauss"][off[s]]]) * 1e-3 + 1e-6
    m["covs"] = covs
    load_model(engine, m)
    assert engine.scorer_in_use() == 1
    load_model(engine, g["model"])


def test_fp16_range_fallback(engine, ref_small):
    """A feature far outside the fp16 range of the default scorer's scaled terms makes the call fall back to the
    bf16x3 kernel: results stay finite and the other frames are unchanged."""
    g = ref_small
    load_model(engine, g["model"])
    feats = g["feats"].astype(np.float32).copy()
    clean = engine.gmm_score(feats, precision=F32)
    feats[7, 3] = 3.0e4
    l0 = engine.launch_count()
    got = engine.gmm_score(feats, precision=F32)
    assert engine.launch_count() - l0 >= 4          # first attempt + the bf16x3 redo (expansion, scorer, transposes)
    assert np.isfinite(got).all() and got[7].max() < -1e5
    keep = np.arange(feats.shape[0]) != 7
    assert np.abs(got[keep] - clean[keep]).max() <= 3e-5
    # and the default kernel is back for the next call
    l0 = engine.launch_count()
    again = engine.gmm_score(g["feats"].astype(np.float32), precision=F32)
    assert engine.launch_count() - l0 == 2 and np.array_equal(again, clean)


def test_full_covariance_pool(engine, ref_full, tmp_path):
    """FullCovarianceGaussian path (exponential form, double): mixed diag/full pool read from .gk,
    and an all-full pool through akugpu_model_load_full."""
    from aaltoasr_b200 import formats
    g = ref_full
    base = str(tmp_path / "full")
    formats.write_model(base, **g["model"])
    engine.model_read(base)
    assert engine.num_states == 8 and engine.num_gaussians == 24
    lik = engine.gmm_score(g["feats"], precision=F64)
    assert (np.abs(lik - g["lik"]) / g["lik"]).max() <= 1e-10
    ll32 = engine.gmm_score(g["feats"], precision=F32)
    assert np.abs(ll32 - np.log(g["lik"])).max() <= 1e-5
    for nb in (2, 4):
        rec = engine.gmm_lna(g["feats"], precision=F64, lnabytes=nb)
        want = g["lna%d" % nb][5:]
        if nb == 4:
            x, y = lna4(rec).reshape(-1), want.view("<f4")
            assert (np.abs(x - y) / np.abs(y)).max() <= 1e-6
        else:
            d = np.abs(codes2(rec).reshape(-1) - want.view(">u2").astype(np.int64))
            assert d.max() <= 1 and (d != 0).mean() <= 1e-3
    # all-full pool through the direct loader vs the oracle
    m = g["model"]
    fm = m["full_mask"]
    idx = np.nonzero(fm)[0]
    sub = dict(mix_offsets=np.arange(0, len(idx) + 1, 3, dtype=np.int32), mix_gauss=np.arange(len(idx), dtype=np.int32),
               mix_weight=np.ones(len(idx)), means=m["means"][idx], covs=m["covs"][idx], full_covs=m["full_covs"][idx],
               full_mask=np.ones(len(idx), dtype=bool))
    engine.model_load_full(sub["mix_offsets"], sub["mix_gauss"], sub["mix_weight"], sub["means"], sub["full_covs"])
    got = engine.gmm_score(g["feats"][:32], precision=F64)
    want = oracle_np.state_likelihoods(sub, g["feats"][:32])
    assert (np.abs(got - want) / want).max() <= 1e-10
    # end to end from PCM with the full-covariance model
    engine.frontend_load_config_text(g["cfg"])
    engine.model_read(base)
    rec4, fo, _ = engine.phone_probs(g["pcm"], lnabytes=4)
    y = g["lna4"][5:].view("<f4").reshape(rec4.shape[0], -1)
    assert (np.abs(lna4(rec4) - y) / np.abs(y)).max() <= 1e-4


@pytest.mark.parametrize("precision", [F32, F64])
def test_phone_probs_end_to_end(engine, ref_small, precision):
    """PCM -> LNA through the fused batch entry point vs the files the literal phone_probs wrote."""
    g = ref_small
    engine.frontend_load_config_text(g["cfg"])
    load_model(engine, g["model"])
    F, S = g["lik"].shape
    rec4, fo, _ = engine.phone_probs(g["pcm"], precision=precision, lnabytes=4)
    assert list(fo) == [0, F] and rec4.shape == (F, S * 4)                 # frame indexing: identical
    want4 = lna4(g["lna4"][5:]).reshape(F, S)
    rel = np.abs(lna4(rec4) - want4) / np.abs(want4)
    assert rel.max() <= REL_TOL, rel.max()
    rec2, _, chk = engine.phone_probs(g["pcm"], precision=precision, lnabytes=2, checksum=True)
    assert chk == int(rec2.astype(np.uint64).sum())
    d = np.abs(codes2(rec2) - codes2(g["lna2"][5:]).reshape(F, S))
    assert d.max() <= 1 and (d != 0).mean() <= 0.03, (d.max(), (d != 0).mean())
    # discard mode (kernel-only timing path) gives the same checksum
    _, _, chk2 = engine.phone_probs(g["pcm"], precision=precision, lnabytes=2, discard=True, checksum=True)
    assert chk2 == chk


def test_phone_probs_batch_and_chunking(engine, ref_small):
    """Utterance batching and frame chunking never change a byte."""
    g = ref_small
    engine.frontend_load_config_text(g["cfg"])
    load_model(engine, g["model"])
    pcm = g["pcm"]
    cuts = [pcm[:7000], pcm, pcm[5000:20000]]
    uo = np.concatenate([[0], np.cumsum([c.size for c in cuts])])
    try:
        engine.set_chunk_frames(128)
        small, fo, _ = engine.phone_probs(np.concatenate(cuts), uo, lnabytes=2)
        engine.set_chunk_frames(0)
        big, fo2, _ = engine.phone_probs(np.concatenate(cuts), uo, lnabytes=2)
    finally:
        engine.set_chunk_frames(0)
    assert np.array_equal(fo, fo2) and np.array_equal(small, big)
    for k, c in enumerate(cuts):
        one, _, _ = engine.phone_probs(c, lnabytes=2)
        assert np.array_equal(big[fo[k]:fo[k + 1]], one)


def test_device_buffers(engine, ref_small):
    """Device-resident inputs/outputs (torch tensors) give the same bytes as host buffers."""
    import torch
    g = ref_small
    engine.frontend_load_config_text(g["cfg"])
    load_model(engine, g["model"])
    host, fo, _ = engine.phone_probs(g["pcm"], lnabytes=2)
    pcm_d = torch.from_numpy(g["pcm"]).cuda()
    out_d = torch.empty(host.shape, dtype=torch.uint8, device="cuda")
    engine.phone_probs(pcm_d, lnabytes=2, out=out_d)
    assert np.array_equal(out_d.cpu().numpy(), host)
    pinned = torch.empty(host.shape, dtype=torch.uint8).pin_memory()
    engine.phone_probs(pcm_d, lnabytes=2, out=pinned)
    assert np.array_equal(pinned.numpy(), host)


# ------------------------------------------------------------------ BASELINE-size properties
@pytest.fixture(scope="module")
def big_case(engine):
    """Config-2 model (5000 states x 16 mixtures x 39 dims) on 20 s of synthetic audio."""
    from aaltoasr_b200 import synth
    cfg = synth.mfcc39_config()
    pcm = np.concatenate([synth.synth_audio(2000 + i, 160000) for i in range(2)])
    uo = np.array([0, 160000, 320000])
    engine.frontend_load_config_text(cfg)
    feats, fo = engine.features(pcm, uo, dtype=np.float64)
    model = synth.synth_diag_model(2999, feats, 5000, 16)
    return dict(cfg=cfg, pcm=pcm, uo=uo, feats=feats, fo=fo, model=model)


def test_full_size_properties(engine, big_case):
    b = big_case
    engine.frontend_load_config_text(b["cfg"])
    load_model(engine, b["model"])
    assert engine.num_states == 5000 and engine.num_gaussians == 80000
    rec4, fo, _ = engine.phone_probs(b["pcm"], b["uo"], lnabytes=4)
    assert list(fo) == [0, 1248, 2496]
    lp = lna4(rec4).astype(np.float64)
    # (a) a normalised frame sums to one (checksum over states)
    assert np.abs(np.exp(lp).sum(axis=1) - 1).max() <= 2e-5
    # (b) throughput vs parity mode on a sample of frames, full model
    idx = np.arange(0, 2496, 96)
    f64 = lna4(engine.gmm_lna(b["feats"][idx], precision=F64, lnabytes=4)).astype(np.float64)
    rel = np.abs(lp[idx] - f64) / np.abs(f64)
    assert rel.max() <= REL_TOL, rel.max()
    # (c) the oracle on a few frames (seconds on a CPU core)
    o_lik = oracle_np.state_likelihoods(b["model"], b["feats"][idx[:4]])
    _, o_lp = oracle_np.lna_records(o_lik, 4)
    assert np.array_equal(lna4(engine.gmm_lna(b["feats"][idx[:4]], precision=F64, lnabytes=4)), o_lp)
    # (d) frame permutation commutes with scoring
    perm = np.random.default_rng(3).permutation(len(idx))
    a = engine.gmm_lna(b["feats"][idx].astype(np.float32), lnabytes=2)
    c = engine.gmm_lna(b["feats"][idx][perm].astype(np.float32), lnabytes=2)
    assert np.array_equal(a[perm], c)
    # (d2) the row-streaming LNA kernel and the column kernel write identical bytes at full width
    import os
    os.environ["AKUGPU_LNA_OLD"] = "1"
    try:
        c_old = engine.gmm_lna(b["feats"][idx][perm].astype(np.float32), lnabytes=2)
    finally:
        del os.environ["AKUGPU_LNA_OLD"]
    assert np.array_equal(c, c_old)
    # (e) both kernel variants within tolerance of each other
    engine.set_scorer_variant(1)
    try:
        v1 = lna4(engine.gmm_lna(b["feats"][idx].astype(np.float32), lnabytes=4)).astype(np.float64)
    finally:
        engine.set_scorer_variant(0)
    assert (np.abs(v1 - f64) / np.abs(f64)).max() <= REL_TOL


# ------------------------------------------------------------------ error behaviour
def test_errors(engine, ref_small):
    with pytest.raises(AkuGpuError, match="SRNormModule: Must set both in_frames and out_frames"):
        engine.frontend_load_config_text("module\n{\n name a\n type audiofile\n sample_rate 16000\n}\nmodule\n{\n name f\n type fft\n sources a\n}\n"
                                         "module\n{\n name s\n type sr_norm\n sources f\n}\n")
    with pytest.raises(AkuGpuError, match="Unknown module type"):
        engine.frontend_load_config_text("module\n{\n name a\n type nonsense\n}\n")
    with pytest.raises(AkuGpuError, match="first module should be a base module"):
        engine.frontend_load_config_text("module\n{\n name a\n type fft\n}\n")
    with pytest.raises(AkuGpuError, match="Must set sample rate"):
        engine.frontend_load_config_text("module\n{\n name a\n type audiofile\n}\n")
    with pytest.raises(AkuGpuError, match="unknown source module"):
        engine.frontend_load_config_text("module\n{\n name a\n type audiofile\n sample_rate 16000\n}\nmodule\n{\n name f\n type fft\n sources b\n}\n")
    engine.frontend_load_config_text(ref_small["cfg"])
    with pytest.raises(AkuGpuError, match="audio shorter than frame"):
        engine.features(np.zeros(100, dtype=np.int16))
    m = dict(ref_small["model"])
    m["means"] = m["means"][:, :20]
    m["covs"] = m["covs"][:, :20]
    load_model(engine, m)
    with pytest.raises(AkuGpuError, match=r"Gaussian dimension is 20 but feature dimension is 39\."):     # aku/phone_probs.cc:119-124
        engine.phone_probs(ref_small["pcm"])
    with pytest.raises(AkuGpuError, match="could not open"):
        engine.model_read("/nonexistent/model")


def test_cmllr_regression_classes(engine, ref_cmllr_units, tmp_path):
    """Regression-class model-level CMLLR (unitmode UNIT_PHONE / UNIT_MIX / UNIT_GAUSSIAN) on the GPU double path: every
    Gaussian sees its class's A f + b (gmm_diag_f64 picks the row by class) and its class's factor; likelihoods equal the
    reference's aku::HmmSet after SpeakerConfig::set_speaker, LNA bytes the literal phone_probs -S files; throughput-mode
    requests are served by the same path; removing the transforms restores the plain model."""
    from aaltoasr_b200 import SpeakerConfig, formats
    g = ref_cmllr_units
    base = str(tmp_path / "m")
    formats.write_model(base, **g["model"])
    open(base + ".ph", "w").write(g["ph"])
    engine.frontend_load_config_text(g["cfg"])
    engine.model_read(base)                                   # UNIT_PHONE needs the phone table of the files
    plain = engine.gmm_score(g["feats"], precision=F64)
    assert (np.abs(plain - g["lik_plain"]) / g["lik_plain"]).max() <= 1e-12
    spkc = str(tmp_path / "x.spkc")
    open(spkc, "w").write(g["spkc"])
    sc = SpeakerConfig(engine)
    sc.read_speaker_file(spkc)
    for spk in ("phone", "mix", "gauss"):
        sc.set_speaker(spk)
        assert engine.scorer_in_use() in (0, 3)              # whatever the image, requests go to the double path
        lik = engine.gmm_score(g["feats"], precision=F64)
        rel = np.abs(lik - g["lik_" + spk]) / g["lik_" + spk]
        assert rel.max() <= 1e-12, (spk, rel.max())
        for nb in (2, 4):
            want = g["lna%d_%s" % (nb, spk)][5:]
            assert np.array_equal(engine.gmm_lna(g["feats"], precision=F64, lnabytes=nb).reshape(-1), want), (spk, nb)
            # F32 requests and float features: same path, features rounded to float
            got = engine.gmm_lna(g["feats"].astype(np.float32), precision=F32, lnabytes=nb).reshape(-1)
            if nb == 2:
                d = np.abs(got.view(">u2").astype(int) - want.view(">u2").astype(int))
                assert d.max() <= 1 and (d != 0).mean() <= 0.02
        ll32 = engine.gmm_score(g["feats"].astype(np.float32), precision=F32)
        assert np.abs(ll32 - np.log(g["lik_" + spk])).max() <= 2e-4
        row = engine.gmm_logprobs(g["feats"][5:6], precision=F32, tiny=1e-30)          # the small-call path declines, general path serves
        assert np.abs(row[0] - np.log(g["lik_" + spk][5])).max() <= 2e-4
    # direct API: Gaussian units on a model loaded from memory; phones are not available there
    load_model(engine, g["model"])
    trs = [([str(u) for u in g["units_gauss_%d" % i]], g["W_gauss_%d" % i]) for i in range(2)]
    engine.model_set_cmllr_units("UNIT_GAUSSIAN", trs)
    lik = engine.gmm_score(g["feats"], precision=F64)
    assert (np.abs(lik - g["lik_gauss"]) / g["lik_gauss"]).max() <= 1e-12
    with pytest.raises(AkuGpuError, match="phone table"):
        engine.model_set_cmllr_units("UNIT_PHONE", [(["a"], g["W_phone_0"])])
    with pytest.raises(AkuGpuError, match="out of range"):
        engine.model_set_cmllr_units("UNIT_MIX", [(["24"], g["W_mix_0"])])
    engine.model_set_cmllr_units("UNIT_GAUSSIAN", [])
    assert (np.abs(engine.gmm_score(g["feats"], precision=F64) - g["lik_plain"]) / g["lik_plain"]).max() <= 1e-12
    assert engine.scorer_in_use() == 3
