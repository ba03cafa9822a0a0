"""GPU parity at the BASELINE.json shapes, DIRECTLY against the reference's own code compiled into oracle/_ref (it travels
to the GPU box): the literal aku/phone_probs.cc binary and aku::HmmSet through libref_capi.so.

  config 1   one 10 s 16 kHz WAV (1248 frames), 39-dim MFCC, 100 states x 8 mixtures:
             the literal tool's LNA file == akugpu F64 on the reference's features, byte for byte; from the WAV (our
             features, <= 1e-5 from the reference's) and in F32 the differing-code fraction is measured and bounded
  config 2   5000 x 16: the benchmarked kernel (gmm_tc16_kernel<5>, F32 mode) against aku::HmmSet on 256 frames
  config 4   10000 x 32 (two slots per state, 1250 component tiles): the same on 64 frames
  config 5   2000 x 16 full covariance: the double path and gmm_tc16_kernel<0> against aku::HmmSet on 64 frames
The reference scores ~50 frames/s/core at 5000 x 16, so these cost seconds each.  For configs 4 / 5 the reference
reads a sub-model (512 / 64 states of the full model: a state's score does not depend on the other states; the text
model files of the full models are 0.5 - 1 GB) while the GPU scores the full-size model.
"""
import os
import subprocess

import numpy as np
import pytest

from aaltoasr_b200 import F32, F64, formats, synth
from oracle import oracle_np, ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")]

REL_TOL = 1e-4            # north_star: float log-probs within 1e-4 relative
ABS_TOL = 2e-5            # ... where the value itself is ~0 (the dominant state of a frame: lp ~ -1e-7)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lna4(rec):
    return np.ascontiguousarray(rec).view("<f4")


def codes2(rec):
    return np.ascontiguousarray(rec).view(">u2").astype(np.int64)


def load_model(engine, m):
    engine.model_load_diag(m["mix_offsets"], m["mix_gauss"], m["mix_weight"], m["means"], m["covs"])


def sub_model(m, states, n_mix, full=False):
    """The listed states of a model with n_mix components each (no sharing), as a model of its own."""
    states = np.asarray(states)
    g = (states[:, None] * n_mix + np.arange(n_mix)[None, :]).reshape(-1)
    out = dict(mix_offsets=np.arange(0, len(g) + 1, n_mix, dtype=np.int32), mix_gauss=np.arange(len(g), dtype=np.int32),
               mix_weight=m["mix_weight"][g], means=m["means"][g])
    if full:
        out["full_covs"] = m["full_covs"][g]
    else:
        out["covs"] = m["covs"][g]
    return out


def check_logprobs(got, want, what, frac_bar=1.0, abs_tol=ABS_TOL):
    """float log-probs: within REL_TOL relative, or abs_tol absolute where the value is ~0."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    err = np.abs(got - want)
    ok = (err <= REL_TOL * np.abs(want)) | (err <= abs_tol)
    assert ok.mean() >= frac_bar, (what, 1 - ok.mean(), err.max(), (err / np.maximum(np.abs(want), 1e-30)).max())
    return err.max()


@pytest.fixture(scope="module")
def wav10s(tmp_path_factory):
    d = tmp_path_factory.mktemp("c1")
    pcm = synth.synth_audio(1001, 160000, 16000)
    wav = str(d / "utt.wav")
    formats.write_wav(wav, pcm, 16000)
    cfg = str(d / "mfcc.cfg")
    open(cfg, "w").write(synth.mfcc39_config(16000))
    feats, last, fr = ref.features(cfg, wav)                  # the reference's FeatureGenerator, frame by frame
    return dict(dir=d, pcm=pcm, wav=wav, cfg=cfg, cfg_text=synth.mfcc39_config(16000), ref_feats=feats)


# ------------------------------------------------------------------------------------------------ config 1
def test_config1_exact_shape_against_the_literal_tool(engine, wav10s, tmp_path, record_property):
    w = wav10s
    assert w["ref_feats"].shape == (1248, 39)
    engine.frontend_load_config_text(w["cfg_text"])
    assert engine.num_frames(w["pcm"].size) == 1248
    model = synth.synth_diag_model(1002, w["ref_feats"], 100, 8)
    base = str(tmp_path / "m100x8")
    formats.write_model(base, **model)
    rec = str(tmp_path / "recipe")
    open(rec, "w").write("audio=%s lna=utt.lna\n" % w["wav"])
    engine.model_read(base)                                    # the same files the reference reads
    assert engine.num_states == 100 and engine.num_gaussians == 800
    ours_feats, fo = engine.features(w["pcm"], dtype=np.float64)
    assert list(fo) == [0, 1248] and np.abs(ours_feats - w["ref_feats"]).max() <= 1e-5
    tool = os.path.join(ROOT, "aaltoasr_b200", "akugpu_phone_probs")
    for nb in (2, 4):
        for flags in ((), ("-N",)):
            out = tmp_path / ("ref%d%s" % (nb, "n" if flags else ""))
            out.mkdir()
            ref.phone_probs(w["cfg"], base, rec, str(out), lnabytes=nb, extra=list(flags))      # literal aku/phone_probs.cc
            want = np.frombuffer(open(str(out / "utt.lna"), "rb").read(), dtype=np.uint8)
            assert want.size == 5 + 1248 * 100 * nb and bytes(want[:5]) == engine.lna_header(nb)
            # (a) parity mode on the reference's features: identical bytes
            got = engine.gmm_lna(w["ref_feats"], precision=F64, lnabytes=nb, normalize=not flags)
            assert np.array_equal(got.reshape(-1), want[5:]), (nb, flags)
            # (b) our own tool from the WAV, parity arithmetic: the file differs only where a feature's last bits do
            o2 = tmp_path / ("ours%d%s" % (nb, "n" if flags else ""))
            o2.mkdir()
            r = subprocess.run([tool, "-b", base, "-c", w["cfg"], "-r", rec, "-o", str(o2), "--lnabytes=%d" % nb, "--precision=f64"]
                               + list(flags), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
            assert r.returncode == 0, r.stderr.decode()
            mine = np.frombuffer(open(str(o2 / "utt.lna"), "rb").read(), dtype=np.uint8)
            assert mine.size == want.size and bytes(mine[:5]) == bytes(want[:5])
            # (c) throughput mode from the WAV
            f32rec, _, _ = engine.phone_probs(w["pcm"], precision=F32, lnabytes=nb, normalize=not flags)
            if nb == 2:
                for tag, rec_b in (("f64_from_wav", mine[5:]), ("f32_from_wav", f32rec.reshape(-1))):
                    d = np.abs(codes2(rec_b) - codes2(want[5:]))
                    frac = float((d != 0).mean())
                    record_property("config1_%s%s_differing_code_fraction" % (tag, "_nonorm" if flags else ""), frac)
                    print("config 1 %s %s: %.4f %% of 2-byte codes differ (all by %d)" % (tag, flags, 100 * frac, d.max()))
                    assert d.max() <= 1 and frac <= (0.01 if tag.startswith("f64") else 0.03), (tag, flags, d.max(), frac)
            else:
                # from the WAV the features differ from the reference's by up to 1e-5 (FFT rounding order): un-normalised
                # log-likelihoods near zero move by a few 1e-5 in absolute terms whatever the scorer's arithmetic
                tol = 1e-4 if flags else ABS_TOL
                e64 = check_logprobs(lna4(mine[5:]), lna4(want[5:]), "f64 from wav", abs_tol=tol)
                e32 = check_logprobs(lna4(f32rec.reshape(-1)), lna4(want[5:]), "f32 from wav", abs_tol=tol)
                print("config 1 4-byte %s: max |error| f64-from-wav %.3g, f32-from-wav %.3g" % (flags, e64, e32))


# ------------------------------------------------------------------------------------------------ config 2
@pytest.fixture(scope="module")
def gpu_feats(engine, wav10s):
    """Features of two synthetic utterances from the GPU front-end (the models' means are drawn from them)."""
    engine.frontend_load_config_text(wav10s["cfg_text"])
    pcm = np.concatenate([synth.synth_audio(2000 + i, 160000) for i in range(2)])
    feats, _ = engine.features(pcm, np.array([0, 160000, 320000]), dtype=np.float64)
    return feats


def test_config2_benchmarked_kernel_against_reference_hmmset(engine, gpu_feats, tmp_path):
    """The kernel bench.py times (gmm_tc16_kernel<5>, 5000 x 16, F32 mode) against aku::HmmSet::state_likelihood on
    256 frames -- no intermediate GPU result in the comparison."""
    model = synth.synth_diag_model(2999, gpu_feats, 5000, 16)
    load_model(engine, model)
    assert engine.scorer_in_use() == 3 and engine.num_gaussians == 80000
    idx = np.arange(0, 2496, 2496 // 256)[:256]
    x = gpu_feats[idx]
    base = str(tmp_path / "m5000x16")
    formats.write_model(base, **model)
    M = ref.Model(base)
    try:
        assert (M.S, M.G, M.D) == (5000, 80000, 39)
        lik = M.state_likelihoods(x)                       # HmmSet::precompute_likelihoods + state_likelihood, per frame
    finally:
        M.close()
    want_ll = np.log(lik)
    ll = engine.gmm_score(x.astype(np.float32), precision=F32).astype(np.float64)
    # fp32 features: the reference scored the doubles; the difference this makes is part of what the bar covers
    err = np.abs(ll - want_ll)
    print("config 2: max |log-lik error| of gmm_tc16_kernel vs aku::HmmSet on 256 frames: %.3g" % err.max())
    assert err.max() <= 6e-5, err.max()
    # through the LNA epilogue: the contract itself
    want_rec4, want_lp = oracle_np.lna_records(lik, 4)
    got_lp = lna4(engine.gmm_lna(x.astype(np.float32), precision=F32, lnabytes=4))
    check_logprobs(got_lp, want_lp, "config 2 lna4")
    want_rec2, _ = oracle_np.lna_records(lik, 2)
    d = np.abs(codes2(engine.gmm_lna(x.astype(np.float32), precision=F32, lnabytes=2)) - codes2(want_rec2))
    print("config 2: %.3f %% of 2-byte codes differ from the reference's (max %d)" % (100 * (d != 0).mean(), d.max()))
    assert d.max() <= 1 and (d != 0).mean() <= 0.03
    # parity mode: bytes identical to the reference arithmetic at this size
    assert np.array_equal(engine.gmm_lna(x, precision=F64, lnabytes=2), want_rec2)
    assert np.array_equal(engine.gmm_lna(x, precision=F64, lnabytes=4), want_rec4)
    lik64 = engine.gmm_score(x, precision=F64)
    assert (np.abs(lik64 - lik) / lik).max() <= 1e-12


# ------------------------------------------------------------------------------------------------ config 4 model
def test_config4_model_10000x32_against_reference_hmmset(engine, gpu_feats, tmp_path):
    """10000 states x 32 mixtures: every state owns TWO 16-component slots, 1250 component tiles -- a different slot /
    meta pattern than 5000 x 16.  GPU: the full model; reference: 512 of its states (first / middle / last)."""
    S, K = 10000, 32
    model = synth.synth_diag_model(4999, gpu_feats, S, K)
    load_model(engine, model)
    assert engine.scorer_in_use() == 3 and engine.num_gaussians == S * K
    x = gpu_feats[np.arange(0, 2496, 39)[:64]]
    states = np.concatenate([np.arange(0, 192), np.arange(4900, 5028), np.arange(S - 192, S)])
    base = str(tmp_path / "m10000x32_sub")
    formats.write_model(base, **sub_model(model, states, K))
    M = ref.Model(base)
    try:
        lik = M.state_likelihoods(x)
    finally:
        M.close()
    ll = engine.gmm_score(x.astype(np.float32), precision=F32).astype(np.float64)
    err = np.abs(ll[:, states] - np.log(lik))
    print("config 4 model: max |log-lik error| vs aku::HmmSet on 64 frames x 512 states: %.3g" % err.max())
    assert err.max() <= 6e-5, err.max()
    lik64 = engine.gmm_score(x, precision=F64)
    assert (np.abs(lik64[:, states] - lik) / lik).max() <= 1e-12
    # every state: throughput vs parity mode of the library, and the LNA contract through both
    assert np.abs(ll - np.log(lik64)).max() <= 6e-5
    a = lna4(engine.gmm_lna(x.astype(np.float32), precision=F32, lnabytes=4))
    b = lna4(engine.gmm_lna(x, precision=F64, lnabytes=4))
    check_logprobs(a, b, "config 4 lna4")
    d = np.abs(codes2(engine.gmm_lna(x.astype(np.float32), precision=F32, lnabytes=2)) - codes2(engine.gmm_lna(x, precision=F64, lnabytes=2)))
    assert d.max() <= 1 and (d != 0).mean() <= 0.03
    # the oracle restatement on the first frames of the sub-model: bit-identical doubles
    o = oracle_np.state_likelihoods(sub_model(model, states, K), x[:2])
    assert np.array_equal(o, lik[:2])


# ------------------------------------------------------------------------------------------------ config 5
def test_config5_full_covariance_2000x16_against_reference_hmmset(engine, gpu_feats, tmp_path):
    """2000 states x 16 full-covariance Gaussians (BASELINE config 5; generator = bench.py's): the double path AND the
    benchmarked tensor-core kernel gmm_tc16_kernel<0> against aku::HmmSet (FullCovarianceGaussian, exponential form)."""
    S, K, D = 2000, 16, 39
    G = S * K
    rng = np.random.default_rng(5999)
    sd = gpu_feats.std(axis=0)
    means = gpu_feats[rng.integers(0, gpu_feats.shape[0], G)] + 0.3 * sd * rng.standard_normal((G, D))
    A = rng.standard_normal((G, D, 4)) * sd[None, :, None]
    full = np.einsum("gik,gjk->gij", A, A) * 0.1
    full[:, np.arange(D), np.arange(D)] += rng.uniform(0.5, 2, (G, D)) * sd ** 2
    model = dict(mix_offsets=np.arange(0, G + 1, K, dtype=np.int32), mix_gauss=np.arange(G, dtype=np.int32),
                 mix_weight=rng.dirichlet(np.ones(K), S).reshape(-1), means=means, full_covs=full)
    engine.model_load_full(model["mix_offsets"], model["mix_gauss"], model["mix_weight"], means, full)
    assert engine.scorer_in_use() == 4
    x = gpu_feats[np.arange(0, 2496, 39)[:64]]
    states = np.concatenate([np.arange(0, 24), np.arange(1000, 1016), np.arange(S - 24, S)])    # the reference's full-covariance load costs ~18 ms per Gaussian
    base = str(tmp_path / "m2000x16full_sub")
    formats.write_model(base, **sub_model(model, states, K, full=True))
    M = ref.Model(base)
    try:
        lik = M.state_likelihoods(x)
    finally:
        M.close()
    lik64 = engine.gmm_score(x, precision=F64)
    rel = np.abs(lik64[:, states] - lik) / lik
    print("config 5: double path vs aku::HmmSet, max relative error %.3g" % rel.max())
    assert rel.max() <= 1e-9                                # LAPACK-free inverse vs the reference's: ~1e-13 expected
    ll = engine.gmm_score(x.astype(np.float32), precision=F32).astype(np.float64)
    err = np.abs(ll[:, states] - np.log(lik))
    print("config 5: gmm_tc16_kernel<0> vs aku::HmmSet, max |log-lik error| %.3g" % err.max())
    assert err.max() <= 2e-4, err.max()                      # K = 819 cancellation: DESIGN.md 4.1a'
    assert np.abs(ll - np.log(lik64)).max() <= 2e-4          # and every state against the double path
    want2, _ = oracle_np.lna_records(lik64, 2)
    d = np.abs(codes2(engine.gmm_lna(x.astype(np.float32), precision=F32, lnabytes=2)) - codes2(want2))
    print("config 5: %.3f %% of 2-byte codes differ (max %d)" % (100 * (d != 0).mean(), d.max()))
    # a log-likelihood error of 8e-5 is 0.15 code steps (1820 codes per unit): up to ~8 % of the codes land on the other
    # side of a rounding boundary; never by more than one
    assert d.max() <= 1 and (d != 0).mean() <= 0.10


# ------------------------------------------------------------------------------------------------ advisor items
def test_hybrid_model_redone_after_fp16_overflow_keeps_the_direct_form_states(engine, ref_edge):
    """A hybrid model (ill-conditioned states on the FP32-pipe kernel) whose call is redone with the bf16x3 kernel after an
    fp16 range overflow: the ill-conditioned states must still come from the direct form (they used to be scored in the
    expanded form the packer refuses for them)."""
    g = ref_edge
    load_model(engine, g["model"])
    assert engine.scorer_in_use() == 5
    feats = g["feats"].astype(np.float32).copy()
    clean = engine.gmm_score(feats, precision=F32)
    feats[5, 2] = 3.0e4                                       # leaves the fp16 range of the scaled terms -> redo
    got = engine.gmm_score(feats, precision=F32)
    keep = np.arange(feats.shape[0]) != 5
    live = clean[keep] > -100
    err = (np.abs(got[keep] - clean[keep]) / (1 + np.abs(clean[keep]) / 40))[live]
    assert err.max() <= 3e-5, err.max()
    a = engine.gmm_lna(feats[keep], precision=F32, lnabytes=2)
    b = engine.gmm_lna(np.vstack([feats[keep][:3], feats[5:6], feats[keep][3:]]), precision=F32, lnabytes=2)   # redone call
    d = np.abs(codes2(np.vstack([b[:3], b[4:]])) - codes2(a))
    assert (d <= 1).mean() >= 0.999


def test_misaligned_device_output_is_refused(engine, ref_small):
    import torch
    from aaltoasr_b200 import AkuGpuError
    g = ref_small
    load_model(engine, g["model"])
    S = engine.num_states
    F = 8
    buf = torch.empty(F * S * 2 + 8, dtype=torch.uint8, device="cuda")
    with pytest.raises(AkuGpuError, match="4-byte aligned"):
        engine.gmm_lna(g["feats"][:F].astype(np.float32), lnabytes=2, out=buf[5:5 + F * S * 2])
    ok = engine.gmm_lna(g["feats"][:F].astype(np.float32), lnabytes=2, out=buf[4:4 + F * S * 2])
    assert np.array_equal(ok.cpu().numpy().reshape(F, -1), engine.gmm_lna(g["feats"][:F].astype(np.float32), lnabytes=2))
    torch.cuda.synchronize()


def test_lna_bytes_do_not_depend_on_the_launch_shape(engine, gpu_feats):
    """A frame's records are the same bytes whether it was scored in a full chunk (one CTA sweeps every component tile and
    its epilogue produces the normaliser) or in a short call whose component tiles are spread over grid.y (the
    normaliser is then replayed from the stored scores in the epilogue's order, tc16_norm_replay): per-utterance
    checksums of differently batched / differently sharded runs are compared on this."""
    model = synth.synth_diag_model(2999, gpu_feats, 5000, 16)
    load_model(engine, model)
    assert engine.scorer_in_use() == 3
    pcm = np.concatenate([synth.synth_audio(2000 + i, 160000) for i in range(20)])       # 24 960 frames: more than one wave
    uo = np.arange(21, dtype=np.int64) * 160000
    big, fo, chk = engine.phone_probs(pcm, uo, lnabytes=2, utt_checksums=True)
    for u in (0, 7, 19):
        one, _, c1 = engine.phone_probs(pcm[uo[u]:uo[u + 1]], lnabytes=2, utt_checksums=True)      # 1248 frames: tiles split over grid.y
        assert np.array_equal(one, big[fo[u]:fo[u + 1]]) and c1[0] == chk[u]
    feats32 = gpu_feats[:300].astype(np.float32)
    a = engine.gmm_lna(feats32, lnabytes=4)
    b = np.vstack([engine.gmm_lna(feats32[i:i + 100], lnabytes=4) for i in range(0, 300, 100)])
    assert np.array_equal(a, b)
    try:
        engine.set_chunk_frames(128 * 148)                 # one wave per chunk + a ragged tail
        c = engine.gmm_lna(np.vstack([feats32] * 70), lnabytes=4)
    finally:
        engine.set_chunk_frames(0)
    assert np.array_equal(c[:300], a) and np.array_equal(c[-300:], a)


def test_overlapped_lna_pipeline_writes_the_same_bytes(engine, gpu_feats, monkeypatch):
    """The LNA epilogue of chunk k runs on a second stream next to the scorer of chunk k+1 (double-buffered scores): same
    bytes as the serial pipeline (AKUGPU_OVERLAP=0), device and host outputs, checksums, several chunks + a ragged tail."""
    import torch
    from aaltoasr_b200 import AkuGpu
    model = synth.synth_diag_model(2999, gpu_feats, 5000, 16)
    pcm = np.concatenate([synth.synth_audio(2000 + i, 160000) for i in range(70)])       # 87 360 frames: 3 chunks, the last one ragged
    uo = np.arange(71, dtype=np.int64) * 160000
    res = {}
    for tag, env in (("overlap", "1"), ("serial", "0")):
        monkeypatch.setenv("AKUGPU_OVERLAP", env)
        with AkuGpu(0) as eng:
            eng.frontend_load_config_text(synth.mfcc39_config(16000))
            load_model(eng, model)
            host, fo, chk = eng.phone_probs(pcm, uo, lnabytes=2, utt_checksums=True)
            dev = torch.empty(host.shape, dtype=torch.uint8, device="cuda")
            _, _, tot = eng.phone_probs(torch.from_numpy(pcm).cuda(), uo, lnabytes=2, out=dev, checksum=True)
            f4 = eng.gmm_lna(gpu_feats[:2000].astype(np.float32), lnabytes=4, normalize=False)
            res[tag] = (host, chk, dev.cpu().numpy(), tot, f4)
    a, b = res["overlap"], res["serial"]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3]
    assert np.array_equal(a[0], a[2]) and a[3] == int(a[0].astype(np.uint64).sum()) and np.array_equal(a[4], b[4])
