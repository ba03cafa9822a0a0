"""GPU tests of the host side above the C ABI: the Python mirrors of FeatureGenerator / HmmSet /
phone_probs, and the C++ tool aaltoasr_b200/akugpu_phone_probs (same flags as aku/phone_probs.cc)."""
import os
import subprocess

import numpy as np
import pytest

from aaltoasr_b200 import F32, F64, FeatureGenerator, HmmSet, PhoneProbs, formats

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "aaltoasr_b200", "akugpu_phone_probs")


def write_case(tmp_path, g, n=3):
    cfg = str(tmp_path / "mfcc.cfg")
    open(cfg, "w").write(g["cfg"])
    base = str(tmp_path / "model")
    formats.write_model(base, **g["model"])
    pcm = g["pcm"]
    cuts = [pcm, pcm[:9000], pcm[4000:20000]][:n]
    lines = []
    for i, c in enumerate(cuts):
        w = str(tmp_path / ("utt%d.wav" % i))
        formats.write_wav(w, c, 16000)
        lines.append("audio=%s lna=utt%d.lna" % (w, i))
    rec = str(tmp_path / "recipe")
    open(rec, "w").write("\n".join(lines) + "\n")
    return cfg, base, rec, cuts


def test_feature_generator_and_hmmset_mirror(engine, ref_small, tmp_path):
    """The reference's per-frame call pattern (aku/phone_probs.cc:217-230) on the mirrors."""
    g = ref_small
    cfg, base, _, _ = write_case(tmp_path, g, 1)
    gen = FeatureGenerator(engine)
    gen.load_configuration(cfg)
    gen.open(str(tmp_path / "utt0.wav"))
    model = HmmSet(engine, precision=F64)
    model.read_all(base)
    assert gen.dim() == model.dim() == 39 and gen.frame_rate() == 125.0
    assert gen.last_frame() == int(g["last_frame"]) and model.num_states() == g["lik"].shape[1]
    model.set_utterance_features(g["feats"])          # score the reference's features: doubles must match
    f = 0
    while True:
        fea = gen.generate(f)
        if gen.eof():
            break
        assert np.abs(fea - g["feats"][f]).max() <= 1e-5
        model.reset_cache()
        model.precompute_likelihoods(f)
        for s in (0, 5, model.num_states() - 1):
            assert abs(model.state_likelihood(s) - g["lik"][f, s]) <= 4.5e-16 * g["lik"][f, s]
        f += 1
    assert f == g["feats"].shape[0]
    assert np.abs(gen.generate(-3) - g["feats_ext"][-3 - int(g["ext_start"])]).max() <= 1e-5
    # a feature vector instead of a frame index: the F=1 case of the same kernel
    model.reset_cache()
    model.precompute_likelihoods(g["feats"][7])
    assert abs(model.state_likelihood(3) - g["lik"][7, 3]) <= 4.5e-16 * g["lik"][7, 3]


@pytest.mark.parametrize("lnabytes", [2, 4])
def test_cpp_tool_matches_python_and_reference(engine, ref_small, tmp_path, lnabytes):
    g = ref_small
    cfg, base, rec, cuts = write_case(tmp_path, g)
    out_cpp, out_py = tmp_path / "cpp", tmp_path / "py"
    out_cpp.mkdir(); out_py.mkdir()
    r = subprocess.run([TOOL, "-b", base, "-c", cfg, "-r", rec, "-o", str(out_cpp), "--lnabytes=%d" % lnabytes, "-i", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert r.returncode == 0, r.stderr.decode()
    pp = PhoneProbs(engine, precision=F32, lnabytes=lnabytes)
    pp.read_configuration(cfg)
    pp.read_models(base)
    assert pp.run_recipe(rec, str(out_py)) == len(cuts)
    for i in range(len(cuts)):
        a = open(str(out_cpp / ("utt%d.lna" % i)), "rb").read()
        b = open(str(out_py / ("utt%d.lna" % i)), "rb").read()
        assert a == b                                     # both hosts drive the same library
    # utterance 0 is the golden case: header identical, records within the throughput-mode bar
    want = g["lna%d" % lnabytes]
    got = np.frombuffer(open(str(out_cpp / "utt0.lna"), "rb").read(), dtype=np.uint8)
    assert bytes(got[:5]) == bytes(want[:5]) and got.size == want.size
    if lnabytes == 4:
        x, y = got[5:].view("<f4"), want[5:].view("<f4")
        assert (np.abs(x - y) / np.abs(y)).max() <= 1e-4
    else:
        d = np.abs(got[5:].view(">u2").astype(int) - want[5:].view(">u2").astype(int))
        assert d.max() <= 1 and (d != 0).mean() <= 0.03
    lp, S, nb = formats.read_lna(str(out_cpp / "utt0.lna"))      # decoder-side reader contract
    assert (S, nb) == (g["lik"].shape[1], lnabytes) and lp.shape == g["lik"].shape
    # ... and the decoder's own reader (decoder/src/LnaReaderCircular.cc, built into oracle/_ref) on our files
    from oracle import ref
    assert ref.available(), "oracle/_ref (prebuilt) did not travel to the GPU box"
    for i in range(len(cuts)):
        dec = ref.lna_read(str(out_cpp / ("utt%d.lna" % i)))
        mine, S2, _ = formats.read_lna(str(out_cpp / ("utt%d.lna" % i)))
        assert dec.shape == (engine.num_frames(cuts[i].size), S) and np.array_equal(dec, np.asarray(mine, dtype=np.float32))


def test_cpp_tool_flags(engine, ref_small, tmp_path):
    g = ref_small
    cfg, base, rec, cuts = write_case(tmp_path, g)
    out = tmp_path / "o"; out.mkdir()
    run = lambda *a: subprocess.run([TOOL] + list(a), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    # -B/-I batching: batch 2 of 2 holds only the last utterance (contiguous split, aku/Recipe.cc:63-115)
    assert run("-b", base, "-c", cfg, "-r", rec, "-o", str(out), "-B", "2", "-I", "2").returncode == 0
    assert sorted(os.listdir(str(out))) == ["utt2.lna"]
    # --no-overwrite keeps an existing file untouched; parity precision and -N work together
    before = open(str(out / "utt2.lna"), "rb").read()
    assert run("-b", base, "-c", cfg, "-r", rec, "-o", str(out), "-n", "-N", "--precision=f64").returncode == 0
    assert open(str(out / "utt2.lna"), "rb").read() == before
    got = np.frombuffer(open(str(out / "utt0.lna"), "rb").read(), dtype=np.uint8)
    d = np.abs(got[5:].view(">u2").astype(int) - g["lna2_nonorm"][5:].view(">u2").astype(int))
    assert d.max() <= 1
    # errors like the reference's
    r = run("-c", cfg, "-r", rec)
    assert r.returncode != 0 and b"Must give either --base or all --gk, --mc and --ph" in r.stderr
    r = run("-b", base, "-c", cfg, "-r", rec, "--lnabytes=3")
    assert r.returncode != 0 and b"Invalid number of LNA bytes" in r.stderr
    r = run("-g", base + ".gk", "-m", base + ".mc", "-p", base + ".ph", "-c", cfg, "-r", rec, "-o", str(out), "-a")
    assert r.returncode == 0
    # output names as io::Stream takes them (aku/io.cc:35-130): *.gz through gzip, |command through a pipe
    import gzip
    rec2 = str(tmp_path / "recipe_gz")
    w0 = open(rec).read().split()[0]
    open(rec2, "w").write("%s lna=%s\n%s lna=|cat>%s\n" % (w0, str(tmp_path / "z.lna.gz"), w0, str(tmp_path / "piped.lna")))
    assert run("-b", base, "-c", cfg, "-r", rec2).returncode == 0
    plain = open(str(out / "utt0.lna"), "rb").read()
    assert gzip.open(str(tmp_path / "z.lna.gz"), "rb").read() == plain
    assert open(str(tmp_path / "piped.lna"), "rb").read() == plain


def test_cpp_tool_gaussian_clustering(engine, ref_clust, tmp_path):
    """akugpu_phone_probs -C x.gcl --eval-ming=0.25 writes the bytes the literal aku/phone_probs.cc wrote with the same
    flags, when fed the reference's features' audio is replaced by the parity path (f64) on the same WAV: the feature
    front-ends differ by ~1e-6, so codes are held to +-1; the GMM stage itself is checked byte-exactly in
    test_gpu_parity.py::test_gaussian_clustering_parity."""
    from aaltoasr_b200 import synth
    g = ref_clust
    cfg = str(tmp_path / "mfcc.cfg")
    open(cfg, "w").write(synth.mfcc39_config())
    base = str(tmp_path / "model")
    formats.write_model(base, **g["model"])
    gcl = str(tmp_path / "c.gcl")
    open(gcl, "w").write(g["gcl"])
    w = str(tmp_path / "utt0.wav")
    formats.write_wav(w, synth.synth_audio(7001, 24000), 16000)
    rec = str(tmp_path / "recipe")
    open(rec, "w").write("audio=%s lna=utt0.lna\n" % w)
    out = tmp_path / "o"; out.mkdir()
    r = subprocess.run([TOOL, "-b", base, "-c", cfg, "-r", rec, "-o", str(out), "-C", gcl, "--eval-ming=0.25", "--precision=f64"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert r.returncode == 0, r.stderr.decode()
    got = np.frombuffer(open(str(out / "utt0.lna"), "rb").read(), dtype=np.uint8)
    want = g["lna2"]
    assert bytes(got[:5]) == bytes(want[:5]) and got.size == want.size
    d = np.abs(got[5:].view(">u2").astype(int) - want[5:].view(">u2").astype(int))
    assert d.max() <= 1 and (d != 0).mean() <= 0.03, (d.max(), (d != 0).mean())
    # and it differs from the exact evaluation (the flag is not ignored)
    out2 = tmp_path / "o2"; out2.mkdir()
    r = subprocess.run([TOOL, "-b", base, "-c", cfg, "-r", rec, "-o", str(out2), "--precision=f64"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert r.returncode == 0
    exact = np.frombuffer(open(str(out2 / "utt0.lna"), "rb").read(), dtype=np.uint8)
    assert (exact != got).mean() > 0.2


def test_cpp_tool_speaker_config(engine, ref_spk, tmp_path):
    """akugpu_phone_probs -S x.spkc (aku/phone_probs.cc:93-94,192-197): per-speaker lin_transform parameters pushed
    through FeatureModule::set_parameters between GPU calls; LNA files against the literal phone_probs -S output
    (end to end from WAV: codes within +-1), the Python mirror writes the same bytes, model-level entries are refused."""
    from aaltoasr_b200 import SpeakerConfig
    g = ref_spk
    cfg = str(tmp_path / "spk.cfg"); open(cfg, "w").write(g["cfg"])
    base = str(tmp_path / "model"); formats.write_model(base, **g["model"])
    spkc = str(tmp_path / "x.spkc"); open(spkc, "w").write(g["spkc"])
    lines = []
    for i, spk in enumerate(g["speakers"]):
        a, b = g["cut_ranges"][i]
        w = str(tmp_path / ("spk%d.wav" % i))
        formats.write_wav(w, g["pcm"][a:b], 16000)
        lines.append("audio=%s lna=spk%d.lna speaker=%s" % (w, i, spk))
    rec = str(tmp_path / "recipe"); open(rec, "w").write("\n".join(lines) + "\n")
    out = tmp_path / "o"; out.mkdir()
    r = subprocess.run([TOOL, "-b", base, "-c", cfg, "-r", rec, "-o", str(out), "-S", spkc, "--precision=f64"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert r.returncode == 0, r.stderr.decode()
    engine.frontend_load_config(cfg)
    engine.model_read(base)
    sc = SpeakerConfig(engine)
    sc.read_speaker_file(spkc)
    for i, spk in enumerate(g["speakers"]):
        got = np.frombuffer(open(str(out / ("spk%d.lna" % i)), "rb").read(), dtype=np.uint8)
        want = g["lna2_%d" % i]
        assert got.size == want.size and bytes(got[:5]) == bytes(want[:5])
        d = np.abs(got[5:].view(">u2").astype(int) - want[5:].view(">u2").astype(int))
        assert d.max() <= 1 and (d != 0).mean() <= 0.03, (spk, d.max(), (d != 0).mean())
        sc.set_speaker(spk)                                  # the Python mirror drives the same library
        a, b = g["cut_ranges"][i]
        mine, _, _ = engine.phone_probs(g["pcm"][a:b], precision=F64, lnabytes=2)
        assert np.array_equal(mine.reshape(-1), got[5:])
    open(spkc, "w").write("speaker default\n{\n  model mllr\n  {\n  }\n}\n")
    r = subprocess.run([TOOL, "-b", base, "-c", cfg, "-r", rec, "-o", str(out), "-S", spkc],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert r.returncode != 0 and b"unknown model module requested: mllr" in r.stderr


def test_cpp_tool_model_cmllr(engine, ref_cmllr, tmp_path):
    """akugpu_phone_probs -S x.spkc with `model cmllr` entries (global model-level CMLLR, aku/ModelModules.cc): LNA files
    for three speakers against the literal phone_probs -S output (2- and 4-byte, normalised and -N: the factor
    |prod diag(A)| shows only without normalisation); the Python mirror writes the same bytes; regression-class unit modes
    are refused."""
    from aaltoasr_b200 import SpeakerConfig
    g = ref_cmllr
    cfg = str(tmp_path / "c.cfg"); open(cfg, "w").write(g["cfg"])
    base = str(tmp_path / "model"); formats.write_model(base, **g["model"])
    spkc = str(tmp_path / "x.spkc"); open(spkc, "w").write(g["spkc"])
    lines = []
    for i, spk in enumerate(g["speakers"]):
        a, b = g["cut_ranges"][i]
        w = str(tmp_path / ("cm%d.wav" % i))
        formats.write_wav(w, g["pcm"][a:b], 16000)
        lines.append("audio=%s lna=cm%d.lna speaker=%s" % (w, i, spk))
    rec = str(tmp_path / "recipe"); open(rec, "w").write("\n".join(lines) + "\n")
    engine.frontend_load_config(cfg)
    engine.model_read(base)
    sc = SpeakerConfig(engine)
    sc.read_speaker_file(spkc)
    for nb, tag, extra in ((2, "", []), (4, "", []), (2, "raw", ["-N"]), (4, "raw", ["-N"])):
        out = tmp_path / ("o%d%s" % (nb, tag)); out.mkdir()
        r = subprocess.run([TOOL, "-b", base, "-c", cfg, "-r", rec, "-o", str(out), "-S", spkc, "--precision=f64",
                            "--lnabytes=%d" % nb] + extra, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        assert r.returncode == 0, r.stderr.decode()
        for i, spk in enumerate(g["speakers"]):
            got = np.frombuffer(open(str(out / ("cm%d.lna" % i)), "rb").read(), dtype=np.uint8)
            want = g["lna%d%s_%d" % (nb, tag, i)]
            assert got.size == want.size and bytes(got[:5]) == bytes(want[:5])
            if nb == 2:      # end to end from WAV: codes within +-1 (the FFT rounds differently from KissFFT)
                d = np.abs(got[5:].view(">u2").astype(int) - want[5:].view(">u2").astype(int))
                assert d.max() <= 1 and (d != 0).mean() <= 0.03, (spk, tag, d.max(), (d != 0).mean())
            else:
                assert np.abs(got[5:].view("<f4") - want[5:].view("<f4")).max() <= 2e-3, (spk, tag)
            sc.set_speaker(spk)                                  # the Python mirror drives the same library
            a, b = g["cut_ranges"][i]
            mine, _, _ = engine.phone_probs(g["pcm"][a:b], precision=F64, lnabytes=nb, normalize=not extra)
            assert np.array_equal(mine.reshape(-1), got[5:])
    engine.model_set_cmllr(None)
    # alice and bob differ from the untransformed model, carol (default speaker: no `w`) does not
    plain = [engine.phone_probs(g["pcm"][a:b], precision=F64, lnabytes=2)[0].reshape(-1) for a, b in g["cut_ranges"]]
    assert not np.array_equal(plain[0], g["lna2_0"][5:]) and not np.array_equal(plain[1], g["lna2_1"][5:])
    d = np.abs(plain[2].view(">u2").astype(int) - g["lna2_2"][5:].view(">u2").astype(int))
    assert d.max() <= 1
    open(spkc, "w").write(g["spkc"].replace("unitmode UNIT_NO", "unitmode UNIT_MIX"))
    r = subprocess.run([TOOL, "-b", base, "-c", cfg, "-r", rec, "-o", str(tmp_path / "o2"), "-S", spkc],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert r.returncode != 0 and b"not enough elements for matrix w1" in r.stderr        # a regression-class entry needs its units


def test_cpp_pptoolbox_mirror(engine, ref_small, ref_clust, tmp_path):
    """akugpu::PPToolbox (C++, aku/PhoneProbsToolbox.hh:13-31 / aku/swig/PPToolbox.i:59-66): generate() from a file,
    generate_to_fd() from open descriptors (WAV and headerless raw PCM), set_clustering(); the LNA stream (5-byte
    header + 2-byte normalised codes) is byte-identical to the library's own phone_probs output and within +-1 code of
    the literal reference tool's file."""
    from test_abi import build_pptoolbox_driver
    exe = build_pptoolbox_driver(tmp_path)
    g = ref_small
    cfg, base, rec, cuts = write_case(tmp_path, g, n=1)
    wav = str(tmp_path / "utt0.wav")
    raw = str(tmp_path / "utt0.raw")
    g["pcm"].astype("<i2").tofile(raw)
    engine.frontend_load_config(cfg)
    engine.model_read(base)
    S = engine.num_states
    run = lambda *a: subprocess.run([exe] + list(a), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    for prec_name, prec in (("f32", F32), ("f64", F64)):
        want, _, _ = engine.phone_probs(g["pcm"], precision=prec, lnabytes=2)
        for mode, src in (("file", wav), ("fd", wav), ("rawfd", raw)):
            out = str(tmp_path / ("o_%s_%s.lna" % (prec_name, mode)))
            r = run(cfg, base, src, out, prec_name, mode)
            assert r.returncode == 0, r.stderr.decode()
            got = np.frombuffer(open(out, "rb").read(), dtype=np.uint8)
            assert bytes(got[:5]) == S.to_bytes(4, "big") + b"\x02" == bytes(g["lna2"][:5])
            assert np.array_equal(got[5:], want.reshape(-1)), (prec_name, mode)
        d = np.abs(got[5:].view(">u2").astype(int) - g["lna2"][5:].view(">u2").astype(int))
        assert d.max() <= 1 and (d != 0).mean() <= 0.03, (prec_name, d.max(), (d != 0).mean())
    # set_clustering: the approximation of phone_probs -C x.gcl --eval-ming=0.25 (fixture of the literal tool)
    gc = ref_clust
    gcl = str(tmp_path / "c.gcl")
    open(gcl, "w").write(gc["gcl"])
    out = str(tmp_path / "o_clust.lna")
    r = run(cfg, base, wav, out, "f32", "file", gcl, "0.0", "0.25")
    assert r.returncode == 0, r.stderr.decode()
    got = np.frombuffer(open(out, "rb").read(), dtype=np.uint8)
    d = np.abs(got[5:].view(">u2").astype(int) - gc["lna2"][5:].view(">u2").astype(int))
    assert got.size == gc["lna2"].size and d.max() <= 1 and (d != 0).mean() <= 0.03
    engine.read_clustering(gcl)
    engine.set_clustering_min_evals(0.0, 0.25)
    want, _, _ = engine.phone_probs(g["pcm"], precision=F32, lnabytes=2)
    engine.use_clustering(False)
    assert np.array_equal(got[5:], want.reshape(-1))
    # errors: a model of another dimension, a missing input
    bad = str(tmp_path / "bad")
    m = g["model"]
    formats.write_model(bad, m["mix_offsets"], m["mix_gauss"], m["mix_weight"], m["means"][:, :13], m["covs"][:, :13])
    r = run(cfg, bad, wav, out)
    assert r.returncode == 1 and b"Gaussian dimension is 13 but feature dimension is 39." in r.stderr
    r = run(cfg, base, str(tmp_path / "missing.wav"), out)
    assert r.returncode == 1 and b"could not open file" in r.stderr


FEACAT = os.path.join(ROOT, "aaltoasr_b200", "akugpu_feacat")


def _ascii_rows(b):
    return np.array([[float(x) for x in l.split()] for l in b.decode().splitlines() if l.strip()])


def test_feacat_tool_aku_scripts(engine, aku_tests, ref_vtln, tmp_path):
    """akugpu_feacat run the way the reference's own regression scripts run feacat
    (aku/tests/mfcc_p_dd.script, mfcc_cms_norm.script, pre_test.script), against their .ref outputs;
    the bar is theirs: equal at the 4 printed decimals up to the reference's float-buffer rounding (5.1e-3)."""
    wav = str(tmp_path / "short.wav")
    formats.write_wav(wav, aku_tests["short_wav"], int(aku_tests["sample_rate"]))
    cfgs = {}
    for k in ("mfcc_p_dd", "mfcc_cms_norm", "pre"):
        cfgs[k] = str(tmp_path / (k + ".feaconf"))
        open(cfgs[k], "w").write(aku_tests[k + "_cfg"])
    run = lambda *a, **kw: subprocess.run([FEACAT] + list(a), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300, **kw)
    # mfcc_p_dd.script: audio on standard input, negative start, --write-config and a second run on the written copy
    tmpcfg = str(tmp_path / "mfcc_p_dd.feaconf.tmp")
    audio = open(wav, "rb").read()
    r1 = run("--start-frame", "-10", "--end-frame", "80", "--write-config", tmpcfg, "-c", cfgs["mfcc_p_dd"], "-", input=audio)
    assert r1.returncode == 0, r1.stderr.decode()
    r2 = run("--start-frame", "-10", "--end-frame", "80", "-c", tmpcfg, "-", input=audio)
    assert r2.returncode == 0, r2.stderr.decode()
    got = _ascii_rows(r1.stdout + r2.stdout)
    gold = aku_tests["mfcc_p_dd_ref"]
    assert got.shape == gold.shape and np.abs(got - gold).max() <= 0.0051
    assert r1.stdout == r2.stdout
    # every row is "%8.4f " per component
    line = r1.stdout.decode().splitlines()[0]
    assert len(line) == 9 * gold.shape[1] and line.endswith(" ")
    # mfcc_cms_norm.script: short options, frames past both ends
    r = run("-c", cfgs["mfcc_cms_norm"], "-s", "-15", "-e", "90", "-", input=audio)
    assert r.returncode == 0, r.stderr.decode()
    got = _ascii_rows(r.stdout)
    assert got.shape == aku_tests["mfcc_cms_norm_ref"].shape and np.abs(got - aku_tests["mfcc_cms_norm_ref"]).max() <= 0.0051
    # pre_test.script: raw output with header, read back through a `pre` configuration
    r = run("--start-frame", "10", "--end-frame", "60", "-c", cfgs["mfcc_p_dd"], "-H", "--raw-output", wav)
    assert r.returncode == 0, r.stderr.decode()
    dim = int(np.frombuffer(r.stdout[:4], "<i4")[0])
    assert dim == 39 and len(r.stdout) == 4 + 51 * 39 * 4
    pre = str(tmp_path / "pre_test.tmp")
    open(pre, "wb").write(r.stdout)
    r = run("-c", cfgs["pre"], pre)
    assert r.returncode == 0, r.stderr.decode()
    got = _ascii_rows(r.stdout)
    assert got.shape == aku_tests["pre_test_ref"].shape and np.abs(got - aku_tests["pre_test_ref"]).max() <= 0.0051
    # whole file (open end), a descending range, and the library's own matrix behind the same rows
    engine.frontend_load_config_text(aku_tests["mfcc_p_dd_cfg"])
    full, _ = engine.features(aku_tests["short_wav"], dtype=np.float64)
    r = run("-c", cfgs["mfcc_p_dd"], "--raw-output", wav)
    assert r.returncode == 0, r.stderr.decode()
    raw = np.frombuffer(r.stdout, "<f4").reshape(-1, 39)
    assert raw.shape == full.shape and np.array_equal(raw, full.astype(np.float32))
    r = run("-c", cfgs["mfcc_p_dd"], "--raw-output", "-s", "5", "-e", "-3", wav)
    assert r.returncode == 0, r.stderr.decode()
    desc = np.frombuffer(r.stdout, "<f4").reshape(-1, 39)
    want = engine.features_range(aku_tests["short_wav"], -3, 6, dtype=np.float64)[::-1]
    assert np.array_equal(desc, want.astype(np.float32))
    # -S / -d: a speaker's feature-module parameters (the vtln warp) reach the front-end; goldens = the reference's
    cfg = str(tmp_path / "vtln.cfg")
    open(cfg, "w").write(ref_vtln["cfg_blin"])
    spkc = str(tmp_path / "v.spkc")
    open(spkc, "w").write(ref_vtln["spkc_blin"])
    w2 = str(tmp_path / "v.wav")
    formats.write_wav(w2, ref_vtln["pcm"], 16000)
    for spk in ("s1", "other", "s2"):
        r = run("-c", cfg, "-S", spkc, "-d", spk, "--raw-output", w2)
        assert r.returncode == 0, r.stderr.decode()
        gold = ref_vtln["feats_blin_" + spk]
        got = np.frombuffer(r.stdout, "<f4").reshape(-1, gold.shape[1])
        assert got.shape == gold.shape and np.abs(got - gold).max() <= 2e-5 + 1e-6 * np.abs(gold).max()
    # errors: -G refused, bad config reported like the reference ("exception: ...", non-zero exit)
    assert run("-c", cfgs["mfcc_p_dd"], "-G", "0.1", wav).returncode != 0
    r = run("-c", str(tmp_path / "missing.cfg"), wav)
    assert r.returncode != 0 and b"exception:" in r.stderr


def test_cpp_tool_regression_class_cmllr(ref_cmllr_units, tmp_path):
    """akugpu_phone_probs -S with `model cmllr` entries in the regression-class unit modes (UNIT_PHONE / UNIT_MIX /
    UNIT_GAUSSIAN): the C++ SpeakerConfig hands units + matrices to akugpu_model_set_cmllr_units; files within +-1 code of
    the literal phone_probs -S output (features from the WAV differ in the last bits)."""
    g = ref_cmllr_units
    base = str(tmp_path / "m")
    formats.write_model(base, **g["model"])
    open(base + ".ph", "w").write(g["ph"])
    cfg = str(tmp_path / "c.cfg"); open(cfg, "w").write(g["cfg"])
    spkc = str(tmp_path / "x.spkc"); open(spkc, "w").write(g["spkc"])
    wav = str(tmp_path / "a.wav"); formats.write_wav(wav, g["pcm"], 16000)
    rec = str(tmp_path / "r")
    open(rec, "w").write("".join("audio=%s lna=%s.lna speaker=%s\n" % (wav, spk, spk) for spk in ("phone", "mix", "gauss")))
    for prec in ("f64", "f32"):
        out = tmp_path / ("o_" + prec); out.mkdir()
        r = subprocess.run([TOOL, "-b", base, "-c", cfg, "-r", rec, "-o", str(out), "-S", spkc, "--precision=" + prec],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        assert r.returncode == 0, r.stderr.decode()
        for spk in ("phone", "mix", "gauss"):
            got = np.frombuffer(open(str(out / (spk + ".lna")), "rb").read(), dtype=np.uint8)
            want = g["lna2_" + spk]
            assert got.size == want.size and bytes(got[:5]) == bytes(want[:5])
            d = np.abs(got[5:].view(">u2").astype(int) - want[5:].view(">u2").astype(int))
            assert d.max() <= 1 and (d != 0).mean() <= 0.03, (prec, spk, d.max(), (d != 0).mean())
