"""GPU tests of the reference's LITERAL tools running on the library: oracle/_ref/ref_feacat_gpu and ref_phone_probs_gpu
are aku/feacat.cc / aku/phone_probs.cc linked with integration/GpuFrontendModule.hh registered in FeatureGenerator and
integration/GpuHmmSetHook.hh hooked into HmmSet (three added lines, oracle/build_ref.sh), against the real libakugpu.so.

The same chain also runs on the CPU against the fake ABI (tests/test_abi.py)."""
import os
import subprocess

import numpy as np
import pytest

from aaltoasr_b200 import formats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(REF, "ref_feacat_gpu")), reason="oracle/_ref not built")]


def gpu_config(tmp_path, inner_text):
    inner = str(tmp_path / "inner.feaconf")
    open(inner, "w").write(inner_text)
    cfg = str(tmp_path / "gpu.feaconf")
    open(cfg, "w").write("module\n{\n  name gpu\n  type gpu_frontend\n  config %s\n}\n" % inner)
    return cfg


def test_reference_feacat_on_the_gpu_module(aku_tests, tmp_path):
    """aku/tests/mfcc_p_dd.script with the reference's own feacat, the GPU chain as its base module."""
    wav = str(tmp_path / "short.wav")
    formats.write_wav(wav, aku_tests["short_wav"], int(aku_tests["sample_rate"]))
    cfg = gpu_config(tmp_path, aku_tests["mfcc_p_dd_cfg"])
    r = subprocess.run([os.path.join(REF, "ref_feacat_gpu"), "-c", cfg, "--start-frame", "-10", "--end-frame", "80", "-"],
                       input=open(wav, "rb").read(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert r.returncode == 0, r.stderr.decode()
    got = np.array([[float(x) for x in ln.split()] for ln in r.stdout.decode().splitlines()])
    gold = aku_tests["mfcc_p_dd_ref"][:91]
    assert got.shape == gold.shape and np.abs(got - gold).max() <= 0.0051


def test_reference_phone_probs_on_module_and_hook(ref_small, tmp_path):
    """The reference's own phone_probs loop, features from the GPU module, likelihoods through the HmmSet hook (double
    path, whole utterance in one call): LNA files within +-1 code of the CPU tool's, and identical to the CPU tool's
    when the hook is off and the configuration is the plain one."""
    g = ref_small
    cfg = gpu_config(tmp_path, g["cfg"])
    base = str(tmp_path / "model")
    formats.write_model(base, **g["model"])
    wav = str(tmp_path / "a.wav")
    formats.write_wav(wav, g["pcm"], 16000)
    rec = str(tmp_path / "recipe")
    open(rec, "w").write("audio=%s lna=a.lna\n" % wav)
    env = dict(os.environ, AKUGPU_HOOK="1")
    for nb, key in ((2, "lna2"), (4, "lna4")):
        out = tmp_path / ("o%d" % nb)
        out.mkdir()
        r = subprocess.run([os.path.join(REF, "ref_phone_probs_gpu"), "-b", base, "-c", cfg, "-r", rec, "-o", str(out),
                            "--lnabytes=%d" % nb], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=600)
        assert r.returncode == 0, r.stderr.decode()
        got = np.frombuffer(open(str(out / "a.lna"), "rb").read(), dtype=np.uint8)
        want = g[key]
        assert got.size == want.size and bytes(got[:5]) == bytes(want[:5])
        if nb == 2:
            d = np.abs(got[5:].view(">u2").astype(int) - want[5:].view(">u2").astype(int))
            assert d.max() <= 1 and (d != 0).mean() <= 0.03, (d.max(), (d != 0).mean())
        else:
            a, b = got[5:].view("<f4"), want[5:].view("<f4")
            assert (np.abs(a - b) / np.abs(b)).max() <= 1e-4
    # hook off, plain configuration: the binary is the CPU tool, byte for byte
    plain = str(tmp_path / "plain.feaconf")
    open(plain, "w").write(g["cfg"])
    out = tmp_path / "cpu"
    out.mkdir()
    r = subprocess.run([os.path.join(REF, "ref_phone_probs_gpu"), "-b", base, "-c", plain, "-r", rec, "-o", str(out)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0 and open(str(out / "a.lna"), "rb").read() == bytes(g["lna2"])


def test_reference_aligner_on_the_hook(ref_small, tmp_path):
    """The reference's Viterbi aligner (aku/align.cc -> Viterbi.cc:249,369: reset_cache() + state_likelihood(s, f) for the
    ACTIVE states only) on the HmmSet hook: the first miss of a frame scores every state on the GPU, the frame's other
    requests are cache hits.  Same segmentation as the CPU tool, with the reference's own features (one library call
    per frame) and with the GPU module as the base module (the whole utterance scored in one call)."""
    g = ref_small
    base = str(tmp_path / "model")
    formats.write_model(base, **g["model"])
    S = len(g["model"]["mix_offsets"]) - 1
    P = S // 3
    with open(base + ".ph", "w") as f:          # three-state left-to-right phones over the fixture's 24 states
        f.write("PHONE\n%d\n" % P)
        for p in range(P):
            f.write("%d 5 p%d\n-1 -2 %d %d %d\n0 1 2 1\n1 0\n2 2 2 0.8 3 0.2\n3 2 3 0.8 4 0.2\n4 2 4 0.8 1 0.2\n" % (p + 1, p, 3 * p, 3 * p + 1, 3 * p + 2))
    wav = str(tmp_path / "a.wav")
    formats.write_wav(wav, g["pcm"], 16000)
    plain = str(tmp_path / "plain.feaconf")
    open(plain, "w").write(g["cfg"])
    gpucfg = gpu_config(tmp_path, g["cfg"])
    phn = str(tmp_path / "a.phn")
    open(phn, "w").write("\n".join("p%d" % p for p in [0, 3, 1, 5, 2, 7, 4, 6, 0, 2]) + "\n")

    def align(exe, cfg, tag, hook):
        out = str(tmp_path / (tag + ".phn"))
        rec = str(tmp_path / (tag + ".recipe"))
        open(rec, "w").write("audio=%s transcript=%s alignment=%s\n" % (wav, phn, out))
        env = dict(os.environ, AKUGPU_HOOK="1" if hook else "")
        r = subprocess.run([os.path.join(REF, exe), "-b", base, "-c", cfg, "-r", rec, "-i", "2"], stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, env=env, timeout=600)
        assert r.returncode == 0, r.stderr.decode()
        ll = [float(ln.split(":")[1]) for ln in r.stderr.decode().splitlines() if ln.startswith("File log likelihood")]
        return open(out).read(), ll[0]

    cpu, ll_cpu = align("ref_align", plain, "cpu", False)
    assert len(cpu.splitlines()) == 30 and cpu.splitlines()[0].split()[2] == "p0.0"
    hooked, ll_h = align("ref_align_gpu", plain, "hook", True)              # reference features, GPU likelihoods per frame
    assert hooked == cpu and abs(ll_h - ll_cpu) <= 1e-6 * abs(ll_cpu)
    full, ll_f = align("ref_align_gpu", gpucfg, "gpu", True)                # GPU features + whole utterance scored at once
    assert abs(ll_f - ll_cpu) <= 1e-4 * abs(ll_cpu)
    a = [ln.split() for ln in full.splitlines()]
    b = [ln.split() for ln in cpu.splitlines()]
    assert [x[2] for x in a] == [x[2] for x in b]
    assert max(abs(int(x[0]) - int(y[0])) for x, y in zip(a, b)) <= 128      # boundaries within a frame (features differ by 1e-5)
    off, ll_o = align("ref_align_gpu", plain, "off", False)                 # hook off: the binary is the CPU tool
    assert off == cpu and ll_o == ll_cpu
