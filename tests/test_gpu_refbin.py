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
