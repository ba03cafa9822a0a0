"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/akugpu.h
declares, and fails loudly (no fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

import aaltoasr_b200
from aaltoasr_b200 import _lib, formats, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "akugpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(akugpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = aaltoasr_b200.load_library()
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libakugpu.so does not export " + n
    assert sorted(_lib.SYMBOLS) == names, "ctypes table and header disagree"


def test_no_torch_types_in_abi():
    text = open(os.path.join(ROOT, "include", "akugpu.h")).read()
    assert 'extern "C"' in text
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)      # declarations only, comments stripped
    assert "torch" not in code and "std::" not in code and "at::" not in code


def test_lna_header_needs_no_gpu():
    lib = aaltoasr_b200.load_library()
    buf = (ctypes.c_uint8 * 5)()
    assert lib.akugpu_lna_header(5000, 2, buf) == 0
    assert bytes(buf) == b"\x00\x00\x13\x88\x02" == formats.lna_header(5000, 2)


def test_create_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(aaltoasr_b200.AkuGpuError, match="no CPU fallback"):
        aaltoasr_b200.AkuGpu(0)


@pytest.mark.parametrize("tool", ["akugpu_phone_probs", "akugpu_feacat"])
def test_host_tools_build_and_fail_loudly_without_cuda(tool, tmp_path):
    """The C++ tools over the ABI (aku/phone_probs.cc, aku/feacat.cc re-hosted) are built by build(), print their
    usage without a device, and exit non-zero with the library's message when there is none."""
    import subprocess
    import torch
    exe = os.path.join(ROOT, "aaltoasr_b200", tool)
    assert os.access(exe, os.X_OK), exe + " missing: run __graft_entry__.build()"
    r = subprocess.run([exe, "--help"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    assert r.returncode == 0 and b"usage: " + tool.encode() in r.stdout and b"--config" in r.stdout
    if torch.cuda.is_available():
        return
    args = ["-c", str(tmp_path / "x.cfg")] + (["-r", str(tmp_path / "x.recipe"), "-b", "m"] if tool.endswith("probs") else ["x.wav"])
    r = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    assert r.returncode != 0 and b"no CPU fallback" in r.stderr


def build_pptoolbox_driver(tmp_path):
    """g++ build of tests/cpp/pptoolbox_main.cc (akugpu::PPToolbox, header-only) against the in-tree libakugpu.so."""
    import subprocess
    exe = str(tmp_path / "pptoolbox_main")
    libdir = os.path.join(ROOT, "aaltoasr_b200")
    subprocess.run(["g++", "-O1", "-std=c++11", "-Wall", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "cpp", "pptoolbox_main.cc"),
                    "-L" + libdir, "-lakugpu", "-Wl,-rpath," + libdir], check=True, timeout=300)
    return exe


def test_cpp_pptoolbox_builds_and_fails_loudly_without_cuda(tmp_path):
    """akugpu::PPToolbox (C++ mirror of aku::PPToolbox, aku/PhoneProbsToolbox.hh) compiles warning-free against the C ABI
    and reports the missing device instead of falling back."""
    import subprocess
    import torch
    exe = build_pptoolbox_driver(tmp_path)
    if torch.cuda.is_available():
        return
    r = subprocess.run([exe, "x.cfg", "model", "in.wav", str(tmp_path / "out.lna")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    assert r.returncode == 1 and b"no CPU fallback" in r.stderr and not os.path.exists(str(tmp_path / "out.lna"))


def test_cpp_speaker_config_matches_python_mirror(tmp_path):
    """akugpu::SpeakerConfig (the C++ adapter, csrc/host/akugpu.hh) compiled against stubs of the C-ABI calls it makes:
    for the speaker files of the goldens it issues the same frontend_set_parameters / model_set_cmllr calls, with the
    same values, as the Python mirror -- feature modules, `model cmllr` (floats through str2float), default speaker,
    sticky parameters, and the reference's error messages.  Also akugpu::read_audio (WAV / raw PCM16, as AudioReader)."""
    import subprocess
    from conftest import load_golden
    from aaltoasr_b200 import parse_speaker_file
    from aaltoasr_b200.hostapi import parse_cmllr_parameters
    exe = str(tmp_path / "spk_harness")
    subprocess.run(["g++", "-O1", "-std=c++11", "-Wall", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "cpp", "spk_harness.cc")],
                   check=True, timeout=300)

    def calls(spkc_text, dim, speakers):
        path = str(tmp_path / "x.spkc")
        open(path, "w").write(spkc_text)
        r = subprocess.run([exe, path, str(dim)] + list(speakers), stdout=subprocess.PIPE, timeout=60)
        return r.returncode, r.stdout.decode()

    # akugpu::read_audio: RIFF/WAVE PCM16 (with an odd-sized extra chunk before `data`), headerless raw, errors
    pcm = synth.synth_audio(11, 4001)
    w = str(tmp_path / "a.wav")
    formats.write_wav(w, pcm, 8000)
    blob = open(w, "rb").read()
    i = blob.index(b"data")
    extra = blob[:i] + b"LIST" + (3).to_bytes(4, "little") + b"abc\x00" + blob[i:]
    extra = extra[:4] + (len(extra) - 8).to_bytes(4, "little") + extra[8:]
    open(str(tmp_path / "b.wav"), "wb").write(extra)
    pcm.astype("<i2").tofile(str(tmp_path / "a.raw"))
    want_sum = int((pcm.astype(np.int64) * (np.arange(pcm.size) % 97 + 1)).sum())
    # WAVE_FORMAT_EXTENSIBLE with a PCM SubFormat (what many recorders write; libsndfile reads it as PCM16)
    fmt_at = blob.index(b"fmt ")
    body = blob[fmt_at + 8:fmt_at + 24]
    ext = (b"\xfe\xff" + body[2:] + (22).to_bytes(2, "little") + (16).to_bytes(2, "little") + (4).to_bytes(4, "little") +
           b"\x01\x00\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71")
    extw = blob[:fmt_at] + b"fmt " + (40).to_bytes(4, "little") + ext + blob[fmt_at + 24:]
    extw = extw[:4] + (len(extw) - 8).to_bytes(4, "little") + extw[8:]
    open(str(tmp_path / "e.wav"), "wb").write(extw)
    got_pcm, got_sr = formats.read_wav(str(tmp_path / "e.wav"))
    assert got_sr == 8000 and np.array_equal(got_pcm, pcm)
    for path, cfg_rate, raw, rate in (("a.wav", 16000, 0, 8000), ("b.wav", 16000, 0, 8000), ("e.wav", 16000, 0, 8000), ("a.raw", 16000, 1, 16000)):
        r = subprocess.run([exe, "audio", str(tmp_path / path), str(cfg_rate), str(raw)], stdout=subprocess.PIPE, timeout=60)
        assert r.returncode == 0 and r.stdout.decode().split() == [str(pcm.size), str(rate), str(want_sum)], (path, r.stdout)
    r = subprocess.run([exe, "audio", str(tmp_path / "nope.wav"), "16000", "0"], stdout=subprocess.PIPE, timeout=60)
    assert r.returncode == 1 and b"could not open file" in r.stdout
    stereo = bytearray(blob)
    stereo[22] = 2
    open(str(tmp_path / "s.wav"), "wb").write(bytes(stereo))
    r = subprocess.run([exe, "audio", str(tmp_path / "s.wav"), "16000", "0"], stdout=subprocess.PIPE, timeout=60)
    assert r.returncode == 1 and b"multiple channels not supported" in r.stdout

    # akugpu::FeatureGenerator: open(FILE*, dont_fclose) / open_fd / open(path) on the same file
    w16 = str(tmp_path / "a16.wav")
    formats.write_wav(w16, pcm, 16000)
    r = subprocess.run([exe, "fgopen", w], stdout=subprocess.PIPE, timeout=60)           # the 8 kHz file: refused like the reference does
    assert r.returncode == 1 and b"Audio file sample rate (8000 Hz) and model configuration (16000 Hz) don't agree." in r.stdout
    r = subprocess.run([exe, "fgopen", w16], stdout=subprocess.PIPE, timeout=60)
    assert r.returncode == 0 and r.stdout.decode().splitlines() == ["FILE* 31 frames, f(3,1)=3.25", "fd 31 frames, eof(last)=0 eof(last+1)=1",
                                                                    "path 31 frames"], r.stdout
    r = subprocess.run([exe, "fgopen", str(tmp_path / "a.raw")], stdout=subprocess.PIPE, timeout=60)      # headerless: read as raw PCM16
    assert r.returncode == 0 and r.stdout.decode().splitlines()[0].startswith("FILE* 31 frames")

    # akugpu::HmmSet: the reference's per-vector signatures over one GPU call per reset_cache() (stubbed scorer, 3 states)
    for prec in ("f64", "f32"):
        r = subprocess.run([exe, "hmm", prec], stdout=subprocess.PIPE, timeout=60)
        assert r.returncode == 0
        assert r.stdout.decode().splitlines() == ["4 12 calls=1", "8 calls=1", "24 12 calls=2", "1e-50 calls=3", "36 102 calls=4"], (prec, r.stdout)

    # akugpu::StreamSession: open on construction, rows of the context returned without a copy, closed on destruction
    r = subprocess.run([exe, "session", "x"], stdout=subprocess.PIPE, timeout=60)
    assert r.returncode == 0
    want = [np.float32(np.log((1 + 1 + 2) * (s + 1) / 1000.0)) for s in range(3)] + [np.float32(np.log((1 + 5 + 6) / 1000.0))]
    assert r.stdout.decode().splitlines() == ["3 states, open=1, rows %.6g %.6g %.6g | %.6g" % tuple(want),
                                              "3 values, %.6g, calls=2" % np.float32(np.log(12 * 3 / 1000.0)), "open=0"], r.stdout

    # model-level CMLLR fixture
    g = load_golden("ref_cmllr")
    spkc = str(g["spkc"])
    rc, out = calls(spkc, 39, ["alice", "bob", "carol", "alice"])
    assert rc == 0, out
    conf = parse_speaker_file(spkc)["speaker"]
    blocks = out.split("speaker ")[1:]
    assert [b.split("\n", 1)[0] for b in blocks] == ["alice", "bob", "carol", "alice"]
    for b in blocks:
        spk, rest = b.split("\n", 1)
        W = parse_cmllr_parameters(conf.get(spk, conf["default"])["model cmllr"], 39)
        if W is None:
            assert rest.strip() == "cmllr none"
        else:
            vals = np.array([float(t) for t in rest.split()[1:]])
            assert rest.startswith("cmllr ") and np.array_equal(vals, W.reshape(-1))
    # feature-module fixture: the parameter text reaches akugpu_frontend_set_parameters unchanged
    g = load_golden("ref_spk")
    spkc = str(g["spkc"])
    rc, out = calls(spkc, 39, ["alice", "carol"])
    assert rc == 0, out
    conf = parse_speaker_file(spkc)["speaker"]
    assert out == "speaker alice\nfeature cmllr\n%s.\nspeaker carol\nfeature cmllr\n%s.\n" % (conf["alice"]["cmllr"], conf["default"]["cmllr"])
    # errors, worded like the reference
    head = "speaker a\n{\n  model cmllr\n  {\n"
    for body, msg in (("    unitmode UNIT_PHONE\n    w1 1 0 0 1 0 0\n", "not enough elements for matrix w1"),      # no unit in front of the matrix
                      ("    w1 1 2 3\n", "not enough elements for matrix w1"),
                      ("    w1 0 1 zz 0 0 1\n", "invalid value: zz"),
                      ("    w1 p 0 1 0 0 0 1\n    w2 q 0 1 0 0 0 1\n", "only contain one transform")):
        rc, out = calls(head + body + "  }\n}\n", 2, ["a"])
        assert rc == 1 and msg in out, (body, out)
    rc, out = calls(head + "    w1 0.5 1 0 -0.5 0 1\n  }\n}\n", 2, ["a"])
    assert rc == 0 and out == "speaker a\ncmllr 0.5 1 0 -0.5 0 1\n"
    # regression classes: units in front of each matrix, handed to the library in the reference's std::map order
    rc, out = calls(head + "    unitmode UNIT_PHONE\n    w1 u a 0.5 1 0 -0.5 0 1\n    w2 b 0.25 1 0 0 0 2\n  }\n}\n", 2, ["a"])
    assert rc == 0 and out == "speaker a\ncmllr_units UNIT_PHONE [b] 0.25 2 [u a] 0.5 1\n", out
    rc, out = calls("speaker a\n{\n  model mllr\n  {\n  }\n}\n", 2, ["a"])
    assert rc == 1 and "SpeakerConfig: error on line 3: unknown model module requested: mllr" in out
    rc, out = calls("utterance u\n{\n  model cmllr\n  {\n  }\n}\n", 2, [])
    assert rc == 1 and "utterance-level" in out
    rc, out = calls("speaker a\n{\n}\n", 2, ["b"])
    assert rc == 1 and "Unknown speaker b, and default speaker settings are missing." in out


@pytest.mark.skipif(not (os.path.isdir("/root/reference/aku") and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libaku_ref.a"))),
                    reason="needs the reference's headers and oracle/_ref/libaku_ref.a")
def test_reference_side_plugin_compiles_and_serves_reference_modules(tmp_path):
    """integration/GpuFrontendModule.hh -- the aku::BaseFeaModule subclass a maintainer adds to the reference -- builds
    against the reference's own headers and library, and behaves as FeatureGenerator expects of a base module: dim /
    rates / config round trip, eof and last_frame, at() forward, past both ends and backward, set_file on a stream, a second
    file after reset, speaker parameters addressed to a module of the GPU chain, and finally the reference's literal
    feacat on a FeatureGenerator with the one-line registration -- with the reference's own DeltaModule computing from it (fake ABI: feature(f, d) = clamp(f) + 0.25 d)."""
    import subprocess
    R = "/root/reference"
    exe = str(tmp_path / "plugin_harness")
    subprocess.run(["g++", "-O1", "-std=gnu++11", "-DKISS_FFT", "-DDLLIMPORT=", "-fpermissive", "-w",
                    "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + R + "/aku", "-I" + R + "/vendor/kiss_fft",
                    "-I" + os.path.join(ROOT, "integration"), "-I" + os.path.join(ROOT, "aaltoasr_b200", "csrc", "host"),
                    "-o", exe, os.path.join(ROOT, "tests", "cpp", "plugin_harness.cc"), os.path.join(ROOT, "tests", "cpp", "stub_akugpu.cc"),
                    os.path.join(ROOT, "oracle", "_ref", "libaku_ref.a"), "-lm"], check=True, timeout=600)
    cfg = str(tmp_path / "f.cfg")
    open(cfg, "w").write("x")
    formats.write_wav(str(tmp_path / "a.wav"), np.zeros(1280, np.int16), 16000)      # 10 frames
    formats.write_wav(str(tmp_path / "b.wav"), np.zeros(2560, np.int16), 16000)      # 20 frames
    r = subprocess.run([exe, cfg, str(tmp_path / "a.wav"), str(tmp_path / "b.wav"), "-"], stdout=subprocess.PIPE,
                       input=open(str(tmp_path / "a.wav"), "rb").read(), timeout=60)
    assert r.returncode == 0, r.stdout
    out = r.stdout.decode().splitlines()
    assert out[:2] == ["type gpu_frontend dim 3 rate 16000 fr 125", "config roundtrip"]

    def block(n):
        last = n - 1
        rows = ["file last_frame %d eof(last) 0 eof(last+1) 1" % last]
        for f in (0, 1, 5, last, last + 3, -2, 3):
            c = min(max(f, 0), last)
            rows.append("base %d: %g %g %g" % (f, c, c + 0.25, c + 0.5))
        for f, d in ((-1, 0.2), (0, 0.5), (1, 0.8), (2, 1)):      # DeltaModule width 2 over clamp(f): sum k (x[t+k] - x[t-k]) / 10
            rows.append("delta %d: %g %g %g" % (f, d, d, d))
        return rows

    extra = ["shifted 4: 6 6.25 6.5", "refused: GpuFrontendModule: parameter keys are <module>.<parameter>: value 1"]
    assert out[2:] == block(10) + extra + block(20) + block(10)
    bad = str(tmp_path / "c.wav")
    formats.write_wav(bad, np.zeros(1280, np.int16), 8000)
    r = subprocess.run([exe, cfg, bad], stdout=subprocess.PIPE, timeout=60)
    assert r.returncode == 1 and b"Audio file sample rate (8000 Hz) and model configuration (16000 Hz) don't agree." in r.stdout

    # The one-line registration, applied to a scratch copy of aku/FeatureGenerator.cc (never stored in the repository),
    # and the reference's LITERAL feacat tool on top: a reference configuration whose base module is the GPU chain and
    # whose delta / merge modules are the reference's own.
    src = open(R + "/aku/FeatureGenerator.cc").read()
    marker = "    else\n      throw std::string(\"Unknown module type '\")"
    assert src.count(marker) == 1 and src.count('#include "FeatureModules.hh"\n') == 1
    src = src.replace('#include "FeatureModules.hh"\n', '#include "FeatureModules.hh"\n#include "GpuFrontendModule.hh"\n')
    src = src.replace(marker, "    else if (type == GpuFrontendModule::type_str())\n      module = new GpuFrontendModule();\n" + marker)
    patched = str(tmp_path / "FeatureGenerator_registered.cc")
    open(patched, "w").write(src)
    flags = ["-O1", "-std=gnu++11", "-DKISS_FFT", "-DDLLIMPORT=", "-fpermissive", "-w",
             "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + R + "/aku", "-I" + R + "/vendor/kiss_fft",
             "-I" + os.path.join(ROOT, "integration"), "-I" + os.path.join(ROOT, "aaltoasr_b200", "csrc", "host")]
    feacat = str(tmp_path / "ref_feacat_with_plugin")
    subprocess.run(["g++"] + flags + ["-o", feacat, R + "/aku/feacat.cc", patched, os.path.join(ROOT, "tests", "cpp", "stub_akugpu.cc"),
                                      os.path.join(ROOT, "oracle", "_ref", "libaku_ref.a"), "-lm"], check=True, timeout=600)
    ref_cfg = str(tmp_path / "ref.feaconf")
    open(ref_cfg, "w").write("module\n{\n  name gpu\n  type gpu_frontend\n  config %s\n}\n"
                             "module\n{\n  name d\n  type delta\n  sources gpu\n  width 2\n}\n"
                             "module\n{\n  name m\n  type merge\n  sources gpu d\n}\n" % cfg)
    r = subprocess.run([feacat, "-c", ref_cfg, "-s", "-1", "-e", "3", str(tmp_path / "a.wav")], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, timeout=60)
    assert r.returncode == 0, r.stderr
    got = np.array([[float(x) for x in ln.split()] for ln in r.stdout.decode().splitlines()])
    want = np.array([[c, c + 0.25, c + 0.5, d, d, d] for c, d in ((0, 0.2), (0, 0.5), (1, 0.8), (2, 1.0), (3, 1.0))])
    assert got.shape == want.shape and np.abs(got - want).max() < 1e-4
    r = subprocess.run([feacat, "-c", ref_cfg, str(tmp_path / "b.wav")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    assert r.returncode == 0 and len(r.stdout.decode().splitlines()) == 20          # to the end of the file: eof from the plugin


@pytest.mark.skipif(not (os.path.isdir("/root/reference/aku") and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libaku_ref.a"))),
                    reason="needs the reference's sources and oracle/_ref/libaku_ref.a")
def test_reference_phone_probs_runs_unmodified_on_plugin_and_hook(tmp_path):
    """The reference's LITERAL phone_probs main, linked with scratch copies of aku/FeatureGenerator.cc (+ the one-line
    registration of integration/GpuFrontendModule.hh) and aku/HmmSet.cc (+ the three hook lines of
    integration/GpuHmmSetHook.hh): with AKUGPU_HOOK=1 every frame's features and state likelihoods come from the C ABI
    (here its fake: closed forms), the reference's own loop normalises and quantises them, and the LNA file equals the
    oracle's encoding of those likelihoods; without the variable the same binary is the CPU tool."""
    import subprocess
    import sys
    sys.path.insert(0, ROOT)
    from oracle import oracle_np
    R = "/root/reference"
    fg = open(R + "/aku/FeatureGenerator.cc").read()
    marker = "    else\n      throw std::string(\"Unknown module type '\")"
    fg = fg.replace('#include "FeatureModules.hh"\n', '#include "FeatureModules.hh"\n#include "GpuFrontendModule.hh"\n')
    fg = fg.replace(marker, "    else if (type == GpuFrontendModule::type_str())\n      module = new GpuFrontendModule();\n" + marker)
    hs = open(R + "/aku/HmmSet.cc").read()
    a = '  read_gk(base + ".gk");\n}\n'
    b = "  // Precompute base distribution likelihoods\n  m_pool.precompute_likelihoods(*f.get_vector());\n"
    c = "    return m_pdf_likelihoods[p];\n\n  m_pdf_likelihoods[p] = m_emission_pdfs[p]->compute_likelihood(*feature.get_vector());\n"
    assert hs.count(a) == 1 and hs.count(b) == 1 and hs.count(c) == 1
    hs = hs.replace('#include "HmmSet.hh"\n', '#include "HmmSet.hh"\n#include "GpuHmmSetHook.hh"\n')
    hs = hs.replace(c, "    return m_pdf_likelihoods[p];\n  if (akugpu_hook::score(this, *feature.get_vector(), m_pdf_likelihoods, m_valid_pdf_likelihoods))\n"
                       "    return m_pdf_likelihoods[p];\n\n  m_pdf_likelihoods[p] = m_emission_pdfs[p]->compute_likelihood(*feature.get_vector());\n")
    hs = hs.replace(a, '  read_gk(base + ".gk");\n  akugpu_hook::attach(this, base);\n}\n')
    hs = hs.replace(b, "  if (akugpu_hook::score(this, *f.get_vector(), m_pdf_likelihoods, m_valid_pdf_likelihoods)) return;\n" + b)
    open(str(tmp_path / "FeatureGenerator_registered.cc"), "w").write(fg)
    open(str(tmp_path / "HmmSet_hooked.cc"), "w").write(hs)
    flags = ["-O1", "-std=gnu++11", "-DKISS_FFT", "-DDLLIMPORT=", "-fpermissive", "-w",
             "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + R + "/aku", "-I" + R + "/vendor/kiss_fft",
             "-I" + os.path.join(ROOT, "integration"), "-I" + os.path.join(ROOT, "aaltoasr_b200", "csrc", "host")]
    exe = str(tmp_path / "ref_phone_probs_gpu")
    subprocess.run(["g++"] + flags + ["-o", exe, R + "/aku/phone_probs.cc", str(tmp_path / "FeatureGenerator_registered.cc"),
                                      str(tmp_path / "HmmSet_hooked.cc"), os.path.join(ROOT, "tests", "cpp", "stub_akugpu.cc"),
                                      os.path.join(ROOT, "oracle", "_ref", "libaku_ref.a"), "-lm"], check=True, timeout=900)
    inner = str(tmp_path / "inner.cfg")
    open(inner, "w").write("x")
    cfg = str(tmp_path / "ref.feaconf")
    open(cfg, "w").write("module\n{\n  name gpu\n  type gpu_frontend\n  config %s\n}\n" % inner)
    rng = np.random.default_rng(3)
    base = str(tmp_path / "m")
    model = dict(mix_offsets=np.arange(5, dtype=np.int32), mix_gauss=np.arange(4, dtype=np.int32), mix_weight=np.ones(4),
                 means=rng.standard_normal((4, 3)) + 3, covs=rng.uniform(0.5, 2, (4, 3)))
    formats.write_model(base, **model)
    wav = str(tmp_path / "a.wav")
    formats.write_wav(wav, np.zeros(1280, np.int16), 16000)           # 10 frames
    rec = str(tmp_path / "r")
    open(rec, "w").write("audio=%s lna=a.lna\n" % wav)
    cfg_t = str(tmp_path / "ref_t.feaconf")          # a CPU-side reference module after the GPU module
    open(cfg_t, "w").write(open(cfg).read() + "module\n{\n  name t\n  type lin_transform\n  sources gpu\n  bias 1 0 0\n}\n")
    env = dict(os.environ)
    out = {}
    for tag, hook, c in (("gpu", "1", cfg), ("cpu", "", cfg), ("gpu_t", "1", cfg_t)):
        log = str(tmp_path / ("log_" + tag))
        od = tmp_path / ("o_" + tag)
        od.mkdir()
        env["AKUGPU_STUB_LOG"] = log
        env["AKUGPU_HOOK"] = hook
        r = subprocess.run([exe, "-b", base, "-c", c, "-r", rec, "-o", str(od), "-N"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           env=env, timeout=120)          # -N: the fake's likelihoods differ between frames only before normalisation
        assert r.returncode == 0, r.stderr
        out[tag] = (open(str(od / "a.lna"), "rb").read(), open(log).read().splitlines())
    f = np.arange(10, dtype=np.float64)
    # vectors straight from the GPU module: the hook scores the whole utterance in ONE call, then serves the frames
    blob, log = out["gpu"]
    assert [ln for ln in log if ln.startswith("gmm_score")] == ["gmm_score frames=10 precision=1"] and "model_read " + base in log
    lik = ((1 + f + (f + 0.25))[:, None] * (np.arange(4) + 1)[None, :]) / 1000.0
    want, _ = oracle_np.lna_records(lik, 2, normalize=False)
    assert blob[:5] == b"\x00\x00\x00\x04\x02" and blob[5:] == want.tobytes()
    # vectors changed by a CPU module downstream: the single-frame path, one call per frame
    blob_t, log_t = out["gpu_t"]
    assert [ln for ln in log_t if ln.startswith("gmm_score")] == ["gmm_score frames=1 precision=1"] * 10
    lik_t = ((1 + (f + 1) + (f + 0.25))[:, None] * (np.arange(4) + 1)[None, :]) / 1000.0
    want_t, _ = oracle_np.lna_records(lik_t, 2, normalize=False)
    assert blob_t[5:] == want_t.tobytes() and blob_t != blob
    blob0, log0 = out["cpu"]
    assert not any(ln.startswith("gmm_score") or ln.startswith("model_read") for ln in log0)        # CPU tool: the library is not asked
    assert len(blob0) == len(blob) and blob0 != blob
    # the CPU path of the same binary is the reference's arithmetic: the oracle on the plugin's features
    feats = np.stack([f, f + 0.25, f + 0.5], axis=1)
    want0, _ = oracle_np.lna_records(oracle_np.state_likelihoods(model, feats), 2, normalize=False)
    assert blob0[5:] == want0.tobytes()


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_feacat_gpu")), reason="oracle/_ref not built")
def test_prebuilt_reference_tools_on_the_library(aku_tests, tmp_path):
    """oracle/_ref/ref_feacat_gpu / ref_phone_probs_gpu = the reference's literal tools with the plugin registered and the
    hook in place, linked against the real libakugpu.so (oracle/build_ref.sh).  Without a `gpu_frontend` module they ARE
    the CPU tools (same bytes as ref_feacat); with one they need the device and say so."""
    import subprocess
    import torch
    wav = str(tmp_path / "short.wav")
    formats.write_wav(wav, aku_tests["short_wav"], int(aku_tests["sample_rate"]))
    cfg = str(tmp_path / "mfcc_p_dd.feaconf")
    open(cfg, "w").write(aku_tests["mfcc_p_dd_cfg"])
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    outs = [subprocess.run([os.path.join(ref_dir, exe), "-c", cfg, "-s", "-10", "-e", "80", wav], stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, timeout=120) for exe in ("ref_feacat", "ref_feacat_gpu")]
    assert outs[0].returncode == 0 and outs[1].returncode == 0 and outs[0].stdout == outs[1].stdout and len(outs[0].stdout) > 1000
    if torch.cuda.is_available():
        return
    gcfg = str(tmp_path / "gpu.feaconf")
    open(gcfg, "w").write("module\n{\n  name gpu\n  type gpu_frontend\n  config %s\n}\n" % cfg)
    r = subprocess.run([os.path.join(ref_dir, "ref_feacat_gpu"), "-c", gcfg, wav], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"no CPU fallback" in r.stderr


def test_library_model_readers_on_the_cpu(ref_small, ref_edge, ref_full, ref_clust, tmp_path):
    """The library's own .gk / .mc / .ph and .gcl readers (csrc/model.cu, pure host code exported from libakugpu.so) run
    without a device: what they read equals the arrays the files were written from (weights normalised like
    Mixture::normalize_weights, full covariances, the .gcl reader's repeated last pair and the moment-matched centres
    of the oracle), the sizes agree with the reference's HmmSet on the same files, and malformed files end in an error
    that names the problem -- never in a crash."""
    import subprocess
    import sys
    sys.path.insert(0, ROOT)
    from oracle import oracle_np, ref
    exe = str(tmp_path / "model_harness")
    libdir = os.path.join(ROOT, "aaltoasr_b200")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-I/usr/local/cuda/include", "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "model_harness.cc"), "-L" + libdir, "-lakugpu", "-Wl,-rpath," + libdir],
                   check=True, timeout=300)

    def load(path, dtypes):
        b, pos, out = open(path, "rb").read(), 0, []
        for dt in dtypes:
            n = int(np.frombuffer(b, "<i8", 1, pos)[0])
            pos += 8
            out.append(np.frombuffer(b, dt, n, pos))
            pos += n * np.dtype(dt).itemsize
        assert pos == len(b)
        return out

    for k, g in enumerate((ref_small, ref_edge, ref_full)):
        m = g["model"]
        base = str(tmp_path / ("m%d" % k))
        formats.write_model(base, **m)
        r = subprocess.run([exe, "read", base, base + ".bin"], stdout=subprocess.PIPE, timeout=60)
        assert r.returncode == 0, r.stdout
        hdr, off, idx, w, mean, cov, fidx, fcov = load(base + ".bin", ["<i4", "<i4", "<i4", "<f8", "<f8", "<f8", "<i4", "<f8"])
        G, D = m["means"].shape
        S = len(m["mix_offsets"]) - 1
        full = np.asarray(m["full_mask"], dtype=bool) if "full_mask" in m else np.zeros(G, dtype=bool)
        assert list(hdr) == [S, G, D, int(full.sum())]
        assert np.array_equal(off, m["mix_offsets"]) and np.array_equal(idx, m["mix_gauss"])
        want_w = np.asarray(m["mix_weight"], dtype=np.float64).copy()
        for s_ in range(S):
            a, b = m["mix_offsets"][s_], m["mix_offsets"][s_ + 1]
            tot = 0.0
            for v in want_w[a:b]:
                tot += v
            want_w[a:b] /= tot
        assert np.array_equal(w, want_w)
        assert np.array_equal(mean.reshape(G, D), m["means"])
        if "covs" in m:
            assert np.array_equal(cov.reshape(G, D)[~full], np.asarray(m["covs"])[~full])
        if full.any():
            assert np.array_equal(fidx >= 0, full)
            assert np.array_equal(fcov.reshape(-1, D, D), np.asarray(m["full_covs"])[full])
        if ref.available():
            M = ref.Model(base)
            assert (M.S, M.G, M.D) == (S, G, D)
            M.close()
    # clustering file: pairs as the reference's loop reads them, centres as the oracle merges them
    m = ref_clust["model"]
    base = str(tmp_path / "mc")
    formats.write_model(base, **m)
    gcl = str(tmp_path / "c.gcl")
    open(gcl, "w").write(ref_clust["gcl"])
    r = subprocess.run([exe, "gcl", base, gcl, base + ".bin"], stdout=subprocess.PIPE, timeout=60)
    assert r.returncode == 0, r.stdout
    hdr, g2c, sizes, cmean, ccov = load(base + ".bin", ["<i4", "<i4", "<i4", "<f8", "<f8"])
    n, gi, ci = oracle_np.parse_clustering(ref_clust["gcl"])
    cm, cc, members = oracle_np.cluster_centers(m, n, gi, ci)
    assert hdr[0] == n == 12 and list(sizes) == [len(L) for L in members] and sum(sizes) == len(gi)      # incl. the repeated last pair
    want_g2c = np.full(m["means"].shape[0], -1)
    for c, L in enumerate(members):
        want_g2c[L] = c
    assert np.array_equal(g2c, want_g2c) and (g2c < 0).sum() == 3
    assert np.allclose(cmean.reshape(n, -1), cm, rtol=1e-14, atol=0) and np.allclose(ccov.reshape(n, -1), cc, rtol=1e-13, atol=0)
    # malformed files
    base = str(tmp_path / "bad")
    formats.write_model(base, **ref_small["model"])
    orig = {e: open(base + e).read() for e in (".gk", ".mc", ".ph")}
    gk, mc = orig[".gk"].splitlines(), orig[".mc"].splitlines()
    cases = [(".gk", "\n".join(gk[:5]) + "\n", "unexpected end of file"),
             (".gk", "\n".join(gk[:3] + [" ".join(gk[3].split()[:20])] + gk[4:]) + "\n", "expected number"),
             (".gk", "\n".join([gk[0], gk[1].replace("diag", "foo")] + gk[2:]) + "\n", "Unknown model type"),
             (".gk", "x y z\n" + "\n".join(gk[1:]) + "\n", "expected integer"),
             (".gk", "", "expected integer"),
             (".mc", "\n".join([mc[0], "1 99999 1.0"] + mc[2:]) + "\n", "refers to Gaussian 99999 of 96"),
             (".mc", "\n".join(mc[:3]) + "\n", "expected integer"),
             (".ph", None, "could not open"),
             (".ph", "HELLO\n" + "\n".join(orig[".ph"].splitlines()[1:]) + "\n", "first token is not PHONE")]
    for ext, text, msg in cases:
        for e in orig:
            open(base + e, "w").write(orig[e])
        if text is None:
            os.remove(base + ext)
        else:
            open(base + ext, "w").write(text)
        r = subprocess.run([exe, "read", base, base + ".bin"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
        assert r.returncode == 1 and r.stdout.startswith(b"error -") and msg.encode() in r.stdout, (ext, msg, r.returncode, r.stdout)
    for e in orig:
        open(base + e, "w").write(orig[e])
    for text, msg in (("99\n0 0\n", "seems insensible"), ("2\n0 0\n5000 1\n", "Gauss index out of bounds"), ("2\n0 7\n", "out of bounds")):
        open(gcl, "w").write(text)
        r = subprocess.run([exe, "gcl", base, gcl, base + ".bin"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
        assert r.returncode == 1 and msg.encode() in r.stdout, (text, r.returncode, r.stdout)


def test_library_expanded_form_and_slot_tables_on_the_cpu(ref_small, ref_edge, ref_full, tmp_path):
    """Host mathematics of the tensor-core packers, run on the CPU from libakugpu.so: tc_expanded_params (csrc/gmm_tc.cu)
    -- <expand(x - c), theta_g> + gconst_g reproduces every Gaussian's log-likelihood (diagonal: the oracle's; full: the
    reference's exponential form) and q_g is the cancelling magnitude the conditioning guard is built on -- and
    tc_build_slots: every component in exactly one slot, states contiguous, never across a warp's group of slots."""
    import subprocess
    import sys
    sys.path.insert(0, ROOT)
    from oracle import oracle_np
    exe = str(tmp_path / "model_harness")
    libdir = os.path.join(ROOT, "aaltoasr_b200")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-I/usr/local/cuda/include", "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "model_harness.cc"), "-L" + libdir, "-lakugpu", "-Wl,-rpath," + libdir],
                   check=True, timeout=300)

    def load(path, dtypes):
        b, pos, out = open(path, "rb").read(), 0, []
        for dt in dtypes:
            n = int(np.frombuffer(b, "<i8", 1, pos)[0])
            pos += 8
            out.append(np.frombuffer(b, dt, n, pos))
            pos += n * np.dtype(dt).itemsize
        assert pos == len(b)
        return out

    # diagonal pools
    for k, g in enumerate((ref_small, ref_edge)):
        m = g["model"]
        base = str(tmp_path / ("d%d" % k))
        formats.write_model(base, **m)
        assert subprocess.run([exe, "expand", base, base + ".bin"], timeout=60).returncode == 0
        cen, theta, gconst, q, qmax = load(base + ".bin", ["<f8"] * 5)
        G, D = m["means"].shape
        theta = theta.reshape(G, 2 * D)
        prec, cst = oracle_np.gaussian_params(m["means"], m["covs"])
        mc = m["means"] - cen[None, :]
        assert np.allclose(q, 0.5 * (prec * mc * mc).sum(axis=1), rtol=1e-12) and qmax[0] == q.max()
        x = g["feats"][::7]
        y = x - cen[None, :]
        ll = np.concatenate([y * y, y], axis=1) @ theta.T + gconst[None, :]
        want = oracle_np._diag_loglik(x, m["means"], prec, cst)
        ok = np.isfinite(want) & (np.abs(want) < 1e6)
        assert ok.mean() > 0.9
        assert (np.abs(ll - want)[ok] <= 1e-9 * (1 + np.abs(want[ok]) + q[None, :].repeat(x.shape[0], 0)[ok])).all()
        if k == 0:
            assert qmax[0] < 200          # the synthetic "realistic" model is well inside the guard (TC_Q_MAX)
        else:
            assert qmax[0] > 200          # the edge model is what the hybrid split exists for
    # an all-full pool: the reference's exponential form around the centre
    m = ref_full["model"]
    idx = np.nonzero(m["full_mask"])[0]
    base = str(tmp_path / "f")
    off = np.arange(0, len(idx) + 1, 3, dtype=np.int32)
    formats.write_model(base, off, np.arange(len(idx), dtype=np.int32), np.ones(len(idx)), m["means"][idx],
                        full_covs=np.asarray(m["full_covs"])[idx])
    assert subprocess.run([exe, "expand", base, base + ".bin"], timeout=60).returncode == 0
    cen, theta, gconst, q, qmax = load(base + ".bin", ["<f8"] * 5)
    D = m["means"].shape[1]
    L = D * (D + 3) // 2
    theta = theta.reshape(len(idx), L)
    x = ref_full["feats"][::9]
    phi = oracle_np.exponential_feature(x - cen[None, :])
    ll = phi @ theta.T + gconst[None, :]
    for j, gi in enumerate(idx[:12]):
        cov = np.asarray(m["full_covs"][gi], dtype=np.float64)
        P = np.linalg.inv(cov)
        d = x - m["means"][gi][None, :]
        want = -0.5 * np.einsum("fi,ij,fj->f", d, P, d) + 0.5 * np.linalg.slogdet(P)[1]
        assert np.abs(ll[:, j] - want).max() <= 1e-7 * (1 + np.abs(want).max() + q[j])
    # slot tables (16 components per slot; a warp owns `group` consecutive slots of a tile)
    m = ref_edge["model"]
    base = str(tmp_path / "d1")
    S = len(m["mix_offsets"]) - 1
    for group in (4, 8):
        assert subprocess.run([exe, "slots", base, str(group), base + ".slots"], timeout=60).returncode == 0
        t = load(base + ".slots", ["<i4"] * 6)
        for (st, k0, fl), skipped in ((t[:3], set()), (t[3:], set(range(0, S, 5)))):
            seen = {}
            for i, s_ in enumerate(st):
                if s_ < 0:
                    assert fl[i] == 3
                    continue
                seen.setdefault(int(s_), []).append(i)
            assert set(seen) == set(range(S)) - skipped
            for s_, slots in seen.items():
                K = m["mix_offsets"][s_ + 1] - m["mix_offsets"][s_]
                assert slots == list(range(slots[0], slots[0] + len(slots))) and len(slots) == max(1, -(-K // 16))
                assert slots[0] // group == slots[-1] // group                      # no state straddles a warp's group
                assert [int(k0[i]) for i in slots] == [16 * j for j in range(len(slots))]
                assert [int(fl[i]) for i in slots] == [((j == 0) << 1) | (j == len(slots) - 1) for j in range(len(slots))]


def test_parsers_survive_mutated_inputs(ref_small, ref_clust, ref_spk, tmp_path):
    """Seeded mutation fuzzing of every host-side parser that takes a user's file: the library's model / clustering readers
    (from libakugpu.so) and the adapters' audio, recipe and speaker-file readers (built with ASan + UBSan): each mutated
    input ends in success or in a reported error -- no signal, no sanitizer report.  (600 + 500 mutations were run by
    hand when this was written; the suite keeps a short deterministic sample.)"""
    import random
    import subprocess
    libdir = os.path.join(ROOT, "aaltoasr_b200")
    mh = str(tmp_path / "model_harness")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I/usr/local/cuda/include", "-o", mh, os.path.join(ROOT, "tests", "cpp", "model_harness.cc"),
                    "-L" + libdir, "-lakugpu", "-Wl,-rpath," + libdir], check=True, timeout=300)
    hz = str(tmp_path / "spk_harness_asan")
    subprocess.run(["g++", "-g", "-O1", "-std=c++11", "-fsanitize=address,undefined", "-o", hz,
                    os.path.join(ROOT, "tests", "cpp", "spk_harness.cc")], check=True, timeout=300)
    rnd = random.Random(20261017)

    def mutate(b, alphabet):
        b = bytearray(b)
        for _ in range(rnd.randint(1, 6)):
            op = rnd.random()
            if op < 0.4 and b:
                b[rnd.randrange(len(b))] = rnd.choice(alphabet)
            elif op < 0.6 and b:
                i = rnd.randrange(len(b))
                del b[i:i + rnd.randint(1, 60)]
            elif op < 0.8:
                i = rnd.randrange(len(b) + 1)
                b[i:i] = bytes(rnd.choice(alphabet) for _ in range(rnd.randint(1, 20)))
            else:
                b = b[:rnd.randrange(len(b) + 1)]
        return bytes(b)

    m = ref_small["model"]
    off = m["mix_offsets"][:7]
    base = str(tmp_path / "fm")
    formats.write_model(base, off, m["mix_gauss"][:off[-1]], m["mix_weight"][:off[-1]], m["means"], m["covs"])
    orig = {e: open(base + e, "rb").read() for e in (".gk", ".mc", ".ph")}
    text = b"0123456789 .-e\nxdiagful"
    outcomes = set()
    for it in range(48):
        for e in orig:
            open(base + e, "wb").write(orig[e])
        if it % 4 < 3:
            e = (".gk", ".mc", ".ph")[it % 4]
            open(base + e, "wb").write(mutate(orig[e], text))
            args = [mh, "read", base, str(tmp_path / "o.bin")]
        else:
            open(str(tmp_path / "f.gcl"), "wb").write(mutate(ref_clust["gcl"].encode(), text))
            args = [mh, "gcl", base, str(tmp_path / "f.gcl"), str(tmp_path / "o.bin")]
        r = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
        assert r.returncode in (0, 1), (it, r.returncode, r.stdout[:200], r.stderr[:200])
        outcomes.add(r.returncode)
    assert outcomes == {0, 1}
    pcm = (np.arange(3000) % 200).astype(np.int16)
    formats.write_wav(str(tmp_path / "seed.wav"), pcm, 16000)
    wav = open(str(tmp_path / "seed.wav"), "rb").read()
    rec = b"audio=a.wav lna=b.lna speaker=s start-time=0.5\naudio=c.wav\n# c\n\naudio=d.wav utterance=u end-time=3\n"
    spk = (ref_spk["spkc"][:3000] + "\nspeaker z\n{\n  model cmllr\n  {\n    unitmode UNIT_NO\n    w1 0 1 0 0 0 1\n  }\n}\n").encode()
    anybyte = bytes(range(256))
    for it in range(45):
        if it % 3 == 0:
            open(str(tmp_path / "f.wav"), "wb").write(mutate(wav, anybyte))
            args = [hz, "audio", str(tmp_path / "f.wav"), "16000", "0"]
        elif it % 3 == 1:
            open(str(tmp_path / "f.recipe"), "wb").write(mutate(rec, anybyte))
            args = [hz, "recipe", str(tmp_path / "f.recipe"), str(rnd.randint(0, 4)), str(rnd.randint(0, 4)), "1"]
        else:
            open(str(tmp_path / "f.spkc"), "wb").write(mutate(spk, anybyte))
            args = [hz, str(tmp_path / "f.spkc"), "2", "alice", "z", "bob", "nobody"]
        r = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
        assert r.returncode in (0, 1) and b"AddressSanitizer" not in r.stderr and b"runtime error" not in r.stderr, (it, r.returncode, r.stderr[:400])


def test_recipe_reader_three_ways(tmp_path):
    """aku::Recipe::read + sort_infos as phone_probs uses them (-B / -I / --sort-recipe, aku/phone_probs.cc:137-142): the
    C++ adapter (akugpu::Recipe), the Python mirror (formats.read_recipe / sort_recipe) and -- when oracle/_ref is built --
    the reference's own class give the same utterance lists for every batch of several recipes: inherited keys (also
    across batch borders), comments, blank lines, tabs, fewer lines than batches, and the reference's errors."""
    import subprocess
    import sys
    sys.path.insert(0, ROOT)
    from oracle import ref
    exe = str(tmp_path / "spk_harness")
    subprocess.run(["g++", "-O1", "-std=c++11", "-Wall", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "cpp", "spk_harness.cc")],
                   check=True, timeout=300)
    rng = np.random.default_rng(5)
    recipes = []
    lines = []
    for i in range(23):
        f = ["audio=/d/a%d.wav" % i]
        if i % 3 != 1:
            f.append("lna=a%d.lna" % i)
        if i % 4 == 0:
            f.append("speaker=spk%d" % rng.integers(0, 4))
        if i % 5 == 2:
            f.append("utterance=u%d" % i)
        if i % 6 == 3:
            f += ["start-time=%.2f" % (0.1 * i), "end-time=%.2f" % (0.1 * i + 1)]
        lines.append((" " if i % 2 else "\t").join(f))
        if i % 7 == 0:
            lines.append("# a comment")
        if i % 8 == 0:
            lines.append("   ")
    recipes.append("\n".join(lines) + "\n")
    recipes.append("audio=x.wav speaker=b\naudio=y.wav speaker=a\naudio=z.wav\n  audio=w.wav speaker=b transcript=t.phn")   # no final newline
    recipes.append("")

    def fmt(rows):
        return ["%s|%s|%s|%s|%g|%g" % r for r in rows]

    for k, text in enumerate(recipes):
        path = str(tmp_path / ("r%d.recipe" % k))
        open(path, "w").write(text)
        for B, I in [(0, 0), (1, 1), (2, 1), (2, 2), (3, 2), (5, 1), (5, 5), (7, 4), (30, 3), (30, 29)]:
            for srt in (0, 1):
                r = subprocess.run([exe, "recipe", path, str(B), str(I), str(srt)], stdout=subprocess.PIPE, timeout=60)
                assert r.returncode == 0, r.stdout
                cpp = r.stdout.decode().splitlines()
                infos = formats.read_recipe(path, B, I)
                if srt:
                    infos = formats.sort_recipe(infos)
                py = fmt([(d.get("audio", ""), d.get("lna", ""), d.get("speaker", ""), d.get("utterance", ""),
                           float(d.get("start-time", 0)), float(d.get("end-time", 0))) for d in infos])
                assert cpp == py, (k, B, I, srt)
                if ref.available():
                    assert cpp == fmt(ref.recipe_read(path, B, I, bool(srt))), (k, B, I, srt)
        # all batches together = the whole recipe, in order
        whole = formats.read_recipe(path)
        for B in (2, 3, 7):
            assert sum((formats.read_recipe(path, B, i) for i in range(1, B + 1)), []) == whole
    bad = str(tmp_path / "bad.recipe")
    for text in ("audio=a.wav lna\n", "audio=a.wav lna=\n", "audio=a.wav x=1=2\n"):
        open(bad, "w").write(text)
        r = subprocess.run([exe, "recipe", bad, "0", "0", "0"], stdout=subprocess.PIPE, timeout=60)
        assert r.returncode == 1 and b"Invalid recipe line: audio=a.wav" in r.stdout
        with pytest.raises(ValueError, match="Invalid recipe line"):
            formats.read_recipe(bad)
        if ref.available():
            with pytest.raises(RuntimeError, match="Invalid recipe line"):
                ref.recipe_read(bad)
    open(bad, "w").write("audio=a.wav\n")
    r = subprocess.run([exe, "recipe", bad, "3", "4", "0"], stdout=subprocess.PIPE, timeout=60)
    assert r.returncode == 1 and b"Invalid batch index" in r.stdout
    with pytest.raises(ValueError, match="Invalid batch index"):
        formats.read_recipe(bad, 3, 4)


def test_formats_roundtrip(tmp_path):
    pcm = synth.synth_audio(5, 4000)
    formats.write_wav(str(tmp_path / "a.wav"), pcm, 16000)
    back, sr = formats.read_wav(str(tmp_path / "a.wav"))
    assert sr == 16000 and np.array_equal(back, pcm)
    rec = np.arange(2 * 3 * 2, dtype=np.uint8).reshape(2, 6)
    formats.write_lna(str(tmp_path / "a.lna"), rec, 3, 2)
    lp, S, nb = formats.read_lna(str(tmp_path / "a.lna"))
    assert (S, nb) == (3, 2) and lp.shape == (2, 3)
    assert np.allclose(lp[0, 1], -(2 * 256 + 3) / 1820.0)
    open(str(tmp_path / "r"), "w").write("audio=a.wav lna=a.lna speaker=x\naudio=b.wav lna=b.lna\n\n# c\n")
    infos = formats.read_recipe(str(tmp_path / "r"))
    assert len(infos) == 2 and infos[1]["speaker"] == "x"      # inherited, as in aku/Recipe.cc:82-90
