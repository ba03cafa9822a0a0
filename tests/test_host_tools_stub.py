"""CPU tests of the host tools' OWN logic -- aaltoasr_b200/csrc/host/phone_probs_main.cc and feacat_main.cc linked
against a FAKE of the C ABI (tests/cpp/stub_akugpu.cc: closed-form "features" and "LNA bytes", every call logged).
What is checked here is what the tools add above the library: the recipe loop of aku/phone_probs.cc:135-267 (output
names, -o / -a / -n, start-time / end-time cropping, -B / -I, --sort-recipe, one GPU call per run of utterances that
share a speaker, speaker files incl. sticky `model cmllr`), the frame ranges and output formats of aku/feacat.cc, and
the error paths.  The real library behind the same tools is covered by tests/test_gpu_host.py on the GPU."""
import gzip
import os
import struct
import subprocess

import numpy as np
import pytest

from aaltoasr_b200 import formats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "aaltoasr_b200", "csrc", "host")
STUB = os.path.join(ROOT, "tests", "cpp", "stub_akugpu.cc")
S = 4


@pytest.fixture(scope="module")
def tools(tmp_path_factory):
    d = tmp_path_factory.mktemp("stubtools")
    out = {}
    for name in ("phone_probs", "feacat"):
        exe = str(d / name)
        subprocess.run(["g++", "-O1", "-std=c++11", "-Wall", "-Werror", "-o", exe, os.path.join(HOST, name + "_main.cc"), STUB],
                       check=True, timeout=300)
        out[name] = exe
    return out


def lna_file(frames, lnabytes=2, normalize=1, cmllr=0, shift=0):
    """What the stub's akugpu_phone_probs writes for the utterance-local frames `frames`, behind the 5-byte header."""
    body = bytes((7 * f + 3 * s + b + 100 * normalize + 50 * cmllr + shift) & 255
                 for f in frames for s in range(S) for b in range(lnabytes))
    return struct.pack(">I", S) + bytes([lnabytes]) + body


SPKC = """speaker default
{
  shift
  {
    value 0
  }
}
speaker s1
{
  feature shift
  {
    value 1
  }
}
speaker s2
{
  shift
  {
    value 2
  }
  model cmllr
  {
    unitmode UNIT_NO
    w1 0.5 1 0 0 -0.5 0 1 0 0.25 0 0 1
  }
}
utterance default
{
  shift
  {
    value 0
  }
}
utterance u1
{
  shift
  {
    value 7
  }
}
"""


@pytest.fixture()
def case(tmp_path):
    n_samples = [1280, 2560, 640, 1920, 1300]          # 10, 20, 5, 15, 10 frames
    for i, n in enumerate(n_samples):
        formats.write_wav(str(tmp_path / ("a%d.wav" % i)), (np.arange(n) % 100).astype(np.int16), 16000)
    rec = str(tmp_path / "r.recipe")
    open(rec, "w").write(
        "audio=%(d)s/a0.wav lna=x0.lna speaker=s1\n"
        "# the speaker carries over\n"
        "audio=%(d)s/a1.wav lna=x1.lna\n"
        "\n"
        "audio=%(d)s/a2.wav speaker=s2 lna=x2.lna start-time=0.02 end-time=0.036\n"
        "audio=%(d)s/a3.wav lna=x3.lna speaker=s1 start-time=0 end-time=0\n"
        "audio=%(d)s/a4.wav speaker=s3 lna=x4.lna\n" % {"d": str(tmp_path)})
    cfg = str(tmp_path / "f.cfg")
    open(cfg, "w").write("module\n{\n  name shift\n  type stub\n}\n")
    spkc = str(tmp_path / "x.spkc")
    open(spkc, "w").write(SPKC)
    return dict(dir=tmp_path, recipe=rec, cfg=cfg, spkc=spkc, frames=[n // 128 for n in n_samples])


def run(exe, args, log=None, env=None, **kw):
    e = dict(os.environ)
    e.pop("AKUGPU_STUB_LOG", None)
    if log:
        e["AKUGPU_STUB_LOG"] = log
    if env:
        e.update(env)
    return subprocess.run([exe] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120, env=e, **kw)


def calls(log, what):
    return [ln for ln in open(log).read().splitlines() if ln.startswith(what)]


def test_phone_probs_recipe_loop(tools, case):
    c, exe = case, tools["phone_probs"]
    out = c["dir"] / "o"
    out.mkdir()
    log = str(c["dir"] / "log1")
    r = run(exe, ["-b", "mdl", "-c", c["cfg"], "-r", c["recipe"], "-o", str(out), "-i", "1"], log)
    assert r.returncode == 0, r.stderr.decode()
    fr = c["frames"]
    for i in (0, 1, 3, 4):
        assert open(str(out / ("x%d.lna" % i)), "rb").read() == lna_file(range(fr[i])), i
    assert open(str(out / "x2.lna"), "rb").read() == lna_file(range(2, 4))            # start-time / end-time in frames
    # one library call per utterance without -S (no reason to split) -- all five in one
    assert calls(log, "phone_probs") == ["phone_probs n_utts=5 samples=7700 precision=0 lnabytes=2 normalize=1"]
    assert calls(log, "model_read") == ["model_read mdl"] and calls(log, "load_config") == ["load_config " + c["cfg"]]
    assert r.stdout.decode().count("Processing file: ") == 5

    # -S: utterances that share a speaker share a call; `model cmllr` stays loaded for speakers without the entry
    log = str(c["dir"] / "log2")
    r = run(exe, ["-b", "mdl", "-c", c["cfg"], "-r", c["recipe"], "-o", str(out), "-S", c["spkc"], "--lnabytes=4", "-N"], log)
    assert r.returncode == 0, r.stderr.decode()
    assert [ln.split()[1] for ln in calls(log, "phone_probs")] == ["n_utts=2", "n_utts=1", "n_utts=1", "n_utts=1"]
    seq = [ln for ln in open(log).read().splitlines() if ln.split()[0] in ("set_parameters", "set_cmllr", "phone_probs")]
    assert seq[:2] == ["set_parameters shift value 1;", "phone_probs n_utts=2 samples=3840 precision=0 lnabytes=4 normalize=0"]
    assert "set_cmllr 0.5 1" in seq and seq.index("set_cmllr 0.5 1") < seq.index("phone_probs n_utts=1 samples=640 precision=0 lnabytes=4 normalize=0")
    assert open(str(out / "x0.lna"), "rb").read() == lna_file(range(fr[0]), 4, 0, 0, 1)
    assert open(str(out / "x2.lna"), "rb").read() == lna_file(range(2, 4), 4, 0, 1, 2)
    assert open(str(out / "x3.lna"), "rb").read() == lna_file(range(fr[3]), 4, 0, 1, 1)      # s1 again: cmllr still on
    assert open(str(out / "x4.lna"), "rb").read() == lna_file(range(fr[4]), 4, 0, 1, 0)      # unknown speaker: default block

    # --sort-recipe groups the speakers (stable): s1 x 3 in one call, then s2, s3
    log = str(c["dir"] / "log3")
    r = run(exe, ["-b", "mdl", "-c", c["cfg"], "-r", c["recipe"], "-o", str(out), "-S", c["spkc"], "--sort-recipe"], log)
    assert r.returncode == 0, r.stderr.decode()
    assert [ln.split()[1] for ln in calls(log, "phone_probs")] == ["n_utts=3", "n_utts=1", "n_utts=1"]
    assert open(str(out / "x3.lna"), "rb").read() == lna_file(range(fr[3]), 2, 1, 0, 1)      # before s2 now: no cmllr yet


def test_phone_probs_flags(tools, case):
    c, exe = case, tools["phone_probs"]
    out = c["dir"] / "o"
    out.mkdir()
    fr = c["frames"]
    base = ["-c", c["cfg"], "-r", c["recipe"], "-o", str(out)]
    # -a / --afname: named by the audio file; -g -m -p instead of -b; clustering flags reach the library
    log = str(c["dir"] / "log")
    r = run(exe, base + ["-g", "m.gk", "-m", "m.mc", "-p", "m.ph", "--afname", "-C", "c.gcl", "--eval-minc", "0.1", "--eval-ming=0.2"], log)
    assert r.returncode == 0, r.stderr.decode()
    assert sorted(os.listdir(str(out))) == ["a%d.lna" % i for i in range(5)]
    assert calls(log, "model_read") == ["model_read_files m.gk m.mc m.ph"]
    assert calls(log, "read_clustering") == ["read_clustering c.gcl"] and calls(log, "min_evals") == ["min_evals 0.1 0.2"]
    # -n: any existing file is skipped with the reference's warning (stat() == 0, aku/phone_probs.cc:180-190), empty or not
    for f in os.listdir(str(out)):
        os.remove(str(out / f))
    open(str(out / "x0.lna"), "wb").write(b"keep")
    open(str(out / "x1.lna"), "wb").close()
    log = str(c["dir"] / "logn")
    r = run(exe, base + ["-b", "mdl", "-n"], log)
    assert r.returncode == 0, r.stderr.decode()
    assert open(str(out / "x0.lna"), "rb").read() == b"keep" and open(str(out / "x1.lna"), "rb").read() == b""
    assert r.stderr.decode().count("WARNING: skipping existing lna file ") == 2
    assert ("WARNING: skipping existing lna file %s\n" % (out / "x1.lna")) in r.stderr.decode()
    assert calls(log, "phone_probs")[0].split()[1] == "n_utts=3"
    # -a keeps a leading dot (the reference strips the extension only at position > 0); a quote in a .gz name is no shell escape
    formats.write_wav(str(c["dir"] / ".hid"), (np.arange(640) % 100).astype(np.int16), 16000)
    rec = str(c["dir"] / "dot.recipe")
    open(rec, "w").write("audio=%s/.hid lna=unused.lna\naudio=%s/a2.wav lna=it's.lna.gz\n" % (c["dir"], c["dir"]))
    r = run(exe, ["-b", "mdl", "-c", c["cfg"], "-r", rec, "-o", str(out), "-a"])
    assert r.returncode == 0 and os.path.exists(str(out / ".hid.lna")), r.stderr.decode()
    r = run(exe, ["-b", "mdl", "-c", c["cfg"], "-r", rec, "-o", str(out)])
    assert r.returncode == 0 and gzip.open(str(out / "it's.lna.gz")).read() == lna_file(range(fr[2])), r.stderr.decode()
    # a negative start-time: frames before the file are generated through features_range + gmm_lna (the reference's loop
    # starts at the negative frame, border-replicated windows), then the file's own records follow; start == end == file
    # end writes a header-only file (used to index one past the buffer)
    rec = str(c["dir"] / "neg.recipe")
    open(rec, "w").write("audio=%s/a0.wav lna=n0.lna start-time=-0.024 end-time=0.016\n"
                         "audio=%s/a2.wav lna=n2.lna start-time=0.04 end-time=0.04\n" % (c["dir"], c["dir"]))
    log = str(c["dir"] / "logneg")
    r = run(exe, ["-b", "mdl", "-c", c["cfg"], "-r", rec, "-o", str(out)], log)
    assert r.returncode == 0, r.stderr.decode()
    assert calls(log, "features_range") == ["features_range -3 0"] and calls(log, "gmm_lna") == ["gmm_lna frames=3 precision=0 lnabytes=2 normalize=1"]
    neg = bytes((200 + 7 * f + 3 * s + b + 100) & 255 for f in range(3) for s in range(S) for b in range(2))
    assert open(str(out / "n0.lna"), "rb").read() == lna_file([])[:5] + neg + lna_file(range(2))[5:]
    assert open(str(out / "n2.lna"), "rb").read() == lna_file([])[:5]
    # -B / -I: the second of two batches is lines 4-5 (3 + 2); both flags are needed
    for f in os.listdir(str(out)):
        os.remove(str(out / f))
    r = run(exe, base + ["-b", "mdl", "-B", "2", "-I", "2"])
    assert r.returncode == 0 and sorted(os.listdir(str(out))) == ["x3.lna", "x4.lna"]
    r = run(exe, base + ["-b", "mdl", "-B", "2"])
    assert r.returncode == 1 and b"Must give both --batch and --bindex" in r.stderr
    r = run(exe, base + ["-b", "mdl", "-B", "2", "-I", "3"])
    assert r.returncode == 1 and b"Invalid batch index" in r.stderr
    # io::Stream output names: a trailing .gz goes through gzip
    rec = str(c["dir"] / "gz.recipe")
    open(rec, "w").write("audio=%s/a2.wav lna=z.lna.gz\n" % c["dir"])
    r = run(exe, ["-b", "mdl", "-c", c["cfg"], "-r", rec, "-o", str(out)])
    assert r.returncode == 0, r.stderr.decode()
    assert gzip.open(str(out / "z.lna.gz")).read() == lna_file(range(fr[2]))
    # errors worded like the reference's
    for args, msg in ((["-r", c["recipe"], "-b", "m"], b"Must give --config"),
                      (["-c", c["cfg"], "-b", "m"], b"Must give --recipe"),
                      (base, b"Must give either --base or all --gk, --mc and --ph"),
                      (base + ["-b", "m", "--lnabytes=3"], b"Invalid number of LNA bytes"),
                      (base + ["-b", "m", "--frobnicate"], b"unknown option --frobnicate"),
                      (["-c", str(c["dir"] / "none.cfg"), "-r", c["recipe"], "-b", "m"], b"could not open file")):
        r = run(exe, args)
        assert r.returncode == 1 and msg in r.stderr, (args, r.stderr)
    r = run(exe, base + ["-b", "m"], env={"AKUGPU_STUB_MODEL_DIM": "5"})
    assert r.returncode == 1 and b"Gaussian dimension is 5 but feature dimension is 3." in r.stderr
    formats.write_wav(str(c["dir"] / "a2.wav"), np.zeros(640, np.int16), 8000)
    r = run(exe, base + ["-b", "m"])
    assert r.returncode == 1 and b"Audio file sample rate (8000 Hz) and model configuration (16000 Hz) don't agree." in r.stderr


def rows_of(frames, shift=0.0):
    return np.array([[f + 0.25 * d + shift for d in range(3)] for f in frames])


def test_feacat_ranges_and_formats(tools, case):
    c, exe = case, tools["feacat"]
    wav = str(c["dir"] / "a1.wav")                      # 20 frames
    r = run(exe, ["-c", c["cfg"], wav])
    assert r.returncode == 0, r.stderr.decode()
    lines = r.stdout.decode().splitlines()
    assert len(lines) == 20 and lines[3] == "  3.0000   3.2500   3.5000 "            # "%8.4f " per component
    # frames outside the file: the library's border handling, one call for the whole range; end frame inclusive
    log = str(c["dir"] / "log")
    r = run(exe, ["-c", c["cfg"], "-s", "-2", "-e", "3", wav], log)
    got = np.array([[float(x) for x in ln.split()] for ln in r.stdout.decode().splitlines()])
    assert np.array_equal(got, rows_of([0, 0, 0, 1, 2, 3])) and calls(log, "features_range") == ["features_range -2 4"]
    r = run(exe, ["-c", c["cfg"], "--start-frame=18", wav])                          # open end: to the end of the file
    assert len(r.stdout.decode().splitlines()) == 2
    r = run(exe, ["-c", c["cfg"], "--start-frame", "25", wav])
    assert r.returncode == 0 and r.stdout == b""
    # raw output with the int32 dimension header; descending range
    r = run(exe, ["-c", c["cfg"], "--raw-output", "-H", "-s", "5", "-e", "2", wav])
    assert r.stdout[:4] == struct.pack("<i", 3)
    assert np.array_equal(np.frombuffer(r.stdout[4:], "<f4").reshape(-1, 3), rows_of([5, 4, 3, 2]).astype(np.float32))
    r = run(exe, ["-c", c["cfg"], "-H", "-s", "1", "-e", "1", wav])
    assert b"Warning: header is only written in raw output mode" in r.stderr and len(r.stdout.decode().splitlines()) == 1
    # audio on standard input; --write-config
    wcfg = str(c["dir"] / "w.cfg")
    r = run(exe, ["-c", c["cfg"], "-w", wcfg, "-e", "1", "-"], input=open(wav, "rb").read())
    assert r.returncode == 0 and len(r.stdout.decode().splitlines()) == 2 and open(wcfg).read() == open(c["cfg"]).read()
    # -S / -d / -u: speaker, then utterance parameters
    r = run(exe, ["-c", c["cfg"], "-S", c["spkc"], "-d", "s2", "-e", "0", wav])
    assert [float(x) for x in r.stdout.decode().split()] == [2.0, 2.25, 2.5]
    r = run(exe, ["-c", c["cfg"], "-S", c["spkc"], "-d", "s1", "-u", "u1", "-e", "0", wav])
    assert [float(x) for x in r.stdout.decode().split()] == [7.0, 7.25, 7.5]
    r = run(exe, ["-c", c["cfg"], "-S", c["spkc"], "-d", "nobody", "-e", "0", wav])     # default speaker block
    assert [float(x) for x in r.stdout.decode().split()] == [0.0, 0.25, 0.5]
    # a `pre` configuration reads the raw file feacat -H --raw-output wrote
    pre_cfg = str(c["dir"] / "pre.cfg")
    open(pre_cfg, "w").write("module\n{\n  name pre\n  type pre\n  dim 3\n}\n")
    raw = str(c["dir"] / "f.raw")
    open(raw, "wb").write(run(exe, ["-c", c["cfg"], "--raw-output", "-H", "-s", "4", "-e", "9", wav]).stdout)
    r = run(exe, ["-c", pre_cfg, "-s", "-1", "-e", "6", raw])
    got = np.array([[float(x) for x in ln.split()] for ln in r.stdout.decode().splitlines()])
    assert np.array_equal(got, rows_of([4, 4, 5, 6, 7, 8, 9, 9]))
    # PreModule::set_file (aku/FeatureModules.cc:602-627): a header dimension other than the configured one is refused
    # (it used to be trusted: rows mis-strided, host buffer over-read), and `legacy_file 1` = a ONE-byte header
    body = open(raw, "rb").read()[4:]
    bad = str(c["dir"] / "bad.raw")
    open(bad, "wb").write(struct.pack("<i", 2) + body)
    r = run(exe, ["-c", pre_cfg, bad])
    assert r.returncode == 1 and b"PreModule: The file has invalid dimension" in r.stderr
    legacy_cfg = str(c["dir"] / "legacy_pre.cfg")
    open(legacy_cfg, "w").write("module\n{\n  name pre\n  type pre\n  dim 3\n  legacy_file 1\n}\n")
    leg = str(c["dir"] / "leg.raw")
    open(leg, "wb").write(bytes([3]) + body + b"\x00\x00")        # + a trailing partial row, never read
    r = run(exe, ["-c", legacy_cfg, "-s", "0", "-e", "1", leg])
    got = np.array([[float(x) for x in ln.split()] for ln in r.stdout.decode().splitlines()])
    assert np.array_equal(got, rows_of([4, 5]))
    r = run(exe, ["-c", legacy_cfg, raw])                              # int32 header read as one byte + misaligned rows
    assert r.returncode == 0 or b"PreModule" in r.stderr
    # errors
    assert run(exe, ["-c", c["cfg"]]).returncode == 1 and run(exe, ["-c", c["cfg"], wav, wav]).returncode == 1
    r = run(exe, [wav])
    assert r.returncode == 1 and b"Must give --config" in r.stderr
    r = run(exe, ["-c", c["cfg"], "-G", "0.1", wav])
    assert r.returncode == 1 and b"not supported" in r.stderr
    r = run(exe, ["-c", c["cfg"], str(c["dir"] / "none.wav")])
    assert r.returncode == 1 and b"exception: AudioReader::open(): could not open file" in r.stderr


class FakeEngine:
    """The stub's formulas behind the Python engine surface PhoneProbs.run_recipe uses."""
    sample_rate, frame_rate, num_states = 16000, 125.0, S

    def phone_probs(self, pcm, uo=None, precision=0, lnabytes=2, normalize=True):
        uo = np.array([0, pcm.size]) if uo is None else uo
        fo = np.concatenate([[0], np.cumsum(np.diff(uo) // 128)]).astype(np.int64)
        rec = np.concatenate([np.frombuffer(lna_file(range(int(fo[k + 1] - fo[k])), lnabytes, 1 if normalize else 0)[5:], np.uint8)
                              for k in range(len(uo) - 1)]).reshape(int(fo[-1]), S * lnabytes)
        return rec, fo, None


def test_python_run_recipe_matches_the_cpp_tool(tools, case):
    """hostapi.PhoneProbs.run_recipe (the Python mirror used by the GPU tests) and the C++ tool write the same files for
    the same recipe: names, cropping, batches, --sort-recipe, --no-overwrite."""
    from aaltoasr_b200 import PhoneProbs
    c = case
    for k, (flags, kw) in enumerate(((["-B", "2", "-I", "1"], dict(batch=2, bindex=1)),
                                     ([], {}),
                                     (["--sort-recipe", "--lnabytes=4", "-N"], dict(sort_recipe=True)),
                                     (["-a", "-n"], dict(audio_ext_lna=True, no_overwrite=True)))):
        o_cpp, o_py = c["dir"] / ("cpp%d" % k), c["dir"] / ("py%d" % k)
        o_cpp.mkdir(); o_py.mkdir()
        if "-n" in flags:
            for o in (o_cpp, o_py):
                open(str(o / "a1.lna"), "wb").write(b"old")
        r = run(tools["phone_probs"], ["-b", "m", "-c", c["cfg"], "-r", c["recipe"], "-o", str(o_cpp)] + flags)
        assert r.returncode == 0, r.stderr.decode()
        four = "--lnabytes=4" in flags
        pp = PhoneProbs(engine=FakeEngine(), lnabytes=4 if four else 2, normalize=not four)
        n = pp.run_recipe(c["recipe"], str(o_py), **kw)
        names = sorted(os.listdir(str(o_cpp)))
        assert names == sorted(os.listdir(str(o_py))) and n == len(names) - (1 if "-n" in flags else 0)
        for f in names:
            assert open(str(o_cpp / f), "rb").read() == open(str(o_py / f), "rb").read(), (flags, f)


def test_python_pptoolbox_surface(case):
    """aaltoasr_b200.PPToolbox (= PhoneProbs): the method names and arguments of the reference's SWIG class
    (aku/swig/PPToolbox.i:59-66) -- generate(in, out, raw_flag), generate_to_fd, set_clustering -- over the engine."""
    import aaltoasr_b200
    c = case

    class Eng(FakeEngine):
        log = []

        def read_clustering(self, path):
            self.log.append(("read_clustering", path))

        def set_clustering_min_evals(self, a, b):
            self.log.append(("min_evals", a, b))

    eng = Eng()
    pp = aaltoasr_b200.PPToolbox(engine=eng)
    pp.set_clustering("c.gcl", 0.1, 0.2)
    assert eng.log == [("read_clustering", "c.gcl"), ("min_evals", 0.1, 0.2)]
    wav = str(c["dir"] / "a1.wav")
    out = str(c["dir"] / "py.lna")
    assert pp.generate(wav, out, False) == 20 and open(out, "rb").read() == lna_file(range(20))
    raw = str(c["dir"] / "a1.raw")
    formats.read_wav(wav)[0].astype("<i2").tofile(raw)
    assert pp.generate(raw, out, True) == 20 and open(out, "rb").read() == lna_file(range(20))
    fin, fout = os.open(wav, os.O_RDONLY), os.open(str(c["dir"] / "fd.lna"), os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o664)
    assert pp.generate_to_fd(fin, fout, False) == 20
    os.close(fin)
    os.close(fout)
    assert open(str(c["dir"] / "fd.lna"), "rb").read() == lna_file(range(20))
    formats.write_wav(str(c["dir"] / "k8.wav"), np.zeros(640, np.int16), 8000)
    with pytest.raises(aaltoasr_b200.AkuGpuError, match="don't agree"):
        pp.generate(str(c["dir"] / "k8.wav"), out)


def test_raw_and_endian_keys_of_the_audiofile_module(tools, case):
    """`raw 1` and `endian big` of the audiofile module (AudioFileModule::set_module_config, aku/FeatureModules.cc:345-356;
    AudioReader::open, aku/AudioReader.cc:94-108) are the host reader's business: headerless big-endian PCM reaches the
    library as the same samples as the WAV file, through both tools; without the keys the bytes are taken little-endian;
    `raw 1` also stops a RIFF header from being interpreted."""
    c = case
    pcm, _ = formats.read_wav(str(c["dir"] / "a1.wav"))
    wsum = int((pcm.astype(np.int64) * (np.arange(pcm.size) % 97 + 1)).sum())
    be = str(c["dir"] / "be.raw")
    pcm.astype(">i2").tofile(be)
    le = str(c["dir"] / "le.raw")
    pcm.astype("<i2").tofile(le)
    cfg_be = str(c["dir"] / "be.cfg")
    open(cfg_be, "w").write("module\n{\n  name audiofile\n  type audiofile\n  sample_rate 16000\n  endian big\n  raw 1\n}\n"
                            "module\n{\n  name x\n  type stub\n  endian little\n  sources audiofile\n}\n")     # only the base module counts
    want = "pcm n=%d wsum=%d" % (pcm.size, wsum)
    swapped = pcm.astype("<i2").view(">i2").astype(np.int64)
    want_swapped = "pcm n=%d wsum=%d" % (pcm.size, int((swapped * (np.arange(pcm.size) % 97 + 1)).sum()))
    for exe, args in ((tools["feacat"], ["-e", "0"]), ):
        for cfg, path, expect in ((c["cfg"], str(c["dir"] / "a1.wav"), want), (c["cfg"], le, want), (cfg_be, be, want),
                                  (c["cfg"], be, want_swapped)):
            log = str(c["dir"] / "log_fc")
            if os.path.exists(log):
                os.remove(log)
            r = run(exe, ["-c", cfg] + args + [path], log)
            assert r.returncode == 0, r.stderr.decode()
            assert calls(log, "pcm") == [expect], (cfg, path)
    assert formats.config_audio_format(open(cfg_be).read()) == (True, True) and formats.config_audio_format(open(c["cfg"]).read()) == (False, False)
    # `raw 1`: the 44 header bytes of a WAV file are samples too
    log = str(c["dir"] / "log_rawwav")
    r = run(tools["feacat"], ["-c", cfg_be, "-e", "0", str(c["dir"] / "a1.wav")], log)
    assert r.returncode == 0 and calls(log, "pcm")[0].startswith("pcm n=%d " % (pcm.size + 22))
    # the recipe tool: configuration keys, and -R on top of a plain configuration
    rec = str(c["dir"] / "raw.recipe")
    out = c["dir"] / "oraw"
    out.mkdir()
    open(rec, "w").write("audio=%s lna=be.lna\n" % be)
    log = str(c["dir"] / "log_pp")
    r = run(tools["phone_probs"], ["-b", "m", "-c", cfg_be, "-r", rec, "-o", str(out)], log)
    assert r.returncode == 0 and calls(log, "pcm") == [want]
    open(rec, "w").write("audio=%s lna=le.lna\n" % le)
    log = str(c["dir"] / "log_pp2")
    r = run(tools["phone_probs"], ["-b", "m", "-c", c["cfg"], "-r", rec, "-o", str(out), "-R"], log)
    assert r.returncode == 0 and calls(log, "pcm") == [want]
