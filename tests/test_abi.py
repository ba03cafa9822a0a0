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
