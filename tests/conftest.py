import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    if "cfg" in d:
        d["cfg"] = str(d["cfg"])
    d["model"] = {k[6:]: d[k] for k in list(d) if k.startswith("model_")}
    return d


@pytest.fixture(scope="session")
def ref_small():
    return load_golden("ref_small")


@pytest.fixture(scope="session")
def ref_edge():
    return load_golden("ref_edge")


@pytest.fixture(scope="session")
def ref_full():
    return load_golden("ref_full")


@pytest.fixture(scope="session")
def ref_clust():
    d = load_golden("ref_clust")
    d["gcl"] = str(d["gcl"])
    return d


@pytest.fixture(scope="session")
def ref_spk():
    d = load_golden("ref_spk")
    d["spkc"] = str(d["spkc"])
    d["speakers"] = [str(s) for s in d["speakers"]]
    return d


@pytest.fixture(scope="session")
def ref_cmllr():
    d = load_golden("ref_cmllr")
    d["spkc"] = str(d["spkc"])
    d["gcl"] = str(d["gcl"])
    d["speakers"] = [str(s) for s in d["speakers"]]
    return d


@pytest.fixture(scope="session")
def ref_cmllr_units():
    z = np.load(os.path.join(GOLDEN, "ref_cmllr_units.npz"))
    d = {k: z[k] for k in z.files}
    for k in ("cfg", "spkc", "ph", "unitmode_phone", "unitmode_mix", "unitmode_gauss"):
        d[k] = str(d[k])
    d["model"] = {k[6:]: d[k] for k in list(d) if k.startswith("model_")}
    return d


@pytest.fixture(scope="session")
def ref_pre():
    return load_golden("ref_pre")


@pytest.fixture(scope="session")
def ref_vtln():
    z = np.load(os.path.join(GOLDEN, "ref_vtln.npz"))
    return {k: (str(z[k]) if k.startswith(("cfg_", "spkc_")) else z[k]) for k in z.files}


@pytest.fixture(scope="session")
def ref_modx():
    z = np.load(os.path.join(GOLDEN, "ref_modx.npz"))
    return {k: (str(z[k]) if k.startswith(("cfg_", "spkc_")) else z[k]) for k in z.files}


@pytest.fixture(scope="session")
def aku_tests():
    z = np.load(os.path.join(GOLDEN, "aku_tests.npz"))
    return {k: (str(z[k]) if k.endswith("_cfg") else z[k]) for k in z.files}


@pytest.fixture(scope="session")
def engine():
    """One AkuGpu context for the GPU tests.  Fails loudly (no skip, no fallback) if the CUDA
    library is missing or no device is visible."""
    from aaltoasr_b200 import AkuGpu
    eng = AkuGpu(0)
    yield eng
    eng.close()
