"""ref.py -- access to the REFERENCE's own code as built by oracle/build_ref.sh into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  `available()` is False when oracle/_ref has not been built (it is
built in the dev container where /root/reference is mounted; the prebuilt files travel to the
GPU box with the snapshot).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_lib = None


def available():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("libref_capi.so", "ref_phone_probs", "ref_feacat"))


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.path.join(REF_DIR, "libref_capi.so"))
        _lib.ref_last_error.restype = C.c_char_p
        _lib.ref_features.restype = C.c_long
        _lib.ref_features.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_long,
                                      C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float)]
        _lib.ref_features_spk.restype = C.c_long
        _lib.ref_features_spk.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_long,
                                          C.POINTER(C.c_int)]
        _lib.ref_module_output.restype = C.c_long
        _lib.ref_module_output.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_long,
                                           C.POINTER(C.c_int)]
        _lib.ref_model_open.restype = C.c_void_p
        _lib.ref_model_open.argtypes = [C.c_char_p]
        _lib.ref_model_close.argtypes = [C.c_void_p]
        for f in ("ref_model_num_states", "ref_model_dim", "ref_model_num_gaussians"):
            getattr(_lib, f).argtypes = [C.c_void_p]
        _lib.ref_state_likelihoods.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p]
        _lib.ref_gaussian_loglik.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p]
        _lib.ref_lna_read.restype = C.c_long
        _lib.ref_lna_read.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_long, C.POINTER(C.c_int), C.c_int]
        _lib.ref_model_read_clustering.argtypes = [C.c_void_p, C.c_char_p]
        _lib.ref_model_set_clustering_min_evals.argtypes = [C.c_void_p, C.c_double, C.c_double]
        _lib.ref_model_set_speaker.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        _lib.ref_recipe_read.restype = C.c_long
        _lib.ref_recipe_read.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_long]
    return _lib


def _err():
    return RuntimeError("reference: " + lib().ref_last_error().decode())


def features(cfg_path, wav_path, start=0, end=-1, max_frames=1 << 20, dim_hint=4096):
    """FeatureGenerator::generate over frames [start,end) (end<0: until eof). float64 [F x dim]."""
    buf = np.empty(max_frames * 64 if end < 0 else (end - start) * dim_hint, dtype=np.float64)
    dim, last, fr = C.c_int(0), C.c_int(0), C.c_float(0)
    cap = max_frames if end < 0 else (end - start)
    # first call to learn dim
    n = lib().ref_features(cfg_path.encode(), wav_path.encode(), start, start + 1 if end >= 0 else -1, None, 0,
                           C.byref(dim), C.byref(last), C.byref(fr))
    if n < 0:
        raise _err()
    buf = np.empty(cap * dim.value, dtype=np.float64)
    n = lib().ref_features(cfg_path.encode(), wav_path.encode(), start, end, buf.ctypes.data, cap, C.byref(dim),
                           C.byref(last), C.byref(fr))
    if n < 0:
        raise _err()
    return buf[:n * dim.value].reshape(n, dim.value).copy(), last.value, fr.value


def features_spk(cfg_path, wav_path, spkc_path, speaker, start=0, end=-1, max_frames=1 << 16):
    """FeatureGenerator::generate with a speaker configuration applied (SpeakerConfig::set_speaker). float64 [F x dim]."""
    dim = C.c_int(0)
    buf = np.empty(max_frames * 128, dtype=np.float64)
    n = lib().ref_features_spk(cfg_path.encode(), wav_path.encode(), spkc_path.encode(), speaker.encode(), start, end,
                               buf.ctypes.data, max_frames, C.byref(dim))
    if n < 0:
        raise _err()
    return buf[:n * dim.value].reshape(n, dim.value).copy()


def module_output(cfg_path, wav_path, module, start, end):
    dim = C.c_int(0)
    n = lib().ref_module_output(cfg_path.encode(), wav_path.encode(), module.encode(), start, start, None, 0, C.byref(dim))
    if n < 0:
        raise _err()
    buf = np.empty((end - start) * dim.value, dtype=np.float64)
    n = lib().ref_module_output(cfg_path.encode(), wav_path.encode(), module.encode(), start, end, buf.ctypes.data,
                                end - start, C.byref(dim))
    if n < 0:
        raise _err()
    return buf.reshape(n, dim.value)


class Model:
    """aku::HmmSet read with read_all(base)."""

    def __init__(self, base):
        self.h = lib().ref_model_open(base.encode())
        if not self.h:
            raise _err()
        self.S = lib().ref_model_num_states(self.h)
        self.D = lib().ref_model_dim(self.h)
        self.G = lib().ref_model_num_gaussians(self.h)

    def state_likelihoods(self, feats):
        feats = np.ascontiguousarray(feats, dtype=np.float64)
        out = np.empty((feats.shape[0], self.S))
        if lib().ref_state_likelihoods(self.h, feats.ctypes.data, feats.shape[0], feats.shape[1], out.ctypes.data):
            raise _err()
        return out

    def read_clustering(self, path):
        if lib().ref_model_read_clustering(self.h, path.encode()):
            raise _err()

    def set_clustering_min_evals(self, min_clusters, min_gaussians):
        if lib().ref_model_set_clustering_min_evals(self.h, float(min_clusters), float(min_gaussians)):
            raise _err()

    def set_speaker(self, spkc_path, speaker):
        """aku::SpeakerConfig(gen, &model).set_speaker: loads the speaker's `model` transformations (the file is read
        on the first call)."""
        if lib().ref_model_set_speaker(self.h, spkc_path.encode(), speaker.encode()):
            raise _err()

    def gaussian_loglik(self, feats):
        feats = np.ascontiguousarray(feats, dtype=np.float64)
        out = np.empty((feats.shape[0], self.G))
        if lib().ref_gaussian_loglik(self.h, feats.ctypes.data, feats.shape[0], feats.shape[1], out.ctypes.data):
            raise _err()
        return out

    def close(self):
        if self.h:
            lib().ref_model_close(self.h)
            self.h = None


def recipe_read(path, num_batches=0, batch_index=0, sort=False):
    """aku::Recipe::read as phone_probs calls it (+ sort_infos): list of (audio, lna, speaker, utterance, start, end)."""
    buf = C.create_string_buffer(1 << 20)
    n = lib().ref_recipe_read(path.encode(), num_batches, batch_index, 1 if sort else 0, buf, len(buf))
    if n < 0:
        raise _err()
    rows = [ln.split("|") for ln in buf.value.decode().splitlines()]
    assert len(rows) == n
    return [(r[0], r[1], r[2], r[3], float(r[4]), float(r[5])) for r in rows]


def lna_read(path, buf_frames=64, backwards=True):
    """The decoder's LnaReaderCircular on an LNA file: float32 log-probs [frames x models] as the decoder sees them."""
    S = C.c_int(0)
    n = lib().ref_lna_read(path.encode(), buf_frames, None, 0, C.byref(S), 0)
    if n < 0:
        raise _err()
    out = np.empty((n, S.value), dtype=np.float32)
    n2 = lib().ref_lna_read(path.encode(), buf_frames, out.ctypes.data, n, C.byref(S), 1 if backwards else 0)
    if n2 != n:
        raise _err() if n2 < 0 else RuntimeError("frame count changed between reads")
    return out


def phone_probs(cfg_path, model_base, recipe_path, out_dir, lnabytes=2, extra=(), timeout=3600):
    """Runs the literal aku/phone_probs.cc binary.  Returns wall seconds."""
    import time
    cmd = [os.path.join(REF_DIR, "ref_phone_probs"), "-b", model_base, "-c", cfg_path, "-r", recipe_path,
           "-o", out_dir, "--lnabytes=%d" % lnabytes] + list(extra)
    t0 = time.time()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
    dt = time.time() - t0
    if p.returncode != 0:
        raise RuntimeError("ref_phone_probs failed (%d): %s" % (p.returncode, p.stderr.decode()[-2000:]))
    return dt
