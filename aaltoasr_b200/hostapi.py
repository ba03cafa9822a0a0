"""Host-side mirror of the reference's call surface for the accelerated path.

Same names, argument meaning and error behaviour as the reference classes, so that code
(and tests) written against them read the same:

  FeatureGenerator  aku/FeatureGenerator.hh:23-123   load_configuration / open / generate(frame) /
                                                      eof / last_frame / dim / frame_rate / sample_rate
  HmmSet            aku/HmmSet.hh:94-571              read_all / reset_cache / precompute_likelihoods /
                                                      state_likelihood / num_states / dim
  PhoneProbs        aku/phone_probs.cc:46-267,        the tool's loop: recipe -> LNA files
                    aku/PhoneProbsToolbox.cc:135-222  (PPToolbox: read_configuration/read_models/generate)

The reference pulls one frame at a time through ring buffers and scores every Gaussian for
that one frame.  Here the per-frame methods are served from whole-utterance results computed
on the GPU (features once per open(), likelihoods once per utterance), which is what lets
the same API run at B200 speed.  Errors surface as AkuGpuError (a RuntimeError), as the
reference's SWIG layer turns `throw std::string` into RuntimeError.
"""
import os

import numpy as np

from . import formats
from ._lib import AkuGpuError
from .engine import AkuGpu, F32, F64


class FeatureGenerator:
    def __init__(self, engine=None, device=0):
        self.engine = engine if engine is not None else AkuGpu(device)
        self._pcm = None
        self._feats = None
        self._n = 0
        self._eof_on_last_frame = False

    def load_configuration(self, path_or_file):
        if hasattr(path_or_file, "read"):
            self.engine.frontend_load_config_text(path_or_file.read())
        else:
            self.engine.frontend_load_config(path_or_file)

    def load_configuration_text(self, text):
        self.engine.frontend_load_config_text(text)

    def open(self, filename):
        pcm, sr = formats.read_wav(filename)
        if sr != self.engine.sample_rate:
            # aku/FeatureModules.cc:254-261
            raise AkuGpuError(-2, "Audio file sample rate (%d Hz) and model configuration (%d Hz) don't agree."
                              % (sr, self.engine.sample_rate))
        self.open_pcm(pcm)

    def open_pcm(self, pcm):
        self._pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        self._n = self.engine.num_frames(self._pcm.size)
        if self._n <= 0:
            raise AkuGpuError(-2, "audio shorter than frame")   # aku/FeatureModules.cc:409
        self._feats, _ = self.engine.features(self._pcm, dtype=np.float64)
        self._eof_on_last_frame = False

    def close(self):
        self._pcm = self._feats = None

    def generate(self, frame):
        """Feature vector of `frame` (any integer; frames outside the file behave as in the
        reference: first/last window replicated at the base module)."""
        if self._pcm is None:
            raise AkuGpuError(-5, "no audio opened")
        self._eof_on_last_frame = frame >= self._n
        if 0 <= frame < self._n:
            return self._feats[frame]
        return self.engine.features_range(self._pcm, frame, frame + 1)[0]

    def features(self):
        """All frames 0..last_frame() at once, [F x dim] float64."""
        return self._feats

    def eof(self):
        return self._eof_on_last_frame

    def last_frame(self):
        return self._n - 1

    def dim(self):
        return self.engine.feature_dim

    def frame_rate(self):
        return self.engine.frame_rate

    def sample_rate(self):
        return self.engine.sample_rate

    def module_output(self, name, start, end):
        return self.engine.features_range(self._pcm, start, end, module=name)


class HmmSet:
    def __init__(self, engine=None, device=0, precision=F64):
        self.engine = engine if engine is not None else AkuGpu(device)
        self.precision = precision
        self._lik = None
        self._row = None
        self._utt = None

    def read_all(self, base):
        self.engine.model_read(base)

    def num_states(self):
        return self.engine.num_states

    def dim(self):
        return self.engine.model_dim

    # Gaussian clustering approximation (aku/HmmSet.cc:1354-1366)
    def read_clustering(self, filename):
        self.engine.read_clustering(filename)

    def set_clustering_min_evals(self, min_clusters=1.0, min_gaussians=1.0):
        self.engine.set_clustering_min_evals(min_clusters, min_gaussians)

    # Whole-utterance entry: score every frame once, then serve the per-frame API from it.
    def set_utterance_features(self, feats):
        feats = np.ascontiguousarray(feats)
        self._utt = feats
        out = self.engine.gmm_score(feats, precision=self.precision)
        self._lik = out if self.precision == F64 else np.maximum(np.exp(out.astype(np.float64)), 1e-50)

    def reset_cache(self):
        self._row = None

    def precompute_likelihoods(self, feature):
        """feature: a frame index into the utterance given to set_utterance_features(), or a
        feature vector (scored on the spot: the F=1 case of the same kernel)."""
        if isinstance(feature, (int, np.integer)):
            self._row = self._lik[int(feature)]
        else:
            v = np.ascontiguousarray(feature, dtype=np.float64).reshape(1, -1)
            out = self.engine.gmm_score(v, precision=self.precision)[0]
            self._row = out if self.precision == F64 else np.maximum(np.exp(out.astype(np.float64)), 1e-50)

    def state_likelihood(self, state, feature=None):
        if self._row is None:
            self.precompute_likelihoods(feature)
        return float(self._row[state])


class StreamSession:
    """The stream decoder's session (decoder/decode-stream.cc:150-207: one acoustic model for the life of the stream, one
    scoring call per frame).  While the object is open the model sits in the shared memory of the SMs
    (akugpu_stream_open) and log_probs() is a message to the resident kernel; the returned rows are a VIEW of the
    context's pinned memory, valid until its next call.  Mirror of akugpu::StreamSession (csrc/host/akugpu.hh)."""

    def __init__(self, engine, tiny=1e-30, idle_ms=100.0):
        self.engine, self.tiny = engine, tiny
        engine.stream_open(idle_ms)
        self._open = True

    def log_probs(self, feats):
        x = np.ascontiguousarray(feats, dtype=np.float32)
        return self.engine.stream_logprobs(x.reshape(1, -1) if x.ndim == 1 else x, tiny=self.tiny)

    def close(self):
        if self._open:
            self._open = False
            self.engine.stream_close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def parse_speaker_file(text):
    """The .spkc format of aku::SpeakerConfig::read_speaker_file (aku/SpeakerConfig.cc:20-147):
    `speaker|utterance <id|default>` / `{` / `[feature] <module>` / `{ key value ... }` / `}`.
    Returns {"speaker": {id: {module: text}}, "utterance": {...}} with the parameter text as `key value` lines."""
    lines = [ln.strip() for ln in text.splitlines()]
    pos = 0

    def nxt():
        nonlocal pos
        while pos < len(lines):
            ln = lines[pos]
            pos += 1
            if ln:
                return ln
        return None

    out = {"speaker": {}, "utterance": {}}
    while True:
        ln = nxt()
        if ln is None:
            break
        f = ln.split()
        if len(f) != 2 or f[0] not in ("speaker", "utterance"):
            raise AkuGpuError(-2, "SpeakerConfig: Syntax error on line %d: %s" % (pos, ln))
        if f[1] in out[f[0]] and f[1] == "default":
            raise AkuGpuError(-2, "SpeakerConfig: Default %s configuration already defined, redefinition on line %d: %s" % (f[0], pos, ln))
        mods = out[f[0]].setdefault(f[1], {})
        if nxt() != "{":
            raise AkuGpuError(-2, "'{' expected in speaker config file")
        while True:
            ln = nxt()
            if ln is None or ln == "}":
                break
            parts = ln.split(None, 1)
            ns, name = ("feature", ln) if len(parts) < 2 else (parts[0], parts[1])
            if ns not in ("model", "feature"):
                raise AkuGpuError(-2, "SpeakerConfig: Unknown module namespace at line %d" % pos)
            if ns == "model":
                # ModelTransformer::get_new_module knows one module (aku/ModelModules.cc:12-18)
                if name != "cmllr":
                    raise AkuGpuError(-2, "SpeakerConfig: error on line %d: unknown model module requested: %s" % (pos, name))
                if f[0] == "utterance":
                    raise AkuGpuError(-2, "SpeakerConfig: error on line %d: model modules are loaded per speaker "
                                      "(aku/SpeakerConfig.cc:248-284); an utterance-level entry is not supported" % pos)
            if nxt() != "{":
                raise AkuGpuError(-2, "SpeakerConfig: Failed reading module parameters around line %d: '{' expected in module config file" % pos)
            body = []
            while True:
                ln = nxt()
                if ln is None:
                    raise AkuGpuError(-2, "SpeakerConfig: Failed reading module parameters around line %d: unexpected end of module config file" % pos)
                if ln == "}":
                    break
                body.append(ln)
            mods[name if ns == "feature" else "model " + name] = "\n".join(body) + ("\n" if body else "")
    return out


def parse_cmllr_parameters(text, dim):
    """The global transform of a `model cmllr` entry (unitmode UNIT_NO): W [dim x (dim+1)] or None (no `w` entry)."""
    um, trs = parse_cmllr_transforms(text, dim)
    if um != "UNIT_NO":
        raise AkuGpuError(-2, "cmllr: unitmode %s holds regression-class transforms: parse_cmllr_transforms" % um)
    return trs[0][1] if trs else None


def parse_cmllr_transforms(text, dim):
    """ConstrainedMllr::set_parameters (aku/ModelModules.cc:62-95): `unitmode UNIT_NO | UNIT_PHONE | UNIT_MIX | UNIT_GAUSSIAN`
    and entries `w<i> [units...] <dim*(dim+1) numbers>` (row-major [dim x (dim+1)], column 0 = bias; regression-class
    modes need at least one unit in front of the matrix).  Returns (unitmode, [(units, W), ...]) in the order w1, w2, ..."""
    params = {}
    for ln in text.splitlines():
        f = ln.split()
        if f:
            params[f[0]] = f[1:]
    um = params.get("unitmode", ["UNIT_NO"])
    um = um[0] if um else "UNIT_NO"
    n = dim * (dim + 1)
    need = n if um == "UNIT_NO" else n + 1
    found = {}
    i = 1
    while "w%d" % i in params:
        parts = params["w%d" % i]
        if len(parts) < need:
            raise AkuGpuError(-2, "ERROR: not enough elements for matrix w%d" % i)
        vals = []
        for t in parts[len(parts) - n:]:
            try:
                vals.append(float(np.float32(float(t))))      # str::str2float returns through a float (aku/str.cc:261-282)
            except ValueError:
                raise AkuGpuError(-2, "invalid value: " + t)
        found[tuple(parts[:len(parts) - n])] = np.array(vals, dtype=np.float64).reshape(dim, dim + 1)     # same units: one map entry
        i += 1
    if um == "UNIT_NO" and len(found) > 1:
        raise AkuGpuError(-2, "ERROR: speaker can only contain one transform when UNIT_NO (global transform) is set")
    return um, [(list(k), w) for k, w in found.items()]


class SpeakerConfig:
    """aku::SpeakerConfig (aku/SpeakerConfig.cc:239-336): set_speaker / set_utterance push the stored parameters of
    feature modules through FeatureModule::set_parameters (akugpu_frontend_set_parameters) and a speaker's global
    `model cmllr` transform through akugpu_model_set_cmllr (the model must be loaded first, as in the reference)."""

    def __init__(self, engine):
        self.engine = engine
        self.conf = {"speaker": {}, "utterance": {}}
        self.cur_speaker = ""
        self.cur_utterance = ""

    def read_speaker_file(self, path):
        self.conf = parse_speaker_file(open(path).read())

    def _apply(self, kind, ident, what):
        table = self.conf[kind]
        if not ident:
            if "default" not in table:
                raise AkuGpuError(-2, "SpeakerConfig: No speaker defined, needs a default speaker." if kind == "speaker"
                                  else "SpeakerConfig: Default utterance is required.")
            mods = table["default"]
        elif ident in table and ident != "default":
            mods = table[ident]
        else:
            if "default" not in table:
                raise AkuGpuError(-2, "SpeakerConfig: Unknown %s %s, and default %s settings are missing." % (what, ident, what))
            mods = table["default"]
        for name in sorted(mods):
            if name == "model cmllr":     # model namespace: ModelTransformer (aku/SpeakerConfig.cc:248-284)
                um, trs = parse_cmllr_transforms(mods[name], self.engine.model_dim)
                if um == "UNIT_NO":
                    self.engine.model_set_cmllr(trs[0][1] if trs else None)
                else:                     # regression classes: phones / mixtures / Gaussians listed in front of each matrix
                    self.engine.model_set_cmllr_units(um, trs)
            else:
                self.engine.frontend_set_parameters(name, mods[name])

    def set_speaker(self, speaker_id):
        if self.cur_utterance:
            self.set_utterance("")
        self._apply("speaker", speaker_id, "speaker")
        self.cur_speaker = speaker_id

    def set_utterance(self, utterance_id):
        self._apply("utterance", utterance_id, "utterance")
        self.cur_utterance = utterance_id


class PhoneProbs:
    """aku/phone_probs as an object (and the PPToolbox method names)."""

    def __init__(self, engine=None, device=0, precision=F32, lnabytes=2, normalize=True):
        self.engine = engine if engine is not None else AkuGpu(device)
        self.precision = precision
        self.lnabytes = lnabytes
        self.normalize = normalize

    # PPToolbox names (aku/swig/PPToolbox.i:59-66)
    def read_configuration(self, cfg):
        self.engine.frontend_load_config(cfg)
        self._cfg_raw, self._cfg_big_endian = formats.config_audio_format(open(cfg).read())

    def read_models(self, base):
        self.engine.model_read(base)

    def set_clustering(self, clfile_name, eval_minc, eval_ming):
        """PPToolbox::set_clustering (aku/PhoneProbsToolbox.cc:50-53)."""
        self.engine.read_clustering(clfile_name)
        self.engine.set_clustering_min_evals(eval_minc, eval_ming)

    def _pcm_of(self, blob, raw_flag, what):
        raw_flag = raw_flag or getattr(self, "_cfg_raw", False)
        if not raw_flag and blob[:4] == b"RIFF" and blob[8:12] == b"WAVE":
            pcm, sr = formats.parse_wav(blob, what)
        else:                                   # headerless PCM16 at the configured rate, like AudioReader's fallback
            dt = ">i2" if getattr(self, "_cfg_big_endian", False) else "<i2"
            pcm, sr = np.frombuffer(blob[:len(blob) // 2 * 2], dtype=dt).astype(np.int16), self.engine.sample_rate
        if sr != self.engine.sample_rate:
            raise AkuGpuError(-2, "Audio file sample rate (%d Hz) and model configuration (%d Hz) don't agree."
                              % (sr, self.engine.sample_rate))
        return pcm

    def _stream(self, pcm):
        rec, _, _ = self.engine.phone_probs(pcm, precision=self.precision, lnabytes=self.lnabytes,
                                            normalize=self.normalize)
        return rec

    def generate(self, audio_path, lna_path, raw_flag=False):
        """PPToolbox::generate(input_name, output_name, raw_flag) (aku/PhoneProbsToolbox.cc:211-222)."""
        rec = self._stream(self._pcm_of(open(audio_path, "rb").read(), raw_flag, audio_path))
        formats.write_lna(lna_path, rec, self.engine.num_states, self.lnabytes)
        return rec.shape[0]

    def generate_to_fd(self, in_fd, out_fd, raw_flag=False):
        """PPToolbox::generate_to_fd (aku/PhoneProbsToolbox.cc:55-133): audio from an open descriptor (read to its end),
        header + records to out_fd; neither descriptor is closed."""
        chunks = []
        while True:
            b = os.read(in_fd, 1 << 20)
            if not b:
                break
            chunks.append(b)
        rec = self._stream(self._pcm_of(b"".join(chunks), raw_flag, "<fd>"))
        data = formats.lna_header(self.engine.num_states, self.lnabytes) + np.ascontiguousarray(rec, dtype=np.uint8).tobytes()
        done = 0
        while done < len(data):
            done += os.write(out_fd, data[done:])
        return rec.shape[0]

    def run_recipe(self, recipe_path, out_dir="", batch=1, bindex=1, no_overwrite=False, audio_ext_lna=False,
                   max_batch_samples=64 << 20, sort_recipe=False):
        """The tool's recipe loop (aku/phone_probs.cc:135-267): utterances of this batch share GPU
        launches; output files are per utterance like the reference's."""
        infos = formats.read_recipe(recipe_path, batch, bindex)
        if sort_recipe:     # --sort-recipe (aku/phone_probs.cc:141-142)
            infos = formats.sort_recipe(infos)
        todo = []
        for info in infos:
            if audio_ext_lna:           # aku/phone_probs.cc:158-176: extension stripped only at a dot past position 0
                name = info["audio"]
                cut = name.rfind("/")
                if 0 <= cut < len(name) - 1:
                    name = name[cut + 1:]
                dot = name.rfind(".")
                name = (name[:dot] if dot > 0 else name) + ".lna"
            else:
                name = info.get("lna", "")
            path = os.path.join(out_dir, name) if out_dir else name
            if no_overwrite and os.path.exists(path):      # stat() == 0, empty files included (aku/phone_probs.cc:180-190)
                import sys
                sys.stderr.write("WARNING: skipping existing lna file %s\n" % path)
                continue
            todo.append((info, path))
        written = 0
        i = 0
        while i < len(todo):
            pcms, j, tot = [], i, 0
            while j < len(todo) and (j == i or tot < max_batch_samples):
                pcm, sr = formats.read_wav(todo[j][0]["audio"])
                if sr != self.engine.sample_rate:
                    raise AkuGpuError(-2, "Audio file sample rate (%d Hz) and model configuration (%d Hz) don't agree."
                                      % (sr, self.engine.sample_rate))
                pcms.append(pcm)
                tot += pcm.size
                j += 1
            uo = np.zeros(len(pcms) + 1, dtype=np.int64)
            uo[1:] = np.cumsum([p.size for p in pcms])
            rec, fo, _ = self.engine.phone_probs(np.concatenate(pcms), uo, precision=self.precision,
                                                 lnabytes=self.lnabytes, normalize=self.normalize)
            fr = self.engine.frame_rate
            for k in range(len(pcms)):
                info, path = todo[i + k]
                a, b = int(fo[k]), int(fo[k + 1])
                s = int(float(info.get("start-time", 0) or 0) * fr)          # aku/phone_probs.cc:199-206
                e = int(float(info.get("end-time", 0) or 0) * fr)
                if e == 0 or e > b - a:
                    e = b - a
                s = min(s, e)
                part = rec[a + max(s, 0):a + max(e, 0)]
                if s < 0:       # frames before the file: the reference's loop generates them (first window replicated)
                    fneg = self.engine.features_range(pcms[k], s, min(e, 0), dtype=np.float64)
                    rneg = self.engine.gmm_lna(fneg, precision=self.precision, lnabytes=self.lnabytes, normalize=self.normalize)
                    part = np.concatenate([rneg.reshape(-1, rec.shape[1]), part])
                formats.write_lna(path, part, self.engine.num_states, self.lnabytes)
                written += 1
            i = j
        return written
