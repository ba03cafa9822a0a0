"""oracle_np.py -- numpy restatement of the reference's hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product may import this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and
only as the checker.

Every function restates one reference routine and cites it (paths relative to the AaltoASR
tree).  The restatement is pinned against the reference itself: tests/test_oracle.py checks
it against the aku/tests golden vectors (via tests/golden/aku_tests.npz) and against outputs
of the reference's own code built into oracle/_ref (features to <=1e-5 absolute -- only the
FFT's internal rounding order differs -- and the GMM+LNA stage bit for bit).

Float/double placement follows the reference exactly: float
pre-emphasis, float Hamming table (cosf), float32 FFT (the KissFFT build), sqrtf/logf,
float mel accumulators, cosf DCT basis with double accumulation, float power accumulator,
double deltas; double Gaussians; float/double hybrid LNA normalisation.
"""
import math

import numpy as np

f32 = np.float32
# libm exp/log element by element: numpy's SIMD exp/log differ from glibc's in the last bit,
# and the reference links glibc.
_exp = np.frompyfunc(lambda v: math.exp(v) if v > -745.2 else 0.0, 1, 1)
_log = np.frompyfunc(math.log, 1, 1)


# ------------------------------------------------------------------------------------
# feature configuration  (FeatureGenerator::load_configuration, aku/FeatureGenerator.cc:97-219;
#                         ModuleConfig::read, aku/ModuleConfig.cc:166-203)
def parse_config(text):
    mods, cur, state = [], None, 0
    for raw in text.splitlines():
        line = raw.strip(" \t\r\n")
        if not line:
            continue
        if state == 0:
            if line != "module":
                raise ValueError("expected keyword 'module': " + line)
            cur, state = {}, 1
        elif state == 1:
            if line != "{":
                raise ValueError("'{' expected in module config file: " + line)
            state = 2
        elif line == "}":
            mods.append(cur)
            state = 0
        else:
            parts = line.split(None, 1)
            if len(parts) < 2:
                raise ValueError("value missing for option: " + line)
            if parts[0] in cur:
                raise ValueError("value redefined: " + line)
            cur[parts[0]] = parts[1].strip()
    if state != 0:
        raise ValueError("unexpected end of module config file")
    return mods


def _fvec(s):
    return np.array([f32(float(x)) for x in s.split()], dtype=f32)


class Pipeline:
    """The module chain evaluated for whole utterances; frame indices follow the reference
    (negative / past-EOF frames replicate the first / last window at the base module only)."""

    def __init__(self, cfg_text, fft_dtype=np.float32):
        self.mods = parse_config(cfg_text)
        self.fft_dtype = fft_dtype
        self.by_name = {}
        for i, m in enumerate(self.mods):
            m["sources"] = m.get("sources", "").split()
            self.by_name[m["name"]] = i
        self._setup()

    # set_module_config of every module type
    def _setup(self):
        for m in self.mods:
            t = m["type"]
            src = [self.mods[self.by_name[s]] for s in m["sources"]]
            sdim = src[-1]["dim"] if src else 0
            m["ctx_l"] = m["ctx_r"] = 0
            if t == "pre":         # PreModule::set_module_config :671-689: stored float32 feature rows instead of audio
                m["dim"] = int(m["dim"])
                m["sr"] = int(m.get("sample_rate", 16000))
                m["frate"] = f32(float(m.get("frame_rate", 125)))
            elif t == "audiofile":   # aku/FeatureModules.cc:328-360
                sr = int(m["sample_rate"])
                m["sr"] = sr
                m["emph"] = f32(float(m.get("pre_emph_coef", 0.97)))
                m["frate"] = f32(float(m.get("frame_rate", 125)))
                m["adv"] = f32(f32(sr) / m["frate"])
                m["W"] = int(m["window_width"]) if "window_width" in m else int(f32(2 * sr) / m["frate"])
                m["copy_borders"] = int(m.get("copy_borders", 1))
                m["dim"] = m["W"]
            elif t == "fft":       # :476-518
                N = sdim
                m["magnitude"] = int(m.get("magnitude", 1))
                m["log"] = int(m.get("log", 0))
                m["dim"] = N // 2 + 1
                m["window"] = np.array(
                    [f32(.54 - .46 * float(_cosf(2 * math.pi * i / (N - 1.0)))) for i in range(N)], dtype=f32)
            elif t == "mel":       # :776-802
                sr = self.mods[0]["sr"]
                m["root"] = int(m.get("root", 0))
                num = f32(f32(21 + 2) * _log10f(1 + sr / 1400.0))
                m["dim"] = int(f32(f32(num / _log10f(1 + 16000 / 1400.0)) - f32(2)))
                edges = m["dim"] + 2
                rate = f32(sr)
                mel_step = f32(f32(f32(2595) * _log10f(1.0 + float(rate) / 1400.0)) / f32(edges))
                m["edges"] = np.array(
                    [f32(1400.0 * (math.pow(10, float(f32(f32(f32(i + 1) * mel_step) / f32(2595)))) - 1) * (sdim - 1)
                         / float(rate)) for i in range(edges)], dtype=f32)
            elif t in ("power", "mel_power"):
                m["dim"] = 1
            elif t == "dct":       # :938-979
                m["dim"] = int(m.get("dim", 12))
                m["zeroth"] = int(m.get("zeroth", 0))
                bias = 1 if m["zeroth"] else 0
                tab = np.ones((m["dim"], sdim), dtype=f32)
                for i in range(m["dim"] - bias):
                    for b in range(sdim):
                        tab[i + bias, b] = _cosf((i + 1) * (b + 0.5) * math.pi / sdim)
                m["table"] = tab
            elif t == "delta":     # :999-1016
                m["dim"] = sdim
                w = int(m.get("width", 2))
                m["width"] = w
                m["norm"] = f32(float(m["normalization"])) if "normalization" in m else f32(2 * w * (w + 1) * (2 * w + 1) // 6)
                m["ctx_l"] = m["ctx_r"] = w
            elif t == "merge":
                m["dim"] = sum(s["dim"] for s in src)
            elif t == "concat":    # :1473-1485
                m["ctx_l"], m["ctx_r"] = int(m.get("left", 0)), int(m.get("right", 0))
                m["dim"] = sdim * (1 + m["ctx_l"] + m["ctx_r"])
            elif t == "normalization":   # :1057-1080
                m["dim"] = sdim
                m["mean_v"] = _fvec(m["mean"]) if "mean" in m else np.zeros(sdim, f32)
                if "var" in m:
                    m["scale_v"] = np.array([f32(f32(1) / np.sqrt(v, dtype=f32)) for v in _fvec(m["var"])], dtype=f32)
                elif "scale" in m:
                    m["scale_v"] = _fvec(m["scale"])
                else:
                    m["scale_v"] = np.ones(sdim, f32)
            elif t == "lin_transform":   # :1167-1241
                m["dim"] = int(m.get("dim", sdim))
                m["matrix_v"] = _fvec(m["matrix"]).reshape(m["dim"], sdim) if "matrix" in m else None
                m["bias_v"] = _fvec(m["bias"]) if "bias" in m else None
            elif t == "mean_subtractor":  # :1385-1405
                m["dim"] = sdim
                m["ctx_l"], m["ctx_r"] = int(m.get("left", 75)), int(m.get("right", 75))
            elif t == "vtln":      # VtlnModule::set_module_config :1530-1573
                m["dim"] = sdim
                m["use_pwlin"] = int(m.get("pwlin_vtln", 0))
                m["turn"] = f32(float(m.get("pwlin_turnpoint", 0.8)))
                m["use_slapt"] = int(m.get("slapt", 0))
                m["rad"] = int(m.get("sinc_interpolation_rad", 8))
                m["all_pass"] = int(m.get("all-pass", 0))
                if m["use_pwlin"] and m["all_pass"]:
                    raise ValueError("VtlnModule: Can not use both pwlin_vtln and all-pass!")
                m["lanczos"] = int(m.get("lanczos_window", 0 if m["all_pass"] else 1)) > 0
                if m["lanczos"] and m["all_pass"]:
                    raise ValueError("VtlnModule: Can not use both lanczos_window and all-pass!")
                m["warp"] = f32(1.0)
                m["slapt_params"] = [f32(0.0)]
                _vtln_tables(m)
            elif t == "sr_norm":   # SRNormModule::set_module_config :1954-1989
                m["in_frames"], m["out_frames"] = int(m["in_frames"]), int(m["out_frames"])
                m["frame_dim"] = sdim // m["in_frames"]
                m["dim"] = m["out_frames"] * m["frame_dim"]
                m["order"] = int(m.get("lanczos_order", 4))
                m["rate"] = f32(float(m.get("speech_rate", 1.0)))
                _srnorm_tables(m)
            elif t == "quanteq":   # QuantEqModule::set_module_config :2086-2093
                m["dim"] = sdim
                m["alpha"] = m["gamma"] = m["qmax"] = None
            else:
                raise ValueError("Unknown module type '%s'" % t)

    def set_parameters(self, name, text):
        """FeatureModule::set_parameters (aku/FeatureModule.hh:107) of one module; text = `key value...` lines."""
        m = self.mods[self.by_name[name]]
        kv = {}
        for ln in text.splitlines():
            p = ln.split(None, 1)
            if len(p) == 2:
                kv[p[0]] = p[1]
        t = m["type"]
        if t == "vtln":            # :1575-1592
            if m["use_slapt"]:
                m["slapt_params"] = list(_fvec(kv["slapt_coef"])) if "slapt_coef" in kv else [f32(0.0)]
            else:
                m["warp"] = f32(float(kv.get("warp_factor", 1.0)))
            _vtln_tables(m)
        elif t == "sr_norm":       # :1991-1996
            m["rate"] = f32(float(kv.get("speech_rate", 1.0)))
            _srnorm_tables(m)
        elif t == "quanteq":       # :2095-2104
            m["alpha"] = _fvec(kv["alpha"]) if "alpha" in kv else None
            m["gamma"] = _fvec(kv["gamma"]) if "gamma" in kv else None
            m["qmax"] = _fvec(kv["quant_max"]) if "quant_max" in kv else None
        elif t == "lin_transform":  # :1187-1195
            sdim = self.mods[self.by_name[m["sources"][-1]]]["dim"]
            m["matrix_v"] = _fvec(kv["matrix"]).reshape(m["dim"], sdim) if "matrix" in kv else None
            m["bias_v"] = _fvec(kv["bias"]) if "bias" in kv else None
        elif t == "normalization":
            sdim = m["dim"]
            m["mean_v"] = _fvec(kv["mean"]) if "mean" in kv else np.zeros(sdim, f32)
            if "var" in kv:
                m["scale_v"] = np.array([f32(f32(1) / np.sqrt(v, dtype=f32)) for v in _fvec(kv["var"])], dtype=f32)
            else:
                m["scale_v"] = _fvec(kv["scale"]) if "scale" in kv else np.ones(sdim, f32)

    @property
    def dim(self):
        return self.mods[-1]["dim"]

    def num_frames(self, n_samples):
        """Frames before eof(): frame f is whole iff (int)(f*adv)+W+1 <= N (aku/FeatureModules.cc:399-404)."""
        a = self.mods[0]
        f = 0
        while int(f32(f32(f) * a["adv"])) + a["W"] + 1 <= n_samples:
            f += 1
        return f

    def run(self, pcm, start=0, end=None, module=None):
        """Output of `module` (default: last) for frames [start,end) (default end: first eof frame)."""
        pcm = np.asarray(pcm, dtype=np.int16)
        n = self.num_frames(pcm.size)
        if n <= 0:
            raise ValueError("audio shorter than frame")
        if end is None:
            end = n
        tgt = self.by_name[module] if module else len(self.mods) - 1
        # frame range each module is needed on (context propagated from the target back to the base),
        # so that every module is evaluated exactly once
        need = {tgt: [start, end]}
        for mi in range(tgt, -1, -1):
            if mi not in need:
                continue
            m = self.mods[mi]
            for sname in m["sources"]:
                si = self.by_name[sname]
                lo, hi = need[mi][0] - m["ctx_l"], need[mi][1] + m["ctx_r"]
                if si in need:
                    need[si] = [min(need[si][0], lo), max(need[si][1], hi)]
                else:
                    need[si] = [lo, hi]
        cache = {}
        for mi in range(0, tgt + 1):
            if mi in need and self.mods[mi]["type"] != "audiofile":
                lo, hi = need[mi]
                cache[mi] = (lo, self._eval(mi, np.arange(lo, hi), pcm, n, cache))
        lo, arr = cache[tgt]
        return arr[start - lo:end - lo]

    def run_pre(self, rows, start=0, end=None, module=None):
        """The same for a configuration whose base module is `pre` (PreModule::generate :705-755): rows = the stored
        float32 features [n x dim]; frames before / after the file repeat the first / last row."""
        rows = np.asarray(rows, dtype=np.float32)
        n = rows.shape[0]
        if end is None:
            end = n
        tgt = self.by_name[module] if module else len(self.mods) - 1
        need = {tgt: [start, end]}
        for mi in range(tgt, -1, -1):
            if mi not in need:
                continue
            m = self.mods[mi]
            for sname in m["sources"]:
                si = self.by_name[sname]
                lo, hi = need[mi][0] - m["ctx_l"], need[mi][1] + m["ctx_r"]
                need[si] = [min(need[si][0], lo), max(need[si][1], hi)] if si in need else [lo, hi]
        lo, hi = need[0]
        cache = {0: (lo, rows[np.clip(np.arange(lo, hi), 0, n - 1)].astype(np.float64))}
        for mi in range(1, tgt + 1):
            if mi in need:
                lo, hi = need[mi]
                cache[mi] = (lo, self._eval(mi, np.arange(lo, hi), None, n, cache))
        lo, arr = cache[tgt]
        return arr[start - lo:end - lo]

    def _eval(self, mi, frames, pcm, n, cache):
        m = self.mods[mi]
        t = m["type"]
        src = [self.by_name[s] for s in m["sources"]]
        def ev(si, fr):
            lo, arr = cache[si]
            return arr[fr[0] - lo:fr[-1] + 1 - lo]
        if t == "fft":
            return self._spectrum(m, self.mods[src[0]], frames, pcm, n)
        if t == "audiofile":
            raise ValueError("audiofile output is only consumed by fft")
        if t == "mel":
            return _mel(m, ev(src[0], frames))
        if t == "power":       # :875-885  float accumulator, natural log
            x = ev(src[0], frames)
            p = np.zeros(x.shape[0], dtype=f32)
            for i in range(x.shape[1]):
                p = (p.astype(np.float64) + x[:, i]).astype(f32)
            return np.log(p.astype(np.float64) + 1e-10)[:, None]
        if t == "mel_power":   # :908-919
            x = ev(src[0], frames)
            p = np.zeros(x.shape[0], dtype=f32)
            for i in range(x.shape[1]):
                p = (p.astype(np.float64) + np.exp(x[:, i])).astype(f32)
            return np.log(p.astype(np.float64) + 1e-10)[:, None]
        if t == "dct":         # :956-979  double accumulate in b order
            x = ev(src[0], frames)
            out = np.zeros((x.shape[0], m["dim"]))
            for b in range(x.shape[1]):
                out += x[:, b:b + 1] * m["table"][:, b].astype(np.float64)[None, :]
            return out
        if t == "delta":       # :1019-1037
            acc = np.zeros((len(frames), m["dim"]))
            for k in range(1, m["width"] + 1):
                acc += k * (ev(src[0], frames + k) - ev(src[0], frames - k))
            return acc / float(m["norm"])
        if t == "merge":
            return np.concatenate([ev(s, frames) for s in src], axis=1)
        if t == "concat":
            return np.concatenate([ev(src[0], frames + k) for k in range(-m["ctx_l"], m["ctx_r"] + 1)], axis=1)
        if t == "normalization":   # :1136-1142
            x = ev(src[0], frames)
            return (x - m["mean_v"].astype(np.float64)) * m["scale_v"].astype(np.float64)
        if t == "lin_transform":   # :1244-1269
            x = ev(src[0], frames)
            if m["matrix_v"] is not None:
                out = np.zeros((x.shape[0], m["dim"]))
                M = m["matrix_v"].astype(np.float64)
                for j in range(x.shape[1]):
                    out += M[:, j][None, :] * x[:, j:j + 1]
            else:
                out = x[:, :m["dim"]].copy()
            if m["bias_v"] is not None:
                out = out + m["bias_v"].astype(np.float64)
            return out
        if t == "vtln":            # VtlnModule::generate :1906-1934
            x = ev(src[0], frames)
            out = np.zeros_like(x)
            if m["rad"] > 0:
                for b in range(m["dim"]):
                    st, cf = m["sinc_start"][b], m["sinc_coef"][b]
                    acc = np.zeros(x.shape[0])
                    for i, c in enumerate(cf):          # double accumulator, float taps
                        acc = acc + x[:, st + i] * float(c)
                    out[:, b] = np.maximum(acc.astype(f32), f32(0)).astype(np.float64)
            else:
                for b in range(m["dim"]):
                    vb = m["bins"][b]
                    p = f32(f32(math.ceil(float(vb))) - vb)
                    out[:, b] = float(p) * x[:, int(math.floor(float(vb)))] + float(f32(f32(1) - p)) * x[:, int(math.ceil(float(vb)))]
            return out
        if t == "sr_norm":         # SRNormModule::generate :2039-2061
            x = ev(src[0], frames)
            fd = m["frame_dim"]
            out = np.zeros((x.shape[0], m["dim"]))
            for i in range(m["out_frames"]):
                st, cf = m["start"][i], m["coef"][i]
                acc = np.zeros((x.shape[0], fd))
                for j, c in enumerate(cf):
                    acc = acc + float(c) * x[:, (st + j) * fd:(st + j + 1) * fd]
                out[:, i * fd:(i + 1) * fd] = np.maximum(acc.astype(f32), f32(0)).astype(np.float64)
            return out
        if t == "quanteq":         # QuantEqModule::generate :2122-2140 (the linear term sits in the exponent, as written there)
            x = ev(src[0], frames)
            if m["alpha"] is None or m["gamma"] is None or m["qmax"] is None:
                return x.copy()
            a, g, q = (m[k][:m["dim"]] for k in ("alpha", "gamma", "qmax"))
            u = x / q.astype(np.float64)
            e = g.astype(np.float64) + (f32(1) - a).astype(np.float64) * u
            return q.astype(np.float64) * (a.astype(np.float64) * np.power(u, e))
        if t == "mean_subtractor":  # :1414-1454 (full-window branch; the recursive branch differs by ~1e-15)
            x = ev(src[0], frames)
            acc = np.zeros_like(x)
            for k in range(-m["ctx_l"], m["ctx_r"] + 1):
                acc += ev(src[0], frames + k)
            return x - acc / float(m["ctx_l"] + m["ctx_r"] + 1)
        raise ValueError(t)

    def _spectrum(self, m, a, frames, pcm, n):
        """AudioFileModule::generate (:371-440) + FFTModule::generate (:521-566)."""
        W, N = a["W"], pcm.size
        out = np.empty((len(frames), m["dim"]))
        x = np.concatenate([pcm.astype(f32), np.zeros(W + 2, f32)])
        uniq = {}
        for i, fr in enumerate(frames):
            fc = min(max(int(fr), 0), n - 1) if a["copy_borders"] else int(fr)
            if fc in uniq:
                out[i] = out[uniq[fc]]
                continue
            uniq[fc] = i
            ws = int(f32(f32(fc) * a["adv"]))
            idx = np.arange(ws, ws + W + 1)
            ok = (idx >= 0) & (idx < N)
            seg = np.where(ok, x[np.clip(idx, 0, N + W)], f32(0)).astype(f32)
            pre = (seg[1:] - (a["emph"] * seg[:-1]).astype(f32)).astype(f32)              # float pre-emphasis
            win = (m["window"].astype(np.float64) * pre.astype(np.float64)).astype(f32)     # float*double -> float
            spec = np.fft.rfft(win.astype(self.fft_dtype))
            if self.fft_dtype == np.float32:
                spec = _fft_float32(win)
            re, im = spec.real.astype(f32), spec.imag.astype(f32)
            p = (re * re + im * im).astype(f32)
            if m["magnitude"]:
                p = np.sqrt(p, dtype=f32)
            if m["log"]:
                p = np.log(p, dtype=f32)
            out[i] = p.astype(np.float64)
        return out


def _sincf(x):
    """util::sinc (aku/util.hh:151-159): float in, double arithmetic, float out."""
    x = f32(x)
    if abs(float(x)) < 1e-8:
        return f32(1)
    y = 3.14159265358979323846 * float(x)
    return f32(math.sin(y) / y)


def _full_conv(a, b):
    """Linear convolution (a * b)[j] = sum_{p+q=j} a[p] b[q], p ascending (the index bookkeeping of
    aku/FeatureModules.cc:1793-1815,1836-1862 spells out exactly this sum)."""
    out = [0.0] * (len(a) + len(b) - 1)
    for j in range(len(out)):
        lo = max(0, j - (len(b) - 1))
        hi = min(j, len(a) - 1)
        t = 0.0
        for p in range(lo, hi + 1):
            t += a[p] * b[j - p]
        out[j] = t
    return out


def _vtln_allpass_tables(m):
    """All-pass VTLN: the warp as a matrix on the cepstrum of the spectrum, wrapped in a DCT / inverse DCT and applied as
    full rows of `sinc` coefficients (create_all_pass_blin_transform :1717-1757, create_all_pass_slapt_transform
    :1759-1868, set_all_pass_transform :1870-1904)."""
    dim = m["dim"]
    T = np.zeros((dim, dim))
    T[0, 0] = 1.0
    if not m["use_slapt"]:
        alpha = float(f32(m["warp"] - f32(1)))               # double alpha = m_warp_factor - 1 (a float expression)
        q1 = [0.0] * dim
        q1[0] = -alpha
        temp = 1 - alpha * alpha
        for i in range(1, dim):
            q1[i] = temp
            temp *= alpha
        q = [0.0] * dim
        q[0] = 1.0
        for i in range(1, dim):                                # column i: the i-th convolution power of q1, truncated
            qn = [0.0] * dim
            for j in range(dim):
                t = 0.0
                for k in range(j + 1):
                    t += q[k] * q1[j - k]
                qn[j] = t
            q = qn
            T[0, i] = 2 * q[0]
            for j in range(1, dim):
                T[j, i] = q[j]
    else:
        sp = [float(x) for x in m["slapt_params"]]
        P = len(sp)
        f1 = [0.0] * (2 * P + 1)
        for i in range(P):
            f1[i] = -sp[P - i - 1] * math.pi / 2
            f1[i + P + 1] = sp[i] * math.pi / 2
        q = [0.0] * (2 * dim + 1)
        cur_f, center, cur_m = [1.0], 0, 1.0
        for i in range(11):                                    # exp of the sequence: sum over i of f1^(*i) / i!
            if i > 0:
                cur_m = cur_m / float(i)
            for j in range(max(0, dim - center), min(2 * dim + 1, dim + center + 1)):
                q[j] = q[j] + cur_m * cur_f[j - dim + center]
            cur_f = _full_conv(cur_f, f1)
            center = (len(cur_f) - 1) // 2
        q = q[:-2]                                             # "make the initial sequence symmetric"
        q1 = list(q)
        for i in range(1, dim):
            T[0, i] = 2 * q[dim - 1]
            for j in range(1, dim):
                T[j, i] = q[dim + j - 1] + q[dim - j - 1]
            full = _full_conv(q, q1)
            q = full[dim - 1:3 * dim - 2]
    # set_all_pass_transform: final = idct * (T * dct), sums over the inner index in ascending order
    dct = np.array([[math.cos(i * (j + 0.5) * math.pi / dim) for j in range(dim)] for i in range(dim)])
    idct = np.array([[(1.0 / dim) if j == 0 else math.cos((i + 0.5) * j * math.pi / dim) * 2 / dim for j in range(dim)] for i in range(dim)])
    tmp = np.zeros((dim, dim))
    for p in range(dim):
        tmp += T[:, p:p + 1] * dct[p:p + 1, :]
    fin = np.zeros((dim, dim))
    for p in range(dim):
        fin += idct[:, p:p + 1] * tmp[p:p + 1, :]
    m["bins"] = np.zeros(dim, dtype=f32)
    m["sinc_start"] = [0] * dim
    m["sinc_coef"] = [[f32(v) for v in fin[i]] for i in range(dim)]


def _vtln_tables(m):
    """create_blin_bins :1662-1675 / create_pwlin_bins :1634-1660 / create_slapt_bins :1677-1695 and
    create_sinc_coef_table :1697-1723: float bins, float taps, the reference's mixed arithmetic."""
    if m.get("all_pass"):
        return _vtln_allpass_tables(m)
    dim = m["dim"]
    bins = np.zeros(dim, dtype=f32)
    if m["use_slapt"]:
        for t in range(dim - 1):
            nf = math.pi * float(t) / (dim - 1)
            v = f32(t)
            for i, sp in enumerate(m["slapt_params"]):
                v = f32(float(v) + float(sp) * math.sin((i + 1) * nf) * (dim - 1))
            bins[t] = v
    elif m["use_pwlin"]:
        border = f32(m["turn"] * f32(dim - 1))
        slope = point = f32(0)
        limit = False
        for t in range(dim - 1):
            bins[t] = f32(m["warp"] * f32(t)) if not limit else f32(f32(slope * f32(t)) + point)
            if not limit and (t >= border or bins[t] >= border):
                slope = f32(f32(f32(f32(dim) - f32(1)) - bins[t]) / f32(f32(f32(dim) - f32(1)) - f32(t)))
                point = f32(f32(f32(1) - slope) * f32(dim - 1))
                limit = True
    else:
        w = m["warp"]
        for t in range(dim - 1):
            nf = math.pi * float(t) / (dim - 1)
            bins[t] = f32(t + 2 * math.atan2(float(f32(w - f32(1))) * math.sin(nf), 1 + float(f32(f32(1) - w)) * math.cos(nf))
                          / math.pi * (dim - 1))
    bins[dim - 1] = f32(dim - 1)
    m["bins"] = bins
    m["sinc_start"], m["sinc_coef"] = [], []
    rad = m["rad"]
    if rad > 0:
        for b in range(dim):
            cent = int(float(bins[b]) + 0.5)
            lo, hi = max(cent - rad, 0), min(cent + rad + 1, dim)
            cf = []
            for i in range(lo, hi):
                d = f32(f32(i) - bins[b])
                t = _sincf(d)
                if m["lanczos"]:
                    t = f32(t * _sincf(f32(d / f32(rad)))) if abs(float(d)) < rad else f32(0)
                cf.append(t)
            m["sinc_start"].append(lo)
            m["sinc_coef"].append(cf)


def _srnorm_tables(m):
    """SRNormModule::set_speech_rate :2004-2036."""
    in_cent = f32(f32(m["in_frames"] - 1) / f32(2))
    out_cent = f32(f32(m["out_frames"] - 1) / f32(2))
    m["start"], m["coef"] = [], []
    for i in range(m["out_frames"]):
        pos = f32(f32(f32(f32(i) - out_cent) / m["rate"]) + in_cent)
        cent = int(math.copysign(math.floor(abs(float(pos)) + 0.5), float(pos)))      # roundf: halves away from zero
        lo, hi = max(cent - m["order"], 0), min(cent + m["order"] + 1, m["in_frames"])
        cf = []
        for j in range(lo, hi):
            d = f32(f32(j) - pos)
            t = _sincf(d)
            t = f32(t * _sincf(f32(d / f32(m["order"])))) if abs(float(d)) < m["order"] else f32(0)
            cf.append(t)
        m["start"].append(lo)
        m["coef"].append(cf)


def _fft_float32(x):
    """Real FFT in float32 arithmetic (radix-2 for powers of two, like the float KissFFT build,
    vendor/kiss_fft; numpy's pocketfft in float64 rounded otherwise)."""
    n = x.size
    if n & (n - 1):
        return np.fft.rfft(x.astype(np.float64)).astype(np.complex64)
    a = x.astype(np.complex64)
    # iterative radix-2 DIT in complex64
    bits = n.bit_length() - 1
    rev = np.array([int(format(i, "0%db" % bits)[::-1], 2) for i in range(n)])
    a = a[rev]
    length = 2
    while length <= n:
        half = length // 2
        w = np.exp(-2j * np.pi * np.arange(half) / length).astype(np.complex64)
        a = a.reshape(-1, length)
        t = (a[:, half:] * w).astype(np.complex64)
        a = np.concatenate([a[:, :half] + t, a[:, :half] - t], axis=1).astype(np.complex64)
        a = a.reshape(-1)
        length *= 2
    return a[:n // 2 + 1]


def _cosf(x):
    return np.cos(f32(x), dtype=f32)


def _log10f(x):
    return np.log10(f32(x), dtype=f32)


def _mel(m, data):
    """MelModule::generate, aku/FeatureModules.cc:806-849 (float val/sum/scale)."""
    edges = m["edges"]
    out = np.empty((data.shape[0], m["dim"]))
    for b in range(m["dim"]):
        val = np.zeros(data.shape[0], dtype=f32)
        s = f32(0)
        beg = f32(edges[b] - f32(1))
        end = edges[b + 1]
        t = int(max(np.ceil(beg), f32(0)))
        while t < end:
            scale = f32(f32(f32(t) - beg) / f32(end - beg))
            val = (val.astype(np.float64) + np.float64(scale) * data[:, t]).astype(f32)
            s = f32(s + scale)
            t += 1
        beg = end
        end = edges[b + 2]
        while t < end:
            scale = f32(f32(end - f32(t)) / f32(end - beg))
            val = (val.astype(np.float64) + np.float64(scale) * data[:, t]).astype(f32)
            s = f32(s + scale)
            t += 1
        if m["root"]:
            out[:, b] = np.power((val / s).astype(np.float64), 0.1)
        else:
            out[:, b] = np.log((val / s).astype(f32) + f32(1), dtype=f32).astype(np.float64)
    return out


# ------------------------------------------------------------------------------------
# acoustic model
def read_model(base):
    """HmmSet::read_all: .mc (aku/HmmSet.cc:157-180), .ph (:209-329), .gk (aku/Distributions.cc:2812-2910)."""
    tok = open(base + ".mc").read().split()
    n = int(tok[0]); p = 1
    mix = []
    for _ in range(n):
        k = int(tok[p]); p += 1
        idx = [int(tok[p + 2 * j]) for j in range(k)]
        w = [float(tok[p + 2 * j + 1]) for j in range(k)]
        p += 2 * k
        mix.append((idx, w))
    tok = open(base + ".ph").read().split()
    assert tok[0] == "PHONE"
    phones = int(tok[1]); p = 2
    S = 0
    for _ in range(phones):
        states = int(tok[p + 1]) - 2
        p += 3 + 2
        for s in range(states):
            S = max(S, int(tok[p]) + 1); p += 1
        for s in range(states + 2):
            ntr = int(tok[p + 1]); p += 2 + 2 * ntr
    tok = open(base + ".gk").read().split()
    G, D, typ = int(tok[0]), int(tok[1]), tok[2]
    p = 3
    means = np.empty((G, D)); covs = np.ones((G, D))
    full_mask = np.zeros(G, dtype=bool); full_covs = np.zeros((G, D, D))
    for g in range(G):
        full = typ == "full_cov"
        if typ == "variable":
            full = tok[p] == "full"
            assert tok[p] in ("diag", "full")
            p += 1
        means[g] = [float(v) for v in tok[p:p + D]]; p += D
        if full:
            full_mask[g] = True
            full_covs[g] = np.array([float(v) for v in tok[p:p + D * D]]).reshape(D, D); p += D * D
        else:
            covs[g] = [float(v) for v in tok[p:p + D]]; p += D
    off = [0]; mg = []; mw = []
    for s in range(S):
        idx, w = mix[s]
        mg += idx; mw += w; off.append(len(mg))
    out = dict(mix_offsets=np.array(off, np.int32), mix_gauss=np.array(mg, np.int32),
               mix_weight=np.array(mw, np.float64), means=means, covs=covs)
    if full_mask.any():
        out["full_mask"], out["full_covs"] = full_mask, full_covs
    return out


def full_gaussian_params(mean, cov):
    """FullCovarianceGaussian::set_covariance (aku/Distributions.cc:1560-1586) +
    recompute_exponential_parameters (:1530-1547) + LinearAlgebra::map_m2v (aku/LinearAlgebra.cc:220-238).
    Returns (theta [D(D+3)/2], normalizer, constant); a non-SPD covariance gives zeros (precision = 0)."""
    D = mean.size
    L = D * (D + 3) // 2
    try:
        np.linalg.cholesky(cov)
        P = np.linalg.inv(cov)
        ch = np.linalg.cholesky(P)
    except np.linalg.LinAlgError:
        return np.zeros(L), 0.0, 0.0
    det = np.prod(np.diag(ch)) ** 2
    cst = math.log(math.sqrt(det))
    Pm = P @ mean
    theta = np.empty(L)
    theta[:D] = Pm
    pos = D
    for i in range(D):
        for j in range(i + 1):
            theta[pos] = -0.5 * (P[i, j] if i == j else math.sqrt(2.0) * P[i, j])
            pos += 1
    return theta, -0.5 * float(Pm @ mean), cst


def exponential_feature(f):
    """PDFPool::precompute_likelihoods (aku/Distributions.cc:2664-2672): [f ; map_m2v(f f^T)]."""
    D = f.shape[1]
    cols = [f]
    for i in range(D):
        for j in range(i + 1):
            m = f[:, i] * f[:, j]
            cols.append((m if i == j else math.sqrt(2.0) * m)[:, None])
    return np.concatenate(cols, axis=1)


def gaussian_params(means, covs):
    """DiagonalGaussian::read (:1132-1150) + set_constant (:1274-1288)."""
    covs = np.asarray(covs, dtype=np.float64)
    prec = np.where(covs > 0, 1.0 / np.where(covs > 0, covs, 1.0), 0.0)
    cst = np.ones(covs.shape[0])
    for i in range(covs.shape[1]):      # sequential product, as the reference multiplies
        cst = cst * prec[:, i]
    ok = cst > 0
    # libm's log element by element (numpy's SIMD log differs from glibc's in the last bit for about one argument in 1e4:
    # seen as a 7e-15 relative difference against the compiled reference on the 10000 x 32 model)
    cst = np.where(ok, _log(np.sqrt(np.where(ok, cst, 1.0))).astype(np.float64), cst)
    return prec, cst


def parse_clustering(text):
    """PDFPool::read_clustering (aku/Distributions.cc:3115-3147): cluster count, then `gauss cluster` pairs read in a
    `while (in) { in >> g >> c; ... }` loop -- after the last pair the stream is still good, the next extraction
    fails without touching g and c, and the last pair is recorded once more.  Returns (n_clusters, gauss[], cluster[])."""
    tok = text.split()
    n = int(tok[0])
    vals = [int(t) for t in tok[1:]]
    gi, ci = vals[0::2], vals[1::2]
    if len(gi) > len(ci):
        gi = gi[:len(ci)]
    if gi:
        gi.append(gi[-1])
        ci.append(ci[-1])
    return n, np.array(gi, dtype=np.int32), np.array(ci, dtype=np.int32)


def cluster_centers(model, n_clusters, gauss, cluster):
    """Centres by moment matching with unit weights, diagonal kept: Gaussian::merge (aku/Distributions.cc:854-897) +
    DiagonalGaussian::set_covariance (:1208-1234), operations in the reference's order.  Returns (means, covs, members)."""
    mu = np.asarray(model["means"], dtype=np.float64)
    cv = np.asarray(model["covs"], dtype=np.float64)
    D = mu.shape[1]
    members = [[] for _ in range(n_clusters)]
    for g, c in zip(gauss, cluster):
        members[int(c)].append(int(g))
    cm = np.zeros((n_clusters, D))
    cc = np.zeros((n_clusters, D))
    for c, L in enumerate(members):
        wsum = 0.0
        for _ in L:
            wsum += 1.0
        for g in L:
            cur = cv[g] + mu[g] * mu[g]
            cc[c] = cc[c] + 1.0 * cur
            cm[c] = cm[c] + 1.0 * mu[g]
        r = 1.0 / wsum
        cm[c] = cm[c] * r
        cc[c] = cc[c] * r
        cc[c] = cc[c] + (-1.0 * cm[c]) * cm[c]
    return cm, cc, members


def _diag_loglik(x, mu, prec, cst):
    ll = np.zeros((x.shape[0], mu.shape[0]))
    for i in range(x.shape[1]):         # ll += d*d*prec, left to right (:1052-1056)
        d = x[:, i:i + 1] - mu[None, :, i]
        ll += d * d * prec[None, :, i]
    ll *= -0.5
    ll += cst[None, :]
    return ll


def cmllr_adapt(W, feats):
    """Global model-level CMLLR (ConstrainedMllr with unitmode UNIT_NO, aku/ModelModules.cc:172-236): W is the `w1`
    matrix [D x (D+1)], column 0 = b, columns 1..D = A (:208-211).  Returns (A f + b for every frame, factor):
    AdaptedFeatureVector::calculate_new_ada_vector (aku/ModelModules.hh:208-212) copies b and adds A f; the factor that
    AdaptedGaussian::compute_likelihood (:170) multiplies every likelihood with is
    |LinearAlgebra::full_matrix_determinant(A)| (:141), and that routine multiplies the diagonal of A ITSELF, not of
    its LU factors (aku/LinearAlgebra.cc:74-86) -- restated as it is."""
    W = np.asarray(W, dtype=np.float64)
    D = W.shape[0]
    A, b = W[:, 1:], W[:, 0]
    feats = np.asarray(feats, dtype=np.float64)
    acc = np.zeros(feats.shape)
    for j in range(D):                        # row sums in index order
        acc += feats[:, j:j + 1] * A[None, :, j]
    det = 1.0
    for i in range(D):
        det *= A[i, i]
    return acc + b[None, :], abs(det)


def center_phone(label):
    """Hmm::get_center_phone (aku/HmmSet.cc:22-40): the b of a-b+c, a-b, b+c or b."""
    p1, p2 = label.rfind("-"), label.find("+")
    if p1 >= 0 and p2 >= 0:
        t = label[p1 + 1:p2] if p2 > p1 + 1 else ""
    elif p1 >= 0:
        t = label[p1 + 1:]
    elif p2 >= 0:
        t = label[:p2]
    else:
        t = label
    if not t:
        raise ValueError("Invalid phone label " + label)
    return t


def cmllr_unit_gaussians(unitmode, units, model, phones=None):
    """The Gaussians a regression-class transform applies to (RegClassTree::Unit*::get_gaussians, aku/RegClassTree.cc:
    302-322, 368-385, 444-454).  units: the strings in front of the matrix in a `w<i>` entry.  phones (UNIT_PHONE): list of
    (label, [state, ...]) from the .ph file; a state's emission pdf is the mixture of the same index."""
    off, mg = model["mix_offsets"], model["mix_gauss"]
    out = set()
    if unitmode == "UNIT_PHONE":
        for label, states in phones:
            if center_phone(label) in units:
                for st in states:
                    out.update(int(g) for g in mg[off[st]:off[st + 1]])
    elif unitmode == "UNIT_MIX":
        for u in units:
            try:
                m = int(u)
            except ValueError:
                continue                       # str2long failed: skipped (:378)
            out.update(int(g) for g in mg[off[m]:off[m + 1]])
    elif unitmode == "UNIT_GAUSSIAN":
        for u in units:
            try:
                out.add(int(u))
            except ValueError:
                out.add(0)                     # str2long leaves 0 and the result is used regardless (:451)
    else:
        raise ValueError("unitmode")
    return sorted(out)


def cmllr_unit_assignment(unitmode, transforms, model, phones=None):
    """ConstrainedMllr::load_transform (aku/ModelModules.cc:172-236): the transforms are visited in the order of their
    std::map key -- the vector of unit strings, compared lexicographically -- and every Gaussian of a transform is wrapped
    anew, so a Gaussian that several transforms claim ends up with the LAST one.  transforms: list of (units, W).
    Returns (gauss -> transform index or -1, transforms in visiting order)."""
    order = sorted(range(len(transforms)), key=lambda i: list(transforms[i][0]))
    g2t = np.full(np.asarray(model["means"]).shape[0], -1, dtype=np.int32)
    for rank, i in enumerate(order):
        for g in cmllr_unit_gaussians(unitmode, transforms[i][0], model, phones):
            g2t[g] = rank
    return g2t, [transforms[i] for i in order]


def state_likelihoods(model, feats, block=256, clustering=None, cmllr=None, cmllr_units=None):
    """HmmSet::precompute_likelihoods (aku/HmmSet.cc:485-501) for every frame: linear state
    likelihoods floored at 1e-50.  Sums run in the reference's order (dims then components).
    clustering = dict(n_clusters, gauss, cluster, min_clusters, min_gaussians) (the two ratios of
    HmmSet::set_clustering_min_evals) switches on the clustered branch of PDFPool::precompute_likelihoods
    (aku/Distributions.cc:2685-2722): clusters in descending order of their centre's likelihood (the reference pops a
    std::priority_queue; ties here go to the lower index), members of the leading clusters evaluated exactly while
    (clusters < min) or (Gaussians < min), the rest take the centre's likelihood -- and, because
    PDFPool::compute_likelihood (:2637-2644) only trusts cached values > 0, a centre likelihood of 0 means the Gaussian is
    evaluated exactly after all.
    cmllr = W [D x (D+1)]: every Gaussian and cluster centre is wrapped in an AdaptedGaussian (see cmllr_adapt).
    cmllr_units = (g2t, [W_0, W_1, ...]): regression-class transforms (see cmllr_unit_assignment): Gaussian g with
    g2t[g] = t >= 0 is evaluated at A_t f + b_t and multiplied by transform t's factor, the others are left alone."""
    feats = np.asarray(feats, dtype=np.float64)
    factor = None
    if cmllr is not None:
        feats, factor = cmllr_adapt(cmllr, feats)
    mu = np.asarray(model["means"], dtype=np.float64)
    prec, cst = gaussian_params(mu, model["covs"])
    off, mg = model["mix_offsets"], model["mix_gauss"]
    w = np.asarray(model["mix_weight"], dtype=np.float64).copy()
    S = len(off) - 1
    for s in range(S):                  # Mixture::normalize_weights (:2068-2075)
        a, b = off[s], off[s + 1]
        tot = 0.0
        for k in range(a, b):
            tot += w[k]
        w[a:b] = w[a:b] / tot
    F, D = feats.shape
    out = np.empty((F, S))
    full_cache = {}
    if clustering is not None:
        C = int(clustering["n_clusters"])
        cm, cc, members = cluster_centers(model, C, clustering["gauss"], clustering["cluster"])
        cprec, ccst = gaussian_params(cm, cc)
        csize = np.array([len(L) for L in members])
        g2c = np.full(mu.shape[0], -1)
        for c, L in enumerate(members):
            g2c[L] = c
        min_c = int(clustering["min_clusters"] * C)
        min_g = int(clustering["min_gaussians"] * mu.shape[0])
    for f0 in range(0, F, block):
        x = feats[f0:f0 + block]
        ll = np.zeros((x.shape[0], mu.shape[0]))
        for i in range(D):              # ll += d*d*prec, left to right (:1052-1056)
            d = x[:, i:i + 1] - mu[None, :, i]
            ll += d * d * prec[None, :, i]
        ll *= -0.5
        ll += cst[None, :]
        if "full_mask" in model:              # FullCovarianceGaussian::compute_log_likelihood_exponential (:1437-1446)
            phi = exponential_feature(x)
            for g in np.nonzero(model["full_mask"])[0]:
                th, nrm, c = full_cache.setdefault(int(g), full_gaussian_params(mu[g], np.asarray(model["full_covs"][g], dtype=np.float64)))
                dot = np.zeros(x.shape[0])
                for l in range(th.size):      # Blas_Dot_Prod, index order
                    dot += phi[:, l] * th[l]
                ll[:, g] = (dot + nrm) + c
        lik = _exp(ll).astype(np.float64)   # DiagonalGaussian::compute_likelihood (:1036)
        if factor is not None:
            lik = lik * factor              # AdaptedGaussian::compute_likelihood
        if cmllr_units is not None:
            g2t, Ws = cmllr_units
            for t, W in enumerate(Ws):
                gs = np.nonzero(np.asarray(g2t) == t)[0]
                if len(gs) == 0:
                    continue
                xa, fac = cmllr_adapt(W, x)
                lik[:, gs] = _exp(_diag_loglik(xa, mu[gs], prec[gs], cst[gs])).astype(np.float64) * fac
        if clustering is not None:
            clik = _exp(_diag_loglik(x, cm, cprec, ccst)).astype(np.float64)
            if factor is not None:
                clik = clik * factor
            for t in range(x.shape[0]):
                order = sorted(range(C), key=lambda c: (-clik[t, c], c))
                sel = np.zeros(C, dtype=bool)
                n_c = n_g = 0
                for c in order:
                    if not (n_c < min_c or n_g < min_g):
                        break
                    sel[c] = True
                    n_c += 1
                    n_g += csize[c]
                use_centre = (g2c >= 0) & ~sel[np.maximum(g2c, 0)] & (clik[t, np.maximum(g2c, 0)] > 0)
                lik[t, use_centre] = clik[t, g2c[use_centre]]
        for s in range(S):              # Mixture::compute_likelihood (:2079-2086)
            acc = np.zeros(x.shape[0])
            for k in range(off[s], off[s + 1]):
                acc += w[k] * lik[:, mg[k]]
            out[f0:f0 + x.shape[0], s] = np.maximum(acc, 1e-50)
    return out


def lna_records(lik, lnabytes=2, normalize=True):
    """The normalise/quantise loop of aku/phone_probs.cc:225-262.  lik: linear state
    likelihoods [F x S] (double).  Returns (bytes [F x S*lnabytes] uint8, float32 log-probs)."""
    with np.errstate(under="ignore"):
        obs = np.asarray(lik, dtype=np.float64).astype(f32)            # obs_log_probs is vector<float>
    F, S = obs.shape
    Z = np.zeros(F)
    for s in range(S):                                                 # double accumulator, state order
        Z += obs[:, s].astype(np.float64)
    if not normalize:
        Z[:] = 1
    Z[Z == 0] = 1
    x = obs.astype(np.float64) / Z[:, None]
    lp = np.where(x < 1e-50, math.log(1e-50), _log(np.where(x < 1e-50, 1.0, x)).astype(np.float64)).astype(f32)   # util::safe_log
    if lnabytes == 4:
        return lp.astype("<f4").view(np.uint8).reshape(F, S * 4), lp
    lpd = lp.astype(np.float64)
    code = np.where(lpd < -36.008, 65535, (-1820.0 * lpd + .5).astype(np.int64))    # (int) truncates toward zero
    hi = (code >> 8) & 255
    lo = code & 255
    rec = np.stack([hi, lo], axis=2).astype(np.uint8).reshape(F, S * 2)
    return rec, lp
