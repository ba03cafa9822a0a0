"""Readers/writers for the reference's on-disk formats.

Host-side I/O only -- nothing here computes features or likelihoods.
  WAV     RIFF PCM16 mono, what AudioReader accepts (aku/AudioReader.cc:86-155)
  .gk     PDFPool::read_gk / write_gk        aku/Distributions.cc:2812-2945
  .mc     HmmSet::read_mc / Mixture::write   aku/HmmSet.cc:157-180, aku/Distributions.cc:2405-2415
  .ph     HmmSet::read_legacy_ph / write_ph  aku/HmmSet.cc:209-329, 379-425
  recipe  Recipe::read                        aku/Recipe.cc:24-147
  LNA     phone_probs                         aku/phone_probs.cc:213-262
"""
import struct

import numpy as np


def write_wav(path, pcm, sample_rate):
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    data = pcm.tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, sample_rate, sample_rate * 2, 2, 16))
        f.write(b"data" + struct.pack("<I", len(data)))
        f.write(data)


def read_wav(path):
    """Returns (int16 samples, sample_rate).  PCM16 mono only, like the reference."""
    return parse_wav(open(path, "rb").read(), path)


def parse_wav(b, path="<bytes>"):
    """read_wav on the bytes of a RIFF/WAVE file."""
    if b[:4] != b"RIFF" or b[8:12] != b"WAVE":
        raise ValueError("%s: not a RIFF/WAVE file" % path)
    p, sr, ch, bits, fmt = 12, None, None, None, None
    while p + 8 <= len(b):
        cid, ln = b[p:p + 4], struct.unpack("<I", b[p + 4:p + 8])[0]
        if cid == b"fmt ":
            fmt, ch, sr = struct.unpack("<HHI", b[p + 8:p + 16])
            bits = struct.unpack("<H", b[p + 22:p + 24])[0]
            if fmt == 0xFFFE and ln >= 40:      # WAVE_FORMAT_EXTENSIBLE: format tag = first field of the SubFormat GUID
                fmt = struct.unpack("<H", b[p + 32:p + 34])[0]
        elif cid == b"data":
            if fmt != 1 or bits != 16:
                raise ValueError("%s: sample format not PCM16" % path)
            if ch != 1:
                raise ValueError("AudioReader: sorry, audio files with multiple channels not supported")
            return np.frombuffer(b, dtype="<i2", count=min(ln, len(b) - p - 8) // 2, offset=p + 8).copy(), sr
        p += 8 + ln + (ln & 1)
    raise ValueError("%s: no data chunk" % path)


def write_model(base, mix_offsets, mix_gauss, mix_weight, means, covs=None, full_covs=None, full_mask=None):
    """Writes base.gk / base.mc / base.ph (legacy PHONE format, one single-state phone per
    mixture so that state k <-> mixture k, as HmmSet::read_legacy_ph binds them).
    covs: [G x D] diagonal variances; full_covs: [G x D x D] for the Gaussians where full_mask is set
    (all of them when full_mask is None and covs is None)."""
    means = np.asarray(means, dtype=np.float64)
    G, D = means.shape
    if full_covs is not None and full_mask is None:
        full_mask = np.ones(G, dtype=bool) if covs is None else np.zeros(G, dtype=bool)
    S = len(mix_offsets) - 1
    with open(base + ".gk", "w") as f:
        f.write("%d %d variable\n" % (G, D))
        for g in range(G):
            if full_mask is not None and full_mask[g]:
                f.write("full " + " ".join(repr(float(v)) for v in means[g]) + " " +
                        " ".join(repr(float(v)) for v in np.asarray(full_covs[g], dtype=np.float64).reshape(-1)) + "\n")
            else:
                f.write("diag " + " ".join(repr(float(v)) for v in means[g]) + " " +
                        " ".join(repr(float(v)) for v in np.asarray(covs[g], dtype=np.float64)) + "\n")
    with open(base + ".mc", "w") as f:
        f.write("%d\n" % S)
        for s in range(S):
            a, b = mix_offsets[s], mix_offsets[s + 1]
            f.write("%d" % (b - a))
            for k in range(a, b):
                f.write(" %d %s" % (mix_gauss[k], repr(float(mix_weight[k]))))
            f.write("\n")
    with open(base + ".ph", "w") as f:
        f.write("PHONE\n%d\n" % S)
        for s in range(S):
            f.write("%d 3 p%d\n-1 -2 %d\n0 1 2 1\n1 0\n2 2 2 0.8 1 0.2\n" % (s + 1, s, s))


def config_audio_format(cfg_text):
    """(raw, big_endian) from the `raw` / `endian` keys of the FIRST module of a feature configuration -- the audiofile
    module's keys that concern the file, not the signal processing (aku/FeatureModules.cc:345-356)."""
    raw, big, depth, module = False, False, 0, 0
    for line in cfg_text.splitlines():
        f = line.split()
        if not f:
            continue
        if f[0] == "{":
            depth += 1
        elif f[0] == "}":
            depth -= 1
            if depth == 0 and module == 1:
                break
        elif depth == 0 and f[0] == "module":
            module += 1
        elif depth == 1 and module == 1 and len(f) >= 2:
            if f[0] == "raw":
                raw = int(f[1]) != 0
            if f[0] == "endian":
                big = f[1] == "big"
    return raw, big


def read_recipe(path, num_batches=0, batch_index=0):
    """aku::Recipe::read (aku/Recipe.cc:24-147): one utterance per line of key=value fields; a key missing on a later
    line inherits the previous line's value (the reference never clears its map, :29,82-90), also across batch borders.
    num_batches > 1: the contiguous part `batch_index` (1-based) of the non-empty, non-comment lines, the first
    (lines % batches) parts one line longer (:63-115 with cluster_speakers = false, as phone_probs calls it)."""
    lines = []
    for line in open(path):
        line = line.strip("\n\t \r")
        if line and not line.startswith("#"):
            lines.append(line)
    if num_batches > 1 and not 1 <= batch_index <= num_batches:
        raise ValueError("Invalid batch index")
    first, last = 0, len(lines)
    if num_batches > 1:
        per, rem = divmod(len(lines), num_batches)
        first = (batch_index - 1) * per + min(batch_index - 1, rem)
        last = first + per + (1 if batch_index - 1 < rem else 0)
    infos, cur = [], {}
    for i, line in enumerate(lines[:last]):
        for field in line.replace("\t", " ").split(" "):
            if not field:
                continue
            kv = field.split("=")
            if len(kv) != 2 or field.endswith("="):          # str::split(field, "=", false) must give two parts
                raise ValueError("Invalid recipe line: " + line)
            cur[kv[0]] = kv[1]
        if i >= first:
            infos.append(dict(cur))
    return infos


def sort_recipe(infos):
    """Recipe::sort_infos (aku/Recipe.hh:89-91,115-117; phone_probs --sort-recipe): stable sort by speaker."""
    return sorted(infos, key=lambda d: d.get("speaker", ""))


def lna_header(num_states, lnabytes):
    return struct.pack(">I", num_states) + bytes([lnabytes])


def write_lna(path, records, num_states, lnabytes):
    with open(path, "wb") as f:
        f.write(lna_header(num_states, lnabytes))
        f.write(np.ascontiguousarray(records, dtype=np.uint8).tobytes())


def read_lna(path):
    """Returns (log_probs [F x S] float32, num_states, lnabytes) decoded as the decoder does
    (decoder/src/LnaReaderCircular.cc:170-198)."""
    b = open(path, "rb").read()
    S, nb = struct.unpack(">I", b[:4])[0], b[4]
    body = b[5:]
    if nb == 2:
        codes = np.frombuffer(body, dtype=">u2").reshape(-1, S)
        return (codes.astype(np.float64) / -1820.0).astype(np.float32), S, nb
    if nb == 4:
        return np.frombuffer(body, dtype="<f4").reshape(-1, S).copy(), S, nb
    raise ValueError("unsupported LNA byte width %d" % nb)
