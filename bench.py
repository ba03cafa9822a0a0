#!/usr/bin/env python
"""bench.py -- acoustic frames/sec of the accelerated path (MFCC front-end + diagonal-GMM
log-likelihoods + LNA encoding) on N B200s, and the reference's own CPU path beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm
  python bench.py --impl reference [--steps K] [--warmup W]       reference arm (CPU, rank 0 only)

Workload (BASELINE.json configs[1]): a 1000-utterance synthetic batch (10 s of 16 kHz audio each,
1248 frames per utterance), 39-dim MFCC+delta+delta-delta, 5000-state x 16-mixture diagonal GMM,
2-byte LNA output.  One step = one pass of the hot path over that batch (per GPU: weak scaling,
utterances are independent, no data-path collective).

  value : frames/s with PCM already resident in HBM and LNA written to HBM (CUDA events)
  e2e   : the same through the host-buffer entry point: PCM from pinned host memory, LNA bytes
          back to pinned host memory, copies inside the timed region
  roofline : the scorer kernel (gmm_tc16_kernel, tcgen05) against the measured bf16 tensor peak
          (MEASURED_PEAKS.json) -- batched scoring is compute bound; the HBM view of the same
          launches is reported next to it
  cpu_baseline : the reference's own FeatureGenerator + HmmSet code (oracle/_ref, built from
          /root/reference) on the host cores, on a bounded sample of the same workload
  sub_records : the other BASELINE.json configs in the same driver-run line (each a small, separately timed run):
          N = 1: parity_mode_f64 (config-2 model in the byte-exact double arithmetic), config3_feature_sweep,
                 config5_full_covariance (+ its own cpu_baseline), config4_model_1gpu (10000 x 32 on one GPU),
                 streaming (the decoder's per-frame feed: microseconds per call, sweep rate against L2 / HBM)
          N > 1: config4 (100 h sharded by utterance over the N GPUs, 10000 x 32: model broadcast, frame-count
                 all-gather, LPT partition, per-rank writers AND the LNA gather to one writer rank -- fused p2p stores
                 and ncclSend/Recv -- with a checksum sink, 1-vs-N per-utterance checksum check), host_d2h (the
                 concurrent PCIe probe that explains the e2e curve)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_UTTS = 1000
UTT_SAMPLES = 160000
N_STATES, N_MIX = 5000, 16
LNABYTES = 2
SAMPLE_RATE = 16000
WORKLOAD = "1000 utterances x 10 s @16 kHz, 39-dim MFCC+d+dd, 5000-state x 16-mix diag GMM, 2-byte LNA"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_audio(rank, n_utts):
    from aaltoasr_b200 import synth
    out = np.empty(n_utts * UTT_SAMPLES, dtype=np.int16)
    block = 50
    distinct = min(n_utts, 1000)          # beyond 1000 utterances the audio repeats (generation time)
    for b0 in range(0, distinct, block):
        nb = min(block, distinct - b0)
        # one long seeded stream per block, cut into utterances (fast; every utterance differs)
        x = synth.synth_audio(2000 + 1000 * rank + b0, nb * UTT_SAMPLES, SAMPLE_RATE)
        out[b0 * UTT_SAMPLES:(b0 + nb) * UTT_SAMPLES] = x
    for b0 in range(distinct, n_utts, distinct):
        nb = min(distinct, n_utts - b0)
        out[b0 * UTT_SAMPLES:(b0 + nb) * UTT_SAMPLES] = out[:nb * UTT_SAMPLES]
    return out


def make_model(eng):
    """Same model on every rank: means drawn from the features of 16 fixed utterances."""
    from aaltoasr_b200 import synth
    pcm = np.concatenate([synth.synth_audio(2000 + i, UTT_SAMPLES, SAMPLE_RATE) for i in range(16)])
    uo = np.arange(17, dtype=np.int64) * UTT_SAMPLES
    feats, _ = eng.features(pcm, uo, dtype=np.float64)
    return synth.synth_diag_model(2999, feats, N_STATES, N_MIX)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        # samples under load only (the idle ones before/after the region pull the median down)
        load = [s for s, p in zip(sm, power) if p > 300] or sm
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own classes from oracle/_ref on the host cores
def _ref_worker(args):
    base, cfg, wav, n_frames = args
    from oracle import oracle_np, ref
    M = ref.Model(base)                       # model load is not timed
    t0 = time.perf_counter()
    feats, _, _ = ref.features(cfg, wav, 0, n_frames)          # FeatureGenerator::generate per frame
    lik = M.state_likelihoods(feats)                           # HmmSet::precompute_likelihoods + state_likelihood
    rec, _ = oracle_np.lna_records(lik, LNABYTES)              # normalise + quantise (aku/phone_probs.cc:225-262)
    dt = time.perf_counter() - t0
    M.close()
    return dt, int(rec.shape[0])


def reference_fixture(tmp, model, n_wavs):
    from aaltoasr_b200 import formats, synth
    cfg = os.path.join(tmp, "mfcc.cfg")
    open(cfg, "w").write(synth.mfcc39_config(SAMPLE_RATE))
    base = os.path.join(tmp, "model")
    t0 = time.time()
    formats.write_model(base, **model)
    log("reference arm: wrote %s.gk/.mc/.ph in %.1f s" % (base, time.time() - t0))
    wavs = []
    for i in range(n_wavs):
        w = os.path.join(tmp, "u%d.wav" % i)
        formats.write_wav(w, synth.synth_audio(2000 + i, UTT_SAMPLES, SAMPLE_RATE), SAMPLE_RATE)
        wavs.append(w)
    return cfg, base, wavs


def run_reference_sample(cfg, base, wavs, cores, frames_per_core):
    """One bounded sample: every core scores `frames_per_core` frames of its own utterance with the
    reference's code.  Returns (frames/s, seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_ref_worker, [(base, cfg, wavs[i % len(wavs)], frames_per_core) for i in range(cores)])
    frames = sum(r[1] for r in res)
    secs = max(r[0] for r in res)
    return frames / secs, secs


def ref_model_from_oracle():
    """Model for the reference arm without a GPU: means from the oracle's features (same seeds)."""
    from aaltoasr_b200 import synth
    from oracle import oracle_np
    P = oracle_np.Pipeline(synth.mfcc39_config(SAMPLE_RATE))
    feats = np.concatenate([P.run(synth.synth_audio(2000 + i, UTT_SAMPLES, SAMPLE_RATE)) for i in range(16)])
    return synth.synth_diag_model(2999, feats, N_STATES, N_MIX)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}))
        return
    cores = os.cpu_count() or 1
    frames_per_core = 384
    with tempfile.TemporaryDirectory() as tmp:
        cfg, base, wavs = reference_fixture(tmp, ref_model_from_oracle(), min(cores, 16))
        for _ in range(args.warmup):
            run_reference_sample(cfg, base, wavs, cores, 32)
        t_total, f_total = 0.0, 0
        for _ in range(args.steps):
            fps, secs = run_reference_sample(cfg, base, wavs, cores, frames_per_core)
            t_total += secs
            f_total += fps * secs
    value = f_total / t_total
    sample = "%d frames per core on %d cores per step (FeatureGenerator + HmmSet from oracle/_ref, model load excluded)" % (
        frames_per_core, cores)
    print(json.dumps({
        "impl": "reference", "metric": "acoustic frames/sec (MFCC+GMM log-lik -> LNA)", "value": value,
        "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def bind_to_gpu_numa(torch, local):
    """Several ranks on one host: keep this rank's threads (and with them its pinned host buffers, first touch) on the
    CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI device).  Best effort: any failure leaves the affinity alone."""
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local).uuid)
        out = subprocess.run(["nvidia-smi", "--query-gpu=uuid,pci.bus_id", "--format=csv,noheader"], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, text=True, timeout=20).stdout
        bus = None
        for ln in out.splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) == 2 and f[0] == uuid:
                bus = f[1]
        if not bus:
            return None
        dom, rest = bus.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/local_cpulist" % (dom[-4:].lower(), rest.lower())
        cpus = set()
        for part in open(path).read().strip().split(","):
            if "-" in part:
                a, z = part.split("-")
                cpus.update(range(int(a), int(z) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# --------------------------------------------------------------------------------------
def _event_time(torch, stream, fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bench_feature_sweep(args, local):
    """BASELINE.json configs[2]: FeatureGenerator only (audiofile -> fft -> mel -> dct + power -> delta -> delta-delta),
    sample rates 8-64 kHz at the reference's default window (sample_rate / 62.5) and window widths 256-2048 at 16 kHz.
    Algorithmic HBM bytes per frame = 2 * hop (PCM in, each sample once) + 4 * dim (features out)."""
    import torch
    from aaltoasr_b200 import AkuGpu, synth
    eng = AkuGpu(local)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    hbm_peak, hbm_src = measured_peaks()
    settings = [(sr, None) for sr in (8000, 16000, 32000, 48000, 64000)] + [(16000, ww) for ww in (256, 512, 1024, 2048)]
    secs, n_utts = 60, 16
    sweep, head = [], None
    for sr, ww in settings:
        cfg = synth.mfcc39_config(sr)
        if ww:
            cfg = cfg.replace("sample_rate %d" % sr, "sample_rate %d\n  window_width %d" % (sr, ww))
        eng.frontend_load_config_text(cfg)
        one = synth.synth_audio(3000 + sr // 1000, secs * sr, sr)
        pcm = np.tile(one, n_utts)
        uo = np.arange(n_utts + 1, dtype=np.int64) * one.size
        fo = eng.frame_offsets(uo)
        F, dim = int(fo[-1]), eng.feature_dim
        pcm_d = torch.from_numpy(pcm).cuda()
        out_d = torch.empty((F, dim), dtype=torch.float32, device="cuda")
        pcm_p = torch.from_numpy(pcm).pin_memory()
        out_p = torch.empty((F, dim), dtype=torch.float32).pin_memory()
        is_head = sr == 16000 and ww is None
        reps = args.steps if is_head else 3
        for _ in range(args.warmup if is_head else 2):
            eng.features(pcm_d, uo, out=out_d)
        l0 = eng.launch_count()
        ms = _event_time(torch, stream, lambda: eng.features(pcm_d, uo, out=out_d), reps)
        launches = eng.launch_count() - l0
        eng.features(pcm_p, uo, out=out_p)            # warm-up of the host-buffer path (staging buffers grow on first use)
        ms_e2e = _event_time(torch, stream, lambda: eng.features(pcm_p, uo, out=out_p), reps)
        hop = sr / 125.0
        alg = (2 * hop + 4 * dim) * F
        row = {"sample_rate": sr, "window": ww or int(sr / 62.5), "frames": F, "frames_per_s": F / (ms * 1e-3),
               "e2e_frames_per_s": F / (ms_e2e * 1e-3), "audio_x_realtime": F / 125.0 / (ms * 1e-3),
               "algorithmic_GBps": alg / (ms * 1e-3) / 1e9, "hbm_frac": alg / (ms * 1e-3) / 1e9 / hbm_peak}
        sweep.append(row)
        log("config 3: %d Hz window %d: %.1f M frames/s resident, %.1f M e2e, %.1f GB/s algorithmic" % (
            sr, row["window"], row["frames_per_s"] / 1e6, row["e2e_frames_per_s"] / 1e6, row["algorithmic_GBps"]))
        if is_head:
            head = dict(row, ms=ms, ms_e2e=ms_e2e, launches=launches, h2d=int(pcm.nbytes), d2h=int(F * dim * 4))
    eng.close()
    return {"metric": "acoustic frames/sec (FeatureGenerator only: MFCC+d+dd)", "value": head["frames_per_s"], "unit": "frames/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
            "config": {"workload": "feature-only sweep; headline = 16 x 60 s @16 kHz, 256-sample window, 39-dim MFCC+d+dd",
                       "l2_policy": "PCM + module buffers per step exceed L2"},
            "e2e": {"value": head["e2e_frames_per_s"], "unit": "frames/s", "h2d_bytes_per_step": head["h2d"],
                    "d2h_bytes_per_step": head["d2h"], "ms_per_step": head["ms_e2e"]},
            "gpu_launches": int(head["launches"]),
            "roofline": {"kernel": "fe_spectrum_bfft<256> (FFT in the registers of one warp per frame, 32 frames per CTA; fused mel + power + dct + merge with lane = frame) + fe_delta2_merge_tiled",
                         "bound": "hbm", "achieved": head["algorithmic_GBps"], "peak": hbm_peak, "unit": "GB/s", "frac": head["hbm_frac"],
                         "traffic": None, "peak_source": hbm_src,
                         "algorithmic": "2 * hop + 4 * dim = 412 B per frame (SURVEY.md 8d) x frames / time of a whole akugpu_features call",
                         "note": "the stage is issue / latency bound, not bandwidth bound: ncu (profiles/r02_fe_spectrum_bfft_ncu_full.txt) shows "
                                 "1043 warp instructions per frame at 68 % issue-slot utilisation, DRAM traffic 260 B per frame at 2 % of "
                                 "DRAM throughput; the operations are the reference's float / double sequence, reproduced exactly"},
            "sweep": sweep, "cpu_baseline": None}


def bench_full_cov(args, local, cpu_baseline=False):
    """BASELINE.json configs[4]: 2000-state x 16-mixture full-covariance pool (FullCovarianceGaussian; the subspace
    classes are dead code in the reference build), features resident -> LNA, tensor-core expanded form."""
    import torch
    from aaltoasr_b200 import AkuGpu, synth
    eng = AkuGpu(local)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    eng.frontend_load_config_text(synth.mfcc39_config(SAMPLE_RATE))
    S, M, D = 2000, 16, 39
    n_utts = args.utts or 100
    base = [synth.synth_audio(2000 + i, UTT_SAMPLES, SAMPLE_RATE) for i in range(8)]
    pcm = np.concatenate([base[i % 8] for i in range(n_utts)])
    uo = np.arange(n_utts + 1, dtype=np.int64) * UTT_SAMPLES
    feats, fo = eng.features(pcm, uo, dtype=np.float32)
    F = int(fo[-1])
    rng = np.random.default_rng(5999)
    G = S * M
    f64 = feats[:20000].astype(np.float64)
    sd = f64.std(axis=0)
    means = f64[rng.integers(0, f64.shape[0], G)] + 0.3 * sd * rng.standard_normal((G, D))
    A = rng.standard_normal((G, D, 4)) * sd[None, :, None]
    full = np.einsum("gik,gjk->gij", A, A) * 0.1
    full[:, np.arange(D), np.arange(D)] += rng.uniform(0.5, 2, (G, D)) * sd ** 2
    t0 = time.time()
    weights = rng.dirichlet(np.ones(M), S).reshape(-1)
    eng.model_load_full(np.arange(0, G + 1, M, dtype=np.int32), np.arange(G, dtype=np.int32), weights, means, full)
    log("config 5: model packed in %.1f s, %d frames" % (time.time() - t0, F))
    feats_d = torch.from_numpy(feats).cuda()
    out_d = torch.empty((F, S * LNABYTES), dtype=torch.uint8, device="cuda")
    feats_p = torch.from_numpy(feats).pin_memory()
    out_p = torch.empty((F, S * LNABYTES), dtype=torch.uint8).pin_memory()
    for _ in range(args.warmup):
        eng.gmm_lna(feats_d, lnabytes=LNABYTES, out=out_d)
    eng.stage_times_reset(True)
    l0 = eng.launch_count()
    ms = _event_time(torch, stream, lambda: eng.gmm_lna(feats_d, lnabytes=LNABYTES, out=out_d), args.steps)
    launches = eng.launch_count() - l0
    st = eng.stage_times()
    eng.stage_times_reset(False)
    ms_e2e = _event_time(torch, stream, lambda: eng.gmm_lna(feats_p, lnabytes=LNABYTES, out=out_p), args.steps)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    L = D * (D + 3) // 2
    Kp = 3 * (-(-(L + 2) // 64) * 64)          # three fp16 split products over whole 64-wide k-blocks
    gmm_ms, gmm_launches = st["gmm"]
    # the stage also holds the feature expansion kernel; the MMA work is what is counted
    mma_flop = 2.0 * Kp * G * (-(-F // 128) * 128) * args.steps
    achieved = mma_flop / (gmm_ms * 1e-3) / 1e12
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    cpu = None
    if cpu_baseline:
        try:
            n_sub = 32
            g = np.arange(n_sub * M)
            sub = dict(mix_offsets=np.arange(0, n_sub * M + 1, M, dtype=np.int32), mix_gauss=np.arange(n_sub * M, dtype=np.int32),
                       mix_weight=weights[:n_sub * M], means=means[g], full_covs=full[g])
            cpu = sub_full_cov_cpu_baseline(sub, feats[:128].astype(np.float64), S)
        except Exception as e:
            cpu = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}
    useful_tf = 2.0 * L * G * F * args.steps / (gmm_ms * 1e-3) / 1e12
    eng.close()
    return {"metric": "acoustic frames/sec (full-covariance GMM log-lik -> LNA)", "value": F / (ms * 1e-3), "unit": "frames/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16x2->f32", "data": "synthetic",
            "config": {"workload": "%d utterances x 10 s, 39-dim features resident, 2000-state x 16-mix full-covariance GMM, 2-byte LNA" % n_utts,
                       "frames_per_gpu": F, "l2_policy": "expanded features + outputs per step exceed L2"},
            "e2e": {"value": F / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(feats.nbytes),
                    "d2h_bytes_per_step": int(F * S * LNABYTES), "ms_per_step": ms_e2e},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "gmm_tc16_kernel<0> (tcgen05 kind::f16, fp16 hi/lo-split exponential form, A' and B' streamed, K' = %d issued for %d useful terms)" % (Kp, L),
                         "bound": "tensor", "achieved": useful_tf, "peak": peak, "unit": "TFLOP/s", "frac": useful_tf / peak,
                         "traffic": None, "launches": int(gmm_launches), "stage_ms": {"gmm+expand": gmm_ms, "lna": st["lna"][0]},
                         "algorithmic_flop_per_frame": 2.0 * L * G,
                         "algorithmic": "2 K G per frame with K = D(D+3)/2 = %d (SURVEY.md 8d, the reference's own exponential form)" % L,
                         "issued_mma": {"achieved": achieved, "unit": "TFLOP/s", "frac_of_peak": achieved / peak, "k_issued": Kp, "k_algorithmic": L}},
            "cpu_baseline": cpu}


# --------------------------------------------------------------------------------------
def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p)) if os.path.exists(p) else {}


def timed_region(torch, dist, world, stream, fn, steps):
    """barrier + synchronize on both sides, CUDA events on the launching stream, MAX over ranks (ms for `steps` calls)."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def scorer_roofline(eng, st, F, steps, ms_total, S, M, D=39):
    """Roofline record of the dominant kernel (gmm_tc16_kernel) from the stage timers of the timed region.
    `achieved` counts ALGORITHMIC flops: SURVEY.md section 8(d)'s GEMM-form figure 2 (2D+1) S M per frame; the MMA work the
    kernel actually issues (three fp16 split products over whole K16 steps) is reported beside it."""
    peaks = _peaks()
    hbm_peak, hbm_src = measured_peaks()
    gmm_ms, gmm_launches = st["gmm"]
    frames_per_launch = F * steps / max(1, gmm_launches)
    avg_ms = gmm_ms / max(1, gmm_launches)
    G = S * M
    alg_flop_per_frame = 2.0 * (2 * D + 1) * G
    achieved = alg_flop_per_frame * frames_per_launch / (avg_ms * 1e-3) / 1e12
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)       # kernel timed inside a long step -> the sustained figure
    nch = -(-(2 * D + 2) // 16)
    Kp = 3 * nch * 16
    comps = -(-S * (-(-M // 16) * 16) // 128) * 128
    frames_pad = -(-int(frames_per_launch) // 128) * 128
    issued = 2.0 * Kp * comps * frames_pad / (avg_ms * 1e-3) / 1e12
    bytes_per_frame = D * 4 + S * 4                                # features in, state log-likelihoods out
    param_bytes = comps * (-(-2 * nch // 4) * 64) * 2
    alg_bytes = bytes_per_frame * frames_per_launch + param_bytes
    hbm_gbs = alg_bytes / (avg_ms * 1e-3) / 1e9
    rates = eng.pipe_rates()
    mufu = G * frames_per_launch / (avg_ms * 1e-3)
    in_use = eng.scorer_in_use()
    r = {"kernel": "gmm_tc16_kernel (tcgen05 kind::f16, fp16 hi/lo-split expanded form, A' resident in shared memory)",
         "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
         "traffic": None,
         "algorithmic_flop_per_frame": alg_flop_per_frame,
         "algorithmic": "2 (2D+1) S M per frame (SURVEY.md 8d, GEMM form) x frames per launch / average launch time",
         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400 (B200_PROFILING.md sustained)",
         "launches": int(gmm_launches), "avg_launch_ms": avg_ms, "frames_per_launch": frames_per_launch,
         "share_of_step": gmm_ms / (ms_total if ms_total > 0 else 1),
         "issued_mma": {"achieved": issued, "unit": "TFLOP/s", "frac_of_peak": issued / peak_tf, "k_issued": Kp, "k_algorithmic": 2 * D + 1,
                        "note": "tensor-pipe occupancy: three fp16 split products (hi.hi, hi.lo, lo.hi) over %d K16 steps" % nch},
         "mufu_view": {"achieved": mufu, "peak": rates["ex2"], "unit": "exp2/s", "frac": mufu / rates["ex2"]},
         "hbm_view": {"bound": "hbm", "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                      "peak_source": hbm_src, "algorithmic_bytes_per_launch": alg_bytes},
         "traffic_note": "DRAM bytes are not measurable inside the run; the ncu --set full capture of this kernel is profiles/r01_gmm_tc16_ncu_full.txt "
                         "(762 MB per launch against 793 MB algorithmic)",
         "fp32_pipe_peak_tflops": 2.0 * rates["tile_ffma2"] / 1e12,
         "scorer_in_use": {0: "double path", 1: "gmm_diag_f32 (FP32 pipe)", 2: "gmm_tc_kernel (bf16x3)",
                           3: "gmm_tc16_kernel (resident A')", 4: "gmm_tc16_kernel<0> (streaming A')",
                           5: "gmm_tc16_kernel + gmm_diag_f32 for ill-conditioned states"}[in_use],
         "expanded_form_q_max": eng.expanded_form_q(),
         "stage_ms": {"frontend": st["frontend"][0], "gmm": gmm_ms, "lna": st["lna"][0]}}
    if in_use != 3:
        r["kernel"] = "WARNING: the arithmetic below assumes gmm_tc16_kernel; in use: " + r["scorer_in_use"]
    return r


def sub_parity_mode(args, torch, eng, stream, pcm_d, pcm_p, uo, fo, n_utts=24):
    """The config-2 model in parity arithmetic (F64: the reference's operations in double, byte-identical LNA)."""
    from aaltoasr_b200 import F64
    n = min(n_utts, len(uo) - 1)
    F = int(fo[n])
    rec = N_STATES * LNABYTES
    out_d = torch.empty((F, rec), dtype=torch.uint8, device="cuda")
    out_p = torch.empty((F, rec), dtype=torch.uint8).pin_memory()
    sub_uo = uo[:n + 1]
    ns = int(uo[n])
    eng.phone_probs(pcm_d[:ns], sub_uo, precision=F64, lnabytes=LNABYTES, out=out_d)
    ms = _event_time(torch, stream, lambda: eng.phone_probs(pcm_d[:ns], sub_uo, precision=F64, lnabytes=LNABYTES, out=out_d), 2)
    ms_e = _event_time(torch, stream, lambda: eng.phone_probs(pcm_p[:ns], sub_uo, precision=F64, lnabytes=LNABYTES, out=out_p), 2)
    # how far the headline (throughput) mode is from these bytes: the same utterances in F32, code by code (outside any timing)
    codes = None
    if LNABYTES == 2:
        from aaltoasr_b200 import F32
        eng.phone_probs(pcm_d[:ns], sub_uo, precision=F64, lnabytes=2, out=out_d)
        out32 = torch.empty_like(out_d)
        eng.phone_probs(pcm_d[:ns], sub_uo, precision=F32, lnabytes=2, out=out32)
        torch.cuda.synchronize()
        a, b = out_d.view(F, N_STATES, 2).to(torch.int32), out32.view(F, N_STATES, 2).to(torch.int32)
        diff = ((a[..., 0] * 256 + a[..., 1]) - (b[..., 0] * 256 + b[..., 1])).abs()
        codes = {"entries": int(F) * N_STATES, "differing_fraction": float((diff != 0).float().mean().item()),
                 "max_abs_code_difference": int(diff.max().item()),
                 "note": "2-byte codes of the throughput mode (the headline) against the parity mode's (= the reference's bytes), end to end from PCM"}
        del a, b, diff, out32
    return {"metric": "acoustic frames/sec (MFCC+GMM log-lik -> LNA), parity mode", "value": F / (ms * 1e-3), "unit": "frames/s",
            "throughput_mode_codes_vs_parity_mode": codes,
            "dtype": "f64", "steps": 2, "ms_per_step": ms,
            "config": {"workload": "%d utterances x 10 s of the headline workload, F64 parity arithmetic (gmm_diag_f64 + lna_f64: the "
                                   "reference's operations in double, LNA bytes identical to the reference's)" % n, "frames": F},
            "e2e": {"value": F / (ms_e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": ns * 2, "d2h_bytes_per_step": F * rec, "ms_per_step": ms_e}}


def sub_streaming(args, local, model2):
    """Streaming regime: the decoder's per-frame feed -- a call of akugpu_gmm_logprobs on F <= 16 frames
    sweeps the whole parameter image.  Latency per call through the C ABI with HOST buffers, the kernel's sweep rate, and
    the L2 / HBM read rates of that very buffer it is measured against."""
    import ctypes as C
    from aaltoasr_b200 import AkuGpu, synth
    hbm_peak, hbm_src = measured_peaks()
    out = {"metric": "microseconds per per-frame scoring call (host features in, host log-probs out)", "unit": "us", "higher_is_better": False,
           "kernel": "gmm_stream_kernel (tcgen05, GEMM turned around: components x frames; one launch over all SMs; results and the "
                     "completion flag written into mapped host memory)", "models": {}}
    eng = AkuGpu(local)
    eng.frontend_load_config_text(synth.mfcc39_config(SAMPLE_RATE))
    pcm = np.concatenate([synth.synth_audio(2000 + i, UTT_SAMPLES, SAMPLE_RATE) for i in range(2)])
    feats, _ = eng.features(pcm, np.arange(3, dtype=np.int64) * UTT_SAMPLES, dtype=np.float32)
    lib, h = eng._lib, eng._h
    for name, model in (("5000x16", model2), ("10000x32", synth.synth_diag_model(4999, feats.astype(np.float64), C4_STATES, C4_MIX))):
        eng.model_load_diag(model["mix_offsets"], model["mix_gauss"], model["mix_weight"], model["means"], model["covs"])
        S = eng.num_states
        rec = {}
        for F in (1, 4, 8, 16):
            x = np.ascontiguousarray(feats[100:100 + F])
            ob = np.empty((F, S), dtype=np.float32)
            px, po = C.c_void_p(x.ctypes.data), C.c_void_p(ob.ctypes.data)
            tiny = C.c_double(1e-30)
            for _ in range(100):
                lib.akugpu_gmm_logprobs(h, px, 0, F, 0, tiny, po)
            n = 2000 if F <= 8 else 500
            t0 = time.perf_counter()
            for _ in range(n):
                lib.akugpu_gmm_logprobs(h, px, 0, F, 0, tiny, po)
            rec["F=%d" % F] = 1e6 * (time.perf_counter() - t0) / n
        # the same calls timed inside the library (no ctypes overhead), one launch per call
        native = {"F=%d" % F: eng.stream_latency(feats[100:100 + F], n_calls=1000) for F in (1, 4, 8)}
        # resident scorer (akugpu_stream_open): the parameter image stays in shared memory, a call is a message
        resident = {"us_per_call": {}, "native": {}}
        t0 = time.perf_counter()
        eng.stream_open(200.0)
        resident["open_ms"] = 1e3 * (time.perf_counter() - t0)    # launch + the CTAs' fill of their resident tiles begun
        try:
            rows_p = C.POINTER(C.c_float)()
            for F in ((1, 4, 8, 16) if name == "5000x16" else (1, 4, 8)):
                x = np.ascontiguousarray(feats[100:100 + F])
                px = C.c_void_p(x.ctypes.data)
                for _ in range(200):
                    lib.akugpu_stream_logprobs(h, px, F, tiny, C.byref(rows_p))
                n = 4000
                t0 = time.perf_counter()
                for _ in range(n):
                    lib.akugpu_stream_logprobs(h, px, F, tiny, C.byref(rows_p))
                resident["us_per_call"]["F=%d" % F] = 1e6 * (time.perf_counter() - t0) / n
                resident["native"]["F=%d" % F] = eng.stream_latency(x, n_calls=4000)
                if F == 1:
                    st = eng.stream_stats()
                    resident["device_us_F=1"] = {"command_to_A_built": st["device_ns_features"] / 1e3, "to_results_stored": st["device_ns_stored"] / 1e3,
                                                 "to_last_cta_counted": st["device_ns_fenced"] / 1e3, "to_rows_fenced": st["device_ns_call"] / 1e3}
            resident["launches"] = eng.stream_stats()["launches"]
        finally:
            eng.stream_close()
        resident["kernel"] = ("gmm_resident_kernel: one CTA per SM keeps its component tiles of B' in shared memory (40 KB per tile; what does not fit is "
                              "streamed from L2 per call); a call = self-validating packets in mapped host memory, one per CTA; results into mapped "
                              "host memory, one system-scope fence by the last CTA")
        eng.set_streaming(False)
        x = np.ascontiguousarray(feats[100:101])
        ob = np.empty((1, S), dtype=np.float32)
        for _ in range(20):
            eng.gmm_logprobs(x, tiny=1e-30, out=ob)
        t0 = time.perf_counter()
        for _ in range(300):
            eng.gmm_logprobs(x, tiny=1e-30, out=ob)
        general = 1e6 * (time.perf_counter() - t0) / 300
        eng.set_streaming(True)
        p = eng.stream_probe()
        out["models"][name] = {
            "us_per_call": rec, "us_per_call_native_loop": native, "resident": resident,
            "us_per_call_general_path_F=1": general, "calls_per_s_F=1": 1e6 / resident["native"]["F=1"]["mean_us"],
            "x_realtime_per_frame_loop": 1e6 / resident["native"]["F=1"]["mean_us"] / 125.0,
            "image_bytes": p["image_bytes"],
            "kernel_us_isolated_launch": 1e6 * p["kernel_s_l2"], "kernel_us_after_l2_flush": 1e6 * p["kernel_s_hbm"],
            "kernel_us_in_launch_train": 1e6 * p["kernel_s_train"],
            "sweep_GBps_in_launch_train": p["kernel_GBps_train"], "sweep_GBps_after_l2_flush": p["kernel_GBps_hbm"],
            "l2_read_probe_GBps": p["probe_GBps_l2"], "hbm_read_probe_GBps_same_buffer": p["probe_GBps_hbm"],
            "sweep_frac_of_l2_probe": p["kernel_GBps_train"] / p["probe_GBps_l2"],
            "sweep_frac_of_hbm_peak_cold": p["kernel_GBps_hbm"] / hbm_peak, "hbm_peak_GBps": hbm_peak, "hbm_peak_source": hbm_src,
            "sm_mhz": p["sm_mhz_train"],
            "regime": "the image is L2-resident from the second call on (126 MB L2): the launch-train figure is an L2 sweep; the "
                      "after-flush figure is the cold (HBM) sweep of a single launch, launch latency included"}
    out["value"] = out["models"]["5000x16"]["resident"]["native"]["F=1"]["mean_us"]
    out["value_is"] = "akugpu_stream_logprobs inside an open session (resident scorer), F = 1, timed inside the library; us_per_call = one launch per call"
    eng.close()
    return out


def sub_full_cov_cpu_baseline(model_sub, feats, n_states_full):
    """The reference's FullCovarianceGaussian path (aku::HmmSet from oracle/_ref) on all host cores, on a sub-model
    (its loader spends ~18 ms per full-covariance Gaussian); cost per frame is linear in the number of Gaussians."""
    import multiprocessing as mp
    from aaltoasr_b200 import formats
    from oracle import ref
    if not ref.available():
        return {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built on this box"}
    cores = os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as tmp:
        base = os.path.join(tmp, "full")
        formats.write_model(base, **model_sub)
        np.save(os.path.join(tmp, "x.npy"), feats)
        with mp.get_context("spawn").Pool(cores) as pool:
            res = pool.map(_ref_fullcov_worker, [(base, os.path.join(tmp, "x.npy"))] * cores)
    n_sub = len(model_sub["mix_offsets"]) - 1
    frames = sum(r[1] for r in res)
    secs = max(r[0] for r in res)
    scale = n_sub / float(n_states_full)
    return {"value": frames / secs * scale, "unit": "frames/s", "cores": cores, "kind": "reference",
            "sample": "%d frames per core on %d cores against %d of the %d states (aku::HmmSet, FullCovarianceGaussian exponential "
                      "form; %.2f s), scaled by %d/%d: cost per frame is linear in the Gaussians; model load excluded" % (
                          res[0][1], cores, n_sub, n_states_full, secs, n_sub, n_states_full)}


def _ref_fullcov_worker(a):
    base, xpath = a
    from oracle import ref
    M = ref.Model(base)
    x = np.load(xpath)
    t0 = time.perf_counter()
    M.state_likelihoods(x)
    dt = time.perf_counter() - t0
    M.close()
    return dt, int(x.shape[0])


# --------------------------------------------------------------------------------------
# BASELINE.json configs[3]: 100 h sharded by utterance over the GPUs of the node, 10000 x 32
C4_STATES, C4_MIX = 10000, 32
C4_CLIPS, C4_CLIP_SAMPLES = 32, 15 * SAMPLE_RATE


def c4_utterance_lengths(n_total):
    """Jittered utterance lengths U(5, 15) s, the same table on every rank."""
    return (np.random.default_rng(4000).uniform(5.0, 15.0, n_total) * SAMPLE_RATE).astype(np.int64)


def c4_audio(clips, lens, ids):
    """Utterance u = a window of clip u % 32: fully determined by the global utterance id, so any rank can rebuild any
    utterance (the 1-vs-N checksum check needs that)."""
    total = int(lens[ids].sum())
    pcm = np.empty(total, dtype=np.int16)
    uo = np.zeros(len(ids) + 1, dtype=np.int64)
    o = 0
    for k, u in enumerate(int(x) for x in ids):
        n = int(lens[u])
        start = (u * 7919) % (C4_CLIP_SAMPLES - n + 1)
        pcm[o:o + n] = clips[u % C4_CLIPS][start:start + n]
        o += n
        uo[k + 1] = o
    return pcm, uo


def bench_config4(args, torch, dist, rank, world, local, utts_per_gpu):
    """One pass of the hot path over this rank's shard in every payload mode; see the module docstring."""
    from contextlib import nullcontext
    from aaltoasr_b200 import AkuGpu, F32, multigpu as mg, synth
    from aaltoasr_b200.engine import DevPtr
    from aaltoasr_b200.partition import gather_utterance_table
    t_setup = time.time()
    eng = AkuGpu(local)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    eng.frontend_load_config_text(synth.mfcc39_config(SAMPLE_RATE))
    n_total = utts_per_gpu * world
    lens = c4_utterance_lengths(n_total)
    clips = [synth.synth_audio(4000 + i, C4_CLIP_SAMPLES, SAMPLE_RATE) for i in range(C4_CLIPS)]
    collectives = []
    # ---- model: built by rank 0, broadcast as one packed buffer
    model = None
    if rank == 0:
        pcm16 = np.concatenate(clips[:16])
        feats, _ = eng.features(pcm16, np.arange(17, dtype=np.int64) * C4_CLIP_SAMPLES, dtype=np.float64)
        model = synth.synth_diag_model(4999, feats, C4_STATES, C4_MIX)
    if world > 1:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        model, nbytes = mg.broadcast_model(model, 0)
        torch.cuda.synchronize()
        collectives.append({"op": "broadcast (packed model arrays, rank 0 -> all)", "bytes": nbytes, "ms": 1e3 * (time.perf_counter() - t0)})
    t0 = time.time()
    eng.model_load_diag(model["mix_offsets"], model["mix_gauss"], model["mix_weight"], model["means"], model["covs"])
    t_pack = time.time() - t0
    S = eng.num_states
    rec = S * LNABYTES
    # ---- frame counts: every rank looks at every world-th utterance, all-gather, partition redundantly
    ids_seen = np.arange(rank, n_total, world)
    counts_seen = np.array([eng.num_frames(int(lens[u])) for u in ids_seen], dtype=np.int64)
    if world > 1:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n_frames = mg.gather_frame_counts(ids_seen, counts_seen, n_total)
        collectives.append({"op": "all_gather (int32 n_frames[utt])", "bytes": int(n_total * 8), "ms": 1e3 * (time.perf_counter() - t0)})
    else:
        n_frames = counts_seen
    parts = mg.partition(n_frames, world, args.split)
    mine = parts[rank]
    pcm, uo = c4_audio(clips, lens, mine)
    my_frames = n_frames[mine]
    fo = np.concatenate([[0], np.cumsum(my_frames)]).astype(np.int64)
    F_mine, F_total = int(fo[-1]), int(n_frames.sum())
    pcm_p = torch.from_numpy(pcm).pin_memory()
    pcm_d = pcm_p.cuda()
    max_frames = 2 * 148 * 128                                     # one chunk of the scorer (two waves)
    plan = mg.GatherPlan(n_frames, parts, max_frames, rec, writer=0)
    slot_bytes = plan.slot_bytes
    log("config 4 rank %d: %d utterances, %d frames (%.1f GB of LNA), %d sub-batches, model packed in %.1f s, setup %.1f s" % (
        rank, len(mine), F_mine, F_mine * rec / 1e9, len(plan.sched[rank]), t_pack, time.time() - t_setup))

    src_chk = np.zeros(len(mine), dtype=np.uint64)

    trace = os.environ.get("AKUGPU_BENCH_TRACE")

    def produce_into(want_chk):
        def produce(u0, u1, out):
            a, b = int(uo[u0]), int(uo[u1])
            t0 = time.perf_counter()
            _, _, uc = eng.phone_probs(pcm_d[a:b], uo[u0:u1 + 1] - uo[u0], precision=F32, lnabytes=LNABYTES, out=out,
                                       utt_checksums=want_chk)
            if trace:
                log("rank %d produce utts [%d, %d) %d frames: %.2f ms" % (rank, u0, u1, int(fo[u1] - fo[u0]), 1e3 * (time.perf_counter() - t0)))
            if want_chk:
                src_chk[u0:u1] = uc
        return produce

    def timed(fn):
        return timed_region(torch, dist, world, stream, fn, 1)

    res = {}
    # ---- (1) device-resident: records into two rotating device slots
    dev_slots = [torch.empty(slot_bytes, dtype=torch.uint8, device="cuda") for _ in range(2)]
    # warm-up: the largest sub-batch first (every scratch buffer of the library reaches its final size), then two more
    sched = plan.sched[rank]
    big = max(range(len(sched)), key=lambda k: sched[k][3])
    for k in [big] + list(range(min(2, len(sched)))):
        produce_into(False)(sched[k][0], sched[k][1], dev_slots[k & 1])
    eng.stage_times_reset(True)
    l0 = eng.launch_count()
    ms_res = timed(lambda: mg.run_writers(produce_into(False), my_frames, max_frames, dev_slots))
    launches = eng.launch_count() - l0
    st = eng.stage_times()
    eng.stage_times_reset(False)
    # ---- (2) per-rank writers end to end: pinned host PCM in, records to two rotating pinned host slots
    e2e_chunks = 4 if world <= 2 else 2            # pinned host memory per rank: two slots of e2e_chunks x 0.76 GB
    host_slots = [torch.empty(e2e_chunks * slot_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)]

    def produce_host(u0, u1, out):
        a, b = int(uo[u0]), int(uo[u1])
        eng.phone_probs(pcm_p[a:b], uo[u0:u1 + 1] - uo[u0], precision=F32, lnabytes=LNABYTES, out=out)
    sched_e = mg.sub_batches(my_frames, e2e_chunks * max_frames)
    big = max(range(len(sched_e)), key=lambda k: sched_e[k][3])
    produce_host(sched_e[big][0], sched_e[big][1], host_slots[0])                                        # warm-up
    ms_e2e = timed(lambda: mg.run_writers(produce_host, my_frames, e2e_chunks * max_frames, host_slots))
    del host_slots
    res.update({"value": F_total / (ms_res * 1e-3), "ms_per_step": ms_res,
                "e2e": {"value": F_total / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(pcm.nbytes),
                        "d2h_bytes_per_step": int(F_mine * rec), "ms_per_step": ms_e2e,
                        "d2h_GBps_per_gpu": F_mine * rec / (ms_e2e * 1e-3) / 1e9,
                        "note": "per-rank writers (the reference's -B/-I model): every rank streams its own records to pinned host memory, "
                                "%d-chunk calls, two rotating host slots" % e2e_chunks}})
    # ---- (3) LNA gather to the writer rank
    gather = None
    table_ok = None
    if world > 1:
        gather = {}
        g_tok, g_free = dist.new_group(), dist.new_group()
        for g in (g_tok, g_free):                                  # communicators created now, by all ranks together
            dist.barrier(group=g)
        sink_eng = AkuGpu(local) if rank == 0 else None
        sink_stream = torch.cuda.Stream() if rank == 0 else None
        if rank == 0:
            sink_eng.set_stream(sink_stream.cuda_stream)
        base = np.concatenate([[0], np.cumsum([int(plan.fo[r][-1]) for r in plan.senders])]).astype(np.int64)
        sink_base = {r: int(base[i]) for i, r in enumerate(plan.senders)}
        sink_fo = np.concatenate([plan.fo[r][:-1] + sink_base[r] for r in plan.senders] + [[int(base[-1])]]).astype(np.int64)
        sink_ids = np.concatenate([plan.parts[r] for r in plan.senders])
        recv_bytes = int(base[-1]) * rec

        def sink(r, slot, f0, n):
            sink_eng.checksum_update(slot, sink_base[r] + f0, n)

        def sink_scope():
            return torch.cuda.stream(sink_stream)
        sunk = {}
        copy_eng, copy_stream, copy_ev, copy_log = None, None, [None, None], []
        if rank != 0:
            copy_eng = AkuGpu(local)
            copy_stream = torch.cuda.Stream()
            copy_eng.set_stream(copy_stream.cuda_stream)
        idx = {r: i for i, r in enumerate(plan.senders)}
        nslots = 2
        for mode in ("p2p_store", "p2p_copy", "nccl"):
            shared = peer_base = recv_slots = None
            if mode != "nccl":
                handle = torch.zeros(64, dtype=torch.uint8, device="cuda")
                if rank == 0:
                    shared, h = eng.shared_alloc(len(plan.senders) * nslots * slot_bytes)
                    handle.copy_(torch.frombuffer(bytearray(h), dtype=torch.uint8))
                dist.broadcast(handle, 0)
                if rank != 0:
                    peer_base = eng.shared_open(handle.cpu().numpy().tobytes())
                root = shared if rank == 0 else peer_base

                def peer_slot(r, j, root=root):
                    return DevPtr(int(root) + (idx[r] * nslots + j) * slot_bytes)
                token = torch.zeros(1, dtype=torch.int64, device="cuda")
                if mode == "p2p_store":
                    prod, tok_scope = produce_into(False), nullcontext
                else:
                    kcount = [0]

                    def prod(u0, u1, peer_out):
                        # records into a local slot (scored on the compute stream), then a copy engine carries them over
                        # NVLink on its own stream while the next sub-batch is scored; the token follows the copy
                        k = kcount[0]
                        kcount[0] += 1
                        if copy_ev[k & 1] is not None:
                            stream.wait_event(copy_ev[k & 1])          # the slot's previous copy has left
                        produce_into(False)(u0, u1, dev_slots[k & 1])
                        nb = int(fo[u1] - fo[u0]) * rec
                        t0e = torch.cuda.Event(enable_timing=True)
                        t0e.record(copy_stream)
                        copy_eng.copy_async(peer_out, dev_slots[k & 1], nb)
                        copy_ev[k & 1] = torch.cuda.Event(enable_timing=True)
                        copy_ev[k & 1].record(copy_stream)
                        copy_log.append((nb, t0e, copy_ev[k & 1]))

                    def tok_scope():
                        return torch.cuda.stream(copy_stream)

                def run(prod=prod, tok_scope=tok_scope, peer_slot=peer_slot, token=token):
                    with torch.cuda.stream(stream):
                        mg.gather_p2p(plan, rank, produce_into(False) if rank == 0 else prod, sink, dev_slots, peer_slot, nslots, g_tok, g_free,
                                      token, sink_scope if rank == 0 else nullcontext, nullcontext if rank == 0 else tok_scope)
                        if rank == 0:
                            stream.wait_stream(sink_stream)          # the timed region ends after the sink's last checksum
                        elif copy_stream is not None:
                            stream.wait_stream(copy_stream)
            else:
                recv_slots = {r: [torch.empty(slot_bytes, dtype=torch.uint8, device="cuda") for _ in range(2)] for r in plan.senders} \
                    if rank == 0 else None

                def run(recv_slots=recv_slots):
                    with torch.cuda.stream(stream):
                        mg.gather_nccl(plan, rank, produce_into(False), sink, dev_slots, recv_slots, None,
                                       sink_scope if rank == 0 else nullcontext)
                        if rank == 0:
                            stream.wait_stream(sink_stream)
            if rank == 0:
                sink_eng.checksum_begin(sink_fo, rec)
            ms = timed(run)
            if rank == 0:
                sunk[mode] = sink_eng.checksum_end()
            gather[mode] = {"value": F_total / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms,
                            "bytes_into_writer": recv_bytes, "payload_GBps_into_writer": recv_bytes / (ms * 1e-3) / 1e9}
            if mode != "nccl":
                torch.cuda.synchronize(); dist.barrier()
                if rank != 0:
                    eng.shared_release(peer_base)
                dist.barrier()
                if rank == 0:
                    eng.shared_release(shared)
            recv_slots = None
        # NVLink rate of the copy-engine transfers themselves (sender side, events around every copy), all ranks' medians
        torch.cuda.synchronize()
        rates = sorted(nb / (a.elapsed_time(b) * 1e-3) / 1e9 for nb, a, b in copy_log if nb > (64 << 20))
        med = torch.tensor([rates[len(rates) // 2] if rates else 0.0], device="cuda")
        allmed = [torch.zeros_like(med) for _ in range(world)]
        dist.all_gather(allmed, med)
        gather["p2p_copy"]["copy_GBps_per_sender_median"] = [float(x.item()) for x in allmed[1:]]
        gather["p2p_copy"]["copy_GBps_sum_over_senders"] = float(sum(x.item() for x in allmed[1:]))
        if copy_eng:
            copy_eng.close()
        gather["p2p_store"]["how"] = ("lna_f32_rows of every rank stores its records straight into the writer's rotating buffer (CUDA-IPC mapped "
                                      "peer memory over NVLink): epilogue + gather in one kernel; NCCL carries two 8-byte tokens per sub-batch")
        gather["p2p_copy"]["how"] = ("records into a local slot, then one asynchronous copy-engine transfer into the writer's mapped rotating buffer "
                                     "while the SMs score the next sub-batch; same tokens")
        gather["nccl"]["how"] = "records into a local send slot, ncclSend -> ncclRecv into the writer's rotating slots"
        gather["note"] = ("payload_GBps_into_writer = bytes / step time: it is bounded by the rate at which the senders PRODUCE records "
                          "(10000 states x 2 bytes per frame), not by NVLink, unless the writer's ingest saturates")
        # ---- source checksums: one more (untimed) local pass with the per-utterance checksums switched on
        mg.run_writers(produce_into(True), my_frames, max_frames, dev_slots)
        # ---- the table: (utterance, n_frames, source checksum) all-gathered; the writer compares what arrived
        torch.cuda.synchronize(); t0 = time.perf_counter()
        tf, tc, owner = gather_utterance_table(mine, my_frames, src_chk, n_total)
        collectives.append({"op": "all_gather (utterance, n_frames, checksum) table", "bytes": int(n_total * 24), "ms": 1e3 * (time.perf_counter() - t0)})
        collectives.append({"op": "LNA gather to rank 0 (p2p stores / ncclSend+Recv), per step", "bytes": recv_bytes})
        if rank == 0:
            gather["sink_checksums_equal_source"] = bool(all(np.array_equal(sunk[m], tc[sink_ids]) for m in sunk))
            table_ok = bool(np.array_equal(tf, n_frames) and all(owner[i] == r for r in range(world) for i in parts[r][:4]))
        sink_eng and sink_eng.close()
    else:
        # one GPU: the per-utterance checksums of the run itself (two differently batched passes must agree)
        mg.run_writers(produce_into(True), my_frames, max_frames, dev_slots)
        tc = src_chk.copy()
    # ---- 1-GPU-vs-N-GPU: rank 0 alone rescoring a sample of ALL utterances, batched differently
    check = None
    if rank == 0:
        sample = np.arange(0, n_total, max(1, n_total // 256))[:256]
        spcm, suo = c4_audio(clips, lens, sample)
        _, _, sc = eng.phone_probs(spcm, suo, precision=F32, lnabytes=LNABYTES, discard=True, utt_checksums=True)
        check = {"checksum_1_vs_N_equal": bool(np.array_equal(sc, tc[sample])), "checked_utterances": int(len(sample)),
                 "how": "rank 0 alone rescoring every %d-th utterance of the whole list in one differently batched call; per-utterance "
                        "order-sensitive checksums (akugpu_phone_probs_ex) against the N-rank run's table" % max(1, n_total // 256)}
    line = None
    if rank == 0:
        line = {"metric": "acoustic frames/sec (MFCC+GMM log-lik -> LNA)", "unit": "frames/s", "n_gpus": world, "steps": 1, "warmup": 1,
                "higher_is_better": True, "scaling": "weak", "dtype": "f16x2->f32", "data": "synthetic",
                "config": {"workload": "%.1f h @16 kHz sharded by utterance: %d utterances of 5-15 s (%d per GPU), 39-dim MFCC+d+dd, "
                                       "10000-state x 32-mix diag GMM, 2-byte LNA" % (lens.sum() / SAMPLE_RATE / 3600.0, n_total, utts_per_gpu),
                           "utterances": int(n_total), "frames": F_total, "split": args.split, "sub_batch_frames": max_frames,
                           "lna_bytes_total": int(F_total * rec), "l2_policy": "every sub-batch writes 0.76 GB of records: far beyond L2"},
                "gpu_launches": int(launches),
                "roofline": scorer_roofline(eng, st, F_mine, 1, ms_res, C4_STATES, C4_MIX),
                "gather": gather, "collectives": collectives, "table_ok": table_ok}
        line.update(res)
        line.update(check)
    eng.close()
    return line


def host_d2h_probe(torch, dist, rank, world, stream):
    """All ranks copy device -> pinned host at the same time: what the host side of PCIe gives N GPUs together."""
    out = {}
    for mb in (64, 1024):
        n = mb << 20
        src = torch.empty(n, dtype=torch.uint8, device="cuda")
        dst = torch.empty(n, dtype=torch.uint8).pin_memory()
        dst.copy_(src, non_blocking=True)
        reps = 4 if mb >= 1024 else 16
        ms = timed_region(torch, dist, world, stream, lambda: dst.copy_(src, non_blocking=True), reps)
        out["d2h_%dMB_aggregate_GBps" % mb] = world * n * reps / (ms * 1e-3) / 1e9
        ms = timed_region(torch, dist, world, stream, lambda: src.copy_(dst, non_blocking=True), reps)
        out["h2d_%dMB_aggregate_GBps" % mb] = world * n * reps / (ms * 1e-3) / 1e9
        del src, dst
    try:
        out["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
        out["cpus_visible"] = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return out


class Watchdog:
    """A sub-record that hangs (a collective waiting for a rank that failed) must not take the headline with it: when the
    timer fires, rank 0 prints the line it has -- the sub-record marked as timed out -- and every rank leaves."""

    def __init__(self, seconds, on_fire):
        self.t = threading.Timer(seconds, self._fire)
        self.t.daemon = True
        self.on_fire = on_fire

    def _fire(self):
        try:
            self.on_fire()
        finally:
            sys.stdout.flush()
            os._exit(0)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.t.cancel()
        return False


def _claim_stdout():
    """stdout carries exactly ONE JSON line: whatever else writes to file descriptor 1 (NCCL's version banner, a library's
    printf) is sent to stderr, and print() keeps a private copy of the original descriptor."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--utts", type=int, default=None, help="utterances per GPU (default: the BASELINE config)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json config to run on its own: 2 = 1000 utts, 5000x16 (default, the metric's config, with the "
                         "other configs as sub_records); 3 = feature-only sweep; 4 = 100 h over the GPUs (4500 utts per GPU), "
                         "10000x32, with the LNA gather; 5 = 2000-state x 16-mix full-covariance pool")
    ap.add_argument("--split", default="lpt", choices=["lpt", "reference"],
                    help="config 4: utterance partition (lpt = by frame count; reference = aku/Recipe.cc:63-115, what -B N -I i gives)")
    ap.add_argument("--c4-utts", type=int, default=None, help="config 4 sub-record: utterances per GPU (default 4500 at N > 1, 300 at N = 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-records", action="store_true")
    ap.add_argument("--sub-timeout", type=int, default=420, help="seconds the multi-GPU sub-records may take before they are abandoned")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from aaltoasr_b200 import AkuGpu, F32, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the accelerated path has no CPU fallback")
    torch.cuda.set_device(local)
    numa_cpus = bind_to_gpu_numa(torch, local) if world > 1 else None
    if world > 1:
        # NCCL prints its version banner / debug lines to stdout; stdout carries exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.config in (3, 5):
        if world > 1:
            raise SystemExit("bench.py --config %d is a single-GPU configuration" % args.config)
        line = (bench_feature_sweep if args.config == 3 else bench_full_cov)(args, local)
        print(json.dumps(line))
        return
    if args.config == 4:
        line = bench_config4(args, torch, dist, rank, world, local, args.utts or args.c4_utts or 4500)
        if rank == 0:
            line["vs_baseline"] = None
            print(json.dumps(line))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    n_utts = args.utts if args.utts else N_UTTS

    eng = AkuGpu(local)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)        # torch.cuda.Event then times the library's launches
    eng.frontend_load_config_text(synth.mfcc39_config(SAMPLE_RATE))
    model = make_model(eng)
    eng.model_load_diag(model["mix_offsets"], model["mix_gauss"], model["mix_weight"], model["means"], model["covs"])
    t0 = time.time()
    pcm = make_audio(rank, n_utts)
    uo = np.arange(n_utts + 1, dtype=np.int64) * UTT_SAMPLES
    fo = eng.frame_offsets(uo)
    F = int(fo[-1])
    rec = N_STATES * LNABYTES
    log("rank %d: %d utterances, %d frames, audio generated in %.1f s" % (rank, n_utts, F, time.time() - t0))

    pcm_d = torch.from_numpy(pcm).cuda()
    out_d = torch.empty((F, rec), dtype=torch.uint8, device="cuda")     # 12.5 GB: the LNA of a step stays in HBM
    # e2e buffers: pinned PCM; LNA drained through one pinned buffer per sub-batch (a writer would stream it out)
    # utterances per e2e call: 250 on one GPU; fewer per rank when several ranks share the host (pinned memory per rank
    # = one sub-batch of LNA: 3.1 GB at 250 utterances)
    sub = min(n_utts, int(os.environ.get("AKUGPU_BENCH_SUB", str(max(50, 250 // max(1, world // 2))))))
    pcm_p = torch.from_numpy(pcm).pin_memory()
    out_p = torch.empty((int(fo[sub]), rec), dtype=torch.uint8).pin_memory()

    def step_resident():
        eng.phone_probs(pcm_d, uo, precision=F32, lnabytes=LNABYTES, out=out_d)

    def step_e2e():
        for u0 in range(0, n_utts, sub):
            u1 = min(n_utts, u0 + sub)
            eng.phone_probs(pcm_p[u0 * UTT_SAMPLES:u1 * UTT_SAMPLES], uo[u0:u1 + 1] - uo[u0], precision=F32,
                            lnabytes=LNABYTES, out=out_p[:int(fo[u1] - fo[u0])])

    def timed(fn, steps):
        return timed_region(torch, dist, world, stream, fn, steps)

    for _ in range(args.warmup):
        step_resident()
    clocks = ClockSampler(local)
    clocks.start()
    eng.stage_times_reset(True)
    l0 = eng.launch_count()
    ms_res = timed(step_resident, args.steps)
    launches = eng.launch_count() - l0
    st = eng.stage_times()
    eng.stage_times_reset(False)
    clk = clocks.stop()
    for _ in range(min(args.warmup, 1)):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # the PCIe link itself (device -> pinned host, one large async copy): the roofline of the e2e number
    probe_bytes = min(int(out_p.numel()), 2 << 30)
    src_probe = torch.empty(probe_bytes, dtype=torch.uint8, device="cuda")
    dst_probe = out_p.view(-1)[:probe_bytes]
    def d2h_probe():
        dst_probe.copy_(src_probe, non_blocking=True)
    d2h_probe()
    ms_link = timed(d2h_probe, 3) / 3
    d2h_gbs = probe_bytes / (ms_link * 1e-3) / 1e9
    del src_probe

    total_frames = F * world
    value = total_frames * args.steps / (ms_res * 1e-3)
    e2e_val = total_frames * args.steps / (ms_e2e * 1e-3)

    line = None
    if rank == 0:
        roofline = scorer_roofline(eng, st, F, args.steps, ms_res, N_STATES, N_MIX)
        line = {
            "metric": "acoustic frames/sec (MFCC+GMM log-lik -> LNA)", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16x2->f32", "data": "synthetic",
            "config": {"workload": WORKLOAD if n_utts == N_UTTS else WORKLOAD.replace("1000 utterances", "%d utterances" % n_utts),
                       "utterances_per_gpu": n_utts, "frames_per_gpu": F, "precision": "F32 throughput mode",
                       "l2_policy": "inputs+outputs per step (%.1f GB) exceed L2; no flush needed" % (F * rec / 1e9),
                       "parallelism": "utterance shards, one process per GPU, no data-path collective"},
            "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": int(pcm.nbytes),
                    "d2h_bytes_per_step": int(F * rec), "ms_per_step": ms_e2e / args.steps,
                    "d2h_GBps": F * rec / (ms_e2e / args.steps * 1e-3) / 1e9, "pcie_d2h_peak_GBps": d2h_gbs,
                    "pcie_frac": (F * rec / (ms_e2e / args.steps * 1e-3) / 1e9) / d2h_gbs,
                    "note": "LNA records leave the device at the PCIe rate; %d-utterance calls%s" % (
                        sub, "; rank bound to the %d CPUs next to its GPU" % numa_cpus if numa_cpus else "")},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": None, "sub_records": {}}

    # ---------------- sub-records: the other BASELINE configs, in the same driver-run line ----------------
    subs = {}
    if not args.no_sub_records:
        def guarded(name, fn):
            try:
                t0 = time.time()
                r = fn()
                if r is not None:
                    r["wall_s"] = time.time() - t0
                return r
            except Exception as e:      # a sub-record never takes the headline down; the failure is recorded
                log("sub-record %s failed: %r" % (name, e))
                return {"failed": repr(e)}
        if world == 1:
            subs["parity_mode_f64"] = guarded("parity_mode_f64", lambda: sub_parity_mode(args, torch, eng, stream, pcm_d, pcm_p, uo, fo))
    del out_d, out_p, pcm_d, pcm_p
    eng.close()
    torch.cuda.empty_cache()
    if not args.no_sub_records:
        sub_args = argparse.Namespace(**vars(args))
        sub_args.steps, sub_args.warmup, sub_args.utts = 3, 3, None
        if world == 1:
            subs["config3_feature_sweep"] = guarded("config3", lambda: bench_feature_sweep(sub_args, local))
            subs["config5_full_covariance"] = guarded("config5", lambda: bench_full_cov(sub_args, local, cpu_baseline=not args.no_cpu_baseline))
            subs["config4_model_1gpu"] = guarded("config4", lambda: bench_config4(sub_args, torch, dist, rank, world, local, args.c4_utts or 300))
            subs["streaming"] = guarded("streaming", lambda: sub_streaming(sub_args, local, model))
        else:
            def bail():
                if rank == 0:
                    subs.setdefault("config4", {"failed": "timed out after %d s (see stderr)" % args.sub_timeout})
                    line["sub_records"] = subs
                    print(json.dumps(line))
            with Watchdog(args.sub_timeout, bail):
                subs["host_d2h"] = guarded("host_d2h", lambda: host_d2h_probe(torch, dist, rank, world, stream))
                subs["config4"] = guarded("config4", lambda: bench_config4(sub_args, torch, dist, rank, world, local, args.c4_utts or 4500))

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import ref
                if ref.available():
                    cores = os.cpu_count() or 1
                    with tempfile.TemporaryDirectory() as tmp:
                        cfg, base, wavs = reference_fixture(tmp, model, min(cores, 16))
                        fps, secs = run_reference_sample(cfg, base, wavs, cores, 384)
                    cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference",
                           "sample": "384 frames per core on %d cores, %.1f s (reference FeatureGenerator + HmmSet from "
                                     "oracle/_ref, model load excluded)" % (cores, secs)}
                else:
                    cpu = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference",
                           "sample": "oracle/_ref not built on this box"}
            except Exception as e:   # the baseline is reported, never required
                cpu = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}
        line["cpu_baseline"] = cpu
        line["sub_records"] = subs
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
