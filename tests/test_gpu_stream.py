"""GPU tests of the streaming-regime scorer (gmm_stream_kernel): the decoder's per-frame feed
(decoder/decode-stream.cc:191-207 -> OneFrameAcoustics::set) -- a few frames against the whole model in one launch."""
import time

import numpy as np
import pytest

from aaltoasr_b200 import F32, F64, synth
from oracle import oracle_np

pytestmark = pytest.mark.gpu


def load_model(engine, m):
    engine.model_load_diag(m["mix_offsets"], m["mix_gauss"], m["mix_weight"], m["means"], m["covs"])


def both_paths(engine, fn):
    engine.set_streaming(True)
    l0 = engine.launch_count()
    a = fn()
    n_stream = engine.launch_count() - l0
    engine.set_streaming(False)
    try:
        l0 = engine.launch_count()
        b = fn()
        n_batch = engine.launch_count() - l0
    finally:
        engine.set_streaming(True)
    return a, b, n_stream, n_batch


@pytest.mark.parametrize("F", [1, 2, 7, 9, 16])
def test_streaming_scorer_equals_batch_path_small_model(engine, ref_small, F):
    g = ref_small
    load_model(engine, g["model"])
    assert engine.scorer_in_use() == 3
    x = g["feats"][3:3 + F].astype(np.float32)
    a, b, ns, nb = both_paths(engine, lambda: engine.gmm_score(x, precision=F32))
    assert ns == 1 and nb >= 2                                   # ONE launch against scorer + transpose
    assert np.abs(a.astype(np.float64) - b).max() <= 2e-6        # same products, another summation order in the epilogue
    assert np.abs(a - np.log(g["lik"][3:3 + F])).max() <= 3e-5   # the reference's likelihoods
    # the decoder feed: (float) log(max(l, 1e-30)), floats and doubles in
    a, b, ns, nb = both_paths(engine, lambda: engine.gmm_logprobs(g["feats"][3:3 + F], precision=F32, tiny=1e-30))
    assert ns == 1 and np.abs(a.astype(np.float64) - b).max() <= 2e-6
    want = np.log(np.maximum(g["lik"][3:3 + F], 1e-30))
    assert np.abs(a - want).max() <= 3e-5


def test_streaming_scorer_floor_and_edge_model(engine, ref_edge):
    """The edge-case model is hybrid (ill-conditioned states): the streaming scorer declines, results are the general
    path's.  A plain model with frames far from every Gaussian: the floor of the decoder feed applies."""
    g = ref_edge
    load_model(engine, g["model"])
    assert engine.scorer_in_use() == 5
    x = g["feats"][:4].astype(np.float32)
    a, b, ns, nb = both_paths(engine, lambda: engine.gmm_logprobs(x, precision=F32, tiny=1e-30))
    assert ns == nb and np.array_equal(a, b)


def test_streaming_scorer_floor(engine, ref_small):
    g = ref_small
    load_model(engine, g["model"])
    x = g["feats"][:5].astype(np.float32).copy()
    x[2] += 40.0                                                  # far from everything: likelihood below 1e-30
    a, b, ns, nb = both_paths(engine, lambda: engine.gmm_logprobs(x, precision=F32, tiny=1e-30))
    assert ns == 1
    floor = np.float32(np.log(1e-30))
    assert (a[2] == floor).all() and (b[2] == floor).all()
    assert np.abs(a.astype(np.float64) - b).max() <= 2e-6
    raw, rawb, _, _ = both_paths(engine, lambda: engine.gmm_score(x, precision=F32))       # no floor in the plain score
    assert (raw[2] < floor).all() and (np.abs(raw[2].astype(np.float64) - rawb[2]) <= 2e-6 * np.abs(rawb[2]) + 2e-6).all()


def test_streaming_scorer_device_buffers_and_overflow(engine, ref_small):
    import torch
    g = ref_small
    load_model(engine, g["model"])
    x = g["feats"][10:13].astype(np.float32)
    host = engine.gmm_logprobs(x, precision=F32, tiny=1e-30)
    xd = torch.from_numpy(x).cuda()
    od = torch.empty((3, engine.num_states), dtype=torch.float32, device="cuda")
    l0 = engine.launch_count()
    engine.gmm_logprobs(xd, precision=F32, tiny=1e-30, out=od)
    assert engine.launch_count() - l0 == 1
    assert np.array_equal(od.cpu().numpy(), host)                 # device features are centred in the kernel: same arithmetic
    x64d = torch.from_numpy(g["feats"][10:13]).cuda()
    engine.gmm_logprobs(x64d, precision=F32, tiny=1e-30, out=od)       # the unrounded doubles: differs by the feature rounding only
    assert np.abs(od.cpu().numpy() - host).max() <= 1e-5
    # a feature outside the fp16 range of the scaled terms: the call falls through to the general path (bf16x3 redo)
    bad = x.copy()
    bad[1, 3] = 3.0e4
    l0 = engine.launch_count()
    got = engine.gmm_score(bad, precision=F32)
    assert engine.launch_count() - l0 >= 4
    assert np.isfinite(got).all() and got[1].max() < -1e5
    assert np.abs(got[[0, 2]] - engine.gmm_score(x, precision=F32)[[0, 2]]).max() <= 3e-5
    again = engine.gmm_logprobs(x, precision=F32, tiny=1e-30)     # and the fast path is back, flag cleared
    assert np.array_equal(again, host)


@pytest.fixture(scope="module")
def feats20s(engine):
    engine.frontend_load_config_text(synth.mfcc39_config())
    pcm = np.concatenate([synth.synth_audio(2000 + i, 160000) for i in range(2)])
    feats, _ = engine.features(pcm, np.array([0, 160000, 320000]), dtype=np.float64)
    return feats


@pytest.mark.parametrize("S,K,seed", [(5000, 16, 2999), (10000, 32, 4999)])
def test_streaming_scorer_full_size_models(engine, feats20s, S, K, seed):
    """Config-2 and config-4 models: every frame of a per-frame loop equals the batch path; latency per call reported."""
    model = synth.synth_diag_model(seed, feats20s, S, K)
    load_model(engine, model)
    assert engine.scorer_in_use() == 3
    idx = np.arange(0, 2496, 96)
    x = feats20s[idx].astype(np.float32)
    engine.set_streaming(False)
    try:
        batch = engine.gmm_logprobs(x, precision=F32, tiny=1e-30)
    finally:
        engine.set_streaming(True)
    rows = np.stack([engine.gmm_logprobs(x[i:i + 1], precision=F32, tiny=1e-30)[0] for i in range(len(idx))])
    assert np.abs(rows.astype(np.float64) - batch).max() <= 4e-6
    blk = engine.gmm_logprobs(x[:16], precision=F32, tiny=1e-30)
    assert np.abs(blk.astype(np.float64) - batch[:16]).max() <= 4e-6
    blk = engine.gmm_logprobs(x[:7], precision=F32, tiny=1e-30)              # ragged
    assert np.abs(blk.astype(np.float64) - batch[:7]).max() <= 4e-6
    want = np.log(np.maximum(oracle_np.state_likelihoods(model, x[:2].astype(np.float64)), 1e-30))
    assert np.abs(rows[:2] - want).max() <= 4e-5
    p = engine.stream_probe()
    print("stream probe %d x %d: image %.1f MB; kernel %.2f us (L2-resident, %.0f GB/s) / %.2f us (after L2 flush, %.0f GB/s); "
          "plain read sweep %.0f GB/s (L2) / %.0f GB/s (HBM)" % (S, K, p["image_bytes"] / 1e6, 1e6 * p["kernel_s_l2"], p["kernel_GBps_l2"],
                                                                 1e6 * p["kernel_s_hbm"], p["kernel_GBps_hbm"], p["probe_GBps_l2"], p["probe_GBps_hbm"]))
    print("stream probe %d x %d: SM clock after isolated launches %.0f MHz; train of 200 launches: %.2f us per launch (%.0f GB/s), SM clock %.0f MHz"
          % (S, K, p["sm_mhz_isolated"], 1e6 * p["kernel_s_train"], p["kernel_GBps_train"], p["sm_mhz_train"]))
    import ctypes as C
    lib, h = engine._lib, engine._h
    for F in (1, 4, 8, 16):
        xb = np.ascontiguousarray(x[:F])
        ob = np.empty((F, S), dtype=np.float32)
        px, po = C.c_void_p(xb.ctypes.data), C.c_void_p(ob.ctypes.data)
        call = lambda: lib.akugpu_gmm_logprobs(h, px, 0, F, 0, C.c_double(1e-30), po)
        for _ in range(50):
            assert call() == 0
        t0 = time.perf_counter()
        n = 1000
        for _ in range(n):
            call()
        us = 1e6 * (time.perf_counter() - t0) / n
        engine.set_streaming(False)
        try:
            for _ in range(20):
                call()
            t0 = time.perf_counter()
            for _ in range(200):
                call()
            us_general = 1e6 * (time.perf_counter() - t0) / 200
        finally:
            engine.set_streaming(True)
        print("streaming scorer %d x %d, F = %d: %.1f us per akugpu_gmm_logprobs call (host buffers, ctypes loop); general path %.1f us"
              % (S, K, F, us, us_general))
        assert us < 1000.0


# ---------------------------------------------------------------------------------------------------------------------
# resident scorer (gmm_resident_kernel): the parameter image stays in shared memory, a call is a message
def _session_pair(engine, fn):
    """fn() through one launch per call, then through the resident kernel: (launch result, session result, launches
    during the session call)."""
    engine.stream_close()
    a = fn()
    engine.stream_open(200.0)
    try:
        fn()                                   # the first message of a session
        l0 = engine.launch_count()
        b = fn()
        n = engine.launch_count() - l0
    finally:
        engine.stream_close()
    return a, b, n


@pytest.mark.parametrize("F", [1, 2, 7, 9, 16])
def test_resident_scorer_returns_the_bits_of_the_launch_path_small_model(engine, ref_small, F):
    g = ref_small
    load_model(engine, g["model"])
    assert engine.scorer_in_use() == 3
    x = g["feats"][3:3 + F].astype(np.float32)
    a, b, n = _session_pair(engine, lambda: engine.gmm_score(x, precision=F32))
    assert n == 0                                                 # no launch on the path of a call
    assert np.array_equal(a, b)
    assert np.abs(b - np.log(g["lik"][3:3 + F])).max() <= 3e-5    # the reference's likelihoods
    a, b, n = _session_pair(engine, lambda: engine.gmm_logprobs(g["feats"][3:3 + F], precision=F32, tiny=1e-30))    # doubles in
    assert n == 0 and np.array_equal(a, b)


def test_resident_scorer_lifecycle(engine, ref_small):
    """Another entry point ends the kernel and the next small call starts it again; the idle timer ends it too; a
    feature outside the fp16 range sends the call down the general path; device buffers are not served by it."""
    import torch
    g = ref_small
    load_model(engine, g["model"])
    x = g["feats"][:5].astype(np.float32).copy()
    x[2] += 40.0                                                  # below the floor of the decoder feed
    engine.stream_close()
    want = engine.gmm_logprobs(x, precision=F32, tiny=1e-30)
    s0 = engine.stream_stats()
    engine.stream_open(30.0)
    try:
        st = engine.stream_stats()
        assert st["want"] and st["live"] and st["launches"] == s0["launches"] + 1
        got = engine.gmm_logprobs(x, precision=F32, tiny=1e-30)
        assert np.array_equal(got, want) and (got[2] == np.float32(np.log(1e-30))).all()
        assert engine.stream_stats()["calls"] == s0["calls"] + 1
        # any other entry point: the kernel is told to end first, the device is whole
        rec = engine.gmm_lna(g["feats"], precision=F64, lnabytes=2)
        assert np.array_equal(rec.reshape(-1), g["lna2"][5:])
        assert not engine.stream_stats()["live"]
        got = engine.gmm_logprobs(x, precision=F32, tiny=1e-30)   # started again by the call
        st = engine.stream_stats()
        assert np.array_equal(got, want) and st["live"] and st["launches"] == s0["launches"] + 2
        # idle timer (30 ms): the kernel is gone, the call notices and starts another
        time.sleep(0.3)
        got = engine.gmm_logprobs(x[:1], precision=F32, tiny=1e-30)
        assert np.array_equal(got, want[:1]) and engine.stream_stats()["launches"] == s0["launches"] + 3
        for i in range(200):                                      # a per-frame loop, every answer checked
            r = engine.gmm_logprobs(x[i % 5:i % 5 + 1], precision=F32, tiny=1e-30)
            assert np.array_equal(r[0], want[i % 5])
        assert engine.stream_stats()["launches"] == s0["launches"] + 3
        # out of the fp16 range: general path (bf16x3 redo), then the session again
        bad = x.copy()
        bad[1, 3] = 3.0e4
        r = engine.gmm_score(bad, precision=F32)
        assert np.isfinite(r).all() and r[1].max() < -1e5
        assert np.array_equal(engine.gmm_logprobs(x, precision=F32, tiny=1e-30), want)
        # device buffers: one launch per call beside a closed kernel
        xd = torch.from_numpy(x).cuda()
        od = torch.empty((5, engine.num_states), dtype=torch.float32, device="cuda")
        engine.gmm_logprobs(xd, precision=F32, tiny=1e-30, out=od)
        assert np.array_equal(od.cpu().numpy(), want) and not engine.stream_stats()["live"]
    finally:
        engine.stream_close()
    assert not engine.stream_stats()["want"] and not engine.stream_stats()["live"]


def test_resident_scorer_refuses_models_it_does_not_serve(engine, ref_edge):
    load_model(engine, ref_edge["model"])                         # hybrid: ill-conditioned states on the FP32 pipe
    with pytest.raises(Exception):
        engine.stream_open(50.0)
    assert not engine.stream_stats()["want"]


@pytest.mark.parametrize("S,K,seed", [(5000, 16, 2999), (10000, 32, 4999)])
def test_resident_scorer_full_size_models(engine, feats20s, S, K, seed):
    """Config-2 model: the whole image is resident (5 tiles per CTA); config-4 model: 3 resident tiles + a ring for the
    other 14.  Every frame of a per-frame loop carries the bits of the launch path; latency per call reported."""
    model = synth.synth_diag_model(seed, feats20s, S, K)
    load_model(engine, model)
    assert engine.scorer_in_use() == 3
    idx = np.arange(0, 2496, 48)
    x = feats20s[idx].astype(np.float32)
    engine.stream_close()
    rows = np.stack([engine.gmm_logprobs(x[i:i + 1], precision=F32, tiny=1e-30)[0] for i in range(len(idx))])
    blk = engine.gmm_logprobs(x[:7], precision=F32, tiny=1e-30)
    want = np.log(np.maximum(oracle_np.state_likelihoods(model, x[:2].astype(np.float64)), 1e-30))
    assert np.abs(rows[:2] - want).max() <= 4e-5
    engine.stream_open(200.0)
    try:
        l0 = engine.launch_count()
        got = np.stack([engine.gmm_logprobs(x[i:i + 1], precision=F32, tiny=1e-30)[0] for i in range(len(idx))])
        assert np.array_equal(got, rows)
        if K == 16:
            assert np.array_equal(engine.gmm_logprobs(x[:7], precision=F32, tiny=1e-30), blk)
        assert engine.launch_count() == l0
        import ctypes as C
        lib, h = engine._lib, engine._h
        for F in ((1, 4, 8, 16) if K == 16 else (1, 4, 8)):
            xb = np.ascontiguousarray(x[:F])
            ob = np.empty((F, S), dtype=np.float32)
            px, po = C.c_void_p(xb.ctypes.data), C.c_void_p(ob.ctypes.data)
            call = lambda: lib.akugpu_gmm_logprobs(h, px, 0, F, 0, C.c_double(1e-30), po)
            for _ in range(200):
                assert call() == 0
            t0 = time.perf_counter()
            n = 5000
            for _ in range(n):
                call()
            us = 1e6 * (time.perf_counter() - t0) / n
            # the zero-copy entry point: a view of the pinned rows
            view = engine.stream_logprobs(xb, tiny=1e-30)
            assert np.array_equal(view, ob)
            rows_p = C.POINTER(C.c_float)()
            vcall = lambda: lib.akugpu_stream_logprobs(h, px, F, C.c_double(1e-30), C.byref(rows_p))
            for _ in range(200):
                assert vcall() == 0
            t0 = time.perf_counter()
            for _ in range(n):
                vcall()
            us_view = 1e6 * (time.perf_counter() - t0) / n
            print("resident scorer %d x %d, F = %d: %.2f us per akugpu_stream_logprobs call (pinned rows returned, ctypes loop)" % (S, K, F, us_view))
            nat = engine.stream_latency(xb, n_calls=5000)
            print("resident scorer %d x %d, F = %d: timed inside the library: mean %.2f us, median %.2f, p99 %.2f, max %.1f"
                  % (S, K, F, nat["mean_us"], nat["median_us"], nat["p99_us"], nat["max_us"]))
            st = engine.stream_stats()
            print("resident scorer %d x %d, F = %d: %.2f us per akugpu_gmm_logprobs call (host buffers, ctypes loop); on the device: "
                  "%.2f us command -> A', %.2f -> this CTA's results stored, %.2f -> every CTA's, %.2f us command -> rows in host memory"
                  % (S, K, F, us, st["device_ns_features"] / 1e3, st["device_ns_stored"] / 1e3, st["device_ns_fenced"] / 1e3, st["device_ns_call"] / 1e3))
            assert us < 500.0
        assert engine.launch_count() == l0
    finally:
        engine.stream_close()


def test_stream_session_mirror(engine, ref_small):
    """The Python mirror of akugpu::StreamSession: a per-frame loop over an utterance, rows returned as views."""
    from aaltoasr_b200 import StreamSession
    g = ref_small
    load_model(engine, g["model"])
    x = g["feats"][:40].astype(np.float32)
    engine.stream_close()
    want = np.stack([engine.gmm_logprobs(x[f:f + 1], precision=F32, tiny=1e-30)[0] for f in range(len(x))])   # one launch per frame
    want6 = engine.gmm_logprobs(x[3:9], precision=F32, tiny=1e-30)
    l0 = engine.launch_count()
    with StreamSession(engine, tiny=1e-30, idle_ms=200.0) as ses:
        for f in range(len(x)):
            assert np.array_equal(ses.log_probs(x[f])[0], want[f])
        assert np.array_equal(ses.log_probs(x[3:9]), want6)
    assert engine.launch_count() - l0 == 1                       # the kernel itself
    assert not engine.stream_stats()["live"]


def test_resident_scorer_stress(engine, ref_small):
    """A few thousand messages in a seeded random order: 1 / 2-frame commands (packets), 3 ... 16-frame commands (relayed by
    CTA 0), pauses longer than the idle timer (the kernel ends and is started again), other entry points in between (the
    kernel is told to end).  Every answer carries the bits of the launch-per-call scorer."""
    g = ref_small
    load_model(engine, g["model"])
    rng = np.random.default_rng(20261017)
    x = g["feats"][:64].astype(np.float32)
    engine.stream_close()
    blocks = [(int(s), int(n)) for n in (1, 2, 3, 4, 7, 16) for s in rng.integers(0, 64 - n, size=6)]
    want = {b: engine.gmm_logprobs(x[b[0]:b[0] + b[1]], precision=F32, tiny=1e-30).copy() for b in blocks}
    s0 = engine.stream_stats()
    engine.stream_open(5.0)                                       # idle timer: 5 ms
    try:
        n_calls = 0
        for it in range(3000):
            b = blocks[int(rng.integers(0, len(blocks)))]
            r = rng.random()
            if r < 0.004:
                time.sleep(0.02)                                  # the kernel ends on its timer
            elif r < 0.008:
                engine.gmm_lna(g["feats"][:8], precision=F64, lnabytes=2)     # another entry point: quit message first
            if it & 1:
                got = engine.stream_logprobs(x[b[0]:b[0] + b[1]], tiny=1e-30)
            else:
                got = engine.gmm_logprobs(x[b[0]:b[0] + b[1]], precision=F32, tiny=1e-30)
            n_calls += 1
            assert np.array_equal(got, want[b]), (it, b)
        st = engine.stream_stats()
        assert st["calls"] - s0["calls"] == n_calls
        assert 2 <= st["launches"] - s0["launches"] <= 60          # restarted after every pause / foreign call, not more
    finally:
        engine.stream_close()


def test_one_resident_scorer_per_device(engine, ref_small):
    """A second context of the process cannot open a session on a device whose resident scorer another context holds."""
    from aaltoasr_b200 import AkuGpu
    g = ref_small
    load_model(engine, g["model"])
    x = g["feats"][:1].astype(np.float32)
    engine.stream_close()
    want = engine.gmm_logprobs(x, precision=F32, tiny=1e-30)
    other = AkuGpu(0)
    try:
        load_model(other, g["model"])
        engine.stream_open(100.0)
        with pytest.raises(Exception):
            other.stream_open(100.0)
        assert np.array_equal(other.gmm_logprobs(x, precision=F32, tiny=1e-30), want)      # served by one launch per call
        assert np.array_equal(engine.stream_logprobs(x, tiny=1e-30), want)
        engine.stream_close()
        other.stream_open(100.0)                                                            # free now
        assert np.array_equal(other.stream_logprobs(x, tiny=1e-30), want)
        other.stream_close()
    finally:
        engine.stream_close()
        other.close()
