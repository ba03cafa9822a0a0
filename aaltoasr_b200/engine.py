"""Thin object wrapper over the C ABI (include/akugpu.h).

Buffers may be numpy arrays (host) or torch tensors (CUDA or pinned host); only their
address crosses the boundary.  All arithmetic happens in libakugpu.so on the GPU.
"""
import ctypes as C

import numpy as np

from ._lib import AkuGpuError, load_library

F32 = 0   # AKUGPU_F32: throughput mode
F64 = 1   # AKUGPU_F64: parity mode (the reference's arithmetic in double)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class DevPtr(int):
    """A raw DEVICE address (e.g. a peer buffer mapped by shared_open, plus an offset) passed where a buffer is expected."""


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, DevPtr):
        return C.c_void_p(int(x))
    if _is_torch(x):
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return C.c_void_p(x.data_ptr())
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return C.c_void_p(x.ctypes.data)
    raise TypeError("expected numpy array or torch tensor, got %r" % type(x))


def _is_f64(x):
    if _is_torch(x):
        import torch
        if x.dtype == torch.float64:
            return 1
        if x.dtype == torch.float32:
            return 0
    else:
        if x.dtype == np.float64:
            return 1
        if x.dtype == np.float32:
            return 0
    raise TypeError("features must be float32 or float64")


class AkuGpu:
    """One context = one GPU = one host thread (not thread-safe, like the reference classes)."""

    def __init__(self, device=0):
        self._lib = load_library()
        self._h = self._lib.akugpu_create(int(device))
        if not self._h:
            raise AkuGpuError(-1, self._lib.akugpu_last_error(None).decode())
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.akugpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise AkuGpuError(rc, self._lib.akugpu_last_error(self._h).decode())

    # ---- context ----
    def set_stream(self, cuda_stream_ptr):
        self._ck(self._lib.akugpu_set_stream(self._h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def synchronize(self):
        self._ck(self._lib.akugpu_synchronize(self._h))

    def launch_count(self):
        return int(self._lib.akugpu_launch_count(self._h))

    def stage_times_reset(self, enable=True):
        self._ck(self._lib.akugpu_stage_times_reset(self._h, 1 if enable else 0))

    def stage_times(self):
        ms = (C.c_double * 3)()
        n = (C.c_int64 * 3)()
        self._ck(self._lib.akugpu_stage_times(self._h, ms, n))
        return {"frontend": (ms[0], n[0]), "gmm": (ms[1], n[1]), "lna": (ms[2], n[2])}

    def set_chunk_frames(self, frames):
        self._ck(self._lib.akugpu_set_chunk_frames(self._h, int(frames)))

    def set_scorer_variant(self, variant):
        self._ck(self._lib.akugpu_set_scorer_variant(self._h, int(variant)))

    def expanded_form_q(self):
        return float(self._lib.akugpu_model_expanded_form_q(self._h))

    def scorer_in_use(self):
        """0 double path, 1 FP32-pipe, 2 bf16x3 tensor-core, 3 fp16x2 tensor-core (resident A'), 4 fp16x2 (streaming A'),
        5 fp16x2 tensor-core + FP32-pipe for the ill-conditioned states."""
        return int(self._lib.akugpu_scorer_in_use(self._h))

    def set_streaming(self, enable=True):
        self._ck(self._lib.akugpu_set_streaming(self._h, 1 if enable else 0))

    def stream_open(self, idle_ms=100.0):
        """Starts the resident scorer (akugpu_stream_open): small host-buffer calls of gmm_score / gmm_logprobs become
        messages to a kernel that keeps the parameter image in shared memory; any other call ends it."""
        self._ck(self._lib.akugpu_stream_open(self._h, float(idle_ms)))

    def stream_logprobs(self, feats, tiny=1e-30):
        """akugpu_stream_logprobs: 1..16 rows of host float32 features -> a VIEW of the pinned result rows [F x S]
        (valid until the next call on this context)."""
        x = np.ascontiguousarray(feats, dtype=np.float32)
        rows = C.POINTER(C.c_float)()
        self._ck(self._lib.akugpu_stream_logprobs(self._h, C.c_void_p(x.ctypes.data), int(x.shape[0]), float(tiny), C.byref(rows)))
        return np.ctypeslib.as_array(rows, shape=(int(x.shape[0]), self.num_states))

    def stream_latency(self, feats, tiny=1e-30, n_calls=2000):
        """akugpu_stream_latency: microseconds per per-frame-loop call, timed inside the library."""
        x = np.ascontiguousarray(feats, dtype=np.float32)
        out = (C.c_double * 4)()
        self._ck(self._lib.akugpu_stream_latency(self._h, C.c_void_p(x.ctypes.data), int(x.shape[0]), float(tiny), int(n_calls), out))
        return {"mean_us": out[0], "median_us": out[1], "p99_us": out[2], "max_us": out[3]}

    def stream_close(self):
        self._ck(self._lib.akugpu_stream_close(self._h))

    def stream_stats(self):
        out = (C.c_int64 * 8)()
        self._ck(self._lib.akugpu_stream_stats(self._h, out))
        return {"want": bool(out[0]), "live": bool(out[1]), "launches": int(out[2]), "calls": int(out[3]),
                "device_ns_features": int(out[4]), "device_ns_call": int(out[5]),
                "device_ns_stored": int(out[6]), "device_ns_fenced": int(out[7])}

    def stream_probe(self):
        """Streaming-regime rates for the loaded model (akugpu_stream_probe)."""
        out = (C.c_double * 8)()
        self._ck(self._lib.akugpu_stream_probe(self._h, out))
        b = out[0]
        return {"image_bytes": b, "kernel_s_l2": out[1], "kernel_s_hbm": out[2], "probe_s_l2": out[3], "probe_s_hbm": out[4],
                "kernel_GBps_l2": b / out[1] / 1e9, "kernel_GBps_hbm": b / out[2] / 1e9,
                "probe_GBps_l2": b / out[3] / 1e9, "probe_GBps_hbm": b / out[4] / 1e9,
                "sm_mhz_isolated": out[5], "kernel_s_train": out[6], "kernel_GBps_train": b / out[6] / 1e9 if out[6] > 0 else None,
                "sm_mhz_train": out[7]}

    def pipe_rates(self):
        out = (C.c_double * 8)()
        self._ck(self._lib.akugpu_pipe_rates(self._h, out))
        return {"ffma": out[0], "ffma2": out[1], "dfma": out[2], "ex2": out[3], "ffma_3reg": out[4],
                "ffma2_3reg": out[5], "tile_ffma": out[6], "tile_ffma2": out[7]}

    # ---- front-end ----
    def frontend_load_config(self, path):
        self._ck(self._lib.akugpu_frontend_load_config(self._h, str(path).encode()))

    def frontend_load_config_text(self, text):
        self._ck(self._lib.akugpu_frontend_load_config_text(self._h, text.encode()))

    def frontend_set_parameters(self, module, text):
        self._ck(self._lib.akugpu_frontend_set_parameters(self._h, module.encode(), text.encode()))

    @property
    def feature_dim(self):
        d = self._lib.akugpu_frontend_dim(self._h)
        if d < 0:
            raise AkuGpuError(d, "no feature configuration loaded")
        return d

    @property
    def sample_rate(self):
        return self._lib.akugpu_frontend_sample_rate(self._h)

    @property
    def frame_rate(self):
        return float(self._lib.akugpu_frontend_frame_rate(self._h))

    def num_frames(self, n_samples):
        n = self._lib.akugpu_frontend_num_frames(self._h, int(n_samples))
        if n < 0:
            raise AkuGpuError(int(n), "no feature configuration loaded")
        return int(n)

    def frame_offsets(self, utt_offsets):
        uo = np.ascontiguousarray(utt_offsets, dtype=np.int64)
        fo = np.zeros(len(uo), dtype=np.int64)
        self._ck(self._lib.akugpu_features(self._h, None, _ptr(uo), len(uo) - 1, None, 0, _ptr(fo)))
        return fo

    def features(self, pcm, utt_offsets=None, dtype=np.float32, out=None):
        """pcm: int16 array/tensor (all utterances concatenated); returns ([F x dim], frame_offsets)."""
        n = int(pcm.numel() if _is_torch(pcm) else pcm.size)
        uo = np.ascontiguousarray(utt_offsets if utt_offsets is not None else [0, n], dtype=np.int64)
        fo = self.frame_offsets(uo)
        if out is None:
            out = np.empty((int(fo[-1]), self.feature_dim), dtype=dtype)
        self._ck(self._lib.akugpu_features(self._h, _ptr(pcm), _ptr(uo), len(uo) - 1, _ptr(out), _is_f64(out), _ptr(fo)))
        return out, fo

    def features_range(self, pcm, start, end, module=None, dtype=np.float64):
        """Frames [start,end) of one utterance in the reference's frame numbering (may leave the file)."""
        n = int(pcm.numel() if _is_torch(pcm) else pcm.size)
        dim = C.c_int(0)
        self._ck(self._lib.akugpu_features_range(self._h, None, n, 0, 0, module.encode() if module else None, None, 0,
                                                 C.byref(dim)))
        out = np.empty((max(0, end - start), dim.value), dtype=dtype)
        self._ck(self._lib.akugpu_features_range(self._h, _ptr(pcm), n, int(start), int(end),
                                                 module.encode() if module else None, _ptr(out), _is_f64(out),
                                                 C.byref(dim)))
        return out

    def features_pre(self, rows, row_offsets=None, dtype=np.float64, out=None):
        """A configuration with a `pre` base module: rows = stored float32 features [n x dim] of all utterances."""
        rows = np.ascontiguousarray(rows, dtype=np.float32) if not _is_torch(rows) else rows
        n = int(rows.shape[0])
        ro = np.ascontiguousarray(row_offsets if row_offsets is not None else [0, n], dtype=np.int64)
        fo = np.zeros(len(ro), dtype=np.int64)
        self._ck(self._lib.akugpu_features_pre(self._h, None, _ptr(ro), len(ro) - 1, None, 0, _ptr(fo)))
        if out is None:
            out = np.empty((int(fo[-1]), self.feature_dim), dtype=dtype)
        self._ck(self._lib.akugpu_features_pre(self._h, _ptr(rows), _ptr(ro), len(ro) - 1, _ptr(out), _is_f64(out), _ptr(fo)))
        return out, fo

    def features_pre_range(self, rows, start, end, module=None, dtype=np.float64):
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        dim = C.c_int(0)
        mod = module.encode() if module else None
        self._ck(self._lib.akugpu_features_pre_range(self._h, None, rows.shape[0], 0, 0, mod, None, 0, C.byref(dim)))
        out = np.empty((max(0, end - start), dim.value), dtype=dtype)
        self._ck(self._lib.akugpu_features_pre_range(self._h, _ptr(rows), rows.shape[0], int(start), int(end), mod, _ptr(out),
                                                     _is_f64(out), C.byref(dim)))
        return out

    # ---- model ----
    def model_read(self, base):
        self._ck(self._lib.akugpu_model_read(self._h, str(base).encode()))

    def model_load_diag(self, mix_offsets, mix_gauss, mix_weight, means, covs):
        mo = np.ascontiguousarray(mix_offsets, dtype=np.int32)
        mg = np.ascontiguousarray(mix_gauss, dtype=np.int32)
        mw = np.ascontiguousarray(mix_weight, dtype=np.float64)
        mu = np.ascontiguousarray(means, dtype=np.float64)
        cv = np.ascontiguousarray(covs, dtype=np.float64)
        if mu.ndim != 2 or mu.shape != cv.shape:
            raise ValueError("means/covs must be [G x D]")
        self._ck(self._lib.akugpu_model_load_diag(self._h, len(mo) - 1, mu.shape[0], mu.shape[1], _ptr(mo), _ptr(mg),
                                                  _ptr(mw), _ptr(mu), _ptr(cv)))

    def model_load_full(self, mix_offsets, mix_gauss, mix_weight, means, full_covs):
        mo = np.ascontiguousarray(mix_offsets, dtype=np.int32)
        mg = np.ascontiguousarray(mix_gauss, dtype=np.int32)
        mw = np.ascontiguousarray(mix_weight, dtype=np.float64)
        mu = np.ascontiguousarray(means, dtype=np.float64)
        cv = np.ascontiguousarray(full_covs, dtype=np.float64)
        if mu.ndim != 2 or cv.shape != (mu.shape[0], mu.shape[1], mu.shape[1]):
            raise ValueError("means must be [G x D], full_covs [G x D x D]")
        self._ck(self._lib.akugpu_model_load_full(self._h, len(mo) - 1, mu.shape[0], mu.shape[1], _ptr(mo), _ptr(mg),
                                                  _ptr(mw), _ptr(mu), _ptr(cv)))

    # ---- Gaussian clustering (phone_probs -C / --eval-minc / --eval-ming) ----
    def read_clustering(self, path):
        """HmmSet::read_clustering (aku/HmmSet.cc:1354)."""
        self._ck(self._lib.akugpu_model_read_clustering(self._h, str(path).encode()))

    def set_clustering(self, n_clusters, gauss_index, cluster_index):
        gi = np.ascontiguousarray(gauss_index, dtype=np.int32)
        ci = np.ascontiguousarray(cluster_index, dtype=np.int32)
        if gi.shape != ci.shape:
            raise ValueError("gauss_index / cluster_index must have the same length")
        self._ck(self._lib.akugpu_model_set_clustering(self._h, int(n_clusters), _ptr(gi), _ptr(ci), gi.size))

    def set_clustering_min_evals(self, min_clusters=1.0, min_gaussians=1.0):
        """HmmSet::set_clustering_min_evals (aku/HmmSet.cc:1360): switches the approximation on."""
        self._ck(self._lib.akugpu_model_set_clustering_min_evals(self._h, float(min_clusters), float(min_gaussians)))

    def use_clustering(self, on=True):
        self._ck(self._lib.akugpu_model_use_clustering(self._h, 1 if on else 0))

    def model_set_cmllr(self, W):
        """Global model-level CMLLR: W = [dim x (dim+1)] (column 0 = bias, the rest = A), None removes it
        (ConstrainedMllr with unitmode UNIT_NO, aku/ModelModules.cc:172-236)."""
        if W is None:
            self._ck(self._lib.akugpu_model_set_cmllr(self._h, None))
            return
        W = np.ascontiguousarray(W, dtype=np.float64)
        D = self.model_dim
        if W.shape != (D, D + 1):
            raise ValueError("W must be [%d x %d]" % (D, D + 1))
        self._ck(self._lib.akugpu_model_set_cmllr(self._h, _ptr(W)))

    def model_set_cmllr_units(self, unitmode, transforms):
        """Regression-class model-level CMLLR: transforms = [(units: list of str, W [dim x (dim+1)]), ...]
        (unitmode UNIT_PHONE / UNIT_MIX / UNIT_GAUSSIAN of aku/ModelModules.cc; [] removes them)."""
        n = len(transforms)
        D = self.model_dim
        arr = (C.c_char_p * max(1, n))(*[(" ".join(u)).encode() for u, _ in transforms])
        W = np.ascontiguousarray(np.stack([np.asarray(w, dtype=np.float64) for _, w in transforms]) if n else np.zeros((0, D, D + 1)))
        if n and W.shape != (n, D, D + 1):
            raise ValueError("every W must be [%d x %d]" % (D, D + 1))
        self._ck(self._lib.akugpu_model_set_cmllr_units(self._h, unitmode.encode(), n, arr, _ptr(W) if n else None))

    @property
    def num_states(self):
        return self._lib.akugpu_model_num_states(self._h)

    @property
    def model_dim(self):
        return self._lib.akugpu_model_dim(self._h)

    @property
    def num_gaussians(self):
        return self._lib.akugpu_model_num_gaussians(self._h)

    # ---- scoring ----
    def gmm_score(self, feats, precision=F32, out=None):
        F = int(feats.shape[0])
        if out is None:
            out = np.empty((F, self.num_states), dtype=np.float64 if precision == F64 else np.float32)
        self._ck(self._lib.akugpu_gmm_score(self._h, _ptr(feats), _is_f64(feats), F, precision, _ptr(out)))
        return out

    def gmm_logprobs(self, feats, precision=F32, tiny=1e-30, out=None):
        """The in-process decoder feed of decoder/decode-stream.cc: (float) log(max(likelihood, tiny)), un-normalised."""
        F = int(feats.shape[0])
        if out is None:
            out = np.empty((F, self.num_states), dtype=np.float32)
        self._ck(self._lib.akugpu_gmm_logprobs(self._h, _ptr(feats), _is_f64(feats), F, precision, float(tiny), _ptr(out)))
        return out

    def gmm_lna(self, feats, precision=F32, lnabytes=2, normalize=True, out=None):
        F = int(feats.shape[0])
        if out is None:
            out = np.empty((F, self.num_states * lnabytes), dtype=np.uint8)
        self._ck(self._lib.akugpu_gmm_lna(self._h, _ptr(feats), _is_f64(feats), F, precision, lnabytes,
                                          1 if normalize else 0, _ptr(out)))
        return out

    def phone_probs(self, pcm, utt_offsets=None, precision=F32, lnabytes=2, normalize=True, out=None, discard=False,
                    checksum=False, utt_checksums=False):
        """PCM -> LNA records for a batch of utterances.  Returns (records [F x S*lnabytes] or None,
        frame_offsets, checksum or None); with utt_checksums=True the third item is the array of per-utterance
        order-sensitive checksums (akugpu_phone_probs_ex) instead.  `out` may be a DevPtr (raw device address)."""
        n = int(pcm.numel() if _is_torch(pcm) else pcm.size)
        uo = np.ascontiguousarray(utt_offsets if utt_offsets is not None else [0, n], dtype=np.int64)
        fo = np.zeros(len(uo), dtype=np.int64)
        if out is None and not discard:
            fo = self.frame_offsets(uo)
            out = np.empty((int(fo[-1]), self.num_states * lnabytes), dtype=np.uint8)
        chk = C.c_uint64(0)
        if utt_checksums:
            uc = np.zeros(len(uo) - 1, dtype=np.uint64)
            self._ck(self._lib.akugpu_phone_probs_ex(self._h, _ptr(pcm), _ptr(uo), len(uo) - 1, precision, lnabytes,
                                                     1 if normalize else 0, _ptr(out), _ptr(fo), None, _ptr(uc)))
            return out, fo, uc
        self._ck(self._lib.akugpu_phone_probs(self._h, _ptr(pcm), _ptr(uo), len(uo) - 1, precision, lnabytes,
                                              1 if normalize else 0, _ptr(out), _ptr(fo),
                                              C.byref(chk) if checksum else None))
        return out, fo, (int(chk.value) if checksum else None)

    # ---- utterance-sharded runs over several GPUs: checksum sink, buffers shared between the processes of a node ----
    def checksum_begin(self, frame_offsets, rec_bytes):
        fo = np.ascontiguousarray(frame_offsets, dtype=np.int64)
        self._chk_n = len(fo) - 1
        self._ck(self._lib.akugpu_checksum_begin(self._h, _ptr(fo), self._chk_n, int(rec_bytes)))

    def checksum_update(self, records, first_frame, n_frames):
        """records: DEVICE buffer (torch CUDA tensor or DevPtr) holding frames [first_frame, first_frame + n_frames)."""
        self._ck(self._lib.akugpu_checksum_update(self._h, _ptr(records), int(first_frame), int(n_frames)))

    def checksum_end(self):
        out = np.zeros(self._chk_n, dtype=np.uint64)
        self._ck(self._lib.akugpu_checksum_end(self._h, _ptr(out)))
        return out

    def shared_alloc(self, nbytes):
        """(DevPtr, 64-byte handle) of a device buffer other processes of this node can map (CUDA IPC)."""
        p = C.c_void_p(0)
        h = np.zeros(64, dtype=np.uint8)
        self._ck(self._lib.akugpu_shared_alloc(self._h, int(nbytes), C.byref(p), _ptr(h)))
        return DevPtr(p.value), h.tobytes()

    def shared_open(self, handle):
        p = C.c_void_p(0)
        h = np.frombuffer(bytes(handle), dtype=np.uint8).copy()
        self._ck(self._lib.akugpu_shared_open(self._h, _ptr(h), C.byref(p)))
        return DevPtr(p.value)

    def copy_async(self, dst, src, nbytes):
        """cudaMemcpyAsync between device buffers (tensors or DevPtr, local or mapped peer memory) on this context's stream."""
        self._ck(self._lib.akugpu_copy_async(self._h, _ptr(dst), _ptr(src), int(nbytes)))

    def shared_release(self, ptr):
        self._ck(self._lib.akugpu_shared_release(self._h, C.c_void_p(int(ptr))))

    def lna_header(self, lnabytes):
        buf = np.zeros(5, dtype=np.uint8)
        self._lib.akugpu_lna_header(self.num_states, lnabytes, _ptr(buf))
        return buf.tobytes()
