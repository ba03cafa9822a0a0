"""Utterance partitioning across GPUs.

The path shards by utterance exactly like the reference's process-level batching
(`phone_probs -B N -I i`, aku/phone_probs.cc:78-79,135-139; contiguous split in aku/Recipe.cc:63-115).
Every rank computes the same partition from the same frame counts, so no communication is needed
for the split itself; the only exchange is a control-plane all-gather of per-utterance results
(frame counts, checksums) for whoever assembles the output table.
"""
import numpy as np


def lpt_partition(n_frames, world_size):
    """Longest-processing-time-first: utterances sorted by frame count (ties by index), each to the
    least-loaded rank (ties to the lowest rank).  Returns a list of index arrays, one per rank,
    each in ascending utterance order."""
    n_frames = np.asarray(n_frames, dtype=np.int64)
    order = sorted(range(len(n_frames)), key=lambda i: (-int(n_frames[i]), i))
    load = [0] * world_size
    parts = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        parts[r].append(i)
        load[r] += int(n_frames[i])
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def reference_partition(n_utts, world_size):
    """The reference's contiguous split: what `-B world_size -I rank+1` would give each process."""
    per, rem = divmod(n_utts, world_size)
    out, start = [], 0
    for r in range(world_size):
        cnt = per + (1 if r < rem else 0)
        out.append(np.arange(start, start + cnt, dtype=np.int64))
        start += cnt
    return out


def gather_utterance_table(local_ids, local_frames, local_checksums, n_utts, group=None):
    """All-gather of (utterance id, frame count, LNA checksum) so that every rank holds the global
    table.  Works with any torch.distributed backend (NCCL on GPUs, gloo on CPU); tensors live on
    the current CUDA device when the backend is NCCL."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    m = max(1, n_utts)                    # fixed-size rows so all_gather needs no size exchange
    buf = torch.full((m, 3), -1, dtype=torch.int64, device=dev)
    k = len(local_ids)
    if k:
        buf[:k, 0] = torch.as_tensor(np.asarray(local_ids, dtype=np.int64), device=dev)
        buf[:k, 1] = torch.as_tensor(np.asarray(local_frames, dtype=np.int64), device=dev)
        buf[:k, 2] = torch.as_tensor(np.asarray(local_checksums, dtype=np.uint64).astype(np.int64), device=dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    frames = np.zeros(n_utts, dtype=np.int64)
    chks = np.zeros(n_utts, dtype=np.uint64)
    owner = np.full(n_utts, -1, dtype=np.int64)
    for r, t in enumerate(out):
        a = t.cpu().numpy()
        a = a[a[:, 0] >= 0]
        frames[a[:, 0]] = a[:, 1]
        chks[a[:, 0]] = a[:, 2].astype(np.uint64)
        owner[a[:, 0]] = r
    return frames, chks, owner
