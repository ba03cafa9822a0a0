"""CPU tests of the multi-GPU host logic: partitioning and the world-size-2 control-plane exchange
(gloo backend, 127.0.0.1 rendezvous)."""
import os
import socket

import numpy as np
import pytest

from aaltoasr_b200 import partition


def test_lpt_partition_is_balanced_and_complete():
    rng = np.random.default_rng(0)
    frames = rng.integers(600, 1900, size=1000)
    parts = partition.lpt_partition(frames, 8)
    allidx = np.concatenate(parts)
    assert sorted(allidx) == list(range(1000))
    loads = np.array([frames[p].sum() for p in parts])
    assert loads.max() - loads.min() <= frames.max()
    again = partition.lpt_partition(frames, 8)                      # deterministic
    assert all(np.array_equal(a, b) for a, b in zip(parts, again))
    assert all(len(p) == 0 for p in partition.lpt_partition([], 4))            # empty input


def test_reference_partition_matches_recipe_split():
    parts = partition.reference_partition(10, 4)       # aku/Recipe.cc contiguous batches
    assert [list(p) for p in parts] == [[0, 1, 2], [3, 4, 5], [6, 7], [8, 9]]


def _worker(rank, world, port, frames, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parts = partition.lpt_partition(frames, world)
    mine = parts[rank]
    chk = [int(frames[i]) * 7 + i for i in mine]                    # stand-in for the LNA checksum of each utterance
    got_frames, got_chk, owner = partition.gather_utterance_table(mine, frames[mine], chk, len(frames))
    q.put((rank, got_frames.tolist(), got_chk.tolist(), owner.tolist()))
    dist.destroy_process_group()


def test_world_size_2_gather_gloo():
    import torch.multiprocessing as mp
    frames = np.array([1248, 600, 1900, 73, 1248, 999, 1500], dtype=np.int64)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    parts = partition.lpt_partition(frames, 2)
    for rank, got_frames, got_chk, owner in res:
        assert got_frames == frames.tolist()
        assert got_chk == [int(frames[i]) * 7 + i for i in range(len(frames))]
        for r in range(2):
            assert all(owner[i] == r for i in parts[r])
