"""CPU tests (gloo, world sizes 2 and 3, 127.0.0.1) of the utterance-sharded run's host logic -- aaltoasr_b200/multigpu.py:
model broadcast, frame-count all-gather, partition + sub-batch schedule, and BOTH LNA gather protocols run with a fake
producer (records = a closed-form function of (utterance, frame, byte)); the "peer memory" of the p2p protocol is a
/dev/shm file mapped by every process.  The per-utterance checksums the writer's sink computes must equal the ones
every rank computed at the source and the ones of a single-process run."""
import os
import socket
import tempfile

import numpy as np
import pytest

from aaltoasr_b200 import multigpu as mg

REC = 12                 # bytes per frame record in the fake (S = 6 states x 2 bytes)


def fake_records(utt_id, n_frames):
    f = np.arange(n_frames, dtype=np.int64)[:, None]
    b = np.arange(REC, dtype=np.int64)[None, :]
    return ((utt_id * 131 + f * 17 + b * 7 + (f * b) % 5) & 255).astype(np.uint8)


def test_sub_batches_and_checksum_definition():
    n = [5, 3, 9, 1, 1, 1, 12, 2]
    sched = mg.sub_batches(n, 10)
    assert sched == [(0, 2, 0, 8), (2, 4, 8, 10), (4, 6, 18, 2), (6, 7, 20, 12), (7, 8, 32, 2)]     # a long utterance gets its own
    assert mg.sub_batches([], 10) == [] and mg.sub_batches([4], 10) == [(0, 1, 0, 4)]
    # the checksum: order-sensitive in frames and in bytes, additive over any split of the frames into updates
    rec = fake_records(3, 7)
    fo = [0, 3, 7]
    c = mg.utt_checksums_host(rec, fo)
    w = rec.view("<u4").astype(object)
    want0 = sum((2 * i + 1) * sum((2 * j + 1) * int(w[i, j]) for j in range(3)) for i in range(3)) % (1 << 64)
    assert int(c[0]) == want0
    swapped = rec.copy(); swapped[[0, 1]] = swapped[[1, 0]]
    assert mg.utt_checksums_host(swapped, fo)[0] != c[0] and mg.utt_checksums_host(swapped, fo)[1] == c[1]
    odd = fake_records(1, 4)[:, :10]                      # record size not a multiple of 4: zero padded words
    assert mg.utt_checksums_host(odd, [0, 4]).shape == (1,)


def _worker(rank, world, port, shm_path, split, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g_tok, g_free = dist.new_group(), dist.new_group()
    try:
        # 1. model broadcast
        rng = np.random.default_rng(7)
        model = None
        if rank == 0:
            model = dict(mix_offsets=np.array([0, 2, 5], np.int32), mix_gauss=np.array([0, 1, 2, 3, 4], np.int32),
                         mix_weight=rng.random(5), means=rng.standard_normal((5, 3)), covs=rng.random((5, 3)) + 0.5)
        got, nbytes = mg.broadcast_model(model, 0)
        want_rng = np.random.default_rng(7)
        w = want_rng.random(5); m = want_rng.standard_normal((5, 3)); c = want_rng.random((5, 3)) + 0.5
        assert np.array_equal(got["mix_weight"], w) and np.array_equal(got["means"], m) and np.array_equal(got["covs"], c)
        assert got["mix_offsets"].dtype == np.int32 and list(got["mix_offsets"]) == [0, 2, 5] and nbytes == (3 + 5 + 5 + 15 + 15) * 8
        # 2. frame counts: every rank looks at ids rank, rank + world, ...
        n_utts = 23
        truth = np.random.default_rng(5).integers(3, 40, n_utts)
        ids = np.arange(rank, n_utts, world)
        frames = mg.gather_frame_counts(ids, truth[ids], n_utts)
        assert np.array_equal(frames, truth)
        # 3. partition + schedule
        parts = mg.partition(frames, world, split)
        plan = mg.GatherPlan(frames, parts, max_frames=50, rec_bytes=REC, writer=0)
        mine = plan.parts[rank]
        src_chk = {}

        def produce(u0, u1, out):
            a = 0
            for u in mine[u0:u1]:
                r = fake_records(int(u), int(frames[u]))
                out[a:a + r.size] = torch.from_numpy(r.reshape(-1)) if isinstance(out, torch.Tensor) else r.reshape(-1)
                a += r.size
                src_chk[int(u)] = int(mg.utt_checksums_host(r, [0, r.shape[0]])[0])

        sunk = {r: [] for r in plan.senders}

        def sink(r, slot, f0, n):
            buf = slot.numpy() if isinstance(slot, torch.Tensor) else slot
            sunk[r].append((f0, np.array(buf[:n * REC]).reshape(n, REC).copy()))

        def sink_result():
            out = {}
            for r in plan.senders:
                pieces = sorted(sunk[r], key=lambda t: t[0])
                assert [p[0] for p in pieces] == [s[2] for s in plan.sched[r]]
                allrec = np.concatenate([p[1] for p in pieces]) if pieces else np.zeros((0, REC), np.uint8)
                chk = mg.utt_checksums_host(allrec, plan.fo[r])
                for u, cv in zip(plan.parts[r], chk):
                    out[int(u)] = int(cv)
                sunk[r] = []
            return out

        # 4a. gather over send / recv
        send_slots = [torch.zeros(plan.slot_bytes, dtype=torch.uint8) for _ in range(2)]
        recv_slots = {r: [torch.zeros(plan.slot_bytes, dtype=torch.uint8) for _ in range(2)] for r in plan.senders} if rank == 0 else None
        mg.gather_nccl(plan, rank, produce, sink, send_slots, recv_slots)
        res_nccl = sink_result() if rank == 0 else None
        # 4b. gather by direct stores into the writer's memory (here: a shared file mapping), tokens over two groups
        nslots = 2
        shm = np.memmap(shm_path, dtype=np.uint8, mode="r+", shape=(world, nslots, plan.slot_bytes))
        mg.gather_p2p(plan, rank, produce, sink, [np.zeros(plan.slot_bytes, np.uint8)] * 2, lambda r, j: shm[r, j], nslots,
                      g_tok, g_free, torch.zeros(1, dtype=torch.int64))
        res_p2p = sink_result() if rank == 0 else None
        # 5. the table
        from aaltoasr_b200 import partition as pt
        tf, tc, owner = pt.gather_utterance_table(mine, frames[mine], [src_chk[int(u)] for u in mine], n_utts)
        q.put((rank, tf.tolist(), [int(x) for x in tc], owner.tolist(), res_nccl, res_p2p, [len(s) for s in plan.sched]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,split", [(2, "lpt"), (3, "reference")])
def test_gather_protocols_over_gloo(world, split):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    n_utts = 23
    truth = np.random.default_rng(5).integers(3, 40, n_utts)
    shm_dir = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    fd, shm_path = tempfile.mkstemp(dir=shm_dir, prefix="akugpu_test_")
    try:
        os.ftruncate(fd, world * 2 * 64 * REC * 4)
        os.close(fd)
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, world, port, shm_path, split, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = [q.get(timeout=180) for _ in procs]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        os.unlink(shm_path)
    single = {u: int(mg.utt_checksums_host(fake_records(u, int(truth[u])), [0, int(truth[u])])[0]) for u in range(n_utts)}
    parts = mg.partition(truth, world, split)
    for rank, tf, tc, owner, res_nccl, res_p2p, nsub in res:
        assert tf == truth.tolist()
        assert tc == [single[u] for u in range(n_utts)]            # N-process table == single-process checksums
        for r in range(world):
            assert all(owner[i] == r for i in parts[r])
        if rank == 0:
            want = {u: single[u] for r in range(1, world) for u in (int(x) for x in parts[r])}
            assert res_nccl == want and res_p2p == want                # what arrived at the writer == what was sent
            assert max(nsub) >= 3                                      # the slots were actually reused
