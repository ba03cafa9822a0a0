"""GPU tests of the device-side pieces of the utterance-sharded run (include/akugpu.h: akugpu_phone_probs_ex,
akugpu_checksum_*, akugpu_shared_*) and of the p2p LNA gather itself: two PROCESSES on one GPU (CUDA IPC works between
processes on the same device; the control plane is gloo here, NCCL refuses two ranks per device), the sender's LNA kernel
storing straight into the writer's buffer, the writer's checksum sink reading it."""
import os
import socket

import numpy as np
import pytest

from aaltoasr_b200 import F32, F64, multigpu as mg

pytestmark = pytest.mark.gpu


def load_model(engine, m):
    engine.model_load_diag(m["mix_offsets"], m["mix_gauss"], m["mix_weight"], m["means"], m["covs"])


def cuts_of(pcm):
    """Eight utterances of unequal length (6 - 40 frames) cut from the fixture's 1.5 s of audio."""
    frac = np.cumsum([0, 3, 9, 2, 5, 1.5, 7, 4, 6.5])
    edges = (frac / frac[-1] * len(pcm)).astype(np.int64)
    return [pcm[a:b] for a, b in zip(edges[:-1], edges[1:])]


def test_per_utterance_checksums(engine, ref_small):
    g = ref_small
    engine.frontend_load_config_text(g["cfg"])
    load_model(engine, g["model"])
    cuts = cuts_of(g["pcm"])
    uo = np.concatenate([[0], np.cumsum([c.size for c in cuts])]).astype(np.int64)
    for nb, prec in ((2, F32), (4, F32), (2, F64)):
        rec, fo, chk = engine.phone_probs(np.concatenate(cuts), uo, precision=prec, lnabytes=nb, utt_checksums=True)
        assert chk.dtype == np.uint64 and len(chk) == len(cuts)
        assert np.array_equal(chk, mg.utt_checksums_host(rec, fo))                  # the definition, on the host bytes
        # an utterance's checksum does not depend on what else was in the call, nor on the chunking
        one = [engine.phone_probs(c, precision=prec, lnabytes=nb, utt_checksums=True)[2][0] for c in cuts]
        assert np.array_equal(chk, np.array(one, dtype=np.uint64))
        try:
            engine.set_chunk_frames(128)
            _, _, small = engine.phone_probs(np.concatenate(cuts), uo, precision=prec, lnabytes=nb, discard=True, utt_checksums=True)
        finally:
            engine.set_chunk_frames(0)
        assert np.array_equal(chk, small)
    # the plain entry point still returns the byte-sum of the whole call
    rec, fo, tot = engine.phone_probs(np.concatenate(cuts), uo, lnabytes=2, checksum=True)
    assert tot == int(rec.astype(np.uint64).sum())


def test_checksum_sink_on_device_records(engine, ref_small):
    import torch
    g = ref_small
    engine.frontend_load_config_text(g["cfg"])
    load_model(engine, g["model"])
    cuts = cuts_of(g["pcm"])
    uo = np.concatenate([[0], np.cumsum([c.size for c in cuts])]).astype(np.int64)
    rec, fo, chk = engine.phone_probs(np.concatenate(cuts), uo, lnabytes=2, utt_checksums=True)
    dev = torch.from_numpy(rec).cuda()
    R = rec.shape[1]
    engine.checksum_begin(fo, R)
    F = rec.shape[0]
    pieces = [(0, 17), (100, F - 100), (17, 83)]                                   # any order, any split
    for a, n in pieces:
        engine.checksum_update(dev[a:a + n], a, n)
    assert np.array_equal(engine.checksum_end(), chk)
    # a stream numbered from an offset (the writer's concatenation of all senders' streams)
    engine.checksum_begin(fo + 1000, R)
    engine.checksum_update(dev, 1000, F)
    assert np.array_equal(engine.checksum_end(), chk)
    from aaltoasr_b200 import AkuGpuError
    with pytest.raises(AkuGpuError, match="without akugpu_checksum_begin"):
        engine.checksum_update(dev, 0, 1)
    engine.checksum_begin(fo, R)
    with pytest.raises(AkuGpuError, match="outside the frame-offset table"):
        engine.checksum_update(dev, F - 1, 2)
    with pytest.raises(AkuGpuError, match="device buffer"):
        engine.checksum_update(rec, 0, 1)
    engine.checksum_end()


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from aaltoasr_b200 import AkuGpu
    from aaltoasr_b200.engine import DevPtr
    from conftest import load_golden
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g_tok, g_free = dist.new_group(), dist.new_group()
    eng = sink_eng = None
    try:
        g = load_golden("ref_small")
        eng = AkuGpu(0)
        eng.frontend_load_config_text(g["cfg"])
        load_model(eng, g["model"])
        cuts = cuts_of(g["pcm"])
        n_frames = np.array([eng.num_frames(c.size) for c in cuts], dtype=np.int64)
        frames = mg.gather_frame_counts(np.arange(rank, len(cuts), world), n_frames[rank::world], len(cuts))
        assert np.array_equal(frames, n_frames)
        parts = mg.partition(frames, world, "lpt")
        R = eng.num_states * 2
        plan = mg.GatherPlan(frames, parts, max_frames=30, rec_bytes=R, writer=0)
        mine = plan.parts[rank]
        pcm = np.concatenate([cuts[u] for u in mine])
        uo = np.concatenate([[0], np.cumsum([cuts[u].size for u in mine])]).astype(np.int64)
        src = np.zeros(len(mine), dtype=np.uint64)

        def produce(u0, u1, out):
            a, b = int(uo[u0]), int(uo[u1])
            _, _, uc = eng.phone_probs(pcm[a:b], uo[u0:u1 + 1] - uo[u0], lnabytes=2, out=out, utt_checksums=True)
            src[u0:u1] = uc

        nslots = 2
        handle = torch.zeros(64, dtype=torch.uint8)
        shared = None
        if rank == 0:
            shared, h = eng.shared_alloc(world * nslots * plan.slot_bytes)
            handle.copy_(torch.frombuffer(bytearray(h), dtype=torch.uint8))
        dist.broadcast(handle, 0)
        root = shared if rank == 0 else eng.shared_open(handle.numpy().tobytes())
        sunk = None
        if rank == 0:
            sink_eng = AkuGpu(0)
            sink_eng.checksum_begin(plan.fo[1], R)

        def sink(r, slot, f0, n):
            sink_eng.checksum_update(slot, f0, n)

        own = [torch.empty(plan.slot_bytes, dtype=torch.uint8, device="cuda") for _ in range(2)]
        mg.gather_p2p(plan, rank, produce, sink, own, lambda r, j: DevPtr(int(root) + (r * nslots + j) * plan.slot_bytes), nslots,
                      g_tok, g_free, torch.zeros(1, dtype=torch.int64))
        if rank == 0:
            sunk = sink_eng.checksum_end()
        dist.barrier()
        if rank != 0:
            eng.shared_release(root)
        dist.barrier()
        if rank == 0:
            eng.shared_release(shared)
        from aaltoasr_b200.partition import gather_utterance_table
        tf, tc, owner = gather_utterance_table(mine, frames[mine], src, len(cuts))
        q.put((rank, [int(x) for x in tc], None if sunk is None else [int(x) for x in sunk], [int(u) for u in plan.parts[1]],
               [len(s) for s in plan.sched]))
    finally:
        if sink_eng:
            sink_eng.close()
        if eng:
            eng.close()
        dist.destroy_process_group()


def test_p2p_gather_between_two_processes_on_one_gpu(engine, ref_small):
    import torch.multiprocessing as mp
    g = ref_small
    engine.frontend_load_config_text(g["cfg"])
    load_model(engine, g["model"])
    cuts = cuts_of(g["pcm"])
    uo = np.concatenate([[0], np.cumsum([c.size for c in cuts])]).astype(np.int64)
    _, _, single = engine.phone_probs(np.concatenate(cuts), uo, lnabytes=2, discard=True, utt_checksums=True)   # the 1-process run
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, table, sunk, sender_ids, nsub in res:
        assert table == [int(x) for x in single]                       # 2-process table == 1-process checksums, per utterance
        if rank == 0:
            assert sunk == [int(single[u]) for u in sender_ids]       # what the sink read in the writer's memory == what was sent
            assert nsub[1] >= 3                                        # the two slots were reused
