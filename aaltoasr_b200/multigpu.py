"""Utterance-sharded phone_probs over the GPUs of one node (BASELINE.json configs[3]).

The path shards by utterance exactly like the reference's process-level batching (`phone_probs -B N -I i`,
aku/phone_probs.cc:78-79,135-139 over aku/Recipe.cc:63-115): one process per GPU, parameters replicated, every rank
scores its own utterances.  What the ranks exchange (torch.distributed: NCCL on GPUs, gloo in the CPU tests):

  1. broadcast_model          the acoustic model arrays, packed into one buffer, from the rank that read the files
  2. gather_frame_counts      all-gather of n_frames[utt] (each rank reads the headers of every world-th file)
  3. partition                redundantly on every rank from that table: LPT by frame count, or the reference's
                              contiguous split (what `-B world -I rank+1` gives)
  4. the LNA payload          "writers": every rank drains its own records (the reference's model, no collective);
                              "gather":  all records to ONE writer rank's rotating device buffer + checksum sink, either
                                 p2p   the LNA kernel of every rank stores straight into the writer's buffer over NVLink
                                       (CUDA-IPC mapped peer memory: epilogue + gather in one kernel; NCCL carries only
                                       two 8-byte tokens per sub-batch: "slot filled", "slot free"), or
                                 nccl  records into a local send slot, ncclSend -> ncclRecv into the writer's slot
  5. gather_utterance_table   all-gather of (utterance, n_frames, checksum): the writer's offset table and the
                              1-GPU-vs-N-GPU per-utterance checksum check

The protocol code takes a *producer* (scores a sub-batch into a buffer) and a *sink* (consumes a received slot) so that
the CPU tests run the very same loops over gloo with a fake producer.
"""
import numpy as np

from .partition import lpt_partition, reference_partition


# ---------------------------------------------------------------------------------------------- host-side helpers
def utt_checksums_host(records, frame_offsets):
    """The library's per-utterance checksum (include/akugpu.h, akugpu_checksum_begin) on host bytes -- for tests and for
    verifying buffers that landed in host memory.  records: [F x rec_bytes] uint8."""
    rec = np.ascontiguousarray(records, dtype=np.uint8)
    F, R = rec.shape
    pad = (-R) % 4
    if pad:
        rec = np.concatenate([rec, np.zeros((F, pad), np.uint8)], axis=1)
    w = rec.view("<u4").astype(np.uint64)
    with np.errstate(over="ignore"):
        row = (w * (2 * np.arange(w.shape[1], dtype=np.uint64) + 1)).sum(axis=1, dtype=np.uint64)
        out = np.zeros(len(frame_offsets) - 1, dtype=np.uint64)
        for u in range(len(out)):
            a, b = int(frame_offsets[u]), int(frame_offsets[u + 1])
            i = np.arange(b - a, dtype=np.uint64)
            out[u] = (row[a:b] * (2 * i + 1)).sum(dtype=np.uint64)
    return out


def partition(n_frames, world, split="lpt"):
    """Utterance ids per rank.  'lpt': balanced by frame count; 'reference': the contiguous split of aku/Recipe.cc:63-115."""
    if split == "lpt":
        return lpt_partition(n_frames, world)
    if split == "reference":
        return reference_partition(len(n_frames), world)
    raise ValueError("split must be 'lpt' or 'reference'")


def sub_batches(n_frames, max_frames):
    """Consecutive utterances of a rank's list grouped into sub-batches of at most max_frames frames (one scorer chunk;
    at least one utterance each).  Deterministic: the writer rank derives every sender's schedule from the global
    table with it.  Returns a list of (first_utt, end_utt, first_frame, n_frames) in the rank's own numbering."""
    out, a, f0, acc = [], 0, 0, 0
    for i, n in enumerate(int(x) for x in n_frames):
        if i > a and acc + n > max_frames:
            out.append((a, i, f0, acc))
            a, f0, acc = i, f0 + acc, 0
        acc += n
    if len(n_frames) > a:
        out.append((a, len(n_frames), f0, acc))
    return out


def _dev(group=None):
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


# ---------------------------------------------------------------------------------------------- collectives
MODEL_KEYS = ("mix_offsets", "mix_gauss", "mix_weight", "means", "covs")


def broadcast_model(model, src=0, group=None):
    """The model arrays of rank `src` on every rank: ONE broadcast of a packed float64 buffer (header: the five array
    lengths and D) -- the ncclBroadcast of packed parameters of SURVEY.md section 8e.  Returns (model dict, bytes)."""
    import torch
    import torch.distributed as dist
    dev = _dev(group)
    rank = dist.get_rank(group)
    head = torch.zeros(8, dtype=torch.float64, device=dev)
    if rank == src:
        arrs = [np.ascontiguousarray(model[k], dtype=np.float64).reshape(-1) for k in MODEL_KEYS]
        D = np.asarray(model["means"]).shape[1]
        head[:6] = torch.tensor([a.size for a in arrs] + [D], dtype=torch.float64)
    dist.broadcast(head, src, group=group)
    sizes = [int(x) for x in head[:5].tolist()]
    D = int(head[5].item())
    buf = torch.empty(sum(sizes), dtype=torch.float64, device=dev)
    if rank == src:
        buf.copy_(torch.from_numpy(np.concatenate(arrs)))
    dist.broadcast(buf, src, group=group)
    flat = buf.cpu().numpy()
    out, o = {}, 0
    for k, n in zip(MODEL_KEYS, sizes):
        out[k] = flat[o:o + n]
        o += n
    out["mix_offsets"] = out["mix_offsets"].astype(np.int32)
    out["mix_gauss"] = out["mix_gauss"].astype(np.int32)
    out["means"] = out["means"].reshape(-1, D)
    out["covs"] = out["covs"].reshape(-1, D)
    return out, int(buf.numel() * 8)


def gather_frame_counts(local_ids, local_counts, n_utts, group=None):
    """All-gather of int32 n_frames[utt]: rank r contributes the utterances it looked at (e.g. ids r, r+world, ...)."""
    import torch
    import torch.distributed as dist
    dev = _dev(group)
    world = dist.get_world_size(group)
    m = (n_utts + world - 1) // world
    buf = torch.full((m, 2), -1, dtype=torch.int32, device=dev)
    k = len(local_ids)
    if k > m:
        raise ValueError("a rank may contribute at most ceil(n_utts / world) utterances")
    if k:
        buf[:k, 0] = torch.as_tensor(np.asarray(local_ids, dtype=np.int32), device=dev)
        buf[:k, 1] = torch.as_tensor(np.asarray(local_counts, dtype=np.int32), device=dev)
    out = torch.empty((world, m, 2), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(out.view(-1, 2), buf, group=group)
    a = out.cpu().numpy().reshape(-1, 2)
    a = a[a[:, 0] >= 0]
    frames = np.zeros(n_utts, dtype=np.int64)
    frames[a[:, 0]] = a[:, 1]
    if len(a) != n_utts:
        raise RuntimeError("frame-count table incomplete: %d of %d utterances" % (len(a), n_utts))
    return frames


# ---------------------------------------------------------------------------------------------- payload: per-rank writers
def run_writers(produce, n_frames, max_frames, slots):
    """Every rank drains its own records: sub-batch k is scored into slots[k % len(slots)] (device slots: records stay
    resident; pinned host slots: the library streams them out while it scores the next chunk).
    produce(u0, u1, out) scores utterances [u0, u1) of this rank's list into `out`.  Returns the number of sub-batches."""
    sched = sub_batches(n_frames, max_frames)
    for k, (u0, u1, _, _) in enumerate(sched):
        produce(u0, u1, slots[k % len(slots)])
    return len(sched)


# ---------------------------------------------------------------------------------------------- payload: gather to a writer rank
class GatherPlan:
    """Everything both sides of the gather derive from the global (utterance, n_frames) table and the partition."""

    def __init__(self, n_frames_global, parts, max_frames, rec_bytes, writer=0):
        self.parts = [np.asarray(p, dtype=np.int64) for p in parts]
        self.world = len(parts)
        self.writer = writer
        self.rec_bytes = int(rec_bytes)
        self.max_frames = int(max_frames)
        self.counts = [np.asarray(n_frames_global, dtype=np.int64)[p] for p in self.parts]
        self.sched = [sub_batches(c, max_frames) for c in self.counts]
        self.fo = [np.concatenate([[0], np.cumsum(c)]).astype(np.int64) for c in self.counts]   # per-rank stream numbering
        self.slot_frames = max([max_frames] + [s[3] for sc in self.sched for s in sc])          # a long utterance may exceed the target
        self.slot_bytes = self.slot_frames * self.rec_bytes
        self.senders = [r for r in range(self.world) if r != writer]

    def rounds(self):
        return max(len(s) for s in self.sched)


class _NoScope:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def gather_nccl(plan, rank, produce, sink, send_slots, recv_slots, group=None, sink_scope=_NoScope):
    """records -> local send slot -> send/recv -> the writer's slot -> sink.
    produce(u0, u1, out_tensor)            scores a sub-batch of THIS rank into a tensor (device)
    sink(r, slot_tensor, first_frame, n)   consumes frames [first_frame, first_frame + n) of sender r's stream (writer only;
                                           the writer's own records never leave its device and are not passed through it)
    send_slots: 2 uint8 tensors of plan.slot_bytes; recv_slots[r]: 2 uint8 tensors per sender r (writer only)
    sink_scope: context manager factory entered around the writer's receive / sink work (a second CUDA stream on
                GPUs, so that receiving overlaps the writer's own scoring)."""
    import torch.distributed as dist
    R = plan.rec_bytes
    mine = plan.sched[rank]
    if rank != plan.writer:
        works = [None, None]
        for k, (u0, u1, f0, n) in enumerate(mine):
            if works[k & 1] is not None:
                works[k & 1].wait()                              # the slot's previous send has left
            produce(u0, u1, send_slots[k & 1])
            works[k & 1] = dist.isend(send_slots[k & 1][:n * R], plan.writer, group=group)
        for w in works:
            if w is not None:
                w.wait()
        return
    for k in range(plan.rounds()):
        pend = []
        with sink_scope():                                       # post this round's receives, then score the own share
            for r in plan.senders:
                if k < len(plan.sched[r]):
                    n = plan.sched[r][k][3]
                    pend.append((r, dist.irecv(recv_slots[r][k & 1][:n * R], r, group=group)))
        if k < len(mine):
            u0, u1, f0, n = mine[k]
            produce(u0, u1, send_slots[k & 1])
        with sink_scope():
            for r, w in pend:
                w.wait()
                _, _, f0, n = plan.sched[r][k]
                sink(r, recv_slots[r][k & 1], f0, n)


def gather_p2p(plan, rank, produce, sink, own_slots, peer_slot, nslots, g_tok, g_free, token, sink_scope=_NoScope,
               tok_scope=_NoScope):
    """The LNA kernel stores into the writer's memory: produce(u0, u1, out) is handed the mapped peer slot.
    peer_slot(r, j)  -> what `produce` / `sink` take for slot j of sender r (a DevPtr into the shared buffer)
    own_slots        the writer's buffers for its own records
    g_tok / g_free   two process groups carrying the 8-byte "slot filled" (sender -> writer) and "slot free"
                     (writer -> sender) tokens: separate communicators, so neither direction can block the other
    token            a small tensor on the right device
    tok_scope        context manager factory entered around a sender's "slot filled" token (copy-engine variant: the
                     stream the asynchronous copy was issued on, so that the token follows the copy, not the scoring)."""
    import torch.distributed as dist
    mine = plan.sched[rank]
    # Tokens are sent asynchronously and collected at the end: with blocking sends the two directions wait for each
    # other under rendezvous semantics (sender in "filled k+1", writer in "free k" -- seen with gloo).
    works = []
    tok_in = token.clone()                                       # received tokens land here, `token` is only ever sent
    if rank != plan.writer:
        for k, (u0, u1, f0, n) in enumerate(mine):
            if k >= nslots:
                dist.recv(tok_in, plan.writer, group=g_free)      # the writer has consumed sub-batch k - nslots
            produce(u0, u1, peer_slot(rank, k % nslots))       # returns when the records are in / on their way to the writer's memory
            with tok_scope():
                works.append(dist.isend(token, plan.writer, group=g_tok))
    else:
        for k in range(plan.rounds()):
            if k < len(mine):
                u0, u1, f0, n = mine[k]
                produce(u0, u1, own_slots[k % len(own_slots)])
            with sink_scope():
                for r in plan.senders:
                    if k < len(plan.sched[r]):
                        dist.recv(tok_in, r, group=g_tok)
                        _, _, f0, n = plan.sched[r][k]
                        sink(r, peer_slot(r, k % nslots), f0, n)
                        if k + nslots < len(plan.sched[r]):
                            works.append(dist.isend(token, r, group=g_free))
    for w in works:
        w.wait()
