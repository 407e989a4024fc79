"""Utterance sharding across the GPUs of one box.

The reference is single-process / single-device; utterances are independent and the
model's results are batch-invariant (MaskConv + packed sequences, model.py:57-58), so the path shards
naturally: every rank holds a replica of the weights (and LM tables), gets a subset of the utterances
and runs them in length-sorted batches.  There is NO collective on the data path -- only a final host
gather of the transcript strings (``torch.distributed.all_gather_object``).
"""


def lpt_shards(lengths, n_ranks):
    """Longest-processing-time-first bin packing on length (cost is proportional to frames).

    Returns ``n_ranks`` lists of utterance indices; deterministic (ties broken by index)."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * n_ranks
    shards = [[] for _ in range(n_ranks)]
    for i in order:
        r = min(range(n_ranks), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += int(lengths[i])
    return shards


def make_batches(indices, lengths, max_batch=64, max_samples=None):
    """Sort a shard by length (descending, the pack_padded_sequence contract) and cut it into batches of at
    most ``max_batch`` utterances and, optionally, ``max_samples`` padded samples."""
    order = sorted(indices, key=lambda i: (-int(lengths[i]), i))
    batches, cur = [], []
    for i in order:
        longest = int(lengths[cur[0]]) if cur else int(lengths[i])
        if cur and (len(cur) >= max_batch or (max_samples and (len(cur) + 1) * longest > max_samples)):
            batches.append(cur)
            cur = []
        cur.append(i)
    if cur:
        batches.append(cur)
    return batches


def gather_transcripts(local, world_size, group=None):
    """``local``: {utterance index: result}.  Returns the merged dict on every rank (host-side gather)."""
    if world_size == 1:
        return dict(local)
    import torch.distributed as dist
    parts = [None] * world_size
    dist.all_gather_object(parts, local, group=group)
    merged = {}
    for p in parts:
        merged.update(p)
    return merged


def transcribe_sharded(recognize_batch, recordings, rank=0, world_size=1, max_batch=64, group=None,
                       recognize_batches=None):
    """Shard ``recordings`` over ``world_size`` ranks, run this rank's share through ``recognize_batch``
    (a callable list-of-audio -> list-of-results) and gather.  The result list is in input order and is
    identical for every world size (shard invariance).  ``recognize_batches`` (list of batches -> list of result
    lists, e.g. ``Recognizer.recognize_batches``) overlaps the host staging of a batch with the previous one."""
    lengths = [len(r) for r in recordings]
    mine = lpt_shards(lengths, world_size)[rank]
    local = {}
    batches = make_batches(mine, lengths, max_batch=max_batch)
    if recognize_batches is not None:
        outs = recognize_batches([[recordings[i] for i in batch] for batch in batches])
    else:
        outs = [recognize_batch([recordings[i] for i in batch]) for batch in batches]
    for batch, out in zip(batches, outs):
        for i, o in zip(batch, out):
            local[i] = o
    merged = gather_transcripts(local, world_size, group=group)
    return [merged[i] for i in range(len(recordings))]
