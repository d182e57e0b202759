"""Multi-GPU plumbing of the path: VFOs are independent consumers of one IQ stream, so the VFO
set is sharded across ranks and the IQ block is broadcast from the ingest rank; no other collective exists.
torch.distributed is plumbing only (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""


def vfo_cost(in_sr, out_sr, bw, design_resampler, design_decim_plan):
    """Approximate FMA per input sample of one VFO (stage 1 dominates): used to balance shards."""
    info, _ = design_resampler(in_sr, out_sr)
    cost, rate = 4.0, 1.0
    if info["mode"] in (0, 1):
        for d, taps in design_decim_plan(info["predec"]):
            cost += 2.0 * len(taps) * rate / d
            rate /= d
    if info["mode"] in (0, 2):
        cost += 2.0 * info["tpp"] * out_sr / in_sr
    if bw != out_sr:
        cost += 2.0 * int(3.8 * out_sr / (bw / 20.0)) * out_sr / in_sr
    return cost


def shard_vfos(costs, world, base_load=None):
    """Greedy longest-processing-time partition: returns, per rank, the list of VFO indices it owns.
    Deterministic (ties by index), every VFO in exactly one shard. base_load: work a rank already has
    (rank 0 ingests, broadcasts and computes the spectrum), in the same units as costs."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    if base_load:
        for r, b in enumerate(base_load[:world]):
            load[r] = float(b)
    shards = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += costs[i]
    return [sorted(s) for s in shards]


def broadcast_block(tensor, src=0):
    """Broadcast one IQ block (any dtype/shape tensor) from the ingest rank to every rank."""
    import torch.distributed as dist
    dist.broadcast(tensor, src=src)
    return tensor


def bucket_slot(i, bucket_blocks, source_blocks):
    """Where block i of a phase lives when the stream is broadcast in buckets of `bucket_blocks` consecutive blocks out
    of a circular source of `source_blocks` blocks (a multiple of the bucket size): (staging bucket 0/1, index inside the
    bucket, index of the bucket's first block in the source)."""
    assert bucket_blocks >= 1 and source_blocks % bucket_blocks == 0
    j = i % bucket_blocks
    return (i // bucket_blocks) % 2, j, (i - j) % source_blocks


def fit_bucket(bucket_blocks, source_blocks):
    """Largest bucket size <= the requested one that divides the source length."""
    kb = max(1, min(int(bucket_blocks), int(source_blocks)))
    while source_blocks % kb:
        kb -= 1
    return kb
