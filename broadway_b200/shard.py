"""Multi-GPU partitioning of the decode path: the path shards by INDEPENDENT
units — whole streams, or IDR-bounded GOP segments of one stream — and has no exchange step,
so there is no data-path collective: every rank (one process per GPU) decodes its own units on
its own engine, and the only cross-rank traffic is the gather of results (digests / frame
counts; frames themselves stay in each rank's pinned host memory).  torch.distributed is
plumbing only: `nccl` on the GPU box, `gloo` in the CPU tests.

Reference analogue: TestBenchMultipleInstance.c:60-350 (independent instances) and the DPB
flush at IDR pictures (h264bsd_dpb.c:675-708) that makes a GOP segment self-contained.
"""


def assign(units_cost, world_size):
    """Greedy longest-processing-time assignment of units (cost = bytes) to ranks.
    Returns a list (per rank) of unit indices, each list in ascending order.  Deterministic."""
    order = sorted(range(len(units_cost)), key=lambda i: (-units_cost[i], i))
    load = [0] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += units_cost[i]
    return [sorted(x) for x in out]


def decode_sharded(units, decode_fn, rank, world_size, dist=None):
    """Decode `units` (list of Annex-B byte strings: streams or GOP segments) across the ranks.

    decode_fn(list_of_units) -> list (per unit) of per-picture results (e.g. MD5 strings).
    Every rank returns the complete per-unit result list in unit order (all_gather_object);
    with dist None (single process) nothing is exchanged."""
    plan = assign([len(u) for u in units], world_size)
    mine = plan[rank]
    local = decode_fn([units[i] for i in mine]) if mine else []
    if dist is None or world_size == 1:
        gathered = [list(zip(mine, local))]
    else:
        gathered = [None] * world_size
        dist.all_gather_object(gathered, list(zip(mine, local)))
    out = [None] * len(units)
    for part in gathered:
        for i, res in part:
            out[i] = res
    return out
