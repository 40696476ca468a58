"""Walker sharding across the GPUs of one box (one process per GPU) for the device-resident VMC.

The reference parallelises by splitting the walkers over a futures client and averaging the
workers' block averages weighted by their walker counts (``vmc_parallel``, ``pyqmc/method/mc.py:
156-173``; ``configs.split/join``, ``coord.py:72-88``).  Here each rank owns its shard for the whole
run (walkers stay resident on its GPU) and the only communication is ONE ``allreduce(sum)`` per
block of a small statistics vector ``[n_r, n_r * avg_k ...]`` over NCCL/NVLink (``gloo`` on CPU
for the tests) -- no data-path collective.  The combined averages equal the reference's weighted
mean: sum_r avg_r * n_r / sum_r n_r.
"""
import numpy as np

SKIP_KEYS = ("block", "nconfig", "move time", "accumulator time")


def shard(configs, rank, world):
    """This rank's walkers: the ``rank``-th piece of ``configs.split(world)`` (np.array_split)."""
    return configs.split(world)[rank]


def pack_block(block_avg, nconf):
    keys = sorted(k for k in block_avg if k not in SKIP_KEYS)
    vec = np.empty(1 + len(keys))
    vec[0] = nconf
    for i, k in enumerate(keys):
        vec[1 + i] = nconf * float(block_avg[k])
    return keys, vec


def unpack_block(keys, vec):
    total = vec[0]
    return {k: vec[1 + i] / total for i, k in enumerate(keys)}, int(round(total))


def allreduce_block(block_avg, nconf, group=None, device=None):
    """One collective per block.  Returns the walker-weighted global averages and the total count."""
    import torch
    import torch.distributed as dist

    keys, vec = pack_block(block_avg, nconf)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return unpack_block(keys, vec)
    if device is None:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.from_numpy(vec).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return unpack_block(keys, t.cpu().numpy())


def vmc_distributed(wf, configs, tstep=0.5, nblocks=10, nsteps_per_block=10, accumulators=None, group=None,
                    block_fn=None, seed=None):
    """Device-resident VMC on this rank's shard with one allreduce per block.

    ``configs`` holds the GLOBAL walkers on every rank (cheap: host numpy); the shard is cut with
    ``shard``.  Returns (df of global block averages -- identical on all ranks, local configs).
    ``block_fn(wf, configs, tstep, nsteps, accumulators) -> (block_avg, configs)`` defaults to the
    device-resident block of ``pyqmc_b200.mc``; each rank draws from its own ``np.random`` stream
    (seeded with ``seed + rank`` when given), as the reference's workers do."""
    import torch.distributed as dist

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if block_fn is None:
        from .mc import vmc_block_device as block_fn
    if seed is not None:
        np.random.seed(seed + rank)
    local = shard(configs, rank, world)
    accumulators = {} if accumulators is None else accumulators
    rows = []
    for block in range(nblocks):
        avg, local = block_fn(wf, local, tstep, nsteps_per_block, accumulators)
        glob, total = allreduce_block(avg, local.configs.shape[0], group)
        glob["block"] = block
        glob["nconfig"] = nsteps_per_block * total
        rows.append(glob)
    df = {k: np.asarray([r[k] for r in rows]) for k in rows[0]} if rows else {}
    return df, local


def gather_configs(local, group=None):
    """All walkers on every rank (rank order = ``configs.split`` order), e.g. for checkpointing."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local.configs.copy()
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, local.configs, group=group)
    return np.concatenate(parts, axis=0)


# ---- DMC: weighted block statistics and global branching ------------------------------------------
def allreduce_dmc_block(block_avg, nconf, group=None, device=None):
    """Combination rule of ``dmc_propagate_parallel`` (dmc.py:238-303) with one allreduce: every rank
    contributes ``[n_p, w_p n_p, <O>_p w_p n_p ...]``; the global weight is ``sum w_p n_p / sum n_p``
    and the observables are averaged with the weights ``w_p n_p``."""
    import torch
    import torch.distributed as dist

    keys = sorted(k for k in block_avg if k not in SKIP_KEYS and k != "weight")
    wn = float(block_avg["weight"]) * nconf
    vec = np.array([nconf, wn] + [float(block_avg[k]) * wn for k in keys])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if device is None:
            device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.from_numpy(vec).to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        vec = t.cpu().numpy()
    out = {k: vec[2 + i] / vec[1] for i, k in enumerate(keys)}
    out["weight"] = vec[1] / vec[0]
    return out, int(round(vec[0]))


def branch_global(local, weights, group=None, base_draw=None):
    """Stochastic-comb branching over the GLOBAL population (``branch``, dmc.py:342-376, applied after
    ``configs.join`` as the reference's parallel driver does): the shards' walkers and weights are
    gathered (a few hundred KB), rank 0's ``np.random.rand()`` fixes the comb offset for everybody,
    every rank computes the same resampling indices and keeps its ``np.array_split`` slice of the
    resampled population.  Returns (local configs, local weights, info)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        from .dmc import branch

        return branch(local, weights, base_draw)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = [None] * world
    payload = (local.configs, getattr(local, "wrap", None), np.asarray(weights))
    dist.all_gather_object(parts, payload, group=group)
    allc = np.concatenate([p[0] for p in parts], axis=0)
    allw = np.concatenate([p[2] for p in parts])
    box = [(np.random.rand() if base_draw is None else base_draw) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    nconfig = len(allw)
    probability = np.cumsum(allw)
    wtot = probability[-1]
    base = box[0] * wtot
    newinds = np.searchsorted(probability, (base + np.linspace(0, wtot, nconfig, endpoint=False)) % wtot)
    unique, counts = np.unique(newinds, return_counts=True)
    mine = np.array_split(newinds, world)[rank]
    local.configs = allc[mine]
    if parts[0][1] is not None:
        local.wrap = np.concatenate([p[1] for p in parts], axis=0)[mine]
    new_w = np.full(len(mine), wtot / nconfig)
    return local, new_w, {"max branches": np.max(counts), "Number of walkers killed": nconfig - unique.shape[0]}
