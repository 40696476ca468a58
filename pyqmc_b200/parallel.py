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
    """[nconf, nconf * value ...] in sorted key order; a complex value (ECP term and total energy of a complex wave
    function) takes two slots, its key listed as ``(name, "imag")`` for the second."""
    keys, vals = [], []
    for k in sorted(k for k in block_avg if k not in SKIP_KEYS):
        v = block_avg[k]
        keys.append(k)
        if np.iscomplexobj(v):
            vals.append(nconf * float(np.real(v)))
            keys.append((k, "imag"))
            vals.append(nconf * float(np.imag(v)))
        else:
            vals.append(nconf * float(v))
    return keys, np.array([float(nconf)] + vals)


def unpack_block(keys, vec):
    total = vec[0]
    out = {}
    for i, k in enumerate(keys):
        if isinstance(k, tuple):
            out[k[0]] = out[k[0]] + 1j * (vec[1 + i] / total)
        else:
            out[k] = vec[1 + i] / total
    return out, int(round(total))


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


def _walker_counts(local, n_local, group, device):
    """Walkers held by every rank.  ``branch_global`` leaves the layout it produced (``np.array_split`` shares of a
    population whose size never changes) on the walker container it returns; a container without that note -- the
    first block, or walkers re-sharded by the caller -- costs one all-reduce.  Every rank runs the same program, so
    all of them take the same branch here."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    note = getattr(local, "_rank_layout", None)
    if note is not None and len(note) == world and note[rank] == n_local:
        return note
    counts_t = torch.zeros(world, dtype=torch.int64, device=device)
    counts_t[rank] = n_local
    dist.all_reduce(counts_t, group=group)
    return counts_t.cpu().numpy()


def branch_global(local, weights, group=None, base_draw=None, block_avg=None):
    """Stochastic-comb branching over the GLOBAL population (``branch``, dmc.py:342-376, applied after
    ``configs.join`` as the reference's parallel driver does) without ever assembling that population:

    1. one all-gather of the WEIGHTS (8 bytes per walker; rank 0's slot also carries its ``np.random.rand()``,
       the comb offset for everybody, and -- with ``block_avg`` -- every rank's block statistics, so the
       combination of ``allreduce_dmc_block`` needs no collective of its own);
    2. every rank computes the same resampling indices (``dmc.comb_indices``) and, from the ``np.array_split``
       layout, which rank holds which walker before and after;
    3. one ``all_to_all_single`` moves only the walkers that change owner (``nelec * 24`` bytes each, 48 with the
       periodic wrap vectors); copies that stay on their rank never leave it.

    Equal to join -> branch -> split of the reference with the same offset (tests/test_parallel_gloo.py).
    Returns (local configs, local weights, info); with ``block_avg`` the info dictionary also holds
    ``"block_avg"``: the globally combined block dictionary (``allreduce_dmc_block``'s first result)."""
    import torch
    import torch.distributed as dist

    from .dmc import branch, comb_indices

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        res = branch(local, weights, base_draw)
        if block_avg is not None:
            res[2]["block_avg"] = allreduce_dmc_block(block_avg, len(weights), group)[0]
        return res
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    weights = np.asarray(weights, dtype=np.float64)
    n_local = len(weights)
    counts = _walker_counts(local, n_local, group, device)
    stat_keys = sorted(k for k in block_avg if k not in SKIP_KEYS and k != "weight") if block_avg is not None else []
    nstat = 2 + len(stat_keys) if block_avg is not None else 0
    width = int(counts.max()) + 1 + nstat  # after the weights: the comb offset (rank 0's draw), then the statistics
    row = np.zeros(width)
    row[:n_local] = weights
    if rank == 0:
        row[int(counts.max())] = np.random.rand() if base_draw is None else base_draw
    if block_avg is not None:
        wn = float(block_avg["weight"]) * n_local
        row[width - nstat:] = [n_local, wn] + [float(block_avg[k]) * wn for k in stat_keys]
    if device == "cuda":  # NCCL: one output tensor, no per-rank pieces to allocate and stack
        gathered_t = torch.empty((world, width), dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(gathered_t, torch.from_numpy(row).to(device), group=group)
        gathered = gathered_t.cpu().numpy()
    else:
        pieces = [torch.empty(width, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(pieces, torch.from_numpy(row), group=group)
        gathered = torch.stack(pieces).numpy()
    allw = np.concatenate([gathered[r, : counts[r]] for r in range(world)])
    picked, total = comb_indices(allw, gathered[0, int(counts.max())])
    combined = None
    if block_avg is not None:  # the rule of allreduce_dmc_block on the gathered per-rank vectors
        vec = gathered[:, width - nstat:].sum(axis=0)
        combined = {k: vec[2 + i] / vec[1] for i, k in enumerate(stat_keys)}
        combined["weight"] = vec[1] / vec[0]
    if np.any(allw > 2.0) and rank == 0:
        import logging

        logging.warning("Some weights are larger than 2")
    # ownership before (contiguous ranges of global ids) and after (array_split of the resampled population: slot i
    # of the new population belongs to the rank whose contiguous slot range holds i)
    first = np.concatenate([[0], np.cumsum(counts)])
    owner = np.searchsorted(first, picked, side="right") - 1
    total_n = len(picked)
    share = np.array([total_n // world + (1 if r < total_n % world else 0) for r in range(world)])  # np.array_split sizes
    bounds = np.concatenate([[0], np.cumsum(share)])
    dest = np.repeat(np.arange(world), share)
    fields = [local.configs.reshape(n_local, -1)]
    if getattr(local, "wrap", None) is not None:
        fields.append(local.wrap.reshape(n_local, -1))
    rows = fields[0] if len(fields) == 1 else np.concatenate(fields, axis=1)
    width_r = rows.shape[1]
    # what this rank sends: its walkers picked for slots of other ranks, in slot order (= grouped by destination)
    leaving = np.flatnonzero((owner == rank) & (dest != rank))
    send_counts = np.bincount(dest[leaving], minlength=world).tolist()
    send = rows[picked[leaving] - first[rank]]
    # what it keeps / receives: its own slot range
    lo, hi = bounds[rank], bounds[rank + 1]
    my_ids, my_owner = picked[lo:hi], owner[lo:hi]
    stay = my_owner == rank
    out = np.empty((hi - lo, width_r))
    out[stay] = rows[my_ids[stay] - first[rank]]  # copies that never leave this rank
    arriving = np.flatnonzero(~stay)
    recv_counts = np.bincount(my_owner[arriving], minlength=world).tolist()
    send_t = torch.from_numpy(np.ascontiguousarray(send)).to(device)
    recv_t = torch.empty((len(arriving), width_r), dtype=torch.float64, device=device)
    dist.all_to_all_single(recv_t, send_t, output_split_sizes=recv_counts, input_split_sizes=send_counts, group=group)
    # the pieces arrive grouped by source rank, each in the sender's slot order
    out[arriving[np.argsort(my_owner[arriving], kind="stable")]] = recv_t.cpu().numpy()
    ncfg = fields[0].shape[1]
    local.configs = np.ascontiguousarray(out[:, :ncfg]).reshape((len(out),) + local.configs.shape[1:])
    if len(fields) > 1:
        local.wrap = np.ascontiguousarray(out[:, ncfg:]).reshape(local.configs.shape)
    copies = np.bincount(picked, minlength=len(picked))
    new_w = np.full(len(out), total / len(picked))
    local._rank_layout = share
    info = {"max branches": np.max(copies), "Number of walkers killed": int(np.count_nonzero(copies == 0))}
    if combined is not None:
        info["block_avg"] = combined
    return local, new_w, info
