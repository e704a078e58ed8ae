"""Ray sharding and the two collectives of the multi-GPU trace loop.

The reference is single-device (one OpenCL queue, iterative_tracer.py:60,74); its only
partitioning is the <=50,000-ray chunk loop (it:218-273), which shows that rays are
independent.  Here the ray axis is sharded over one process per GPU:

  * initial rays are split into contiguous blocks (``shard_bounds``), the scene is replicated,
    children stay on the GPU that produced them - no data-path collective;
  * per iteration ONE all-reduce of (live ray count, live power) implements the reference's
    global termination tests (it:372-391) - 16 bytes, latency bound;
  * at the end ONE padded all-gather of the measured records (x, y, z, power) + lineage
    implements ``get_measured_rays`` (it:395-411); merging by (iteration, lineage) restores the
    reference's order, so 1/2/4/8-GPU runs return identical arrays;
  * the 64-bit lineage holds PATH_BITS branch bits above a 32-bit index.  Before they run out
    (every PATH_BITS iterations) ``renumber_lineage`` replaces every live ray's id by its rank in
    the generation's global reference order (one all-gather of the live ids) and the branch bits
    start over - traces of any depth keep the reference order (reference default: 100 iterations,
    it:77).

All helpers work on CPU tensors with the ``gloo`` backend too (used by the CPU tests).
"""

from __future__ import annotations

import os
import time

import torch


def _td():
    import torch.distributed as td

    return td


def resolve_group(process_group=None):
    """Return a marker for "distributed is active": the given group, or the string "world" if
    torch.distributed is initialised, else None (single process)."""
    td = _td()
    if process_group is not None:
        return process_group
    if td.is_available() and td.is_initialized():
        return "world"
    return None


def _pg(group):
    return None if group == "world" else group


def world(group):
    if group is None:
        return 1, 0
    td = _td()
    return td.get_world_size(_pg(group)), td.get_rank(_pg(group))


def _comm_device(group, device):
    """NCCL needs CUDA tensors; gloo works on CPU tensors."""
    td = _td()
    backend = str(td.get_backend(_pg(group))).lower()
    return torch.device(device) if "nccl" in backend else torch.device("cpu")


PATH_BITS = 31  # branch bits of a 64-bit lineage: id = (path << 32) | index, bit k of path = "refracted at epoch iteration k"


def shard_bounds(n, world_size, rank):
    """Contiguous block [lo, hi) of rank ``rank`` out of ``n`` rays (sizes differ by at most 1)."""
    return (n * rank) // world_size, (n * (rank + 1)) // world_size


def all_gather_int(value, group, device):
    td = _td()
    ws, _ = world(group)
    dev = _comm_device(group, device)
    mine = torch.tensor([int(value)], dtype=torch.int64, device=dev)
    out = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(ws)]
    td.all_gather(out, mine, group=_pg(group))
    return [int(t.item()) for t in out]


def all_reduce_sum(value, group, device):
    td = _td()
    t = torch.tensor([float(value)], dtype=torch.float64, device=_comm_device(group, device))
    td.all_reduce(t, op=td.ReduceOp.SUM, group=_pg(group))
    return float(t.item())


def all_reduce_array(arr, group, device):
    """all-reduce(sum) of a NumPy float64 array (the per-rank histograms of the binning step)."""
    import numpy as np

    td = _td()
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(_comm_device(group, device))
    td.all_reduce(t, op=td.ReduceOp.SUM, group=_pg(group))
    return t.cpu().numpy()


def all_reduce_pair(live, power, group, device):
    """One all-reduce(sum) of {live ray count, live power} (counts are exact in float64 < 2^53)."""
    td = _td()
    t = torch.tensor([float(live), float(power)], dtype=torch.float64, device=_comm_device(group, device))
    td.all_reduce(t, op=td.ReduceOp.SUM, group=_pg(group))
    live_sum, power_sum = t.tolist()  # one device -> host read for both numbers
    return int(round(live_sum)), float(power_sum)


def renumber_lineage(keys, group, device):
    """Start a new lineage epoch: replace every live ray's id by its rank in the GLOBAL reference
    order of the current generation (a dense index < 2^32, branch bits cleared).

    ``keys``: this rank's ids in its local order.  Inside one generation the reference order is the
    ascending id order and a rank's rays keep it (the compaction is order preserving), so ``keys`` is
    strictly increasing and the global rank of a ray is its local index plus, for every other rank,
    the number of that rank's ids below it - one padded all-gather and world_size searches.
    Collective: every rank of the group calls it at the same iteration, also with zero rays."""
    n = int(keys.shape[0])
    if group is None:
        return torch.arange(n, dtype=torch.int64, device=keys.device)
    td = _td()
    ws, rank = world(group)
    dev = _comm_device(group, device)
    counts = all_gather_int(n, group, device)
    if sum(counts) >= (1 << 32):
        raise ValueError("lineage renumbering needs < 2^32 live rays in a generation (got %d)" % sum(counts))
    kmax = max(max(counts), 1)
    mine = torch.full((kmax,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=dev)
    local = keys.to(dev)
    if n:
        mine[:n] = local
    everyone = [torch.empty_like(mine) for _ in range(ws)]
    td.all_gather(everyone, mine, group=_pg(group))
    new = torch.arange(n, dtype=torch.int64, device=dev)
    for r in range(ws):
        if r != rank and counts[r] and n:
            new += torch.searchsorted(everyone[r][: counts[r]], local)
    return new.to(keys.device)


def merge_measured(records, lineages, iter_ids):
    """Order records by (iteration, lineage): the reference's get_measured_rays order.
    Two stable sorts (least-significant key first)."""
    order = torch.sort(lineages, stable=True).indices
    order = order[torch.sort(iter_ids[order], stable=True).indices]
    return records[order], lineages[order], iter_ids[order]


GATHER_TIMINGS = {}  # LPCL_GATHER_TIMING=1: seconds per phase of the last gather_measured call (synchronising: diagnostics only)


def gather_measured(pos_pwr, lineage, per_iter, group, device):
    """All ranks receive every rank's measured records, merged into the reference order.

    pos_pwr (K,4) f32, lineage (K,) i64, per_iter: this rank's measured count per iteration (every rank ran the
    same number of iterations: the termination tests are global).  Two padded all-gathers (records, lineage) and
    one small one (the per-iteration counts, from which every rank rebuilds the iteration of each record)."""
    td = _td()
    ws, _ = world(group)
    dev = _comm_device(group, device)
    timing = os.environ.get("LPCL_GATHER_TIMING") == "1"
    marks = []

    def mark(name):
        if timing:
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            marks.append((name, time.perf_counter()))

    mark("start")
    k = int(pos_pwr.shape[0])
    n_iter = len(per_iter)
    mine = torch.tensor(list(per_iter) + [k], dtype=torch.int64, device=dev)  # per-iteration counts, then the total
    table = [torch.empty_like(mine) for _ in range(ws)]
    td.all_gather(table, mine, group=_pg(group))
    table = torch.stack(table).cpu()
    mark("counts (first collective: waits for the slowest rank)")
    counts = [int(c) for c in table[:, n_iter]]
    kmax = max(max(counts), 1)
    rec = torch.zeros((kmax, 4), dtype=torch.float32, device=dev)
    lin = torch.zeros(kmax, dtype=torch.int64, device=dev)
    if k:
        rec[:k] = pos_pwr.to(dev)
        lin[:k] = lineage.to(dev)
    recs = [torch.empty_like(rec) for _ in range(ws)]
    lins = [torch.empty_like(lin) for _ in range(ws)]
    td.all_gather(recs, rec, group=_pg(group))
    td.all_gather(lins, lin, group=_pg(group))
    mark("all-gather of records and lineages")
    rec_all = torch.cat([r[:c] for r, c in zip(recs, counts)], dim=0)
    lin_all = torch.cat([r[:c] for r, c in zip(lins, counts)], dim=0)
    # records are iteration-major inside a rank: rank r contributes table[r, i] records of iteration i, in that order.
    # The iteration ids are expanded on the communication device (no host pass over the records), as int16: the
    # second, stable sort then needs a quarter of the radix passes of a 64-bit key
    total = int(sum(counts))
    iters = torch.arange(n_iter, dtype=torch.int16, device=dev).repeat(ws)
    itr_all = torch.repeat_interleave(iters, table[:, :n_iter].reshape(-1).to(dev), output_size=total)
    order = torch.sort(lin_all, stable=True).indices
    order = order[torch.sort(itr_all[order], stable=True).indices]
    out = rec_all[order]
    mark("merge (two stable sorts)")
    if timing:
        GATHER_TIMINGS.clear()
        GATHER_TIMINGS.update({b[0]: b[1] - a[1] for a, b in zip(marks, marks[1:])})
    return out
