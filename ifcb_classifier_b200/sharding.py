"""Bin sharding for multi-GPU RUN: bins are independent units (each bin -> its own
dataset, forward passes and output file: reference neuston_net.py:233-268), so ranks take
disjoint bins and no collective touches the data path.  Assignment is deterministic
(sorted enumeration, longest-processing-time-first on ROI count) so every rank computes
the same partition without communicating."""
import os


def bin_cost(basepath):
    """Cheap proxy for a bin's work = number of ROI rows (newline count of the .adc)."""
    try:
        with open(basepath + '.adc', 'rb') as f:
            return max(1, f.read().count(b'\n'))
    except OSError:
        return 1


def assign(costs, world):
    """LPT: sort by (-cost, index), give each item to the least-loaded rank (ties -> lowest
    rank).  Returns list of rank per item."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * world
    owner = [0] * len(costs)
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += costs[i]
    return owner


def my_bins(basepaths, rank, world):
    basepaths = sorted(basepaths)
    if world <= 1:
        return list(basepaths)
    owner = assign([bin_cost(b) for b in basepaths], world)
    return [b for b, o in zip(basepaths, owner) if o == rank]


def gather_summary(summary, world, backend_group=None):
    """All ranks -> rank 0: list of per-rank summaries (n_bins, n_rois, seconds, error_bins).
    The only collective of RUN, after the data path has finished."""
    if world <= 1:
        return [summary]
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, summary, group=backend_group)
    return out


def env_rank_world():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))
