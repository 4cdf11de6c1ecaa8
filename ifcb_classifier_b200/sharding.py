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


# ---------------------------------------------------------------------------------------------
# TRAIN: data-parallel gradient exchange (Lightning DDP in the reference, neuston_net.py:101-107:
# per-rank batches, gradient MEAN over ranks, per-rank BatchNorm statistics).
# ---------------------------------------------------------------------------------------------
def plan_buckets(offsets_backward, n_params, bucket_elems):
    """Contiguous gradient-arena ranges to all-reduce, in the order the backward pass completes them.

    ``offsets_backward``: for each backward unit in execution order, the arena offset of its first
    parameter (parameters are laid out in forward order, so offsets decrease), or None for units
    without parameters.  Returns [(unit index after which the range is final, lo, hi), ...]; the
    ranges tile [0, n_params) exactly and each (but possibly the last) holds >= bucket_elems."""
    marks, hi = [], n_params
    for i, lo in enumerate(offsets_backward):
        if lo is None:
            continue
        assert 0 <= lo <= hi, 'parameters must be registered in forward order'
        if hi - lo >= bucket_elems:
            marks.append((i, lo, hi))
            hi = lo
    if hi > 0:
        marks.append((len(offsets_backward) - 1, 0, hi))
    return marks


class GradReducer(object):
    """Launches one asynchronous SUM all-reduce per finished bucket (NCCL runs them on its own stream
    while the remaining backward kernels execute) and waits for all of them before the optimizer;
    the 1/world factor of the mean is applied inside the Adam kernel (``grad_scale``)."""

    def __init__(self, grads, group=None):
        import torch.distributed as dist
        self.dist, self.grads, self.group = dist, grads, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.works = []

    def __call__(self, lo, hi):
        if self.world > 1 and hi > lo:
            self.works.append(self.dist.all_reduce(self.grads[lo:hi], op=self.dist.ReduceOp.SUM, group=self.group, async_op=True))

    def wait(self):
        for w in self.works:
            w.wait()
        self.works = []
        return 1.0 / self.world
