"""Multi-GPU plumbing for the one-process-per-GPU launch (torchrun): sample sharding
and the single collective of the path -- a sum all-reduce of the uint64 counters
(n_pass, n_total, fail_per_spec[], hist[]).  Samples are independent units; the
Philox counter carries the GLOBAL sample index, so the combined counters do not
depend on the number of ranks.  The frequency axis is never split across ranks.
"""
import os


def shard_range(n_samples, rank, world):
    """Contiguous global sample range [lo, hi) of `rank`: lo = floor(rank*N/W)."""
    lo = n_samples * rank // world
    hi = n_samples * (rank + 1) // world
    return lo, hi


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend):
    """Initialise torch.distributed from the torchrun environment (MASTER_ADDR/PORT, RANK, WORLD_SIZE)."""
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group(backend=backend)
    return dist


def allreduce_counters(counters):
    """Sum the int64 counter tensor over all ranks in place (NCCL on GPUs, gloo on CPU).
    Integer sums keep the yield bit-reproducible and independent of the rank count."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def split_counters(counters, nspec, hist_bins):
    c = [int(x) for x in counters.tolist()]
    return dict(n_pass=c[0], n_total=c[1], fail_per_spec=c[2:2 + nspec], hist=c[2 + nspec:2 + nspec + hist_bins])
