"""Frame sharding across the GPUs of one box (replaces the reference's MPI/OpenMP drivers, main_mpi.c / main_openmp.c).

Frames are independent, so rank g of G owns the contiguous range [floor(N g / G), floor(N (g+1) / G)) and there is
NO collective on the estimation path.  The only exchange is an optional all-reduce (sum / max) of the per-shard
error statistics [sum|H-Href|^2, sum|Href|^2, count, max|H-Href|] after the estimators have run.
"""
import math


def shard_range(n_frames, rank, world_size):
    """Contiguous shard of rank `rank`: [lo, hi)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    lo = (n_frames * rank) // world_size
    hi = (n_frames * (rank + 1)) // world_size
    return lo, hi


class ShardedEstimator:
    """Runs `fn(first_frame, n_local)` on this rank's shard and reduces statistics over the process group.

    `fn` returns a 4-element float64 tensor [se, sr, count, maxerr] living on the device the process group
    reduces on (CUDA for NCCL, CPU for gloo)."""

    def __init__(self, n_frames, rank=0, world_size=1, group=None):
        self.n_frames, self.rank, self.world_size, self.group = n_frames, rank, world_size, group
        self.lo, self.hi = shard_range(n_frames, rank, world_size)

    @property
    def n_local(self):
        return self.hi - self.lo

    def run(self, fn):
        return fn(self.lo, self.n_local)

    def reduce_stats(self, stats):
        """all-reduce: sums for stats[0:3], max for stats[3]; returns dict with nmse and max error."""
        if self.world_size > 1:
            import torch.distributed as dist
            sums = stats[:3].clone()
            mx = stats[3:4].clone()
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
            se, sr, cnt, mxv = (float(sums[0]), float(sums[1]), float(sums[2]), float(mx[0]))
        else:
            se, sr, cnt, mxv = (float(x) for x in stats)
        return {"sum_sq_err": se, "sum_sq_ref": sr, "count": cnt, "max_abs_err": mxv,
                "nmse": se / sr if sr > 0 else math.nan}
