"""Region sharding across the GPUs of one box, the Python twin of bvhost::shard_region.

Every genomic position is computed independently (src/basetype_caller.cpp:586-611); the reference cuts a calling
interval into 100-kb tasks (src/basetype_caller.cpp:474-510).  A shard is a contiguous run of whole tasks; shard
sizes differ by at most one task; the merge is the concatenation in coordinate order.  There is no collective on the
data path: torch.distributed is used only for the barrier and for the max-over-ranks time of the benchmark.
"""
STEP_REGION_LEN = 100000  # src/basetype_caller.cpp:474


def shard_region(reg_beg, reg_end, n_shards, step=STEP_REGION_LEN):
    """[(beg, end), ...] half-open site ranges, at most n_shards of them, in coordinate order."""
    if reg_end <= reg_beg or n_shards <= 0:
        return []
    n_steps = (reg_end - reg_beg + step - 1) // step
    g = min(n_shards, n_steps)
    out, s0 = [], 0
    for i in range(g):
        cnt = n_steps // g + (1 if i < n_steps % g else 0)
        out.append((reg_beg + s0 * step, min(reg_end, reg_beg + (s0 + cnt) * step)))
        s0 += cnt
    return out


def rank_site_range(rank, world, sites_per_gpu):
    """Weak scaling of the benchmark: rank r owns sites [r*S, (r+1)*S) of the synthetic genome."""
    return rank * sites_per_gpu, (rank + 1) * sites_per_gpu


def max_over_ranks(value, device=None):
    """max of a Python float over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
