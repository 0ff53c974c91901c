"""Row sharding of independent streams across ranks (one process per GPU).

The path needs no data-path collective: every stream is an independent FIR chain with private carry state
(SURVEY.md §8e). torch.distributed is used only as plumbing: a barrier around the timed region and a
MAX/SUM reduction of per-rank timings/counts so that rank 0 can report whole-job throughput.
"""
from __future__ import annotations


def partition(total_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition; the first `total_rows % world` ranks take one extra row."""
    if world < 1 or not (0 <= rank < world) or total_rows < 0:
        raise ValueError("bad partition arguments")
    base, extra = divmod(total_rows, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def job_throughput(outputs_local: int, seconds_local: float, group=None, device=None) -> tuple[float, int, float]:
    """Whole-job (outputs/s, total outputs, max seconds): SUM of outputs over ranks / MAX of time over ranks."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return outputs_local / seconds_local, outputs_local, seconds_local
    t = torch.tensor([seconds_local], dtype=torch.float64, device=device)
    n = torch.tensor([outputs_local], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
    return float(n.item()) / float(t.item()), int(n.item()), float(t.item())
