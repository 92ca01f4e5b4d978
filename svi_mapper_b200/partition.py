"""Frame-level data parallelism: stereo pairs are independent (addNewLandmarks keeps no state between
pairs and CTriangulator is const), so a batch is cut into contiguous frame ranges, one per GPU, with no
collective on the data path.  Only the timing uses a reduction (max over ranks)."""
from __future__ import annotations


def frame_range(n_frames: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous range [g*F/G, (g+1)*F/G) of rank g; ranges are disjoint and cover [0, F)."""
    if world < 1 or not (0 <= rank < world) or n_frames < 0:
        raise ValueError("bad partition arguments")
    return (rank * n_frames) // world, ((rank + 1) * n_frames) // world


def max_over_ranks(value: float, device=None) -> float:
    """Max of a scalar over all ranks of the default process group (identity without one)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
