"""Multi-GPU sharding of independent mode problems.

Each (plane, frequency) pair is an independent eigenproblem (the reference loops them serially with no carried state,
mode_solver.py:665-671), so the path shards with no data-path collective: work items are block-partitioned over ranks
(contiguous, so a rank's frequencies of one plane stay adjacent), every rank solves its shard on its own GPU through
the C ABI, and ONE collective at the end gathers ``n_complex`` (and optionally the fields) -- NCCL over NVLink on GPUs,
gloo in the CPU tests.  One process per GPU (torchrun); ``torch.distributed`` is plumbing only.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np


def partition(n_items: int, world: int, rank: int) -> range:
    """Contiguous block partition; the first ``n_items % world`` ranks take one extra item."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def solve_sharded(
    problems: Sequence[dict],
    solve_fn: Optional[Callable[[Sequence[dict]], list]] = None,
    gather_fields: bool = False,
    device: Optional[str] = None,
):
    """Solve ``problems`` across the ranks of the default process group.

    Returns on every rank the full list of ``n_complex`` arrays in the original order (and the list of fields when
    ``gather_fields``; otherwise fields stay on the rank that computed them and the second return value holds only the
    local ones keyed by global index).  ``solve_fn(problems) -> [(fields, n_complex, eps_spec), ...]`` defaults to
    ``tidy3d_b200.compute_modes_batch`` on this rank's GPU.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if solve_fn is None:
        from .solver import compute_modes_batch

        local = int(__import__("os").environ.get("LOCAL_RANK", rank))
        solve_fn = lambda ps: compute_modes_batch(ps, device=local)  # noqa: E731
    mine = partition(len(problems), world, rank)
    local_out = solve_fn([problems[i] for i in mine]) if len(mine) else []
    m_modes = [int(p["mode_spec"].num_modes) for p in problems]
    if world == 1:
        return [o[1] for o in local_out], {i: o[0] for i, o in zip(mine, local_out)}
    dev = device or ("cuda" if dist.get_backend() == "nccl" else "cpu")
    # n_complex: pad every rank's block to the same length, one all_gather
    counts = [sum(m_modes[i] for i in partition(len(problems), world, r)) for r in range(world)]
    buf = torch.zeros(2 * max(counts), dtype=torch.float64, device=dev)
    if local_out:
        flat = np.concatenate([np.asarray(o[1], dtype=np.complex128) for o in local_out]).view(np.float64)
        buf[: flat.size] = torch.from_numpy(flat).to(dev)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    n_all: List[np.ndarray] = []
    for r in range(world):
        arr = gathered[r].cpu().numpy()[: 2 * counts[r]].view(np.complex128)
        off = 0
        for i in partition(len(problems), world, r):
            n_all.append(arr[off : off + m_modes[i]].copy())
            off += m_modes[i]
    fields = {i: o[0] for i, o in zip(mine, local_out)}
    if gather_fields:
        sizes = [int(np.prod(problems[i]["eps_cross"][0].shape)) * 6 * m_modes[i] for i in range(len(problems))]
        for r in range(world):
            for i in partition(len(problems), world, r):
                t = torch.zeros(2 * sizes[i], dtype=torch.float64, device=dev)
                if r == rank:
                    t.copy_(torch.from_numpy(np.ascontiguousarray(fields[i], dtype=np.complex128).view(np.float64).ravel()).to(dev))
                dist.broadcast(t, src=r)
                if r != rank:
                    nx, ny = problems[i]["eps_cross"][0].shape
                    fields[i] = t.cpu().numpy().view(np.complex128).reshape(2, 3, nx, ny, 1, m_modes[i])
    return n_all, fields
