"""Multi-GPU sharding of independent mode problems.

Each (plane, frequency) pair is an independent eigenproblem (the reference loops them serially with no carried state,
mode_solver.py:665-671), so the path shards with no data-path collective: work items are block-partitioned over ranks
(contiguous, so a rank's frequencies of one plane stay adjacent), every rank solves its shard on its own GPU through
the C ABI, and ONE exchange at the end gathers the resulting data -- ``n_complex`` with one ``all_gather`` of equal
(padded) blocks, the fields with one grouped send/recv batch (``batch_isend_irecv``) of exactly-sized blocks.  On GPUs
(NCCL over NVLink) the fields never visit the host on the sending ranks: the library writes them into a device buffer
(``b200ms_result.fields`` may be device memory) which is what NCCL sends; the destination rank copies the gathered
buffer to pinned host memory once.  gloo runs the same code on CPU tensors (tests).  One process per GPU (torchrun);
``torch.distributed`` is plumbing only.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence

import numpy as np


def partition(n_items: int, world: int, rank: int) -> range:
    """Contiguous block partition; the first ``n_items % world`` ranks take one extra item."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def _field_meta(p):
    """(shape, dtype, nbytes) of the packed fields of one problem (solver.py:257-267)."""
    nx, ny = len(p["coords"][0]) - 1, len(p["coords"][1]) - 1  # (a problem may carry a `section` instead of `eps_cross`)
    m = int(p["mode_spec"].num_modes)
    dt = np.complex64 if getattr(p["mode_spec"], "precision", "single") == "single" else np.complex128
    shape = (2, 3, nx, ny, 1, m)
    return shape, np.dtype(dt), int(np.prod(shape)) * np.dtype(dt).itemsize


def solve_sharded(
    problems: Sequence[dict],
    solve_fn: Optional[Callable[[Sequence[dict]], list]] = None,
    gather_fields: bool = False,
    dst: Optional[int] = None,
    info: Optional[dict] = None,
):
    """Solve ``problems`` across the ranks of the default process group.

    Returns ``(n_all, fields)``: ``n_all`` is the list of ``n_complex`` arrays of ALL problems in the original order, on
    every rank.  ``fields`` maps global problem index -> packed field array: with ``gather_fields=False`` only this
    rank's own problems; with ``gather_fields=True`` all problems on rank ``dst`` (``dst=None``: on every rank); other ranks keep
    their own problems' fields only when those were produced in host memory (on NCCL they stay in HBM and are sent).  ``solve_fn(problems) -> [(fields, n_complex, eps_spec), ...]`` replaces the device solve
    (tests); by default this rank's GPU (``LOCAL_RANK``) is used through ``tidy3d_b200.compute_modes_batch``.
    ``info`` (optional dict) receives timing / byte counters.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    nccl = dist.is_initialized() and dist.get_backend() == "nccl"
    local = int(os.environ.get("LOCAL_RANK", rank if nccl else 0))
    if nccl:
        torch.cuda.set_device(local)  # collectives' buffers must live on THIS rank's GPU
    dev = torch.device("cuda", local) if nccl else torch.device("cpu")
    n = len(problems)
    mine = partition(n, world, rank)
    metas = [_field_meta(p) for p in problems] if (gather_fields or nccl) else None
    m_modes = [int(p["mode_spec"].num_modes) for p in problems]

    # ---- local solve -------------------------------------------------------------------------------------------
    loc_dev = None  # flat byte tensor holding this rank's fields when they stay on the device
    offs = []
    if solve_fn is None:
        from .solver import compute_modes_batch

        if gather_fields and nccl and world > 1:
            # fields stay in HBM: the library writes straight into the buffer NCCL will send
            tot, offs = 0, []
            for i in mine:
                offs.append(tot)
                tot += (metas[i][2] + 255) & ~255
            loc_dev = torch.empty(max(tot, 1), dtype=torch.uint8, device=dev)
            ptrs = [loc_dev.data_ptr() + o for o in offs]
            local_out = compute_modes_batch([problems[i] for i in mine], device=local, fields_ptrs=ptrs) if len(mine) else []
        else:
            local_out = compute_modes_batch([problems[i] for i in mine], device=local) if len(mine) else []
    else:
        local_out = solve_fn([problems[i] for i in mine]) if len(mine) else []
    fields = {i: o[0] for i, o in zip(mine, local_out) if o[0] is not None}
    if world == 1:
        return [o[1] for o in local_out], fields

    # ---- n_complex: equal (padded) blocks, one all_gather ---------------------------------------------------------
    counts = [sum(m_modes[i] for i in partition(n, world, r)) for r in range(world)]
    buf = torch.zeros(2 * max(max(counts), 1), dtype=torch.float64, device=dev)
    if local_out:
        flat = np.concatenate([np.asarray(o[1], dtype=np.complex128) for o in local_out]).view(np.float64)
        buf[: flat.size] = torch.from_numpy(flat).to(dev)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    n_all: List[np.ndarray] = []
    for r in range(world):
        arr = gathered[r].cpu().numpy()[: 2 * counts[r]].view(np.complex128)
        off = 0
        for i in partition(n, world, r):
            n_all.append(arr[off : off + m_modes[i]].copy())
            off += m_modes[i]
    if info is not None:
        info["n_complex_gather_bytes"] = int(buf.numel() * 8 * world)
    if not gather_fields:
        return n_all, fields

    # ---- fields: exactly-sized blocks, one grouped send/recv batch -------------------------------------------------
    def block_bytes(r):
        return sum((metas[i][2] + 255) & ~255 for i in partition(n, world, r))

    if loc_dev is None:  # host fields (gloo tests, or a custom solve_fn): stage this rank's block
        tot, offs = 0, []
        for i in mine:
            offs.append(tot)
            tot += (metas[i][2] + 255) & ~255
        stage = np.zeros(max(tot, 1), dtype=np.uint8)
        for i, o in zip(mine, offs):
            a = np.ascontiguousarray(fields[i], dtype=metas[i][1]).view(np.uint8).ravel()
            stage[o : o + a.size] = a
        loc_dev = torch.from_numpy(stage).to(dev)
    receivers = list(range(world)) if dst is None else [dst]
    recv = {}
    ops = []
    if rank in receivers:
        for r in range(world):
            if r == rank or block_bytes(r) == 0:
                continue
            recv[r] = torch.empty(block_bytes(r), dtype=torch.uint8, device=dev)
            ops.append(dist.P2POp(dist.irecv, recv[r], r))
    for d in receivers:
        if d != rank and block_bytes(rank) > 0:
            ops.append(dist.P2POp(dist.isend, loc_dev[: block_bytes(rank)], d))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    moved = 0
    if rank in receivers:
        for r in range(world):
            src_t = loc_dev if r == rank else recv.get(r)
            if src_t is None:
                continue
            nb = block_bytes(r)
            if r != rank:
                moved += nb
            if src_t.is_cuda:
                host = torch.empty(nb, dtype=torch.uint8, pin_memory=True)  # torch's caching pinned allocator recycles these
                host[:nb].copy_(src_t[:nb], non_blocking=True)
                torch.cuda.current_stream().synchronize()
                arr = host[:nb].numpy()
            else:
                arr = src_t[:nb].numpy()
            off = 0
            for i in partition(n, world, r):
                shape, dt, nbytes = metas[i]
                fields[i] = arr[off : off + nbytes].view(dt).reshape(shape)
                off += (nbytes + 255) & ~255
    if info is not None:
        info["field_gather_bytes"] = int(moved)
        info["field_gather"] = "batch_isend_irecv (grouped send/recv), exactly-sized blocks" + (" over NCCL" if nccl else "")
    return n_all, fields
