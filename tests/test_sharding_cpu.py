"""CPU (gloo, world_size 2): the multi-GPU sharding logic -- partitioning and the gather of n_complex / fields.
The per-rank solve is replaced by a deterministic stand-in (no GPU here); the GPU path is exercised by bench.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_solve(problems):
    out = []
    for p in problems:
        m = p["mode_spec"].num_modes
        nx, ny = p["eps_cross"][0].shape
        n = (p["freq"] * 1e-14 + np.arange(m)) * (1 + 0.5j)
        dt = np.complex64 if p["mode_spec"].precision == "single" else np.complex128
        f = (np.full((2, 3, nx, ny, 1, m), p["freq"] * 1e-14, dtype=np.complex128) * (1 + np.arange(m))).astype(dt)
        out.append((f, n, "diagonal"))
    return out


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from tidy3d_b200 import workloads as W
    from tidy3d_b200.sharding import partition, solve_sharded

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    wl = W.si_strip(8, 3, W.sweep_freqs(7))
    wl2 = W.si_strip(6, 2, W.sweep_freqs(4))
    wl2.mode_spec.precision = "single"  # complex64 fields: the gather must preserve the dtype
    probs = [dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=f, mode_spec=wl.mode_spec) for f in wl.freqs]
    probs += [dict(eps_cross=wl2.eps_cross, coords=wl2.coords, freq=f, mode_spec=wl2.mode_spec) for f in wl2.freqs]
    n_all, fields = solve_sharded(probs, solve_fn=_fake_solve, gather_fields=True)
    ref = _fake_solve(probs)
    ok = len(n_all) == len(probs)
    ok &= all(np.array_equal(a, r[1]) for a, r in zip(n_all, ref))
    ok &= all(np.array_equal(fields[i], ref[i][0]) and fields[i].dtype == ref[i][0].dtype for i in range(len(probs)))
    # gather to rank 0 only (what the bench does on NCCL): rank 0 holds everything, rank 1 its own shard
    info = {}
    n_all, fields = solve_sharded(probs, solve_fn=_fake_solve, gather_fields=True, dst=0, info=info)
    ok &= all(np.array_equal(a, r[1]) for a, r in zip(n_all, ref))
    if rank == 0:
        ok &= all(np.array_equal(fields[i], ref[i][0]) for i in range(len(probs)))
        ok &= info["field_gather_bytes"] > 0
    else:
        ok &= sorted(fields) == list(partition(len(probs), world, rank))
    ok &= list(partition(11, 2, 0)) == list(range(0, 6)) and list(partition(11, 2, 1)) == list(range(6, 11))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_partition_properties():
    sys.path.insert(0, ROOT)
    from tidy3d_b200.sharding import partition

    for n in (0, 1, 7, 256, 4096):
        for world in (1, 2, 3, 8):
            parts = [list(partition(n, world, r)) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_gloo_world2_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_field_meta_of_a_section_problem():
    """Problems described by a geometric cross-section have no eps_cross array: the field block size comes from coords."""
    import types

    from tidy3d_b200.sharding import _field_meta

    spec = types.SimpleNamespace(num_modes=3, precision="single")
    shape, dt, nbytes = _field_meta(dict(section=object(), coords=[np.zeros(11), np.zeros(8)], mode_spec=spec))
    assert shape == (2, 3, 10, 7, 1, 3) and dt == np.complex64 and nbytes == 2 * 3 * 10 * 7 * 3 * 8
    spec.precision = "double"
    assert _field_meta(dict(eps_cross=np.zeros((9, 10, 7)), coords=[np.zeros(11), np.zeros(8)], mode_spec=spec))[1] == np.complex128
