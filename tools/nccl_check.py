"""Multi-GPU check of tidy3d_b200.sharding.solve_sharded on NCCL (run with torchrun --nproc-per-node N under gpurun --gpus N):
the sharded result -- n_complex on every rank, fields gathered on rank 0 through device buffers -- must equal the
single-GPU result of the same problems."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, "/root/repo")
from tidy3d_b200 import compute_modes_batch  # noqa: E402
from tidy3d_b200 import workloads as W  # noqa: E402
from tidy3d_b200.sharding import solve_sharded  # noqa: E402

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
wl = W.c2(nf=11, n=96)
probs = [dict(eps_cross=np.ascontiguousarray(np.stack(wl.eps_cross)), coords=wl.coords, freq=f, mode_spec=wl.mode_spec) for f in wl.freqs]
wl2 = W.si_strip(64, 2, W.sweep_freqs(5))
wl2.mode_spec.precision = "single"
probs += [dict(eps_cross=wl2.eps_cross, coords=wl2.coords, freq=f, mode_spec=wl2.mode_spec) for f in wl2.freqs]
info = {}
n_all, fields = solve_sharded(probs, gather_fields=True, dst=0, info=info)
ok = True
if rank == 0:
    ref = compute_modes_batch(probs, device=local)
    for i, (f, n, s) in enumerate(ref):
        ok &= bool(np.array_equal(n_all[i], n))
        ok &= bool(fields[i].dtype == f.dtype and np.array_equal(fields[i], f))
    print("nccl_check rank0: problems", len(probs), "world", world, "ok", ok, info, flush=True)
t = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
# n_complex must be identical on every rank
chk = torch.tensor(np.concatenate(n_all).view(np.float64), device="cuda")
ref_t = chk.clone()
dist.broadcast(ref_t, src=0)
ok2 = bool(torch.equal(chk, ref_t))
print(f"rank {rank}: n_complex identical to rank 0: {ok2}", flush=True)
dist.destroy_process_group()
sys.exit(0 if (t.item() == 1.0 and ok2) else 1)
