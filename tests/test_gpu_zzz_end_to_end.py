"""GPU: ``ModeSolver.data_raw`` of the unmodified reference reproduced END TO END on the device (SURVEY 8 rows f-2 -> a -> f-1):
the plane description of the f-2 seam (per-site medium map + media tensors per frequency) rasterised on the device, the
eigenproblems of the sweep solved in one batch, gauge / finite-grid correction / flux normalisation / modal overlaps evaluated
in HBM, mode tracking on the host from the M x M overlap matrices -- against tests/golden/e2e_strip.npz, which the reference's
OWN code produced (oracle/ref_solver.py: its permittivity sampling, its compute_modes, its ModeSolverData arithmetic).
The comparison itself is tests/e2e_case.check; tests/test_end_to_end_cpu.py runs the same check with the device call emulated
by the restatements (measured there: n 1e-8, fields 2e-6, overlaps 1e-8 -- the reference's own ARPACK tolerance)."""
import pytest

from tests import e2e_case as E

pytestmark = pytest.mark.gpu


def test_device_chain_reproduces_the_reference_mode_solver_data():
    from tidy3d_b200 import compute_modes_batch
    from tidy3d_b200.solver import get_handle

    h = get_handle(tolerance="tight")

    def device(problems, post):
        return compute_modes_batch(problems, handle=h, post=post, return_info=True)

    worst = E.check(device)
    print("end-to-end vs the reference's ModeSolver.data_raw:", worst)
