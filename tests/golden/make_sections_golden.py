"""Generates tests/golden/sections_ref.npz: ``ModeSolver._solver_eps(freq)`` of the scene in
tests/section_cases.py for the three plane normals, computed by the UNMODIFIED reference's own sampling code
(oracle/ref_sections.py: ``Simulation.epsilon_on_grid``, ``Geometry.inside_meshgrid``, ``Box / Sphere / Cylinder.inside``,
``_tensorial_material_profile_modal_plane_tranform`` cut out of the reference at run time).

Run in the build container (needs /root/reference):   python tests/golden/make_sections_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import section_cases as SC  # noqa: E402


def reference_eps(normal):
    from oracle import ref_sections as RS

    structures = [(RS.geometry(kind, **kw), RS.TensorMedium(t, slope)) for kind, kw, t, slope in SC.STRUCTURES]
    ms = RS.solver(normal, SC.edges(normal), structures, RS.TensorMedium(*SC.BACKGROUND))
    return [np.array(ms._solver_eps(f)) for f in SC.FREQS]


def reference_masks():
    """``inside_meshgrid`` of the reference's Box / Sphere / Cylinder at the sites (sx, sy, z = 0) for the primitive cuts of
    tests/section_cases.PRIMITIVES."""
    from oracle import ref_sections as RS

    out = {}
    for name, (kind, kw, _) in SC.PRIMITIVES.items():
        out[f"mask_{name}"] = RS.geometry(kind, **kw).inside_meshgrid(SC.SX, SC.SY, np.array([0.0]))[:, :, 0]
    return out


def main():
    out = reference_masks()
    for normal in (0, 1, 2):
        for k, eps in enumerate(reference_eps(normal)):
            out[f"eps_n{normal}_f{k}"] = eps
    path = os.path.join(ROOT, "tests", "golden", "sections_ref.npz")
    np.savez_compressed(path, **out)
    print(path, f"{os.path.getsize(path) / 1024:.1f} KB", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
