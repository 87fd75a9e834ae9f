"""The solver dictionaries the REFERENCE produces, as a fixture (run from the repo root in the build container:
python tests/golden/make_reference_parameters.py).

oracle/refshim.py loads alfi/solver.py from the reference tree over stand-ins and calls
`<Solver>.get_parameters()` (alfi/solver.py:305-510, configure_patch_solver :599-602 / :655-659) for the
BASELINE.json configurations; the nested dictionaries go to tests/golden/reference_parameters.json.  The
plugin tests feed them — unchanged — to alfi_b200.PatchPC / fieldsplit0_config.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402

# BASELINE.json configs -> arguments of get_parameters' owner
VARIANTS = {
    "ldc2d-sv-k2": dict(solver="ScottVogeliusSolver", tdim=2, patch="macro"),
    "ldc2d-pkp0": dict(solver="ConstantPressureSolver", tdim=2, patch="star"),
    "bfs2d-sv-k2": dict(solver="ScottVogeliusSolver", tdim=2, patch="macro"),
    "ldc3d-pkp0": dict(solver="ConstantPressureSolver", tdim=3, patch="star"),
    "ldc3d-sv-k3": dict(solver="ScottVogeliusSolver", tdim=3, patch="macro"),
    "ldc3d-sv-k3-multiplicative": dict(solver="ScottVogeliusSolver", tdim=3, patch="macro",
                                       patch_composition="multiplicative"),
    "ldc2d-pkp0-star-multiplicative": dict(solver="ConstantPressureSolver", tdim=2, patch="star",
                                           patch_composition="multiplicative"),
}


def run(name):
    outer, side, smoothing = refshim.reference_solver_parameters(**VARIANTS[name])
    return {"outer": outer, "firedrake_parameters": side, "smoothing": smoothing}


if __name__ == "__main__":
    blob = {name: run(name) for name in VARIANTS}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_parameters.json")
    with open(path, "w") as fh:
        json.dump(blob, fh, indent=1, sort_keys=True)
    print(path, os.path.getsize(path))
