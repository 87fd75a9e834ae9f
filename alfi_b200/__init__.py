"""alfi_b200 — B200-native velocity-block multigrid for alfi's augmented-Lagrangian preconditioner.

Drop-in names (same as the reference's `alfi` package where they exist):
    Star, MacroStar                         patch constructors        (alfi/relaxation.py)
    CoarseCellPatches, CoarseCellMacroPatches, SVSchoeberlTransfer, PkP0SchoeberlTransfer,
    NullTransfer                            robust transfer           (alfi/transfer.py)
    PatchPC, VelocityMGPC, ALFieldsplitPC   petsc4py python PCs       (replace firedrake.PatchPC / fieldsplit_0 / the outer fieldsplit)
The compute path is the CUDA library `libalfib.so` (include/alfib.h); there is no CPU fallback.
"""
from .relaxation import MacroStar, OrderedRelaxation, Star, select_entity  # noqa: F401
from .transfer import (AutoSchoeberlTransfer, CoarseCellMacroPatches, CoarseCellPatches,  # noqa: F401
                       NullTransfer, PkP0SchoeberlTransfer, SVSchoeberlTransfer)
from .pc import ALFieldsplitPC, HostAdapter, PatchPC, VelocityMGPC  # noqa: F401

__version__ = "0.1.0"
