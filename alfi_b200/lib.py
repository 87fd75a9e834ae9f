"""ctypes binding of libalfib.so — the thin C-ABI hand-over the north star asks for.

`Context` owns one `alfib_ctx` (one process <-> one GPU).  Vector arguments may be numpy arrays
(host; copied by the library inside the call) or torch CUDA tensors (device; passed as raw
``data_ptr()`` — PyTorch is used for buffer ownership only).  There is **no CPU fallback**: if
the shared library is missing or no GPU is visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

__all__ = ["load_library", "Context", "AlfibError", "EVENT_NAMES", "LIB_PATH"]

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libalfib.so")
_lib = None

# PETSc event names alfi reports (alfi/driver.py:80) in ALFIB_EV_* order
EVENT_NAMES = ["PCPATCHApply", "MatMult", "SchoeberlProlong", "SchoeberlRestrict",
               "KSPGMRESOrthog", "MGCoarseSolve", "PCSetUp_PATCH", "SFBcastReduce"]

PATCHES_SMOOTHER, PATCHES_TRANSFER = 0, 1
OPT_DETERMINISTIC, OPT_SYNC_ALWAYS, OPT_ROBUST_RESTRICT = 1, 2, 3

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.c_void_p          # host or device pointer to double

# every symbol include/alfib.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "alfib_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "alfib_destroy": (C.c_int, [C.c_void_p]),
    "alfib_last_error": (C.c_char_p, [C.c_void_p]),
    "alfib_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "alfib_set_deterministic": (C.c_int, [C.c_void_p, C.c_int]),
    "alfib_synchronize": (C.c_int, [C.c_void_p]),
    "alfib_launch_count": (C.c_int64, [C.c_void_p]),
    "alfib_stream": (C.c_void_p, [C.c_void_p]),
    "alfib_host_register": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "alfib_host_unregister": (C.c_int, [C.c_void_p, C.c_void_p]),
    "alfib_comm_unique_id": (C.c_int, [C.c_void_p]),
    "alfib_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "alfib_comm_peer_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "alfib_comm_peer_open": (C.c_int, [C.c_void_p, C.c_void_p]),
    "alfib_level_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "alfib_level_set_halo": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int32, C.c_int32, C.c_int32, _i32p, _i64p, _i32p,
                                       _i64p, _i32p, _i64p, _i64p]),
    "alfib_level_set_bsr_pattern": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, _i32p, _i32p]),
    "alfib_level_set_bsr_values": (C.c_int, [C.c_void_p, C.c_int, _f64p, C.c_int]),
    "alfib_level_set_bc": (C.c_int, [C.c_void_p, C.c_int, C.c_int32, _i32p]),
    "alfib_spmv": (C.c_int, [C.c_void_p, C.c_int, _f64p, _f64p]),
    "alfib_residual": (C.c_int, [C.c_void_p, C.c_int, _f64p, _f64p, _f64p]),
    "alfib_level_set_patches": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int32, _i64p, _i32p, C.c_int32,
                                          _i32p, _i32p]),
    "alfib_level_set_patch_blocks": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _i32p]),
    "alfib_level_set_sweep_stages": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int32, _i32p, C.c_int32, C.c_int]),
    "alfib_level_set_patch_corrections": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _i64p, _i32p, _i32p]),
    "alfib_level_set_patch_correction_values": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _f64p]),
    "alfib_patch_apply_bytes": (C.c_int64, [C.c_void_p, C.c_int, C.c_int]),
    "alfib_patch_storage_bytes": (C.c_int64, [C.c_void_p, C.c_int, C.c_int]),
    "alfib_patch_storage_form": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "alfib_patch_bind_storage": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64]),
    "alfib_level_factor": (C.c_int, [C.c_void_p, C.c_int]),
    "alfib_smoother_apply": (C.c_int, [C.c_void_p, C.c_int, _f64p, _f64p]),
    "alfib_get_colours": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _i32p]),
    "alfib_get_patch_inverse": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int32, _f64p]),
    "alfib_transfer_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int32, C.c_int32, _i32p, _i32p, _f64p, C.c_int32,
                                     _i32p, C.c_int]),
    "alfib_transfer_update": (C.c_int, [C.c_void_p, C.c_int, _f64p, _f64p, C.c_int]),
    "alfib_prolong": (C.c_int, [C.c_void_p, C.c_int, _f64p, _f64p]),
    "alfib_restrict": (C.c_int, [C.c_void_p, C.c_int, _f64p, _f64p]),
    "alfib_smooth": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _f64p, _f64p]),
    "alfib_coarse_factor": (C.c_int, [C.c_void_p]),
    "alfib_coarse_solve": (C.c_int, [C.c_void_p, _f64p, _f64p]),
    "alfib_cycle_setup": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "alfib_cycle_apply": (C.c_int, [C.c_void_p, _f64p, _f64p]),
    "alfib_schur_set": (C.c_int, [C.c_void_p, C.c_int32, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p, C.c_int]),
    "alfib_schur_apply": (C.c_int, [C.c_void_p, C.c_double, C.c_double, _f64p, _f64p]),
    "alfib_jacobian_apply": (C.c_int, [C.c_void_p, _f64p, _f64p]),
    "alfib_outer_solve": (C.c_int, [C.c_void_p, C.c_double, C.c_double, _f64p, _f64p, C.c_double, C.c_double, C.c_int32,
                                    C.c_int32, _i32p, C.POINTER(C.c_double), C.c_int32]),
    "alfib_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "alfib_profile_get": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), _i64p]),
    "alfib_profile_reset": (C.c_int, [C.c_void_p]),
}


class AlfibError(RuntimeError):
    pass


def load_library(path: str | None = None):
    """dlopen libalfib.so and bind every declared symbol.  Raises if the library is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise AlfibError("libalfib.so not found at %s — run `python -m alfi_b200.build` "
                         "(there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


class _Vec:
    """Resolve a numpy array / torch tensor to a raw pointer, keeping the owner alive."""

    def __init__(self, obj, n=None, writable=False):
        self.owner = obj
        if isinstance(obj, np.ndarray):
            if obj.dtype != np.float64 or not obj.flags.c_contiguous:
                if writable:
                    raise AlfibError("output arrays must be C-contiguous float64")
                obj = np.ascontiguousarray(obj, dtype=np.float64)
                self.owner = obj
            self.ptr = obj.ctypes.data
            size = obj.size
        elif hasattr(obj, "data_ptr"):         # torch tensor (device or pinned host)
            import torch
            if obj.dtype != torch.float64 or not obj.is_contiguous():
                raise AlfibError("tensors must be contiguous float64")
            self.ptr = obj.data_ptr()
            size = obj.numel()
        else:
            raise AlfibError("expected a numpy array or a torch tensor")
        if n is not None and size != n:
            raise AlfibError("vector has %d entries, expected %d" % (size, n))


class Context:
    """One GPU, one alfib_ctx.  Mirrors include/alfib.h one to one."""

    def __init__(self, device: int = 0, deterministic: bool = False):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.alfib_create(device, C.byref(h))
        if rc != 0 or not h:
            raise AlfibError("alfib_create(device=%d) failed with code %d — a CUDA device is required "
                             "(no CPU fallback)" % (device, rc))
        self.h = h
        self.device = device
        self._sizes = {}
        self._keep = []
        if deterministic:
            self.set_deterministic(True)

    # -- plumbing
    def _check(self, rc):
        if rc != 0:
            raise AlfibError("alfib error %d: %s" % (rc, self.lib.alfib_last_error(self.h).decode()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.alfib_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:       # noqa: BLE001
            pass

    def set_deterministic(self, flag=True):
        self._check(self.lib.alfib_set_deterministic(self.h, int(flag)))

    def set_option(self, key, value):
        self._check(self.lib.alfib_set_option(self.h, key, int(value)))

    def synchronize(self):
        self._check(self.lib.alfib_synchronize(self.h))

    @property
    def launches(self):
        return int(self.lib.alfib_launch_count(self.h))

    @property
    def stream(self):
        return int(self.lib.alfib_stream(self.h) or 0)

    def host_register(self, arr):
        """Page-lock a numpy array that will be handed over repeatedly (operator values, vectors); the array is kept
        alive by the context until `host_unregister` / `close`."""
        if not (isinstance(arr, np.ndarray) and arr.flags.c_contiguous):
            raise AlfibError("host_register takes a C-contiguous numpy array")
        self._check(self.lib.alfib_host_register(self.h, arr.ctypes.data, arr.nbytes))
        self._keep.append(arr)

    def host_unregister(self, arr):
        self._check(self.lib.alfib_host_unregister(self.h, arr.ctypes.data))

    @staticmethod
    def comm_unique_id() -> bytes:
        """128-byte ncclUniqueId (call on rank 0, broadcast to the others)."""
        buf = C.create_string_buffer(128)
        rc = load_library().alfib_comm_unique_id(buf)
        if rc != 0:
            raise AlfibError("alfib_comm_unique_id failed (%d)" % rc)
        return buf.raw

    def comm_init(self, unique_id: bytes | None, rank: int, nranks: int):
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        self._check(self.lib.alfib_comm_init(self.h, buf, rank, nranks))
        self.rank, self.nranks = rank, nranks

    def comm_peer_handle(self) -> bytes:
        """64-byte CUDA IPC handle of this rank's symmetric buffer (all-gather, then comm_peer_open)."""
        buf = C.create_string_buffer(64)
        self._check(self.lib.alfib_comm_peer_handle(self.h, buf))
        return buf.raw

    def comm_peer_open(self, handles: bytes):
        buf = C.create_string_buffer(handles, len(handles))
        self._check(self.lib.alfib_comm_peer_open(self.h, buf))

    def level_sizes(self):
        return dict(self._sizes)

    # -- level operator
    def level_create(self, level, n_nodes, bs):
        self._check(self.lib.alfib_level_create(self.h, level, n_nodes, bs))
        self._sizes[level] = n_nodes * bs

    @staticmethod
    def halo_peer_list(send: dict, recv: dict):
        """(peers ascending, send offsets, recv offsets) — the packed layout alfib_level_set_halo uses."""
        peers = sorted(set(send) | set(recv))
        s_off = np.zeros(len(peers) + 1, np.int64)
        r_off = np.zeros(len(peers) + 1, np.int64)
        for k, q in enumerate(peers):
            s_off[k + 1] = s_off[k] + np.asarray(send.get(q, ())).size
            r_off[k + 1] = r_off[k] + np.asarray(recv.get(q, ())).size
        return peers, s_off, r_off

    def set_halo(self, level, n_owned, n_local, send: dict, recv: dict, which=0, peer_offsets=None):
        """Distributed vectors on `level` (which = 0) or the transfer halo of `level` on level-1 (which = 1):
        `send[peer]` = owned local dofs the peer holds as ghosts, `recv[peer]` = the matching ghost local dofs
        (alfi_b200.halo.LocalLevel.send / .recv).  `peer_offsets` = {peer: (offset of this rank's segment in the
        peer's send list, in its recv list)} enables the NVLink peer-memory transport.  See include/alfib.h."""
        peers = sorted(set(send) | set(recv))
        empty = np.empty(0, np.int32)
        s_off = np.zeros(len(peers) + 1, np.int64)
        r_off = np.zeros(len(peers) + 1, np.int64)
        for k, q in enumerate(peers):
            s_off[k + 1] = s_off[k] + np.asarray(send.get(q, empty)).size
            r_off[k + 1] = r_off[k] + np.asarray(recv.get(q, empty)).size
        s_idx = _i32(np.concatenate([np.asarray(send.get(q, empty)).ravel() for q in peers])) if peers else empty
        r_idx = _i32(np.concatenate([np.asarray(recv.get(q, empty)).ravel() for q in peers])) if peers else empty
        pa = _i32(peers)
        pso = pro = None
        if peer_offsets is not None:
            pso = np.ascontiguousarray([peer_offsets[q][0] for q in peers], dtype=np.int64)
            pro = np.ascontiguousarray([peer_offsets[q][1] for q in peers], dtype=np.int64)
        self._check(self.lib.alfib_level_set_halo(self.h, level, which, int(n_owned), int(n_local), len(peers),
                                                  _ptr(pa, C.c_int32), _ptr(s_off, C.c_int64), _ptr(s_idx, C.c_int32),
                                                  _ptr(r_off, C.c_int64), _ptr(r_idx, C.c_int32),
                                                  None if pso is None else _ptr(pso, C.c_int64),
                                                  None if pro is None else _ptr(pro, C.c_int64)))

    def set_bsr_pattern(self, level, rowptr, colidx):
        rowptr, colidx = _i32(rowptr), _i32(colidx)
        self._check(self.lib.alfib_level_set_bsr_pattern(self.h, level, colidx.size, _ptr(rowptr, C.c_int32),
                                                         _ptr(colidx, C.c_int32)))

    def set_bsr_values(self, level, vals, block_col_major=False):
        v = _Vec(vals)
        self._check(self.lib.alfib_level_set_bsr_values(self.h, level, v.ptr, int(block_col_major)))

    def set_bc(self, level, bc_dofs):
        bc = _i32(bc_dofs)
        self._check(self.lib.alfib_level_set_bc(self.h, level, bc.size, _ptr(bc, C.c_int32)))

    def spmv(self, level, x, y):
        n = self._sizes[level]
        vx, vy = _Vec(x, n), _Vec(y, n, True)
        self._check(self.lib.alfib_spmv(self.h, level, vx.ptr, vy.ptr))
        return y

    def residual(self, level, b, x, r):
        n = self._sizes[level]
        vb, vx, vr = _Vec(b, n), _Vec(x, n), _Vec(r, n, True)
        self._check(self.lib.alfib_residual(self.h, level, vb.ptr, vx.ptr, vr.ptr))
        return r

    # -- patches
    def set_patches(self, level, offsets, dofs, order=None, colours=None, which=PATCHES_SMOOTHER):
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        dofs = _i32(dofs)
        npatch = offsets.size - 1
        order_a = _i32(order) if order is not None else None
        col_a = _i32(colours) if colours is not None else None
        self._check(self.lib.alfib_level_set_patches(
            self.h, level, which, npatch, _ptr(offsets, C.c_int64), _ptr(dofs, C.c_int32),
            0 if order_a is None else order_a.size,
            None if order_a is None else _ptr(order_a, C.c_int32),
            None if col_a is None else _ptr(col_a, C.c_int32)))

    def set_patch_blocks(self, level, blocks, which=PATCHES_SMOOTHER):
        """Block/separator structure of the patches set with set_patches (one int32 per patch dof:
        < 0 separator, else a block label); None returns to dense inverses.  See include/alfib.h."""
        if blocks is None:
            self._check(self.lib.alfib_level_set_patch_blocks(self.h, level, which, None))
            return
        b = _i32(blocks)
        self._check(self.lib.alfib_level_set_patch_blocks(self.h, level, which, _ptr(b, C.c_int32)))

    def set_sweep_stages(self, level, stage_of_visit, symmetric=False, which=PATCHES_SMOOTHER):
        """Multiplicative composition of the patch set (`alfi_b200.patches.sweep_stages`); None: additive."""
        if stage_of_visit is None:
            self._check(self.lib.alfib_level_set_sweep_stages(self.h, level, which, 0, None, 0, 0))
            return
        st = _i32(stage_of_visit)
        nstage = int(st.max()) + 1 if st.size else 0
        self._check(self.lib.alfib_level_set_sweep_stages(self.h, level, which, st.size, _ptr(st, C.c_int32), nstage,
                                                          int(bool(symmetric))))

    def set_patch_corrections(self, level, corr_off, rows, cols, which=PATCHES_SMOOTHER):
        """A_i = A[I_i, I_i] + C_i: COO pattern of the C_i in patch-local indices (Burman stabilisation; include/alfib.h).
        None removes them."""
        if corr_off is None:
            self._check(self.lib.alfib_level_set_patch_corrections(self.h, level, which, None, None, None))
            return
        off = np.ascontiguousarray(corr_off, dtype=np.int64)
        r, cidx = _i32(rows), _i32(cols)
        self._check(self.lib.alfib_level_set_patch_corrections(self.h, level, which, _ptr(off, C.c_int64),
                                                               _ptr(r, C.c_int32), _ptr(cidx, C.c_int32)))

    def set_patch_correction_values(self, level, vals, which=PATCHES_SMOOTHER):
        v = _Vec(np.ascontiguousarray(vals, dtype=np.float64) if isinstance(vals, np.ndarray) else vals)
        self._check(self.lib.alfib_level_set_patch_correction_values(self.h, level, which, v.ptr))

    def patch_apply_bytes(self, level, which=PATCHES_SMOOTHER):
        """Algorithmic bytes of one application of the patch set (factors + indices + 16 N)."""
        return int(self.lib.alfib_patch_apply_bytes(self.h, level, which))

    def patch_storage_form(self, level, which=PATCHES_SMOOTHER):
        """0 dense, 1 condensed per (patch, block), 2 condensed with shared blocks."""
        return int(self.lib.alfib_patch_storage_form(self.h, level, which))

    def patch_storage_bytes(self, level, which=PATCHES_SMOOTHER):
        return int(self.lib.alfib_patch_storage_bytes(self.h, level, which))

    def bind_patch_storage(self, level, tensor, which=PATCHES_SMOOTHER):
        """Hand a torch CUDA uint8/float64 tensor to the library as factor storage."""
        nbytes = tensor.numel() * tensor.element_size()
        self._keep.append(tensor)
        self._check(self.lib.alfib_patch_bind_storage(self.h, level, which, tensor.data_ptr(), nbytes))

    def factor(self, level):
        self._check(self.lib.alfib_level_factor(self.h, level))

    def smoother_apply(self, level, x, y):
        n = self._sizes[level]
        vx, vy = _Vec(x, n), _Vec(y, n, True)
        self._check(self.lib.alfib_smoother_apply(self.h, level, vx.ptr, vy.ptr))
        return y

    def colours(self, level, npatch, which=PATCHES_SMOOTHER):
        out = np.empty(npatch, dtype=np.int32)
        self._check(self.lib.alfib_get_colours(self.h, level, which, _ptr(out, C.c_int32)))
        return out

    def patch_inverse(self, level, patch, n, which=PATCHES_SMOOTHER):
        out = np.empty((n, n), dtype=np.float64)
        self._check(self.lib.alfib_get_patch_inverse(self.h, level, which, patch, out.ctypes.data))
        return out

    # -- transfer
    def set_transfer(self, level, P, cb_dofs, dof_level=False):
        """P: scipy CSR, scalar per node (fine nodes x coarse nodes) or, with dof_level, on scalar dofs."""
        P = P.tocsr()
        P.sort_indices()
        rp, ci = _i32(P.indptr), _i32(P.indices)
        pv = np.ascontiguousarray(P.data, dtype=np.float64)
        cb = _i32(cb_dofs)
        self._check(self.lib.alfib_transfer_set(self.h, level, P.shape[0], P.shape[1], _ptr(rp, C.c_int32),
                                                _ptr(ci, C.c_int32), pv.ctypes.data, cb.size, _ptr(cb, C.c_int32),
                                                int(dof_level)))

    def transfer_update(self, level, a0_vals, d_vals, block_col_major=False):
        a0 = _Vec(a0_vals) if a0_vals is not None else None
        dv = _Vec(d_vals) if d_vals is not None else None
        self._check(self.lib.alfib_transfer_update(self.h, level, None if a0 is None else a0.ptr,
                                                   None if dv is None else dv.ptr, int(block_col_major)))

    def prolong(self, level, coarse, fine):
        vc, vf = _Vec(coarse, self._sizes[level - 1]), _Vec(fine, self._sizes[level], True)
        self._check(self.lib.alfib_prolong(self.h, level, vc.ptr, vf.ptr))
        return fine

    def restrict(self, level, fine, coarse):
        vf, vc = _Vec(fine, self._sizes[level]), _Vec(coarse, self._sizes[level - 1], True)
        self._check(self.lib.alfib_restrict(self.h, level, vf.ptr, vc.ptr))
        return coarse

    # -- smoother / cycle
    def smooth(self, level, m, b, x):
        n = self._sizes[level]
        vb, vx = _Vec(b, n), _Vec(x, n, True)
        self._check(self.lib.alfib_smooth(self.h, level, m, vb.ptr, vx.ptr))
        return x

    def coarse_factor(self):
        self._check(self.lib.alfib_coarse_factor(self.h))

    def coarse_solve(self, b, x):
        vb, vx = _Vec(b, self._sizes[0]), _Vec(x, self._sizes[0], True)
        self._check(self.lib.alfib_coarse_solve(self.h, vb.ptr, vx.ptr))
        return x

    def cycle_setup(self, nlevels, smoothing):
        self._check(self.lib.alfib_cycle_setup(self.h, nlevels, smoothing))
        self._nlevels = nlevels

    def cycle_apply(self, b, x):
        n = self._sizes[self._nlevels - 1]
        vb, vx = _Vec(b, n), _Vec(x, n, True)
        self._check(self.lib.alfib_cycle_apply(self.h, vb.ptr, vx.ptr))
        return x

    # -- outer Schur-complement fieldsplit (alfi/solver.py:15-38, 405-421, 463-474)
    def schur_set(self, B, Minv, remove_constant=True):
        """B: scipy sparse (pressure dofs x finest-level velocity dofs), Dirichlet columns removed; Minv: the
        (block-diagonal) inverse pressure mass matrix."""
        B = B.tocsr()
        B.sort_indices()
        Minv = Minv.tocsr()
        Minv.sort_indices()
        n_p = B.shape[0]
        if B.shape[1] != self._sizes[self._nlevels - 1] or Minv.shape != (n_p, n_p):
            raise AlfibError("B / M_p^-1 shapes do not match the finest level")
        brp, bci, bv = _i32(B.indptr), _i32(B.indices), np.ascontiguousarray(B.data, dtype=np.float64)
        mrp, mci, mv = _i32(Minv.indptr), _i32(Minv.indices), np.ascontiguousarray(Minv.data, dtype=np.float64)
        self._check(self.lib.alfib_schur_set(self.h, n_p, _ptr(brp, C.c_int32), _ptr(bci, C.c_int32), bv.ctypes.data,
                                             _ptr(mrp, C.c_int32), _ptr(mci, C.c_int32), mv.ctypes.data,
                                             int(bool(remove_constant))))
        self._n_outer = B.shape[1] + n_p

    def schur_apply(self, nu, gamma, r, y):
        vr, vy = _Vec(r, self._n_outer), _Vec(y, self._n_outer, True)
        self._check(self.lib.alfib_schur_apply(self.h, float(nu), float(gamma), vr.ptr, vy.ptr))
        return y

    def jacobian_apply(self, z, out):
        vz, vo = _Vec(z, self._n_outer), _Vec(out, self._n_outer, True)
        self._check(self.lib.alfib_jacobian_apply(self.h, vz.ptr, vo.ptr))
        return out

    def outer_solve(self, nu, gamma, rhs, x, rtol, atol, maxit=500, restart=30):
        """Returns (x, iterations, residual history)."""
        vb, vx = _Vec(rhs, self._n_outer), _Vec(x, self._n_outer, True)
        its = C.c_int32(0)
        hist = (C.c_double * (maxit + 1))()
        self._check(self.lib.alfib_outer_solve(self.h, float(nu), float(gamma), vb.ptr, vx.ptr, float(rtol), float(atol),
                                               int(maxit), int(restart), C.byref(its), hist, maxit + 1))
        return x, int(its.value), [hist[i] for i in range(its.value + 1)]

    # -- instrumentation
    def profile(self, enable=True):
        self._check(self.lib.alfib_profile(self.h, int(enable)))

    def profile_reset(self):
        self._check(self.lib.alfib_profile_reset(self.h))

    def profile_get(self, level=-1):
        """{PETSc event name: (milliseconds, calls)} accumulated since the last reset."""
        ms = (C.c_double * len(EVENT_NAMES))()
        calls = (C.c_int64 * len(EVENT_NAMES))()
        self._check(self.lib.alfib_profile_get(self.h, level, ms, calls))
        return {name: (ms[i], int(calls[i])) for i, name in enumerate(EVENT_NAMES)}
