"""The Firedrake adapter cannot run here (no firedrake / petsc4py), but its pure index arithmetic
can: BAIJ `getValuesCSR()` -> block CSR + (nnzb, bs, bs) values, exercised through a fake Mat."""
import numpy as np
import pytest
import scipy.sparse as sp

from alfi_b200.firedrake_adapter import FiredrakeAdapter


class FakeBAIJ:
    def __init__(self, A_bsr):
        self.A = A_bsr

    def getBlockSize(self):
        return self.A.blocksize[0]

    def getValuesCSR(self):
        # PETSc returns the scalar CSR of a BAIJ matrix with every block fully stored
        rows = np.repeat(np.arange(self.A.shape[0] // self.A.blocksize[0]), np.diff(self.A.indptr))
        bs = self.A.blocksize[0]
        n = self.A.shape[0]
        indptr = [0]
        indices, data = [], []
        for i in range(n):
            bi, r = divmod(i, bs)
            for k in range(self.A.indptr[bi], self.A.indptr[bi + 1]):
                cj = self.A.indices[k]
                indices.extend(range(cj * bs, cj * bs + bs))
                data.extend(self.A.data[k][r])
            indptr.append(len(indices))
        return np.array(indptr, np.int32), np.array(indices, np.int32), np.array(data)


class FakePC:
    def __init__(self, mat):
        self.mat = mat

    def getOperators(self):
        return None, self.mat


@pytest.mark.parametrize("bs", [2, 3])
def test_baij_csr_to_block_csr(bs):
    rng = np.random.default_rng(bs)
    nb = 9
    pat = (sp.random(nb, nb, density=0.3, random_state=bs) + sp.identity(nb)).tocsr()
    pat.sort_indices()
    vals = rng.standard_normal((pat.indices.size, bs, bs))
    A = sp.bsr_matrix((vals, pat.indices, pat.indptr), shape=(nb * bs, nb * bs))
    rowptr, colidx, v, colmajor = FiredrakeAdapter().operator(FakePC(FakeBAIJ(A)))
    assert not colmajor
    assert np.array_equal(rowptr, pat.indptr) and np.array_equal(colidx, pat.indices)
    assert np.array_equal(v, vals)
