"""`SparseMatrixCSC{Float64,Int}` as the assembler returns it: 1-based Int64 `colptr` / `rowval`, rows ascending and
unique per column (src/Algebra/SparseMatrixCSC.jl:264-283)."""
import numpy as np


class SparseMatrixCSC:
    def __init__(self, m, n, colptr, rowval, nzval):
        self.m, self.n = int(m), int(n)
        self.colptr, self.rowval, self.nzval = colptr, rowval, nzval

    @property
    def shape(self):
        return (self.m, self.n)

    def nnz(self):
        return len(self.nzval)

    def getindex(self, i, j):
        """A[i,j], 1-based (nz_index, src/Algebra/SparseMatrixCSC.jl:14-22)."""
        lo, hi = self.colptr[j - 1] - 1, self.colptr[j] - 1
        k = lo + np.searchsorted(self.rowval[lo:hi], i)
        return self.nzval[k] if k < hi and self.rowval[k] == i else 0.0

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.nzval, self.rowval - 1, self.colptr - 1), shape=(self.m, self.n))

    def toarray(self):
        return self.to_scipy().toarray()

    def findnz(self):
        J = np.repeat(np.arange(1, self.n + 1), np.diff(self.colptr))
        return self.rowval.copy(), J, self.nzval.copy()


class BlockMatrix:
    """`mortar(blocks)` of BlockArrays as `create_from_nz(::ArrayBlock)` returns it
    (src/MultiField/BlockSparseMatrixAssemblers.jl:222-227): blocks[i][j] is the SparseMatrixCSC of field block (i, j)."""

    def __init__(self, blocks):
        self.blocks = blocks

    def blocksize(self):
        return (len(self.blocks), len(self.blocks[0]))

    @property
    def shape(self):
        return (sum(r[0].m for r in self.blocks), sum(b.n for b in self.blocks[0]))

    def nnz(self):
        return sum(b.nnz() for r in self.blocks for b in r)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.bmat([[b.to_scipy() for b in r] for r in self.blocks], format="csc")

    def toarray(self):
        return self.to_scipy().toarray()


class BlockVector:
    """mortar of per-field vectors; `blocks[i]` are views of one contiguous array (`array`)."""

    def __init__(self, array, sizes):
        self.array = array
        ofs = np.concatenate([[0], np.cumsum(sizes)])
        self.blocks = [array[ofs[i]:ofs[i + 1]] for i in range(len(sizes))]

    def __array__(self, dtype=None, copy=None):
        return self.array if dtype is None else self.array.astype(dtype)

    def __len__(self):
        return len(self.array)


class SparseMatrixCSR:
    """`SparseMatrixCSR{Bi,Float64,Int}` of SparseMatricesCSR.jl as Gridap's CSR builder returns it
    (src/Algebra/SparseMatrixCSR.jl:31-75): `rowptr` / `colval` with index base Bi, columns ascending inside a row.
    `SparseMatrixCSR[Bi]` is the type to hand to `SparseMatrixAssembler(mat_type, vec_type, U, V)`."""

    Bi = 1

    def __class_getitem__(cls, bi):
        if bi not in (0, 1):
            raise ValueError("SparseMatrixCSR{Bi}: Bi must be 0 or 1")
        return type("SparseMatrixCSR_%d" % bi, (cls,), {"Bi": int(bi)})

    def __init__(self, m, n, rowptr, colval, nzval):
        self.m, self.n = int(m), int(n)
        self.rowptr, self.colval, self.nzval = rowptr, colval, nzval

    @property
    def shape(self):
        return (self.m, self.n)

    def nnz(self):
        return len(self.nzval)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.nzval, self.colval - self.Bi, self.rowptr - self.Bi), shape=(self.m, self.n))

    def toarray(self):
        return self.to_scipy().toarray()


class SymSparseMatrixCSR(SparseMatrixCSR):
    """`SymSparseMatrixCSR{Bi,Float64,Int}` (src/Algebra/SymSparseMatrixCSR.jl:1-50): the upper triangle (col >= row) of a symmetric
    matrix in CSR form; the builder drops the entries below the diagonal (`is_entry_stored(::Type{<:SymSparseMatrixCSR},i,j) = i<=j`).
    `SymSparseMatrixCSR[Bi]` is the type to hand to `SparseMatrixAssembler(mat_type, vec_type, U, V)`."""

    def __class_getitem__(cls, bi):
        if bi not in (0, 1):
            raise ValueError("SymSparseMatrixCSR{Bi}: Bi must be 0 or 1")
        return type("SymSparseMatrixCSR_%d" % bi, (cls,), {"Bi": int(bi)})

    def to_scipy(self):
        import scipy.sparse as sp
        U = sp.csr_matrix((self.nzval, self.colval - self.Bi, self.rowptr - self.Bi), shape=(self.m, self.n))
        return (U + sp.triu(U, 1).T).tocsr()
