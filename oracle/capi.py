"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the plain-C oracle (oracle/ref_assembly.c).

High-level entry points take numpy arrays in Gridap's conventions (1-based signed ids) and return
`SparseMatrixCSC`-like triples (colptr, rowval, nzval) with 1-based Int64 indices.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libref_assembly.so")

MASS, LAPLACIAN, ELASTICITY, STOKES, NEOHOOKEAN_JAC = 1, 2, 3, 4, 5
SOURCE, NEOHOOKEAN_RES = 10, 11
FACET, FACET_VEC = 20, 21


def build(force=False):
    src = os.path.join(_HERE, "ref_assembly.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB


class _Geom(C.Structure):
    _fields_ = [("D", C.c_int32), ("nn", C.c_int32), ("np", C.c_int32), ("ncells", C.c_int64), ("nnodes", C.c_int64),
                ("X", C.c_void_p), ("cell_nodes", C.c_void_p), ("w", C.c_void_p), ("Ng", C.c_void_p), ("dNg", C.c_void_p), ("Dr", C.c_int32),
                ("lface", C.c_void_p), ("nref", C.c_void_p)]


class _Field(C.Structure):
    _fields_ = [("nds", C.c_int32), ("ncomp", C.c_int32), ("N", C.c_void_p), ("dN", C.c_void_p), ("cell_dofs", C.c_void_p),
                ("free_values", C.c_void_p), ("dirichlet_values", C.c_void_p), ("offset", C.c_int64), ("fq", C.c_void_p), ("src", C.c_void_p)]


class _Problem(C.Structure):
    _fields_ = [("form_mat", C.c_int32), ("form_vec", C.c_int32), ("nfields", C.c_int32), ("fields", C.c_void_p),
                ("touched", C.c_void_p), ("params", C.c_void_p), ("fq", C.c_void_p), ("state_field", C.c_int32),
                ("lift_dirichlet", C.c_int32), ("nrows", C.c_int64), ("ncols", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_assemble.restype = C.c_int64
        _lib.orc_assemble_inplace.restype = C.c_int64
        _lib.orc_assemble_const.restype = C.c_int64
        _lib.orc_builder_from_counts.restype = C.c_void_p
        _lib.orc_builder_finish.restype = C.c_int64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Field:
    """One FE field: scalar Lagrangian tabulation N[p][a], dN[p][a][d], ncomp, signed cell dof ids."""

    def __init__(self, N, dN, ncomp, cell_dofs, offset=0, free_values=None, dirichlet_values=None, fq=None, src=None):
        self.N = _f64(N)
        self.dN = _f64(dN)
        self.ncomp = int(ncomp)
        self.cell_dofs = np.ascontiguousarray(cell_dofs, dtype=np.int32)
        self.offset = int(offset)
        self.free_values = _f64(free_values)
        self.dirichlet_values = _f64(dirichlet_values)
        self.fq, self.src = _f64(fq), _f64(src)   # per-field source of a multi-field linear form
        assert self.cell_dofs.shape[1] == self.N.shape[1] * self.ncomp


class Problem:
    def __init__(self, X, cell_nodes, w, Ng, dNg, fields, form_mat=0, form_vec=0, params=None, fq=None, touched=None,
                 state_field=0, lift_dirichlet=False, nrows=None, ncols=None, lface=None, nref=None):
        """lface / nref: facet-of-cell glue (the cells are the cells adjacent to boundary facets; lface 0-based local face per
        cell; the tabulations hold one block of quadrature points per local face; w has len = points per facet x faces)"""
        self.X = _f64(X)
        self.cell_nodes = np.ascontiguousarray(cell_nodes, dtype=np.int32)
        self.w = _f64(w)
        self.Ng = _f64(Ng)
        self.dNg = _f64(dNg)
        self.fields = fields
        self.params = _f64(params if params is not None else [0.0, 0.0, 0.0])
        self.fq = _f64(fq)
        self.touched = None if touched is None else np.ascontiguousarray(touched, dtype=np.uint8)
        D = self.X.shape[1]
        self.lface = None if lface is None else np.ascontiguousarray(lface, dtype=np.int32)
        self.nref = _f64(nref)
        npts = len(self.w) if lface is None else len(self.w) // len(self.nref)
        self.g = _Geom(D, self.cell_nodes.shape[1], npts, self.cell_nodes.shape[0], self.X.shape[0], _p(self.X),
                       _p(self.cell_nodes), _p(self.w), _p(self.Ng), _p(self.dNg), self.dNg.shape[2],   # Dr < D: boundary facets
                       _p(self.lface), _p(self.nref))
        self.farr = (_Field * len(fields))()
        for k, f in enumerate(fields):
            self.farr[k] = _Field(f.N.shape[1], f.ncomp, _p(f.N), _p(f.dN), _p(f.cell_dofs), _p(f.free_values),
                                  _p(f.dirichlet_values), f.offset, _p(f.fq), _p(f.src))
        if nrows is None:
            nrows = max(int(f.cell_dofs.max(initial=0)) for f in fields)
        if ncols is None:
            ncols = nrows
        self.nrows, self.ncols = int(nrows), int(ncols)
        self.pb = _Problem(form_mat, form_vec, len(fields), C.cast(self.farr, C.c_void_p), _p(self.touched), _p(self.params),
                           _p(self.fq), state_field, int(bool(lift_dirichlet)), self.nrows, self.ncols)

    # assemble_matrix / assemble_matrix_and_vector (from scratch)
    def assemble(self, with_vector=False):
        L = lib()
        colptr = np.zeros(self.ncols + 1, dtype=np.int64)
        rv = C.c_void_p()
        nz = C.c_void_p()
        b = np.zeros(self.nrows) if with_vector else None
        nnz = L.orc_assemble(C.byref(self.g), C.byref(self.pb), _p(colptr), C.byref(rv), C.byref(nz), _p(b))
        rowval = np.ctypeslib.as_array(C.cast(rv, C.POINTER(C.c_int64)), shape=(max(nnz, 1),))[:nnz].copy()
        nzval = np.ctypeslib.as_array(C.cast(nz, C.POINTER(C.c_double)), shape=(max(nnz, 1),))[:nnz].copy()
        L.orc_free(rv)
        L.orc_free(nz)
        return (colptr, rowval, nzval, b) if with_vector else (colptr, rowval, nzval)

    def assemble_vector(self, b=None, add=False):
        L = lib()
        if b is None:
            b = np.zeros(self.nrows)
        L.orc_assemble_inplace(C.byref(self.g), C.byref(self.pb), None, None, None, _p(b), int(add))
        return b

    def assemble_inplace(self, colptr, rowval, nzval, b=None, add=False):
        L = lib()
        missing = L.orc_assemble_inplace(C.byref(self.g), C.byref(self.pb), _p(colptr), _p(rowval), _p(nzval), _p(b), int(add))
        assert missing == 0, "entries outside the pattern"
        return nzval, b

    def symbolic_count(self):
        out = np.zeros(self.ncols, dtype=np.int64)
        lib().orc_symbolic_count(C.byref(self.g), C.byref(self.pb), _p(out))
        return out

    def cell_local(self, cell):
        nf = len(self.fields)
        nd = [f.cell_dofs.shape[1] for f in self.fields]
        K = [[np.zeros((nd[j], nd[i])) for j in range(nf)] for i in range(nf)]  # K[bi][bj] stored col-major -> shape (nj,ni) C-order
        b = [np.zeros(nd[i]) for i in range(nf)]
        Kp = (C.c_void_p * (nf * nf))(*[_p(K[i][j]) for i in range(nf) for j in range(nf)])
        bp = (C.c_void_p * nf)(*[_p(x) for x in b])
        lib().orc_cell_local(C.byref(self.g), C.byref(self.pb), C.c_int64(cell), Kp, bp)
        return [[K[i][j].T.copy() for j in range(nf)] for i in range(nf)], b

    def quadrature_only(self, nthreads=1):
        """CONTEXT ONLY (not the reference algorithm): per-cell quadrature without insertion on `nthreads` POSIX threads -> checksum"""
        L = lib()
        L.orc_quadrature_only.restype = C.c_double
        return L.orc_quadrature_only(C.byref(self.g), C.byref(self.pb), C.c_int32(int(nthreads)))

    def quadrature_points(self):
        xq = np.zeros((self.cell_nodes.shape[0], self.g.np, self.X.shape[1]))
        lib().orc_quadrature_points(C.byref(self.g), _p(xq))
        return xq


def assemble_const(cell_dofs, Ke, nrows, ncols):
    L = lib()
    cell_dofs = np.ascontiguousarray(cell_dofs, dtype=np.int32)
    KeF = np.asfortranarray(Ke, dtype=np.float64)
    colptr = np.zeros(ncols + 1, dtype=np.int64)
    rv, nz = C.c_void_p(), C.c_void_p()
    nnz = L.orc_assemble_const(C.c_int64(cell_dofs.shape[0]), C.c_int32(cell_dofs.shape[1]), _p(cell_dofs),
                               KeF.ctypes.data_as(C.c_void_p), C.c_int64(nrows), C.c_int64(ncols), _p(colptr), C.byref(rv), C.byref(nz))
    rowval = np.ctypeslib.as_array(C.cast(rv, C.POINTER(C.c_int64)), shape=(max(nnz, 1),))[:nnz].copy()
    nzval = np.ctypeslib.as_array(C.cast(nz, C.POINTER(C.c_double)), shape=(max(nnz, 1),))[:nnz].copy()
    L.orc_free(rv)
    L.orc_free(nz)
    return colptr, rowval, nzval


class Builder:
    """nz_counter -> nz_allocation -> add_entry! -> create_from_nz on raw (i,j,v) triples."""

    def __init__(self, nrows, ncols):
        self.nrows, self.ncols = nrows, ncols
        self.colnnzmax = np.zeros(ncols, dtype=np.int64)
        self.h = None

    def count(self, i, j):  # CounterCSC add_entry! (SparseMatrixCSC.jl:82-85); add_entries! skips ids <= 0
        if i > 0 and j > 0:
            self.colnnzmax[j - 1] += 1

    def allocate(self):
        self.h = C.c_void_p(lib().orc_builder_from_counts(C.c_int64(self.nrows), C.c_int64(self.ncols), _p(self.colnnzmax)))

    def add(self, v, i, j):
        lib().orc_builder_add(self.h, int(v is not None), C.c_double(0.0 if v is None else v), C.c_int64(i), C.c_int64(j))

    def state(self):
        colptr = np.zeros(self.ncols + 1, dtype=np.int64)
        colnnz = np.zeros(self.ncols, dtype=np.int64)
        lib().orc_builder_state(self.h, _p(colptr), _p(colnnz))
        return colptr, colnnz

    def finish(self):
        cap = int(self.colnnzmax.sum())
        colptr = np.zeros(self.ncols + 1, dtype=np.int64)
        rowval = np.zeros(cap + 1, dtype=np.int64)
        nzval = np.zeros(cap + 1)
        nnz = lib().orc_builder_finish(self.h, _p(colptr), _p(rowval), _p(nzval))
        return colptr, rowval[:nnz], nzval[:nnz]
