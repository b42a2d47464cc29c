"""`B200SparseMatrixAssembler`: the drop-in for Gridap's `SparseMatrixAssembler` on the B200 path.

Mirrors (same names / argument meaning / error behaviour) the reference interface
  src/FESpaces/Assemblers.jl:155-257        allocate_* / assemble_*! / assemble_*_add! / assemble_*
  src/FESpaces/Assemblers.jl:288-400        assemble_matrix(f,U,V), assemble_vector(f,V), assemble_matrix_and_vector
  src/FESpaces/Assemblers.jl:432-541        collect_cell_matrix / collect_cell_vector / collect_cell_matrix_and_vector
  src/FESpaces/AffineFEOperators.jl:23-54   AffineFEOperator
  src/FESpaces/FEOperatorsFromWeakForm.jl   FEOperator(res,jac,U,V): residual!, jacobian!
Every numeric method is a call into libgridap_b200.so (lib.DevicePlan); there is no CPU path.
"""
import numpy as np

from . import celldata as cd
from . import lib
from . import reffes as rf
from .algebra import BlockMatrix, BlockVector, SparseMatrixCSC, SparseMatrixCSR
from .geometry import BoundaryTriangulation
from .fespaces import BlockMultiFieldStyle, FEFunction, FESpace, MultiFieldFESpace, TrialFESpace


def _base(space):
    return space.space if isinstance(space, TrialFESpace) else space


def _fields(space):
    if isinstance(space, MultiFieldFESpace):
        return [s for s in space.spaces]
    return [space]


def get_fe_basis(V):
    """test basis (src/FESpaces/FESpaceInterface.jl:164-177)."""
    if isinstance(V, MultiFieldFESpace):
        return tuple(cd.Basis("test", _base(s), k) for k, s in enumerate(V.spaces))
    return cd.Basis("test", _base(V))


def get_trial_fe_basis(U):
    """trial basis = transpose of the test basis (src/FESpaces/FESpaceInterface.jl:186-191)."""
    if isinstance(U, MultiFieldFESpace):
        return tuple(cd.Basis("trial", _base(s), k) for k, s in enumerate(U.spaces))
    return cd.Basis("trial", _base(U))


class MatData:
    """What `collect_cell_matrix` returns: per-triangulation (cell matrices, rows, cols) -- here the cell matrices
    stay symbolic (recognised terms) or are one constant local matrix (the `Fill(K_e,ncells)` case)."""

    def __init__(self, terms, measure, const_Ke=None, extra=()):
        self.terms, self.measure, self.const_Ke = terms, measure, const_Ke
        self.extra = list(extra)   # the same for further triangulations of the form (a = int_Omega ... + int_Gamma ...)


class VecData:
    def __init__(self, terms, measure, extra=()):
        self.terms, self.measure = terms, measure
        self.extra = list(extra)


def _by_measure(contrib):
    """[(measure, Sum of its integrands)] in order of first appearance: one entry per triangulation / quadrature, like the
    per-triangulation lists of `collect_cell_matrix` (src/FESpaces/Assemblers.jl:432-448).  The bulk measure goes first: its
    plan owns the sparsity pattern, boundary contributions are merged into it."""
    groups = {}
    for e, m in contrib.terms:
        groups.setdefault(id(m), (m, []))[1].append(e)
    out = [(m, cd.Sum(es)) for m, es in groups.values()]
    out.sort(key=lambda t: isinstance(t[0].trian, BoundaryTriangulation))
    return out


def collect_cell_matrix(U, V, contrib):
    parts = [MatData(cd.recognise_matrix(e), m) for m, e in _by_measure(contrib)]
    if isinstance(parts[0].measure.trian, BoundaryTriangulation):
        raise NotImplementedError("a bilinear form with boundary terms only: on the B200 path the bulk term defines the sparsity pattern")
    parts[0].extra = parts[1:]
    return parts[0]


def collect_cell_vector(V, contrib):
    parts = [VecData(cd.recognise_vector(e), m) for m, e in _by_measure(contrib)]
    parts[0].extra = parts[1:]
    return parts[0]


def collect_cell_matrix_and_vector(U, V, mat_contrib, vec_contrib, uhd=None):
    m = collect_cell_matrix(U, V, mat_contrib)
    v = collect_cell_vector(V, vec_contrib)
    if m.measure.degree != v.measure.degree:
        raise NotImplementedError("matrix and vector terms paired in AffineFEOperator must share the quadrature degree")
    return (m, v, uhd)


def fill_cell_matrix(Ke, measure):
    """matdata whose cell-matrix array is Fill(K_e, ncells) (src/Arrays/LazyArrays.jl:302-322): scatter only."""
    return MatData([], measure, const_Ke=np.asarray(Ke, dtype=np.float64))


class B200SparseMatrixAssembler:
    """SparseMatrixAssembler(U, V): matrix type SparseMatrixCSC{Float64,Int}, vector type Vector{Float64},
    DefaultAssemblyStrategy (src/FESpaces/SparseMatrixAssemblers.jl:127-160)."""

    def __init__(self, U, V, ctx=None, deterministic=False, col_range=None):
        self.U, self.V = U, V
        self.ctx = ctx if ctx is not None else lib.default_context(None, deterministic)
        self.trial_fields = [_base(s) for s in _fields(U)]
        self.test_fields = [_base(s) for s in _fields(V)]
        if len(self.trial_fields) != len(self.test_fields):
            raise NotImplementedError("different numbers of trial and test fields")
        self.row_offsets = V.offsets if isinstance(V, MultiFieldFESpace) else [0]
        self.col_offsets = U.offsets if isinstance(U, MultiFieldFESpace) else [0]
        self.nrows = V.num_free_dofs()
        self.ncols = U.num_free_dofs()
        self.col_range = col_range  # (lo, hi) 0-based half-open range of owned columns (multi-GPU), None = all
        self._plans = {}

    # -- Assembler interface
    def get_rows(self):
        return range(1, self.nrows + 1)

    def get_cols(self):
        return range(1, self.ncols + 1)

    def num_rows(self):
        return self.nrows

    def num_cols(self):
        return self.ncols

    def get_assembly_strategy(self):
        return "DefaultAssemblyStrategy" if self.col_range is None else ("OwnedColumnsStrategy", self.col_range)

    def get_matrix_type(self):
        return SparseMatrixCSC

    def get_vector_type(self):
        return np.ndarray

    # -- plan management (symbolic phase, persistent on the device)
    def _touched(self, terms):
        nf = len(self.test_fields)
        if nf == 1:
            return None
        if any(t.form == lib.FORM_STOKES for t in terms):
            return np.array([[1, 1], [1, 0]], dtype=np.uint8)
        raise NotImplementedError("multi-field forms other than Stokes are not on the B200 path")

    def plan(self, measure, touched=None):
        trian = measure.trian
        on_boundary = isinstance(trian, BoundaryTriangulation)
        key = (measure.degree, None if touched is None else touched.tobytes(), id(trian) if on_boundary else None)
        if key in self._plans:
            return self._plans[key]
        test_fields, trial_fields = self.test_fields, self.trial_fields
        if on_boundary:   # facet-wise DoF tables of the same spaces (same global numbering)
            if self.col_range is not None:
                raise NotImplementedError("boundary terms with column ownership (multi-GPU)")
            test_fields = [trian.restrict(s) for s in test_fields]
            trial_fields = [trian.restrict(s) for s in trial_fields]
        model = test_fields[0].model
        mesh = model.device_mesh(self.ctx)
        xq, w = measure.points, measure.weights
        Ng, dNg = rf.tabulate_lagrangian(model.ptype, 1, xq)
        geo = lib.DeviceRefEl(self.ctx, w, Ng, dNg, 1)
        tests, trials = [], []
        ncols_local = self.ncols
        for k, (t, u) in enumerate(zip(test_fields, trial_fields)):
            if t.reffe.order != u.reffe.order or t.ncomp != u.ncomp:
                raise NotImplementedError("trial and test reference FEs must coincide on the B200 path")
            N, dN = rf.tabulate_lagrangian(model.ptype, t.reffe.order, xq)
            refel = lib.DeviceRefEl(self.ctx, w, N, dN, t.ncomp)
            ts = t.device_space(self.ctx, (measure.degree, "test"), refel)
            if self.col_range is not None:
                if len(self.test_fields) != 1:
                    raise NotImplementedError("column ownership with multi-field spaces")
                lo, hi = self.col_range
                cache = u.__dict__.setdefault("_owned_col_ids", {})   # the rank-local trial numbering, built once per space
                if (lo, hi) not in cache:
                    ids = u.get_cell_dof_ids().copy()
                    pos = ids > 0
                    owned = pos & (ids > lo) & (ids <= hi)
                    ids[pos & ~owned] = 0          # masked: neither free nor Dirichlet (AssemblyStrategy col_mask)
                    ids[owned] -= lo
                    cache[(lo, hi)] = ids
                ids = cache[(lo, hi)]
                self._masked_ids = ids
                us = lib.DeviceSpace(self.ctx, mesh, refel, ids, hi - lo, u.num_dirichlet_dofs())
                ncols_local = hi - lo
            elif u is t:
                us = ts
            else:
                us = u.device_space(self.ctx, (measure.degree, "trial"), refel)
            tests.append(ts)
            trials.append(us)
        p = lib.DevicePlan(self.ctx, mesh, geo, tests, trials, touched, self.row_offsets, self.col_offsets if self.col_range is None else [0],
                           self.nrows, ncols_local)
        self._plans[key] = p
        return p

    def _set_dirichlet(self, plan, state=None):
        for k, u in enumerate(_fields(self.U)):
            dv = getattr(u, "dirichlet_values", None)
            fv = None
            if state is not None:
                uh = state[k] if isinstance(state, (list, tuple)) else state
                fv, dv = uh.free_values, uh.dirichlet_values
            plan.set_state(k, fv, dv)

    def _fq(self, plan, term):
        if term.fq is None:
            return None, term.params
        xq = plan.quadrature_points()  # physical points from the device; f(x) evaluated on the host
        nc, np_, D = xq.shape
        vals = np.asarray(term.fq(xq.reshape(-1, D)), dtype=np.float64)
        ncomp = self.test_fields[0].ncomp
        vals = vals.reshape(nc, np_, ncomp) * term.params[0]
        return np.ascontiguousarray(vals), ()

    # -- allocate
    def allocate_matrix(self, matdata, zero=True, wait=True):
        plan = self.plan(matdata.measure, self._touched(matdata.terms))
        colptr, rowval = plan.pattern(wait)
        nzval = self.ctx.pinned_empty(plan.nnz, np.float64)  # page-locked: D2H of the values at full PCIe rate
        if zero:
            nzval[:] = 0.0
        return SparseMatrixCSC(self.nrows, plan.ncols, colptr, rowval, nzval)

    def allocate_vector(self, vecdata):
        return np.zeros(self.nrows)

    def allocate_matrix_and_vector(self, data, wait=True, zero=True):
        return self.allocate_matrix(data[0], zero=zero, wait=wait), self.allocate_vector(data[1])

    # -- numeric
    def _check(self, A, plan):
        if len(A.nzval) != plan.nnz or A.n != plan.ncols or A.m != self.nrows:
            raise ValueError("matrix was not allocated by this assembler for this form")

    def _assemble_extra_matrices(self, plan, matdata):
        """further triangulations of the form (boundary terms): assembled on their own plan, merged into the bulk plan's device matrix"""
        for e in matdata.extra:
            eplan = self.plan(e.measure, self._touched(e.terms))
            for j, t in enumerate(e.terms):
                if t.state is not None:
                    self._set_dirichlet(eplan, t.state)
                eplan.assemble_matrix(t.form, t.params, None, j > 0)
            if e.terms:
                plan.add_matrix_from(eplan)

    def assemble_matrix_add_(self, A, matdata, add=True):
        plan = self.plan(matdata.measure, self._touched(matdata.terms))
        self._check(A, plan)
        if matdata.extra:
            if add or matdata.const_Ke is not None:
                raise NotImplementedError("assemble_matrix_add! / Fill cell matrices for forms over several triangulations")
            for k, t in enumerate(matdata.terms):
                if t.state is not None:
                    self._set_dirichlet(plan, t.state)
                plan.assemble_matrix(t.form, t.params, None, k > 0)   # device-resident: the boundary terms are merged on the device
            self._assemble_extra_matrices(plan, matdata)
            plan.download_into(A.nzval, None)
            return A
        if matdata.const_Ke is not None:
            plan.assemble_matrix_const(matdata.const_Ke, A.nzval, add)
            return A
        if not matdata.terms and not add:
            A.nzval[:] = 0.0
        for k, t in enumerate(matdata.terms):
            if t.state is not None:
                self._set_dirichlet(plan, t.state)
            plan.assemble_matrix(t.form, t.params, A.nzval, add or k > 0)
        return A

    def assemble_matrix_(self, A, matdata):
        return self.assemble_matrix_add_(A, matdata, add=False)

    def assemble_vector_add_(self, b, vecdata, add=True):
        plan = self.plan(vecdata.measure, None if len(self.test_fields) == 1 else np.array([[1, 1], [1, 0]], dtype=np.uint8))
        if not vecdata.terms and not add:
            b[:] = 0.0
        for k, t in enumerate(vecdata.terms):
            if t.state is not None:
                self._set_dirichlet(plan, t.state)
            fq, params = self._fq(plan, t)
            plan.assemble_vector(t.form, params, fq, b, add or k > 0)
        for e in vecdata.extra:   # further triangulations (Neumann terms on a BoundaryTriangulation): accumulate into the same vector
            self.assemble_vector_add_(b, VecData(e.terms, e.measure), add=True)
        return b

    def assemble_vector_(self, b, vecdata):
        return self.assemble_vector_add_(b, vecdata, add=False)

    def assemble_matrix_and_vector_add_(self, A, b, data, add=True):
        matdata, vecdata, uhd = data
        plan = self.plan(matdata.measure, self._touched(matdata.terms))
        self._check(A, plan)
        self._set_dirichlet(plan, uhd)
        if len(matdata.terms) == 1 and len(vecdata.terms) == 1:
            fq, vparams = self._fq(plan, vecdata.terms[0])
            if not matdata.extra:
                plan.assemble_matrix_and_vector(matdata.terms[0].form, matdata.terms[0].params, vecdata.terms[0].form, vparams, fq, A.nzval, b, add)
            else:
                if add:
                    raise NotImplementedError("assemble_matrix_and_vector_add! for forms over several triangulations")
                plan.assemble_matrix_and_vector(matdata.terms[0].form, matdata.terms[0].params, vecdata.terms[0].form, vparams, fq, None, b, False)
                for e in matdata.extra:   # boundary matrix terms (Robin): K_Gamma merged on the device, lifting b -= K_Gamma u_D added to b
                    eplan = self.plan(e.measure, self._touched(e.terms))
                    if len(e.terms) != 1:
                        raise NotImplementedError("several boundary matrix terms on one triangulation in an AffineFEOperator")
                    self._set_dirichlet(eplan, uhd)
                    zero = (0.0,) * self.test_fields[0].ncomp
                    lift = np.zeros(self.nrows)   # add = False: the facet plan's device matrix is overwritten, lift = -K_Gamma u_D
                    eplan.assemble_matrix_and_vector(e.terms[0].form, e.terms[0].params, lib.FORM_SOURCE, zero, None, None, lift, False)
                    b += lift
                    plan.add_matrix_from(eplan)
                plan.download_into(A.nzval, None)
            for e in vecdata.extra:
                self.assemble_vector_add_(b, VecData(e.terms, e.measure), add=True)
            return A, b
        raise NotImplementedError("AffineFEOperator with several matrix / vector terms")

    def assemble_matrix_and_vector_(self, A, b, data):
        return self.assemble_matrix_and_vector_add_(A, b, data, add=False)

    def assemble_matrix(self, matdata):
        # the numeric phase overwrites every stored entry: no need to zero the freshly allocated values first
        # ... and the download of the pattern overlaps the numeric phase (completed by the numeric call's synchronisation)
        empty = not matdata.terms and matdata.const_Ke is None
        return self.assemble_matrix_(self.allocate_matrix(matdata, zero=empty, wait=empty), matdata)

    def assemble_vector(self, vecdata):
        return self.assemble_vector_(self.allocate_vector(vecdata), vecdata)

    def assemble_matrix_and_vector(self, data):
        A, b = self.allocate_matrix_and_vector(data, wait=False, zero=False)  # the numeric call overwrites and synchronises
        return self.assemble_matrix_and_vector_(A, b, data)


class B200BlockSparseMatrixAssembler(B200SparseMatrixAssembler):
    """BlockSparseMatrixAssembler (src/MultiField/BlockSparseMatrixAssemblers.jl:19-33): trial and test spaces with
    BlockMultiFieldStyle(); matrices come back as a BlockMatrix of SparseMatrixCSC (block-local ids), vectors as a BlockVector.
    The numeric phase is the same single device assembly; the blocks are extracted on the device
    (gb200_plan_get_block_pattern / gb200_plan_download_block)."""

    def __init__(self, U, V, **kw):
        super().__init__(U, V, **kw)
        self.row_sizes = [s.num_free_dofs() for s in self.test_fields]
        self.col_sizes = [s.num_free_dofs() for s in self.trial_fields]

    def get_rows(self):
        return [range(1, n + 1) for n in self.row_sizes]

    def get_cols(self):
        return [range(1, n + 1) for n in self.col_sizes]

    def get_matrix_type(self):
        return BlockMatrix

    def get_vector_type(self):
        return BlockVector

    def allocate_matrix(self, matdata, zero=True, wait=True):
        plan = self.plan(matdata.measure, self._touched(matdata.terms))
        blocks = []
        for i, m in enumerate(self.row_sizes):
            row = []
            for j, n in enumerate(self.col_sizes):
                colptr, rowval = plan.block_pattern(i, j, n)
                row.append(SparseMatrixCSC(m, n, colptr, rowval, np.zeros(len(rowval))))
            blocks.append(row)
        return BlockMatrix(blocks)

    def allocate_vector(self, vecdata):
        return BlockVector(np.zeros(self.nrows), self.row_sizes)

    def _check(self, A, plan):
        if not isinstance(A, BlockMatrix) or A.shape != (self.nrows, plan.ncols) or A.nnz() != plan.nnz:
            raise ValueError("matrix was not allocated by this assembler for this form")

    def _download_blocks(self, A, plan):
        for i, row in enumerate(A.blocks):
            for j, blk in enumerate(row):
                plan.download_block(i, j, blk.nzval)
        return A

    def assemble_matrix_add_(self, A, matdata, add=True):
        plan = self.plan(matdata.measure, self._touched(matdata.terms))
        self._check(A, plan)
        if add:
            raise NotImplementedError("assemble_matrix_add! on a BlockMatrix")
        for k, t in enumerate(matdata.terms):
            if t.state is not None:
                self._set_dirichlet(plan, t.state)
            plan.assemble_matrix(t.form, t.params, None, k > 0)   # device-resident; the blocks are downloaded below
        if matdata.terms:
            self._assemble_extra_matrices(plan, matdata)
        if not matdata.terms:
            for row in A.blocks:
                for blk in row:
                    blk.nzval[:] = 0.0
            return A
        return self._download_blocks(A, plan)

    def assemble_vector_add_(self, b, vecdata, add=True):
        super().assemble_vector_add_(b.array, vecdata, add)
        return b

    def assemble_matrix_and_vector_add_(self, A, b, data, add=True):
        raise NotImplementedError("AffineFEOperator with BlockMultiFieldStyle is not on the B200 path (Stokes has no source term here)")


class B200CSRSparseMatrixAssembler(B200SparseMatrixAssembler):
    """SparseMatrixAssembler(SparseMatrixCSR{Bi,Float64,Int}, Vector{Float64}, U, V) (src/FESpaces/SparseMatrixAssemblers.jl:127-153
    with the CSR builder of src/Algebra/SparseMatrixCSR.jl:31-75): same device assembly, results delivered in CSR order."""

    def __init__(self, U, V, mat_type, **kw):
        super().__init__(U, V, **kw)
        self.mat_type = mat_type

    def get_matrix_type(self):
        return self.mat_type

    def allocate_matrix(self, matdata, zero=True, wait=True):
        plan = self.plan(matdata.measure, self._touched(matdata.terms))
        rowptr, colval = plan.csr_pattern(self.mat_type.Bi)
        return self.mat_type(self.nrows, plan.ncols, rowptr, colval, np.zeros(plan.nnz))

    def _check(self, A, plan):
        if not isinstance(A, SparseMatrixCSR) or len(A.nzval) != plan.nnz or A.shape != (self.nrows, plan.ncols):
            raise ValueError("matrix was not allocated by this assembler for this form")

    def assemble_matrix_add_(self, A, matdata, add=True):
        plan = self.plan(matdata.measure, self._touched(matdata.terms))
        self._check(A, plan)
        if add:
            raise NotImplementedError("assemble_matrix_add! on a SparseMatrixCSR")
        if matdata.const_Ke is not None:
            plan.assemble_matrix_const(matdata.const_Ke, None, False)
        for k, t in enumerate(matdata.terms):
            if t.state is not None:
                self._set_dirichlet(plan, t.state)
            plan.assemble_matrix(t.form, t.params, None, k > 0)   # device-resident; values come back in CSR order below
        if matdata.terms:
            self._assemble_extra_matrices(plan, matdata)   # boundary terms merged on the device
        if not matdata.terms and matdata.const_Ke is None:
            A.nzval[:] = 0.0
            return A
        plan.download_csr(A.nzval)
        return A

    def assemble_matrix_and_vector_add_(self, A, b, data, add=True):
        matdata, vecdata, uhd = data
        plan = self.plan(matdata.measure, self._touched(matdata.terms))
        self._check(A, plan)
        if add or len(matdata.terms) != 1 or len(vecdata.terms) != 1 or matdata.extra:
            raise NotImplementedError("AffineFEOperator on a SparseMatrixCSR: one bulk matrix term and one bulk vector term, no _add!")
        self._set_dirichlet(plan, uhd)
        fq, vparams = self._fq(plan, vecdata.terms[0])
        plan.assemble_matrix_and_vector(matdata.terms[0].form, matdata.terms[0].params, vecdata.terms[0].form, vparams, fq, None, b, False)
        plan.download_csr(A.nzval)
        for e in vecdata.extra:   # Neumann terms
            self.assemble_vector_add_(b, VecData(e.terms, e.measure), add=True)
        return A, b


def SparseMatrixAssembler(*args, **kw):
    """SparseMatrixAssembler(U, V) | SparseMatrixAssembler(mat_type, vec_type, U, V) (src/FESpaces/SparseMatrixAssemblers.jl:127-160)."""
    if len(args) == 4:
        mat_type, vec_type, U, V = args
        if isinstance(mat_type, type) and issubclass(mat_type, SparseMatrixCSR):
            return B200CSRSparseMatrixAssembler(U, V, mat_type, **kw)
        if mat_type is not SparseMatrixCSC:
            raise NotImplementedError("matrix type %r: SparseMatrixCSC and SparseMatrixCSR{Bi} are on the B200 path" % (mat_type,))
    else:
        U, V = args
    bu = isinstance(getattr(U, "style", None), BlockMultiFieldStyle)
    bv = isinstance(getattr(V, "style", None), BlockMultiFieldStyle)
    if bu != bv:
        raise NotImplementedError("trial and test spaces must both have BlockMultiFieldStyle (BlockSparseMatrixAssemblers.jl:104-106)")
    if bu:
        return B200BlockSparseMatrixAssembler(U, V, **kw)
    return B200SparseMatrixAssembler(U, V, **kw)


# ---- function-taking sugar (src/FESpaces/Assemblers.jl:288-400)
def _split_args(args):
    if isinstance(args[0], B200SparseMatrixAssembler):
        return args[0], args[1:]
    return None, args


def assemble_matrix(f, *args):
    """assemble_matrix(f, U, V) | assemble_matrix(f, assem, U, V) | assemble_matrix(assem, matdata)."""
    if isinstance(f, B200SparseMatrixAssembler):
        return f.assemble_matrix(args[0])
    a, (U, V) = _split_args(args)
    a = a or SparseMatrixAssembler(U, V)
    return a.assemble_matrix(collect_cell_matrix(U, V, f(get_trial_fe_basis(U), get_fe_basis(V))))


def assemble_vector(f, *args):
    if isinstance(f, B200SparseMatrixAssembler):
        return f.assemble_vector(args[0])
    a, (V,) = _split_args(args)
    a = a or SparseMatrixAssembler(V, V)
    return a.assemble_vector(collect_cell_vector(V, f(get_fe_basis(V))))


def assemble_matrix_and_vector(f, b, *args):
    if isinstance(f, B200SparseMatrixAssembler):
        return f.assemble_matrix_and_vector(b)
    a, (U, V) = _split_args(args)
    a = a or SparseMatrixAssembler(U, V)
    uhd = FEFunction(U, np.zeros(U.num_free_dofs()))
    data = collect_cell_matrix_and_vector(U, V, f(get_trial_fe_basis(U), get_fe_basis(V)), b(get_fe_basis(V)), uhd)
    return a.assemble_matrix_and_vector(data)


class AffineFEOperator:
    """AffineFEOperator(a, l, U, V[, assem]) (src/FESpaces/AffineFEOperators.jl:23-54): assembles A and
    b = l(v) - a(u_D, v) in one fused pass with Dirichlet lifting."""

    def __init__(self, a, l, U, V, assem=None):
        self.trial, self.test = U, V
        self.assem = assem or SparseMatrixAssembler(U, V)
        uhd = FEFunction(U, np.zeros(U.num_free_dofs()))
        data = collect_cell_matrix_and_vector(U, V, a(get_trial_fe_basis(U), get_fe_basis(V)), l(get_fe_basis(V)), uhd)
        self.matrix, self.vector = self.assem.assemble_matrix_and_vector(data)

    def get_matrix(self):
        return self.matrix

    def get_vector(self):
        return self.vector


def get_matrix(op):
    return op.matrix


def get_vector(op):
    return op.vector


class FEOperator:
    """FEOperator(res, jac, U, V[, assem]) (src/FESpaces/FEOperatorsFromWeakForm.jl:24-27,50-103).
    A Jacobian must be given explicitly: the ForwardDiff path of the reference is outside the GPU path."""

    def __init__(self, res, jac, U, V, assem=None):
        if jac is None:
            raise NotImplementedError("FEOperator without an explicit Jacobian (automatic differentiation) is not on the B200 path")
        self.res, self.jac, self.trial, self.test = res, jac, U, V
        self.assem = assem or SparseMatrixAssembler(U, V)

    def allocate_residual(self, uh):
        return np.zeros(self.test.num_free_dofs())

    def residual_(self, b, uh):
        vecdata = collect_cell_vector(self.test, self.res(uh, get_fe_basis(self.test)))
        return self.assem.assemble_vector_(b, vecdata)

    def residual(self, uh):
        return self.residual_(self.allocate_residual(uh), uh)

    def _matdata(self, uh):
        return collect_cell_matrix(self.trial, self.test, self.jac(uh, get_trial_fe_basis(self.trial), get_fe_basis(self.test)))

    def allocate_jacobian(self, uh):
        return self.assem.allocate_matrix(self._matdata(uh))

    def jacobian_(self, A, uh):
        return self.assem.assemble_matrix_(A, self._matdata(uh))

    def jacobian(self, uh):
        return self.jacobian_(self.allocate_jacobian(uh), uh)

    def residual_and_jacobian(self, uh):
        """residual_and_jacobian! (src/FESpaces/FEOperatorsFromWeakForm.jl:85-103): one fused pass over the cells when both
        forms are single recognised terms evaluated at the same u_h, else two passes."""
        matdata = self._matdata(uh)
        vecdata = collect_cell_vector(self.test, self.res(uh, get_fe_basis(self.test)))
        if len(matdata.terms) == 1 and len(vecdata.terms) == 1 and matdata.measure.degree == vecdata.measure.degree \
                and matdata.terms[0].state is not None and vecdata.terms[0].fq is None:
            A = self.assem.allocate_matrix(matdata, zero=False, wait=False)
            b = self.allocate_residual(uh)
            self.assem.assemble_matrix_and_vector_(A, b, (matdata, vecdata, matdata.terms[0].state))
            return b, A
        return self.residual(uh), self.jacobian(uh)


def test_assembler(a, matdata, vecdata, data):
    """The conformance checker of the reference (`test_assembler` / `test_sparse_matrix_assembler`,
    src/FESpaces/Assemblers.jl:261-286, src/FESpaces/SparseMatrixAssemblers.jl:110-114): every entry point of the Assembler interface
    is called once; sizes must agree with num_rows / num_cols.  Beyond the reference's checks, `!` must overwrite and `_add!` must
    accumulate (the values after `assemble_matrix!` + `assemble_matrix_add!` are twice those of `assemble_matrix`)."""
    def close(x, y):
        return np.abs(np.asarray(x) - np.asarray(y)).max() <= 1e-12 * max(np.abs(np.asarray(y)).max(), 1e-300)
    A = a.allocate_matrix(matdata)
    assert a.num_cols() == A.shape[1] and a.num_rows() == A.shape[0]
    a.assemble_matrix_(A, matdata)
    a.assemble_matrix_add_(A, matdata)
    A1 = a.assemble_matrix(matdata)
    assert a.num_cols() == A1.shape[1] and a.num_rows() == A1.shape[0]
    assert np.array_equal(A.colptr, A1.colptr) and np.array_equal(A.rowval, A1.rowval) and close(A.nzval, 2.0 * A1.nzval)
    b = a.allocate_vector(vecdata)
    assert a.num_rows() == len(b)
    a.assemble_vector_(b, vecdata)
    a.assemble_vector_add_(b, vecdata)
    b1 = a.assemble_vector(vecdata)
    assert a.num_rows() == len(b1) and close(b, 2.0 * b1)
    A, b = a.allocate_matrix_and_vector(data)
    a.assemble_matrix_and_vector_(A, b, data)
    a.assemble_matrix_and_vector_add_(A, b, data)
    assert a.num_cols() == A.shape[1] and a.num_rows() == A.shape[0] and a.num_rows() == len(b)
    A2, b2 = a.assemble_matrix_and_vector(data)
    assert a.num_cols() == A2.shape[1] and a.num_rows() == A2.shape[0] and a.num_rows() == len(b2)
    assert close(A.nzval, 2.0 * A2.nzval) and close(b, 2.0 * b2)
    return True


test_sparse_matrix_assembler = test_assembler
test_assembler.__test__ = False   # (not a pytest test: a checker the tests call)
