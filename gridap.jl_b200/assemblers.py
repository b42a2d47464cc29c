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
from .algebra import BlockMatrix, BlockVector, SparseMatrixCSC, SparseMatrixCSR, SymSparseMatrixCSR
from .geometry import BoundaryTriangulation, DiscreteModel, SkeletonTriangulation, _FacetSpace
from .fespaces import BlockMultiFieldStyle, FEFunction, FESpace, FESpaceWithLinearConstraints, MultiFieldFESpace, TrialFESpace, has_constraints


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

    def __init__(self, terms, measure, const_Ke=None, extra=(), glued=False):
        self.terms, self.measure, self.const_Ke = terms, measure, const_Ke
        self.extra = list(extra)   # the same for further triangulations of the form (a = int_Omega ... + int_Gamma ...)
        self.glued = glued         # boundary terms with normals / cell-basis gradients: integrated on the cells adjacent to the facets


class VecData:
    def __init__(self, terms, measure, extra=(), glued=False):
        self.terms, self.measure = terms, measure
        self.extra = list(extra)
        self.glued = glued


def _to_facet_of_cell(t):
    """a plain boundary term (mass / source on the facet's own DoFs) as a facet-of-cell term: discontinuous spaces have no facet-wise
    DoF table (the trace of a cell basis on a facet is not shared with the neighbour), so every boundary term is integrated on the
    adjacent cell (FaceToCellGlue)"""
    if t.glued:
        return t
    if t.form == lib.FORM_MASS:
        return cd.Term(lib.FORM_FACET, (t.params[0], 0, 0), glued=True)
    if t.form == lib.FORM_SOURCE and t.fields is None:
        if t.fq is not None:
            c = float(t.params[0])
            return cd.Term(lib.FORM_FACET_VEC, (c, 0, 0), fq=t.fq, glued=True)
        gv = np.asarray(t.params, dtype=np.float64)
        return cd.Term(lib.FORM_FACET_VEC, (1.0, 0, 0), fq=(lambda x, gv=gv: np.broadcast_to(gv, (len(x), len(gv)))), glued=True)
    raise NotImplementedError("this boundary term on a discontinuous (L2) space")


def _split_glued(cls, terms, measure, space=None):
    """one part per (triangulation, kind of plan): the terms that need the adjacent cell of a boundary facet (FaceToCellGlue) are
    integrated on their own plan; the jump / mean terms of a SkeletonTriangulation on a skeleton plan (plus and minus cells)"""
    if isinstance(measure.trian, SkeletonTriangulation):
        if any(t.glued != "skeleton" for t in terms):
            raise NotImplementedError("on a SkeletonTriangulation only jump / mean terms are on the B200 path")
        return [cls(terms, measure, glued="skeleton")]
    if any(t.glued == "skeleton" for t in terms):
        raise NotImplementedError("jump / mean need a SkeletonTriangulation")
    if isinstance(measure.trian, BoundaryTriangulation) and space is not None and getattr(_base(_fields(space)[0]), "conformity", "H1") == "L2":
        terms = [_to_facet_of_cell(t) for t in terms]
    plain = [t for t in terms if not t.glued]
    glued = [t for t in terms if t.glued]
    if glued and not isinstance(measure.trian, BoundaryTriangulation):
        raise NotImplementedError("normal vectors / traces of FE functions inside a bulk integral")
    out = []
    if plain or not glued:
        out.append(cls(plain, measure))
    if glued:
        out.append(cls(glued, measure, glued=True))
    return out


def _by_measure(contrib):
    """[(measure, Sum of its integrands)] in order of first appearance: one entry per triangulation / quadrature, like the
    per-triangulation lists of `collect_cell_matrix` (src/FESpaces/Assemblers.jl:432-448).  The bulk measure goes first: its
    plan owns the sparsity pattern, boundary contributions are merged into it."""
    groups = {}
    for e, m in contrib.terms:
        groups.setdefault(id(m), (m, []))[1].append(e)
    out = [(m, cd.Sum(es)) for m, es in groups.values()]
    out.sort(key=lambda t: 2 if isinstance(t[0].trian, (BoundaryTriangulation, SkeletonTriangulation)) else
             (1 if getattr(t[0].trian, "cells", None) is not None else 0))      # bulk, views of the bulk, facets
    return out


def collect_cell_matrix(U, V, contrib):
    parts = [part for m, e in _by_measure(contrib) for part in _split_glued(MatData, cd.recognise_matrix(e), m, V)]
    if isinstance(parts[0].measure.trian, (BoundaryTriangulation, SkeletonTriangulation)):
        raise NotImplementedError("a bilinear form with boundary terms only: on the B200 path the bulk term defines the sparsity pattern")
    if getattr(parts[0].measure.trian, "cells", None) is not None and any(p.measure.trian is not parts[0].measure.trian for p in parts[1:]):
        raise NotImplementedError("a bilinear form over several triangulations none of which is the whole bulk: on the B200 path the first "
                                  "triangulation defines the sparsity pattern")
    parts[0].extra = parts[1:]
    return parts[0]


def collect_cell_vector(V, contrib):
    parts = [part for m, e in _by_measure(contrib) for part in _split_glued(VecData, cd.recognise_vector(e), m, V)]
    parts[0].extra = parts[1:]
    return parts[0]


def _same_domain(m1, m2):
    """two measures integrate over the same cells with the same quadrature: the pairing condition of `pair_arrays` in
    collect_cell_matrix_and_vector (src/FESpaces/Assemblers.jl:496-524 pairs contributions per triangulation)."""
    if m1 is m2:
        return True
    t1, t2 = m1.trian, m2.trian
    if isinstance(t1, BoundaryTriangulation) or isinstance(t2, BoundaryTriangulation):
        return t1 is t2 and m1.degree == m2.degree
    return t1.model is t2.model and m1.degree == m2.degree


def collect_cell_matrix_and_vector(U, V, mat_contrib, vec_contrib, uhd=None):
    """(matdata, vecdata, uhd): matrix and vector contributions on the same triangulation + quadrature are fused in one pass with
    the Dirichlet lifting; every other vector contribution (another quadrature degree, a BoundaryTriangulation) is assembled on
    its own plan into the same vector (the un-paired `vecdata` of the reference)."""
    m = collect_cell_matrix(U, V, mat_contrib)
    v = collect_cell_vector(V, vec_contrib)
    return (m, v, uhd)


def fill_cell_matrix(Ke, measure):
    """matdata whose cell-matrix array is Fill(K_e, ncells) (src/Arrays/LazyArrays.jl:302-322): scatter only."""
    return MatData([], measure, const_Ke=np.asarray(Ke, dtype=np.float64))


def skels_measure(matdata):
    return next(e.measure for e in matdata.extra if e.glued == "skeleton")


def plan_nl(assem):
    return assem.test_fields[0].get_cell_dof_ids().shape[1]


class DefaultAssemblyStrategy:
    """src/FESpaces/Assemblers.jl:120-132: identity maps, nothing masked."""


class GenericAssemblyStrategy:
    """GenericAssemblyStrategy(row_map, col_map, row_mask, col_mask) (src/FESpaces/Assemblers.jl:134-150).  The four callables
    act on numpy arrays of positive global ids (vectorised); an id is assembled at `map(id)` when `mask(id)` holds and skipped
    otherwise (`map_rows!` / `map_cols!`, :31-55).  On the wire a masked id is 0: neither free nor Dirichlet."""

    def __init__(self, row_map, col_map, row_mask, col_mask, ncols=None):
        self.row_map, self.col_map, self.row_mask, self.col_mask = row_map, col_map, row_mask, col_mask
        self.ncols = ncols     # number of columns of the assembled matrix when col_map renumbers into a smaller range
        self.identity_rows = False
        self._memo = {}

    def map_ids(self, ids, kind, offset=0):
        """ids (signed cell DoF table of one field) -> the table the device sees; memoised per table (a partition's strategy
        serves many assemblers: the mapping is host work that must not be repeated per assembly)"""
        key = (id(ids), kind, int(offset))
        hit = self._memo.get(key)
        if hit is not None and hit[0] is ids:
            return hit[1]
        out = self._map_ids(ids, kind, offset)
        if len(self._memo) > 16:
            self._memo.clear()
        self._memo[key] = (ids, out)
        return out

    def _map_ids(self, ids, kind, offset):
        fmap, fmask = (self.row_map, self.row_mask) if kind == "rows" else (self.col_map, self.col_mask)
        if offset:
            ids = ids.copy()
            ids[ids > 0] += offset
        if kind == "rows" and self.identity_rows:
            return ids
        out = ids.copy()
        pos = ids > 0
        gid = ids[pos].astype(np.int64)
        keep = np.asarray(fmask(gid), dtype=bool)
        mapped = np.zeros(len(gid), dtype=np.int64)
        mapped[keep] = np.asarray(fmap(gid[keep]), dtype=np.int64)
        out[pos] = mapped.astype(ids.dtype)
        return out


class OwnedColumns(GenericAssemblyStrategy):
    """Column ownership of the multi-GPU path: the rank assembles the columns `owned` (bool per global free trial DoF, or a
    half-open 0-based range (lo, hi)), renumbered 1..n_owned in ascending global order; rows keep their global ids."""

    def __init__(self, owned, nfree=None):
        if isinstance(owned, tuple):
            lo, hi = owned
            self.range = (int(lo), int(hi))
            super().__init__(lambda r: r, lambda c: c - lo, lambda r: np.ones(len(r), dtype=bool), lambda c: (c > lo) & (c <= hi), ncols=hi - lo)
            self.owned_ids = np.arange(lo + 1, hi + 1, dtype=np.int64)
            self.identity_rows = True
        else:
            owned = np.asarray(owned, dtype=bool)
            self.range = None
            local = np.cumsum(owned)
            super().__init__(lambda r: r, lambda c: local[c - 1], lambda r: np.ones(len(r), dtype=bool), lambda c: owned[c - 1], ncols=int(owned.sum()))
            self.owned_ids = np.nonzero(owned)[0] + 1
            self.identity_rows = True


class B200SparseMatrixAssembler:
    """SparseMatrixAssembler(U, V): matrix type SparseMatrixCSC{Float64,Int}, vector type Vector{Float64},
    DefaultAssemblyStrategy (src/FESpaces/SparseMatrixAssemblers.jl:127-160); an AssemblyStrategy may be given
    (GenericAssemblyStrategy, :145-153)."""

    def __init__(self, U, V, ctx=None, deterministic=False, col_range=None, strategy=None):
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
        if col_range is not None:
            if strategy is not None:
                raise ValueError("give either col_range or strategy")
            strategy = OwnedColumns(tuple(col_range))
        if isinstance(strategy, DefaultAssemblyStrategy):
            strategy = None
        if strategy is not None and not isinstance(strategy, GenericAssemblyStrategy):
            raise NotImplementedError("assembly strategy %r: DefaultAssemblyStrategy and GenericAssemblyStrategy are on the B200 path" % (strategy,))
        self.strategy = strategy
        self.col_range = getattr(strategy, "range", None)  # (lo, hi) 0-based half-open range of owned columns, when contiguous
        self.ncols_assembled = self.ncols if strategy is None or strategy.ncols is None else int(strategy.ncols)
        self._plans = {}
        self._mapped = {}

    # -- Assembler interface
    def get_rows(self):
        return range(1, self.nrows + 1)

    def get_cols(self):
        return range(1, self.ncols + 1)

    def num_rows(self):
        return self.nrows

    def num_cols(self):
        return self.ncols_assembled

    def get_assembly_strategy(self):
        return DefaultAssemblyStrategy() if self.strategy is None else self.strategy

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

    def _vector_touched(self):
        return None if len(self.test_fields) == 1 else np.array([[1, 1], [1, 0]], dtype=np.uint8)

    def _strategy_space(self, space, kind, offset, refel, mesh):
        """DeviceSpace whose ids went through the assembly strategy (global ids: the field offset is applied first, as
        get_cell_dof_ids of a MultiFieldFESpace does, src/MultiField/MultiFieldFESpaces.jl:460-488)."""
        key = (id(space), kind, id(refel))
        if key not in self._mapped:
            ids = self.strategy.map_ids(space.get_cell_dof_ids(), kind, offset)
            nfree = self.ncols_assembled if kind == "cols" else self.nrows
            self._mapped[key] = (lib.DeviceSpace(self.ctx, mesh, refel, ids, nfree, space.num_dirichlet_dofs()), space)
        return self._mapped[key][0]

    def plan(self, measure, touched=None, glued=False):
        trian = measure.trian
        if glued == "skeleton":
            return self._skeleton_plan(measure)
        on_boundary = isinstance(trian, BoundaryTriangulation)
        is_view = not on_boundary and getattr(trian, "cells", None) is not None      # Triangulation(model, cell_ids)
        key = (measure.degree, None if touched is None else touched.tobytes(), id(trian) if (on_boundary or is_view) else None, bool(glued))
        if key in self._plans:
            return self._plans[key][0]
        test_fields, trial_fields = self.test_fields, self.trial_fields
        space_model = test_fields[0].model
        if glued:         # FaceToCellGlue: the cells adjacent to the facets with their full cell DoF tables
            if not on_boundary:
                raise NotImplementedError("facet-of-cell terms need a BoundaryTriangulation")
            test_fields = [trian.glue_space(s) for s in test_fields]
            trial_fields = [trian.glue_space(s) for s in trial_fields]
        elif on_boundary:   # facet-wise DoF tables of the same spaces (same global numbering)
            test_fields = [trian.restrict(s) for s in test_fields]
            trial_fields = [trian.restrict(s) for s in trial_fields]
        elif is_view:       # the same spaces on a subset of the cells (same global numbering)
            if self.strategy is not None:
                raise NotImplementedError("view triangulations with a non-default AssemblyStrategy")
            test_fields = [trian.restrict(s) for s in test_fields]
            trial_fields = [trian.restrict(s) for s in trial_fields]
        elif trian.model is not space_model and trian.model is not getattr(space_model, "_partition_parent", None):
            raise ValueError("the Measure lives on another model than the FE spaces of this assembler")
        model = test_fields[0].model
        mesh = model.device_mesh(self.ctx)
        xq, w = measure.points, measure.weights
        if glued:   # the facet rule mapped onto every local face of the reference cell: one block of points per local face
            pts, wf, nref = rf.facet_glue(model.ptype, measure.degree)
            xq, w = pts.reshape(-1, pts.shape[2]), np.tile(wf, pts.shape[0])
        Ng, dNg = rf.tabulate_lagrangian(model.ptype, 1, xq)
        geo = lib.DeviceRefEl(self.ctx, w, Ng, dNg, 1)
        tests, trials, full_trials = [], [], []
        for k, (t, u) in enumerate(zip(test_fields, trial_fields)):
            if t.reffe.order != u.reffe.order or t.ncomp != u.ncomp:
                raise NotImplementedError("trial and test reference FEs must coincide on the B200 path")
            N, dN = rf.tabulate_lagrangian(model.ptype, t.reffe.order, xq)
            refel = lib.DeviceRefEl(self.ctx, w, N, dN, t.ncomp)
            if self.strategy is None:
                ts = t.device_space(self.ctx, (measure.degree, "test"), refel)
                us = ts if u is t else u.device_space(self.ctx, (measure.degree, "trial"), refel)
                full_trials.append(None)
            else:
                ts = self._strategy_space(t, "rows", self.row_offsets[k], refel, mesh)
                us = self._strategy_space(u, "cols", self.col_offsets[k], refel, mesh)
                full_trials.append((u, refel))
            tests.append(ts)
            trials.append(us)
        zero = [0] * len(tests)
        p = lib.DevicePlan(self.ctx, mesh, geo, tests, trials, touched, self.row_offsets if self.strategy is None else zero,
                           self.col_offsets if self.strategy is None else zero, self.nrows, self.ncols_assembled)
        if glued:
            p.set_facets(trian.lfaces + 1, nref)
        p._full_trials = full_trials
        p._has_state_space = False
        self._plans[key] = (p, trian)    # the triangulation stays alive with its plan: id(trian) cannot be recycled
        return p

    def _pair_plan(self, key, measure, plus_model, minus_model, ids_plus, ids_minus, skeleton=None):
        """two-field plan over pairs of cells (field 0 = the space on the first cell of every pair, field 1 = on the second, both in the
        same global numbering, all four blocks touched): its pattern holds the cross couplings of the pairs"""
        if key in self._plans:
            return self._plans[key][0]
        if len(self.test_fields) != 1 or self.strategy is not None:
            raise NotImplementedError("SkeletonTriangulation terms: single-field spaces with the default AssemblyStrategy")
        t, u = self.test_fields[0], self.trial_fields[0]
        if t.reffe.order != u.reffe.order or t.ncomp != u.ncomp:
            raise NotImplementedError("trial and test reference FEs must coincide on the B200 path")
        ptype = plus_model.ptype
        pts, wf, nref = rf.facet_glue(ptype, measure.degree)
        xq, w = pts.reshape(-1, pts.shape[2]), np.tile(wf, pts.shape[0])
        Ng, dNg = rf.tabulate_lagrangian(ptype, 1, xq)
        geo = lib.DeviceRefEl(self.ctx, w, Ng, dNg, 1)
        N, dN = rf.tabulate_lagrangian(ptype, t.reffe.order, xq)
        refel = lib.DeviceRefEl(self.ctx, w, N, dN, t.ncomp)
        tests, trials = [], []
        for model, (ti, ui) in ((plus_model, ids_plus), (minus_model, ids_minus)):
            mesh = model.device_mesh(self.ctx)
            ts = lib.DeviceSpace(self.ctx, mesh, refel, ti, t.num_free_dofs(), t.num_dirichlet_dofs())
            us = ts if ui is ti else lib.DeviceSpace(self.ctx, mesh, refel, ui, u.num_free_dofs(), u.num_dirichlet_dofs())
            tests.append(ts)
            trials.append(us)
        p = lib.DevicePlan(self.ctx, plus_model.device_mesh(self.ctx), geo, tests, trials, np.ones((2, 2), dtype=np.uint8), [0, 0], [0, 0],
                           self.nrows, self.ncols_assembled)
        if skeleton is not None:
            p.set_skeleton(skeleton.lfaces_plus + 1, skeleton.lfaces_minus + 1, skeleton.point_permutation(pts), nref)
        p._full_trials = [None, None]
        p._has_state_space = False
        self._plans[key] = (p, (skeleton, plus_model, minus_model))
        return p

    def _skeleton_plan(self, measure):
        """the plan of the jump / mean terms on the interior facets: plus and minus cells with their full cell DoF tables"""
        trian = measure.trian
        t, u = self.test_fields[0], self.trial_fields[0]
        if t.model is not trian.parent:
            raise ValueError("the FE space lives on another model than the SkeletonTriangulation")
        tp, tm = t.get_cell_dof_ids()[trian.cells_plus], t.get_cell_dof_ids()[trian.cells_minus]
        if u is t:
            up, um = tp, tm
        else:
            up, um = u.get_cell_dof_ids()[trian.cells_plus], u.get_cell_dof_ids()[trian.cells_minus]
        return self._pair_plan(("skeleton", measure.degree, id(trian)), measure, trian.glue_model("plus"), trian.glue_model("minus"), (tp, up), (tm, um), trian)

    def _union_plan(self, matdata):
        """the pattern of a form with skeleton terms = couplings inside every cell + couplings across every interior facet (the
        symbolic loop of the reference runs over all contributions, src/FESpaces/SparseMatrixAssemblers.jl:174-210): a pair plan
        over (cell, cell) for every cell and (plus, minus) for every interior facet.  It only carries the pattern and the
        accumulated values (gb200_plan_add_matrix_from); no integrand is evaluated on it."""
        skels = [e.measure.trian for e in matdata.extra if e.glued == "skeleton"]
        key = ("union", matdata.measure.degree) + tuple(id(tr) for tr in skels)
        if key in self._plans:
            return self._plans[key][0]
        t, u = self.test_fields[0], self.trial_fields[0]
        m = t.model
        allc = np.arange(m.num_cells(), dtype=np.int64)
        first = np.concatenate([allc] + [tr.cells_plus for tr in skels])
        second = np.concatenate([allc] + [tr.cells_minus for tr in skels])
        pm = DiscreteModel(m.node_coordinates, m.cell_node_ids[first], m.ptype)
        mm = DiscreteModel(m.node_coordinates, m.cell_node_ids[second], m.ptype)
        ti, ui = t.get_cell_dof_ids(), u.get_cell_dof_ids()
        tp, tm = ti[first], ti[second]
        up, um = (tp, tm) if u is t else (ui[first], ui[second])
        plan = self._pair_plan(key, skels_measure(matdata), pm, mm, (tp, up), (tm, um), None)
        self._plans[key] = (plan, (skels, pm, mm))
        return plan

    def _pattern_plan(self, matdata):
        """the plan that owns the sparsity pattern of the assembled matrix: the bulk plan, or the union plan of a form with skeleton terms"""
        if any(e.glued == "skeleton" for e in matdata.extra):
            return self._union_plan(matdata)
        return self.plan(matdata.measure, self._touched(matdata.terms))

    def _finish_matrix(self, plan, matdata, uhd=None):
        """skeleton parts: assembled on their own plan; the result is accumulated in the union plan (bulk + boundary + skeleton)"""
        skel = [e for e in matdata.extra if e.glued == "skeleton"]
        if not skel:
            return plan
        dv = getattr(uhd, "dirichlet_values", None)
        if dv is not None and np.any(np.asarray(dv) != 0.0):
            raise NotImplementedError("Dirichlet lifting through skeleton terms")
        union = self._union_plan(matdata)
        union.assemble_matrix_const(np.zeros((2 * plan_nl(self), 2 * plan_nl(self))), None, False)   # zero-fill
        union.add_matrix_from(plan)
        for e in skel:
            eplan = self._skeleton_plan(e.measure)
            for j, t in enumerate(e.terms):
                eplan.assemble_matrix(t.form, t.params, None, j > 0)
            if e.terms:
                union.add_matrix_from(eplan)
        return union

    def _set_dirichlet(self, plan, state=None):
        if state is not None and self.strategy is not None and not plan._has_state_space:
            # u_h lives on the global trial space: gather it through the unmasked ids (the plan's trial ids are mapped / masked)
            for k, ft in enumerate(plan._full_trials):
                u, refel = ft
                plan.set_state_space(k, u.device_space(self.ctx, ("state", id(refel)), refel))
            plan._has_state_space = True
        for k, u in enumerate(_fields(self.U)):
            dv = getattr(u, "dirichlet_values", None)
            fv = None
            if state is not None:
                uh = state[k] if isinstance(state, (list, tuple)) else state
                fv, dv = uh.free_values, uh.dirichlet_values
            plan.set_state(k, fv, dv)

    def _fq(self, plan, term):
        """(fq array or None, constant parameters) of a source term; multi-field: one source per field, field after field"""
        nf = len(self.test_fields)
        srcs = term.fields if term.fields is not None else {0: (term.params, term.fq)}
        if nf == 1 and term.fq is None:
            return None, term.params
        ncomps = [t.ncomp for t in self.test_fields]
        if all(fq is None for _, fq in srcs.values()):
            params = []
            for k in range(nf):
                pk = srcs.get(k, ((0.0,) * ncomps[k], None))[0]
                params += list(pk)
            return None, tuple(params)
        xq = plan.quadrature_points()  # physical points from the device; f(x) evaluated on the host
        nc, np_, D = xq.shape
        blocks = []
        for k in range(nf):
            pk, fq = srcs.get(k, ((0.0,) * ncomps[k], None))
            if fq is None:
                vals = np.broadcast_to(np.asarray(pk, dtype=np.float64), (nc, np_, ncomps[k]))
            else:
                vals = np.asarray(fq(xq.reshape(-1, D)), dtype=np.float64).reshape(nc, np_, ncomps[k]) * pk[0]
            blocks.append(np.ascontiguousarray(vals).ravel())
        return np.concatenate(blocks), ()

    # -- allocate
    def allocate_matrix(self, matdata, zero=True, wait=True):
        plan = self._pattern_plan(matdata)
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

    def _vec(self, b):
        return b

    def _fetch_matrix(self, A, plan, add):
        """the plan's device matrix -> A (layout of the matrix type); add: `_add!` semantics, A += device matrix"""
        if not plan.nnz:
            return A
        if add:
            tmp = self.ctx.pinned_empty(plan.nnz, np.float64)
            plan.download_into(tmp, None)
            A.nzval += tmp
        else:
            plan.download_into(A.nzval, None)
        return A

    def _assemble_extra_matrices(self, plan, matdata, uhd=None, lift_into=None):
        """further triangulations of the form (boundary terms): assembled on their own plan, merged into the bulk plan's device
        matrix; with `lift_into` the Dirichlet lifting -K_Gamma u_D of an AffineFEOperator is added to that vector"""
        for e in matdata.extra:
            if not e.terms or e.glued == "skeleton":   # (skeleton parts: _finish_matrix)
                continue
            eplan = self.plan(e.measure, self._touched(e.terms), e.glued)
            for j, t in enumerate(e.terms):
                if t.state is not None:
                    self._set_dirichlet(eplan, t.state)
                elif lift_into is not None and j == 0:
                    self._set_dirichlet(eplan, uhd)
                if lift_into is not None:   # the matrix term and its share of the lifting (zero local vector), device-resident
                    if e.glued:
                        eplan.assemble_matrix_and_vector(t.form, t.params, lib.FORM_FACET_VEC, (0.0, 0, 0), None, None, None, j > 0)
                    else:
                        zero = (0.0,) * sum(f.ncomp for f in self.test_fields)
                        eplan.assemble_matrix_and_vector(t.form, t.params, lib.FORM_SOURCE, zero, None, None, None, j > 0)
                else:
                    eplan.assemble_matrix(t.form, t.params, None, j > 0)
            if lift_into is not None:
                lift = np.zeros(self.nrows)
                eplan.download_into(None, lift)
                lift_into += lift
            plan.add_matrix_from(eplan)

    def _device_matrix(self, plan, matdata):
        """all matrix terms of all triangulations into the plan's device matrix (overwritten)"""
        if matdata.const_Ke is not None:
            if matdata.extra:
                raise NotImplementedError("Fill cell matrices for forms over several triangulations")
            plan.assemble_matrix_const(matdata.const_Ke, None, False)
        for k, t in enumerate(matdata.terms):
            if t.state is not None:
                self._set_dirichlet(plan, t.state)
            plan.assemble_matrix(t.form, t.params, None, k > 0)
        self._assemble_extra_matrices(plan, matdata)

    def assemble_matrix_add_(self, A, matdata, add=True):
        plan = self.plan(matdata.measure, self._touched(matdata.terms))
        self._check(A, self._pattern_plan(matdata))
        if not matdata.terms and matdata.const_Ke is None:
            if not add:
                self._zero_matrix(A)
            return A
        direct = type(self)._fetch_matrix is B200SparseMatrixAssembler._fetch_matrix and not matdata.extra
        if direct and matdata.const_Ke is not None:
            plan.assemble_matrix_const(matdata.const_Ke, A.nzval, add)
            return A
        if direct and len(matdata.terms) == 1:   # one term straight into the caller's array (add: uploaded first, accumulated on the device)
            t = matdata.terms[0]
            if t.state is not None:
                self._set_dirichlet(plan, t.state)
            plan.assemble_matrix(t.form, t.params, A.nzval, add)
            return A
        self._device_matrix(plan, matdata)
        return self._fetch_matrix(A, self._finish_matrix(plan, matdata), add)

    def _zero_matrix(self, A):
        A.nzval[:] = 0.0

    def assemble_matrix_(self, A, matdata):
        return self.assemble_matrix_add_(A, matdata, add=False)

    def assemble_vector_add_(self, b, vecdata, add=True):
        bb = self._vec(b)
        plan = self.plan(vecdata.measure, self._vector_touched(), vecdata.glued)
        if not vecdata.terms and not add:
            bb[:] = 0.0
        for k, t in enumerate(vecdata.terms):
            if t.state is not None:
                self._set_dirichlet(plan, t.state)
            if t.form == lib.FORM_FACET_VEC:   # params = (coef, test kind, data kind); g at the facet points when data kind 0
                fq, params = None, t.params
                if t.fq is not None:
                    xq = plan.quadrature_points()
                    fq = np.ascontiguousarray(np.asarray(t.fq(xq.reshape(-1, xq.shape[2])), dtype=np.float64).reshape(xq.shape[0], xq.shape[1], -1))
            else:
                fq, params = self._fq(plan, t)
            plan.assemble_vector(t.form, params, fq, bb, add or k > 0)
        for e in vecdata.extra:   # further triangulations (Neumann terms on a BoundaryTriangulation): accumulate into the same vector
            self.assemble_vector_add_(b, VecData(e.terms, e.measure, glued=e.glued), add=True)
        return b

    def assemble_vector_(self, b, vecdata):
        return self.assemble_vector_add_(b, vecdata, add=False)

    def assemble_matrix_and_vector_add_(self, A, b, data, add=True):
        """numeric_loop_matrix_and_vector! (src/FESpaces/SparseMatrixAssemblers.jl:365-405): the paired (matrix, vector) terms in one
        fused pass with the Dirichlet lifting, then the un-paired matrix terms and vector terms."""
        matdata, vecdata, uhd = data
        plan = self.plan(matdata.measure, self._touched(matdata.terms))
        self._check(A, self._pattern_plan(matdata))
        bb = self._vec(b)
        if not matdata.terms:
            raise NotImplementedError("AffineFEOperator without a bulk matrix term")
        state_form = matdata.terms[0].state is not None   # residual_and_jacobian: u_h is in the forms, no lifting
        self._set_dirichlet(plan, matdata.terms[0].state if state_form else uhd)
        vparts = [VecData(vecdata.terms, vecdata.measure, glued=vecdata.glued)] + list(vecdata.extra)
        paired, rest = None, []
        for part in vparts:
            terms = list(part.terms)
            if paired is None and terms and _same_domain(part.measure, matdata.measure):
                paired = terms.pop(0)
            if terms:
                rest.append(VecData(terms, part.measure, glued=part.glued))
        direct = type(self)._fetch_matrix is B200SparseMatrixAssembler._fetch_matrix
        if paired is not None:
            fq, vparams = self._fq(plan, paired)
            vform = paired.form
        else:   # no vector term on the bulk quadrature: the lifting alone (zero source)
            fq, vparams, vform = None, (0.0,) * sum(f.ncomp for f in self.test_fields), lib.FORM_SOURCE
        t0 = matdata.terms[0]
        if direct and len(matdata.terms) == 1 and not matdata.extra:
            plan.assemble_matrix_and_vector(t0.form, t0.params, vform, vparams, fq, A.nzval, bb, add)
        else:
            # device-resident: every matrix term with its share of the lifting, boundary matrices merged, then one download
            vec0 = bb.copy() if add else None
            tmpb = np.zeros(self.nrows)
            plan.assemble_matrix_and_vector(t0.form, t0.params, vform, vparams, fq, None, tmpb, False)
            for t in matdata.terms[1:]:
                if state_form:
                    raise NotImplementedError("residual_and_jacobian with several Jacobian terms")
                plan.assemble_matrix_and_vector(t.form, t.params, lib.FORM_SOURCE, (0.0,) * len(vparams), None, None, None, True)
            if len(matdata.terms) > 1:
                plan.download_into(None, tmpb)
            self._assemble_extra_matrices(plan, matdata, uhd=uhd, lift_into=None if state_form else tmpb)
            self._fetch_matrix(A, self._finish_matrix(plan, matdata, uhd=uhd), add)
            bb[:] = tmpb if vec0 is None else vec0 + tmpb
        for part in rest:
            self.assemble_vector_add_(b, part, add=True)
        return A, b

    def assemble_matrix_and_vector_(self, A, b, data):
        return self.assemble_matrix_and_vector_add_(A, b, data, add=False)

    def assemble_matrix(self, matdata):
        # the numeric phase overwrites every stored entry: no need to zero the freshly allocated values first
        # ... and the download of the pattern overlaps the numeric phase (completed by the numeric call's synchronisation)
        empty = not matdata.terms and matdata.const_Ke is None
        A = self.allocate_matrix(matdata, zero=empty, wait=empty)
        try:
            return self.assemble_matrix_(A, matdata)
        except Exception:
            self.ctx.synchronize_quiet()   # the asynchronous pattern download must not outlive the pinned arrays it writes
            raise

    def assemble_vector(self, vecdata):
        return self.assemble_vector_(self.allocate_vector(vecdata), vecdata)

    def assemble_matrix_and_vector(self, data):
        A, b = self.allocate_matrix_and_vector(data, wait=False, zero=False)  # the numeric call overwrites and synchronises
        try:
            return self.assemble_matrix_and_vector_(A, b, data)
        except Exception:
            self.ctx.synchronize_quiet()
            raise


class B200ConstrainedSparseMatrixAssembler(B200SparseMatrixAssembler):
    """SparseMatrixAssembler(U, V) on spaces with linear constraints (FESpaceWithLinearConstraints).  The reference multiplies every
    cell matrix / vector by the cell-wise constraint matrices before the scatter (attach_constraints_rows / _cols,
    src/FESpaces/FESpaceInterface.jl:361-387).  Here the unconstrained space (free and Dirichlet DoFs in one numbering) is assembled
    by the ordinary device path -- every fast kernel applies -- and the assembled arrays are folded on the device,
    A_c = T^T A T, b_c = T^T b - A_c[:, Dirichlet masters] u_D (gb200_plan_fold_constraints), into a plan whose pattern comes from the
    master DoF tables of the cells: the same matrix, summed in another order."""

    def __init__(self, U, V, ctx=None, deterministic=False, **kw):
        if kw.get("strategy") is not None or kw.get("col_range") is not None:
            raise NotImplementedError("constrained spaces with a non-default AssemblyStrategy")
        cu, cv = _base(U), _base(V)
        if not (isinstance(cu, FESpaceWithLinearConstraints) and cu is cv):
            raise NotImplementedError("trial and test spaces must be built on the same FESpaceWithLinearConstraints")
        self.cspace = cv
        self.inner = B200SparseMatrixAssembler(cv.extended, cv.extended, ctx=ctx, deterministic=deterministic)
        self.U, self.V = U, V
        self.ctx = self.inner.ctx
        self.test_fields, self.trial_fields = [cv], [cu]
        self.nrows = self.ncols = self.ncols_assembled = cv.num_free_dofs()
        self.strategy = None
        self.row_offsets = self.col_offsets = [0]
        self._plans = {}
        self._mapped = {}

    # -- the plan that carries the constrained pattern and receives the folded arrays (no integrand is evaluated on it)
    def _cplan(self, matdata):
        skels = [e.measure.trian for e in (matdata.extra if matdata is not None else []) if e.glued == "skeleton"]
        key = ("constrained",) + tuple(id(tr) for tr in skels)
        if key in self._plans:
            return self._plans[key][0]
        cs = self.cspace
        ext = cs.extended.cell_dof_ids
        m = cs.model
        first = np.arange(m.num_cells(), dtype=np.int64)
        second = first
        for tr in skels:
            first, second = np.concatenate([first, tr.cells_plus]), np.concatenate([second, tr.cells_minus])
        pair = np.concatenate([ext[first], ext[second]], axis=1) if skels else ext
        table = cs.master_table(np.ascontiguousarray(pair))
        # a "reference element" with one scalar shape function per padded master slot: only the symbolic phase looks at this plan
        mesh = (m if not skels else DiscreteModel(m.node_coordinates, m.cell_node_ids[first], m.ptype)).device_mesh(self.ctx)
        xq, w = rf.Quadrature(m.ptype, 1)
        Ng, dNg = rf.tabulate_lagrangian(m.ptype, 1, xq)
        geo = lib.DeviceRefEl(self.ctx, w, Ng, dNg, 1)
        width = table.shape[1]
        refel = lib.DeviceRefEl(self.ctx, w, np.zeros((len(w), width)), np.zeros((len(w), width, m.D)), 1)
        sp = lib.DeviceSpace(self.ctx, mesh, refel, table, cs.num_free_dofs(), cs.num_dirichlet_dofs())
        p = lib.DevicePlan(self.ctx, mesh, geo, [sp], [sp], None, [0], [0], self.nrows, self.ncols)
        self._plans[key] = (p, (skels, table))
        return p

    def _pattern_plan(self, matdata):
        return self._cplan(matdata)

    def plan(self, measure, touched=None, glued=False):
        return self.inner.plan(measure, touched, glued)

    def _touched(self, terms):
        return self.inner._touched(terms)

    def _dirichlet_master_values(self, uhd):
        dv = getattr(uhd, "dirichlet_values", None)
        if dv is None:
            dv = getattr(self.U, "dirichlet_values", None)
        return np.zeros(self.cspace.num_dirichlet_dofs()) if dv is None else np.asarray(dv, dtype=np.float64)

    def _fold(self, matdata, src_plan, uhd, with_matrix, with_vector):
        cs = self.cspace
        cp = self._cplan(matdata)
        cp.fold_constraints_from(src_plan, cs.DOF_to_mDOFs_ptrs, cs.DOF_to_mdofs, cs.DOF_to_coeffs,
                                 self._dirichlet_master_values(uhd) if with_vector else None, with_matrix, with_vector)
        return cp

    def _inner_matrix(self, matdata):
        """all matrix terms of the unconstrained space, device-resident; returns the plan that holds them"""
        if matdata.const_Ke is not None:
            raise NotImplementedError("Fill cell matrices on a constrained space")
        if any(t.state is not None for t in matdata.terms):
            raise NotImplementedError("forms that carry u_h on a constrained space")
        plan = self.inner.plan(matdata.measure, self.inner._touched(matdata.terms))
        self.inner._device_matrix(plan, matdata)
        return self.inner._finish_matrix(plan, matdata)

    def allocate_vector(self, vecdata):
        return np.zeros(self.nrows)

    def assemble_matrix_add_(self, A, matdata, add=True):
        self._check(A, self._cplan(matdata))
        if not matdata.terms:
            if not add:
                self._zero_matrix(A)
            return A
        cp = self._fold(matdata, self._inner_matrix(matdata), None, True, False)
        self._fetch_matrix(A, cp, add)
        self.ctx.synchronize()
        return A

    def assemble_vector_add_(self, b, vecdata, add=True):
        bext = self.inner.assemble_vector(vecdata)           # accumulated over the triangulations of the form
        src = self.inner.plan(vecdata.measure, self.inner._vector_touched(), vecdata.glued)
        if src.nrows != len(bext):
            raise NotImplementedError("vector terms of this kind on a constrained space")
        src.upload_vector(bext)
        cp = self._cplan(None)   # (the vector fold needs no pattern: the bulk constrained plan serves)
        cs = self.cspace
        cp.fold_constraints_from(src, cs.DOF_to_mDOFs_ptrs, cs.DOF_to_mdofs, cs.DOF_to_coeffs, None, False, True)
        tmp = np.zeros(self.nrows)
        cp.download_into(None, tmp)
        if add:
            b += tmp
        else:
            b[:] = tmp
        return b

    def assemble_matrix_and_vector_add_(self, A, b, data, add=True):
        matdata, vecdata, uhd = data
        self._check(A, self._cplan(matdata))
        src = self._inner_matrix(matdata)
        bext = self.inner.assemble_vector(vecdata)
        src.upload_vector(bext)
        cp = self._fold(matdata, src, uhd, True, True)
        self._fetch_matrix(A, cp, add)
        tmp = np.zeros(self.nrows)
        cp.download_into(None, tmp)
        self.ctx.synchronize()   # (an asynchronous pattern download of allocate_matrix completes here)
        if add:
            b += tmp
        else:
            b[:] = tmp
        return A, b


class B200BlockSparseMatrixAssembler(B200SparseMatrixAssembler):
    """BlockSparseMatrixAssembler (src/MultiField/BlockSparseMatrixAssemblers.jl:19-33): trial and test spaces with
    BlockMultiFieldStyle(); matrices come back as a BlockMatrix of SparseMatrixCSC (block-local ids), vectors as a BlockVector.
    The numeric phase is the same single device assembly; the blocks are extracted on the device
    (gb200_plan_get_block_pattern / gb200_plan_download_block)."""

    def __init__(self, U, V, **kw):
        super().__init__(U, V, **kw)
        if self.strategy is not None:
            raise NotImplementedError("BlockSparseMatrixAssembler with a non-default AssemblyStrategy")
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

    def _vec(self, b):
        return b.array

    def _zero_matrix(self, A):
        for row in A.blocks:
            for blk in row:
                blk.nzval[:] = 0.0

    def _fetch_matrix(self, A, plan, add):
        for i, row in enumerate(A.blocks):
            for j, blk in enumerate(row):
                if add:   # assemble_matrix_add! (the stage loops of src/ODEs/ODEOpsFromTFEOps.jl:124-405): block += device block
                    blk.nzval += plan.download_block(i, j, np.zeros(len(blk.nzval)))
                else:
                    plan.download_block(i, j, blk.nzval)
        return A


class B200CSRSparseMatrixAssembler(B200SparseMatrixAssembler):
    """SparseMatrixAssembler(SparseMatrixCSR{Bi,Float64,Int}, Vector{Float64}, U, V) (src/FESpaces/SparseMatrixAssemblers.jl:127-153
    with the CSR builder of src/Algebra/SparseMatrixCSR.jl:31-75) and SymSparseMatrixCSR{Bi} (src/Algebra/SymSparseMatrixCSR.jl:1-50:
    the upper triangle of the CSR, entries below the diagonal are skipped by `add_entry!`): same device assembly, results
    delivered in CSR order."""

    def __init__(self, U, V, mat_type, **kw):
        super().__init__(U, V, **kw)
        self.mat_type = mat_type
        self.sym = issubclass(mat_type, SymSparseMatrixCSR)
        self._upper = {}

    def get_matrix_type(self):
        return self.mat_type

    def _upper_of(self, plan, rowptr=None, colval=None):
        """positions of the upper triangle (col >= row) inside the CSR arrays of the plan (SymSparseMatrixCSR keeps only those)"""
        if id(plan) not in self._upper:
            if rowptr is None:
                rowptr, colval = plan.csr_pattern(self.mat_type.Bi)
            rows = np.repeat(np.arange(self.nrows, dtype=np.int64), np.diff(rowptr)) + self.mat_type.Bi
            self._upper[id(plan)] = np.nonzero(colval >= rows)[0]
        return self._upper[id(plan)]

    def allocate_matrix(self, matdata, zero=True, wait=True):
        plan = self.plan(matdata.measure, self._touched(matdata.terms))
        rowptr, colval = plan.csr_pattern(self.mat_type.Bi)
        if not self.sym:
            return self.mat_type(self.nrows, plan.ncols, rowptr, colval, np.zeros(plan.nnz))
        if self.nrows != plan.ncols:
            raise ValueError("SymSparseMatrixCSR needs a square system")
        up = self._upper_of(plan, rowptr, colval)
        rows = np.repeat(np.arange(self.nrows, dtype=np.int64), np.diff(rowptr))
        counts = np.bincount(rows[up], minlength=self.nrows)
        rp = np.concatenate([[0], np.cumsum(counts)]) + self.mat_type.Bi
        return self.mat_type(self.nrows, plan.ncols, rp.astype(np.int64), colval[up].copy(), np.zeros(len(up)))

    def _check(self, A, plan):
        n = len(self._upper_of(plan)) if self.sym else plan.nnz
        if not isinstance(A, SparseMatrixCSR) or len(A.nzval) != n or A.shape != (self.nrows, plan.ncols):
            raise ValueError("matrix was not allocated by this assembler for this form")

    def _fetch_matrix(self, A, plan, add):
        if not plan.nnz:
            return A
        vals = plan.download_csr(np.zeros(plan.nnz))
        if self.sym:
            vals = vals[self._upper_of(plan)]
        if add:
            A.nzval += vals
        else:
            A.nzval[:] = vals
        return A


def SparseMatrixAssembler(*args, **kw):
    """SparseMatrixAssembler(U, V) | SparseMatrixAssembler(mat_type, vec_type, U, V[, strategy])
    (src/FESpaces/SparseMatrixAssemblers.jl:127-160)."""
    if len(args) in (4, 5):
        mat_type, vec_type, U, V = args[:4]
        if len(args) == 5:
            kw = dict(kw, strategy=args[4])
        if isinstance(mat_type, type) and issubclass(mat_type, SparseMatrixCSR):
            return B200CSRSparseMatrixAssembler(U, V, mat_type, **kw)
        if mat_type is not SparseMatrixCSC:
            raise NotImplementedError("matrix type %r: SparseMatrixCSC, SparseMatrixCSR{Bi} and SymSparseMatrixCSR{Bi} are on the B200 path" % (mat_type,))
    else:
        U, V = args
    bu = isinstance(getattr(U, "style", None), BlockMultiFieldStyle)
    bv = isinstance(getattr(V, "style", None), BlockMultiFieldStyle)
    if bu != bv:
        raise NotImplementedError("trial and test spaces must both have BlockMultiFieldStyle (BlockSparseMatrixAssemblers.jl:104-106)")
    if bu:
        return B200BlockSparseMatrixAssembler(U, V, **kw)
    if has_constraints(U) or has_constraints(V):
        return B200ConstrainedSparseMatrixAssembler(U, V, **kw)
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
    uhd = _zero_trial_function(U)
    data = collect_cell_matrix_and_vector(U, V, f(get_trial_fe_basis(U), get_fe_basis(V)), b(get_fe_basis(V)), uhd)
    return a.assemble_matrix_and_vector(data)


def _zero_trial_function(U):
    """uhd = zero(trial) of AffineFEOperator (src/FESpaces/AffineFEOperators.jl:50-54): free values 0, the trial space's Dirichlet
    values.  Multi-field: None -- the assembler then takes every field's own Dirichlet values."""
    if isinstance(U, MultiFieldFESpace):
        return None
    return FEFunction(U, np.zeros(U.num_free_dofs()))


class AffineFEOperator:
    """AffineFEOperator(a, l, U, V[, assem]) (src/FESpaces/AffineFEOperators.jl:23-54): assembles A and
    b = l(v) - a(u_D, v) in one fused pass with Dirichlet lifting."""

    def __init__(self, a, l, U, V, assem=None):
        self.trial, self.test = U, V
        self.assem = assem or SparseMatrixAssembler(U, V)
        uhd = _zero_trial_function(U)
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
        if len(matdata.terms) == 1 and len(vecdata.terms) == 1 and _same_domain(matdata.measure, vecdata.measure) \
                and not matdata.extra and not vecdata.extra and matdata.terms[0].state is not None and vecdata.terms[0].fq is None:
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
