"""Multi-GPU partition of the assembly (one process per GPU).

Cells are independent units of work; the only coupling is the scatter into shared columns.  The partition is by
*column ownership*: cells are cut into `world` contiguous index ranges (Cartesian numbering is z-slowest, so a range is a
z-slab; 6 tets of a hexahedron are consecutive), a free DoF belongs to the rank of its lowest-numbered incident cell (the
interface node-plane goes to the lower slab), and rank r assembles exactly the CSC columns it owns, from every
cell that touches one of them (its own slab of cells + one ghost layer).  This is the column-mask / column-map of Gridap's
`AssemblyStrategy` (src/FESpaces/Assemblers.jl:31-55) -- the same idea GridapDistributed calls a fully-assembled strategy --
and it needs no exchange of partial nnz values: each rank's result is its column set of the global CSC, complete and
identical to the single-GPU result.  Works for any Lagrangian space of the package (Q1, Q2, P1, P2; scalar / vector) and for
multi-field spaces (every field of the consecutive numbering is cut by the same cell ranges: per-field column sets).
`gather_csc_owned` interleaves the ranks' columns back into the global matrix.
"""
import numpy as np

from .algebra import SparseMatrixCSC
from .fespaces import FEFunction, MultiFieldFESpace, TrialFESpace
from .geometry import DiscreteModel, Triangulation


def column_ranges(nfree, world):
    """contiguous, balanced ownership ranges: rank r owns 1-based ids lo < id <= hi."""
    cuts = [(nfree * r) // world for r in range(world + 1)]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def cell_ranges(ncells, world, group=1):
    """contiguous cell ranges [lo, hi) per rank; `group` consecutive cells stay together (the 6 tets of a hexahedron)"""
    ng = ncells // group
    cuts = [((ng * r) // world) * group for r in range(world)] + [ncells]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


class _LocalSpace:
    """The part of an FE space a rank needs: ids of its local cells (global numbering)."""

    def __init__(self, V, model, cell_dof_ids):
        self.model, self.reffe, self.ncomp, self.order = model, V.reffe, V.ncomp, V.reffe.order
        self.cell_dof_ids = np.ascontiguousarray(cell_dof_ids)
        self.nfree, self.ndirichlet = V.nfree, V.ndirichlet
        self.dirichlet_values = getattr(V, "dirichlet_values", None)
        self._device = {}

    def num_free_dofs(self):
        return self.nfree

    def num_dirichlet_dofs(self):
        return self.ndirichlet

    def get_cell_dof_ids(self):
        return self.cell_dof_ids

    def device_space(self, ctx, refel_key, refel, ids=None):
        from . import lib
        key = (id(ctx), refel_key)
        if key not in self._device:
            self._device[key] = lib.DeviceSpace(ctx, self.model.device_mesh(ctx), refel, self.cell_dof_ids, self.nfree, self.ndirichlet)
        return self._device[key]


def _local_model(model, local_cells):
    cn = model.cell_node_ids[local_cells].astype(np.int64) - 1
    used = np.unique(cn)
    remap = np.full(model.num_nodes(), -1, dtype=np.int64)
    remap[used] = np.arange(len(used))
    lm = DiscreteModel(model.node_coordinates[used], (remap[cn] + 1).astype(np.int32), model.ptype)
    lm._partition_parent = model   # measures of the global model are accepted by assemblers on the local one
    return lm


class SlabPartition:
    """Single-field partition by contiguous, balanced COLUMN ranges (Cartesian Q1 numbering: z-slabs of nodes)."""

    def __init__(self, model, V, world, rank):
        base = V.space if hasattr(V, "space") else V
        self.world, self.rank = world, rank
        self.col_range = column_ranges(base.nfree, world)[rank]
        lo, hi = self.col_range
        ids = base.cell_dof_ids
        touching = ((ids > lo) & (ids <= hi)).any(axis=1)
        self.local_cells = np.nonzero(touching)[0]
        self.local_model = _local_model(model, self.local_cells)
        self.local_space = _LocalSpace(base, self.local_model, ids[self.local_cells])
        self.ncells_owned = model.num_cells() / world  # work share (cells are shared at the interfaces)
        self.nrows = base.nfree

    def assembler(self, U, V, ctx=None):
        from .assemblers import B200SparseMatrixAssembler
        return B200SparseMatrixAssembler(self.local_space, self.local_space, ctx=ctx, col_range=self.col_range)


def slab_partition(model, V, world, rank):
    return SlabPartition(model, V, world, rank)


class Partition:
    """General partition (any supported space, single- or multi-field): see the module docstring.

    owned[k]      bool per free DoF of field k: this rank assembles that column
    owned_ids     1-based global column ids (consecutive multi-field numbering) this rank assembles, ascending: local column
                  j of the rank's matrix is global column owned_ids[j]
    local_cells   the cells the rank needs (own range + ghosts), ascending -> the per-column summation order is the serial one
    """

    def __init__(self, model, U, V, world, rank):
        from .assemblers import OwnedColumns, _base, _fields
        self.world, self.rank, self.model = world, rank, model
        tests = [_base(s) for s in _fields(V)]
        trials = _fields(U)
        nc = model.num_cells()
        group = 6 if model.ptype == "TET" and nc % 6 == 0 else 2 if model.ptype == "TRI" and nc % 2 == 0 else 1
        self.cell_range = cell_ranges(nc, world, group)[rank]
        cell_rank = np.empty(nc, dtype=np.int32)
        for r, (lo, hi) in enumerate(cell_ranges(nc, world, group)):
            cell_rank[lo:hi] = r
        self.owned = []
        touching = np.zeros(nc, dtype=bool)
        for t in tests:
            ids = t.cell_dof_ids
            owner = np.full(t.nfree, world, dtype=np.int32)
            pos = ids > 0
            np.minimum.at(owner, ids[pos] - 1, np.broadcast_to(cell_rank[:, None], ids.shape)[pos])
            mine = owner == rank
            self.owned.append(mine)
            hit = np.zeros(ids.shape, dtype=bool)
            hit[pos] = mine[ids[pos] - 1]
            touching |= hit.any(axis=1)
        self.local_cells = np.nonzero(touching)[0]
        self.local_model = _local_model(model, self.local_cells)
        locals_ = []
        for t, u in zip(tests, trials):
            ls = _LocalSpace(t, self.local_model, t.cell_dof_ids[self.local_cells])
            ls.dirichlet_values = getattr(u, "dirichlet_values", None)
            locals_.append(ls)
        multi = isinstance(V, MultiFieldFESpace)
        self.local_space = MultiFieldFESpace(locals_, style=V.style) if multi else locals_[0]
        self.strategy = OwnedColumns(np.concatenate(self.owned))
        self.owned_ids = self.strategy.owned_ids
        self.nrows = V.num_free_dofs()
        self.ncells_owned = self.cell_range[1] - self.cell_range[0]

    def assembler(self, ctx=None, deterministic=False):
        from .assemblers import B200SparseMatrixAssembler
        return B200SparseMatrixAssembler(self.local_space, self.local_space, ctx=ctx, deterministic=deterministic, strategy=self.strategy)

    def local_function(self, uh):
        """the FE function on the local spaces (free / Dirichlet vectors stay global: u_h is gathered through the unmasked ids)"""
        return FEFunction(self.local_space, uh.free_values, uh.dirichlet_values)

    def owned_rows(self):
        """0-based global rows that are complete on this rank (every cell touching them is local)"""
        return self.owned_ids - 1


def partition(model, U, V, world, rank):
    return Partition(model, U, V, world, rank)


def gather_csc(slabs, nrows):
    """concatenate the ranks' column slabs [(colptr, rowval, nzval), ...] into the global SparseMatrixCSC."""
    colptr = [np.array([1], dtype=np.int64)]
    off = 0
    for cp, _, _ in slabs:
        colptr.append(cp[1:] + off)
        off += cp[-1] - 1
    rowval = np.concatenate([s[1] for s in slabs])
    nzval = np.concatenate([s[2] for s in slabs])
    cp = np.concatenate(colptr)
    return SparseMatrixCSC(nrows, len(cp) - 1, cp, rowval, nzval)


def gather_csc_owned(slabs, owned_ids, nrows):
    """global SparseMatrixCSC from the ranks' matrices [(colptr, rowval, nzval), ...] whose local column j is the global column
    owned_ids[rank][j] (1-based): the columns are interleaved back into ascending global order."""
    cat = gather_csc(slabs, nrows)
    gcol = np.concatenate(owned_ids)
    order = np.argsort(gcol, kind="stable")
    counts = np.diff(cat.colptr)[order]
    colptr = np.concatenate([[1], 1 + np.cumsum(counts)]).astype(np.int64)
    starts = (cat.colptr[:-1] - 1)[order]
    idx = np.repeat(starts - (colptr[:-1] - 1), counts) + np.arange(colptr[-1] - 1)
    return SparseMatrixCSC(nrows, len(order), colptr, cat.rowval[idx], cat.nzval[idx])


def gather_vector(slabs, col_ranges):
    """global RHS from the ranks' vectors: rank r's rows (lo, hi] are complete on r (every cell touching an owned DoF is local)."""
    return np.concatenate([np.asarray(b)[lo:hi] for b, (lo, hi) in zip(slabs, col_ranges)])


def gather_vector_owned(vecs, owned_ids, nrows):
    out = np.zeros(nrows)
    for b, ids in zip(vecs, owned_ids):
        out[ids - 1] = np.asarray(b)[ids - 1]
    return out
