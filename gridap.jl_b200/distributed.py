"""Multi-GPU partition of the assembly (one process per GPU).

Cells are independent units of work; the only coupling is the scatter into shared columns.  The partition is by
*column ownership*: the free DoFs are split into `world` contiguous ranges (Cartesian numbering is z-slowest, so a range
is a z-slab of nodes); rank r assembles exactly the CSC columns it owns, from every cell that touches one of them
(its own slab of cells + one ghost layer on each side).  This is the column-mask of Gridap's `AssemblyStrategy`
(src/FESpaces/Assemblers.jl:31-55) -- the same idea GridapDistributed calls a fully-assembled strategy -- and it needs
no exchange of partial nnz values: each rank's result is its column slab of the global CSC, complete and bit-identical
to the single-GPU result.  The global matrix is the concatenation of the slabs (`gather_csc`).
"""
import numpy as np

from .algebra import SparseMatrixCSC
from .geometry import DiscreteModel


def column_ranges(nfree, world):
    """contiguous, balanced ownership ranges: rank r owns 1-based ids lo < id <= hi."""
    cuts = [(nfree * r) // world for r in range(world + 1)]
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


class SlabPartition:
    def __init__(self, model, V, world, rank):
        base = V.space if hasattr(V, "space") else V
        self.world, self.rank = world, rank
        self.col_range = column_ranges(base.nfree, world)[rank]
        lo, hi = self.col_range
        ids = base.cell_dof_ids
        touching = ((ids > lo) & (ids <= hi)).any(axis=1)
        self.local_cells = np.nonzero(touching)[0]
        cn = model.cell_node_ids[self.local_cells].astype(np.int64) - 1
        used = np.unique(cn)
        remap = np.full(model.num_nodes(), -1, dtype=np.int64)
        remap[used] = np.arange(len(used))
        self.local_model = DiscreteModel(model.node_coordinates[used], (remap[cn] + 1).astype(np.int32), model.ptype)
        self.local_space = _LocalSpace(base, self.local_model, ids[self.local_cells])
        self.ncells_owned = model.num_cells() / world  # work share (cells are shared at the interfaces)
        self.nrows = base.nfree

    def assembler(self, U, V, ctx=None):
        from .assemblers import B200SparseMatrixAssembler
        return B200SparseMatrixAssembler(self.local_space, self.local_space, ctx=ctx, col_range=self.col_range)


def slab_partition(model, V, world, rank):
    return SlabPartition(model, V, world, rank)


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


def gather_vector(slabs, col_ranges):
    """global RHS from the ranks' vectors: rank r's rows (lo, hi] are complete on r (every cell touching an owned DoF is local)."""
    return np.concatenate([np.asarray(b)[lo:hi] for b, (lo, hi) in zip(slabs, col_ranges)])
