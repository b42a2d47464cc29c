"""FE spaces: the *input producer* of the hot path (`cell_dof_ids`, free / Dirichlet DoF numbering).

Vectorised restatement of Gridap's numbering so that a model + reference FE gives the ids Gridap gives:
  * order-1 H1 Lagrangian spaces  -> CLagrangianFESpace: node sweep, components interleaved per node, free ids
    positive / Dirichlet ids negative in the same sweep (src/FESpaces/CLagrangianFESpaces.jl:155-288,356-380;
    chosen by the factory src/FESpaces/FESpaceFactories.jl:61-89)
  * order-2 spaces                -> face-based conforming numbering: sweep d = 0..D, faces by id, own DoFs of a
    face component-major (src/FESpaces/ConformingFESpaces.jl:367-423,543-636,823-864)
  * MultiFieldFESpace (consecutive style) offsets (src/MultiField/MultiFieldFESpaces.jl:356-364,482-488)
"""
import numpy as np

from . import reffes as rf
from .geometry import Triangulation, local_face_vertices


class FESpace:
    """FESpace(model, reffe; dirichlet_tags, dirichlet_masks) == TestFESpace(...)."""

    def __init__(self, model, reffe, conformity="H1", dirichlet_tags=(), dirichlet_masks=None, constraint=None):
        if isinstance(model, Triangulation):
            model = model.model
        if constraint is not None:
            raise NotImplementedError("constrained spaces (constraint=%r) are outside the B200 path" % (constraint,))
        if conformity not in ("H1", None):
            raise NotImplementedError("conformity %r: only H1 Lagrangian spaces are on the B200 path" % (conformity,))
        self.model, self.reffe = model, reffe
        self.ncomp, self.order = reffe.ncomp, reffe.order
        tags = list(dirichlet_tags) if isinstance(dirichlet_tags, (list, tuple)) else [dirichlet_tags]
        if dirichlet_masks is None:
            masks = np.ones((len(tags), self.ncomp), dtype=bool)
        else:
            masks = np.array([np.broadcast_to(np.asarray(m, dtype=bool), (self.ncomp,)) for m in dirichlet_masks]).reshape(len(tags), self.ncomp)
        self.dirichlet_tags = tags
        if self.order == 1:
            self._build_clagrangian(tags, masks)
        else:
            self._build_conforming(tags, masks)
        self._device = {}

    # -- numbering
    def _split(self, tag_index, masks):
        """tag_index[n] (0 = UNSET) -> ids[n, ncomp] signed, in sweep order (entity-major, component-minor)."""
        n = len(tag_index)
        isdir = np.zeros((n, self.ncomp), dtype=bool)
        tagged = tag_index > 0
        if len(masks):
            isdir[tagged] = masks[tag_index[tagged] - 1]
        flat = isdir.ravel()
        free_id = np.cumsum(~flat)
        dir_id = np.cumsum(flat)
        ids = np.where(flat, -dir_id, free_id).reshape(n, self.ncomp)
        return ids, int((~flat).sum()), int(flat.sum())

    def _build_clagrangian(self, tags, masks):
        m = self.model
        tag_index = m.face_tag_index(0, tags) if tags else np.zeros(m.num_nodes(), dtype=np.int32)
        ids, self.nfree, self.ndirichlet = self._split(tag_index, masks)
        self.node_and_comp_to_dof = ids.astype(np.int32)
        cn = m.cell_node_ids.astype(np.int64) - 1  # [nc, nl]
        nc, nl = cn.shape
        # local dof k = lnode + nl*comp (component-major)
        self.cell_dof_ids = np.ascontiguousarray(np.transpose(ids[cn], (0, 2, 1)).reshape(nc, nl * self.ncomp).astype(np.int32))
        # coordinates of the DoF nodes
        X = m.node_coordinates
        self._dof_nodes_X = X  # per entity (node)
        self._entity_ids = ids

    def _build_conforming(self, tags, masks):
        m = self.model
        D = m.D
        simplex = m.ptype in ("TET", "TRI")
        dims = [0, 1] if simplex else list(range(D + 1))
        ent_ids = []
        ent_X = []
        offset_free = 0
        offset_dir = 0
        cell_cols = []
        for d in dims:
            c2f, fverts = m.faces(d)
            nf = len(fverts)
            if d < D and tags:
                tag_index = m.face_tag_index(d, tags)
            else:
                tag_index = np.zeros(nf, dtype=np.int32)
            ids, nfree, ndir = self._split(tag_index, masks)
            ids = np.where(ids > 0, ids + offset_free, ids - offset_dir)
            offset_free += nfree
            offset_dir += ndir
            ent_ids.append(ids)
            ent_X.append(m.node_coordinates[fverts].mean(axis=1))
            cell_cols.append(ids[c2f])  # [nc, nlf, ncomp]
        self.nfree, self.ndirichlet = offset_free, offset_dir
        allc = np.concatenate(cell_cols, axis=1)  # [nc, nlnodes, ncomp]
        nc, nl, _ = allc.shape
        self.cell_dof_ids = np.ascontiguousarray(np.transpose(allc, (0, 2, 1)).reshape(nc, nl * self.ncomp).astype(np.int32))
        self._entity_ids = np.concatenate(ent_ids, axis=0)
        self._dof_nodes_X = np.concatenate(ent_X, axis=0)

    # -- Gridap.FESpaces API names
    def num_free_dofs(self):
        return self.nfree

    def num_dirichlet_dofs(self):
        return self.ndirichlet

    def get_cell_dof_ids(self):
        return self.cell_dof_ids

    def get_triangulation(self):
        return Triangulation(self.model)

    def dof_coordinates(self):
        """(free_X [nfree, D], free_comp, dir_X [ndir, D], dir_comp): node of every DoF (Lagrangian dof basis)."""
        ids = self._entity_ids
        X = self._dof_nodes_X
        D = X.shape[1]
        fx, fc = np.zeros((self.nfree, D)), np.zeros(self.nfree, dtype=np.int64)
        dx, dc = np.zeros((self.ndirichlet, D)), np.zeros(self.ndirichlet, dtype=np.int64)
        for c in range(self.ncomp):
            col = ids[:, c]
            f = col > 0
            fx[col[f] - 1] = X[f]
            fc[col[f] - 1] = c
            d = col < 0
            dx[-col[d] - 1] = X[d]
            dc[-col[d] - 1] = c
        return fx, fc, dx, dc

    def _evaluate(self, g, X, comp):
        """values of g (callable x -> scalar / vector, or constant) at points X for component comp[i]."""
        if callable(g):
            vals = np.asarray(g(X))
        else:
            vals = np.asarray(g, dtype=np.float64)
        if self.ncomp == 1:
            return np.broadcast_to(vals.reshape(-1) if vals.ndim else vals, (len(X),)).astype(np.float64).copy()
        vals = np.broadcast_to(vals, (len(X), self.ncomp)) if vals.ndim <= 1 else vals
        return np.ascontiguousarray(vals[np.arange(len(X)), comp], dtype=np.float64)

    def interpolate_dirichlet_values(self, g):
        _, _, dx, dc = self.dof_coordinates()
        return self._evaluate(g, dx, dc)

    def interpolate_free_values(self, g):
        fx, fc, _, _ = self.dof_coordinates()
        return self._evaluate(g, fx, fc)

    def device_space(self, ctx, refel_key, refel, ids=None):
        key = (id(ctx), refel_key, None if ids is None else id(ids))
        if key not in self._device:
            from . import lib
            mesh = self.model.device_mesh(ctx)
            self._device[key] = lib.DeviceSpace(ctx, mesh, refel, self.cell_dof_ids if ids is None else ids, self.nfree, self.ndirichlet)
        return self._device[key]


TestFESpace = FESpace


class TrialFESpace:
    """TrialFESpace(V, g): same DoFs as V plus Dirichlet values interpolated from g (src/FESpaces/TrialFESpaces.jl)."""

    def __init__(self, V, g=None):
        self.space = V
        self.dirichlet_values = np.zeros(V.ndirichlet) if g is None else V.interpolate_dirichlet_values(g)

    def __getattr__(self, name):
        return getattr(self.space, name)


class FEFunction:
    """FEFunction(U, free_values): u_h with the trial space's Dirichlet values (PosNegReindex of the reference)."""

    def __init__(self, U, free_values, dirichlet_values=None):
        self.space = U
        self.free_values = np.ascontiguousarray(free_values, dtype=np.float64)
        if dirichlet_values is None:
            dirichlet_values = getattr(U, "dirichlet_values", None)
        if dirichlet_values is None:
            dirichlet_values = np.zeros(U.num_dirichlet_dofs())
        self.dirichlet_values = np.ascontiguousarray(dirichlet_values, dtype=np.float64)

    def get_free_dof_values(self):
        return self.free_values


def interpolate(g, U):
    base = U.space if isinstance(U, TrialFESpace) else U
    return FEFunction(U, base.interpolate_free_values(g))


def zero(U):
    return FEFunction(U, np.zeros(U.num_free_dofs()))


class ConsecutiveMultiFieldStyle:
    """src/MultiField/MultiFieldFESpaces.jl:14-22: one global numbering, field k after field k-1."""


class BlockMultiFieldStyle:
    """BlockMultiFieldStyle() (src/MultiField/MultiFieldFESpaces.jl:24-75): one block per field; the assembler returns a
    BlockMatrix / BlockVector (src/MultiField/BlockSparseMatrixAssemblers.jl).  Merged / permuted blocks (NB, SB, P) are not
    on the B200 path."""

    def __init__(self, *args):
        if args:
            raise NotImplementedError("BlockMultiFieldStyle(NB, SB, P) with merged or permuted blocks is not on the B200 path")


class MultiFieldFESpace:
    """MultiFieldFESpace([U1, U2]; style) -- ConsecutiveMultiFieldStyle (default): field k's positive ids are shifted by
    sum_{m<k} num_free_dofs(m) (src/MultiField/MultiFieldFESpaces.jl:356-364,460-488); BlockMultiFieldStyle: same cell ids
    on the device, block-structured results."""

    def __init__(self, spaces, style=None):
        self.spaces = list(spaces)
        self.style = style if style is not None else ConsecutiveMultiFieldStyle()
        if len(self.spaces) > 2:
            raise NotImplementedError("more than 2 fields")
        n = [s.num_free_dofs() for s in self.spaces]
        self.offsets = [int(sum(n[:k])) for k in range(len(n))]
        self.nfree = int(sum(n))

    def num_free_dofs(self):
        return self.nfree

    def __len__(self):
        return len(self.spaces)

    def __getitem__(self, k):
        return self.spaces[k]

    def get_cell_dof_ids(self):
        out = []
        for s, o in zip(self.spaces, self.offsets):
            ids = s.get_cell_dof_ids().copy()
            ids[ids > 0] += o
            out.append(ids)
        return out
