"""FE spaces: the *input producer* of the hot path (`cell_dof_ids`, free / Dirichlet DoF numbering).

Vectorised restatement of Gridap's numbering so that a model + reference FE gives the ids Gridap gives:
  * order-1 H1 Lagrangian spaces  -> CLagrangianFESpace: node sweep, components interleaved per node, free ids
    positive / Dirichlet ids negative in the same sweep (src/FESpaces/CLagrangianFESpaces.jl:155-288,356-380;
    chosen by the factory src/FESpaces/FESpaceFactories.jl:61-89)
  * order-2 / order-3 spaces      -> face-based conforming numbering: sweep d = 0..D, faces by id, own DoFs of a
    face component-major (src/FESpaces/ConformingFESpaces.jl:367-423,543-636,823-864); faces that own several nodes
    (order 3) are read through the vertex order of the cell's local face against the face's own frame
  * MultiFieldFESpace (consecutive style) offsets (src/MultiField/MultiFieldFESpaces.jl:356-364,482-488)
"""
import numpy as np

from . import reffes as rf
from .geometry import Triangulation, local_face_vertices


class FESpace:
    """FESpace(model, reffe; dirichlet_tags, dirichlet_masks) == TestFESpace(...)."""

    def __init__(self, model, reffe, conformity="H1", dirichlet_tags=(), dirichlet_masks=None, constraint=None):
        if isinstance(model, Triangulation):
            if getattr(model, "cells", None) is not None and type(model) is Triangulation:
                raise NotImplementedError("FE spaces on a view Triangulation(model, cell_ids): build the space on the model and integrate "
                                          "on the view")
            model = model.model
        if constraint not in (None, "zeromean", ":zeromean"):
            raise NotImplementedError("constraint=%r: valid values are nothing and :zeromean (src/FESpaces/FESpaceFactories.jl:130-143); linear "
                                      "constraints: FESpaceWithLinearConstraints" % (constraint,))
        self.zero_mean = constraint is not None
        self._fixed_dof = 0
        conformity = {None: "H1", "H1": "H1", ":H1": "H1", "L2": "L2", ":L2": "L2"}.get(conformity, conformity)
        if conformity not in ("H1", "L2"):
            raise NotImplementedError("conformity %r: H1 and L2 Lagrangian spaces are on the B200 path" % (conformity,))
        self.model, self.reffe, self.conformity = model, reffe, conformity
        self.ncomp, self.order = reffe.ncomp, reffe.order
        if conformity == "L2":
            if dirichlet_tags:
                raise NotImplementedError("Dirichlet tags on a discontinuous space (impose the data weakly: Nitsche terms)")
            self.dirichlet_tags = []
            self._build_discontinuous()
            self._device = {}
            if self.zero_mean:
                self._fix_constant()
            return
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
        if self.zero_mean:
            self._fix_constant()

    # -- numbering
    def _split(self, tag_index, masks, nown=None):
        """tag_index[n] (0 = UNSET) -> ids[n, ncomp] signed, in sweep order (entity-major, component-minor); with `nown` own nodes
        per entity -> ids[n, ncomp, nown], entity-major, then component, then node"""
        n = len(tag_index)
        isdir = np.zeros((n, self.ncomp), dtype=bool)
        tagged = tag_index > 0
        if len(masks):
            isdir[tagged] = masks[tag_index[tagged] - 1]
        shape = (n, self.ncomp)
        if nown is not None:
            isdir = np.repeat(isdir[:, :, None], nown, axis=2)
            shape = (n, self.ncomp, nown)
        flat = isdir.ravel()
        free_id = np.cumsum(~flat)
        dir_id = np.cumsum(flat)
        ids = np.where(flat, -dir_id, free_id).reshape(shape)
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

    def _fix_constant(self):
        """constraint = :zeromean -> ZeroMeanFESpace(space, Measure(trian, order)) (src/FESpaces/FESpaceFactories.jl:130-137,
        ZeroMeanFESpaces.jl:11-19): the space with its LAST free DoF fixed (FESpaceWithConstantFixed(space, true, num_free_dofs(space)),
        FESpacesWithConstantFixed.jl:14-25,124-163: cell id == dof_to_fix -> -1, smaller ids unchanged), only when the space has no
        Dirichlet DoF; FE functions are shifted to zero mean afterwards (zero_mean_values)."""
        if self.ndirichlet != 0:     # (DoNotFixConstant)
            return
        fix = self.nfree
        self._fixed_dof = fix
        ids = self.cell_dof_ids
        ids[ids == fix] = -1
        if self._entity_ids is not None:
            self._entity_ids = np.where(self._entity_ids == fix, -1, self._entity_ids)
        if getattr(self, "node_and_comp_to_dof", None) is not None:
            self.node_and_comp_to_dof = np.where(self.node_and_comp_to_dof == fix, -1, self.node_and_comp_to_dof).astype(np.int32)
        self.nfree, self.ndirichlet = fix - 1, 1

    def zero_mean_values(self, free_values, dirichlet_values):
        """FEFunction(f::ZeroMeanFESpace, fv, dv) (ZeroMeanFESpaces.jl:40-75): the constant c = -(sum_i u_i vol_i) / vol is added to
        all DoF values, vol_i = assemble_vector(v -> int(v) dOmega, unconstrained space) -- assembled on the device, once."""
        if not self._fixed_dof:
            return free_values, dirichlet_values
        if getattr(self, "_vol_i", None) is None:
            from .assemblers import assemble_vector
            from .celldata import Integral, Measure
            twin = FESpace(self.model, self.reffe, conformity=self.conformity)
            dO = Measure(Triangulation(self.model), self.order)
            self._vol_i = assemble_vector(lambda v: Integral(v * 1.0) * dO, twin)
        vol_i = self._vol_i
        fix = self._fixed_dof
        c = -(np.dot(free_values, vol_i[:fix - 1]) + dirichlet_values[0] * vol_i[fix - 1]) / vol_i.sum()
        return np.asarray(free_values) + c, np.asarray(dirichlet_values) + c

    def _build_discontinuous(self):
        """conformity = :L2 (src/FESpaces/DiscontinuousFESpaces.jl, compute_discontinuous_cell_dofs): the DoFs of a cell are its own,
        numbered cell after cell in the local order of the reference FE; no Dirichlet DoFs."""
        h1 = FESpace(self.model, self.reffe)   # (for the DoF nodes: same reference FE, same local order)
        nc, nl = h1.cell_dof_ids.shape
        self.cell_dof_ids = np.ascontiguousarray((np.arange(nc * nl, dtype=np.int64) + 1).reshape(nc, nl).astype(np.int32))
        self.nfree, self.ndirichlet = nc * nl, 0
        fx, fc, _, _ = h1.dof_coordinates()
        src = h1.cell_dof_ids.astype(np.int64).ravel() - 1
        self._l2_X, self._l2_comp = fx[src], fc[src]
        self._entity_ids = None

    def _build_conforming(self, tags, masks):
        m = self.model
        D = m.D
        if self.order >= 3:
            return self._build_conforming_high_order(tags, masks)
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

    def _build_conforming_high_order(self, tags, masks):
        """Faces that own several nodes (order >= 3).  The own DoFs of a face live in the face's own frame -- its vertices in the
        local order of the first cell that holds it -- component-major, node-minor (`_generate_face_own_dofs`,
        src/ReferenceFEs/LagrangianRefFEs.jl:254-272), numbered in the sweep d = 0..D, faces by id, free / Dirichlet per component
        (src/FESpaces/ConformingFESpaces.jl:574-636).  A cell whose local face lists the vertices in another order reads them through
        the node permutation of that vertex permutation (`CellDofsNonOriented`, :844-864; `compute_cell_permutations`,
        src/Geometry/GridTopologies.jl:515-549; `_compute_node_permutations`, src/ReferenceFEs/CLagrangianRefFEs.jl:549-577).
        Here the permutation is not looked up by index: the own node i of the cell's local face has lattice weights on the face's
        vertices; carried to the face's frame through the vertex match they name the face's own node directly."""
        m = self.model
        D, k, ncomp = m.D, self.order, self.ncomp
        lat, own = rf.lagrangian_lattice(m.ptype, k)       # reference lattice [nl, D], {d: [nlf, nown] local node ids}
        nl = len(lat)
        simplex = m.ptype in ("TET", "TRI")
        bary = np.concatenate([k - lat.sum(axis=1, keepdims=True), lat], axis=1) if simplex else None
        cn = m.cell_node_ids.astype(np.int64) - 1
        nc = len(cn)
        cell_ids = np.zeros((nc, ncomp, nl), dtype=np.int64)
        ent_ids, ent_X = [], []
        offset_free = offset_dir = 0
        for d in range(D + 1):
            nown = own[d].shape[1]
            if nown == 0:
                continue
            c2f, _ = m.faces(d)
            frames = m.face_frames(d)                       # [nf, nv] vertices of every face in its own frame
            nf, nv = frames.shape
            tag_index = m.face_tag_index(d, tags) if (d < D and tags) else np.zeros(nf, dtype=np.int32)
            ids, nfree, ndir = self._split(tag_index, masks, nown)             # [nf, ncomp, nown]
            ids = np.where(ids > 0, ids + offset_free, ids - offset_dir)
            offset_free += nfree
            offset_dir += ndir
            lfv = np.array(local_face_vertices(m.ptype, d)) if d < D else np.arange(cn.shape[1])[None, :]
            # weights (integers, sum = k^dim for n-cubes / k for simplices) of the own nodes on the vertices of their local face
            wts = np.zeros((len(lfv), nown, nv), dtype=np.int64)
            for lf in range(len(lfv)):
                for i, ln in enumerate(own[d][lf]):
                    wts[lf, i] = self._vertex_weights(lat, bary, ln, lfv[lf], k, simplex)
            # the same weights for the own nodes of the FRAME (local face 0 stands for every face of this dimension: the weights of
            # own node j on the face's vertices in their local order do not depend on which face it is)
            frame_w = wts[0]                                                   # [nown, nv]
            # physical position of the frame's own nodes: the face's linear map (:516-538 of CLagrangianRefFEs.jl)
            scale = float(frame_w[0].sum())
            ent_X.append(np.einsum("jv,fvx->fjx", frame_w / scale, m.node_coordinates[frames]).reshape(nf * nown, -1))
            ent_ids.append(np.transpose(ids, (0, 2, 1)).reshape(nf * nown, ncomp))
            for lf in range(len(lfv)):
                face = c2f[:, lf] if d < D else np.arange(nc)
                cv = cn[:, lfv[lf]]                                            # [nc, nv] the cell's vertices of this local face
                fr = frames[face]                                              # [nc, nv] the same vertices in the face's frame
                pos = np.argmax(fr[:, :, None] == cv[:, None, :], axis=2)      # frame vertex -> position in the cell's local face
                for i, ln in enumerate(own[d][lf]):
                    w_in_frame = wts[lf, i][pos]                               # [nc, nv] weights of the node on the frame's vertices
                    hit = (w_in_frame[:, None, :] == frame_w[None, :, :]).all(axis=2)                   # [nc, nown]
                    if not hit.any(axis=1).all():
                        raise ValueError("inconsistent mesh topology: a cell's local face does not list the vertices of its global face")
                    j = np.argmax(hit, axis=1)                                                           # the frame's own node
                    cell_ids[:, :, ln] = ids[face, :, j]
        self.nfree, self.ndirichlet = offset_free, offset_dir
        self.cell_dof_ids = np.ascontiguousarray(cell_ids.reshape(nc, ncomp * nl).astype(np.int32))
        self._entity_ids = np.concatenate(ent_ids, axis=0)
        self._dof_nodes_X = np.concatenate(ent_X, axis=0)

    @staticmethod
    def _vertex_weights(lat, bary, ln, face_vertices, k, simplex):
        """integer weights of reference node `ln` on the vertices of the local face that owns it (multilinear on n-cube faces,
        barycentric on simplex faces)"""
        if simplex:
            return bary[ln][face_vertices]
        D = lat.shape[1]
        out = []
        varying = [ax for ax in range(D) if len({(v >> ax) & 1 for v in face_vertices}) == 2]
        for v in face_vertices:
            w = 1
            for ax in varying:
                w *= lat[ln, ax] if (v >> ax) & 1 else k - lat[ln, ax]
            out.append(w)
        return np.array(out, dtype=np.int64)

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
        if self.conformity == "L2":
            D = self._l2_X.shape[1]
            k = len(self._l2_X) - (1 if self._fixed_dof else 0)   # (zero-mean: the last DoF is the fixed one)
            return self._l2_X[:k], self._l2_comp[:k], self._l2_X[k:], self._l2_comp[k:]
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


class _ExtendedSpace:
    """The unconstrained space with its free and Dirichlet DoFs in ONE positive numbering, DOF = dof if dof > 0 else n_fdofs - dof
    (`_dof_to_DOF`, src/FESpaces/FESpacesWithLinearConstraints.jl:356-362): what the device assembles before the constraints are
    applied (gb200_plan_fold_constraints) -- no Dirichlet DoFs, hence no lifting inside."""

    def __init__(self, space):
        n = space.nfree
        ext = lambda ids: np.where(ids > 0, ids, n - ids).astype(np.int32)   # noqa: E731
        self.model, self.reffe, self.ncomp, self.order = space.model, space.reffe, space.ncomp, space.order
        self.conformity = getattr(space, "conformity", "H1")
        self.cell_dof_ids = np.ascontiguousarray(ext(space.cell_dof_ids))
        self.nfree, self.ndirichlet = space.nfree + space.ndirichlet, 0
        self._entity_ids = None if space._entity_ids is None else ext(space._entity_ids)
        self.dirichlet_tags = []
        self._device = {}

    def num_free_dofs(self):
        return self.nfree

    def num_dirichlet_dofs(self):
        return 0

    def get_cell_dof_ids(self):
        return self.cell_dof_ids

    device_space = FESpace.device_space


class FESpaceWithLinearConstraints:
    """FESpaceWithLinearConstraints(sDOF_to_dof, sDOF_to_dofs, sDOF_to_coeffs, space)
    (src/FESpaces/FESpacesWithLinearConstraints.jl:40-120,171-290): slave DoF sDOF (signed id sDOF_to_dof[s] of `space`) equals
    sum_k coeffs[s][k] * (DoF dofs[s][k]); one level of constraints; Dirichlet slaves depend on Dirichlet masters only.
    The free DoFs of this space are the free masters (ascending DOF), its Dirichlet DoFs the Dirichlet masters; its cell DoF ids are
    the masters of every cell (`cell_to_lmdof_to_mdof`; the local order inside a cell is the iteration order of a Set in the reference
    and has no influence on the assembled arrays: ascending here)."""

    def __init__(self, sDOF_to_dof, sDOF_to_dofs, sDOF_to_coeffs, space):
        base = space.space if isinstance(space, TrialFESpace) else space
        self.space = base
        nf, nd = base.nfree, base.ndirichlet
        n = nf + nd
        to_DOF = lambda d: np.where(np.asarray(d) > 0, np.asarray(d), nf - np.asarray(d)).astype(np.int64)   # noqa: E731
        # DOF_to_DOFs / DOF_to_coeffs (_prepare_DOF_to_DOFs, :76-120): identity rows, slave rows replaced
        lens = np.ones(n, dtype=np.int64)
        sD = to_DOF(np.asarray(sDOF_to_dof, dtype=np.int64)) - 1
        lens[sD] = [len(r) for r in sDOF_to_dofs]
        ptrs = np.concatenate([[0], np.cumsum(lens)])
        data = np.zeros(ptrs[-1], dtype=np.int64)
        coef = np.ones(ptrs[-1])
        data[ptrs[:-1]] = np.arange(1, n + 1)
        for s_, D_ in enumerate(sD):
            row = to_DOF(np.asarray(sDOF_to_dofs[s_], dtype=np.int64))
            data[ptrs[D_]:ptrs[D_] + len(row)] = row
            coef[ptrs[D_]:ptrs[D_] + len(row)] = np.asarray(sDOF_to_coeffs[s_], dtype=np.float64)
        # masters (_find_master_dofs, :171-192): every DOF that appears on a right-hand side; they must be unconstrained themselves
        ismaster = np.zeros(n, dtype=bool)
        ismaster[data - 1] = True
        mast = np.nonzero(ismaster)[0]
        if np.any(lens[mast] != 1) or np.any(data[ptrs[mast]] != mast + 1):
            raise ValueError("recursive constraints are not allowed")
        self.mDOF_to_DOF = mast + 1
        self.n_fdofs, self.n_fmdofs = nf, int(ismaster[:nf].sum())
        n_m = len(mast)
        DOF_to_mDOF = np.zeros(n, dtype=np.int64)
        DOF_to_mDOF[mast] = np.arange(1, n_m + 1)
        mD = DOF_to_mDOF[data - 1]                               # (_renumber_constraints!, :194-203)
        self.DOF_to_mDOFs_ptrs = ptrs + 1                        # 1-based, as a Table
        self.DOF_to_mdofs = np.where(mD > self.n_fmdofs, -(mD - self.n_fmdofs), mD).astype(np.int32)   # signed master ids
        self.DOF_to_coeffs = coef
        if np.any((np.repeat(np.arange(n), lens) >= nf) & (self.DOF_to_mdofs > 0)):
            raise ValueError("Dirichlet dofs can only depend on Dirichlet dofs")
        self.nfree, self.ndirichlet = self.n_fmdofs, n_m - self.n_fmdofs
        self.model, self.reffe, self.ncomp, self.order = base.model, base.reffe, base.ncomp, base.order
        self.conformity = getattr(base, "conformity", "H1")
        self.dirichlet_tags = getattr(base, "dirichlet_tags", [])
        self.extended = _ExtendedSpace(base)
        self._master_tables = {}

    def has_constraints(self):
        return True

    def num_free_dofs(self):
        return self.nfree

    def num_dirichlet_dofs(self):
        return self.ndirichlet

    def master_table(self, cell_dofs_ext):
        """cell table of extended DOFs [n, nl] -> the master DoFs of every row (signed, ascending positive then negative... unique),
        padded with 0 to the longest row: `_setup_cell_to_lmdof_to_mdof` (:205-262) for any cell-like table (cells, facet pairs)"""
        key = id(cell_dofs_ext)
        hit = self._master_tables.get(key)
        if hit is not None and hit[0] is cell_dofs_ext:
            return hit[1]
        ids = np.asarray(cell_dofs_ext, dtype=np.int64)
        p0 = self.DOF_to_mDOFs_ptrs - 1
        lens = (p0[1:] - p0[:-1])[ids - 1]                       # [n, nl]
        rows = []
        width = 0
        for r in range(ids.shape[0]):                            # (host preparation, once per table)
            q = np.concatenate([np.arange(p0[d - 1], p0[d]) for d in ids[r]])
            m = np.unique(self.DOF_to_mdofs[q])
            rows.append(m)
            width = max(width, len(m))
        out = np.zeros((ids.shape[0], width), dtype=np.int32)
        for r, m in enumerate(rows):
            out[r, :len(m)] = m
        del lens
        self._master_tables[key] = (cell_dofs_ext, out)
        return out

    def get_cell_dof_ids(self):
        return self.master_table(self.extended.cell_dof_ids)

    @property
    def cell_dof_ids(self):
        return self.get_cell_dof_ids()

    # -- values (gather / scatter_free_and_dirichlet_values, :305-420)
    def scatter_free_and_dirichlet_values(self, fmdof_to_val, dmdof_to_val):
        """values of the masters -> (free, Dirichlet) values of the underlying space"""
        mvals = np.concatenate([np.asarray(fmdof_to_val, dtype=np.float64), np.asarray(dmdof_to_val, dtype=np.float64)])
        md = self.DOF_to_mdofs.astype(np.int64)
        idx = np.where(md > 0, md - 1, self.n_fmdofs - md - 1)
        contrib = mvals[idx] * self.DOF_to_coeffs
        p0 = self.DOF_to_mDOFs_ptrs - 1
        vals = np.add.reduceat(contrib, p0[:-1])
        return vals[:self.n_fdofs], vals[self.n_fdofs:]

    def _master_values(self, fvals, dvals):
        allv = np.concatenate([fvals, dvals])[self.mDOF_to_DOF - 1]
        return allv[:self.n_fmdofs], allv[self.n_fmdofs:]

    def interpolate_free_values(self, g):
        return self._master_values(self.space.interpolate_free_values(g), self.space.interpolate_dirichlet_values(g))[0]

    def interpolate_dirichlet_values(self, g):
        return self._master_values(self.space.interpolate_free_values(g), self.space.interpolate_dirichlet_values(g))[1]


def has_constraints(space):
    """has_constraints(space) (src/FESpaces/FESpaceInterface.jl:330-340)"""
    base = space.space if isinstance(space, TrialFESpace) else space
    return isinstance(base, FESpaceWithLinearConstraints)


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
        base = U.space if isinstance(U, TrialFESpace) else U
        if getattr(base, "_fixed_dof", 0):   # ZeroMeanFESpace: shift to zero mean
            fv, dv = base.zero_mean_values(self.free_values, self.dirichlet_values)
            self.free_values, self.dirichlet_values = np.ascontiguousarray(fv), np.ascontiguousarray(dv)

    def get_free_dof_values(self):
        return self.free_values


def interpolate(g, U):
    base = U.space if isinstance(U, TrialFESpace) else U
    if getattr(base, "_fixed_dof", 0):   # interpolate!(object, free_values, fs::ZeroMeanFESpace): everywhere, then subtract the mean
        return FEFunction(U, base.interpolate_free_values(g), base.interpolate_dirichlet_values(g))
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
