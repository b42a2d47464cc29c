"""Host-side geometry inputs of the assembly path (vectorised numpy; mirrors the names of Gridap.Geometry).

What the hot path needs from a `DiscreteModel` is `get_node_coordinates`, `get_cell_node_ids` and the face labeling
that decides Dirichlet DoFs (`src/Geometry`): everything else of Gridap's geometry layer is out of
scope.  Numbering conventions follow the reference (file:line in each function) so that `cell_dof_ids` built on top
are the ones Gridap would hand to the assembler.
"""
import numpy as np

from . import lib

_DIM = {"SEG": 1, "QUAD": 2, "HEX": 3, "TRI": 2, "TET": 3}
_CELLTYPE = {"SEG": lib.SEG2, "QUAD": lib.QUAD4, "HEX": lib.HEX8, "TRI": lib.TRI3, "TET": lib.TET4}

# simplexify tables (src/ReferenceFEs/ExtrusionPolytopes.jl:290-299), 0-based local vertices
_HEX_TO_TETS = np.array([[0, 1, 2, 6], [0, 1, 4, 6], [1, 2, 3, 6], [1, 3, 6, 7], [1, 4, 5, 6], [1, 5, 6, 7]])
_QUAD_TO_TRIS = np.array([[0, 1, 2], [1, 2, 3]])


def ncube_faces(D):
    """(dim, extrusion bits, anchor bits) of the n-faces of the D-cube in Gridap's order: by dimension, then
    extrusion, then anchor, last axis most significant (src/ReferenceFEs/ExtrusionPolytopes.jl:460-474)."""
    faces = [(bin(e).count("1"), e, a) for e in range(2 ** D) for a in range(2 ** D) if not (a & e)]
    faces.sort()
    return faces


def local_face_vertices(ptype, d):
    """0-based local vertex ids of the local d-faces of a polytope (HEX/QUAD generated; TET/TRI tables)."""
    if ptype in ("HEX", "QUAD", "SEG"):
        D = _DIM[ptype]
        out = []
        for (dim, e, a) in ncube_faces(D):
            if dim != d:
                continue
            axes = [k for k in range(D) if (e >> k) & 1]
            vs = []
            for bits in range(2 ** len(axes)):
                v = a
                for t, k in enumerate(axes):
                    if (bits >> t) & 1:
                        v |= 1 << k
                vs.append(v)
            out.append(sorted(vs))
        return out
    if ptype == "TET":
        return {0: [[0], [1], [2], [3]], 1: [[0, 1], [0, 2], [1, 2], [0, 3], [1, 3], [2, 3]],
                2: [[0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]], 3: [[0, 1, 2, 3]]}[d]
    if ptype == "TRI":
        return {0: [[0], [1], [2]], 1: [[0, 1], [0, 2], [1, 2]], 2: [[0, 1, 2]]}[d]
    raise NotImplementedError(ptype)


def _cartesian_index(nodes, partition):
    """node ids (0-based, any shape) -> Cartesian multi-index, shape nodes.shape + (D,), first axis fastest."""
    nodes = np.asarray(nodes, dtype=np.int64)
    out = np.empty(nodes.shape + (len(partition),), dtype=np.int64)
    r = nodes.copy()
    for d, p in enumerate(partition):
        out[..., d] = r % (p + 1)
        r //= (p + 1)
    return out


class DiscreteModel:
    """Body-fitted model of one cell type: node coordinates, cell connectivity (1-based Int32) and face labels."""

    def __init__(self, coords, cell_node_ids, ptype, partition=None):
        self.node_coordinates = np.ascontiguousarray(coords, dtype=np.float64)
        self.cell_node_ids = np.ascontiguousarray(cell_node_ids, dtype=np.int32)
        self.ptype = ptype
        self.D = _DIM[ptype]
        self.partition = None if partition is None else tuple(int(p) for p in partition)
        self._faces = {}
        self._face_first = {}
        self._device = {}

    # -- Gridap.Geometry API names
    def num_cells(self):
        return self.cell_node_ids.shape[0]

    def num_nodes(self):
        return self.node_coordinates.shape[0]

    def get_node_coordinates(self):
        return self.node_coordinates

    def get_cell_node_ids(self):
        return self.cell_node_ids

    def celltype(self):
        return _CELLTYPE[self.ptype]

    # -- topology: global d-faces numbered by first touch (src/Geometry/GridTopologies.jl:1184-1251)
    def faces(self, d):
        """returns (cell_to_faces [ncells, nlf] 0-based, face_vertices [nfaces, nv] sorted 0-based node ids)"""
        if d in self._faces:
            return self._faces[d]
        if d == 0:
            res = (self.cell_node_ids - 1, np.arange(self.num_nodes(), dtype=np.int64)[:, None])
        elif d == self.D:
            nc = self.num_cells()
            res = (np.arange(nc, dtype=np.int64)[:, None], np.sort(self.cell_node_ids - 1, axis=1))
        else:
            lf = np.array(local_face_vertices(self.ptype, d))  # [nlf, nv]
            cn = self.cell_node_ids.astype(np.int64) - 1
            fv = np.sort(cn[:, lf], axis=2)  # [nc, nlf, nv]
            nc, nlf, nv = fv.shape
            flat = fv.reshape(nc * nlf, nv)
            nn = self.num_nodes()
            key = np.zeros(nc * nlf, dtype=np.int64) if nv <= 2 else None
            if nv <= 2:
                key = flat[:, 0] * nn + flat[:, 1]
                uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
            else:
                uniq, first, inv = np.unique(flat, axis=0, return_index=True, return_inverse=True)
            order = np.argsort(first, kind="stable")  # first-touch order
            rank = np.empty_like(order)
            rank[order] = np.arange(len(order))
            ids = rank[inv.ravel()].reshape(nc, nlf)
            res = (ids, flat[first[order]])
            self._face_first[d] = first[order]      # (cell * nlf + local face) of the first touch
        self._faces[d] = res
        return res

    def face_frames(self, d):
        """face_vertices [nfaces, nv] 0-based with the vertices of every face in the local order of the face in the FIRST cell that
        holds it (`_face_to_vertices_fill!`, src/Geometry/GridTopologies.jl:1565-1598): the frame in which the own DoFs of the face
        are numbered when the face owns more than one node (order >= 3)."""
        cn = self.cell_node_ids.astype(np.int64) - 1
        if d == 0:
            return np.arange(self.num_nodes(), dtype=np.int64)[:, None]
        if d == self.D:
            return cn
        self.faces(d)
        lf = np.array(local_face_vertices(self.ptype, d))
        first = self._face_first[d]
        return cn[first // len(lf)][np.arange(len(first))[:, None], lf[first % len(lf)]]

    def face_entities(self, d):
        raise NotImplementedError("face labeling is only available for Cartesian models")

    def tag_entities(self, tag):
        raise NotImplementedError

    def device_mesh(self, ctx):
        key = id(ctx)
        if key not in self._device:
            self._device[key] = lib.DeviceMesh(ctx, self.node_coordinates, self.cell_node_ids, self.celltype())
        return self._device[key]


class CartesianDiscreteModel(DiscreteModel):
    """CartesianDiscreteModel(domain, partition): nodes x0 + (I-1)*dx, first axis fastest; cells first axis fastest
    with local nodes first axis fastest (src/Geometry/CartesianGrids.jl:59-70,116-124,156-165); face labeling with one
    entity per box n-face (src/Geometry/CartesianDiscreteModels.jl:133-189)."""

    def __init__(self, domain, partition):
        D = len(partition)
        partition = tuple(int(p) for p in partition)
        x0 = np.array([float(domain[2 * d]) for d in range(D)])
        dx = np.array([(float(domain[2 * d + 1]) - float(domain[2 * d])) / partition[d] for d in range(D)])
        shape = [p + 1 for p in partition]
        axes = [x0[d] + np.arange(shape[d], dtype=np.float64) * dx[d] for d in range(D)]
        grids = np.meshgrid(*axes, indexing="ij")
        coords = np.stack([g.ravel(order="F") for g in grids], axis=1)
        strides = np.cumprod([1] + shape[:-1]).astype(np.int64)
        cidx = np.meshgrid(*[np.arange(p, dtype=np.int64) for p in partition], indexing="ij")
        base = sum(c.ravel(order="F") * s for c, s in zip(cidx, strides))
        loc = np.array([sum(((v >> d) & 1) * strides[d] for d in range(D)) for v in range(2 ** D)], dtype=np.int64)
        cells = (base[:, None] + loc[None, :] + 1).astype(np.int32)
        super().__init__(coords, cells, "HEX" if D == 3 else "QUAD", partition)
        self.domain = tuple(domain)
        self.origin, self.sizes = x0, dx

    def _entity_from_status(self, spans, at_hi):
        """spans/at_hi: bool arrays [n, D] -> 1-based entity id of the minimal box n-face."""
        D = self.D
        faces = ncube_faces(D)
        table = np.zeros((2 ** D, 2 ** D), dtype=np.int32)
        for k, (_, e, a) in enumerate(faces):
            table[e, a] = k + 1
        e = sum((spans[:, d].astype(np.int64) << d) for d in range(D))
        a = sum(((at_hi[:, d] & ~spans[:, d]).astype(np.int64) << d) for d in range(D))
        return table[e, a]

    def face_entities(self, d):
        """entity id (1-based) of every d-face: the box n-face of minimal dimension that contains it."""
        D = self.D
        _, fverts = self.faces(d)
        idx = _cartesian_index(fverts, self.partition)  # [nfaces, nv, D]
        lo, hi = idx.min(axis=1), idx.max(axis=1)
        part = np.array(self.partition, dtype=np.int64)
        spans = (lo != hi) | ((lo > 0) & (lo < part))
        at_hi = (lo == part)
        return self._entity_from_status(spans, at_hi)

    def tag_entities(self, tag):
        nfaces = 3 ** self.D
        if isinstance(tag, (int, np.integer)):
            return [int(tag)]
        if tag == "boundary":
            return list(range(1, nfaces))
        if tag == "interior":
            return [nfaces]
        if isinstance(tag, str) and tag.startswith("tag_"):
            return [int(tag[4:])]
        raise KeyError("unknown tag %r" % (tag,))

    def face_tag_index(self, d, tags):
        """get_face_tag_index(labels,tags,d): position (1-based) of the last tag containing the face, 0 = UNSET."""
        ent = self.face_entities(d)
        out = np.zeros(len(ent), dtype=np.int32)
        for i, tag in enumerate(tags):
            out[np.isin(ent, self.tag_entities(tag))] = i + 1
        return out


class UnstructuredDiscreteModel(DiscreteModel):
    """UnstructuredDiscreteModel(model): same numbering, explicit coordinate / connectivity arrays -- the form in
    which the reference's own benchmark feeds Cartesian meshes to the assembler (benchmark/bm/bm_assembly.jl:29)."""

    def __init__(self, model):
        super().__init__(model.node_coordinates, model.cell_node_ids, model.ptype, model.partition)
        self._parent = model

    def face_entities(self, d):
        return self._parent.face_entities(d)

    def tag_entities(self, tag):
        return self._parent.tag_entities(tag)

    def face_tag_index(self, d, tags):
        ent = self.face_entities(d)
        out = np.zeros(len(ent), dtype=np.int32)
        for i, tag in enumerate(tags):
            out[np.isin(ent, self.tag_entities(tag))] = i + 1
        return out


class _SimplexifiedModel(UnstructuredDiscreteModel):
    def __init__(self, model):
        table = _HEX_TO_TETS if model.ptype == "HEX" else _QUAD_TO_TRIS
        cells = model.cell_node_ids[:, table].reshape(-1, table.shape[1])
        DiscreteModel.__init__(self, model.node_coordinates, cells, "TET" if model.ptype == "HEX" else "TRI", model.partition)
        self._parent = model
        self._cart = model if isinstance(model, CartesianDiscreteModel) else getattr(model, "_parent", None)

    def face_entities(self, d):
        # minimal box face containing the face, from the Cartesian index of its vertices
        cart = self._cart
        _, fverts = self.faces(d)
        idx = _cartesian_index(fverts, cart.partition)
        lo, hi = idx.min(axis=1), idx.max(axis=1)
        part = np.array(cart.partition, dtype=np.int64)
        spans = (lo != hi) | ((lo > 0) & (lo < part))
        at_hi = (lo == part)
        return cart._entity_from_status(spans, at_hi)

    def tag_entities(self, tag):
        return self._cart.tag_entities(tag)


def simplexify(model):
    """simplexify(model): each hex -> 6 tets, cell 6(h-1)+t (src/Geometry/Grids.jl:487-530)."""
    return _SimplexifiedModel(model)


class Triangulation:
    """Triangulation(model): the body-fitted bulk triangulation; Triangulation(model, cell_ids) / Triangulation(model, mask): the view
    on a subset of its cells (`Triangulation(model, cell_to_parent_cell)`, src/Geometry/Triangulations.jl, BodyFittedTriangulation with
    a `tface_to_mface` glue: the reference's own assembly benchmark runs every case on the bulk and on such a view,
    benchmark/bm/bm_assembly.jl:30-33).  `cell_ids` are Gridap's 1-based cell ids, in the order given."""

    def __init__(self, model, cells=None):
        self.cells = None
        if cells is None:
            self.model = model
            return
        cells = np.asarray(cells)
        if cells.dtype == bool:
            if len(cells) != model.num_cells():
                raise ValueError("cell mask of length %d on a model of %d cells" % (len(cells), model.num_cells()))
            cells = np.nonzero(cells)[0]
        else:
            cells = cells.astype(np.int64) - 1
            if len(cells) and (cells.min() < 0 or cells.max() >= model.num_cells()):
                raise ValueError("cell ids out of range (1-based ids of the model's cells are expected)")
        self.cells = cells
        self.parent = model
        self.model = DiscreteModel(model.node_coordinates, model.cell_node_ids[cells], model.ptype)
        self._spaces = {}

    def num_cells(self):
        return self.model.num_cells()

    def restrict(self, space):
        """the cell-wise DoF ids of `space` on the cells of the view (same global numbering)"""
        key = id(space)
        if key not in self._spaces:
            if space.model is not self.parent and getattr(space.model, "_parent", None) is not self.parent \
                    and getattr(self.parent, "_parent", None) is not space.model:
                raise ValueError("the FE space lives on another model than the Triangulation")
            self._spaces[key] = (_FacetSpace(space, self, space.get_cell_dof_ids()[self.cells]), space)   # (the space stays alive: id())
        return self._spaces[key][0]


def get_triangulation(model):
    return Triangulation(model)


class _FacetSpace:
    """An FE space restricted to the facets of a BoundaryTriangulation: the trace of a Lagrangian basis on a facet is the facet's
    own Lagrangian basis on the DoFs that lie on it (the other shape functions of the cell vanish there), so the facet's local
    vector / matrix only carries those DoFs.  Gridap keeps the zero rows of the off-facet DoFs of the adjacent cell
    (FaceToCellGlue, src/Geometry/BoundaryTriangulations.jl:13-70); adding zeros changes nothing."""

    def __init__(self, space, trian, cell_dof_ids):
        self.model, self.reffe, self.ncomp, self.order = trian.model, space.reffe, space.ncomp, space.reffe.order
        self.cell_dof_ids = np.ascontiguousarray(cell_dof_ids, dtype=np.int32)
        self.nfree, self.ndirichlet = space.nfree, space.ndirichlet
        self._device = {}

    def num_free_dofs(self):
        return self.nfree

    def num_dirichlet_dofs(self):
        return self.ndirichlet

    def get_cell_dof_ids(self):
        return self.cell_dof_ids

    def device_space(self, ctx, refel_key, refel, ids=None):
        key = (id(ctx), refel_key)
        if key not in self._device:
            self._device[key] = lib.DeviceSpace(ctx, self.model.device_mesh(ctx), refel, self.cell_dof_ids, self.nfree, self.ndirichlet)
        return self._device[key]


class BoundaryTriangulation(Triangulation):
    """BoundaryTriangulation(model; tags) (src/Geometry/BoundaryTriangulations.jl:152-203): the boundary facets of the model (those
    with one incident cell), optionally only those labelled with `tags`, in ascending facet id; facet nodes in the local order of
    the adjacent cell's face (tensor order for n-cube facets).  HEX -> QUAD, TET -> TRI, QUAD / TRI -> SEG facets."""

    def __init__(self, model, tags=None):
        facet_ptype = {"HEX": "QUAD", "QUAD": "SEG", "TET": "TRI", "TRI": "SEG"}[model.ptype]
        D = model.D
        c2f, fverts = model.faces(D - 1)
        nf = len(fverts)
        sel = np.bincount(c2f.ravel(), minlength=nf) == 1
        if tags is not None:
            taglist = list(tags) if isinstance(tags, (list, tuple)) else [tags]
            sel &= model.face_tag_index(D - 1, taglist) > 0
        nc, nlf = c2f.shape
        flat = c2f.ravel()
        first = np.full(nf, nc * nlf, dtype=np.int64)
        np.minimum.at(first, flat, np.arange(nc * nlf, dtype=np.int64))   # first (cell, local face) touching every facet
        self.face_ids = np.nonzero(sel)[0]
        self.cells = first[self.face_ids] // nlf
        self.lfaces = first[self.face_ids] % nlf
        lf = np.array(local_face_vertices(model.ptype, D - 1))             # [nlf, nv] local vertices, ascending = tensor order
        face_nodes = model.cell_node_ids[self.cells[:, None], lf[self.lfaces]]
        self.parent = model
        self.model = DiscreteModel(model.node_coordinates, face_nodes, facet_ptype)
        self._spaces = {}

    def num_cells(self):
        return len(self.face_ids)

    # -- FaceToCellGlue (src/Geometry/BoundaryTriangulations.jl:13-70): the cell adjacent to every facet and the facet's local face
    def glue_model(self):
        """the mesh of the cells ADJACENT to the facets (one "cell" per facet): what a term with normals / cell-basis gradients is
        integrated on (facet quadrature mapped into the cell's reference space)"""
        if getattr(self, "_glue_model", None) is None:
            m = self.parent
            self._glue_model = DiscreteModel(m.node_coordinates, m.cell_node_ids[self.cells], m.ptype)
            self._glue_spaces = {}
        return self._glue_model

    def glue_space(self, space):
        """the cell DoF tables of the adjacent cells (all DoFs of the cell: the gradient of every cell shape function is seen on
        the facet)"""
        gm = self.glue_model()
        key = id(space)
        hit = self._glue_spaces.get(key)
        if hit is not None and hit[1] is space:
            return hit[0]
        if space.model is not self.parent and getattr(space.model, "_parent", None) is not self.parent:
            raise ValueError("the FE space lives on another model than the BoundaryTriangulation")
        fs = _FacetSpace(space, self, space.get_cell_dof_ids()[self.cells])
        fs.model = gm
        self._glue_spaces[key] = (fs, space)
        return fs

    def restrict(self, space):
        """cell-wise (facet-wise) DoF ids of `space` on the facets, component-major, facet-local Lagrangian node order
        (vertices, then edges in the facet's local edge order, then the facet interior)."""
        key = id(space)
        if key in self._spaces:
            return self._spaces[key]
        if space.model is not self.parent and getattr(space.model, "_parent", None) is not self.parent:
            raise ValueError("the FE space lives on another model than the BoundaryTriangulation")
        if space.order > 2:
            self._spaces[key] = _FacetSpace(space, self, self._restrict_through_cells(space))
            return self._spaces[key]
        fn = self.model.cell_node_ids.astype(np.int64) - 1               # [nfacets, nv]
        ent = space._entity_ids                                          # [entities, ncomp]: vertices (| edges | faces | cells)
        cols = [ent[fn]]                                                 # vertices: [nfacets, nv, ncomp]
        if space.order == 2:
            m, D = self.parent, self.parent.D
            nn = m.num_nodes()
            ofs = nn
            if D == 3:   # edges of the QUAD facet: local edges of the 2-cube in the facet's tensor order
                _, everts = m.faces(1)
                ekey = everts[:, 0] * nn + everts[:, 1]
                order = np.argsort(ekey)
                le = np.array(local_face_vertices(self.model.ptype, 1))  # QUAD: [[0,1],[2,3],[0,2],[1,3]]; TRI: [[0,1],[0,2],[1,2]]
                pair = np.sort(fn[:, le], axis=2)                        # [nfacets, 4, 2]
                eid = order[np.searchsorted(ekey[order], pair[..., 0] * nn + pair[..., 1])]
                cols.append(ent[ofs + eid])
                ofs += len(everts)
            if self.model.ptype != "TRI":                                # (a P2 triangle has no interior node)
                cols.append(ent[ofs + self.face_ids][:, None, :])        # the facet's own interior node
        allc = np.concatenate(cols, axis=1)                              # [nfacets, nl, ncomp]
        nfac, nl, nc_ = allc.shape
        ids = np.transpose(allc, (0, 2, 1)).reshape(nfac, nl * nc_)
        self._spaces[key] = _FacetSpace(space, self, ids)
        return self._spaces[key]


BoundaryTriangulation._restrict_through_cells = lambda self, space: _facet_ids_through_cells(self, space)


def _facet_ids_through_cells(trian, space):
    """facet-wise DoF ids of a Lagrangian space of any order, read from the adjacent cell: node j of the facet's own Lagrangian element
    (lattice of the facet polytope, facet vertices = the vertices of the cell's local face in their local order) is the node of the
    cell's element with the same position (the trace of the cell basis on the facet is the facet basis on those nodes)"""
    from . import reffes as rf
    m = trian.parent
    k, ncomp = space.order, space.ncomp
    D = m.D
    simplex = m.ptype in ("TET", "TRI")
    clat, _ = rf.lagrangian_lattice(m.ptype, k)
    nl = len(clat)
    where = {tuple(int(x) for x in row): a for a, row in enumerate(clat)}
    flat, _ = rf.lagrangian_lattice(trian.model.ptype, k)                # [nfn, D-1]
    lfv = local_face_vertices(m.ptype, D - 1)
    node_map = np.zeros((len(lfv), len(flat)), dtype=np.int64)
    for lf, vs in enumerate(lfv):
        for j, fl in enumerate(flat):
            if simplex:
                fb = [k - int(fl.sum())] + [int(x) for x in fl]           # barycentric lattice on the facet's vertices
                cb = [0] * (D + 1)
                for w, v in zip(fb, vs):
                    cb[v] = w
                idx = cb[1:]
            else:
                axes = [ax for ax in range(D) if len({(v >> ax) & 1 for v in vs}) == 2]
                idx = [k * ((vs[0] >> ax) & 1) for ax in range(D)]
                for t, ax in enumerate(axes):
                    idx[ax] = int(fl[t])
            node_map[lf, j] = where[tuple(idx)]
    cell_ids = space.get_cell_dof_ids()[trian.cells].reshape(len(trian.cells), ncomp, nl)
    nodes = node_map[trian.lfaces]                                          # [nfacets, nfn]
    ids = np.take_along_axis(cell_ids, np.broadcast_to(nodes[:, None, :], (len(nodes), ncomp, nodes.shape[1])), axis=2)
    return ids.reshape(len(nodes), ncomp * nodes.shape[1])


Boundary = BoundaryTriangulation


class SkeletonTriangulation(Triangulation):
    """SkeletonTriangulation(model) (src/Geometry/SkeletonTriangulations.jl:7-32,54-99): the interior facets of the model (two
    incident cells) in ascending facet id; plus = the first incident cell (lower cell id), minus = the second -- the order of
    `get_faces(topo, D-1, D)` the reference takes its `SkeletonPair` from.  Facet nodes in the local order of the plus cell's face."""

    def __init__(self, model):
        facet_ptype = {"HEX": "QUAD", "QUAD": "SEG", "TET": "TRI", "TRI": "SEG"}[model.ptype]
        D = model.D
        c2f, fverts = model.faces(D - 1)
        nf = len(fverts)
        nc, nlf = c2f.shape
        flat = c2f.ravel()
        pos = np.arange(nc * nlf, dtype=np.int64)
        first = np.full(nf, nc * nlf, dtype=np.int64)
        last = np.full(nf, -1, dtype=np.int64)
        np.minimum.at(first, flat, pos)
        np.maximum.at(last, flat, pos)
        self.face_ids = np.nonzero(np.bincount(flat, minlength=nf) == 2)[0]
        self.cells_plus, self.lfaces_plus = first[self.face_ids] // nlf, first[self.face_ids] % nlf
        self.cells_minus, self.lfaces_minus = last[self.face_ids] // nlf, last[self.face_ids] % nlf
        lf = np.array(local_face_vertices(model.ptype, D - 1))
        face_nodes = model.cell_node_ids[self.cells_plus[:, None], lf[self.lfaces_plus]]
        self.parent = model
        self.model = DiscreteModel(model.node_coordinates, face_nodes, facet_ptype)
        self._glue = {}

    def num_cells(self):
        return len(self.face_ids)

    def glue_model(self, side):
        """the mesh of the plus / minus cells of the facets (one "cell" per facet)"""
        key = ("model", side)
        if key not in self._glue:
            m = self.parent
            cells = self.cells_plus if side == "plus" else self.cells_minus
            self._glue[key] = DiscreteModel(m.node_coordinates, m.cell_node_ids[cells], m.ptype)
        return self._glue[key]

    def glue_space(self, space, side):
        key = ("space", side, id(space))
        hit = self._glue.get(key)
        if hit is not None and hit[1] is space:
            return hit[0]
        if space.model is not self.parent:
            raise ValueError("the FE space lives on another model than the SkeletonTriangulation")
        cells = self.cells_plus if side == "plus" else self.cells_minus
        fs = _FacetSpace(space, self, space.get_cell_dof_ids()[cells])
        fs.model = self.glue_model(side)
        self._glue[key] = (fs, space)
        return fs

    def point_permutation(self, pts):
        """pts [nlf, npf, D]: the facet rule on every local face of the reference cell.  -> perm [nfacets, npf]: the point of the
        minus cell's local-face block that coincides (physically) with point p of the plus side -- what the vertex permutation of
        FaceToCellGlue (cell_to_lface_to_pindex, src/Geometry/BoundaryTriangulations.jl:42-70) does to the quadrature points."""
        from . import reffes as rf
        m = self.parent
        nlf, npf, D = pts.shape
        Ng, _ = rf.tabulate_lagrangian(m.ptype, 1, pts.reshape(-1, D))
        Ng = Ng.reshape(nlf, npf, -1)
        X = m.node_coordinates
        xp = np.einsum("fpa,fad->fpd", Ng[self.lfaces_plus], X[m.cell_node_ids[self.cells_plus].astype(np.int64) - 1])
        xm = np.einsum("fpa,fad->fpd", Ng[self.lfaces_minus], X[m.cell_node_ids[self.cells_minus].astype(np.int64) - 1])
        d2 = ((xp[:, :, None, :] - xm[:, None, :, :]) ** 2).sum(axis=3)
        perm = d2.argmin(axis=2)
        scale = ((xp.max(axis=1) - xp.min(axis=1)) ** 2).sum(axis=1).max() if npf > 1 else 1.0
        if len(perm) and d2.min(axis=2).max() > 1e-16 * max(scale, 1e-300) + 1e-24:
            raise ValueError("the facet quadrature points of the plus and minus cells do not coincide (non-conforming mesh?)")
        return np.ascontiguousarray(perm, dtype=np.int32)


Skeleton = SkeletonTriangulation


class NormalVector:
    """get_normal_vector(trian) (src/Geometry/BoundaryTriangulations.jl:244-283): the outward unit normal of the facets, as a symbol of
    the weak-form language (evaluated on the device at the facet quadrature points)."""

    def __init__(self, trian):
        if not isinstance(trian, (BoundaryTriangulation, SkeletonTriangulation)):
            raise NotImplementedError("get_normal_vector: BoundaryTriangulation / SkeletonTriangulation")
        self.trian = trian   # (skeleton: the plus normal n+; n- = -n+, jump(v n) = (v+ - v-) n+)


def get_normal_vector(trian):
    return NormalVector(trian)
