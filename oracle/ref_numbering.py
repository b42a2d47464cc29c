"""TEST INFRASTRUCTURE ONLY -- oracle restatement of Gridap's mesh / DoF numbering.

Loop-by-loop restatement (1-based ids kept, python loops, small meshes only) of:
  * Cartesian nodes / cells      src/Geometry/CartesianGrids.jl:59-70,116-124,156-165
  * n-cube / simplex local faces src/ReferenceFEs/ExtrusionPolytopes.jl:460-531
  * simplexify                   src/ReferenceFEs/ExtrusionPolytopes.jl:290-299, src/Geometry/Grids.jl:487-530
  * global face numbering        src/Geometry/GridTopologies.jl:1184-1251 (first touch)
  * Cartesian face labeling      src/Geometry/CartesianDiscreteModels.jl:133-267
  * CLagrangian DoFs             src/FESpaces/CLagrangianFESpaces.jl:155-288,356-380
  * conforming (face-based) DoFs src/FESpaces/ConformingFESpaces.jl:367-423,543-636,823-864; any order: own-node
    permutations src/ReferenceFEs/CLagrangianRefFEs.jl:549-577, cell permutation indices src/Geometry/GridTopologies.jl:515-690
  * multi-field offsets          src/MultiField/MultiFieldFESpaces.jl:356-364,482-488
The product has its own vectorised generators (gridap.jl_b200/geometry.py, fespaces.py);
tests compare the two.
"""
import itertools
import numpy as np

UNSET = 0


# ----------------------------------------------------------------------------- Cartesian grid
def cartesian_descriptor(domain, partition):
    """CartesianDescriptor(domain,partition): origin, sizes (CartesianGrids.jl:59-70)."""
    D = len(partition)
    origin = [float(domain[2 * d]) for d in range(D)]
    sizes = [(float(domain[2 * d + 1]) - float(domain[2 * d])) / partition[d] for d in range(D)]
    return origin, sizes


def cartesian_node_coordinates(domain, partition):
    """x[node] = x0 + (I-1)*dx, node = LinearIndices(partition.+1)[I], first axis fastest
    (CartesianGrids.jl:116-124)."""
    D = len(partition)
    x0, dx = cartesian_descriptor(domain, partition)
    shape = [p + 1 for p in partition]
    n = int(np.prod(shape))
    X = np.zeros((n, D))
    for node, I in enumerate(itertools.product(*[range(1, s + 1) for s in reversed(shape)])):
        I = I[::-1]  # first axis fastest
        for d in range(D):
            X[node, d] = x0[d] + (I[d] - 1) * dx[d]
    return X


def cartesian_cell_node_ids(partition):
    """cell -> 2^D node ids (1-based), cells first axis fastest, local nodes first axis fastest
    (CartesianGrids.jl:156-165)."""
    D = len(partition)
    shape = [p + 1 for p in partition]
    strides = [int(np.prod(shape[:d])) for d in range(D)]
    cells = []
    for ci in itertools.product(*[range(1, p + 1) for p in reversed(partition)]):
        ci = ci[::-1]
        v = []
        for ln in itertools.product(*[range(1, 3) for _ in range(D)]):
            ln = ln[::-1]
            k = [ln[d] + ci[d] - 1 for d in range(D)]
            v.append(1 + sum((k[d] - 1) * strides[d] for d in range(D)))
        cells.append(v)
    return np.array(cells, dtype=np.int32)


# ----------------------------------------------------------------------------- polytopes
def ncube_faces(D):
    """Local n-faces of the D-cube as (dim, extrusion bits, anchor bits), in Gridap's order:
    stable sort by dimension, then extrusion, then anchor, last axis most significant
    (ExtrusionPolytopes.jl:460-474).  Pinned by QUAD edges [1,2],[3,4],[1,3],[2,4]
    (test/ReferenceFEsTests/ExtrusionPolytopesTests.jl:17)."""
    faces = []
    for e in range(2 ** D):
        for a in range(2 ** D):
            if a & e:
                continue
            faces.append((bin(e).count("1"), e, a))
    faces.sort()
    return faces


def ncube_face_vertices(D, d):
    """Vertex ids (1-based, x fastest) of each local d-face of the D-cube."""
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
            vs.append(v + 1)
        out.append(sorted(vs))
    return out


# TET local faces (ExtrusionPolytopes.jl:460-531, run for extrusion (1,2,2)); see SURVEY App. B
TET_VERTS = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)]
TET_EDGES = [[1, 2], [1, 3], [2, 3], [1, 4], [2, 4], [3, 4]]
TET_FACES = [[1, 2, 3], [1, 2, 4], [1, 3, 4], [2, 3, 4]]
TRI_EDGES = [[1, 2], [1, 3], [2, 3]]
# simplexify(HEX): 6 tets per hex in this order (ExtrusionPolytopes.jl:290-299)
HEX_TO_TETS = [[1, 2, 3, 7], [1, 2, 5, 7], [2, 3, 4, 7], [2, 4, 7, 8], [2, 5, 6, 7], [2, 6, 7, 8]]
QUAD_TO_TRIS = [[1, 2, 3], [2, 3, 4]]


def local_face_vertices(ptype, d):
    if ptype == "HEX":
        return ncube_face_vertices(3, d)
    if ptype == "QUAD":
        return ncube_face_vertices(2, d)
    if ptype == "TET":
        return {0: [[1], [2], [3], [4]], 1: TET_EDGES, 2: TET_FACES, 3: [[1, 2, 3, 4]]}[d]
    if ptype == "TRI":
        return {0: [[1], [2], [3]], 1: TRI_EDGES, 2: [[1, 2, 3]]}[d]
    raise ValueError(ptype)


def simplexify(cell_nodes, ptype):
    """cell 6(h-1)+t (Grids.jl:487-530)."""
    table = HEX_TO_TETS if ptype == "HEX" else QUAD_TO_TRIS
    out = []
    for nodes in cell_nodes:
        for lt in table:
            out.append([nodes[k - 1] for k in lt])
    return np.array(out, dtype=np.int32)


# ----------------------------------------------------------------------------- topology
def global_faces(cell_nodes, ptype, d):
    """cell -> global d-face ids, numbered by first touch sweeping cells then local faces
    (GridTopologies.jl:1184-1251).  Returns (cell_to_faces[ncells][nlf], face_to_vertices)."""
    if d == 0:
        # vertices keep their node ids (vertex_to_node is the identity for grids with num_nodes == num_vertices,
        # src/Geometry/UnstructuredGridTopologies.jl:145-149); only edges / faces are generated by first touch
        nn = int(np.max(cell_nodes))
        return np.array(cell_nodes, dtype=np.int32), [(v,) for v in range(1, nn + 1)]
    lfaces = local_face_vertices(ptype, d)
    seen = {}
    face_vertices = []
    cell_faces = []
    for nodes in cell_nodes:
        row = []
        for lf in lfaces:
            key = tuple(sorted(int(nodes[k - 1]) for k in lf))
            if key not in seen:
                seen[key] = len(seen) + 1
                face_vertices.append(key)
            row.append(seen[key])
        cell_faces.append(row)
    return np.array(cell_faces, dtype=np.int32), face_vertices


# ----------------------------------------------------------------------------- labels
def cartesian_entity_of_vertices(partition, vertex_ids):
    """Entity id of the box n-face of minimal dimension containing a mesh face given by its
    vertex node ids (geometric statement of _fill_cartesian_entities!,
    CartesianDiscreteModels.jl:140-267; pinned by num_entities==27 and the 2-D tag tests)."""
    D = len(partition)
    shape = [p + 1 for p in partition]
    faces = ncube_faces(D)
    e = 0
    a = 0
    idx = []
    for v in vertex_ids:
        r = v - 1
        I = []
        for d in range(D):
            I.append(r % shape[d])
            r //= shape[d]
        idx.append(I)
    for d in range(D):
        vals = set(I[d] for I in idx)
        if len(vals) > 1:
            e |= 1 << d  # the face spans this axis
        else:
            c = vals.pop()
            if c == 0:
                pass
            elif c == partition[d]:
                a |= 1 << d
            else:
                e |= 1 << d  # interior along this axis
    dim = bin(e).count("1")
    return faces.index((dim, e, a)) + 1


def cartesian_tag_entities(D, tag):
    """tag name / int -> list of entity ids (CartesianDiscreteModels.jl:178-189)."""
    nfaces = 3 ** D
    if isinstance(tag, (int, np.integer)):
        return [int(tag)]
    if tag == "boundary":
        return list(range(1, nfaces))
    if tag == "interior":
        return [nfaces]
    if tag.startswith("tag_"):
        return [int(tag[4:])]
    raise KeyError(tag)


def face_tag_index(entities, D, tags):
    """get_face_tag_index(labels,tags,d): for each face the LAST position in `tags` whose tag
    contains the face's entity, else UNSET (src/Geometry/FaceLabelings.jl get_face_tag_index)."""
    out = []
    for ent in entities:
        idx = UNSET
        for i, tag in enumerate(tags):
            if ent in cartesian_tag_entities(D, tag):
                idx = i + 1
        out.append(idx)
    return out


# ----------------------------------------------------------------------------- CLagrangian DoFs
def clagrangian_dofs(node_to_tag, tag_to_masks, ncomp):
    """node sweep, components interleaved per node; free -> +(++nfree), Dirichlet -> -(++ndiri)
    (CLagrangianFESpaces.jl:155-288).  Returns node_and_comp_to_dof [nnodes][ncomp], nfree, ndiri,
    dirichlet_dof_to_node, dirichlet_dof_to_comp."""
    nfree = 0
    ndiri = 0
    out = []
    d2n, d2c = [], []
    for node, tag in enumerate(node_to_tag, start=1):
        m = []
        for comp in range(ncomp):
            if tag == UNSET:
                isdiri = False
            else:
                masks = tag_to_masks[tag - 1]
                isdiri = bool(masks[comp]) if ncomp > 1 or isinstance(masks, (list, tuple)) else bool(masks)
            if isdiri:
                ndiri += 1
                m.append(-ndiri)
                d2n.append(node)
                d2c.append(comp + 1)
            else:
                nfree += 1
                m.append(nfree)
        out.append(m)
    return np.array(out, dtype=np.int32), nfree, ndiri, d2n, d2c


def clagrangian_cell_dofs(cell_nodes, node_and_comp_to_dof):
    """local dof (lnode, comp) -> k = lnode + nlnodes*(comp-1) (component-major)
    (CLagrangianFESpaces.jl:356-380, LagrangianDofBases.jl:77-96)."""
    nc, nl = cell_nodes.shape
    ncomp = node_and_comp_to_dof.shape[1]
    out = np.zeros((nc, nl * ncomp), dtype=np.int32)
    for c in range(nc):
        for comp in range(ncomp):
            for ln in range(nl):
                out[c, ln + nl * comp] = node_and_comp_to_dof[cell_nodes[c, ln] - 1, comp]
    return out


# ----------------------------------------------------------------------------- conforming DoFs (order 2)
def conforming_dofs_order2(cell_nodes, ptype, ncomp, dface_to_tag, tag_to_masks):
    """Face-based numbering for Lagrangian order-2 spaces (Q2 on n-cubes, P2 on simplices):
    sweep d = 0..D, faces by id, own DoFs of a face component-major (one node per face for
    order 2), split free/Dirichlet in the same sweep (ConformingFESpaces.jl:543-636);
    cell ids via CellDofsNonOriented (:844-864).  `dface_to_tag[d]` = tag index per d-face
    (UNSET = free).  Returns cell_dofs [ncells][nlnodes*ncomp], nfree, ndiri, and for each
    d the (cell_to_faces, face_vertices)."""
    D = {"HEX": 3, "QUAD": 2, "TET": 3, "TRI": 2}[ptype]
    simplex = ptype in ("TET", "TRI")
    dims_with_nodes = [0, 1] if simplex else list(range(D + 1))
    topo = {d: global_faces(cell_nodes, ptype, d) for d in dims_with_nodes}
    nfree = 0
    ndiri = 0
    face_own = {}
    for d in dims_with_nodes:
        nfaces = len(topo[d][1])
        tags = dface_to_tag.get(d) if dface_to_tag is not None else None
        own = np.zeros((nfaces, ncomp), dtype=np.int64)
        for f in range(nfaces):
            tag = UNSET if (tags is None or d == D) else tags[f]
            for comp in range(ncomp):
                if tag == UNSET:
                    isdiri = False
                else:
                    masks = tag_to_masks[tag - 1]
                    isdiri = bool(masks[comp]) if isinstance(masks, (list, tuple, np.ndarray)) else bool(masks)
                if isdiri:
                    ndiri += 1
                    own[f, comp] = -ndiri
                else:
                    nfree += 1
                    own[f, comp] = nfree
        face_own[d] = own
    nl = sum(len(local_face_vertices(ptype, d)) for d in dims_with_nodes)
    nc = len(cell_nodes)
    cell_dofs = np.zeros((nc, nl * ncomp), dtype=np.int32)
    for c in range(nc):
        ln = 0
        for d in dims_with_nodes:
            for f in topo[d][0][c]:
                for comp in range(ncomp):
                    cell_dofs[c, ln + nl * comp] = face_own[d][f - 1, comp]
                ln += 1
    return cell_dofs, nfree, ndiri, topo


# ----------------------------------------------------------------------------- conforming DoFs (any order)
def vertex_permutations(ptype):
    """get_vertex_permutations(p) (ExtrusionPolytopes.jl:793-858): dimension 0 / 1 and simplices of dimension 2 -> all the
    permutations in lexicographic order (Combinatorics.permutations); the square -> those lexicographic permutations whose
    auxiliary Jacobian sum_i grad_i (x) x_perm(i) has the |determinant| of the identity permutation; dimension 3 -> identity only.
    Pinned by test/ReferenceFEsTests/ExtrusionPolytopesTests.jl:33-36,58-75."""
    if ptype == "VERTEX":
        return [[1]]
    if ptype == "SEG":
        return [[1, 2], [2, 1]]
    if ptype == "TRI":
        return [list(p) for p in itertools.permutations([1, 2, 3])]
    if ptype == "QUAD":
        anchors = [[(v >> d) & 1 for d in range(2)] for v in range(4)]
        grads = [[-1 if x[di] == 0 else 1 for di in range(2)] for x in anchors]      # _setup_aux_grads (:860-877)
        out, vol = [], -1
        for perm in itertools.permutations([1, 2, 3, 4]):
            m = [[0, 0], [0, 0]]
            for i, cj in enumerate(perm):                                             # _setup_aux_jacobian (:879-893)
                x = anchors[cj - 1]
                for di in range(2):
                    for dj in range(2):
                        m[di][dj] += x[dj] * grads[i][di]
            vol_i = abs(m[0][0] * m[1][1] - m[0][1] * m[1][0])
            if vol < 0:
                vol = vol_i
            if vol_i == vol:
                out.append(list(perm))
        return out
    if ptype in ("HEX", "TET"):
        return [list(range(1, 9 if ptype == "HEX" else 5))]
    raise ValueError(ptype)


def own_nodes_permutations(ptype, interior_nodes, linear_shapefuns):
    """_compute_node_permutations(p, interior_nodes) (CLagrangianRefFEs.jl:549-577): for every vertex permutation the own nodes
    are mapped by the linear shape functions onto the polytope with permuted vertex coordinates, pvertex_to_coord[perm[v]] =
    vertex_to_coord[v]; node_to_pnode[node] = the first mapped node that coincides with `node` (0 = INVALID_PERM).
    `linear_shapefuns(points)` -> [npoints][nvertices] (the caller passes the tabulation restatement)."""
    from . import ref_tabulation as rt
    interior_nodes = np.asarray(interior_nodes, dtype=float)
    if ptype == "VERTEX":
        return [[1]]
    vertex_to_coord = rt.vertex_coordinates(ptype)
    if len(interior_nodes) == 0:
        return [[] for _ in vertex_permutations(ptype)]
    shp = linear_shapefuns(interior_nodes)
    out = []
    for perm in vertex_permutations(ptype):
        pcoord = np.zeros_like(vertex_to_coord)
        for v, pv in enumerate(perm):
            pcoord[pv - 1] = vertex_to_coord[v]
        pnodes = shp @ pcoord
        node_to_pnode = [0] * len(interior_nodes)
        for node, x in enumerate(interior_nodes):
            for i, y in enumerate(pnodes):
                if np.linalg.norm(y - x) < 1.0e-10:
                    node_to_pnode[node] = i + 1
                    break
        out.append(node_to_pnode)
    return out


def global_faces_oriented(cell_nodes, ptype, d):
    """(cell_to_faces, face_to_vertices) with the vertices of every face in the local order of the face in the FIRST cell that
    holds it (_face_to_vertices_fill!, GridTopologies.jl:1565-1598) -- the frame the own DoFs of the face are numbered in."""
    c2f, fv_sorted = global_faces(cell_nodes, ptype, d)
    if d == 0:
        return c2f, [list(v) for v in fv_sorted]
    lfaces = local_face_vertices(ptype, d)
    fverts = [None] * len(fv_sorted)
    for c, nodes in enumerate(cell_nodes):
        for lf, lv in enumerate(lfaces):
            f = c2f[c][lf] - 1
            if fverts[f] is None:
                fverts[f] = [int(nodes[k - 1]) for k in lv]
    return c2f, fverts


def cell_permutations(cell_nodes, ptype, d, c2f, fverts):
    """compute_cell_permutations(top, d) (GridTopologies.jl:515-549,649-690): pindex[cell][lface] = the first vertex permutation
    `cfvertex_to_fvertex` of the face's polytope with face_vertices[perm[cfvertex]] == cell_vertices[lface_vertices[cfvertex]]
    for all cfvertex; 1 for d = 0 and d = D."""
    D = {"HEX": 3, "QUAD": 2, "TET": 3, "TRI": 2}[ptype]
    nc = len(cell_nodes)
    lfaces = local_face_vertices(ptype, d) if d < D else [None]
    out = np.ones((nc, len(lfaces)), dtype=np.int32)
    if d == 0 or d == D:
        return out
    from .ref_tabulation import _FACE_PTYPE
    perms = vertex_permutations(_FACE_PTYPE[(ptype, d)])
    for c in range(nc):
        for lf, lv in enumerate(lfaces):
            fv = fverts[c2f[c][lf] - 1]
            found = False
            for pindex, perm in enumerate(perms):
                if all(fv[perm[k] - 1] == cell_nodes[c][lv[k] - 1] for k in range(len(lv))):
                    out[c, lf] = pindex + 1
                    found = True
                    break
            assert found, "Valid pindex not found"
    return out


def conforming_dofs(cell_nodes, ptype, order, ncomp, dface_to_tag, tag_to_masks):
    """compute_conforming_cell_dofs for a Lagrangian space of any order (ConformingFESpaces.jl:367-423,543-636,823-864):
      * the own DoFs of a face = own nodes of the face, component-major / node-minor (_generate_face_own_dofs,
        LagrangianRefFEs.jl:254-272), numbered in the sweep d = 0..D, faces by id, free / Dirichlet per component of the tag
        (_split_face_own_dofs_into_free_and_dirichlet_with_components!);
      * a cell reads the own DoFs of its faces through the permutation of its local face against the face's own frame:
        dofs[own_ldofs[i]] = face_own_dofs[face][pdofs[i]], pdofs = own-DoF permutation number pindex
        (CellDofsNonOriented.getindex!, :844-864; _generate_face_own_dofs_permutations, LagrangianRefFEs.jl:283-317).
    Returns cell_dofs [ncells][nlnodes*ncomp] (local DoF = lnode + nlnodes*comp), nfree, ndiri."""
    from . import ref_tabulation as rt
    D = {"HEX": 3, "QUAD": 2, "TET": 3, "TRI": 2}[ptype]
    nodes, face_own = rt.lagrangian_nodes_and_face_own_nodes(ptype, order)
    nl = len(nodes)
    offs = [0]
    for d in range(D + 1):
        offs.append(offs[-1] + (1 if d == D else len(local_face_vertices(ptype, d))))
    nfree = ndiri = 0
    d_own, d_topo, d_pidx, d_perms = {}, {}, {}, {}
    for d in range(D + 1):
        lf_own = face_own[offs[d]:offs[d + 1]]          # own nodes (1-based) of every local d-face
        nown = len(lf_own[0])
        if nown == 0:
            continue
        if d == D:
            c2f = np.arange(1, len(cell_nodes) + 1).reshape(-1, 1)
            fverts = [list(map(int, n)) for n in cell_nodes]
        else:
            c2f, fverts = global_faces_oriented(cell_nodes, ptype, d)
        tags = dface_to_tag.get(d) if (dface_to_tag is not None and d < D) else None
        own = np.zeros((len(fverts), ncomp * nown), dtype=np.int64)
        for f in range(len(fverts)):
            tag = UNSET if tags is None else tags[f]
            for comp in range(ncomp):
                for j in range(nown):
                    if tag == UNSET:
                        isdiri = False
                    else:
                        masks = tag_to_masks[tag - 1]
                        isdiri = bool(masks[comp]) if isinstance(masks, (list, tuple, np.ndarray)) else bool(masks)
                    if isdiri:
                        ndiri += 1
                        own[f, comp * nown + j] = -ndiri
                    else:
                        nfree += 1
                        own[f, comp * nown + j] = nfree
        d_own[d], d_topo[d] = own, (c2f, fverts)
        d_pidx[d] = cell_permutations(cell_nodes, ptype, d, c2f, fverts)
        # own-node permutations of the face's Lagrangian element (interior nodes in the face's reference space)
        if d == 0:
            d_perms[d] = [[1]]
        else:
            fp = ptype if d == D else rt._FACE_PTYPE[(ptype, d)]
            d_perms[d] = own_nodes_permutations(fp, rt.interior_nodes(fp, order), lambda x, fp=fp: rt.lagrangian_tabulate(fp, 1, x)[0])
    nc = len(cell_nodes)
    cell_dofs = np.zeros((nc, nl * ncomp), dtype=np.int32)
    for c in range(nc):
        for d in d_own:
            lf_own = face_own[offs[d]:offs[d + 1]]
            nown = len(lf_own[0])
            c2f = d_topo[d][0]
            for lf, own_nodes in enumerate(lf_own):
                face = c2f[c][lf] - 1
                inode_to_pinode = d_perms[d][d_pidx[d][c, lf] - 1]
                for comp in range(ncomp):
                    for i, lnode in enumerate(own_nodes):
                        j = inode_to_pinode[i]
                        cell_dofs[c, (lnode - 1) + nl * comp] = d_own[d][face, comp * nown + (j - 1)]
    return cell_dofs, nfree, ndiri


def multifield_offsets(nfrees):
    """offset_k = sum_{m<k} num_free_dofs(m) (MultiFieldFESpaces.jl:356-364)."""
    offs = [0]
    for n in nfrees[:-1]:
        offs.append(offs[-1] + n)
    return offs


def multifield_cell_dofs(cell_dofs_list, nfrees):
    """positive ids shifted by the field offset, negative ids untouched (:482-488)."""
    offs = multifield_offsets(nfrees)
    out = []
    for ids, o in zip(cell_dofs_list, offs):
        s = ids.copy()
        s[s > 0] += o
        out.append(s)
    return out
