"""TEST INFRASTRUCTURE ONLY -- builds oracle `Problem`s for the benchmark configs from the
line-by-line numbering / tabulation restatements (small meshes: python loops)."""
import numpy as np

from . import capi
from . import ref_numbering as rn
from . import ref_tabulation as rt


def cartesian_mesh(domain, partition, simplex=False):
    D = len(partition)
    X = rn.cartesian_node_coordinates(domain, partition)
    cells = rn.cartesian_cell_node_ids(partition)
    ptype = "HEX" if D == 3 else "QUAD"
    if simplex:
        cells = rn.simplexify(cells, ptype)
        ptype = "TET" if D == 3 else "TRI"
    return X, cells, ptype


def node_tags(partition, nnodes, dirichlet_tags):
    D = len(partition)
    ents = [rn.cartesian_entity_of_vertices(partition, [n]) for n in range(1, nnodes + 1)]
    return rn.face_tag_index(ents, D, dirichlet_tags)


def lagrangian_space(partition, cells, ptype, order, ncomp, dirichlet_tags, dirichlet_masks=None, nnodes=None):
    """FESpace(model, ReferenceFE(lagrangian,T,order); dirichlet_tags, dirichlet_masks)
    -> (cell_dofs, nfree, ndiri) following FESpaceFactories.jl:61-89."""
    D = len(partition)
    tags = list(dirichlet_tags) if isinstance(dirichlet_tags, (list, tuple)) else [dirichlet_tags]
    if dirichlet_masks is None:
        masks = [[True] * ncomp if ncomp > 1 else True for _ in tags]
    else:
        masks = dirichlet_masks
    if order == 1:
        n2t = node_tags(partition, nnodes, tags)
        nd, nfree, ndiri, _, _ = rn.clagrangian_dofs(n2t, masks, ncomp)
        return rn.clagrangian_cell_dofs(cells, nd), nfree, ndiri
    simplex = ptype in ("TET", "TRI")
    dims = [0, 1] if simplex else list(range(D))
    dface_to_tag = {}
    for d in dims:
        _, fverts = rn.global_faces(cells, ptype, d)
        ents = [rn.cartesian_entity_of_vertices(partition, list(v)) for v in fverts]
        dface_to_tag[d] = rn.face_tag_index(ents, D, tags)
    if order >= 3:
        if simplex:                       # (a P3 triangle face owns a node: the tags of the 2-faces of a tetrahedral mesh are needed too)
            for d in range(2, D):
                _, fverts = rn.global_faces(cells, ptype, d)
                ents = [rn.cartesian_entity_of_vertices(partition, list(v)) for v in fverts]
                dface_to_tag[d] = rn.face_tag_index(ents, D, tags)
        return rn.conforming_dofs(cells, ptype, order, ncomp, dface_to_tag, masks)
    cell_dofs, nfree, ndiri, _ = rn.conforming_dofs_order2(cells, ptype, ncomp, dface_to_tag, masks)
    return cell_dofs, nfree, ndiri


def tabulate(ptype, order, degree):
    xq, w = rt.quadrature(ptype, degree)
    N, dN = rt.lagrangian_tabulate(ptype, order, xq)
    Ng, dNg = rt.lagrangian_tabulate(ptype, 1, xq)
    return xq, w, N, dN, Ng, dNg


def single_field_problem(domain, partition, order=1, ncomp=1, degree=None, dirichlet_tags="boundary", dirichlet_masks=None,
                         form_mat=capi.LAPLACIAN, form_vec=0, params=None, fq=None, simplex=False, X=None,
                         dirichlet_values=None, free_values=None, lift=False):
    Xc, cells, ptype = cartesian_mesh(domain, partition, simplex)
    if X is None:
        X = Xc
    if degree is None:
        degree = 2 * order
    cell_dofs, nfree, ndiri = lagrangian_space(partition, cells, ptype, order, ncomp, dirichlet_tags, dirichlet_masks, nnodes=len(X))
    xq, w, N, dN, Ng, dNg = tabulate(ptype, order, degree)
    fld = capi.Field(N, dN, ncomp, cell_dofs, 0, free_values, dirichlet_values)
    pb = capi.Problem(X, cells, w, Ng, dNg, [fld], form_mat, form_vec, params, fq, None, 0, lift, nfree, nfree)
    pb.nfree, pb.ndiri, pb.ptype, pb.cells, pb.cell_dofs = nfree, ndiri, ptype, cells, cell_dofs
    pb.tab = (xq, w, N, dN, Ng, dNg)
    return pb


def stokes_problem(domain, partition, degree=4, simplex=True):
    """Taylor-Hood P2/P1 (Q2/Q1 if simplex=False), velocity Dirichlet on the boundary,
    consecutive multi-field style (test/GridapTests/StokesTaylorHoodTests.jl:59)."""
    X, cells, ptype = cartesian_mesh(domain, partition, simplex)
    D = len(partition)
    vd, nfu, ndu = lagrangian_space(partition, cells, ptype, 2, D, "boundary", None, nnodes=len(X))
    pd, nfp, ndp = lagrangian_space(partition, cells, ptype, 1, 1, [], None, nnodes=len(X))
    vd2, pd2 = rn.multifield_cell_dofs([vd, pd], [nfu, nfp])
    xq, w = rt.quadrature(ptype, degree)
    N2, dN2 = rt.lagrangian_tabulate(ptype, 2, xq)
    N1, dN1 = rt.lagrangian_tabulate(ptype, 1, xq)
    fu = capi.Field(N2, dN2, D, vd2, 0)
    fp = capi.Field(N1, dN1, 1, pd2, nfu)
    touched = np.array([[1, 1], [1, 0]], dtype=np.uint8)
    pb = capi.Problem(X, cells, w, N1, dN1, [fu, fp], capi.STOKES, 0, None, None, touched, 0, False, nfu + nfp, nfu + nfp)
    pb.nfree = (nfu, nfp)
    pb.ptype, pb.cells, pb.cell_dofs = ptype, cells, (vd, pd)
    pb.tab = (xq, w, N2, dN2, N1, dN1)
    return pb


def csc_to_dense(colptr, rowval, nzval, nrows, ncols):
    A = np.zeros((nrows, ncols))
    for j in range(ncols):
        for k in range(colptr[j] - 1, colptr[j + 1] - 1):
            A[rowval[k] - 1, j] += nzval[k]
    return A
