"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, python loops: small meshes) of the reference's assembly of terms on a
SkeletonTriangulation; never imported by the product code.

What it follows:
  * interior facets, plus = first / minus = second incident cell: src/Geometry/SkeletonTriangulations.jl:54-99
    (SkeletonTriangulation(model) -> BoundaryTriangulation(model, face_to_mask, 1 | 2))
  * facet quadrature mapped into the reference space of the adjacent cell: FaceToCellGlue /
    compute_face_to_cell_reference_map, src/Geometry/BoundaryTriangulations.jl:13-70,320-340 (the vertex permutation between the
    facet and the cell's local face is realised here by matching the physical points of the two sides)
  * unit normal n = invJt . nref / |...| (push_normal, :310-318), facet measure
  * jump(a) = a+ - a-, jump(a n) = a+ n+ + a- n- = (a+ - a-) n+, mean(a) = (a+ + a-)/2: src/CellData/CellFields.jl:643-652
  * local 2x2 block matrix [plus, minus] x [plus, minus] added at the plus / minus cell DoF ids (BlockMap over SkeletonPair)
"""
import numpy as np

from . import ref_numbering as rn
from . import ref_tabulation as rt

FACET_PTYPE = {"HEX": "QUAD", "QUAD": "SEG", "TET": "TRI", "TRI": "SEG"}
REF_VERTS = {
    "QUAD": np.array([[0, 0], [1, 0], [0, 1], [1, 1]], dtype=float),
    "HEX": np.array([[x, y, z] for z in (0, 1) for y in (0, 1) for x in (0, 1)], dtype=float),
    "TRI": np.array([[0, 0], [1, 0], [0, 1]], dtype=float),
    "TET": np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=float),
}


def interior_facets(cell_nodes, ptype):
    """[(plus cell, plus local face, minus cell, minus local face)], 0-based, ascending facet id"""
    D = REF_VERTS[ptype].shape[1]
    c2f, fverts = rn.global_faces(cell_nodes, ptype, D - 1)
    touch = {}
    for c, row in enumerate(c2f):
        for lf, f in enumerate(row):
            touch.setdefault(int(f), []).append((c, lf))
    return [(t[0][0], t[0][1], t[1][0], t[1][1]) for f, t in sorted(touch.items()) if len(t) == 2]


def face_points(ptype, degree):
    """facet rule on every local face of the reference cell: pts [nlf][npf][D], w [npf], nref [nlf][D] (outward, scaled by the
    ratio of the reference measures)"""
    D = REF_VERTS[ptype].shape[1]
    fp = FACET_PTYPE[ptype]
    if fp == "SEG":
        xf, wf = rt.tensor_quadrature(1, degree)
        Nf = np.stack([1 - xf[:, 0], xf[:, 0]], axis=1)
    else:
        xf, wf = rt.quadrature(fp, degree)
        Nf, _ = rt.lagrangian_tabulate(fp, 1, xf)
    verts = REF_VERTS[ptype]
    centre = verts.mean(axis=0)
    pts, nref = [], []
    for lf in rn.local_face_vertices(ptype, D - 1):
        fv = verts[[k - 1 for k in lf]]
        pts.append(Nf @ fv)
        if D == 3:
            n = np.cross(fv[1] - fv[0], fv[2] - fv[0])
        else:
            t = fv[1] - fv[0]
            n = np.array([t[1], -t[0]])
        if np.dot(n, fv.mean(axis=0) - centre) < 0:
            n = -n
        nref.append(n)
    return np.array(pts), np.asarray(wf, dtype=float), np.array(nref)


def _side(X, nodes, ptype, order, pts):
    """physical points, inv(Jt), det at the reference points `pts` of one cell; N [np][nd], physical gradients [np][nd][D]"""
    Ng, dNg = rt.lagrangian_tabulate(ptype, 1, pts)
    N, dN = rt.lagrangian_tabulate(ptype, order, pts)
    Xc = X[[n - 1 for n in nodes]]
    x = Ng @ Xc
    Jt = np.einsum("pad,ae->pde", dNg, Xc)             # Jt[p][i][j] = sum_a d_i N_a x_a,j
    iJ = np.linalg.inv(Jt)
    det = np.linalg.det(Jt)
    G = np.einsum("pim,pam->pai", iJ, dN)              # grad phi_a = inv(Jt) . dN_a
    return x, iJ, det, N, G


def assemble_skeleton_dense(X, cell_nodes, ptype, order, ncomp, cell_dofs, degree, terms, nrows, ncols):
    """dense matrix of sum over terms (coef, T kind, w+, w-, U kind, z+, z-) of
    int_Lambda coef [w+ T(v+) + w- T(v-)] [z+ U(u+) + z- U(u-)], kinds: 0 value, 1 derivative along n+; equal components only"""
    A = np.zeros((nrows, ncols))
    pts, wf, nref = face_points(ptype, degree)
    nd = cell_dofs.shape[1] // ncomp
    for cp, lp, cm, lm in interior_facets(cell_nodes, ptype):
        xp, iJp, detp, Np, Gp = _side(X, cell_nodes[cp], ptype, order, pts[lp])
        xm, iJm, detm, Nm, Gm = _side(X, cell_nodes[cm], ptype, order, pts[lm])
        for p in range(len(wf)):
            q = int(np.argmin(((xm - xp[p]) ** 2).sum(axis=1)))
            assert np.allclose(xm[q], xp[p], atol=1e-12)
            v = iJp[p] @ nref[lp]
            m = np.linalg.norm(v)
            n = v / m
            dV = abs(detp[p]) * m * wf[p]
            val = [Np[p], Nm[q]]
            dn = [Gp[p] @ n, Gm[q] @ n]
            ids = [cell_dofs[cp], cell_dofs[cm]]
            for coef, tk, w0, w1, uk, z0, z1 in terms:
                for si, ws in ((0, w0), (1, w1)):
                    T = val[si] if int(tk) == 0 else dn[si]
                    for sj, zs in ((0, z0), (1, z1)):
                        U = val[sj] if int(uk) == 0 else dn[sj]
                        K = coef * ws * zs * dV * np.outer(T, U)     # [a][b]
                        for c in range(ncomp):
                            for a in range(nd):
                                r = ids[si][a + nd * c]
                                if r <= 0:
                                    continue
                                for b in range(nd):
                                    col = ids[sj][b + nd * c]
                                    if col > 0:
                                        A[r - 1, col - 1] += K[a, b]
    return A


def coupling_mask(cell_nodes, ptype, cell_dofs, nrows, ncols):
    """stored positions of the skeleton contribution: all (row, col) pairs of the DoFs of the two cells of every interior facet"""
    M = np.zeros((nrows, ncols), dtype=bool)
    for cp, _, cm, _ in interior_facets(cell_nodes, ptype):
        ids = np.concatenate([cell_dofs[cp], cell_dofs[cm]])
        ids = ids[ids > 0] - 1
        M[np.ix_(ids, ids)] = True
    return M
