"""TEST INFRASTRUCTURE ONLY -- oracle restatement of Gridap's reference-element tabulation.

  * Quadrature(HEX/QUAD, degree)  tensor-product Gauss-Legendre mapped to [0,1], first axis fastest
        src/ReferenceFEs/TensorProductQuadratures.jl:68-79, Quadratures.jl:191-223
        (1-D rule: QuadGK.gauss, third-party, Project.toml compat "2.4"; restated with
        numpy.polynomial.legendre.leggauss -- agreement at the ulp level)
  * Quadrature(TET/TRI, degree)   Witherden-Vincent tables (published rule, constants are data)
        src/ReferenceFEs/WitherdenVincentQuadratures.jl:367-518,1290-1299, Quadratures.jl:224-241
  * Lagrangian shape functions    change = inv(dofs(prebasis)); shapefuns = change^T * prebasis
        src/ReferenceFEs/ReferenceFEInterfaces.jl:563-583, CLagrangianRefFEs.jl:246-275,466-545,
        src/Polynomials/MonomialBases.jl:33-58
"""
import itertools
import numpy as np

from . import ref_numbering as rn


def gauss_legendre_01(npts):
    x, w = np.polynomial.legendre.leggauss(npts)
    return (x + 1.0) / 2.0, w / 2.0


def tensor_quadrature(D, degree):
    """points [np][D], weights [np]; point index: first axis fastest (Quadratures.jl:204-223)."""
    n = degree // 2 + 1  # _npoints_from_degree (Quadratures.jl:191)
    x1, w1 = gauss_legendre_01(n)
    pts, ws = [], []
    for ci in itertools.product(*[range(n) for _ in range(D)]):
        ci = ci[::-1]
        w = 1.0
        p = []
        for d in range(D):
            p.append(x1[ci[d]])
            w *= w1[ci[d]]
        pts.append(p)
        ws.append(w)
    return np.array(pts), np.array(ws)


# Witherden-Vincent symmetric-orbit data on the [-1,1] simplex, degrees used by the configs.
_WV_TET = {
    1: dict(d1=1.3333333333333333333333333333333333333),
    2: dict(d2=[(0.33333333333333333333333333333333333333, -0.72360679774997896964091736687312762354,
                 0.17082039324993690892275210061938287063)]),
    3: dict(d2=[(0.18162379004944980942342872025562069427, -0.34367339496723662642072827083693243093,
                 -0.9689798150982901207378151874892027072),
                (0.15170954328388352390990461307771263906, -0.78390550020314279176487322158837338344,
                 0.35171650060942837529461966476512015033)]),
    4: dict(d2=[(0.15025056762402113439891420311104844508, -0.37822816147339878040530853247308433401,
                 -0.86531551557980365878407440258074699796),
                (0.097990724155149266058280273981770004697, -0.81452949937821754719535217252593878951,
                 0.44358849813465264158605651757781636853)],
            d3=[(0.05672802770277528858409257082700992237, -0.90899259174870070101623894744132112187,
                 -0.091007408251299298983761052558678878131)]),
}
_WV_TET[0] = _WV_TET[1]
_WV_TET[5] = _WV_TET[4]


def wv_tet_quadrature(degree):
    data = _WV_TET[degree]
    rows = []
    if "d1" in data:
        rows.append((data["d1"], -0.5, -0.5, -0.5))
    for (w, s, t) in data.get("d2", []):
        for (x, y, z) in ((s, s, t), (s, t, s), (t, s, s), (s, s, s)):
            rows.append((w, x, y, z))
    for (w, s, t) in data.get("d3", []):
        for (x, y, z) in ((s, t, t), (t, s, t), (s, s, t), (s, t, s), (t, s, s), (t, t, s)):
            rows.append((w, x, y, z))
    wx = np.array(rows)
    wx[:, 0] /= 2.0
    wx[:, 1:] = (wx[:, 1:] + 1.0) / 2.0
    w = wx[:, 0] * ((1.0 / 6.0) / wx[:, 0].sum())  # scale = get_measure(p)/sum(weights)
    return wx[:, 1:].copy(), w


# Witherden-Vincent data on the [-1,1] triangle (src/ReferenceFEs/WitherdenVincentQuadratures.jl:72-97,330-362): d1 = weight of the
# centroid (-1/3,-1/3), d2 = (w, s, t) orbits (s,t) (t,s) (s,s)
_WV_TRI = {
    1: dict(d1=2.0),
    2: dict(d2=[(0.66666666666666666666666666666666666667, -0.66666666666666666666666666666666666667, 0.33333333333333333333333333333333333333)]),
    4: dict(d2=[(0.44676317935602293139001401686624560874, -0.1081030181680702273633414922338960232, -0.7837939636638595452733170155322079536),
                (0.21990348731064373527665264980042105793, -0.81684757298045851308085707319559698429, 0.63369514596091702616171414639119396858)]),
}
_WV_TRI[0] = _WV_TRI[1]
_WV_TRI[3] = _WV_TRI[4]


def wv_tri_quadrature(degree):
    data = _WV_TRI[degree]
    rows = []
    if "d1" in data:
        rows.append((data["d1"], -1.0 / 3.0, -1.0 / 3.0))
    for (w, s, t) in data.get("d2", []):
        for (x, y) in ((s, t), (t, s), (s, s)):
            rows.append((w, x, y))
    wx = np.array(rows)
    wx[:, 0] /= 2.0                       # _geometric_map_to_01! (:1290-1299)
    wx[:, 1:] = (wx[:, 1:] + 1.0) / 2.0
    w = wx[:, 0] * (0.5 / wx[:, 0].sum())  # scale = get_measure(p)/sum(weights)
    return wx[:, 1:].copy(), w


def quadrature(ptype, degree):
    if ptype in ("HEX", "QUAD", "SEG"):
        return tensor_quadrature({"HEX": 3, "QUAD": 2, "SEG": 1}[ptype], degree)
    if ptype == "TET":
        return wv_tet_quadrature(degree)
    if ptype == "TRI":
        return wv_tri_quadrature(degree)
    raise NotImplementedError(ptype)


# ----------------------------------------------------------------------------- Lagrangian nodes
_DIMS = {"HEX": 3, "QUAD": 2, "TET": 3, "TRI": 2, "SEG": 1}
_FACE_PTYPE = {("HEX", 1): "SEG", ("HEX", 2): "QUAD", ("QUAD", 1): "SEG", ("TET", 1): "SEG", ("TET", 2): "TRI", ("TRI", 1): "SEG"}


def vertex_coordinates(ptype):
    D = _DIMS[ptype]
    if ptype in ("HEX", "QUAD", "SEG"):
        return np.array([[(v >> d) & 1 for d in range(D)] for v in range(2 ** D)], dtype=float)
    return np.array([list(v[:D]) for v in (rn.TET_VERTS if ptype == "TET" else [(0, 0), (1, 0), (0, 1)])], dtype=float)


def interior_nodes(ptype, order):
    """compute_own_nodes(p::ExtrusionPolytope, orders) = _interior_nodes(extrusion, orders) (CLagrangianRefFEs.jl:625-632,689-698):
    the terms of _add_terms!(terms, term, extrusion, orders, D, k=1) (:727-745) -- dimension D outermost, dimension 1 innermost,
    i = k .. orders[dim]-k, and on a TET_AXIS every step i != 0 lowers ALL the orders by one (cumulatively) -- turned into
    coordinates (t-1)/order (:763-778)."""
    D = _DIMS[ptype]
    simplex = ptype in ("TET", "TRI")
    terms = []

    def add_terms(term, orders, dim):
        term = list(term)
        orders = list(orders)
        for i in range(1, orders[dim - 1] - 1 + 1):
            term[dim - 1] = i
            if dim > 1:
                if simplex and i != 0:          # (extrusion[dim] == TET_AXIS for every dim >= 2 of a simplex)
                    orders = [o - 1 for o in orders]
                add_terms(term, orders, dim - 1)
            else:
                terms.append(tuple(term))

    add_terms([0] * D, [order] * D, D)
    return np.array([[t / order for t in term] for term in terms], dtype=float).reshape(len(terms), D)


def lagrangian_nodes_and_face_own_nodes(ptype, order):
    """compute_nodes(p, orders) -> (node coordinates, face_own_nodes) (CLagrangianRefFEs.jl:466-545,662-670): the vertices, then for
    d = 1 .. D-1 and every d-face of the polytope the interior nodes of the face's own Lagrangian element mapped by the face's
    LINEAR shape functions onto the face (face vertices in the local order of `get_faces(p,d,0)`), then the interior nodes of the
    polytope.  face_own_nodes: one list (1-based node ids) per face, faces ordered by dimension."""
    D = _DIMS[ptype]
    verts = vertex_coordinates(ptype)
    nodes = [v for v in verts]
    face_own = [[k + 1] for k in range(len(verts))]
    if order == 1:                                     # _compute_linear_nodes (:487-494)
        for d in range(1, D + 1):
            face_own += [[] for _ in ([0] if d == D else rn.local_face_vertices(ptype, d))]
        return verts, face_own
    for d in range(1, D):                              # _compute_high_order_nodes_dim_d! (:516-538)
        fp = _FACE_PTYPE[(ptype, d)]
        ref = interior_nodes(fp, order)
        shp, _ = lagrangian_tabulate(fp, 1, ref) if len(ref) else (np.zeros((0, 0)), None)
        for lf in rn.local_face_vertices(ptype, d):
            face_x = verts[[k - 1 for k in lf]]
            own = []
            for row in (shp @ face_x if len(ref) else []):
                nodes.append(row)
                own.append(len(nodes))
            face_own.append(own)
    own = []
    for x in interior_nodes(ptype, order):             # _compute_high_order_nodes_dim_D! (:540-548)
        nodes.append(x)
        own.append(len(nodes))
    face_own.append(own)
    nodes = np.round(np.array(nodes) * order) / order  # _coords_to_terms / _terms_to_coords (:747-778)
    return nodes, face_own


def lagrangian_nodes(ptype, order):
    """vertices, then interior nodes of each edge, each face, then the cell interior (CLagrangianRefFEs.jl:493-545)"""
    return lagrangian_nodes_and_face_own_nodes(ptype, order)[0]


def monomial_exponents(ptype, order):
    D = {"HEX": 3, "QUAD": 2, "TET": 3, "TRI": 2, "SEG": 1}[ptype]
    exps = []
    for e in itertools.product(*[range(order + 1) for _ in range(D)]):
        e = e[::-1]
        if ptype in ("TET", "TRI") and sum(e) > order:
            continue
        exps.append(e)
    return exps


def monomials(exps, x):
    """values [npts][nmono] and gradients [npts][nmono][D] of the monomial prebasis."""
    x = np.atleast_2d(x)
    D = x.shape[1]
    V = np.ones((x.shape[0], len(exps)))
    G = np.zeros((x.shape[0], len(exps), D))
    for m, e in enumerate(exps):
        for d in range(D):
            V[:, m] *= x[:, d] ** e[d]
        for k in range(D):
            if e[k] == 0:
                continue
            g = e[k] * x[:, k] ** (e[k] - 1)
            for d in range(D):
                if d != k:
                    g = g * x[:, d] ** e[d]
            G[:, m, k] = g
    return V, G


def lagrangian_tabulate(ptype, order, points):
    """N[p][a], dN[p][a][d] of the scalar Lagrangian basis at `points` (reference gradients)."""
    nodes = lagrangian_nodes(ptype, order)
    exps = monomial_exponents(ptype, order)
    A, _ = monomials(exps, nodes)  # A[node][mono] = dofs(prebasis)
    change = np.linalg.inv(A)  # [mono][shape]
    V, G = monomials(exps, points)
    N = V @ change
    dN = np.einsum("pmd,ma->pad", G, change)
    return N, dN
