"""Reference elements and quadratures tabulated on the host, once per reference element (a21 in SURVEY.md):
`ReferenceFE(lagrangian, T, order)`, `Quadrature(polytope, degree)`.

The tabulated arrays N[p,a], dN[p,a,:], w[p] are what the C ABI takes (`gb200_refel_create`).  Shape functions are
written in closed form (tensor products of 1-D Lagrange polynomials on n-cubes, barycentric formulas on simplices) in
Gridap's node order: vertices, then edge / face / interior nodes in local-face order
(src/ReferenceFEs/CLagrangianRefFEs.jl:493-545).
"""
import numpy as np

from .geometry import _DIM, local_face_vertices

lagrangian = "lagrangian"


class VectorValue:
    """Type tag: VectorValue{D,Float64}."""

    def __init__(self, D):
        self.D = int(D)

    def __eq__(self, o):
        return isinstance(o, VectorValue) and o.D == self.D

    def __hash__(self):
        return hash(("VectorValue", self.D))


class ReferenceFE:
    """ReferenceFE(lagrangian, T, order) -- the (name, args) tuple of Gridap (src/ReferenceFEs/ReferenceFEInterfaces.jl)."""

    def __init__(self, name, T, order):
        if name != lagrangian:
            raise NotImplementedError("only `lagrangian` reference FEs are on the B200 path (got %r)" % (name,))
        if order not in (1, 2, 3):
            raise NotImplementedError("Lagrangian order %r: the B200 path covers orders 1, 2 and 3" % (order,))
        self.name, self.T, self.order = name, T, int(order)
        self.ncomp = T.D if isinstance(T, VectorValue) else 1


# ------------------------------------------------------------------------------------------------ quadratures
def _gauss01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return (x + 1.0) / 2.0, w / 2.0


def tensor_product_quadrature(D, degree):
    """Quadrature(QUAD/HEX, degree): n = div(degree,2)+1 Gauss-Legendre points per axis mapped to [0,1], point index
    first axis fastest (src/ReferenceFEs/TensorProductQuadratures.jl:68-79, Quadratures.jl:191-223)."""
    n = degree // 2 + 1
    x1, w1 = _gauss01(n)
    idx = np.stack(np.unravel_index(np.arange(n ** D), (n,) * D, order="F"), axis=1)
    return x1[idx], np.prod(w1[idx], axis=1)


# Witherden-Vincent orbit data on the [-1,1] tetrahedron (published rule; src/ReferenceFEs/WitherdenVincentQuadratures.jl:367-518)
_WV_TET = {
    1: ([1.3333333333333333333333333333333333333], [], []),
    2: (None, [(0.33333333333333333333333333333333333333, -0.72360679774997896964091736687312762354,
                0.17082039324993690892275210061938287063)], []),
    3: (None, [(0.18162379004944980942342872025562069427, -0.34367339496723662642072827083693243093,
                -0.9689798150982901207378151874892027072),
               (0.15170954328388352390990461307771263906, -0.78390550020314279176487322158837338344,
                0.35171650060942837529461966476512015033)], []),
    4: (None, [(0.15025056762402113439891420311104844508, -0.37822816147339878040530853247308433401,
                -0.86531551557980365878407440258074699796),
               (0.097990724155149266058280273981770004697, -0.81452949937821754719535217252593878951,
                0.44358849813465264158605651757781636853)],
        [(0.05672802770277528858409257082700992237, -0.90899259174870070101623894744132112187,
          -0.091007408251299298983761052558678878131)]),
}
_WV_TET[0] = _WV_TET[1]
_WV_TET[5] = _WV_TET[4]


def witherden_vincent_tet(degree):
    if degree not in _WV_TET:
        raise NotImplementedError("Witherden-Vincent TET rule of degree %d is not tabulated here (0..5 are)" % degree)
    d1, d2, d3 = _WV_TET[degree]
    rows = []
    if d1:
        rows.append((d1[0], -0.5, -0.5, -0.5))
    for (w, s, t) in d2:
        rows += [(w, s, s, t), (w, s, t, s), (w, t, s, s), (w, s, s, s)]
    for (w, s, t) in d3:
        rows += [(w, s, t, t), (w, t, s, t), (w, s, s, t), (w, s, t, s), (w, t, s, s), (w, t, t, s)]
    wx = np.array(rows)
    x = (wx[:, 1:] + 1.0) / 2.0
    w = wx[:, 0] / 2.0
    return x, w * ((1.0 / 6.0) / w.sum())


# Witherden-Vincent orbit data on the [-1,1] triangle (published rule; src/ReferenceFEs/WitherdenVincentQuadratures.jl:72-97)
_WV_TRI = {
    1: (2.0, []),
    2: (None, [(0.66666666666666666666666666666666666667, -0.66666666666666666666666666666666666667, 0.33333333333333333333333333333333333333)]),
    4: (None, [(0.44676317935602293139001401686624560874, -0.1081030181680702273633414922338960232, -0.7837939636638595452733170155322079536),
               (0.21990348731064373527665264980042105793, -0.81684757298045851308085707319559698429, 0.63369514596091702616171414639119396858)]),
}
_WV_TRI[0] = _WV_TRI[1]
_WV_TRI[3] = _WV_TRI[4]


def witherden_vincent_tri(degree):
    if degree not in _WV_TRI:
        raise NotImplementedError("Witherden-Vincent TRI rule of degree %d is not tabulated here (0..4 are)" % degree)
    d1, d2 = _WV_TRI[degree]
    rows = []
    if d1:
        rows.append((d1, -1.0 / 3.0, -1.0 / 3.0))
    for (w, s, t) in d2:
        rows += [(w, s, t), (w, t, s), (w, s, s)]
    wx = np.array(rows)
    x = (wx[:, 1:] + 1.0) / 2.0
    w = wx[:, 0] / 2.0
    return x, w * (0.5 / w.sum())


def Quadrature(ptype, degree):
    """Quadrature(p::Polytope, degree) (src/ReferenceFEs/Quadratures.jl:156-176)."""
    if ptype in ("HEX", "QUAD", "SEG"):
        return tensor_product_quadrature(_DIM[ptype], degree)
    if ptype == "TET":
        return witherden_vincent_tet(degree)
    if ptype == "TRI":
        return witherden_vincent_tri(degree)
    raise NotImplementedError("quadratures on %s" % ptype)


# ------------------------------------------------------------------------------------------------ shape functions
def lagrangian_lattice(ptype, order):
    """Lattice positions (integers, coordinate = index / order) of the Lagrangian nodes in Gridap's order, and who owns them:
    vertices, then the interior nodes of every edge, every face, then of the cell (src/ReferenceFEs/CLagrangianRefFEs.jl:493-545).
    Interior nodes of an n-cube face run over the face's own axes, first axis fastest (:689-745); those of an edge (a, b) of a simplex
    from a towards b, a triangle of order 3 owns its centroid.
    -> lattice [nl, D] (n-cubes: Cartesian index; simplices: Cartesian index, lam_0 = order - sum), own {d: [nlf_d, nown_d] local node ids}"""
    D = _DIM[ptype]
    k = int(order)
    if ptype in ("HEX", "QUAD", "SEG"):
        from .geometry import ncube_faces
        nodes, own = [], {d: [] for d in range(D + 1)}
        for (dim, e, a) in ncube_faces(D):
            axes = [ax for ax in range(D) if (e >> ax) & 1]
            mine = []
            for flat in range((k - 1) ** dim):
                idx = [k * ((a >> ax) & 1) for ax in range(D)]
                r = flat
                for ax in axes:                      # first own axis fastest
                    idx[ax] = 1 + r % (k - 1)
                    r //= (k - 1)
                mine.append(len(nodes))
                nodes.append(idx)
            own[dim].append(mine)
        return np.array(nodes, dtype=np.int64), {d: np.array(v, dtype=np.int64).reshape(len(v), -1) for d, v in own.items()}
    if k > 3:
        raise NotImplementedError("simplices of order %d" % k)
    verts = np.concatenate([np.zeros((1, D), dtype=np.int64), np.eye(D, dtype=np.int64)], axis=0) * k
    nodes = [v for v in verts]
    own = {0: [[i] for i in range(D + 1)]}
    for d in range(1, D + 1):
        own[d] = []
        for lf in local_face_vertices(ptype, d):
            mine = []
            if d == 1:
                va, vb = verts[lf[0]], verts[lf[1]]
                for j in range(1, k):
                    mine.append(len(nodes))
                    nodes.append(va + (vb - va) * j // k)
            elif d == 2 and k == 3:
                mine.append(len(nodes))
                nodes.append(verts[lf].sum(axis=0) // 3)
            own[d].append(mine)
    return np.array(nodes, dtype=np.int64), {d: np.array(v, dtype=np.int64).reshape(len(v), -1) for d, v in own.items()}


def lagrangian_node_multiindex(ptype, order):
    """n-cubes: per node a multi-index in {0,1,2=midpoint}^D, Gridap node order."""
    D = _DIM[ptype]
    verts = [[(v >> d) & 1 for d in range(D)] for v in range(2 ** D)]
    nodes = [list(v) for v in verts]
    if order == 2:
        for d in range(1, D + 1):
            for lf in local_face_vertices(ptype, d):
                mi = []
                for k in range(D):
                    vals = {verts[v][k] for v in lf}
                    mi.append(vals.pop() if len(vals) == 1 else 2)
                nodes.append(mi)
    return np.array(nodes)


def _lagrange_1d(order, x):
    """values L[k](x), derivatives dL[k](x) for node k in (0, 1[, 2 = 0.5])"""
    if order == 1:
        return np.stack([1 - x, x]), np.stack([-np.ones_like(x), np.ones_like(x)])
    L = np.stack([2 * (x - 0.5) * (x - 1), 2 * x * (x - 0.5), 4 * x * (1 - x)])
    dL = np.stack([4 * x - 3, 4 * x - 1, 4 - 8 * x])
    return L, dL


def _lagrange_1d_lattice(order, x):
    """L[j](x), dL[j](x) of the 1-D Lagrange polynomials on the equispaced nodes j / order, j = 0..order (product form)"""
    k = int(order)
    t = k * np.asarray(x, dtype=np.float64)
    L = np.ones((k + 1,) + t.shape)
    dL = np.zeros((k + 1,) + t.shape)
    for j in range(k + 1):
        others = [m for m in range(k + 1) if m != j]
        for m in others:
            L[j] *= (t - m) / (j - m)
        for s in others:
            term = np.full(t.shape, k / (j - s))
            for m in others:
                if m != s:
                    term = term * (t - m) / (j - m)
            dL[j] += term
    return L, dL


def _simplex_factor(k, i, lam):
    """phi_i(lam) = prod_{t<i} (k lam - t) / (t + 1) and its derivative (the barycentric factor of the simplex Lagrange basis)"""
    v = np.ones_like(lam)
    dv = np.zeros_like(lam)
    for s in range(i):
        term = np.full(lam.shape, k / (s + 1.0))
        for t in range(i):
            if t != s:
                term = term * (k * lam - t) / (t + 1.0)
        dv += term
    for t in range(i):
        v = v * (k * lam - t) / (t + 1.0)
    return v, dv


def tabulate_lagrangian(ptype, order, points):
    """N[p,a], dN[p,a,d] (reference gradients) of the scalar Lagrangian basis of `order` on `ptype`."""
    points = np.atleast_2d(np.asarray(points, dtype=np.float64))
    D = _DIM[ptype]
    npts = points.shape[0]
    if ptype in ("HEX", "QUAD", "SEG"):
        if order <= 2:
            mi = lagrangian_node_multiindex(ptype, order)
            one_d = _lagrange_1d
        else:
            mi, _ = lagrangian_lattice(ptype, order)
            one_d = _lagrange_1d_lattice
        L = [None] * D
        dL = [None] * D
        for d in range(D):
            L[d], dL[d] = one_d(order, points[:, d])
        nd = len(mi)
        N = np.ones((npts, nd))
        dN = np.ones((npts, nd, D))
        for a in range(nd):
            for d in range(D):
                N[:, a] *= L[d][mi[a, d]]
                for k in range(D):
                    dN[:, a, k] *= dL[d][mi[a, d]] if k == d else L[d][mi[a, d]]
        return N, dN
    # simplices: barycentric coordinates lam_0 = 1 - sum x, lam_i = x_{i-1}
    lam = np.concatenate([1.0 - points.sum(axis=1, keepdims=True), points], axis=1)  # [p, D+1]
    dlam = np.concatenate([-np.ones((1, D)), np.eye(D)], axis=0)  # [D+1, D]
    if order == 1:
        return lam.copy(), np.broadcast_to(dlam, (npts, D + 1, D)).copy()
    if order >= 3:   # N_a = prod_m phi_{i_m}(lam_m) on the barycentric lattice (i_0, .., i_D), sum = order
        lat, _ = lagrangian_lattice(ptype, order)
        bary = np.concatenate([order - lat.sum(axis=1, keepdims=True), lat], axis=1)
        nd = len(bary)
        N = np.ones((npts, nd))
        dN = np.zeros((npts, nd, D))
        for a in range(nd):
            fac = [_simplex_factor(order, int(bary[a, m]), lam[:, m]) for m in range(D + 1)]
            for m in range(D + 1):
                N[:, a] *= fac[m][0]
                g = fac[m][1].copy()
                for n in range(D + 1):
                    if n != m:
                        g = g * fac[n][0]
                dN[:, a, :] += g[:, None] * dlam[m][None, :]
        return N, dN
    edges = local_face_vertices(ptype, 1)
    nd = D + 1 + len(edges)
    N = np.zeros((npts, nd))
    dN = np.zeros((npts, nd, D))
    for i in range(D + 1):
        N[:, i] = lam[:, i] * (2 * lam[:, i] - 1)
        dN[:, i, :] = (4 * lam[:, i] - 1)[:, None] * dlam[i][None, :]
    for e, (a, b) in enumerate(edges):
        N[:, D + 1 + e] = 4 * lam[:, a] * lam[:, b]
        dN[:, D + 1 + e, :] = 4 * (lam[:, a][:, None] * dlam[b][None, :] + lam[:, b][:, None] * dlam[a][None, :])
    return N, dN


def reference_nodes(ptype, order):
    """reference coordinates of the Lagrangian nodes (Gridap order)."""
    D = _DIM[ptype]
    if order >= 3:
        return lagrangian_lattice(ptype, order)[0] / float(order)
    if ptype in ("HEX", "QUAD", "SEG"):
        mi = lagrangian_node_multiindex(ptype, order)
        return np.where(mi == 2, 0.5, mi.astype(float))
    verts = np.concatenate([np.zeros((1, D)), np.eye(D)], axis=0)
    if order == 1:
        return verts
    mids = [verts[list(e)].mean(axis=0) for e in local_face_vertices(ptype, 1)]
    return np.concatenate([verts, np.array(mids)], axis=0)


_FACET_PTYPE = {"HEX": "QUAD", "QUAD": "SEG", "TET": "TRI", "TRI": "SEG"}


def reference_vertices(ptype):
    """vertex coordinates of the reference polytope in Gridap's order (n-cubes: first axis fastest; simplices: origin, e_1, ...)"""
    D = _DIM[ptype]
    if ptype in ("HEX", "QUAD", "SEG"):
        return np.array([[(v >> d) & 1 for d in range(D)] for v in range(2 ** D)], dtype=np.float64)
    return np.vstack([np.zeros((1, D)), np.eye(D)])


def facet_glue(ptype, degree):
    """FaceToCellGlue data of a reference cell (src/Geometry/BoundaryTriangulations.jl:13-70,320-340): the facet quadrature of
    `degree` mapped onto every local face of the reference cell.
    -> xq [nlf, npf, D] points in the cell's reference space, w [npf], nref [nlf, D] outward reference normals scaled by the ratio
    of the reference measures (face of the reference cell / facet reference polytope; get_facet_normal gives the unit normals)."""
    D = _DIM[ptype]
    fp = _FACET_PTYPE[ptype]
    xf, wf = Quadrature(fp, degree)
    Nf, _ = tabulate_lagrangian(fp, 1, xf)          # the facet's own order-1 map: [npf, nv]
    verts = reference_vertices(ptype)
    centre = verts.mean(axis=0)
    lfv = local_face_vertices(ptype, D - 1)
    pts, nref = [], []
    for vs in lfv:
        fv = verts[vs]
        pts.append(Nf @ fv)
        if D == 3:
            n = np.cross(fv[1] - fv[0], fv[2] - fv[0])
        else:
            t = fv[1] - fv[0]
            n = np.array([t[1], -t[0]])
        if np.dot(n, fv.mean(axis=0) - centre) < 0:
            n = -n
        nref.append(n)
    return np.array(pts), wf, np.array(nref)
