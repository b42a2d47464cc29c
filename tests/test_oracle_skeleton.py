"""CPU pins of the oracle's skeleton restatement (oracle/ref_skeleton.py): the reference holds no golden matrix for jump / mean terms
(test/GridapTests/PoissonDGTests.jl checks the solution of the DG Poisson problem, replayed on the device in test_gpu_skeleton.py),
so the restatement is pinned by identities of the operators (src/CellData/CellFields.jl:643-652) on a perturbed mesh."""
import numpy as np
import pytest

from oracle import problems
from oracle import ref_skeleton as rs
from oracle import ref_tabulation as rt


def _l2_space(cells, ptype, order):
    nl = len(rt.lagrangian_nodes(ptype, order))
    nc = len(cells)
    ids = (np.arange(nc * nl) + 1).reshape(nc, nl).astype(np.int32)
    # DoF coordinates: the physical image of the reference nodes
    return ids, nl


def _dof_points(X, cells, ptype, order):
    Ng, _ = rt.lagrangian_tabulate(ptype, 1, rt.lagrangian_nodes(ptype, order))
    return np.concatenate([Ng @ X[[n - 1 for n in nodes]] for nodes in cells])


@pytest.mark.parametrize("ptype,order,part", [("QUAD", 1, (3, 3)), ("QUAD", 2, (3, 2)), ("TRI", 1, (3, 3)), ("HEX", 1, (2, 2, 2)), ("TET", 1, (2, 2, 2))])
def test_jump_and_mean_identities(ptype, order, part):
    D = len(part)
    X, cells, pt = problems.cartesian_mesh((0, 1) * D, part, simplex=ptype in ("TRI", "TET"))
    assert pt == ptype
    rng = np.random.default_rng(3)
    X = X.copy()
    inner = np.all((X > 1e-9) & (X < 1 - 1e-9), axis=1)
    if ptype in ("TRI", "TET", "QUAD") and order == 1 or ptype in ("TRI", "TET"):
        X[inner] += 0.1 / max(part) * rng.uniform(-1, 1, size=(int(inner.sum()), D))   # (simplices / bilinear maps: linears stay in the space)
    ids, nl = _l2_space(cells, ptype, order)
    n = ids.size
    P = _dof_points(X, cells, ptype, order)
    deg = 2 * order
    jj = rs.assemble_skeleton_dense(X, cells, ptype, order, 1, ids, deg, [(1.0, 0, 1.0, -1.0, 0, 1.0, -1.0)], n, n)
    jm = rs.assemble_skeleton_dense(X, cells, ptype, order, 1, ids, deg, [(1.0, 0, 1.0, -1.0, 1, 0.5, 0.5)], n, n)
    mj = rs.assemble_skeleton_dense(X, cells, ptype, order, 1, ids, deg, [(1.0, 1, 0.5, 0.5, 0, 1.0, -1.0)], n, n)
    # a continuous function has no jump: jump(v n).jump(u n) and mean(grad v).jump(u n) vanish on u = 1 + 2x - y (in every space here)
    ulin = 1.0 + 2.0 * P[:, 0] - P[:, 1]
    assert np.abs(jj @ ulin).max() <= 1e-12 and np.abs(mj @ ulin).max() <= 1e-12
    # symmetry of the penalty term, transposition of the consistency terms
    assert np.abs(jj - jj.T).max() <= 1e-13 and np.abs(jm - mj.T).max() <= 1e-13
    # v = indicator of cell K (discontinuous), u linear: int_{dK, interior} (n_K . grad u) = jm[K dofs].sum @ u;  by the divergence
    # theorem the integral over ALL of dK vanishes, so the interior part = - the boundary part of dK (computed from the geometry)
    g = np.array([2.0, -1.0, 0.0][:D])
    facets = rs.interior_facets(cells, ptype)
    pts, wf, nref = rs.face_points(ptype, deg)
    for K in range(len(cells)):
        lhs = (jm[ids[K] - 1] @ ulin).sum()
        # flux of the constant field g through the interior facets of K, by direct integration with outward normals
        flux = 0.0
        for cp, lp, cm, lm in facets:
            for c, lf in ((cp, lp), (cm, lm)):
                if c != K:
                    continue
                x, iJ, det, N, G = rs._side(X, cells[c], ptype, order, pts[lf])
                for p in range(len(wf)):
                    v = iJ[p] @ nref[lf]
                    flux += abs(det[p]) * wf[p] * (v @ g)     # |det| |invJt nref| w (n . g), n = v / |v|
        assert abs(lhs - flux) <= 1e-12 * max(1.0, abs(flux))
