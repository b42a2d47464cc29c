"""Helpers of the mid / large parity tests: the CUDA path against the CPU oracle (oracle/ref_assembly.c) ENTRY BY ENTRY on meshes
large enough to exercise the real control flow of the kernels (persistent warps taking several trips, run-length rows, chunked
pipelines, partial CTA batches).  The inputs (numbering, connectivity) come from the host package's vectorised spaces -- checked
against the oracle's line-by-line numbering on small meshes in test_gpu_forms.py / test_host_logic.py -- so that the oracle's
python-loop numbering does not limit the mesh size; tabulations come from the oracle's own restatement."""
import os

import numpy as np

import gridap_b200 as g
from oracle import capi
from oracle import ref_tabulation as rt

E, NU = 2.1e4, 0.3
LAM, MU = E * NU / ((1 + NU) * (1 - 2 * NU)), E / (2 * (1 + NU))


def relerr(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def perturb(model, frac, seed):
    """interior nodes displaced by frac * h * U(-1,1)^D (general-geometry variant), in place"""
    X = model.node_coordinates
    h = 1.0 / max(model.partition) if model.partition else 1.0 / round(model.num_cells() ** (1.0 / X.shape[1]))
    rng = np.random.default_rng(seed)
    lo, hi = X.min(axis=0), X.max(axis=0)
    inner = np.all((X > lo + 1e-9) & (X < hi - 1e-9), axis=1)
    X[inner] += frac * h * rng.uniform(-1, 1, size=(int(inner.sum()), X.shape[1]))
    model._device.clear()
    return model


def shear(model, A, b=(0.0, 0.0, 0.0)):
    """x -> A x + b: cells stay affine (parallelepipeds) but the Jacobian is full, in place"""
    model.node_coordinates[:] = model.node_coordinates @ np.asarray(A, dtype=np.float64).T + np.asarray(b)
    model._device.clear()
    return model


def oracle_field(model, V, degree, offset=0, free_values=None, dirichlet_values=None, ids=None, **kw):
    xq, w = rt.quadrature(model.ptype, degree)
    N, dN = rt.lagrangian_tabulate(model.ptype, V.reffe.order, xq)
    return capi.Field(N, dN, V.ncomp, V.cell_dof_ids if ids is None else ids, offset, free_values, dirichlet_values, **kw)


def oracle_problem(model, fields, degree, form_mat=0, form_vec=0, params=None, fq=None, touched=None, lift=False, nrows=None, ncols=None):
    xq, w = rt.quadrature(model.ptype, degree)
    Ng, dNg = rt.lagrangian_tabulate(model.ptype, 1, xq)
    return capi.Problem(model.node_coordinates, model.cell_node_ids, w, Ng, dNg, fields, form_mat, form_vec, params, fq, touched, 0, lift,
                        nrows, ncols)


def check_csc(A, ref, tol=1e-12):
    colptr, rowval, nzval = ref[:3]
    assert np.array_equal(A.colptr, colptr) and np.array_equal(A.rowval, rowval), "sparsity pattern differs from the oracle"
    assert relerr(A.nzval, nzval) <= tol


def hex_model(part, unstructured=True):
    m = g.CartesianDiscreteModel((0, 1) * len(part), part)
    return g.UnstructuredDiscreteModel(m) if unstructured else m


def facet_problem(G, V, degree, form_mat=0, form_vec=capi.SOURCE, params=None, fq=None, dirichlet_values=None, lift=False):
    """oracle Problem on the facets of a BoundaryTriangulation: facet mesh + facet DoF table from the host mirror, tabulation of the
    facet's own Lagrangian element by the oracle (rt), measure sqrt(det(Jt J))."""
    fm, fs = G.model, G.restrict(V)
    xq, w = rt.quadrature(fm.ptype, degree)
    N, dN = rt.lagrangian_tabulate(fm.ptype, V.reffe.order, xq)
    Ng, dNg = rt.lagrangian_tabulate(fm.ptype, 1, xq)
    fld = capi.Field(N, dN, V.ncomp, fs.cell_dof_ids, 0, None, dirichlet_values)
    return capi.Problem(fm.node_coordinates, fm.cell_node_ids, w, Ng, dNg, [fld], form_mat, form_vec, params, fq, None, 0, lift, V.nfree, V.nfree)


class env:
    """library tunables read from the environment per plan / per call (e.g. GB200_NO_AFFINE_GATHER, GB200_GATHER_DIAG)"""

    def __init__(self, **kw):
        self.kw = {k: str(v) for k, v in kw.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def facet_glue_oracle(ptype, degree):
    """FaceToCellGlue of a reference cell restated with the oracle's own tabulation (src/Geometry/BoundaryTriangulations.jl:13-70,
    320-340): facet rule mapped onto every local face -> (xq [nlf, npf, D], w [npf], nref [nlf, D] scaled outward normals)"""
    from oracle import ref_numbering as rn
    D = {"HEX": 3, "TET": 3, "QUAD": 2, "TRI": 2}[ptype]
    fp = {"HEX": "QUAD", "QUAD": "SEG", "TET": "TRI", "TRI": "SEG"}[ptype]
    xf, wf = rt.quadrature(fp, degree)
    Nf, _ = rt.lagrangian_tabulate(fp, 1, xf)
    if ptype in ("HEX", "QUAD"):
        verts = np.array([[(v >> d) & 1 for d in range(D)] for v in range(2 ** D)], dtype=np.float64)
    else:
        verts = np.vstack([np.zeros((1, D)), np.eye(D)])
    centre = verts.mean(axis=0)
    pts, nref = [], []
    for vs in rn.local_face_vertices(ptype, D - 1):
        fv = verts[[v - 1 for v in vs]]   # (the oracle keeps Gridap's 1-based local ids)
        pts.append(np.asarray(Nf) @ fv)
        n = np.cross(fv[1] - fv[0], fv[2] - fv[0]) if D == 3 else np.array([(fv[1] - fv[0])[1], -(fv[1] - fv[0])[0]])
        nref.append(n if np.dot(n, fv.mean(axis=0) - centre) > 0 else -n)
    return np.array(pts), np.asarray(wf), np.array(nref)


def glued_facet_problem(G, V, degree, form_mat=0, form_vec=0, params=None, fq=None, free_values=None, dirichlet_values=None, lift=False):
    """oracle Problem of a facet-of-cell term: the cells adjacent to the facets of the BoundaryTriangulation G, the facet rule mapped
    onto the local faces, the full cell DoF tables"""
    m = G.parent
    pts, wf, nref = facet_glue_oracle(m.ptype, degree)
    nlf, npf, D = pts.shape
    xq = pts.reshape(-1, D)
    N, dN = rt.lagrangian_tabulate(m.ptype, V.reffe.order, xq)
    Ng, dNg = rt.lagrangian_tabulate(m.ptype, 1, xq)
    fld = capi.Field(N, dN, V.ncomp, V.cell_dof_ids[G.cells], 0, free_values, dirichlet_values)
    return capi.Problem(m.node_coordinates, m.cell_node_ids[G.cells], np.tile(wf, nlf), Ng, dNg, [fld], form_mat, form_vec, params, fq, None, 0, lift,
                        V.nfree, V.nfree, lface=G.lfaces, nref=nref)
