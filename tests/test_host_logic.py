"""CPU tests of the host side: numbering / tabulation of the product vs the oracle, the integrand recogniser, the C-ABI
exports, and the multi-GPU partition logic (world_size 2, gloo)."""
import ctypes
import os
import re

import numpy as np
import pytest

import gridap_b200 as g
from gridap_b200 import celldata as cd
from gridap_b200 import distributed as gd
from gridap_b200 import lib
from oracle import capi, problems
from oracle import ref_numbering as rn
from oracle import ref_tabulation as rt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gridap_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(gb200_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 25
    L = ctypes.CDLL(lib.LIB_PATH)
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(lib.SYMBOLS) == declared  # the Python binding covers the whole ABI
    assert b"sm_100a" in lib.load().gb200_version()


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.GridapB200Error) as e:
        lib.Context(0)
    assert "no CPU path" in str(e.value)


@pytest.mark.parametrize("part", [(3, 4), (3, 2, 4), (100, 60), (13, 9, 11)])
def test_mesh_and_dof_numbering_match_oracle(part):
    D = len(part)
    m = g.CartesianDiscreteModel((0, 1) * D, part)
    assert np.array_equal(m.cell_node_ids, rn.cartesian_cell_node_ids(part))
    assert np.array_equal(m.node_coordinates, rn.cartesian_node_coordinates((0, 1) * D, part))
    X, cells, ptype = problems.cartesian_mesh((0, 1) * D, part)
    for order in (1, 2):
        for ncomp in (1, D):
            for tags in ([], ["boundary"], [5] if D == 2 else [25], [1, 7] if D == 2 else [3, 12, 22]):
                T = float if ncomp == 1 else g.VectorValue(D)
                V = g.FESpace(m, g.ReferenceFE(g.lagrangian, T, order), dirichlet_tags=tags)
                cdofs, nf, ndr = problems.lagrangian_space(part, cells, ptype, order, ncomp, tags, None, nnodes=len(X))
                assert (nf, ndr) == (V.nfree, V.ndirichlet)
                assert np.array_equal(cdofs, V.cell_dof_ids)
    ms = g.simplexify(m)
    Xs, cells_s, ptype_s = problems.cartesian_mesh((0, 1) * D, part, simplex=True)
    assert np.array_equal(ms.cell_node_ids, cells_s)
    if D == 3:
        for order, ncomp, tags in ((2, 3, ["boundary"]), (1, 1, []), (2, 1, [21])):
            T = float if ncomp == 1 else g.VectorValue(D)
            V = g.FESpace(ms, g.ReferenceFE(g.lagrangian, T, order), dirichlet_tags=tags)
            cdofs, nf, ndr = problems.lagrangian_space(part, cells_s, ptype_s, order, ncomp, tags, None, nnodes=len(Xs))
            assert nf == V.nfree and np.array_equal(cdofs, V.cell_dof_ids)


def test_dirichlet_masks_match_reference_golden():
    # test/FESpacesTests/CLagrangianFESpacesTests.jl:52-72
    m = g.CartesianDiscreteModel((0, 1, 0, 1), (2, 2))
    tags = [1, 2, 4, 5, 8]
    V = g.FESpace(m, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags=tags, dirichlet_masks=[True, True, False, True, True])
    assert V.cell_dof_ids.tolist() == [[-1, -2, 1, 2], [-2, -3, 2, -4], [1, 2, 3, 4], [2, -4, 4, 5]]
    masks2 = [(True, True), (True, False), (False, False), (False, True), (True, True)]
    V = g.FESpace(m, g.ReferenceFE(g.lagrangian, g.VectorValue(2), 1), dirichlet_tags=tags, dirichlet_masks=masks2)
    assert V.cell_dof_ids.tolist() == [[-1, 1, 3, 5, -2, -3, 4, 6], [1, -4, 5, -5, -3, 2, 6, -6], [3, 5, 7, 9, 4, 6, 8, 10], [5, -5, 9, 11, 6, -6, 10, 12]]
    V = g.FESpace(m, g.ReferenceFE(g.lagrangian, g.VectorValue(2), 1))
    assert V.cell_dof_ids.tolist() == [[1, 3, 7, 9, 2, 4, 8, 10], [3, 5, 9, 11, 4, 6, 10, 12], [7, 9, 13, 15, 8, 10, 14, 16], [9, 11, 15, 17, 10, 12, 16, 18]]


def test_tabulation_matches_oracle():
    for ptype, order, deg in (("HEX", 1, 2), ("HEX", 2, 4), ("QUAD", 1, 2), ("TET", 2, 4), ("TET", 1, 4), ("QUAD", 2, 4), ("HEX", 1, 3)):
        xq, w = g.Quadrature(ptype, deg)
        xo, wo = rt.quadrature(ptype, deg)
        assert np.allclose(xq, xo, atol=1e-15, rtol=0) and np.allclose(w, wo, atol=1e-16, rtol=0)
        N, dN = g.reffes.tabulate_lagrangian(ptype, order, xq)
        No, dNo = rt.lagrangian_tabulate(ptype, order, xo)
        assert np.allclose(N, No, atol=1e-13, rtol=0) and np.allclose(dN, dNo, atol=1e-12, rtol=0)


def test_dirichlet_value_interpolation():
    m = g.CartesianDiscreteModel((0, 1, 0, 2), (2, 2))
    V = g.FESpace(m, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
    U = g.TrialFESpace(V, lambda x: x[:, 0] + 10 * x[:, 1])
    X = m.node_coordinates
    ids = V.node_and_comp_to_dof[:, 0]
    for node in range(9):
        if ids[node] < 0:
            assert U.dirichlet_values[-ids[node] - 1] == X[node, 0] + 10 * X[node, 1]
    V2 = g.FESpace(m, g.ReferenceFE(g.lagrangian, g.VectorValue(2), 2), dirichlet_tags=[7])
    U2 = g.TrialFESpace(V2, lambda x: np.stack([x[:, 1], -x[:, 0] + 1], axis=1))
    fx, fc, dx, dc = V2.dof_coordinates()
    assert np.allclose(dx[:, 0], 0.0) and np.allclose(U2.dirichlet_values, np.where(dc == 0, dx[:, 1], 1.0))


def _spaces():
    m = g.CartesianDiscreteModel((0, 1) * 3, (2, 2, 2))
    V = g.FESpace(m, g.ReferenceFE(g.lagrangian, float, 1))
    W = g.FESpace(m, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 1))
    dO = g.Measure(g.Triangulation(m), 2)
    return m, V, W, dO


def test_recogniser_maps_forms_to_kernels():
    m, V, W, dO = _spaces()
    u, v = g.get_trial_fe_basis(V), g.get_fe_basis(V)
    t = cd.recognise_matrix(g.inner(g.grad(v), g.grad(u)))
    assert [(x.form, x.params) for x in t] == [(lib.FORM_LAPLACIAN, (1.0,))]
    t = cd.recognise_matrix(g.dot(g.grad(u), g.grad(v)) * 3.0 + u * v)
    assert [(x.form, x.params) for x in t] == [(lib.FORM_LAPLACIAN, (3.0,)), (lib.FORM_MASS, (1.0,))]
    law = g.IsotropicLinearElasticity.from_E_nu(2.1e4, 0.3)
    uu, vv = g.get_trial_fe_basis(W), g.get_fe_basis(W)
    t = cd.recognise_matrix(g.inner(g.eps(vv), law(g.eps(uu))))
    assert t[0].form == lib.FORM_ELASTICITY and np.allclose(t[0].params, (law.lam, law.mu))
    Y = g.MultiFieldFESpace([W, V])
    (uf, pf), (vf, qf) = g.get_trial_fe_basis(Y), g.get_fe_basis(Y)
    t = cd.recognise_matrix(g.inner(g.grad(vf), g.grad(uf)) - g.div(vf) * pf + qf * g.div(uf))
    assert [x.form for x in t] == [lib.FORM_STOKES]
    nh = g.NeoHookean(100.0, 1.0)
    uh = g.FEFunction(W, np.zeros(W.num_free_dofs()))
    t = cd.recognise_matrix(nh.jac(uh, uu, vv))
    assert t[0].form == lib.FORM_NEOHOOKEAN_JAC and t[0].state is uh
    t = cd.recognise_vector(nh.res(uh, vv))
    assert t[0].form == lib.FORM_NEOHOOKEAN_RES and t[0].state is uh
    t = cd.recognise_vector(v * 2.0)
    assert (t[0].form, t[0].params) == (lib.FORM_SOURCE, (2.0,))
    t = cd.recognise_vector(g.dot(vv, (0.0, 0.0, -1.0)))
    assert (t[0].form, t[0].params) == (lib.FORM_SOURCE, (0.0, 0.0, -1.0))
    f = lambda x: x[:, 0]  # noqa: E731
    t = cd.recognise_vector(v * f)
    assert t[0].form == lib.FORM_SOURCE and t[0].fq is f


def test_recogniser_rejects_everything_else():
    m, V, W, dO = _spaces()
    u, v = g.get_trial_fe_basis(V), g.get_fe_basis(V)
    for bad in (g.inner(g.grad(v), u), g.inner(g.grad(g.grad(v)), g.grad(g.grad(u))), g.inner(v, v)):
        with pytest.raises(NotImplementedError):
            cd.recognise_matrix(bad)
    with pytest.raises(NotImplementedError):
        cd.recognise_vector(g.inner(g.grad(v), (1.0, 0.0, 0.0)))
    Y = g.MultiFieldFESpace([W, V])
    (uf, pf), (vf, qf) = g.get_trial_fe_basis(Y), g.get_fe_basis(Y)
    with pytest.raises(NotImplementedError):
        cd.recognise_matrix(g.inner(g.grad(vf), g.grad(uf)) + qf * pf)  # touches the (q,p) block: not Stokes
    with pytest.raises(NotImplementedError):
        g.FEOperator(lambda u, v: None, None, V, V)  # no AD Jacobian on the GPU path


def _slab_worker(rank, world, port, n, out):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    model = g.CartesianDiscreteModel((0, 1) * 3, (n, n, n))
    V = g.FESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
    part = gd.slab_partition(model, V, world, rank)
    lo, hi = part.col_range
    ids = part.local_space.cell_dof_ids.copy()
    pos = ids > 0
    owned = pos & (ids > lo) & (ids <= hi)
    cols = ids.copy()
    cols[pos & ~owned] = 0
    cols[owned] -= lo
    # the CPU oracle stands in for the device here: rows = global ids, cols = masked local ids
    xq, w = rt.quadrature("HEX", 2)  # the oracle's own tabulation on both sides -> bitwise comparison is meaningful
    N, dN = rt.lagrangian_tabulate("HEX", 1, xq)
    lm = part.local_model
    slab = _assemble_rows_cols(lm, w, N, dN, ids, cols, V.nfree, hi - lo)
    # the rank's RHS over its local cells (own slab + ghost layer): complete on the rows it owns
    pbv = capi.Problem(lm.node_coordinates, lm.cell_node_ids, w, N, dN, [capi.Field(N, dN, 1, ids)], 0, capi.SOURCE, [1.0], nrows=V.nfree, ncols=V.nfree)
    bloc = pbv.assemble_vector()
    gathered = [None] * world
    dist.all_gather_object(gathered, (slab, bloc, part.col_range))
    if rank == 0:
        A = gd.gather_csc([t[0] for t in gathered], V.nfree)
        b = gd.gather_vector([t[1] for t in gathered], [t[2] for t in gathered])
        out.put((A.colptr, A.rowval, A.nzval, b))
    dist.barrier()
    dist.destroy_process_group()


def _assemble_rows_cols(lm, w, N, dN, rows, cols, nrows, ncols):
    """serial reference loop with distinct row / column tables (ids <= 0 skipped), local matrices from the oracle."""
    pb = capi.Problem(lm.node_coordinates, lm.cell_node_ids, w, N, dN, [capi.Field(N, dN, 1, rows)], capi.LAPLACIAN, 0, nrows=nrows, ncols=nrows)
    b = capi.Builder(nrows, ncols)
    nc = rows.shape[0]
    for c in range(nc):
        for j in cols[c]:
            for i in rows[c]:
                b.count(i, j)
    b.allocate()
    for c in range(nc):
        Ke = pb.cell_local(c)[0][0][0]
        for lj, j in enumerate(cols[c]):
            for li, i in enumerate(rows[c]):
                if i > 0 and j > 0:
                    b.add(Ke[li, lj], int(i), int(j))
    return b.finish()


def test_two_rank_column_slabs_reproduce_the_serial_matrix():
    import torch.multiprocessing as mp
    n, world = 4, 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    colptr, rowval, nzval, bvec = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pb = problems.single_field_problem((0, 1) * 3, (n, n, n), form_mat=capi.LAPLACIAN)
    cp, rv, nz = pb.assemble()
    assert np.array_equal(colptr, cp) and np.array_equal(rowval, rv)
    assert np.array_equal(nzval, nz)  # same cells in the same order per column -> bitwise equal
    pbs = problems.single_field_problem((0, 1) * 3, (n, n, n), form_mat=capi.LAPLACIAN, form_vec=capi.SOURCE, params=[1.0])
    assert np.array_equal(bvec, pbs.assemble(with_vector=True)[3])   # owned rows: same cells in the same order


def test_column_ranges_cover_everything():
    for nfree, world in ((27, 2), (1000, 8), (7, 8)):
        r = gd.column_ranges(nfree, world)
        assert r[0][0] == 0 and r[-1][1] == nfree and all(a[1] == b[0] for a, b in zip(r, r[1:]))


def _facet_problem(G, V, degree, form_mat=0, form_vec=capi.SOURCE, params=None, fq=None, dirichlet_values=None, lift=False):
    """oracle Problem on the facets of a BoundaryTriangulation: facet mesh + facet DoF table from the host mirror, tabulation of the
    facet's own Lagrangian element by the oracle (rt), measure sqrt(det(Jt J))."""
    fm, fs = G.model, G.restrict(V)
    xq, w = rt.quadrature(fm.ptype, degree)
    N, dN = rt.lagrangian_tabulate(fm.ptype, V.reffe.order, xq)
    Ng, dNg = rt.lagrangian_tabulate(fm.ptype, 1, xq)
    fld = capi.Field(N, dN, V.ncomp, fs.cell_dof_ids, 0, None, dirichlet_values)
    return capi.Problem(fm.node_coordinates, fm.cell_node_ids, w, Ng, dNg, [fld], form_mat, form_vec, params, fq, None, 0, lift, V.nfree, V.nfree)


@pytest.mark.parametrize("order", [1, 2])
def test_boundary_triangulation_facets_and_measure(order):
    # BoundaryTriangulation(model; tags) (src/Geometry/BoundaryTriangulations.jl:194-203): facets, facet DoFs, surface measure
    model = g.CartesianDiscreteModel((0, 2, 0, 1, 0, 3), (3, 2, 4))
    X = model.node_coordinates
    G = g.BoundaryTriangulation(model)
    assert G.num_cells() == 2 * (3 * 2 + 3 * 4 + 2 * 4)
    top = g.BoundaryTriangulation(model, tags=[22])          # face z = 3 of the box (entity 22)
    assert top.num_cells() == 3 * 2 and np.all(X[top.model.cell_node_ids - 1][:, :, 2] == 3.0)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, order))
    fs = top.restrict(V)
    assert fs.cell_dof_ids.shape == (6, 4 if order == 1 else 9) and (fs.cell_dof_ids > 0).all()
    fx = V.dof_coordinates()[0]
    assert np.all(fx[fs.cell_dof_ids - 1][:, :, 2] == 3.0)    # every facet DoF lies on the facet
    b = _facet_problem(top, V, 2 * order, params=[1.0]).assemble_vector()
    assert abs(b.sum() - 2.0) < 1e-13                         # sum_i int_Gamma N_i = |Gamma| = 2 x 1
    # perturbed mesh, whole boundary of the unit cube: the facets are bilinear patches, the oracle's measure is |t1 x t2|
    bw = _facet_problem(G, V, 2 * order, params=[1.0]).assemble_vector()
    assert abs(bw.sum() - 2 * (2 * 1 + 2 * 3 + 1 * 3)) < 1e-12
    # 2-D: SEG facets
    m2 = g.CartesianDiscreteModel((0, 1, 0, 2), (4, 3))
    G2 = g.BoundaryTriangulation(m2)
    assert G2.num_cells() == 14 and G2.model.ptype == "SEG"


@pytest.mark.parametrize("order", [1, 2])
def test_boundary_triangulation_of_simplex_models(order):
    # TET model -> TRI facets, TRI model -> SEG facets; measure through the oracle
    model = g.simplexify(g.CartesianDiscreteModel((0, 1, 0, 2, 0, 1), (2, 3, 2)))
    G = g.BoundaryTriangulation(model)
    assert G.model.ptype == "TRI" and G.num_cells() == 2 * 2 * (2 * 3 + 2 * 2 + 3 * 2)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, order))
    assert G.restrict(V).cell_dof_ids.shape[1] == (3 if order == 1 else 6)
    b = _facet_problem(G, V, 2 * order, params=[1.0]).assemble_vector()
    assert abs(b.sum() - 2 * (1 * 2 + 1 * 1 + 2 * 1)) < 1e-12
    m2 = g.simplexify(g.CartesianDiscreteModel((0, 1, 0, 2), (3, 2)))
    G2 = g.BoundaryTriangulation(m2, tags=[8])            # right edge x = 1
    V2 = g.TestFESpace(m2, g.ReferenceFE(g.lagrangian, float, order))
    assert G2.model.ptype == "SEG" and G2.num_cells() == 2
    b2 = _facet_problem(G2, V2, 2 * order, params=[1.0]).assemble_vector()
    assert abs(b2.sum() - 2.0) < 1e-13


def test_julia_shim_binds_only_exported_symbols():
    # every `ccall((:gb200_x, LIB), ...)` of the Julia shim a maintainer would add (julia/GridapB200.jl) names an
    # entry point that include/gridap_b200.h declares and the built library exports
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = set()
    for f in ("julia/GridapB200.jl", "INTEGRATION.md"):
        names |= set(re.findall(r"\(:(gb200_[a-z_0-9]+),\s*LIB\)", open(os.path.join(root, f)).read()))
    assert len(names) >= 20
    header = open(os.path.join(root, "include", "gridap_b200.h")).read()
    L = lib.load()
    for n in sorted(names):
        assert re.search(r"\b%s\s*\(" % n, header), n
        assert hasattr(L, n), n


def test_host_matrix_containers():
    # SparseMatrixCSC / SparseMatrixCSR{Bi} / BlockMatrix / BlockVector as the assemblers return them (1-based Julia conventions)
    import scipy.sparse as sp
    rng = np.random.default_rng(5)
    D = sp.random(7, 5, density=0.4, random_state=3, format="csc")
    D.sort_indices()
    A = g.SparseMatrixCSC(7, 5, D.indptr.astype(np.int64) + 1, D.indices.astype(np.int64) + 1, D.data.copy())
    assert A.shape == (7, 5) and A.nnz() == D.nnz and np.array_equal(A.toarray(), D.toarray())
    i, j = int(D.indices[0]) + 1, 1
    assert A.getindex(i, j) == D[i - 1, j - 1] and A.getindex(7, 5) == D[6, 4]
    I, J, V = A.findnz()
    assert np.array_equal(sp.csc_matrix((V, (I - 1, J - 1)), shape=(7, 5)).toarray(), D.toarray())
    R = D.tocsr()
    R.sort_indices()
    for bi in (0, 1):
        T = g.SparseMatrixCSR[bi]
        assert T.Bi == bi and issubclass(T, g.SparseMatrixCSR)
        M = T(7, 5, R.indptr.astype(np.int64) + bi, R.indices.astype(np.int64) + bi, R.data.copy())
        assert np.array_equal(M.toarray(), D.toarray()) and M.nnz() == D.nnz
    with pytest.raises(ValueError):
        g.SparseMatrixCSR[2]
    B = g.BlockMatrix([[A, A], [A, A]])
    assert B.blocksize() == (2, 2) and B.shape == (14, 10) and B.nnz() == 4 * D.nnz
    assert np.array_equal(B.toarray(), np.block([[D.toarray()] * 2] * 2))
    v = g.BlockVector(rng.standard_normal(9), [4, 5])
    v.blocks[1][:] = 0.0                       # the blocks are views of one array
    assert len(v) == 9 and np.all(np.asarray(v)[4:] == 0.0) and np.all(np.asarray(v)[:4] != 0.0)
    # styles
    assert isinstance(g.BlockMultiFieldStyle(), g.BlockMultiFieldStyle)
    with pytest.raises(NotImplementedError):
        g.BlockMultiFieldStyle(2, (1, 1))


@pytest.mark.parametrize("simplex", [False, True])
@pytest.mark.parametrize("order", [1, 2])
def test_vector_valued_facet_spaces(simplex, order):
    # facet DoF tables of vector-valued spaces (component-major, k = a + nd*c): int_Gamma v.t sums to |Gamma| t per component
    model = g.CartesianDiscreteModel((0, 2, 0, 1, 0, 3), (2, 2, 3))
    if simplex:
        model = g.simplexify(model)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), order), dirichlet_tags=[21])
    G = g.BoundaryTriangulation(model, tags=[22])                      # face z = 3, area 2
    b = _facet_problem(G, V, 2 * order, params=[1.0, -2.0, 0.5]).assemble_vector()
    fx, fc, _, _ = V.dof_coordinates()
    assert np.allclose([b[fc == c].sum() for c in range(3)], [2.0, -4.0, 1.0], atol=1e-12)
    assert np.all(b[~np.isclose(fx[:, 2], 3.0)] == 0.0)                # nothing off the facet


def test_owned_column_ids_helper_matches_the_python_strategy():
    # gb200_owned_column_ids (host helper of the C ABI, no device): the same masking / renumbering as OwnedColumns.map_ids
    model = g.simplexify(g.CartesianDiscreteModel((0, 1) * 3, (3, 2, 4)))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 2), dirichlet_tags="boundary")
    Q = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1))
    Y = g.MultiFieldFESpace([V, Q])
    X = g.MultiFieldFESpace([g.TrialFESpace(V, (0.0, 0.0, 0.0)), g.TrialFESpace(Q)])
    world = 3
    seen = np.zeros(Y.nfree, dtype=int)
    for rank in range(world):
        part = gd.partition(model, X, Y, world, rank)
        owned = np.concatenate(part.owned)
        for k, ls in enumerate(part.local_space.spaces):
            ids = ls.cell_dof_ids.copy()
            ids[ids > 0] += Y.offsets[k]
            ref = part.strategy.map_ids(ls.cell_dof_ids, "cols", Y.offsets[k])
            out, oid = lib.owned_column_ids(ids, owned)
            assert np.array_equal(out, ref) and np.array_equal(oid, part.owned_ids)
        seen[part.owned_ids - 1] += 1
        # every cell that touches an owned DoF is local, and the local cells are ascending (serial summation order per column)
        assert np.all(np.diff(part.local_cells) > 0)
    assert np.all(seen == 1)   # every column has exactly one owner


# ---- round 2: host logic of skeleton triangulations, discontinuous / constrained / zero-mean spaces (no device call) ----------
def test_linear_constraints_numbering_matches_the_reference_goldens():
    # test/FESpacesTests/FESpacesWithLinearConstraintsTests.jl:30-59: n_fdofs == 6, n_fmdofs == 4, and the cell DoF values of the
    # constrained FE function with master values 1..4 / -1..-2
    model = g.CartesianDiscreteModel((0, 1, 0, 1), (2, 2))
    V = g.FESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags=[1, 2, 5])
    Vc = g.FESpaceWithLinearConstraints([1, 5, -2], [[-1, 4], [4, 6], [-1, -3]], [[0.5, 0.5]] * 3, V)
    assert g.has_constraints(Vc) and not g.has_constraints(V)
    assert (Vc.n_fdofs, Vc.n_fmdofs, Vc.num_free_dofs(), Vc.num_dirichlet_dofs()) == (6, 4, 4, 2)
    fv, dv = Vc.scatter_free_and_dirichlet_values(np.arange(1.0, 5.0), -np.arange(1.0, 3.0))
    ids = V.cell_dof_ids
    vals = np.concatenate([fv, dv])[np.where(ids > 0, ids - 1, V.nfree - ids - 1)]
    assert np.allclose(vals, [[-1.0, -1.5, 1.0, 1.0], [-1.5, -2.0, 1.0, 2.0], [1.0, 1.0, 3.0, 3.5], [1.0, 2.0, 3.5, 4.0]])
    # cell tables of master DoFs: every master of every DoF of the cell, once; the extended (unconstrained) numbering is positive
    tab = Vc.get_cell_dof_ids()
    assert tab.shape[0] == 4 and set(np.unique(tab)) <= set(range(-2, 5))
    assert Vc.extended.cell_dof_ids.min() >= 1 and Vc.extended.num_free_dofs() == V.nfree + V.ndirichlet
    with pytest.raises(ValueError):   # recursive constraints: a master that is itself a slave
        g.FESpaceWithLinearConstraints([1, 4], [[4, 6], [5, 6]], [[0.5, 0.5]] * 2, V)


def test_zero_mean_and_discontinuous_numbering():
    model = g.CartesianDiscreteModel((0, 1, 0, 1), (4, 4))
    # conformity = :L2 (src/FESpaces/DiscontinuousFESpaces.jl): cell after cell, no Dirichlet DoFs
    V = g.FESpace(model, g.ReferenceFE(g.lagrangian, float, 2), conformity="L2")
    assert V.num_free_dofs() == 16 * 9 and V.num_dirichlet_dofs() == 0
    assert np.array_equal(V.cell_dof_ids, np.arange(1, 16 * 9 + 1).reshape(16, 9))
    fx = V.dof_coordinates()[0]
    assert np.allclose(fx[:4], [[0, 0], [0.25, 0], [0, 0.25], [0.25, 0.25]])          # the vertices of cell 1 come first (Q2 local order)
    # constraint = :zeromean = FESpaceWithConstantFixed(space, true, num_free_dofs(space)) (FESpacesWithConstantFixed.jl:14-25,146-163)
    V0 = g.FESpace(model, g.ReferenceFE(g.lagrangian, float, 2), conformity="L2", constraint="zeromean")
    assert (V0.num_free_dofs(), V0.num_dirichlet_dofs()) == (16 * 9 - 1, 1)
    ref = V.cell_dof_ids.copy()
    ref[ref == 16 * 9] = -1
    assert np.array_equal(V0.cell_dof_ids, ref)
    # a space that already has Dirichlet DoFs is left alone (DoNotFixConstant)
    Vd = g.FESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags=[1], constraint="zeromean")
    assert Vd.num_dirichlet_dofs() == 1 and Vd._fixed_dof == 0
    with pytest.raises(NotImplementedError):
        g.FESpace(model, g.ReferenceFE(g.lagrangian, float, 1), conformity="L2", dirichlet_tags="boundary")


def test_skeleton_triangulation_topology_and_point_permutation():
    from gridap_b200 import reffes as rf
    from oracle import ref_skeleton as rs
    for model in (g.CartesianDiscreteModel((0, 1, 0, 1), (4, 3)), g.CartesianDiscreteModel((0, 1) * 3, (3, 2, 2)),
                  g.simplexify(g.CartesianDiscreteModel((0, 1) * 3, (2, 2, 2))), g.simplexify(g.CartesianDiscreteModel((0, 1, 0, 1), (3, 3)))):
        L = g.SkeletonTriangulation(model)
        want = rs.interior_facets(model.cell_node_ids, model.ptype)        # the oracle's line-by-line facet sweep
        got = list(zip(L.cells_plus.tolist(), L.lfaces_plus.tolist(), L.cells_minus.tolist(), L.lfaces_minus.tolist()))
        assert got == [tuple(int(v) for v in w) for w in want]
        assert np.all(L.cells_plus < L.cells_minus)                       # plus = the first incident cell
        pts, wf, nref = rf.facet_glue(model.ptype, 3)
        perm = L.point_permutation(pts)
        assert perm.shape == (L.num_cells(), len(wf)) and np.all(np.sort(perm, axis=1) == np.arange(len(wf)))
    # Cartesian 2D: (nx - 1) ny + nx (ny - 1) interior facets
    assert g.SkeletonTriangulation(g.CartesianDiscreteModel((0, 1, 0, 1), (4, 3))).num_cells() == 3 * 3 + 4 * 2


def test_recogniser_of_jump_and_mean_terms():
    from gridap_b200 import celldata as cd
    model = g.CartesianDiscreteModel((0, 1, 0, 1), (2, 2))
    V = g.FESpace(model, g.ReferenceFE(g.lagrangian, float, 1), conformity="L2")
    v, u = g.get_fe_basis(V), g.get_trial_fe_basis(V)
    n = g.get_normal_vector(g.SkeletonTriangulation(model))
    # test/GridapTests/PoissonDGTests.jl:42-45 and FESpacesWithLinearConstraintsTests.jl:77 (jump(u)*jump(v))
    e = 40.0 * g.dot(g.jump(v * n), g.jump(u * n)) - g.dot(g.jump(v * n), g.mean(g.grad(u))) - g.dot(g.mean(g.grad(v)), g.jump(u * n)) \
        + g.jump(u) * g.jump(v) + 0.5 * g.jump(g.dot(n, g.grad(v))) * g.jump(g.dot(n, g.grad(u)))
    terms = cd.recognise_matrix(cd._wrap(e))
    assert [t.params for t in terms] == [(40.0, 0, 1.0, -1.0, 0, 1.0, -1.0), (-1.0, 0, 1.0, -1.0, 1, 0.5, 0.5), (-1.0, 1, 0.5, 0.5, 0, 1.0, -1.0),
                                         (1.0, 0, 1.0, -1.0, 0, 1.0, -1.0), (0.5, 1, 1.0, -1.0, 1, 1.0, -1.0)]
    assert all(t.form == lib.FORM_SKELETON and t.glued == "skeleton" for t in terms)
    with pytest.raises(NotImplementedError):   # a vector along the normal against a scalar
        cd.recognise_matrix(cd._wrap(g.jump(v * n) * g.jump(u)))


# ------------------------------------------------------------------------------------------------ order 3, views, graddiv
def _rotated_cells_model(part, seed):
    """a Cartesian mesh whose cells list their vertices in randomly rotated local frames (proper rotations of the n-cube): shared
    edges / faces are then seen with different vertex orders by their cells, which is what the own-DoF permutations are for"""
    import itertools
    D = len(part)
    m = g.CartesianDiscreteModel((0, 1) * D, part)
    rots = []
    for pa in itertools.permutations(range(D)):
        for fl in itertools.product([0, 1], repeat=D):
            if np.linalg.det(np.eye(D)[list(pa)]) * (-1) ** sum(fl) > 0:
                relabel = []
                for v in range(2 ** D):
                    o = [0] * D
                    for d in range(D):
                        o[pa[d]] = ((v >> d) & 1) ^ fl[d]
                    relabel.append(sum(o[d] << d for d in range(D)))
                rots.append(relabel)
    rng = np.random.default_rng(seed)
    cells = m.cell_node_ids.copy()
    for c in range(len(cells)):
        cells[c] = cells[c][rots[rng.integers(len(rots))]]
    return g.DiscreteModel(m.node_coordinates, cells, m.ptype)


def _assert_conforming(model, V):
    """every DoF id names ONE (physical node, component) whichever cell it is read from, and different ids different ones"""
    from gridap_b200 import reffes as rf
    nodes = rf.reference_nodes(model.ptype, V.order)
    Ng, _ = rf.tabulate_lagrangian(model.ptype, 1, nodes)
    P = np.einsum("av,cvx->cax", Ng, model.node_coordinates[model.cell_node_ids.astype(np.int64) - 1])
    nl = len(nodes)
    seen = {}
    for c in range(model.num_cells()):
        for comp in range(V.ncomp):
            for a in range(nl):
                val = (comp,) + tuple(np.round(P[c, a], 10))
                assert seen.setdefault(int(V.cell_dof_ids[c, a + nl * comp]), val) == val
    assert len(set(seen.values())) == len(seen) == V.nfree + V.ndirichlet


@pytest.mark.parametrize("part,simplex", [((3, 4), False), ((3, 2), True), ((2, 3, 2), False), ((2, 2, 2), True)])
def test_order3_numbering_matches_oracle(part, simplex):
    D = len(part)
    m = g.CartesianDiscreteModel((0, 1) * D, part)
    if simplex:
        m = g.simplexify(m)
    X, cells, ptype = problems.cartesian_mesh((0, 1) * D, part, simplex)
    for ncomp in (1, D):
        for tags in ([], ["boundary"], [5, 6] if D == 2 else [21, 22]):
            masks = None
            if ncomp > 1 and len(tags) == 2:
                masks = [[True, False, True][:ncomp], [False, True, True][:ncomp]]
            T = float if ncomp == 1 else g.VectorValue(D)
            V = g.FESpace(m, g.ReferenceFE(g.lagrangian, T, 3), dirichlet_tags=tags, dirichlet_masks=masks)
            cdofs, nf, ndr = problems.lagrangian_space(part, cells, ptype, 3, ncomp, tags, masks, nnodes=len(X))
            assert (nf, ndr) == (V.nfree, V.ndirichlet) and np.array_equal(cdofs, V.cell_dof_ids)
            if ncomp == 1 and not tags:
                _assert_conforming(m, V)
                assert V.nfree == np.prod([3 * p + 1 for p in part])
                # the DoF nodes (interpolation points): a linear function is reproduced at every DoF of every cell
                f = V.interpolate_free_values(lambda x: 1.0 + x @ np.arange(1, D + 1))
                from gridap_b200 import reffes as rf
                Ng, _ = rf.tabulate_lagrangian(m.ptype, 1, rf.reference_nodes(m.ptype, 3))
                P = np.einsum("av,cvx->cax", Ng, m.node_coordinates[m.cell_node_ids.astype(np.int64) - 1])
                assert np.allclose(f[V.cell_dof_ids - 1], 1.0 + P @ np.arange(1, D + 1), atol=1e-12)


@pytest.mark.parametrize("part", [(3, 3), (2, 2, 2)])
def test_order3_numbering_on_rotated_cells(part):
    # non-identity permutations of edges AND of quadrilateral faces: the product (lattice weights carried to the face's frame)
    # against the oracle (the reference's permutation tables), plus the geometric meaning of conformity
    D = len(part)
    nonid = 0
    for seed in range(3):
        m = _rotated_cells_model(part, seed)
        cells = [list(map(int, r)) for r in m.cell_node_ids]
        for d in range(1, D):
            c2f, fv = rn.global_faces_oriented(cells, m.ptype, d)
            nonid += int((rn.cell_permutations(cells, m.ptype, d, c2f, fv) != 1).sum())
        for ncomp in (1, D):
            V = g.FESpace(m, g.ReferenceFE(g.lagrangian, float if ncomp == 1 else g.VectorValue(D), 3))
            cdofs, nf, _ = rn.conforming_dofs(cells, m.ptype, 3, ncomp, None, [])
            assert nf == V.nfree and np.array_equal(cdofs, V.cell_dof_ids)
            _assert_conforming(m, V)
    assert nonid > 10


def test_order3_tabulation():
    from gridap_b200 import reffes as rf
    rng = np.random.default_rng(0)
    for ptype, D in (("SEG", 1), ("QUAD", 2), ("HEX", 3), ("TRI", 2), ("TET", 3)):
        assert np.allclose(rf.reference_nodes(ptype, 3), rt.lagrangian_nodes(ptype, 3), atol=1e-15)
        lat, own = rf.lagrangian_lattice(ptype, 3)
        _, fo = rt.lagrangian_nodes_and_face_own_nodes(ptype, 3)
        assert [list(r + 1) for d in sorted(own) for r in own[d]] == [list(a) for a in fo]
        x = rng.uniform(0, 1, (9, D)) / (1 if ptype in ("SEG", "QUAD", "HEX") else D + 1)
        N, dN = rf.tabulate_lagrangian(ptype, 3, x)
        N2, dN2 = rt.lagrangian_tabulate(ptype, 3, x)
        # (the oracle inverts the monomial Vandermonde matrix like the reference: ~1e-12 on the 64 x 64 Q3 hexahedron)
        assert np.abs(N - N2).max() < 1e-11 and np.abs(dN - dN2).max() < 1e-10
        Nn, _ = rf.tabulate_lagrangian(ptype, 3, rf.reference_nodes(ptype, 3))
        assert np.allclose(Nn, np.eye(len(Nn)), atol=1e-13)
        assert np.allclose(N.sum(axis=1), 1.0, atol=1e-13) and np.allclose(dN.sum(axis=1), 0.0, atol=1e-12)
        # exact for cubics: the interpolant of a cubic is the cubic
        f = lambda y: (y ** 3).sum(axis=1) + y[:, 0] * y[:, -1] ** 2 - 2.0 * y[:, 0]   # noqa: E731
        assert np.allclose(N @ f(rf.reference_nodes(ptype, 3)), f(x), atol=1e-13)
    with pytest.raises(NotImplementedError):
        g.ReferenceFE(g.lagrangian, float, 4)


def test_view_triangulation_host_logic():
    m = g.UnstructuredDiscreteModel(g.CartesianDiscreteModel((0, 1) * 2, (4, 3)))
    V = g.FESpace(m, g.ReferenceFE(g.lagrangian, float, 2))
    view = g.Triangulation(m, np.arange(1, 7))            # collect(1:div(n,2)), 1-based like the reference
    assert view.num_cells() == 6 and np.array_equal(view.model.cell_node_ids, m.cell_node_ids[:6])
    assert np.array_equal(view.restrict(V).get_cell_dof_ids(), V.cell_dof_ids[:6])
    assert view.restrict(V).num_free_dofs() == V.nfree
    mask = np.zeros(12, dtype=bool)
    mask[[1, 5]] = True
    assert np.array_equal(g.Triangulation(m, mask).cells, [1, 5])
    with pytest.raises(ValueError):
        g.Triangulation(m, [0, 1])                          # 1-based ids
    with pytest.raises(NotImplementedError):
        g.FESpace(view, g.ReferenceFE(g.lagrangian, float, 1))
    other = g.CartesianDiscreteModel((0, 1) * 2, (4, 3))
    with pytest.raises(ValueError):
        g.Triangulation(other, [1, 2]).restrict(V)
    # a bilinear form over two different views has no pattern owner on this path
    from gridap_b200 import assemblers as asm
    d1, d2 = g.Measure(view, 2), g.Measure(g.Triangulation(m, [7, 8]), 2)
    u, v = cd.Basis("trial", V), cd.Basis("test", V)
    with pytest.raises(NotImplementedError):
        asm.collect_cell_matrix(V, V, g.Integral(u * v) * d1 + g.Integral(u * v) * d2)
    # bulk first, then the view, whatever the order in the form
    dO = g.Measure(g.Triangulation(m), 2)
    md = asm.collect_cell_matrix(V, V, g.Integral(u * v) * d1 + g.Integral(u * v) * dO)
    assert md.measure is dO and [e.measure for e in md.extra] == [d1]


def test_graddiv_is_the_lambda_part_of_elasticity():
    # benchmark/bm/bm_assembly.jl:9: graddiv(u,v,dΩ) = ∫((∇⋅u)⋅(∇⋅v))dΩ
    m = g.CartesianDiscreteModel((0, 1) * 2, (2, 2))
    V = g.FESpace(m, g.ReferenceFE(g.lagrangian, g.VectorValue(2), 1))
    u, v = cd.Basis("trial", V), cd.Basis("test", V)
    terms = cd.recognise_matrix(2.0 * (g.div(u) * g.div(v)))
    assert len(terms) == 1 and terms[0].form == lib.FORM_ELASTICITY and tuple(terms[0].params) == (2.0, 0.0)
    terms = cd.recognise_matrix(g.dot(g.div(v), g.div(u)))
    assert terms[0].form == lib.FORM_ELASTICITY and tuple(terms[0].params) == (1.0, 0.0)


@pytest.mark.parametrize("D,n", [(2, 10), (3, 6)])
def test_bm_protocol_spaces_match_oracle(D, n):
    # the spaces of benchmark/bm/bm_assembly.jl:27-41 (no Dirichlet tags), orders 1-3, scalar and vector-valued: host numbering vs the
    # oracle's loop-by-loop restatement at the benchmark's own sizes
    part = (n,) * D
    m = g.UnstructuredDiscreteModel(g.CartesianDiscreteModel((0, 1) * D, part))
    X, cells, ptype = problems.cartesian_mesh((0, 1) * D, part)
    for order in (1, 2, 3):
        for ncomp in (1, D):
            V = g.TestFESpace(m, g.ReferenceFE(g.lagrangian, float if ncomp == 1 else g.VectorValue(D), order))
            cdofs, nf, ndr = problems.lagrangian_space(part, cells, ptype, order, ncomp, [], None, nnodes=len(X))
            assert (nf, ndr) == (V.nfree, 0) == (ncomp * (order * n + 1) ** D, 0)
            assert np.array_equal(cdofs, V.cell_dof_ids)
            view = g.Triangulation(m, np.arange(1, n ** D // 2 + 1))
            assert np.array_equal(view.restrict(V).get_cell_dof_ids(), cdofs[: n ** D // 2])
