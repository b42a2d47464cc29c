"""Pins the CPU oracle against every golden value the reference's own tests hold for the path
(SURVEY.md section 8c).  Each test cites the reference test file:line it restates."""
import numpy as np

from oracle import capi, problems
from oracle import ref_numbering as rn
from oracle import ref_tabulation as rt


def test_cartesian_grid_nodes_and_cells():
    # test/GeometryTests/CartesianGridsTests.jl:24-26,52-61
    domain = (0.0, 1.0, -1.0, 2.0)
    partition = (3, 4)
    x = rn.cartesian_node_coordinates(domain, partition)
    assert len(x) == 20
    assert np.allclose(x[13 - 1], (0.0, 1.25)) and tuple(x[13 - 1]) == (0.0, 1.25)
    assert tuple(x[4 - 1]) == (1.0, -1.0)
    assert tuple(x[0]) == (0.0, -1.0)
    assert tuple(x[-1]) == (1.0, 2.0)
    t = rn.cartesian_cell_node_ids(partition)
    assert list(t[0]) == [1, 2, 5, 6]
    assert list(t[11 - 1]) == [14, 15, 18, 19]


def test_polytope_face_tables():
    # test/ReferenceFEsTests/ExtrusionPolytopesTests.jl:16-26 (QUAD), SURVEY App. B (HEX)
    assert rn.ncube_face_vertices(2, 1) == [[1, 2], [3, 4], [1, 3], [2, 4]]
    assert rn.ncube_face_vertices(3, 1) == [[1, 2], [3, 4], [5, 6], [7, 8], [1, 3], [2, 4], [5, 7], [6, 8], [1, 5], [2, 6], [3, 7], [4, 8]]
    assert rn.ncube_face_vertices(3, 2) == [[1, 2, 3, 4], [5, 6, 7, 8], [1, 2, 5, 6], [3, 4, 7, 8], [1, 3, 5, 7], [2, 4, 6, 8]]
    # num_entities == 27 for a 3-D Cartesian model (test/GeometryTests/CartesianDiscreteModelsTests.jl:28-29)
    assert len(rn.ncube_faces(3)) == 27


def test_simplexify_hex():
    # test/ReferenceFEsTests/ExtrusionPolytopesTests.jl:129-133
    assert rn.HEX_TO_TETS == [[1, 2, 3, 7], [1, 2, 5, 7], [2, 3, 4, 7], [2, 4, 7, 8], [2, 5, 6, 7], [2, 6, 7, 8]]
    t = rn.simplexify(np.array([[1, 2, 3, 4, 5, 6, 7, 8]]), "HEX")
    assert t.shape == (6, 4)


def _space_2x2(ncomp, tags, masks):
    partition = (2, 2)
    cells = rn.cartesian_cell_node_ids(partition)
    n2t = problems.node_tags(partition, 9, tags)
    nd, nfree, ndiri, d2n, d2c = rn.clagrangian_dofs(n2t, masks, ncomp)
    return rn.clagrangian_cell_dofs(cells, nd), nd, nfree, ndiri, d2n, d2c


def test_clagrangian_cell_dof_ids():
    # test/FESpacesTests/CLagrangianFESpacesTests.jl:23-26 (no tags: dofs == node ids)
    cd, nd, nfree, ndiri, _, _ = _space_2x2(1, [], [])
    assert (cd == rn.cartesian_cell_node_ids((2, 2))).all() and nfree == 9 and ndiri == 0
    # :40-43 vector valued, no tags
    cd, nd, *_ = _space_2x2(2, [], [])
    assert cd.tolist() == [[1, 3, 7, 9, 2, 4, 8, 10], [3, 5, 9, 11, 4, 6, 10, 12], [7, 9, 13, 15, 8, 10, 14, 16], [9, 11, 15, 17, 10, 12, 16, 18]]
    assert nd.tolist() == [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16], [17, 18]]
    # :52-63 scalar with tags and masks
    tags = [1, 2, 4, 5, 8]
    cd, nd, nfree, ndiri, d2n, d2c = _space_2x2(1, tags, [True, True, False, True, True])
    assert cd.tolist() == [[-1, -2, 1, 2], [-2, -3, 2, -4], [1, 2, 3, 4], [2, -4, 4, 5]]
    assert nd[:, 0].tolist() == [-1, -2, -3, 1, 2, -4, 3, 4, 5]
    assert d2n == [1, 2, 3, 6] and d2c == [1, 1, 1, 1]
    # :65-72 vector with component masks
    masks2 = [(True, True), (True, False), (False, False), (False, True), (True, True)]
    cd, nd, nfree, ndiri, d2n, d2c = _space_2x2(2, tags, masks2)
    assert cd.tolist() == [[-1, 1, 3, 5, -2, -3, 4, 6], [1, -4, 5, -5, -3, 2, 6, -6], [3, 5, 7, 9, 4, 6, 8, 10], [5, -5, 9, 11, 6, -6, 10, 12]]
    assert nd.tolist() == [[-1, -2], [1, -3], [-4, 2], [3, 4], [5, 6], [-5, -6], [7, 8], [9, 10], [11, 12]]
    assert d2n == [1, 1, 2, 3, 6, 6] and d2c == [1, 2, 2, 1, 1, 2]
    # :85-101 factory paths
    cd, nd, *_ = _space_2x2(1, tags, [True] * 5)
    assert nd[:, 0].tolist() == [-1, -2, -3, 1, 2, -4, 3, 4, -5]
    cd, nd, *_ = _space_2x2(2, tags, [(True, True)] * 5)
    assert nd.tolist() == [[-1, -2], [-3, -4], [-5, -6], [1, 2], [3, 4], [-7, -8], [5, 6], [7, 8], [-9, -10]]


def test_csc_builder_protocol():
    # test/AlgebraTests/AlgebraInterfacesTests.jl:106-152 (MinMemory CSC builder)
    a = capi.Builder(6, 9)
    for (i, j) in [(1, 1), (1, 1), (3, 1), (2, 1), (4, 9)]:
        a.count(i, j)
    assert a.colnnzmax.tolist() == [4, 0, 0, 0, 0, 0, 0, 0, 1]
    a.allocate()
    colptr, colnnz = a.state()
    assert colnnz.tolist() == [0] * 9
    assert colptr.tolist() == [1, 5, 5, 5, 5, 5, 5, 5, 5, 6]
    a.add(1.0, 1, 1)
    a.add(None, 1, 1)
    a.add(4.0, 3, 1)
    a.add(2.0, 3, 1)
    a.add(8.0, 2, 1)
    a.add(3.0, 4, 9)
    _, colnnz = a.state()
    assert colnnz.tolist() == [3, 0, 0, 0, 0, 0, 0, 0, 1]
    colptr, rowval, nzval = a.finish()
    J = np.repeat(np.arange(1, 10), np.diff(colptr))
    assert rowval.tolist() == [1, 2, 3, 4] and J.tolist() == [1, 1, 1, 9] and nzval.tolist() == [1.0, 8.0, 6.0, 3.0]
    # :143-151 negative ids are skipped by add_entries!
    a = capi.Builder(6, 9)
    for (i, j) in [(1, -1), (-1, -1), (1, 1), (-1, 1), (1, 1), (1, -1), (1, 1), (1, -1)]:
        a.count(i, j)
    assert a.colnnzmax.tolist() == [3, 0, 0, 0, 0, 0, 0, 0, 0]


def _poisson_2x2():
    # test/FESpacesTests/SparseMatrixAssemblersTests.jl:16-40: 2x2 Q1, dirichlet_tags=[1,2,3,4,6,5], degree 2, b(x)=x[2]
    pb = problems.single_field_problem((0, 1, 0, 1), (2, 2), order=1, degree=2, dirichlet_tags=[1, 2, 3, 4, 6, 5],
                                       form_mat=capi.LAPLACIAN, form_vec=capi.SOURCE)
    xq = pb.quadrature_points()
    fq = xq[:, :, 1].copy()
    return problems.single_field_problem((0, 1, 0, 1), (2, 2), order=1, degree=2, dirichlet_tags=[1, 2, 3, 4, 6, 5],
                                         form_mat=capi.LAPLACIAN, form_vec=capi.SOURCE, fq=fq, lift=True)


def test_sparse_matrix_assembler_golden():
    # test/FESpacesTests/SparseMatrixAssemblersTests.jl:104-152
    pb = _poisson_2x2()
    assert pb.nfree == 3
    colptr, rowval, nzval, vec = pb.assemble(with_vector=True)
    A = problems.csc_to_dense(colptr, rowval, nzval, 3, 3)
    assert np.allclose(vec, [0.0625, 0.125, 0.0625], rtol=0, atol=1e-14)
    assert abs(A[0, 0] - 1.333333333333333) < 1e-14
    assert abs(A[1, 0] + 0.33333333333333) < 1e-13
    assert abs(A[0, 1] + 0.33333333333333) < 1e-13
    assert abs(A[1, 1] - 2.666666666666666) < 1e-14
    assert abs(A[2, 1] + 0.33333333333333) < 1e-13
    assert abs(A[1, 2] + 0.33333333333333) < 1e-13
    assert abs(A[2, 2] - 1.333333333333333) < 1e-14
    # in-place re-assembly twice gives the same (assemble_matrix_and_vector! x2, :124-130)
    nz2 = nzval.copy()
    b2 = vec.copy()
    pb.assemble_inplace(colptr, rowval, nz2, b2, add=False)
    pb.assemble_inplace(colptr, rowval, nz2, b2, add=False)
    assert np.array_equal(nz2, nzval) and np.array_equal(b2, vec)
    # rows sorted & unique per column (canonical CSC)
    for j in range(3):
        r = rowval[colptr[j] - 1:colptr[j + 1] - 1]
        assert (np.diff(r) > 0).all()


def test_attach_dirichlet():
    # test/CellDataTests/AttachDirichletTests.jl:13-31: (mat, vec - mat*vals) on Dirichlet cells only
    dv = np.array([1.0, 2.0, 3.0, 4.0, 5.0, 6.0])
    pb0 = problems.single_field_problem((0, 1, 0, 1), (2, 2), order=1, degree=2, dirichlet_tags=[1, 2, 3, 4, 6, 5],
                                        form_mat=capi.LAPLACIAN, form_vec=capi.SOURCE, params=[1.0], dirichlet_values=dv, lift=False)
    pb1 = problems.single_field_problem((0, 1, 0, 1), (2, 2), order=1, degree=2, dirichlet_tags=[1, 2, 3, 4, 6, 5],
                                        form_mat=capi.LAPLACIAN, form_vec=capi.SOURCE, params=[1.0], dirichlet_values=dv, lift=True)
    for cell in range(4):
        K0, b0 = pb0.cell_local(cell)
        K1, b1 = pb1.cell_local(cell)
        ids = pb0.cell_dofs[cell]
        vals = np.array([dv[-i - 1] if i < 0 else 0.0 for i in ids])
        assert np.array_equal(K0[0][0], K1[0][0])
        assert np.allclose(b1[0], b0[0] - K0[0][0] @ vals, rtol=0, atol=1e-15)


def test_quadrature_weights_sum_to_measure():
    # test/CellDataTests/CellQuadraturesTests.jl:47-62
    for D in (2, 3):
        for degree in (1, 2, 3, 4):
            x, w = rt.tensor_quadrature(D, degree)
            assert abs(w.sum() - 1.0) < 1e-14 and len(w) == (degree // 2 + 1) ** D
    for degree in (1, 2, 3, 4):
        x, w = rt.wv_tet_quadrature(degree)
        assert abs(w.sum() - 1.0 / 6.0) < 1e-15
        # exactness on monomials of total degree <= degree: int x^a y^b z^c = a! b! c! / (a+b+c+3)!
        from math import factorial as f
        for a in range(degree + 1):
            for b in range(degree + 1 - a):
                for c in range(degree + 1 - a - b):
                    ex = f(a) * f(b) * f(c) / f(a + b + c + 3)
                    assert abs((w * x[:, 0] ** a * x[:, 1] ** b * x[:, 2] ** c).sum() - ex) < 1e-14
    assert len(rt.wv_tet_quadrature(4)[1]) == 14


def test_lagrangian_basis_kronecker_and_q2_layout():
    # test/ReferenceFEsTests/CLagrangianRefFEsTests.jl:121-129 (Q2 has 27 nodes: 8 + 12 + 6 + 1)
    for ptype, order, n in (("HEX", 1, 8), ("HEX", 2, 27), ("TET", 1, 4), ("TET", 2, 10), ("QUAD", 2, 9)):
        nodes = rt.lagrangian_nodes(ptype, order)
        assert len(nodes) == n
        N, dN = rt.lagrangian_tabulate(ptype, order, nodes)
        assert np.allclose(N, np.eye(n), atol=1e-12)
        assert np.allclose(dN.sum(axis=1), 0.0, atol=1e-11)  # partition of unity


def test_conforming_space_counts():
    # test/FESpacesTests/ConformingFESpacesTests.jl:43-46,66-69 style counts:
    # Q2 scalar on a 2x2 model has (2*2+1)^2 = 25 dofs; with boundary Dirichlet 9 free
    X, cells, ptype = problems.cartesian_mesh((0, 1, 0, 1), (2, 2))
    cd, nfree, ndiri = problems.lagrangian_space((2, 2), cells, ptype, 2, 1, "boundary", nnodes=len(X))
    assert nfree == 9 and ndiri == 16
    cd, nfree, ndiri = problems.lagrangian_space((2, 2), cells, ptype, 2, 1, [], nnodes=len(X))
    assert nfree == 25 and ndiri == 0 and sorted(set(cd.ravel())) == list(range(1, 26))
    # 3-D: Q2 on 2x2x2 -> 125 dofs, P2 on the simplexified 2x2x2 -> same 125 nodes (vertices + edges incl. diagonals)
    X, cells, ptype = problems.cartesian_mesh((0, 1, 0, 1, 0, 1), (2, 2, 2))
    cd, nfree, ndiri = problems.lagrangian_space((2, 2, 2), cells, ptype, 2, 1, [], nnodes=len(X))
    assert nfree == 125
    cd, nfree, ndiri = problems.lagrangian_space((2, 2, 2), cells, ptype, 2, 3, "boundary", nnodes=len(X))
    assert nfree == 3 * 27 and ndiri == 3 * 98


def test_manufactured_poisson_solution():
    # test/GridapTests/PoissonTests.jl style: exact for u = x + 2y (Q1 reproduces linears), f = 0, u on the boundary
    n = 4
    X = rn.cartesian_node_coordinates((0, 1, 0, 1), (n, n))
    pb = problems.single_field_problem((0, 1, 0, 1), (n, n), form_mat=capi.LAPLACIAN, form_vec=capi.SOURCE, params=[0.0])
    u = X[:, 0] + 2 * X[:, 1]
    n2t = problems.node_tags((n, n), len(X), ["boundary"])
    nd, nfree, ndiri, d2n, _ = rn.clagrangian_dofs(n2t, [True], 1)
    dv = np.array([u[k - 1] for k in d2n])
    pb = problems.single_field_problem((0, 1, 0, 1), (n, n), form_mat=capi.LAPLACIAN, form_vec=capi.SOURCE, params=[0.0],
                                       dirichlet_values=dv, lift=True)
    colptr, rowval, nzval, b = pb.assemble(with_vector=True)
    A = problems.csc_to_dense(colptr, rowval, nzval, nfree, nfree)
    x = np.linalg.solve(A, b)
    free_nodes = [k for k in range(len(X)) if nd[k, 0] > 0]
    assert np.allclose(x, u[free_nodes], atol=1e-12)


def test_stokes_taylor_hood_manufactured_solution():
    """test/GridapTests/StokesTaylorHoodTests.jl:6-80 replayed through the oracle: Q2/Q1 on the 3x3 mesh of (0,2)^2, velocity
    Dirichlet on tags [1,2,5], l((v,q)) = int(v.f + q*g)dOmega + int(v.(n.grad u) - (n.v)p)dGamma on tags [6,7,8], degree 2;
    the reference asserts ||u-uh||_L2, ||u-uh||_H1, ||p-ph||_L2 < 1e-9.  The manufactured solution lies in the discrete spaces, so
    the discrete solution is its interpolant: this pins the Stokes blocks, the per-field sources, the lifting of the inhomogeneous
    velocity data and the facet source term by an answer the reference holds."""
    import gridap_b200 as g
    from parity_helpers import facet_problem
    domain, part, tags = (0, 2, 0, 2), (3, 3), [1, 2, 5]
    model = g.CartesianDiscreteModel(domain, part)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(2), 2), dirichlet_tags=tags)
    Q = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1))
    X, cells, ptype = problems.cartesian_mesh(domain, part)
    vd, nfu, ndu = problems.lagrangian_space(part, cells, ptype, 2, 2, tags, None, nnodes=len(X))
    pd, nfp, ndp = problems.lagrangian_space(part, cells, ptype, 1, 1, [], None, nnodes=len(X))
    assert np.array_equal(vd, V.cell_dof_ids) and np.array_equal(pd, Q.cell_dof_ids) and (nfu, nfp) == (V.nfree, Q.nfree)

    def u(x):
        return np.stack([x[:, 0] ** 2 + 2 * x[:, 1] ** 2, -x[:, 0] ** 2], axis=1)

    def p(x):
        return x[:, 0] + 3 * x[:, 1]

    f = [-6.0 + 1.0, 2.0 + 3.0]                       # -Laplace(u) + grad(p)
    dv = V.interpolate_dirichlet_values(u)
    vd2, pd2 = rn.multifield_cell_dofs([vd, pd], [nfu, nfp])
    xq, w = rt.quadrature(ptype, 2)
    N2, dN2 = rt.lagrangian_tabulate(ptype, 2, xq)
    N1, dN1 = rt.lagrangian_tabulate(ptype, 1, xq)
    touched = np.array([[1, 1], [1, 0]], dtype=np.uint8)
    geo = capi.Problem(X, cells, w, N1, dN1, [capi.Field(N1, dN1, 1, pd, 0)], capi.MASS, 0, None, None, None, 0, False, nfp, nfp)
    xphys = geo.quadrature_points()                   # [nc, np, 2]
    gq = 2.0 * xphys[:, :, 0:1]                       # g = div(u) = 2x
    fu = capi.Field(N2, dN2, 2, vd2, 0, None, dv, src=f)
    fp = capi.Field(N1, dN1, 1, pd2, nfu, fq=gq)
    pb = capi.Problem(X, cells, w, N1, dN1, [fu, fp], capi.STOKES, capi.SOURCE, None, None, touched, 0, True, nfu + nfp, nfu + nfp)
    colptr, rowval, nzval, b = pb.assemble(with_vector=True)
    # boundary term on the Neumann tags: t = n.grad(u) - p n at the facet quadrature points
    G = g.BoundaryTriangulation(model, tags=[6, 7, 8])
    fpb = facet_problem(G, V, 2)
    xf = fpb.quadrature_points()                      # [nfacets, np, 2]
    mid = xf.mean(axis=1)
    n = np.zeros_like(mid)
    n[np.isclose(mid[:, 0], 0.0)] = (-1.0, 0.0)
    n[np.isclose(mid[:, 0], 2.0)] = (1.0, 0.0)
    n[np.isclose(mid[:, 1], 2.0)] = (0.0, 1.0)
    assert np.all(np.abs(n).sum(axis=1) == 1.0)
    x, y = xf[..., 0], xf[..., 1]
    gu = np.stack([np.stack([2 * x, -2 * x], axis=-1), np.stack([4 * y, 0 * y], axis=-1)], axis=-2)   # gu[..., i, j] = d_i u_j
    t = np.einsum("fi,fpij->fpj", n, gu) - (x + 3 * y)[..., None] * n[:, None, :]
    b[:nfu] += facet_problem(G, V, 2, fq=t).assemble_vector()
    A = problems.csc_to_dense(colptr, rowval, nzval, nfu + nfp, nfu + nfp)
    sol = np.linalg.solve(A, b)
    fx, fc, _, _ = V.dof_coordinates()
    assert np.abs(sol[:nfu] - u(fx)[np.arange(nfu), fc]).max() < 1e-9
    px = Q.dof_coordinates()[0]
    assert np.abs(sol[nfu:] - p(px)).max() < 1e-9


def test_vertex_and_own_node_permutations_reference_goldens():
    # test/ReferenceFEsTests/ExtrusionPolytopesTests.jl:33-36,58-59,74-75
    assert rn.vertex_permutations("QUAD") == [[1, 2, 3, 4], [1, 3, 2, 4], [2, 1, 4, 3], [2, 4, 1, 3],
                                              [3, 1, 4, 2], [3, 4, 1, 2], [4, 2, 3, 1], [4, 3, 2, 1]]
    assert rn.vertex_permutations("SEG") == [[1, 2], [2, 1]]
    assert rn.vertex_permutations("TRI") == [[1, 2, 3], [1, 3, 2], [2, 1, 3], [2, 3, 1], [3, 1, 2], [3, 2, 1]]
    lin = lambda fp: (lambda x: rt.lagrangian_tabulate(fp, 1, x)[0])   # noqa: E731
    # test/ReferenceFEsTests/CLagrangianRefFEsTests.jl:112-114: SEGMENT of order 4
    assert rn.own_nodes_permutations("SEG", rt.interior_nodes("SEG", 4), lin("SEG")) == [[1, 2, 3], [3, 2, 1]]
    # :117-119: QUAD with orders (2, 3): the own nodes (1/2, 1/3), (1/2, 2/3); 0 = INVALID_PERM
    own = np.array([[0.5, 1.0 / 3.0], [0.5, 2.0 / 3.0]])
    assert rn.own_nodes_permutations("QUAD", own, lin("QUAD")) == [[1, 2], [0, 0], [1, 2], [0, 0], [0, 0], [2, 1], [0, 0], [2, 1]]


def test_high_order_face_own_nodes_reference_goldens():
    # test/ReferenceFEsTests/CLagrangianRefFEsTests.jl:84-89: LagrangianRefFE(VectorValue{2,Float64}, TRI, 3):
    # get_face_own_dofs == [[1,11],[2,12],[3,13],[4,5,14,15],[6,7,16,17],[8,9,18,19],[10,20]] (DoF = node + 10 * component)
    nodes, own = rt.lagrangian_nodes_and_face_own_nodes("TRI", 3)
    assert len(nodes) == 10
    dofs = [[n + 10 * c for c in range(2) for n in face] for face in own]
    assert dofs == [[1, 11], [2, 12], [3, 13], [4, 5, 14, 15], [6, 7, 16, 17], [8, 9, 18, 19], [10, 20]]
    # :80-82: SEGMENT of order 2, two components: [[1, 4], [2, 5], [3, 6]]
    nodes, own = rt.lagrangian_nodes_and_face_own_nodes("SEG", 2)
    assert [[n + 3 * c for c in range(2) for n in face] for face in own] == [[1, 4], [2, 5], [3, 6]]
    # :121-129: QUAD of order 2: nodes 5..8 on the edges, 9 inside
    nodes, own = rt.lagrangian_nodes_and_face_own_nodes("QUAD", 2)
    assert own == [[1], [2], [3], [4], [5], [6], [7], [8], [9]]
    # the order-0 interior node of compute_own_nodes(TRI, (0,0)) aside, order 3 / 4 interior nodes of TRI follow _add_terms!
    assert np.allclose(rt.interior_nodes("TRI", 3), [[1 / 3, 1 / 3]])
    assert np.allclose(rt.interior_nodes("TRI", 4), [[0.25, 0.25], [0.5, 0.25], [0.25, 0.5]])
    # counts: Q3 hexahedron 8 + 12*2 + 6*4 + 8 = 64, P3 tetrahedron 4 + 6*2 + 4*1 + 0 = 20
    for ptype, n in (("HEX", 64), ("TET", 20), ("QUAD", 16), ("TRI", 10)):
        nodes, own = rt.lagrangian_nodes_and_face_own_nodes(ptype, 3)
        assert len(nodes) == n and sorted(k for f in own for k in f) == list(range(1, n + 1))
        N, dN = rt.lagrangian_tabulate(ptype, 3, nodes)
        assert np.allclose(N, np.eye(n), atol=1e-10) and np.allclose(dN.sum(axis=1), 0.0, atol=1e-9)


def test_general_order_numbering_reduces_to_the_pinned_order2_numbering():
    for part, simplex in (((3, 2), False), ((3, 2), True), ((2, 2, 2), False), ((2, 2, 2), True)):
        D = len(part)
        X, cells, ptype = problems.cartesian_mesh((0, 1) * D, part, simplex)
        for ncomp, tags in ((1, []), (D, ["boundary"]), (D, ["tag_5", "tag_6"] if D == 2 else ["tag_21", "tag_22"])):
            masks = [[True] * ncomp if ncomp > 1 else True for _ in tags]
            if len(tags) == 2:
                masks = [[True, False, True][:ncomp], [False, True, True][:ncomp]]
            dft = {}
            for d in range(D):
                _, fv = rn.global_faces(cells, ptype, d)
                dft[d] = rn.face_tag_index([rn.cartesian_entity_of_vertices(part, list(v)) for v in fv], D, tags)
            a = rn.conforming_dofs_order2(cells, ptype, ncomp, dft, masks)
            b = rn.conforming_dofs(cells, ptype, 2, ncomp, dft, masks)
            assert np.array_equal(a[0], b[0]) and a[1:3] == b[1:3]


def test_compute_conforming_cell_dofs_reference_goldens():
    # test/FESpacesTests/ConformingFESpacesTests.jl:16-69: 3x3 quadrilaterals, dirichlet_tags = ["tag_1", "tag_6"]
    import gridap_b200 as g
    part, tags = (3, 3), ["tag_1", "tag_6"]
    X, cells, ptype = problems.cartesian_mesh((0, 1, 0, 1), part)
    dft = {}
    for d in range(2):
        _, fv = rn.global_faces(cells, ptype, d)
        dft[d] = rn.face_tag_index([rn.cartesian_entity_of_vertices(part, list(v)) for v in fv], 2, tags)
    m = g.CartesianDiscreteModel((0, 1, 0, 1), part)
    # :20-46 order 2, scalar
    r2 = [[-1, 1, 4, 5, 14, 15, 16, 17, 35], [1, 2, 5, 6, 18, 19, 17, 20, 36], [2, 3, 6, 7, 21, 22, 20, 23, 37],
          [4, 5, 8, 9, 15, 24, 25, 26, 38], [5, 6, 9, 10, 19, 27, 26, 28, 39], [6, 7, 10, 11, 22, 29, 28, 30, 40],
          [8, 9, 12, -2, 24, -4, 31, 32, 41], [9, 10, -2, -3, 27, -5, 32, 33, 42], [10, 11, -3, 13, 29, -6, 33, 34, 43]]
    for cd, nfree, ndiri in (rn.conforming_dofs_order2(cells, ptype, 1, dft, [True, True])[:3], rn.conforming_dofs(cells, ptype, 2, 1, dft, [True, True])):
        assert cd.tolist() == r2 and (nfree, ndiri) == (43, 6)
    V = g.FESpace(m, g.ReferenceFE(g.lagrangian, float, 2), dirichlet_tags=tags)
    assert V.cell_dof_ids.tolist() == r2 and (V.nfree, V.ndirichlet) == (43, 6)
    # :48-69 order 1, VectorValue{2}, dirichlet_components = [(true,true), (false,true)]
    r1 = [[-1, 1, 7, 9, -2, 2, 8, 10], [1, 3, 9, 11, 2, 4, 10, 12], [3, 5, 11, 13, 4, 6, 12, 14],
          [7, 9, 15, 17, 8, 10, 16, 18], [9, 11, 17, 19, 10, 12, 18, 20], [11, 13, 19, 21, 12, 14, 20, 22],
          [15, 17, 23, 25, 16, 18, 24, -3], [17, 19, 25, 26, 18, 20, -3, -4], [19, 21, 26, 27, 20, 22, -4, 28]]
    masks = [(True, True), (False, True)]
    cd, nfree, ndiri = rn.conforming_dofs(cells, ptype, 1, 2, dft, masks)          # the general-order restatement at order 1
    assert cd.tolist() == r1 and (nfree, ndiri) == (28, 4)
    V = g.FESpace(m, g.ReferenceFE(g.lagrangian, g.VectorValue(2), 1), dirichlet_tags=tags, dirichlet_masks=masks)
    assert V.cell_dof_ids.tolist() == r1 and (V.nfree, V.ndirichlet) == (28, 4)
