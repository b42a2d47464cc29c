"""Small instances of every affine-gather instance (for compute-sanitizer)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import gridap_b200 as g

def run(name, f):
    try:
        f()
        print("ok", name, flush=True)
    except Exception as e:
        print("FAIL", name, repr(e)[:300], flush=True)

def stokes():
    model = g.simplexify(g.CartesianDiscreteModel((0, 1) * 3, (2, 2, 2)))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 2), dirichlet_tags="boundary")
    Q = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1))
    Y = g.MultiFieldFESpace([V, Q])
    dO = g.Measure(g.Triangulation(model), 4)
    def a(up, vq):
        (u, p), (v, q) = up, vq
        return g.Integral(g.inner(g.grad(v), g.grad(u)) - g.div(v) * p + q * g.div(u)) * dO
    A = g.assemble_matrix(a, Y, Y)
    print(A.nnz(), np.abs(A.nzval).max())

def elas(order, tets=False):
    def f():
        model = g.CartesianDiscreteModel((0, 1) * 3, (3, 2, 2))
        if tets:
            model = g.simplexify(model)
        V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), order), dirichlet_tags=[25])
        dO = g.Measure(g.Triangulation(model), 2 * order)
        sigma = g.IsotropicLinearElasticity(2.0, 1.0)
        A = g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.eps(v), sigma(g.eps(u)))) * dO, V, V)
        print(A.nnz(), np.abs(A.nzval).max())
    return f

def scalar(order, tets):
    def f():
        model = g.CartesianDiscreteModel((0, 1) * 3, (3, 2, 2))
        if tets:
            model = g.simplexify(model)
        V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, order), dirichlet_tags=[25])
        dO = g.Measure(g.Triangulation(model), 2 * order)
        A = g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO, V, V)
        print(A.nnz(), np.abs(A.nzval).max())
    return f

run("elas q2", elas(2))
run("elas q1", elas(1))
run("elas p2", elas(2, True))
run("elas p1", elas(1, True))
run("scalar q2", scalar(2, False))
run("scalar p2", scalar(2, True))
run("scalar p1", scalar(1, True))
run("stokes", stokes)


# ---- round 2 additions: staged gather (non-affine / neo-Hookean), mirror pairs, skeleton plans, constraint fold, headline graph replay
def neohookean(perturbed):
    def f():
        model = g.CartesianDiscreteModel((0, 1) * 3, (5, 3, 3))
        if perturbed:
            rng = np.random.default_rng(1)
            X = model.node_coordinates
            inner = np.all((X > 1e-9) & (X < 1 - 1e-9), axis=1)
            X[inner] += 0.03 * rng.uniform(-1, 1, size=(int(inner.sum()), 3))
        V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 1), dirichlet_tags="boundary")
        U = g.TrialFESpace(V, (0.0, 0.0, 0.0))
        dO = g.Measure(g.Triangulation(model), 2)
        nh = g.NeoHookean(100.0, 1.0)
        uh = g.interpolate(lambda x: 0.05 * np.sin(np.pi * x[:, :1]) * np.ones((1, 3)), U)
        op = g.FEOperator(lambda u, v: g.Integral(nh.res(u, v)) * dO, lambda u, du, v: g.Integral(nh.jac(u, du, v)) * dO, U, V)
        b, A = op.residual_and_jacobian(uh)
        print(A.nnz(), np.abs(A.nzval).max(), np.abs(b).max())
    return f


def mirror():
    import os
    os.environ["GB200_MIRROR"] = "1"
    try:
        elas(1)()
        stokes()
    finally:
        del os.environ["GB200_MIRROR"]


def skeleton_and_constraints():
    model = g.CartesianDiscreteModel((0, 1, 0, 1), (3, 3))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 2), conformity="L2")
    dO = g.Measure(g.Triangulation(model), 4)
    L = g.SkeletonTriangulation(model)
    dL = g.Measure(L, 4)
    n = g.get_normal_vector(L)
    A = g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO + g.Integral(
        4.0 * g.dot(g.jump(v * n), g.jump(u * n)) - g.dot(g.jump(v * n), g.mean(g.grad(u))) - g.dot(g.mean(g.grad(v)), g.jump(u * n))) * dL, V, V)
    print(A.nnz(), np.abs(A.nzval).max())
    W = g.FESpace(g.CartesianDiscreteModel((0, 1, 0, 1), (2, 2)), g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags=[1, 2, 5])
    Wc = g.FESpaceWithLinearConstraints([1, 5, -2], [[-1, 4], [4, 6], [-1, -3]], [[0.5, 0.5]] * 3, W)
    Uc = g.TrialFESpace(Wc, lambda x: x[:, 0] + 2 * x[:, 1])
    dW = g.Measure(g.Triangulation(W.model), 2)
    op = g.AffineFEOperator(lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dW, lambda v: g.Integral(v * 1.0) * dW, Uc, Wc)
    print(op.get_matrix().nnz(), np.abs(op.get_vector()).max())


def headline_graph():
    model = g.UnstructuredDiscreteModel(g.CartesianDiscreteModel((0, 1) * 3, (6, 5, 4)))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
    dO = g.Measure(g.Triangulation(model), 2)
    assem = g.SparseMatrixAssembler(V, V)
    md = g.collect_cell_matrix(V, V, g.Integral(g.inner(g.grad(g.get_fe_basis(V)), g.grad(g.get_trial_fe_basis(V)))) * dO)
    A = assem.allocate_matrix(md)
    for _ in range(5):   # from the third identical device-resident call on: one CUDA graph per step
        assem.plan(dO).assemble_matrix(md.terms[0].form, md.terms[0].params, None, False)
    assem.assemble_matrix_(A, md)
    print(A.nnz(), np.abs(A.nzval).max())


run("neo-Hookean q1 (nhq1 staged)", neohookean(False))
run("neo-Hookean q1 perturbed", neohookean(True))
run("mirror pairs", mirror)
run("skeleton + constraints", skeleton_and_constraints)
run("headline graph replay", headline_graph)
