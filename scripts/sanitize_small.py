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
