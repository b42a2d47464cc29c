"""The reference's own assembly benchmark (benchmark/bm/bm_assembly.jl:7-57) on the B200 path, case by case:

    driver(trian, reffe, qdegree, biform):  dΩ = Measure(trian, qdegree); V = TestFESpace(model, reffe); A = assemble_matrix(a, V, V)

for (D, n) in [(2, 10), (3, 6)], trian in [bulk, view on the first half of the cells], order in [1, 2, 3], scalar / vector-valued Lagrangian
elements (raviart_thomas: out of scope), biform in [mass (2*order), laplacian (2*(order-1)), graddiv (2*(order-1), vector-valued only)].

Per case one JSON line with
  driver_ms     the whole `driver` call through the public API, host arrays in, host SparseMatrixCSC out (what BenchmarkTools times in the
                reference: space construction + symbolic + numeric phase), median of `--samples` runs after one warm-up
  assemble_ms   `assemble_matrix(a, V, V)` alone on a fresh assembler (symbolic + numeric + download)
  numeric_ms    re-assembly on the persistent plan, device-resident (the Newton / time-loop cost)
  cpu_port_ms   the oracle port of the reference algorithm on one host core (symbolic + numeric, numbering not included)
These meshes hold 50 - 216 cells: every GPU figure here is launch- and latency-bound, the table documents that the protocol runs, not a
bandwidth claim.  Usage: python scripts/bm_assembly.py [--samples 5] [--no-cpu] > profiles/bm_assembly_r02.jsonl"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gridap_b200 as g  # noqa: E402


def cases():
    for D, n in ((2, 10), (3, 6)):
        for trian_name in ("bulk", "view"):
            for order in (1, 2, 3):
                for vector in (False, True):
                    for biform, qdegree in (("mass", 2 * order), ("laplacian", 2 * (order - 1)), ("graddiv", 2 * (order - 1))):
                        if biform == "graddiv" and not vector:
                            continue
                        yield D, n, trian_name, order, vector, biform, qdegree


def biform_fn(biform, dO):
    if biform == "mass":
        return lambda u, v: g.Integral(g.dot(u, v)) * dO
    if biform == "laplacian":
        return lambda u, v: g.Integral(g.inner(g.grad(u), g.grad(v))) * dO
    return lambda u, v: g.Integral(g.div(u) * g.div(v)) * dO


def driver(model, trian, reffe, qdegree, biform):
    dO = g.Measure(trian, qdegree)
    V = g.TestFESpace(model, reffe)
    return g.assemble_matrix(biform_fn(biform, dO), V, V), V, dO


def median_ms(fn, samples):
    fn()
    ts = []
    for _ in range(samples):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts) * 1e3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    models = {}
    for case in cases():
        D, n, trian_name, order, vector, biform, qdegree = case
        if (D, n) not in models:
            models[(D, n)] = g.UnstructuredDiscreteModel(g.CartesianDiscreteModel((0, 1) * D, (n,) * D))
        model = models[(D, n)]
        trian = g.Triangulation(model) if trian_name == "bulk" else g.Triangulation(model, np.arange(1, n ** D // 2 + 1))
        reffe = g.ReferenceFE(g.lagrangian, g.VectorValue(D) if vector else float, order)
        A, V, dO = driver(model, trian, reffe, qdegree, biform)
        out = {"case": "assembly_%dD_%s_%s_%s_%d" % (D, trian_name, "vector_lagrangian" if vector else "lagrangian", biform, order),
               "ncells": int(trian.model.num_cells()), "free_dofs": int(V.nfree), "nnz": int(A.nnz()), "qdegree": qdegree}
        out["driver_ms"] = median_ms(lambda: driver(model, trian, reffe, qdegree, biform), args.samples)
        a = biform_fn(biform, dO)
        out["assemble_ms"] = median_ms(lambda: g.assemble_matrix(a, V, V), args.samples)
        assem = g.SparseMatrixAssembler(V, V)
        data = g.collect_cell_matrix(V, V, a(g.get_trial_fe_basis(V), g.get_fe_basis(V)))
        plan = assem.plan(dO)
        A2 = assem.assemble_matrix(data)

        def numeric():
            for t in data.terms:
                plan.assemble_matrix(t.form, t.params, None, False)
            assem.ctx.synchronize()

        out["numeric_ms"] = median_ms(numeric, max(args.samples, 10))
        out["kernel_path"] = plan.kernel_path(data.terms[0].form)
        assert np.array_equal(A2.rowval, A.rowval)
        if not args.no_cpu:
            from test_gpu_bm_protocol import bm_oracle
            t0 = time.perf_counter()
            ref = bm_oracle(model, V, trian_name, n, order, biform, qdegree)
            out["cpu_port_ms"] = (time.perf_counter() - t0) * 1e3
            out["max_rel_diff_vs_cpu_port"] = float(np.abs(A.nzval - ref[2]).max() / np.abs(ref[2]).max())
            assert np.array_equal(A.rowval, ref[1]) and np.array_equal(A.colptr, ref[0])
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
