#!/usr/bin/env python
"""bench.py -- `assemble_matrix` throughput of the B200 assembly engine on the BASELINE.json configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 2|2rhs|2general|1|3|4|5] [--n SIZE]

Default (`--config 2`): the headline, 3D Q1 Poisson `assemble_matrix` on a 256^3-cell mesh (BASELINE.json configs[1]).
Own arm (`--impl b200`): one process per GPU (torchrun for N > 1).  A step = one full numeric assembly of the workload (geometry
from the node coordinates + all nnz values), device-resident: mesh, ids and the symbolic plan are already in HBM, the result stays
in HBM.  N > 1: strong scaling, cells partitioned in z-slabs, every rank assembles the CSC columns it owns from its cells + one
ghost layer (no data-path collective).  `e2e` = the public call (`assemble_matrix(a,U,V)` / `AffineFEOperator` /
`residual_and_jacobian`) from host arrays to host results (H2D of mesh and ids, symbolic phase, numeric phase, D2H inside the
timed region).
Reference arm (`--impl reference`): the reference is Julia (no `julia` in this image) -> the CPU oracle port of its algorithm
(oracle/ref_assembly.c, single-threaded like the reference) on a bounded sample of the same workload.
The other configs follow the reference's own benchmark protocol (benchmark/bm/bm_assembly.jl:7-57: assemble on the
UnstructuredDiscreteModel of a Cartesian mesh, per element / form) at the BASELINE sizes.  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG_PER_CELL = 524.1    # coords 24.3 + dof ids 32 + Int32 slot map 256 + nzval 211.8 B/cell
COMPULSORY_PER_CELL = 268.1  # the same without any slot map: what has to cross the HBM interface at least once (config 2)
E_MOD, NU = 2.1e4, 0.3
LAM, MU = E_MOD * NU / ((1 + NU) * (1 - 2 * NU)), E_MOD / (2 * (1 + NU))


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workloads
class Workload:
    """One BASELINE.json config: host inputs (model, spaces, forms) through the public API of the package."""

    def __init__(self, key, n):
        import gridap_b200 as g
        self.key, self.n, self.g = key, n, g
        self.uh, self.kind = None, "matrix"
        if key in ("2", "2rhs", "2general"):
            self.model = g.UnstructuredDiscreteModel(g.CartesianDiscreteModel((0, 1) * 3, (n, n, n)))
            if key == "2general":   # same connectivity, interior nodes displaced by 0.2 dx U(-1,1)^3, default_rng(12345)
                X = self.model.node_coordinates
                inner = np.all((X > 1e-9) & (X < 1 - 1e-9), axis=1)
                X[inner] += 0.2 / n * np.random.default_rng(12345).uniform(-1, 1, size=(int(inner.sum()), 3))
            self.V = g.TestFESpace(self.model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
            self.U = g.TrialFESpace(self.V, (lambda x: x[:, 0] + 2.0 * x[:, 1]) if key == "2rhs" else 0.0)
            self.degree = 2
            if key == "2rhs":
                self.kind = "matrix+rhs"
            self.metric = "assemble_matrix cells/s, 3D Q1 Poisson (device-resident numeric assembly)"
        elif key == "1":
            self.model = g.UnstructuredDiscreteModel(g.CartesianDiscreteModel((0, 1, 0, 1), (n, n)))
            self.V = g.TestFESpace(self.model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
            self.U = g.TrialFESpace(self.V, 0.0)
            self.degree, self.kind = 2, "matrix+rhs"
            self.metric = "assemble_matrix_and_vector cells/s, 2D Q1 Poisson"
        elif key == "3":
            self.model = g.UnstructuredDiscreteModel(g.CartesianDiscreteModel((0, 1) * 3, (n, n, n)))
            self.V = g.TestFESpace(self.model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 2), dirichlet_tags=[25, 1, 3, 5, 7, 13, 15, 17, 19])
            self.U = g.TrialFESpace(self.V, (0.0, 0.0, 0.0))
            self.degree = 4
            self.metric = "assemble_matrix cells/s, 3D Q2 linear elasticity"
        elif key == "4":
            self.model = g.simplexify(g.CartesianDiscreteModel((0, 1) * 3, (n, n, n)))
            Vv = g.TestFESpace(self.model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 2), dirichlet_tags="boundary")
            Q = g.TestFESpace(self.model, g.ReferenceFE(g.lagrangian, float, 1))
            self.V = g.MultiFieldFESpace([Vv, Q])
            self.U = g.MultiFieldFESpace([g.TrialFESpace(Vv, (0.0, 0.0, 0.0)), g.TrialFESpace(Q)])
            self.degree = 4
            self.metric = "assemble_matrix cells/s, Stokes Taylor-Hood P2/P1 tets"
        elif key == "5":
            self.model = g.UnstructuredDiscreteModel(g.CartesianDiscreteModel((0, 1) * 3, (n, n, n)))
            self.V = g.TestFESpace(self.model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 1), dirichlet_tags="boundary")
            self.U = g.TrialFESpace(self.V, (0.0, 0.0, 0.0))
            self.degree, self.kind = 2, "res+jac"
            self.uh = g.interpolate(lambda x: 0.05 * (np.sin(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1]) * np.sin(np.pi * x[:, 2]))[:, None] * np.ones((1, 3)), self.U)
            self.metric = "residual_and_jacobian cells/s, 3D Q1 neo-Hookean"
        else:
            raise SystemExit("unknown --config %r" % key)
        self.name = workload_title(key, n)
        self.ncells = self.model.num_cells()
        self.nfree = self.V.num_free_dofs()
        self._bind()

    def _bind(self):
        """the weak forms on this workload's model (a rank's local workload binds them to its local measure)"""
        g, key = self.g, self.key
        dO = self.dO = g.Measure(g.Triangulation(self.model), self.degree)
        self.l = self.res = self.jac = None
        if key in ("1", "2", "2rhs", "2general"):
            self.a = lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO
            if self.kind == "matrix+rhs":
                self.l = lambda v: g.Integral(v * 1.0) * dO
        elif key == "3":
            sigma = g.IsotropicLinearElasticity(LAM, MU)
            self.a = lambda u, v: g.Integral(g.inner(g.eps(v), sigma(g.eps(u)))) * dO
        elif key == "4":
            def a(up, vq):
                (u, p), (v, q) = up, vq
                return g.Integral(g.inner(g.grad(v), g.grad(u)) - g.div(v) * p + q * g.div(u)) * dO
            self.a = a
        else:
            nh = g.NeoHookean(100.0, 1.0)
            self.res = lambda u, v: g.Integral(nh.res(u, v)) * dO
            self.jac = lambda u, du, v: g.Integral(nh.jac(u, du, v)) * dO

    def localized(self, part):
        """this workload on a rank's part of the mesh (own cells + ghost layer, global DoF numbering)"""
        import copy
        w = copy.copy(self)
        w.model = part.local_model
        w.U = w.V = part.local_space
        w.ncells = w.model.num_cells()
        if self.uh is not None:
            w.uh = part.local_function(self.uh)
        w._bind()
        return w

    # ---- device-resident step on a persistent plan
    def make_step(self, assem):
        g, lib = self.g, self.g.lib
        if self.kind == "res+jac":
            matdata = g.collect_cell_matrix(self.U, self.V, self.jac(self.uh, g.get_trial_fe_basis(self.U), g.get_fe_basis(self.V)))
            plan = assem.plan(matdata.measure, None)
            plan.set_state(0, self.uh.free_values, self.uh.dirichlet_values)
            prm = matdata.terms[0].params
            return plan, matdata.terms[0].form, lambda: plan.assemble_matrix_and_vector(lib.FORM_NEOHOOKEAN_JAC, prm, lib.FORM_NEOHOOKEAN_RES, prm, None, None, None)
        matdata = g.collect_cell_matrix(self.U, self.V, self.a(g.get_trial_fe_basis(self.U), g.get_fe_basis(self.V)))
        plan = assem.plan(matdata.measure, assem._touched(matdata.terms))
        term = matdata.terms[0]
        if self.kind == "matrix+rhs":
            assem._set_dirichlet(plan, None)
            return plan, term.form, lambda: plan.assemble_matrix_and_vector(term.form, term.params, lib.FORM_SOURCE, (1.0,), None, None, None)
        return plan, term.form, lambda: plan.assemble_matrix(term.form, term.params, None, False)

    # ---- end to end through the public API: host arrays in, host results out
    def e2e_call(self, assem):
        g = self.g
        if self.kind == "res+jac":
            op = g.FEOperator(self.res, self.jac, self.U, self.V, assem)
            b, A = op.residual_and_jacobian(self.uh)
            return A, b
        if self.kind == "matrix+rhs":
            op = g.AffineFEOperator(self.a, self.l, self.U, self.V, assem)
            return op.get_matrix(), op.get_vector()
        return g.assemble_matrix(self.a, assem, self.U, self.V), None

    def compulsory_bytes(self, plan):
        """bytes that must cross the HBM interface at least once per step: every stored value written once, the plan's slot map
        (none on the owner-computes gather path), cell ids, node coordinates (+ the state read / vector written)"""
        fields = self.V.spaces if hasattr(self.V, "spaces") else [self.V]
        nl = sum(f.cell_dof_ids.shape[1] for f in fields)
        b = 8.0 * plan.nnz + 4.0 * self.ncells * nl + self.model.cell_node_ids.nbytes + self.model.node_coordinates.nbytes
        if self.key not in ("2", "2rhs", "2general"):
            b += 2.0 * self.ncells * nl * nl
        if self.kind != "matrix":
            b += 8.0 * self.nfree
        if self.kind == "res+jac":
            b += 8.0 * self.nfree
        return b


DEFAULT_N = {"1": 100, "2": 256, "2rhs": 256, "2general": 256, "3": 128, "4": 70, "5": 192}
SAMPLE_N = {"1": 100, "2": 128, "2rhs": 96, "2general": 96, "3": 8, "4": 14, "5": 24}


# ------------------------------------------------------------------------------------------------ reference arm / CPU baseline
def oracle_sample(key, n):
    """the oracle port of the reference algorithm (two passes, per-entry binary-search insertion), 1 thread, on a sample of `key`"""
    from oracle import capi
    from oracle import ref_tabulation as rt
    w = Workload(key, n)
    model = w.model
    fields_host = w.V.spaces if hasattr(w.V, "spaces") else [w.V]
    degree = w.dO.degree
    xq, wq = rt.quadrature(model.ptype, degree)
    Ng, dNg = rt.lagrangian_tabulate(model.ptype, 1, xq)
    ids = w.V.get_cell_dof_ids() if hasattr(w.V, "spaces") else [w.V.cell_dof_ids]
    offs = w.V.offsets if hasattr(w.V, "spaces") else [0]
    flds = []
    for k, f in enumerate(fields_host):
        N, dN = rt.lagrangian_tabulate(model.ptype, f.reffe.order, xq)
        fv = w.uh.free_values if w.uh is not None else None
        dv = w.uh.dirichlet_values if w.uh is not None else getattr(_fields(w.U)[k], "dirichlet_values", None)
        flds.append(capi.Field(N, dN, f.ncomp, ids[k], offs[k], fv, dv))
    form_mat = {"1": capi.LAPLACIAN, "2": capi.LAPLACIAN, "2rhs": capi.LAPLACIAN, "2general": capi.LAPLACIAN, "3": capi.ELASTICITY,
                "4": capi.STOKES, "5": capi.NEOHOOKEAN_JAC}[key]
    form_vec = capi.SOURCE if w.kind == "matrix+rhs" else capi.NEOHOOKEAN_RES if w.kind == "res+jac" else 0
    params = {"3": [LAM, MU], "5": [100.0, 1.0]}.get(key, [1.0])
    touched = np.array([[1, 1], [1, 0]], dtype=np.uint8) if key == "4" else None
    pb = capi.Problem(model.node_coordinates, model.cell_node_ids, wq, Ng, dNg, flds, form_mat, form_vec, params, None, touched, 0,
                      w.kind == "matrix+rhs", w.nfree, w.nfree)
    return w, pb, (lambda: pb.assemble(with_vector=form_vec != 0))


def _fields(space):
    return space.spaces if hasattr(space, "spaces") else [space]


def quadrature_only_context(pb, ncells):
    """CONTEXT, not the reference algorithm: the per-cell quadrature alone (no sparse insertion) on every host core (POSIX threads)."""
    nt = os.cpu_count() or 1
    t = time.perf_counter()
    pb.quadrature_only(nt)
    dt = time.perf_counter() - t
    return {"value": ncells / dt, "unit": "cells/s", "threads": nt,
            "note": "NOT the reference algorithm (Gridap's loop is serial): local matrices only, no CSC insertion, all host cores"}


def sample_text(w, secs=None):
    t = "" if secs is None else ", %.1f s" % secs
    return "%d-cell sample of the workload (%s; same element, quadrature, boundary conditions)%s, oracle/ref_assembly.c (serial, like the reference)" \
        % (w.ncells, w.name.split(",")[0] + " n=%d" % w.n, t)


def cpu_baseline(key, n_sample):
    w, pb, run = oracle_sample(key, n_sample)
    t = time.perf_counter()
    run()
    dt = time.perf_counter() - t
    return {"value": w.ncells / dt, "unit": "cells/s", "cores": 1, "kind": "port", "sample": sample_text(w, dt) + "; host has %d cores" % os.cpu_count(),
            "host_cores": os.cpu_count(),
            "julia_threads": "n/a (julia is not installed in this image; Gridap's assembly loop is single-threaded by construction)",
            "context_quadrature_only_all_cores": quadrature_only_context(pb, w.ncells)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    key = args.config
    n = args.sample_n or {"2": 96}.get(key, SAMPLE_N[key])
    w, pb, run = oracle_sample(key, n)
    for _ in range(min(args.warmup, 1)):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    value = args.steps * w.ncells / dt
    sample = sample_text(w) + " per step"
    wl = workload_title(key, args.n or DEFAULT_N[key])
    print(json.dumps({
        "impl": "reference", "metric": w.metric, "value": value, "unit": "cells/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": wl, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": 1, "kind": "port", "sample": sample, "host_cores": os.cpu_count(),
                         "julia_threads": "n/a (julia is not installed in this image; Gridap's assembly loop is single-threaded by construction)",
                         "context_quadrature_only_all_cores": quadrature_only_context(pb, w.ncells),
                         "note": "reference is Julia (not installed); oracle/ref_assembly.c restates its serial algorithm; host has %d cores" % os.cpu_count()},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_title(key, n):
    """the workload string without building the (possibly multi-GB) inputs"""
    return {
        "1": "2D Poisson Q1 on CartesianDiscreteModel %dx%d, assemble_matrix + assemble_vector" % (n, n),
        "2": "3D Poisson Q1 hex, %d^3 cells, FP64 matrix assembly (UnstructuredDiscreteModel of a Cartesian mesh, Dirichlet boundary)" % n,
        "2rhs": "3D Poisson Q1 hex, %d^3 cells, FP64 matrix + RHS (AffineFEOperator, Dirichlet lifting) assembly (UnstructuredDiscreteModel of a Cartesian mesh, Dirichlet boundary)" % n,
        "2general": "3D Poisson Q1 hex, %d^3 cells, FP64 matrix assembly (UnstructuredDiscreteModel of a Cartesian mesh, Dirichlet boundary, general (perturbed, non-affine) geometry)" % n,
        "3": "3D linear elasticity Q2 vector-valued hex, %d^3 cells (FP64 tensor-core local contraction)" % n,
        "4": "Stokes Taylor-Hood P2/P1 multifield on a tetrahedral mesh, %d cells (simplexified %d^3, block assembly)" % (6 * n ** 3, n),
        "5": "3D neo-Hookean hyperelasticity Q1, Newton residual + Jacobian assembly, %d^3 cells" % n,
    }[key]


# ------------------------------------------------------------------------------------------------ own arm
def run_b200(args):
    import torch
    import gridap_b200 as g
    from gridap_b200 import distributed as gd
    from gridap_b200 import lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device: libgridap_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        # NCCL writes its banner / debug lines to stdout by default; stdout carries exactly one JSON line (bench contract)
        # (the "NCCL version ..." banner of communicator creation): file descriptor 1 points at stderr until the communicator exists
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    key = args.config
    n = args.n or DEFAULT_N[key]
    ctx = lib.Context(local_rank, deterministic=False)

    # ---- inputs (host, untimed): mesh, space, weak form
    t_host = time.perf_counter()
    w = Workload(key, n)
    if world > 1:
        part = gd.partition(w.model, w.U, w.V, world, rank)
        assem = part.assembler(ctx)
        wl = w.localized(part)
        ncells_local = part.ncells_owned
    else:
        part, wl = None, w
        assem = g.SparseMatrixAssembler(w.U, w.V, ctx=ctx)
        ncells_local = w.ncells
    t_host = time.perf_counter() - t_host
    plan, form, step = wl.make_step(assem)   # H2D + symbolic phase (once; reused by every step)
    sym = dict(plan.symbolic_timings)

    stream = torch.cuda.ExternalStream(ctx.stream(), device=local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    path = plan.kernel_path(form)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.timings()  # reset the per-kernel event timers after warm-up
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()  # asynchronous: device-resident re-assemblies queue back to back on the library's stream
    e1.record(stream)
    barrier()
    kern = {k: float(v) for k, v in ctx.timings().items()}  # mean device time per region over the K timed steps
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    if key in ("2", "2rhs") and "k:q1hex_gather" not in kern:
        # the timed steps replayed the two launches of a step as one CUDA graph; per-kernel attribution from a short untimed pass
        # with the graph off (same kernels, same arguments, timed one by one)
        os.environ["GB200_GRAPH"] = "0"
        for _ in range(3):
            step()
        barrier()
        split = {k: float(v) for k, v in ctx.timings().items() if k.startswith("k:")}
        del os.environ["GB200_GRAPH"]
        kern.update({k + " (untimed attribution pass, graph off)": v for k, v in split.items()})
        kern_attr = split
    else:
        kern_attr = kern
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = w.ncells / (ms_step * 1e-3)

    # ---- roofline: compulsory bytes of one step / device time of the step's kernels (CUDA events on the library's stream)
    peak, peak_src = read_peaks()
    step_kernel_ms = kern.get("kernels", ms_step)
    comp_bytes = wl.compulsory_bytes(plan)
    if key == "2":
        comp_bytes = COMPULSORY_PER_CELL * ncells_local   # SURVEY 8(d)'s per-cell figure x the cells one launch processes
    knames = [k for k in kern_attr if k.startswith("k:")]
    dom = max(knames, key=lambda k: kern_attr[k]) if knames else "kernels"
    dom_ms = kern_attr.get(dom, step_kernel_ms)
    pipeline = dom == "k:q1hex_pipeline"
    achieved = comp_bytes / (step_kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": ("q1hex_gather_kernel + cell_geom_kernel (L2 chunk pipeline, one CUDA graph per step)" if pipeline else dom[2:] + "_kernel"),
                "kernel_ms": dom_ms, "step_kernels_ms": step_kernel_ms, "algorithmic_bytes_per_launch": comp_bytes, "peak_source": peak_src,
                "note": "achieved = compulsory bytes of one step (every nnz value written once + cell ids + node coordinates"
                        + (" = 268.1 B/cell, SURVEY 8d" if key == "2" else " + the plan's slot map" if not key.startswith("2") else "")
                        + ") / device time of ALL kernels of the step; frac is physical (<= 1)",
                "all_kernels_ms": kern}
    if key == "2":
        roofline["b_alg_note"] = "SURVEY 8(d)'s B_alg = 524.1 B/cell also counts a 256 B/cell Int32 slot map that this path never reads (stencil " \
                                 "classification instead); B_alg x cells / step time = %.0f GB/s is NOT a fraction of peak" % (B_ALG_PER_CELL * ncells_local / (step_kernel_ms * 1e-3) / 1e9)
    traffic_file = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if os.path.exists(traffic_file) and world == 1:
        with open(traffic_file) as f:
            tr = json.load(f)
        ent = tr.get(key, {})
        if ent.get("n") == n:
            roofline["traffic"] = ent.get("dram_bytes_per_step")
            roofline["traffic_source"] = ent.get("source")

    # ---- end to end through the public API (host arrays in, host results out)
    # N > 1: every rank does the same for its column slab (its cells + ghost layer in, its columns out, over its own PCIe
    # link); the timed region is bracketed by barriers, bytes are summed over the ranks.
    e2e = None
    big = plan.nnz * 24 > 40e9   # host CSC (Int64 colptr/rowval + values) beyond ~40 GB: not attempted
    if not args.no_e2e and not big:
        e2e_steps = max(1, min(args.steps, 3))

        def e2e_step():
            asm = g.SparseMatrixAssembler(w.U, w.V, ctx=ctx) if world == 1 else part.assembler(ctx)
            wl.model._device.clear()
            for sp in _fields(wl.V):
                getattr(sp, "space", sp)._device.clear()
            return wl.e2e_call(asm)
        for _ in range(2):  # warm-up: page-locked result buffers and device blocks are pooled and reused from here on
            A, b = e2e_step()
            del A, b
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            A = b = None  # the previous result is released before the next call, as a Newton / time loop would
            A, b = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / e2e_steps
        nptr = (wl.ncells + 1) * 4
        h2d = wl.model.node_coordinates.nbytes + wl.model.cell_node_ids.nbytes + nptr + sum(s.cell_dof_ids.nbytes + nptr for s in _fields(wl.V)) * (1 if world == 1 else 2)
        if wl.uh is not None:
            h2d += wl.uh.free_values.nbytes + wl.uh.dirichlet_values.nbytes
        # bytes that cross the link: the row indices travel as the Int32 the device holds and are widened to Int64 by host threads
        # (gb200_plan_get_pattern_async; GB200_HOST_WIDEN=0 or fewer than 2^22 entries: widened on the device, Int64 on the link)
        host_widen = os.environ.get("GB200_HOST_WIDEN", "1") != "0" and len(A.rowval) >= (1 << 22) and \
            (int(os.environ.get("LOCAL_WORLD_SIZE", "1")) <= 2 or os.environ.get("GB200_HOST_WIDEN") == "1")
        d2h = A.colptr.nbytes + (A.rowval.nbytes // 2 if host_widen else A.rowval.nbytes) + A.nzval.nbytes + (0 if b is None else np.asarray(b).nbytes)
        tt = torch.tensor([dt, float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
        if world > 1:
            tmax = tt.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            dt = float(tmax[0].item())
        h2d, d2h = int(tt[1].item()), int(tt[2].item())
        e2e = {"value": w.ncells / dt, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": dt * 1e3, "steps": e2e_steps,
               "note": "public API call from host arrays to host results: H2D mesh + ids, symbolic phase, numeric phase, D2H colptr/rowval/nzval (row indices as Int32 on the link, widened to Int64 by host threads while the values arrive)"
                       + ("" if world == 1 else " (each rank its column slab; max over ranks)")
                       + "; results land in pooled page-locked buffers (gb200_host_alloc, warmed by two untimed calls: a first call from a "
                         "cold process additionally pays the pinning of the result arrays)"}
        del A, b
    nnz_t = torch.tensor([float(plan.nnz)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(nnz_t, op=dist.ReduceOp.SUM)
    nnz = int(nnz_t.item())

    out = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(key, args.sample_n or SAMPLE_N[key])
        out = {
            "metric": w.metric, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w.name, "ncells": w.ncells, "free_dofs": w.nfree, "nnz": nnz,
                       "dofs_per_s": w.nfree / (ms_step * 1e-3), "kernel_path": path,
                       "l2_policy": "inputs+outputs per step exceed the 126 MB L2; no explicit flush" if comp_bytes > 4e8 else
                                    "working set below the L2 size: steps run back to back, inputs may stay in L2",
                       "parallelism": "1 GPU" if world == 1 else "%d GPUs: z-slab cell partition, owner-computes columns + 1 ghost layer, no collective" % world,
                       "symbolic_ms": sym, "host_input_build_s": t_host},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="2", choices=list(DEFAULT_N), help="BASELINE.json config (2 = headline; 2rhs / 2general: its RHS and general-geometry variants)")
    ap.add_argument("--n", type=int, default=0, help="cells per axis (default: the BASELINE size of the config)")
    ap.add_argument("--sample-n", type=int, default=0, help="cells per axis of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
