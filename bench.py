#!/usr/bin/env python
"""bench.py -- headline benchmark: `assemble_matrix` for 3D Q1 Poisson on a 256^3-cell mesh (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n 256]

Own arm (`--impl b200`): one process per GPU (torchrun for N > 1).  A step = one full numeric assembly of the
256^3 problem (geometry factors from the node coordinates + all nnz values), device-resident: mesh, ids and the
symbolic plan are already in HBM, the result stays in HBM.  N > 1: strong scaling, cells partitioned in z-slabs,
every rank assembles the CSC columns it owns from its cells + one ghost layer (no data-path collective).
`e2e` = the public call `assemble_matrix(a, U, V)` from host arrays to a host SparseMatrixCSC (H2D of mesh and ids,
symbolic phase, numeric phase, D2H of colptr/rowval/nzval all inside the timed region).
Reference arm (`--impl reference`): the reference is Julia (no `julia` in this image) -> the CPU oracle port of its
algorithm (oracle/ref_assembly.c, single-threaded like the reference) on a bounded sample of the same workload.
Prints ONE JSON line.
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

B_ALG_PER_CELL = 524.1  # coords 24.3 + dof ids 32 + Int32 slot map 256 + nzval 211.8 B/cell
METRIC = "assemble_matrix cells/s, 3D Q1 Poisson (device-resident numeric assembly)"


def workload_name(n):
    return "3D Poisson Q1 hex, %d^3 cells, FP64 matrix assembly (UnstructuredDiscreteModel of a Cartesian mesh, Dirichlet boundary)" % n


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


# ------------------------------------------------------------------------------------------------ reference arm
def oracle_problem(n):
    import gridap_b200 as g
    from oracle import capi
    model = g.CartesianDiscreteModel((0, 1) * 3, (n, n, n))
    V = g.FESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
    xq, w = g.Quadrature("HEX", 2)
    N, dN = g.reffes.tabulate_lagrangian("HEX", 1, xq)
    fld = capi.Field(N, dN, 1, V.cell_dof_ids)
    return capi.Problem(model.node_coordinates, model.cell_node_ids, w, N, dN, [fld], capi.LAPLACIAN, 0, None, None, None, 0, False, V.nfree, V.nfree)


def quadrature_only_context(pb, n_sample):
    """CONTEXT, not the reference algorithm: the per-cell quadrature alone (no sparse insertion) on every host core (POSIX threads)."""
    nt = os.cpu_count() or 1
    t = time.perf_counter()
    pb.quadrature_only(nt)
    dt = time.perf_counter() - t
    return {"value": n_sample ** 3 / dt, "unit": "cells/s", "threads": nt,
            "note": "NOT the reference algorithm (Gridap's loop is serial): local matrices only, no CSC insertion, all host cores"}


def cpu_baseline(n_sample, repeats=1):
    """the oracle port of the reference algorithm (two passes, per-entry binary-search insertion), 1 thread."""
    pb = oracle_problem(n_sample)
    best = None
    for _ in range(repeats):
        t = time.perf_counter()
        pb.assemble()
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    cpu_baseline.context = quadrature_only_context(pb, n_sample)
    return n_sample ** 3 / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.sample_n or 96
    pb = oracle_problem(n)
    for _ in range(min(args.warmup, 1)):
        pb.assemble()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pb.assemble()
    dt = time.perf_counter() - t0
    context = quadrature_only_context(pb, n)
    value = args.steps * n ** 3 / dt
    sample = "%d^3-cell sample of the workload per step (same element, quadrature, boundary conditions)" % n
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": workload_name(args.n), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": 1, "kind": "port", "sample": sample, "host_cores": os.cpu_count(),
                         "julia_threads": "n/a (julia is not installed in this image; Gridap's assembly loop is single-threaded by construction)",
                         "context_quadrature_only_all_cores": context,
                         "note": "reference is Julia (not installed); oracle/ref_assembly.c restates its serial algorithm; host has %d cores" % os.cpu_count()},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


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
    n = args.n
    ctx = lib.Context(local_rank, deterministic=False)

    # ---- inputs (host, untimed): mesh, space, weak form
    t_host = time.perf_counter()
    model = g.UnstructuredDiscreteModel(g.CartesianDiscreteModel((0, 1) * 3, (n, n, n)))
    reffe = g.ReferenceFE(g.lagrangian, float, 1)
    V = g.TestFESpace(model, reffe, dirichlet_tags="boundary")
    U = g.TrialFESpace(V, 0.0)
    dO = g.Measure(g.Triangulation(model), 2)

    def a(u, v):
        return g.Integral(g.inner(g.grad(v), g.grad(u))) * dO

    if world > 1:
        part = gd.slab_partition(model, V, world, rank)
        assem = part.assembler(U, V, ctx)
        ncells_local = part.ncells_owned
    else:
        part = None
        assem = g.SparseMatrixAssembler(U, V, ctx=ctx)
        ncells_local = model.num_cells()
    matdata = g.collect_cell_matrix(U, V, a(g.get_trial_fe_basis(U), g.get_fe_basis(V)))
    t_host = time.perf_counter() - t_host
    plan = assem.plan(matdata.measure, None)   # H2D + symbolic phase (once; reused by every step)
    sym = dict(plan.symbolic_timings)
    term = matdata.terms[0]

    stream = torch.cuda.ExternalStream(ctx.stream(), device=local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        plan.assemble_matrix(term.form, term.params, None, False)   # nzval stays in HBM

    for _ in range(args.warmup):
        step()
    path = plan.kernel_path(term.form)
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
    kern = {k: [v] for k, v in ctx.timings().items()}  # mean device time per kernel over the K timed steps
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = n ** 3 / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (device time from CUDA events on the library's stream)
    peak, peak_src = read_peaks()
    dom_name = next((nm for nm in ("k:q1hex_fused", "k:q1hex_gather", "k:generic") if nm in kern), "kernels")
    dom_ms = float(np.mean(kern[dom_name]))
    step_kernel_ms = float(np.mean(kern.get("kernels", [ms_step])))
    alg_bytes = B_ALG_PER_CELL * ncells_local
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": dom_name[2:] + "_kernel", "kernel_ms": dom_ms, "step_kernels_ms": step_kernel_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "frac_compulsory": 268.1 * ncells_local / (dom_ms * 1e-3) / 1e9 / peak,
                "step_frac": alg_bytes / (step_kernel_ms * 1e-3) / 1e9 / peak,
                "note": "achieved = 524.1 B/cell (SURVEY 8d, incl. a 256 B/cell Int32 slot map this kernel replaces by a stencil "
                        "classification) x cells / kernel time; frac_compulsory uses the stricter 268.1 B/cell; step_frac = whole step "
                        "(cell_geom + gather)",
                "all_kernels_ms": {k: float(np.mean(v)) for k, v in kern.items()}}
    traffic_file = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(traffic_file) and world == 1:
        with open(traffic_file) as f:
            tr = json.load(f)
        if tr.get("n") == n:
            roofline["traffic"] = tr.get("dram_bytes_per_launch")

    # ---- end to end through the public API (host arrays in, host SparseMatrixCSC out)
    # N > 1: every rank does the same for its column slab (its cells + ghost layer in, its columns out, over its own PCIe
    # link); the timed region is bracketed by barriers, bytes are summed over the ranks.
    e2e_steps = max(1, min(args.steps, 3))
    em, es = (model, V) if world == 1 else (part.local_model, part.local_space)

    def e2e_step():
        asm = g.SparseMatrixAssembler(U, V, ctx=ctx) if world == 1 else part.assembler(U, V, ctx)
        em._device.clear()
        es._device.clear()
        return g.assemble_matrix(a, asm, U, V)
    for _ in range(2):  # warm-up: page-locked result buffers and device blocks are pooled and reused from here on
        A = e2e_step()
        del A
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        A = None  # the previous result is released before the next call, as a Newton / time loop would
        A = e2e_step()
    barrier()
    dt = (time.perf_counter() - t0) / e2e_steps
    nptr = (em.num_cells() + 1) * 4
    h2d = em.node_coordinates.nbytes + em.cell_node_ids.nbytes + nptr + (1 if world == 1 else 2) * (es.cell_dof_ids.nbytes + nptr)
    d2h = A.colptr.nbytes + A.rowval.nbytes + A.nzval.nbytes
    tt = torch.tensor([dt, float(h2d), float(d2h), float(A.nnz())], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        dt = float(tmax[0].item())
    h2d, d2h, nnz = int(tt[1].item()), int(tt[2].item()), int(tt[3].item())
    e2e = {"value": n ** 3 / dt, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": dt * 1e3, "steps": e2e_steps,
           "note": "assemble_matrix(a,U,V): H2D mesh+ids, symbolic phase, numeric phase, D2H colptr/rowval/nzval"
                   + ("" if world == 1 else " (each rank its column slab; max over ranks)")}
    del A

    out = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ns = args.sample_n or 128
            v, secs = cpu_baseline(ns)
            cpu = {"value": v, "unit": "cells/s", "cores": 1, "kind": "port",
                   "sample": "%d^3-cell sample of the workload, %.1f s, oracle/ref_assembly.c (serial, like the reference); host has %d cores"
                             % (ns, secs, os.cpu_count()),
                   "host_cores": os.cpu_count(),
                   "julia_threads": "n/a (julia is not installed in this image; Gridap's assembly loop is single-threaded by construction)",
                   "context_quadrature_only_all_cores": cpu_baseline.context}
        out = {
            "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n), "ncells": n ** 3, "free_dofs": (n - 1) ** 3, "nnz": nnz,
                       "dofs_per_s": (n - 1) ** 3 / (ms_step * 1e-3), "kernel_path": path,
                       "l2_policy": "inputs+outputs per step (>= 5 GB) exceed the 126 MB L2; no explicit flush",
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
    ap.add_argument("--n", type=int, default=256, help="cells per axis (headline: 256)")
    ap.add_argument("--sample-n", type=int, default=0, help="cells per axis of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
