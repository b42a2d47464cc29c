"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Gridap assembly hot path.

This package is a CPU restatement (numpy + plain C) of the reference algorithm
(Gridap.jl v0.20.8, /root/reference) for the path in SURVEY.md section 8:
cell-wise quadrature of the weak form and the scatter into the global CSC matrix / vector.

Nothing in the product (`gridap.jl_b200/`) may import it.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs use it,
and only as the checker / the CPU baseline.

Parity status: the reference is Julia and `julia` is not installed in this image, so the
reference itself cannot be executed here.  The oracle is pinned against every golden
value the reference's own tests hold for this path (tests/test_oracle_golden.py lists
them with file:line).  The neo-Hookean law and the exact linear-elasticity / Stokes
matrices are NOT in the reference's tests: for those forms "parity unpinned" applies
(the oracle pins them by construction: patch tests + finite-difference Jacobian checks).
"""
