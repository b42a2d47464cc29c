"""Generates tests/golden/<case>.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).
See tests/golden_cases.py for why the generator is the oracle and not the (Julia) reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_cases  # noqa: E402

only = sys.argv[1:]          # python tests/golden/make_golden.py [case ...]: regenerate only the named fixtures
for name in golden_cases.CASES:
    if only and name not in only:
        continue
    pb, extra, with_vector = golden_cases.build(name)
    out = pb.assemble(with_vector=with_vector)
    d = {"colptr": out[0], "rowval": out[1], "nzval": out[2]}
    if with_vector:
        d["b"] = out[3]
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **d)
    print(name, "nnz", len(out[1]), os.path.getsize(path), "bytes")
