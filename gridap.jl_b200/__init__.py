"""gridap_b200 -- B200-native FE assembly engine behind Gridap's SparseMatrixAssembler interface.

The directory is called `gridap.jl_b200`; import it as `gridap_b200` (alias module at the repo root).
Layout: `csrc/` hand-written CUDA (sm_100a) + the C ABI of `lib/libgridap_b200.so`; the Python modules are the
host-side mirror of the reference interface for this path (Julia is not available in this image).
"""
from . import lib  # noqa: F401
