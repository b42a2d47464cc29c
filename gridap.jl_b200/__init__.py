"""gridap_b200 -- B200-native FE assembly engine behind Gridap's SparseMatrixAssembler interface.

The directory is called `gridap.jl_b200`; import it as `gridap_b200` (alias module at the repo root).
Layout: `csrc/` hand-written CUDA (sm_100a) + the C ABI of `lib/libgridap_b200.so`; the Python modules are the
host-side mirror of the reference interface for this path (Julia is not available in this image):
geometry / reffes / fespaces produce the inputs (`node_coordinates`, `cell_node_ids`, `cell_dof_ids`, tabulations),
celldata recognises the weak form, assemblers is the `SparseMatrixAssembler` drop-in.
"""
from . import lib  # noqa: F401
from .algebra import BlockMatrix, BlockVector, SparseMatrixCSC, SparseMatrixCSR, SymSparseMatrixCSR  # noqa: F401
from .assemblers import test_assembler, test_sparse_matrix_assembler  # noqa: F401
from .assemblers import (AffineFEOperator, B200SparseMatrixAssembler, DefaultAssemblyStrategy, FEOperator, GenericAssemblyStrategy, OwnedColumns, SparseMatrixAssembler,  # noqa: F401
                         assemble_matrix, assemble_matrix_and_vector, assemble_vector, collect_cell_matrix,
                         collect_cell_matrix_and_vector, collect_cell_vector, fill_cell_matrix, get_fe_basis, get_matrix,
                         get_trial_fe_basis, get_vector)
from .celldata import (Integral, IsotropicLinearElasticity, Measure, NeoHookean, div, dot, eps, grad, inner, jump, mean, nabla, ε)  # noqa: F401
from .fespaces import (BlockMultiFieldStyle, ConsecutiveMultiFieldStyle, FEFunction, FESpace, FESpaceWithLinearConstraints, has_constraints, MultiFieldFESpace, TestFESpace, TrialFESpace, interpolate, zero)  # noqa: F401
from .geometry import (Boundary, BoundaryTriangulation, Skeleton, SkeletonTriangulation, get_normal_vector, CartesianDiscreteModel, DiscreteModel, Triangulation, UnstructuredDiscreteModel, get_triangulation,  # noqa: F401
                       simplexify)
from .reffes import Quadrature, ReferenceFE, VectorValue, lagrangian  # noqa: F401
