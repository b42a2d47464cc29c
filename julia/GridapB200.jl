# GridapB200.jl -- host-side Julia shim: Gridap's SparseMatrixAssembler interface over libgridap_b200.so.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: `julia` is not installed in the build image.  Every ccall below is mirrored,
# argument for argument, by gridap.jl_b200/lib.py, which the GPU tests exercise.  Reference interface being implemented:
#   src/FESpaces/Assemblers.jl:155-257, src/FESpaces/SparseMatrixAssemblers.jl:4-106
module GridapB200

using Gridap
using Gridap.Arrays, Gridap.Fields, Gridap.Geometry, Gridap.ReferenceFEs, Gridap.CellData, Gridap.FESpaces, Gridap.Algebra
using SparseArrays, FillArrays

const LIB = get(ENV, "GRIDAP_B200_LIB", joinpath(@__DIR__, "..", "gridap.jl_b200", "lib", "libgridap_b200.so"))

# form ids (include/gridap_b200.h)
const FORM_MASS, FORM_LAPLACIAN, FORM_ELASTICITY, FORM_STOKES, FORM_NEOHOOKEAN_JAC = Int32(1), Int32(2), Int32(3), Int32(4), Int32(5)
const FORM_SOURCE, FORM_NEOHOOKEAN_RES = Int32(10), Int32(11)
const FORM_FACET, FORM_FACET_VEC = Int32(20), Int32(21)   # facet-of-cell plans: Nitsche / normal-flux terms
const ERR_UNSUPPORTED = Int32(-2)

last_error(ctx) = unsafe_string(ccall((:gb200_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
function check(ctx, rc)
  rc == 0 && return nothing
  rc == ERR_UNSUPPORTED && error("not implemented on the B200 path (no CPU fallback): " * last_error(ctx))
  error(last_error(ctx))
end

celltype_id(p::Polytope) = p == QUAD ? Int32(1) : p == HEX ? Int32(2) : p == TRI ? Int32(3) : p == TET ? Int32(4) :
  error("cell type $p is not supported by the B200 assembler")

mutable struct B200SparseMatrixAssembler <: SparseMatrixAssembler
  ctx::Ptr{Cvoid}
  mesh::Ptr{Cvoid}
  trial
  test
  rows::Base.OneTo{Int}
  cols::Base.OneTo{Int}
  plans::Dict{Any,Any}      # degree => (plan, refels, spaces)
  strategy::AssemblyStrategy   # DefaultAssemblyStrategy, or any strategy whose row/col map + mask are applied to the id tables
  meshes::Dict{UInt,Ptr{Cvoid}}  # view triangulations Triangulation(model, cell_ids): one mesh over the cells of each view
end

# gb200_mesh_create from a Grid: node coordinates + the cell -> node table of its cells (a view hands over its own cells only)
function mesh_create(ctx::Ptr{Cvoid}, grid)
  @assert length(get_reffes(grid)) == 1 "one cell type per mesh"
  x = get_node_coordinates(grid)
  c2n = Table(get_cell_node_ids(grid))
  mesh = Ref{Ptr{Cvoid}}(C_NULL)
  check(ctx, ccall((:gb200_mesh_create, LIB), Int32,
    (Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}),
    ctx, num_point_dims(grid), length(x), reinterpret(Float64, collect(x)), num_cells(grid), c2n.data, c2n.ptrs,
    celltype_id(get_polytope(first(get_reffes(grid)))), mesh))
  mesh[]
end

# the mesh a quadrature lives on: the space's own triangulation, or a view of it (benchmark/bm/bm_assembly.jl:30-33 runs every
# case on `Triangulation(model, collect(1:div(n^D,2)))` too) -- same node coordinates and global DoF ids, fewer cells
function mesh_for!(a, trian)
  trian === get_triangulation(a.test) && return a.mesh
  get!(() -> mesh_create(a.ctx, get_grid(trian)), a.meshes, objectid(trian))
end

# SparseMatrixAssembler(mat, vec, U, V, strategy) (src/FESpaces/SparseMatrixAssemblers.jl:127-153): a non-default strategy is
# honoured, never silently ignored -- its maps and masks are applied to the cell DoF tables on the host (mapped_ids below).
function B200SparseMatrixAssembler(U, V; device::Integer=0, deterministic::Bool=false, strategy::AssemblyStrategy=DefaultAssemblyStrategy())
  ctx = Ref{Ptr{Cvoid}}(C_NULL)
  rc = ccall((:gb200_init, LIB), Int32, (Int32, UInt32, Ref{Ptr{Cvoid}}), device, deterministic ? 1 : 0, ctx)
  rc == 0 || error(last_error(C_NULL))
  mesh = Ref(mesh_create(ctx[], get_grid(get_triangulation(V))))
  a = B200SparseMatrixAssembler(ctx[], mesh[], U, V, Base.OneTo(num_free_dofs(V)), Base.OneTo(num_free_dofs(U)), Dict(), strategy, Dict{UInt,Ptr{Cvoid}}())
  finalizer(free!, a)
end

function free!(a::B200SparseMatrixAssembler)
  for (_, (plan, refels, spaces)) in a.plans
    ccall((:gb200_plan_destroy, LIB), Int32, (Ptr{Cvoid},), plan)
    foreach(s -> ccall((:gb200_space_destroy, LIB), Int32, (Ptr{Cvoid},), s), spaces)
    foreach(r -> ccall((:gb200_refel_destroy, LIB), Int32, (Ptr{Cvoid},), r), refels)
  end
  foreach(m -> ccall((:gb200_mesh_destroy, LIB), Int32, (Ptr{Cvoid},), m), values(a.meshes))
  ccall((:gb200_mesh_destroy, LIB), Int32, (Ptr{Cvoid},), a.mesh)
  ccall((:gb200_finalize, LIB), Int32, (Ptr{Cvoid},), a.ctx)
  nothing
end

FESpaces.get_rows(a::B200SparseMatrixAssembler) = a.rows
FESpaces.get_cols(a::B200SparseMatrixAssembler) = a.cols
FESpaces.get_assembly_strategy(a::B200SparseMatrixAssembler) = a.strategy
FESpaces.get_matrix_builder(::B200SparseMatrixAssembler) = SparseMatrixBuilder(SparseMatrixCSC{Float64,Int})
FESpaces.get_vector_builder(::B200SparseMatrixAssembler) = ArrayBuilder(Vector{Float64})

# ---- tabulation: get_shapefuns / Quadrature evaluated once per reference element (a21)
function refel_create(a, reffe, quad, ncomp)
  xq, w = get_coordinates(quad), get_weights(quad)
  sreffe = ncomp == 1 ? reffe : LagrangianRefFE(Float64, get_polytope(reffe), get_orders(reffe))   # scalar basis; k = a + nd*(c-1)
  shapes = get_shapefuns(sreffe)
  N = evaluate(shapes, xq)                                   # Matrix{Float64} [np, nd]
  dN = evaluate(Broadcasting(∇)(shapes), xq)                 # Matrix{VectorValue{D,Float64}} [np, nd]
  D = num_dims(get_polytope(reffe))
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(a.ctx, ccall((:gb200_refel_create, LIB), Int32,
    (Ptr{Cvoid}, Int32, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
    a.ctx, D, length(w), size(N, 2), ncomp, w, N, reinterpret(Float64, dN), h))
  h[]
end

# `map_rows!` / `map_cols!` of the strategy (src/FESpaces/Assemblers.jl:31-55) on a whole id table.  On the wire a masked id is 0
# (neither free nor Dirichlet); negative (Dirichlet) ids are kept, the lifting needs them (the reference maps on the fly and keeps
# the Dirichlet values in the AttachDirichletMap).  `offset`: the field offset of a MultiFieldFESpace (the strategy sees global ids).
function mapped_ids(s::AssemblyStrategy, data::Vector{Int32}, rows::Bool, offset::Integer)
  s isa DefaultAssemblyStrategy && return data
  out = copy(data)
  for k in eachindex(data)
    id = Int(data[k])
    id > 0 || continue
    g = id + offset
    keep = rows ? FESpaces.row_mask(s, g) : FESpaces.col_mask(s, g)
    out[k] = keep ? Int32(rows ? FESpaces.row_map(s, g) : FESpaces.col_map(s, g)) : Int32(0)
  end
  out
end

# multi-GPU column ownership without host loops in Julia: the C helper masks / renumbers a trial id table for the columns a rank
# owns (`owned[j] == 1`); local column j of the rank's matrix is global column owned_ids[j]
function owned_column_ids(ids::Vector{Int32}, owned::Vector{UInt8})
  out = similar(ids); n = Ref{Int64}(0); oid = Vector{Int64}(undef, count(!iszero, owned))
  rc = ccall((:gb200_owned_column_ids, LIB), Int32, (Ptr{Int32}, Int64, Ptr{UInt8}, Int64, Ptr{Int32}, Ref{Int64}, Ptr{Int64}),
             ids, length(ids), owned, length(owned), out, n, oid)
  rc == 0 || error(last_error(C_NULL))
  out, oid
end

function space_create(a, refel, space; rows::Bool=true, offset::Integer=0, trian=get_triangulation(space))
  ids = Table(get_cell_dof_ids(space, trian))                # Table{Int32}: free > 0, Dirichlet < 0 (on a view: its cells only)
  ids = Table(mapped_ids(a.strategy, ids.data, rows, offset), ids.ptrs)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(a.ctx, ccall((:gb200_space_create, LIB), Int32,
    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Int64, Int64, Ref{Ptr{Cvoid}}),
    a.ctx, mesh_for!(a, trian), refel, ids.data, ids.ptrs, num_free_dofs(space), num_dirichlet_dofs(space), h))
  h[]
end

fields(space) = space isa MultiFieldFESpace ? collect(space.spaces) : [space]
foffs(s) = s isa MultiFieldFESpace ? Int64[0; cumsum(num_free_dofs.(s.spaces))[1:end-1]] : Int64[0]
ncomps(space) = num_components(eltype(get_free_dof_values(zero(space)))) # 1 or D

"symbolic phase, cached per quadrature (src/FESpaces/SparseMatrixAssemblers.jl:174-210 -> gb200_plan_create)"
function plan!(a::B200SparseMatrixAssembler, quad::CellQuadrature, touched::Matrix{UInt8})
  key = (objectid(quad), touched)
  haskey(a.plans, key) && return a.plans[key][1]
  q = first(quad.cell_quad.value isa Quadrature ? [quad.cell_quad.value] : quad.cell_quad)
  trian = quad.trian                                          # the space's triangulation or a view of it
  grid = get_grid(trian)
  geo = refel_create(a, first(get_reffes(grid)), q, 1)
  tests, trials, refels = Ptr{Cvoid}[], Ptr{Cvoid}[], Ptr{Cvoid}[geo]
  for (t, u) in zip(fields(a.test), fields(a.trial))
    @notimplementedif has_constraints(t) || has_constraints(u) "constrained spaces are not on the B200 path"
    r = refel_create(a, first(get_fe_basis(t).cell_basis.value.fields isa Any ? get_reffes(t) : get_reffes(t)), q, ncomps(t))
    default = a.strategy isa DefaultAssemblyStrategy
    k = length(tests) + 1
    push!(refels, r)
    push!(tests, space_create(a, r, t; rows=true, offset=default ? 0 : foffs(a.test)[k], trian=trian))
    push!(trials, (t === u && default) ? tests[end] : space_create(a, r, u; rows=false, offset=default ? 0 : foffs(a.trial)[k], trian=trian))
  end
  # with a strategy the ids on the wire already are global (offset added before the map): the plan gets zero offsets
  offs(s) = a.strategy isa DefaultAssemblyStrategy ? foffs(s) : zeros(Int64, length(fields(s)))
  plan = Ref{Ptr{Cvoid}}(C_NULL)
  check(a.ctx, ccall((:gb200_plan_create, LIB), Int32,
    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}, Int32, Ptr{Ptr{Cvoid}}, Ptr{UInt8}, Ptr{Int64}, Ptr{Int64}, Int64, Int64, Ref{Ptr{Cvoid}}),
    a.ctx, mesh_for!(a, trian), geo, length(tests), tests, length(trials), trials, touched, offs(a.test), offs(a.trial), length(a.rows), length(a.cols), plan))
  a.plans[key] = (plan[], refels, unique(vcat(tests, trials)))
  plan[]
end

# ---- integrand recogniser: walks the lazy tree like print_op_tree (src/Arrays/PrintOpTrees.jl:55-68)
struct Recognised
  form::Int32
  params::Vector{Float64}
  quad::CellQuadrature
  touched::Matrix{UInt8}
end

"""
cellmat is `lazy_map(IntegrationMap(), bx, w, Jtx)` (src/CellData/CellQuadratures.jl:157-160).  `bx.maps.value` is the
BroadcastingFieldOpMap of the integrand; its `.op` and the roots of its arguments identify the form:
  ⊙ / ⋅ of (∇v, ∇u)            -> LAPLACIAN        * of (v, u)               -> MASS
  ⊙ of (ε(v), σ∘ε(u))           -> ELASTICITY       blocks {∇v⊙∇u, -(∇⋅v)p, q(∇⋅u)} -> STOKES
Anything else: error -- the B200 assembler never evaluates the lazy array on the CPU.
"""
function recognise(cellmat)::Recognised
  cellmat isa Fill && return Recognised(Int32(0), vec(collect(cellmat.value)), nothing, ones(UInt8, 1, 1))  # Fill(K_e): scatter only
  cellmat isa LazyArray && cellmat.maps.value isa IntegrationMap || error("B200 assembler: unrecognised cell array $(typeof(cellmat))")
  bx = cellmat.args[1]
  match_form(bx)   # implemented per form in forms.jl: pattern match on bx.maps.value.op and the argument trees
end

function FESpaces.allocate_matrix(a::B200SparseMatrixAssembler, matdata)
  r = recognise(matdata[1][1])
  plan = plan!(a, r.quad, r.touched)
  nnz = Ref{Int64}(0)
  check(a.ctx, ccall((:gb200_plan_nnz, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), plan, nnz))
  colptr = Vector{Int}(undef, length(a.cols) + 1)
  rowval = Vector{Int}(undef, nnz[])
  check(a.ctx, ccall((:gb200_plan_get_pattern, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), plan, colptr, rowval))
  SparseMatrixCSC(length(a.rows), length(a.cols), colptr, rowval, zeros(Float64, nnz[]))
end

function _assemble_matrix!(A, a, matdata, add::Integer)
  @assert length(matdata[1]) == 1 "one triangulation per form on the B200 path"
  r = recognise(matdata[1][1])
  plan = plan!(a, r.quad, r.touched)
  if r.form == 0
    check(a.ctx, ccall((:gb200_assemble_matrix_const, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32), plan, r.params, nonzeros(A), add))
  else
    check(a.ctx, ccall((:gb200_assemble_matrix, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Int32),
      plan, r.form, r.params, length(r.params), nonzeros(A), add))
  end
  A
end
FESpaces.assemble_matrix!(A, a::B200SparseMatrixAssembler, matdata) = _assemble_matrix!(A, a, matdata, 0)
FESpaces.assemble_matrix_add!(A, a::B200SparseMatrixAssembler, matdata) = _assemble_matrix!(A, a, matdata, 1)

FESpaces.allocate_vector(a::B200SparseMatrixAssembler, vecdata) = zeros(Float64, length(a.rows))

function _assemble_vector!(b, a, vecdata, add::Integer)
  r = recognise_vector(vecdata[1][1])                       # source term: f evaluated at x_q on the host
  plan = plan!(a, r.quad, r.touched)
  fq = r.f === nothing ? C_NULL : begin
    np = length(get_weights(first(r.quad.cell_quad))); nc = num_cells(get_triangulation(a.test)); D = num_point_dims(get_triangulation(a.test))
    xq = Vector{Float64}(undef, D * np * nc)
    check(a.ctx, ccall((:gb200_quadrature_points, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), plan, xq))
    collect(reinterpret(Float64, r.f.(reinterpret(Point{D,Float64}, xq))))
  end
  check(a.ctx, ccall((:gb200_assemble_vector, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Int32),
    plan, r.form, r.params, length(r.params), fq, b, add))
  b
end
FESpaces.assemble_vector!(b, a::B200SparseMatrixAssembler, vecdata) = _assemble_vector!(b, a, vecdata, 0)
FESpaces.assemble_vector_add!(b, a::B200SparseMatrixAssembler, vecdata) = _assemble_vector!(b, a, vecdata, 1)

function FESpaces.allocate_matrix_and_vector(a::B200SparseMatrixAssembler, data)
  (allocate_matrix(a, (map(first ∘ unpair, data[1][1]), data[1][2], data[1][3])), zeros(Float64, length(a.rows)))
end

function _assemble_matrix_and_vector!(A, b, a, data, add::Integer)
  matvecdata, matdata, vecdata = data
  # root map AttachDirichletMap over lazy_map(tuple, cellmat, cellvec) (src/CellData/AttachDirichlet.jl:5-8, src/Arrays/ArrayPairs.jl:3-5)
  cellmat, cellvec, dirichlet_values = unpack_attach_dirichlet(matvecdata[1][1])
  rm, rv = recognise(cellmat), recognise_vector(cellvec)
  plan = plan!(a, rm.quad, rm.touched)
  check(a.ctx, ccall((:gb200_plan_set_state, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}), plan, 0, C_NULL, dirichlet_values))
  check(a.ctx, ccall((:gb200_assemble_matrix_and_vector, LIB), Int32,
    (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32),
    plan, rm.form, rm.params, length(rm.params), rv.form, rv.params, length(rv.params), C_NULL, nonzeros(A), b, add))
  isempty(matdata[1]) || _assemble_matrix!(A, a, matdata, 1)      # leftover un-paired terms (SparseMatrixAssemblers.jl:399-403)
  isempty(vecdata[1]) || _assemble_vector!(b, a, vecdata, 1)
  A, b
end
FESpaces.assemble_matrix_and_vector!(A, b, a::B200SparseMatrixAssembler, data) = _assemble_matrix_and_vector!(A, b, a, data, 0)
FESpaces.assemble_matrix_and_vector_add!(A, b, a::B200SparseMatrixAssembler, data) = _assemble_matrix_and_vector!(A, b, a, data, 1)

# assemble_matrix(a, matdata) = allocate + numeric (src/FESpaces/SparseMatrixAssemblers.jl:70-77): the pattern download is only
# enqueued (second stream) and completes inside the numeric call's synchronisation, so it overlaps the numeric phase.
function FESpaces.assemble_matrix(a::B200SparseMatrixAssembler, matdata)
  r = recognise(matdata[1][1])
  plan = plan!(a, r.quad, r.touched)
  nnz = Ref{Int64}(0)
  check(a.ctx, ccall((:gb200_plan_nnz, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), plan, nnz))
  colptr = pinned_vector(a, Int, length(a.cols) + 1)       # gb200_host_alloc + unsafe_wrap (freed by a finalizer)
  rowval = pinned_vector(a, Int, nnz[])
  nzval = pinned_vector(a, Float64, nnz[])
  check(a.ctx, ccall((:gb200_plan_get_pattern_async, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), plan, colptr, rowval))
  A = SparseMatrixCSC(length(a.rows), length(a.cols), colptr, rowval, nzval)
  _assemble_matrix!(A, a, matdata, 0)                      # synchronises both streams
end

# BlockMultiFieldStyle(): a BlockMatrix with one SparseMatrixCSC per field block, cut out of the single device matrix
# (src/MultiField/BlockSparseMatrixAssemblers.jl:19-33,197-230).  Same plan, same numeric call (device-resident, nzval = C_NULL).
struct B200BlockSparseMatrixAssembler <: SparseMatrixAssembler
  inner::B200SparseMatrixAssembler
  row_sizes::Vector{Int}; col_sizes::Vector{Int}
end
FESpaces.get_rows(a::B200BlockSparseMatrixAssembler) = map(Base.OneTo, a.row_sizes)
FESpaces.get_cols(a::B200BlockSparseMatrixAssembler) = map(Base.OneTo, a.col_sizes)

function block_csc(a::B200BlockSparseMatrixAssembler, plan, bi, bj; values=true)
  nnz = Ref{Int64}(0)
  check(a.inner.ctx, ccall((:gb200_plan_block_nnz, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ref{Int64}), plan, bi - 1, bj - 1, nnz))
  colptr = Vector{Int}(undef, a.col_sizes[bj] + 1); rowval = Vector{Int}(undef, nnz[]); nzval = zeros(Float64, nnz[])
  check(a.inner.ctx, ccall((:gb200_plan_get_block_pattern, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int64}, Ptr{Int64}), plan, bi - 1, bj - 1, colptr, rowval))
  values && nnz[] > 0 && check(a.inner.ctx, ccall((:gb200_plan_download_block, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}), plan, bi - 1, bj - 1, nzval))
  SparseMatrixCSC(a.row_sizes[bi], a.col_sizes[bj], colptr, rowval, nzval)
end

function FESpaces.assemble_matrix(a::B200BlockSparseMatrixAssembler, matdata)
  r = recognise(matdata[1][1])
  plan = plan!(a.inner, r.quad, r.touched)
  check(a.inner.ctx, ccall((:gb200_assemble_matrix, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Int32),
    plan, r.form, r.params, length(r.params), C_NULL, 0))
  mortar([block_csc(a, plan, i, j) for i in eachindex(a.row_sizes), j in eachindex(a.col_sizes)])   # BlockArrays.mortar
end

# Forms that carry u_h (residual / Jacobian) on an assembler with a strategy: the plan's trial ids are mapped / masked, u_h lives on the
# global trial space -> gather it through an unmapped space (same mesh, same reference element)
function set_state_space!(a::B200SparseMatrixAssembler, plan, field::Integer, refel, trial_space)
  a.strategy isa DefaultAssemblyStrategy && return nothing
  ids = Table(get_cell_dof_ids(trial_space))
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(a.ctx, ccall((:gb200_space_create, LIB), Int32,
    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Int64, Int64, Ref{Ptr{Cvoid}}),
    a.ctx, a.mesh, refel, ids.data, ids.ptrs, num_free_dofs(trial_space), num_dirichlet_dofs(trial_space), h))
  check(a.ctx, ccall((:gb200_plan_set_state_space, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}), plan, field, h[]))
  h[]
end

# Nitsche / normal-flux terms on a BoundaryTriangulation Γ (test/GridapTests/PoissonTests.jl:99-107): the plan lives on the cells
# adjacent to the facets.  mesh: gb200_mesh_create with get_cell_node_ids(model)[Γ.glue.face_to_cell]; spaces: the cell DoF tables of
# those cells; tabulations: the facet rule mapped onto every local face (compute_face_to_cell_reference_map,
# src/Geometry/BoundaryTriangulations.jl:320-340), one block of points per local face; then
function set_facets!(a::B200SparseMatrixAssembler, plan, Γ::BoundaryTriangulation, nref::Vector{Float64})
  lface = Int32.(Γ.glue.face_to_lface)                       # 1-based local face of every facet (FaceToCellGlue)
  nlf = div(length(nref), num_point_dims(Γ))
  check(a.ctx, ccall((:gb200_plan_set_facets, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Int32, Ptr{Float64}), plan, lface, nlf, nref))
end
# and the terms are GB200_FORM_FACET {coef, test kind, trial kind} / GB200_FORM_FACET_VEC {coef, test kind, data kind} through
# gb200_assemble_matrix / gb200_assemble_vector on that plan, merged into the bulk matrix by gb200_plan_add_matrix_from.

# jump / mean terms on a SkeletonTriangulation Λ (test/GridapTests/PoissonDGTests.jl:42-45): a two-field plan, field 1 = the space on
# the plus cells (Λ.plus.glue.face_to_cell), field 2 = the same space on the minus cells (two gb200_mesh_create / gb200_space_create
# calls, offsets 0, all four blocks touched); tabulations as for set_facets!.  perm[p, facet] (0-based) = the point of the minus
# cell's local-face block that coincides with point p of the plus side: the plus / minus rows of get_cell_points(Λ) mapped by
# FaceToCellGlue (cell_to_lface_to_pindex, src/Geometry/BoundaryTriangulations.jl:42-70) compared in physical space.
function set_skeleton!(a::B200SparseMatrixAssembler, plan, Λ::SkeletonTriangulation, perm::Matrix{Int32}, nref::Vector{Float64})
  lplus, lminus = Int32.(Λ.plus.glue.face_to_lface), Int32.(Λ.minus.glue.face_to_lface)
  nlf = div(length(nref), num_point_dims(Λ))
  check(a.ctx, ccall((:gb200_plan_set_skeleton, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Float64}),
                     plan, lplus, lminus, perm, nlf, nref))
end
# terms: GB200_FORM_SKELETON {coef, T kind, w+, w-, U kind, z+, z-} (= 22) through gb200_assemble_matrix on that plan; the matrix of a
# form with skeleton terms lives in a pattern-only pair plan over (cell, cell) for every cell and (plus, minus) for every interior
# facet (the symbolic loop over all contributions, src/FESpaces/SparseMatrixAssemblers.jl:174-210), filled by gb200_plan_add_matrix_from.

# FESpaceWithLinearConstraints (src/FESpaces/FESpacesWithLinearConstraints.jl): the unconstrained space f.space is assembled with its
# free and Dirichlet DoFs in one numbering (DOF = dof > 0 ? dof : n_fdofs - dof, :356-362) on `src`; `dst` is a plan created from
# get_cell_dof_ids(f) (masters per cell, padded with 0); the tables are f.DOF_to_mDOFs / f.DOF_to_coeffs with master ids signed by
# _DOF_to_dof(mDOF, f.n_fmdofs) (:364-372)
function fold_constraints!(a::B200SparseMatrixAssembler, dst, src, f::FESpaceWithLinearConstraints, dirichlet_master_values::Vector{Float64};
                           matrix::Bool=true, vector::Bool=true)
  ptrs = Int64.(f.DOF_to_mDOFs.ptrs)
  mdofs = Int32[m > f.n_fmdofs ? -(m - f.n_fmdofs) : m for m in f.DOF_to_mDOFs.data]
  check(a.ctx, ccall((:gb200_plan_fold_constraints, LIB), Int32,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Int64, Int32, Int32),
                     dst, src, ptrs, mdofs, Float64.(f.DOF_to_coeffs.data), dirichlet_master_values, length(dirichlet_master_values), matrix, vector))
end
# (a vector accumulated over several triangulations on the host is handed back first: gb200_plan_upload_vector(src, b))
upload_vector!(a::B200SparseMatrixAssembler, plan, b::Vector{Float64}) =
  check(a.ctx, ccall((:gb200_plan_upload_vector, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), plan, b))

# residual_and_jacobian! (src/FESpaces/FEOperatorsFromWeakForm.jl:85-103) for the neo-Hookean pair: one fused pass, no lifting
function fused_residual_and_jacobian!(b, A, a::B200SparseMatrixAssembler, plan, params, free_values, dirichlet_values)
  check(a.ctx, ccall((:gb200_plan_set_state, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}), plan, 0, free_values, dirichlet_values))
  check(a.ctx, ccall((:gb200_assemble_matrix_and_vector, LIB), Int32,
    (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32),
    plan, 5, params, length(params), 11, params, length(params), C_NULL, nonzeros(A), b, 0))   # NEOHOOKEAN_JAC = 5, NEOHOOKEAN_RES = 11
  b, A
end

export B200SparseMatrixAssembler, B200BlockSparseMatrixAssembler
end # module
