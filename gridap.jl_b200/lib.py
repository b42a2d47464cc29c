"""ctypes binding of libgridap_b200.so -- the same entry points the Julia shim `ccall`s.

The product never computes on the CPU: if the library is missing, or no CUDA device is present,
every call raises.
"""
import ctypes as C
import json
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgridap_b200.so")

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_STATE = 0, -1, -2, -3, -4
QUAD4, HEX8, TRI3, TET4, SEG2 = 1, 2, 3, 4, 5
FORM_NONE, FORM_MASS, FORM_LAPLACIAN, FORM_ELASTICITY, FORM_STOKES, FORM_NEOHOOKEAN_JAC = 0, 1, 2, 3, 4, 5
FORM_SOURCE, FORM_NEOHOOKEAN_RES = 10, 11
FORM_SKELETON = 22   # skeleton plans: coef [w+ T(v+) + w- T(v-)] [z+ U(u+) + z- U(u-)]
FORM_FACET, FORM_FACET_VEC = 20, 21   # facet-of-cell plans: coef T(v) U(u) / coef T(v) d, kinds 0 = value, 1 = normal derivative
FLAG_DETERMINISTIC = 1

SYMBOLS = [
    "gb200_init", "gb200_finalize", "gb200_last_error", "gb200_version", "gb200_get_timings", "gb200_launch_count",
    "gb200_stream", "gb200_synchronize", "gb200_host_alloc", "gb200_host_free", "gb200_host_register", "gb200_host_unregister", "gb200_trim", "gb200_mesh_create", "gb200_mesh_destroy", "gb200_mesh_is_affine",
    "gb200_refel_create", "gb200_refel_destroy", "gb200_space_create", "gb200_space_destroy", "gb200_plan_create",
    "gb200_plan_destroy", "gb200_plan_set_facets", "gb200_plan_set_skeleton", "gb200_plan_fold_constraints", "gb200_plan_upload_vector", "gb200_plan_nnz", "gb200_plan_get_pattern", "gb200_plan_get_pattern_async", "gb200_plan_set_state", "gb200_plan_set_state_device", "gb200_plan_set_state_space", "gb200_assemble_matrix",
    "gb200_assemble_matrix_const", "gb200_assemble_vector", "gb200_assemble_matrix_and_vector", "gb200_quadrature_points",
    "gb200_plan_add_matrix_from", "gb200_plan_get_csr_pattern", "gb200_plan_download_csr", "gb200_plan_block_nnz", "gb200_plan_get_block_pattern", "gb200_plan_download_block", "gb200_plan_device_nzval", "gb200_plan_device_pattern", "gb200_plan_device_vector", "gb200_plan_download", "gb200_plan_kernel_path", "gb200_owned_column_ids",
]


class GridapB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libgridap_b200 error %d: %s" % (code, msg))
        self.code = code


class UnsupportedError(GridapB200Error, NotImplementedError):
    """The `@notimplemented` of the reference: integrand / space outside the supported set (never a CPU fallback)."""


_lib = None


def load():
    """dlopen the CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GridapB200Error(ERR_STATE, "%s not found: build it with `python gridap.jl_b200/build.py` "
                              "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
    pvp = C.POINTER(C.c_void_p)
    L.gb200_init.argtypes = [i32, u32, pvp]
    L.gb200_finalize.argtypes = [vp]
    L.gb200_last_error.argtypes = [vp]
    L.gb200_last_error.restype = C.c_char_p
    L.gb200_version.restype = C.c_char_p
    L.gb200_get_timings.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.gb200_launch_count.argtypes = [vp]
    L.gb200_launch_count.restype = i64
    L.gb200_stream.argtypes = [vp]
    L.gb200_stream.restype = vp
    L.gb200_synchronize.argtypes = [vp]
    L.gb200_host_alloc.argtypes = [vp, C.c_size_t, pvp]
    L.gb200_host_free.argtypes = [vp, vp]
    L.gb200_host_register.argtypes = [vp, vp, C.c_size_t]
    L.gb200_host_unregister.argtypes = [vp, vp]
    L.gb200_trim.argtypes = [vp]
    L.gb200_mesh_create.argtypes = [vp, i32, i64, vp, i64, vp, vp, i32, pvp]
    L.gb200_mesh_destroy.argtypes = [vp]
    L.gb200_mesh_is_affine.argtypes = [vp, C.POINTER(i32)]
    L.gb200_refel_create.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, pvp]
    L.gb200_refel_destroy.argtypes = [vp]
    L.gb200_space_create.argtypes = [vp, vp, vp, vp, vp, i64, i64, pvp]
    L.gb200_space_destroy.argtypes = [vp]
    L.gb200_plan_create.argtypes = [vp, vp, vp, i32, pvp, i32, pvp, vp, vp, vp, i64, i64, pvp]
    L.gb200_plan_destroy.argtypes = [vp]
    L.gb200_plan_set_facets.argtypes = [vp, vp, i32, vp]
    L.gb200_plan_set_skeleton.argtypes = [vp, vp, vp, vp, i32, vp]
    L.gb200_plan_fold_constraints.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, i32]
    L.gb200_plan_upload_vector.argtypes = [vp, vp]
    L.gb200_plan_nnz.argtypes = [vp, C.POINTER(i64)]
    L.gb200_plan_get_pattern.argtypes = [vp, vp, vp]
    L.gb200_plan_get_pattern_async.argtypes = [vp, vp, vp]
    L.gb200_plan_add_matrix_from.argtypes = [vp, vp]
    L.gb200_plan_get_csr_pattern.argtypes = [vp, i32, vp, vp]
    L.gb200_plan_download_csr.argtypes = [vp, vp]
    L.gb200_plan_block_nnz.argtypes = [vp, i32, i32, C.POINTER(i64)]
    L.gb200_plan_get_block_pattern.argtypes = [vp, i32, i32, vp, vp]
    L.gb200_plan_download_block.argtypes = [vp, i32, i32, vp]
    L.gb200_plan_set_state.argtypes = [vp, i32, vp, vp]
    L.gb200_plan_set_state_device.argtypes = [vp, i32, vp, vp]
    L.gb200_plan_set_state_space.argtypes = [vp, i32, vp]
    L.gb200_assemble_matrix.argtypes = [vp, i32, vp, i32, vp, i32]
    L.gb200_assemble_matrix_const.argtypes = [vp, vp, vp, i32]
    L.gb200_assemble_vector.argtypes = [vp, i32, vp, i32, vp, vp, i32]
    L.gb200_assemble_matrix_and_vector.argtypes = [vp, i32, vp, i32, i32, vp, i32, vp, vp, vp, i32]
    L.gb200_quadrature_points.argtypes = [vp, vp]
    L.gb200_plan_device_nzval.argtypes = [vp, pvp, C.POINTER(i64)]
    L.gb200_plan_device_pattern.argtypes = [vp, pvp, pvp]
    L.gb200_plan_device_vector.argtypes = [vp, pvp, C.POINTER(i64)]
    L.gb200_plan_download.argtypes = [vp, vp, vp]
    L.gb200_plan_kernel_path.argtypes = [vp, i32]
    L.gb200_owned_column_ids.argtypes = [vp, i64, vp, i64, vp, C.POINTER(i64), vp]
    L.gb200_plan_kernel_path.restype = C.c_char_p
    for name in SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is C.c_int:  # default
            fn.restype = i32
    L.gb200_last_error.restype = C.c_char_p
    L.gb200_version.restype = C.c_char_p
    L.gb200_plan_kernel_path.restype = C.c_char_p
    L.gb200_launch_count.restype = i64
    L.gb200_stream.restype = vp
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def check(rc, ctx=None):
    if rc == OK:
        return
    msg = load().gb200_last_error(ctx).decode()
    if rc == ERR_UNSUPPORTED:
        raise UnsupportedError(rc, msg)
    raise GridapB200Error(rc, msg)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """gb200_init / gb200_finalize: one per GPU."""

    def __init__(self, device=0, deterministic=False):
        L = load()
        h = C.c_void_p()
        check(L.gb200_init(device, FLAG_DETERMINISTIC if deterministic else 0, C.byref(h)))
        self.h = h
        self.device = device
        self.deterministic = deterministic

    def timings(self):
        buf = C.create_string_buffer(4096)
        load().gb200_get_timings(self.h, buf, 4096)
        return json.loads(buf.value.decode() or "{}")

    def launch_count(self):
        return int(load().gb200_launch_count(self.h))

    def synchronize(self):
        check(load().gb200_synchronize(self.h), self.h)

    def synchronize_quiet(self):
        """error paths: wait for pending asynchronous copies (into pooled pinned arrays) without raising a second error"""
        try:
            load().gb200_synchronize(self.h)
        except Exception:
            pass

    def stream(self):
        return load().gb200_stream(self.h)

    def trim(self):
        """return the pooled (freed) device blocks to the driver"""
        check(load().gb200_trim(self.h), self.h)

    # -- page-locked host memory (full-rate PCIe copies): a small pool of pinned blocks reused across calls
    def pinned_empty(self, n, dtype):
        """1-D numpy array of `n` items in page-locked memory; the block returns to the pool when the array dies."""
        dtype = np.dtype(dtype)
        nbytes = max(int(n) * dtype.itemsize, 1)
        pool = self.__dict__.setdefault("_pinned_pool", {})
        free = pool.setdefault(nbytes, [])
        if free:
            ptr = free.pop()
        else:
            p = C.c_void_p()
            check(load().gb200_host_alloc(self.h, nbytes, C.byref(p)), self.h)
            ptr = p.value
        buf = (C.c_char * nbytes).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=int(n))
        weakref.finalize(buf, free.append, ptr)
        return arr

    def pin(self, arr):
        """cudaHostRegister a numpy array in place (unregistered again when the array is garbage collected)."""
        if arr is None or arr.nbytes == 0:
            return arr
        reg = self.__dict__.setdefault("_registered", set())
        key = (arr.ctypes.data, arr.nbytes)
        if key in reg:
            return arr
        rc = load().gb200_host_register(self.h, arr.ctypes.data_as(C.c_void_p), arr.nbytes)
        if rc == OK:
            reg.add(key)
            h, L = self.h, load()
            base = arr if arr.base is None else arr.base

            def _unpin(ptr=arr.ctypes.data, key=key):
                reg.discard(key)
                if self.h:
                    L.gb200_host_unregister(h, C.c_void_p(ptr))
            try:
                weakref.finalize(base, _unpin)
            except TypeError:
                pass
        return arr

    def close(self):
        if self.h:
            load().gb200_finalize(self.h)
            self.h = None


def owned_column_ids(ids, owned):
    """gb200_owned_column_ids: (masked / renumbered ids, 1-based global ids of the owned columns) -- the host helper a binding without
    numpy (the Julia shim) uses for the multi-GPU column ownership"""
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    owned = np.ascontiguousarray(owned, dtype=np.uint8)
    out = np.empty_like(ids)
    n = C.c_int64(0)
    oid = np.zeros(int(owned.sum()), dtype=np.int64)
    check(load().gb200_owned_column_ids(_ptr(ids), ids.size, _ptr(owned), owned.size, _ptr(out), C.byref(n), _ptr(oid)))
    return out, oid[:n.value]


_default_ctx = {}


def default_context(device=None, deterministic=False):
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", os.environ.get("GB200_DEVICE", "0")))
    key = (device, bool(deterministic))
    if key not in _default_ctx:
        _default_ctx[key] = Context(device, deterministic)
    return _default_ctx[key]


_ptrs_cache = {}


def table_ptrs(nrows, rowlen):
    """`ptrs` of a Table{Int32} with `rowlen` entries in every row (src/Arrays/Tables.jl:21-28).  Gridap's Table carries
    its ptrs; the numpy mirror keeps cell arrays as 2-D arrays, so the ptrs are built once per shape and kept."""
    key = (int(nrows), int(rowlen))
    if key not in _ptrs_cache:
        if len(_ptrs_cache) > 16:
            _ptrs_cache.clear()
        _ptrs_cache[key] = (1 + rowlen * np.arange(nrows + 1, dtype=np.int64)).astype(np.int32)
    return _ptrs_cache[key]


class DeviceMesh:
    def __init__(self, ctx, coords, cell_nodes, celltype):
        coords = f64(coords)
        cell_nodes = np.ascontiguousarray(cell_nodes, dtype=np.int32)
        nc, nn = cell_nodes.shape
        ptrs = table_ptrs(nc, nn)
        ctx.pin(coords)
        ctx.pin(cell_nodes)
        h = C.c_void_p()
        check(load().gb200_mesh_create(ctx.h, coords.shape[1], coords.shape[0], _ptr(coords), nc, _ptr(cell_nodes), _ptr(ptrs),
                                       celltype, C.byref(h)), ctx.h)
        self.h, self.ctx = h, ctx
        self.ncells, self.D = nc, coords.shape[1]

    def is_affine(self):
        r = C.c_int32(0)
        check(load().gb200_mesh_is_affine(self.h, C.byref(r)), self.ctx.h)
        return bool(r.value)

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                load().gb200_mesh_destroy(self.h)
        except Exception:
            pass


class DeviceRefEl:
    def __init__(self, ctx, w, N, dN, ncomp=1):
        """w[np], N[np,nd], dN[np,nd,D] (row-major numpy) -> Julia column-major layout on the wire."""
        w, N, dN = f64(w), f64(N), f64(dN)
        np_, nd, D = dN.shape
        Nw = np.asfortranarray(N)  # N[p + np*a]
        dNw = np.ascontiguousarray(np.transpose(dN, (1, 0, 2)))  # [a][p][d] -> d + D*(p + np*a)
        h = C.c_void_p()
        check(load().gb200_refel_create(ctx.h, D, np_, nd, ncomp, _ptr(w), Nw.ctypes.data_as(C.c_void_p), _ptr(dNw), C.byref(h)), ctx.h)
        self.h, self.ctx = h, ctx
        self.np, self.nd, self.ncomp, self.D = np_, nd, ncomp, D

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                load().gb200_refel_destroy(self.h)
        except Exception:
            pass


class DeviceSpace:
    def __init__(self, ctx, mesh, refel, cell_dofs, nfree, ndir):
        cell_dofs = np.ascontiguousarray(cell_dofs, dtype=np.int32)
        nc, nld = cell_dofs.shape
        ptrs = table_ptrs(nc, nld)
        ctx.pin(cell_dofs)
        h = C.c_void_p()
        check(load().gb200_space_create(ctx.h, mesh.h, refel.h, _ptr(cell_dofs), _ptr(ptrs), nfree, ndir, C.byref(h)), ctx.h)
        self.h, self.ctx, self.mesh, self.refel = h, ctx, mesh, refel
        self.nfree, self.ndir = nfree, ndir

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                load().gb200_space_destroy(self.h)
        except Exception:
            pass


class DeviceArray:
    """A 1-D device array owned by a plan, exported through the CUDA array interface (version 3)."""

    def __init__(self, ptr, n, typestr, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr or 0), False), "version": 3,
                                         "stream": int(owner.ctx.stream() or 0) or None}


class DevicePlan:
    """gb200_plan_*: symbolic phase + persistent device state for re-assembly."""

    def __init__(self, ctx, mesh, geo, test_spaces, trial_spaces, touched, row_offsets, col_offsets, nrows, ncols):
        nf = len(test_spaces)
        ts = (C.c_void_p * nf)(*[s.h for s in test_spaces])
        us = (C.c_void_p * nf)(*[s.h for s in trial_spaces])
        tch = None if touched is None else np.asfortranarray(np.asarray(touched, dtype=np.uint8))
        ro = np.asarray(row_offsets, dtype=np.int64)
        co = np.asarray(col_offsets, dtype=np.int64)
        h = C.c_void_p()
        check(load().gb200_plan_create(ctx.h, mesh.h, geo.h, nf, ts, nf, us, None if tch is None else tch.ctypes.data_as(C.c_void_p),
                                       _ptr(ro), _ptr(co), nrows, ncols, C.byref(h)), ctx.h)
        self.h, self.ctx = h, ctx
        self.keep = (mesh, geo, list(test_spaces), list(trial_spaces))
        self.nrows, self.ncols = nrows, ncols
        n = C.c_int64(0)
        check(load().gb200_plan_nnz(h, C.byref(n)), ctx.h)
        self.nnz = n.value
        self.symbolic_timings = ctx.timings()
        self.ncells, self.np, self.D = mesh.ncells, geo.np, mesh.D   # D: space dimension (quadrature points are physical points)

    def pattern(self, wait=True):
        """colptr, rowval (1-based Int64) in page-locked memory.  wait=False: the copy is only enqueued; the next numeric call
        with a host array (or ctx.synchronize()) completes it."""
        colptr = self.ctx.pinned_empty(self.ncols + 1, np.int64)
        rowval = self.ctx.pinned_empty(self.nnz, np.int64)
        fn = load().gb200_plan_get_pattern if wait else load().gb200_plan_get_pattern_async
        check(fn(self.h, _ptr(colptr), _ptr(rowval)), self.ctx.h)
        return colptr, rowval

    def set_facets(self, lface, nref):
        """facet-of-cell plan (gb200_plan_set_facets): lface[ncells] 1-based local face of every facet, nref[nlf, D]"""
        lf = np.ascontiguousarray(lface, dtype=np.int32)
        nr = f64(nref)
        check(load().gb200_plan_set_facets(self.h, _ptr(lf), nr.shape[0], _ptr(nr)), self.ctx.h)
        self.np //= nr.shape[0]

    def set_skeleton(self, lface_plus, lface_minus, perm, nref):
        """skeleton plan (gb200_plan_set_skeleton): 1-based local faces of every interior facet in its plus / minus cell, the point
        permutation of the minus side, nref[nlf, D]"""
        lp = np.ascontiguousarray(lface_plus, dtype=np.int32)
        lm = np.ascontiguousarray(lface_minus, dtype=np.int32)
        pm = np.ascontiguousarray(perm, dtype=np.int32)
        nr = f64(nref)
        check(load().gb200_plan_set_skeleton(self.h, _ptr(lp), _ptr(lm), _ptr(pm), nr.shape[0], _ptr(nr)), self.ctx.h)
        self.np //= nr.shape[0]

    def upload_vector(self, b):
        bb = f64(b)
        check(load().gb200_plan_upload_vector(self.h, _ptr(bb)), self.ctx.h)

    def fold_constraints_from(self, other, ptrs, mdofs, coeffs, dirichlet_master_values, with_matrix, with_vector):
        """device arrays := T^T (arrays of `other`) T (gb200_plan_fold_constraints): `other` is the plan of the unconstrained space
        with free and Dirichlet DoFs in one numbering; ptrs (1-based Int64), mdofs (signed i32), coeffs: DOF -> master DoFs"""
        p = np.ascontiguousarray(ptrs, dtype=np.int64)
        m = np.ascontiguousarray(mdofs, dtype=np.int32)
        c = f64(coeffs)
        dv = None if dirichlet_master_values is None else f64(dirichlet_master_values)
        check(load().gb200_plan_fold_constraints(self.h, other.h, _ptr(p), _ptr(m), _ptr(c), None if dv is None else _ptr(dv),
                                                 0 if dv is None else len(dv), int(with_matrix), int(with_vector)), self.ctx.h)

    def add_matrix_from(self, other):
        """device matrix += the device matrix of `other` (a plan on another triangulation whose pattern is contained in this one)"""
        check(load().gb200_plan_add_matrix_from(self.h, other.h), self.ctx.h)

    # -- SparseMatrixCSR view
    def csr_pattern(self, index_base):
        rowptr = np.zeros(self.nrows + 1, dtype=np.int64)
        colval = np.zeros(self.nnz, dtype=np.int64)
        check(load().gb200_plan_get_csr_pattern(self.h, index_base, _ptr(rowptr), _ptr(colval)), self.ctx.h)
        return rowptr, colval

    def download_csr(self, nzval):
        check(load().gb200_plan_download_csr(self.h, _ptr(nzval)), self.ctx.h)
        return nzval

    # -- BlockMultiFieldStyle views (one CSC per field block)
    def block_nnz(self, bi, bj):
        n = C.c_int64(0)
        check(load().gb200_plan_block_nnz(self.h, bi, bj, C.byref(n)), self.ctx.h)
        return n.value

    def block_pattern(self, bi, bj, ncols_b):
        colptr = np.zeros(ncols_b + 1, dtype=np.int64)
        rowval = np.zeros(self.block_nnz(bi, bj), dtype=np.int64)
        check(load().gb200_plan_get_block_pattern(self.h, bi, bj, _ptr(colptr), _ptr(rowval)), self.ctx.h)
        return colptr, rowval

    def download_block(self, bi, bj, nzval):
        if len(nzval):
            check(load().gb200_plan_download_block(self.h, bi, bj, _ptr(nzval)), self.ctx.h)
        return nzval

    def set_state(self, field, free_values, dirichlet_values):
        fv = None if free_values is None else f64(free_values)
        dv = None if dirichlet_values is None else f64(dirichlet_values)
        check(load().gb200_plan_set_state(self.h, field, _ptr(fv), _ptr(dv)), self.ctx.h)

    def set_state_space(self, field, space):
        """gather u_h through the (unmasked, global) ids of `space` instead of the plan's trial ids (owned-column plans)"""
        check(load().gb200_plan_set_state_space(self.h, field, None if space is None else space.h), self.ctx.h)
        self.keep = self.keep + (space,)

    def set_state_device(self, field, d_free=None, d_dirichlet=None):
        """device pointers (ints) or objects with `__cuda_array_interface__` / `.data_ptr()`; copied device-to-device on the
        context stream (the producer's stream must have been synchronised with it)"""
        def ptr(x):
            if x is None:
                return None
            if isinstance(x, int):
                return C.c_void_p(x)
            if hasattr(x, "data_ptr"):
                return C.c_void_p(x.data_ptr())
            return C.c_void_p(x.__cuda_array_interface__["data"][0])
        check(load().gb200_plan_set_state_device(self.h, field, ptr(d_free), ptr(d_dirichlet)), self.ctx.h)

    def assemble_matrix(self, form, params=(), nzval=None, add=False):
        p = f64(list(params))
        check(load().gb200_assemble_matrix(self.h, form, _ptr(p), len(p), _ptr(nzval), int(add)), self.ctx.h)
        return nzval

    def assemble_matrix_const(self, Ke, nzval=None, add=False):
        KeF = np.asfortranarray(Ke, dtype=np.float64)
        check(load().gb200_assemble_matrix_const(self.h, KeF.ctypes.data_as(C.c_void_p), _ptr(nzval), int(add)), self.ctx.h)
        return nzval

    def assemble_vector(self, form, params=(), fq=None, b=None, add=False):
        p = f64(list(params))
        fq = None if fq is None else f64(fq)
        check(load().gb200_assemble_vector(self.h, form, _ptr(p), len(p), _ptr(fq), _ptr(b), int(add)), self.ctx.h)
        return b

    def assemble_matrix_and_vector(self, form_mat, mat_params, form_vec, vec_params, fq=None, nzval=None, b=None, add=False):
        mp, vp = f64(list(mat_params)), f64(list(vec_params))
        fq = None if fq is None else f64(fq)
        check(load().gb200_assemble_matrix_and_vector(self.h, form_mat, _ptr(mp), len(mp), form_vec, _ptr(vp), len(vp), _ptr(fq),
                                                      _ptr(nzval), _ptr(b), int(add)), self.ctx.h)
        return nzval, b

    def quadrature_points(self):
        xq = np.zeros((self.ncells, self.np, self.D))
        check(load().gb200_quadrature_points(self.h, _ptr(xq)), self.ctx.h)
        return xq

    def download_into(self, nzval=None, b=None):
        """gb200_plan_download into caller arrays (either may be None)"""
        check(load().gb200_plan_download(self.h, _ptr(nzval), _ptr(b)), self.ctx.h)

    def download(self, nzval=True, b=True):
        nz = np.zeros(self.nnz) if nzval else None
        bb = np.zeros(self.nrows) if b else None
        check(load().gb200_plan_download(self.h, _ptr(nz), _ptr(bb)), self.ctx.h)
        return nz, bb

    def device_nzval(self):
        p, n = C.c_void_p(), C.c_int64()
        check(load().gb200_plan_device_nzval(self.h, C.byref(p), C.byref(n)), self.ctx.h)
        return p.value, n.value

    def device_pattern(self):
        """(colptr device pointer [Int64, ncols+1], rowval device pointer [Int32, nnz]), 0-based"""
        cp, rv = C.c_void_p(), C.c_void_p()
        check(load().gb200_plan_device_pattern(self.h, C.byref(cp), C.byref(rv)), self.ctx.h)
        return cp.value, rv.value

    def device_arrays(self):
        """the device-resident CSC (colptr, rowval, nzval) and vector as objects with `__cuda_array_interface__` (zero-copy:
        `torch.as_tensor(x, device="cuda")`, CuPy, Numba ... wrap them without a copy; they stay valid while the plan lives).
        The consumer works on its own stream: synchronise it before the next assembly call overwrites the arrays, and call
        `ctx.synchronize()` after a device-resident assembly before reading them."""
        cp, rv = self.device_pattern()
        nz, _ = self.device_nzval()
        bv, _ = self.device_vector()
        return (DeviceArray(cp, self.ncols + 1, "<i8", self), DeviceArray(rv, self.nnz, "<i4", self),
                DeviceArray(nz, self.nnz, "<f8", self), DeviceArray(bv, self.nrows, "<f8", self))

    def device_vector(self):
        p, n = C.c_void_p(), C.c_int64()
        check(load().gb200_plan_device_vector(self.h, C.byref(p), C.byref(n)), self.ctx.h)
        return p.value, n.value

    def kernel_path(self, form):
        return load().gb200_plan_kernel_path(self.h, form).decode()

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                load().gb200_plan_destroy(self.h)
        except Exception:
            pass
