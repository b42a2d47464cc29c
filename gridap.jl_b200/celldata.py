"""Weak forms: a small expression language mirroring Gridap's `∫( ... )dΩ` syntax, and the integrand recogniser.

In the reference the form closure is evaluated into a lazy tree of maps (`integrate`, src/CellData/CellQuadratures.jl:138-170;
`IntegrationMap`, src/Fields/FieldsInterfaces.jl:627-776) which the CPU assembler then evaluates cell by cell.  Here the
tree is *recognised* and mapped to one of the hand-written kernels; anything else raises
`NotImplementedError` -- there is no CPU fallback.

    a(u, v) = Integral(inner(grad(v), grad(u))) * dΩ         # ∫( ∇(v)⊙∇(u) )dΩ
    l(v)    = Integral(v * f) * dΩ                            # ∫( v*f )dΩ
"""
import numpy as np

from . import lib
from . import reffes as rf
from .geometry import NormalVector, Triangulation


# ------------------------------------------------------------------------------------------------ expressions
class Expr:
    def __add__(self, o):
        return Sum([self, _wrap(o)])

    __radd__ = __add__

    def __sub__(self, o):
        return Sum([self, Scaled(-1.0, _wrap(o))])

    def __rsub__(self, o):
        return Sum([_wrap(o), Scaled(-1.0, self)])

    def __neg__(self):
        return Scaled(-1.0, self)

    def __mul__(self, o):
        if isinstance(o, (int, float)):
            return Scaled(float(o), self)
        return Mul(self, _wrap(o))

    def __rmul__(self, o):
        if isinstance(o, (int, float)):
            return Scaled(float(o), self)
        return Mul(_wrap(o), self)


def _wrap(o):
    if isinstance(o, Expr):
        return o
    if isinstance(o, NormalVector):
        return Normal(o.trian)
    if hasattr(o, "free_values") and hasattr(o, "dirichlet_values"):   # an FEFunction inside an integrand
        return State(o)
    if isinstance(o, (int, float)):
        return Const(float(o))
    if callable(o):
        return Coef(o)
    if isinstance(o, (tuple, list, np.ndarray)):
        return Const(np.asarray(o, dtype=np.float64))
    raise TypeError("cannot use %r in a weak form" % (o,))


class Const(Expr):
    def __init__(self, value):
        self.value = value


class Coef(Expr):
    """A Julia closure `f(x)` cannot run on the GPU: it is evaluated on the host at the physical quadrature points the
    library exports (gb200_quadrature_points) and shipped as values (hard parts)."""

    def __init__(self, fn):
        self.fn = fn


class Normal(Expr):
    """the unit normal n_Gamma of a BoundaryTriangulation inside an integrand"""

    def __init__(self, trian):
        self.trian = trian


class Basis(Expr):
    def __init__(self, kind, space, field=None):
        self.kind, self.space, self.field = kind, space, field  # kind: 'test' | 'trial'


class State(Expr):
    """An FE function u_h inside an integrand (residual / Jacobian)."""

    def __init__(self, uh):
        self.uh = uh


class Grad(Expr):
    def __init__(self, a):
        self.a = a


class Div(Expr):
    def __init__(self, a):
        self.a = a


class SymGrad(Expr):
    def __init__(self, a):
        self.a = a


class Inner(Expr):
    def __init__(self, a, b):
        self.a, self.b = a, b


class Dot(Inner):
    pass


class Mul(Inner):
    pass


class Scaled(Expr):
    def __init__(self, c, a):
        self.c, self.a = c, a


class Sum(Expr):
    def __init__(self, terms):
        self.terms = terms


class Jump(Expr):
    """jump(a) = a.plus - a.minus on a SkeletonTriangulation (src/CellData/CellFields.jl `jump`); with the normal inside,
    jump(v*n) = v+ n+ + v- n- = (v+ - v-) n+"""

    def __init__(self, a):
        self.a = a


class Mean(Expr):
    """mean(a) = 0.5 (a.plus + a.minus)"""

    def __init__(self, a):
        self.a = a


class LawTerm(Expr):
    """Application of a constitutive law, e.g. sigma∘eps(u)."""

    def __init__(self, law, name, args):
        self.law, self.name, self.args = law, name, args


def grad(a):
    return Grad(_state(a))


nabla = grad


def div(a):
    return Div(_state(a))


def eps(a):
    return SymGrad(_state(a))


ε = eps


def inner(a, b):
    return Inner(_wrap(_state(a)), _wrap(_state(b)))


def dot(a, b):
    return Dot(_wrap(_state(a)), _wrap(_state(b)))


def jump(a):
    return Jump(_wrap(a))


def mean(a):
    return Mean(_wrap(a))


def _state(a):
    from .fespaces import FEFunction
    if isinstance(a, NormalVector):
        return Normal(a.trian)
    return State(a) if isinstance(a, FEFunction) else a


class IsotropicLinearElasticity:
    """sigma(eps) = lambda tr(eps) I + 2 mu eps (test/GridapTests/IsotropicDamageTests.jl:14-18)."""

    def __init__(self, lam, mu):
        self.lam, self.mu = float(lam), float(mu)

    @classmethod
    def from_E_nu(cls, E, nu):
        return cls(E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu)))

    def __call__(self, e):
        return LawTerm(self, "sigma", (e,))


class NeoHookean:
    """Compressible neo-Hookean law of the Gridap hyperelasticity tutorial (not in the
    reference repository, parity unpinned):  S = mu (I - C^-1) + lambda ln(J) C^-1."""

    def __init__(self, lam, mu):
        self.lam, self.mu = float(lam), float(mu)

    def S(self, gu):
        return LawTerm(self, "S", (gu,))

    def dE(self, gdu, gu):
        return LawTerm(self, "dE", (gdu, gu))

    def dS(self, gdu, gu):
        return LawTerm(self, "dS", (gdu, gu))

    def res(self, u, v):
        return inner(self.dE(grad(v), grad(u)), self.S(grad(u)))

    def jac(self, u, du, v):
        return inner(self.dE(grad(v), grad(u)), self.dS(grad(du), grad(u))) + inner(grad(v), dot(self.S(grad(u)), grad(du)))


# ------------------------------------------------------------------------------------------------ measures
class Measure:
    """Measure(Ω, degree): CellQuadrature of `degree` on every cell (src/CellData/CellQuadratures.jl:195-202)."""

    def __init__(self, trian, degree):
        if not isinstance(trian, Triangulation):
            trian = Triangulation(trian)
        self.trian, self.degree = trian, int(degree)
        self.points, self.weights = rf.Quadrature(trian.model.ptype, self.degree)


class Integrand:
    def __init__(self, expr):
        self.expr = _wrap(expr)

    def __mul__(self, measure):
        if not isinstance(measure, Measure):
            raise TypeError("Integral(...) must be multiplied by a Measure")
        return DomainContribution([(self.expr, measure)])


def Integral(expr):
    """`∫(expr)`; multiply by a Measure: Integral(expr)*dΩ."""
    return Integrand(expr)


class DomainContribution:
    """Sum of integrals over triangulations (src/CellData/DomainContributions.jl:7-10)."""

    def __init__(self, terms):
        self.terms = terms

    def __add__(self, o):
        return DomainContribution(self.terms + o.terms)

    def __sub__(self, o):
        return DomainContribution(self.terms + [(Scaled(-1.0, e), m) for e, m in o.terms])


# ------------------------------------------------------------------------------------------------ recogniser
class Term:
    def __init__(self, form, params=(), fq=None, state=None, fields=None, glued=False):
        self.form, self.params, self.fq, self.state, self.fields = form, tuple(params), fq, state, fields
        self.glued = glued   # a term on boundary facets that needs the adjacent cell (normal, cell-basis gradients)


def _flatten(e, c=1.0):
    """-> list of (coef, expr) with Sum / Scaled removed."""
    if isinstance(e, Sum):
        out = []
        for t in e.terms:
            out += _flatten(t, c)
        return out
    if isinstance(e, Scaled):
        return _flatten(e.a, c * e.c)
    if type(e) is Mul:
        for x, y in ((e.a, e.b), (e.b, e.a)):
            if isinstance(x, Const) and np.ndim(x.value) == 0 and _count_basis(y) == 2:
                return _flatten(y, c * float(x.value))
    if isinstance(e, Inner):   # products are bilinear: scalar factors of the operands move out, (c v) * u = c (v * u)
        a, b, moved = e.a, e.b, False
        while isinstance(a, Scaled):
            c, a, moved = c * a.c, a.a, True
        while isinstance(b, Scaled):
            c, b, moved = c * b.c, b.a, True
        if moved:
            return _flatten(type(e)(a, b), c)
    return [(c, e)]


def _has_basis(e):
    return _count_basis(e) > 0


def _count_basis(e):
    if isinstance(e, Basis):
        return 1
    n = 0
    for k in ("a", "b"):
        if hasattr(e, k):
            n += _count_basis(getattr(e, k))
    if isinstance(e, LawTerm):
        n += sum(_count_basis(x) for x in e.args)
    if isinstance(e, Sum):
        n += sum(_count_basis(x) for x in e.terms)
    return n


def _basis(e, kind):
    return e if isinstance(e, Basis) and e.kind == kind else None


def _pair(e, fa, fb):
    """match a binary node whose operands satisfy (fa, fb) in either order; returns (xa, xb) or None"""
    if not isinstance(e, Inner):
        return None
    for x, y in ((e.a, e.b), (e.b, e.a)):
        ra, rb = fa(x), fb(y)
        if ra is not None and rb is not None:
            return ra, rb
    return None


def _grad_of(kind):
    return lambda x: _basis(x.a, kind) if isinstance(x, Grad) else None


def _div_of(kind):
    return lambda x: _basis(x.a, kind) if isinstance(x, Div) else None


def _sym_of(kind):
    return lambda x: _basis(x.a, kind) if isinstance(x, SymGrad) else None


def _nderiv(x, inner_match):
    """n . grad(y) (either operand order) -> inner_match(y), else None"""
    if isinstance(x, Inner):
        for p, q in ((x.a, x.b), (x.b, x.a)):
            if isinstance(p, Normal) and isinstance(q, Grad):
                return inner_match(q.a)
    return None


def _facet_factor(x, kind):
    """test / trial factor of a facet term: (basis, 0) for the value, (basis, 1) for the normal derivative n.grad"""
    b = _basis(x, kind)
    if b is not None:
        return (b, 0)
    b = _nderiv(x, lambda y: _basis(y, kind))
    return None if b is None else (b, 1)


def _skeleton_factor(x, kind):
    """test / trial factor of a term on a SkeletonTriangulation -> (basis, op kind, w+, w-, is_vector) or None
    (src/CellData/CellFields.jl:643-652: jump(a) = a+ - a-; on a SkeletonPair, i.e. with the normal inside, jump(a n) = a+ n+ + a- n-
    = (a+ - a-) n+; mean(a) = (a+ + a-)/2).  op kind 0: value, 1: derivative along n+.
    jump(v) -> (0, 1, -1, scalar);  jump(v*n) -> (0, 1, -1, vector along n+);  mean(v) -> (0, .5, .5, scalar);
    mean(grad v) / jump(grad v) -> (1, ., ., vector: dotted with a vector along n+);  jump(n.grad v) -> (1, 1, -1, scalar)"""
    if not isinstance(x, (Jump, Mean)):
        return None
    is_jump = isinstance(x, Jump)
    w = (1.0, -1.0) if is_jump else (0.5, 0.5)
    a = x.a
    b = _basis(a, kind)
    if b is not None:
        return (b, 0, w[0], w[1], False)
    if isinstance(a, Grad) and _basis(a.a, kind) is not None:
        return (a.a, 1, w[0], w[1], True)
    if not is_jump:
        return None   # (mean of a SkeletonPair is not defined in the reference either)
    if type(a) in (Mul, Inner, Dot):   # v * n
        for y, z in ((a.a, a.b), (a.b, a.a)):
            if _basis(y, kind) is not None and isinstance(z, Normal):
                return (y, 0, 1.0, -1.0, True)
    b = _nderiv(a, lambda y: _basis(y, kind))   # n . grad(v)
    if b is not None:
        return (b, 1, 1.0, -1.0, False)
    return None


def _facet_data(x):
    """data factor of a facet vector term: ("g", Const | Coef), ("u", FEFunction) for u_h, ("dn", FEFunction) for n.grad(u_h)"""
    if isinstance(x, (Const, Coef)):
        return ("g", x)
    if isinstance(x, State):
        return ("u", x.uh)
    uh = _nderiv(x, lambda y: y.uh if isinstance(y, State) else None)
    return None if uh is None else ("dn", uh)


def _unsupported(what):
    return NotImplementedError("%s is not in the supported integrand set {mass, laplacian, linear elasticity, Stokes blocks, "
                               "neo-Hookean residual/Jacobian, source}; the B200 assembler never falls back to the CPU" % what)


def _nh_state(x):
    """Grad(State) -> FEFunction"""
    if isinstance(x, Grad) and isinstance(x.a, State):
        return x.a.uh
    return None


def recognise_matrix(expr):
    """Bilinear integrand -> list of Term (single-field forms may be summed; Stokes is recognised as a whole)."""
    terms = _flatten(expr)
    out = []
    stokes = {}
    for c, e in terms:
        m = _pair(e, _grad_of("test"), _grad_of("trial"))
        if m:
            v, u = m
            if v.field is not None and u.field is not None:
                if (v.field, u.field) == (0, 0) and c == 1.0:
                    stokes["vu"] = True
                    continue
                raise _unsupported("a multi-field gradient term on fields %r" % ((v.field, u.field),))
            out.append(Term(lib.FORM_LAPLACIAN, (c,)))
            continue
        m = _pair(e, lambda x: _basis(x, "test"), lambda x: _basis(x, "trial"))
        if m:
            v, u = m
            if v.field is not None:
                raise _unsupported("a multi-field mass term")
            out.append(Term(lib.FORM_MASS, (c,)))
            continue
        m = _pair(e, lambda x: _facet_factor(x, "test"), lambda x: _facet_factor(x, "trial"))
        if m:   # v (n.grad u), (n.grad v) u, (n.grad v)(n.grad u): Nitsche-type terms on a BoundaryTriangulation
            (v, tk), (u, uk) = m
            if v.field is not None or u.field is not None:
                raise _unsupported("a multi-field boundary term with normal derivatives")
            out.append(Term(lib.FORM_FACET, (c, tk, uk), glued=True))
            continue
        m = _pair(e, lambda x: _skeleton_factor(x, "test"), lambda x: _skeleton_factor(x, "trial"))
        if m:   # jump / mean terms on a SkeletonTriangulation (test/GridapTests/PoissonDGTests.jl:42-45)
            (v, tk, tp, tm, tvec), (u, uk, up, um, uvec) = m
            if v.field is not None or u.field is not None:
                raise _unsupported("a multi-field skeleton term")
            if tvec != uvec:
                raise _unsupported("a skeleton term pairing a scalar with a vector along the normal")
            out.append(Term(lib.FORM_SKELETON, (c, tk, tp, tm, uk, up, um), glued="skeleton"))
            continue
        m = _pair(e, _div_of("test"), _div_of("trial"))
        if m:   # (div v)(div u), the `graddiv` form of the reference's assembly benchmark (benchmark/bm/bm_assembly.jl:9): the
            #     lambda-part of isotropic linear elasticity, sigma = lambda tr(eps) I with mu = 0
            if m[0].field is not None or m[1].field is not None:
                raise _unsupported("a multi-field div-div term")
            out.append(Term(lib.FORM_ELASTICITY, (c, 0.0)))
            continue
        m = _pair(e, _div_of("test"), lambda x: _basis(x, "trial"))
        if m and m[0].field == 0 and m[1].field == 1 and c == -1.0:
            stokes["vp"] = True
            continue
        m = _pair(e, lambda x: _basis(x, "test"), _div_of("trial"))
        if m and m[0].field == 1 and m[1].field == 0 and c == 1.0:
            stokes["qu"] = True
            continue
        if isinstance(e, Inner):
            for x, y in ((e.a, e.b), (e.b, e.a)):
                if isinstance(x, SymGrad) and _basis(x.a, "test") and isinstance(y, LawTerm) and isinstance(y.law, IsotropicLinearElasticity) \
                        and isinstance(y.args[0], SymGrad) and _basis(y.args[0].a, "trial"):
                    out.append(Term(lib.FORM_ELASTICITY, (c * y.law.lam, c * y.law.mu)))
                    break
            else:
                nh = _match_nh_jac(e)
                if nh is None or c != 1.0:
                    raise _unsupported("this bilinear term (%s)" % type(e).__name__)
                out.append(nh)
            continue
        raise _unsupported("this bilinear term (%s)" % type(e).__name__)
    if stokes:
        if set(stokes) != {"vu", "vp", "qu"} or out:
            raise _unsupported("this combination of multi-field terms (Stokes needs exactly grad(v):grad(u) - div(v)*p + q*div(u))")
        return [Term(lib.FORM_STOKES, ())]
    # the two neo-Hookean Jacobian terms come as a pair
    nh = [t for t in out if isinstance(t, tuple)]
    if nh:
        kinds = sorted(k for k, _, _ in nh)
        if kinds != ["geo", "mat"] or len(out) != 2 or nh[0][1] is not nh[1][1]:
            raise _unsupported("this neo-Hookean Jacobian (needs dE(∇v,∇u)⊙dS(∇du,∇u) + ∇v⊙(S(∇u)⋅∇du))")
        law, uh = nh[0][1], nh[0][2]
        return [Term(lib.FORM_NEOHOOKEAN_JAC, (law.lam, law.mu), state=uh)]
    return out


def _match_nh_jac(e):
    a, b = e.a, e.b
    for x, y in ((a, b), (b, a)):
        # material part: dE(∇v,∇u) ⊙ dS(∇du,∇u)
        if isinstance(x, LawTerm) and x.name == "dE" and isinstance(y, LawTerm) and y.name == "dS" and x.law is y.law \
                and _grad_of("test")(x.args[0]) and _grad_of("trial")(y.args[0]):
            uh = _nh_state(x.args[1])
            if uh is not None and _nh_state(y.args[1]) is uh:
                return ("mat", x.law, uh)
        # geometric part: ∇v ⊙ (S(∇u)⋅∇du)
        if _grad_of("test")(x) and isinstance(y, Dot) and isinstance(y.a, LawTerm) and y.a.name == "S" and _grad_of("trial")(y.b):
            uh = _nh_state(y.a.args[0])
            if uh is not None:
                return ("geo", y.a.law, uh)
    return None


def _merge_field_sources(srcs, ncomps):
    """[(field, coef, Const | Coef)] -> {field: (params, fq)}: one source per field (several terms on a field are summed)"""
    out = {}
    for k in sorted({k for k, _, _ in srcs}):
        mine = [(c, f) for kk, c, f in srcs if kk == k]
        if all(isinstance(f, Const) for _, f in mine):
            tot = sum(c * np.broadcast_to(np.atleast_1d(np.asarray(f.value, dtype=np.float64)), (ncomps[k],)) for c, f in mine)
            out[k] = (tuple(float(x) for x in tot), None)
        else:
            def fq(x, mine=mine, nck=ncomps[k]):
                tot = 0.0
                for c, f in mine:
                    v = np.asarray(f.fn(x), dtype=np.float64) if isinstance(f, Coef) else np.broadcast_to(np.atleast_1d(np.asarray(f.value, dtype=np.float64)), (len(x), nck))
                    tot = tot + c * v.reshape(len(x), nck)
                return tot
            out[k] = ((1.0,), fq)
    return out


def recognise_vector(expr):
    """Linear integrand -> list of Term.  Multi-field forms l((v,q)) = int(v.f + q*g) (test/GridapTests/StokesTaylorHoodTests.jl:61)
    become ONE source term carrying a source per field (the block vector of src/Arrays/AlgebraMaps.jl:154-271)."""
    out = []
    mf = []      # (field, coef, Const | Coef) of multi-field source terms
    ncomps = {}
    for c, e in _flatten(expr):
        if _basis(e, "test") is not None:  # v*c with a plain number c
            if e.field is not None:
                mf.append((e.field, c, Const(1.0)))
                ncomps[e.field] = e.space.ncomp
                continue
            out.append(Term(lib.FORM_SOURCE, (c,) * e.space.ncomp))
            continue
        if isinstance(e, Inner):
            fm = _pair(e, lambda x: _facet_factor(x, "test"), _facet_data)
            if fm and (fm[0][1] == 1 or fm[1][0] != "g"):
                # (n.grad v) g, v u_h, (n.grad v) u_h, v (n.grad u_h): boundary terms that need the adjacent cell
                (v, tk), (dkind, dat) = fm
                if v.field is not None:
                    raise _unsupported("a multi-field boundary term with normal derivatives")
                if dkind == "g":
                    if isinstance(dat, Const):
                        gv = np.broadcast_to(np.atleast_1d(np.asarray(dat.value, dtype=np.float64)), (v.space.ncomp,)).copy()
                        out.append(Term(lib.FORM_FACET_VEC, (c, tk, 0), fq=(lambda x, gv=gv: np.broadcast_to(gv, (len(x), len(gv)))), glued=True))
                    else:
                        out.append(Term(lib.FORM_FACET_VEC, (c, tk, 0), fq=dat.fn, glued=True))
                else:
                    out.append(Term(lib.FORM_FACET_VEC, (c, tk, 1 if dkind == "u" else 2), state=dat, glued=True))
                continue
            m = None
            for x, y in ((e.a, e.b), (e.b, e.a)):
                if _basis(x, "test") and isinstance(y, (Const, Coef)):
                    m = (x, y)
            if m:
                v, f = m
                if v.field is not None:
                    mf.append((v.field, c, f))
                    ncomps[v.field] = v.space.ncomp
                    continue
                if isinstance(f, Const):
                    out.append(Term(lib.FORM_SOURCE, tuple(c * np.atleast_1d(f.value))))
                else:
                    out.append(Term(lib.FORM_SOURCE, (c,), fq=f.fn))
                continue
            for x, y in ((e.a, e.b), (e.b, e.a)):
                if isinstance(x, LawTerm) and x.name == "dE" and _grad_of("test")(x.args[0]) and isinstance(y, LawTerm) and y.name == "S" \
                        and x.law is y.law and _nh_state(x.args[1]) is not None and _nh_state(y.args[0]) is _nh_state(x.args[1]) and c == 1.0:
                    out.append(Term(lib.FORM_NEOHOOKEAN_RES, (x.law.lam, x.law.mu), state=_nh_state(y.args[0])))
                    break
            else:
                raise _unsupported("this linear term (%s)" % type(e).__name__)
            continue
        raise _unsupported("this linear term (%s)" % type(e).__name__)
    if mf:
        if out:
            raise _unsupported("a mix of single-field and multi-field linear terms")
        out.append(Term(lib.FORM_SOURCE, (), fields=_merge_field_sources(mf, ncomps)))
    return out
