"""Host-side mirror of the reference's user API for the assembly path.

Same names, argument meaning and error behaviour as GalerkinToolkit
(problems.jl:4-34, 244-274, 319-404; assembly.jl:11-25) so tests read like the
reference's own (test/problems_tests.jl, test/assembly_tests.jl):

    mesh = GT.cartesian_mesh((0,1,0,1),(64,64))
    Ω = GT.interior(mesh);  Γ = GT.boundary(mesh)
    V = GT.lagrange_space(Ω, 1, dirichlet_boundary=Γ)
    dΩ = GT.measure(Ω, 2)
    a = lambda u, v: GT.integrate(lambda x: GT.dot(GT.grad(u, x), GT.grad(v, x)), dΩ)     # ∫(x->∇(u,x)⋅∇(v,x),dΩ)
    l = lambda v:    GT.integrate(lambda x: f(x) * v(x), dΩ)
    A = GT.assemble_matrix(a, float, V, V)
    b = GT.assemble_vector(l, float, V)

The integrand lambdas run on *symbolic quantities* and build a small term tree
(the role of compiler.jl:372-390, 487-531, 1544-1606).  `recognise_*` is the
stand-in for compiler.jl/passes.jl's form recognition: the
supported shapes (mass, Laplacian, isotropic elasticity, source terms) are
dispatched to the fused CUDA kernels; anything else raises
:class:`UnsupportedFormError` — there is no CPU fallback.

The cell loop, scatter and compression all run in libgtkasm (CUDA); numpy is
only used here to prepare the flat input arrays (hostprep.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

from . import engine as _eng
from . import hostprep as _hp
from .engine import DIRICHLET, FREE, UnsupportedFormError  # noqa: F401

Float64 = float

# ---------------------------------------------------------------------------
# mesh / domains / spaces / measures
# ---------------------------------------------------------------------------
cartesian_mesh = _hp.cartesian_mesh


@dataclass
class Domain:
    """interior(mesh) / boundary(mesh; group_names) (mesh.jl:223-271)."""
    mesh: _hp.Mesh
    kind: str                      # "interior" | "boundary"
    sides: Optional[Sequence[int]] = None


def interior(mesh) -> Domain:
    return Domain(mesh, "interior")


def boundary(mesh, group_names: Optional[Sequence[str]] = None) -> Domain:
    """group names are the reference's "<D-1>-face-<i>" (cartesian_mesh.jl:169-178)."""
    sides = None
    if group_names is not None:
        sides = []
        for g in group_names:
            d, _, i = g.split("-")
            if int(d) != mesh.D - 1:
                raise ValueError(f"only (D-1)-face groups can be boundaries, got {g}")
            sides.append(int(i))
    return Domain(mesh, "boundary", sides)


def skeleton(mesh) -> Domain:
    """GT.skeleton(mesh): the interior (D-1)-faces, two cells around each (domain.jl; accessors.jl:394-473)."""
    return Domain(mesh, "skeleton")


@dataclass
class Measure:
    domain: Domain
    degree: int


def measure(domain: Domain, degree: int) -> Measure:
    """GT.measure(Ω | Γ, degree) (problems.jl:12-25).  Boundary measures serve linear forms ∫_Γ g v dΓ (Neumann terms)."""
    return Measure(domain, degree)


class Space:
    """lagrange_space(Ω, order; dirichlet_boundary, tensor_size) (space.jl:1702-1735)."""

    def __init__(self, domain: Domain, order: int, dirichlet_boundary: Optional[Domain] = None, tensor_size=None, continuous: bool = True):
        if domain.kind != "interior":
            raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "spaces on boundary domains are out of scope")
        n_comp = 1 if tensor_size is None else int(np.prod(tensor_size))
        bc = None
        if dirichlet_boundary is not None:
            bc = "boundary" if dirichlet_boundary.sides is None else list(dirichlet_boundary.sides)
        self.domain = domain
        self.continuous = bool(continuous)
        if not continuous:
            if bc is not None:
                raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "discontinuous spaces take their boundary conditions weakly (no dirichlet_boundary)")
            self.data = _hp.discontinuous_lagrange_space(domain.mesh, order, n_comp)
        else:
            self.data = _hp.lagrange_space(domain.mesh, order, bc, n_comp)
        self._tab = {}

    # -- reference accessors ---------------------------------------------------
    def face_dofs(self):
        return self.data.cell_dofs

    def num_free_dofs(self):
        return self.data.n_free

    def num_dirichlet_dofs(self):
        return self.data.n_dirichlet

    def tabulation(self, degree: int):
        if degree not in self._tab:
            self._tab[degree] = _hp.measure_tabulation(self.data, degree)
        return self._tab[degree]

    def face_problem(self, meas: "Measure"):
        """faces of the boundary domain of `meas` + this space's dofs on them + tabulation on the reference face"""
        key = ("Γ", None if meas.domain.sides is None else tuple(meas.domain.sides), meas.degree)
        if key not in self._tab:
            self._tab[key] = _hp.face_problem(self.data, meas.domain.sides, meas.degree)
        return self._tab[key]


    def __mul__(self, other):
        """V × Q (cartesian_product, space.jl): Python spells × as *"""
        return cartesian_product(self, other)


def lagrange_space(domain, order, dirichlet_boundary=None, tensor_size=None, continuous: bool = True) -> Space:
    return Space(domain, order, dirichlet_boundary, tensor_size, continuous)


class ProductSpace:
    """CartesianProductSpace: fields(V × Q) = (V, Q); free / Dirichlet dofs of the fields are numbered block after block
    (assembly.jl:321-333)."""

    def __init__(self, *spaces):
        flat = []
        for sp in spaces:
            flat += list(sp.fields) if isinstance(sp, ProductSpace) else [sp]
        if any(sp.domain.mesh is not flat[0].domain.mesh for sp in flat):
            raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "fields of a product space must share one mesh")
        self.fields = tuple(flat)
        self.domain = flat[0].domain

    def __mul__(self, other):
        return ProductSpace(self, other)

    def num_free_dofs(self):
        return sum(f.num_free_dofs() for f in self.fields)

    def num_dirichlet_dofs(self):
        return sum(f.num_dirichlet_dofs() for f in self.fields)


def cartesian_product(*spaces) -> ProductSpace:
    return ProductSpace(*spaces)


def fields(space):
    return space.fields if isinstance(space, ProductSpace) else (space,)


# ---------------------------------------------------------------------------
# symbolic quantities (tiny stand-in for compiler.jl's term IR)
# ---------------------------------------------------------------------------
class Term:
    def __mul__(self, o): return Call("*", self, _lift(o))
    def __rmul__(self, o): return Call("*", _lift(o), self)
    def __add__(self, o): return Call("+", self, _lift(o))
    def __radd__(self, o): return Call("+", _lift(o), self)
    def __sub__(self, o): return Call("-", self, _lift(o))
    def __rsub__(self, o): return Call("-", _lift(o), self)
    def __matmul__(self, o): return Call("dot", self, _lift(o))
    def __truediv__(self, o): return Call("/", self, _lift(o))
    def __rtruediv__(self, o): return Call("/", _lift(o), self)


@dataclass
class Const(Term):
    value: object


@dataclass
class Coordinate(Term):        # CoordinateTerm (compiler.jl:60-1028)
    pass


@dataclass
class FormArg(Term):           # FormArgumentTerm: arg 1 = test, arg 2 = trial (problems.jl:324-327)
    arg: int
    op: str                    # "value" | "gradient" | "divergence"
    field: int = 0             # field of a product space (0-based)
    side: int = 0              # skeleton integrals: u[1] / u[2] = the cell around (1-based; 0 = no restriction)


@dataclass
class Call(Term):              # CallTerm
    fn: object
    a: Term
    b: Optional[Term] = None


@dataclass
class Named(Term):             # named integrand whose identity is matched
    name: str
    params: dict
    args: tuple


@dataclass
class Normal(Term):            # unit_normal(mesh, D-1)[side](x) on a skeleton face (accessors.jl:1009-1035)
    side: int


@dataclass
class FaceDiameter(Term):      # face_diameter_field(Λ)(x) (field.jl:488-492)
    inverse: bool = False


def _lift(o):
    return o if isinstance(o, Term) else Const(o)


class _UnitNormal:
    """n = GT.unit_normal(mesh, D-1): n[1](x), n[2](x) = the outward unit normals of the two cells around a skeleton face"""

    def __getitem__(self, k: int):
        if k not in (1, 2):
            raise IndexError("a skeleton face has two cells around: n[1], n[2]")
        return lambda x: Normal(k)

    def __call__(self, x):
        """on a boundary face: the outward unit normal of the one cell around"""
        return Normal(0)


def unit_normal(mesh, d: int) -> _UnitNormal:
    if d != mesh.D - 1:
        raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "unit normals are defined on (D-1)-faces")
    return _UnitNormal()


def face_diameter_field(domain: Domain):
    """h_Λ = GT.face_diameter_field(Λ): h_Λ(x) is the diameter of the face x lies on (field.jl:488-492)"""
    if domain.kind not in ("skeleton", "boundary"):
        raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "face diameters are supported on skeleton and boundary domains on the GPU path")
    return lambda x: FaceDiameter()


def uniform_quantity(v):
    return v


class FormArgument:
    """form_argument_quantity(space, arg) (compiler.jl:487-531)."""

    def __init__(self, space: Space, arg: int, field: int = 0, side: int = 0):
        self.space, self.arg, self.field, self.side = space, arg, field, side

    def __call__(self, x):
        return FormArg(self.arg, "value", self.field, self.side)

    def __getitem__(self, k: int):
        """u[1], u[2]: the restriction to the first / second cell around a skeleton face (compiler.jl:507-531, SkeletonTerm)"""
        if k not in (1, 2):
            raise IndexError("a skeleton face has two cells around: u[1], u[2]")
        return FormArgument(self.space, self.arg, self.field, k)


def grad(u, x):
    """∇(u,x) = ForwardDiff.gradient(u,x) on a form argument → tabulated gradient (compiler.jl:573-589); on a
    DiscreteField → Σ_i u_i ∇φ_i (DiscreteFieldTerm, accessors.jl:1549-1556); on an AnalyticalField with a known
    gradient → that gradient sampled by the host."""
    if isinstance(u, FormArgument):
        return FormArg(u.arg, "gradient", u.field, u.side)
    if isinstance(u, DiscreteField):
        return FieldTerm(u, "gradient")
    if isinstance(u, AnalyticalField) and u.gradient is not None:
        return Call(u.gradient, x)
    raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "∇ of a non form-argument quantity is not supported on the GPU path")


def dot(a, b):
    return Call("dot", _lift(a), _lift(b))


def div(u, x):
    """div(u,x) = tr(ForwardDiff.jacobian(u,x)) of a vector-valued form argument (docs/src/src_jl/example_stokes.jl)"""
    if isinstance(u, FormArgument):
        return FormArg(u.arg, "divergence", u.field, u.side)
    raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "div of a non form-argument quantity is not supported on the GPU path")


@dataclass
class FieldTerm(Term):         # DiscreteFieldTerm (compiler.jl:60-1028): value or gradient of a DiscreteField at x
    field: "DiscreteField"
    op: str                    # "value" | "gradient"


class DiscreteField:
    """GT.DiscreteField (field.jl:93-125): a space plus free and Dirichlet values.  Passed as `parameters=(uh,)` it is
    uploaded to the engine's field slot before every re-assembly (gtk_field_set_values)."""

    def __init__(self, space: "Space", free_values, dirichlet_values):
        self.space = space
        self.free_values = np.ascontiguousarray(free_values, dtype=np.float64)
        self.dirichlet_values = np.ascontiguousarray(dirichlet_values, dtype=np.float64)

    def __call__(self, x):
        return FieldTerm(self, "value")


def discrete_field(space, free_values, dirichlet_values) -> DiscreteField:
    return DiscreteField(space, free_values, dirichlet_values)


def free_values(uh: DiscreteField):
    return uh.free_values


def dirichlet_values(uh: DiscreteField):
    return uh.dirichlet_values


def zero_field(T, space: "Space") -> DiscreteField:
    """GT.zero_field(T, V) (field.jl:202-204)"""
    return DiscreteField(space, np.zeros(space.num_free_dofs()), np.zeros(space.num_dirichlet_dofs()))


def rand_field(T, space: "Space", rng=None) -> DiscreteField:
    """GT.rand_field(T, V) (field.jl:196-200): random free values in [0,1), ZERO Dirichlet values"""
    rng = np.random.default_rng() if rng is None else rng
    return DiscreteField(space, rng.random(space.num_free_dofs()), np.zeros(space.num_dirichlet_dofs()))


class AnalyticalField:
    """GT.analytical_field(f, Ω) (field.jl:17-58): evaluated by the host at x_q, enters the engine as data.
    `gradient` (optional) plays ForwardDiff.gradient(f, x) for error norms ∇(u,q) − ∇(uh,q)."""

    def __init__(self, f: Callable, domain: Optional[Domain] = None, gradient: Optional[Callable] = None):
        self.f = f
        self.gradient = gradient

    def __call__(self, x):
        if isinstance(x, Coordinate):
            return Call(self.f, x)
        return self.f(x)


analytical_field = AnalyticalField


def call(fn, *args):
    """GT.call(f, args...) (compiler.jl:372-390): apply an external function to quantities.  Only the named functions
    the engine has a fused kernel for can be recognised afterwards (plaplacian_flux / plaplacian_dflux)."""
    if len(args) == 1:
        return Call(fn, _lift(args[0]))
    if len(args) == 2:
        return Call(fn, _lift(args[0]), _lift(args[1]))
    raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "GT.call with more than two arguments is not recognised by the GPU engine")


def abs2(t):
    return Call("abs2", _lift(t))


class plaplacian_flux:
    """flux(∇u) = norm(∇u)^(q-2) * ∇u (test/problems_ext_tests.jl:160, test/assembly_tests.jl:678) as a NAMED callable:
    the reference passes an opaque closure to GT.call, which no recogniser can look into."""

    def __init__(self, q):
        self.q = q

    def __call__(self, gu):
        gu = np.asarray(gu, dtype=np.float64)
        return np.linalg.norm(gu, axis=0) ** (self.q - 2) * gu


class plaplacian_dflux:
    """dflux(∇du,∇u) = (q-2)*norm(∇u)^(q-4)*(∇u⋅∇du)*∇u + norm(∇u)^(q-2)*∇du (test/problems_ext_tests.jl:161)"""

    def __init__(self, q):
        self.q = q

    def __call__(self, gdu, gu):
        gu = np.asarray(gu, dtype=np.float64); gdu = np.asarray(gdu, dtype=np.float64)
        n = np.linalg.norm(gu, axis=0)
        return (self.q - 2) * n ** (self.q - 4) * (gu * gdu).sum(0) * gu + n ** (self.q - 2) * gdu


def isotropic_elasticity(lam: float, mu: float):
    """Named integrand σ(ε(u)):ε(v) with σ = λ tr(ε) I + 2μ ε.  The docs example builds this from opaque
    closures through GT.external (docs/src/src_jl/example_linear_elasticity.jl:76-105), which cannot be
    pattern-matched; the shim ships it as a named integrand instead."""
    def integrand(u, v, x):
        return Named("isotropic_elasticity", dict(lam=lam, mu=mu), (u.arg, v.arg))
    return integrand


@dataclass
class Integral:
    """∫(f, dΩ) → DomainContribution (problems.jl:29-125): list of (term, measure, scale)."""
    contributions: list

    def __add__(self, o):
        return Integral(self.contributions + o.contributions)

    def __rmul__(self, s):
        return Integral([(t, m, s * a) for (t, m, a) in self.contributions])

    def __sub__(self, o):
        return self + (-1.0) * o

    def sum(self):
        """`∫(...) |> sum` = assemble_scalar (problems.jl:173-199)"""
        return assemble_scalar(self)


def integrate(f: Callable, measure: Measure) -> Integral:
    term = f(Coordinate())
    return Integral([(_lift(term), measure, 1.0)])


# ---------------------------------------------------------------------------
# form recognition (compiler.jl/passes.jl stand-in)
# ---------------------------------------------------------------------------
def _flatten_product(t):
    """a*b*c → ([factors], scalar)"""
    if isinstance(t, Call) and t.fn == "*":
        fa, sa = _flatten_product(t.a)
        fb, sb = _flatten_product(t.b)
        return fa + fb, sa * sb
    if isinstance(t, Const) and np.isscalar(t.value):
        return [], float(t.value)
    return [t], 1.0


def recognise_bilinear(term, space: Optional["Space"] = None, meas: Optional["Measure"] = None):
    """→ (form_id, params).  Raises UnsupportedFormError for anything that is not mass / Laplacian / elasticity.
    A scalar AnalyticalField factor κ(x) in front of ∇u·∇v or u v becomes a host-sampled coefficient (coef_qp)."""
    factors, scale = _flatten_product(term)
    coefs = [f for f in factors if isinstance(f, Call) and callable(f.fn) and isinstance(f.a, Coordinate)]
    if len(coefs) == 1 and space is not None and meas is not None:
        rest = [f for f in factors if f is not coefs[0]]
        inner = rest[0] if len(rest) == 1 else None
        if len(rest) == 2:
            inner = Call("*", rest[0], rest[1])
        if inner is not None:
            form, params = recognise_bilinear(inner)
            if form in (_eng.FORM_LAPLACE, _eng.FORM_MASS):
                xq = quadrature_point_coordinates(space, meas)
                vals = np.asarray(coefs[0].fn(np.moveaxis(xq, -1, 0)), dtype=np.float64)
                params["alpha"] = params.get("alpha", 1.0) * scale
                params["coef_qp"] = np.ascontiguousarray(np.broadcast_to(vals, xq.shape[:2]))
                return form, params
    if len(factors) == 1 and isinstance(factors[0], Call) and factors[0].fn == "dot":
        # ∇(v,x)⋅GT.call(dflux, ∇(du,x), ∇(u,x)): Jacobian of the p-Laplacian about the DiscreteField u
        for a, b in ((factors[0].a, factors[0].b), (factors[0].b, factors[0].a)):
            if isinstance(a, FormArg) and a.arg == 1 and a.op == "gradient" and isinstance(b, Call) \
                    and isinstance(b.fn, plaplacian_dflux) and isinstance(b.a, FormArg) and b.a.arg == 2 \
                    and b.a.op == "gradient" and isinstance(b.b, FieldTerm) and b.b.op == "gradient":
                return _eng.FORM_PLAPLACE_JACOBIAN, dict(alpha=scale, exponent=float(b.fn.q), field=b.b.field)
    if len(factors) == 1 and isinstance(factors[0], Named) and factors[0].name == "isotropic_elasticity":
        return _eng.FORM_ELASTICITY_ISO, dict(alpha=scale, **factors[0].params)
    if len(factors) == 1 and isinstance(factors[0], Call) and factors[0].fn == "dot":
        a, b = factors[0].a, factors[0].b
        if isinstance(a, FormArg) and isinstance(b, FormArg) and {a.arg, b.arg} == {1, 2}:
            if a.op == b.op == "gradient":
                return _eng.FORM_LAPLACE, dict(alpha=scale)
            if a.op == b.op == "value":
                return _eng.FORM_MASS, dict(alpha=scale)
    if len(factors) == 2 and all(isinstance(f, FormArg) and f.op == "value" for f in factors) \
            and {factors[0].arg, factors[1].arg} == {1, 2}:
        return _eng.FORM_MASS, dict(alpha=scale)
    raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM,
                               "bilinear form not recognised by the GPU engine (supported: ∫∇u·∇v, ∫u v, "
                               "isotropic_elasticity); refusing to fall back to a CPU loop")


def recognise_linear(term, space: Space, meas: Measure):
    """→ (form_id, params) for ∫ f·v with f a constant or a host-evaluated analytical field, and for the p-Laplacian
    residual ∫ ∇v⋅flux(∇u_h) − f v about a DiscreteField u_h."""
    if isinstance(term, Call) and term.fn == "-" and isinstance(term.a, Call) and term.a.fn == "dot":
        for a, b in ((term.a.a, term.a.b), (term.a.b, term.a.a)):
            if isinstance(a, FormArg) and a.arg == 1 and a.op == "gradient" and isinstance(b, Call) \
                    and isinstance(b.fn, plaplacian_flux) and isinstance(b.a, FieldTerm) and b.a.op == "gradient" and b.b is None:
                sform, sparams = recognise_linear(term.b, space, meas)      # the source part f v
                if space.data.n_comp != 1 or sparams.get("alpha", 1.0) != 1.0:
                    break
                params = dict(alpha=1.0, exponent=float(b.fn.q), field=b.a.field)
                if sform == _eng.FORM_SOURCE_CONST:
                    params["f_const"] = sparams["f_const"]
                else:
                    params["f_qp"] = sparams["f_qp"]
                return _eng.FORM_PLAPLACE_RESIDUAL, params
    factors, scale = _flatten_product(term)
    if isinstance(term, Call) and term.fn == "dot":
        factors = [term.a, term.b]
    tests = [f for f in factors if isinstance(f, FormArg)]
    others = [f for f in factors if not isinstance(f, FormArg)]
    if len(tests) != 1 or tests[0].arg != 1 or tests[0].op != "value" or len(others) > 1:
        raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM,
                                   "linear form not recognised by the GPU engine (supported: ∫ f v with f constant "
                                   "or an analytical field); refusing to fall back to a CPU loop")
    n_comp = space.data.n_comp
    if not others:
        return _eng.FORM_SOURCE_CONST, dict(alpha=scale, f_const=np.ones(n_comp))
    f = others[0]
    if isinstance(f, Const):
        return _eng.FORM_SOURCE_CONST, dict(alpha=scale, f_const=np.broadcast_to(np.asarray(f.value, dtype=float), (n_comp,)))
    if isinstance(f, Call) and callable(f.fn) and isinstance(f.a, Coordinate):
        xq = quadrature_point_coordinates(space, meas)                 # [nc, nq, D]
        vals = np.asarray(f.fn(np.moveaxis(xq, -1, 0)), dtype=np.float64)   # user f gets x[0], x[1], … arrays
        if n_comp == 1:
            vals = np.broadcast_to(vals, xq.shape[:2])[..., None]
        else:
            vals = np.moveaxis(np.broadcast_to(vals, (n_comp,) + xq.shape[:2]), 0, -1)
        return _eng.FORM_SOURCE_QP, dict(alpha=scale, f_qp=np.ascontiguousarray(vals))
    raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "source term not recognised by the GPU engine")


def quadrature_point_coordinates(space: Space, meas: Measure) -> np.ndarray:
    """x_q = Σ_i x_i M_i(ξ_q) (accessors.jl:983-988), all cells (or boundary faces) at once (host input preparation
    for f(x_q))."""
    mesh = space.domain.mesh
    if meas.domain.kind == "boundary":
        fp = space.face_problem(meas)
        X = mesh.node_coordinates[fp.face_nodes.astype(np.int64) - 1]     # [nf, nfn, D]
        return np.einsum("qn,cnd->cqd", fp.tab.M, X)
    tab = space.tabulation(meas.degree)
    X = mesh.node_coordinates[mesh.cell_nodes.astype(np.int64) - 1]       # [nc, nln, D]
    return np.einsum("qn,cnd->cqd", tab.M, X)


# ---------------------------------------------------------------------------
# containers
# ---------------------------------------------------------------------------
@dataclass
class SparseMatrixCSC:
    """Julia's SparseMatrixCSC{Float64,Int32}: 1-based colptr/rowval."""
    m: int
    n: int
    colptr: np.ndarray
    rowval: np.ndarray
    nzval: np.ndarray

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.nzval, self.rowval.astype(np.int64) - 1, self.colptr.astype(np.int64) - 1), shape=(self.m, self.n))

    def sum(self):
        return float(self.nzval.sum())


@dataclass
class AssemblyCache:
    """What `reuse=Val(true)` returns next to the matrix/vector (problems.jl:267-273, 343-349):
    here the engine context holding pattern + plan on the device."""
    engine: _eng.Engine
    form: int
    params: dict
    kind: str


# ---------------------------------------------------------------------------
# assemble_* (problems.jl:244-404)
# ---------------------------------------------------------------------------
_default_device = 0


def set_device(device: int):
    global _default_device
    _default_device = device


def _setup_engine(space: Space, meas: Measure, engine: Optional[_eng.Engine] = None) -> _eng.Engine:
    eng = engine or _eng.Engine(_default_device)
    key = (id(space), meas.domain.kind, None if meas.domain.sides is None else tuple(meas.domain.sides), meas.degree)
    if getattr(eng, "_gt_setup", None) == key:
        return eng          # same mesh / space / measure already resident: keep patterns, plans and the field
    eng._gt_setup = key
    mesh = space.domain.mesh
    if meas.domain.kind == "boundary":
        # the faces of Γ as a mesh of (D-1)-cells embedded in D dimensions (gtk_set_manifold_dim)
        fp = space.face_problem(meas)
        eng.set_mesh(mesh.node_coordinates, fp.face_nodes)
        eng.set_manifold_dim(mesh.D - 1)
        eng.set_space(fp.face_dofs, space.data.n_free, space.data.n_dirichlet, space.data.n_comp)
        eng.set_tabulation(fp.tab.w, fp.tab.N, fp.tab.dN, fp.tab.M, fp.tab.dM)
        return eng
    tab = space.tabulation(meas.degree)
    eng.set_mesh(mesh.node_coordinates, mesh.cell_nodes)
    eng.set_space(space.data.cell_dofs, space.data.n_free, space.data.n_dirichlet, space.data.n_comp)
    eng.set_tabulation(tab.w, tab.N, tab.dN, tab.M, tab.dM)
    return eng


def _require_same_measure(meas_a: Measure, meas_l: Measure):
    """the fused calls set the engine up from the bilinear form's measure: the linear form must integrate over the same
    cells with the same rule, otherwise it would silently be integrated over the wrong domain"""
    da, dl = meas_a.domain, meas_l.domain
    sides = lambda d: None if d.sides is None else tuple(d.sides)
    if meas_a.degree != meas_l.degree or da.kind != dl.kind or da.mesh is not dl.mesh or sides(da) != sides(dl):
        raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM,
                                   "matrix and vector must share one measure (same domain, same degree) in the fused call; "
                                   "assemble them separately with assemble_matrix / assemble_vector")


def _single_contribution(integral: Integral):
    if len(integral.contributions) != 1:
        raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "sums of integrals are assembled one call at a time on the GPU path")
    return integral.contributions[0]


def _upload_field(eng: _eng.Engine, params: dict) -> dict:
    """`parameters=(uh,)`: the DiscreteField a recognised form depends on goes to the engine's field slot
    (problems.jl:276-285, 352-361); returns the params the engine call takes"""
    fld = params.get("field")
    if fld is not None:
        eng.field_set_values(fld.free_values, fld.dirichlet_values)
    return {k: v for k, v in params.items() if k != "field"}


# ---------------------------------------------------------------------------
# product spaces and skeleton integrals: block recognition
# ---------------------------------------------------------------------------
def _expand(t):
    """term -> list of (scalar, factors, dotted): the integrand as a sum of products of form-argument factors"""
    if isinstance(t, Const) and np.isscalar(t.value):
        return [(float(t.value), (), False)]
    if isinstance(t, (FormArg, Normal, FaceDiameter)):
        return [(1.0, (t,), False)]
    if isinstance(t, Call) and t.fn == "/" and t.b is not None:
        den = _expand(t.b)
        if len(den) == 1 and den[0][1] == ():                                   # a / scalar
            return [(c / den[0][0], f, d) for (c, f, d) in _expand(t.a)]
        if len(den) == 1 and len(den[0][1]) == 1 and isinstance(den[0][1][0], FaceDiameter) and not den[0][1][0].inverse:
            return [(c / den[0][0], f + (FaceDiameter(True),), d) for (c, f, d) in _expand(t.a)]   # a / h(x)
        raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "division by anything but a constant or the face diameter is not recognised; no CPU fallback")
    if isinstance(t, Call) and t.fn in ("+", "-") and t.b is not None:
        sign = 1.0 if t.fn == "+" else -1.0
        return _expand(t.a) + [(sign * c, f, d) for (c, f, d) in _expand(t.b)]
    if isinstance(t, Call) and t.fn in ("*", "dot") and t.b is not None:
        return [(ca * cb, fa + fb, da or db or t.fn == "dot") for (ca, fa, da) in _expand(t.a) for (cb, fb, db) in _expand(t.b)]
    raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM,
                               "integrand not recognised by the GPU engine (sums of products of u, v, ∇u, ∇v, div u, div v "
                               "with constant factors); refusing to fall back to a CPU loop")


def _has_face_terms(term) -> bool:
    """unit normals, face diameters or gradients of form arguments somewhere in the term"""
    if isinstance(term, (Normal, FaceDiameter)):
        return True
    if isinstance(term, FormArg):
        return term.op != "value"
    if isinstance(term, Call):
        return _has_face_terms(term.a) or (term.b is not None and _has_face_terms(term.b))
    return False


def _is_blocks_case(space, meas: "Measure", term=None) -> bool:
    """product spaces, skeleton integrals, and boundary integrals that need the cell around the face (Nitsche terms: normals,
    gradients, face diameters) or live on a discontinuous space go to the block kernels"""
    if isinstance(space, ProductSpace) or meas.domain.kind == "skeleton":
        return True
    if meas.domain.kind == "boundary":
        return any(not f.continuous for f in fields(space)) or (term is not None and _has_face_terms(term))
    return False


def _block_problem(space, meas: "Measure"):
    from . import multifield as _mf
    data = [f.data for f in fields(space)]
    if meas.domain.kind == "skeleton":
        return _mf.skeleton_problem(data, meas.degree, gradients=True)
    if meas.domain.kind == "interior":
        return _mf.volume_problem(data, meas.degree)
    if meas.domain.kind == "boundary":
        return _mf.boundary_problem(data, meas.domain.sides, meas.degree)
    raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "block kernels run on interior, skeleton and boundary measures")


def _setup_block_engine(space, meas: "Measure", engine: Optional[_eng.Engine] = None, bp=None):
    bp = bp if bp is not None else _block_problem(space, meas)
    eng = engine or _eng.Engine(_default_device)
    eng.set_mesh(bp.node_coordinates, bp.face_nodes)
    if bp.manifold_dim != bp.node_coordinates.shape[1]:
        eng.set_manifold_dim(bp.manifold_dim)
    eng.set_space(bp.super_dofs, bp.n_free, bp.n_dirichlet, 1)
    eng.set_parts(bp.w, bp.M, bp.dM, bp.parts, bp.n_sides, bp.face_var)
    if bp.dM_cell is not None:
        eng.set_skeleton_cells(bp.cell_nodes, bp.side_cells, bp.dM_cell, bp.ref_normals)
    eng._gt_setup = None
    return eng, bp


def _kind(skeleton_measure) -> str:
    """measure kind of a block problem: "skeleton" (two cells around), "boundary" (one) or "interior"; True / False are the
    skeleton / interior shorthands"""
    if isinstance(skeleton_measure, str):
        return skeleton_measure
    return "skeleton" if skeleton_measure else "interior"


def _part_of(bp, fa: FormArg, skeleton_measure) -> int:
    if _kind(skeleton_measure) == "skeleton":
        if fa.side not in (1, 2):
            raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "on a skeleton measure form arguments are restricted to a cell around: u[1](x), u[2](x)")
        return bp.part_index(fa.field, fa.side - 1)
    if fa.side:
        raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "u[1] / u[2] are meaningful on skeleton measures only")
    return bp.part_index(fa.field, 0)


def recognise_blocks(term, bp, skeleton_measure):
    """-> [(part_u, part_v, block form, alpha)]: one recognised term per (part of u, part of v) block."""
    out = {}
    ip = {}                                  # interior-penalty blocks: key -> [c0, c1, c2]
    for coef, factors, dotted in _expand(term):
        us = [f for f in factors if isinstance(f, FormArg) and f.arg == 2]
        vs = [f for f in factors if isinstance(f, FormArg) and f.arg == 1]
        others = [f for f in factors if not isinstance(f, FormArg)]
        if len(us) != 1 or len(vs) != 1 or len(factors) != 2 + len(others):
            raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "every term of a bilinear integrand must hold exactly one trial and one test factor; no CPU fallback")
        u, v = us[0], vs[0]
        ops = (u.op, v.op)
        if others:
            # terms with unit normals / the face diameter: (1/h)(v n_sv)⋅(u n_su), (v n_sv)⋅∇u, ∇v⋅(u n_su)
            normals = sorted(f.side for f in others if isinstance(f, Normal))
            invh = [f for f in others if isinstance(f, FaceDiameter)]
            mk = _kind(skeleton_measure)
            if mk == "interior" or len(normals) + len(invh) != len(others) or any(not f.inverse for f in invh) or (normals and not dotted):
                raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "normal / face-diameter term not recognised by the GPU engine; no CPU fallback")
            if ops == ("value", "value") and len(invh) == 1 and (normals == sorted([u.side, v.side]) or (mk == "boundary" and not normals)):
                kind = 0          # boundary: (γ/h) v u = (γ/h)(v n)⋅(u n) with the one normal
            elif ops == ("gradient", "value") and not invh and normals == [v.side]:
                kind = 1
            elif ops == ("value", "gradient") and not invh and normals == [u.side]:
                kind = 2
            else:
                raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "normal / face-diameter term not recognised by the GPU engine; no CPU fallback")
            key = (_part_of(bp, u, mk), _part_of(bp, v, mk))
            ip.setdefault(key, [0.0, 0.0, 0.0])[kind] += coef
            continue
        if ops == ("value", "value"):
            form = _eng.BLOCK_MASS
        elif ops == ("gradient", "gradient") and dotted:
            form = _eng.BLOCK_LAPLACE
        elif ops == ("value", "divergence"):
            form = _eng.BLOCK_VALU_DIVV
        elif ops == ("divergence", "value"):
            form = _eng.BLOCK_DIVU_VALV
        else:
            raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, f"block term ({u.op} of u) x ({v.op} of v) is not recognised by the GPU engine; no CPU fallback")
        if form != _eng.BLOCK_MASS and _kind(skeleton_measure) != "interior":
            raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "on faces, gradients enter through normal terms only ((v n)⋅∇u, ∇v⋅(u n)); no CPU fallback")
        key = (_part_of(bp, u, skeleton_measure), _part_of(bp, v, skeleton_measure))
        if key in out:
            if out[key][0] != form:
                raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "two different terms in one block are not recognised; no CPU fallback")
            out[key] = (form, out[key][1] + coef)
        else:
            out[key] = (form, coef)
    merged = []
    for key in sorted(set(ip) & set(out)):
        # v u next to normal terms in one block (Nitsche without the 1/h scaling, test/issue_224.jl:73-76): on a boundary face
        # v u = (v n)⋅(u n), so it is the c0 term of the interior-penalty form WITHOUT the division by h
        form, alpha = out[key]
        if form != _eng.BLOCK_MASS or ip[key][0] != 0.0 or _kind(skeleton_measure) != "boundary":
            raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "two different terms in one block are not recognised; no CPU fallback")
        merged.append((key[0], key[1], _eng.BLOCK_IP_NOH, 1.0, (alpha, ip[key][1], ip[key][2])))
        del out[key], ip[key]
    return [(pu, pv, form, alpha) for (pu, pv), (form, alpha) in out.items()] + \
           [(pu, pv, _eng.BLOCK_IP, 1.0, tuple(c)) for (pu, pv), c in ip.items()] + merged


def recognise_vblocks(term, bp, skeleton_measure, space):
    """linear forms -> ([(part, alpha, f_const or (c0, c1, c2))], g_qp or None):
    constant multiples of v(x) (f_const), or — with an analytical field g sampled at the face points —
    g (c0 v + (c1/h) v + c2 n⋅∇v): volume sources v f and the Nitsche right-hand side (γ/h) v g - n⋅∇v g"""
    mk = _kind(skeleton_measure)
    flds = fields(space)
    const, data, fn = {}, {}, None
    for coef, factors, dotted in _expand_with_data(term):
        vs = [f for f in factors if isinstance(f, FormArg)]
        normals = [f for f in factors if isinstance(f, Normal)]
        invh = [f for f in factors if isinstance(f, FaceDiameter)]
        datas = [f for f in factors if isinstance(f, Call)]
        if len(vs) != 1 or vs[0].arg != 1 or len(vs) + len(normals) + len(invh) + len(datas) != len(factors) or len(datas) > 1:
            raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "linear block term not recognised by the GPU engine; no CPU fallback")
        v = vs[0]
        part = _part_of(bp, v, mk)
        if not datas:
            if v.op != "value" or normals or invh:
                raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "linear block terms without data are constant multiples of v(x); no CPU fallback")
            const[part] = const.get(part, 0.0) + coef
            continue
        if fn is not None and datas[0].fn is not fn:
            raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "one analytical field per linear block integral; no CPU fallback")
        fn = datas[0].fn
        if v.op == "value" and not normals and not invh:
            kind = 0
        elif v.op == "value" and not normals and len(invh) == 1 and invh[0].inverse and mk != "interior":
            kind = 1
        elif v.op == "gradient" and len(normals) == 1 and normals[0].side == v.side and not invh and dotted and mk != "interior":
            kind = 2
        else:
            raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "linear block term not recognised by the GPU engine; no CPU fallback")
        data.setdefault(part, [0.0, 0.0, 0.0])[kind] += coef
    if const and data:
        raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "constant and data terms in one linear block integral are not recognised; no CPU fallback")
    if data:
        from . import multifield as _mf
        xq = _mf.face_point_coordinates(bp)                                     # [n_faces, nq, D]
        g = np.asarray(fn(np.moveaxis(xq, -1, 0)), dtype=np.float64)
        g = np.ascontiguousarray(np.broadcast_to(g, xq.shape[:2]))
        return [(part, 1.0, tuple(c)) for part, c in data.items()], g
    return [(part, alpha, np.ones(flds[bp.parts[part]["field"]].data.n_comp)) for part, alpha in const.items()], None


def _expand_with_data(t):
    """_expand that also accepts analytical-field factors f(x) (Call(fn, Coordinate)) as opaque data factors"""
    if isinstance(t, Call) and callable(t.fn) and isinstance(t.a, Coordinate):
        return [(1.0, (t,), False)]
    if isinstance(t, Call) and t.fn in ("+", "-") and t.b is not None:
        sign = 1.0 if t.fn == "+" else -1.0
        return _expand_with_data(t.a) + [(sign * c, f, d) for (c, f, d) in _expand_with_data(t.b)]
    if isinstance(t, Call) and t.fn in ("*", "dot") and t.b is not None:
        return [(ca * cb, fa + fb, da or db or t.fn == "dot") for (ca, fa, da) in _expand_with_data(t.a) for (cb, fb, db) in _expand_with_data(t.b)]
    if isinstance(t, Call) and t.fn == "/" and t.b is not None:
        den = _expand(t.b)
        num = _expand_with_data(t.a)
        if len(den) == 1 and den[0][1] == ():
            return [(c / den[0][0], f, d) for (c, f, d) in num]
        if len(den) == 1 and len(den[0][1]) == 1 and isinstance(den[0][1][0], FaceDiameter) and not den[0][1][0].inverse:
            return [(c / den[0][0], f + (FaceDiameter(True),), d) for (c, f, d) in num]
    return _expand(t)


def _form_arguments(space, arg: int):
    if isinstance(space, ProductSpace):
        return tuple(FormArgument(f, arg, k) for k, f in enumerate(space.fields))
    return FormArgument(space, arg)


def _assemble_matrix_blocks(a, U, V, reuse, free_or_dirichlet, engine, index_type):
    if U is not V:
        raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "trial and test spaces must be the same object on the GPU path")
    term, meas, scale = _single_contribution(a(_form_arguments(U, 2), _form_arguments(V, 1)))
    bp = _block_problem(V, meas)
    blocks = [(b[0], b[1], b[2], b[3] * scale) + tuple(b[4:]) for b in recognise_blocks(term, bp, meas.domain.kind)]   # raises before any engine exists
    eng, bp = _setup_block_engine(V, meas, engine, bp)
    eng.matrix_symbolic(*free_or_dirichlet)
    colptr, rowval = eng.matrix_pattern_i64() if index_type in (int, np.int64) else eng.matrix_pattern()
    nzval = eng.matrix_numeric_blocks(blocks)
    A = SparseMatrixCSC(eng.n_rows, eng.n_cols, colptr, rowval, nzval)
    if reuse:
        return A, AssemblyCache(eng, "blocks", dict(blocks=blocks), "matrix")
    eng.close()
    return A


def _assemble_vector_blocks(l, V, reuse, free_or_dirichlet, engine):
    term, meas, scale = _single_contribution(l(_form_arguments(V, 1)))
    bp = _block_problem(V, meas)
    vb, g_qp = recognise_vblocks(term, bp, meas.domain.kind, V)
    eng, bp = _setup_block_engine(V, meas, engine, bp)
    vblocks = [(part, alpha * scale, f) for (part, alpha, f) in vb]
    eng.vector_symbolic(free_or_dirichlet)
    b = eng.vector_assemble_blocks(vblocks, g_qp=g_qp)
    if reuse:
        return b, AssemblyCache(eng, "blocks", dict(vblocks=vblocks, g_qp=g_qp), "vector")
    eng.close()
    return b


def _assemble_matrix_sum(contributions, U, V, reuse, free_or_dirichlet, index_type):
    """a(u,v) = ∫_Ω … + ∫_Γ … + ∫_Λ …: the reference pushes every contribution into one COO allocation and compresses once
    (problems.jl:319-350).  One engine context per integral (its own integration faces), merged on the device
    (gtk_matrix_sum_*): union pattern, values summed in the order of the contributions."""
    parts = []
    try:
        for term, meas, scale in contributions:
            if _is_blocks_case(V, meas, term):
                bp = _block_problem(V, meas)
                blocks = [(b[0], b[1], b[2], b[3] * scale) + tuple(b[4:]) for b in recognise_blocks(term, bp, meas.domain.kind)]
                eng, bp = _setup_block_engine(V, meas, None, bp)
                run = (lambda e=eng, b=blocks: e.matrix_numeric_blocks_device(b))
            else:
                form, params = recognise_bilinear(term, V, meas)
                if meas.domain.kind == "boundary" and form != _eng.FORM_MASS:
                    raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "on a boundary measure only ∫_Γ u v dΓ (Robin term) is assembled by the GPU engine")
                if "field" in params:
                    raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "forms with parameters are assembled one integral at a time on the GPU path")
                params["alpha"] = params.get("alpha", 1.0) * scale
                eng = _setup_engine(V, meas)
                run = (lambda e=eng, f=form, p=params: e.matrix_numeric_device(f, **p))
            parts.append((eng, run))
            eng.matrix_symbolic(*free_or_dirichlet)
            run()
        total = _eng.Engine(_default_device)
        sources = [e for e, _ in parts]
        total.matrix_sum_symbolic(sources)
        colptr, rowval = total.matrix_pattern_i64() if index_type in (int, np.int64) else total.matrix_pattern()
        nzval = total.matrix_sum_numeric(sources)
    except Exception:
        for e, _ in parts:
            e.close()
        raise
    A = SparseMatrixCSC(total.n_rows, total.n_cols, colptr, rowval, nzval)
    if reuse:
        return A, AssemblyCache(total, "sum", dict(parts=parts), "matrix")
    for e, _ in parts:
        e.close()
    total.close()
    return A


def assemble_matrix(a: Callable, T, U: Space, V: Space, *, reuse: bool = False, parameters=(),
                    free_or_dirichlet=(FREE, FREE), engine: Optional[_eng.Engine] = None, assembly_options=None):
    """GT.assemble_matrix(a, T, U, V; reuse, free_or_dirichlet) (problems.jl:319-350).
    Rows enumerate V (test), columns U (trial); only U is V is supported on the GPU path."""
    if T not in (float, np.float64):
        raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "the engine assembles Float64 only")
    if U is not V:
        raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "trial and test spaces must be the same object on the GPU path")
    probe = a(_form_arguments(U, 2), _form_arguments(V, 1))
    if len(probe.contributions) == 1 and _is_blocks_case(V, probe.contributions[0][1], probe.contributions[0][0]):
        # product spaces / skeleton integrals: block kernels on the super element
        opts = dict(assembly_options or {})
        index_type = opts.pop("index_type", np.int32)
        if opts.pop("eltype", np.float64) not in (float, np.float64) or opts or parameters:
            raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "assembly_options / parameters not supported for product spaces on the GPU engine")
        return _assemble_matrix_blocks(a, U, V, reuse, free_or_dirichlet, engine, index_type)
    if len(probe.contributions) > 1:
        # a sum of integrals, possibly over different domains: one context per integral, merged on the device
        opts = dict(assembly_options or {})
        index_type = opts.pop("index_type", np.int32)
        if opts.pop("eltype", np.float64) not in (float, np.float64) or opts or parameters or engine is not None:
            raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "assembly_options / parameters / engine are not supported for sums of integrals on the GPU engine")
        return _assemble_matrix_sum(probe.contributions, U, V, reuse, free_or_dirichlet, index_type)
    term, meas, scale = _single_contribution(probe)
    form, params = recognise_bilinear(term, V, meas)
    if meas.domain.kind == "boundary" and form != _eng.FORM_MASS:
        raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "on a boundary measure only ∫_Γ u v dΓ (Robin term) is assembled by the GPU engine")
    params["alpha"] = params.get("alpha", 1.0) * scale
    # assembly_options (assembly.jl:434-445): index_type Int32 (default) | Int64; eltype / matrix_type other than
    # Float64 / SparseMatrixCSC are not something this engine produces: explicit error
    opts = dict(assembly_options or {})
    index_type = opts.pop("index_type", np.int32)
    if opts.pop("eltype", np.float64) not in (float, np.float64) or opts:
        raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, f"assembly_options {assembly_options} are not supported by the GPU engine")
    eng = _setup_engine(V, meas, engine)
    eng.matrix_symbolic(*free_or_dirichlet)
    colptr, rowval = eng.matrix_pattern_i64() if index_type in (int, np.int64) else eng.matrix_pattern()
    nzval = eng.matrix_numeric(form, **_upload_field(eng, params))
    A = SparseMatrixCSC(eng.n_rows, eng.n_cols, colptr, rowval, nzval)
    if reuse or parameters:
        return A, AssemblyCache(eng, form, params, "matrix")
    eng.close()
    return A


def update_matrix(A: SparseMatrixCSC, cache: AssemblyCache, parameters=(), **new_params):
    """GT.update_matrix!(A, cache; parameters) (problems.jl:352-361): numeric re-assembly on the cached pattern; a
    DiscreteField in `parameters` replaces the one the form was recognised with."""
    if cache.form == "blocks":
        cache.engine.matrix_numeric_blocks(cache.params["blocks"], out=A.nzval)
        return A
    if cache.form == "sum":
        for _, run in cache.params["parts"]:
            run()
        cache.engine.matrix_sum_numeric([e for e, _ in cache.params["parts"]], out=A.nzval)
        return A
    cache.params.update(new_params)
    if parameters:
        cache.params["field"] = parameters[0]
    cache.engine.matrix_numeric(cache.form, out=A.nzval, **_upload_field(cache.engine, cache.params))
    return A


def assemble_vector(l: Callable, T, V: Space, *, reuse: bool = False, parameters=(), free_or_dirichlet=FREE,
                    engine: Optional[_eng.Engine] = None):
    """GT.assemble_vector(l, T, V; reuse) (problems.jl:244-274)."""
    if T not in (float, np.float64):
        raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "the engine assembles Float64 only")
    v = _form_arguments(V, 1)
    contributions = l(v).contributions
    if len(contributions) == 1 and _is_blocks_case(V, contributions[0][1], contributions[0][0]):
        if parameters:
            raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "parameters are not supported for product spaces on the GPU engine")
        return _assemble_vector_blocks(l, V, reuse, free_or_dirichlet, engine)
    if len(contributions) > 1:
        # ∫_Ω f v dΩ + ∫_Γ g v dΓ + …: the reference pushes every integral into ONE COO vector, contribution after
        # contribution (problems.jl:258-266); each integral is one engine pass that continues the sums of the previous
        if reuse or engine is not None:
            raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "reuse is available for single-integral linear forms")
        b = None
        for term, meas, scale in contributions:
            if _is_blocks_case(V, meas, term):          # skeleton / Nitsche / product-space integral: block kernels, same COO vector
                bp = _block_problem(V, meas)
                vb, g_qp = recognise_vblocks(term, bp, meas.domain.kind, V)
                eng, bp = _setup_block_engine(V, meas, None, bp)
                eng.vector_symbolic(free_or_dirichlet)
                if b is not None:
                    eng.set_vector(b)
                b = eng.vector_assemble_blocks([(part, alpha * scale, f) for (part, alpha, f) in vb], accumulate=b is not None, g_qp=g_qp)
                eng.close()
                continue
            form, params = recognise_linear(term, V, meas)
            params["alpha"] = params.get("alpha", 1.0) * scale
            eng = _setup_engine(V, meas)
            eng.vector_symbolic(free_or_dirichlet)
            if b is not None:
                eng.set_vector(b)
                params["accumulate"] = True
            b = eng.vector_assemble(form, **params)
            eng.close()
        return b
    term, meas, scale = contributions[0]
    form, params = recognise_linear(term, V, meas)
    params["alpha"] = params.get("alpha", 1.0) * scale
    eng = _setup_engine(V, meas, engine)
    eng.vector_symbolic(free_or_dirichlet)
    b = eng.vector_assemble(form, **_upload_field(eng, params))
    if reuse or parameters:
        return b, AssemblyCache(eng, form, params, "vector")
    eng.close()
    return b


def update_vector(b: np.ndarray, cache: AssemblyCache, parameters=(), **new_params):
    """GT.update_vector!(b, cache; parameters) (problems.jl:276-285)."""
    if cache.form == "blocks":
        cache.engine.vector_assemble_blocks(cache.params["vblocks"], out=b, g_qp=cache.params.get("g_qp"))
        return b
    cache.params.update(new_params)
    if parameters:
        cache.params["field"] = parameters[0]
    cache.engine.vector_assemble(cache.form, out=b, **_upload_field(cache.engine, cache.params))
    return b


def assemble_matrix_and_vector(a: Callable, l: Callable, T, U: Space, V: Space, *, reuse: bool = False):
    """GT.assemble_matrix_and_vector(a, l, T, U, V) (problems.jl:391-404): one fused pass over the cells."""
    if U is not V:
        raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "trial and test spaces must be the same object on the GPU path")
    u, v = FormArgument(U, 2), FormArgument(V, 1)
    term_a, meas_a, sa = _single_contribution(a(u, v))
    term_l, meas_l, sl = _single_contribution(l(v))
    _require_same_measure(meas_a, meas_l)
    mform, mparams = recognise_bilinear(term_a)
    vform, vparams = recognise_linear(term_l, V, meas_l)
    mparams["alpha"] = mparams.get("alpha", 1.0) * sa
    vparams["alpha"] = vparams.get("alpha", 1.0) * sl
    eng = _setup_engine(V, meas_a)
    eng.matrix_symbolic(FREE, FREE)
    colptr, rowval = eng.matrix_pattern()
    nzval, b = eng.assemble_matrix_and_vector(mform, mparams, vform, vparams)
    A = SparseMatrixCSC(eng.n_rows, eng.n_cols, colptr, rowval, nzval)
    if reuse:
        return A, b, AssemblyCache(eng, mform, dict(m=mparams, v=vparams, vform=vform), "both")
    eng.close()
    return A, b


# ---------------------------------------------------------------------------
# free + Dirichlet columns and the linear-problem right-hand side (problems.jl:363-387, 413-453)
# ---------------------------------------------------------------------------
def assemble_matrix_and_vector_with_free_and_dirichlet_columns(a: Callable, l: Callable, T, U: Space, V: Space, *,
                                                               reuse: bool = False, engine: Optional[_eng.Engine] = None):
    """GT.assemble_matrix_and_vector_with_free_and_dirichlet_columns(a, l, T, U, V) (problems.jl:413-430):
    A = free x free, Ad = free x Dirichlet, b.  The reference runs the cell loop twice for the matrices ("this can be
    optimized", problems.jl:369); here both live in one engine context (matrix slots 0 and 1) next to b."""
    if U is not V:
        raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "trial and test spaces must be the same object on the GPU path")
    u, v = FormArgument(U, 2), FormArgument(V, 1)
    term_a, meas_a, sa = _single_contribution(a(u, v))
    term_l, meas_l, sl = _single_contribution(l(v))
    _require_same_measure(meas_a, meas_l)
    mform, mparams = recognise_bilinear(term_a)
    vform, vparams = recognise_linear(term_l, V, meas_l)
    mparams["alpha"] = mparams.get("alpha", 1.0) * sa
    vparams["alpha"] = vparams.get("alpha", 1.0) * sl
    eng = _setup_engine(V, meas_a, engine)
    eng.select_matrix(0)
    eng.matrix_symbolic(FREE, FREE)
    colptr, rowval = eng.matrix_pattern()
    nzval, b = eng.assemble_matrix_and_vector(mform, mparams, vform, vparams)
    A = SparseMatrixCSC(eng.n_rows, eng.n_cols, colptr, rowval, nzval)
    eng.select_matrix(1)
    eng.matrix_symbolic(FREE, DIRICHLET)
    cpd, rvd = eng.matrix_pattern()
    Ad = SparseMatrixCSC(eng.n_rows, eng.n_cols, cpd, rvd, eng.matrix_numeric(mform, **mparams))
    eng.select_matrix(0)
    if reuse:
        return A, Ad, b, AssemblyCache(eng, mform, dict(m=mparams, v=vparams, vform=vform), "both+dirichlet")
    eng.close()
    return A, Ad, b


def linear_problem(dirichlet_values: np.ndarray, a: Callable, l: Callable, U: Space, V: Optional[Space] = None):
    """PartitionedSolvers_linear_problem(uhd, a, l) (problems.jl:439-453): returns (x0, A, b) with
    b = l - Ad*xd computed on the device exactly like `mul!(b, Ad, xd, -1, 1)`, x0 = zeros."""
    V = U if V is None else V
    xd = np.ascontiguousarray(dirichlet_values, dtype=np.float64)
    A, Ad, b, cache = assemble_matrix_and_vector_with_free_and_dirichlet_columns(a, l, np.float64, U, V, reuse=True)
    eng = cache.engine
    eng.select_matrix(1)
    b = eng.matvec_add(-1.0, xd, 1.0)
    eng.select_matrix(0)
    eng.close()
    return np.zeros(A.n, dtype=np.float64), A, b


# ---------------------------------------------------------------------------
# 0-forms: assemble_scalar (problems.jl:173-199)
# ---------------------------------------------------------------------------
def _recognise_scalar(term, space_hint=None):
    """→ (kind, field, g_fn or None): abs2(uh(x)), abs2(u(x) − uh(x)), ∇e⋅∇e with e = u − uh (or uh alone), 1"""
    def split_diff(t, op):
        # → (field, analytic callable or None, sign of the field) for `uh`, `u − uh`, `uh − u`
        if isinstance(t, FieldTerm) and t.op == op:
            return t.field, None
        if isinstance(t, Call) and t.fn == "-":
            for f, g in ((t.a, t.b), (t.b, t.a)):
                if isinstance(f, FieldTerm) and f.op == op and isinstance(g, Call) and callable(g.fn) and isinstance(g.a, Coordinate):
                    return f.field, g.fn
        return None
    if isinstance(term, Const) and term.value == 1:
        return _eng.SCALAR_VOLUME, None, None
    if isinstance(term, Call) and term.fn == "abs2":
        r = split_diff(term.a, "value")
        if r:
            return _eng.SCALAR_L2SQ, r[0], r[1]
    if isinstance(term, Call) and term.fn == "dot":
        ra, rb = split_diff(term.a, "gradient"), split_diff(term.b, "gradient")
        if ra and rb and ra[0] is rb[0] and ra[1] is rb[1]:
            return _eng.SCALAR_H1SQ, ra[0], ra[1]
    raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM,
                               "scalar integrand not recognised by the GPU engine (supported: 1, abs2(uh), abs2(u - uh), "
                               "∇e⋅∇e with e = uh or u - uh); refusing to fall back to a CPU loop")


def assemble_scalar(integral: Integral, space: Optional[Space] = None) -> float:
    """GT.assemble_scalar(∫(...)) = `∫(...) |> sum` (problems.jl:173-199): Σ over contributions of coefficient · Σ_cells Σ_q
    integrand·dV, each on the device (gtk_scalar_assemble)."""
    total = 0.0
    for term, meas, scale in integral.contributions:
        kind, fld, g = _recognise_scalar(term)
        V = fld.space if fld is not None else space
        if V is None:
            raise UnsupportedFormError(_eng.GTK_ERR_UNSUPPORTED_FORM, "∫ 1 needs a space to pick the cells from (pass space=)")
        eng = _setup_engine(V, meas)
        params = {}
        if fld is not None:
            eng.field_set_values(fld.free_values, fld.dirichlet_values)
        if g is not None:
            xq = quadrature_point_coordinates(V, meas)
            vals = np.asarray(g(np.moveaxis(xq, -1, 0)), dtype=np.float64)
            if kind == _eng.SCALAR_H1SQ:
                vals = np.moveaxis(np.broadcast_to(vals, (xq.shape[-1],) + xq.shape[:2]), 0, -1)
            else:
                vals = np.broadcast_to(vals, xq.shape[:2])
            params["f_qp"] = np.ascontiguousarray(vals)
        total += scale * eng.scalar_assemble(kind, **params)
        eng.close()
    return total


# ---------------------------------------------------------------------------
# Dirichlet data and solution fields (space.jl:1876-1897, 2000-2060; problems.jl:501-526)
# ---------------------------------------------------------------------------
def _reference_node_tabulation(V: Space) -> np.ndarray:
    """tabulator(refface)(value, node_coordinates(reffe)) (space.jl:1918-1922): geometry shape functions at the reference
    nodes of the space, [n_lnodes_space, n_lnodes_mesh]"""
    d = V.data
    return _hp.tabulate(d.mesh.D, 1, d.kind, _hp.reference_nodes(d.mesh.D, d.order, d.kind))[0]


def dof_coordinates(V: Space):
    """(x_free, x_dirichlet): node_coordinates(V) seen through free_dof_node / dirichlet_dof_node (space.jl:1876-1897,
    1960-1998), computed on the device (gtk_space_dof_coordinates) and cached on the space"""
    if "xdof" not in V._tab:
        mesh = V.domain.mesh
        eng = _eng.Engine(_default_device)
        eng.set_mesh(mesh.node_coordinates, mesh.cell_nodes)
        eng.set_space(V.data.cell_dofs, V.data.n_free, V.data.n_dirichlet, V.data.n_comp)
        V._tab["xdof"] = eng.space_dof_coordinates(_reference_node_tabulation(V))
        eng.close()
    return V._tab["xdof"]


def _dof_component(V: Space, free: bool) -> np.ndarray:
    """component of every free / Dirichlet dof: local dof = node * n_comp + c in every cell (space.jl:1267-1271)"""
    d = V.data.cell_dofs
    nc = V.data.n_comp
    lc = np.tile(np.arange(d.shape[1]) % nc, (d.shape[0], 1))
    if free:
        comp = np.zeros(V.data.n_free, dtype=np.int64)
        sel = d > 0
        comp[d[sel] - 1] = lc[sel]
    else:
        comp = np.zeros(V.data.n_dirichlet, dtype=np.int64)
        sel = d < 0
        comp[-d[sel] - 1] = lc[sel]
    return comp


def _evaluate_at(g, X: np.ndarray, V: Space, free: bool) -> np.ndarray:
    g = g.f if isinstance(g, AnalyticalField) else g
    vals = np.asarray(g(np.moveaxis(X, -1, 0)), dtype=np.float64)
    nc = V.data.n_comp
    if nc == 1:
        return np.ascontiguousarray(np.broadcast_to(vals, X.shape[:1]))
    vals = np.broadcast_to(vals, (nc,) + X.shape[:1])
    return np.ascontiguousarray(vals[_dof_component(V, free), np.arange(X.shape[0])])


def interpolate_dirichlet(g: Callable, V):
    """GT.interpolate_dirichlet(g, V) / interpolate_dirichlet!(g, uh) (field.jl:352-373; space.jl:2000-2060): the Dirichlet
    value of a dof is g at its node (component c of g for vector spaces).  Given a space: → xd [n_dirichlet] (what
    `dirichlet_values(interpolate_dirichlet(g, V))` holds); given a DiscreteField: updates it in place and returns it."""
    if isinstance(V, DiscreteField):
        V.dirichlet_values[:] = _evaluate_at(g, dof_coordinates(V.space)[1], V.space, False)
        return V
    return _evaluate_at(g, dof_coordinates(V)[1], V, False)


def interpolate_free(g: Callable, V):
    """GT.interpolate_free(g, V) / interpolate_free!(g, uh) (field.jl:352-360)"""
    if isinstance(V, DiscreteField):
        V.free_values[:] = _evaluate_at(g, dof_coordinates(V.space)[0], V.space, True)
        return V
    return _evaluate_at(g, dof_coordinates(V)[0], V, True)


def interpolate(g: Callable, V: Space) -> DiscreteField:
    """GT.interpolate(g, V) (field.jl:335-345)"""
    xf, xd = dof_coordinates(V)
    return DiscreteField(V, _evaluate_at(g, xf, V, True), _evaluate_at(g, xd, V, False))


def solution_field(U, x, xd=None):
    """GT.solution_field(uhd | U, x | problem) (problems.jl:501-563): the DiscreteField with free values x and the
    Dirichlet values of uhd (zero for a space).  Legacy form solution_field(V, x, xd) -> the value of every local dof of
    every cell [n_cells, n_ldofs] (kept for the linear-problem tests)."""
    if isinstance(x, NonlinearProblem):
        x = x.x
    if xd is not None and isinstance(U, Space):
        d = U.data.cell_dofs.astype(np.int64)
        x = np.asarray(x, dtype=np.float64)
        xd = np.asarray(xd, dtype=np.float64)
        out = np.empty(d.shape, dtype=np.float64)
        pos = d > 0
        out[pos] = x[d[pos] - 1]
        out[~pos] = xd[-d[~pos] - 1]
        return out
    if isinstance(U, DiscreteField):
        return DiscreteField(U.space, np.array(x, dtype=np.float64), U.dirichlet_values)
    return DiscreteField(U, np.array(x, dtype=np.float64), np.zeros(U.num_dirichlet_dofs()))


# ---------------------------------------------------------------------------
# Nonlinear problems (problems.jl:465-497)
# ---------------------------------------------------------------------------
class NonlinearProblem:
    """PartitionedSolvers_nonlinear_problem(uh, r, j) (problems.jl:465-478): x = free values of uh, b = residual,
    A = Jacobian, both assembled on ONE engine context (same mesh/space/measure) and re-assembled by `update`
    (nonlinear_problem_update, problems.jl:480-497) with u_h resident in HBM."""

    def __init__(self, uh: DiscreteField, r: Callable, j: Callable, V: Optional[Space] = None):
        U = uh.space
        V = U if V is None else V
        self.uh = uh
        self.x = uh.free_values.copy()
        self.b, self.residual_cache = assemble_vector(r(uh), np.float64, V, parameters=(uh,))
        self.A, self.jacobian_cache = assemble_matrix(j(uh), np.float64, U, V, parameters=(uh,),
                                                      engine=self.residual_cache.engine)

    def update(self, x, residual=True, jacobian=True):
        """solution_field!(uh, x) then update_vector! / update_matrix! with parameters=(uh,)"""
        self.x = np.array(x, dtype=np.float64)
        self.uh.free_values[:] = self.x
        if residual:
            update_vector(self.b, self.residual_cache, parameters=(self.uh,))
        if jacobian:
            update_matrix(self.A, self.jacobian_cache, parameters=(self.uh,))
        return self

    def close(self):
        self.residual_cache.engine.close()


def nonlinear_problem(uh: DiscreteField, r: Callable, j: Callable, V: Optional[Space] = None) -> NonlinearProblem:
    return NonlinearProblem(uh, r, j, V)


def newton_solve(p: NonlinearProblem, *, rtol=1e-12, atol=1e-13, maxiter=60, verbose=False) -> NonlinearProblem:
    """Stand-in for `NonlinearSolve.solve(prob)` / `PS.NLsolve_nlsolve(p; method=:newton)` of the reference's tests: Newton
    with a backtracking line search on ‖r‖₂; the linear solves run on the host (SciPy sparse LU — solvers are out of
    scope), every residual and Jacobian is a device re-assembly."""
    import scipy.sparse.linalg as spl
    r0 = None
    for it in range(maxiter):
        nr = float(np.linalg.norm(p.b))
        r0 = nr if r0 is None else r0
        if verbose:
            print(f"newton {it}: |r| = {nr:.3e}")
        if nr <= atol or nr <= rtol * r0:
            return p
        dx = spl.spsolve(p.A.to_scipy().tocsc(), -p.b)
        x0, t = p.x.copy(), 1.0
        while True:
            p.update(x0 + t * dx, jacobian=False)
            if float(np.linalg.norm(p.b)) < (1.0 - 1e-4 * t) * nr or t < 1e-8:
                break
            t *= 0.5
        p.update(p.x, residual=False)
    raise RuntimeError("newton_solve: no convergence")
