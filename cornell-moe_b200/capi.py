"""ctypes binding of the C ABI declared in include/cmoe_b200.h.

This is the thinnest possible host layer: numpy arrays in, numpy arrays out, every call goes straight to
libcornell_moe_b200.so (hand-written sm_100a kernels).  If the library is missing or no CUDA device is visible the
calls raise — there is deliberately no fallback.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CMOE_B200_LIB", os.path.join(_HERE, "libcornell_moe_b200.so"))

OK, ERR_SINGULAR, ERR_BOUNDS, ERR_INVALID_VALUE, ERR_RUNTIME, ERR_NO_DEVICE = range(6)
SQUARE_EXPONENTIAL, MATERN_NU_2P5 = 0, 1

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


class CmoeError(RuntimeError):
    """Base error (reference: OptimalLearningException)."""

    def __init__(self, code, msg, info=0):
        super().__init__(msg)
        self.code = code
        self.info = info


class SingularMatrixError(CmoeError):
    """reference: SingularMatrixException; ``info`` = leading minor index."""


class BoundsError(CmoeError):
    """reference: BoundsException."""


class InvalidValueError(CmoeError):
    """reference: InvalidValueException."""


class NoDeviceError(CmoeError):
    """No CUDA device: the product has no CPU path."""


_ERRS = {ERR_SINGULAR: SingularMatrixError, ERR_BOUNDS: BoundsError, ERR_INVALID_VALUE: InvalidValueError,
         ERR_RUNTIME: CmoeError, ERR_NO_DEVICE: NoDeviceError}


class GDParams(ctypes.Structure):
    """GradientDescentParameters (gpp_optimizer_parameters.hpp:81)."""
    _fields_ = [("num_multistarts", ctypes.c_int), ("max_num_steps", ctypes.c_int), ("max_num_restarts", ctypes.c_int),
                ("num_steps_averaged", ctypes.c_int), ("gamma", ctypes.c_double), ("pre_mult", ctypes.c_double),
                ("max_relative_change", ctypes.c_double), ("tolerance", ctypes.c_double)]

    @classmethod
    def from_seq(cls, p):
        if isinstance(p, cls):
            return p
        return cls(int(p[0]), int(p[1]), int(p[2]), int(p[3]), float(p[4]), float(p[5]), float(p[6]), float(p[7]))


class KGStats(ctypes.Structure):
    _fields_ = [("mc_samples", ctypes.c_uint64), ("posterior_evals", ctypes.c_uint64),
                ("line_search_steps", ctypes.c_uint64), ("point_evals", ctypes.c_uint64),
                ("line_batches", ctypes.c_uint64)]


_lib = None


def lib():
    """Loads libcornell_moe_b200.so (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `make -C cornell-moe_b200` (or __graft_entry__.build()); "
                "there is no fallback implementation")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.cmoe_last_error.restype = ctypes.c_char_p
        _lib.cmoe_version.restype = ctypes.c_char_p
    return _lib


def _check(rc, info=0):
    if rc != OK:
        msg = lib().cmoe_last_error().decode()
        raise _ERRS.get(rc, CmoeError)(rc, msg, info)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.zeros(0, dtype=np.int32) if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def set_option(name, value):
    """cmoe_set_option: 'legacy_linalg' / 'cov_tma' (0 or 1)."""
    _check(lib().cmoe_set_option(name.encode(), int(value)))


def device_count():
    return lib().cmoe_device_count()


def version():
    return lib().cmoe_version().decode()


def cholesky(A, device=0):
    """Lower Cholesky of a symmetric [n, n] array through the device path; returns L as [row, col] array."""
    A = np.array(A, dtype=np.float64)
    n = A.shape[0]
    buf = np.ascontiguousarray(A.T)
    info = ctypes.c_int(0)
    rc = lib().cmoe_cholesky(n, _d(buf), int(device), ctypes.byref(info))
    _check(rc, info.value)
    return buf.T.copy()


def potrs(L, B, device=0):
    Lc = np.ascontiguousarray(np.array(L, dtype=np.float64).T)
    B = np.array(B, dtype=np.float64)
    B2 = B.reshape(B.shape[0], -1)
    buf = np.ascontiguousarray(B2.T)
    _check(lib().cmoe_potrs(Lc.shape[0], B2.shape[1], _d(Lc), _d(buf), int(device)))
    return buf.T.copy().reshape(B.shape)


def philox_normals(seed, first_draw, num_draws, per_draw, device=0):
    out = np.empty(num_draws * per_draw)
    _check(lib().cmoe_philox_normals(ctypes.c_uint64(seed), ctypes.c_uint64(first_draw), int(num_draws),
                                     int(per_draw), _d(out), int(device)))
    return out.reshape(num_draws, per_draw)


class GaussianProcess:
    """Device-resident GP (reference: optimal_learning::GaussianProcess, gpp_math.hpp:275)."""

    def __init__(self, kernel, alpha, lengths, X, y, noise, derivs=None, device=0):
        X = _f64(X)
        self.N, self.dim = X.shape
        self.derivs = _i32(derivs)
        self.g = len(self.derivs)
        self.n = self.N * (1 + self.g)
        y = _f64(y).ravel()
        noise = _f64(noise).ravel()
        lengths = _f64(lengths).ravel()
        assert y.size == self.n and noise.size == 1 + self.g and lengths.size == self.dim
        self.kernel = kernel
        self.device = device
        h = ctypes.c_void_p()
        info = ctypes.c_int(0)
        rc = lib().cmoe_gp_create(int(kernel), ctypes.c_double(alpha), _d(lengths), _d(X), _d(y), _d(noise),
                                  _i(self.derivs), self.g, self.dim, self.N, int(device), ctypes.byref(h),
                                  ctypes.byref(info))
        self.h = None
        _check(rc, info.value)
        self.h = h

    def __del__(self):
        try:
            if self.h:
                lib().cmoe_gp_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def state(self):
        K = np.empty(self.n * self.n)
        kinvy = np.empty(self.n)
        mean = ctypes.c_double()
        _check(lib().cmoe_gp_get_state(self.h, _d(K), _d(kinvy), ctypes.byref(mean)))
        return K.reshape(self.n, self.n).T.copy(), kinvy, mean.value

    def fit_timings_usec(self):
        t = np.zeros(3)
        lib().cmoe_gp_fit_timings(self.h, _d(t))
        return t

    def add_sampled_points(self, pts, vals):
        pts = _f64(pts).reshape(-1, self.dim)
        vals = _f64(vals).ravel()
        info = ctypes.c_int(0)
        rc = lib().cmoe_gp_add_sampled_points(self.h, _d(pts), _d(vals), pts.shape[0], ctypes.byref(info))
        _check(rc, info.value)
        self.N += pts.shape[0]
        self.n = self.N * (1 + self.g)

    def bench_cov_build(self, repeats=10):
        t = ctypes.c_double()
        _check(lib().cmoe_bench_cov_build(self.h, int(repeats), ctypes.byref(t)))
        return t.value

    def bench_cholesky(self, repeats=3):
        t = ctypes.c_double()
        _check(lib().cmoe_bench_cholesky(self.h, int(repeats), ctypes.byref(t)))
        return t.value

    def posterior(self, sets, derivs_s=None, want=("mean", "var")):
        """sets: [num_sets, num_pts, dim].  Returns dict of arrays with a leading num_sets axis (reference layouts)."""
        sets = _f64(sets)
        if sets.ndim == 2:
            sets = sets[None]
        ns, num, _ = sets.shape
        ds = _i32(derivs_s)
        Q = num * (1 + len(ds))
        shapes = {"mean": (ns, Q), "grad_mean": (ns, Q * self.dim), "var": (ns, Q * Q), "chol_var": (ns, Q * Q),
                  "grad_var": (ns, num * Q * Q * self.dim), "grad_chol": (ns, num * Q * Q * self.dim)}
        bufs = {k: (np.empty(shapes[k]) if k in want else None) for k in shapes}
        info = ctypes.c_int(0)
        rc = lib().cmoe_gp_posterior(self.h, _d(sets), ns, num, _i(ds), len(ds), _d(bufs["mean"]),
                                     _d(bufs["grad_mean"]), _d(bufs["var"]), _d(bufs["chol_var"]),
                                     _d(bufs["grad_var"]), _d(bufs["grad_chol"]), ctypes.byref(info))
        _check(rc, info.value)
        return {k: v for k, v in bufs.items() if v is not None}

    def ei(self, candidates, Xp, num_mc, best_so_far, seed=0, table=None, grad=False):
        cand = _f64(candidates)
        if cand.ndim == 2:
            cand = cand[None]
        nc, q, _ = cand.shape
        Xp = _f64(Xp).reshape(-1, self.dim) if Xp is not None and len(Xp) else np.zeros((0, self.dim))
        table = _f64(table).ravel() if table is not None else None
        if table is not None:
            assert table.size >= num_mc * (q + Xp.shape[0])
        ei = np.empty(nc)
        g = np.empty((nc, q, self.dim)) if grad else None
        info = ctypes.c_int(0)
        rc = lib().cmoe_ei_eval(self.h, _d(cand), nc, q, _d(Xp), Xp.shape[0], int(num_mc),
                                ctypes.c_double(best_so_far), ctypes.c_uint64(seed), _d(table), _d(ei), _d(g),
                                ctypes.byref(info))
        _check(rc, info.value)
        return (ei, g) if grad else ei

    def kg(self, candidates, Xp, num_mc, best_so_far, inner, inner_bounds, discrete_pts, num_fidelity=0, seed=0,
           table=None, grad=False, stats=False):
        cand = _f64(candidates)
        if cand.ndim == 2:
            cand = cand[None]
        nc, q, _ = cand.shape
        Xp = _f64(Xp).reshape(-1, self.dim) if Xp is not None and len(Xp) else np.zeros((0, self.dim))
        table = _f64(table).ravel() if table is not None else None
        inner = GDParams.from_seq(inner)
        inner_bounds = _f64(inner_bounds).ravel()
        disc = _f64(discrete_pts).reshape(-1, self.dim - num_fidelity)
        kg = np.empty(nc)
        g = np.empty((nc, q, self.dim)) if grad else None
        st = KGStats()
        info = ctypes.c_int(0)
        rc = lib().cmoe_kg_eval(self.h, int(num_fidelity), ctypes.byref(inner), _d(inner_bounds), _d(disc),
                                disc.shape[0], _d(cand), nc, q, _d(Xp), Xp.shape[0], int(num_mc),
                                ctypes.c_double(best_so_far), ctypes.c_uint64(seed), _d(table), _d(kg), _d(g),
                                ctypes.byref(st), ctypes.byref(info))
        _check(rc, info.value)
        res = [kg]
        if grad:
            res.append(g)
        if stats:
            res.append({"mc_samples": st.mc_samples, "posterior_evals": st.posterior_evals,
                        "line_search_steps": st.line_search_steps, "point_evals": st.point_evals,
                        "line_batches": st.line_batches})
        return res[0] if len(res) == 1 else tuple(res)


def log_marginal_likelihood(kernel, alpha, lengths, X, y, noise, derivs=None, device=0):
    """cmoe_log_marginal_likelihood (gpp_model_selection.cpp:540-612): one device fit + two host reductions."""
    X = _f64(X)
    N, dim = X.shape
    derivs = _i32(derivs)
    val = ctypes.c_double()
    info = ctypes.c_int()
    rc = lib().cmoe_log_marginal_likelihood(int(kernel), ctypes.c_double(alpha), _d(_f64(lengths).ravel()), _d(X),
                                            _d(_f64(y).ravel()), _d(_f64(noise).ravel()), _i(derivs), len(derivs), dim, N,
                                            int(device), ctypes.byref(val), ctypes.byref(info))
    _check(rc, info.value)
    return val.value


def grad_log_marginal_likelihood(kernel, alpha, lengths, X, y, noise, derivs=None, device=0):
    """cmoe_grad_log_marginal_likelihood: d log p / d (alpha, lengths..., noise by observation type)."""
    X = _f64(X)
    N, dim = X.shape
    derivs = _i32(derivs if derivs is not None else [])
    grad = np.zeros(dim + 2 + derivs.size)
    info = ctypes.c_int(0)
    rc = lib().cmoe_grad_log_marginal_likelihood(int(kernel), ctypes.c_double(alpha), _d(_f64(lengths).ravel()), _d(X),
                                                 _d(_f64(y).ravel()), _d(_f64(noise).ravel()), _i(derivs), derivs.size,
                                                 dim, N, int(device), _d(grad), ctypes.byref(info))
    _check(rc, info.value)
    return grad


def fp64_peaks(device=0):
    """Measured FP64 peaks (TFLOP/s): (DFMA vector pipe, DMMA tensor pipe)."""
    t = np.zeros(2)
    _check(lib().cmoe_bench_fp64_peaks(int(device), _d(t)))
    return float(t[0]), float(t[1])


def fp64_mixed(device=0):
    """cmoe_bench_fp64_mixed: total TFLOP/s with DFMA and DMMA interleaved in every warp."""
    out = np.zeros(1)
    _check(lib().cmoe_bench_fp64_mixed(int(device), _d(out)))
    return float(out[0])


class KGPlan:
    """Device-resident q-KG evaluation plan (cmoe_kg_plan_*): create -> upload -> run -> sync -> download."""

    def __init__(self, gp, num_mc, best_so_far, inner, inner_bounds, discrete_pts, max_candidates, q, Xp=None,
                 num_fidelity=0, seed=0, want_grad=True, table=None):
        self.gp = gp
        self.q = q
        self.dim = gp.dim
        self.want_grad = want_grad
        Xp = _f64(Xp).reshape(-1, gp.dim) if Xp is not None and len(Xp) else np.zeros((0, gp.dim))
        inner = GDParams.from_seq(inner)
        inner_bounds = _f64(inner_bounds).ravel()
        disc = _f64(discrete_pts).reshape(-1, gp.dim - num_fidelity)
        h = ctypes.c_void_p()
        rc = lib().cmoe_kg_plan_create(gp.h, int(num_fidelity), ctypes.byref(inner), _d(inner_bounds), _d(disc),
                                       disc.shape[0], int(max_candidates), int(q), _d(Xp), Xp.shape[0], int(num_mc),
                                       ctypes.c_double(best_so_far), ctypes.c_uint64(seed), int(bool(want_grad)),
                                       ctypes.byref(h))
        self.h = None
        _check(rc)
        self.h = h
        self.nc = 0
        if table is not None:
            table = _f64(table).ravel()
            _check(lib().cmoe_kg_plan_set_table(self.h, _d(table), table.size))

    def __del__(self):
        try:
            if self.h:
                lib().cmoe_kg_plan_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_stale_union(self, points):
        """Reference-driver compatibility (cmoe_kg_plan_set_stale_union): the discretisation set keeps these q points."""
        pts = _f64(points).reshape(self.q, self.dim) if points is not None else None
        _check(lib().cmoe_kg_plan_set_stale_union(self.h, _d(pts)))

    def upload(self, candidates):
        cand = _f64(candidates).reshape(-1, self.q, self.dim)
        self.nc = cand.shape[0]
        _check(lib().cmoe_kg_plan_upload(self.h, _d(cand), self.nc))

    def run(self):
        _check(lib().cmoe_kg_plan_run(self.h))

    def sync(self):
        info = ctypes.c_int(0)
        _check(lib().cmoe_kg_plan_sync(self.h, ctypes.byref(info)), info.value)

    def download(self):
        kg = np.empty(self.nc)
        g = np.empty((self.nc, self.q, self.dim)) if self.want_grad else None
        st = KGStats()
        _check(lib().cmoe_kg_plan_download(self.h, _d(kg), _d(g), ctypes.byref(st)))
        return kg, g, {"mc_samples": st.mc_samples, "posterior_evals": st.posterior_evals,
                       "line_search_steps": st.line_search_steps, "point_evals": st.point_evals,
                       "line_batches": st.line_batches}

    def timings(self):
        total, mc = ctypes.c_double(), ctypes.c_double()
        launches = ctypes.c_int()
        _check(lib().cmoe_kg_plan_timings(self.h, ctypes.byref(total), ctypes.byref(mc), ctypes.byref(launches)))
        return total.value, mc.value, launches.value


def _gd_args(gp, outer, inner, domain_bounds, inner_bounds, discrete_pts, num_fidelity):
    return (GDParams.from_seq(outer), GDParams.from_seq(inner) if inner is not None else None,
            _f64(domain_bounds).ravel(), _f64(inner_bounds).ravel() if inner_bounds is not None else None,
            _f64(discrete_pts).reshape(-1, gp.dim - num_fidelity) if discrete_pts is not None else None)


class MultistartOpts(ctypes.Structure):
    _fields_ = [("normals_table", ctypes.POINTER(ctypes.c_double)), ("table_len", ctypes.c_size_t),
                ("devices", ctypes.POINTER(ctypes.c_int)), ("num_devices", ctypes.c_int),
                ("domain_type", ctypes.c_int), ("fresh_discretisation", ctypes.c_int),
                ("stale_union", ctypes.POINTER(ctypes.c_double))]


TENSOR_PRODUCT, SIMPLEX = 0, 1


def limit_update(domain_type, bounds, mrc, x, upd):
    """cmoe_limit_update: one LimitUpdate of the outer optimiser's domain (host arithmetic)."""
    x = _f64(x).ravel()
    upd = _f64(upd).ravel().copy()
    _check(lib().cmoe_limit_update(int(domain_type), _d(_f64(bounds).ravel()), x.size, ctypes.c_double(mrc), _d(x),
                                   _d(upd)))
    return upd


def _multistart_opts(table, devices, domain_type=0, fresh=False, stale=None):
    """(opts struct or None, keep-alive tuple) for cmoe_multistart_{kg,ei}_ex."""
    if table is None and not devices and not domain_type and not fresh and stale is None:
        return None, ()
    t = _f64(table).ravel() if table is not None else None
    dv = _i32(devices) if devices else None
    o = MultistartOpts(_d(t) if t is not None else None, t.size if t is not None else 0,
                       _i(dv) if dv is not None else None, dv.size if dv is not None else 0, int(domain_type),
                       int(bool(fresh)), None)
    st = None
    if stale is not None:
        st = _f64(stale).ravel()
        o.stale_union = _d(st)
    return o, (t, dv, st)


def multistart_kg(gp, starts, Xp, num_mc, best_so_far, outer, inner, domain_bounds, inner_bounds, discrete_pts,
                  num_fidelity=0, seed=0, table=None, devices=None, fresh_discretisation=False, stale_union=None):
    """cmoe_multistart_kg(_ex): returns (best_point [q, dim], best_value, found_flag, start_values).
    table: normals replayed by every evaluation (NormalRNGSimulator semantics); devices: GPUs to shard the starts over."""
    starts = _f64(starts)
    ns, q, dim = starts.shape
    Xp = _f64(Xp).reshape(-1, dim) if Xp is not None and len(Xp) else np.zeros((0, dim))
    outer, inner, db, ib, disc = _gd_args(gp, outer, inner, domain_bounds, inner_bounds, discrete_pts, num_fidelity)
    vals = np.empty(ns)
    best = np.empty((q, dim))
    bv = ctypes.c_double()
    found = ctypes.c_int()
    info = ctypes.c_int()
    opts, keep = _multistart_opts(table, devices, 0, fresh_discretisation, stale_union)
    rc = lib().cmoe_multistart_kg_ex(gp.h, int(num_fidelity), ctypes.byref(outer), ctypes.byref(inner), _d(db), _d(ib),
                                     _d(disc), disc.shape[0], _d(starts), ns, q, _d(Xp), Xp.shape[0], int(num_mc),
                                     ctypes.c_double(best_so_far), ctypes.c_uint64(seed),
                                     ctypes.byref(opts) if opts is not None else None, _d(vals), _d(best),
                                     ctypes.byref(bv), ctypes.byref(found), ctypes.byref(info))
    del keep
    _check(rc, info.value)
    return best, bv.value, bool(found.value), vals


def multistart_ei(gp, starts, Xp, num_mc, best_so_far, outer, domain_bounds, seed=0, table=None, devices=None,
                  domain_type=0):
    starts = _f64(starts)
    ns, q, dim = starts.shape
    Xp = _f64(Xp).reshape(-1, dim) if Xp is not None and len(Xp) else np.zeros((0, dim))
    outer = GDParams.from_seq(outer)
    db = _f64(domain_bounds).ravel()
    vals = np.empty(ns)
    best = np.empty((q, dim))
    bv = ctypes.c_double()
    found = ctypes.c_int()
    info = ctypes.c_int()
    opts, keep = _multistart_opts(table, devices, domain_type)
    rc = lib().cmoe_multistart_ei_ex(gp.h, ctypes.byref(outer), _d(db), _d(starts), ns, q, _d(Xp), Xp.shape[0],
                                     int(num_mc), ctypes.c_double(best_so_far), ctypes.c_uint64(seed),
                                     ctypes.byref(opts) if opts is not None else None, _d(vals), _d(best),
                                     ctypes.byref(bv), ctypes.byref(found), ctypes.byref(info))
    del keep
    _check(rc, info.value)
    return best, bv.value, bool(found.value), vals


def kg_gradient_descent(gp, starts, Xp, num_mc, best_so_far, outer, inner, domain_bounds, inner_bounds, discrete_pts,
                        num_fidelity=0, seed=0, stale_union=None):
    starts = _f64(starts)
    ns, q, dim = starts.shape
    Xp = _f64(Xp).reshape(-1, dim) if Xp is not None and len(Xp) else np.zeros((0, dim))
    outer, inner, db, ib, disc = _gd_args(gp, outer, inner, domain_bounds, inner_bounds, discrete_pts, num_fidelity)
    vals = np.empty(ns)
    pts = np.empty((ns, q, dim))
    info = ctypes.c_int()
    st = _f64(stale_union).ravel() if stale_union is not None else None
    rc = lib().cmoe_kg_gradient_descent_ex(gp.h, int(num_fidelity), ctypes.byref(outer), ctypes.byref(inner), _d(db),
                                           _d(ib), _d(disc), disc.shape[0], _d(starts), ns, q, _d(Xp), Xp.shape[0],
                                           int(num_mc), ctypes.c_double(best_so_far), ctypes.c_uint64(seed), _d(st),
                                           _d(vals), _d(pts), ctypes.byref(info))
    _check(rc, info.value)
    return vals, pts


def ei_gradient_descent(gp, starts, Xp, num_mc, best_so_far, outer, domain_bounds, seed=0):
    starts = _f64(starts)
    ns, q, dim = starts.shape
    Xp = _f64(Xp).reshape(-1, dim) if Xp is not None and len(Xp) else np.zeros((0, dim))
    outer = GDParams.from_seq(outer)
    db = _f64(domain_bounds).ravel()
    vals = np.empty(ns)
    pts = np.empty((ns, q, dim))
    info = ctypes.c_int()
    rc = lib().cmoe_ei_gradient_descent(gp.h, ctypes.byref(outer), _d(db), _d(starts), ns, q, _d(Xp), Xp.shape[0],
                                        int(num_mc), ctypes.c_double(best_so_far), ctypes.c_uint64(seed), _d(vals),
                                        _d(pts), ctypes.byref(info))
    _check(rc, info.value)
    return vals, pts


def posterior_mean_optimization(gp, initial_guess, params, domain_bounds, num_fidelity=0):
    """cmoe_posterior_mean_optimization: returns (best_point [dim - num_fidelity], -min mu, found)."""
    ps = gp.dim - num_fidelity
    x0 = _f64(initial_guess).ravel()
    assert x0.size == ps
    params = GDParams.from_seq(params)
    db = _f64(domain_bounds).ravel()
    best = np.zeros(ps)
    val = ctypes.c_double()
    found = ctypes.c_int()
    _check(lib().cmoe_posterior_mean_optimization(gp.h, int(num_fidelity), ctypes.byref(params), _d(db), _d(x0),
                                                  _d(best), ctypes.byref(val), ctypes.byref(found)))
    return best, val.value, bool(found.value)


def ei_analytic(gp, points, best_so_far, grad=False):
    """cmoe_ei_analytic: closed-form 1-EI (and gradient) at points [n, dim]."""
    pts = _f64(points).reshape(-1, gp.dim)
    vals = np.empty(pts.shape[0])
    g = np.empty_like(pts) if grad else None
    info = ctypes.c_int()
    _check(lib().cmoe_ei_analytic(gp.h, _d(pts), pts.shape[0], ctypes.c_double(best_so_far), _d(vals), _d(g),
                                  ctypes.byref(info)), info.value)
    return (vals, g) if grad else vals


class GaussianProcessEnsemble:
    """An array of GP handles over the same data, one per hyper-parameter sample — the reference's GaussianProcessMCMC
    (gpp_knowledge_gradient_mcmc_optimization.cpp:24-48).  hypers[M][1+dim] = (alpha, lengths); noises[M][1+g]."""

    def __init__(self, hypers, noises, X, y, derivs=None, kernel=MATERN_NU_2P5, device=0):
        hypers, noises = _f64(hypers), _f64(noises)
        self.members = [GaussianProcess(kernel, hypers[m, 0], hypers[m, 1:], X, y, noises[m], derivs, device)
                        for m in range(hypers.shape[0])]
        self.dim = self.members[0].dim

    def __len__(self):
        return len(self.members)

    def _handles(self):
        return (ctypes.c_void_p * len(self.members))(*[gp.h for gp in self.members])

    def _cands(self, candidates, Xp):
        cand = _f64(candidates)
        if cand.ndim == 2:
            cand = cand[None]
        Xp = _f64(Xp).reshape(-1, self.dim) if Xp is not None and len(Xp) else np.zeros((0, self.dim))
        return cand, Xp

    def kg(self, candidates, Xp, num_mc, best_so_far, inner, inner_bounds, discrete_pts, num_fidelity=0, seed=0,
           table=None, grad=False):
        """cmoe_kg_eval_mcmc.  discrete_pts [M, num_pts, dim - nf]; best_so_far [M]."""
        cand, Xp = self._cands(candidates, Xp)
        nc, q, _ = cand.shape
        M = len(self.members)
        disc = _f64(discrete_pts).reshape(M, -1, self.dim - num_fidelity)
        best = _f64(best_so_far).ravel()
        assert best.size == M
        table = _f64(table).ravel() if table is not None else None
        inner = GDParams.from_seq(inner)
        ib = _f64(inner_bounds).ravel()
        vals = np.empty(nc)
        g = np.empty((nc, q, self.dim)) if grad else None
        info = ctypes.c_int()
        rc = lib().cmoe_kg_eval_mcmc(self._handles(), M, int(num_fidelity), ctypes.byref(inner), _d(ib), _d(disc),
                                     disc.shape[1], _d(cand), nc, q, _d(Xp), Xp.shape[0], int(num_mc), _d(best),
                                     ctypes.c_uint64(seed), _d(table), _d(vals), _d(g), ctypes.byref(info))
        _check(rc, info.value)
        return (vals, g) if grad else vals

    def ei(self, candidates, Xp, num_mc, best_so_far, seed=0, table=None, grad=False, analytic_single=False):
        """cmoe_ei_eval_mcmc."""
        cand, Xp = self._cands(candidates, Xp)
        nc, q, _ = cand.shape
        M = len(self.members)
        best = _f64(best_so_far).ravel()
        assert best.size == M
        table = _f64(table).ravel() if table is not None else None
        vals = np.empty(nc)
        g = np.empty((nc, q, self.dim)) if grad else None
        info = ctypes.c_int()
        rc = lib().cmoe_ei_eval_mcmc(self._handles(), M, _d(cand), nc, q, _d(Xp), Xp.shape[0], int(num_mc), _d(best),
                                     ctypes.c_uint64(seed), _d(table), int(bool(analytic_single)), _d(vals), _d(g),
                                     ctypes.byref(info))
        _check(rc, info.value)
        return (vals, g) if grad else vals

    def multistart_kg(self, starts, Xp, num_mc, best_so_far, outer, inner, domain_bounds, inner_bounds, discrete_pts,
                      num_fidelity=0, seed=0):
        """cmoe_multistart_kg_mcmc: returns (best_point [q, dim], best_value, found_flag, start_values)."""
        starts = _f64(starts)
        ns, q, dim = starts.shape
        _, Xp = self._cands(starts, Xp)
        M = len(self.members)
        disc = _f64(discrete_pts).reshape(M, -1, self.dim - num_fidelity)
        best_in = _f64(best_so_far).ravel()
        outer, inner = GDParams.from_seq(outer), GDParams.from_seq(inner)
        db, ib = _f64(domain_bounds).ravel(), _f64(inner_bounds).ravel()
        vals, best = np.empty(ns), np.empty((q, dim))
        bv, found, info = ctypes.c_double(), ctypes.c_int(), ctypes.c_int()
        rc = lib().cmoe_multistart_kg_mcmc(self._handles(), M, int(num_fidelity), ctypes.byref(outer),
                                           ctypes.byref(inner), _d(db), _d(ib), _d(disc), disc.shape[1], _d(starts), ns,
                                           q, _d(Xp), Xp.shape[0], int(num_mc), _d(best_in), ctypes.c_uint64(seed),
                                           _d(vals), _d(best), ctypes.byref(bv), ctypes.byref(found),
                                           ctypes.byref(info))
        _check(rc, info.value)
        return best, bv.value, bool(found.value), vals

    def multistart_ei(self, starts, Xp, num_mc, best_so_far, outer, domain_bounds, seed=0):
        """cmoe_multistart_ei_mcmc: returns (best_point [q, dim], best_value, found_flag, start_values)."""
        starts = _f64(starts)
        ns, q, dim = starts.shape
        _, Xp = self._cands(starts, Xp)
        M = len(self.members)
        best_in = _f64(best_so_far).ravel()
        outer = GDParams.from_seq(outer)
        db = _f64(domain_bounds).ravel()
        vals, best = np.empty(ns), np.empty((q, dim))
        bv, found, info = ctypes.c_double(), ctypes.c_int(), ctypes.c_int()
        rc = lib().cmoe_multistart_ei_mcmc(self._handles(), M, ctypes.byref(outer), _d(db), _d(starts), ns, q, _d(Xp),
                                           Xp.shape[0], int(num_mc), _d(best_in), ctypes.c_uint64(seed), _d(vals),
                                           _d(best), ctypes.byref(bv), ctypes.byref(found), ctypes.byref(info))
        _check(rc, info.value)
        return best, bv.value, bool(found.value), vals

    def kg_gradient_descent(self, starts, Xp, num_mc, best_so_far, outer, inner, domain_bounds, inner_bounds,
                            discrete_pts, num_fidelity=0, seed=0):
        """cmoe_kg_gradient_descent_mcmc: returns (values [ns], points [ns, q, dim])."""
        starts = _f64(starts)
        ns, q, dim = starts.shape
        _, Xp = self._cands(starts, Xp)
        M = len(self.members)
        disc = _f64(discrete_pts).reshape(M, -1, self.dim - num_fidelity)
        best_in = _f64(best_so_far).ravel()
        outer, inner = GDParams.from_seq(outer), GDParams.from_seq(inner)
        db, ib = _f64(domain_bounds).ravel(), _f64(inner_bounds).ravel()
        vals, pts = np.empty(ns), np.empty((ns, q, dim))
        info = ctypes.c_int()
        rc = lib().cmoe_kg_gradient_descent_mcmc(self._handles(), M, int(num_fidelity), ctypes.byref(outer),
                                                 ctypes.byref(inner), _d(db), _d(ib), _d(disc), disc.shape[1],
                                                 _d(starts), ns, q, _d(Xp), Xp.shape[0], int(num_mc), _d(best_in),
                                                 ctypes.c_uint64(seed), _d(vals), _d(pts), ctypes.byref(info))
        _check(rc, info.value)
        return vals, pts

    def ei_gradient_descent(self, starts, Xp, num_mc, best_so_far, outer, domain_bounds, seed=0):
        """cmoe_ei_gradient_descent_mcmc: returns (values [ns], points [ns, q, dim])."""
        starts = _f64(starts)
        ns, q, dim = starts.shape
        _, Xp = self._cands(starts, Xp)
        M = len(self.members)
        best_in = _f64(best_so_far).ravel()
        outer = GDParams.from_seq(outer)
        db = _f64(domain_bounds).ravel()
        vals, pts = np.empty(ns), np.empty((ns, q, dim))
        info = ctypes.c_int()
        rc = lib().cmoe_ei_gradient_descent_mcmc(self._handles(), M, ctypes.byref(outer), _d(db), _d(starts), ns, q,
                                                 _d(Xp), Xp.shape[0], int(num_mc), _d(best_in), ctypes.c_uint64(seed),
                                                 _d(vals), _d(pts), ctypes.byref(info))
        _check(rc, info.value)
        return vals, pts
