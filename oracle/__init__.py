"""TEST INFRASTRUCTURE — ctypes bindings for the CPU checkers.

Two interchangeable back ends with the same method surface:

* ``load_oracle()``    -> oracle/libmoe_oracle.so  (plain-C restatement, always buildable; ``make -C oracle oracle``)
* ``load_reference()`` -> oracle/_ref/libmoe_ref.so (the unmodified Cornell-MOE C++ core; ``make -C oracle ref``,
  only buildable where /root/reference exists, but the built .so travels to the GPU box)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package.  The product (``cornell_moe_b200``) never does.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    if a is None:
        return np.zeros(0, dtype=np.int32)
    return np.ascontiguousarray(a, dtype=np.int32)


class CpuGP:
    """Handle on a CPU GaussianProcess (oracle or reference)."""

    def __init__(self, backend, handle, kernel, dim, N, derivs):
        self.b = backend
        self.h = handle
        self.kernel = kernel
        self.dim = dim
        self.N = N
        self.derivs = derivs
        self.g = len(derivs)
        self.n = N * (1 + self.g)

    def __del__(self):
        try:
            if self.h:
                self.b._fn("gp_destroy")(self.h)
                self.h = None
        except Exception:
            pass

    def state(self):
        K = np.empty(self.n * self.n)
        kinvy = np.empty(self.n)
        mean = ctypes.c_double()
        self.b._fn("gp_get_state")(self.h, _d(K), _d(kinvy), ctypes.byref(mean))
        return K.reshape(self.n, self.n).T.copy(), kinvy, mean.value  # K as a proper [row, col] array

    def posterior(self, pts, derivs_s=None, want=("mean", "var")):
        """Returns dict with any of mean, grad_mean, var, chol_var, grad_var, grad_chol (raw reference layouts)."""
        pts = _f64(pts).reshape(-1, self.dim)
        num = pts.shape[0]
        ds = _i32(derivs_s)
        Q = num * (1 + len(ds))
        out = {}
        bufs = {
            "mean": np.empty(Q),
            "grad_mean": np.empty(self.dim * Q),
            "var": np.empty(Q * Q),
            "chol_var": np.empty(Q * Q),
            "grad_var": np.empty(self.dim * Q * Q * num),
            "grad_chol": np.empty(self.dim * Q * Q * num),
        }
        args = [(_d(bufs[k]) if k in want else None) for k in ("mean", "grad_mean", "var", "chol_var", "grad_var", "grad_chol")]
        f = self.b._fn("gp_posterior")
        f.restype = ctypes.c_int
        rc = f(self.h, _d(pts), num, _i(ds), len(ds), *args)
        out["rc"] = rc
        for k in want:
            out[k] = bufs[k]
        return out

    def mean_additional(self, pts):
        pts = _f64(pts).reshape(-1, self.dim)
        out = np.empty(pts.shape[0])
        self.b._fn("gp_mean_additional")(self.h, _d(pts), pts.shape[0], _d(out))
        return out

    def posterior_mean_optimization(self, initial_guess, gd, bounds, num_fidelity=0):
        x0 = _f64(initial_guess).ravel()
        gd = _f64(gd)
        bounds = _f64(bounds).ravel()
        best = np.zeros(self.dim - num_fidelity)
        f = self.b._fn("posterior_mean_optimization")
        f.restype = ctypes.c_double
        v = f(self.h, int(num_fidelity), _d(gd), _d(bounds), _d(x0), _d(best))
        return best, v

    def ei(self, Xq, Xp, num_mc, best_so_far, table, grad=False):
        Xq = _f64(Xq).reshape(-1, self.dim)
        Xp = _f64(Xp).reshape(-1, self.dim) if Xp is not None and len(Xp) else np.zeros((0, self.dim))
        table = _f64(table).ravel()
        g = np.empty(Xq.size) if grad else None
        f = self.b._fn("ei")
        f.restype = ctypes.c_double
        v = f(self.h, _d(Xq), _d(Xp), Xq.shape[0], Xp.shape[0], int(num_mc), ctypes.c_double(best_so_far), _d(table),
              table.size, _d(g))
        return (v, g.reshape(Xq.shape)) if grad else v

    def kg(self, Xq, Xp, num_mc, best_so_far, table, gd, inner_bounds, discrete_pts, num_fidelity=0, grad=False,
           want_best_points=False):
        Xq = _f64(Xq).reshape(-1, self.dim)
        Xp = _f64(Xp).reshape(-1, self.dim) if Xp is not None and len(Xp) else np.zeros((0, self.dim))
        table = _f64(table).ravel()
        gd = _f64(gd)
        inner_bounds = _f64(inner_bounds).ravel()
        discrete_pts = _f64(discrete_pts).reshape(-1, self.dim - num_fidelity)
        g = np.empty(Xq.size) if grad else None
        bp = np.empty(num_mc * self.dim) if want_best_points else None
        f = self.b._fn("kg")
        f.restype = ctypes.c_double
        v = f(self.h, int(num_fidelity), _d(gd), _d(inner_bounds), _d(discrete_pts), discrete_pts.shape[0], _d(Xq),
              _d(Xp), Xq.shape[0], Xp.shape[0], int(num_mc), ctypes.c_double(best_so_far), _d(table), table.size,
              _d(g), _d(bp))
        res = [v]
        if grad:
            res.append(g.reshape(Xq.shape))
        if want_best_points:
            res.append(bp.reshape(num_mc, self.dim))
        return res[0] if len(res) == 1 else tuple(res)


    def kg_reused_state(self, Xfirst, Xq, Xp, num_mc, best_so_far, table, gd, inner_bounds, discrete_pts, num_fidelity=0,
                        grad=False):
        """q-KG at Xq through a state that was constructed with Xfirst (reference back end only): what one evaluation
        inside the reference's multistart drivers computes — the inner optimiser's start set keeps Xfirst."""
        Xq = _f64(Xq).reshape(-1, self.dim)
        Xf = _f64(Xfirst).reshape(-1, self.dim)
        Xp = _f64(Xp).reshape(-1, self.dim) if Xp is not None and len(Xp) else np.zeros((0, self.dim))
        table = _f64(table).ravel()
        discrete_pts = _f64(discrete_pts).reshape(-1, self.dim - num_fidelity)
        g = np.empty(Xq.size) if grad else None
        f = self.b._fn("kg_reused_state")
        f.restype = ctypes.c_double
        v = f(self.h, int(num_fidelity), _d(_f64(gd)), _d(_f64(inner_bounds).ravel()), _d(discrete_pts),
              discrete_pts.shape[0], _d(Xf), _d(Xq), _d(Xp), Xq.shape[0], Xp.shape[0], int(num_mc),
              ctypes.c_double(best_so_far), _d(table), table.size, _d(g))
        return (v, g.reshape(Xq.shape)) if grad else v


class CpuBackend:
    def __init__(self, path, prefix, name):
        self.lib = ctypes.CDLL(path)
        self.prefix = prefix
        self.name = name

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def covariance(self, kernel, alpha, lengths, p1, d1, p2, d2, grad=False):
        lengths = _f64(lengths)
        p1, p2 = _f64(p1), _f64(p2)
        d1, d2 = _i32(d1), _i32(d2)
        dim = p1.size
        size = (1 + len(d1)) * (1 + len(d2)) * (dim if grad else 1)
        out = np.empty(size)
        self._fn("grad_covariance" if grad else "covariance")(
            int(kernel), dim, ctypes.c_double(alpha), _d(lengths), _d(p1), _i(d1), len(d1), _d(p2), _i(d2), len(d2),
            _d(out))
        return out

    def cholesky(self, A):
        """A: [n, n] symmetric (row/col array). Returns (rc, L as [row, col] with the reference's untouched upper part)."""
        A = np.array(A, dtype=np.float64)
        n = A.shape[0]
        buf = np.ascontiguousarray(A.T)  # column-major buffer
        f = self._fn("cholesky")
        f.restype = ctypes.c_int
        rc = f(n, _d(buf))
        return rc, buf.T.copy()

    def potrs(self, L, B):
        L = np.ascontiguousarray(np.array(L, dtype=np.float64).T)
        B = np.array(B, dtype=np.float64)
        B2 = B.reshape(B.shape[0], -1)
        buf = np.ascontiguousarray(B2.T)
        self._fn("potrs")(_d(L), L.shape[0], B2.shape[1], _d(buf))
        return buf.T.copy().reshape(B.shape)

    def limit_update(self, bounds, mrc, x, upd):
        bounds = _f64(bounds).ravel()
        x = _f64(x)
        upd = np.array(upd, dtype=np.float64)
        self._fn("limit_update")(_d(bounds), x.size, ctypes.c_double(mrc), _d(x), _d(upd))
        return upd

    def gp(self, kernel, alpha, lengths, X, y, noise, derivs=None):
        X = _f64(X)
        N, dim = X.shape
        y = _f64(y).ravel()
        derivs = _i32(derivs)
        noise = _f64(noise).ravel()
        lengths = _f64(lengths)
        assert y.size == N * (1 + len(derivs)) and noise.size == 1 + len(derivs) and lengths.size == dim
        lm = ctypes.c_int(0)
        f = self._fn("gp_create")
        f.restype = ctypes.c_void_p
        h = f(int(kernel), ctypes.c_double(alpha), _d(lengths), _d(X), _d(y), _d(noise), _i(derivs), len(derivs), dim,
              N, ctypes.byref(lm))
        if not h:
            return None, lm.value
        return CpuGP(self, ctypes.c_void_p(h), kernel, dim, N, derivs), 0

    # ---- MCMC-averaged acquisition over an ensemble of Matern-5/2 GPs (hypers[M][1+dim], noises[M][1+g]) ----
    def _ensemble(self, hypers, noises, X, y, derivs):
        hypers = _f64(hypers)
        noises = _f64(noises)
        gps = []
        for m in range(hypers.shape[0]):
            gp, lm = self.gp(1, hypers[m, 0], hypers[m, 1:], X, y, noises[m], derivs)
            assert gp is not None, f"ensemble member {m} singular (leading minor {lm})"
            gps.append(gp)
        return gps

    def kg_mcmc(self, hypers, noises, X, y, derivs, Xq, Xp, num_mc, best_so_far, table, gd, inner_bounds,
                discrete_pts, num_fidelity=0, grad=False):
        hypers, noises, X, y = _f64(hypers), _f64(noises), _f64(X), _f64(y)
        derivs = _i32(derivs if derivs is not None else [])
        dim = X.shape[1]
        Xq = _f64(Xq).reshape(-1, dim)
        Xp = _f64(Xp if Xp is not None else np.zeros((0, dim)))
        q, p = Xq.shape[0], (Xp.size // dim)
        best = _f64(best_so_far).ravel()
        table = _f64(table).ravel()
        gd = _f64(gd)
        ib = _f64(inner_bounds).ravel()
        D = _f64(discrete_pts)
        M = hypers.shape[0]
        num_pts = D.size // (M * (dim - num_fidelity))
        g = np.zeros((q, dim)) if grad else None
        gptr = _d(g) if grad else None
        if self.prefix == "ref_":
            f = self._fn("kg_mcmc")
            f.restype = ctypes.c_double
            v = f(_d(hypers), _d(noises), M, _d(X), _d(y), _i(derivs), derivs.size, dim, X.shape[0],
                  int(num_fidelity), _d(gd), _d(ib), _d(D), num_pts, _d(Xq), _d(Xp), q, p, int(num_mc), _d(best),
                  _d(table), table.size, gptr)
        else:
            gps = self._ensemble(hypers, noises, X, y, derivs)
            arr = (ctypes.c_void_p * M)(*[gp.h for gp in gps])
            f = self._fn("kg_mcmc")
            f.restype = ctypes.c_double
            v = f(arr, M, int(num_fidelity), _d(gd), _d(ib), _d(D), num_pts, _d(Xq), _d(Xp), q, p, int(num_mc),
                  _d(best), _d(table), table.size, gptr)
        return (v, g) if grad else v

    def ei_mcmc(self, hypers, noises, X, y, derivs, Xq, Xp, num_mc, best_so_far, table, grad=False):
        hypers, noises, X, y = _f64(hypers), _f64(noises), _f64(X), _f64(y)
        derivs = _i32(derivs if derivs is not None else [])
        dim = X.shape[1]
        Xq = _f64(Xq).reshape(-1, dim)
        Xp = _f64(Xp if Xp is not None else np.zeros((0, dim)))
        q, p = Xq.shape[0], (Xp.size // dim)
        best = _f64(best_so_far).ravel()
        table = _f64(table).ravel()
        M = hypers.shape[0]
        g = np.zeros((q, dim)) if grad else None
        gptr = _d(g) if grad else None
        f = self._fn("ei_mcmc")
        f.restype = ctypes.c_double
        if self.prefix == "ref_":
            v = f(_d(hypers), _d(noises), M, _d(X), _d(y), _i(derivs), derivs.size, dim, X.shape[0], _d(Xq), _d(Xp),
                  q, p, int(num_mc), _d(best), _d(table), table.size, gptr)
        else:
            gps = self._ensemble(hypers, noises, X, y, derivs)
            arr = (ctypes.c_void_p * M)(*[gp.h for gp in gps])
            v = f(arr, M, _d(Xq), _d(Xp), q, p, int(num_mc), _d(best), _d(table), table.size, gptr)
        return (v, g) if grad else v

    def log_marginal_likelihood(self, kernel, alpha, lengths, X, y, noise, derivs=None):
        X = _f64(X)
        N, dim = X.shape
        derivs = _i32(derivs if derivs is not None else [])
        f = self._fn("log_marginal_likelihood")
        f.restype = ctypes.c_double
        return f(int(kernel), ctypes.c_double(alpha), _d(_f64(lengths)), _d(X), _d(_f64(y).ravel()), _d(_f64(noise).ravel()),
                 _i(derivs), derivs.size, dim, N)

    def grad_log_marginal_likelihood(self, kernel, alpha, lengths, X, y, noise, derivs=None):
        X = _f64(X)
        N, dim = X.shape
        derivs = _i32(derivs if derivs is not None else [])
        grad = np.zeros(dim + 1 + 1 + derivs.size)
        self._fn("grad_log_marginal_likelihood")(int(kernel), ctypes.c_double(alpha), _d(_f64(lengths)), _d(X),
                                                 _d(_f64(y).ravel()), _d(_f64(noise).ravel()), _i(derivs), derivs.size,
                                                 dim, N, _d(grad))
        return grad

    def max_threads(self):
        f = self._fn("max_threads")
        f.restype = ctypes.c_int
        return f()


def oracle_path():
    return os.path.join(_HERE, "libmoe_oracle.so")


def reference_path():
    return os.path.join(_HERE, "_ref", "libmoe_ref.so")


def load_oracle():
    return CpuBackend(oracle_path(), "oracle_", "port")


def load_reference():
    return CpuBackend(reference_path(), "ref_", "reference")


def have_reference():
    return os.path.exists(reference_path())


def philox_normals(seed, first_draw, num_draws, per_draw):
    """Host restatement of the CUDA path's Philox4x32-10 + Box-Muller stream: [num_draws, per_draw]."""
    lib = ctypes.CDLL(oracle_path())
    out = np.empty(num_draws * per_draw)
    lib.oracle_philox_normals(ctypes.c_uint64(seed), ctypes.c_uint64(first_draw), int(num_draws), int(per_draw), _d(out))
    return out.reshape(num_draws, per_draw)


def kg_grad_at_point_list(backend, gp, candidates, Xp, num_mc, best_so_far, gd, inner_bounds, discrete_pts,
                          num_threads, seed=1, num_fidelity=0, want_grad=True):
    """The CPU baseline call: value (+gradient) of q-KG for a list of candidates, OpenMP over candidates.

    reference back end: the reference's own evaluator/state classes with one NormalRNG per thread;
    port back end: the C oracle with the Philox table."""
    cand = _f64(candidates)
    nc, q, dim = cand.shape
    Xp = _f64(Xp).reshape(-1, dim) if Xp is not None and len(Xp) else np.zeros((0, dim))
    gd = _f64(gd)
    ib = _f64(inner_bounds).ravel()
    disc = _f64(discrete_pts).reshape(-1, dim - num_fidelity)
    vals = np.empty(nc)
    grads = np.empty((nc, q, dim)) if want_grad else None
    if backend.prefix == "ref_":
        if want_grad:
            backend.lib.ref_kg_grad_at_point_list(gp.h, int(num_fidelity), _d(gd), _d(ib), _d(disc), disc.shape[0],
                                                  _d(cand), _d(Xp), nc, q, Xp.shape[0], int(num_mc),
                                                  ctypes.c_double(best_so_far), int(num_threads),
                                                  ctypes.c_uint(seed), _d(vals), _d(grads))
        else:
            dom = np.tile([0.0, 1.0], dim)
            backend.lib.ref_evaluate_kg_at_point_list(gp.h, int(num_fidelity), _d(gd), _d(dom), _d(ib), _d(disc),
                                                      disc.shape[0], _d(cand), _d(Xp), nc, q, Xp.shape[0],
                                                      int(num_mc), ctypes.c_double(best_so_far), int(num_threads),
                                                      ctypes.c_uint(seed), _d(vals), None)
    else:
        backend.lib.oracle_kg_at_point_list(gp.h, int(num_fidelity), _d(gd), _d(ib), _d(disc), disc.shape[0], _d(cand),
                                            _d(Xp), nc, q, Xp.shape[0], int(num_mc), ctypes.c_double(best_so_far),
                                            int(num_threads), ctypes.c_uint64(seed), _d(vals), _d(grads))
    return vals, grads


def ei_grad_at_point_list(backend, gp, candidates, Xp, num_mc, best_so_far, num_threads, seed=1):
    """CPU baseline for q-EI: value + gradient for a list of candidates, OpenMP static over candidates
    (reference back end: ExpectedImprovementEvaluator + one NormalRNG per thread; port: the C oracle)."""
    cand = _f64(candidates)
    nc, q, dim = cand.shape
    Xp = _f64(Xp).reshape(-1, dim) if Xp is not None and len(Xp) else np.zeros((0, dim))
    vals = np.empty(nc)
    grads = np.empty((nc, q, dim))
    if backend.prefix == "ref_":
        backend.lib.ref_ei_grad_at_point_list(gp.h, _d(cand), _d(Xp), nc, q, Xp.shape[0], int(num_mc),
                                              ctypes.c_double(best_so_far), int(num_threads), ctypes.c_uint(seed),
                                              _d(vals), _d(grads))
    else:
        backend.lib.oracle_ei_at_point_list(gp.h, _d(cand), _d(Xp), nc, q, Xp.shape[0], int(num_mc),
                                            ctypes.c_double(best_so_far), int(num_threads), ctypes.c_uint64(seed),
                                            _d(vals), _d(grads))
    return vals, grads


def normal_draws(seed, count):
    """First `count` draws of the reference's NormalRNG(seed) — the table every evaluation of a multistart driver call
    replays (each evaluation rewinds the generator).  Reference back end only."""
    lib = load_reference().lib
    out = np.empty(count)
    lib.ref_normal_draws(ctypes.c_uint(seed), int(count), _d(out))
    return out


def ref_multistart_kg(gp, starts, Xp, num_mc, best_so_far, outer, inner, domain_bounds, inner_bounds, discrete_pts,
                      seed, num_fidelity=0):
    """The reference's ComputeKGOptimalPointsToSampleViaMultistartGradientDescent, unmodified, one thread.
    Returns (best_point [q, dim], found_flag)."""
    lib = load_reference().lib
    starts = _f64(starts)
    ns, q, dim = starts.shape
    Xp = _f64(Xp).reshape(-1, dim) if Xp is not None and len(Xp) else np.zeros((0, dim))
    disc = _f64(discrete_pts).reshape(-1, dim - num_fidelity)
    best = np.empty((q, dim))
    lib.ref_multistart_kg.restype = ctypes.c_int
    found = lib.ref_multistart_kg(gp.h, int(num_fidelity), _d(_f64(outer)), _d(_f64(inner)),
                                  _d(_f64(domain_bounds).ravel()), _d(_f64(inner_bounds).ravel()), _d(disc),
                                  disc.shape[0], _d(starts), ns, q, _d(Xp), Xp.shape[0], int(num_mc),
                                  ctypes.c_double(best_so_far), ctypes.c_uint(seed), _d(best))
    return best, bool(found)


def ref_multistart_ei(gp, starts, Xp, num_mc, best_so_far, outer, domain_bounds, seed):
    """The reference's ComputeOptimalPointsToSampleViaMultistartGradientDescent, unmodified, one thread."""
    lib = load_reference().lib
    starts = _f64(starts)
    ns, q, dim = starts.shape
    Xp = _f64(Xp).reshape(-1, dim) if Xp is not None and len(Xp) else np.zeros((0, dim))
    best = np.empty((q, dim))
    lib.ref_multistart_ei(gp.h, _d(_f64(outer)), _d(_f64(domain_bounds).ravel()), _d(starts), ns, q, _d(Xp),
                          Xp.shape[0], int(num_mc), ctypes.c_double(best_so_far), ctypes.c_uint(seed), _d(best))
    return best


def ref_limit_update_simplex(bounds, mrc, x, upd):
    """SimplexIntersectTensorProductDomain::LimitUpdate of the compiled reference; returns the limited update."""
    lib = load_reference().lib
    x = _f64(x).ravel()
    upd = _f64(upd).ravel().copy()
    lib.ref_limit_update_simplex(_d(_f64(bounds).ravel()), x.size, ctypes.c_double(mrc), _d(x), _d(upd))
    return upd


def ref_multistart_ei_simplex(gp, starts, Xp, num_mc, best_so_far, outer, domain_bounds, seed):
    lib = load_reference().lib
    starts = _f64(starts)
    ns, q, dim = starts.shape
    Xp = _f64(Xp).reshape(-1, dim) if Xp is not None and len(Xp) else np.zeros((0, dim))
    best = np.empty((q, dim))
    lib.ref_multistart_ei_simplex(gp.h, _d(_f64(outer)), _d(_f64(domain_bounds).ravel()), _d(starts), ns, q, _d(Xp),
                                  Xp.shape[0], int(num_mc), ctypes.c_double(best_so_far), ctypes.c_uint(seed), _d(best))
    return best
