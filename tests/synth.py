"""Fixed-seed synthetic inputs shared by the parity tests."""
import numpy as np

EXAMPLE_INNER_GD = [1, 6, 1, 3, 0.0, 1.0, 0.1, 1e-10]  # examples/main.py:123-130 in the reference
DISCRETE_ONLY_GD = [1, 0, 1, 3, 0.0, 1.0, 0.1, 1e-10]  # inner max_num_steps = 0


def make_problem(N, dim, g_idx=(), seed=0, noise=1e-2, length=0.5, alpha=1.0, func="sin"):
    rng = np.random.default_rng(seed)
    X = rng.uniform(0.0, 1.0, size=(N, dim))
    g = len(g_idx)
    f = np.sin(3.0 * X).sum(axis=1)
    cols = [f]
    for a in g_idx:
        cols.append(3.0 * np.cos(3.0 * X[:, a]))
    y = np.stack(cols, axis=1) + np.sqrt(noise) * rng.standard_normal((N, 1 + g))
    lengths = np.full(dim, length) * np.sqrt(dim / 8.0) if dim >= 8 else np.full(dim, length)
    noise_v = np.full(1 + g, noise)
    return dict(X=X, y=y.ravel(), lengths=lengths, alpha=alpha, noise=noise_v, derivs=np.array(g_idx, dtype=np.int32),
                dim=dim, N=N)


def unit_bounds(dim):
    return np.tile(np.array([0.0, 1.0]), dim)
