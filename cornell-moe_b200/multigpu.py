"""One-process-per-GPU sharding of the multistart axis (the path's only data-parallel axis).

The reference parallelises multistart candidates with an OpenMP ``parallel for`` and reduces with a critical section
(gpp_optimization.hpp:1472-1546).  Here every rank owns a full replica of the (small) GP state on its GPU and a strided
share of the candidates; there is NO data-path collective.  The only exchange is one all-gather of per-candidate values
(8 bytes each) so that every rank selects the identical top-20 / arg-max, and one all-gather of the 20 optimised
points.  Works with any ``torch.distributed`` backend: NCCL over NVLink on the GPU box (tensors on the rank's device),
gloo on CPU (used by the world_size-2 tests, where the compute callbacks are stand-ins).

Determinism: candidate c always lives on rank c % world, its MC stream depends only on the seed (common random numbers),
values are gathered in candidate order, ties break on the lowest index -> the selected index is identical for 1/2/4/8 ranks.
"""
import numpy as np

TOP_K = 20  # hard-coded in the reference: gpp_knowledge_gradient_optimization.hpp:901, gpp_math.hpp:1765


def _dist():
    import torch.distributed as dist
    return dist


def world_info(group=None):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_indices(n, rank, world):
    """Indices owned by `rank`: candidate c belongs to rank c % world."""
    return np.arange(rank, n, world)


def _gather_padded(local, per_rank, width, device, group=None):
    """All-gather a [len(local), width] float64 array padded to per_rank rows; returns [world, per_rank, width]."""
    import torch
    dist = _dist()
    rank, world = world_info(group)
    buf = torch.full((per_rank, width), float("nan"), dtype=torch.float64, device=device)
    if len(local):
        buf[: len(local)] = torch.as_tensor(np.asarray(local, dtype=np.float64).reshape(len(local), width), device=device)
    if world == 1:
        return buf.cpu().numpy()[None]
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.stack(out).cpu().numpy()


def gather_values(local_values, n_total, device="cpu", group=None):
    """Reassemble the length-n_total vector of per-candidate values from every rank's strided share."""
    rank, world = world_info(group)
    per_rank = (n_total + world - 1) // world
    g = _gather_padded(np.asarray(local_values).reshape(-1, 1), per_rank, 1, device, group)  # [world, per_rank, 1]
    full = g[:, :, 0].T.reshape(-1)  # index = local * world + rank
    return full[:n_total]


def top_k_indices(values, k=TOP_K):
    """The reference's priority-queue selection of the k largest values, in the order it feeds them to the optimiser
    (ascending value; ties between equal values put the larger index first) — gpp_knowledge_gradient_optimization.hpp:900-921."""
    import heapq
    heap = []  # max-heap on (-value, index) emulated with a min-heap of negated keys
    k = min(k, len(values))
    for i, v in enumerate(values):
        key = (-float(v), i)
        if i < k:
            heapq.heappush(heap, (-key[0], -key[1]))
        else:
            top = (-heap[0][0], -heap[0][1])  # largest (-value, index) currently kept
            if top[0] > key[0]:
                heapq.heapreplace(heap, (-key[0], -key[1]))
    order = []
    while heap:
        a, b = heapq.heappop(heap)
        order.append(-b)
    return order


def sharded_multistart(evaluate_fn, descend_fn, starts, init_best=-np.inf, device="cpu", group=None):
    """Sharded version of the multistart driver.

    evaluate_fn(starts_subset) -> values [m]            (device batch evaluation of this rank's share)
    descend_fn(starts_subset)  -> (values [m], points [m, ...])   (restarted gradient descent on this rank's share)
    Returns (best_point, best_value, found_flag, start_values) — identical on every rank.
    """
    starts = np.asarray(starts, dtype=np.float64)
    n = starts.shape[0]
    rank, world = world_info(group)
    mine = shard_indices(n, rank, world)
    local_vals = evaluate_fn(starts[mine]) if len(mine) else np.zeros(0)
    values = gather_values(local_vals, n, device, group)
    top = top_k_indices(values)
    k = len(top)
    my_slots = shard_indices(k, rank, world)
    width = int(np.prod(starts.shape[1:]))
    if len(my_slots):
        v, p = descend_fn(starts[[top[s] for s in my_slots]])
        local = np.concatenate([np.asarray(v, dtype=np.float64).reshape(-1, 1),
                                np.asarray(p, dtype=np.float64).reshape(len(my_slots), width)], axis=1)
    else:
        local = np.zeros((0, 1 + width))
    per_rank = (k + world - 1) // world
    g = _gather_padded(local, per_rank, 1 + width, device, group)  # [world, per_rank, 1+width]
    flat = g.transpose(1, 0, 2).reshape(-1, 1 + width)[:k]       # slot = local * world + rank
    best_value, found = init_best, False
    best_point = starts[top[0]].copy()
    for s in range(k):  # strict `<` update in slot order (gpp_optimization.hpp:1511, 1540), lowest slot wins ties
        if best_value < flat[s, 0]:
            best_value, found = flat[s, 0], True
            best_point = flat[s, 1:].reshape(starts.shape[1:])
    return best_point, float(best_value), found, values


def multistart_kg(gp, starts, Xp, num_mc, best_so_far, outer, inner, domain_bounds, inner_bounds, discrete_pts,
                  num_fidelity=0, seed=0, device="cpu", group=None):
    """multistart_knowledge_gradient_optimization sharded over the ranks of `group` (each rank passes its own `gp`)."""
    from . import capi

    # the reference driver's state keeps the FIRST start's points in the inner optimiser's discretisation set
    # (cmoe_kg_plan_set_stale_union); every rank uses the global first start so that the shards reproduce the one-process
    # driver bit for bit
    starts = np.asarray(starts, dtype=np.float64)
    stale = starts[0]
    rank, world = world_info(group)
    share = max(1, -(-starts.shape[0] // world))
    plan = capi.KGPlan(gp, num_mc, best_so_far, inner, inner_bounds, discrete_pts, share, starts.shape[1], Xp=Xp,
                       num_fidelity=num_fidelity, seed=seed, want_grad=False)
    plan.set_stale_union(stale)

    def evaluate(sub):
        plan.upload(sub)
        plan.run()
        plan.sync()
        return plan.download()[0]

    def descend(sub):
        return capi.kg_gradient_descent(gp, sub, Xp, num_mc, best_so_far, outer, inner, domain_bounds, inner_bounds,
                                        discrete_pts, num_fidelity=num_fidelity, seed=seed, stale_union=stale)

    return sharded_multistart(evaluate, descend, starts, -np.inf, device, group)


def multistart_ei(gp, starts, Xp, num_mc, best_so_far, outer, domain_bounds, seed=0, device="cpu", group=None):
    from . import capi

    analytic = np.asarray(starts).shape[1] == 1 and (Xp is None or len(Xp) == 0)

    def evaluate(sub):
        if analytic:  # closed-form 1-EI, as the reference and cmoe_multistart_ei screen (gpp_math.hpp:1703-1749)
            return capi.ei_analytic(gp, np.asarray(sub).reshape(-1, np.asarray(sub).shape[-1]), best_so_far)
        return gp.ei(sub, Xp, num_mc, best_so_far, seed=seed)

    def descend(sub):
        return capi.ei_gradient_descent(gp, sub, Xp, num_mc, best_so_far, outer, domain_bounds, seed=seed)

    return sharded_multistart(evaluate, descend, starts, -1.0, device, group)


def multistart_kg_mcmc(ens, starts, Xp, num_mc, best_so_far, outer, inner, domain_bounds, inner_bounds, discrete_pts,
                       num_fidelity=0, seed=0, device="cpu", group=None):
    """multistart_knowledge_gradient_mcmc_optimization sharded over the ranks of `group`: every rank passes its own
    `capi.GaussianProcessEnsemble` (all members on its GPU) and evaluates / descends its strided share of the starts."""

    def evaluate(sub):
        return ens.kg(sub, Xp, num_mc, best_so_far, inner, inner_bounds, discrete_pts, num_fidelity=num_fidelity,
                      seed=seed)

    def descend(sub):
        return ens.kg_gradient_descent(sub, Xp, num_mc, best_so_far, outer, inner, domain_bounds, inner_bounds,
                                       discrete_pts, num_fidelity=num_fidelity, seed=seed)

    return sharded_multistart(evaluate, descend, starts, -np.inf, device, group)


def multistart_ei_mcmc(ens, starts, Xp, num_mc, best_so_far, outer, domain_bounds, seed=0, device="cpu", group=None):
    """multistart_expected_improvement_mcmc_optimization sharded over ranks (initial best 0.0 as in the reference)."""

    def evaluate(sub):
        return ens.ei(sub, Xp, num_mc, best_so_far, seed=seed, analytic_single=True)

    def descend(sub):
        return ens.ei_gradient_descent(sub, Xp, num_mc, best_so_far, outer, domain_bounds, seed=seed)

    return sharded_multistart(evaluate, descend, starts, 0.0, device, group)
