"""Host-side plumbing of the multi-GPU path: the env axis is the only thing that shards.

Every environment's solve is independent (the reference expresses this as ``jax.vmap``: examples/multi_env.py:46-49,
mpx/utils/mpc_wrapper_srbd.py:77), so rank r of G owns a contiguous slice of the environments together with their
warm-start state, which never leaves its GPU.  There is no collective on the solve path; the only exchanges are one
gather of the controls ``tau [E/G, nj]`` per tick and the max-over-ranks reduction of the device time in the
benchmark.  The helpers work on any ``torch.distributed`` backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def env_slice(n_env_total, world, rank):
    """Contiguous [lo, hi) slice of the env axis owned by ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world) or n_env_total < 0:
        raise ValueError(f"bad shard request: n_env_total={n_env_total} world={world} rank={rank}")
    base, rem = divmod(n_env_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_controls(tau_local, n_env_total=None, dst=0, group=None):
    """Gather the per-rank controls ``[E_r, nj]`` into ``[E, nj]`` on ``dst`` (None elsewhere), in env order.

    Slices may be ragged (``env_slice``); they are padded to the largest slice for the collective and trimmed after."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tau_local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [tau_local.shape[0]] * world
    if n_env_total is not None:
        sizes = [hi - lo for lo, hi in (env_slice(n_env_total, world, r) for r in range(world))]
        if sizes[rank] != tau_local.shape[0]:
            raise ValueError(f"rank {rank} holds {tau_local.shape[0]} envs, its slice has {sizes[rank]}")
    width = max(sizes)
    send = tau_local
    if tau_local.shape[0] != width:
        send = torch.zeros((width,) + tuple(tau_local.shape[1:]), dtype=tau_local.dtype, device=tau_local.device)
        send[:tau_local.shape[0]] = tau_local
    send = send.contiguous()
    out = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, out, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def max_over_ranks(value, device=None, group=None):
    """Max of a host scalar over the ranks (device-timed milliseconds in the benchmark)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
