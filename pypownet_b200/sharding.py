"""Env-batch sharding over the GPUs of one box: envs are independent, so rank r owns a
contiguous block of the global batch and steps it with its own VecRunEnv; the only exchange is one all-gather per
step of the packed (five sub-rewards, done, flag) rows, over NCCL/NVLink on GPUs (gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist

PACK_WIDTH = 7          # reward[5] | done | flag, as float64 (exact for these integer codes)


def shard_bounds(n_envs_total, rank, world_size):
    """[lo, hi) of the global env indices owned by `rank`; blocks differ by at most one env."""
    base, extra = divmod(int(n_envs_total), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def env_starts(n_chronics, n_rows, lo, hi):
    """Starting chronic / first row of global envs lo..hi-1: env e plays chronic e mod n_chronics from row
    (e // n_chronics) mod (n_rows - 1), identical whatever the number of shards."""
    e = np.arange(lo, hi)
    return (e % n_chronics).astype(np.int32), ((e // n_chronics) % max(n_rows - 1, 1)).astype(np.int32)


def strided_env_ids(n_envs_per_rank, rank, world_size):
    """Global env indices of a rank when the batch is dealt round-robin (env e -> GPU e mod n_gpu): every
    shard then samples the same mix of chronics and rows, so no rank carries a systematically heavier batch -- the
    step time of a synchronous run is the maximum over ranks."""
    return rank + world_size * np.arange(int(n_envs_per_rank))


def env_starts_of(n_chronics, n_rows, env_ids):
    """Starting chronic / first row of the given global env indices (same rule as env_starts)."""
    e = np.asarray(env_ids)
    return (e % n_chronics).astype(np.int32), ((e // n_chronics) % max(n_rows - 1, 1)).astype(np.int32)


def pack_results(reward, done, flag, out=None):
    """[B, 7] float64 rows from reward [B,5] f64, done [B] u8, flag [B] i32 (any device)."""
    B = reward.shape[0]
    if out is None:
        out = torch.empty((B, PACK_WIDTH), dtype=torch.float64, device=reward.device)
    out[:, :5] = reward
    out[:, 5] = done
    out[:, 6] = flag
    return out


def unpack_results(packed):
    return packed[:, :5], packed[:, 5].to(torch.uint8), packed[:, 6].to(torch.int32)


def gather_results(packed, world_size=None, out=None):
    """All-gather of equally sized packs -> [world * B, 7] on every rank (rank-major = global env order)."""
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if world_size == 1:
        return packed
    if out is None:
        out = torch.empty((world_size * packed.shape[0], packed.shape[1]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed.contiguous())
    return out
