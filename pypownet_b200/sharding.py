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


class PipelinedGather(object):
    """One all-gather per step whose communication overlaps the next step: step t writes its packed rows into buffer
    t mod 2 and starts the gather asynchronously; the gather of step t-1 is waited for right after step t has been
    enqueued (a stream wait on GPUs), so a buffer is never rewritten while its gather is in flight.  `wait_all` before
    reading the last results.  bench.py's sharded loop follows the same discipline over NCCL (the two pack buffers are
    registered with the handle through ppn_set_result_pack); tests/test_sharding_gloo.py covers it on gloo."""

    def __init__(self, n_local, world_size, device='cpu'):
        self.world = int(world_size)
        self.packs = [torch.zeros((n_local, PACK_WIDTH), dtype=torch.float64, device=device) for _ in range(2)]
        self.gathered = [torch.zeros((self.world * n_local, PACK_WIDTH), dtype=torch.float64, device=device)
                         for _ in range(2)]
        self.works = [None, None]
        self.step = 0

    def buffer(self):
        """The pack buffer the step about to run must fill."""
        return self.packs[self.step & 1]

    def launch(self):
        """Call once the step that fills `buffer()` has been enqueued.  Returns the gathered rows of the PREVIOUS step
        (complete), or None on the first call."""
        buf = self.step & 1
        prev = None
        if self.works[buf ^ 1] is not None:
            self.works[buf ^ 1].wait()
            self.works[buf ^ 1] = None
            prev = self.gathered[buf ^ 1]
        if self.world > 1:
            self.works[buf] = dist.all_gather_into_tensor(self.gathered[buf], self.packs[buf], async_op=True)
        else:
            self.gathered[buf].copy_(self.packs[buf])
        self.step += 1
        return prev

    def wait_all(self):
        for k in (0, 1):
            if self.works[k] is not None:
                self.works[k].wait()
                self.works[k] = None
        return self.gathered[(self.step - 1) & 1] if self.step else None
