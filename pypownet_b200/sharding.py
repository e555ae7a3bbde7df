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


def env_starts_spread(n_chronics, n_rows, n_local, rank=0, rank_offset=97):
    """Starting chronic / first row of the n_local envs of a rank such that EVERY rank's batch samples the whole data
    set: local env k plays chronic k mod n_chronics from a row spread evenly over the chronic's rows, and rank r is
    shifted by r * rank_offset rows.  A step lasts as long as its slowest env, and that is decided by which (chronic,
    row) pairs the batch holds: with contiguous windows (env_starts) the batch of one GPU can miss or contain the data
    set's worst chains for hundreds of steps in a row (measured: 0.32 ms against 0.45 ms per step for two windows of the
    same IEEE-14 chronics), so per-GPU work is not the same on every GPU.  Spread starts make the per-GPU workload
    statistically identical whatever the number of GPUs -- which is what weak scaling assumes.  Local env 0 of rank 0
    starts on chronic 0, row 0 (the reference's own starting point)."""
    k = np.arange(int(n_local))
    per_chronic = -(-int(n_local) // n_chronics)
    span = max(n_rows - 1, 1)
    rows = ((k // n_chronics) * span) // per_chronic + rank * rank_offset
    return (k % n_chronics).astype(np.int32), (rows % span).astype(np.int32)


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
    """All-gather of equally sized packs -> [world * B, 7] on every rank, rank-major: rows [r * B, (r + 1) * B) are rank
    r's envs in its local order (with strided_env_ids sharding, local env k of rank r is global env r + world * k)."""
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
            if self.step:
                prev = self.gathered[buf ^ 1]          # one rank: the previous rows are simply the previous buffer
        self.step += 1
        return prev

    def wait_all(self):
        for k in (0, 1):
            if self.works[k] is not None:
                self.works[k].wait()
                self.works[k] = None
        return self.gathered[(self.step - 1) & 1] if self.step else None


class PeerGather(object):
    """The rewards / dones / flags of every shard on the collecting rank WITHOUT a collective between two steps.

    Rank `root` owns, on its GPU, a ring of `ring` result buffers [world * n_local, 7] float64 plus one step counter per
    rank and a `consumed` counter (ppn_peer_alloc); every other rank maps that memory through CUDA IPC (ppn_peer_open;
    the 64-byte handle travels once through torch.distributed).  Step t of rank r:

        before_step(t)   r != root: wait (on the GPU, ppn_peer_wait over NVLink) until the ring slot t % ring has been
                         consumed; point the handle's result pack (ppn_set_result_pack) at rank r's rows of that slot
        env.step(...)    the step kernel stores its 56-byte rows straight into the root's memory (NVLink stores)
        after_step(t)    release-store t + 1 into the root's counter of rank r (ppn_peer_signal)

    and on the root, collect(t) -- issued on a side stream, typically one step behind -- waits until all `world`
    counters reached t + 1, copies the slot into pinned host memory and bumps `consumed`.  The ranks never run in
    lock-step: a rank can be up to `ring` steps ahead of the slowest one.  world == 1: a local buffer, no waits."""

    FLAG_BYTES = 1024          # the counters sit in front of the ring: [world] step counters, then `consumed` at word 64

    def __init__(self, env, rank, world, root=0, ring=4, group=None):
        import ctypes as C

        from pypownet_b200 import _lib
        self.lib = _lib.load()
        self.env, self.rank, self.world, self.root, self.ring = env, int(rank), int(world), int(root), int(ring)
        self.n_local = env.n_envs
        self.device = env.device_index
        self.slot_doubles = self.world * self.n_local * PACK_WIDTH
        self.bytes = self.FLAG_BYTES + self.ring * self.slot_doubles * 8
        self.base = C.c_void_p()
        self.is_root = self.rank == self.root
        handle = torch.zeros(64, dtype=torch.uint8)
        if self.is_root:
            buf = (C.c_uint8 * 64)()
            self._ck(self.lib.ppn_peer_alloc(self.device, self.bytes, C.byref(self.base), buf))
            handle = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
        if self.world > 1:
            h = handle.to(env.device)
            dist.broadcast(h, src=self.root, group=group)
            if not self.is_root:
                hb = (C.c_uint8 * 64).from_buffer_copy(bytes(h.cpu().numpy().tobytes()))
                self._ck(self.lib.ppn_peer_open(self.device, hb, C.byref(self.base)))
            dist.barrier(group=group)
        self.side = torch.cuda.Stream(device=env.device) if self.is_root else None
        self.host = [torch.empty((self.world * self.n_local, PACK_WIDTH), dtype=torch.float64).pin_memory()
                     for _ in range(self.ring)] if self.is_root else None
        self.collected = 0                # steps whose rows the root has copied out

    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.ppn_last_error(None)
            raise RuntimeError('%s (code %d)' % (msg.decode() if msg else 'peer memory call failed', rc))

    def _addr(self, byte_off):
        import ctypes as C
        return C.c_void_p(self.base.value + byte_off)

    def _stream(self):
        import ctypes as C
        return C.c_void_p(torch.cuda.current_stream(self.env.device).cuda_stream)

    def rows_address(self, t):
        """Device address (on the root's GPU) of this rank's rows of step t."""
        slot = t % self.ring
        return self.base.value + self.FLAG_BYTES + 8 * (slot * self.slot_doubles + self.rank * self.n_local * PACK_WIDTH)

    def before_step(self, t):
        import ctypes as C
        if not self.is_root and t >= self.ring:
            self._ck(self.lib.ppn_peer_wait(self.device, self._addr(8 * 64), 1, t - self.ring + 1, self._stream()))
        self.env._check(self.lib.ppn_set_result_pack(self.env.handle, C.c_void_p(self.rows_address(t))))

    def after_step(self, t):
        self._ck(self.lib.ppn_peer_signal(self.device, self._addr(8 * self.rank), t + 1, self._stream()))

    def collect(self, t):
        """Root only: enqueue, on the side stream, `wait for the rows of step t -> copy to pinned host memory -> mark the
        slot consumed`.  Returns the pinned tensor (complete after wait_all / a synchronisation of the side stream)."""
        import ctypes as C
        if not self.is_root:
            return None
        if self.ring > 1 and t >= self.collected + self.ring:
            raise RuntimeError('collect(%d): the ring of %d buffers has been overrun' % (t, self.ring))
        s = C.c_void_p(self.side.cuda_stream)
        slot = t % self.ring
        self._ck(self.lib.ppn_peer_wait(self.device, self._addr(0), self.world, t + 1, s))
        self._ck(self.lib.ppn_peer_read(self.device, C.c_void_p(self.host[slot].data_ptr()),
                                        self._addr(self.FLAG_BYTES + 8 * slot * self.slot_doubles),
                                        8 * self.slot_doubles, s))
        self._ck(self.lib.ppn_peer_signal(self.device, self._addr(8 * 64), t + 1, s))
        self.collected = t + 1
        return self.host[slot]

    def wait_all(self):
        if self.side is not None:
            self.side.synchronize()

    def close(self):
        if self.base.value:
            torch.cuda.synchronize(self.env.device)
            if self.is_root:
                self.lib.ppn_peer_free(self.device, self.base)
            else:
                self.lib.ppn_peer_close(self.device, self.base)
            self.base.value = None
