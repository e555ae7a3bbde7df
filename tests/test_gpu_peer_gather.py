"""Env-sharded runs on two GPUs of one box: the step kernels store their packed result rows straight into rank 0's GPU
memory over NVLink (sharding.PeerGather, ppn_peer_* of the C ABI); the rows rank 0 collects must be exactly what each
rank's own reward / done / flag tensors hold.  Needs two GPUs (skipped on a one-GPU box): run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_peer_gather.py -m gpu`."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ring, steps):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
    from golden_util import Fixture
    from pypownet_b200 import sharding
    from pypownet_b200.vec_env import VecRunEnv
    fx = Fixture('d14_ac_random')
    B = 6
    rows = (np.arange(B) * 7 + 13 * rank).astype(np.int32)
    env = VecRunEnv(fx.case, fx.config, fx.chronics, B, device=rank, game_over_mode=fx.mode,
                    reward_constant=fx.reward_constant, thermal_limits=fx.thermal_limits,
                    start_chronics=np.zeros(B, dtype=np.int32), start_rows=rows)
    pg = sharding.PeerGather(env, rank, world, ring=ring)
    mine, got = [], []
    for t in range(steps):
        pg.before_step(t)
        obs, reward, done, flag = env.step(np.repeat(fx.actions[t][None], B, axis=0), auto_reset=True)
        pg.after_step(t)
        mine.append(sharding.pack_results(reward, done, flag).clone())
        if pg.is_root and t >= 1:
            h = pg.collect(t - 1)
            pg.wait_all()
            got.append(h.clone())
    if pg.is_root:
        h = pg.collect(steps - 1)
        pg.wait_all()
        got.append(h.clone())
    torch.cuda.synchronize()
    # every rank's own rows to rank 0 through NCCL, as the reference to compare with
    local = torch.stack(mine)                                   # [steps, B, 7]
    allr = torch.zeros((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(allr, local)
    if pg.is_root:
        allr = allr.cpu()
        for t in range(steps):
            for r in range(world):
                assert torch.equal(got[t][r * B:(r + 1) * B], allr[r, t]), 'step %d rank %d' % (t, r)
    dist.barrier()
    pg.close()
    dist.destroy_process_group()


@pytest.mark.parametrize('ring', [2, 4])
def test_rows_written_over_nvlink_equal_each_ranks_own_results(ring):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, ring, 14), nprocs=2, join=True)
