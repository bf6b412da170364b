"""world_size-2 gloo tests (CPU) of the host-side sharding logic used by
parallel_map=GPUShards(): block bounds, row gathering, scalar reduction."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from krotov_b200.parallelization import ShardComm, shard_bounds


def test_shard_bounds_cover_everything():
    for K in (1, 2, 7, 128, 129):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(K, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == K
            for (a, b), (c, d) in zip(blocks, blocks[1:]):
                assert b == c
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, K, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        comm = ShardComm(dist, None, torch.device('cpu'))
        lo, hi = shard_bounds(K, world, rank)
        full = torch.arange(K * 3, dtype=torch.float64).reshape(K, 3)
        cfull = torch.complex(full, -full)
        got = comm.all_gather_rows(cfull[lo:hi].clone(), K)
        assert torch.equal(got, cfull)
        got = comm.all_gather_rows(full[lo:hi, 0].clone(), K)
        assert torch.equal(got, full[:, 0])
        # equally sized padded blocks gathered straight into a byte buffer
        # (the packed results of a hooked, sharded iteration)
        kmax = -(-K // world)
        n = kmax * 4
        loc = torch.zeros(n, dtype=torch.complex128)
        loc[:(hi - lo) * 3] = cfull[lo:hi].reshape(-1)
        loc[kmax * 3:kmax * 3 + (hi - lo)] = cfull[lo:hi, 0] * 2
        raw = torch.zeros(world * n * 16 + 32, dtype=torch.uint8)
        region = raw[16:16 + world * n * 16].view(torch.complex128)
        comm.all_gather_flat(region, loc)
        G = region.numpy().reshape(world, n)
        rows, taus = [], []
        for r in range(world):
            a, b = shard_bounds(K, world, r)
            rows.append(G[r, :kmax * 3].reshape(kmax, 3)[:b - a])
            taus.append(G[r, kmax * 3:kmax * 3 + (b - a)])
        assert np.array_equal(np.concatenate(rows), cfull.numpy())
        assert np.array_equal(np.concatenate(taus), 2 * cfull[:, 0].numpy())
        t = torch.tensor([complex(rank + 1, -rank)], dtype=torch.complex128)
        comm.all_reduce_sum(t)
        want = sum(complex(r + 1, -r) for r in range(world))
        assert abs(t.item() - want) < 1e-15
        out.put((rank, 'ok'))
    except Exception as exc:  # pragma: no cover
        out.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('K', [5, 128])
def test_gather_and_reduce_world2(K):
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, K, out))
             for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, 'ok'), (1, 'ok')]
