"""Device-timed Krotov iterations of ONE ensemble sharded over the GPUs of a
torchrun job (GPUShards mode 'sharded': kq_krotov_iteration with a kq_comm)
against the same problem on one GPU.  Prints one JSON line per case (rank 0).

    torchrun --nproc-per-node N tools/multigpu_probe.py [K ...]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import krotov_b200 as krotov  # noqa: E402
from krotov_b200.compiler import compile_problem, initialize_controls  # noqa: E402
from krotov_b200.engine import SweepEngine  # noqa: E402
from krotov_b200.parallelization import ShardComm, shard_bounds  # noqa: E402


def time_case(K, nt, sharded, steps=20, warmup=5):
    rank, world = dist.get_rank(), dist.get_world_size()
    wl = krotov.workloads.tls_ensemble(K=K, nt=nt)
    objectives = wl.objectives(krotov.Objective)
    controls, _, guess, mapping, lam, shp = initialize_controls(
        objectives, wl.pulse_options, wl.tlist)
    lo, hi = shard_bounds(K, world, rank) if sharded else (0, K)
    cp = compile_problem(objectives[lo:hi], controls, mapping[lo:hi], wl.tlist)
    eng = SweepEngine(cp, shp, lam)
    shard = None
    if sharded:
        shard = ShardComm(dist, None, eng.device).attach(eng)
        eng.K_total = K
    stream = torch.cuda.current_stream()
    g = eng.pulses_to_device(guess)
    o = g.clone()
    phiT = eng.propagate_forward(g)
    tau = eng.overlaps(eng.t_targets, phiT)
    sp, st = eng.new_states(), torch.empty_like(tau)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32,
                        device=eng.device)
    times = []
    hint = False
    for i in range(warmup + steps):
        flush.zero_()
        dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.krotov_iteration('re', g, o, phiT, tau, sp, st,
                             prev_guess_t=o if hint else None)
        e1.record(stream)
        e1.synchronize()
        hint = True
        phiT, sp = sp, phiT
        tau, st = st, tau
        g, o = o, g
        if i >= warmup:
            times.append(e0.elapsed_time(e1))
    # back-to-back (no flush, no barrier): steady-state pipeline rate
    torch.cuda.synchronize()
    dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        eng.krotov_iteration('re', g, o, phiT, tau, sp, st, prev_guess_t=o)
        phiT, sp = sp, phiT
        tau, st = st, tau
        g, o = o, g
    e1.record(stream)
    e1.synchronize()
    b2b = e0.elapsed_time(e1) / steps
    if eng.status() != 0 or eng.first_failed_epoch() != 0:
        raise RuntimeError("fused iteration failed (status %d, epoch %d)"
                           % (eng.status(), eng.first_failed_epoch()))
    _, rounds = eng.sweep_diagnostics()
    t = torch.tensor([float(np.mean(times)), b2b], dtype=torch.float64,
                     device=eng.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pulses = g.clone()
    if shard is not None:
        shard.close()
    return float(t[0]), float(t[1]), rounds, pulses


def main():
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    cases = [int(a) for a in sys.argv[1:]] or [128, 1024 * world]
    for K in cases:
        ms1, b1, r1, p1 = time_case(K, 1000, sharded=False) \
            if K <= 1184 else (None, None, None, None)
        msN, bN, rN, pN = time_case(K, 1000, sharded=True)
        line = dict(K=K, nt=1000, n_gpus=world, ms_sharded=msN,
                    ms_sharded_back_to_back=bN, rounds_sharded=rN,
                    ms_one_gpu=ms1, ms_one_gpu_back_to_back=b1,
                    rounds_one_gpu=r1)
        if p1 is not None:
            line['rel_dev_vs_one_gpu'] = float(
                (pN - p1).abs().max() / p1.abs().max())
        if rank == 0:
            print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
