"""CPU: the N>1 launch path of bench.py with gloo, world_size 2 (no GPU needed):
 * the rank-timing reduction is MAX over ranks;
 * under torchrun the reference arm prints exactly one JSON line on rank 0 and the other rank exits 0."""
import json
import os
import socket
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bench
    vals = bench.reduce_max_over_ranks([1.0 + rank, 10.0 - rank], dist, torch.device('cpu'))
    if rank == 0:
        out.put(vals)
    dist.destroy_process_group()


def test_rank_timings_are_max_reduced():
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert q.get() == [2.0, 10.0]


def test_reference_arm_under_torchrun_two_ranks():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', str(_free_port()), os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
           '--gpus', '2', '--steps', '1', '--warmup', '0']
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    line = lines[0]
    assert line['impl'] == 'reference' and line['n_gpus'] == 2 and line['unit'] == 'ms/img'
    assert line['cpu_baseline']['kind'] == 'port' and line['value'] > 0 and line['higher_is_better'] is False
    assert line['e2e']['h2d_bytes_per_step'] == 0
