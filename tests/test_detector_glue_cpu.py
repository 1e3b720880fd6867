"""f3: parse_losses == BaseDetector._parse_losses (mmdet/models/detectors/base.py:176-219), single process and gloo world 2."""
import os
from collections import OrderedDict

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from boxinstseg_b200.models.detectors import parse_losses


def _reference_parse(losses):
    """The reference's loop, restated: mean per key, sum of list means, 'loss' = sum of the keys containing 'loss', every value
    averaged over ranks, floats."""
    log_vars = OrderedDict()
    for k, v in losses.items():
        log_vars[k] = v.mean() if isinstance(v, torch.Tensor) else sum(x.mean() for x in v)
    loss = sum(v for k, v in log_vars.items() if 'loss' in k)
    log_vars['loss'] = loss
    out = OrderedDict()
    for k, v in log_vars.items():
        v = v.data.clone()
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(v.div_(dist.get_world_size()))
        out[k] = v.item()
    return loss, out


def _losses(seed):
    g = torch.Generator().manual_seed(seed)
    return OrderedDict(loss_cls=torch.rand(7, generator=g), loss_bbox=torch.rand((), generator=g),
                       loss_prj=[torch.rand(3, generator=g), torch.rand(2, 2, generator=g)], acc=torch.rand(5, generator=g),
                       loss_pairwise=torch.rand(1, generator=g))


def test_parse_losses_single_process():
    losses = _losses(0)
    for v in losses.values():
        for t in (v if isinstance(v, list) else [v]):
            t.requires_grad_(True)
    loss, log_vars = parse_losses(losses)
    ref_loss, ref_vars = _reference_parse(losses)
    assert torch.equal(loss, ref_loss) and loss.requires_grad
    assert list(log_vars.keys()) == list(ref_vars.keys())
    for k in ref_vars:
        assert log_vars[k] == pytest.approx(ref_vars[k], rel=1e-6)
    assert 'acc' in log_vars and len(log_vars) == 6
    _, eager = parse_losses(losses, defer=False)
    assert dict(eager.items()) == dict(log_vars.items())
    with pytest.raises(TypeError):
        parse_losses({'loss_x': 1.0})


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    losses = _losses(10 + rank)
    loss, log_vars = parse_losses(losses)
    ref_loss, ref_vars = _reference_parse(losses)
    ok = torch.equal(loss, ref_loss) and all(abs(log_vars[k] - ref_vars[k]) < 1e-6 for k in ref_vars)
    # a rank with a different key set must fail with the reference's assertion (base.py:199-207), not hang
    bad = _losses(10 + rank)
    if rank == 1:
        bad['loss_extra'] = torch.ones(1)
    caught = False
    try:
        _, lv = parse_losses(bad)
        lv['loss']                                   # the check happens where the values are read
    except AssertionError as exc:
        caught = 'different across GPUs' in str(exc)
    ret[rank] = (ok, caught)
    dist.destroy_process_group()


def test_parse_losses_gloo_world2():
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0][0] and ret[1][0]
    assert ret[0][1] and ret[1][1]
