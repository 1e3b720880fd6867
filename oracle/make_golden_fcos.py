"""Golden vectors for the FCOS point assignment, minted from the REFERENCE's own methods (authoring container only):
``CondInstBoxHead.get_targets`` and ``_get_target_single`` are AST-extracted from mmdet/models/dense_heads/condinst_head.py,
bound to a stand-in ``self`` holding the attributes they read, and run on seeded inputs (every image has ground truth: the
reference's empty-image branch returns two values and cannot be unpacked).  Asserts the oracle restatement == the
reference, writes tests/golden/fcos_targets.npz.      python -m oracle.make_golden_fcos"""
import ast
import os
import types

import numpy as np
import torch

from oracle import fcos_targets as oft

REF = os.environ.get('BXS_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'fcos_targets.npz')
# configs/boxinst/boxinst_r50_fpn_1x_coco.py:24-36 + the defaults of condinst_head.py:250-260
CFG = dict(regress_ranges=((-1, 64), (64, 128), (128, 256), (256, 512), (512, oft.INF)), strides=(8, 16, 32, 64, 128),
           num_classes=80)


def multi_apply(func, *args, **kwargs):                   # mmdet.core.utils.misc.multi_apply (third party to this file)
    from functools import partial
    pfunc = partial(func, **kwargs) if kwargs else func
    return tuple(map(list, zip(*map(pfunc, *args))))


def reference_methods():
    src = open(os.path.join(REF, 'mmdet/models/dense_heads/condinst_head.py')).read()
    cls = [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == 'CondInstBoxHead'][0]
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ('get_targets', '_get_target_single')]
    g = dict(torch=torch, multi_apply=multi_apply, INF=oft.INF)
    exec(compile(ast.Module(body=fns, type_ignores=[]), 'condinst_head.py', 'exec'), g)
    return g['get_targets'], g['_get_target_single']


def case(seed, B=2, H=160, W=224, G=(7, 12), center_sampling=True, norm_on_bbox=True):
    gen = torch.Generator().manual_seed(seed)
    sizes = [((H + s - 1) // s, (W + s - 1) // s) for s in CFG['strides']]
    points = oft.grid_points(sizes, CFG['strides'])
    boxes, labels = [], []
    for b in range(B):
        g = G[b % len(G)]
        x1 = torch.rand(g, generator=gen) * (W - 16)
        y1 = torch.rand(g, generator=gen) * (H - 16)
        bw = 6 + torch.rand(g, generator=gen) ** 2 * W
        bh = 6 + torch.rand(g, generator=gen) ** 2 * H
        bx = torch.stack([x1, y1, torch.clamp(x1 + bw, max=float(W)), torch.clamp(y1 + bh, max=float(H))], 1)
        bx[0] = torch.tensor([8.0, 8.0, 40.0, 40.0])                   # edges on location coordinates: the > 0 boundaries
        if g > 2:
            bx[2] = bx[1]                                               # two ground truths of equal area: first one wins
        boxes.append(bx)
        labels.append(torch.randint(0, CFG['num_classes'], (g,), generator=gen))
    return points, boxes, labels, dict(center_sampling=center_sampling, norm_on_bbox=norm_on_bbox)


CASES = [(0, dict()), (1, dict(B=3, G=(1, 20, 5))), (2, dict(center_sampling=False)), (3, dict(norm_on_bbox=False, H=96, W=128))]


def main():
    get_targets, single = reference_methods()
    out = {}
    for seed, kw in CASES:
        points, boxes, labels, flags = case(seed, **kw)
        me = types.SimpleNamespace(center_sample_radius=1.5, **CFG, **flags)
        me._get_target_single = types.MethodType(single, me)
        r = get_targets(me, points, [b.clone() for b in boxes], [l.clone() for l in labels])
        o = oft.get_targets(points, boxes, labels, CFG['regress_ranges'], CFG['strides'], CFG['num_classes'],
                            flags['center_sampling'], 1.5, flags['norm_on_bbox'])
        for a_list, b_list in zip(r, o):
            for a, b in zip(a_list, b_list):
                assert torch.equal(a, b), 'oracle restatement != reference'
        out[f's{seed}_labels'] = torch.cat(r[0]).numpy().astype(np.int16)
        out[f's{seed}_targets'] = torch.cat(r[1]).numpy()
        out[f's{seed}_inds'] = torch.cat(r[2]).numpy().astype(np.int16)
        pos = int((torch.cat(r[2]) >= 0).sum())
        assert pos > 20, pos
    np.savez_compressed(OUT, **out)
    print('oracle == reference on', len(out), 'arrays;', os.path.getsize(OUT) // 1024, 'KiB ->', OUT)


if __name__ == '__main__':
    main()
