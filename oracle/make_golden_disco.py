"""Golden vectors for DiscoBox's SOLO target builders, minted from the REFERENCE's own methods (authoring container only):
``DiscoBoxSOLOv2Head.solov2_target_single`` / ``best_target_single`` and the module-level ``center_of_mass`` are AST-extracted
from mmdet/models/dense_heads/discobox_head.py, bound to a stand-in ``self`` with the attributes they read, and run on seeded
inputs; ``mmcv.imrescale`` (third party, absent) is the restated ``cv2.resize`` call of oracle/solo_targets.py and
``BitmapMasks`` a three-method stand-in (index, iterate, ``to_ndarray``).  Asserts the oracle restatement == the reference,
writes tests/golden/disco_targets.npz.      python -m oracle.make_golden_disco"""
import ast
import os
import types

import numpy as np
import torch

from oracle import solo_targets as ost

REF = os.environ.get('BXS_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'disco_targets.npz')
# configs/discobox/discobox_solov2_coco_r50_fpn_3x.py:33-38 (scale ranges shrunk with the 160x192 test image)
CFG = dict(scale_ranges=((1, 24), (12, 48), (24, 96), (48, 192), (96, 512)), strides=[8, 8, 16, 32, 32],
           seg_num_grids=[40, 36, 24, 16, 12], sigma=0.2, num_classes=80)


class Masks:
    """What the two methods use of mmdet.core.mask.structures.BitmapMasks."""

    def __init__(self, arr):
        self.masks = arr

    def __getitem__(self, idx):
        return Masks(self.masks[idx])

    def __iter__(self):
        return iter(self.masks)

    def to_ndarray(self):
        return self.masks


def reference_methods():
    src = open(os.path.join(REF, 'mmdet/models/dense_heads/discobox_head.py')).read()
    tree = ast.parse(src)
    com = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'center_of_mass'][0]
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'DiscoBoxSOLOv2Head'][0]
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ('solov2_target_single', 'best_target_single')]
    g = dict(torch=torch, np=np, mmcv=types.SimpleNamespace(imrescale=ost.imrescale))
    exec(compile(ast.Module(body=[com] + fns, type_ignores=[]), 'discobox_head.py', 'exec'), g)
    return g['solov2_target_single'], g['best_target_single']


def case(seed, H=160, W=192, G=9):
    gen = torch.Generator().manual_seed(seed)
    x1 = torch.rand(G, generator=gen) * (W - 24)
    y1 = torch.rand(G, generator=gen) * (H - 24)
    bw = 8 + torch.rand(G, generator=gen) ** 2 * (W - 8)
    bh = 8 + torch.rand(G, generator=gen) ** 2 * (H - 8)
    boxes = torch.stack([x1, y1, torch.minimum(x1 + bw, torch.tensor(float(W - 1))),
                         torch.minimum(y1 + bh, torch.tensor(float(H - 1)))], 1)
    labels = torch.randint(0, 80, (G,), generator=gen)
    masks = np.zeros((G, H, W), np.uint8)
    for g_, b in enumerate(boxes):
        masks[g_, int(b[1]):int(b[3]) + 1, int(b[0]):int(b[2]) + 1] = 1      # box-shaped masks (box supervision)
    masks[1, ::3, :] = 0                                                # not a rectangle
    masks[2] = 0                                                        # an empty mask: skipped (valid_mask_flag)
    boxes[4] = boxes[3]                                                 # two ground truths on the same cells: both are
    masks[4] = masks[3]                                                 # appended, the later one owns cate_label
    return boxes, labels, masks, (H // 4, W // 4)


def main():
    general, best = reference_methods()
    me = types.SimpleNamespace(**CFG)
    me.scale_mids = torch.tensor(np.array(CFG['scale_ranges']))
    me.scale_mids = (me.scale_mids[:, 0] * me.scale_mids[:, 1]) ** 0.5                  # discobox_head.py:701-702
    out = {}
    for seed in (0, 1, 2):
        boxes, labels, masks, fsize = case(seed)
        for name, fn, flag in (('gen', general, False), ('best', best, True)):
            r = fn(me, boxes, labels, Masks(masks), mask_feat_size=fsize)
            o = ost.disco_target_single(boxes, labels, masks, fsize, best=flag, **CFG)
            for a_list, b_list in zip(r[:3], o[:3]):
                for a, b in zip(a_list, b_list):
                    assert a.shape == b.shape and torch.equal(a, b), 'oracle restatement != reference'
            assert [list(map(int, x)) for x in r[3]] == o[3], 'oracle restatement != reference (grid order)'
            assert sum(len(x) for x in r[3]) >= 4
            for lvl in range(5):
                out[f's{seed}_{name}_cate{lvl}'] = r[1][lvl].numpy().astype(np.int16)
                out[f's{seed}_{name}_ind{lvl}'] = r[2][lvl].numpy()
                out[f's{seed}_{name}_order{lvl}'] = np.asarray(r[3][lvl], dtype=np.int32)
                out[f's{seed}_{name}_ins{lvl}'] = np.packbits(r[0][lvl].numpy(), axis=-1)
                out[f's{seed}_{name}_insshape{lvl}'] = np.asarray(r[0][lvl].shape)
    np.savez_compressed(OUT, **out)
    print('oracle == reference on', len(out), 'arrays;', os.path.getsize(OUT) // 1024, 'KiB ->', OUT)


if __name__ == '__main__':
    main()
