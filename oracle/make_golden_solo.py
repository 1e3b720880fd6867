"""Golden vectors for the SOLO grid targets, minted from the REFERENCE's own method (authoring container only):
``BoxSOLOv2Head.solo_target_single`` is AST-extracted from mmdet/models/dense_heads/box_solov2_head.py, bound to a stand-in
``self`` holding the five attributes it reads, and run on seeded inputs; ``mmcv.imrescale`` (third party, absent) is the restated
``cv2.resize`` call of oracle/solo_targets.py.  Asserts the oracle restatement == the reference, writes
tests/golden/solo_targets.npz.      python -m oracle.make_golden_solo"""
import ast
import os
import types

import numpy as np
import torch
import torch.nn.functional as F
from scipy import ndimage

from oracle import solo_targets as ost

REF = os.environ.get('BXS_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'solo_targets.npz')
# configs/boxlevelset/box_levelset_coco_r50_fpn_3x.py:23-26 (scale ranges shrunk with the 160x192 test image)
CFG = dict(scale_ranges=((1, 24), (12, 48), (24, 96), (48, 192), (96, 512)), strides=(8, 8, 16, 32, 32),
           seg_num_grids=[40, 36, 24, 16, 12], sigma=0.2, num_classes=80)


def reference_method():
    src = open(os.path.join(REF, 'mmdet/models/dense_heads/box_solov2_head.py')).read()
    cls = [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == 'BoxSOLOv2Head'][0]
    fn = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == 'solo_target_single'][0]
    g = dict(torch=torch, F=F, np=np, ndimage=ndimage, mmcv=types.SimpleNamespace(imrescale=ost.imrescale))
    exec(compile(ast.Module(body=[fn], type_ignores=[]), 'box_solov2_head.py', 'exec'), g)
    return g['solo_target_single']


def case(seed, H=160, W=192, G=9):
    gen = torch.Generator().manual_seed(seed)
    x1 = torch.rand(G, generator=gen) * (W - 24)
    y1 = torch.rand(G, generator=gen) * (H - 24)
    bw = 8 + torch.rand(G, generator=gen) ** 2 * (W - 8)
    bh = 8 + torch.rand(G, generator=gen) ** 2 * (H - 8)
    boxes = torch.stack([x1, y1, torch.minimum(x1 + bw, torch.tensor(float(W))), torch.minimum(y1 + bh, torch.tensor(float(H)))], 1)
    boxes[0] = torch.tensor([3.0, 4.0, 6.5, 7.0])                       # fewer than 10 mask pixels: skipped (:441-442)
    labels = torch.randint(0, 80, (G,), generator=gen)
    masks = np.zeros((G, H, W), np.uint8)
    for g_, b in enumerate(boxes):
        masks[g_, int(b[1]):int(b[3]) + 1, int(b[0]):int(b[2]) + 1] = 1      # box-shaped masks (box supervision)
    masks[1, ::3, :] = 0                                                # and one that is not a rectangle
    img = torch.randn(3, H, W, generator=gen)
    lst = torch.randn(5, H // 4, W // 4, generator=gen)
    fs = [(H // 4, W // 4), (H // 4, W // 4), (H // 8, W // 8), (H // 16, W // 16), (H // 16, W // 16)]     # ins_pred sizes: 2 x (H / stride)
    return boxes, labels, masks, img, lst, fs


def main():
    ref = reference_method()
    me = types.SimpleNamespace(**CFG)
    out = {}
    for seed in (0, 1, 2):
        boxes, labels, masks, img, lst, fs = case(seed)
        r = ref(me, boxes, labels, masks, img, lst, featmap_sizes=fs)
        o = ost.solo_target_single(boxes, labels, masks, img, lst, fs, **CFG)
        for a_list, b_list in zip(r, o):
            for a, b in zip(a_list, b_list):
                assert torch.equal(a, b), 'oracle restatement != reference'
        out[f's{seed}_boxes'] = boxes.numpy(); out[f's{seed}_labels'] = labels.numpy(); out[f's{seed}_masks'] = np.packbits(masks, axis=-1)
        out[f's{seed}_shape'] = np.asarray(masks.shape)
        for lvl in range(5):
            out[f's{seed}_cate{lvl}'] = r[1][lvl].numpy().astype(np.int16)
            out[f's{seed}_ind{lvl}'] = r[2][lvl].numpy()
            pos = r[2][lvl].nonzero().flatten()
            out[f's{seed}_ins{lvl}'] = np.packbits(r[0][lvl][pos].numpy(), axis=-1)       # positives only (the rest is zero)
            out[f's{seed}_insshape{lvl}'] = np.asarray(r[0][lvl][pos].shape)
            assert int(r[0][lvl].sum()) == int(r[0][lvl][pos].sum())
    np.savez_compressed(OUT, **out)
    print('oracle == reference on', len(out), 'arrays;', os.path.getsize(OUT) // 1024, 'KiB ->', OUT)


if __name__ == '__main__':
    main()
