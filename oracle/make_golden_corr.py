"""Golden vectors for DiscoBox's semantic-correspondence path, minted from the REFERENCE's own code (authoring container
only): ``relu_and_l2_norm_feat``, ``ObjectFactory``, ``ObjectElements``, ``ObjectQueues``, ``SemanticCorrSolver`` and
``DiscoBoxSOLOv2Head.superres_T`` are AST-extracted from mmdet/models/dense_heads/discobox_head.py and run on seeded inputs (CPU
tensors); the six transfer lines of ``corr_loss`` (:1080-1096), which are inline in a 240-line method, are evaluated on the
reference's own ``solve`` / ``superres_T`` outputs.  Asserts the oracle restatement (oracle/corr.py) == the reference, writes
tests/golden/corr.npz.      python -m oracle.make_golden_corr"""
import ast
import math
import os
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import corr as oc

REF = os.environ.get('BXS_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'corr.npz')
# configs/discobox/discobox_solov2_coco_r50_fpn_3x.py:65-93
SOLVER = dict(exp=1.0, eps=0.05, gaussian_filter_size=3, low_score=0.3, num_iter=10, num_smooth_iter=1, dist_kernel=9)
BANK = dict(len_queue=6, fg_iou_thresh=0.7, bg_iou_thresh=0.7, ratio_range=[0.9, 1.2], appear_thresh=0.7, max_retrieval_objs=5)
FEAT, MASK, CH = 7, 28, 16


class _NoAutocast:
    def __init__(self, enabled=True):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def reference_code():
    src = open(os.path.join(REF, 'mmdet/models/dense_heads/discobox_head.py')).read()
    tree = ast.parse(src)
    want = ('relu_and_l2_norm_feat', 'ObjectFactory', 'ObjectElements', 'ObjectQueues', 'SemanticCorrSolver')
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in want]
    head = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'DiscoBoxSOLOv2Head'][0]
    body += [n for n in head.body if isinstance(n, ast.FunctionDef) and n.name == 'superres_T']
    g = dict(torch=torch, F=F, nn=nn, math=math, np=np, autocast=_NoAutocast)
    exec(compile(ast.Module(body=body, type_ignores=[]), 'discobox_head.py', 'exec'), g)
    return g


def blob(gen, n, size, cy, cx, ry, rx, soft=0.15):
    """n soft elliptical masks in [0,1] (what RoIAlign of a sigmoid mask looks like), jittered around one shape."""
    yy, xx = torch.meshgrid(torch.arange(size, dtype=torch.float32), torch.arange(size, dtype=torch.float32), indexing='ij')
    out = []
    for _ in range(n):
        j = torch.randn(4, generator=gen) * 0.6
        d = ((yy - cy - j[0]) / (ry + j[2])) ** 2 + ((xx - cx - j[1]) / (rx + j[3])) ** 2
        out.append(torch.sigmoid((1 - d) / soft))
    return torch.stack(out)


def case(seed, K=5, h=FEAT, w=FEAT, Hm=MASK, C=CH):
    gen = torch.Generator().manual_seed(seed)
    base = torch.randn(1, C, h, w, generator=gen)
    f0 = oc.relu_and_l2_norm_feat(base + 0.3 * torch.randn(1, C, h, w, generator=gen))
    f1 = oc.relu_and_l2_norm_feat(base + 0.5 * torch.randn(K, C, h, w, generator=gen))
    m0 = blob(gen, 1, Hm, Hm / 2, Hm / 2, Hm * 0.33, Hm * 0.28)
    m1 = blob(gen, K, Hm, Hm / 2, Hm / 2, Hm * 0.33, Hm * 0.28)
    return f0, f1, m0, m1


def bank_case(seed, n=9, C=CH):
    gen = torch.Generator().manual_seed(100 + seed)
    base = torch.randn(1, C, FEAT, FEAT, generator=gen)
    feats = oc.relu_and_l2_norm_feat(base + 0.25 * torch.randn(n, C, FEAT, FEAT, generator=gen))
    masks = blob(gen, n, MASK, MASK / 2, MASK / 2, MASK * 0.36, MASK * 0.3)
    wh = 40 + torch.rand(n, 2, generator=gen) * 6
    xy = torch.rand(n, 2, generator=gen) * 50
    boxes = torch.cat([xy, xy + wh], 1)
    masks[3] = blob(gen, 1, MASK, 6.0, 6.0, 4.0, 4.0)[0]                # a different shape: fails the IoU tests
    boxes[5, 3] = boxes[5, 1] + 2 * wh[5, 1]                            # a different aspect ratio: fails the ratio test
    return feats, masks, boxes


def main():
    g = reference_code()
    out = {}
    solver = g['SemanticCorrSolver'](**SOLVER)
    me = types.SimpleNamespace(corr_feat_height=FEAT, corr_feat_width=FEAT, corr_mask_height=MASK, corr_mask_width=MASK)
    for seed in (0, 1, 2):
        f0, f1, m0, m1 = case(seed)
        q = types.SimpleNamespace(mask=m0)
        Cu, T, fg_mask, bg_mask = solver.solve(q, dict(feature=f1, mask=m1), f0)
        o_Cu = oc.cosine_table(f0, f1)
        o_T = oc.solve_votes(o_Cu, FEAT, FEAT, SOLVER['dist_kernel'], SOLVER['num_iter'], SOLVER['num_smooth_iter'])
        assert torch.equal(Cu, o_Cu), 'oracle cosine table != reference'
        assert (T - o_T).abs().max() <= 2e-6 * T.abs().max(), 'oracle solve != reference'
        # corr_loss :1080-1096 on the reference's own outputs
        assignment = T.argmax(2).reshape(-1)
        Cs = F.softmax(Cu.float(), 2).reshape(-1, Cu.shape[2])
        nce = nn.CrossEntropyLoss()(Cs, assignment)
        with torch.no_grad():
            T2 = T * Cs.reshape(T.shape)
        T2 = T2 / (T2.sum(2, keepdim=True) + 1e-5)
        Ts = g['superres_T'](me, T2)
        fg = torch.matmul(Ts * (fg_mask > 0.5).float(), torch.clamp(m1, min=0.1, max=0.9).reshape(Ts.shape[0], Ts.shape[2], 1)).mean(0).reshape(MASK, MASK)
        bg = torch.matmul(Ts * (bg_mask > 0.5).float(), torch.clamp(1 - m1, min=0.1, max=0.9).reshape(Ts.shape[0], Ts.shape[2], 1)).mean(0).reshape(MASK, MASK)
        o_asg, o_fg, o_bg = oc.transfer(T, Cu, m0, m1, FEAT, FEAT)
        assert torch.equal(o_asg, assignment)
        assert (o_fg - fg).abs().max() <= 1e-6 and (o_bg - bg).abs().max() <= 1e-6, 'oracle transfer != reference'
        assert abs(float(oc.nce_loss(Cu, assignment)) - float(nce)) < 1e-6
        assert float(fg.max()) > 0.05 and float(bg.max()) > 0.05
        for name, t in (('Cu', Cu), ('T', T), ('fg', fg), ('bg', bg)):
            out[f's{seed}_{name}'] = t.numpy()
        out[f's{seed}_nce'] = np.asarray(float(nce), dtype=np.float32)
    # the object bank: fill a ring past its length, then retrieve
    for seed in (0, 1):
        feats, masks, boxes = bank_case(seed)
        ref_q = g['ObjectQueues'](num_class=3, **BANK)
        my_q = oc.Queues(num_class=3, **BANK)
        for i in range(feats.shape[0]):
            a = ref_q.append(1, i, feats, masks, boxes, None, device='cpu')
            b = my_q.append(1, i, feats, masks, boxes)
            assert bool(a) == bool(b)
        rb = ref_q.queues[1]
        assert torch.equal(rb.feature, my_q.banks[1].feature) and torch.equal(rb.mask, my_q.banks[1].mask) and rb.ptr == my_q.banks[1].ptr
        gen = torch.Generator().manual_seed(7 + seed)
        qf = oc.relu_and_l2_norm_feat(feats[:1] + 0.1 * torch.randn(1, CH, FEAT, FEAT, generator=gen))
        qobj = g['ObjectFactory'].create_one(mask=masks[7:8].clone(), feature=qf.clone(), box=boxes[7:8].clone(), category=1, img=None)
        got = ref_q.get_similar_obj(qobj)
        idx = my_q.similar(1, masks[7:8], oc.relu_and_l2_norm_feat(qf), boxes[7:8])
        assert torch.equal(got['mask'], my_q.banks[1].mask[idx]) and torch.equal(got['feature'], my_q.banks[1].feature[idx])
        assert 0 < len(idx) <= BANK['max_retrieval_objs'] and len(idx) < BANK['len_queue']
        out[f'b{seed}_idx'] = idx.numpy()
        assert ref_q.get_similar_obj(g['ObjectFactory'].create_one(mask=masks[7:8].clone(), feature=qf.clone(), box=boxes[7:8].clone(), category=2, img=None)) is None
    np.savez_compressed(OUT, **out)
    print('oracle == reference on', len(out), 'arrays;', os.path.getsize(OUT) // 1024, 'KiB ->', OUT)


if __name__ == '__main__':
    main()


def levels_case(seed, B=2, H=48, W=64, C=CH, per_level=(5, 4)):
    """Inputs of the per-level body of corr_loss (oracle.corr.corr_loss_levels / DiscoCorr.levels): two levels of objects of
    one class with similar boxes (so that retrieval succeeds once the bank holds five of them), one object of another class,
    one empty target; logits that are high on an ellipse inside the box (so that the RoI masks have both foreground and
    background: the two IoU tests of the retrieval are 0 / 0 on a mask without background)."""
    gen = torch.Generator().manual_seed(1000 + seed)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing='ij')
    pattern = torch.stack([torch.sin(yy / 3 + c) + torch.cos(xx / 4 + 2 * c) for c in range(C)])          # shared texture
    s_feat = pattern[None] + 0.03 * torch.randn(B, C, H, W, generator=gen)
    t_feat = pattern[None] + 0.03 * torch.randn(B, C, H, W, generator=gen)
    color = torch.randn(B, 3, H, W, generator=gen)
    s_list, img_list, tgt_list, lab_list = [], [], [], []
    for n in per_level:
        tg = torch.zeros(n, H, W, dtype=torch.uint8)
        imgs = torch.randint(0, B, (n,), generator=gen)
        obj = []
        for i in range(n):
            y0, x0 = 8 + int(torch.randint(0, 4, (1,), generator=gen)), 14 + int(torch.randint(0, 4, (1,), generator=gen))
            bh, bw = 24 + int(torch.randint(0, 2, (1,), generator=gen)), 26 + int(torch.randint(0, 2, (1,), generator=gen))
            tg[i, y0:y0 + bh, x0:x0 + bw] = 1
            inside = ((yy - (y0 + bh / 2 - 0.5)) / (0.40 * bh)) ** 2 + ((xx - (x0 + bw / 2 - 0.5)) / (0.40 * bw)) ** 2 < 1
            obj.append(inside)
        logits = (torch.stack(obj).float() * 2 - 1) * 4 + 0.7 * torch.randn(n, H, W, generator=gen)
        labels = torch.full((n,), 1, dtype=torch.int64)
        labels[-1] = 2
        s_list.append(logits)
        img_list.append(imgs)
        tgt_list.append(tg)
        lab_list.append(labels)
    tgt_list[1][0] = 0                                                   # an all-zero target: removed (:1024-1028)
    return s_feat, t_feat, color, s_list, img_list, tgt_list, lab_list
