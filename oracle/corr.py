"""TEST INFRASTRUCTURE (oracle): CPU restatement of DiscoBox's semantic-correspondence path
(mmdet/models/dense_heads/discobox_head.py): ``relu_and_l2_norm_feat`` (:16-20), ``ObjectElements`` / ``ObjectFactory`` /
``ObjectQueues`` (:23-226: the per-class ring buffer of RoI features / masks / boxes and the retrieval of similar objects),
``SemanticCorrSolver.solve`` with ``pass_message`` (:347-411), ``superres_T`` (:851-865) and the transfer lines of
``corr_loss`` (:1080-1096).  Plain torch on CPU, float32 like the reference.  Pinned on the reference's own classes /
methods (AST-extracted, ``oracle/make_golden_corr.py``) through ``tests/golden/corr.npz``.
Third party and absent: ``mmcv.ops.RoIAlign`` (mmcv-full 1.3.17-1.6.0; ``aligned=True``, adaptive sampling) -- the RoI
features and masks are INPUTS of everything restated here."""
import torch
import torch.nn.functional as F


def relu_and_l2_norm_feat(feat, dim=1):
    """:16-20 (without the in-place relu on the caller's tensor)."""
    feat = F.relu(feat)
    norm = ((feat ** 2).sum(dim=dim, keepdim=True) + 1e-6) ** 0.5
    return feat / (norm + 1e-6)


def pass_message(T, h, w):
    """:347-366.  T [K,P,P] -> [K,P,P]: every entry becomes the mean over the (dy,dx) shifts of BOTH cells that stay inside
    the h x w grid.  Zero padding adds exact zeros in the reference's accumulation order (dx outer, dy inner)."""
    K = T.shape[0]
    T5 = T.reshape(K, h, w, h, w)
    Tp = F.pad(T5, (1, 1, 1, 1, 1, 1, 1, 1))
    Op = F.pad(torch.ones_like(T5), (1, 1, 1, 1, 1, 1, 1, 1))
    acc, cnt = torch.zeros_like(T5), torch.zeros_like(T5)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            sl = (slice(None), slice(1 - dy, 1 - dy + h), slice(1 - dx, 1 - dx + w), slice(1 - dy, 1 - dy + h),
                  slice(1 - dx, 1 - dx + w))
            acc = acc + Tp[sl]
            cnt = cnt + Op[sl]
    return (acc / cnt).reshape(K, h * w, h * w)


def window_mask(h, w, dist_kernel):
    """:393-395: [1,P,P], 1 where two cells are within dist_kernel // 2 of each other in both coordinates."""
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
    yy, xx = yy.reshape(-1), xx.reshape(-1)
    r = dist_kernel // 2
    near = ((yy[:, None] - yy[None, :]).abs() <= r) & ((xx[:, None] - xx[None, :]).abs() <= r)
    return near.float()[None]


def cosine_table(f0, f1):
    """:386-391: f0 [1,C,h,w] (query), f1 [K,C,h,w] (retrieved) -> Cu [K,P,P]."""
    a = f0.reshape(f0.shape[0], f0.shape[1], -1).transpose(2, 1)
    b = f1.reshape(f1.shape[0], f1.shape[1], -1)
    a = a / (torch.norm(a, p=2, dim=2, keepdim=True) + 1e-4)
    b = b / (torch.norm(b, p=2, dim=1, keepdim=True) + 1e-4)
    return torch.matmul(a, b)


def solve_votes(Cu, h, w, dist_kernel, num_iter, num_smooth):
    """:396-410 given Cu: the regularised table C."""
    C = Cu * window_mask(h, w, dist_kernel)
    for _ in range(num_iter):
        votes = C
        for _ in range(num_smooth):
            votes = pass_message(votes, h, w)
            votes = votes / (votes.sum(2, keepdim=True) + 1e-4)
        C = Cu + votes
        C = C / (C.sum(2, keepdim=True) + 1e-4)
    return C


def superres(T, h, w, Hm, Wm):
    """:851-865: [K,P,P] -> [K,M,M]: bilinear over the target cell, then over the source cell, times P / M."""
    K, P, M = T.shape[0], h * w, Hm * Wm
    t = F.interpolate(T.reshape(K, P, h, w), (Hm, Wm), mode='bilinear', align_corners=False)          # target cell
    t = F.interpolate(t.reshape(K, 1, h, w, M), (Hm, Wm, M), mode='trilinear', align_corners=False)   # source cell
    return t.reshape(K, M, M) * (1.0 * h * w / Hm / Wm)


def transfer(T, Cu, m0, m1, h, w):
    """:1080-1096.  T, Cu [K,P,P]; m0 [1,Hm,Wm] query RoI mask; m1 [K,Hm,Wm] masks of the retrieved objects.  Returns
    (assignment [K*P] int64 = T.argmax(2), fg_ci [Hm,Wm], bg_ci [Hm,Wm])."""
    K, Hm, Wm = m1.shape
    assignment = T.argmax(2).reshape(-1)
    T2 = T * F.softmax(Cu, 2)
    T2 = T2 / (T2.sum(2, keepdim=True) + 1e-5)
    Ts = superres(T2, h, w, Hm, Wm)
    fg_mask = m0.reshape(1, -1, 1) * m1.reshape(K, 1, -1)                                              # :378
    bg_mask = (1 - m0).reshape(1, -1, 1) * (1 - m1).reshape(K, 1, -1)                                  # :379
    fg = torch.matmul(Ts * (fg_mask > 0.5).float(), torch.clamp(m1, min=0.1, max=0.9).reshape(K, -1, 1)).mean(0)
    bg = torch.matmul(Ts * (bg_mask > 0.5).float(), torch.clamp(1 - m1, min=0.1, max=0.9).reshape(K, -1, 1)).mean(0)
    return assignment, fg.reshape(Hm, Wm), bg.reshape(Hm, Wm)


def nce_loss(Cu, assignment):
    """:1080-1084: cross entropy of softmax(Cu) (sic: a softmax in front of CrossEntropyLoss) against the assignment."""
    logits = F.softmax(Cu.float(), 2).reshape(-1, Cu.shape[2])
    return F.cross_entropy(logits, assignment)


# -----------------------------------------------------------------------------------------------------------------------
# the object bank (:23-226)
# -----------------------------------------------------------------------------------------------------------------------
class Bank:
    """One class's ring buffer (ObjectElements :69-95 as filled by ObjectFactory.create_queue_by_one :47-66 and
    ObjectQueues.append :146-172)."""

    def __init__(self, size, n_channel, feat_size, mask_size):
        self.mask = torch.zeros(size, mask_size, mask_size)
        self.feature = torch.zeros(size, n_channel, feat_size, feat_size)
        self.box = torch.zeros(size, 4)
        self.ptr = 0


class Queues:
    def __init__(self, num_class, len_queue, fg_iou_thresh, bg_iou_thresh, ratio_range, appear_thresh, max_retrieval_objs):
        self.banks = [None] * num_class
        self.len_queue = len_queue
        self.fg, self.bg, self.app = fg_iou_thresh, bg_iou_thresh, appear_thresh
        self.ratio_range = ratio_range
        self.max_objs = max_retrieval_objs

    def append(self, cls, idx, feature, mask, box):
        """:146-172.  True when a new bank was created."""
        created = self.banks[cls] is None
        if created:
            self.banks[cls] = Bank(self.len_queue, feature.shape[1], feature.shape[2], mask.shape[1])
        b = self.banks[cls]
        b.feature[b.ptr] = feature[idx]
        b.mask[b.ptr] = mask[idx]
        b.box[b.ptr] = box[idx]
        b.ptr = (b.ptr + 1) % self.len_queue
        return created

    def similar(self, cls, q_mask, q_feature, q_box):
        """get_similar_obj (:205-226) for one query (q_mask [1,Hm,Wm], q_feature [1,C,h,w], q_box [1,4]): the indices
        (ascending, at most max_retrieval_objs) of the bank entries that pass the four tests, or None without a bank."""
        b = self.banks[cls]
        if b is None:
            return None
        A, Bm = q_mask, b.mask
        fg_iou = (A * Bm).sum([1, 2]) / ((A + Bm) >= 1).float().sum([1, 2])                           # :174-180
        bg_iou = ((1 - A) * (1 - Bm)).sum([1, 2]) / ((2 - A - Bm) >= 1).float().sum([1, 2])           # :182-186
        fs = q_feature.shape[2:]
        a_small = F.interpolate(A[:, None], fs, mode='bilinear', align_corners=False)[:, 0]           # :188-199
        b_small = F.interpolate(Bm[:, None], fs, mode='bilinear', align_corners=False)[:, 0]
        appear = (q_feature * b.feature * a_small[:, None] * b_small[:, None]).sum([1, 2, 3]) / ((a_small * b_small).sum([1, 2]) + 1e-6)
        ratio = lambda x: (x[:, 2] - x[:, 0]) / (x[:, 3] - x[:, 1] + 1e-5)                            # :103-105
        rr = (ratio(q_box)[:, None] / ratio(b.box)[None, :])[0]                                       # :201-205
        keep = (fg_iou > self.fg) & (bg_iou > self.bg) & (appear > self.app) & (rr >= self.ratio_range[0]) & (rr <= self.ratio_range[1])
        return torch.where(keep)[0][:self.max_objs]


def corr_objects(queues, roi_s_feat, roi_t_feat, roi_s_mask, roi_t_mask, boxes, kernel_labels, iiu, solver_cfg, min_size,
                 min_objs=5, state=None):
    """The per-object loop of corr_loss (:1056-1125) on CPU tensors with the pieces above; boxes integer valued [n,4].
    Returns (sum of the InfoNCE terms, number of terms); fills iiu [2n,H,W] and updates `queues`."""
    n = roi_s_feat.shape[0]
    h, w = roi_s_feat.shape[2:]
    loss, num = roi_s_feat.new_zeros(()), 0
    state = {'first': True} if state is None else state       # the head's query holder exists after its very first object
    for i in range(n):
        x1, y1, x2, y2 = [int(v) for v in boxes[i]]
        cls = int(kernel_labels[i])
        qf = roi_s_feat[i:i + 1].detach()
        if state['first']:                          # ObjectFactory.create_one re-normalises the very first query (:43)
            qf, state['first'] = relu_and_l2_norm_feat(qf), False
        idx = queues.similar(cls, roi_s_mask[i:i + 1], qf, boxes[i:i + 1])
        if idx is not None and len(idx) >= min_objs:
            bank = queues.banks[cls]
            Cu = cosine_table(roi_s_feat[i:i + 1], bank.feature[idx])
            T = solve_votes(Cu.detach(), h, w, solver_cfg['dist_kernel'], solver_cfg['num_iter'], solver_cfg['num_smooth_iter'])
            asg, fg, bg = transfer(T, Cu.detach(), roi_s_mask[i:i + 1], bank.mask[idx], h, w)
            loss = loss + nce_loss(Cu, asg)
            num += 1
            if y2 > y1 and x2 > x1:
                iiu[2 * i, y1:y2, x1:x2] = F.interpolate(bg[None, None], (y2 - y1, x2 - x1), mode='bilinear', align_corners=False)[0, 0]
                iiu[2 * i + 1, y1:y2, x1:x2] = F.interpolate(fg[None, None], (y2 - y1, x2 - x1), mode='bilinear', align_corners=False)[0, 0]
        if (x2 - x1) > min_size and (y2 - y1) > min_size:
            queues.append(cls, i, roi_t_feat, roi_t_mask, boxes)
    return loss, num


def corr_loss_levels(queues, s_ins_pred_list, img_ind_list, ins_labels, kernel_label_list, s_feat, t_feat, color_feats,
                     solver_cfg, bank_cfg, mf_cfg, roi_align, state=None):
    """corr_loss :1013-1139 (no independent teacher) on CPU tensors, object by object like the reference.  ``roi_align(x,
    rois, out_size)`` stands for mmcv's RoIAlign (third party).  mf_cfg: kernel_size, theta0, theta1, alpha0, iter, base,
    gamma; the mean field of image b uses color_feats[b:b+1].  Returns (corr_loss / (num + 1e-4), [dice terms per (level,
    image) chunk, concatenated per level])."""
    from oracle import levelset as ol
    total, num, loss_ts = s_feat.new_zeros(()), 0, []
    state = {'first': True} if state is None else state
    B = color_feats.shape[0]
    kernels = [ol.meanfield_kernel(color_feats[b:b + 1], mf_cfg['kernel_size'], mf_cfg['theta0'], mf_cfg['theta1'], mf_cfg['alpha0'])
               for b in range(B)]
    for s_in, img_inds, target, klabels in zip(s_ins_pred_list, img_ind_list, ins_labels, kernel_label_list):
        if s_in is None:
            continue
        s = torch.sigmoid(s_in)
        keep = torch.tensor([bool(t.sum()) for t in target])
        if keep.sum() == 0:
            continue
        s, img_inds, target, klabels = s[keep], img_inds[keep], target[keep], klabels[keep]
        pos = [torch.where(t) for t in target]
        boxes = torch.tensor([[int(p[1].min()), int(p[0].min()), int(p[1].max()) + 1, int(p[0].max()) + 1] for p in pos]).float()
        rois = torch.cat([img_inds.float()[:, None], boxes], 1)
        roi_s_feat = relu_and_l2_norm_feat(roi_align(s_feat, rois, (bank_cfg['feat_height'], bank_cfg['feat_width'])))
        roi_t_feat = relu_and_l2_norm_feat(roi_align(t_feat.detach(), rois, (bank_cfg['feat_height'], bank_cfg['feat_width']))).detach()
        own = torch.cat([torch.arange(len(target)).float()[:, None], boxes], 1)
        roi_mask = roi_align(s.detach()[:, None], own, (bank_cfg['mask_height'], bank_cfg['mask_width']))[:, 0]
        iiu = torch.zeros(2 * len(target), *s.shape[1:])
        l, k = corr_objects(queues, roi_s_feat, roi_t_feat, roi_mask, roi_mask, boxes, klabels, iiu, solver_cfg,
                            bank_cfg['min_size'], state=state)
        total, num = total + l, num + k
        iiu = iiu.reshape(-1, 2, *iiu.shape[1:])
        enlarged = F.max_pool2d(target.float()[:, None], kernel_size=3, stride=1, padding=1)[:, 0]
        chunks = []
        for b in range(B):
            sel = img_inds == b
            if sel.sum() > 0:
                pseudo, _ = ol.meanfield_forward(kernels[b], s[sel][:, None].detach(), target[sel][:, None].float(), mf_cfg['kernel_size'],
                                                 mf_cfg['iter'], mf_cfg['base'], inter=iiu[sel], gamma=mf_cfg['gamma'])
                cropped = s[sel] * enlarged[sel]
                cropped = cropped * mf_cfg['gamma'] + cropped.detach() * (1 - mf_cfg['gamma'])
                chunks.append((sel.nonzero().flatten(), ol.disco_dice_loss(cropped, pseudo)))
        order = torch.cat([c[0] for c in chunks])
        vals = torch.cat([c[1] for c in chunks])
        loss_ts.append(vals[torch.argsort(order)])
    return total / (num + 1e-4), loss_ts


def corr_inputs(s_kernel_preds_raw, s_ins_pred, gt_bbox_list, gt_label_list, gt_mask_list, scale_ranges, strides, seg_num_grids,
                sigma, num_classes, best=True):
    """discobox_head.py:917-1003 (no independent teacher): per-image SOLO targets, the kernels of the covered cells gathered by
    grid_order, one F.conv2d per (level, image) with objects.  gt_mask_list: per image uint8 NUMPY [G,H,W].  Returns
    (s_ins_pred_list, img_ind_list, ins_labels, kernel_label_list); None for a level without objects."""
    from oracle import solo_targets as ost
    fsize = tuple(s_ins_pred.shape[-2:])
    tg = [ost.disco_target_single(b, l, m, fsize, scale_ranges, strides, seg_num_grids, sigma, num_classes, best=best)
          for b, l, m in zip(gt_bbox_list, gt_label_list, gt_mask_list)]
    L_, B = len(seg_num_grids), len(tg)
    ins_labels = [torch.cat([tg[b][0][lv] for b in range(B)], 0) for lv in range(L_)]
    klabels = [torch.cat([tg[b][1][lv].reshape(-1)[tg[b][3][lv]] for b in range(B)], 0) for lv in range(L_)]
    s_list, img_list = [], []
    for lv in range(L_):
        preds, inds = [], []
        for b in range(B):
            kern = s_kernel_preds_raw[lv][b].view(s_kernel_preds_raw[lv].shape[1], -1)[:, tg[b][3][lv]]
            if kern.size(-1) == 0:
                continue
            C, I = kern.shape
            H, W = s_ins_pred.shape[-2:]
            preds.append(F.conv2d(s_ins_pred[b][None], kern.permute(1, 0).view(I, -1, 1, 1), stride=1).view(-1, H, W))
            inds.append(torch.ones(I) * b)
        s_list.append(torch.cat(preds, 0) if preds else None)
        img_list.append(torch.cat(inds, 0) if inds else None)
    return s_list, img_list, ins_labels, klabels
