"""f4: DiscoBox's semantic-correspondence path (mmdet/models/dense_heads/discobox_head.py): the per-class object bank
(``ObjectElements`` / ``ObjectQueues`` :69-226), ``SemanticCorrSolver.solve`` (:369-411) and the per-object body of
``corr_loss`` (:1059-1125: retrieval, solve, InfoNCE term, transfer of the retrieved masks into the query's box, bank update).

What runs where:
  * cosine table ``Cu`` (:386-391): normalise + ``torch.matmul`` ([49,C] x [C,49] per retrieved object: a library GEMM, and the
    only part with a gradient -- to the query's RoI feature);
  * the regularised table ``T`` (:393-410): ONE launch of ``bxs_corr_solve`` (the reference: ~25 slice-assign / element-wise
    kernels per round, 10 rounds, per query object);
  * the transfer (:1086-1096 + ``superres_T`` :851-865): ``bxs_corr_transfer`` -- none of the reference's six [K,784,784]
    intermediates (12 MB each at K = 5) exists;
  * resizing the two 28 x 28 maps to the query's box (:1098-1103): the a18 kernel;
  * the bank: device ring buffers, retrieval tests as a handful of batched reductions; ONE host read per query object (the
    retrieved indices: the reference's control flow -- ``if kobjs['mask'].shape[0] >= 5`` -- is decided on the host there too,
    after ~12 implicit synchronisations).

Third party and absent from the reference checkout: ``mmcv.ops.RoIAlign`` -- the RoI features / masks are inputs here
(``torchvision.ops.roi_align(x, rois, size, 1.0, 0, True)`` is the same operator).
"""
import torch
import torch.nn.functional as F

from ... import _lib as L
from ...ops.resize import bilinear_resize


def relu_and_l2_norm_feat(feat, dim=1):
    """:16-20."""
    feat = F.relu(feat)
    norm = ((feat ** 2).sum(dim=dim, keepdim=True) + 1e-6) ** 0.5
    return feat / (norm + 1e-6)


class ObjectElements:
    """:69-130: ``size`` slots of (mask [ms,ms], feature [C,fs,fs], box [4]) of one category."""

    def __init__(self, size=100, img_size=0, feat_size=28, mask_size=56, n_channel=256, device='cpu', category=None):
        self.mask = torch.zeros(size, mask_size, mask_size, device=device)
        self.feature = torch.zeros(size, n_channel, feat_size, feat_size, device=device)
        self.box = torch.zeros(size, 4, device=device)
        self.img = None
        self.category = int(category)
        self.ptr = 0

    def get_category(self):
        return self.category

    def get_feature(self):
        return self.feature

    def get_mask(self):
        return self.mask

    def get_box(self):
        return self.box

    def get_ratio(self):
        b = self.box
        return (b[:, 2] - b[:, 0]) / (b[:, 3] - b[:, 1] + 1e-5)

    def __len__(self):
        return len(self.feature)

    def __getitem__(self, idx):
        if isinstance(idx, int):
            idx = slice(idx, idx + 1)
        elif torch.is_tensor(idx):
            idx = idx.to(self.mask.device).long()
        return dict(img=None, mask=self.mask[idx], feature=self.feature[idx], box=self.box[idx], category=self.category)


def create_one(mask, feature, box, category):
    """ObjectFactory.create_one (:26-44): a one-slot holder of the query; the feature is re-normalised like there."""
    q = ObjectElements(size=1, feat_size=feature.shape[2], mask_size=mask.shape[1], n_channel=feature.shape[1],
                       device=mask.device, category=category)
    q.mask[...] = mask
    q.feature[...] = relu_and_l2_norm_feat(feature[0:1])
    q.box[...] = box
    return q


class ObjectQueues:
    """:132-226, same constructor arguments.  All banks live on the device (the reference spills to host memory after
    ``num_gpu_bank`` classes: 80 classes x 100 slots x (128 x 7 x 7 + 28 x 28) floats = 226 MB -- nothing on a 180 GB part)."""

    def __init__(self, num_class, len_queue, fg_iou_thresh, bg_iou_thresh, ratio_range, appear_thresh, max_retrieval_objs):
        self.num_class = num_class
        self.queues = [None] * num_class
        self.len_queue = len_queue
        self.fg_iou_thresh = fg_iou_thresh
        self.bg_iou_thresh = bg_iou_thresh
        self.appear_thresh = appear_thresh
        self.ratio_range = ratio_range
        self.max_retrieval_objs = max_retrieval_objs

    @torch.no_grad()
    def append(self, class_idx, idx, feature, mask, box, img=None, device=None):
        """:146-172: entry ``idx`` of the three batched tensors goes to the class's write pointer.  True when the bank was
        created by this call."""
        created = self.queues[class_idx] is None
        if created:
            self.queues[class_idx] = ObjectElements(size=self.len_queue, feat_size=feature.shape[2], mask_size=mask.shape[1],
                                                    n_channel=feature.shape[1], device=device or mask.device,
                                                    category=class_idx)
        q = self.queues[class_idx]
        q.feature[q.ptr] = feature[idx]
        q.mask[q.ptr] = mask[idx]
        q.box[q.ptr] = box[idx]
        q.ptr = (q.ptr + 1) % self.len_queue
        return created

    @torch.no_grad()
    def similar_indices(self, qobj):
        """The slots of the query's class that pass the four tests of get_similar_obj (:205-221), ascending, at most
        ``max_retrieval_objs`` -- as a device tensor (no synchronisation here), or None without a bank."""
        k = self.queues[qobj.get_category()]
        if k is None:
            return None
        A, B = qobj.get_mask(), k.get_mask().to(qobj.get_mask())
        fg = (A * B).sum([1, 2]) / ((A + B) >= 1).float().sum([1, 2])                                        # :174-180
        bg = ((1 - A) * (1 - B)).sum([1, 2]) / ((2 - A - B) >= 1).float().sum([1, 2])                        # :182-186
        f0, f1 = qobj.get_feature(), k.get_feature().to(qobj.get_feature())
        a = F.interpolate(A[:, None], f0.shape[2:], mode='bilinear', align_corners=False)[:, 0]              # :188-199
        b = F.interpolate(B[:, None], f1.shape[2:], mode='bilinear', align_corners=False)[:, 0]
        appear = (f0 * f1 * a[:, None] * b[:, None]).sum([1, 2, 3]) / ((a * b).sum([1, 2]) + 1e-6)
        ratio = (qobj.get_ratio()[:, None] / k.get_ratio()[None, :].to(A))[0]                                # :201-205
        keep = ((fg > self.fg_iou_thresh) & (bg > self.bg_iou_thresh) & (appear > self.appear_thresh) &
                (ratio >= self.ratio_range[0]) & (ratio <= self.ratio_range[1]))
        return torch.where(keep)[0][:self.max_retrieval_objs]

    def get_similar_obj(self, qobj):
        """:205-226: dict(mask, feature, box, category) of the retrieved slots, or None."""
        idx = self.similar_indices(qobj)
        return None if idx is None else self.queues[qobj.get_category()][idx]


class SemanticCorrSolver:
    """:229-411, same constructor arguments (the optimal-transport and Hough-space helpers of the reference class are dead
    code on this path and are not rebuilt)."""

    def __init__(self, exp, eps, gaussian_filter_size, low_score, num_iter, num_smooth_iter, dist_kernel):
        self.exp = exp
        self.eps = eps
        self.gaussian_filter_size = gaussian_filter_size
        self.low_score = low_score
        self.num_iter = num_iter
        self.num_smooth_iter = num_smooth_iter
        self.dist_kernel = dist_kernel

    def cosine_table(self, f0, f1):
        """:386-391: f0 [1,C,h,w], f1 [K,C,h,w] -> Cu [K,P,P] (differentiable w.r.t. f0)."""
        a = f0.float().reshape(f0.shape[0], f0.shape[1], -1).transpose(2, 1)
        b = f1.float().reshape(f1.shape[0], f1.shape[1], -1)
        a = a / (torch.norm(a, p=2, dim=2, keepdim=True) + 1e-4)
        b = b / (torch.norm(b, p=2, dim=1, keepdim=True) + 1e-4)
        with torch.autocast('cuda', enabled=False):
            return torch.matmul(a, b)

    @torch.no_grad()
    def votes(self, Cu, h, w):
        """:393-410: the regularised table for a given Cu [K,P,P] (one launch).  No gradient: the reference uses T only
        through argmax and inside no_grad (:1081,1086-1088)."""
        Cu = Cu.detach().contiguous().float()
        L.require_cuda(Cu)
        T = torch.empty_like(Cu)
        with torch.cuda.device(Cu.device):
            L.check(L.lib().bxs_corr_solve(L.ptr(Cu), L.ptr(T), Cu.shape[0], h, w, int(self.dist_kernel), int(self.num_iter),
                                           int(self.num_smooth_iter), L.stream()), 'corr_solve')
        return T

    def solve(self, qobjs, kobjs, f0, return_masks=False):
        """The reference's signature.  Returns (Cu, C, fg_mask, bg_mask); the two [K,M,M] mask products (:378-379) are only
        built on request -- ``transfer`` does not need them."""
        m0 = qobjs.mask.float()
        f1 = kobjs['feature'].to(m0).float()
        m1 = kobjs['mask'].to(m0).float()
        Cu = self.cosine_table(f0, f1)
        C = self.votes(Cu, f0.shape[2], f0.shape[3])
        if not return_masks:
            return Cu, C, None, None
        fg_mask = m0.reshape(m0.shape[0], -1, 1) * m1.reshape(m1.shape[0], 1, -1)
        bg_mask = (1 - m0).reshape(m0.shape[0], -1, 1) * (1 - m1).reshape(m1.shape[0], 1, -1)
        return Cu, C, fg_mask, bg_mask

    @torch.no_grad()
    def transfer(self, T, Cu, qmask, kmask, h, w):
        """:1086-1096 + superres_T: T, Cu [K,P,P], qmask [1,Hm,Wm], kmask [K,Hm,Wm] -> (fg_ci, bg_ci) [Hm,Wm]."""
        T, Cu = T.contiguous().float(), Cu.detach().contiguous().float()
        m0, m1 = qmask.detach().contiguous().float(), kmask.detach().contiguous().float()
        L.require_cuda(T, Cu, m0, m1)
        K, Hm, Wm = m1.shape
        fg = torch.empty((Hm, Wm), dtype=torch.float32, device=T.device)
        bg = torch.empty_like(fg)
        lib = L.lib()
        ws = torch.empty(lib.bxs_corr_transfer_workspace_bytes(K, Hm, Wm), dtype=torch.uint8, device=T.device)
        with torch.cuda.device(T.device):
            L.check(lib.bxs_corr_transfer(L.ptr(T), L.ptr(Cu), L.ptr(m0), L.ptr(m1), L.ptr(fg), L.ptr(bg), L.ptr(ws), K, h, w,
                                          Hm, Wm, L.stream()), 'corr_transfer')
        return fg, bg


def corr_objects(solver, queues, qobj, roi_s_feat, roi_t_feat, roi_s_mask, roi_t_mask, boxes, kernel_labels, iiu,
                 objbank_min_size, min_objs=5):
    """The per-object loop of ``corr_loss`` (:1056-1125) for the n objects of one level.

    roi_s_feat [n,C,h,w] (student RoI features, normalised, carries the gradient), roi_t_feat / roi_s_mask / roi_t_mask
    (detached RoI tensors), boxes [n,4] (x1,y1,x2,y2 on the mask-feature grid, integer valued), kernel_labels [n], iiu
    [2n,H,W] zeros.  Returns (sum of the InfoNCE terms, number of terms, qobj); fills ``iiu`` and updates the bank.
    ``qobj`` is the head's one-slot query holder (None on the first call, :1060-1066)."""
    n = roi_s_feat.shape[0]
    box_l = [[int(v) for v in b] for b in boxes.tolist()]                 # ONE host read for all boxes of the level
    labels = [int(v) for v in kernel_labels.tolist()]
    loss = roi_s_feat.new_zeros(())
    num = 0
    h, w = roi_s_feat.shape[2:]
    for i in range(n):
        x1, y1, x2, y2 = box_l[i]
        if qobj is None:
            qobj = create_one(roi_s_mask[i:i + 1].detach(), roi_s_feat[i:i + 1].detach(), boxes[i:i + 1].detach(), labels[i])
        else:                                                             # :1067-1071 (no re-normalisation there)
            qobj.mask[...] = roi_s_mask[i:i + 1].detach()
            qobj.feature[...] = roi_s_feat[i:i + 1].detach()
            qobj.box[...] = boxes[i:i + 1].detach()
            qobj.category = labels[i]
        idx = queues.similar_indices(qobj)
        if idx is not None and idx.numel() >= min_objs:                   # the host decision of :1075 (one read)
            kobjs = queues.queues[labels[i]][idx]
            Cu, T, _, _ = solver.solve(qobj, kobjs, roi_s_feat[i:i + 1])
            assignment = T.argmax(2).reshape(-1)                          # :1081-1084
            loss = loss + F.cross_entropy(F.softmax(Cu.float(), 2).reshape(-1, Cu.shape[2]), assignment)
            num += 1
            fg, bg = solver.transfer(T, Cu, qobj.mask, kobjs['mask'], h, w)
            if y2 > y1 and x2 > x1:                                       # :1098-1107
                iiu[2 * i, y1:y2, x1:x2] = bilinear_resize(bg[None, None], (y2 - y1, x2 - x1))[0, 0]
                iiu[2 * i + 1, y1:y2, x1:x2] = bilinear_resize(fg[None, None], (y2 - y1, x2 - x1))[0, 0]
        if (x2 - x1) > objbank_min_size and (y2 - y1) > objbank_min_size:   # :1056-1057, 1113-1124
            queues.append(labels[i], i, roi_t_feat, roi_t_mask, boxes.detach())
    return loss, num, qobj


def _torchvision_roi_align(out_size):
    """mmcv.ops.RoIAlign(out_size) with its defaults (spatial_scale 1, adaptive sampling, aligned=True) -- third party to the
    reference; torchvision ships the same operator."""
    from torchvision.ops import roi_align

    def run(x, rois):
        return roi_align(x, rois, out_size, 1.0, 0, True)
    return run


def mask_boxes(target):
    """The tight box of every non-empty target mask, (min_x, min_y, max_x + 1, max_y + 1) as float (:1029-1035, where the
    reference loops over the objects with four ``.min()/.max()`` host reads each): four reductions, no synchronisation."""
    n, H, W = target.shape
    on = target > 0
    rows, cols = on.any(2), on.any(1)                                        # [n,H], [n,W]
    ys = torch.arange(H, device=target.device)
    xs = torch.arange(W, device=target.device)
    min_y = torch.where(rows, ys, ys.new_full((), H)).amin(1)
    max_y = torch.where(rows, ys, ys.new_full((), -1)).amax(1) + 1
    min_x = torch.where(cols, xs, xs.new_full((), W)).amin(1)
    max_x = torch.where(cols, xs, xs.new_full((), -1)).amax(1) + 1
    return torch.stack([min_x, min_y, max_x, max_y], 1).float()


class DiscoCorr:
    """The state and the per-level body of ``DiscoBoxSOLOv2Head.corr_loss`` (:900-1139) after the dynamic convolutions: per
    level, drop the empty targets, box every target, RoI-align the student / teacher features and masks, run the per-object
    loop (``corr_objects``), then the mean field WITH the transferred inter-image maps and the dice against its pseudo labels.
    Built from the head's ``loss_corr`` config dict (:720-749)."""

    def __init__(self, num_classes, loss_corr, roi_align=_torchvision_roi_align):
        bank = loss_corr['obj_bank']
        self.solver = SemanticCorrSolver(loss_corr['corr_exp'], loss_corr['corr_eps'], loss_corr['gaussian_filter_size'],
                                         loss_corr['low_score'], loss_corr['corr_num_iter'], loss_corr['corr_num_smooth_iter'],
                                         dist_kernel=loss_corr['dist_kernel'])
        self.object_queues = ObjectQueues(num_class=num_classes, len_queue=bank['len_object_queues'],
                                          fg_iou_thresh=bank['fg_iou_thresh'], bg_iou_thresh=bank['bg_iou_thresh'],
                                          ratio_range=bank['ratio_range'], appear_thresh=bank['appear_thresh'],
                                          max_retrieval_objs=bank['max_retrieval_objs'])
        self.feat_roi_align = roi_align((bank['feat_height'], bank['feat_width']))
        self.mask_roi_align = roi_align((bank['mask_height'], bank['mask_width']))
        self.objbank_min_size = bank['min_size']
        self.min_objs = 5                                                    # the literal of :1075
        self.corr_loss_weight = loss_corr['loss_weight']
        self.qobj = None

    def levels(self, s_ins_pred_list, t_ins_pred_list, img_ind_list, ins_labels, kernel_label_list, s_feat, t_feat, mean_field,
               use_ind_teacher=False):
        """Per level l: s_ins_pred_list[l] [n,H,W] student mask LOGITS (None: no object), t_ins_pred_list[l] the teacher's
        (ignored unless use_ind_teacher), img_ind_list[l] [n], ins_labels[l] [n,H,W] targets, kernel_label_list[l] [n];
        s_feat / t_feat [B,C,H,W]; mean_field: ONE ``MeanField`` over the batch's colour features (``obj_img`` selects the
        image; the reference keeps one module per image).  Returns (corr_loss / (num + 1e-4), [per-level dice terms [n_l]])."""
        total = s_feat.new_zeros(())
        num = 0
        loss_ts = []
        for s_in, t_in, img_inds, target, klabels in zip(s_ins_pred_list, t_ins_pred_list, img_ind_list, ins_labels,
                                                         kernel_label_list):
            if s_in is None or s_in.shape[0] == 0:
                continue
            keep = (target.flatten(1) != 0).any(1).nonzero().flatten()       # remove all-zero targets (:1024-1028); one sync
            if keep.numel() == 0:
                continue
            s = torch.sigmoid(s_in).index_select(0, keep)
            t = torch.sigmoid(t_in).index_select(0, keep) if use_ind_teacher else s
            img_inds = img_inds.to(s.device).index_select(0, keep)
            target = target.index_select(0, keep)
            klabels = klabels.index_select(0, keep)
            n = s.shape[0]
            boxes = mask_boxes(target)
            rois = torch.cat([img_inds.to(s_feat)[:, None], boxes], 1)
            roi_s_feat = relu_and_l2_norm_feat(self.feat_roi_align(s_feat, rois))
            with torch.no_grad():
                roi_t_feat = relu_and_l2_norm_feat(self.feat_roi_align(t_feat.detach(), rois))
                own = torch.cat([torch.arange(n, device=s.device).to(s)[:, None], boxes], 1)
                roi_s_mask = self.mask_roi_align(s.detach()[:, None], own)[:, 0]
                roi_t_mask = self.mask_roi_align(t.detach()[:, None], own)[:, 0]
                iiu = s.new_zeros((2 * n,) + tuple(s.shape[1:]))
            lvl_loss, lvl_num, self.qobj = corr_objects(self.solver, self.object_queues, self.qobj, roi_s_feat, roi_t_feat,
                                                        roi_s_mask, roi_t_mask, boxes, klabels, iiu, self.objbank_min_size,
                                                        self.min_objs)
            total = total + lvl_loss
            num += lvl_num
            iiu = iiu.view(n, 2, *iiu.shape[1:])
            tf = target.float()
            enlarged = F.max_pool2d(tf[:, None], kernel_size=3, stride=1, padding=1)[:, 0]                  # :1128
            pseudo, _ = mean_field(((t + s) / 2)[:, None], tf[:, None], iiu, obj_img=img_inds)              # :1130-1132
            cropped = s * enlarged
            cropped = cropped * mean_field.gamma + cropped.detach() * (1 - mean_field.gamma)                # :1134
            x, p = cropped.flatten(1).float(), pseudo.flatten(1).float()                                    # dice_loss :542-550
            loss_ts.append(1 - 2 * (x * p).sum(1) / ((x * x).sum(1) + 0.001 + (p * p).sum(1) + 0.001))
        return total / (num + 1e-4), loss_ts
