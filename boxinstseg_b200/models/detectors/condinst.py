"""f3: the detector-level glue around the mask-loss path, without host synchronisation.

* ``mask_branch_step`` is the tail of ``CondInst.forward_train`` (mmdet/models/detectors/condinst.py:66-75): instance sampling ->
  dynamic mask head -> mask loss, given what the backbone / FCOS head / mask branch produced.  From ``mask_head(...)`` on,
  every call is issued on the current stream without a device->host read, so that part of a step can be captured in a CUDA
  graph (``tests/test_detector_glue_gpu.py`` captures head forward + loss + backward and replays it).
* ``parse_losses`` is ``BaseDetector._parse_losses`` (mmdet/models/detectors/base.py:176-219) with the same results:
  per key the mean (or the sum of the means of a list), ``loss`` = the sum of the keys containing "loss", every log variable
  averaged over the ranks.  The reference pays one ``all_reduce`` and one ``.item()`` (a device->host round trip that drains
  the stream) PER KEY plus an ``all_reduce`` + host assert for the key count, every iteration; here the log variables are
  stacked into ONE fixed-size tensor (64 slots, the key count in the last one) and reduced by one ``all_reduce`` -- fixed size, so
  that ranks with different key sets still run the same collective and fail with the reference's assertion when the values
  are read, instead of hanging; and the host copy is deferred until somebody reads the values (``LogVars``: one non-blocking
  copy into pinned memory, one event wait) -- the training step itself never waits for the device.
"""
from collections import OrderedDict

import torch
import torch.distributed as dist


class LogVars(OrderedDict):
    """The ``log_vars`` of ``_parse_losses``: an ordered mapping name -> float whose values are fetched from the device
    on first access (all at once).  ``names`` keeps the reference's key order, ``values_device`` is the stacked tensor."""

    def __init__(self, names, values_device, mean_key_count=False):
        super().__init__()
        self.names = list(names)
        self.values_device = values_device
        self._counted = mean_key_count          # the last slot holds the mean over ranks of the number of keys
        self._event = None
        self._host = None
        # (inside a CUDA-graph capture no pinned buffer is allocated and nothing is copied: the values stay on the device and
        # are fetched when -- if ever -- they are read after a replay)
        if values_device.is_cuda and not torch.cuda.is_current_stream_capturing():
            self._host = torch.empty(values_device.shape, dtype=values_device.dtype, device='cpu', pin_memory=True)
            self._host.copy_(values_device, non_blocking=True)
            self._event = torch.cuda.Event()
            self._event.record(torch.cuda.current_stream(values_device.device))
        self._ready = False

    def _materialise(self):
        if self._ready:
            return
        if self._event is not None:
            self._event.synchronize()
            vals = self._host.tolist()
        else:
            vals = self.values_device.tolist()
        if self._counted:
            # base.py:199-207: the mean of the key counts equals the local count on every rank iff all ranks agree
            message = f'len(log_vars): {len(self.names) - 1} keys: ' + ','.join(self.names[:-1])
            assert abs(vals[-1] - (len(self.names) - 1)) < 1e-3, 'loss log variables are different across GPUs!\n' + message
        for k, v in zip(self.names, vals):
            OrderedDict.__setitem__(self, k, v)
        self._ready = True

    def __getitem__(self, key):
        self._materialise()
        return OrderedDict.__getitem__(self, key)

    def __iter__(self):
        return iter(self.names)

    def __len__(self):
        return len(self.names)

    def __contains__(self, key):
        return key in self.names

    def keys(self):
        return list(self.names)

    def items(self):
        self._materialise()
        return [(k, OrderedDict.__getitem__(self, k)) for k in self.names]

    def values(self):
        self._materialise()
        return [OrderedDict.__getitem__(self, k) for k in self.names]


_SLOTS = 64                # fixed size of the reduced buffer: ranks with different key sets still run the SAME collective


def parse_losses(losses, defer=True):
    """-> (loss, log_vars) exactly as BaseDetector._parse_losses (base.py:176-219) computes them.
    ``defer=False`` fills ``log_vars`` with floats at once (one host round trip instead of one per key)."""
    names, vals = [], []
    for loss_name, loss_value in losses.items():
        if isinstance(loss_value, torch.Tensor):
            vals.append(loss_value.mean())
        elif isinstance(loss_value, list):
            vals.append(sum(_loss.mean() for _loss in loss_value))
        else:
            raise TypeError(f'{loss_name} is not a tensor or list of tensors')
        names.append(loss_name)
    loss = sum(v for k, v in zip(names, vals) if 'loss' in k)
    names.append('loss')
    vals.append(loss if isinstance(loss, torch.Tensor) else torch.as_tensor(float(loss)))
    stacked = torch.stack([v.detach().float().reshape(()) for v in vals])
    distributed = dist.is_available() and dist.is_initialized()
    if distributed:
        if len(names) >= _SLOTS:
            raise ValueError(f'more than {_SLOTS - 1} log variables')
        buf = stacked.new_zeros(_SLOTS)
        buf[:len(names)] = stacked
        buf[_SLOTS - 1] = float(len(names) - 1)                          # the reference counts the keys before adding 'loss'
        buf /= dist.get_world_size()
        dist.all_reduce(buf)                                             # ONE collective for all log variables + the count
        stacked = torch.cat([buf[:len(names)], buf[_SLOTS - 1:]])
    log_vars = LogVars(names, stacked, mean_key_count=distributed)
    if not defer:
        log_vars._materialise()
    return loss, log_vars


def mask_branch_step(mask_head, mask_feat, cls_score, centerness, param_pred, coors, level_inds, img_inds, gt_inds, img,
                     img_metas, gt_bboxes, gt_masks=None, gt_labels=None):
    """condinst.py:66-75: ``training_sample`` -> ``mask_head(mask_feat, ...)`` -> ``mask_head.loss(...)``; returns the loss
    dict that ``forward_train`` merges into the detector's losses."""
    param_pred, coors, level_inds, img_inds, gt_inds = mask_head.training_sample(
        cls_score, centerness, param_pred, coors, level_inds, img_inds, gt_inds)
    mask_pred = mask_head(mask_feat, param_pred, coors, level_inds, img_inds)
    return mask_head.loss(img, img_metas, mask_pred, gt_inds, gt_bboxes, gt_masks, gt_labels)
