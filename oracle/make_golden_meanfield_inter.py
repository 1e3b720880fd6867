"""Golden vectors for the mean field WITH the inter-image term of corr_loss (``MeanField.forward(x, targets, inter_img_mask)``,
mmdet/models/dense_heads/discobox_head.py:616-651), minted from the reference's own class (AST-extracted like
oracle/make_golden.py does); asserts the oracle restatement is bit-exact, writes tests/golden/meanfield_inter.npz.
      python -m oracle.make_golden_meanfield_inter"""
import os

import numpy as np
import torch

from oracle import levelset as ol
from oracle.make_golden import OUT, extract

CFG = dict(kernel_size=3, theta0=0.5, theta1=30, theta2=10, alpha0=2, iter=10, base=0.1)


def case(seed, n=4, h=14, w=18):
    gen = torch.Generator().manual_seed(seed)
    fm = torch.randn(1, 3, h, w, generator=gen)
    x = torch.rand(n, 1, h, w, generator=gen)
    t = torch.zeros(n, 1, h, w)
    t[0, 0, 1:11, 2:15] = 1
    t[1, 0, :, :] = 1
    t[2, 0, 4:9, 4:12] = 1
    t[3, 0, 2:13, 1:9] = 1
    iiu = torch.zeros(n, 2, h, w)
    iiu[0, :, 1:11, 2:15] = torch.rand(2, 10, 13, generator=gen)           # pasted into the box, like corr_loss :1104-1107
    iiu[2, :, 4:9, 4:12] = torch.rand(2, 5, 8, generator=gen)
    iiu[3, 1, 2:13, 1:9] = 0.9                                             # a strong foreground prior
    return fm, x, t, iiu


def main():
    db = extract('mmdet/models/dense_heads/discobox_head.py', ['MeanField'])
    out = {}
    for seed, gamma in ((0, 0.01), (1, 0.5), (2, 2.0)):
        fm, x, t, iiu = case(seed)
        mf = db.MeanField(fm, gamma=gamma, **CFG)
        pseudo, valid = mf(x, t, iiu.clone())
        plain, _ = mf(x, t)
        k = ol.meanfield_kernel(fm, 3, 0.5, 30.0, 2.0)
        o_ps, o_va = ol.meanfield_forward(k, x, t, 3, 10, 0.1, inter=iiu, gamma=gamma)
        assert torch.equal(o_ps, pseudo) and torch.equal(o_va, valid), 'oracle restatement != reference'
        out[f's{seed}_pseudo'] = np.packbits(pseudo.numpy().astype(np.uint8), axis=-1)
        out[f's{seed}_valid'] = valid.numpy()
        out[f's{seed}_flips'] = np.asarray(int((pseudo != plain).sum()))
        print('seed', seed, 'gamma', gamma, 'pixels changed by the inter-image term:', int((pseudo != plain).sum()))
    assert int(out['s1_flips']) > 0 and int(out['s2_flips']) > 0
    np.savez_compressed(os.path.join(OUT, 'meanfield_inter.npz'), **out)


if __name__ == '__main__':
    main()
