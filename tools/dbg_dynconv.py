import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boxinstseg_b200 import _lib as L
lib = L.lib()
h = ctypes.CDLL(L.LIB_PATH)
dev = 'cuda:0'
B, C, hh, ww, I = 1, 256, 16, 32, 16
feat = torch.arange(B * C * hh * ww, dtype=torch.float32, device=dev).reshape(B, C, hh, ww) % 97 / 16
kern = (torch.arange(B * I * C, dtype=torch.float32, device=dev).reshape(B, I, C) % 13) / 8
dbg = torch.zeros(256, device=dev)
h.bxs_dynconv1x1_set_debug(ctypes.c_void_p(dbg.data_ptr()))
out = torch.full((B, I, hh, ww), -7.0, device=dev)
rc = lib.bxs_dynconv1x1_forward(L.ptr(feat), L.ptr(kern), L.ptr(out), B, C, hh * ww, I, L.stream())
torch.cuda.synchronize()
print('rc', rc)
d = dbg.cpu()
print('smem A[0:16]', d[:16].tolist())
print('feat[c=0, p=0:8]', feat[0, 0].flatten()[:8].tolist(), ' feat[c=1,p=0:4]', feat[0, 1].flatten()[:4].tolist())
print('smem A[32:48]', d[32:48].tolist())
print('smem B[0:16]', d[64:80].tolist())
print('kern[i=0,c=0:8]', kern[0, 0, :8].tolist())
print('tmem_base %x idesc %x a_base %x b_base %x' % tuple(int(x) for x in d[128:132].view(torch.int32).tolist()))
for wq in range(4): print('tmem warp', wq + 2, d[160 + 16 * wq:176 + 16 * wq].tolist())
print('poison readback', d[224:228].tolist(), ' late read', d[232:248].tolist())
ref = torch.einsum('bic,bchw->bihw', kern.double(), feat.double())
print('ref[0,0:4,0,0]', ref[0, :4, 0, 0].tolist(), 'out', out[0, :4, 0, 0].tolist())
print('out stats', out.min().item(), out.max().item(), 'ref max', ref.abs().max().item())
