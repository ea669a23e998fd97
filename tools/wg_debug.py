import sys, torch, torch.nn.functional as F
sys.path.insert(0, '.')
from rpg_ramnet_b200 import ops
dev = torch.device('cuda', 0)
def nhwc(t): return t.to(dev).contiguous(memory_format=torch.channels_last)
torch.manual_seed(0)
N, H, W, C0, Cout, k, stride = 1, 8, 8, 32, 32, 1, 1
x = torch.randn(N, C0, H, W); dz = torch.randn(N, Cout, H, W)
mode = sys.argv[1] if len(sys.argv) > 1 else 'rand'
if mode == 'ones':
    x = torch.ones_like(x); dz = torch.ones_like(dz)
if mode == 'chan':   # x[c] = c+1, dz[c] = 1  -> dW[co][ci] = 64*(ci+1)
    x = (torch.arange(C0).float() + 1).view(1, C0, 1, 1).expand(N, C0, H, W).contiguous(); dz = torch.ones_like(dz)
if mode == 'chan2':  # dz[c] = c+1, x = 1 -> dW[co][ci] = 64*(co+1)
    dz = (torch.arange(Cout).float() + 1).view(1, Cout, 1, 1).expand(N, Cout, H, W).contiguous(); x = torch.ones_like(x)
w = torch.zeros(Cout, C0, k, k, requires_grad=True)
F.conv2d(x, w, None, stride=stride, padding=k // 2).backward(dz)
ref = w.grad[:, :, 0, 0]
dw = torch.zeros(Cout, C0, k, k, device=dev)
ops.conv_wgrad(nhwc(dz), nhwc(x), None, Cout, k, stride, dw, None, ops.MMA_TF32)
torch.cuda.synchronize()
a = dw.cpu()[:, :, 0, 0]
print('mode', mode, 'err', ((a - ref).norm() / ref.norm()).item(), 'errT', ((a.t() - ref).norm() / ref.norm()).item())
print('ours[0,:8]', a[0, :8].tolist()); print('ours[:8,0]', a[:8, 0].tolist()); print('ref[0,:8]', ref[0, :8].tolist()); print('ref[:8,0]', ref[:8, 0].tolist())
print('nonzero frac', (a != 0).float().mean().item())
