import sys, torch
sys.path.insert(0, '.')
import torch.nn.functional as F
from deepsee_b200 import ops
variant = sys.argv[1]
B,H,W,Cin,Cout = 1,8,16,128,128
g = torch.Generator().manual_seed(0)
x = torch.randn(B,Cin,H,W,generator=g).cuda().requires_grad_(True)
w = (torch.randn(Cout,Cin,3,3,generator=g)/30).cuda().requires_grad_(True)
dy = torch.randn(B,Cout,H,W,generator=g).cuda()
F.conv2d(x,w,None,padding=1).backward(dy)
nhwc = lambda t: t.permute(0,2,3,1).contiguous()
a = ops.split_f16(nhwc(x.detach()))
if variant == 'bf16':
    gp,_ = ops.grad_prep(nhwc(dy))
elif variant == 'f16':
    gp = ops.split_f16(nhwc(dy))
elif variant == 'bothbf16':
    gp,_ = ops.grad_prep(nhwc(dy)); a,_ = ops.grad_prep(nhwc(x.detach()))
dw = ops.conv3x3_wgrad(gp, a, passes=3)
torch.cuda.synchronize()
print(variant, 'err', (dw-w.grad).abs().max().item(), 'scale', w.grad.abs().max().item())
