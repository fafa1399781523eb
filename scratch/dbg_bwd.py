import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from oracle import deepsee_oracle as O
from test_generator_gpu import _build_G
from deepsee_b200 import ops
from deepsee_b200.config import config
train = sys.argv[1] == 'train'
name, over = "8x_independent_256x256", dict(ngf=8, start_size=8, crop_size=64, load_size=64)
o = O.make_opt(name, is_train=True, **over)
sd = O.make_generator_state(o, 0)
sd_ref = {k: v.clone() for k, v in sd.items()}
for k, v in sd_ref.items():
    if v.is_floating_point() and not O._is_buffer(k): v.requires_grad_(True)
batch = 2
d = O.preprocess(o, O.synthetic_batch(o, batch, seed=31))
g = torch.Generator().manual_seed(32)
z_ref = (torch.rand(batch, 19, 128, generator=g) * 2 - 1).requires_grad_(True)
proj = torch.randn(batch, 3, o.crop_size, o.crop_size, generator=g)
noises = {}
def noise_fn(nm, shape):
    noises[nm] = torch.randn(shape, generator=torch.Generator().manual_seed(len(noises)))
    return noises[nm]
taps = {}
class T(dict):
    def __setitem__(self, k, v):
        v.retain_grad(); super().__setitem__(k, v)
taps = T()
ref = O.generator_forward(sd_ref, o, d["image_lr"], d["input_semantics"], z_ref, train, noise_fn, taps)
(ref * proj).sum().backward()
config.passes = 3
G = _build_G(o, sd); G.train(train)
if o.add_noise and train:
    for pfx, _, _ in O.generator_layout(o):
        blk = G.get_submodule(pfx[:-1])
        for nm in ("noise_in", "noise_skip", "noise_middle"):
            n = noises[pfx + nm].permute(0, 2, 3, 1).contiguous().cuda()
            getattr(blk, nm).sample = (lambda t: (lambda B, H, W: t))(n)
# record intermediates
rec = []
orig_conv = ops.conv3x3
def conv_rec(*a, **k):
    r = orig_conv(*a, **k)
    if k.get('tag') == 'dgrad': rec.append(('dt', r[0].clone()))
    return r
ops.conv3x3 = conv_rec
import deepsee_b200.deepsee_models.networks.architecture as A
orig_gp = ops.grad_prep
def gp_rec(dy, *a, **k):
    rec.append(('gp_in', dy.clone())); return orig_gp(dy, *a, **k)
ops.grad_prep = gp_rec
orig_bn = ops.bn_bwd
def bn_rec(*a, **k):
    r = orig_bn(*a, **k); rec.append(('bn_dx', r[0].clone())); return r
ops.bn_bwd = bn_rec
z = z_ref.detach().clone().cuda().requires_grad_(True)
out = G(d["image_lr"].cuda(), seg=d["input_semantics"].cuda(), z=z)
(out * proj.cuda()).sum().backward()
torch.cuda.synchronize()
# order of rec: blocks in reverse; per block: dt1, bn_dx(ddx1), dt0, bn_dx(dx)
lay = [p for p, _, _ in O.generator_layout(o)][::-1]
def rel(a, b):
    e = (a - b).abs(); s_ = b.abs().max()
    return "max %.1e l2 %.1e frac>1e-3 %.1e" % ((e.max() / s_).item(), ((a - b).norm() / b.norm()).item(), (e > 1e-3 * s_).float().mean().item())
i = 0
nchw = lambda t: t.permute(0, 3, 1, 2).cpu()
for pfx in lay:
    dout, dt1, ddx1, _gp0, dt0, dx = [r[1] for r in rec[i:i + 6]]; i += 6
    a1, a0 = taps[pfx + 'act_1'], taps[pfx + 'act_0']
    # reference dt = da * lrelu'(.)
    m1 = torch.where(a1 > 0, 1.0, 0.2); m0 = torch.where(a0 > 0, 1.0, 0.2)
    print(pfx, '\n   dout', rel(nchw(dout), taps[pfx+'out'].grad), '\n   dt1 ', rel(nchw(dt1), a1.grad * m1), '\n   dt0 ', rel(nchw(dt0), a0.grad * m0))
