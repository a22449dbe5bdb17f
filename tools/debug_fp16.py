"""Debug aid: per-module input-gradient error of a block under fp16 / bf16 autocast against the fp64 oracle."""
import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import ref_nd as R
from test_gpu_parity import _randomise, _sd64, rel_err
from multimodal_neuroimage_b200.modules import swin_v2_module as v2, swinfusion_module as fu, crossmodal_transformer as cm

grid, C, nH, B = (8, 8, 8), 96, 3, 2
blk = v2.SwinTransformerBlock(C, grid, nH, window_size=4, shift_size=2)
cross = fu.Cross_SwinTransformerBlock(C, grid, nH, window_size=4, shift_size=2)
for i, m in enumerate((blk, cross)):
    _randomise(m, 20 + i)
g = torch.Generator().manual_seed(2)
x, y = torch.randn(B, math.prod(grid), C, generator=g), torch.randn(B, math.prod(grid), C, generator=g)
xo, yo = (t.double().requires_grad_(True) for t in (x, y))
w1 = R.swin_v2_block(xo, _sd64(blk), grid, 4, 2, nH)
w2a, w2b = R.cross_block(xo, yo, _sd64(cross), grid, grid, 4, 2, nH)
g1 = torch.autograd.grad(w1.mean(), xo)[0]
g2 = torch.autograd.grad(w2a.mean() + w2b.mean(), (xo, yo))
blk, cross = blk.cuda(), cross.cuda()
for dt in (torch.float16, torch.bfloat16):
    for scale in (1.0, 65536.0):
        xc, yc = (t.cuda().requires_grad_(True) for t in (x, y))
        with torch.autocast("cuda", dtype=dt):
            o1 = blk(xc)
            o2a, o2b = cross(xc, yc, grid)
        d1 = torch.autograd.grad(o1.float().mean() * scale, xc)[0] / scale
        d2 = torch.autograd.grad((o2a.float().mean() + o2b.float().mean()) * scale, (xc, yc))
        print(dt, scale, "swinv2 dx", rel_err(d1, g1), "cross dx", rel_err(d2[0] / scale, g2[0]), "cross dy", rel_err(d2[1] / scale, g2[1]),
              "sum dx", rel_err(d1 + d2[0] / scale, g1 + g2[0]), "out", rel_err(o1, w1), rel_err(o2a, w2a))
