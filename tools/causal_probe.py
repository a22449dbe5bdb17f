import os, sys, torch
sys.path.insert(0, "/root/repo")
from multimodal_neuroimage_b200 import ops
nH, d, T, B = 12, 64, 2048, 32
E = nH * d
q, k, v = (torch.randn(T, B, E, device="cuda", dtype=torch.bfloat16) for _ in range(3))
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for mk, diag in ((0, 1), (3, 1), (3, 100000), (3, -100000), (3, 1025), (3, -1023)):
    t = timeit(lambda: ops.mha_fwd(q, k, v, None, nH, mk, diag, d ** -0.5, 0.0, 0, 0))
    print(f"mask_kind {mk} diag {diag}: fwd {t:.3f} ms")
