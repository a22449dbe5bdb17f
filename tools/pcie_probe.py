import torch, time
n = 402653184
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / it * 1e3
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()
def both_chunked():
    c = n // 8
    for i in range(8):
        with torch.cuda.stream(s1): d_in[i*c:(i+1)*c].copy_(h_in[i*c:(i+1)*c], non_blocking=True)
        with torch.cuda.stream(s2): h_out[i*c:(i+1)*c].copy_(d_out[i*c:(i+1)*c], non_blocking=True)
print("h2d ms", t(h2d), "GB/s", n / t(h2d) / 1e6)
print("d2h ms", t(d2h), "GB/s", n / t(d2h) / 1e6)
print("both ms", t(both))
print("both chunked ms", t(both_chunked))
