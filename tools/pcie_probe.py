"""PCIe rates of this host: 1-D pinned copies vs the 2-D (pitched) D2H copies the streaming parse makes.
python tools/pcie_probe.py"""
import time, json, ctypes as C
import torch
rt = C.CDLL("libcudart.so.12")
S, V, n = 2504, 1_100_000, 26_000
d = torch.zeros(S * 32768, dtype=torch.uint8, device="cuda")          # device planes: [S][stride 32768]
h = torch.empty(S * V, dtype=torch.uint8).pin_memory()
h1 = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
d1 = torch.zeros(1 << 30, dtype=torch.uint8, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
out = {}
dt = t(lambda: h1.copy_(d1, non_blocking=True)); out["d2h_1d_GBs"] = (1 << 30) / dt / 1e9
dt = t(lambda: d1.copy_(h1, non_blocking=True)); out["h2d_1d_GBs"] = (1 << 30) / dt / 1e9
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
for width in (4096, 26000, 32768):
    def f():
        rc = rt.cudaMemcpy2DAsync(h.data_ptr(), V, d.data_ptr(), 32768, width, S, 2, None)
        assert rc == 0
    dt = t(f); out["d2h_2d_w%d_GBs" % width] = width * S / dt / 1e9
# both directions at once
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2): d.copy_(h[:d.numel()], non_blocking=True)
dt = t(both); out["bidir_GBs_each"] = [(1 << 30) / dt / 1e9, d.numel() / dt / 1e9]
print(json.dumps(out))
