import torch
x = torch.empty(1 << 31, dtype=torch.uint8, device="cuda")
y = torch.empty(1 << 31, dtype=torch.uint8, device="cuda")
def t(fn, n=5):
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
for _ in range(2): x.zero_(); y.copy_(x)
ms = t(lambda: x.zero_()); print("memset  2 GiB: %.3f ms -> %.0f GB/s (write only)" % (ms, (1 << 31) / ms / 1e6))
ms = t(lambda: y.copy_(x)); print("copy    2 GiB: %.3f ms -> %.0f GB/s (read + write)" % (ms, 2 * (1 << 31) / ms / 1e6))
ms = t(lambda: x.sum(dtype=torch.int64)); print("reduce  2 GiB: %.3f ms -> %.0f GB/s (read only)" % (ms, (1 << 31) / ms / 1e6))
