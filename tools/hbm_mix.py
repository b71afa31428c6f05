"""HBM bandwidth by access mix: write-only, read-only, copy (torch kernels; size >> L2)."""
import torch, json
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(True); b = torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
N = 4 * 1024 ** 3
x = torch.empty(N, dtype=torch.uint8, device="cuda"); y = torch.empty(N, dtype=torch.uint8, device="cuda")
xi = x.view(torch.int32)
res = {}
ms = t(lambda: x.fill_(3)); res["write_only_fill_GBs"] = N / ms / 1e6
ms = t(lambda: x.zero_()); res["write_only_memset_GBs"] = N / ms / 1e6
ms = t(lambda: y.copy_(x)); res["copy_GBs(read+write)"] = 2 * N / ms / 1e6
ms = t(lambda: xi.sum()); res["read_only_sum_GBs"] = N / ms / 1e6
# 1 read : 6 write like 16->96 expansion
a = torch.empty(N // 8, dtype=torch.uint8, device="cuda"); 
ms = t(lambda: torch.add(x, 1, out=y)); res["add_scalar_GBs(read+write)"] = 2 * N / ms / 1e6
print(json.dumps(res))
