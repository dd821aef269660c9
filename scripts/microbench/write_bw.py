"""HBM bandwidth by access mix on one B200 (torch kernels, CUDA events): pure write, copy (1 read : 1 write), pure read,
and 1 read : 4 writes (the mix of k_scan_lanes: u in, X and Xs out).  Prints GB/s per mix."""
import torch

dev = torch.device("cuda:0")
n = 1 << 29                                    # 4 GiB of fp64
a = torch.empty(n, dtype=torch.float64, device=dev)
b = torch.empty(n, dtype=torch.float64, device=dev)
small = torch.empty(n // 4, dtype=torch.float64, device=dev).normal_()


def timed(fn, bytes_moved, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return bytes_moved * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


print("pure write  (fill)            %.0f GB/s" % timed(lambda: a.fill_(1.5), 8.0 * n))
print("copy        (1 read : 1 write) %.0f GB/s" % timed(lambda: b.copy_(a), 16.0 * n))
print("pure read   (sum)             %.0f GB/s" % timed(lambda: a.sum(), 8.0 * n))
av = a.view(4, n // 4)
print("1 read : 4 writes             %.0f GB/s" % timed(lambda: torch.add(small.view(1, -1), 1.0, out=av[0:1]) if False else av.copy_(small.view(1, -1).expand(4, -1)), 8.0 * n + 8.0 * n / 4))
print("2 reads : 3 writes (filter)   see bench kernels_ms")
