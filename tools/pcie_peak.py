"""Device-to-host copy rate of this box (pinned memory, one cudaMemcpyAsync per size), to put the end-to-end bench
number in context: it returns 19 B per photon over this link."""
import torch

for mb in (1, 4, 19, 64, 256):
    n = mb * (1 << 20)
    d = torch.empty(n, dtype=torch.uint8, device='cuda')
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for _ in range(3):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(5, 2048 // mb)
    e0.record()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    e1.record()
    e1.synchronize()
    print('D2H %4d MiB x %4d: %.1f GB/s' % (mb, reps, reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9))
