"""Pure-write, pure-read and copy bandwidth of this GPU (torch kernels, CUDA events, best of 10) — the training-mode forward and
dgrad kernels are write streams (8.1 / 7.6 GB per fine launch), the wgrad kernel a read stream (16 GB): which ceiling applies?"""
import torch
n = 1 << 31          # 2 Gi floats = 8 GiB
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n // 2, dtype=torch.float32, device="cuda")
def best(fn, bytes_, reps=10):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return bytes_ / (min(t) * 1e-3) / 1e9, min(t)
w, tw = best(lambda: a.zero_(), a.numel() * 4)
f, tf = best(lambda: a.fill_(1.5), a.numel() * 4)
r, tr = best(lambda: a.sum(), a.numel() * 4)
c, tc = best(lambda: b.copy_(a[: n // 2]), b.numel() * 8)
print(f"pure write (memset 8 GiB)      {w:8.1f} GB/s  ({tw:.3f} ms)")
print(f"pure write (fill kernel 8 GiB) {f:8.1f} GB/s  ({tf:.3f} ms)")
print(f"pure read  (sum 8 GiB)         {r:8.1f} GB/s  ({tr:.3f} ms)")
print(f"copy 4 GiB -> 4 GiB (r+w)      {c:8.1f} GB/s  ({tc:.3f} ms)")
