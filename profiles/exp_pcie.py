"""PCIe rate of the box: 1.25 GB pinned copies in both directions (the decoder returns that much per clip)."""
import torch, time
n = 1253376000
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
    for _ in range(2): fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(name, f"{n / ms / 1e6:.1f} GB/s ({ms:.1f} ms)")
# chunked d2h on a side stream in 20 pieces
s = torch.cuda.Stream()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(s):
    e0.record()
    step = n // 30
    for i in range(30):
        h[i * step:(i + 1) * step].copy_(d[i * step:(i + 1) * step], non_blocking=True)
    e1.record()
torch.cuda.synchronize()
print("d2h in 30 chunks", f"{n / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
