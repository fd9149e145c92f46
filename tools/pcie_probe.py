"""tools/pcie_probe.py — development aid: device->host copy rate of this box for the e2e path's transfer sizes
(one 24.9 MB frame at 24 bit/pixel as 1, 12 and 24 pinned copies; CUDA events)."""
import torch
n = 3840 * 2160 * 3
src = torch.empty(n, dtype=torch.uint8, device="cuda:0")
dst = torch.empty(n, dtype=torch.uint8).pin_memory()
s = torch.cuda.Stream()
for parts in (1, 2, 12, 24):
    edges = [n * i // parts for i in range(parts + 1)]
    with torch.cuda.stream(s):
        for rep in range(3):
            for i in range(parts):
                dst[edges[i]:edges[i + 1]].copy_(src[edges[i]:edges[i + 1]], non_blocking=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for rep in range(20):
            for i in range(parts):
                dst[edges[i]:edges[i + 1]].copy_(src[edges[i]:edges[i + 1]], non_blocking=True)
        e1.record(s)
    s.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{parts:3d} copies per frame: {ms * 1e3:7.1f} us  {n / ms / 1e6:6.1f} GB/s")
