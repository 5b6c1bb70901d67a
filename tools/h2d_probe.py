"""PCIe probe: pinned host -> device and back, at the sizes bench.py's e2e arm moves per step (133 MB of frames up)."""
import torch
for mb in (8, 33, 133, 512):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    for direction in ("h2d", "d2h"):
        for _ in range(2):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{direction} {mb:4d} MB: {ms:.3f} ms = {mb * 1.048576 / ms:.1f} GB/s", flush=True)
