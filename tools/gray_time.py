"""Device time of the grayscale conversion on a batch of 1080p frames (HBM roofline: 4 B / pixel for BGR, 5 B for BGRA)."""
import sys
sys.path.insert(0, ".")
import ctypes as C
import json
import torch
from stella_vslam_b200 import feature
from stella_vslam_b200._lib import check, lib
B, H, W = 64, 1080, 1920
peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6573.8) if __import__("os").path.exists("MEASURED_PEAKS.json") else 6573.8
ex = feature.orb_extractor(feature.orb_params(), 800, max_batch=B)
stream = torch.cuda.current_stream()
check(lib().b200_orb_set_stream(ex._h, C.c_void_p(stream.cuda_stream), 0))
for ch in (3, 4):
    src = torch.randint(0, 256, (B, H, W, ch), dtype=torch.uint8, device="cuda")
    dst = torch.empty((B, H, W), dtype=torch.uint8, device="cuda")
    run = lambda: check(lib().b200_convert_to_grayscale_device(ex._h, src.data_ptr(), W, H, W * ch, H * W * ch, ch, 0, dst.data_ptr(), W, H * W, B))
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(10):
        flush.zero_()
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    gb = B * H * W * (ch + 1) / 1e9
    print(f"{ch} channels: {ms:.3f} ms per {B} frames, {gb / ms * 1e3:.0f} GB/s = {gb / ms * 1e3 / peak:.2f} of the measured HBM peak ({peak:.0f} GB/s)")
