"""Write-only HBM bandwidth on this GPU (what bounds the 80 GB expansion): torch fill_ / zero_ of a 32 GB buffer."""
import torch
x = torch.empty(4 * 1024**3, dtype=torch.float64, device="cuda")
for name, fn in (("zero_ (memset)", lambda: x.zero_()), ("fill_ (store kernel)", lambda: x.fill_(0.5)), ("copy_ (read+write)", lambda: x[: x.numel() // 2].copy_(x[x.numel() // 2:]))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    nbytes = x.numel() * 8 if "copy" not in name else x.numel() * 8     # copy: half read + half written = the whole buffer
    print("%-22s %.2f ms  %.0f GB/s" % (name, ms, nbytes / ms / 1e6))
