"""Host<->device copy bandwidth of every visible GPU, one at a time and all at once (pinned memory): tells whether the GPUs of a
box share a PCIe link / root (the end-to-end leg of a decomposed run moves 1/N of the arrays per rank, which only helps if they do not)."""
import sys, time, torch
n = torch.cuda.device_count()
mb = 48
bufs = []
for d in range(n):
    torch.cuda.set_device(d)
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    g = torch.empty(mb << 20, dtype=torch.uint8, device=f"cuda:{d}")
    h2 = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    bufs.append((h, g, h2, torch.cuda.Stream(device=d), torch.cuda.Stream(device=d)))
def run(devs, duplex):
    for d in devs:
        torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    for _ in range(10):
        for d in devs:
            h, g, h2, s1, s2 = bufs[d]
            with torch.cuda.stream(s1):
                g.copy_(h, non_blocking=True)
            if duplex:
                with torch.cuda.stream(s2):
                    h2.copy_(g, non_blocking=True)
    for d in devs:
        torch.cuda.synchronize(d)
    dt = time.perf_counter() - t0
    return 10 * mb / 1024 / dt  # GB/s per device, per direction
for d in range(n):
    run([d], False)
    print(f"gpu {d} alone: H2D {run([d], False):.1f} GB/s, duplex {run([d], True):.1f} GB/s per direction")
if n > 1:
    print(f"all {n} together: H2D {run(list(range(n)), False):.1f} GB/s per GPU, duplex {run(list(range(n)), True):.1f} GB/s per GPU per direction")
