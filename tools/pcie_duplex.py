"""PCIe: H2D alone, D2H alone, both at once on two streams (pinned host memory, 1 GiB each) — does the link run full duplex?"""
import torch

dev = torch.device("cuda", 0)
n = 1 << 27   # 1 GiB of float64
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device=dev)
d_out = torch.zeros(n, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run(do_in, do_out, chunks=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream(dev)
    e0.record(main)
    s1.wait_stream(main); s2.wait_stream(main)
    step = n // chunks
    for c in range(chunks):
        a, b = c * step, (c + 1) * step
        if do_in:
            with torch.cuda.stream(s1):
                d_in[a:b].copy_(h_in[a:b], non_blocking=True)
        if do_out:
            with torch.cuda.stream(s2):
                h_out[a:b].copy_(d_out[a:b], non_blocking=True)
    main.wait_stream(s1); main.wait_stream(s2)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for _ in range(2):
    run(True, True)
gib = n * 8 / 1e9
for name, a, b, ch in (("H2D alone", True, False, 1), ("D2H alone", False, True, 1), ("both, 1 chunk", True, True, 1), ("both, 16 chunks", True, True, 16)):
    ms = min(run(a, b, ch) for _ in range(3))
    print(f"{name:18s} {ms:8.2f} ms  {gib * (a + b) / ms * 1e3:7.1f} GB/s total")
