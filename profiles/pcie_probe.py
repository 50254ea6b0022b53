"""Upper bound for the e2e leg: concurrent H2D + D2H copies of 0.54 GB each from / to pinned host memory (torch streams), by chunk count.
`python profiles/pcie_probe.py [modes]`; prints GB/s per direction.  Run once per GPU at the same time (CUDA_VISIBLE_DEVICES) it shows
what the host side of an 8-GPU box sustains.  Not a bench value."""
import sys
import time

import torch

n = 16777216 * 4
dev = torch.device("cuda:0")
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device=dev)
d_out = torch.empty(n, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ("h2d", "d2h", "both")
for mode in modes:
    for chunks in ((8,) if len(sys.argv) > 1 else (1, 4, 8, 16)):
        c = n // chunks
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(chunks):
                sl = slice(k * c, (k + 1) * c)
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        d_in[sl].copy_(h_in[sl], non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        h_out[sl].copy_(d_out[sl], non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        print(f"{mode:5s} chunks {chunks:2d}: {dt * 1e3:7.2f} ms  {n * 8 / dt / 1e9:6.1f} GB/s per direction", flush=True)
