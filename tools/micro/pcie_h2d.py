"""Pinned host -> device copy rate of this box (the ceiling of bench.py's end-to-end number).  python tools/micro/pcie_h2d.py"""
import json
import torch

n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.fill_(7)
d = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
s = [torch.cuda.Stream() for _ in range(2)]
for i in range(4):
    with torch.cuda.stream(s[i & 1]):
        d[i & 1].copy_(h, non_blocking=True)
torch.cuda.synchronize()
out = {}
for label, streams in (("one_stream", 1), ("two_streams", 2)):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    reps = 40
    for i in range(reps):
        st = s[i % streams]
        st.wait_event(a) if i < streams else None
        with torch.cuda.stream(st):
            d[i % streams].copy_(h, non_blocking=True)
    for st in s:
        torch.cuda.current_stream().wait_stream(st)
    b.record()
    torch.cuda.synchronize()
    out[label + "_GBps"] = round(reps * n / (a.elapsed_time(b) * 1e-3) / 1e9, 2)
print(json.dumps({"what": "pinned H2D copies of 256 MiB", **out}))
