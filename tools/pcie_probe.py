"""Host<->device copy rates of the box from pinned memory: one 65 MB copy each way (the bench's per-step state), the same bytes as
ten array-sized copies (what cntmc_kubo_step_host_state issues per slice), and both directions at once on two streams."""
import json, torch
dev = torch.device("cuda:0")
n = 65_000_000
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = {}
out["h2d_ms"] = t(lambda: d.copy_(h, non_blocking=True))
out["d2h_ms"] = t(lambda: h.copy_(d, non_blocking=True))
sizes = [4, 8, 8, 8, 8, 8, 8, 8, 1, 4]  # bytes per exciton of the ten state arrays
offs = [0]
for s in sizes: offs.append(offs[-1] + s * 1_000_000)
def ten_h2d():
    for i in range(10): d[offs[i]:offs[i + 1]].copy_(h[offs[i]:offs[i + 1]], non_blocking=True)
def ten_d2h():
    for i in range(10): h[offs[i]:offs[i + 1]].copy_(d[offs[i]:offs[i + 1]], non_blocking=True)
out["h2d_ten_copies_ms"] = t(ten_h2d); out["d2h_ten_copies_ms"] = t(ten_d2h)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
out["both_directions_ms"] = t(both)
out["h2d_gbs"] = n / out["h2d_ms"] / 1e6; out["d2h_gbs"] = n / out["d2h_ms"] / 1e6
print(json.dumps(out))
