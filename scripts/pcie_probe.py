"""PCIe yard-stick for the e2e figure: pinned H2D alone, D2H alone, both at once (torch, no product code).
    python scripts/pcie_probe.py [MiB]"""
import sys, json, torch
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = mb << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down, reps=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for _ in range(2): run(True, True, 2)
t_up, t_dn, t_both = run(True, False), run(False, True), run(True, True)
print(json.dumps({"MiB_in": mb, "MiB_out": mb // 2, "h2d_alone_GBps": round(n / t_up / 1e6, 2), "d2h_alone_GBps": round(n / 2 / t_dn / 1e6, 2),
                  "both_ms": round(t_both, 3), "h2d_in_duplex_GBps": round(n / t_both / 1e6, 2)}))
