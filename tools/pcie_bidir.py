import torch, time
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, frac=1.0):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m = int(n * frac)
    if h2d:
        with torch.cuda.stream(s1): d_a[:m].copy_(h_in[:m], non_blocking=True)
    if d2h:
        with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t0
for _ in range(2): run(1, 1)
t = min(run(1, 0) for _ in range(3)); print("h2d only", n / t / 1e9)
t = min(run(0, 1) for _ in range(3)); print("d2h only", n / t / 1e9)
t = min(run(1, 1) for _ in range(3)); print("both 1:1 total", 2 * n / t / 1e9, "time ms", t * 1e3)
t = min(run(1, 1, 0.32) for _ in range(3)); print("both 0.32:1 total", 1.32 * n / t / 1e9, "d2h-rate", n / t / 1e9)
