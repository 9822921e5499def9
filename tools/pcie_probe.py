import torch, time
n = 58720256
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=20, chunks=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c = n // chunks
    for _ in range(reps):
        for k in range(chunks):
            if h2d:
                with torch.cuda.stream(s1): d_in[k*c:(k+1)*c].copy_(h_in[k*c:(k+1)*c], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out[k*c:(k+1)*c].copy_(d_out[k*c:(k+1)*c], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return n * reps / dt / 1e9
for ch in (1, 8):
    print("chunks", ch, "H2D only %.1f GB/s" % run(True, False, chunks=ch), "D2H only %.1f GB/s" % run(False, True, chunks=ch), "both: %.1f GB/s each way" % run(True, True, chunks=ch))
