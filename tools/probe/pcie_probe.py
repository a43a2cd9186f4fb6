import torch, time
n = 393216000
d = torch.empty(n, dtype=torch.uint8, device='cuda')
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(29000000, dtype=torch.uint8).pin_memory(); d2 = torch.empty(29000000, dtype=torch.uint8, device='cuda')
s = torch.cuda.Stream(); s2 = torch.cuda.Stream()
for it in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print('D2H alone %.2f ms %.1f GB/s' % (dt * 1e3, n / dt / 1e9))
for it in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s): h.copy_(d, non_blocking=True)
        with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print('D2H + 29 MB H2D concurrently %.2f ms %.1f GB/s' % (dt * 1e3, n / dt / 1e9))
# chunked D2H
for chunks in (4, 16):
    torch.cuda.synchronize(); t = time.perf_counter()
    c = n // chunks
    for _ in range(5):
        with torch.cuda.stream(s):
            for k in range(chunks): h[k*c:(k+1)*c].copy_(d[k*c:(k+1)*c], non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print('D2H in %d chunks %.2f ms %.1f GB/s' % (chunks, dt * 1e3, n / dt / 1e9))
