import torch, time
n = 4 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device='cuda'); d_out = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, chunk):
    torch.cuda.synchronize(); t = time.time()
    for o in range(0, n, chunk):
        if h2d:
            with torch.cuda.stream(s1): d_in[o:o+chunk].copy_(h_in[o:o+chunk], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out[o:o+chunk].copy_(d_out[o:o+chunk], non_blocking=True)
    torch.cuda.synchronize(); dt = time.time() - t
    return (n * (h2d + d2h)) / dt / 1e9
for chunk in (256 << 20, 64 << 20):
    for rep in range(2):
        print("chunk %d MB: h2d %.1f GB/s, d2h %.1f GB/s, both %.1f GB/s aggregate" % (chunk >> 20, run(1, 0, chunk), run(0, 1, chunk), run(1, 1, chunk)))
