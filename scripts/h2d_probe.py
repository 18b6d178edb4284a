import time, torch
n, sz = 296, 640 * 480
host = torch.empty(64, sz, dtype=torch.uint8).pin_memory()
dev = torch.empty(n, sz, dtype=torch.uint8, device="cuda")
big_h = torch.empty(n * sz, dtype=torch.uint8).pin_memory()
st = torch.cuda.Stream()
def many():
    with torch.cuda.stream(st):
        for i in range(n):
            dev[i].copy_(host[(5 * i) % 64], non_blocking=True)
    st.synchronize()
def one():
    with torch.cuda.stream(st):
        dev.view(-1).copy_(big_h, non_blocking=True)
    st.synchronize()
for f, name in ((many, "296 x 307KB"), (one, "1 x 91MB")):
    for _ in range(3): f()
    t = time.perf_counter()
    for _ in range(10): f()
    dt = (time.perf_counter() - t) / 10
    print(name, f"{dt*1e3:.2f} ms  {n*sz/dt/1e9:.1f} GB/s")
