import torch, time
dev = torch.device("cuda:0")
n = 40 * 4096 * 4096
host = torch.empty((n,), dtype=torch.float32, pin_memory=True); host.fill_(1.0)
devb = torch.empty((n,), dtype=torch.float32, device=dev)
for rep in range(3):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); devb.copy_(host, non_blocking=True); e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e); print("H2D pinned 2.68 GB:", round(ms, 2), "ms", round(n * 4 / ms / 1e6, 1), "GB/s", flush=True)
# two halves on two streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): devb[: n // 2].copy_(host[: n // 2], non_blocking=True)
with torch.cuda.stream(s2): devb[n // 2 :].copy_(host[n // 2 :], non_blocking=True)
torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
print("H2D 2 streams:", round(ms, 2), "ms", round(n * 4 / ms / 1e6, 1), "GB/s")
