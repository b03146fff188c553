"""Host link probe: pinned H2D bandwidth of one 284 MB episode batch, alone / in 4 concurrent chunks / while a kernel runs."""
import os, time, torch
dev = torch.device("cuda:0")
n = 284497920 // 4
host = torch.empty(n, dtype=torch.float32).pin_memory()
host.uniform_()
dst = torch.empty(n, device=dev)
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))


def timed(fn, reps=5):
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return sorted(out)[len(out) // 2]


ms = timed(lambda: dst.copy_(host, non_blocking=True))
print("single copy: %.2f ms  %.1f GB/s" % (ms, n * 4 / ms / 1e6))
streams = [torch.cuda.Stream() for _ in range(4)]
ch = n // 4


def chunked():
    evs = []
    for i, s in enumerate(streams):
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            dst[i * ch:(i + 1) * ch].copy_(host[i * ch:(i + 1) * ch], non_blocking=True)
        torch.cuda.current_stream().wait_stream(s)


ms = timed(chunked)
print("4 chunks on 4 streams: %.2f ms  %.1f GB/s" % (ms, n * 4 / ms / 1e6))
# while a bandwidth-heavy kernel runs on the default stream
big = torch.empty(1 << 30, device=dev, dtype=torch.float32)
cs = torch.cuda.Stream()


def overlapped():
    cs.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cs):
        dst.copy_(host, non_blocking=True)
    for _ in range(20):
        big.mul_(1.0001)
    torch.cuda.current_stream().wait_stream(cs)


ms_k = timed(lambda: [big.mul_(1.0001) for _ in range(20)])
ms = timed(overlapped)
print("20 x 8 GB kernels alone %.2f ms, with the copy alongside %.2f ms" % (ms_k, ms))
t0 = time.time(); tmp = torch.empty(n, dtype=torch.float32); tmp.copy_(host); print("host memcpy 284 MB: %.1f ms" % ((time.time() - t0) * 1e3))
