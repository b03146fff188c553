"""Host -> device episode feeder: overlaps the H2D copy of the next packed episode batch with the compute of the
current one (two device buffers, a dedicated copy stream, CUDA events for the hand-over).  The reference moves every
episode with a blocking ``.cuda()`` inside the loop (methods/DKT.py:120-124); at B200 step times the 284 MB / step copy
would otherwise serialise with the kernels."""
import torch


class DevicePrefetcher:
    def __init__(self, iterable, device, timing=False, start=True):
        self.it = iter(iterable)
        self.timing = timing                 # keep (start, end) CUDA events of every H2D copy (bench.py reads them)
        self.copy_events = []
        self.dev = torch.device(device)
        self.cuda = self.dev.type == "cuda"
        self.copy_stream = torch.cuda.Stream(self.dev) if self.cuda else None
        self.bufs = [None, None]
        self.ready = [None, None]
        self.free = [None, None]
        self.k = 0
        self._next = None
        self._started = False
        if start:
            self.start()

    def preallocate(self, shape):
        """Allocate both device buffers up front (keeps cudaMalloc out of the first steps)."""
        if self.cuda:
            for i in range(2):
                self.bufs[i] = torch.empty(tuple(shape), device=self.dev, dtype=torch.float32)

    def start(self):
        if not self._started:
            self._started = True
            self._prefetch()

    def _prefetch(self):
        try:
            x = next(self.it)
        except StopIteration:
            self._next = None
            return
        if not self.cuda:
            self._next = (x.to(self.dev).float().contiguous(), None)
            return
        i = self.k & 1
        if self.bufs[i] is None or self.bufs[i].shape != x.shape:
            self.bufs[i] = torch.empty(x.shape, device=self.dev, dtype=torch.float32)
        with torch.cuda.stream(self.copy_stream):
            if self.free[i] is not None:
                self.copy_stream.wait_event(self.free[i])      # the step that last read this buffer has finished
            if self.timing:
                t0 = torch.cuda.Event(enable_timing=True)
                t0.record(self.copy_stream)
            self.bufs[i].copy_(x, non_blocking=True)
            ev = torch.cuda.Event(enable_timing=self.timing)
            ev.record(self.copy_stream)
            if self.timing:
                self.copy_events.append((t0, ev))
        self._next = (self.bufs[i], ev)
        self.k += 1

    def __iter__(self):
        return self

    def __next__(self):
        self.start()
        if self._next is None:
            raise StopIteration
        buf, ev = self._next
        if ev is not None:
            torch.cuda.current_stream(self.dev).wait_event(ev)
        self._prefetch()
        return buf

    def release(self, buf):
        """Call after the kernels reading ``buf`` have been enqueued: marks the buffer reusable once they finish."""
        if not self.cuda:
            return
        for i in range(2):
            if self.bufs[i] is buf:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.dev))
                self.free[i] = ev


def bind_to_gpu_numa(device_index=0):
    """Pin this process (and therefore its later pinned-host allocations, first-touch) to the CPUs of the NUMA node the
    GPU hangs off: on a two-socket host a feeder running on the far socket pushes every episode batch across the
    inter-socket link first.  Best effort -- returns the CPU set used, or None if the topology cannot be read."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device_index
        if visible:
            tok = visible.split(",")[device_index].strip()
            if tok.isdigit():
                idx = int(tok)
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # nvml prints an 8-digit PCI domain, sysfs uses 4
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None
