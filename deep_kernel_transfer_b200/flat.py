"""Flat fp32 parameter / gradient buffers.  Every trainable tensor of the model is a VIEW into one
contiguous device buffer (and its ``.grad`` a view into a second one of the same layout), so that a
meta-step needs exactly one NCCL all-reduce and one fused Adam launch per learning-rate group
(SURVEY.md 8e).  Segments start on 16-byte boundaries (the kernels read biases / affine parameters
as float4)."""
import torch

ALIGN = 4  # floats


class FlatPack:
    @staticmethod
    def size_of(tensors):
        off = 0
        for _, t, align in tensors:
            off = (off + align - 1) // align * align + t.numel()
        return (max(off, 1) + ALIGN - 1) // ALIGN * ALIGN

    def __init__(self, tensors, device, with_grad=True, flat_storage=None, grad_storage=None):
        """tensors: ordered list of (name, torch.Tensor or nn.Parameter, start alignment in floats).  flat_storage /
        grad_storage: optional pre-allocated fp32 buffers of ``size_of(tensors)`` elements to live in (so that several
        packs can share one contiguous communication buffer)."""
        self.names, self.offsets, self.shapes = [], [], []
        off = 0
        for name, t, align in tensors:
            off = (off + align - 1) // align * align
            self.names.append(name)
            self.offsets.append(off)
            self.shapes.append(tuple(t.shape))
            off += t.numel()
        self.numel = (max(off, 1) + ALIGN - 1) // ALIGN * ALIGN
        self.flat = flat_storage if flat_storage is not None else torch.zeros(self.numel, device=device, dtype=torch.float32)
        self.grad = None
        if with_grad:
            self.grad = grad_storage if grad_storage is not None else torch.zeros(self.numel, device=device, dtype=torch.float32)
        assert self.flat.numel() == self.numel and (self.grad is None or self.grad.numel() == self.numel)
        self._tensors = [t for _, t, _ in tensors]
        for t, o, s in zip(self._tensors, self.offsets, self.shapes):
            n = 1
            for d in s:
                n *= d
            view = self.flat[o:o + n].view(s)
            view.copy_(t.data.to(device=device, dtype=torch.float32))
            t.data = view
            if with_grad and isinstance(t, torch.nn.Parameter) and t.requires_grad:
                t.grad = self.grad[o:o + n].view(s)

    def intact(self):
        base = self.flat.data_ptr()
        return all(t.data_ptr() == base + 4 * o and t.device == self.flat.device
                   for t, o in zip(self._tensors, self.offsets))

    def view(self, name, grad=False):
        i = self.names.index(name)
        o, s = self.offsets[i], self.shapes[i]
        n = 1
        for d in s:
            n *= d
        return (self.grad if grad else self.flat)[o:o + n].view(s)

