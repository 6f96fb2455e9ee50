"""Out-of-bounds WRITE check with guard bands (compute-sanitizer is closed on this GPU pool, so the bounds check is our own).

Every device buffer the model / ops layer allocates while the unit workload of profiles/sanitizer_unit.py runs — both
encoder layouts, training + inference, tcgen05 and block-loop attention, three optimizers, static and dynamic GEMM
scheduling — is carved out of a larger allocation whose 4 KiB margins on both sides are filled with a byte pattern; after
the device has drained, every margin must still hold the pattern. Catches any kernel (TMA stores and red.global included)
writing up to 4 KiB before or after a buffer it was handed.

    python profiles/guard_check.py            # prints buffers checked / violations, exit code 1 on a violation
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GUARD_BYTES = 4096
PATTERN = 0xA5


class GuardedTorch:
    """Stands in for the `torch` module inside nbest_b200.model / ops / optim: empty / zeros / full on a CUDA device return
    views into guard-banded allocations; everything else is forwarded."""

    def __init__(self):
        self.records = []

    def __getattr__(self, name):
        return getattr(torch, name)

    def _alloc(self, shape, dtype, device, fill=None):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        dtype = dtype or torch.float32
        isz = torch.empty(0, dtype=dtype).element_size()
        n = 1
        for s in shape:
            n *= int(s)
        body = (n * isz + 255) // 256 * 256
        raw = torch.empty(body + 2 * GUARD_BYTES, dtype=torch.uint8, device=device)
        raw.fill_(PATTERN)
        inner = raw[GUARD_BYTES:GUARD_BYTES + n * isz].view(dtype).view(tuple(int(s) for s in shape))
        if fill is not None:
            inner.fill_(fill)
        self.records.append((raw, n * isz))
        return inner

    def empty(self, *shape, dtype=None, device=None, pin_memory=False, **kw):
        if device is None or torch.device(device).type != "cuda" or pin_memory:
            return torch.empty(*shape, dtype=dtype, device=device, pin_memory=pin_memory, **kw)
        return self._alloc(shape, dtype, device)

    def zeros(self, *shape, dtype=None, device=None, **kw):
        if device is None or torch.device(device).type != "cuda":
            return torch.zeros(*shape, dtype=dtype, device=device, **kw)
        return self._alloc(shape, dtype, device, fill=0)

    def check(self):
        torch.cuda.synchronize()
        bad = 0
        for raw, nbytes in self.records:
            head = raw[:GUARD_BYTES]
            tail = raw[GUARD_BYTES + nbytes:]
            if not bool((head == PATTERN).all()) or not bool((tail == PATTERN).all()):
                bad += 1
        return len(self.records), bad


def main():
    import nbest_b200.model as M
    import nbest_b200.ops as O
    import nbest_b200.optim as P
    g = GuardedTorch()
    M.torch = O.torch = P.torch = g
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("sanitizer_unit", os.path.join(ROOT, "profiles", "sanitizer_unit.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)          # runs the unit workload
    finally:
        M.torch = O.torch = P.torch = torch
    n, bad = g.check()
    print("guard check: %d device buffers with 2 x %d-byte guard bands, %d violated" % (n, GUARD_BYTES, bad))
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
