#!/usr/bin/env python
"""HBM bandwidth of torch's own kernels for traffic mixes other than a 1:1 copy (development aid; needs a B200):
pure write (fill), 1 byte read : 4 bytes written (uint8 -> float32 cast, the float observation mode's mix)."""
import torch


def timed(fn, k=20, w=3):
    for _ in range(w):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


def main():
    n = 1 << 30
    a = torch.empty(n, dtype=torch.uint8, device="cuda").random_(0, 255)
    b = torch.empty(n, dtype=torch.uint8, device="cuda")
    f = torch.empty(n // 4, dtype=torch.float32, device="cuda")
    u = a[: n // 4]
    print("copy u8 1 GiB        : %.0f GB/s" % (2 * n / timed(lambda: b.copy_(a)) / 1e6))
    print("fill 1 GiB (write)   : %.0f GB/s" % (n / timed(lambda: b.fill_(7)) / 1e6))
    print("u8 -> f32 cast 1:4   : %.0f GB/s" % ((n // 4 + n) / timed(lambda: f.copy_(u)) / 1e6))
    print("read-only sum 1 GiB  : %.0f GB/s" % (n / timed(lambda: a.view(torch.int64).sum()) / 1e6))


if __name__ == "__main__":
    main()
