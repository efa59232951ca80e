#!/usr/bin/env python
"""The windowed-GAE kernel alone at 2^18 windows x 128 steps (406 MB of algorithmic traffic), for an `ncu --set full`
capture of `gae_full_kernel` and a CUDA-event timing: the size at which the kernel is HBM-bound rather than launch-bound."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from surreal_b200 import ops  # noqa: E402


def main():
    B, n = (int(sys.argv[1]) if len(sys.argv) > 1 else 262144), 128
    dev = torch.device('cuda:0')
    g = torch.Generator(device=dev).manual_seed(B)
    r = torch.randn(B, n, device=dev, generator=g)
    v = torch.randn(B, n + 1, device=dev, generator=g)
    d = (torch.rand(B, n, device=dev, generator=g) < 0.01).float()
    adv, ret = torch.empty(B, 1, device=dev), torch.empty(B, 1, device=dev)
    flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)
    ms = []
    for _ in range(8):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gae_window(r, v, d, 0.995, 0.97, adv=adv, ret=ret)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms = sorted(ms[2:])
    byts = B * ((3 * n + 1) * 4 + 8)
    print('gae %d windows x %d: %.1f us, %.0f GB/s algorithmic' % (B, n, ms[len(ms) // 2] * 1e3, byts / (ms[len(ms) // 2] / 1e3) / 1e9))


if __name__ == '__main__':
    main()
