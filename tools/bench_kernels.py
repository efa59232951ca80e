#!/usr/bin/env python
"""Micro-benchmarks of individual kernels (CUDA events, warm L2 unless --flush): run on the GPU box.
    python tools/bench_kernels.py critic | small | all
Prints one JSON line per measurement; used to pick tile variants (SB200_FWD_TM must be set per process)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from surreal_b200 import ops  # noqa: E402


def timeit(fn, reps=50, flush=None):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        ts.append((e0, e1))
    torch.cuda.synchronize()
    v = sorted(a.elapsed_time(b) for a, b in ts)
    return v[len(v) // 2] * 1e3, v[0] * 1e3        # median, min (us)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(0)
    D, H, A = 64, 256, 8
    flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)

    def net(out):
        n = ops.FlatNet([D, H, H, out], [ops.ACT_RELU, ops.ACT_RELU, ops.ACT_NONE], dev)
        n.params.copy_(torch.randn(n.size, generator=g) * 0.05)
        return n
    zs = torch.cat([torch.zeros(D), torch.ones(D), torch.ones(1)]).to(dev)
    tag = dict(tm=os.environ.get('SB200_FWD_TM', 'auto'), no_skinny=os.environ.get('SB200_NO_SKINNY', '0'))
    if what in ('critic', 'all'):
        rows = 1024 * 129
        x = torch.randn(rows, D, device=dev)
        out = torch.empty(rows, 1, device=dev)
        c = net(1)
        med, mn = timeit(lambda: ops.mlp_forward(c, x, zf_stats=zs, out=out), flush=flush)
        flop = 2.0 * rows * (D * H + H * H + H)
        print(json.dumps(dict(kernel='critic_pass', rows=rows, median_us=med, min_us=mn, tflops=flop / med / 1e6, **tag)))
        for v, name in ((1, 'mma<2>'), (2, 'ffma<4,16>')):
            med, mn = timeit(lambda: ops.mlp_forward(c, x, zf_stats=zs, out=out, variant=v), flush=flush)
            print(json.dumps(dict(kernel='critic_pass_variant', variant=name, median_us=med, min_us=mn, tflops=flop / med / 1e6)))
        for frac in (0.35, 0.45, 0.5, 0.55, 0.65):
            med, mn = timeit(lambda: ops.mlp_forward_dual(c, x, zf_stats=zs, out=out, frac=frac), flush=flush)
            print(json.dumps(dict(kernel='critic_pass_dual', frac_tensor=frac, median_us=med, min_us=mn, tflops=flop / med / 1e6)))
    if what in ('small', 'all'):
        for rows in (1024, 4096):
            x = torch.randn(rows, D, device=dev)
            a = net(A)
            out = torch.empty(rows, A, device=dev)
            med, mn = timeit(lambda: ops.mlp_forward(a, x, zf_stats=zs, out=out), reps=200)
            flop = 2.0 * rows * (D * H + H * H + H * A)
            print(json.dumps(dict(kernel='actor_forward', rows=rows, median_us=med, min_us=mn, tflops=flop / med / 1e6, **tag)))
            # back-to-back inside a CUDA graph: what the rollout actually sees
            gr = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(gr):
                for _ in range(64):
                    ops.mlp_forward(a, x, zf_stats=zs, out=out)
            gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                gr.replay()
            e1.record()
            torch.cuda.synchronize()
            print(json.dumps(dict(kernel='actor_forward_in_graph', rows=rows, us_per_launch=e0.elapsed_time(e1) * 1e3 / 320, **tag)))
            pk = ops.PackedWeights(a).refresh()
            med, mn = timeit(lambda: ops.mlp_forward_packed(pk, x, zf_stats=zs, out=out), reps=200)
            print(json.dumps(dict(kernel='actor_forward_packed', rows=rows, median_us=med, min_us=mn, tflops=flop / med / 1e6)))
            gr = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(gr):
                for _ in range(64):
                    ops.mlp_forward_packed(pk, x, zf_stats=zs, out=out)
            gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                gr.replay()
            e1.record()
            torch.cuda.synchronize()
            print(json.dumps(dict(kernel='actor_forward_packed_in_graph', rows=rows, us_per_launch=e0.elapsed_time(e1) * 1e3 / 320)))


if __name__ == '__main__':
    main()
