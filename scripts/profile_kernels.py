"""Kernel-level time table (CUPTI via torch.profiler) of chosen steps of the T=50 trajectory (development aid).
Usage: python scripts/profile_kernels.py [first step] [count]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dev = torch.device("cuda", 0)
    scan, start, g = bench.build_inputs(dev, 0)
    pipe = bench.build_pipeline(dev, scan)
    eng = pipe.engine()
    noise = torch.randn((50, bench.N_POINTS, 3), device=dev, generator=g)
    x_feats = (scan + start).float()
    st = eng.start(scan, x_feats)
    for i in range(3):
        eng.advance(st, noise[i])
    st = eng.start(scan, x_feats)
    for i in range(first):
        eng.advance(st, noise[i])
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(first, first + count):
            eng.advance(st, noise[i])
        torch.cuda.synchronize()
    rows = [(e.key, e.count, e.device_time_total / 1e3 / count) for e in prof.key_averages() if e.device_time_total > 0]
    tot = sum(r[2] for r in rows)
    print(f"steps {first}..{first + count - 1}: {tot:.2f} ms of kernels per step")
    for k, c, ms in sorted(rows, key=lambda r: -r[2])[:40]:
        print(f"  {ms:8.3f} ms  {c / count:6.1f} launches  {k[:110]}")


if __name__ == "__main__":
    main()
