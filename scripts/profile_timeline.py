"""Timeline of one denoising step (CUPTI via torch.profiler, CUDA graphs on): which kernels run while no convolution kernel is running,
i.e. what the step spends outside the conv layers on its critical path (development aid).
Usage: python scripts/profile_timeline.py [first step] [count]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device("cuda", 0)
    scan, start, g = bench.build_inputs(dev, 0)
    pipe = bench.build_pipeline(dev, scan)
    eng = pipe.engine()
    noise = torch.randn((50, bench.N_POINTS, 3), device=dev, generator=g)
    x_feats = (scan + start).float()
    for rep in range(2):                                       # second pass replays the captured graphs
        st = eng.start(scan, x_feats)
        for i in range(first):
            eng.advance(st, noise[i])
        torch.cuda.synchronize()
        if rep == 0:
            for i in range(first, first + count):
                eng.advance(st, noise[i])
            torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(first, first + count):
            eng.advance(st, noise[i])
        torch.cuda.synchronize()
    ev = []
    for e in prof.profiler.kineto_results.events():
        if "cuda" not in str(e.device_type()).lower() or e.duration_ns() <= 0:
            continue
        ev.append((e.start_ns(), e.start_ns() + e.duration_ns(), e.name(), e.device_resource_id()))
    ev.sort()
    t0, t1 = ev[0][0], max(e[1] for e in ev)
    print(f"{len(ev)} device activities over {(t1 - t0) / 1e6:.3f} ms for {count} steps = {(t1 - t0) / 1e6 / count:.3f} ms/step; streams: {sorted(set(e[3] for e in ev))}")
    conv = [(a, b) for a, b, n, s in ev if "spconv" in n]
    conv.sort()
    merged = []
    for a, b in conv:
        if merged and a <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], b)
        else:
            merged.append([a, b])
    busy = sum(b - a for a, b in merged)
    print(f"conv kernels cover {busy / 1e6 / count:.3f} ms/step; gaps without any conv kernel: {(t1 - t0 - busy) / 1e6 / count:.3f} ms/step")
    gaps = [(t0, merged[0][0])] + [(merged[i][1], merged[i + 1][0]) for i in range(len(merged) - 1)] + [(merged[-1][1], t1)]
    import collections
    share = collections.Counter()
    idle = 0
    big = []
    for a, b in gaps:
        if b - a <= 0:
            continue
        inside = [(max(a, x), min(b, y), n) for x, y, n, s in ev if y > a and x < b and "spconv" not in n]
        cover = 0
        cur = a
        for x, y, n in sorted(inside):
            if y <= cur:
                continue
            share[n.split("(")[0].replace("void ", "")[:60]] += y - max(x, cur)
            cover += y - max(x, cur)
            cur = max(cur, y)
        idle += (b - a) - cover
        if b - a > 20000:
            big.append((b - a, (a - t0) / 1e6, [n.split("(")[0].replace("void ", "")[:40] for x, y, n in sorted(inside)][:8]))
    print(f"  of which no kernel at all (launch gaps / dependencies): {idle / 1e6 / count:.3f} ms/step")
    for n, v in share.most_common(20):
        print(f"  {v / 1e6 / count:7.3f} ms/step  {n}")
    nnk = [(a, b - a, s_) for a, b, n, s_ in ev if "nn_match" in n or "compose" in n]
    print("NN kernels of the first profiled step (start ms, us, stream):", [(round((a - t0) / 1e6, 3), round(d / 1e3, 1), s_) for a, d, s_ in nnk[:len(nnk) // count]])
    nshow = int(os.environ.get("TIMELINE_HEAD", "0"))
    if nshow:                                                    # the window between the last conv of a step and the first conv after the stem of the next
        k0 = max(0, next(i for i, e in enumerate(ev) if "guidance_dpm" in e[2]) - 8)
        print("activities around the end of the first profiled step (start ms, us, stream, name):")
        for a, b, n, s_ in ev[k0:k0 + nshow]:
            print(f"  {(a - t0) / 1e6:8.3f} {(b - a) / 1e3:7.1f} {s_:4d} {n.split('(')[0].replace('void ', '')[:50]}")
    print("largest gaps (us, at ms, kernels inside):")
    for d, at, names in sorted(big, reverse=True)[:25]:
        print(f"  {d / 1e3:8.1f} us at {at:8.3f} ms: {names}")


if __name__ == "__main__":
    main()
