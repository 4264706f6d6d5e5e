"""Times the phases of one denoising step of the fused engine on a 180k-point synthetic scan
(development aid; bench.py is the contract).  Usage: python scripts/profile_step.py [n_part] [steps] [algo]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidiff_b200.engine import DenoiseEngine            # noqa: E402
from lidiff_b200.preprocess import farthest_point_sample  # noqa: E402
from lidiff_b200.synth import range_filter, synthetic_scan  # noqa: E402
from lidiff_b200.weights import random_state_dict         # noqa: E402


def main():
    n_part = int(sys.argv[1]) if len(sys.argv) > 1 else 18000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    algo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    dev = "cuda:0"
    raw = torch.tensor(range_filter(synthetic_scan(0)), device=dev)
    torch.cuda.synchronize()
    t0 = time.time()
    sel = farthest_point_sample(raw, n_part)
    torch.cuda.synchronize()
    print(f"FPS {raw.shape[0]} -> {n_part}: {time.time() - t0:.3f} s")
    scan = raw[sel].repeat(10, 1)
    N = scan.shape[0]
    sd_e, sd_d = random_state_dict("enc", 0), random_state_dict("diff", 1)
    eng = DenoiseEngine(sd_e, sd_d, device=dev, n_points=N, denoising_steps=50, conv_algo=algo)
    g = torch.Generator(device=dev).manual_seed(0)
    x = (scan + torch.randn(scan.shape, device=dev, generator=g, dtype=torch.float64)).float()
    noise = torch.randn((steps + 3, N, 3), device=dev, generator=g)
    torch.cuda.synchronize()
    t0 = time.time()
    eng.set_condition(scan)
    torch.cuda.synchronize()
    print(f"set_condition (cond encoder, once per scan): {time.time() - t0:.3f} s; cond level sizes {eng.geom_cond.sizes()}")
    xa, xb = x.contiguous(), torch.empty_like(x)
    ca, cb = torch.zeros(N, 4, device=dev), torch.zeros(N, 4, device=dev)
    ca[:, 1:] = torch.round(xa * 20.0)
    x0s = torch.zeros(N, 3, dtype=torch.float64, device=dev)
    for sigma_tag, first in (("sigma=1 (first steps)", 0),):
        l0 = eng.h.launch_count()
        for i in range(3):
            eng.step(i, xa, xb, ca, cb, scan, noise[i], x0s)
        torch.cuda.synchronize()
        print("level sizes", eng.geom.sizes(), "launches/step", (eng.h.launch_count() - l0) // 3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        w0 = time.time()
        for i in range(steps):
            eng.step(3 + i, xa, xb, ca, cb, scan, noise[3 + i], x0s)     # same geometry every time (xa not advanced)
        e1.record()
        torch.cuda.synchronize()
        print(f"{sigma_tag}: {e0.elapsed_time(e1) / steps:.2f} ms/step on device, host wall {1e3 * (time.time() - w0) / steps:.2f} ms/step")
    # phase breakdown with events
    ev = lambda: torch.cuda.Event(enable_timing=True)
    g_ = eng.geom
    marks = [ev() for _ in range(6)]
    marks[0].record()
    g_.build(ca, N)
    marks[1].record()
    for l in range(5):
        eng.h.nn_match_grid(g_.C[l], g_.d_n[l], N, eng.part_C, eng.part_dn, eng.part_cap, eng.part_grid, 16, 4, eng.buf(f"nn{l}", (N,), torch.int32))
    marks[2].record()
    eng._gate_tables(eng.A_cond, eng.part_cap, eng.part_dn, 0, "c")
    marks[3].record()
    torch.cuda.synchronize()
    print(f"geometry build {marks[0].elapsed_time(marks[1]):.2f} ms | nn_match x5 {marks[1].elapsed_time(marks[2]):.2f} ms | "
          f"gate tables {marks[2].elapsed_time(marks[3]):.2f} ms")
    print("status", eng.h.read_status(), "mem GB", torch.cuda.max_memory_allocated() / 2**30)


if __name__ == "__main__":
    main()
