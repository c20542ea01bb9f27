#!/usr/bin/env python
"""Host link probe (GPU box): GB/s of pinned H2D, D2H, and both at once, one stream each vs. split over several streams;
what the end-to-end leg of bench.py can expect.   python scripts/pcie_probe.py [--mb 1900]"""
import argparse
import json
import torch


def timed(fn, iters=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1900)
    args = ap.parse_args()
    n = args.mb * 2**20 // 4
    h_in, h_out = torch.empty(n).pin_memory(), torch.empty(n).pin_memory()
    d_in, d_out = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
    gb = n * 4 / 1e9
    cur = torch.cuda.current_stream()
    streams = [torch.cuda.Stream() for _ in range(8)]

    def split(parts, h2d=True, d2h=True):
        def run():
            step = n // parts
            for s in streams[: 2 * parts]:
                s.wait_stream(cur)
            for i in range(parts):
                sl = slice(i * step, n if i == parts - 1 else (i + 1) * step)
                if h2d:
                    with torch.cuda.stream(streams[2 * i]):
                        d_in[sl].copy_(h_in[sl], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(streams[2 * i + 1]):
                        h_out[sl].copy_(d_out[sl], non_blocking=True)
            for s in streams[: 2 * parts]:
                cur.wait_stream(s)
        return run
    out = {"GB_each_way": round(gb, 3)}
    out["h2d_only_gbs"] = round(gb / timed(split(1, True, False)) * 1e3, 1)
    out["d2h_only_gbs"] = round(gb / timed(split(1, False, True)) * 1e3, 1)
    for parts in (1, 2, 4):
        ms = timed(split(parts))
        out["both_%d_streams_each_gbs_per_direction" % parts] = round(gb / ms * 1e3, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
