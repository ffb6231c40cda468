"""HBM bandwidth of a B200 by access mix and transfer size (torch ops, CUDA events, best of 10): what denominator a
write-only kernel (conv0 forward: 3.1 GB written, ~0.1 GB read) or a small transfer (LayerNorm: 74 MB) can be held to.
  copy   b.copy_(a)   N bytes read + N bytes written  (MEASURED_PEAKS.json hbm_gbs is this at 2 GiB)
  write  a.zero_() / a.fill_(1)  N bytes written
  read   a.sum()      N bytes read
usage: python tools/micro/hbm_rw.py > gpurun_out/<tag>_hbm_rw.txt"""
import json

import torch


def best(fn, n=10, flush=None):
    torch.cuda.synchronize()
    t = []
    for _ in range(n):
        if flush is not None:
            flush.zero_()          # every timed run starts from an L2 that holds none of the operands
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        t.append(a.elapsed_time(b))
    return min(t)


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for mb in (37, 74, 147, 512, 1536, 3072):
        n = mb << 20
        a = torch.empty(n // 2, dtype=torch.bfloat16, device="cuda").normal_()
        b = torch.empty_like(a)
        rec = {"MB": mb}
        for name, fn, moved in (("copy", lambda: b.copy_(a), 2 * n), ("write_zero", lambda: a.zero_(), n),
                                ("write_fill", lambda: a.fill_(1.0), n), ("read_sum", lambda: a.sum(), n)):
            ms = best(fn, flush=flush)
            rec[name + "_GBs"] = round(moved / ms / 1e6, 1)
            rec[name + "_us"] = round(ms * 1e3, 1)
            ms = best(fn)
            rec[name + "_warm_GBs"] = round(moved / ms / 1e6, 1)
        print(json.dumps(rec), flush=True)
        del a, b


if __name__ == "__main__":
    main()
