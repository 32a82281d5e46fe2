"""Timing of spi_rows_outer_sum (decoder-gradient reduction) against the batched-GEMM formulation it replaced.  Not a benchmark of the step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from spi_b200 import _lib


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    lib = _lib.load()
    torch.backends.cuda.matmul.allow_tf32 = True
    for rows in (1 << 20, 1 << 22):
        for cu, cv in ((64, 32), (36, 64)):
            u = torch.randn(rows, cu, device='cuda')
            v = torch.randn(rows, cv, device='cuda')
            out, usum = torch.empty(cu, cv, device='cuda'), torch.empty(cu, device='cuda')
            mb = (u.numel() + v.numel()) * 4 / 1e6
            t1 = timed(lambda: _lib.check(lib.spi_rows_outer_sum(_lib.ptr(u), _lib.ptr(v), rows, cu, cv, _lib.ptr(out), _lib.ptr(usum), _lib.stream())))
            t2 = timed(lambda: _lib.check(lib.spi_rows_outer_sum(_lib.ptr(u), _lib.ptr(v), rows, cu, cv, _lib.ptr(out), None, _lib.stream())))
            t3 = timed(lambda: torch.bmm(u.view(512, rows // 512, cu).transpose(1, 2), v.view(512, rows // 512, cv)).sum(0))
            s = torch.empty(cu, device='cuda')
            t4 = timed(lambda: _lib.check(lib.spi_column_sums(_lib.ptr(u), rows, cu, _lib.ptr(s), _lib.stream())))
            print(f'rows {rows} u[{cu}] v[{cv}] ({mb:.0f} MB): outer+colsum {1e3 * t1:.1f} us ({mb / t1 / 1e3:.2f} TB/s)  outer only {1e3 * t2:.1f} us  '
                  f'bmm512+sum {1e3 * t3:.1f} us  column_sums {1e3 * t4:.1f} us', flush=True)


if __name__ == '__main__':
    main()
