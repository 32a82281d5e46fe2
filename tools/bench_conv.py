"""Correctness + timing of the tcgen05 implicit-GEMM convolution (spi_conv2d_tc) against cuDNN TF32 and an fp64 reference.
Not a benchmark of the product path; run under gpurun."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

from spi_b200 import _lib

L = _lib.load()


def conv_tc(x, w, per_sample, bias=None, noise=None, strength=None, act=0, slope=0.2, gain=1.0, clamp=-1.0, flags=0):
    n, ci, h, wd = x.shape
    co, kh, kw = w.shape[-4], w.shape[-3], w.shape[-2]
    y = torch.empty(n, co, h, wd, device=x.device).contiguous(memory_format=torch.channels_last)
    _lib.check(L.spi_conv2d_tc(_lib.ptr(x), _lib.ptr(w), _lib.ptr(y), n, h, wd, ci, co, kh, kw, int(per_sample), _lib.ptr(bias), _lib.ptr(noise),
                               _lib.ptr(strength), act, slope, gain, clamp, flags, _lib.stream()))
    return y


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def time_ms(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def case(n, ci, co, h, k, per_sample, check=True):
    torch.manual_seed(0)
    x = torch.randn(n, ci, h, h, device='cuda').contiguous(memory_format=torch.channels_last)
    g = n if per_sample else 1
    w = torch.randn(g, co, k, k, ci, device='cuda') / (k * k * ci) ** 0.5          # [G][O][KH][KW][I]
    w_oikk = w.permute(0, 1, 4, 2, 3)                                              # logical [G,O,I,kh,kw]
    out = {}
    for flags, name in ((0, 'rna'), (1, 'trunc')):
        y = conv_tc(x, w, per_sample, flags=flags)
        err = L.spi_conv2d_tc_error()
        out[name] = y
        if err:
            print(f'  !! barrier time-out flag {err}  ({n} {ci}->{co} {h}^2 k={k})', flush=True)
            return
    msg = ''
    if check:
        xs, ws = x.double(), w_oikk.double()
        ref = torch.cat([F.conv2d(xs[i:i + 1], ws[i if per_sample else 0], padding=k // 2) for i in range(n)])
        torch.backends.cudnn.allow_tf32 = True
        yc = torch.cat([F.conv2d(x[i:i + 1], w_oikk[i if per_sample else 0].contiguous(memory_format=torch.channels_last), padding=k // 2)
                        for i in range(n)])
        sc = ref.abs().mean()
        msg = (f"rel-L2 vs fp64: tc05(rna) {rel(out['rna'], ref):.2e} tc05(trunc) {rel(out['trunc'], ref):.2e} cudnn-tf32 {rel(yc, ref):.2e}"
               f" | signed bias of |y|: rna {float(((out['rna'].double().abs() - ref.abs())).mean() / sc):.1e}"
               f" trunc {float(((out['trunc'].double().abs() - ref.abs())).mean() / sc):.1e} cudnn {float(((yc.double().abs() - ref.abs())).mean() / sc):.1e}")
    gf = 2 * n * h * h * k * k * ci * co / 1e9
    t_tc = time_ms(lambda: conv_tc(x, w, per_sample))
    wc = [w_oikk[i].contiguous(memory_format=torch.channels_last) for i in range(g)]
    torch.backends.cudnn.allow_tf32 = True
    y2 = torch.empty(n, co, h, h, device='cuda').contiguous(memory_format=torch.channels_last)

    def cudnn():
        if per_sample:
            for i in range(n):
                torch.ops.aten.cudnn_convolution.out(x[i:i + 1], wc[i], [k // 2, k // 2], [1, 1], [1, 1], 1, False, False, True, out=y2[i:i + 1])
        else:
            torch.ops.aten.cudnn_convolution.out(x, wc[0], [k // 2, k // 2], [1, 1], [1, 1], 1, False, False, True, out=y2)
    t_cd = time_ms(cudnn)
    print(f'n={n} {ci:4d}->{co:4d} {h:4d}^2 k={k} per_sample={int(per_sample)} {gf:7.1f} GF  tc05 {t_tc:7.3f} ms ({gf / t_tc:6.1f} TF/s)  '
          f'cudnn {t_cd:7.3f} ms ({gf / t_cd:6.1f} TF/s)  {msg}', flush=True)


def epilogue_case():
    torch.manual_seed(1)
    n, ci, co, h = 2, 64, 96, 40
    x = torch.randn(n, ci, h, h, device='cuda').contiguous(memory_format=torch.channels_last)
    w = torch.randn(1, co, 3, 3, ci, device='cuda') / (9 * ci) ** 0.5
    b = torch.randn(co, device='cuda')
    nz = torch.randn(h, h, device='cuda')
    st = torch.tensor(0.7, device='cuda')
    y = conv_tc(x, w, False, bias=b, noise=nz, strength=st, act=2, slope=0.2, gain=2 ** 0.5, clamp=1.5)
    ref = F.conv2d(x.double(), w[0].permute(0, 3, 1, 2).double(), padding=1) + nz.double() * 0.7 + b.double().view(1, -1, 1, 1)
    ref = (F.leaky_relu(ref, 0.2) * 2 ** 0.5).clamp(-1.5, 1.5)
    print('epilogue (noise+bias+lrelu+gain+clamp, ragged 40x40, O=96): rel-L2', rel(y, ref), 'err flag', L.spi_conv2d_tc_error(), flush=True)


if __name__ == '__main__':
    print(torch.cuda.get_device_name(0), flush=True)
    case(1, 32, 32, 16, 3, False)
    case(1, 64, 64, 32, 1, False)
    case(2, 64, 128, 40, 3, True)
    epilogue_case()
    case(1, 128, 128, 512, 3, True)
    case(1, 256, 256, 256, 3, True)
    case(1, 128, 128, 256, 3, True)
    case(1, 256, 256, 128, 3, True)
    case(1, 512, 512, 64, 3, True)
    case(4, 128, 128, 512, 3, False, check=False)
    case(1, 64, 64, 256, 3, False)
    case(1, 128, 96, 256, 1, True)
    case(2, 512, 512, 32, 3, False)
