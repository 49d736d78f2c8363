"""Mainloop probe of the tcgen05 GEMM on the res-64 3x3 conv (8192 x 320 x 2880, 45 k-blocks, one tile per CTA on 128 SMs):
time per launch against the smem pipeline depth (mos_gemm_args.stages) and against the same reduction as a plain 2-D TMA
GEMM.  If us/launch scales with 1/stages the mainloop is bound by the bytes in flight per SM (latency), not by a bandwidth
ceiling.  Back-to-back launches (operands L2-resident, as the activations are in the real step), CUDA events.
    python tools/gemm_stage_sweep.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'mix-of-show_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402

from mos_b200 import ops  # noqa: E402

dev = torch.device('cuda')
g = torch.Generator().manual_seed(0)
f16 = torch.float16
xa = torch.randn(2, 64, 64, 320, generator=g).to(dev, f16)
wc = (torch.randn(320, 2880, generator=g) * 2880 ** -0.5).to(dev, f16)
ap = torch.randn(8192, 2880, generator=g).to(dev, f16)
out = torch.empty(8192, 320, device=dev, dtype=f16)
bias = torch.zeros(320, device=dev)


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


fl = 2.0 * 8192 * 320 * 2880
for mode, name in ((2, '1-CTA 128x160'), (1, '2-CTA pair 256x160')):
    for st in (2, 3, 4, 5, 6):
        try:
            t = timeit(lambda: ops.gemm(xa, wc, out, bias=bias, conv=(2, 64, 64, 320), stages=st, pair_mode=mode))
            print(f'conv  {name:20s} stages {st}: {t:7.2f} us  {fl / t / 1e6:7.1f} TFLOP/s  {t / 45 * 1e3:6.1f} ns per k-block (incl. fixed cost)')
        except Exception as e:
            print(f'conv  {name} stages {st}: {e}')
for st in (2, 3, 4):
    t = timeit(lambda: ops.gemm(ap, wc, out, bias=bias, stages=st, pair_mode=2))
    print(f'plain 1-CTA 128x160       stages {st}: {t:7.2f} us  {fl / t / 1e6:7.1f} TFLOP/s')
# fixed cost: the same tile grid with 5 k-blocks
a5 = torch.randn(8192, 320, generator=g).to(dev, f16)
w5 = (torch.randn(320, 320, generator=g) * 320 ** -0.5).to(dev, f16)
t = timeit(lambda: ops.gemm(a5, w5, out, bias=bias, pair_mode=2))
print(f'plain 8192x320x320 (5 k-blocks): {t:7.2f} us back to back')
