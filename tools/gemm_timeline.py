"""In-kernel timeline of representative GEMM launches (uses mos_debug_set_timeline)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'mix-of-show_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402
from mos_b200 import _lib, ops  # noqa: E402

dev = torch.device('cuda')
tl = torch.zeros(64, dtype=torch.int64, device=dev)
lib = _lib.lib()


def mk(shape, scale=1.0):
    return (torch.randn(shape, device=dev) * scale).to(torch.bfloat16)


def run(name, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    back2back = e0.elapsed_time(e1) / 20 * 1e3
    lib.mos_debug_set_timeline(ctypes.c_void_p(tl.data_ptr()))
    tl.zero_()
    fn()
    torch.cuda.synchronize()
    lib.mos_debug_set_timeline(None)
    t = tl.view(8, 8).cpu()
    t0 = t[:, 0].min()
    rel = (t - t0).float() / 1e3
    names = ['start', 'setup', 'pdlwait', 'tma0', 'epi_pref', 'acc_rdy', 'acc_drn', 'written']
    print(f'{name}: back-to-back {back2back:.1f} us/launch')
    for b in (0, 1, 7):
        print('   cta', b, ' '.join(f'{n}={rel[b, i]:.2f}' for i, n in enumerate(names)))


M, N, K = 8192, 320, 320
A, W = mk((M, K)), mk((N, K), K ** -0.5)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
res = mk((M, N))
bias = torch.randn(N, device=dev)
run('plain 8192x320x320 +bias+res', lambda: ops.gemm(A, W, out, bias=bias, residual=res))
run('plain 8192x320x320', lambda: ops.gemm(A, W, out))
x = mk((2, 64, 64, 320))
wc = mk((320, 2880), 2880 ** -0.5)
run('conv 2x64x64 320->320', lambda: ops.gemm(x, wc, out, bias=bias, conv=(2, 64, 64, 320)))
A2, W2 = mk((8192, 320)), mk((2560, 320), 320 ** -0.5)
o2 = torch.empty(8192, 1280, device=dev, dtype=torch.bfloat16)
b2 = torch.randn(2560, device=dev)
run('geglu 8192x2560x320', lambda: ops.gemm(A2, W2, o2, bias=b2, geglu=True))
A3, W3 = mk((2048, 640)), mk((640, 640), 640 ** -0.5)
o3 = torch.empty(2048, 640, device=dev, dtype=torch.bfloat16)
run('plain 2048x640x640', lambda: ops.gemm(A3, W3, o3))
