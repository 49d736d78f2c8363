"""Differential timing of one captured denoise step: replay the CUDA graph with one kernel family removed at a time.
(No nsys on this image; with PDL the families overlap slightly, so the parts do not sum exactly to the whole.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'mix-of-show_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from mos_b200.engine import UNetEngine, ehs_to_layer_major  # noqa: E402

sd, lora, lat, ehs, cfg = bench.build_workload(False)
eng = UNetEngine(sd, 2, 64, 64, lora=lora)
eng.in_ehs.copy_(ehs_to_layer_major(ehs.cuda(), 16))
eng.in_latents.normal_()
eng.in_t.fill_(981.0)


def timeit(skip):
    eng.skip = set(skip)
    eng.graph = None
    eng.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20


full = timeit([])
print(f'full step            {full:7.3f} ms   ({eng.launches} launches)')
for fam in ('gemm', 'splitk', 'attn', 'gn', 'ln'):
    t = timeit([fam])
    print(f'without {fam:8s}     {t:7.3f} ms   -> {fam} ~ {full - t:6.3f} ms')
t = timeit(['gemm', 'splitk', 'attn', 'gn', 'ln'])
print(f'only misc kernels    {t:7.3f} ms')
