"""Per-layer-shape timing of the tcgen05 GEMM in both tile modes (1-CTA 128 x 160 tiles vs 2-CTA pair 256 x 160 tiles,
mos_gemm_args.pair_mode): replays every DISTINCT mos_gemm_bf16 call of one denoise step (same tensors, same epilogue
options) in isolation, L2 flushed before every timed launch (the step streams 1.7 GB of weights, so weights are cold in the
real step), CUDA events.  Writes gpurun_out/gemm_shape_bench.csv and prints the per-step totals.
    python tools/gemm_shape_bench.py [--reps 10]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'mix-of-show_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--reps', type=int, default=10)
ap.add_argument('--warm', action='store_true', help='no L2 flush between launches')
args = ap.parse_args()
from mos_b200 import ops  # noqa: E402
from mos_b200.engine import UNetEngine, ehs_to_layer_major  # noqa: E402

sd, lora, lat, ehs, cfg = bench.build_workload(False)
eng = UNetEngine(sd, 2, 64, 64, lora=lora, use_graph=False)
eng.in_ehs.copy_(ehs_to_layer_major(ehs.cuda(), 16))
eng.in_latents.normal_()
eng.in_t.fill_(981.0)
eng._run()
torch.cuda.synchronize()
calls = []
_orig = ops.gemm


def _rec(A, W, out=None, **kw):
    calls.append((A, W, out, dict(kw)))
    return _orig(A, W, out, **kw)


ops.gemm = _rec
eng._run()
torch.cuda.synchronize()
ops.gemm = _orig


def key(c):
    A, W, out, kw = c
    conv = kw.get('conv')
    M = conv[0] * conv[1] * conv[2] if conv is not None else (kw.get('M') or A.shape[0])
    return (M, W.shape[0], W.shape[1], int(conv is not None), int(kw.get('lora_down') is not None), int(bool(kw.get('geglu'))),
            int(kw.get('heads') is not None), kw.get('splits') or 1, int(kw.get('residual') is not None))


groups = {}
for c in calls:
    groups.setdefault(key(c), []).append(c)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def time_call(c, mode):
    A, W, out, kw = c
    kw = dict(kw, pair_mode=mode)
    ts = []
    for _ in range(args.reps + 2):
        if not args.warm:
            flush.fill_(1)
            # the activation operand comes from the previous kernel in the real step: touch it back into L2
            A.view(-1)[:1].add_(0) if False else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _orig(A, W, out, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    return ts[len(ts) // 2]


rows = []
tot = {1: 0.0, 2: 0.0, 'best': 0.0}
for k, cs in sorted(groups.items(), key=lambda kv: -len(kv[1])):
    M, N, K = k[0], k[1], k[2]
    mt = (M + 127) // 128
    t2 = time_call(cs[0], 2)
    t1 = time_call(cs[0], 1) if mt % 2 == 0 else float('nan')
    n = len(cs)
    fl = 2.0 * M * N * K
    rows.append(k + (n, t2, t1, fl / (t2 * 1e-6) / 1e12))
    tot[2] += n * t2
    tot[1] += n * (t1 if t1 == t1 else t2)
    tot['best'] += n * (min(t1, t2) if t1 == t1 else t2)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
with open(os.path.join(ROOT, 'gpurun_out', 'gemm_shape_bench.csv'), 'w') as f:
    f.write('M,N,K,conv,lora,geglu,heads,splits,residual,launches_per_step,us_single,us_pair,tflops_single\n')
    for r in rows:
        f.write(','.join(f'{v:.2f}' if isinstance(v, float) else str(v) for v in r) + '\n')
for r in rows:
    print(' '.join(f'{v:9.2f}' if isinstance(v, float) else f'{v:6d}' for v in r))
print(f'per-step sum (isolated launches, L2 {"warm" if args.warm else "flushed"}): single {tot[2] / 1e3:.3f} ms, pair where possible '
      f'{tot[1] / 1e3:.3f} ms, best of both {tot["best"] / 1e3:.3f} ms over {len(calls)} launches')
