"""Debug aid: run the tiny inference engine eagerly and, for every GroupNorm call, compare the one-pass cluster kernel with
the two-launch path on the SAME input (prints shape / pitch / pointer alignment of the first mismatches)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'mix-of-show_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from mos_b200 import _lib, ops  # noqa: E402
from mos_b200.engine import UNetEngine, ehs_to_layer_major  # noqa: E402

tiny = '--full' not in sys.argv
sd, lora, lat, ehs, cfg = bench.build_workload(tiny)
kw = dict(block_out=cfg['block_out_channels'], layers=cfg['layers_per_block']) if cfg else {}
H = 32 if tiny else 64
for dt in (torch.float16, torch.bfloat16):
    eng = UNetEngine(sd, 2, H, H, lora=lora, use_graph=False, act_dtype=dt, **kw)
    nx = len(eng.xattn_names)
    eng.in_ehs.copy_(ehs_to_layer_major(ehs[:, :nx].cuda(), nx, dt))
    eng.in_latents.normal_()
    eng.in_t.fill_(981.0)
    orig = ops.groupnorm
    n = {'calls': 0, 'bad': 0}

    def both(x, gamma, beta, y, partial, **k):
        _lib.lib().mos_debug_set_gn_twopass(1)
        y2 = torch.empty_like(y)
        orig(x, gamma, beta, y2, partial, **dict(k, ldy=y2.stride(-2)))
        torch.cuda.synchronize()
        _lib.lib().mos_debug_set_gn_twopass(0)
        orig(x, gamma, beta, y, partial, **k)
        torch.cuda.synchronize()
        n['calls'] += 1
        d = (y.float() - y2.float()).abs().max().item()
        if d > 2e-2:
            n['bad'] += 1
            if n['bad'] <= 6:
                rows = (y.float() - y2.float()).abs().amax(-1).flatten()
                print(f'  MISMATCH {dt} call {n["calls"]}: B={k["B"]} HW={k["HW"]} C={k["C"]} ldx={k.get("ldx")} ldy={k.get("ldy")} '
                      f'x.ptr%16={x.data_ptr() % 16} y.ptr%16={y.data_ptr() % 16} silu={k["silu"]} max|d|={d:.3f} '
                      f'bad rows {int((rows > 2e-2).sum())}/{rows.numel()} first {int((rows > 2e-2).nonzero()[0])}', flush=True)
        return y

    ops.groupnorm = both
    eng._run()
    torch.cuda.synchronize()
    ops.groupnorm = orig
    print(f'{dt}: {n["calls"]} GroupNorm calls, {n["bad"]} mismatching')
