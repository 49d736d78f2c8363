"""In-kernel timeline of the flash-attention kernel (mos_debug_set_attn_timeline): where one kv tile's time goes.
  python tools/attn_timeline.py [d nq nk]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'mix-of-show_b200')]
import torch  # noqa: E402

from mos_b200 import _lib, ops  # noqa: E402

lib = _lib.lib()
cases = [(40, 4096, 4096), (80, 1024, 1024)]
if len(sys.argv) == 4:
    cases = [tuple(int(v) for v in sys.argv[1:4])]
B, H = 2, 8
for d, nq, nk in cases:
    dp, dv, nk8 = (d + 63) // 64 * 64, (d + 15) // 16 * 16, (nk + 7) // 8 * 8
    Q = torch.zeros(B * H, nq, dp, device='cuda', dtype=torch.bfloat16)
    K = torch.zeros(B * H, nk, dp, device='cuda', dtype=torch.bfloat16)
    Vt = torch.zeros(B * H, dv, nk8, device='cuda', dtype=torch.bfloat16)
    Q[..., :d].normal_()
    K[..., :d].normal_()
    Vt[:, :d, :nk].normal_()
    out = torch.empty(B, nq, H * d, device='cuda', dtype=torch.bfloat16)
    tl = torch.zeros(512, dtype=torch.int64, device='cuda')
    for _ in range(2):
        ops.attention(Q, K, Vt, out, batch=B, heads=H, head_dim=d, nq=nq, nk=nk)
    torch.cuda.synchronize()
    lib.mos_debug_set_attn_timeline(ctypes.c_void_p(tl.data_ptr()))
    ops.attention(Q, K, Vt, out, batch=B, heads=H, head_dim=d, nq=nq, nk=nk)
    torch.cuda.synchronize()
    lib.mos_debug_set_attn_timeline(None)
    full = tl.cpu()
    t = full[:256].view(2, 32, 4)
    t0 = int(t[t > 0].min())
    print(f'--- d={d} nq={nq} nk={nk}: cycles since the first stamp (CTA 0,0)')
    print(' j | softmax: wait_S  S_seen  pass_done  published | mma: kv_landed  S_buf_free  before_P_wait  P_seen')
    for j in range(min(12, -(-nk // (128 if d <= 80 else 64)))):
        a = [int(v) - t0 if int(v) else -1 for v in t[0, j]]
        b = [int(v) - t0 if int(v) else -1 for v in t[1, j]]
        print(f'{j:2d} | {a[0]:8d} {a[1]:8d} {a[2]:8d} {a[3]:8d} | {b[0]:8d} {b[1]:8d} {b[2]:8d} {b[3]:8d}')
