"""One launch of every kernel family that matters for the step / the training step / the fusion, at its dominant shape — the
command `ncu --set full` wraps for the per-kernel evidence in profiles/ (round 2):

  ncu --set full --clock-control none --import-source on -k regex:'attn_kernel|attn_bwd|gn_group|gn_stats|gn_apply|layernorm|lora_grad|dgemm_mixed|gemm_kernel|splitk|softmax_rows' \
      -o gpurun_out/r2_kernels python tools/ncu_targets.py

Every family runs twice (first = warm-up, set kernel attributes); ncu's -s / launch-skip is not needed: the capture keeps
all instances and profiles/README.md quotes the second of each."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'mix-of-show_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402

from mos_b200 import ops  # noqa: E402

dev = torch.device('cuda')
g = torch.Generator().manual_seed(0)


def rnd(shape, dtype=torch.float16, scale=1.0):
    return (torch.randn(shape, generator=g) * scale).to(dev).to(dtype)


def heads(q, d):
    BH, n, _ = q.shape
    dp = (d + 63) // 64 * 64
    Q = torch.zeros(BH, n, dp, device=dev, dtype=q.dtype)
    Q[..., :d] = q
    return Q


def run(name, fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    print('ran', name, flush=True)


# ---- attention forward, res-64 self-attention of the CFG step: B*H = 16, 4096 x 4096, d = 40, fp16
B, H, d, n = 2, 8, 40, 4096
q, k, v = (rnd((B * H, n, d)) for _ in range(3))
Q, K = heads(q, d), heads(k, d)
Vt = torch.zeros(B * H, 48, n, device=dev, dtype=torch.float16)
Vt[:, :d] = v.transpose(1, 2)
out = torch.empty(B, n, H * d, device=dev, dtype=torch.float16)
run('attn_fwd d40 4096^2 fp16', lambda: ops.attention(Q, K, Vt, out, batch=B, heads=H, head_dim=d, nq=n, nk=n))
# cross attention 4096 x 77
kc, vc = rnd((B * H, 77, d)), rnd((B * H, 77, d))
Kc = heads(kc, d)
Vc = torch.zeros(B * H, 48, 80, device=dev, dtype=torch.float16)
Vc[:, :d, :77] = vc.transpose(1, 2)
run('attn_fwd cross d40 4096x77 fp16', lambda: ops.attention(Q, Kc, Vc, out, batch=B, heads=H, head_dim=d, nq=n, nk=77))

# ---- attention backward (training, bf16), batch 4: B*H = 32, 4096 x 4096, d = 40
Bt = 4
bf = torch.bfloat16
qb, kb, vb, dob = (rnd((Bt * H, n, d), bf) for _ in range(4))
Qb, Kb, Vb, dOb = heads(qb, d), heads(kb, d), heads(vb, d), heads(dob, d)
Vtb = torch.zeros(Bt * H, 48, n, device=dev, dtype=bf)
ops.heads_transpose(Vb, Vtb)
ob = torch.empty(Bt, n, H * d, device=dev, dtype=bf)
lse = torch.empty(Bt * H, n, device=dev)
ops.attention_train(Qb, Kb, Vtb, ob, lse, batch=Bt, heads=H, head_dim=d, nq=n, nk=n)
Qt, Kt, dOt = (torch.zeros(Bt * H, 48, n, device=dev, dtype=bf) for _ in range(3))
ops.heads_transpose(Qb, Qt)
ops.heads_transpose(Kb, Kt)
ops.heads_transpose(dOb, dOt)
delta = torch.empty(Bt * H, n, device=dev)
ops.attn_delta(dOb, ob.view(Bt * n, H * d), delta, batch=Bt, heads=H, head_dim=d, N=n, ldo=H * d)
dqkv = torch.empty(Bt * n, 3 * H * d, device=dev, dtype=bf)
C = H * d
run('attn_bwd d40 4096^2 bf16 batch 4',
    lambda: ops.attention_bwd(Qb, Kb, Vb, dOb, Qt, Kt, dOt, lse, delta, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], batch=Bt,
                              heads=H, head_dim=d, nq=n, nk=n, lddq=3 * C, lddk=3 * C, lddv=3 * C))

# ---- GroupNorm (res-64, 320 channels, CFG batch 2): one-pass cluster kernel and the two-launch path; LayerNorm
from mos_b200 import _lib  # noqa: E402
x = rnd((2, 4096, 320))
gam, bet = torch.randn(320, device=dev), torch.randn(320, device=dev)
y = torch.empty_like(x)
part = torch.zeros(2 * 592 * 64, device=dev)
run('groupnorm one-pass 2x4096x320', lambda: ops.groupnorm(x, gam, bet, y, part, B=2, HW=4096, C=320, eps=1e-5, silu=True))
_lib.lib().mos_debug_set_gn_twopass(1)
run('groupnorm two-pass 2x4096x320', lambda: ops.groupnorm(x, gam, bet, y, part, B=2, HW=4096, C=320, eps=1e-5, silu=True))
_lib.lib().mos_debug_set_gn_twopass(0)
xl = rnd((8192, 320))
yl = torch.empty_like(xl)
run('layernorm 8192x320', lambda: ops.layernorm(xl, gam, bet, yl, M=8192, C=320))

# ---- GEMM: res-64 3x3 conv 320 -> 320 (45 k-blocks), K = 320 projection with LoRA + residual, GEGLU ff1, split-K conv
xa = rnd((2, 64, 64, 320))
wc = rnd((320, 9 * 320), scale=(9 * 320) ** -0.5)
oc = torch.empty(8192, 320, device=dev, dtype=torch.float16)
bias = torch.randn(320, device=dev)
run('gemm conv3x3 8192x320x2880', lambda: ops.gemm(xa, wc, oc, bias=bias, conv=(2, 64, 64, 320)))
run('gemm conv3x3 8192x320x2880 cta_group::2 pair (opt-in mode)', lambda: ops.gemm(xa, wc, oc, bias=bias, conv=(2, 64, 64, 320), pair_mode=1))
a2 = rnd((8192, 320))
w2 = rnd((320, 320), scale=320 ** -0.5)
d16 = torch.zeros(16, 320, device=dev, dtype=torch.float16)
d16[:4] = rnd((4, 320), scale=320 ** -0.5)
up = torch.randn(320, 4, device=dev) * 0.02
res = rnd((8192, 320))
run('gemm 8192x320x320 +lora +residual', lambda: ops.gemm(a2, w2, oc, bias=bias, residual=res, lora_down=d16, lora_up=up, lora_seg=320))
wg = rnd((2560, 320), scale=320 ** -0.5)
og = torch.empty(8192, 1280, device=dev, dtype=torch.float16)
bg = torch.randn(2560, device=dev)
run('gemm geglu 8192x2560x320', lambda: ops.gemm(a2, wg, og, bias=bg, geglu=True))
x16 = rnd((2, 16, 16, 1280))
w16 = rnd((1280, 9 * 1280), scale=(9 * 1280) ** -0.5)
o16 = torch.empty(512, 1280, device=dev, dtype=torch.float16)
b16 = torch.randn(1280, device=dev)
partial = torch.empty(4 * 512 * 1280, device=dev)


def splitk():
    ops.gemm(x16, w16, None, conv=(2, 16, 16, 1280), splits=4, partial=partial)
    ops.splitk_finalize(partial, 4, 512, 1280, o16, bias=b16)


run('gemm conv3x3 split-K 4: 512x1280x11520 + finalize', splitk)

# ---- LoRA gradient (training): res-64 projection, batch 4
xg, dyg = rnd((16384, 320), bf), rnd((16384, 320), bf)
Dn, Up = torch.randn(4, 320, device=dev), torch.randn(320, 4, device=dev)
ws = torch.empty(128 * 4 * 640, device=dev)
gD, gU = torch.empty(4, 320, device=dev), torch.empty(320, 4, device=dev)
run('lora_grad 16384x320x320', lambda: ops.lora_grad(xg, dyg, Dn, Up, 1.0, ws, gD, gU, M=16384, K=320, N=320))

# ---- gradient-fusion closure: Y (fp64) = W (fp32 [320, 768]) G (fp64 [768, 768])
Wf = torch.randn(320, 768, device=dev)
Gf = torch.randn(768, 768, device=dev, dtype=torch.float64)
Yf = torch.empty(320, 768, device=dev, dtype=torch.float64)
run('dgemm_mixed 320x768x768', lambda: ops.dgemm_mixed(Wf, Gf, Yf))

# ---- L-BFGS direction (25 pairs, n = 320 x 768): 51 launches of lbfgs_step_kernel
nv = 320 * 768
Sv = [torch.randn(nv, device=dev) * 0.1 for _ in range(25)]
Yv = [Sv[i] * 1.1 + 0.01 * torch.randn(nv, device=dev) for i in range(25)]
gv = torch.randn(nv, device=dev)
dv = torch.empty_like(gv)
workv = torch.zeros(64, device=dev, dtype=torch.float64)
partv = torch.zeros(260, device=dev)
gtdv = torch.zeros(1, device=dev)
rhov = [1.0 / float((Yv[i] * Sv[i]).sum()) for i in range(25)]
run('lbfgs_direction k=25 n=245760', lambda: ops.lbfgs_direction(Sv, Yv, rhov, gv, 0.9, dv, workv, partv, gtdv))

# ---- VAE attention softmax (4096 x 4096 fp32 logits)
S = torch.randn(4096, 4160, device=dev)
P = torch.empty(4096, 4096, device=dev, dtype=torch.float16)
run('softmax_rows 4096x4096', lambda: ops.softmax_rows(S, P, rows=4096, cols=4096, scale=512 ** -0.5))
print('done')
