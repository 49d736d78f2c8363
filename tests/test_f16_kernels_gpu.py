"""fp16-activation instantiations of the sampling kernels (the engine's default activation type, mos_b200/engine.py) vs
plain PyTorch fp32 references of the same op, plus the BASELINE config-4 attention size (18432 x 18432 keys, d = 40).

Tolerances: inputs are exact in both paths and one final fp16 rounding remains (eps 4.9e-4): rel-L2 <= 6e-4; attention
additionally rounds P to fp16 before the PV product: rel-L2 <= 1e-3.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
H16 = torch.float16


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def mk(shape, dev, scale=1.0, seed=0, dtype=H16):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev).to(dtype)


def pack_heads(q, k, v, d):
    B, H, nq, _ = q.shape
    nk = k.shape[2]
    dp, dv, nk8 = (d + 63) // 64 * 64, (d + 15) // 16 * 16, (nk + 7) // 8 * 8
    Q = torch.zeros(B * H, nq, dp, device=q.device, dtype=q.dtype)
    K = torch.zeros(B * H, nk, dp, device=q.device, dtype=q.dtype)
    Vt = torch.zeros(B * H, dv, nk8, device=q.device, dtype=q.dtype)
    Q[..., :d] = q.reshape(B * H, nq, d)
    K[..., :d] = k.reshape(B * H, nk, d)
    Vt[:, :d, :nk] = v.reshape(B * H, nk, d).transpose(1, 2)
    return Q, K, Vt


@pytest.mark.parametrize('d,nq,nk', [(40, 4096, 4096), (40, 1000, 300), (80, 1024, 1024), (160, 256, 256),
                                     (40, 4096, 77), (80, 1024, 77), (160, 64, 77)])
def test_attention_f16(cuda, d, nq, nk):
    from mos_b200 import ops
    B, H = 2, 8
    q, k, v = mk((B, H, nq, d), cuda, seed=1), mk((B, H, nk, d), cuda, seed=2), mk((B, H, nk, d), cuda, seed=3)
    Q, K, Vt = pack_heads(q, k, v, d)
    out = torch.full((B, nq, H * d), float('nan'), device=cuda, dtype=H16)
    probs = torch.empty(B * H, nq, nk, device=cuda) if nk == 77 else None
    ops.attention(Q, K, Vt, out, batch=B, heads=H, head_dim=d, nq=nq, nk=nk, probs=probs)
    torch.cuda.synchronize()
    ref = F.scaled_dot_product_attention(q.float(), k.float(), v.float()).permute(0, 2, 1, 3).reshape(B, nq, H * d)
    assert rel_l2(out, ref) < 1e-3
    if probs is not None:
        pref = ((q.float() @ k.float().transpose(-1, -2)) * d ** -0.5).softmax(-1).reshape(B * H, nq, nk)
        assert rel_l2(probs, pref) < 1e-4


@pytest.mark.parametrize('dtype', [H16, torch.bfloat16])
def test_attention_config4_size(cuda, dtype):
    """BASELINE config 4 (768 x 1536 regional sampling): res-64 self-attention is 18432 queries x 18432 keys, d = 40, CFG
    batch 2 x 8 heads.  Reference: fp32 softmax(QK^T) V per head (1.4 GB of scores per head).  Tolerances: fp16 1e-3, bf16
    8e-3 (P is rounded to the activation type before the PV product)."""
    from mos_b200 import ops
    B, H, d, n = 2, 8, 40, 18432
    q, k, v = (mk((B, H, n, d), cuda, seed=s, dtype=dtype) for s in (1, 2, 3))
    Q, K, Vt = pack_heads(q, k, v, d)
    out = torch.full((B, n, H * d), float('nan'), device=cuda, dtype=dtype)
    ops.attention(Q, K, Vt, out, batch=B, heads=H, head_dim=d, nq=n, nk=n)
    torch.cuda.synchronize()
    num = den = 0.0
    for b in range(B):
        for h in range(H):
            p = ((q[b, h].float() @ k[b, h].float().t()) * d ** -0.5).softmax(-1)
            ref = p @ v[b, h].float()
            got = out[b, :, h * d:(h + 1) * d].float()
            num += (got - ref).pow(2).sum().item()
            den += ref.pow(2).sum().item()
            del p
    e = (num / den) ** 0.5
    print(f'attention 18432^2 d=40 {dtype}: rel-L2 {e:.2e}')
    assert e < (1e-3 if dtype == H16 else 8e-3)


@pytest.mark.parametrize('B,HW,C,ld,silu', [(2, 4096, 320, 320, True), (2, 1024, 640, 1280, True),
                                             (2, 64, 1280, 1280, False), (1, 4096, 960, 960, True)])
def test_groupnorm_f16(cuda, B, HW, C, ld, silu):
    from mos_b200 import ops
    buf = mk((B, HW, ld), cuda, seed=1) * 1.5 + 0.3
    gamma, beta = torch.randn(C, device=cuda), torch.randn(C, device=cuda)
    y = torch.empty((B, HW, C), device=cuda, dtype=H16)
    partial = torch.zeros(B * 592 * 64, device=cuda)
    ops.groupnorm(buf, gamma, beta, y, partial, B=B, HW=HW, C=C, eps=1e-5, silu=silu, ldx=ld)
    ref = F.group_norm(buf[..., :C].float().transpose(1, 2), 32, gamma, beta, 1e-5)
    if silu:
        ref = F.silu(ref)
    assert rel_l2(y, ref.transpose(1, 2)) < 6e-4


def test_layernorm_conv_edge_add_region_f16(cuda):
    from mos_b200 import ops
    x = mk((2048, 640), cuda, seed=1) * 2 + 0.5
    gamma, beta = torch.randn(640, device=cuda), torch.randn(640, device=cuda)
    y = torch.empty_like(x)
    ops.layernorm(x, gamma, beta, y, M=2048, C=640)
    assert rel_l2(y, F.layer_norm(x.float(), (640,), gamma, beta, 1e-5)) < 6e-4
    B, H, W = 2, 32, 32
    xi = torch.randn(B, 4, H, W, device=cuda)
    w = torch.randn(320, 4, 3, 3, device=cuda) * 0.2
    b = torch.randn(320, device=cuda)
    yi = torch.empty(B, H, W, 320, device=cuda, dtype=H16)
    ops.conv_in(xi, w.permute(2, 3, 1, 0).reshape(36, 320).contiguous(), b, yi)
    assert rel_l2(yi, F.conv2d(xi, w, b, padding=1).permute(0, 2, 3, 1)) < 6e-4
    xo = mk((B, H, W, 320), cuda, seed=4)
    wo = torch.randn(4, 320, 3, 3, device=cuda) * 0.05
    bo = torch.randn(4, device=cuda)
    yo = torch.empty(B, 4, H, W, device=cuda)
    ops.conv_out(xo, wo.permute(0, 2, 3, 1).reshape(4, 9, 320).contiguous(), bo, yo, B=B, H=H, W=W, C=320)
    assert rel_l2(yo, F.conv2d(xo.float().permute(0, 3, 1, 2), wo, bo, padding=1)) < 1e-5
    a, r = mk((100, 640), cuda, seed=5), mk((100, 320), cuda, seed=6)
    a0 = a.clone()
    ops.add_rows(a, r, M=100, C=320, ldx=640, ldr=320)
    assert rel_l2(a[:, :320], a0[:, :320].float() + r.float()) < 6e-4 and torch.equal(a[:, 320:], a0[:, 320:])
    FH, FW, C = 12, 24, 320
    glob = mk((B, FH * FW, C), cuda, seed=1)
    regs = [mk((B, FH * FW, C), cuda, seed=2 + i) for i in range(2)]
    boxes = [(0, 1, 12, 9), (1, 7, 12, 16)]
    ptrs = torch.tensor([t.data_ptr() for t in regs], dtype=torch.int64, device=cuda)
    out = torch.empty_like(glob)
    ops.region_combine(glob, ptrs, boxes, out, B=B, FH=FH, FW=FW, C=C, ld=C)
    count = torch.zeros(FH, FW, device=cuda)
    acc = torch.zeros(B, FH, FW, C, device=cuda)
    for t, (sh, sw, eh, ew) in zip(regs, boxes):
        count[sh:eh, sw:ew] += 1
        acc[:, sh:eh, sw:ew] += t.float().view(B, FH, FW, C)[:, sh:eh, sw:ew]
    ref = torch.where(count[None, :, :, None] == 0, glob.float().view(B, FH, FW, C), acc / count.clamp_min(1)[None, :, :, None])
    assert rel_l2(out.view(B, FH, FW, C), ref) < 6e-4
