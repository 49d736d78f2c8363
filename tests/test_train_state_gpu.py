"""Fused flat AdamW (mos_flat_adamw_step) vs torch.optim.AdamW with the reference's three param groups
(train_edlora.py:57: lr 1e-3 / 1e-5 / 1e-4, weight decay 0.01, betas 0.9 / 0.999) over several steps, including the
1/world gradient scaling and Norm_mean of the concept rows (train_edlora.py:138-140).  fp32: rel-L2 <= 1e-6."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_flat_adamw_matches_torch(cuda):
    from mos_b200.dp import FlatTrainState, optimizer_step
    st = FlatTrainState(32, 768, 3000, 5000, device=cuda)
    torch.manual_seed(0)
    st.params.copy_(torch.randn(st.n, device=cuda) * 0.1)
    e0, e1 = st.group_end[0], st.group_end[1]
    ref = [st.params[:e0].clone().requires_grad_(), st.params[e0:e1].clone().requires_grad_(),
           st.params[e1:].clone().requires_grad_()]
    opt = torch.optim.AdamW([{'params': [ref[0]], 'lr': 1e-3}, {'params': [ref[1]], 'lr': 1e-5},
                             {'params': [ref[2]], 'lr': 1e-4}], weight_decay=0.01, betas=(0.9, 0.999))
    norm = torch.zeros(1, device=cuda)
    for step in range(5):
        g = torch.randn(st.n, device=cuda)
        st.grads[:st.n] = g * 2.0                      # "sum over 2 ranks"
        for r, (a, b) in zip(ref, [(0, e0), (e0, e1), (e1, st.n)]):
            r.grad = g[a:b].clone()
        opt.step()
        optimizer_step(st, grad_scale=0.5, norm_out=norm)
    torch.cuda.synchronize()
    want = torch.cat([r.detach() for r in ref])
    assert ((st.params - want).norm() / want.norm()).item() < 1e-6
    rows = want[:32 * 768].view(32, 768)
    assert abs(norm.item() - rows.norm(dim=1).mean().item()) < 1e-5
