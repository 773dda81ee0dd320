"""Fused train-step tail: loss kernel and multi-tensor optimizers against the torch code the
reference's step body runs (utils/utils.py:166-184,362-400; train_adamml.py:250-257)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def ref_policy_loss(selection, cost_weights, gammas, logits, targets):
    """utils/utils.py:166-184 'blockdrop' verbatim semantics (incl. the [N] x [N,1] broadcast)"""
    M = selection.shape[-1]
    loss = torch.tensor(0.0, device=selection.device)
    correct = (torch.argmax(logits.detach(), dim=-1) == targets).type_as(logits)
    sel = torch.mean(selection, dim=1)
    sel = sel * sel
    for w, pl in zip(cost_weights, sel.chunk(M, dim=-1)):
        loss = loss + w * torch.mean(correct * pl)
    return loss + torch.mean((torch.ones_like(correct) - correct) * gammas)


@pytest.mark.parametrize("use_policy", [True, False])
@pytest.mark.parametrize("N,C,S,M", [(72, 31, 5, 2), (7, 31, 10, 3), (300, 5, 2, 1)])
def test_loss_tail_matches_torch(cuda, N, C, S, M, use_policy):
    from adamml_b200.optim import loss_tail
    g = torch.Generator().manual_seed(N + C)
    logits = torch.randn(N, C, generator=g).to(cuda)
    # make some predictions correct so that both branches of the policy term are exercised
    target = torch.where(torch.rand(N, generator=g) < 0.4, logits.cpu().argmax(-1), torch.randint(0, C, (N,), generator=g))
    target = target.to(cuda)
    sel = (torch.rand(N, S, M, generator=g) > 0.5).float().to(cuda)
    cw = (torch.rand(M, generator=g) + 0.5).to(cuda)
    l1, s1 = logits.clone().requires_grad_(True), sel.clone().requires_grad_(True)
    ref = F.cross_entropy(l1, target)
    if use_policy:
        ref = ref + ref_policy_loss(s1, cw, 10.0, l1, target)
    (ref * 1.7).backward()
    l2, s2 = logits.clone().requires_grad_(True), sel.clone().requires_grad_(True)
    out = loss_tail(l2, target, s2, cw, 10.0, use_policy)
    (out * 1.7).backward()
    assert abs(out.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
    assert torch.allclose(l2.grad, l1.grad, rtol=1e-5, atol=1e-7)
    if use_policy:
        assert torch.allclose(s2.grad, s1.grad, rtol=1e-5, atol=1e-8)
    else:
        assert s2.grad.abs().max() == 0


def _params(cuda, seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(64, 3, 7, 7), (64,), (64,), (256, 64, 1, 1), (31, 2048), (31,), (1,), (20000,), (3, 5)]
    return [torch.nn.Parameter(torch.randn(s, generator=g).to(cuda)) for s in shapes]


@pytest.mark.parametrize("kind", ["sgd", "adam"])
def test_fused_optimizers_match_torch(cuda, kind):
    """5 steps with fresh gradients (re-allocated each step, like zero_grad(set_to_none=True)), one parameter frozen
    midway (train_adamml.py freeze phases): every tensor equals torch.optim's to fp32 round-off."""
    from adamml_b200.optim import FusedAdam, FusedSGD
    pa, pb = _params(cuda, 1), _params(cuda, 1)
    if kind == "sgd":   # train_adamml.py:254-257
        oa = torch.optim.SGD(pa, 0.01, momentum=0.9, weight_decay=1e-4)
        ob = FusedSGD(pb, 0.01, momentum=0.9, weight_decay=1e-4)
    else:               # train_adamml.py:250-253
        oa = torch.optim.Adam(pa, 0.001, weight_decay=1e-4)
        ob = FusedAdam(pb, 0.001, weight_decay=1e-4)
    g = torch.Generator().manual_seed(7)
    for it in range(5):
        for i, (a, b) in enumerate(zip(pa, pb)):
            if it >= 3 and i == 3:          # frozen: no gradient -> skipped by both
                a.grad = b.grad = None
                continue
            gr = torch.randn(a.shape, generator=g).to(cuda)
            a.grad, b.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
        if it == 2:                          # an LR-scheduler step
            for o in (oa, ob):
                o.param_groups[0]["lr"] *= 0.5
    for i, (a, b) in enumerate(zip(pa, pb)):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (kind, i, (a - b).abs().max().item())


def test_fused_tail_in_cuda_graph(cuda):
    """the whole tail (loss kernel + both fused optimizers) captured into a CUDA graph replays like eager torch"""
    from adamml_b200.optim import FusedAdam, FusedSGD, loss_tail
    torch.manual_seed(0)
    N, C, S, M = 16, 31, 5, 2

    def make():
        g = torch.Generator().manual_seed(3)
        w = torch.nn.Parameter(torch.randn(C, 40, generator=g).to(cuda) * 0.1)
        v = torch.nn.Parameter(torch.randn(M, 40, generator=g).to(cuda) * 0.1)
        return w, v
    x = torch.randn(N, 40, device=cuda)
    y = torch.randint(0, C, (N,), device=cuda)
    cw = torch.ones(M, device=cuda)

    def fwd(w, v, fused):
        logits = x @ w.t()
        sel = torch.sigmoid((x @ v.t())).unsqueeze(1).expand(N, S, M).contiguous()
        if fused:
            return loss_tail(logits, y, sel, cw, 10.0, True)
        return F.cross_entropy(logits, y) + ref_policy_loss(sel, cw, 10.0, logits, y)

    wa, va = make()
    wb, vb = make()
    oa = [torch.optim.SGD([wa], 0.05, momentum=0.9, weight_decay=1e-4), torch.optim.Adam([va], 0.01, weight_decay=1e-4)]
    ob = [FusedSGD([wb], 0.05, momentum=0.9, weight_decay=1e-4), FusedAdam([vb], 0.01, weight_decay=1e-4)]

    def step(w, v, opts, fused):
        for o in opts:
            o.zero_grad(set_to_none=True)
        loss = fwd(w, v, fused)
        loss.backward()
        for o in opts:
            o.step()
        return loss

    for _ in range(2):                      # warm-up: optimizer state exists before capture
        step(wa, va, oa, False)
        step(wb, vb, ob, True)
    for o in ob:
        o.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        gl = step(wb, vb, ob, True)
    # the capture pass itself does not execute: run 3 replays against 3 eager torch steps
    for i in range(3):
        le = step(wa, va, oa, False).item()
        graph.replay()
        torch.cuda.synchronize()
        assert abs(gl.item() - le) < 1e-5 * max(1.0, abs(le)), (i, gl.item(), le)
    assert torch.allclose(wa, wb, rtol=1e-5, atol=1e-6) and torch.allclose(va, vb, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("max_norm", [0.5, 5.0, 1e6])
def test_fused_clip_grad_norm_matches_torch(cuda, max_norm):
    """clip_grad_norm_ (utils/utils.py:390-391, --clip_gradient) as two multi-tensor launches: same total norm, same
    clipped gradients as torch.nn.utils.clip_grad_norm_ (clipping active, marginal, inactive); parameters without a
    gradient are skipped; the second call goes through the cached tables."""
    from adamml_b200.optim import clip_grad_norm_
    pa, pb = _params(cuda, 2), _params(cuda, 2)
    g = torch.Generator().manual_seed(11)
    for rep in range(2):
        for i, (a, b) in enumerate(zip(pa, pb)):
            if i == 2:
                a.grad = b.grad = None
                continue
            gr = torch.randn(a.shape, generator=g).to(cuda) * (0.3 + rep)
            if a.grad is None:
                a.grad, b.grad = gr.clone(), gr.clone()
            else:  # same storage as in the first round: exercises the table cache
                a.grad.copy_(gr)
                b.grad.copy_(gr)
        want = torch.nn.utils.clip_grad_norm_(pa, max_norm)
        got = clip_grad_norm_(pb, max_norm)
        assert got.shape == want.shape and got.device == want.device
        assert abs(got.item() - want.item()) <= 2e-6 * want.item()
        for i, (a, b) in enumerate(zip(pa, pb)):
            if a.grad is None:
                assert b.grad is None
            else:
                assert torch.allclose(a.grad, b.grad, rtol=3e-6, atol=1e-9), (i, (a.grad - b.grad).abs().max().item())


def test_fused_clip_grad_norm_in_cuda_graph(cuda):
    """the clipping pair captures (device-side coefficient, pinned address table) and replays on fresh gradient values"""
    from adamml_b200.optim import clip_grad_norm_
    ps = _params(cuda, 4)
    for p in ps:
        p.grad = torch.randn_like(p)
    clip_grad_norm_(ps, 1.0)                 # eager warm-up: chunk geometry exists before the capture
    for p in ps:                             # new gradient storage, as zero_grad(set_to_none=True) + backward gives
        p.grad = torch.randn_like(p)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        total = clip_grad_norm_(ps, 1.0)
    for rep in range(2):
        fresh = [torch.randn_like(p) * (rep + 1) for p in ps]
        for p, f in zip(ps, fresh):
            p.grad.copy_(f)
        graph.replay()
        ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
        for r, f in zip(ref, fresh):
            r.grad = f.clone()
        want = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        assert abs(total.item() - want.item()) <= 2e-6 * want.item()
        for r, p in zip(ref, ps):
            assert torch.allclose(r.grad, p.grad, rtol=3e-6, atol=1e-9)
