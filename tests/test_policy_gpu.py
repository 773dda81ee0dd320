"""Policy head + fusion kernels against the oracle equations (fp32, bit-exact selections)."""
import pytest
import torch

from util import O, rel

pytestmark = pytest.mark.gpu


def test_gate_fuse_fwd_bwd(cuda):
    from adamml_b200.models.joint_resnet_mobilenetv2 import _GateFuse
    g = torch.Generator().manual_seed(1)
    for M, lf_on in ((2, True), (3, True), (2, False)):
        S, N, C = 3, 4, 31
        logits = torch.randn(M, S, N, C, generator=g, requires_grad=True)
        dec = (torch.rand(S, M, N, generator=g) > 0.4).float().requires_grad_(True)
        lf = (torch.rand(M - 1, generator=g) * 0.5).requires_grad_(True) if lf_on else None
        # oracle: per segment main_forward fusion then mean over segments
        outs = []
        for s in range(S):
            t = torch.stack([logits[m, s] * dec[s, m].view(N, 1) for m in range(M)])
            if lf is not None:
                w = torch.cat((lf, torch.ones(1) - lf.sum(0, keepdim=True)))
                outs.append((t * w.view(-1, 1, 1)).sum(0))
            else:
                outs.append(t.mean(0))
        ref = torch.stack(outs, 1).mean(1)
        go = torch.randn(N, C, generator=g)
        ref.backward(go)
        l2 = logits.detach().to(cuda).requires_grad_(True)
        d2 = dec.detach().to(cuda).requires_grad_(True)
        f2 = lf.detach().to(cuda).requires_grad_(True) if lf is not None else None
        out = _GateFuse.apply(l2, d2, f2)
        out.backward(go.to(cuda))
        assert rel(out, ref) < 1e-6
        assert rel(l2.grad, logits.grad) < 1e-6
        assert rel(d2.grad, dec.grad) < 1e-5
        if lf is not None:
            assert rel(f2.grad, lf.grad) < 1e-5


@pytest.mark.parametrize("M", [2, 3])
def test_policy_head_matches_oracle(cuda, M):
    from adamml_b200.models.policy_net import PolicyNet, _PolicyHead
    S, N = 4, 5
    g = torch.Generator().manual_seed(2)

    class _J(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.last_channels = 2048
            self.nets = torch.nn.ModuleList()
            self.joint = torch.nn.Sequential(torch.nn.Linear(1280 * M, 2048), torch.nn.ReLU(True),
                                             torch.nn.Linear(2048, 2048), torch.nn.ReLU(True))
    pn = PolicyNet(_J(), ["m%d" % i for i in range(M)])
    shapes = {"policy_net." + k: v.shape for k, v in pn.state_dict().items()}
    sd0 = O.fill_state_dict(shapes, seed=3)
    pn.load_state_dict({k[len("policy_net."):]: v for k, v in sd0.items()})
    pn = pn.to(cuda)
    feats = [torch.randn(S * N, 1280, generator=g) for _ in range(M)]
    expo = [torch.empty(M * N, 2).exponential_(generator=g) for _ in range(S)]
    d_dec = torch.randn(S, M, N, generator=g)

    # ---- oracle: same equations as O.policy_forward after the backbones ----
    sd = O.clone_sd(sd0)
    fo = [f.clone().requires_grad_(True) for f in feats]
    pre = "policy_net."
    x = torch.cat(fo, 1)
    x = torch.relu(torch.nn.functional.linear(x, sd[pre + "joint_net.joint.0.weight"], sd[pre + "joint_net.joint.0.bias"]))
    x = torch.relu(torch.nn.functional.linear(x, sd[pre + "joint_net.joint.2.weight"], sd[pre + "joint_net.joint.2.bias"]))
    outs = x.view(S, N, -1)
    decs = []
    h = torch.zeros(N, 256); c = torch.zeros(N, 256); logits = None
    for s in range(S):
        fb = torch.zeros(N, 2 * M) if s == 0 else logits.view(M, -1, 2).permute(1, 0, 2).contiguous().view(-1, 2 * M)
        h, c = O.lstm_cell(torch.cat((outs[s], fb), -1), h, c, sd[pre + "lstm.weight_ih"], sd[pre + "lstm.weight_hh"],
                           sd[pre + "lstm.bias_ih"], sd[pre + "lstm.bias_hh"])
        logits = torch.cat([torch.nn.functional.linear(h, sd[f"{pre}fcs.{m}.weight"], sd[f"{pre}fcs.{m}.bias"])
                            for m in range(M)])
        decs.append(O.gumbel_hard(logits, expo[s], 5.0))
    o_dec = torch.stack(decs).view(S, M, N)
    (o_dec * d_dec).sum().backward()

    # ---- product ----
    fp = [f.clone().to(cuda).requires_grad_(True) for f in feats]
    j = pn.joint_net.joint
    params = [j[0].weight, j[0].bias, j[2].weight, j[2].bias, pn.lstm.weight_ih, pn.lstm.weight_hh, pn.lstm.bias_ih,
              pn.lstm.bias_hh]
    for fc in pn.fcs:
        params += [fc.weight, fc.bias]
    dec, lg = _PolicyHead.apply(pn, torch.stack(expo).to(cuda), 5.0, True, M, *fp, *params)
    (dec * d_dec.to(cuda)).sum().backward()
    assert torch.equal(dec.detach().cpu(), o_dec.detach()), "selections must be bit-exact"
    assert set(dec.detach().unique().tolist()) <= {0.0, 1.0}
    for a, b in zip(fp, fo):
        assert rel(a.grad, b.grad) < 1e-3
    for k, p in pn.named_parameters():
        assert rel(p.grad, sd[pre + k].grad) < 1e-3, k
