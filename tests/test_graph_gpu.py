"""The benchmarked execution mode: the whole training step (multi-stream backbones, forward + loss + backward +
optimizers) captured ONCE into a CUDA graph and replayed must reproduce the eager path step for step."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import O, namespace, noise_for_model  # noqa: E402

pytestmark = pytest.mark.gpu


def _setup(cuda, dtype):
    from adamml_b200.models import build_model
    case = dict(kind="adamml", modality=["rgb", "sound"], N=2, S=2, hw=64, training=True)
    cfg = O.make_cfg(case["modality"], num_segments=2)
    model, _ = build_model(namespace(case, compute_dtype=dtype))
    model.load_state_dict(O.fill_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=0))
    model = model.to(cuda).train()
    xs, y = O.make_inputs(cfg, 2, 2, hw=64)
    noise = noise_for_model(O.draw_noise(1, cfg, 2, 2, True), cuda)   # static noise tensors: deterministic steps
    xs, y = [x.to(cuda) for x in xs], y.to(cuda)
    p_opt = torch.optim.Adam(model.policy_net.parameters(), 1e-3, capturable=True)
    opt = torch.optim.SGD(model.main_net.parameters(), 1e-2, momentum=0.9)

    def step():
        p_opt.zero_grad(set_to_none=True)
        opt.zero_grad(set_to_none=True)
        logits, dec = model(xs, noise=noise)
        loss = F.cross_entropy(logits, y) + (dec.mean(1) ** 2).mean()
        loss.backward()
        p_opt.step()
        opt.step()
        return loss

    return model, step


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_graph_replay_matches_eager(cuda, dtype):
    from adamml_b200.graph import GraphedTrainStep
    torch.manual_seed(0)
    m_e, step_e = _setup(cuda, dtype)
    m_g, step_g = _setup(cuda, dtype)
    # two eager warm-up steps on both (initialises optimizer state), then 3 eager steps vs 3 graph replays
    for _ in range(2):
        step_e()
        step_g()
    graphed = GraphedTrainStep(step_g).capture()
    assert graphed.launches > 1000
    # Each comparison starts from IDENTICAL parameters / BN buffers (copied in place: the graph keeps its addresses),
    # so the loss of the replay must equal the eager loss up to atomic-order noise; trajectories are not compared
    # (a 2-clip batch is chaotic: the two copies drift apart within a few steps even in eager mode).
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    for i in range(3):
        with torch.no_grad():
            for (k, a), (_, b) in zip(m_e.state_dict().items(), m_g.state_dict().items()):
                b.copy_(a)
        le = step_e().item()
        lg = graphed().item()
        print(f"{dtype} step {i}: eager loss {le:.6f}, graph replay {lg:.6f}")
        assert abs(le - lg) <= tol * max(1.0, abs(le)), (i, le, lg)
    torch.cuda.synchronize()
    for (k, a), (_, b) in zip(m_e.state_dict().items(), m_g.state_dict().items()):
        if not a.is_floating_point():
            assert torch.equal(a, b), k      # num_batches_tracked advanced identically
        else:
            assert torch.isfinite(b).all(), k
