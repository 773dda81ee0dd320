"""The reference's own training-step body driving this package through its DDP entry point (§8 a14).

`train_adamml.py:111-134` wraps the model as SyncBatchNorm.convert_sync_batchnorm + DistributedDataParallel(
find_unused_parameters=True) and `utils/utils.py:349-400` runs, per iteration: model(images) -> compute_policy_loss ->
CE -> metric all-reduces -> loss.item() -> backward -> the two optimizers gated by model.module.update_*_net.  The
loop below restates exactly that body (eager, one process, world size 1 over NCCL) for the three training stages the
script walks through (warm-up: policy frozen; alternating: main epoch / policy epoch), and checks it against the same
steps taken on the bare module: the wrapper must not change a single gradient, and frozen halves must stay frozen."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.nn.functional as F

from util import O, compare_grads, namespace, noise_for_model

pytestmark = pytest.mark.gpu


def accuracy(output, target, topk=(1, 5)):
    """utils/utils.py:42-56"""
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.view(1, -1).expand_as(pred.t()))
    return [correct[:k].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / target.size(0)) for k in topk]


def compute_policy_loss(selection, cost_weights, gammas, cls_logits, cls_targets):
    """utils/utils.py:166-184, penalty_type='blockdrop'"""
    num_modality = selection.shape[-1]
    policy_loss = torch.tensor(0.0, device=selection.device)
    _, pred = cls_logits.detach().max(1)
    correct = (pred == cls_targets).type_as(cls_logits)
    selection = torch.mean(selection, dim=1)
    selection = selection ** 2
    for w, pl in zip(cost_weights, selection.chunk(num_modality, dim=-1)):
        policy_loss = policy_loss + w * torch.mean(correct * pl)
    return policy_loss + torch.mean((torch.ones_like(correct) - correct) * gammas)


def reference_step(model, images, target, optimizer, p_optimizer, cost_weights, gammas, noise):
    """the body of the loop at utils/utils.py:349-400 (model is the DDP wrapper, `noise` pins the RNG draws)"""
    output, selection = model(images, noise=noise)
    policy_loss = compute_policy_loss(selection, cost_weights, gammas, output, target)
    selection_ratio = selection.detach().mean(0).mean(0)
    cls_loss = F.cross_entropy(output, target)
    prec1, prec5 = accuracy(output, target)
    if dist.is_initialized():
        world_size = dist.get_world_size()
        dist.all_reduce(prec1)
        dist.all_reduce(prec5)
        prec1 /= world_size
        prec5 /= world_size
        dist.all_reduce(selection_ratio)
        selection_ratio /= world_size
    loss = cls_loss
    if model.module.update_policy_net:
        loss = loss + policy_loss
    loss_value = loss.item()
    loss.backward()
    grads = {k: p.grad.clone() for k, p in model.module.named_parameters() if p.grad is not None}
    if model.module.update_policy_net:
        p_optimizer.step()
        p_optimizer.zero_grad()
    if model.module.update_main_net:
        optimizer.step()
        optimizer.zero_grad()
    return loss_value, grads, selection_ratio


def bare_step(model, images, target, noise):
    output, selection = model(images, noise=noise)
    loss = F.cross_entropy(output, target)
    if model.update_policy_net:
        loss = loss + compute_policy_loss(selection, torch.tensor([1.0, 1.0], device=target.device),
                                          torch.tensor(10.0, device=target.device), output, target)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    model.zero_grad(set_to_none=True)
    return loss.item(), grads


def test_reference_train_loop_under_ddp(cuda):
    from adamml_b200.models import build_model
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(29700 + os.getpid() % 200))
    own_pg = not dist.is_initialized()
    if own_pg:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=cuda)
    try:
        case = dict(kind="adamml", modality=["rgb", "sound"], N=2, S=2, hw=64, training=True)
        cfg = O.make_cfg(case["modality"], num_segments=2)
        xs, y = O.make_inputs(cfg, 2, 2, hw=64)
        xs, y = [x.to(cuda) for x in xs], y.to(cuda)

        def make():
            model, _ = build_model(namespace(case))  # default precision mode
            model.load_state_dict(O.fill_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=0))
            return model.to(cuda).train()

        # train_adamml.py:111-134
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(make())
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[cuda.index], find_unused_parameters=True)
        # train_adamml.py:250-257
        p_optimizer = torch.optim.Adam(ddp.module.policy_net.parameters(), 0.001, weight_decay=1e-4)
        optimizer = torch.optim.SGD(ddp.module.main_net.parameters(), 0.01, momentum=0.9, weight_decay=1e-4)
        cost_weights = torch.tensor([1.0, 1.0], device=cuda)
        gammas = torch.tensor(10.0, device=cuda)
        bare = make()
        stages = [("warmup", "freeze_policy_net"), ("main_epoch", None), ("policy_epoch", "freeze_main_net")]
        for it, (stage, freeze) in enumerate(stages):
            for m in (ddp.module, bare):
                m.unfreeze_policy_net()
                m.unfreeze_main_net()
                if freeze:
                    getattr(m, freeze)()
            # identical parameters / buffers on both sides before the step
            with torch.no_grad():
                for (k, a), (_, b) in zip(ddp.module.state_dict().items(), bare.state_dict().items()):
                    b.copy_(a)
            noise = noise_for_model(O.draw_noise(10 + it, cfg, 2, 2, True), cuda)
            loss_d, g_d, ratio = reference_step(ddp, xs, y, optimizer, p_optimizer, cost_weights, gammas, noise)
            loss_b, g_b = bare_step(bare, xs, y, noise)
            assert abs(loss_d - loss_b) <= 1e-4 * max(1.0, abs(loss_b)), (stage, loss_d, loss_b)
            assert set(g_d) == set(g_b), stage
            frozen = {"freeze_policy_net": "policy_net.", "freeze_main_net": "main_net."}.get(freeze)
            for k in g_d:
                assert not (frozen and k.startswith(frozen)), (stage, k)
            bad = compare_grads(g_d, g_b, tol=2e-2)  # (robust to mathematically-zero gradients, see util.py)
            assert not bad, (stage, bad[:5])
            assert ratio.shape == (2,) and bool(((ratio >= 0) & (ratio <= 1)).all())
            for p in ddp.module.parameters():
                assert torch.isfinite(p).all()
            print(f"{stage}: loss {loss_d:.5f} (bare module {loss_b:.5f}), {len(g_d)} gradient tensors, "
                  f"selection ratio {ratio.tolist()}")
        # the stage epilogue of the alternating schedule (train_adamml.py:516)
        t0 = ddp.module.policy_net.temperature
        ddp.module.decay_temperature()
        assert abs(ddp.module.policy_net.temperature - 0.965 * t0) < 1e-9
    finally:
        if own_pg:
            dist.destroy_process_group()
