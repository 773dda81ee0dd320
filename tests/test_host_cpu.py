"""Host-side logic that needs no GPU: clip selection for the inference skip path, the gating of that path, and the
uint8-normalisation constants handed to the data-layer kernels."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import namespace  # noqa: E402


def test_select_clips_keeps_frames_of_a_clip_together():
    from adamml_b200 import ops
    clips, T = 6, 4
    x = torch.arange(clips * T * 2 * 3 * 5, dtype=torch.float32).view(clips * T, 2, 3, 5)
    idx = torch.tensor([1, 4, 5])
    sel = ops.select_clips(x, idx, clips)
    assert sel.shape == (3 * T, 2, 3, 5)
    for j, c in enumerate(idx.tolist()):
        assert torch.equal(sel[j * T:(j + 1) * T], x[c * T:(c + 1) * T])
    s2d = ops.S2D(x, 1, 4, 6, 7)
    sel2 = ops.select_clips(s2d, idx, clips)
    assert isinstance(sel2, ops.S2D) and torch.equal(sel2.t, sel) and (sel2.C, sel2.H, sel2.W, sel2.R) == (1, 4, 6, 7)
    with pytest.raises(ValueError):
        ops.select_clips(x, idx, 5)


@pytest.fixture(scope="module")
def model():
    from adamml_b200.models import build_model
    case = dict(kind="adamml", modality=["rgb", "sound"], S=2)
    torch.manual_seed(0)
    m, _ = build_model(namespace(case, compute_dtype=torch.bfloat16))
    return m


def test_skip_path_is_confined_to_eval_without_tape(model):
    model.eval()
    assert model.skip_unselected
    assert not model._can_skip()                       # grad mode on: the straight-through gradient needs every pass
    with torch.no_grad():
        assert model._can_skip()
        model.skip_unselected = False
        assert not model._can_skip()
        model.skip_unselected = True
        model.main_net.nets[1].train()                 # batch-statistics BN couples the clips of a batch
        assert not model._can_skip()
        model.eval()
        assert model._can_skip()
    model.train()
    with torch.no_grad():
        assert not model._can_skip()


def test_u8_normalisation_constants_follow_group_normalize(model):
    """GroupNormalize repeats the per-modality mean/std over the channel planes of a frame
    (utils/video_transforms.py:77-78); values from AdaMML.mean()/std() (adamml.py:93-99)."""
    mean, std = model._input_norm("rgb", 3, torch.device("cpu"))
    assert torch.allclose(mean, torch.tensor([0.485, 0.456, 0.406])) and torch.allclose(std, torch.tensor([0.229, 0.224, 0.225]))
    mean, std = model._input_norm("flow", 10, torch.device("cpu"))
    assert mean.shape == (10,) and torch.all(mean == 0.5) and torch.allclose(std, torch.full((10,), 0.226))
    mean, std = model._input_norm("rgbdiff", 15, torch.device("cpu"))
    assert torch.allclose(mean, torch.tensor([0.485, 0.456, 0.406] * 5))
    with pytest.raises(ValueError):
        model._input_norm("rgb", 4, torch.device("cpu"))


def test_u8_frames_without_norm_are_rejected():
    from adamml_b200 import ops
    x = torch.zeros(1, 3, 4, 4, dtype=torch.uint8)
    with pytest.raises(ValueError):
        ops._u8_norm(x, 3, None)


def test_reference_checkpoint_with_module_prefix_loads(tmp_path, model):
    """Resume / fine-tune compatibility (SURVEY.md §5 checkpoint row): train_adamml.py saves the DDP-wrapped
    state_dict, i.e. every key carries a `module.` prefix (train_adamml.py:373-383), together with the temperature;
    the unimodal loader strips that prefix and loads strictly (joint_resnet_mobilenetv2.py:141-155).  A checkpoint
    file written that way must load into the product model key for key."""
    from adamml_b200.models import build_model
    sd = {"module." + k: v.clone() for k, v in model.state_dict().items()}
    path = tmp_path / "checkpoint.pth.tar"
    torch.save({"epoch": 3, "arch": "adamml", "state_dict": sd, "temperature": 4.2, "stage": "alternative_training"},
               path)
    ckpt = torch.load(path, map_location="cpu")
    # (a) the reference's resume path: DataParallel / DDP wrapper around the model, load_state_dict(strict)
    wrapped = torch.nn.DataParallel(model) if False else None  # DataParallel needs a GPU; emulate the wrapper keys
    m2, _ = build_model(namespace(dict(kind="adamml", modality=["rgb", "sound"], S=2), compute_dtype=torch.bfloat16))
    holder = torch.nn.Module()
    holder.module = m2
    missing, unexpected = holder.load_state_dict(ckpt["state_dict"], strict=True)
    assert not missing and not unexpected
    m2.policy_net.set_temperature(ckpt["temperature"])
    assert m2.policy_net.temperature == 4.2
    for (k, a), (_, b) in zip(model.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    # (b) the unimodal loader of the main net (joint_resnet_mobilenetv2.py:146-155): per-backbone checkpoint files
    paths = []
    for i, net in enumerate(model.main_net.nets):
        p = tmp_path / f"uni{i}.pth.tar"
        torch.save({"state_dict": {"module." + k: v for k, v in net.state_dict().items()}}, p)
        paths.append(str(p))
    m3, _ = build_model(namespace(dict(kind="adamml", modality=["rgb", "sound"], S=2), compute_dtype=torch.bfloat16,
                                  unimodality_pretrained=paths))
    for i, net in enumerate(model.main_net.nets):
        for (k, a), (_, b) in zip(net.state_dict().items(), m3.main_net.nets[i].state_dict().items()):
            assert torch.equal(a, b), (i, k)


def test_graph_step_guard_sees_host_state_changes(model):
    """GraphedTrainStep's guard (adamml_b200/graph.py): temperature decay, freeze / unfreeze, train / eval and a
    Python-float learning rate are baked into captured launches, so each of them must change the snapshot."""
    from adamml_b200.graph import step_guard
    opt = torch.optim.SGD(model.main_net.parameters(), 0.01, momentum=0.9)
    snap = step_guard(model, opt)
    s0 = snap()
    assert snap() == s0
    model.decay_temperature()
    s1 = snap()
    assert s1 != s0
    model.freeze_policy_net()
    s2 = snap()
    assert s2 != s1
    model.unfreeze_policy_net()
    assert snap() == s1
    opt.param_groups[0]["lr"] = 0.001
    s3 = snap()
    assert s3 != s1
    was = model.training
    model.train(not was)
    assert snap() != s3
    model.train(was)
    model.policy_net.set_temperature(5.0)
