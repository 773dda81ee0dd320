"""CPU: the C-ABI library loads and exports every symbol include/adamml_b200.h declares
(no compute calls without a GPU), and the host-side mirror has the reference's surface."""
import ctypes
import os

import pytest

from util import ROOT, namespace


def test_library_exports_every_declared_symbol():
    from adamml_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 38
    assert os.path.exists(_lib.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in protos if not hasattr(cdll, n)]
    assert not missing, missing
    L = _lib.lib()
    assert L.cdll.adamml_abi_version() == 1
    assert L.last_error() == ""


def test_header_cites_reference_sites():
    src = open(os.path.join(ROOT, "include", "adamml_b200.h")).read()
    for needle in ("models/adamml.py:42-67", "policy_net.py:283-290", "resnet.py", "sound_mobilenet_v2.py",
                   "joint_resnet_mobilenetv2.py:92-97"):
        assert needle in src


def test_argument_validation_without_gpu():
    """Bad geometry is rejected on the host before any launch, with a message."""
    from adamml_b200 import _lib
    L = _lib.lib()
    rc = L.cdll.adamml_simt_conv_fwd(None, None, None, 1, 8, 8, 4, 4, 3, 3, 1, 1, 5, 5, 0, 0, 0, 0, None)
    assert rc == 1 and "Ho/Wo" in L.last_error()
    assert L.cdll.adamml_tc_supported(128, 64, 64, 0, 0, 0) == 1
    assert L.cdll.adamml_tc_supported(128, 63, 64, 0, 0, 0) == 0
    assert L.cdll.adamml_tc_supported(128, 64, 3, 0, 0, 0) == 0


@pytest.mark.parametrize("mods", [["rgb", "sound"], ["rgb", "flow", "rgbdiff"], ["rgb", "sound", "flow", "rgbdiff"]])
def test_model_surface_matches_reference_contract(mods):
    from adamml_b200.models import MODEL_TABLE, build_model
    assert set(MODEL_TABLE) == {"adamml", "resnet", "sound_mobilenet_v2"}
    case = dict(kind="adamml", modality=mods, S=5)
    model, arch = build_model(namespace(case))
    assert arch.startswith("kinetics-sounds-" + "-".join(mods) + "-adamml-j_mobilenet_v2-lstm-joint_resnet-50")
    for attr in ("policy_net", "main_net", "update_policy_net", "update_main_net", "freeze_policy_net",
                 "unfreeze_policy_net", "freeze_main_net", "unfreeze_main_net", "decay_temperature", "mean", "std",
                 "network_name", "data_layer"):
        assert hasattr(model, attr), attr
    assert model.policy_net.temperature == 5.0
    model.decay_temperature()
    assert abs(model.policy_net.temperature - 5.0 * 0.965) < 1e-12
    model.freeze_policy_net()
    assert not any(p.requires_grad for p in model.policy_net.parameters()) and not model.update_policy_net
    assert all(p.requires_grad for p in model.main_net.parameters())
    model.unfreeze_policy_net(); model.freeze_main_net()
    assert not any(p.requires_grad for p in model.main_net.parameters())
    m_expected = len(mods) - (1 if ("flow" in mods and "rgbdiff" in mods) else 0)
    assert model.num_modality == m_expected
    assert model.main_net.lf_weights.shape == (m_expected - 1,)
    assert model.policy_net.modality == [m for m in mods if m != "flow" or "rgbdiff" not in mods]
    assert model.main_net.modality == [m for m in mods if m != "rgbdiff" or "flow" not in mods]


def test_unknown_backbone_and_pretrained_errors():
    from adamml_b200.models import build_model
    ns = namespace(dict(kind="resnet", modality=["rgb"], S=1))
    ns.backbone_net = "s3d"
    with pytest.raises(KeyError):
        build_model(ns)
    ns.backbone_net = "resnet"
    ns.imagenet_pretrained = True
    with pytest.raises(RuntimeError):
        build_model(ns)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under adamml_b200/ may reference it."""
    pkg = os.path.join(ROOT, "adamml_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("oracle/", "oracle/").lower() or f == "__none__", (dp, f)


def test_every_call_site_matches_the_header_arity():
    """Static ABI check without a GPU: every `call("op", ...)` in the package passes exactly the parameters the header
    declares for `adamml_op` (minus the trailing stream, which `_lib.call` appends)."""
    import ast
    import glob
    from adamml_b200 import _lib
    protos = _lib.parse_header()
    pkg = os.path.join(ROOT, "adamml_b200")
    files = (glob.glob(os.path.join(pkg, "**", "*.py"), recursive=True) + [os.path.join(ROOT, "bench.py")] +
             glob.glob(os.path.join(ROOT, "scripts", "*.py")) + glob.glob(os.path.join(ROOT, "tests", "*.py")))
    seen, bad = set(), []
    for f in files:
        tree = ast.parse(open(f).read(), f)
        for node in ast.walk(tree):
            if not isinstance(node, ast.Call) or not node.args:
                continue
            fn = node.func
            name = fn.id if isinstance(fn, ast.Name) else (fn.attr if isinstance(fn, ast.Attribute) else None)
            if name != "call" or not isinstance(node.args[0], ast.Constant) or not isinstance(node.args[0].value, str):
                continue
            if any(isinstance(a, ast.Starred) for a in node.args):
                continue
            op = "adamml_" + node.args[0].value
            assert op in protos, f"{f}:{node.lineno}: {op} is not declared in include/adamml_b200.h"
            params = protos[op][1]
            want = len(params) - (1 if params and params[-1][0].startswith("cudaStream_t") else 0)
            got = len(node.args) - 1
            seen.add(op)
            if got != want:
                bad.append((os.path.relpath(f, ROOT), node.lineno, op, got, want))
    assert not bad, bad
    assert len(seen) > 40, len(seen)   # the walk really found the call sites
