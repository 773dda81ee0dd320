"""CPU oracle for the AdaMML hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (adamml_b200/) never does and has no CPU fallback.

What it is
----------
A from-scratch functional restatement (fp32, CPU, explicit state_dict + explicit noise) of
the reference's algorithm for `AdaMML.forward()` and everything below it:

  reference function                                 restated here as
  -------------------------------------------------  ---------------------------
  models/adamml.py:42-67   AdaMML.data_layer          data_layer()
  models/adamml.py:69-91   AdaMML.forward             adamml_forward()
  models/policy_net.py:312-373 PolicyNet.forward      policy_forward()
  models/policy_net.py:283-290 wrapper_gumbel_softmax gumbel_hard()
  models/policy_net.py:235-247 JointMobileNetV2.features  (inside policy_forward)
  models/policy_net.py:54-149  policy MobileNetV2     policy_mobilenet_features()
  models/common.py:4-33    TemporalPooling            temporal_pool()
  models/joint_resnet_mobilenetv2.py:84-128 forward   main_forward()
  models/resnet.py:77-113,195-223 Bottleneck/ResNet   resnet_forward()
  models/sound_mobilenet_v2.py:42-69,143-162          sound_mobilenet_forward()
  utils/utils.py:166-184   compute_policy_loss        policy_loss()

The arithmetic itself lives in a third-party dependency that is NOT under /root/reference:
PyTorch (unpinned by the reference, README.md:19-21; this image: torch 2.11.0+cu128, CPU =
oneDNN/MKL).  The restatement therefore calls torch's own published CPU primitives
(F.conv2d, F.batch_norm, F.max_pool2d/3d, F.interpolate, torch.nn LSTMCell equations) at
the reference's call sites, but owns all structure: layer wiring, segment loops, temporal
pooling, gating, fusion, the Gumbel straight-through estimator and the RNG draw order.

Pinning
-------
The reference has no tests or golden vectors ("parity unpinned by reference tests",
SURVEY.md §8c).  The oracle is instead pinned against outputs of the reference itself run
in the build container: tests/golden/make_golden.py imports /root/reference, runs it on
seeded inputs and commits logits / decisions / gradient fingerprints under tests/golden/;
tests/test_oracle_golden.py replays them through this file.

Noise: the reference draws from torch's global generator.  `draw_noise()` replays the
exact draw order (S x exponential_ for Gumbel, then per segment per modality the dropout
bernoulli_) so that oracle, reference and CUDA path can consume identical numbers.
"""
import math
import zlib

import torch
import torch.nn.functional as F

RESNET_LAYERS = {18: [2, 2, 2, 2], 34: [3, 4, 6, 3], 50: [3, 4, 6, 3], 101: [3, 4, 23, 3], 152: [3, 8, 36, 3]}
MBV2_CFG = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]
INPUT_CHANNELS = {"rgb": 3, "flow": 10, "rgbdiff": 15, "sound": 1}  # train_adamml.py:86-95


# ----------------------------------------------------------------------------- config
def make_cfg(modality, groups=8, num_segments=5, depth=50, num_classes=31, dropout=0.5, pooling_method="max",
             without_t_stride=False, learnable_lf_weights=True, causality_modeling="lstm"):
    modality = list(modality)
    if "rgbdiff" in modality and "flow" in modality:  # adamml.py:143-147
        p_mod = [m for m in modality if m != "flow"]
        m_mod = [m for m in modality if m != "rgbdiff"]
    else:
        p_mod, m_mod = modality, modality
    return dict(modality=modality, p_modality=p_mod, m_modality=m_mod, groups=groups, num_segments=num_segments,
                depth=depth, num_classes=num_classes, dropout=dropout, pooling_method=pooling_method,
                without_t_stride=without_t_stride, learnable_lf_weights=learnable_lf_weights,
                causality_modeling=causality_modeling, p_frames=max(1, groups // 2))


# ----------------------------------------------------------------------------- noise
def draw_noise(seed, cfg, N, S, training, generator=None):
    """Replay of the reference's RNG consumption for one forward (CPU generator).

    Order (probed against the reference): S draws of exponential_ on
    [M*N, 2] (F.gumbel_softmax, policy_net.py:288), then for each segment, for each main
    modality, one bernoulli_(1-p) on the dropout input ([N*T', 2048] ResNet / [N, 1280]
    sound) — dropout only in training mode, Gumbel always.
    """
    g = generator
    if g is None:
        g = torch.Generator()
        g.manual_seed(seed)
    M = len(cfg["p_modality"])
    if cfg["causality_modeling"] is None:  # one F.gumbel_softmax call over all (modality, segment, video) rows
        expo = [torch.empty(M * S * N, 2).exponential_(generator=g)]
    else:
        expo = [torch.empty(M * N, 2).exponential_(generator=g) for _ in range(S)]
    drop = []
    if training and cfg["dropout"] > 0:
        p = cfg["dropout"]
        tprime = resnet_out_frames(cfg)
        for _ in range(S):
            per_mod = []
            for m in cfg["m_modality"]:
                shape = (N, 1280) if m == "sound" else (N * tprime, 2048 if cfg["depth"] >= 50 else 512)
                per_mod.append(torch.empty(shape).bernoulli_(1 - p, generator=g).div_(1 - p))
            drop.append(per_mod)
    return dict(expo=expo, drop=drop)


def resnet_out_frames(cfg):
    t = cfg["groups"]
    if not cfg["without_t_stride"]:
        for _ in range(3):
            t = max(1, t // 2)
    return t


# ----------------------------------------------------------------------------- primitives
def bn(sd, key, x, training, momentum=0.1, eps=1e-5):
    """nn.BatchNorm2d: batch stats (biased var) in training + running update (unbiased var)."""
    w, b = sd[key + ".weight"], sd[key + ".bias"]
    rm, rv = sd[key + ".running_mean"], sd[key + ".running_var"]
    if training and (key + ".num_batches_tracked") in sd:
        sd[key + ".num_batches_tracked"] += 1
    return F.batch_norm(x, rm, rv, w, b, training, momentum, eps)


def temporal_pool(x, frames, mode="max"):
    """common.py:28-33: [N*T,C,H,W] -> pool k3 s2 p1 over T -> [N*T',C,H,W]."""
    nt, c, h, w = x.shape
    v = x.view(-1, frames, c, h, w).transpose(1, 2)
    if mode == "max":
        v = F.max_pool3d(v, (3, 1, 1), (2, 1, 1), (1, 0, 0))
    else:
        v = F.avg_pool3d(v, (3, 1, 1), (2, 1, 1), (1, 0, 0))
    return v.transpose(1, 2).contiguous().view(-1, c, h, w)


# ----------------------------------------------------------------------------- ResNet
def resnet_forward(sd, pre, x, cfg, training, drop_mask):
    """resnet.py:195-223.  x: [N, F*C, H, W] -> logits [N, classes]."""
    depth = cfg["depth"]
    frames = cfg["groups"]
    n, ct, h, w = x.shape
    if ct != 1:
        x = x.view(n * frames, ct // frames, h, w)
    x = F.conv2d(x, sd[pre + "conv1.weight"], None, 2, 3)
    x = F.relu(bn(sd, pre + "bn1", x, training))
    x = F.max_pool2d(x, 3, 2, 1)
    bottleneck = depth >= 50
    for li, nblocks in enumerate(RESNET_LAYERS[depth]):
        for bi in range(nblocks):
            bp = f"{pre}layer{li + 1}.{bi}."
            stride = 2 if (li > 0 and bi == 0) else 1
            identity = x
            if bottleneck:  # resnet.py:93-113 (stride on the 3x3)
                out = F.relu(bn(sd, bp + "bn1", F.conv2d(x, sd[bp + "conv1.weight"]), training))
                out = F.relu(bn(sd, bp + "bn2", F.conv2d(out, sd[bp + "conv2.weight"], None, stride, 1), training))
                out = bn(sd, bp + "bn3", F.conv2d(out, sd[bp + "conv3.weight"]), training)
            else:  # resnet.py:59-74
                out = F.relu(bn(sd, bp + "bn1", F.conv2d(x, sd[bp + "conv1.weight"], None, stride, 1), training))
                out = bn(sd, bp + "bn2", F.conv2d(out, sd[bp + "conv2.weight"], None, 1, 1), training)
            if (bp + "downsample.0.weight") in sd:
                identity = bn(sd, bp + "downsample.1", F.conv2d(x, sd[bp + "downsample.0.weight"], None, stride),
                              training)
            x = F.relu(out + identity)
        if li < 3 and not cfg["without_t_stride"]:
            x = temporal_pool(x, frames, cfg["pooling_method"])
            frames = max(1, frames // 2)
    x = x.mean((2, 3))
    if training and drop_mask is not None:
        x = x * drop_mask
    x = F.linear(x, sd[pre + "fc.weight"], sd[pre + "fc.bias"])
    return x.view(n, -1, x.shape[-1]).mean(1)


# ----------------------------------------------------------------------------- MobileNetV2 (sound main)
def _sound_block_keys(t):
    # sound_mobilenet_v2.py:52-64: [pw ConvBNReLU] + dw ConvBNReLU + pw-linear conv + bn
    if t != 1:
        return ("conv.0.0", "conv.0.1"), ("conv.1.0", "conv.1.1"), ("conv.2", "conv.3")
    return None, ("conv.0.0", "conv.0.1"), ("conv.1", "conv.2")


def sound_mobilenet_features(sd, pre, x, training):
    x = F.relu6(bn(sd, pre + "features.0.1", F.conv2d(x, sd[pre + "features.0.0.weight"], None, 2, 1), training))
    idx, cin = 1, 32
    for t, c, n, s in MBV2_CFG:
        for i in range(n):
            stride = s if i == 0 else 1
            bp = f"{pre}features.{idx}."
            pw, dw, pl = _sound_block_keys(t)
            out = x
            if pw:
                out = F.relu6(bn(sd, bp + pw[1], F.conv2d(out, sd[bp + pw[0] + ".weight"]), training))
            hid = out.shape[1]
            out = F.relu6(bn(sd, bp + dw[1], F.conv2d(out, sd[bp + dw[0] + ".weight"], None, stride, 1, 1, hid),
                             training))
            out = bn(sd, bp + pl[1], F.conv2d(out, sd[bp + pl[0] + ".weight"]), training)
            x = x + out if (stride == 1 and cin == c) else out
            cin = c
            idx += 1
    x = F.relu6(bn(sd, f"{pre}features.{idx}.1", F.conv2d(x, sd[f"{pre}features.{idx}.0.weight"]), training))
    return x


def sound_mobilenet_forward(sd, pre, x, training, drop_mask):
    """sound_mobilenet_v2.py:152-162."""
    x = sound_mobilenet_features(sd, pre, x, training).mean((2, 3))
    if training and drop_mask is not None:
        x = x * drop_mask
    return F.linear(x, sd[pre + "classifier.1.weight"], sd[pre + "classifier.1.bias"])


# ----------------------------------------------------------------------------- MobileNetV2 (policy)
def policy_mobilenet_features(sd, pre, x, num_frames, training):
    """policy_net.py:142-149 feature_extraction: [N, T*C, h, w] -> [N*T', 1280] (T' = 1 for T in {1,4})."""
    n, ct, h, w = x.shape
    x = x.view(n * num_frames, ct // num_frames, h, w)
    x = F.relu6(bn(sd, pre + "features.0.1", F.conv2d(x, sd[pre + "features.0.0.weight"], None, 2, 1), training))
    idx, cin, frames = 1, 32, num_frames
    for t, c, nrep, s in MBV2_CFG:
        has_tp = c in (64, 160)  # policy_net.py:121
        for i in range(nrep):
            stride = s if i == 0 else 1
            bp = f"{pre}features.{idx}.conv."
            if i == 0 and has_tp and frames not in (0, 1):  # policy_net.py:124-125,57,89-90
                x = temporal_pool(x, frames, "max")
            out = x
            if t == 1:  # policy_net.py:63-72
                hid = out.shape[1]
                out = F.relu6(bn(sd, bp + "1", F.conv2d(out, sd[bp + "0.weight"], None, stride, 1, 1, hid), training))
                out = bn(sd, bp + "4", F.conv2d(out, sd[bp + "3.weight"]), training)
            else:  # policy_net.py:74-86
                out = F.relu6(bn(sd, bp + "1", F.conv2d(out, sd[bp + "0.weight"]), training))
                hid = out.shape[1]
                out = F.relu6(bn(sd, bp + "4", F.conv2d(out, sd[bp + "3.weight"], None, stride, 1, 1, hid), training))
                out = bn(sd, bp + "7", F.conv2d(out, sd[bp + "6.weight"]), training)
            x = x + out if (stride == 1 and cin == c) else out
            cin = c
            idx += 1
        if has_tp:
            frames //= 2  # policy_net.py:130-131 (applies even when num_frames == 1: 1//2 = 0 is never used)
    x = F.relu6(bn(sd, pre + "conv.1", F.conv2d(x, sd[pre + "conv.0.weight"]), training))
    return x.mean((2, 3))


def gumbel_hard(logits, expo, tau):
    """F.gumbel_softmax(hard=True)[:, -1] with injected Exp(1) samples (policy_net.py:283-290)."""
    gumbels = -expo.log()
    y_soft = ((logits + gumbels) / tau).softmax(-1)
    index = y_soft.max(-1, keepdim=True)[1]
    y_hard = torch.zeros_like(logits).scatter_(-1, index, 1.0)
    ret = y_hard - y_soft.detach() + y_soft
    return ret[:, -1]


def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    gates = F.linear(x, w_ih, b_ih) + F.linear(h, w_hh, b_hh)
    i, f, g, o = gates.chunk(4, 1)
    i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
    c2 = f * c + i * g
    return o * torch.tanh(c2), c2


def policy_forward(sd, p_x, cfg, training, expo, tau):
    """policy_net.py:312-373 (lstm and None branches). p_x: list over policy modalities of [S,N,FC,h,w]."""
    pre = "policy_net."
    mods = cfg["p_modality"]
    M, S = len(mods), p_x[0].shape[0]
    outs = []
    for s in range(S):
        feats = []
        for mi, m in enumerate(mods):
            frames = 1 if m == "sound" else cfg["p_frames"]
            feats.append(policy_mobilenet_features(sd, f"{pre}joint_net.nets.{mi}.", p_x[mi][s], frames, training))
        f = torch.cat(feats, 1)
        f = F.relu(F.linear(f, sd[pre + "joint_net.joint.0.weight"], sd[pre + "joint_net.joint.0.bias"]))
        f = F.relu(F.linear(f, sd[pre + "joint_net.joint.2.weight"], sd[pre + "joint_net.joint.2.bias"]))
        outs.append(f)
    N = outs[0].shape[0]
    if cfg["causality_modeling"] is None:  # policy_net.py:330-339
        o = torch.stack(outs).view(S * N, -1)
        logits = torch.cat([F.linear(o, sd[f"{pre}fcs.{mi}.weight"], sd[f"{pre}fcs.{mi}.bias"]) for mi in range(M)])
        dec = gumbel_hard(logits, expo[0], tau)
        return dec.view(M, S, -1).transpose(0, 1), logits.view(M, S, -1, 2).transpose(0, 1)
    decisions, all_logits = [], []
    h = c = logits = None
    for s in range(S):
        if s == 0:
            lstm_in = torch.cat((outs[s], torch.zeros(N, 2 * M)), -1)
            h = torch.zeros(N, 256)
            c = torch.zeros(N, 256)
        else:
            fb = logits.view(M, -1, 2).permute(1, 0, 2).contiguous().view(-1, 2 * M)
            lstm_in = torch.cat((outs[s], fb), -1)
        h, c = lstm_cell(lstm_in, h, c, sd[pre + "lstm.weight_ih"], sd[pre + "lstm.weight_hh"],
                         sd[pre + "lstm.bias_ih"], sd[pre + "lstm.bias_hh"])
        logits = torch.cat([F.linear(h, sd[f"{pre}fcs.{mi}.weight"], sd[f"{pre}fcs.{mi}.bias"]) for mi in range(M)])
        all_logits.append(logits.view(M, -1, 2))
        decisions.append(gumbel_hard(logits, expo[s], tau))
    return torch.stack(decisions).view(S, M, -1), torch.stack(all_logits)


# ----------------------------------------------------------------------------- main net + wrapper
def main_forward(sd, xs, decisions, cfg, training, drop_masks):
    """joint_resnet_mobilenetv2.py:84-128 (fusion_point='logits'). xs: list over main modalities of [N,FC,H,W]."""
    pre = "main_net."
    out = []
    for mi, m in enumerate(cfg["m_modality"]):
        dm = drop_masks[mi] if drop_masks else None
        if m == "sound":
            t = sound_mobilenet_forward(sd, f"{pre}nets.{mi}.", xs[mi], training, dm)
        else:
            t = resnet_forward(sd, f"{pre}nets.{mi}.", xs[mi], cfg, training, dm)
        if decisions is not None:
            t = t * decisions[mi].view(t.shape[0], 1)
        out.append(t)
    out = torch.stack(out)
    if (pre + "lf_weights") in sd:
        lf = sd[pre + "lf_weights"]
        w = torch.cat((lf, torch.ones(1) - lf.sum(0, keepdim=True)))
        return (out * w.view(-1, 1, 1)).sum(0)
    return out.mean(0)


def data_layer(x, cfg, S, p_size=(160, 160)):
    """adamml.py:42-67."""
    p_x, m_x = [], []
    F_ = cfg["groups"]
    for x_, m in zip(x, cfg["modality"]):
        if m == "sound":
            if x_.shape[-1] != x_.shape[-2]:
                t = torch.stack(x_.chunk(S, dim=-1), 0).contiguous()
            else:
                t = x_.view(x_.shape[0], S, -1, *x_.shape[-2:]).transpose(0, 1).contiguous()
            p_x.append(t)
            m_x.append(t)
            continue
        if m in cfg["p_modality"]:
            b = x_.shape[0]
            t = F.interpolate(x_, size=p_size, mode="bilinear")
            t = t.view(b, S, F_, -1, *p_size)[:, :, list(range(0, F_, 2))]
            p_x.append(t.reshape(b, S, -1, *p_size).transpose(0, 1).contiguous())
        if m in cfg["m_modality"]:
            m_x.append(x_.view(x_.shape[0], S, -1, *x_.shape[-2:]).transpose(0, 1).contiguous())
    return p_x, m_x


def adamml_forward(sd, x, cfg, training, noise, tau=5.0, num_segments=None):
    """adamml.py:69-91 -> (logits [N,classes], decisions [N,S,M])."""
    S = num_segments or cfg["num_segments"]
    p_x, m_x = data_layer(x, cfg, S)
    decisions, _ = policy_forward(sd, p_x, cfg, training, noise["expo"], tau)
    all_logits = []
    for s in range(S):
        xs = [m_x[mi][s] for mi in range(len(cfg["m_modality"]))]
        dm = noise["drop"][s] if (training and noise.get("drop")) else None
        all_logits.append(main_forward(sd, xs, decisions[s], cfg, training, dm))
    return torch.stack(all_logits, 1).mean(1), decisions.permute(2, 0, 1)


def policy_loss(selection, cost_weights, gammas, cls_logits, cls_targets):
    """utils/utils.py:166-184, 'blockdrop' (incl. its [N]x[N,1] -> [N,N] broadcast)."""
    M = selection.shape[-1]
    loss = torch.tensor(0.0)
    correct = (cls_logits.detach().argmax(-1) == cls_targets).type_as(cls_logits)
    sel = selection.mean(1)
    sel = sel * sel
    for w, pl in zip(cost_weights, sel.chunk(M, dim=-1)):
        loss = loss + w * torch.mean(correct * pl)
    return loss + torch.mean((torch.ones_like(correct) - correct) * gammas)


# ----------------------------------------------------------------------------- deterministic params / inputs
def _gen(key, seed):
    g = torch.Generator()
    g.manual_seed((zlib.crc32(key.encode()) + 7919 * seed) % (2 ** 31))
    return g


def fill_state_dict(shapes, seed=0):
    """Deterministic, reference-independent parameter values keyed by state_dict name.

    shapes: {key: torch.Size}.  Returns {key: tensor}.  BN running stats are non-trivial so that
    eval-mode folding bugs show; weights are fan-in scaled so activations stay O(1).
    """
    sd = {}
    for key in sorted(shapes):
        shp = tuple(shapes[key])
        g = _gen(key, seed)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            sd[key] = torch.zeros(shp, dtype=torch.long)
        elif leaf == "running_mean":
            sd[key] = 0.1 * torch.randn(shp, generator=g)
        elif leaf == "running_var":
            sd[key] = 0.5 + torch.rand(shp, generator=g)
        elif leaf == "lf_weights":
            sd[key] = torch.full(shp, 0.8 / (shp[0] + 1)) + 0.05 * torch.rand(shp, generator=g)
        elif len(shp) == 1 and leaf == "weight":  # BN gamma
            sd[key] = 0.5 + torch.rand(shp, generator=g)
        elif len(shp) == 1:  # biases (BN beta, linear, lstm)
            sd[key] = 0.2 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            sd[key] = torch.randn(shp, generator=g) * (1.4 / math.sqrt(fan_in))
    return sd


def make_inputs(cfg, N, S, seed=123, hw=224, sound_hw=256):
    """Synthetic clips of the BASELINE shape: list aligned with cfg['modality']."""
    g = torch.Generator()
    g.manual_seed(seed)
    xs = []
    for m in cfg["modality"]:
        if m == "sound":
            xs.append(torch.randn(N, S, sound_hw, sound_hw, generator=g))
        else:
            xs.append(torch.randn(N, S * cfg["groups"] * INPUT_CHANNELS[m], hw, hw, generator=g))
    y = torch.randint(0, cfg["num_classes"], (N,), generator=g)
    return xs, y


def clone_sd(sd, requires_grad=True):
    out = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if requires_grad and t.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            t.requires_grad_(True)
        out[k] = t
    return out
