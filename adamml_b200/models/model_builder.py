"""Plugin registry + builder with the reference's contract (models/model_builder.py:3-38)."""
from .adamml import adamml
from .resnet import resnet
from .sound_mobilenet_v2 import sound_mobilenet_v2

MODEL_TABLE = {
    "adamml": adamml,
    "resnet": resnet,
    "sound_mobilenet_v2": sound_mobilenet_v2,
}


def build_model(args, test_mode=False):
    """args: the flat opts.py namespace (+ num_classes, input_channels).  -> (model, arch_name)."""
    model = MODEL_TABLE[args.backbone_net](**vars(args))
    network_name = model.network_name if hasattr(model, "network_name") else args.backbone_net
    modality = "-".join(args.modality) if isinstance(args.modality, list) else args.modality
    arch_name = "{}-{}-{}-f{}".format(args.dataset, modality, network_name, args.groups)
    if args.dense_sampling:
        arch_name += "-s{}".format(args.frames_per_group)
    if not test_mode:
        arch_name += "-{}{}-bs{}{}-e{}".format(args.lr_scheduler, "-syncbn" if args.sync_bn else "", args.batch_size,
                                               "-" + args.prefix if args.prefix else "", args.epochs)
    return model, arch_name
