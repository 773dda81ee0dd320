"""Host-side mirror of the reference's `models` package (models/__init__.py)."""
from .adamml import AdaMML, adamml
from .model_builder import MODEL_TABLE, build_model
from .resnet import ResNet, resnet
from .sound_mobilenet_v2 import MobileNetV2, sound_mobilenet_v2

__all__ = ["adamml", "resnet", "sound_mobilenet_v2", "build_model", "MODEL_TABLE", "AdaMML", "ResNet", "MobileNetV2"]
