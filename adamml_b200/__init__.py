"""adamml_b200 — B200-native (sm_100a) implementation of the AdaMML data-parallel hot path.

Host-side mirror of the reference interface lives in `adamml_b200.models`
(`build_model`, `MODEL_TABLE`, `AdaMML.forward`); the arithmetic lives in
`adamml_b200/csrc/*.cu` behind the C-ABI declared in `include/adamml_b200.h`.
"""
__version__ = "0.1.0"
