"""Host-side mirror of Open-Sora's STDiT2 (ppdiffusers/examples/Open-Sora/models/stdit/stdit2.py)."""
from .stdit2 import STDiT2, STDiT2Config  # noqa: F401
