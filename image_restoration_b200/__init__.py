"""image_restoration_b200 — B200-native (sm_100a) hot path of ChuRuaNh0/Image_Restoration's Car_Plate-Restoration:
the GFPGANv1OCR forward pass and the pyblur degradation, behind the reference's own arch / operator interface.
See DESIGN.md for the design and INTEGRATION.md for the drop-in recipe."""
from .registry import ARCH_REGISTRY, Registry, build_network  # noqa: F401
from .arch import GFPGANv1OCR, register_into  # noqa: F401

__all__ = ['ARCH_REGISTRY', 'Registry', 'build_network', 'GFPGANv1OCR', 'register_into']
