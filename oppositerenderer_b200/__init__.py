"""B200-native light-transport hot path of Opposite Renderer (PPM + vertex-connection BDPT).

The product is ``csrc/liboppositerenderer_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/opr.h``).  This package is the thin Python mirror of the reference's renderer-facing surface
used by the tests and the benchmark; there is no CPU fallback.
"""
from . import capi
from .renderer import (OptixRenderer, ComputeDevice, RenderServerRenderRequestDetails, RenderMethod, Camera, Light,
                       Cornell, CornellSmall, Scene, SceneDescription, ppm_radius_schedule)

__all__ = ["capi", "OptixRenderer", "ComputeDevice", "RenderServerRenderRequestDetails", "RenderMethod", "Camera", "Light",
           "Cornell", "CornellSmall", "Scene", "SceneDescription", "ppm_radius_schedule"]
