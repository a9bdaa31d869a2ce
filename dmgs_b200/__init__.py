"""dmgs_b200 -- B200-native rasteriser + mesh binding for DMGS."""
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, check_async, configure  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "configure", "check_async"]
