"""Drop-in shim: `from diff_gaussian_rasterization import GaussianRasterizationSettings,
GaussianRasterizer` (the import DMGS performs at gaussian_renderer/__init__.py:14) resolves to the
B200-native implementation in dmgs_b200 when this repository is on sys.path."""
from dmgs_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                  rasterize_gaussians)
