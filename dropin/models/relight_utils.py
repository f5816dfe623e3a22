"""models.relight_utils of the reference, served by tensoir_b200 (same names, same argument meaning).
Importers in the reference: renderer.py:6, dataLoader/tensoIR_rotation_setting.py:13, tensoIR_simple.py:12,
tensoIR_general_multi_lights.py:13, scripts/relight_importance.py:21-24 (star import + Environment_Light)."""
import os  # noqa: F401  (scripts/relight_importance.py takes os / np / torch / F from its star import)

import numpy as np  # noqa: F401
import torch  # noqa: F401
import torch.nn.functional as F  # noqa: F401

from tensoir_b200.relight_utils import (GGX_specular, brdf_specular, render_with_BRDF, compute_transmittance,  # noqa: F401
                                        compute_radiance, compute_secondary_shading_effects, linear2srgb_torch,
                                        safe_l2_normalize, grid_sample, sample_ray_equally, compute_visibility,
                                        compute_visibility_and_indirect_light, sph2cart, _convert_sph_conventions,
                                        _clip_0to1_warn_torch)
from tensoir_b200.relight import read_hdr, Environment_Light  # noqa: F401
from tensoir_b200.tensorbase import raw2alpha  # noqa: F401
