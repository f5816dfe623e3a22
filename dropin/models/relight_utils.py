"""models.relight_utils of the reference, served by tensoir_b200.relight_utils."""
from tensoir_b200.relight_utils import *  # noqa: F401,F403
from tensoir_b200.relight_utils import (GGX_specular, brdf_specular, render_with_BRDF, compute_transmittance,  # noqa: F401
                                        compute_radiance, compute_secondary_shading_effects, linear2srgb_torch,
                                        safe_l2_normalize)
