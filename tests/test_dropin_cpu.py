"""The drop-in tree exposes the reference's import surface (train_tensoIR.py:11-12)."""
import importlib
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dropin_import_surface():
    sys.path.insert(0, os.path.join(REPO, "dropin"))
    try:
        for m in ("models", "models.tensoRF_rotated_lights", "models.relight_utils", "renderer"):
            sys.modules.pop(m, None)
        rot = importlib.import_module("models.tensoRF_rotated_lights")
        for n in ("raw2alpha", "TensorVMSplit", "AlphaGridMask"):
            assert hasattr(rot, n)
        gen = importlib.import_module("models.tensoRF_general_multi_lights")
        assert issubclass(gen.TensorVMSplit, rot.TensorVMSplit)
        ini = importlib.import_module("models.tensoRF_init")
        assert hasattr(ini, "TensorVMSplit")
        ru = importlib.import_module("models.relight_utils")
        for n in ("render_with_BRDF", "compute_transmittance", "compute_radiance", "GGX_specular",
                  "linear2srgb_torch", "compute_secondary_shading_effects"):
            assert hasattr(ru, n)
        rnd = importlib.import_module("renderer")
        assert hasattr(rnd, "Renderer_TensoIR_train") and hasattr(rnd, "OctreeRender_trilinear_fast")
    finally:
        sys.path.remove(os.path.join(REPO, "dropin"))
        for m in ("models", "models.tensoRF_rotated_lights", "models.tensoRF_general_multi_lights",
                  "models.tensoRF_init", "models.relight_utils", "renderer"):
            sys.modules.pop(m, None)
