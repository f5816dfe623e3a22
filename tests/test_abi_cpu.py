"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/tensoir_b200.h declares; the product refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from tensoir_b200 import _lib
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(REPO, "include", "tensoir_b200.h")).read()
    names = sorted(set(re.findall(r"^int\s+(tir_\w+)\s*\(", hdr, flags=re.M)))
    assert len(names) >= 20
    from tensoir_b200 import _lib
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.EXPORTS, f"{n} has no ctypes binding"
    assert lib.tir_abi_version() == 1


def test_struct_sizes_match_header(lib):
    from tensoir_b200 import _lib
    assert (ctypes.sizeof(_lib.TirField), ctypes.sizeof(_lib.TirMlp), ctypes.sizeof(_lib.TirMarchCfg)) == (224, 88, 48)
    assert _lib.APP_SAMPLE_BYTES == 24


def test_null_arguments_are_rejected(lib):
    assert lib.tir_pack_channels_last(None, None, 1, 1, 1, None) == -1
    assert lib.tir_density_points(None, None, 5, None, None, None) == -1
    assert lib.tir_density_points(None, None, 0, None, None, None) == 0      # empty input is a no-op, not an error


def test_no_cpu_fallback():
    from tensoir_b200 import _lib
    with pytest.raises(_lib.TirError):
        _lib.dptr(torch.zeros(3))


def test_model_surface_matches_reference_keys(golden_rotated, golden_general, golden_init):
    """state_dict keys / shapes are the reference's (checkpoint compatibility)."""
    from gpu_helpers import model_from_fixture
    for fx in (golden_rotated, golden_general, golden_init):
        m = model_from_fixture(fx, "cpu")
        sd = m.state_dict()
        assert set(sd.keys()) == set(fx["state_dict"].keys())
        for k, v in fx["state_dict"].items():
            assert tuple(sd[k].shape) == tuple(v.shape)
        assert fx["kind"] != "rotated" or m.nSamples == fx["nSamples"]
    m = model_from_fixture(golden_rotated, "cpu")
    groups = m.get_optparam_groups(0.02, 0.001)
    assert len(groups) == 10 and groups[0]["lr"] == 0.02 and groups[5]["lr"] == 0.001
    kw = m.get_kwargs()
    assert kw["light_rotation"] == [0, 120] and kw["gridSize"] == [24, 24, 24]


def test_light_model_matches_oracle(golden_rotated):
    from gpu_helpers import model_from_fixture
    m = model_from_fixture(golden_rotated, "cpu")
    fx = golden_rotated
    assert torch.equal(m.gen_light_incident_dirs(method='fixed_envirmap'), fx["fixed_dirs"])
    assert torch.equal(m.light_area_weight, fx["area_weight"])
    assert torch.equal(m.get_light_rgbs(fx["fixed_dirs"], device='cpu').detach(), fx["light_rgbs"])
    from tensoir_b200.relight_utils import GGX_specular
    nrm, v2c, rough, fres = fx["ggx_in"]
    assert torch.equal(GGX_specular(nrm, v2c, fx["fixed_dirs"][None].repeat(9, 1, 1), rough, fres), fx["ggx"])
