"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/mavmap_b200.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from mavmap_b200 import _abi, _lib


def header_symbols():
    src = open(os.path.join(ROOT, "include", "mavmap_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 28
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
    # and the Python prototype table covers exactly the header
    assert sorted(_abi.PROTOTYPES) == syms


def test_struct_layouts_match_header_sizes():
    # offsets that the C side relies on (LP64): catches accidental field reordering in _abi.py
    # numbers printed by gcc (sizeof / offsetof on include/mavmap_b200.h)
    assert C.sizeof(_abi.MatchOptions) == 32
    assert C.sizeof(_abi.BAProblem) == 136 and _abi.BAProblem.rot_prior.offset == 120
    assert _abi.BAOptions.pcg_tolerance.offset == 112 and C.sizeof(_abi.BAOptions) == 144 and _abi.BAOptions.pcg_preconditioner.offset == 128 and _abi.BAOptions.tile_cholesky_tolerance.offset == 136
    assert _abi.BASummary.trace_cost.offset == 48 and C.sizeof(_abi.BASummary) == 16480


def test_host_only_entry_points():
    lib = _lib.lib()
    assert lib.mm_abi_version() == 1
    assert lib.mm_camera_model_name_to_code(b"PINHOLE") == 1       # camera_models.cc:12-21
    assert lib.mm_camera_model_name_to_code(b"OPENCV") == 2
    assert lib.mm_camera_model_name_to_code(b"CATA") == 3
    assert lib.mm_camera_model_name_to_code(b"FISHEYE") == -1
    assert [lib.mm_camera_model_num_params(c) for c in (1, 2, 3, 4)] == [4, 8, 9, -1]
    p = np.array([1000.0, 1200.0, 0, 0, 0, 0, 0, 0, 0])
    assert lib.mm_camera_image2world_threshold(2.2, 1, p.ctypes.data_as(_abi.p_f64)) == pytest.approx(2.2 / 1100.0)   # camera_models.cc:47-52
    o = _abi.BAOptions(); lib.mm_ba_options_default(C.byref(o))
    assert (o.max_num_iterations, o.function_tolerance, o.gradient_tolerance, o.loss_scale) == (100, 1e-4, 1e-8, 1.0)   # bundle_adjustment.h:40-45
    m = _abi.MatchOptions(); lib.mm_match_options_default(C.byref(m))
    assert (m.ratio_test, m.max_ratio, m.max_distance) == (1, 0.6, -1.0)     # feature.h:107-109


def test_no_cpu_fallback_without_device():
    lib = _lib.lib()
    if lib.mm_device_count() > 0:
        pytest.skip("a CUDA device is present")
    xyz = np.ones((4, 3)); uv = np.zeros((4, 2)); p = np.array([1.0, 1, 0, 0, 0, 0, 0, 0, 0])
    rc = lib.mm_camera_world2image(1, p.ctypes.data_as(_abi.p_f64), 4, xyz.ctypes.data_as(_abi.p_f64), uv.ctypes.data_as(_abi.p_f64))
    assert rc == _abi.MM_ERR_NO_DEVICE and not uv.any()
    from mavmap_b200 import synthetic
    from mavmap_b200.ba import default_c_options, solve_flat
    flat, _ = synthetic.make_ba_problem(n_img=4, n_obs_target=200, track_len=3, seed=1)
    with pytest.raises(_lib.MavmapB200Error) as e:
        solve_flat(flat, default_c_options())
    assert e.value.code == _abi.MM_ERR_NO_DEVICE
    import mavmap_b200 as mm
    d = np.zeros((5, 64), np.float32)
    with pytest.raises(_lib.MavmapB200Error):
        mm.match_brute_force(None, d, None, d)
