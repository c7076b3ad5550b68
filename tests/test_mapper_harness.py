"""the reference's UNMODIFIED mapper layer (src/mapper.cc, src/sfm/sequential_mapper.cc and every other translation
unit the shim does not replace) compiled where it lies under /root/reference and LINKED against the drop-in shim +
libmavmap_b200.so.  Eigen / OpenCV / Boost / glog are not in the image: tests/mapper_harness/stubs holds stand-ins for the
parts of them the reference touches.  Only possible where the reference tree is mounted (the build container)."""
import os
import subprocess

import pytest

from conftest import ROOT

REF = "/root/reference/src"
HARNESS = os.path.join(ROOT, "tests", "mapper_harness")
OUT = os.path.join(ROOT, "build", "mapper_harness")

# what sequential_mapper.cc / mapper.cc call on the hot path (SURVEY 8b "what calls it"): all must come from the shim objects
HOT_PATH_SYMBOLS = ["bundle_adjustment(", "pose_refinement(", "match_brute_force(", "triangulate_point(", "triangulate_points(",
                    "calc_tri_angles(", "calc_reproj_errors(", "calc_depth("]


@pytest.fixture(scope="module")
def mapper_exe():
    if not os.path.exists(os.path.join(REF, "mapper.cc")):
        pytest.skip("reference tree not present")
    if not os.path.exists(os.path.join(ROOT, "mavmap_b200", "libmavmap_b200.so")):
        pytest.skip("libmavmap_b200.so not built")
    subprocess.check_call(["sh", os.path.join(HARNESS, "build_mapper.sh"), REF, OUT])
    return os.path.join(OUT, "mapper")


def _nm(path, *flags):
    return subprocess.run(["nm", "-C"] + list(flags) + [path], stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()


def test_reference_mapper_links_against_the_shim(mapper_exe):
    assert os.path.getsize(mapper_exe) > 0
    # the callers reference the hot-path functions as undefined symbols with the reference's signatures ...
    undefined = [l for l in _nm(os.path.join(OUT, "ref_sfm_sequential_mapper.o"), "-u")]
    for sym in ["bundle_adjustment(", "pose_refinement(", "match_brute_force(", "triangulate_points(", "calc_tri_angles(", "calc_reproj_errors(", "calc_depth("]:
        assert any(sym in l for l in undefined), sym
    # ... and in the linked program each of them is a strong definition that comes from a shim object
    shim_defs = []
    for o in ("shim_base3d_bundle_adjustment.o", "shim_base3d_triangulation.o", "shim_base2d_feature_match.o"):
        shim_defs += [l for l in _nm(os.path.join(OUT, o), "--defined-only") if " T " in l]
    linked = [l for l in _nm(mapper_exe, "--defined-only")]
    for sym in HOT_PATH_SYMBOLS:
        assert any(sym in l for l in shim_defs), "shim does not define " + sym
        hits = [l for l in linked if sym in l and " T " in l]
        assert hits, "not a strong definition in the linked mapper: " + sym
    # the library behind the shim is a dependency of the program
    needed = subprocess.run(["readelf", "-d", mapper_exe], stdout=subprocess.PIPE, text=True, check=True).stdout
    assert "libmavmap_b200.so" in needed


def test_reference_mapper_runs_its_own_option_parser(mapper_exe):
    out = subprocess.run([mapper_exe, "--help"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
    assert out.returncode == 1                     # mapper.cc:922-925 prints the options and returns 1
    for opt in ("--input-path", "--refine-camera-params", "--local-ba-refine-camera-params", "--constrain-rotation", "--match-max-ratio"):
        assert opt in out.stdout, opt
    missing = subprocess.run([mapper_exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
    assert missing.returncode == 1 and "required" in missing.stdout


@pytest.mark.gpu
def test_reference_sequential_mapper_runs_global_ba_through_the_shim(mm):
    """The reference's own SequentialMapper::adjust_bundle / adjust_global_bundle (compiled from /root/reference in the build
    container; the binary travels to the GPU box under build/) on a synthetic 20-image sequence: mapper -> shim -> C ABI -> CUDA."""
    exe = os.path.join(OUT, "global_ba_driver")
    if os.path.exists(os.path.join(REF, "mapper.cc")):
        subprocess.check_call(["sh", os.path.join(HARNESS, "build_mapper.sh"), REF, OUT])
    if not os.path.exists(exe):
        pytest.skip("harness binary not built (needs the reference tree at build time)")
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    print(out.stdout)
    assert out.returncode == 0, out.stdout
    assert out.stdout.count("ok  ") >= 7 and "FAIL" not in out.stdout
