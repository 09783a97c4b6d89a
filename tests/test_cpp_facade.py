"""include/linemod_b200.hpp: compiled with g++ against liblinemod_b200.so and driven like the reference drives
cv::linemod::Detector (tests/cpp/facade_test.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "linemod_pose_estimation_b200")


@pytest.fixture(scope="module")
def facade_binary(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("facade") / "facade_test")
    cmd = ["g++", "-std=c++11", "-O1", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "tests", "cpp", "facade_test.cpp"),
           "-o", out, "-L" + PKG, "-llinemod_b200", "-Wl,-rpath," + PKG]
    subprocess.check_call(cmd)
    return out


def test_facade_host_side(facade_binary, tmp_path):
    res = subprocess.run([facade_binary, "host", str(tmp_path)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert "ok host" in res.stdout


@pytest.mark.gpu
def test_facade_match_on_gpu(facade_binary, tmp_path):
    res = subprocess.run([facade_binary, "gpu", str(tmp_path)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert "ok gpu" in res.stdout


@pytest.fixture(scope="module")
def compat_binary(tmp_path_factory):
    """The reference's call sites (readLinemod, writeLinemod, linemod_detection, detector construction) verbatim, against
    the LINEMOD_B200_WITH_OPENCV branch of the facade and the OpenCV 2.4 API stub under tests/cpp/opencv_stub."""
    out = str(tmp_path_factory.mktemp("compat") / "opencv_compat_test")
    cmd = ["g++", "-std=c++11", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "tests", "cpp", "opencv_stub"),
           os.path.join(ROOT, "tests", "cpp", "opencv_compat_test.cpp"), "-o", out, "-L" + PKG, "-llinemod_b200", "-Wl,-rpath," + PKG]
    subprocess.check_call(cmd)
    return out


def test_reference_call_sites_compile_and_round_trip(compat_binary, tmp_path):
    res = subprocess.run([compat_binary, "host", str(tmp_path)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert "ok host" in res.stdout


@pytest.mark.gpu
def test_reference_call_sites_on_gpu(compat_binary, tmp_path):
    res = subprocess.run([compat_binary, "gpu", str(tmp_path)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert "ok gpu" in res.stdout
