"""The C-ABI library loads and exports every symbol include/linemod_b200.h declares; compute entry points fail loudly
(LM_E_CUDA) instead of falling back when no GPU is present.  No GPU needed."""
import os
import re

import numpy as np
import pytest

import common
from linemod_pose_estimation_b200 import Detector, LinemodError, _capi

HEADER = os.path.join(common.ROOT, "include", "linemod_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lm_[A-Za-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    declared = _declared_functions()
    assert len(declared) >= 35
    assert sorted(_capi.EXPORTS) == declared


def test_library_exports_every_declared_symbol():
    lib = _capi.lib()
    missing = [n for n in _declared_functions() if not hasattr(lib, n)]
    assert missing == []


def test_header_cites_reference_call_sites():
    text = open(HEADER).read()
    for cite in ("src/rgbdDetector.cpp:31-34", "src/renderer.cpp:308", "src/rgbdDetector.cpp:1668-1680",
                 "src/renderer.cpp:56-70", "src/renderer.cpp:179-185"):
        assert cite in text, cite


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may import, link or name it."""
    pkg = os.path.join(common.ROOT, "linemod_pose_estimation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liblinemod_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_compute_entry_points_fail_loudly_without_gpu():
    det = Detector()
    bgr = np.zeros((480, 640, 3), np.uint8)
    depth = np.zeros((480, 640), np.uint16)
    with pytest.raises(LinemodError) as e:
        det.match([bgr, depth], 90.0)
    assert e.value.code == _capi.LM_E_CUDA and "no CPU path" in str(e.value)
    with pytest.raises(LinemodError) as e:
        det.addTemplate([bgr, depth], "obj", np.full((480, 640), 255, np.uint8))
    assert e.value.code == _capi.LM_E_CUDA
    with pytest.raises(LinemodError) as e:
        det.open_stream([(90.0, [])])
    assert e.value.code == _capi.LM_E_CUDA
    with pytest.raises(LinemodError) as e:
        det.match_batch_multi([[bgr, depth]] * 3, [(90.0, [])])
    assert e.value.code == _capi.LM_E_CUDA
    with pytest.raises(LinemodError):
        det.build_front([bgr, depth])


def test_argument_validation_matches_reference_asserts():
    det = Detector()
    bgr = np.zeros((480, 640, 3), np.uint8)
    with pytest.raises(LinemodError) as e:  # CV_Assert(sources.size() == modalities.size())
        det.match([bgr], 90.0)
    assert e.value.code == _capi.LM_E_INVALID and "sources.size()" in str(e.value)
    with pytest.raises(LinemodError) as e:  # wrong depth type
        det.match([bgr, np.zeros((480, 640), np.uint8)], 90.0)
    assert e.value.code == _capi.LM_E_INVALID
    f64 = np.zeros((64, 3), np.int32)
    with pytest.raises(LinemodError) as e:  # CV_Assert(features.size() <= 63)
        det.addSyntheticTemplate([(10, 10, 0, f64)] * 4, "obj")
    assert "63" in str(e.value)
    with pytest.raises(LinemodError):       # pyramid size must be levels * modalities
        det.addSyntheticTemplate([(10, 10, 0, np.zeros((3, 3), np.int32))], "obj")
    with pytest.raises(LinemodError):       # labels are 0..7
        det.addSyntheticTemplate([(10, 10, 0, np.array([[1, 1, 9]], np.int32))] * 4, "obj")
    bad = np.zeros(256, np.uint8)
    bad[5] = 5
    with pytest.raises(LinemodError):       # responses must stay <= 4 for exact u8 accumulation
        det.set_similarity_lut(bad)
    with pytest.raises(LinemodError):
        Detector(T=(5, 8, 8, 8, 8))


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_template_generation_fails_loudly_without_gpu():
    """Rendering / batched addTemplate / depth check need the GPU and say so; the view sphere is host code."""
    from linemod_pose_estimation_b200 import Mesh, ViewSphere, camera, training
    from common import synth
    det = Detector()
    bgr = np.zeros((480, 640, 3), np.uint8)
    depth = np.zeros((480, 640), np.uint16)
    T, up = ViewSphere(4, 80, 0.5, 0.5, 0.1).views()
    assert len(T) == 4 * 3   # angles -80, 0, 80
    mesh = Mesh(synth.box_mesh())
    for call in (lambda: training.render_views(det, mesh, camera(), T, up),
                 lambda: det.trainViews(mesh, camera(), T, up, "obj"),
                 lambda: det.addTemplates([([bgr, depth], np.full((480, 640), 255, np.uint8))], "obj"),
                 lambda: training.depth_diff(det, depth, mesh, camera(), T[:1], up[:1], [0], [0])):
        with pytest.raises(LinemodError) as e:
            call()
        assert e.value.code == _capi.LM_E_CUDA, str(e.value)


def test_modality_process_fails_loudly_without_a_gpu_and_checks_arguments():
    """Modality::process quantises on the GPU: no CPU path; argument errors are reported before any device work."""
    import torch
    from linemod_pose_estimation_b200 import ColorGradient, DepthNormal, process
    bgr = np.zeros((96, 128, 3), np.uint8)
    with pytest.raises(LinemodError) as e:
        process(ColorGradient(), bgr, np.zeros((90, 128), np.uint8))       # mask.size() != src.size()
    assert e.value.code == _capi.LM_E_INVALID
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    for mod, src in ((ColorGradient(), bgr), (DepthNormal(), np.zeros((96, 128), np.uint16))):
        with pytest.raises(LinemodError) as e:
            process(mod, src)
        assert e.value.code == _capi.LM_E_CUDA, str(e.value)


def test_device_group_grid_argument_check():
    import ctypes as C
    det = Detector()
    h = C.c_void_p()
    devs = (C.c_int * 3)(0, 0, 0)
    assert _capi.lib().lm_group_create_grid(det._h, devs, 3, 2, C.byref(h)) == _capi.LM_E_INVALID   # 2 does not divide 3
    assert b"divide" in _capi.lib().lm_last_error()


def test_device_group_fails_loudly_without_a_gpu():
    """lm_group_create has no CPU path either: without a CUDA device it returns LM_E_CUDA."""
    import ctypes as C

    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from linemod_pose_estimation_b200 import Detector, LinemodError, DetectorGroup
    with pytest.raises(LinemodError) as e:
        DetectorGroup(Detector(), [0], "frames")
    assert e.value.code == -2
