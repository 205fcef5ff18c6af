"""The C-ABI library loads and exports every symbol include/sift_b200.h declares (no GPU, no
compute calls), and fails loudly -- never falls back -- without a device."""
import ctypes
import os
import subprocess

import pytest

import sift_project_b200 as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(S.library_path()):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "sift_project_b200", "csrc"), "-j8"])
    return S.load_library()


def test_every_declared_symbol_is_exported(lib):
    names = S.declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n


def test_default_params_are_the_reference_defaults(lib):
    p = S.SiftParams()
    lib.sift_b200_default_params(ctypes.byref(p))
    assert (p.double_image_size, p.init_sigma, p.intervals, p.window_size) == (1, 1.6, 3, 3)
    assert (p.contrast_threshold, p.eigen_ratio, p.num_bins, p.peak_ratio) == (0.04, 10.0, 36.0, 0.8)
    assert (p.ori_sigma_factor, p.desc_scale_factor, p.max_octaves) == (1.5, 3.0, 0)


def test_keypoint_record_is_168_bytes():
    assert S.KP_DTYPE.itemsize == 168
    assert S.KP_DTYPE.fields["desc"][1] == 40


def test_no_device_is_a_loud_error_not_a_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(S.SiftError) as e:
        S.SiftContext(64, 64)
    assert e.value.code == 2 and "no CPU path" in str(e.value)


def test_oversized_context_is_rejected_before_touching_a_device(lib):
    """Plane offsets are 32-bit inside the scale-space kernels: a context whose doubled base plane would reach
    2^31 pixels is refused with E_UNSUPPORTED (argument validation runs before the device query)."""
    with pytest.raises(S.SiftError) as e:
        S.SiftContext(40000, 40000)
    assert e.value.code == 5 and "too large" in str(e.value)
    with pytest.raises(S.SiftError) as e:
        S.SiftContext(1, 64)
    assert e.value.code != 0


def test_product_never_references_the_oracle():
    """oracle/ is test infrastructure: nothing under sift_project_b200/ or include/ may name it."""
    bad = []
    for base in ("sift_project_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            if "build" in dp:
                continue
            for f in fs:
                if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py", ".hh")):
                    t = open(os.path.join(dp, f), errors="ignore").read()
                    if "liboracle" in t or "oracle/" in t or "import oracle" in t or "from oracle" in t:
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
