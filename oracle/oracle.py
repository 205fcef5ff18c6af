"""TEST INFRASTRUCTURE ONLY -- ctypes front end for the CPU oracles.

Two libraries share one interface here:
  * ``oracle/liboracle.so``        -- our restatement (oracle/sift_oracle.cpp), prefix ``oracle_``
  * ``oracle/_ref/libsift_ref.so`` -- the real reference compiled from /root/reference/src
                                      (oracle/Makefile), prefix ``ref_``
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (sift_project_b200) never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# 168-byte record, identical to the reference's struct Keypoint (sift.hh:15-23)
KP_DTYPE = np.dtype(
    [("x", "<f8"), ("y", "<f8"), ("octave", "<i4"), ("layer", "<i4"), ("size", "<f8"),
     ("pori", "<f8"), ("desc", "u1", (128,))]
)
assert KP_DTYPE.itemsize == 168


def synth_image(h, w, seed=1234):
    """Generator "D" of SURVEY.md 8(d): sum of unit-variance Gaussian-filtered noise fields at
    sigma 2,4,8,16,32 (wrap), mapped affinely to [0,255], rounded to u8."""
    from scipy.ndimage import gaussian_filter

    rng = np.random.default_rng(seed)
    acc = np.zeros((h, w), dtype=np.float64)
    for s in (2, 4, 8, 16, 32):
        n = gaussian_filter(rng.standard_normal((h, w)), s, mode="wrap")
        acc += n / n.std()
    acc = (acc - acc.min()) / (acc.max() - acc.min()) * 255.0
    return np.rint(acc).astype(np.uint8)


def synth_descriptors(n, seed=1234):
    """Config-5 style synthetic descriptors: |N(0,1)| pushed through the reference's
    normalise -> clamp 0.2 -> renormalise -> floor(512 x) -> min 255 (sift.cpp:582-602)."""
    rng = np.random.default_rng(seed)
    h = np.abs(rng.standard_normal((n, 128)))
    h /= np.sqrt((h * h).sum(1, keepdims=True))
    h = np.minimum(h, 0.2)
    h /= np.sqrt((h * h).sum(1, keepdims=True))
    return np.minimum(np.floor(512.0 * h), 255).astype(np.uint8)


class Params(C.Structure):
    """OracleParams / RefParams: the tunable arguments of sift.hh:65-71."""
    _fields_ = [("double_image_size", C.c_int32), ("init_sigma", C.c_double), ("intervals", C.c_int32),
                ("contrast_threshold", C.c_double), ("eigen_ratio", C.c_double), ("peak_ratio", C.c_double),
                ("ori_sigma_factor", C.c_double), ("desc_scale_factor", C.c_double),
                ("window_size", C.c_int32), ("num_bins", C.c_double)]

    def __init__(self, double_image_size=True, init_sigma=1.6, intervals=3, contrast_threshold=0.04,
                 eigen_ratio=10.0, peak_ratio=0.8, ori_sigma_factor=1.5, desc_scale_factor=3.0, window_size=3,
                 num_bins=36):
        super().__init__(int(bool(double_image_size)), init_sigma, intervals, contrast_threshold, eigen_ratio,
                         peak_ratio, ori_sigma_factor, desc_scale_factor, window_size, num_bins)


class _Lib:
    def __init__(self, path, prefix):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.p = prefix
        f = self._f
        f("run_create", C.c_void_p, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int])
        f("run_create_ex", C.c_void_p, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Params), C.c_int])
        f("run_destroy", None, [C.c_void_p])
        f("run_octaves", C.c_int, [C.c_void_p])
        f("run_sigmas", C.c_int, [C.c_void_p, C.c_void_p, C.c_int])
        f("run_layer_dims", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)])
        f("run_gaussian", C.POINTER(C.c_double), [C.c_void_p, C.c_int, C.c_int])
        f("run_dog", C.POINTER(C.c_double), [C.c_void_p, C.c_int, C.c_int])
        f("run_extrema", C.c_int, [C.c_void_p, C.c_void_p, C.c_int])
        f("run_keypoints", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int])
        f("run_orient_given", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int])
        f("run_describe_given", C.c_int, [C.c_void_p, C.c_void_p, C.c_int])

    def _f(self, name, res, args):
        fn = getattr(self.lib, f"{self.p}_{name}")
        fn.restype = res
        fn.argtypes = args
        setattr(self, name, fn)
        return fn


class Run:
    """All stage outputs of one detect call on the CPU (FP64)."""

    def __init__(self, lib, image, double_image_size=True, keep_pyramid=True, params=None):
        self._l = lib
        self.params = params
        img = np.ascontiguousarray(image, dtype=np.float64)
        if img.ndim == 2:
            h, w = img.shape
            c = 1
        else:
            h, w, c = img.shape
        self.h, self.w, self.c = h, w, c
        if params is None:
            self._h = lib.run_create(img.ctypes.data, w, h, c, int(double_image_size), int(keep_pyramid))
        else:
            self._h = lib.run_create_ex(img.ctypes.data, w, h, c, C.byref(params), int(keep_pyramid))
        self.keep = keep_pyramid

    def close(self):
        if self._h:
            self._l.run_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def octaves(self):
        return self._l.run_octaves(self._h)

    def sigmas(self):
        out = np.zeros(8)
        n = self._l.run_sigmas(self._h, out.ctypes.data, 8)
        return out[:n]

    def dims(self, o):
        w, h = C.c_int(), C.c_int()
        self._l.run_layer_dims(self._h, o, C.byref(w), C.byref(h))
        return w.value, h.value

    def _plane(self, fn, o, l):
        w, h = self.dims(o)
        p = fn(self._h, o, l)
        return np.ctypeslib.as_array(p, shape=(h, w)).copy()

    def gaussian(self, o, l):
        return self._plane(self._l.run_gaussian, o, l)

    def dog(self, o, l):
        return self._plane(self._l.run_dog, o, l)

    def extrema(self):
        n = self._l.run_extrema(self._h, None, 0)
        out = np.zeros((n, 4))
        self._l.run_extrema(self._h, out.ctypes.data, n)
        return out

    def keypoints(self, stage=2):
        n = self._l.run_keypoints(self._h, stage, None, 0)
        out = np.zeros(n, dtype=KP_DTYPE)
        self._l.run_keypoints(self._h, stage, out.ctypes.data, n)
        return out

    def orient_given(self, raw):
        """compute_orientations (sift.cpp:447-533) on caller-supplied raw keypoints (doubled-image frame, as
        stage 0 returns them) over this run's pyramid."""
        raw = np.ascontiguousarray(raw, dtype=KP_DTYPE)
        cap = 4 * len(raw) + 16
        out = np.zeros(cap, dtype=KP_DTYPE)
        n = self._l.run_orient_given(self._h, raw.ctypes.data, len(raw), out.ctypes.data, cap)
        if n < 0:
            raise RuntimeError("run was created without keep_pyramid")
        return out[:n].copy()

    def describe_given(self, kps):
        """compute_descriptors (sift.cpp:610-682) on caller-supplied oriented keypoints; returns a copy with
        .desc filled."""
        out = np.ascontiguousarray(kps, dtype=KP_DTYPE).copy()
        n = self._l.run_describe_given(self._h, out.ctypes.data, len(out))
        if n < 0:
            raise RuntimeError("run was created without keep_pyramid")
        return out


_cache = {}


def port():
    """Our restatement (always buildable: oracle/Makefile target liboracle.so)."""
    if "port" not in _cache:
        l = _Lib(os.path.join(HERE, "liboracle.so"), "oracle")
        l.lib.oracle_match.restype = C.c_int
        l.lib.oracle_match.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        l.lib.oracle_gaussian_taps.restype = C.c_int
        l.lib.oracle_gaussian_taps.argtypes = [C.c_double, C.c_void_p, C.c_int]
        l.lib.oracle_blur.restype = None
        l.lib.oracle_blur.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p]
        l.match_fn = l.lib.oracle_match
        _cache["port"] = l
    return _cache["port"]


def have_ref(asshipped=False):
    name = "libsift_ref_asshipped.so" if asshipped else "libsift_ref.so"
    return os.path.exists(os.path.join(HERE, "_ref", name))


def ref(asshipped=False):
    """The real reference (prebuilt into oracle/_ref by oracle/Makefile when /root/reference exists)."""
    key = "ref_asshipped" if asshipped else "ref"
    if key not in _cache:
        name = "libsift_ref_asshipped.so" if asshipped else "libsift_ref.so"
        l = _Lib(os.path.join(HERE, "_ref", name), "ref")
        l.lib.ref_match_public.restype = C.c_int
        l.lib.ref_match_public.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        l.lib.ref_detect_public.restype = C.c_int
        l.lib.ref_detect_public.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        l.lib.ref_load_image.restype = C.c_int
        l.lib.ref_load_image.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(C.c_int), C.c_void_p, C.c_long]
        l.match_fn = l.lib.ref_match_public
        _cache[key] = l
    return _cache[key]


def best():
    """The strongest oracle available: the real reference if prebuilt, else the pinned port."""
    return ref() if have_ref() else port()


def match(lib, desc_a, desc_b, ratio=0.75):
    a = np.ascontiguousarray(desc_a, dtype=np.uint8).reshape(-1, 128)
    b = np.ascontiguousarray(desc_b, dtype=np.uint8).reshape(-1, 128)
    cap = max(len(a), 1)
    ia = np.zeros(cap, np.int32)
    ib = np.zeros(cap, np.int32)
    d = np.zeros(cap, np.float64)
    n = lib.match_fn(a.ctypes.data, len(a), b.ctypes.data, len(b), ratio, ia.ctypes.data,
                     ib.ctypes.data, d.ctypes.data, cap)
    return ia[:n].copy(), ib[:n].copy(), d[:n].copy()


def ref_load_image(path):
    """Decode with the reference's own vendored stb (image_io.cpp:20-35) -> float64 HxWxC."""
    l = ref()
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    rc = l.lib.ref_load_image(path.encode(), C.byref(w), C.byref(h), C.byref(c), None, 0)
    if rc != 0:
        raise RuntimeError(f"stb failed to load {path}")
    out = np.zeros((h.value, w.value, c.value))
    l.lib.ref_load_image(path.encode(), C.byref(w), C.byref(h), C.byref(c), out.ctypes.data, out.size)
    return out if c.value > 1 else out[:, :, 0]


def ref_detect_public(image):
    """The reference's public entry point, untouched (writes ./keypoints.png like the original)."""
    l = ref()
    img = np.ascontiguousarray(image, dtype=np.float64)
    h, w = img.shape[:2]
    c = 1 if img.ndim == 2 else img.shape[2]
    cap = max(100000, (w * h) // 20)
    out = np.zeros(cap, dtype=KP_DTYPE)
    n = l.lib.ref_detect_public(img.ctypes.data, w, h, c, out.ctypes.data, cap)
    if n < 0 or n > cap:
        raise RuntimeError(f"reference detect failed or overflowed ({n})")
    return out[:n].copy()
