"""ctypes binding of include/sift_b200.h.

Mirrors the reference's public API (src/sift.hh:65-75):
  detect_keypoints_and_descriptors(image, double_image_size=True, init_sigma=1.6, intervals=3,
      window_size=3, contrast_threshold=0.04, eigen_ratio=10, num_bins=36, peak_ratio=0.8,
      ori_sigma_factor=1.5, desc_scale_factor=3.0) -> records with the reference's Keypoint layout
  match_keypoints(kps1, kps2, ratio_threshold=0.75) -> (idx1, idx2, distance)
Errors surface as SiftError (the C++ shim rethrows std::runtime_error the same way).
"""
import ctypes as C
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

# byte-for-byte the reference's struct Keypoint (sift.hh:15-23)
KP_DTYPE = np.dtype(
    [("x", "<f8"), ("y", "<f8"), ("octave", "<i4"), ("layer", "<i4"), ("size", "<f8"),
     ("pori", "<f8"), ("desc", "u1", (128,))]
)
assert KP_DTYPE.itemsize == 168

STATUS = {0: "OK", 1: "E_INVALID", 2: "E_NO_DEVICE", 3: "E_CUDA", 4: "E_CAPACITY", 5: "E_UNSUPPORTED",
          6: "E_TOO_LARGE"}
E_CAPACITY = 4


class SiftError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"sift_b200: {STATUS.get(code, code)}: {text}")
        self.code = code


class SiftParams(C.Structure):
    """sift_b200_params; defaults are the reference's (sift.hh:65-71)."""
    _fields_ = [
        ("double_image_size", C.c_int32), ("init_sigma", C.c_double), ("intervals", C.c_int32),
        ("window_size", C.c_int32), ("contrast_threshold", C.c_double), ("eigen_ratio", C.c_double),
        ("num_bins", C.c_double), ("peak_ratio", C.c_double), ("ori_sigma_factor", C.c_double),
        ("desc_scale_factor", C.c_double), ("max_octaves", C.c_int32),
    ]


class _Stats(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("octaves", "extrema", "raw_keypoints", "oriented_keypoints",
                                          "final_keypoints", "base_width", "base_height")]


def library_path():
    # SIFT_B200_LIB: another build of the same library (kernel experiments); never a different implementation
    return os.environ.get("SIFT_B200_LIB") or os.path.join(HERE, "libsift_b200.so")


def declared_symbols():
    """Every function include/sift_b200.h declares (used by the load/export test)."""
    text = open(os.path.join(ROOT, "include", "sift_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sift_b200_[a-z0-9_]+)\s*\(", text)))


_lib = None


def load_library():
    """Loads libsift_b200.so; raises if it has not been built (there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise SiftError(2, f"{path} is missing -- build it with `make -C sift_project_b200/csrc` "
                           "(or __graft_entry__.build()); this package has no CPU fallback")
    L = C.CDLL(path)
    vp, i32, i32p = C.c_void_p, C.c_int, C.POINTER(C.c_int)
    PP = C.POINTER(SiftParams)
    sig = {
        "sift_b200_default_params": (None, [PP]),
        "sift_b200_create": (i32, [i32, i32, i32, C.POINTER(vp)]),
        "sift_b200_destroy": (None, [vp]),
        "sift_b200_last_error": (C.c_char_p, [vp]),
        "sift_b200_detect_u8": (i32, [vp, vp, i32, i32, i32, PP, vp, i32, i32p]),
        "sift_b200_detect_f32": (i32, [vp, vp, i32, i32, i32, PP, vp, i32, i32p]),
        "sift_b200_detect_enqueue_u8": (i32, [vp, vp, i32, i32, i32, PP]),
        "sift_b200_detect_finish": (i32, [vp, i32p]),
        "sift_b200_result_device": (i32, [vp, C.POINTER(vp), C.POINTER(vp), i32p]),
        "sift_b200_get_stats": (i32, [vp, C.POINTER(_Stats)]),
        "sift_b200_match": (i32, [vp, vp, i32, vp, i32, C.c_double, vp, vp, vp, i32, i32p]),
        "sift_b200_match_enqueue": (i32, [vp, vp, i32, vp, i32, vp, vp, vp]),
        "sift_b200_sync": (i32, [vp]),
        "sift_b200_stream": (vp, [vp]),
        "sift_b200_match_path": (i32, [i32, i32]),
        "sift_b200_debug_options": (i32, [vp, i32, i32]),
        "sift_b200_debug_plane_dims": (i32, [vp, i32, i32p, i32p]),
        "sift_b200_debug_plane": (i32, [vp, i32, i32, i32, vp]),
        "sift_b200_debug_extrema": (i32, [vp, vp, i32, i32p]),
        "sift_b200_debug_keypoints": (i32, [vp, i32, vp, i32, i32p]),
        "sift_b200_launch_count": (C.c_long, [vp]),
        "sift_b200_graphs_built": (C.c_long, [vp]),
        "sift_b200_host_alloc": (i32, [C.c_size_t, C.POINTER(vp)]),
        "sift_b200_host_free": (None, [vp]),
        "sift_b200_detect_batch_u8": (i32, [vp, i32, vp, i32, i32, i32, i32, PP, vp, vp, vp]),
        "sift_b200_comm_unique_id": (i32, [vp]),
        "sift_b200_comm_attach": (i32, [vp, vp, i32, i32]),
        "sift_b200_comm_attach_all": (i32, [vp, i32]),
        "sift_b200_comm_info": (i32, [vp, i32p, i32p]),
        "sift_b200_collection_match": (i32, [vp, i32, vp, vp, i32, i32p]),
        "sift_b200_collection_match_all": (i32, [vp, i32, i32, vp, vp, i32]),
        "sift_b200_collection_pairs": (i32, [vp, vp, vp, vp, i32, i32p]),
        "sift_b200_collection_fetch": (i32, [vp, i32, C.c_double, vp, vp, vp, i32, i32p]),
        "sift_b200_collection_device": (i32, [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), i32p]),
        "sift_b200_collection_digest": (i32, [vp, C.c_double, i32, C.POINTER(C.c_int64), C.POINTER(C.c_uint64)]),
        "sift_b200_partition_pairs": (i32, [vp, i32, i32, i32, i32, vp, vp, i32, i32p]),
        "sift_b200_debug_orient": (i32, [vp, vp, i32, vp, i32, i32p]),
        "sift_b200_debug_describe": (i32, [vp, vp, i32]),
        "sift_b200_debug_launch_plan": (i32, [vp, i32, i32, i32]),
        "sift_b200_debug_tail": (i32, [vp, i32]),
        "sift_b200_debug_canary_arm": (i32, [vp]),
        "sift_b200_debug_canary_check": (i32, [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
        "sift_b200_result_copy": (i32, [vp, vp, i32, i32p]),
        "sift_b200_set_profiling": (i32, [vp, i32]),
        "sift_b200_get_profile": (i32, [vp, vp, vp]),
        "sift_b200_get_profile_marks": (i32, [vp, vp, vp, i32, i32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def make_params(double_image_size=True, init_sigma=1.6, intervals=3, window_size=3,
                contrast_threshold=0.04, eigen_ratio=10.0, num_bins=36, peak_ratio=0.8,
                ori_sigma_factor=1.5, desc_scale_factor=3.0, max_octaves=0):
    return SiftParams(int(bool(double_image_size)), init_sigma, intervals, window_size, contrast_threshold,
                      eigen_ratio, num_bins, peak_ratio, ori_sigma_factor, desc_scale_factor, max_octaves)


def _ptr(x):
    """Device pointer of a torch CUDA tensor, host pointer of a numpy array, or a raw int."""
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(type(x))


class SiftContext:
    """One GPU + one stream + one workspace (sift_b200_create)."""

    def __init__(self, max_width, max_height, device=0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.sift_b200_create(device, int(max_width), int(max_height), C.byref(h))
        if rc:
            raise SiftError(rc, self._L.sift_b200_last_error(None).decode())
        self._h = h
        self.device = device
        self.max_width, self.max_height = max_width, max_height

    def close(self):
        if getattr(self, "_h", None):
            self._L.sift_b200_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc:
            raise SiftError(rc, self._L.sift_b200_last_error(self._h).decode())

    def _after_torch(self, *tensors):
        """Stream contract of the C ABI (sift_b200.h): device pointers are consumed in the CONTEXT's stream order.
        A torch CUDA tensor was produced on torch's current stream, so the context's stream is made to wait for
        an event recorded there before any call that reads it (no host synchronisation)."""
        if not any(hasattr(t, "is_cuda") and t.is_cuda for t in tensors):
            return
        import torch
        dev = next(t.device for t in tensors if hasattr(t, "is_cuda") and t.is_cuda)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        torch.cuda.ExternalStream(self.stream, device=dev).wait_event(ev)

    # ---- detect_keypoints_and_descriptors ----
    def detect(self, image, capacity=None, **params):
        """image: HxW or HxWx3, uint8 or float (0..255), numpy (host) or torch CUDA tensor."""
        p = make_params(**params)
        is_np = isinstance(image, np.ndarray)
        shape = tuple(image.shape)
        h, w = shape[0], shape[1]
        ch = 1 if len(shape) == 2 else shape[2]
        if is_np:
            if image.dtype == np.uint8:
                arr, fn = np.ascontiguousarray(image), self._L.sift_b200_detect_u8
            else:
                arr, fn = np.ascontiguousarray(image, dtype=np.float32), self._L.sift_b200_detect_f32
        else:
            import torch
            arr = image.contiguous()
            if arr.dtype == torch.uint8:
                fn = self._L.sift_b200_detect_u8
            else:
                arr, fn = arr.float(), self._L.sift_b200_detect_f32
            self._after_torch(arr)
        cap = capacity if capacity is not None else max(4096, (w * h) // 16)
        while True:
            out = np.zeros(cap, dtype=KP_DTYPE)
            n = C.c_int(0)
            rc = fn(self._h, _ptr(arr), w, h, ch, C.byref(p), out.ctypes.data, cap, C.byref(n))
            if rc == E_CAPACITY and capacity is None and n.value > cap:
                cap = n.value
                continue
            self._check(rc)
            return out[: n.value].copy()

    def detect_enqueue(self, image_u8, width, height, channels=1, **params):
        """image_u8: torch CUDA tensor, (pinned) host array / tensor, or a raw pointer."""
        p = make_params(**params)
        self._after_torch(image_u8)
        self._check(self._L.sift_b200_detect_enqueue_u8(self._h, _ptr(image_u8), width, height, channels,
                                                        C.byref(p)))

    def result_copy(self, out):
        """Wait and copy the last detect's records into `out` (numpy KP_DTYPE array); returns count."""
        n = C.c_int(0)
        self._check(self._L.sift_b200_result_copy(self._h, out.ctypes.data, len(out), C.byref(n)))
        return n.value

    STAGES = ("input", "pyramid", "extrema", "refine", "orient", "sort", "describe")

    def set_profiling(self, on):
        self._check(self._L.sift_b200_set_profiling(self._h, int(on)))

    def profile(self):
        ms = np.zeros(len(self.STAGES), np.float32)
        nl = np.zeros(len(self.STAGES), np.int32)
        self._check(self._L.sift_b200_get_profile(self._h, ms.ctypes.data, nl.ctypes.data))
        return dict(zip(self.STAGES, ms.tolist())), dict(zip(self.STAGES, nl.tolist()))

    def profile_marks(self):
        """[(stage name, ms)] per bracketed group of launches of the last profiled detect, in issue order."""
        n = C.c_int(0)
        self._check(self._L.sift_b200_get_profile_marks(self._h, None, None, 0, C.byref(n)))
        st = np.zeros(n.value, np.int32)
        ms = np.zeros(n.value, np.float32)
        self._check(self._L.sift_b200_get_profile_marks(self._h, st.ctypes.data, ms.ctypes.data, n.value, C.byref(n)))
        return [(self.STAGES[s] if 0 <= s < len(self.STAGES) else "other", float(t)) for s, t in zip(st, ms)]

    def detect_finish(self):
        n = C.c_int(0)
        self._check(self._L.sift_b200_detect_finish(self._h, C.byref(n)))
        return n.value

    def result_device(self):
        rec, desc, n = C.c_void_p(), C.c_void_p(), C.c_int(0)
        self._check(self._L.sift_b200_result_device(self._h, C.byref(rec), C.byref(desc), C.byref(n)))
        return rec.value, desc.value, n.value

    def stats(self):
        s = _Stats()
        self._check(self._L.sift_b200_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in _Stats._fields_}

    # ---- match_keypoints ----
    def match(self, desc_a, desc_b, ratio_threshold=0.75):
        """desc_*: n x 128 uint8 (numpy host arrays, torch CUDA tensors) -> (idx_a, idx_b, dist)."""
        na, nb = int(desc_a.shape[0]), int(desc_b.shape[0])
        if isinstance(desc_a, np.ndarray):
            desc_a = np.ascontiguousarray(desc_a, dtype=np.uint8)
        if isinstance(desc_b, np.ndarray):
            desc_b = np.ascontiguousarray(desc_b, dtype=np.uint8)
        cap = max(na, 1)
        ia, ib, d = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.float64)
        n = C.c_int(0)
        self._after_torch(desc_a, desc_b)
        self._check(self._L.sift_b200_match(self._h, _ptr(desc_a) if na else None, na,
                                            _ptr(desc_b) if nb else None, nb, float(ratio_threshold),
                                            ia.ctypes.data, ib.ctypes.data, d.ctypes.data, cap, C.byref(n)))
        return ia[: n.value].copy(), ib[: n.value].copy(), d[: n.value].copy()

    def match_enqueue(self, d_a, na, d_b, nb, d_best_idx, d_best_d2, d_second_d2):
        self._after_torch(d_a, d_b, d_best_idx, d_best_d2, d_second_d2)
        self._check(self._L.sift_b200_match_enqueue(self._h, _ptr(d_a), na, _ptr(d_b), nb, _ptr(d_best_idx),
                                                    _ptr(d_best_d2), _ptr(d_second_d2)))

    def sync(self):
        self._check(self._L.sift_b200_sync(self._h))

    @property
    def stream(self):
        return self._L.sift_b200_stream(self._h)

    @property
    def launches(self):
        return self._L.sift_b200_launch_count(self._h)

    def match_path(self, na, nb):
        return self._L.sift_b200_match_path(na, nb)

    # ---- multi-GPU: NCCL communicator + collection matching behind the C ABI ----
    def comm_attach(self, unique_id, world, rank):
        """Collective: every rank passes the bytes rank 0 got from comm_unique_id()."""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id)) if unique_id is not None else None
        self._check(self._L.sift_b200_comm_attach(self._h, buf, world, rank))

    def comm_info(self):
        w, r = C.c_int(1), C.c_int(0)
        self._check(self._L.sift_b200_comm_info(self._h, C.byref(w), C.byref(r)))
        return w.value, r.value

    def collection_match(self, n_images, local_descs, both_directions=False):
        """local_descs: the (n_i, 128) uint8 matrices (numpy / torch CUDA) of the images this rank owns
        (image % world == rank), ascending image index.  Enqueues the exchange and every owned pair; returns the
        number of owned pairs."""
        keep = [np.ascontiguousarray(d, dtype=np.uint8) if isinstance(d, np.ndarray) else d.contiguous() for d in local_descs]
        ptrs = (C.c_void_p * max(len(keep), 1))(*[(_ptr(d) if d.shape[0] else None) for d in keep])
        counts = (C.c_int32 * max(len(keep), 1))(*[int(d.shape[0]) for d in keep])
        n = C.c_int(0)
        self._after_torch(*keep)
        self._check(self._L.sift_b200_collection_match(self._h, n_images, ptrs, counts, int(both_directions), C.byref(n)))
        self._coll_keep = keep   # host sources must outlive the asynchronous copies
        return n.value

    def collection_pairs(self):
        n = C.c_int(0)
        self._check(self._L.sift_b200_collection_pairs(self._h, None, None, None, 0, C.byref(n)))
        pi, pj, rows = (np.zeros(max(n.value, 1), np.int32) for _ in range(3))
        self._check(self._L.sift_b200_collection_pairs(self._h, pi.ctypes.data, pj.ctypes.data, rows.ctypes.data,
                                                       n.value, C.byref(n)))
        return pi[: n.value], pj[: n.value], rows[: n.value]

    def collection_fetch(self, pair, rows, ratio_threshold=0.75):
        cap = max(int(rows), 1)
        ia, ib, d = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.float64)
        n = C.c_int(0)
        self._check(self._L.sift_b200_collection_fetch(self._h, pair, float(ratio_threshold), ia.ctypes.data,
                                                       ib.ctypes.data, d.ctypes.data, cap, C.byref(n)))
        return ia[: n.value].copy(), ib[: n.value].copy(), d[: n.value].copy()

    def collection_digest(self, ratio_threshold=0.75, all_ranks=True):
        """(matches, checksum) of this rank's pairs, summed over all ranks when all_ranks (collective)."""
        m, h = C.c_int64(0), C.c_uint64(0)
        self._check(self._L.sift_b200_collection_digest(self._h, float(ratio_threshold), int(all_ranks),
                                                        C.byref(m), C.byref(h)))
        return m.value, h.value

    # ---- introspection for the parity tests ----
    def debug_options(self, keep_all_planes=False, unfused_pyramid=False):
        self._check(self._L.sift_b200_debug_options(self._h, int(keep_all_planes), int(unfused_pyramid)))

    def plane_dims(self, octave):
        w, h = C.c_int(), C.c_int()
        self._check(self._L.sift_b200_debug_plane_dims(self._h, octave, C.byref(w), C.byref(h)))
        return w.value, h.value

    def gaussian(self, octave, layer):
        w, h = self.plane_dims(octave)
        out = np.zeros((h, w), np.float32)
        self._check(self._L.sift_b200_debug_plane(self._h, 0, octave, layer, out.ctypes.data))
        return out

    def dog(self, octave, layer):
        w, h = self.plane_dims(octave)
        out = np.zeros((h, w), np.float32)
        self._check(self._L.sift_b200_debug_plane(self._h, 1, octave, layer, out.ctypes.data))
        return out

    def extrema(self):
        n = C.c_int(0)
        self._check(self._L.sift_b200_debug_extrema(self._h, None, 0, C.byref(n)))
        out = np.zeros((max(n.value, 1), 4), np.int32)
        self._check(self._L.sift_b200_debug_extrema(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out[: n.value]

    def launch_plan(self, use_graph=None, centred=None, extrema_form=None):
        self._check(self._L.sift_b200_debug_launch_plan(self._h, -1 if use_graph is None else int(use_graph),
                                                        -1 if centred is None else int(centred),
                                                        -1 if extrema_form is None else int(extrema_form)))

    def tail_kernel(self, one_launch=True):
        """The small octaves in one cascade launch + one extrema launch (default) or octave by octave."""
        self._check(self._L.sift_b200_debug_tail(self._h, int(bool(one_launch))))

    def canary_arm(self):
        self._check(self._L.sift_b200_debug_canary_arm(self._h))

    def canary_check(self):
        """(stray writes, missing writes) of the scale-space kernels since canary_arm(); both must be 0."""
        a, b = C.c_int64(-1), C.c_int64(-1)
        self._check(self._L.sift_b200_debug_canary_check(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def graphs_built(self):
        return self._L.sift_b200_graphs_built(self._h)

    def orient_given(self, raw):
        """compute_orientations on caller-supplied raw keypoints over the last detect's scale space."""
        raw = np.ascontiguousarray(raw, dtype=KP_DTYPE)
        cap = 4 * len(raw) + 16
        out = np.zeros(cap, dtype=KP_DTYPE)
        n = C.c_int(0)
        self._check(self._L.sift_b200_debug_orient(self._h, raw.ctypes.data, len(raw), out.ctypes.data, cap, C.byref(n)))
        return out[: n.value].copy()

    def describe_given(self, kps):
        """compute_descriptors on caller-supplied oriented keypoints; returns a copy with .desc filled."""
        out = np.ascontiguousarray(kps, dtype=KP_DTYPE).copy()
        self._check(self._L.sift_b200_debug_describe(self._h, out.ctypes.data, len(out)))
        return out

    def stage_keypoints(self, stage):
        n = C.c_int(0)
        self._check(self._L.sift_b200_debug_keypoints(self._h, stage, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=KP_DTYPE)
        self._check(self._L.sift_b200_debug_keypoints(self._h, stage, out.ctypes.data, n.value, C.byref(n)))
        return out[: n.value]


def pinned_array(shape, dtype=np.uint8):
    """numpy array in page-locked host memory (sift_b200_host_alloc), freed when the last view goes away."""
    import weakref
    L = load_library()
    count = int(np.prod(shape))
    n = max(count * np.dtype(dtype).itemsize, 1)
    p = C.c_void_p()
    rc = L.sift_b200_host_alloc(n, C.byref(p))
    if rc:
        raise SiftError(rc, L.sift_b200_last_error(None).decode())
    buf = (C.c_uint8 * n).from_address(p.value)
    weakref.finalize(buf, L.sift_b200_host_free, p.value)
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


def comm_unique_id():
    """NCCL unique id (bytes) for SiftContext.comm_attach; call on rank 0 and ship to every rank."""
    L = load_library()
    buf = (C.c_uint8 * 128)()
    rc = L.sift_b200_comm_unique_id(buf)
    if rc:
        raise SiftError(rc, L.sift_b200_last_error(None).decode())
    return bytes(buf)


def comm_attach_all(ctxs):
    """One process driving several GPUs: contexts become ranks 0..n-1 of one communicator."""
    L = load_library()
    arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    rc = L.sift_b200_comm_attach_all(arr, len(ctxs))
    if rc:
        raise SiftError(rc, L.sift_b200_last_error(ctxs[0]._h).decode())


def collection_match_all(ctxs, descs, both_directions=False):
    """descs[i]: (n_i, 128) uint8 of image i, on the GPU of ctxs[i % n] or on the host."""
    L = load_library()
    arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    keep = [np.ascontiguousarray(d, dtype=np.uint8) if isinstance(d, np.ndarray) else d.contiguous() for d in descs]
    ptrs = (C.c_void_p * max(len(keep), 1))(*[(_ptr(d) if d.shape[0] else None) for d in keep])
    counts = (C.c_int32 * max(len(keep), 1))(*[int(d.shape[0]) for d in keep])
    rc = L.sift_b200_collection_match_all(arr, len(ctxs), len(keep), ptrs, counts, int(both_directions))
    if rc:
        raise SiftError(rc, L.sift_b200_last_error(ctxs[0]._h).decode())
    for c in ctxs:
        c._coll_keep = keep


def partition_pairs_native(counts, world, rank, both_directions=False):
    """sift_b200_partition_pairs: the C++ deal of image pairs to ranks (pure host code)."""
    L = load_library()
    cnt = np.ascontiguousarray(counts, dtype=np.int32)
    n = C.c_int(0)
    L.sift_b200_partition_pairs(cnt.ctypes.data, len(cnt), world, int(both_directions), rank, None, None, 0, C.byref(n))
    pi, pj = np.zeros(max(n.value, 1), np.int32), np.zeros(max(n.value, 1), np.int32)
    rc = L.sift_b200_partition_pairs(cnt.ctypes.data, len(cnt), world, int(both_directions), rank, pi.ctypes.data,
                                     pj.ctypes.data, n.value, C.byref(n))
    if rc:
        raise SiftError(rc, "partition_pairs")
    return list(zip(pi[: n.value].tolist(), pj[: n.value].tolist()))


def detect_batch(ctxs, images, capacity=None, **params):
    """sift_b200_detect_batch_u8: image k on ctxs[k % len(ctxs)] (several contexts per GPU and / or several GPUs).
    images: equally sized uint8 arrays (numpy host, ideally pinned, or torch CUDA tensors on the right GPU)."""
    L = load_library()
    p = make_params(**params)
    n = len(images)
    shape = tuple(images[0].shape)
    h, w = shape[0], shape[1]
    ch = 1 if len(shape) == 2 else shape[2]
    keep = [np.ascontiguousarray(im) if isinstance(im, np.ndarray) else im.contiguous() for im in images]
    cap = capacity if capacity is not None else max(4096, (w * h) // 16)
    outs = [np.zeros(cap, dtype=KP_DTYPE) for _ in range(n)]
    arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    img_p = (C.c_void_p * max(n, 1))(*[_ptr(k) for k in keep])
    out_p = (C.c_void_p * max(n, 1))(*[o.ctypes.data for o in outs])
    caps = (C.c_int32 * max(n, 1))(*([cap] * n))
    counts = (C.c_int32 * max(n, 1))()
    rc = L.sift_b200_detect_batch_u8(arr, len(ctxs), img_p, n, w, h, ch, C.byref(p), out_p, caps, counts)
    if rc:
        raise SiftError(rc, L.sift_b200_last_error(ctxs[0]._h).decode())
    return [o[: counts[k]].copy() for k, o in enumerate(outs)]


def detect_keypoints_and_descriptors(image, double_image_size=True, init_sigma=1.6, intervals=3,
                                     window_size=3, contrast_threshold=0.04, eigen_ratio=10.0, num_bins=36,
                                     peak_ratio=0.8, ori_sigma_factor=1.5, desc_scale_factor=3.0, device=0):
    """One-shot form of sift.hh:65-71 (creates and destroys a context)."""
    h, w = image.shape[0], image.shape[1]
    with SiftContext(w, h, device) as ctx:
        return ctx.detect(image, double_image_size=double_image_size, init_sigma=init_sigma,
                          intervals=intervals, window_size=window_size,
                          contrast_threshold=contrast_threshold, eigen_ratio=eigen_ratio, num_bins=num_bins,
                          peak_ratio=peak_ratio, ori_sigma_factor=ori_sigma_factor,
                          desc_scale_factor=desc_scale_factor)


def match_keypoints(keypoints1, keypoints2, ratio_threshold=0.75, device=0):
    """One-shot form of sift.hh:73-75 on keypoint records; returns (idx1, idx2, distance)."""
    with SiftContext(64, 64, device) as ctx:
        return ctx.match(np.ascontiguousarray(keypoints1["desc"]), np.ascontiguousarray(keypoints2["desc"]),
                         ratio_threshold)
