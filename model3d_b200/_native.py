"""ctypes binding of libm3dgpu.so (the C ABI declared in include/m3d.h).

The library is the product: hand-written sm_100a kernels behind a C ABI.  There is
no Python or CPU fallback -- if the shared library is missing or no CUDA device is
present, calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# M3D_LIB: another build of the same library (kernel tuning runs, scripts/build_variant.sh)
LIB_PATH = os.environ.get("M3D_LIB") or os.path.join(_HERE, "libm3dgpu.so")

M3D_OK = 0
ERR_NAMES = {1: "INVALID_ARG", 2: "UNSUPPORTED", 3: "CUDA", 4: "NCCL", 5: "OOM"}

MESH_BUILD_HOST_SAH, MESH_BUILD_DEVICE_LBVH, MESH_BUILD_DEVICE_COLLAPSE = 0, 1, 2
TRACE_COUNTERS = 1
TRACE_NO_REFINE = 2
TRACE_SHARED_ORIGIN = 4

MAT_LAMBERT, MAT_PHONG, MAT_REFRACT, MAT_JOINED = 0, 1, 2, 3
MAT_NO_FLUX_CORRECTION, MAT_CHECKER, MAT_Z_GRADIENT = 1, 2, 4
OBJ_FLIP_NORMAL = 1
FOCUS_PHONG, FOCUS_SPHERE = 0, 1


class M3DError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("m3d error %s: %s" % (ERR_NAMES.get(code, code), msg))
        self.code = code


class Stats(C.Structure):
    _fields_ = [("rays", C.c_int64), ("hits", C.c_int64), ("nodes_visited", C.c_int64),
                ("tris_tested", C.c_int64), ("kernel_ms", C.c_double), ("h2d_ms", C.c_double),
                ("d2h_ms", C.c_double), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("launches", C.c_int64), ("samples", C.c_int64)]


class MeshInfo(C.Structure):
    _fields_ = [("num_triangles", C.c_int64), ("num_nodes", C.c_int64), ("node_bytes", C.c_int64),
                ("tri_bytes", C.c_int64), ("device_bytes", C.c_int64), ("max_depth", C.c_int32),
                ("build_ms", C.c_double), ("sah_cost", C.c_double)]


class Camera(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("screen_x", C.c_double * 3),
                ("screen_y", C.c_double * 3), ("field_of_view", C.c_double)]


class PointLight(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("color", C.c_double * 3),
                ("quad_dropoff", C.c_int32), ("_pad", C.c_int32)]


class MaterialDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("flags", C.c_uint32),
                ("diffuse", C.c_double * 3), ("specular", C.c_double * 3),
                ("emission", C.c_double * 3), ("ambient", C.c_double * 3),
                ("refract", C.c_double * 3), ("alpha", C.c_double),
                ("index_of_refraction", C.c_double), ("diffuse2", C.c_double * 3),
                ("proc_param", C.c_double), ("num_sub", C.c_int32),
                ("sub", C.c_int32 * 4), ("sub_prob", C.c_double * 4)]


class Transform(C.Structure):
    _fields_ = [("matrix", C.c_double * 9), ("offset", C.c_double * 3)]


PART_ATOMIC = 1
IPC_HANDLE_BYTES = 64


class Partition(C.Structure):
    _fields_ = [("row_begin", C.c_int32), ("row_end", C.c_int32), ("sample_begin", C.c_int64),
                ("flags", C.c_uint32), ("_pad", C.c_uint32)]


class FocusPoint(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("target", C.c_double * 3),
                ("alpha", C.c_double), ("radius", C.c_double),
                ("material_mask", C.c_uint64), ("prob", C.c_double)]


class PathParams(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("num_samples", C.c_int32),
                ("min_samples", C.c_int32), ("num_focus_points", C.c_int32),
                ("max_stddev", C.c_double), ("oversaturated_stddevs", C.c_double),
                ("cutoff", C.c_double), ("antialias", C.c_double), ("epsilon", C.c_double),
                ("focus", FocusPoint * 4), ("seed", C.c_uint64)]


class AreaLight(C.Structure):
    _fields_ = [("object", C.c_int32), ("_pad", C.c_int32), ("emission", C.c_double * 3)]


class BidirParams(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("max_light_depth", C.c_int32),
                ("min_depth", C.c_int32), ("num_samples", C.c_int32),
                ("roulette_delta", C.c_double), ("power_heuristic", C.c_double),
                ("cutoff", C.c_double), ("antialias", C.c_double), ("epsilon", C.c_double),
                ("seed", C.c_uint64), ("min_samples", C.c_int32), ("_pad", C.c_int32),
                ("max_stddev", C.c_double), ("oversaturated_stddevs", C.c_double)]


# every symbol include/m3d.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "m3d_abi_version", "m3d_last_error", "m3d_ctx_create", "m3d_ctx_destroy", "m3d_ctx_device",
    "m3d_ctx_synchronize", "m3d_ctx_trim", "m3d_mesh_create", "m3d_mesh_destroy", "m3d_mesh_get_info",
    "m3d_mesh_bounds", "m3d_mesh_first_ray_collisions", "m3d_mesh_first_ray_collisions_device",
    "m3d_mesh_ray_collision_counts", "m3d_mesh_ray_collisions", "m3d_mesh_contains", "m3d_mesh_sdf", "m3d_mesh_sphere_collisions",
    "m3d_scene_builder_create", "m3d_scene_builder_destroy", "m3d_scene_add_material",
    "m3d_scene_add_mesh", "m3d_scene_add_instance", "m3d_scene_add_sphere", "m3d_scene_add_rect", "m3d_scene_add_cylinder",
    "m3d_scene_build", "m3d_scene_destroy", "m3d_scene_bounds", "m3d_scene_get_info", "m3d_scene_cast",
    "m3d_render_raycast", "m3d_render_raycast_device", "m3d_render_raycast_views", "m3d_render_path", "m3d_render_path_device",
    "m3d_render_bidir", "m3d_render_bidir_device", "m3d_finalize_image_device",
    "m3d_measure_l2_bandwidth", "m3d_ctx_create_multi", "m3d_ctx_num_devices",
    "m3d_host_alloc", "m3d_host_free", "m3d_host_register", "m3d_host_unregister",
    "m3d_device_alloc", "m3d_device_free", "m3d_ipc_export", "m3d_ipc_open", "m3d_ipc_close",
]

_lib = None


def lib():
    """Load libm3dgpu.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C model3d_b200/csrc`. model3d_b200 has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.m3d_last_error.restype = C.c_char_p
        for name in SYMBOLS:
            fn = getattr(L, name)
            if name not in ("m3d_last_error", "m3d_ctx_destroy", "m3d_mesh_destroy",
                            "m3d_scene_destroy", "m3d_scene_builder_destroy"):
                fn.restype = C.c_int32
        for name in ("m3d_ctx_destroy", "m3d_mesh_destroy", "m3d_scene_destroy",
                     "m3d_scene_builder_destroy"):
            getattr(L, name).restype = None
        _lib = L
    return _lib


def check(rc):
    if rc != M3D_OK:
        raise M3DError(rc, lib().m3d_last_error().decode("utf-8", "replace"))


class Context:
    """m3d_ctx: one CUDA device."""

    def __init__(self, device=-1):
        self.h = C.c_void_p()
        check(lib().m3d_ctx_create(C.c_int32(device), C.byref(self.h)))

    def close(self):
        if self.h:
            lib().m3d_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def synchronize(self):
        check(lib().m3d_ctx_synchronize(self.h))

    def trim(self):
        """Release the scratch buffers the renderers keep between calls (m3d_ctx_trim)."""
        check(lib().m3d_ctx_trim(self.h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiContext(Context):
    """m3d_ctx over several devices of this node (m3d_ctx_create_multi): meshes and scenes built on
    it are replicated, first-hit batches are sliced and renders sharded over the devices, and the
    per-pixel sums meet in device 0's accumulator inside the flush kernels (NVLink red.add)."""

    def __init__(self, devices=None):
        self.h = C.c_void_p()
        if devices is None:
            check(lib().m3d_ctx_create_multi(None, C.c_int32(0), C.byref(self.h)))
        else:
            arr = (C.c_int32 * len(devices))(*[int(d) for d in devices])
            check(lib().m3d_ctx_create_multi(arr, C.c_int32(len(devices)), C.byref(self.h)))

    @property
    def num_devices(self):
        return int(lib().m3d_ctx_num_devices(self.h))


class _PinnedBlock:
    """Owner of one m3d_host_alloc block; numpy views keep it alive through .base."""

    def __init__(self, nbytes):
        self.nbytes = max(int(nbytes), 1)
        p = C.c_void_p()
        check(lib().m3d_host_alloc(C.c_int64(self.nbytes), C.byref(p)))
        self.ptr = p.value

    @property
    def __array_interface__(self):
        return {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 3}

    def __del__(self):
        try:
            if self.ptr:
                lib().m3d_host_free(C.c_void_p(self.ptr))
                self.ptr = None
        except Exception:
            pass


def host_empty(shape, dtype):
    """numpy array in pinned (page-locked) host memory from m3d_host_alloc: what the host-buffer
    calls need to reach the PCIe rate.  Freed when the array and all of its views are gone."""
    import numpy as np
    dt = np.dtype(dtype)
    count = int(np.prod(shape))
    block = _PinnedBlock(count * dt.itemsize)
    return np.asarray(block)[:count * dt.itemsize].view(dt).reshape(shape)


def device_alloc(ctx, nbytes):
    """Zero-filled device buffer owned by the library (a plain cudaMalloc, exportable over IPC)."""
    p = C.c_void_p()
    check(lib().m3d_device_alloc(ctx.h, C.c_int64(nbytes), C.byref(p)))
    return p.value


def device_free(ctx, ptr):
    check(lib().m3d_device_free(ctx.h, C.c_void_p(ptr)))


def ipc_export(ctx, ptr):
    h = (C.c_uint8 * IPC_HANDLE_BYTES)()
    check(lib().m3d_ipc_export(ctx.h, C.c_void_p(ptr), h))
    return bytes(h)


def ipc_open(ctx, handle):
    h = (C.c_uint8 * IPC_HANDLE_BYTES).from_buffer_copy(handle)
    p = C.c_void_p()
    check(lib().m3d_ipc_open(ctx.h, h, C.byref(p)))
    return p.value


def ipc_close(ctx, ptr):
    check(lib().m3d_ipc_close(ctx.h, C.c_void_p(ptr)))


_default_ctx = {}


def default_context(device=-1):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
