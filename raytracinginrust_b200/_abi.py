"""ctypes mirror of include/rtb200.h (the C ABI of the device library).

Only plain structures live here; they are shared by the product binding
(`raytracinginrust_b200`) and by the oracle's test wrapper (`oracle/oracle_py.py`),
which consumes the same RtSceneDesc.
"""
import ctypes as C

RT_NONE = 0xFFFFFFFF
ABI_VERSION = 1

# RtStatus
RT_OK, RT_ERR_BAD_ARGUMENT, RT_ERR_EMPTY_SCENE, RT_ERR_UNSUPPORTED, RT_ERR_CUDA, RT_ERR_NO_LIGHTS, RT_ERR_INTERNAL = range(7)
STATUS_NAMES = ["RT_OK", "RT_ERR_BAD_ARGUMENT", "RT_ERR_EMPTY_SCENE", "RT_ERR_UNSUPPORTED", "RT_ERR_CUDA",
                "RT_ERR_NO_LIGHTS", "RT_ERR_INTERNAL"]

# RtNodeKind
(NODE_SPHERE, NODE_MOVING_SPHERE, NODE_RECT, NODE_TRIANGLE, NODE_CUBE, NODE_LIST, NODE_BVH, NODE_TRANSLATE,
 NODE_ROTATE, NODE_FLIP, NODE_MEDIUM) = range(11)
PLANE_YZ, PLANE_XZ, PLANE_XY = 0, 1, 2
AXIS_X, AXIS_Y, AXIS_Z = 0, 1, 2
MAT_LAMBERTIAN, MAT_METAL, MAT_DIELECTRIC, MAT_DIFFUSE_LIGHT, MAT_ISOTROPIC, MAT_PBR = range(6)
TEX_CONSTANT, TEX_CHECKER, TEX_NOISE, TEX_IMAGE = range(4)
INTEGRATOR_HEAD, INTEGRATOR_LEGACY = 0, 1
FLAG_TRACE_ZERO_THROUGHPUT = 1
FLAG_WAVEFRONT = 2   # generate / extend / shade stages over HBM-resident ray queues
FLAG_MEGAKERNEL = 4  # one persistent kernel, path state in registers
CREATE_GPU_BVH = 1   # rt_scene_create_ex: build large trees on the GPU


class RtNode(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("material", C.c_uint32), ("child", C.c_uint32), ("count", C.c_uint32),
                ("axis", C.c_uint32), ("reserved", C.c_uint32), ("v", C.c_double * 10)]


class RtMaterial(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("texture", C.c_uint32), ("albedo", C.c_double * 3), ("fuzz", C.c_double),
                ("ir", C.c_double), ("pbr", C.c_double * 10)]


class RtTexture(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("a", C.c_uint32), ("b", C.c_uint32), ("reserved", C.c_uint32),
                ("color", C.c_double * 3), ("scale", C.c_double)]


class RtPerlin(C.Structure):
    _fields_ = [("ranvec", C.c_double * 768), ("perm_x", C.c_uint32 * 256), ("perm_y", C.c_uint32 * 256),
                ("perm_z", C.c_uint32 * 256)]


class RtImage(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("offset", C.c_uint64)]


class RtSceneDesc(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("world", C.c_uint32), ("lights", C.c_uint32), ("reserved", C.c_uint32),
                ("background", C.c_double * 3),
                ("nodes", C.POINTER(RtNode)), ("n_nodes", C.c_uint64),
                ("child_index", C.POINTER(C.c_uint32)), ("n_child_index", C.c_uint64),
                ("materials", C.POINTER(RtMaterial)), ("n_materials", C.c_uint64),
                ("textures", C.POINTER(RtTexture)), ("n_textures", C.c_uint64),
                ("perlin", C.POINTER(RtPerlin)), ("n_perlin", C.c_uint64),
                ("images", C.POINTER(RtImage)), ("n_images", C.c_uint64),
                ("texels", C.POINTER(C.c_uint8)), ("n_texel_bytes", C.c_uint64)]


class RtCamera(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("lower_left_corner", C.c_double * 3), ("horizontal", C.c_double * 3),
                ("vertical", C.c_double * 3), ("cu", C.c_double * 3), ("cv", C.c_double * 3),
                ("lens_radius", C.c_double), ("time0", C.c_double), ("time1", C.c_double)]


class RtRenderOpts(C.Structure):
    _fields_ = [("seed", C.c_uint32), ("integrator", C.c_uint32), ("sample_begin", C.c_uint32),
                ("sample_count", C.c_uint32), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class RtStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays", C.c_uint64), ("nonfinite_samples", C.c_uint64),
                ("render_ms", C.c_double), ("total_ms", C.c_double), ("kernel_launches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


class RtRay(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("direction", C.c_double * 3), ("time", C.c_double)]


class RtHit(C.Structure):
    _fields_ = [("node", C.c_int32), ("face", C.c_int32), ("material", C.c_int32), ("front_face", C.c_int32),
                ("t", C.c_double), ("position", C.c_double * 3), ("normal", C.c_double * 3), ("u", C.c_double),
                ("v", C.c_double)]


# numpy views of the two array-of-struct types the parity hooks exchange
RAY_DTYPE = [("origin", "<f8", 3), ("direction", "<f8", 3), ("time", "<f8")]
HIT_DTYPE = [("node", "<i4"), ("face", "<i4"), ("material", "<i4"), ("front_face", "<i4"), ("t", "<f8"),
             ("position", "<f8", 3), ("normal", "<f8", 3), ("u", "<f8"), ("v", "<f8")]

# every symbol include/rtb200.h declares (checked by tests/test_abi.py)
EXPORTS = ["rt_device_count", "rt_scene_create", "rt_scene_destroy", "rt_release_cached_memory", "rt_scene_device_bytes", "rt_render",
           "rt_render_device", "rt_render_wait", "rt_trace_first_hit", "rt_path_radiance", "rt_camera_rays",
           "rt_measure_fp64_peak", "rt_render_info", "rt_last_error", "rt_version",
           "rt_scene_group_create", "rt_scene_group_destroy", "rt_scene_group_size", "rt_scene_group_scene",
           "rt_render_multi", "rt_encode_rgb8", "rt_encode_ppm",
           "rt_compile", "rt_compiled_data", "rt_compiled_size", "rt_compiled_hash", "rt_compiled_destroy",
           "rt_scene_create_compiled", "rt_scene_create_ex"]
