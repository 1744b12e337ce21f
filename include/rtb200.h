/*
 * rtb200.h — C ABI of the B200-native path-tracing hot path for the
 * RayTracingInRust engine.
 *
 * This header is the drop-in boundary.  The reference has no FFI today: the
 * code this library replaces is the nested pixel/sample loop that is inlined in
 * main() (reference src/main.rs:767-834) together with ray_color
 * (src/main.rs:41-120) and everything ray_color calls.  A Rust host keeps scene
 * construction, the camera parameters and the PPM writer, serialises its
 * `Box<dyn Hittable>` graph into the plain-old-data RtSceneDesc below with a
 * `flatten()` visitor, and calls rt_render().  INTEGRATION.md shows the
 * `extern "C"` block and build.rs a maintainer would add.
 *
 * Every structure here is plain C: fixed-width integers, doubles, pointers and
 * counts.  No C++ types, no torch types, no CUDA types cross this boundary.
 * All arithmetic the reference performs is f64, so every real number in this
 * interface is a double; the only fp32 quantity is the accumulated image
 * (W*H*3 sums), which is what ncclReduce combines across GPUs.
 *
 * Error model (reference: panics, src/bvh.rs:28,55,61, src/hit.rs:95): every
 * entry point returns an RtStatus; nothing unwinds across the boundary;
 * rt_last_error() returns a thread-local message for the last failure.
 */
#ifndef RTB200_H
#define RTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTB200_ABI_VERSION 1u
#define RT_NONE 0xFFFFFFFFu

/* ------------------------------------------------------------------------ */
/* Status codes                                                              */
/* ------------------------------------------------------------------------ */
typedef enum RtStatus {
    RT_OK = 0,
    RT_ERR_BAD_ARGUMENT = 1,  /* null pointer, zero size, index out of range   */
    RT_ERR_EMPTY_SCENE = 2,   /* reference: panic "no object in the scene"     */
    RT_ERR_UNSUPPORTED = 3,   /* scene-graph shape the device compiler rejects */
    RT_ERR_CUDA = 4,          /* a CUDA runtime call failed / no usable device */
    RT_ERR_NO_LIGHTS = 5,     /* HEAD integrator with an empty light list
                                 (reference: unwrap() panic, src/hit.rs:94-96) */
    RT_ERR_INTERNAL = 6
} RtStatus;

/* ------------------------------------------------------------------------ */
/* Scene description: a serialised scene graph                               */
/* ------------------------------------------------------------------------ */

/* Node kinds mirror the reference's Hittable implementors one to one. */
typedef enum RtNodeKind {
    RT_NODE_SPHERE = 0,        /* src/sphere.rs:38-120   v = cx cy cz r                       */
    RT_NODE_MOVING_SPHERE = 1, /* src/sphere.rs:122-201  v = c0x c0y c0z c1x c1y c1z t0 t1 r  */
    RT_NODE_RECT = 2,          /* src/rect.rs:15-111     v = a0 a1 b0 b1 k ; axis = RtPlane   */
    RT_NODE_TRIANGLE = 3,      /* src/tri.rs:9-71        v = v0xyz v1xyz v2xyz                */
    RT_NODE_CUBE = 4,          /* src/cube.rs:7-46       v = minxyz maxxyz (six AARects)      */
    RT_NODE_LIST = 5,          /* src/hit.rs:47-97       child = first slot in child_index[]  */
    RT_NODE_BVH = 6,           /* src/bvh.rs:12-96       same as LIST; v = time0 time1        */
    RT_NODE_TRANSLATE = 7,     /* src/translate.rs:6-40  child ; v = offset xyz               */
    RT_NODE_ROTATE = 8,        /* src/rotate.rs:23-110   child ; axis = RtAxis ; v = degrees  */
    RT_NODE_FLIP = 9,          /* src/hit.rs:99-133      child (FlipNormal)                   */
    RT_NODE_MEDIUM = 10        /* src/medium.rs:10-65    child = boundary ; v = density ;
                                  material = the Isotropic phase function                    */
} RtNodeKind;

/* src/rect.rs:8-13,26-32: which axis is constant, and the (a,b) axes. */
typedef enum RtPlane { RT_PLANE_YZ = 0, RT_PLANE_XZ = 1, RT_PLANE_XY = 2 } RtPlane;
/* src/rotate.rs:8-21 */
typedef enum RtAxis { RT_AXIS_X = 0, RT_AXIS_Y = 1, RT_AXIS_Z = 2 } RtAxis;

typedef struct RtNode {
    uint32_t kind;     /* RtNodeKind                                             */
    uint32_t material; /* material index for primitives / cube / medium, else RT_NONE */
    uint32_t child;    /* wrapper kinds: child node index.  LIST/BVH: first slot in
                          RtSceneDesc.child_index                                 */
    uint32_t count;    /* LIST/BVH: number of children                            */
    uint32_t axis;     /* RECT: RtPlane.  ROTATE: RtAxis                          */
    uint32_t reserved;
    double v[10];      /* parameters, meaning per kind (see RtNodeKind)           */
} RtNode;

/* src/mat.rs:86-422. */
typedef enum RtMaterialKind {
    RT_MAT_LAMBERTIAN = 0,    /* texture = albedo                      */
    RT_MAT_METAL = 1,         /* albedo[3], fuzz                       */
    RT_MAT_DIELECTRIC = 2,    /* ir                                    */
    RT_MAT_DIFFUSE_LIGHT = 3, /* texture = emit                        */
    RT_MAT_ISOTROPIC = 4,     /* texture = albedo (legacy integrator)  */
    RT_MAT_PBR = 5            /* texture = base_color, pbr[10]: the Disney-style material of
                                 src/mat.rs:86-197, sampled through PDF::BRDF (src/pdf.rs:20-60,97-130,151-160) */
} RtMaterialKind;

/* Index of each PBR::new argument (src/mat.rs:101-115) in RtMaterial.pbr */
enum { RT_PBR_METALLIC = 0, RT_PBR_SUBSURFACE = 1, RT_PBR_SPECULAR = 2, RT_PBR_ROUGHNESS = 3,
       RT_PBR_SPECULAR_TINT = 4, RT_PBR_ANISOTROPIC = 5, RT_PBR_SHEEN = 6, RT_PBR_SHEEN_TINT = 7,
       RT_PBR_CLEARCOAT = 8, RT_PBR_CLEARCOAT_GLOSS = 9 };

typedef struct RtMaterial {
    uint32_t kind;
    uint32_t texture;
    double albedo[3];
    double fuzz;
    double ir;
    double pbr[10];
} RtMaterial;

/* src/texture.rs */
typedef enum RtTextureKind {
    RT_TEX_CONSTANT = 0, /* color[3]                                        */
    RT_TEX_CHECKER = 1,  /* a = odd texture, b = even texture               */
    RT_TEX_NOISE = 2,    /* a = index into perlin[], scale                  */
    RT_TEX_IMAGE = 3     /* a = index into images[]                         */
} RtTextureKind;

typedef struct RtTexture {
    uint32_t kind;
    uint32_t a;
    uint32_t b;
    uint32_t reserved;
    double color[3];
    double scale;
} RtTexture;

/* src/perlin.rs:60-75: 256 in-ball gradient vectors and three permutations. */
typedef struct RtPerlin {
    double ranvec[256 * 3];
    uint32_t perm_x[256];
    uint32_t perm_y[256];
    uint32_t perm_z[256];
} RtPerlin;

/* src/texture.rs:83-121: tightly packed RGB8, row 0 at the top. */
typedef struct RtImage {
    uint32_t width;
    uint32_t height;
    uint64_t offset; /* byte offset of texel (0,0) in RtSceneDesc.texels */
} RtImage;

typedef struct RtSceneDesc {
    uint32_t abi_version; /* RTB200_ABI_VERSION */
    uint32_t world;       /* root node of the world (src/main.rs:41 `world`)  */
    uint32_t lights;      /* root node of the light list (`lights`), a LIST; may be empty */
    uint32_t reserved;
    double background[3]; /* src/main.rs:118 */

    const RtNode *nodes;
    uint64_t n_nodes;
    const uint32_t *child_index; /* node indices, referenced by LIST/BVH nodes */
    uint64_t n_child_index;
    const RtMaterial *materials;
    uint64_t n_materials;
    const RtTexture *textures;
    uint64_t n_textures;
    const RtPerlin *perlin;
    uint64_t n_perlin;
    const RtImage *images;
    uint64_t n_images;
    const uint8_t *texels;
    uint64_t n_texel_bytes;
} RtSceneDesc;

/* The fields of the reference's Camera after Camera::new (src/camera.rs:5-49). */
typedef struct RtCamera {
    double origin[3];
    double lower_left_corner[3];
    double horizontal[3];
    double vertical[3];
    double cu[3];
    double cv[3];
    double lens_radius;
    double time0;
    double time1;
} RtCamera;

/* Which ray_color the sample loop runs. */
typedef enum RtIntegrator {
    RT_INTEGRATOR_HEAD = 0,  /* src/main.rs:86-110: scatter_mc_method + light/BSDF mixture */
    RT_INTEGRATOR_LEGACY = 1 /* src/main.rs:84-85 (commented): Material::scatter           */
} RtIntegrator;

typedef struct RtRenderOpts {
    uint32_t seed;         /* Philox stream seed; key = (pixel, sample), see DESIGN.md */
    uint32_t integrator;   /* RtIntegrator */
    uint32_t sample_begin; /* this call renders samples [sample_begin, sample_begin+sample_count) */
    uint32_t sample_count; /* 0 = all of spp.  Multi-GPU: disjoint ranges per GPU        */
    uint32_t flags;        /* RT_FLAG_* */
    uint32_t reserved;
} RtRenderOpts;

#define RT_FLAG_NONE 0u
/* Keep tracing a path whose throughput is exactly zero, as the reference does
 * (src/main.rs:97 evaluates the recursive call even when scattering_pdf == 0).
 * Off by default: the contribution is zero either way. */
#define RT_FLAG_TRACE_ZERO_THROUGHPUT 1u
/* Pipeline choice.  Both run the same device arithmetic and give bit-identical images:
 *   wavefront  - generate / extend / shade stages over HBM-resident ray queues
 *   megakernel - one persistent kernel, one path per lane, the path state in registers
 * With neither flag the library picks per scene (DESIGN.md "Kernels"). */
#define RT_FLAG_WAVEFRONT 2u
#define RT_FLAG_MEGAKERNEL 4u

typedef struct RtStats {
    uint64_t paths;             /* (pixel, sample) paths started                        */
    uint64_t rays;              /* path segments: world.hit calls from the integrator   */
    uint64_t nonfinite_samples; /* samples whose radiance had a NaN/Inf component       */
    double render_ms;           /* device time of the render kernels (CUDA events)      */
    double total_ms;            /* host wall time of the call                           */
    uint64_t kernel_launches;   /* kernels launched by this call                        */
    uint64_t h2d_bytes;         /* bytes copied host->device by this call               */
    uint64_t d2h_bytes;         /* bytes copied device->host by this call               */
} RtStats;

/* A ray (src/ray.rs:3-7).  Direction is never normalised. */
typedef struct RtRay {
    double origin[3];
    double direction[3];
    double time;
} RtRay;

/* What HitRecord (src/hit.rs:9-24) carries, plus the ids the reference lacks. */
typedef struct RtHit {
    int32_t node;        /* RtSceneDesc node index of the primitive / medium hit, -1 = miss */
    int32_t face;        /* CUBE: which of the six rects (src/cube.rs:17-25 order), else 0   */
    int32_t material;    /* material index                                                   */
    int32_t front_face;  /* 0/1                                                              */
    double t;
    double position[3];
    double normal[3];
    double u, v;
} RtHit;

/* ------------------------------------------------------------------------ */
/* Entry points                                                              */
/* ------------------------------------------------------------------------ */

typedef struct RtScene RtScene; /* opaque: device-resident compiled scene */

/* Number of usable CUDA devices (0 when there is none).  Never fails. */
int rt_device_count(void);

/* Compile `desc` (copied; the caller keeps ownership of its buffers) into the
 * device representation on CUDA device `device` and upload it once.
 * Replaces: the scene borrow held by the loop at src/main.rs:624,827. */
RtStatus rt_scene_create(const RtSceneDesc *desc, int device, RtScene **out_scene);
/* The same with options.  RT_CREATE_GPU_BVH: BVH::new (src/bvh.rs:18-73) runs ON THE GPU for every
 * tree of at least 4096 primitives - a linear BVH (Morton codes, radix sort, Karras hierarchy,
 * bottom-up refit: about a millisecond for the 394k triangles of config 5) instead of the host's
 * binned-SAH build (0.2-0.3 s).  Rays find the same hits on either tree, so images are
 * bit-identical; the linear tree is slower to traverse, so this is for callers that want the
 * first image sooner (previews, short renders), not the default.   */
enum { RT_CREATE_GPU_BVH = 1 };
RtStatus rt_scene_create_ex(const RtSceneDesc *desc, int device, uint32_t create_flags, RtScene **out_scene);
void rt_scene_destroy(RtScene *scene);
/* Device blocks (sample planes, the wavefront pool, the tables, the scratch of rt_encode_*) and the
 * few pinned host blocks of a scene are not returned to the driver when their owner is destroyed but
 * parked for the next request of the same device and size class, up to RTB200_SCRATCH_CACHE_MB
 * (default 24576, 0 = off) per process: a host that renders scene after scene pays cudaMalloc /
 * cudaFree once.  This returns the parked blocks to the driver. */
void rt_release_cached_memory(void);

/* Bytes of device memory the compiled scene occupies (h2d traffic of create). */
uint64_t rt_scene_device_bytes(const RtScene *scene);

/* render(world, camera, width, height, spp, max_depth) -> pixels.
 * Replaces src/main.rs:772-834.  Writes W*H*3 fp32 radiance SUMS (not means)
 * over the rendered sample range into caller-owned host memory, rows top-down
 * (first row is j = H-1, src/main.rs:772), columns left to right, RGB
 * interleaved.  The host divides by spp, applies format_color
 * (src/vec.rs:125-131) and writes the PPM.  out_rgb_sum may be NULL: the image then
 * only stays resident on the device, for rt_encode_rgb8 / rt_encode_ppm.
 * RT_ERR_UNSUPPORTED: the scene has a MovingSphere with (time0, time1) != (0, 1) - its
 * centre extrapolates, src/sphere.rs:144-146 - and the camera's shutter leaves [0, 1],
 * the range the culling bounds were built for (every camera of src/main.rs is 0..1). */
RtStatus rt_render(const RtScene *scene, const RtCamera *camera, uint32_t width, uint32_t height,
                   uint32_t spp, uint32_t max_depth, const RtRenderOpts *opts,
                   float *out_rgb_sum, RtStats *stats);

/* Same, but `out_rgb_sum_device` is device memory on the scene's device and the
 * work is enqueued on `cuda_stream` (a cudaStream_t passed as void*; NULL = the
 * default stream).  Returns after enqueueing; stats (if non-NULL) are complete
 * after rt_render_wait().  This is the entry the multi-GPU host uses so that
 * the fp32 sums can go straight into ncclReduce.
 * ONE render in flight per scene: the scene owns the scratch the kernels write (sample
 * planes, counters, the wavefront pool and its CUDA graph), so a second rt_render_device
 * or rt_render before rt_render_wait returns RT_ERR_BAD_ARGUMENT ("render pending").
 * The image is OVERWRITTEN with the sums of the requested sample range, not added to. */
RtStatus rt_render_device(const RtScene *scene, const RtCamera *camera, uint32_t width,
                          uint32_t height, uint32_t spp, uint32_t max_depth,
                          const RtRenderOpts *opts, float *out_rgb_sum_device, void *cuda_stream);
/* Block until the last rt_render_device on this scene finished; fill stats. */
RtStatus rt_render_wait(const RtScene *scene, RtStats *stats);

/* Parity hook 1: closest hit of world.hit(ray, 1e-5, +inf) (src/main.rs:48) for
 * n caller-supplied rays.  Media are skipped (their hit draws a random number,
 * src/medium.rs:42); they are covered by rt_path_radiance. */
RtStatus rt_trace_first_hit(const RtScene *scene, const RtRay *rays, uint64_t n, RtHit *hits);

/* Parity hook 2: radiance of individual paths.  For i in [0,n): the path of
 * pixel (px[i], py[i]) — py counted bottom-up like j in src/main.rs:772 — and
 * sample index sample[i].  rgb receives 3 doubles per path, segments (optional)
 * the number of world.hit calls the path made. */
RtStatus rt_path_radiance(const RtScene *scene, const RtCamera *camera, uint32_t width,
                          uint32_t height, uint32_t max_depth, const RtRenderOpts *opts,
                          const uint32_t *px, const uint32_t *py, const uint32_t *sample,
                          uint64_t n, double *rgb, uint32_t *segments);

/* Parity hook 3: the primary rays the render kernel generates (src/main.rs:813-820
 * + src/camera.rs:51-59) for the same (pixel, sample) addressing. */
RtStatus rt_camera_rays(const RtScene *scene, const RtCamera *camera, uint32_t width,
                        uint32_t height, const RtRenderOpts *opts, const uint32_t *px,
                        const uint32_t *py, const uint32_t *sample, uint64_t n, RtRay *rays);

/* ------------------------------------------------------------------------ */
/* Compile once, create many (one process per GPU: SURVEY §8(e))              */
/* ------------------------------------------------------------------------ */

/* rt_scene_create = compile on the host (graph walk, reference order, SAH BVH build: 0.5 s for
 * the 394k triangles of config 5) + upload.  With one process per GPU every rank would repeat
 * the compile on the same host cores; instead one rank compiles into a RELOCATABLE blob (plain
 * bytes, no pointers), the blob travels (ncclBroadcast, a file, a pipe) and every rank creates
 * its device scene from it.  Replaces BVH::new (src/bvh.rs:18-73) running once per process.
 * Needs no GPU.  The blob is only valid for the library version that made it (checked). */
typedef struct RtCompiled RtCompiled;
RtStatus rt_compile(const RtSceneDesc *desc, RtCompiled **out_compiled);
const void *rt_compiled_data(const RtCompiled *compiled);
uint64_t rt_compiled_size(const RtCompiled *compiled);
/* FNV-1a over the tables (what two ranks compare to know they render the same scene). */
uint64_t rt_compiled_hash(const void *data, uint64_t size);
void rt_compiled_destroy(RtCompiled *compiled);
/* Same result as rt_scene_create on the description the blob was compiled from.  A truncated,
 * foreign or corrupted blob is answered with RT_ERR_BAD_ARGUMENT. */
RtStatus rt_scene_create_compiled(const void *data, uint64_t size, int device, RtScene **out_scene);

/* ------------------------------------------------------------------------ */
/* One host thread, N GPUs (SURVEY §8(b) "Threading", §8(e))                  */
/* ------------------------------------------------------------------------ */

/* A scene replicated on several GPUs of one box.  The Rust host is one process
 * (src/main.rs has one main thread above rayon); this is the handle it holds
 * when render() should use every GPU.  The scene graph is compiled ONCE on the
 * host and the tables are uploaded to each device. */
typedef struct RtSceneGroup RtSceneGroup;

/* devices = NULL means devices 0..n_devices-1; n_devices = 0 means all visible devices. */
RtStatus rt_scene_group_create(const RtSceneDesc *desc, const int *devices, uint32_t n_devices,
                               RtSceneGroup **out_group);
void rt_scene_group_destroy(RtSceneGroup *group);
uint32_t rt_scene_group_size(const RtSceneGroup *group);
/* The per-device scene (i = 0 is the root that holds the combined image). */
const RtScene *rt_scene_group_scene(const RtSceneGroup *group, uint32_t i);

/* rt_render over all GPUs of the group.  The sample range of `opts` is cut into
 * contiguous blocks, one per GPU, every GPU renders its block for ALL pixels with
 * the same Philox keys (the union of samples is identical for any GPU count), and
 * the fp32 sum images are added on the root GPU in device order by ONE kernel that
 * reads the peers' images through NVLink peer mappings (staged with
 * cudaMemcpyPeerAsync where two devices cannot map each other).  The result is
 * bit-identical to rendering the same blocks one after the other and adding them
 * in fp32.  out_rgb_sum (host, W*H*3) may be NULL: the combined image then stays on
 * the root GPU for rt_encode_rgb8 / rt_encode_ppm. */
RtStatus rt_render_multi(const RtSceneGroup *group, const RtCamera *camera, uint32_t width, uint32_t height,
                         uint32_t spp, uint32_t max_depth, const RtRenderOpts *opts, float *out_rgb_sum,
                         RtStats *stats);

/* ------------------------------------------------------------------------ */
/* Output on the device: Vec3::format_color + the P3 body                    */
/* ------------------------------------------------------------------------ */

/* Vec3::format_color (src/vec.rs:125-131) for every pixel, on the GPU:
 * channel = (256 * clamp(sqrt(sum / samples_per_pixel), 0, 0.999)) as u64, with Rust's
 * saturating cast (NaN -> 0).  rgb_sum_device = NULL takes the image the last
 * rt_render / rt_render_multi on `scene` produced (it stays resident on the device);
 * otherwise it is a W*H*3 fp32 device buffer on the scene's device (e.g. the
 * ncclReduce result).  out_rgb8 is HOST memory, W*H*3 bytes. */
RtStatus rt_encode_rgb8(const RtScene *scene, const float *rgb_sum_device, uint32_t width, uint32_t height,
                        uint64_t samples_per_pixel, uint8_t *out_rgb8);

/* The whole P3 file the reference prints (src/main.rs:767-769,832): "P3\nW H\n255\n"
 * and one "r g b\n" line per pixel, formatted on the GPU (format_color, decimal
 * digits, a prefix sum over the line lengths) and copied out as text.  `out` is HOST
 * memory of `capacity` bytes; 32 + 12*W*H always suffices.  *length receives the file
 * size (no terminating NUL).  Byte-identical to the host writer. */
RtStatus rt_encode_ppm(const RtScene *scene, const float *rgb_sum_device, uint32_t width, uint32_t height,
                       uint64_t samples_per_pixel, char *out, uint64_t capacity, uint64_t *length);

/* Diagnostic: which pipeline build the last rt_render / rt_render_device on this scene ran, e.g.
 * "pipeline=megakernel variant=vflat blocks_per_sm=6".  The string lives until the next render
 * on the scene.  "" before the first render. */
const char *rt_render_info(const RtScene *scene);

/* Diagnostic: measured FP64 FMA throughput of `device` in TFLOP/s (a dependent-chain DFMA
 * micro-kernel, best of 5).  bench.py uses it as the denominator of the FP64-issue roofline,
 * the bound that actually applies to this path (DESIGN.md "Rooflines"). */
RtStatus rt_measure_fp64_peak(int device, double *tflops_out);

/* Thread-local description of the last error returned on this thread. */
const char *rt_last_error(void);

/* Library build info, e.g. "rtb200 abi 1 sm_100a f64". */
const char *rt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RTB200_H */
