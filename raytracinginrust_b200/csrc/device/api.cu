// api.cu — the C ABI of include/rtb200.h: scene upload, render, parity hooks.
// There is no CPU fallback: every entry point that computes needs a CUDA device
// and fails with RT_ERR_CUDA when there is none.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/rtb200.h"
#include "compile.h"
#include "kernels.h"
#include "variants.h"
#include "wavefront.h"

using namespace rtb200dev;

namespace {
thread_local std::string g_err;

RtStatus fail(RtStatus s, const std::string &m) {
    g_err = m;
    return s;
}
RtStatus cuda_fail(cudaError_t e, const char *what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return RT_ERR_CUDA;
}
#define CU(call)                                             \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
}  // namespace

// ---------------------------------------------------------------------------
// Device memory that outlives its scene.  A host that renders scene after scene (bench.py's end-to-end leg, an
// animation) creates and destroys gigabytes per job: the f64 sample planes (1.1 GB for the Cornell box at 1000 spp,
// up to 4 GiB), the wavefront pool (0.7 GB), the tables of a mesh.  cudaMalloc / cudaFree of such blocks cost the job
// 5-30 ms of a 126 ms render, and occasionally hundreds (profiles/r1_h_whole_job_phases_cornell.txt; the smoke leg of
// profiles/r2_u_bench.json after the mesh leg had freed 4 GiB; single calls of 0.6-1.1 s in profiles/r2_v_*).  Device
// blocks are therefore parked here when their owner goes and handed to the next request of their size class, up to
// RTB200_SCRATCH_CACHE_MB (default 24576 - the planes of the Next Week final scene at 2000 spp are 15 GB; 0 = off)
// per process; rt_release_cached_memory() returns them to the driver.
// ---------------------------------------------------------------------------
namespace {
struct ScratchCache {
    struct Entry {
        int device;
        size_t bytes;
        void *p;
    };
    std::mutex m;
    std::vector<Entry> idle;
    size_t held = 0;
    static constexpr size_t kMinBytes = 256;
    // what a request is rounded up to, so that blocks find new owners: powers of two below 1 MiB, multiples of 2 MiB above
    static size_t size_class(size_t bytes) {
        if (bytes <= kMinBytes) return kMinBytes;
        if (bytes < (1u << 20)) {
            size_t c = kMinBytes;
            while (c < bytes) c <<= 1;
            return c;
        }
        return (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
    }

    static size_t cap() {
        static const size_t c = [] {
            long long mb = 24576;
            if (const char *v = std::getenv("RTB200_SCRATCH_CACHE_MB")) mb = std::atoll(v);
            return mb > 0 ? (size_t)mb << 20 : (size_t)0;
        }();
        return c;
    }
    // the smallest parked block of this device that holds `bytes` (a size class): of exactly that class below 1 MiB,
    // without wasting more than half of itself above
    void *take(int device, size_t bytes, size_t &capacity) {
        std::lock_guard<std::mutex> lock(m);
        const size_t most = bytes < (1u << 20) ? bytes : 2 * bytes + (64u << 20);
        size_t best = idle.size();
        for (size_t i = 0; i < idle.size(); ++i)
            if (idle[i].device == device && idle[i].bytes >= bytes && idle[i].bytes <= most &&
                (best == idle.size() || idle[i].bytes < idle[best].bytes))
                best = i;
        if (best == idle.size()) return nullptr;
        void *p = idle[best].p;
        capacity = idle[best].bytes;
        held -= capacity;
        idle.erase(idle.begin() + (long)best);
        return p;
    }
    // false: not kept (too small, the cache is off or full) - the caller frees it
    bool give(int device, void *p, size_t bytes) {
        std::lock_guard<std::mutex> lock(m);
        if (held + bytes > cap()) return false;
        idle.push_back(Entry{device, bytes, p});
        held += bytes;
        return true;
    }
    void release_all() {
        std::lock_guard<std::mutex> lock(m);
        int before = 0;
        cudaGetDevice(&before);
        for (const Entry &e : idle) {
            cudaSetDevice(e.device);
            cudaFree(e.p);
        }
        cudaSetDevice(before);
        idle.clear();
        held = 0;
    }
};
ScratchCache g_scratch;

// cudaMalloc through the cache; `capacity` is what scratch_free must be told
cudaError_t scratch_alloc(int device, size_t bytes, void **p, size_t *capacity) {
    bytes = ScratchCache::size_class(bytes);
    size_t cap = bytes;
    if (void *q = g_scratch.take(device, bytes, cap)) {
        *p = q;
        *capacity = cap;
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation && g_scratch.held) {  // the parked blocks are in the way: let them go and retry
        cudaGetLastError();
        g_scratch.release_all();
        e = cudaMalloc(p, bytes);
    }
    *capacity = bytes;
    return e;
}
// the caller has made sure that no work is in flight on the block (the scene's destructor synchronises the device)
void scratch_free(int device, void *p, size_t capacity) {
    if (!p) return;
    if (!g_scratch.give(device, p, capacity)) cudaFree(p);
}
struct Block {
    void *p;
    size_t capacity;
};

// The same for the few small pinned host blocks of a scene (counters, wavefront status): cudaFreeHost synchronises.
struct PinnedCache {
    std::mutex m;
    std::vector<void *> idle;  // blocks of kBytes
    static constexpr size_t kBytes = 256;
    cudaError_t alloc(void **p) {
        {
            std::lock_guard<std::mutex> lock(m);
            if (!idle.empty()) {
                *p = idle.back();
                idle.pop_back();
                return cudaSuccess;
            }
        }
        return cudaMallocHost(p, kBytes);
    }
    void free(void *p) {
        if (!p) return;
        std::lock_guard<std::mutex> lock(m);
        if (ScratchCache::cap() && idle.size() < 64) idle.push_back(p);
        else cudaFreeHost(p);
    }
    void release_all() {
        std::lock_guard<std::mutex> lock(m);
        for (void *p : idle) cudaFreeHost(p);
        idle.clear();
    }
};
PinnedCache g_pinned;
static_assert(sizeof(WfCtl) <= PinnedCache::kBytes && sizeof(unsigned long long) * kNumCounters <= PinnedCache::kBytes, "pinned block size");
}  // namespace

struct RtScene {
    int device = 0;
    DScene ds{};
    std::vector<Block> allocations;
    uint64_t device_bytes = 0;
    uint32_t n_lights = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint32_t features = 0;     // Feat bits of the scene (the integrator bit is added per render)
    uint32_t gpu_built_trees = 0;  // trees built by gpu_bvh.cu (RT_CREATE_GPU_BVH)
    int render_variant = 0;    // bits 0-1: register budget of the megakernel, bit 2: media
    // scratch reused across render calls (the handle is thread-compatible, not thread-safe)
    double *planes = nullptr;
    size_t planes_bytes = 0;     // what a render may use
    size_t planes_capacity = 0;  // what the block holds (scratch_alloc)
    float *out_dev = nullptr;
    size_t out_bytes = 0, out_capacity = 0;
    uint32_t out_width = 0, out_height = 0;  // the image out_dev holds (rt_render / rt_render_multi), for rt_encode_*
    unsigned long long *counters = nullptr;
    unsigned long long *counters_host = nullptr;  // pinned
    // wavefront pipeline: path pool + queues, allocated on first use
    WfPool wf{};
    std::vector<Block> wf_allocations;
    unsigned *wf_status_host = nullptr;  // pinned
    int sms = 148;
    bool has_media = false;
    bool shutter_limited = false;  // CompiledScene::shutter_limited
    bool wavefront_default = false;
    bool wavefront_if_long = false;  // use_wavefront: above kWavefrontLongPaths
    std::string render_info;
    cudaGraphExec_t wf_exec = nullptr;  // the wavefront round loop as a CUDA graph (kept alive until the next render)
    cudaGraph_t wf_graph = nullptr;
    cudaStream_t wf_capture_stream = nullptr;
    WfCtl *wf_ctl_host = nullptr;  // pinned copy of the control block (rounds_done for the stats)
    int wf_rounds_per_launches = 0;  // graph mode: kernels per round, to complete the launch count after the fact
    // last async render
    bool pending = false;
    cudaStream_t pending_stream = nullptr;
    double t_call0 = 0.0;
    uint64_t pending_launches = 0;

    ~RtScene() {
        cudaSetDevice(device);
        cudaDeviceSynchronize();  // nothing of this scene is in flight when its blocks are parked (cudaFree used to imply it)
        for (const Block &b : allocations) scratch_free(device, b.p, b.capacity);
        if (planes) scratch_free(device, planes, planes_capacity);
        if (out_dev) scratch_free(device, out_dev, out_capacity);
        if (counters) scratch_free(device, counters, ScratchCache::size_class(sizeof(unsigned long long) * kNumCounters));
        g_pinned.free(counters_host);
        if (wf_exec) cudaGraphExecDestroy(wf_exec);
        if (wf_graph) cudaGraphDestroy(wf_graph);
        if (wf_capture_stream) cudaStreamDestroy(wf_capture_stream);
        g_pinned.free(wf_ctl_host);
        for (const Block &b : wf_allocations) scratch_free(device, b.p, b.capacity);
        g_pinned.free(wf_status_host);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

template <class T>
RtStatus upload(RtScene &s, const std::vector<T> &v, const T *&dev) {
    size_t bytes = v.size() * sizeof(T), capacity = 0;
    void *p = nullptr;
    CU(scratch_alloc(s.device, bytes, &p, &capacity));
    s.allocations.push_back(Block{p, capacity});
    if (bytes) CU(cudaMemcpy(p, v.data(), bytes, cudaMemcpyHostToDevice));
    s.device_bytes += bytes;
    dev = (const T *)p;
    return RT_OK;
}

uint32_t wf_pool_capacity() {
    if (const char *v = std::getenv("RTB200_WF_POOL")) {
        long long n = std::atoll(v);
        if (n >= 1024 && n <= (1ll << 26)) return (uint32_t)n;
    }
    return 1u << 22;  // 4 Mi slots (640 MB): per-round fixed costs and the extend tail amortise; measured best of 128 Ki..4 Mi
}

// Which pipeline a render runs: the flags win, then RTB200_PIPELINE, then the scene's default.
// `paths`: what this render traces.  A tree over spheres and boxes without media (RTiOW) is faster on the wavefront
// pipeline once the render is long enough to pay its fixed cost (pool initialisation, the latency of ~150 rounds, the
// tail): 500x500 at 800 spp 158.7 ms against 176.3, at 100 spp 27.4 against 24.4 - the lines cross at ~6e7 paths (r2-ai).
constexpr uint64_t kWavefrontLongPaths = 80000000ull;
bool use_wavefront(const RtScene &s, const RtRenderOpts *opts, uint32_t max_depth, uint64_t paths) {
    if (max_depth == 0) return false;  // nothing to trace: the megakernel returns black
    uint32_t flags = opts ? opts->flags : 0u;
    if (flags & RT_FLAG_WAVEFRONT) return true;
    if (flags & RT_FLAG_MEGAKERNEL) return false;
    if (const char *v = std::getenv("RTB200_PIPELINE")) {
        if (!std::strcmp(v, "wavefront")) return true;
        if (!std::strcmp(v, "megakernel")) return false;
    }
    return s.wavefront_default || (s.wavefront_if_long && paths >= kWavefrontLongPaths);
}
uint64_t render_paths(uint32_t width, uint32_t height, uint32_t spp, const RtRenderOpts *opts) {
    const uint32_t begin = opts ? opts->sample_begin : 0u;
    const uint64_t count = (opts && opts->sample_count) ? opts->sample_count : (spp > begin ? spp - begin : 0u);
    return (uint64_t)width * height * count;
}

// A path whose throughput is exactly zero contributes nothing - unless something later in it is
// NaN, because 0 * NaN is NaN and the reference has no guard (§Q10, §Q11).  With the reference's
// PBR material that is common (0/0 at main.rs:104 for directions sampled below the surface), so
// scenes that use it keep tracing zero-throughput paths, exactly like the reference.
uint32_t pbr_flags(const RtScene &s) { return (s.features & F_PBR) ? RT_FLAG_TRACE_ZERO_THROUGHPUT : 0u; }

RtStatus make_params(const RtScene &s, uint32_t width, uint32_t height, uint32_t spp, uint32_t max_depth,
                     const RtRenderOpts *opts, RenderParams &P) {
    if (width < 2 || height < 2) return fail(RT_ERR_BAD_ARGUMENT, "width and height must be at least 2 (main.rs:817-818 divides by W-1, H-1)");
    if ((uint64_t)width * height > (1ull << 31)) return fail(RT_ERR_BAD_ARGUMENT, "image too large");
    RtRenderOpts o{};
    if (opts) o = *opts;
    if (o.integrator > RT_INTEGRATOR_LEGACY) return fail(RT_ERR_BAD_ARGUMENT, "unknown integrator");
    if ((o.flags & RT_FLAG_WAVEFRONT) && (o.flags & RT_FLAG_MEGAKERNEL)) return fail(RT_ERR_BAD_ARGUMENT, "RT_FLAG_WAVEFRONT and RT_FLAG_MEGAKERNEL exclude each other");
    if (o.integrator == RT_INTEGRATOR_HEAD && s.n_lights == 0)
        return fail(RT_ERR_NO_LIGHTS, "HEAD integrator needs a non-empty light list (reference: unwrap() panic at hit.rs:94-96)");
    uint32_t begin = o.sample_begin;
    uint32_t count = o.sample_count ? o.sample_count : (spp > begin ? spp - begin : 0);
    if (count == 0) return fail(RT_ERR_BAD_ARGUMENT, "empty sample range");
    if ((uint64_t)begin + count > 0xFFFFFFFFull) return fail(RT_ERR_BAD_ARGUMENT, "sample range overflows");
    std::memset(&P, 0, sizeof(P));
    P.width = width;
    P.height = height;
    P.max_depth = max_depth;
    P.seed = o.seed;
    P.integrator = o.integrator;
    P.flags = o.flags | pbr_flags(s);
    P.sample_begin = begin;
    P.sample_end = begin + count;
    P.tiles_x = (width + 7) / 8;
    P.tiles_y = (height + 3) / 4;
    P.items_per_chunk = (uint64_t)P.tiles_x * P.tiles_y * 32ull;
    // Work items are (chunk of consecutive samples, pixel).  Small items keep the end of a render short: the last item
    // of every lane / pool slot is run to its end while the others idle; items that are too small hammer the one work
    // counter (an atomic per item) and multiply the f64 planes.  How many samples an item should hold depends on the
    // scene (r2-ac / r2-ad, profiles/r2_ad_samples_per_item.md):
    //  * flat scenes (no tree: the Cornell boxes) gain nothing from lanes that restart together and pay the item fetch
    //    at 2 active lanes: 32 samples per item (Cornell x1000 117.1 -> 113.9 ms, smoke 137.6 -> 135.2);
    //  * trees over spheres and boxes (RTiOW, the Next Week final scene): lanes / slots that ask for work together get
    //    neighbouring pixels, so short items keep a warp's rays together in the tree - ONE sample per item where the
    //    planes allow it (RTiOW x800 208.0 -> 191.7 ms at 2 or 1 per item, x100 - what one of eight GPUs renders - 24.9 ->
    //    22.8 ms at 1; final x512 271.9 -> 249.4 at 2); this is where the planes grow, so the budget is 16 GiB there;
    //  * triangle meshes: 8 as before (1 .. 16 samples per item: 489.6 .. 498 ms at 64 spp, the 4K planes are 199 MB each).
    // At least 2^23 items when the image is small or the render short, and never more planes than the budget.
    // The partition depends only on the scene, the image and the sample range, not on the pipeline or the GPU, so the
    // f64 summation order - and with it every bit of the image - is the same whichever way it is rendered.
    const bool tree = (s.features & F_BVH) != 0u, tris = (s.features & F_TRI) != 0u;
    const uint64_t kSamplesPerItem = !tree ? 32 : (tris ? 8 : 1), kMinItems = 1ull << 23;
    const uint64_t plane_budget = (tree && !tris) ? (16ull << 30) : (4ull << 30);
    uint64_t chunks = (count + kSamplesPerItem - 1) / kSamplesPerItem;
    const uint64_t for_balance = (kMinItems + P.items_per_chunk - 1) / P.items_per_chunk;
    if (chunks < for_balance) chunks = for_balance;
    uint64_t plane_cap = plane_budget / ((uint64_t)width * height * 3 * sizeof(double));
    if (plane_cap < 1) plane_cap = 1;
    if (chunks > plane_cap) chunks = plane_cap;
    if (chunks > 1024) chunks = 1024;
    if (chunks > count) chunks = count;
    if (chunks < 1) chunks = 1;
    if (const char *v = std::getenv("RTB200_CHUNKS")) {  // tuning: force the number of sample chunks
        long long n = std::atoll(v);
        if (n >= 1) chunks = (uint64_t)n < count ? (uint64_t)n : count;
        if (chunks > plane_cap) chunks = plane_cap;
    }
    P.chunk_size = (uint32_t)((count + chunks - 1) / chunks);
    P.n_chunks = (count + P.chunk_size - 1) / P.chunk_size;
    P.n_items = P.items_per_chunk * P.n_chunks;
    P.inv_items_per_chunk = 1.0 / (double)P.items_per_chunk;
    P.inv_tiles_x = 1.0 / (double)P.tiles_x;
    return RT_OK;
}

// The bounds of a MovingSphere whose centre extrapolates were built for shutter times in [0, 1] (compile.h).
RtStatus check_shutter(const RtScene &s, const RtCamera &cam) {
    if (!s.shutter_limited) return RT_OK;
    const double lo = cam.time0 < cam.time1 ? cam.time0 : cam.time1, hi = cam.time0 < cam.time1 ? cam.time1 : cam.time0;
    if (lo >= 0.0 && hi <= 1.0) return RT_OK;
    return fail(RT_ERR_UNSUPPORTED, "a MovingSphere with (time0, time1) != (0, 1) needs a camera shutter inside [0, 1]");
}

uint32_t scene_features(const CompiledScene &cs) {
    uint32_t f = 0;
    for (const DPrim &p : cs.prims) {
        switch (p.kind) {
            case PRIM_SPHERE: f |= F_SPHERE; break;
            case PRIM_MSPHERE: f |= F_MSPHERE; break;
            case PRIM_RECT: f |= F_RECT; break;
            case PRIM_TRI: f |= F_TRI; break;
            default: f |= F_BOX; break;
        }
    }
    if (!cs.nodes.empty()) f |= F_BVH;
    for (const DTexture &t : cs.textures)
        if (t.kind != RT_TEX_CONSTANT) f |= F_TEX;
    for (const DLight &l : cs.lights)
        if (l.kind == LIGHT_SPHERE) f |= F_SPHERE_LIGHT;
    for (const DMaterial &m : cs.materials) {
        if (m.kind == RT_MAT_METAL) f |= F_METAL;
        if (m.kind == RT_MAT_DIELECTRIC) f |= F_DIELECTRIC;
        if (m.kind == RT_MAT_PBR) f |= F_PBR;
    }
    return f;
}

// The pipeline build a render runs on: the most specific variant that covers the scene's features
// and the integrator (RTB200_VARIANT=<name> forces one, e.g. vall for an A/B).
const PipelineVariant *pick_variant(const RtScene &s, uint32_t integrator) {
    if (const char *v = std::getenv("RTB200_VARIANT"))
        if (const PipelineVariant *pv = find_variant_by_name(v)) return pv;
    return find_variant(s.features | (integrator == RT_INTEGRATOR_LEGACY ? F_LEGACY : F_HEAD));
}

RtStatus ensure_scratch(RtScene &s, const RenderParams &P, bool need_out) {
    size_t plane_bytes = (size_t)P.n_chunks * P.width * P.height * 3 * sizeof(double);
    if (plane_bytes > s.planes_capacity) {
        if (s.planes) scratch_free(s.device, s.planes, s.planes_capacity);  // (no render of this scene is pending here)
        s.planes = nullptr;
        s.planes_bytes = s.planes_capacity = 0;
        CU(scratch_alloc(s.device, plane_bytes, (void **)&s.planes, &s.planes_capacity));
    }
    s.planes_bytes = plane_bytes;
    size_t out_bytes = (size_t)P.width * P.height * 3 * sizeof(float);
    if (need_out && out_bytes > s.out_bytes) {
        if (s.out_dev) scratch_free(s.device, s.out_dev, s.out_capacity);
        s.out_dev = nullptr;
        s.out_bytes = s.out_capacity = 0;
        CU(scratch_alloc(s.device, out_bytes, (void **)&s.out_dev, &s.out_capacity));
        s.out_bytes = out_bytes;
    }
    return RT_OK;
}

template <class T>
RtStatus wf_alloc(RtScene &s, T *&ptr, size_t count) {
    void *p = nullptr;
    size_t capacity = 0;
    CU(scratch_alloc(s.device, count * sizeof(T), &p, &capacity));
    s.wf_allocations.push_back(Block{p, capacity});
    ptr = (T *)p;
    return RT_OK;
}

RtStatus ensure_wavefront_pool(RtScene &s, const RenderParams &P) {
    const uint32_t cap = wf_pool_capacity();
    if (s.wf.capacity != cap) {
        for (const Block &b : s.wf_allocations) scratch_free(s.device, b.p, b.capacity);
        s.wf_allocations.clear();
        s.wf = WfPool{};
        WfPool w{};
#define WA(field, count)                              \
    do {                                              \
        RtStatus a__ = wf_alloc(s, w.field, (count)); \
        if (a__ != RT_OK) return a__;                 \
    } while (0)
        WA(slots, cap); WA(sum, cap); WA(state, cap); WA(defer_q, cap); WA(ctl, 1);
#undef WA
        w.capacity = cap;
        s.wf = w;
        if (!s.wf_status_host) CU(g_pinned.alloc((void **)&s.wf_status_host));
        if (!s.wf_ctl_host) CU(g_pinned.alloc((void **)&s.wf_ctl_host));
    }
    // never more slots than there are work items
    uint64_t items = (uint64_t)P.width * P.height * P.n_chunks;
    s.wf.n_slots = (uint32_t)(items < cap ? items : cap);
    return RT_OK;
}

// The wavefront pipeline: rounds of shade / generate / extend / control until no path is alive.
// The number of rounds depends on the paths.  Default: the round is the body of a CUDA-graph WHILE
// node whose condition the control kernel sets on the device (cudaGraphSetConditional), so the whole
// loop is ONE graph launch - no host round trip, and rt_render_device stays asynchronous.
// Fallback (RTB200_WF_GRAPH=0, or a driver without conditional nodes): the host enqueues 8 rounds at
// a time and reads the live count back (an empty round is a few kernels that find nothing to do).
uint32_t wf_leave_threshold(const RtScene &s) {
    // When a warp of the extend stage goes back for new rays (wavefront.inl, phase C): with long,
    // uneven traversals (triangle BVHs) as soon as half of its rays wait; otherwise never - every
    // ray of a media / sphere scene walks the same sequence of queries, and a warp that stays in
    // lockstep runs those steps with all lanes (measured: profiles/r1_f_pipeline_ab.md).
    uint32_t leave = (s.features & F_TRI) ? 16u : 33u;
    if (const char *v = std::getenv("RTB200_WF_LEAVE")) leave = (uint32_t)std::atoi(v);
    leave = leave > 33u ? 33u : leave;  // 0: batch mode (refill only when the whole warp is idle); 33: the simple extend kernel
    return leave;
}

cudaError_t build_wavefront_graph(RtScene &s, const PipelineVariant &pv, const RtCamera &cam, const RenderParams &P) {
    if (s.wf_exec) cudaGraphExecDestroy(s.wf_exec);
    if (s.wf_graph) cudaGraphDestroy(s.wf_graph);
    s.wf_exec = nullptr;
    s.wf_graph = nullptr;
    cudaError_t e = cudaSuccess;
    if (!s.wf_capture_stream) e = cudaStreamCreateWithFlags(&s.wf_capture_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return e;
    if ((e = cudaGraphCreate(&s.wf_graph, 0)) != cudaSuccess) return e;
    cudaGraphConditionalHandle handle;
    if ((e = cudaGraphConditionalHandleCreate(&handle, s.wf_graph, 1, cudaGraphCondAssignDefault)) != cudaSuccess) return e;
    cudaGraphNodeParams np{};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = handle;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t node;
    if ((e = cudaGraphAddNode(&node, s.wf_graph, nullptr, 0, &np)) != cudaSuccess) return e;
    cudaGraph_t body = np.conditional.phGraph_out[0];
    if ((e = cudaStreamBeginCaptureToGraph(s.wf_capture_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed)) != cudaSuccess) return e;
    cudaError_t le = pv.wf_launch_round(s.ds, cam, P, s.wf, s.planes, s.counters, s.has_media, s.sms, wf_leave_threshold(s),
                                        (unsigned long long)handle, s.wf_capture_stream);
    e = cudaStreamEndCapture(s.wf_capture_stream, nullptr);
    if (le != cudaSuccess) return le;
    if (e != cudaSuccess) return e;
    return cudaGraphInstantiate(&s.wf_exec, s.wf_graph, 0);
}

RtStatus run_wavefront(RtScene &s, const PipelineVariant &pv, const RtCamera &cam, const RenderParams &P, cudaStream_t st) {
    int per_round = kWfLaunchesPerRound + ((pv.mask & F_TEX) ? 1 : 0);
    CU(pv.wf_launch_init(s.wf, st));
    s.pending_launches += 1;
    const char *g = std::getenv("RTB200_WF_GRAPH");
    if (!g || std::atoi(g) != 0) {
        cudaError_t e = build_wavefront_graph(s, pv, cam, P);
        if (e == cudaSuccess) e = cudaGraphLaunch(s.wf_exec, st);
        if (e == cudaSuccess) {
            // rounds_done comes back with the counters; the launch count is completed in finish_render
            CU(cudaMemcpyAsync(s.wf_ctl_host, s.wf.ctl, sizeof(WfCtl), cudaMemcpyDeviceToHost, st));
            s.wf_rounds_per_launches = per_round;
            s.render_info += " loop=graph";
            return RT_OK;
        }
        cudaGetLastError();  // conditional graph nodes unavailable: the host drives the loop
    }
    const int kRoundsPerCheck = 8;
    const uint32_t leave = wf_leave_threshold(s);
    s.wf_rounds_per_launches = 0;
    s.render_info += " loop=host";
    for (;;) {
        for (int k = 0; k < kRoundsPerCheck; ++k) CU(pv.wf_launch_round(s.ds, cam, P, s.wf, s.planes, s.counters, s.has_media, s.sms, leave, 0ull, st));
        s.pending_launches += (uint64_t)kRoundsPerCheck * per_round;
        CU(cudaMemcpyAsync(s.wf_status_host, &s.wf.ctl->status_live, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (*s.wf_status_host == 0u) break;
    }
    return RT_OK;
}

RtStatus enqueue_render(RtScene &s, const RtCamera &cam, const RenderParams &P, bool wavefront, float *out_dev, cudaStream_t st) {
    CU(cudaMemsetAsync(s.counters, 0, sizeof(unsigned long long) * kNumCounters, st));
    CU(cudaEventRecord(s.ev0, st));
    s.pending_launches = 0;
    s.wf_rounds_per_launches = 0;
    const PipelineVariant &pv = *pick_variant(s, P.integrator);
    if (wavefront) {
        s.render_info = std::string("pipeline=wavefront variant=") + pv.name + " pool_slots=" + std::to_string(s.wf.n_slots);
        RtStatus w = run_wavefront(s, pv, cam, P, st);
        if (w != RT_OK) return w;
    } else {
        const int variant = s.render_variant;
        int blocks = 0;
        CU(pv.render_grid_size(s.device, variant, &blocks));
        s.render_info = std::string("pipeline=megakernel variant=") + pv.name +
                        " blocks_per_sm=" + std::to_string(blocks / (s.sms > 0 ? s.sms : 1));
        CU(pv.launch_render(s.ds, cam, P, variant, blocks, s.planes, s.counters, st));
        s.pending_launches += 1;
    }
    s.render_info += " chunks=" + std::to_string(P.n_chunks) + " chunk_size=" + std::to_string(P.chunk_size);
    if (s.gpu_built_trees) s.render_info += " gpu_built_trees=" + std::to_string(s.gpu_built_trees);
    CU(launch_reduce_planes(s.planes, out_dev, (uint64_t)P.width * P.height * 3, P.n_chunks, st));
    CU(cudaEventRecord(s.ev1, st));
    CU(cudaMemcpyAsync(s.counters_host, s.counters, sizeof(unsigned long long) * kNumCounters, cudaMemcpyDeviceToHost, st));
    s.pending_launches += 1;
    return RT_OK;
}

RtStatus finish_render(RtScene &s, cudaStream_t st, RtStats *stats, uint64_t d2h_bytes) {
    CU(cudaStreamSynchronize(st));
    if (stats) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
        std::memset(stats, 0, sizeof(*stats));
        stats->paths = s.counters_host[kCounterPaths];
        stats->rays = s.counters_host[kCounterRays];
        stats->nonfinite_samples = s.counters_host[kCounterNonFinite];
        stats->render_ms = ms;
        stats->total_ms = now_ms() - s.t_call0;
        if (s.wf_rounds_per_launches) {  // the graph ran rounds_done rounds on its own
            s.pending_launches += (uint64_t)s.wf_ctl_host->rounds_done * s.wf_rounds_per_launches;
            s.wf_rounds_per_launches = 0;
        }
        stats->kernel_launches = s.pending_launches;
        stats->h2d_bytes = sizeof(DScene) + sizeof(RtCamera) + sizeof(RenderParams);  // kernel arguments only
        stats->d2h_bytes = d2h_bytes + sizeof(unsigned long long) * kNumCounters;
    }
    return RT_OK;
}

// rt_render's first half in two steps.  prepare: parameters and scratch memory (may allocate or free, which
// can synchronise with other devices of the process when peer mappings exist); launch: the kernels, enqueued
// on the scene's own stream into its own resident image (s.out_dev).  rt_render_multi prepares every GPU
// before it launches on any, so that no allocation falls between two launches.
struct PreparedRender {
    RenderParams P;
    bool wavefront = false;
};
RtStatus prepare_render_own(RtScene &s, uint32_t width, uint32_t height, uint32_t spp, uint32_t max_depth,
                            const RtRenderOpts *opts, PreparedRender &pr) {
    CU(cudaSetDevice(s.device));
    pr.wavefront = use_wavefront(s, opts, max_depth, render_paths(width, height, spp, opts));
    RtStatus st = make_params(s, width, height, spp, max_depth, opts, pr.P);
    if (st != RT_OK) return st;
    st = ensure_scratch(s, pr.P, true);
    if (st == RT_OK && pr.wavefront) st = ensure_wavefront_pool(s, pr.P);
    return st;
}
RtStatus launch_render_own(RtScene &s, const RtCamera &camera, const PreparedRender &pr) {
    s.t_call0 = now_ms();
    CU(cudaSetDevice(s.device));
    s.out_width = s.out_height = 0;
    RtStatus st = enqueue_render(s, camera, pr.P, pr.wavefront, s.out_dev, s.stream);
    if (st != RT_OK) return st;
    s.out_width = pr.P.width;
    s.out_height = pr.P.height;
    return RT_OK;
}
RtStatus start_render_own(RtScene &s, const RtCamera &camera, uint32_t width, uint32_t height, uint32_t spp,
                          uint32_t max_depth, const RtRenderOpts *opts) {
    PreparedRender pr;
    RtStatus st = prepare_render_own(s, width, height, spp, max_depth, opts, pr);
    if (st != RT_OK) return st;
    return launch_render_own(s, camera, pr);
}

// Upload a compiled scene to one device (the second half of rt_scene_create; a scene group compiles
// once and calls this per GPU).
RtStatus create_on_device(const CompiledScene &cs, int device, RtScene **out_scene) {
    *out_scene = nullptr;
    int n = rt_device_count();
    if (n == 0) return fail(RT_ERR_CUDA, "no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= n) return fail(RT_ERR_BAD_ARGUMENT, "device ordinal out of range");
    CU(cudaSetDevice(device));
    std::unique_ptr<RtScene> s(new RtScene());
    s->device = device;
    s->n_lights = (uint32_t)cs.lights.size();
    DScene &d = s->ds;
#define UP(field)                                   \
    do {                                            \
        RtStatus u__ = upload(*s, cs.field, d.field); \
        if (u__ != RT_OK) return u__;               \
    } while (0)
    UP(prims);
    UP(ops);
    UP(chains);
    UP(groups);
    UP(nodes);
    UP(media);
    UP(lights);
    UP(materials);
    UP(textures);
    UP(images);
    UP(perlin);
    UP(texels);
#undef UP
    // trees the compiler left for the device (RT_CREATE_GPU_BVH): built in place in the uploaded tables
    if (!cs.pending_bvh.empty()) {
        float *boxes_dev = nullptr;
        CU(cudaMalloc((void **)&boxes_dev, cs.pending_boxes.size() * sizeof(float)));
        cudaError_t e = cudaMemcpy(boxes_dev, cs.pending_boxes.data(), cs.pending_boxes.size() * sizeof(float), cudaMemcpyHostToDevice);
        for (size_t k = 0; e == cudaSuccess && k < cs.pending_bvh.size(); ++k) {
            const CompiledScene::PendingBvh &pb = cs.pending_bvh[k];
            e = build_bvh_on_device(const_cast<DPrim *>(d.prims) + pb.first_prim, pb.n_prims, boxes_dev + 6 * pb.first_box,
                                    const_cast<DBvhNode *>(d.nodes) + pb.node_base, pb.node_base, pb.first_prim, pb.lo, pb.hi, nullptr);
        }
        cudaFree(boxes_dev);
        if (e != cudaSuccess) return fail(RT_ERR_CUDA, std::string("GPU BVH build: ") + cudaGetErrorString(e));
        s->gpu_built_trees = (uint32_t)cs.pending_bvh.size();
    }
    d.n_world_groups = cs.n_world_groups;
    d.n_media = (uint32_t)cs.media.size();
    d.n_lights = (uint32_t)cs.lights.size();
    d.n_prims = (uint32_t)cs.prims.size();
    for (int a = 0; a < 3; ++a) d.background[a] = cs.background[a];
    CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&s->ev0));
    CU(cudaEventCreate(&s->ev1));
    {
        size_t cap = 0;
        CU(scratch_alloc(device, sizeof(unsigned long long) * kNumCounters, (void **)&s->counters, &cap));
    }
    CU(g_pinned.alloc((void **)&s->counters_host));
    // register budget of the megakernel (megakernel.inl: 0 = 6 blocks/SM, 1 = 8, 2 = 12) + media bit
    s->features = scene_features(cs);
    {
        const bool bvh = !cs.nodes.empty(), media = !cs.media.empty(), tris = (s->features & F_TRI) != 0u;
        int budget = (!bvh && !media) ? 0 : ((bvh && !media && !tris) ? 2 : 1);
        if (const char *v = std::getenv("RTB200_RENDER_VARIANT")) budget = std::atoi(v) & 3;
        s->render_variant = budget | (media ? 4 : 0);
    }
    CU(cudaDeviceGetAttribute(&s->sms, cudaDevAttrMultiProcessorCount, device));
    s->has_media = !cs.media.empty();
    s->shutter_limited = cs.shutter_limited;
    // Measured per scene class (profiles/r1_e_pipeline_ab.md): the wavefront stages beat the megakernel
    // only where world.hit is a long chain of queries - media over BVH scenes (the Next Week final
    // scene, +21 %); flat scenes and plain BVH scenes run faster with the path state in registers.
    s->wavefront_default = !cs.media.empty() && !cs.nodes.empty();
    s->wavefront_if_long = cs.media.empty() && !cs.nodes.empty() && !(scene_features(cs) & F_TRI);
    *out_scene = s.release();
    return RT_OK;
}

}  // namespace

extern "C" {

const char *rt_last_error(void) { return g_err.c_str(); }
const char *rt_version(void) { return "rtb200 abi 1 sm_100a f64"; }

int rt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

RtStatus rt_measure_fp64_peak(int device, double *tflops_out) {
    if (!tflops_out) return fail(RT_ERR_BAD_ARGUMENT, "null argument");
    int n = rt_device_count();
    if (n == 0) return fail(RT_ERR_CUDA, "no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= n) return fail(RT_ERR_BAD_ARGUMENT, "device ordinal out of range");
    CU(cudaSetDevice(device));
    CU(measure_fp64_peak(device, tflops_out));
    return RT_OK;
}

RtStatus rt_scene_create_ex(const RtSceneDesc *desc, int device, uint32_t create_flags, RtScene **out_scene) {
    if (!desc || !out_scene) return fail(RT_ERR_BAD_ARGUMENT, "null argument");
    *out_scene = nullptr;
    if (create_flags & ~(uint32_t)RT_CREATE_GPU_BVH) return fail(RT_ERR_BAD_ARGUMENT, "unknown create flag");
    CompiledScene cs;
    std::string err;
    CompileOptions opts;
    if (create_flags & RT_CREATE_GPU_BVH) opts.gpu_bvh_min_prims = 4096;  // smaller trees: the host build is microseconds
    RtStatus st = compile_scene(*desc, cs, err, opts);
    if (st != RT_OK) return fail(st, err);
    return create_on_device(cs, device, out_scene);
}

RtStatus rt_scene_create(const RtSceneDesc *desc, int device, RtScene **out_scene) {
    uint32_t flags = 0;
    if (const char *v = std::getenv("RTB200_GPU_BVH"))  // A/B switch for the probes
        if (std::atoi(v) != 0) flags |= RT_CREATE_GPU_BVH;
    return rt_scene_create_ex(desc, device, flags, out_scene);
}

void rt_scene_destroy(RtScene *scene) { delete scene; }

void rt_release_cached_memory(void) {
    g_scratch.release_all();
    g_pinned.release_all();
}

// ---------------------------------------------------------------------------
// Compile once, create many: CompiledScene <-> a relocatable blob
// ---------------------------------------------------------------------------
}  // extern "C"

struct RtCompiled {  // (not a std::vector: 64 MB of zero fill before every byte is overwritten cost 25 ms on config 5)
    std::unique_ptr<unsigned char[]> data;
    uint64_t size = 0;
};

namespace {
constexpr uint64_t kBlobMagic = 0x3230424c42425452ull;  // "RTBBLB02" (02: four-lane checksum)
constexpr uint32_t kBlobTables = 12;
struct BlobHeader {
    uint64_t magic;
    uint32_t abi_version, header_bytes;
    uint64_t total_bytes, hash;             // hash: FNV-1a over everything after the header
    uint64_t count[kBlobTables];            // elements per table, in the order of CompiledScene
    uint32_t elem_bytes[kBlobTables];       // sizeof of each table's element in the library that wrote the blob
    uint32_t n_world_groups, max_bvh_depth, shutter_limited, pad;
    double background[3];
};
// FNV-1a over 64-bit words in four interleaved lanes (the multiply is a 3-4 cycle dependency: one lane hashes 64 MB in
// ~25 ms, four in ~8), the lanes folded at the end, the tail byte by byte.
uint64_t fnv1a(const unsigned char *p, uint64_t n) {
    const uint64_t kPrime = 1099511628211ull;
    uint64_t h0 = 1469598103934665603ull, h1 = h0 ^ 1u, h2 = h0 ^ 2u, h3 = h0 ^ 3u;
    uint64_t i = 0;
    for (; i + 32 <= n; i += 32) {
        uint64_t w[4];
        std::memcpy(w, p + i, 32);
        h0 = (h0 ^ w[0]) * kPrime;
        h1 = (h1 ^ w[1]) * kPrime;
        h2 = (h2 ^ w[2]) * kPrime;
        h3 = (h3 ^ w[3]) * kPrime;
    }
    uint64_t h = h0;
    h = (h ^ h1) * kPrime;
    h = (h ^ h2) * kPrime;
    h = (h ^ h3) * kPrime;
    for (; i < n; ++i) h = (h ^ p[i]) * kPrime;
    return h;
}
// memcpy with a few threads for the large tables (a fresh 64 MB destination is mostly page faults)
void copy_bytes(unsigned char *dst, const void *src, uint64_t bytes) {
    const uint64_t kPiece = 8ull << 20;
    if (bytes < 2 * kPiece) {
        if (bytes) std::memcpy(dst, src, bytes);
        return;
    }
    const unsigned n = (unsigned)std::min<uint64_t>(4, bytes / kPiece);
    const uint64_t per = ((bytes + n - 1) / n + 63) & ~63ull;
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < n; ++t) {
        const uint64_t a = t * per, b = std::min(bytes, a + per);
        if (a < b) pool.emplace_back([=] { std::memcpy(dst + a, (const unsigned char *)src + a, b - a); });
    }
    std::memcpy(dst, src, std::min(bytes, per));
    for (std::thread &th : pool) th.join();
}
uint64_t pad16(uint64_t n) { return (n + 15u) & ~15ull; }

template <class F>
void for_each_table(CompiledScene &cs, F f) {
    int k = 0;
    f(k++, cs.prims); f(k++, cs.ops); f(k++, cs.chains); f(k++, cs.groups); f(k++, cs.nodes); f(k++, cs.media);
    f(k++, cs.lights); f(k++, cs.materials); f(k++, cs.textures); f(k++, cs.images); f(k++, cs.perlin); f(k++, cs.texels);
}

void serialize(CompiledScene &cs, RtCompiled &dst) {
    BlobHeader h;
    std::memset(&h, 0, sizeof(h));
    h.magic = kBlobMagic;
    h.abi_version = RTB200_ABI_VERSION;
    h.header_bytes = (uint32_t)pad16(sizeof(BlobHeader));
    uint64_t total = h.header_bytes;
    for_each_table(cs, [&](int k, auto &v) {
        h.count[k] = v.size();
        h.elem_bytes[k] = (uint32_t)sizeof(v[0]);
        total += pad16(v.size() * sizeof(v[0]));
    });
    h.total_bytes = total;
    h.n_world_groups = cs.n_world_groups;
    h.max_bvh_depth = cs.max_bvh_depth;
    h.shutter_limited = cs.shutter_limited ? 1u : 0u;
    for (int a = 0; a < 3; ++a) h.background[a] = cs.background[a];
    dst.data.reset(new unsigned char[total]);
    dst.size = total;
    unsigned char *blob = dst.data.get();
    std::memset(blob, 0, h.header_bytes);
    uint64_t off = h.header_bytes;
    for_each_table(cs, [&](int, auto &v) {
        const uint64_t bytes = v.size() * sizeof(v[0]);
        copy_bytes(blob + off, v.data(), bytes);
        std::memset(blob + off + bytes, 0, pad16(bytes) - bytes);
        off += pad16(bytes);
    });
    h.hash = fnv1a(blob + h.header_bytes, total - h.header_bytes);
    std::memcpy(blob, &h, sizeof(h));
}

// false: not a blob of this library (message in err)
bool deserialize(const void *data, uint64_t size, CompiledScene &cs, std::string &err) {
    BlobHeader h;
    if (!data || size < sizeof(BlobHeader)) return err = "compiled scene: truncated header", false;
    std::memcpy(&h, data, sizeof(h));
    if (h.magic != kBlobMagic || h.abi_version != RTB200_ABI_VERSION || h.header_bytes != pad16(sizeof(BlobHeader)))
        return err = "compiled scene: not a blob of this library version", false;
    if (h.total_bytes != size) return err = "compiled scene: size does not match the header (truncated?)", false;
    uint64_t need = h.header_bytes;
    bool sizes_ok = true;
    for_each_table(cs, [&](int k, auto &v) {
        if (h.elem_bytes[k] != sizeof(v[0]) || h.count[k] > (1ull << 34)) sizes_ok = false;
        else need += pad16(h.count[k] * sizeof(v[0]));
    });
    if (!sizes_ok || need != size) return err = "compiled scene: table sizes do not match this library", false;
    const unsigned char *p = (const unsigned char *)data;
    if (fnv1a(p + h.header_bytes, size - h.header_bytes) != h.hash) return err = "compiled scene: checksum mismatch (corrupted)", false;
    uint64_t off = h.header_bytes;
    for_each_table(cs, [&](int k, auto &v) {
        v.resize(h.count[k]);
        const uint64_t bytes = h.count[k] * sizeof(v[0]);
        copy_bytes((unsigned char *)v.data(), p + off, bytes);
        off += pad16(bytes);
    });
    cs.n_world_groups = h.n_world_groups;
    cs.max_bvh_depth = h.max_bvh_depth;
    if (cs.max_bvh_depth >= (uint32_t)kStackSize) return err = "compiled scene: tree deeper than the traversal stack", false;
    cs.shutter_limited = h.shutter_limited != 0u;
    for (int a = 0; a < 3; ++a) cs.background[a] = h.background[a];
    // what the kernels index without checking must stay inside the tables even for a blob that was made by hand
    if (cs.n_world_groups > cs.groups.size()) return err = "compiled scene: group count out of range", false;
    for (const DGroup &g : cs.groups)
        if ((uint64_t)g.first_prim + g.n_prims > cs.prims.size() || (g.bvh_root >= 0 && (uint64_t)g.bvh_root >= cs.nodes.size()))
            return err = "compiled scene: group out of range", false;
    for (const DBvhNode &n : cs.nodes)
        for (const int32_t c : {n.child0, n.child1})
            if (c >= 0 ? (uint64_t)c >= cs.nodes.size() : (uint64_t)((~(uint32_t)c) >> 3) + ((~(uint32_t)c) & 7u) + 1u > cs.prims.size())
                return err = "compiled scene: node child out of range", false;
    return true;
}
}  // namespace

extern "C" {

RtStatus rt_compile(const RtSceneDesc *desc, RtCompiled **out_compiled) {
    if (!desc || !out_compiled) return fail(RT_ERR_BAD_ARGUMENT, "null argument");
    *out_compiled = nullptr;
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) { if (getenv("RTB200_COMPILE_TIMING")) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "[rt_compile] %s %.3f s\n", what, std::chrono::duration<double>(t - T0).count()); T0 = t; } };
    CompiledScene cs;
    std::string err;
    RtStatus st = compile_scene(*desc, cs, err);
    lap("compile_scene");
    if (st != RT_OK) return fail(st, err);
    std::unique_ptr<RtCompiled> c(new RtCompiled());
    serialize(cs, *c);
    lap("serialize");
    *out_compiled = c.release();
    return RT_OK;
}
const void *rt_compiled_data(const RtCompiled *c) { return c ? c->data.get() : nullptr; }
uint64_t rt_compiled_size(const RtCompiled *c) { return c ? c->size : 0; }
uint64_t rt_compiled_hash(const void *data, uint64_t size) {
    BlobHeader h;
    if (!data || size < sizeof(h)) return 0;
    std::memcpy(&h, data, sizeof(h));
    return h.magic == kBlobMagic ? h.hash : 0;
}
void rt_compiled_destroy(RtCompiled *c) { delete c; }

RtStatus rt_scene_create_compiled(const void *data, uint64_t size, int device, RtScene **out_scene) {
    if (!data || !out_scene) return fail(RT_ERR_BAD_ARGUMENT, "null argument");
    *out_scene = nullptr;
    CompiledScene cs;
    std::string err;
    if (!deserialize(data, size, cs, err)) return fail(RT_ERR_BAD_ARGUMENT, err);
    return create_on_device(cs, device, out_scene);
}

uint64_t rt_scene_device_bytes(const RtScene *scene) { return scene ? scene->device_bytes : 0; }

const char *rt_render_info(const RtScene *scene) { return scene ? scene->render_info.c_str() : ""; }

RtStatus rt_render_device(const RtScene *scene, const RtCamera *camera, uint32_t width, uint32_t height, uint32_t spp,
                          uint32_t max_depth, const RtRenderOpts *opts, float *out_rgb_sum_device, void *cuda_stream) {
    if (!scene || !camera || !out_rgb_sum_device) return fail(RT_ERR_BAD_ARGUMENT, "null argument");
    RtScene &s = *const_cast<RtScene *>(scene);
    // One render in flight per scene: the planes, counters, events and the wavefront pool (and its graph) are the
    // scene's own scratch; a second enqueue would resize or overwrite them under the kernels still running.
    if (s.pending) return fail(RT_ERR_BAD_ARGUMENT, "render pending: call rt_render_wait before the next rt_render_device on this scene");
    if (check_shutter(s, *camera) != RT_OK) return RT_ERR_UNSUPPORTED;
    s.t_call0 = now_ms();
    CU(cudaSetDevice(s.device));
    RenderParams P;
    const bool wavefront = use_wavefront(s, opts, max_depth, render_paths(width, height, spp, opts));
    RtStatus st = make_params(s, width, height, spp, max_depth, opts, P);
    if (st != RT_OK) return st;
    st = ensure_scratch(s, P, false);
    if (st == RT_OK && wavefront) st = ensure_wavefront_pool(s, P);
    if (st != RT_OK) return st;
    st = enqueue_render(s, *camera, P, wavefront, out_rgb_sum_device, (cudaStream_t)cuda_stream);
    if (st != RT_OK) return st;
    s.pending = true;
    s.pending_stream = (cudaStream_t)cuda_stream;
    return RT_OK;
}

RtStatus rt_render_wait(const RtScene *scene, RtStats *stats) {
    if (!scene) return fail(RT_ERR_BAD_ARGUMENT, "null argument");
    RtScene &s = *const_cast<RtScene *>(scene);
    if (!s.pending) return fail(RT_ERR_BAD_ARGUMENT, "no render pending");
    CU(cudaSetDevice(s.device));
    s.pending = false;
    return finish_render(s, s.pending_stream, stats, 0);
}

RtStatus rt_render(const RtScene *scene, const RtCamera *camera, uint32_t width, uint32_t height, uint32_t spp,
                   uint32_t max_depth, const RtRenderOpts *opts, float *out_rgb_sum, RtStats *stats) {
    if (!scene || !camera) return fail(RT_ERR_BAD_ARGUMENT, "null argument");
    RtScene &s = *const_cast<RtScene *>(scene);
    if (s.pending) return fail(RT_ERR_BAD_ARGUMENT, "render pending: call rt_render_wait before rt_render on this scene");
    if (check_shutter(s, *camera) != RT_OK) return RT_ERR_UNSUPPORTED;
    RtStatus st = start_render_own(s, *camera, width, height, spp, max_depth, opts);
    if (st != RT_OK) return st;
    size_t out_bytes = out_rgb_sum ? (size_t)width * height * 3 * sizeof(float) : 0;
    if (out_rgb_sum) CU(cudaMemcpyAsync(out_rgb_sum, s.out_dev, out_bytes, cudaMemcpyDeviceToHost, s.stream));
    return finish_render(s, s.stream, stats, out_bytes);
}

// ---------------------------------------------------------------------------
// One host thread, N GPUs
// ---------------------------------------------------------------------------
}  // extern "C"

struct RtSceneGroup {
    std::vector<RtScene *> scenes;   // scenes[0] is the root
    std::vector<cudaEvent_t> done;   // per scene: its image is complete (recorded on its stream)
    std::vector<char> mapped;        // per scene: the root can read its memory directly (same device or peer access)
    std::vector<float *> staging;    // per scene, on the root device: copy target when not mapped
    size_t staging_bytes = 0;
    cudaEvent_t g0 = nullptr, g1 = nullptr;  // on the root stream: begin of the call / image combined

    ~RtSceneGroup() {
        if (!scenes.empty() && scenes[0]) {
            cudaSetDevice(scenes[0]->device);
            for (float *p : staging)
                if (p) cudaFree(p);
            if (g0) cudaEventDestroy(g0);
            if (g1) cudaEventDestroy(g1);
        }
        for (size_t i = 0; i < scenes.size(); ++i) {
            if (i < done.size() && done[i] && scenes[i]) {
                cudaSetDevice(scenes[i]->device);
                cudaEventDestroy(done[i]);
            }
            delete scenes[i];
        }
    }
};

namespace {
// Contiguous, balanced split of `count` samples over n devices (the first count % n get one more):
// the same partition as raytracinginrust_b200/multi_gpu.py: sample_partition.
void sample_block(uint32_t count, uint32_t i, uint32_t n, uint32_t *begin, uint32_t *len) {
    const uint32_t base = count / n, extra = count % n;
    *len = base + (i < extra ? 1u : 0u);
    *begin = i * base + (i < extra ? i : extra);
}
}  // namespace

extern "C" {

RtStatus rt_scene_group_create(const RtSceneDesc *desc, const int *devices, uint32_t n_devices, RtSceneGroup **out_group) {
    if (!desc || !out_group) return fail(RT_ERR_BAD_ARGUMENT, "null argument");
    *out_group = nullptr;
    CompiledScene cs;
    std::string err;
    RtStatus st = compile_scene(*desc, cs, err);  // once, whatever the number of GPUs
    if (st != RT_OK) return fail(st, err);
    const int visible = rt_device_count();
    if (visible == 0) return fail(RT_ERR_CUDA, "no CUDA device (this library has no CPU path)");
    if (n_devices == 0) {
        if (devices) return fail(RT_ERR_BAD_ARGUMENT, "a device list with n_devices = 0");
        n_devices = (uint32_t)visible;
    }
    if (n_devices > kMaxGroupDevices) return fail(RT_ERR_BAD_ARGUMENT, "at most 16 devices per group");
    std::vector<int> dev(n_devices);
    for (uint32_t i = 0; i < n_devices; ++i) {
        dev[i] = devices ? devices[i] : (int)i;
        if (dev[i] < 0 || dev[i] >= visible) return fail(RT_ERR_BAD_ARGUMENT, "device ordinal out of range");
    }
    std::unique_ptr<RtSceneGroup> g(new RtSceneGroup());
    g->scenes.assign(n_devices, nullptr);
    g->done.assign(n_devices, nullptr);
    g->mapped.assign(n_devices, 0);
    g->staging.assign(n_devices, nullptr);
    // context creation and the uploads run per device in parallel (a context alone is ~0.2 s)
    std::vector<RtStatus> status(n_devices, RT_OK);
    std::vector<std::string> message(n_devices);
    {
        std::vector<std::thread> workers;
        for (uint32_t i = 0; i < n_devices; ++i)
            workers.emplace_back([&, i]() {
                status[i] = create_on_device(cs, dev[i], &g->scenes[i]);
                if (status[i] == RT_OK && cudaEventCreateWithFlags(&g->done[i], cudaEventDisableTiming) != cudaSuccess) {
                    status[i] = RT_ERR_CUDA;
                    g_err = "cudaEventCreate";
                }
                if (status[i] != RT_OK) message[i] = g_err;
            });
        for (std::thread &w : workers) w.join();
    }
    for (uint32_t i = 0; i < n_devices; ++i)
        if (status[i] != RT_OK) return fail(status[i], "device " + std::to_string(dev[i]) + ": " + message[i]);
    // the root maps its peers (NVLink / NVSwitch on a B200 box)
    CU(cudaSetDevice(dev[0]));
    CU(cudaEventCreate(&g->g0));
    CU(cudaEventCreate(&g->g1));
    g->mapped[0] = 1;
    for (uint32_t i = 1; i < n_devices; ++i) {
        if (dev[i] == dev[0]) {
            g->mapped[i] = 1;
            continue;
        }
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, dev[0], dev[i]) != cudaSuccess) can = 0;
        if (can && !std::getenv("RTB200_NO_PEER")) {
            cudaError_t e = cudaDeviceEnablePeerAccess(dev[i], 0);
            if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) g->mapped[i] = 1;
        }
        cudaGetLastError();
    }
    *out_group = g.release();
    return RT_OK;
}

void rt_scene_group_destroy(RtSceneGroup *group) { delete group; }
uint32_t rt_scene_group_size(const RtSceneGroup *group) { return group ? (uint32_t)group->scenes.size() : 0u; }
const RtScene *rt_scene_group_scene(const RtSceneGroup *group, uint32_t i) {
    return (group && i < group->scenes.size()) ? group->scenes[i] : nullptr;
}

RtStatus rt_render_multi(const RtSceneGroup *group, const RtCamera *camera, uint32_t width, uint32_t height, uint32_t spp,
                         uint32_t max_depth, const RtRenderOpts *opts, float *out_rgb_sum, RtStats *stats) {
    if (!group || !camera) return fail(RT_ERR_BAD_ARGUMENT, "null argument");
    RtSceneGroup &g = *const_cast<RtSceneGroup *>(group);
    const double t0 = now_ms();
    RtRenderOpts o{};
    if (opts) o = *opts;
    const uint32_t begin = o.sample_begin;
    const uint32_t count = o.sample_count ? o.sample_count : (spp > begin ? spp - begin : 0);
    if (count == 0) return fail(RT_ERR_BAD_ARGUMENT, "empty sample range");
    const uint32_t n = (uint32_t)g.scenes.size();
    RtScene &root = *g.scenes[0];
    if (check_shutter(root, *camera) != RT_OK) return RT_ERR_UNSUPPORTED;
    const size_t out_bytes = (size_t)width * height * 3 * sizeof(float);
    // every GPU gets its block.  First all parameters and scratch memory, then the launches: those only
    // enqueue, so the GPUs run side by side
    const bool timing = std::getenv("RTB200_MULTI_TIMING") != nullptr;
    std::vector<uint32_t> active;
    std::vector<PreparedRender> prepared;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t b, len;
        sample_block(count, i, n, &b, &len);
        if (len == 0) continue;
        RtRenderOpts oi = o;
        oi.sample_begin = begin + b;
        oi.sample_count = len;
        PreparedRender pr;
        RtStatus st = prepare_render_own(*g.scenes[i], width, height, spp, max_depth, &oi, pr);
        if (st != RT_OK) return st;
        active.push_back(i);
        prepared.push_back(pr);
    }
    CU(cudaSetDevice(root.device));
    for (uint32_t i : active) {  // staging buffers on the root for peers it cannot map
        if (g.mapped[i]) continue;
        if (g.staging_bytes < out_bytes) {
            for (float *&p : g.staging) {
                if (p) cudaFree(p);
                p = nullptr;
            }
            g.staging_bytes = out_bytes;
        }
        if (!g.staging[i]) CU(cudaMalloc((void **)&g.staging[i], g.staging_bytes));
    }
    if (timing) std::fprintf(stderr, "[multi] prepared %zu devices at %.3f ms\n", active.size(), now_ms() - t0);
    CU(cudaEventRecord(g.g0, root.stream));
    for (size_t k = 0; k < active.size(); ++k) {
        RtScene &s = *g.scenes[active[k]];
        RtStatus st = launch_render_own(s, *camera, prepared[k]);
        if (st == RT_OK && cudaEventRecord(g.done[active[k]], s.stream) != cudaSuccess) st = fail(RT_ERR_CUDA, "cudaEventRecord");
        if (st != RT_OK) {
            for (size_t j = 0; j <= k; ++j) cudaStreamSynchronize(g.scenes[active[j]]->stream);
            return st;
        }
        if (timing) std::fprintf(stderr, "[multi] device %d enqueued at %.3f ms\n", s.device, now_ms() - t0);
    }
    // combine on the root: one kernel, the peers' images read in place over NVLink
    CU(cudaSetDevice(root.device));
    PeerImages peers{};
    for (uint32_t i : active) {
        if (i == 0) continue;
        RtScene &s = *g.scenes[i];
        CU(cudaStreamWaitEvent(root.stream, g.done[i], 0));
        const float *src = s.out_dev;
        if (!g.mapped[i]) {  // no peer mapping between the two devices: stage the image on the root
            CU(cudaMemcpyPeerAsync(g.staging[i], root.device, s.out_dev, s.device, out_bytes, root.stream));
            src = g.staging[i];
        }
        peers.image[peers.n++] = src;
    }
    CU(launch_sum_peers(root.out_dev, peers, (uint64_t)width * height * 3, root.sms, root.stream));
    CU(cudaEventRecord(g.g1, root.stream));
    if (out_rgb_sum) CU(cudaMemcpyAsync(out_rgb_sum, root.out_dev, out_bytes, cudaMemcpyDeviceToHost, root.stream));
    RtStats total{};
    for (size_t k = active.size(); k-- > 0;) {  // the root last: its stream ends with the combined image
        RtScene &s = *g.scenes[active[k]];
        CU(cudaSetDevice(s.device));
        RtStats one{};
        RtStatus st = finish_render(s, s.stream, &one, 0);
        if (st != RT_OK) return st;
        if (timing) std::fprintf(stderr, "[multi] device %d: %.3f ms on device, %llu paths, done at %.3f ms\n", s.device, one.render_ms,
                                 (unsigned long long)one.paths, now_ms() - t0);
        total.paths += one.paths;
        total.rays += one.rays;
        total.nonfinite_samples += one.nonfinite_samples;
        total.kernel_launches += one.kernel_launches;
        total.h2d_bytes += one.h2d_bytes;
        total.d2h_bytes += one.d2h_bytes;
    }
    if (stats) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, g.g0, g.g1));
        total.render_ms = ms;  // on the root's stream: first enqueue -> combined image
        total.total_ms = now_ms() - t0;
        total.kernel_launches += peers.n ? 1 : 0;
        total.d2h_bytes += out_rgb_sum ? out_bytes : 0;
        *stats = total;
    }
    root.render_info += " gpus=" + std::to_string(active.size());
    return RT_OK;
}

// ---------------------------------------------------------------------------
// Output on the device
// ---------------------------------------------------------------------------
static RtStatus encode_source(RtScene &s, const float *rgb_sum_device, uint32_t width, uint32_t height,
                              uint64_t samples_per_pixel, const float **src) {
    if (width == 0 || height == 0 || samples_per_pixel == 0) return fail(RT_ERR_BAD_ARGUMENT, "zero size");
    if ((uint64_t)width * height > (1ull << 31)) return fail(RT_ERR_BAD_ARGUMENT, "image too large");
    if (rt_device_count() == 0) return fail(RT_ERR_CUDA, "no CUDA device (this library has no CPU path)");
    if (rgb_sum_device) {
        if ((uintptr_t)rgb_sum_device & 15u) return fail(RT_ERR_BAD_ARGUMENT, "rgb_sum_device must be 16-byte aligned");
        *src = rgb_sum_device;
    } else {
        if (!s.out_dev || s.out_width != width || s.out_height != height)
            return fail(RT_ERR_BAD_ARGUMENT, "no resident image of this size (call rt_render / rt_render_multi first)");
        *src = s.out_dev;
    }
    return RT_OK;
}

RtStatus rt_encode_rgb8(const RtScene *scene, const float *rgb_sum_device, uint32_t width, uint32_t height,
                        uint64_t samples_per_pixel, uint8_t *out_rgb8) {
    if (!scene || !out_rgb8) return fail(RT_ERR_BAD_ARGUMENT, "null argument");
    RtScene &s = *const_cast<RtScene *>(scene);
    const float *src = nullptr;
    RtStatus st = encode_source(s, rgb_sum_device, width, height, samples_per_pixel, &src);
    if (st != RT_OK) return st;
    CU(cudaSetDevice(s.device));
    const uint64_t n_values = (uint64_t)width * height * 3;
    uint8_t *d = nullptr;
    size_t d_cap = 0;
    CU(scratch_alloc(s.device, n_values + 16, (void **)&d, &d_cap));
    cudaError_t e = launch_format_rgb8(src, d, n_values, (double)samples_per_pixel, s.sms, s.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_rgb8, d, n_values, cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
    if (e != cudaSuccess) cudaStreamSynchronize(s.stream);
    scratch_free(s.device, d, d_cap);
    if (e != cudaSuccess) return cuda_fail(e, "rt_encode_rgb8");
    return RT_OK;
}

RtStatus rt_encode_ppm(const RtScene *scene, const float *rgb_sum_device, uint32_t width, uint32_t height,
                       uint64_t samples_per_pixel, char *out, uint64_t capacity, uint64_t *length) {
    if (!scene || !out || !length) return fail(RT_ERR_BAD_ARGUMENT, "null argument");
    *length = 0;
    RtScene &s = *const_cast<RtScene *>(scene);
    const float *src = nullptr;
    RtStatus st = encode_source(s, rgb_sum_device, width, height, samples_per_pixel, &src);
    if (st != RT_OK) return st;
    char header[48];
    const int hl = std::snprintf(header, sizeof(header), "P3\n%u %u\n255\n", width, height);  // main.rs:767-769
    CU(cudaSetDevice(s.device));
    const uint64_t n_pixels = (uint64_t)width * height;
    const uint32_t nb = ppm_block_count(n_pixels);
    // one allocation: body (12 B/pixel at most) | packed | block_off | total | block_len
    const size_t body_bytes = (size_t)((n_pixels * 12 + 255) & ~255ull);
    const size_t packed_bytes = (size_t)((n_pixels * 4 + 255) & ~255ull);
    const size_t off_bytes = (size_t)(((uint64_t)nb * 8 + 255) & ~255ull);
    const size_t len_bytes = (size_t)(((uint64_t)nb * 4 + 255) & ~255ull);
    char *d = nullptr;
    size_t d_cap = 0;
    CU(scratch_alloc(s.device, body_bytes + packed_bytes + off_bytes + 256 + len_bytes, (void **)&d, &d_cap));
    char *body = d;
    uint32_t *packed = (uint32_t *)(d + body_bytes);
    uint64_t *block_off = (uint64_t *)(d + body_bytes + packed_bytes);
    uint64_t *total_dev = (uint64_t *)(d + body_bytes + packed_bytes + off_bytes);
    uint32_t *block_len = (uint32_t *)(d + body_bytes + packed_bytes + off_bytes + 256);
    uint64_t total = 0;
    cudaError_t e = launch_ppm_measure(src, packed, block_len, block_off, total_dev, n_pixels, (double)samples_per_pixel, s.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&total, total_dev, sizeof(total), cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
    if (e == cudaSuccess && (uint64_t)hl + total > capacity) {
        scratch_free(s.device, d, d_cap);
        *length = (uint64_t)hl + total;  // what it would have taken
        return fail(RT_ERR_BAD_ARGUMENT, "output buffer too small for the PPM (32 + 12*W*H always suffices)");
    }
    if (e == cudaSuccess) e = launch_ppm_write(packed, block_off, body, n_pixels, s.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out + hl, body, total, cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
    if (e != cudaSuccess) cudaStreamSynchronize(s.stream);
    scratch_free(s.device, d, d_cap);
    if (e != cudaSuccess) return cuda_fail(e, "rt_encode_ppm");
    std::memcpy(out, header, (size_t)hl);
    *length = (uint64_t)hl + total;
    return RT_OK;
}

RtStatus rt_trace_first_hit(const RtScene *scene, const RtRay *rays, uint64_t n, RtHit *hits) {
    if (!scene || (n && (!rays || !hits))) return fail(RT_ERR_BAD_ARGUMENT, "null argument");
    if (n == 0) return RT_OK;
    RtScene &s = *const_cast<RtScene *>(scene);
    CU(cudaSetDevice(s.device));
    RtRay *d_rays = nullptr;
    RtHit *d_hits = nullptr;
    CU(cudaMalloc((void **)&d_rays, n * sizeof(RtRay)));
    cudaError_t e = cudaMalloc((void **)&d_hits, n * sizeof(RtHit));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_rays, rays, n * sizeof(RtRay), cudaMemcpyHostToDevice, s.stream);
    if (e == cudaSuccess) e = launch_first_hit(s.ds, d_rays, n, d_hits, s.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hits, d_hits, n * sizeof(RtHit), cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
    cudaFree(d_rays);
    if (d_hits) cudaFree(d_hits);
    if (e != cudaSuccess) return cuda_fail(e, "rt_trace_first_hit");
    return RT_OK;
}

static RtStatus hook_params(const RtScene &s, uint32_t width, uint32_t height, uint32_t max_depth,
                            const RtRenderOpts *opts, bool need_lights, RenderParams &P) {
    if (width < 2 || height < 2) return fail(RT_ERR_BAD_ARGUMENT, "width and height must be at least 2");
    RtRenderOpts o{};
    if (opts) o = *opts;
    if (o.integrator > RT_INTEGRATOR_LEGACY) return fail(RT_ERR_BAD_ARGUMENT, "unknown integrator");
    if (need_lights && o.integrator == RT_INTEGRATOR_HEAD && s.n_lights == 0)
        return fail(RT_ERR_NO_LIGHTS, "HEAD integrator needs a non-empty light list (reference: unwrap() panic at hit.rs:94-96)");
    std::memset(&P, 0, sizeof(P));
    P.width = width;
    P.height = height;
    P.max_depth = max_depth;
    P.seed = o.seed;
    P.integrator = o.integrator;
    P.flags = o.flags | pbr_flags(s);
    return RT_OK;
}

static RtStatus upload_u32x3(cudaStream_t st, const uint32_t *px, const uint32_t *py, const uint32_t *sample, uint64_t n,
                             uint32_t **d_out) {
    uint32_t *d = nullptr;
    CU(cudaMalloc((void **)&d, 3 * n * sizeof(uint32_t)));
    cudaError_t e = cudaMemcpyAsync(d, px, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + n, py, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + 2 * n, sample, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) {
        cudaFree(d);
        return cuda_fail(e, "upload path ids");
    }
    *d_out = d;
    return RT_OK;
}

RtStatus rt_path_radiance(const RtScene *scene, const RtCamera *camera, uint32_t width, uint32_t height,
                          uint32_t max_depth, const RtRenderOpts *opts, const uint32_t *px, const uint32_t *py,
                          const uint32_t *sample, uint64_t n, double *rgb, uint32_t *segments) {
    if (!scene || !camera || (n && (!px || !py || !sample || !rgb))) return fail(RT_ERR_BAD_ARGUMENT, "null argument");
    if (n == 0) return RT_OK;
    RtScene &s = *const_cast<RtScene *>(scene);
    if (check_shutter(s, *camera) != RT_OK) return RT_ERR_UNSUPPORTED;
    CU(cudaSetDevice(s.device));
    RenderParams P;
    RtStatus st = hook_params(s, width, height, max_depth, opts, true, P);
    if (st != RT_OK) return st;
    for (uint64_t k = 0; k < n; ++k)
        if (px[k] >= width || py[k] >= height) return fail(RT_ERR_BAD_ARGUMENT, "pixel out of range");
    uint32_t *d_ids = nullptr;
    st = upload_u32x3(s.stream, px, py, sample, n, &d_ids);
    if (st != RT_OK) return st;
    double *d_rgb = nullptr;
    uint32_t *d_seg = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_rgb, 3 * n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_seg, n * sizeof(uint32_t));
    if (e == cudaSuccess) e = launch_path_radiance(s.ds, *camera, P, d_ids, d_ids + n, d_ids + 2 * n, n, d_rgb, d_seg, s.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(rgb, d_rgb, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess && segments) e = cudaMemcpyAsync(segments, d_seg, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
    cudaFree(d_ids);
    if (d_rgb) cudaFree(d_rgb);
    if (d_seg) cudaFree(d_seg);
    if (e != cudaSuccess) return cuda_fail(e, "rt_path_radiance");
    return RT_OK;
}

RtStatus rt_camera_rays(const RtScene *scene, const RtCamera *camera, uint32_t width, uint32_t height,
                        const RtRenderOpts *opts, const uint32_t *px, const uint32_t *py, const uint32_t *sample,
                        uint64_t n, RtRay *rays) {
    if (!scene || !camera || (n && (!px || !py || !sample || !rays))) return fail(RT_ERR_BAD_ARGUMENT, "null argument");
    if (n == 0) return RT_OK;
    RtScene &s = *const_cast<RtScene *>(scene);
    CU(cudaSetDevice(s.device));
    RenderParams P;
    RtStatus st = hook_params(s, width, height, 1, opts, false, P);
    if (st != RT_OK) return st;
    uint32_t *d_ids = nullptr;
    st = upload_u32x3(s.stream, px, py, sample, n, &d_ids);
    if (st != RT_OK) return st;
    RtRay *d_rays = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_rays, n * sizeof(RtRay));
    if (e == cudaSuccess) e = launch_camera_rays(*camera, P, d_ids, d_ids + n, d_ids + 2 * n, n, d_rays, s.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(rays, d_rays, n * sizeof(RtRay), cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
    cudaFree(d_ids);
    if (d_rays) cudaFree(d_rays);
    if (e != cudaSuccess) return cuda_fail(e, "rt_camera_rays");
    return RT_OK;
}

}  // extern "C"
