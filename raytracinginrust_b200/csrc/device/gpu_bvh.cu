// gpu_bvh.cu — BVH::new (src/bvh.rs:18-73) on the GPU: SURVEY §8(f) rank 4.
//
// The host compiler builds a binned-SAH tree (compile.cpp, 0.2-0.3 s for the 394k triangles of config 5, on top of
// the graph walk).  This file is the alternative for callers that would rather have the scene sooner than the best
// tree (RT_CREATE_GPU_BVH / RTB200_GPU_BVH=1): a linear BVH built by six kernels in about a millisecond -
//
//   morton_keys_kernel   centroid of each primitive's box -> 30-bit Morton code | primitive index (unique 62-bit keys)
//   cub::DeviceRadixSort keys ascending (the one library call: a sort)
//   gather_prims_kernel  the 128-byte primitive records into Morton order (8 x 128-bit loads and stores each)
//   hierarchy_kernel     Karras 2012: every inner node finds its key range and split by binary search over the
//                        common-prefix length of neighbouring keys; children and parents in one pass, no atomics
//   refit_kernel         bottom-up boxes: a thread per leaf climbs; the second arrival at a node (one atomic counter
//                        per node) owns it and unions its children's boxes
//   emit_kernel          the 64-byte traversal nodes of tables.h: the two child boxes (fp32, already rounded outward
//                        by the host, unions stay outward) and the child codes
//
// The boxes only cull and every primitive stays reachable, so a render on this tree finds the same winner for every
// ray as one on the SAH tree: images are bit-identical (tests/test_gpu_parity.py), only the traversal is longer.
// A radix tree over 62-bit keys is at most 62 levels deep: below the traversal stack (kStackSize = 64).
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include <vector>

#include "kernels.h"
#include "tables.h"

namespace rtb200dev {

namespace {

struct Box6 {
    float lo[3], hi[3];
};

__device__ __forceinline__ uint32_t spread10(uint32_t v) {  // 10 bits -> every third bit
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__global__ void morton_keys_kernel(const Box6 *__restrict__ boxes, uint32_t n, float3 lo, float3 inv_extent,
                                   unsigned long long *__restrict__ keys) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const Box6 b = boxes[k];
    const float cx = (0.5f * (b.lo[0] + b.hi[0]) - lo.x) * inv_extent.x;
    const float cy = (0.5f * (b.lo[1] + b.hi[1]) - lo.y) * inv_extent.y;
    const float cz = (0.5f * (b.lo[2] + b.hi[2]) - lo.z) * inv_extent.z;
    const uint32_t x = (uint32_t)fminf(fmaxf(cx * 1024.0f, 0.0f), 1023.0f);
    const uint32_t y = (uint32_t)fminf(fmaxf(cy * 1024.0f, 0.0f), 1023.0f);
    const uint32_t z = (uint32_t)fminf(fmaxf(cz * 1024.0f, 0.0f), 1023.0f);
    const uint32_t code = (spread10(x) << 2) | (spread10(y) << 1) | spread10(z);
    keys[k] = ((unsigned long long)code << 32) | k;
}

__global__ void gather_prims_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst,
                                    const unsigned long long *__restrict__ keys, uint32_t n) {
    // eight threads per 128-byte record
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t k = t >> 3;
    if (k >= n) return;
    const uint32_t from = (uint32_t)keys[k];
    dst[k * 8 + (t & 7)] = src[(uint64_t)from * 8 + (t & 7)];
}

// length of the common prefix of keys i and j (-1 outside the array); the keys are unique
__device__ __forceinline__ int delta(const unsigned long long *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll((long long)(keys[i] ^ keys[j]));
}

// children: >= 0 inner node, < 0 leaf ~k.  parent[] is indexed [0, n-1) for inner nodes and [n-1, 2n-1) for leaves.
__global__ void hierarchy_kernel(const unsigned long long *__restrict__ keys, int n, int2 *__restrict__ children,
                                 int *__restrict__ parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0, t = l;
    do {  // steps of ceil(l / 2), ceil(l / 4), ..., 1
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + (d < 0 ? d : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const int left = lo == gamma ? ~gamma : gamma;
    const int right = hi == gamma + 1 ? ~(gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    parent[left >= 0 ? left : (n - 1) + gamma] = i;
    parent[right >= 0 ? right : (n - 1) + gamma + 1] = i;
    if (i == 0) parent[0] = -1;
}

// FRESH: the inner box was written by another thread of this launch (refit): read it past the L1
template <bool FRESH>
__device__ __forceinline__ Box6 child_box(int c, const Box6 *node_box, const Box6 *boxes, const unsigned long long *keys) {
    if (c < 0) return boxes[(uint32_t)keys[~c]];
    if (!FRESH) return node_box[c];
    Box6 b;
    const float *src = reinterpret_cast<const float *>(&node_box[c]);
    for (int k = 0; k < 3; ++k) {
        b.lo[k] = __ldcg(src + k);
        b.hi[k] = __ldcg(src + 3 + k);
    }
    return b;
}

__global__ void refit_kernel(const Box6 *__restrict__ boxes, const unsigned long long *__restrict__ keys, int n,
                             const int2 *__restrict__ children, const int *__restrict__ parent, Box6 *node_box,
                             unsigned *__restrict__ arrived) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int cur = parent[(n - 1) + k];
    while (cur >= 0) {
        __threadfence();  // the box this thread wrote below is visible before its arrival is counted
        if (atomicAdd(&arrived[cur], 1u) == 0u) return;  // the sibling subtree is not finished: its thread takes over
        __threadfence();
        const int2 ch = children[cur];
        const Box6 a = child_box<true>(ch.x, node_box, boxes, keys), b = child_box<true>(ch.y, node_box, boxes, keys);
        Box6 u;
        for (int ax = 0; ax < 3; ++ax) {
            u.lo[ax] = fminf(a.lo[ax], b.lo[ax]);
            u.hi[ax] = fmaxf(a.hi[ax], b.hi[ax]);
        }
        // plain stores to memory other threads read after their atomic: volatile keeps them out of registers
        volatile float *dst = reinterpret_cast<volatile float *>(&node_box[cur]);
        for (int ax = 0; ax < 3; ++ax) {
            dst[ax] = u.lo[ax];
            dst[3 + ax] = u.hi[ax];
        }
        cur = parent[cur];
    }
}

__global__ void emit_kernel(const Box6 *__restrict__ boxes, const unsigned long long *__restrict__ keys, int n,
                            const int2 *__restrict__ children, const Box6 *__restrict__ node_box, DBvhNode *__restrict__ nodes,
                            uint32_t node_base, uint32_t first_prim) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int2 ch = children[i];
    const Box6 a = child_box<false>(ch.x, node_box, boxes, keys), b = child_box<false>(ch.y, node_box, boxes, keys);
    DBvhNode dn;
    for (int ax = 0; ax < 3; ++ax) {
        dn.lo0[ax] = a.lo[ax];
        dn.hi0[ax] = a.hi[ax];
        dn.lo1[ax] = b.lo[ax];
        dn.hi1[ax] = b.hi[ax];
    }
    // inner child: its index in the scene's node table; leaf ~k: one primitive, (first_prim + k) << 3 | (count - 1 = 0)
    dn.child0 = ch.x >= 0 ? (int32_t)(node_base + (uint32_t)ch.x) : (int32_t) ~((first_prim + (uint32_t)~ch.x) << 3);
    dn.child1 = ch.y >= 0 ? (int32_t)(node_base + (uint32_t)ch.y) : (int32_t) ~((first_prim + (uint32_t)~ch.y) << 3);
    dn.pad0 = dn.pad1 = 0;
    nodes[i] = dn;
}

#define GB(x)                            \
    do {                                 \
        cudaError_t e__ = (x);           \
        if (e__ != cudaSuccess) {        \
            for (void *p : temps) cudaFree(p); \
            return e__;                  \
        }                                \
    } while (0)

}  // namespace

// prims: the group's first record on the device (n records, permuted in place into Morton order); boxes: n fp32
// boxes in the records' current order, already rounded outward; nodes: the group's slice of the node table (n - 1
// nodes; the root is nodes[0], i.e. node_base in the scene's numbering); lo / hi: the group's bounds.
cudaError_t build_bvh_on_device(DPrim *prims, uint32_t n, const float *boxes_dev, DBvhNode *nodes, uint32_t node_base,
                                uint32_t first_prim, const double lo[3], const double hi[3], cudaStream_t st) {
    if (n < 2) return cudaErrorInvalidValue;
    std::vector<void *> temps;
    auto alloc = [&](void **p, size_t bytes) {
        cudaError_t e = cudaMalloc(p, bytes);
        if (e == cudaSuccess) temps.push_back(*p);
        return e;
    };
    unsigned long long *keys_in = nullptr, *keys = nullptr;
    DPrim *gathered = nullptr;
    int2 *children = nullptr;
    int *parent = nullptr;
    Box6 *node_box = nullptr;
    unsigned *arrived = nullptr;
    void *cub_tmp = nullptr;
    size_t cub_bytes = 0;
    GB(alloc((void **)&keys_in, sizeof(unsigned long long) * n));
    GB(alloc((void **)&keys, sizeof(unsigned long long) * n));
    GB(alloc((void **)&gathered, sizeof(DPrim) * (size_t)n));
    GB(alloc((void **)&children, sizeof(int2) * (n - 1)));
    GB(alloc((void **)&parent, sizeof(int) * (2 * (size_t)n - 1)));
    GB(alloc((void **)&node_box, sizeof(Box6) * (n - 1)));
    GB(alloc((void **)&arrived, sizeof(unsigned) * (n - 1)));
    GB(cub::DeviceRadixSort::SortKeys(nullptr, cub_bytes, keys_in, keys, (int)n, 0, 62, st));
    GB(alloc(&cub_tmp, cub_bytes ? cub_bytes : 16));
    const Box6 *boxes = reinterpret_cast<const Box6 *>(boxes_dev);
    const float3 flo = make_float3((float)lo[0], (float)lo[1], (float)lo[2]);
    const float3 inv = make_float3(hi[0] > lo[0] ? (float)(1.0 / (hi[0] - lo[0])) : 0.f, hi[1] > lo[1] ? (float)(1.0 / (hi[1] - lo[1])) : 0.f,
                                   hi[2] > lo[2] ? (float)(1.0 / (hi[2] - lo[2])) : 0.f);
    const unsigned tb = 256, gn = (n + tb - 1) / tb, gi = (n - 1 + tb - 1) / tb;
    morton_keys_kernel<<<gn, tb, 0, st>>>(boxes, n, flo, inv, keys_in);
    GB(cudaGetLastError());
    GB(cub::DeviceRadixSort::SortKeys(cub_tmp, cub_bytes, keys_in, keys, (int)n, 0, 62, st));
    gather_prims_kernel<<<(unsigned)(((uint64_t)n * 8 + tb - 1) / tb), tb, 0, st>>>(reinterpret_cast<const uint4 *>(prims),
                                                                                reinterpret_cast<uint4 *>(gathered), keys, n);
    GB(cudaGetLastError());
    GB(cudaMemcpyAsync(prims, gathered, sizeof(DPrim) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    GB(cudaMemsetAsync(arrived, 0, sizeof(unsigned) * (n - 1), st));
    hierarchy_kernel<<<gi, tb, 0, st>>>(keys, (int)n, children, parent);
    GB(cudaGetLastError());
    refit_kernel<<<gn, tb, 0, st>>>(boxes, keys, (int)n, children, parent, node_box, arrived);
    GB(cudaGetLastError());
    emit_kernel<<<gi, tb, 0, st>>>(boxes, keys, (int)n, children, node_box, nodes, node_base, first_prim);
    GB(cudaGetLastError());
    GB(cudaStreamSynchronize(st));
    for (void *p : temps) cudaFree(p);
    return cudaSuccess;
}

}  // namespace rtb200dev
