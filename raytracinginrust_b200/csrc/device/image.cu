// image.cu — what happens to the fp32 sum image after the render kernels (sm_100a):
//   * sum_peers_kernel: the multi-GPU combine of rt_render_multi.  ONE kernel on the root GPU adds the
//     peers' images, read through NVLink peer mappings, in device order (a fixed fp32 summation order).
//   * format_rgb8_kernel: Vec3::format_color (src/vec.rs:125-131) for every pixel.
//   * ppm_measure / ppm_scan / ppm_write: the P3 body the reference prints one println! at a time
//     (src/main.rs:832) - format_color, decimal digits and a prefix sum over the line lengths - so the
//     host receives finished text.  These are HBM-bound byte kernels: 128-bit loads of the sums, the text
//     staged in shared memory and stored with 128-bit writes.
#include <cuda_runtime.h>

#include "kernels.h"

namespace rtb200dev {

// ---------------------------------------------------------------------------
// multi-GPU combine
// ---------------------------------------------------------------------------
__global__ void sum_peers_kernel(float *__restrict__ out, const __grid_constant__ PeerImages peers, uint64_t n_values) {
    const uint64_t n4 = n_values / 4;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float4 *out4 = reinterpret_cast<float4 *>(out);
    for (uint64_t q = k; q < n4; q += stride) {
        float4 acc = out4[q];
        for (uint32_t d = 0; d < peers.n; ++d) {  // device order: ((root + peer1) + peer2) + ...
            const float4 v = reinterpret_cast<const float4 *>(peers.image[d])[q];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        out4[q] = acc;
    }
    for (uint64_t q = n4 * 4 + k; q < n_values; q += stride) {
        float acc = out[q];
        for (uint32_t d = 0; d < peers.n; ++d) acc += peers.image[d][q];
        out[q] = acc;
    }
}

cudaError_t launch_sum_peers(float *out, const PeerImages &peers, uint64_t n_values, int sms, cudaStream_t stream) {
    if (peers.n == 0 || n_values == 0) return cudaSuccess;
    uint64_t want = (n_values / 4 + 255) / 256;
    uint64_t cap = (uint64_t)(sms > 0 ? sms : 148) * 8;
    int blocks = (int)(want < cap ? (want ? want : 1) : cap);
    sum_peers_kernel<<<blocks, 256, 0, stream>>>(out, peers, n_values);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// format_color
// ---------------------------------------------------------------------------
// (256.0 * (sum/spp).sqrt().clamp(0.0, 0.999)) as u64 - f64::clamp keeps NaN, `NaN as u64` is 0 (§Q10)
__device__ __forceinline__ uint32_t format_channel(float sum, double spp) {
    double x = sqrt((double)sum / spp);
    if (x < 0.0) x = 0.0;
    if (x > 0.999) x = 0.999;
    const double y = 256.0 * x;
    return (y == y && y > 0.0) ? (uint32_t)y : 0u;
}

__global__ void format_rgb8_kernel(const float *__restrict__ sum, uint8_t *__restrict__ out, uint64_t n_values, double spp) {
    const uint64_t n4 = n_values / 4;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t q = k; q < n4; q += stride) {
        const float4 v = reinterpret_cast<const float4 *>(sum)[q];
        const uint32_t w = format_channel(v.x, spp) | (format_channel(v.y, spp) << 8) | (format_channel(v.z, spp) << 16) |
                           (format_channel(v.w, spp) << 24);
        reinterpret_cast<uint32_t *>(out)[q] = w;
    }
    for (uint64_t q = n4 * 4 + k; q < n_values; q += stride) out[q] = (uint8_t)format_channel(sum[q], spp);
}

cudaError_t launch_format_rgb8(const float *sum, uint8_t *out, uint64_t n_values, double spp, int sms, cudaStream_t stream) {
    if (n_values == 0) return cudaSuccess;
    uint64_t want = (n_values / 4 + 255) / 256;
    uint64_t cap = (uint64_t)(sms > 0 ? sms : 148) * 8;
    int blocks = (int)(want < cap ? (want ? want : 1) : cap);
    format_rgb8_kernel<<<blocks, 256, 0, stream>>>(sum, out, n_values, spp);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// P3 body
// ---------------------------------------------------------------------------
constexpr int kPpmThreads = 256;
constexpr int kPpmPixPerThread = 4;
constexpr int kPpmPixPerBlock = kPpmThreads * kPpmPixPerThread;  // 1024 pixels, at most 12 KB of text

__device__ __forceinline__ uint32_t digits(uint32_t v) { return v >= 100u ? 3u : (v >= 10u ? 2u : 1u); }

// one pixel -> r | g<<8 | b<<16 | line length<<24   ("r g b\n": 6..12 bytes)
__device__ __forceinline__ uint32_t pack_pixel(float r, float g, float b, double spp) {
    const uint32_t cr = format_channel(r, spp), cg = format_channel(g, spp), cb = format_channel(b, spp);
    return cr | (cg << 8) | (cb << 16) | ((digits(cr) + digits(cg) + digits(cb) + 3u) << 24);
}

// block-wide exclusive scan of one value per thread (256 threads); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums, uint32_t *total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (uint32_t)d) inc += o;
    }
    if (lane == 31u) warp_sums[warp] = inc;
    __syncthreads();
    uint32_t base = 0, all = 0;
#pragma unroll
    for (uint32_t w = 0; w < kPpmThreads / 32; ++w) {
        const uint32_t s = warp_sums[w];
        if (w < warp) base += s;
        all += s;
    }
    *total = all;
    return base + inc - v;
}

// pass 1: format_color per pixel, packed values to HBM (4 B/pixel), text bytes per block
__global__ void __launch_bounds__(kPpmThreads) ppm_measure_kernel(const float *__restrict__ sum, uint32_t *__restrict__ packed,
                                                                  uint32_t *__restrict__ block_len, uint64_t n_pixels, double spp) {
    __shared__ uint32_t warp_sums[kPpmThreads / 32];
    const uint64_t p0 = ((uint64_t)blockIdx.x * kPpmThreads + threadIdx.x) * kPpmPixPerThread;
    uint32_t px[kPpmPixPerThread] = {0u, 0u, 0u, 0u};
    if (p0 + kPpmPixPerThread <= n_pixels) {
        // 4 pixels = 12 floats = three 128-bit loads (p0 is a multiple of 4, so 3*p0 floats is 16-byte aligned)
        const float4 *s4 = reinterpret_cast<const float4 *>(sum + 3 * p0);
        const float4 a = s4[0], b = s4[1], c = s4[2];
        px[0] = pack_pixel(a.x, a.y, a.z, spp);
        px[1] = pack_pixel(a.w, b.x, b.y, spp);
        px[2] = pack_pixel(b.z, b.w, c.x, spp);
        px[3] = pack_pixel(c.y, c.z, c.w, spp);
        *reinterpret_cast<uint4 *>(packed + p0) = make_uint4(px[0], px[1], px[2], px[3]);
    } else {
        for (int i = 0; i < kPpmPixPerThread; ++i)
            if (p0 + i < n_pixels) {
                const float *s = sum + 3 * (p0 + i);
                px[i] = pack_pixel(s[0], s[1], s[2], spp);
                packed[p0 + i] = px[i];
            }
    }
    const uint32_t mine = (px[0] >> 24) + (px[1] >> 24) + (px[2] >> 24) + (px[3] >> 24);
    uint32_t total;
    block_exclusive_scan(mine, warp_sums, &total);
    if (threadIdx.x == 0) block_len[blockIdx.x] = total;
}

// pass 2: exclusive scan of the per-block byte counts (one block; chunks of 1024 with a running carry)
__global__ void __launch_bounds__(1024) ppm_scan_kernel(const uint32_t *__restrict__ block_len, uint64_t *__restrict__ block_off,
                                                        uint32_t n_blocks, uint64_t *__restrict__ total_out) {
    __shared__ uint64_t warp_sums[32];
    __shared__ uint64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n_blocks; base += 1024u) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t v = i < n_blocks ? (uint64_t)block_len[i] : 0ull;
        uint64_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (uint32_t)d) inc += o;
        }
        if (lane == 31u) warp_sums[warp] = inc;
        __syncthreads();
        uint64_t wbase = 0, all = 0;
        for (uint32_t w = 0; w < 32u; ++w) {
            const uint64_t s = warp_sums[w];
            if (w < warp) wbase += s;
            all += s;
        }
        const uint64_t carry = carry_s;
        if (i < n_blocks) block_off[i] = carry + wbase + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + all;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry_s;
}

__device__ __forceinline__ uint32_t put_number(char *dst, uint32_t v) {
    uint32_t n = 0;
    if (v >= 100u) dst[n++] = (char)('0' + v / 100u);
    if (v >= 10u) dst[n++] = (char)('0' + (v / 10u) % 10u);
    dst[n++] = (char)('0' + v % 10u);
    return n;
}

// pass 3: the text.  The block's lines are laid out in shared memory at the same offset modulo 16 as
// their place in the output, so whole 16-byte units move with 128-bit stores.
__global__ void __launch_bounds__(kPpmThreads) ppm_write_kernel(const uint32_t *__restrict__ packed,
                                                                const uint64_t *__restrict__ block_off, char *__restrict__ body,
                                                                uint64_t n_pixels) {
    __shared__ uint32_t warp_sums[kPpmThreads / 32];
    __shared__ __align__(16) char text[16 + kPpmPixPerBlock * 12];
    const uint64_t p0 = ((uint64_t)blockIdx.x * kPpmThreads + threadIdx.x) * kPpmPixPerThread;
    uint32_t px[kPpmPixPerThread] = {0u, 0u, 0u, 0u};
    if (p0 + kPpmPixPerThread <= n_pixels) {
        const uint4 v = *reinterpret_cast<const uint4 *>(packed + p0);
        px[0] = v.x; px[1] = v.y; px[2] = v.z; px[3] = v.w;
    } else {
        for (int i = 0; i < kPpmPixPerThread; ++i)
            if (p0 + i < n_pixels) px[i] = packed[p0 + i];
    }
    const uint32_t mine = (px[0] >> 24) + (px[1] >> 24) + (px[2] >> 24) + (px[3] >> 24);
    uint32_t total;
    uint32_t off = block_exclusive_scan(mine, warp_sums, &total);
    const uint64_t gbase = block_off[blockIdx.x];
    const uint32_t pad = (uint32_t)(gbase & 15ull);  // body is 16-byte aligned (cudaMalloc)
    char *dst = text + pad + off;
#pragma unroll
    for (int i = 0; i < kPpmPixPerThread; ++i) {
        if ((px[i] >> 24) == 0u) continue;  // past the end of the image
        dst += put_number(dst, px[i] & 255u);
        *dst++ = ' ';
        dst += put_number(dst, (px[i] >> 8) & 255u);
        *dst++ = ' ';
        dst += put_number(dst, (px[i] >> 16) & 255u);
        *dst++ = '\n';
    }
    __syncthreads();
    // [pad, pad+total) of `text` goes to body[gbase, gbase+total)
    const uint32_t end = pad + total;
    const uint32_t first_full = pad ? 16u : 0u;          // first 16-byte unit that starts inside the range
    const uint32_t last_full = end & ~15u;               // end of the last complete unit
    char *gal = body + (gbase - pad);                    // 16-byte aligned
    if (last_full > first_full) {
        for (uint32_t u = first_full / 16u + threadIdx.x; u < last_full / 16u; u += kPpmThreads)
            reinterpret_cast<uint4 *>(gal)[u] = reinterpret_cast<const uint4 *>(text)[u];
        for (uint32_t b = pad + threadIdx.x; b < first_full && b < end; b += kPpmThreads) gal[b] = text[b];      // head
        for (uint32_t b = (last_full > pad ? last_full : pad) + threadIdx.x; b < end; b += kPpmThreads) gal[b] = text[b];  // tail
    } else {
        for (uint32_t b = pad + threadIdx.x; b < end; b += kPpmThreads) gal[b] = text[b];
    }
}

uint32_t ppm_block_count(uint64_t n_pixels) { return (uint32_t)((n_pixels + kPpmPixPerBlock - 1) / kPpmPixPerBlock); }

cudaError_t launch_ppm_measure(const float *sum, uint32_t *packed, uint32_t *block_len, uint64_t *block_off, uint64_t *total,
                               uint64_t n_pixels, double spp, cudaStream_t stream) {
    const uint32_t nb = ppm_block_count(n_pixels);
    ppm_measure_kernel<<<nb, kPpmThreads, 0, stream>>>(sum, packed, block_len, n_pixels, spp);
    ppm_scan_kernel<<<1, 1024, 0, stream>>>(block_len, block_off, nb, total);
    return cudaGetLastError();
}

cudaError_t launch_ppm_write(const uint32_t *packed, const uint64_t *block_off, char *body, uint64_t n_pixels, cudaStream_t stream) {
    ppm_write_kernel<<<ppm_block_count(n_pixels), kPpmThreads, 0, stream>>>(packed, block_off, body, n_pixels);
    return cudaGetLastError();
}

}  // namespace rtb200dev
