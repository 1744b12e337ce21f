// Generates tests/golden/philox_curand.json from NVIDIA's own Philox4x32-10
// (curand_Philox4x32_10 in <curand_philox4x32_x.h>, run on the host).
//   nvcc -o /tmp/gen_philox gen_philox_golden.cu && /tmp/gen_philox > philox_curand.json
#define QUALIFIERS static inline __host__ __device__
#include <cuda_runtime.h>
#include <curand_philox4x32_x.h>
#include <cstdint>
#include <cstdio>

int main() {
    // the three Random123 kat_vectors inputs, then a deterministic sweep
    uint32_t fixed[3][6] = {
        {0u, 0u, 0u, 0u, 0u, 0u},
        {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu},
        {0x243f6a88u, 0x85a308d3u, 0x13198a2eu, 0x03707344u, 0xa4093822u, 0x299f31d0u}};
    std::printf("[\n");
    uint64_t x = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < 67; ++i) {
        uint32_t v[6];
        if (i < 3) {
            for (int k = 0; k < 6; ++k) v[k] = fixed[i][k];
        } else {
            for (int k = 0; k < 6; ++k) {  // splitmix64
                x += 0x9E3779B97F4A7C15ull;
                uint64_t z = x;
                z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
                z ^= z >> 31;
                v[k] = (i < 35) ? (uint32_t)(z % 1024u) : (uint32_t)z;  // small counters like ours, then full-range
            }
        }
        uint4 c = make_uint4(v[0], v[1], v[2], v[3]);
        uint2 k = make_uint2(v[4], v[5]);
        uint4 o = curand_Philox4x32_10(c, k);
        std::printf("  {\"ctr\": [%u, %u, %u, %u], \"key\": [%u, %u], \"out\": [%u, %u, %u, %u]}%s\n", v[0], v[1], v[2],
                    v[3], v[4], v[5], o.x, o.y, o.z, o.w, i == 66 ? "" : ",");
    }
    std::printf("]\n");
    return 0;
}
