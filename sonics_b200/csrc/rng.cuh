// Counter-based RNG for the PCR simulation kernels: Philox4x32-10
// (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11).
//
// The reference draws everything from numpy's process-global MT19937 and even
// reseeds it from itself inside the inner loop (sonics.pyx:428,450); that
// stream is irreproducible by design, so parity is distributional.
// Here every simulation owns the stream
//     key     = (seed_lo, seed_hi)
//     counter = (block, sim_lo, genotype_uid, stream | sim_hi << 8)
// so a simulation's draws depend only on (seed, uid, sim index, stream) and
// never on launch geometry, batch splits, or the GPU that runs it.  That is
// what lets the replay kernel regenerate the best simulation bit for bit.
#pragma once
#include <stdint.h>

namespace sonics {

enum : uint32_t { STREAM_MAIN = 0u, STREAM_READOUT = 1u };

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t &o0, uint32_t &o1, uint32_t &o2, uint32_t &o3)
{
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    o0 = c0;
    o1 = c1;
    o2 = c2;
    o3 = c3;
}

struct Rng {
    uint32_t k0, k1;
    uint32_t blk, c1, c2, c3;
    uint32_t b0, b1, b2, b3;
    int left;

    __device__ __forceinline__ void init(uint64_t seed, uint64_t sim, uint32_t uid, uint32_t stream)
    {
        k0 = (uint32_t)seed;
        k1 = (uint32_t)(seed >> 32);
        blk = 0u;
        c1 = (uint32_t)sim;
        c2 = uid;
        c3 = stream | ((uint32_t)(sim >> 32) << 8);
        left = 0;
        b0 = b1 = b2 = b3 = 0u;
    }
    __device__ __forceinline__ uint32_t next()
    {
        if (left == 0) {
            philox4x32_10(blk, c1, c2, c3, k0, k1, b0, b1, b2, b3);
            ++blk;
            left = 4;
        }
        const uint32_t r = b0;
        b0 = b1;
        b1 = b2;
        b2 = b3;
        --left;
        return r;
    }
    // uniform on the open interval: (2m+1) * 2^-24, m in [0, 2^23): never 0, never 1
    __device__ __forceinline__ float uniform_open()
    {
        return __uint2float_rn(((next() >> 9) << 1) | 1u) * 5.9604644775390625e-08f;
    }
    // uniform in [0, 1 - 2^-24] with 2^-32 resolution near 0 (inversion samplers)
    __device__ __forceinline__ float uniform_fine() { return __uint2float_rz(next()) * 2.3283064365386963e-10f; }
    // 53-bit double in [0,1): same construction as numpy's legacy next_double
    __device__ __forceinline__ double uniform53()
    {
        const uint32_t a = next() >> 5, b = next() >> 6;
        return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
    }
};

} // namespace sonics
