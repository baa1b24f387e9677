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
//
// Consumption is in whole 128-bit blocks requested at warp-convergent points of
// the kernel (one block serves the proposal pairs of two rejection trials), not
// through a per-lane buffer: a buffered next() desynchronises the lanes' refill
// points and the 10-round Philox body then runs at 3-4 active lanes (ncu, r01).
#pragma once
#include <stdint.h>

namespace sonics {

enum : uint32_t { STREAM_MAIN = 0u, STREAM_READOUT = 1u };

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1)
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
    return make_uint4(c0, c1, c2, c3);
}

// Philox2x32-10 (same paper, same round structure with one multiplier): 64 random bits per call at about
// half the instructions of the 4x32 variant.
__device__ __forceinline__ uint2 philox2x32_10(uint32_t c0, uint32_t c1, uint32_t k)
{
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi = __umulhi(0xD256D193u, c0), lo = 0xD256D193u * c0;
        c0 = hi ^ k ^ c1;
        c1 = lo;
        k += 0x9E3779B9u;
    }
    return make_uint2(c0, c1);
}

// uniform on the open interval from the top 23 bits: (2m+1) * 2^-24, never 0, never 1.
// Pure ALU (no I2F on the XU pipe): 1.m - (1 - 2^-24).
__device__ __forceinline__ float u01_open(uint32_t x)
{
    return __uint_as_float(0x3f800000u | (x >> 9)) - 0.99999994f;
}
// uniform in [0, 1 - 2^-24] with 2^-32 resolution near 0 (inversion samplers)
__device__ __forceinline__ float u01_fine(uint32_t x) { return __uint2float_rz(x) * 2.3283064365386963e-10f; }
// 53-bit double in [0,1): same construction as numpy's legacy next_double
__device__ __forceinline__ double u01_53(uint32_t a, uint32_t b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

struct Rng {
    uint32_t k0, k1;
    uint32_t blk, c1, c2, c3;

    __device__ __forceinline__ void init(uint64_t seed, uint64_t sim, uint32_t uid, uint32_t stream)
    {
        k0 = (uint32_t)seed;
        k1 = (uint32_t)(seed >> 32);
        blk = 0u;
        c1 = (uint32_t)sim;
        c2 = uid;
        c3 = stream | ((uint32_t)(sim >> 32) << 8);
    }
    // the next 4 x 32 random bits of this simulation's stream
    __device__ __forceinline__ uint4 block4()
    {
        const uint4 r = philox4x32_10(blk, c1, c2, c3, k0, k1);
        ++blk;
        return r;
    }
};

// The draw stream of K1a's state machine.  A rejection trial consumes two words and an inversion one, so
// the hot loop takes 64-bit blocks from Philox2x32-10 instead of 128-bit ones (ncu, r01: the 4x32 body was
// 9 % of all warp-instructions with half of every block unused):
//     key     = sim_lo ^ seed_lo
//     counter = (block | sim_hi << 24, genotype_uid ^ seed_hi)
// -- a bijection of (uid, sim) for a fixed seed, so no two simulations of a run share a stream; the
// priors and start alleles of the simulation still come from the 4x32 stream above.
struct PairRng {
    uint32_t k, blk, c1;
    __device__ __forceinline__ void init(uint64_t seed, uint64_t sim, uint32_t uid)
    {
        k = (uint32_t)sim ^ (uint32_t)seed;
        blk = (uint32_t)(sim >> 32) << 24;
        c1 = uid ^ (uint32_t)(seed >> 32);
    }
    __device__ __forceinline__ uint2 pair()
    {
        const uint2 r = philox2x32_10(blk, c1, k);
        ++blk;
        return r;
    }
};

// Hands out (a, b) pairs of 32-bit words: first the pair it was constructed with, then fresh
// blocks from the stream (two pairs per block).  Used by the rejection samplers for retries.
struct PairSource {
    Rng &rng;
    uint32_t a, b, sa, sb;
    int spare;
    __device__ __forceinline__ PairSource(Rng &r, uint32_t a0, uint32_t b0) : rng(r), a(a0), b(b0), sa(0), sb(0), spare(0)
    {
    }
    __device__ __forceinline__ void advance()
    {
        if (spare) {
            a = sa;
            b = sb;
            spare = 0;
        } else {
            const uint4 r = rng.block4();
            a = r.x;
            b = r.y;
            sa = r.z;
            sb = r.w;
            spare = 1;
        }
    }
};

} // namespace sonics
