// Counter-based Philox4x32-10 (Salmon et al., SC'11), host + device.
//
// Replaces the reference's stateful generators: rng.normal (hmc.py:82-87, sgld.py:41-46,
// sghmc.py:31) and the global np.random.rand (hmc.py:46,61).  A draw is addressed by
//   key     = 64-bit seed
//   counter = (element_block, global_chain_id, stream_lo, stream_hi)
// so results do not depend on how chains are sharded over GPUs or on launch geometry.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace bhmc {

struct U4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
  lo = a * b;
  hi = __umulhi(a, b);
#else
  uint64_t r = (uint64_t)a * (uint64_t)b;
  lo = (uint32_t)r;
  hi = (uint32_t)(r >> 32);
#endif
}

__host__ __device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo(M0, c.x, hi0, lo0);
    mulhilo(M1, c.z, hi1, lo1);
    U4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += W0;
    k1 += W1;
  }
  return c;
}

// stream tags (stream_hi high byte) -- one namespace per consumer
enum : uint32_t {
  TAG_MOMENTUM = 0x01000000u,  // hmc.py:41  N(0,1) momentum; stream_lo = step
  TAG_PATH = 0x02000000u,      // hmc.py:46  uniform for the path length
  TAG_ACCEPT = 0x03000000u,    // hmc.py:61  uniform for the Metropolis test
  TAG_NOISE = 0x04000000u,     // sgld.py:45 / sghmc.py:31 N(0, 2 eps) noise; low bits = iteration*vars+var
  TAG_DROPOUT = 0x05000000u,   // mlp.py:29-31 dropout masks
  TAG_TEST = 0x7f000000u
};

// uniform in (0,1): 23 random bits.  (k + 0.5) * 2^-23 with k < 2^23 is exact in fp32 (k + 0.5 needs 24 significant
// bits), so the value is never 0 or 1 (safe under log) and every one of the 2^23 grid points is equally likely.
// (With 24 bits, k + 0.5 is not representable for k >= 2^23: it rounded to even, and k = 2^24 - 1 gave exactly 1.0f.)
__host__ __device__ __forceinline__ float u01_open(uint32_t r) { return ((float)(r >> 9) + 0.5f) * (1.0f / 8388608.0f); }

// uniform in [0,1) with 53 bits, identical on host and device
__host__ __device__ __forceinline__ double u01_double(uint32_t hi, uint32_t lo) {
  uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
  return (double)v * (1.0 / 9007199254740992.0);
}

__host__ __device__ __forceinline__ double philox_uniform(uint64_t seed, int64_t chain, uint32_t slo, uint32_t shi) {
  U4 c{0u, (uint32_t)chain, slo, shi ^ (uint32_t)((uint64_t)chain >> 32)};
  U4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  return u01_double(r.x, r.y);
}

#ifdef __CUDACC__
// four N(0,1) for element block `blk` (elements 4*blk .. 4*blk+3) of a chain: Box-Muller
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, int64_t chain, uint32_t blk, uint32_t slo,
                                                 uint32_t shi) {
  U4 c{blk, (uint32_t)chain, slo, shi ^ (uint32_t)((uint64_t)chain >> 32)};
  U4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  float u0 = u01_open(r.x), u1 = u01_open(r.y), u2 = u01_open(r.z), u3 = u01_open(r.w);
  float r0 = sqrtf(-2.0f * __logf(u0)), r1 = sqrtf(-2.0f * __logf(u2));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u1, &s0, &c0);
  __sincosf(6.283185307179586f * u3, &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}
#endif

}  // namespace bhmc
