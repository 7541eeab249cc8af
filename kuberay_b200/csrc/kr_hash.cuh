// kr_hash.cuh — SHA-1 + base32hex of the muted-spec JSON (GenerateJsonHash, utils/util.go:628-640).
// Part of the sm_100a kernel set of the batched reconcile engine; see kr_kernels.cuh for the pipeline overview.
#pragma once

#include "kr_common.cuh"

namespace kr {

// ------------------------------------------------------------------------------------------------ k_hash
// base32hex(sha1(json)) per RayCluster (utils/util.go:628-640).  One lane per message (SHA-1 is a serial chain per
// message); the warp stages 128 bytes of each of its 32 messages per step with coalesced 16-byte loads into an
// XOR-swizzled shared tile, so the per-lane reads are conflict-free LDS.128.

__device__ __forceinline__ uint32_t rol(uint32_t x, int k) { return __funnelshift_l(x, x, k); }
__device__ __forceinline__ uint32_t bswap(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// ------------------------------------------------------------------------------------------------ k_hash2
// Second-generation hash kernel.  Same one-lane-per-message mapping, but
//  (1) the 128-byte chunks are fetched with cp.async (LDGSTS) straight into a double-buffered, XOR-swizzled shared tile:
//      no register staging, and the fetch of chunk i+1 is in flight during the 160 rounds of chunk i by construction;
//  (2) each round is written so that the only operation on the serial a->a chain is rol5(a)+s (one LEA.HI); s = f+e+K+w is
//      formed off the chain;
//  (3) VARIANT 1 forms s with IMADs (multiply by an opaque 1 from the constant bank) so those adds issue on the FMA pipe
//      while LOP3/SHF/LEA keep the ALU pipe.  Measured on B200 (tools/hash_bench.cu, profiles/r1_hash_variants.txt): with one
//      warp per scheduler (10k messages) VARIANT 0 wins (82 us vs 94 us; 8 ALU-pipe instructions per round at 2 cycles
//      each is the floor), with many warps per scheduler (100k messages) VARIANT 1 wins (1.04 vs 0.95 TB/s).
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, uint32_t src_bytes) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t mad1(uint32_t a, uint32_t one, uint32_t c) {
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(c));
  return d;
}

// rol(x, n) on the FMA pipe: x * 2^n as a 64-bit product puts x << n in the low word and x >> (32 - n) in the high word; the
// two halves have no bit in common, so lo * 1 + hi is the rotation.  `pow2` and `one` are opaque (derived from a kernel
// parameter), otherwise ptxas strength-reduces both back to ALU-pipe shifts.
__device__ __forceinline__ uint32_t rol_fma(uint32_t x, uint32_t pow2, uint32_t one) {
  uint32_t r;
  asm("{\n\t.reg .u64 t;\n\t.reg .u32 lo, hi;\n\tmul.wide.u32 t, %1, %2;\n\tmov.b64 {lo, hi}, t;\n\tmad.lo.u32 %0, lo, %3, hi;\n\t}" : "=r"(r) : "r"(x), "r"(pow2), "r"(one));
  return r;
}

// VARIANT: 0 = every round operation on the ALU pipe; 1 = s formed by two IMADs; 2..5 = experiments that move off-chain work
// to the FMA pipe (5: w+K; 2: w+K and rol30(b); 3: w+K and the schedule's rol1; 4: all three) hoping a lone warp would
// alternate pipes.  It does not pay: at 10k messages 0 -> 82 us, 5 -> 96, 2 -> 121, 3 -> 125, 4 -> 143 us; at 100k messages
// only VARIANT 1 beats 0 (345 vs 375 us).  The engine uses 0 (latency regime) and 1 (throughput regime); 2..5 stay for
// tools/hash_bench.cu, which reproduces the table (profiles/r1_hash_variants.txt).
template <int VARIANT>
__device__ __forceinline__ void sha1_rounds2(uint32_t (&w)[16], uint32_t (&h)[5], uint32_t one) {
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
  const uint32_t two = one << 1, two30 = one << 30;
  constexpr bool kFmaRol30 = VARIANT == 2 || VARIANT == 4, kFmaRol1 = VARIANT == 3 || VARIANT == 4;
#pragma unroll
  for (int i = 0; i < 80; i++) {
    uint32_t wi;
    if (i < 16) wi = w[i];
    else {
      const uint32_t x = w[(i - 3) & 15] ^ w[(i - 8) & 15] ^ w[(i - 14) & 15] ^ w[i & 15];
      wi = kFmaRol1 ? rol_fma(x, two, one) : rol(x, 1);
      w[i & 15] = wi;
    }
    const uint32_t k = i < 20 ? 0x5A827999u : (i < 40 ? 0x6ED9EBA1u : (i < 60 ? 0x8F1BBCDCu : 0xCA62C1D6u));
    uint32_t f;
    if (i < 20) f = (b & c) | (~b & d);
    else if (i < 40) f = b ^ c ^ d;
    else if (i < 60) f = (b & c) | (b & d) | (c & d);
    else f = b ^ c ^ d;
    uint32_t s;
    if (VARIANT == 0) s = f + e + (wi + k);            // lone warp per scheduler (latency regime): fewest instructions wins
    else if (VARIANT == 1) s = mad1(f, one, mad1(e, one, wi + k));  // many warps per scheduler (throughput regime): adds on the FMA pipe
    else s = f + e + mad1(wi, one, k);                  // w + K is far off the chain: FMA pipe
    asm volatile("" : "+r"(s));  // keep s a value of its own: the a->a chain below is then a single rol5(a)+s
    uint32_t t = rol(a, 5) + s;
    e = d; d = c; c = kFmaRol30 ? rol_fma(b, two30, one) : rol(b, 30); b = a; a = t;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e;
}

template <int WARPS, int VARIANT>
__global__ void __launch_bounds__(WARPS * 32) k_hash2(const uint8_t *__restrict__ bytes, const uint64_t *__restrict__ off,
                                                      const uint32_t *__restrict__ len32, const uint64_t *__restrict__ off_end,
                                                      uint32_t n, char *__restrict__ out, uint32_t one = 1) {
  KR_TL(7);
  __shared__ uint4 s_tile[2][WARPS][32][8];  // [buffer][warp][message lane][16-byte piece ^ (lane & 7)]
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid-stride over groups of WARPS*32 messages: the engine caps the grid for large n so that the hash leaves room on every
  // SM for the main chain's blocks (each warp owns its shared tile, so the trips need no block-wide barrier)
  for (uint32_t grp_i = blockIdx.x; (uint64_t)grp_i * (WARPS * 32) < n; grp_i += gridDim.x) {
  const uint32_t m = (grp_i * WARPS + warp) * 32 + lane;
  const bool have = m < n;
  uint64_t moff = 0;
  uint32_t mlen = 0;
  if (have) { moff = off[m]; mlen = len32 ? len32[m] : (uint32_t)(off_end[m] - moff); }
  const uint32_t nblocks = have ? (mlen + 8) / 64 + 1 : 0;
  uint32_t max_blocks = nblocks;
#pragma unroll
  for (int d = 16; d; d >>= 1) max_blocks = max(max_blocks, __shfl_xor_sync(0xFFFFFFFFu, max_blocks, d));
  uint32_t h[5] = {0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u};
  const uint32_t sub = lane & 7, grp = lane >> 3;
  // this lane fetches piece `sub` of messages 4r+grp, r = 0..7: keep their base pointers and padded lengths
  const uint8_t *src[8];
  uint32_t lim[8];
#pragma unroll
  for (int r = 0; r < 8; r++) {
    uint32_t sl = 4 * r + grp;
    uint64_t o = __shfl_sync(0xFFFFFFFFu, moff, sl);
    uint32_t l = __shfl_sync(0xFFFFFFFFu, mlen, sl);
    src[r] = bytes + o + sub * 16;
    lim[r] = (l + 15) & ~15u;  // the arena pads every message to 16 bytes
  }
  auto fetch = [&](uint32_t chunk, int buf) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      uint32_t sl = 4 * r + grp;
      uint32_t pos = chunk * 128 + sub * 16;
      bool in = pos < lim[r];
      cp_async16(&s_tile[buf][warp][sl][sub ^ (sl & 7)], in ? (const void *)(src[r] + (size_t)chunk * 128) : (const void *)bytes, in ? 16u : 0u);
    }
    cp_async_commit();
  };
  const uint32_t nchunks = (max_blocks + 1) / 2;
  if (nchunks) fetch(0, 0);
  for (uint32_t chunk = 0; chunk < nchunks; chunk++) {
    const int buf = chunk & 1;
    if (chunk + 1 < nchunks) { fetch(chunk + 1, buf ^ 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncwarp();
#pragma unroll
    for (int half = 0; half < 2; half++) {
      uint32_t blk = chunk * 2 + half;
      if (blk >= nblocks) continue;
      uint32_t w[16];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        uint4 v = s_tile[buf][warp][lane][(half * 4 + q) ^ (lane & 7)];
        w[4 * q] = bswap(v.x); w[4 * q + 1] = bswap(v.y); w[4 * q + 2] = bswap(v.z); w[4 * q + 3] = bswap(v.w);
      }
      const uint32_t bstart = blk * 64;
      if (bstart + 64 > mlen) {  // tail block(s): 0x80, zero fill, 64-bit big-endian bit length (FIPS 180-4 §5.1.1)
#pragma unroll
        for (int q = 0; q < 16; q++) {
          uint32_t wpos = bstart + 4 * q;
          uint32_t v = w[q];
          if (wpos >= mlen) v = (wpos == mlen) ? 0x80000000u : 0u;
          else if (wpos + 4 > mlen) {
            uint32_t keep = mlen - wpos;  // 1..3 message bytes in this word
            v = (v & (0xFFFFFFFFu << (8 * (4 - keep)))) | (0x80u << (8 * (3 - keep)));
          }
          w[q] = v;
        }
        if (blk == nblocks - 1) { w[14] = mlen >> 29; w[15] = mlen << 3; }
      }
      sha1_rounds2<VARIANT>(w, h, one);
    }
    __syncwarp();  // every lane is done reading this buffer before the fetch two iterations ahead overwrites it
  }
  if (have) {
  uint32_t o32[8];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    uint64_t v;
    switch (j) {
      case 0: v = ((uint64_t)h[0] << 8) | (h[1] >> 24); break;
      case 1: v = ((uint64_t)(h[1] & 0xFFFFFFu) << 16) | (h[2] >> 16); break;
      case 2: v = ((uint64_t)(h[2] & 0xFFFFu) << 24) | (h[3] >> 8); break;
      default: v = ((uint64_t)(h[3] & 0xFFu) << 32) | h[4]; break;
    }
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
      uint32_t cc = (uint32_t)(v >> (35 - 5 * kk)) & 31u;
      uint32_t ch = cc < 10 ? ('0' + cc) : ('A' + cc - 10);
      if (kk < 4) lo |= ch << (8 * kk); else hi |= ch << (8 * (kk - 4));
    }
    o32[2 * j] = lo; o32[2 * j + 1] = hi;
  }
  uint4 *dst = reinterpret_cast<uint4 *>(out + 32 * (size_t)m);
  dst[0] = make_uint4(o32[0], o32[1], o32[2], o32[3]);
  dst[1] = make_uint4(o32[4], o32[5], o32[6], o32[7]);
  }
  __syncwarp();
  }  // next group of messages
}

}  // namespace kr
