// rng.cuh — per-path random numbers.
//   REF    : the reference stream — TEA-16 seeding and a 24-bit LCG
//            (MinimalOptiX/utils_device.h:8-34, Camera.cu:24, folkPayload :192-198).
//   PHILOX : Philox4x32-10 keyed by (pixel, launch seed); counter = (draw/4, depth); word draw%4.
// Per-path carried state is ONE int in both modes (`state`): the LCG seed, or the draw counter.
#pragma once
#include "vec.cuh"

MOX_HD uint32_t tea16(uint32_t v0, uint32_t v1) {
  uint32_t s0 = 0;
#pragma unroll
  for (int n = 0; n < 16; n++) {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  return v0;
}

MOX_D void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// MODE is a compile-time parameter of the kernels so the REF instantiation carries a single
// register of generator state.
template <int MODE>
struct RngT;

template <>
struct RngT<0> {  // REF
  int state;
  MOX_D uint32_t next24() {
    state = (int)(1664525u * (uint32_t)state + 1013904223u);
    return (uint32_t)state & 0x00FFFFFFu;
  }
  MOX_D float rnd() { return (float)next24() / (float)0x01000000; }
  // State of the child path spawned at depth `childDepth` (computed from the CURRENT state).
  MOX_D int forkState(int childDepth) const { return (int)tea16((uint32_t)state, (uint32_t)childDepth); }
};

template <>
struct RngT<1> {  // PHILOX
  int state;  // draw counter within the current depth stream
  uint32_t pixel, launchSeed, depth;
  uint32_t c0, c1, c2, c3;
  uint32_t cachedBlock;
  MOX_D uint32_t next24() {
    uint32_t ctr = (uint32_t)state;
    uint32_t blk = ctr >> 2;
    if (blk != cachedBlock) {
      uint32_t o[4];
      philox4x32_10(blk, depth, 0u, 0u, pixel, launchSeed, o);
      c0 = o[0]; c1 = o[1]; c2 = o[2]; c3 = o[3];
      cachedBlock = blk;
    }
    uint32_t w = (ctr & 3u) == 0 ? c0 : (ctr & 3u) == 1 ? c1 : (ctr & 3u) == 2 ? c2 : c3;
    state = (int)(ctr + 1u);
    return w >> 8;
  }
  MOX_D float rnd() { return (float)next24() / (float)0x01000000; }
  MOX_D int forkState(int) const { return 0; }
};

template <int MODE>
MOX_D RngT<MODE> makeRng(int state, uint32_t pixel, uint32_t launchSeed, uint32_t depth);
template <>
MOX_D RngT<0> makeRng<0>(int state, uint32_t, uint32_t, uint32_t) { RngT<0> r; r.state = state; return r; }
template <>
MOX_D RngT<1> makeRng<1>(int state, uint32_t pixel, uint32_t launchSeed, uint32_t depth) {
  RngT<1> r;
  r.state = state; r.pixel = pixel; r.launchSeed = launchSeed; r.depth = depth;
  r.c0 = r.c1 = r.c2 = r.c3 = 0; r.cachedBlock = 0xffffffffu;
  return r;
}

// utils_device.h:36-52 with the draw order pinned x, y, z.
template <class R>
MOX_D float3 randInUnitSphere(R& r) {
  float3 p;
  do {
    float a = r.rnd(), b = r.rnd(), c = r.rnd();
    p = mk3(a, b, c) * 2.0f - mk3(1.f, 1.f, 1.f);
  } while (length(p) >= 1.0f);
  return p;
}
template <class R>
MOX_D float3 randInUnitDisk(R& r) {
  float3 p;
  do {
    float a = r.rnd(), b = r.rnd();
    p = mk3(a, b, 0.f) * 2.0f - mk3(1.f, 1.f, 0.f);
  } while (length(p) >= 1.f);
  return p;
}
