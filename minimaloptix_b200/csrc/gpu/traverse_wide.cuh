// traverse_wide.cuh — persistent-thread traversal of the compressed 8-wide BVH (BvhNode8,
// gpu_types.h).  Same contract as traverseWarpPersistent (traverse.cuh): identical primitive tests,
// (t, id) lexicographic closest hit, order-independent shadow transmittance; only the hierarchy
// walked differs, so results are bit-identical to the binary traversal.
//
// Per lane: a node group G = (childBase, hit bits 31..24 | primitive bits that exist 23..8 | imask 7..0) and a
// primitive group T = (primBase, 16 hit bits), plus a stack of postponed node groups (Ylitie et al. 2017).  One node
// step pops the front-most child of G (highest bit of slot ^ octant order), pushes the rest of G, fetches the 80-byte
// node and tests its 8 quantised child boxes: one FMA per plane on a grid local to the node, near/far planes picked
// per ray sign for four children at a time, one more FMA + IMAD per child for its hit bits (fixed positions per slot,
// gpu_types.h).  Warp phase voting as in the binary kernel — a node step or a primitive step (up to two primitives of
// the lane) per iteration — but weighted by cost (MOX_VOTE_TRI_WEIGHT): the cheaper primitive phase runs as soon as a
// third as many lanes want it.
#pragma once
#include "traverse.cuh"

#define MOX_WIDE_STACK MOX_TRAVERSAL_STACK

// Four small savings of the node / primitive step, adopted together at the very end of round 2 (1 662 -> 1 720 Mrays/s in
// one A/B call, all GPU tests green on that library); MOX_NODE_STEP_R2A restores the forms measured until then:
//   MOX_ADDEND_DIRECTED   PRMT-plane addends by directed rounding instead of an extra error bound (planeTerms)
//   MOX_BFIND             highest set bit by FLO directly
//   MOX_SWAP_PREDICATED   the octant swaps predicated on the signs of 1/d the node step tests anyway
//   MOX_ONE_FMA           the 1.0f pattern of the PRMT conversion from one FMA
#ifndef MOX_NODE_STEP_R2A
#define MOX_ADDEND_DIRECTED
#define MOX_BFIND
#define MOX_SWAP_PREDICATED
#define MOX_ONE_FMA
#endif


// Quantised plane byte -> float, two ways, mixed per plane so that neither pipe is the bottleneck:
//  * I2F.U8 on the conversion pipe (15.4 results/clk/SM measured, scripts/microbench/pipe_rates.cu, against 124
//    FFMA): with all 48 conversions of a node step on it that pipe was ~70 % busy and 15-19 % of the stall samples
//    were mio_throttle;
//  * one PRMT on the ALU pipe: m = 1 + b * 2^-15 — byte b dropped into mantissa bits 8..15 of 1.0f — with the
//    affine map folded back into the plane FMA:   b * ia + on  ==  m * (2^15 ia) + (on - 2^15 ia).
//    The addend's rounding error (<= 2^-24 of 2^15 |ia|, i.e. 2^-9 of one quantisation step) is added to the
//    conservative margin, so boxes only ever grow.
// MOX_BYTE_MIX: bit 0..2 = near plane x, y, z, bit 3..5 = far plane x, y, z; a set bit converts that plane's eight
// bytes with PRMT.  Measured (1 M-triangle bench, Mrays/s, with the select form of the hit mask): 0 (all I2F) 1 582,
// 0x3f (all PRMT) 1 599, 0x3b 1 591, 0x38 (far planes) 1 617, 0x39 1 625, 0x30 1 633, 0x18 1 632, 0x10 1 616;
// with MOX_HIT_SIGN=1: 0x3f 1 612, 0x10 1 627, 0x39 1 632, 0x30 1 645, 0x28 1 647, 0x33 1 648 (0x38 and 0x31 spill).
// MOX_HIT_SIGN: how a child's box test becomes its hit bits.  0: FSETP + SEL of the child's constant, OR-ed three at
// a time (all on the ALU pipe, which also runs the 32 FMNMX, the PRMTs and the near/far selects of a node step).
// 1 (default): the far-side widening and the comparison are one FMA, tf * 1.00001 - tn, whose sign bit times the
// child's constant is subtracted from an all-hit mask by one IMAD — FFMA + SHF + IMAD instead of FMUL + FSETP + SEL +
// half an IADD3, and only the shift is on the ALU pipe.  2: the same with an arithmetic shift and one LOP3 (1 588).
// 3: no ALU-pipe instruction at all — s = saturate(dd * -3e38) is 1.0 for a miss and 0.0 for a hit (FMUL.SAT,
// flush-to-zero so that a denormal difference counts as a hit), and the mask is accumulated as a float, all 24 bits
// minus s * constant per child (integers below 2^24: exact), and converted once per node.
// (the default is set in gpu_types.h: the builder lays out the node's valid word to match)
#ifndef MOX_BYTE_MIX
#ifdef MOX_BYTE_PRMT
#define MOX_BYTE_MIX 0x3f
#else
#define MOX_BYTE_MIX 0x33
#endif
#endif
// `one` is 1.0f's bit pattern held in a register (see traverseWidePersistent): with the constant as an immediate
// ptxas needs the byte selector in a register and re-materialises four selectors per node.
template <bool PRMT>
MOX_D float byteToFloat(uint32_t w, int i, uint32_t one) {
  if (!PRMT) return (float)((w >> (8 * i)) & 0xffu);
  uint32_t r;
  if (i == 0) asm("prmt.b32 %0, %1, %2, 0x7604;" : "=r"(r) : "r"(w), "r"(one));
  else if (i == 1) asm("prmt.b32 %0, %1, %2, 0x7614;" : "=r"(r) : "r"(w), "r"(one));
  else if (i == 2) asm("prmt.b32 %0, %1, %2, 0x7624;" : "=r"(r) : "r"(w), "r"(one));
  else asm("prmt.b32 %0, %1, %2, 0x7634;" : "=r"(r) : "r"(w), "r"(one));
  return __uint_as_float(r);
}
// Scale and addends of one axis for the plane FMAs  t = float(byte) * scale + addend  (near: rounded down, far:
// rounded up): every axis gets its own rounding bound, 2^-21 of the local origin term (+ 2^-22 of the PRMT
// form's 2^15 ia), folded into the addends so a plane costs one conversion and one FMA.
template <bool NEAR_PRMT, bool FAR_PRMT>
MOX_D void planeTerms(float ia, float oa, float& sn, float& an, float& sf, float& af) {
  const float k = ia * 32768.f;
  // (I2F form: |oa| * 2^-21 is exact, so one FMA rounds the same sum once)
  const float an0 = fmaf(-4.76837158203125e-07f, fabsf(oa), oa), af0 = fmaf(4.76837158203125e-07f, fabsf(oa), oa);
#ifdef MOX_ADDEND_DIRECTED
  // PRMT form: the only new rounding is the subtraction of 2^15 ia from the addend — rounded towards the outside of the
  // box (FADD.RM / FADD.RP) it needs no bound of its own: two instructions per plane instead of four
  if (NEAR_PRMT) { sn = k; an = __fsub_rd(an0, k); } else { sn = ia; an = an0; }
  if (FAR_PRMT) { sf = k; af = __fsub_ru(af0, k); } else { sf = ia; af = af0; }
#else
  const float e = fmaf(2.384185791015625e-07f, fabsf(k), 4.76837158203125e-07f * fabsf(oa));
  if (NEAR_PRMT) { sn = k; an = (oa - e) - k; } else { sn = ia; an = an0; }
  if (FAR_PRMT) { sf = k; af = (oa + e) - k; } else { sf = ia; af = af0; }
#endif
}
// index of the highest set bit (x != 0): FLO directly — written as 31 - __clz(x), ptxas keeps both subtractions
MOX_D uint32_t highestBit(uint32_t x) {
#ifdef MOX_BFIND
  uint32_t r;
  asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
  return r;
#else
  return 31u - (uint32_t)__clz(x);
#endif
}

// (a & m) | (b & ~m) as one LOP3 (written as two ANDs and an OR, ptxas spends two)
MOX_D uint32_t bitSelect(uint32_t a, uint32_t b, uint32_t m) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(r) : "r"(a), "r"(b), "r"(m));
  return r;
}

// CLASSIFY (closest hit of the render path): the hit record is (t, primitive id | shade class << 28) — the class
// comes from the winning primitive's packed record — and beta / gamma are not carried (two registers less per
// lane); the shade kernels recompute them from the same operands.  The raw query keeps the four-word record.
// WT: the packed triangle records hold the raw vertices and the watertight test runs (MOX_ACCEL_WATERTIGHT).
template <bool ANYHIT, bool COUNT, bool CLASSIFY = false, bool WT = false>
__device__ __forceinline__ void traverseWidePersistent(const SceneView& s, const TraceJob& job) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned ltMask = (1u << lane) - 1u;
  const uint32_t jobCount = job.countPtr ? __ldg(job.countPtr) : job.count;
#ifdef MOX_TOP_SMEM
  // Measured option: the first MOX_TOP_SMEM nodes — the array is in level order, 73 = root + two
  // levels — staged once per CTA in shared memory and read from there instead of through L1.
  __shared__ float4 sTop[MOX_TOP_SMEM * 5];
  {
    const float4* src = (const float4*)s.nodes8;
    const uint32_t nTop = min((uint32_t)MOX_TOP_SMEM, s.nNodes8) * 5u;
    for (uint32_t k = threadIdx.x; k < nTop; k += blockDim.x) sTop[k] = __ldg(src + k);
    __syncthreads();
  }
#endif
  // 0x3f800000 that ptxas cannot fold into an immediate (a launch never has 2^31 rays)
#ifndef MOX_ONE_FMA
  const uint32_t one = 0x3f800000u | (job.count >> 31);
#endif
  // Per-lane state.  A lane is busy exactly while it has a primitive group or a node group pending (a ray with
  // neither pops its stack or finishes in the same iteration), so there is no separate "active" flag.  CLASSIFY keeps
  // the winner's shade class in bits 28..30 of bPrim (the form the hit record has anyway).  A shadow ray's
  // transmittance lives in the last two entries of the lane's stack array (local memory, no L1 footprint until
  // touched): it is needed only when a ray meets tinting glass.  (Shared memory for it cost L1: the carve-out is
  // permanent; 2.5 KB per CTA lowered the shadow kernel's L1 hit rate by a point.)
  uint2 stack[MOX_WIDE_STACK];
  int sp = 0;
  uint32_t gBase = 0, gBits = 0;  // node group
  uint32_t tBase = 0, tBits = 0;  // primitive group
  bool exhausted = false;
  uint32_t rayId = 0, octinv = 0;
  float3 o = mk3(0.f), d = mk3(0.f), idir = mk3(0.f);
  float tmin = 0.f, tBest = 0.f, bBeta = 0.f, bGamma = 0.f;
  int bPrim = -1;
  bool tinted = false;               // ANYHIT: the top of the stack array holds a product of glass colours for this ray
  uint32_t shadowEnd = 0;            // ANYHIT && COUNT: 1 blocked, 2 tinted
  uint32_t nv = 0, np = 0;
#define MOX_LANE_BUSY() (tBits != 0u || (gBits & 0xff000000u) != 0u)
  WtRay wr;
  wr.kx = wr.ky = wr.kz = 0; wr.Sx = wr.Sy = wr.Sz = 0.f;

  // Shadow rays (CHUNKED): ray ids are claimed from the global cursor 32 at a time and kept one per lane (poolId); a
  // refill hands them out by shuffle.  The claim — one atomic, then one coalesced queue load whose result nobody
  // touches yet — is issued when a pool runs dry, so what a refill waits for is one memory level (the ray itself)
  // instead of three (cursor, queue entry, ray); the whole warp sits in that wait, busy lanes included.  Measured:
  // shadow 90.6 -> 90.0 ms per step; the closest-hit kernel lost with it (69.3 -> 70.1 ms) and keeps one atomic
  // per refill.
#ifdef MOX_REFILL_PER_EVENT
  constexpr bool CHUNKED = false;
#elif defined(MOX_REFILL_CHUNK_ALL)
  constexpr bool CHUNKED = true;
#else
  constexpr bool CHUNKED = ANYHIT;
#endif
  uint32_t poolId = 0, poolPos = 0, poolCount = 0;   // poolPos / poolCount are warp-uniform
  bool moreChunks = true;
  while (true) {
    // ---------------- refill idle lanes
    if (!exhausted) {
      const bool idleLane = !MOX_LANE_BUSY();
      unsigned idle = __ballot_sync(FULL, idleLane);
      if (idle) {
        bool take;
        uint32_t base = 0;
        if (CHUNKED) {
          if (poolPos >= poolCount && moreChunks) {   // first visit, or the last refill emptied the pool
            if (lane == 0) base = atomicAdd(job.cursor, 32u);
            base = __shfl_sync(FULL, base, 0);
            poolCount = base < jobCount ? min(32u, jobCount - base) : 0u;
            poolPos = 0;
            if ((uint32_t)lane < poolCount) poolId = job.queue ? MOX_LD_STREAM(job.queue + base + lane) : base + lane;
            if (base + 32u >= jobCount) moreChunks = false;
          }
          const uint32_t slot = poolPos + (uint32_t)__popc(idle & ltMask);
          const uint32_t got = __shfl_sync(FULL, poolId, slot & 31u);
          take = idleLane && slot < poolCount;
          poolPos = min(poolPos + (uint32_t)__popc(idle), poolCount);
          if (take) rayId = got;
        } else {
          const int leader = __ffs(idle) - 1;
          if (lane == leader) base = atomicAdd(job.cursor, (uint32_t)__popc(idle));
          base = __shfl_sync(FULL, base, leader);
          const uint32_t i = base + __popc(idle & ltMask);
          take = idleLane && i < jobCount;
          if (take) rayId = job.queue ? MOX_LD_STREAM(job.queue + i) : i;
        }
        if (take) {
          const uint32_t oId = originIndex(job, rayId);
          float4 ro = MOX_LD_STREAM(job.rayO + oId), rd = MOX_LD_STREAM(job.rayD + rayId);
          if (!(ANYHIT && rd.w < 0.f)) {
            RayPre r = prepRay(mk3(ro), mk3(rd), ro.w);
            o = r.o; d = r.d; idir = r.idir; tmin = r.tmin;
            if (WT) wr = wtPrep(d);
#ifdef MOX_SWAP_PREDICATED
            octinv = 7u ^ ((idir.x < 0.f ? 4u : 0u) | (idir.y < 0.f ? 2u : 0u) | (idir.z < 0.f ? 1u : 0u));   // (-0.0 counts as negative here)
#else
            octinv = 7u ^ ((d.x < 0.f ? 4u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 1u : 0u));
#endif
            tBest = rd.w; bPrim = -1;
            if (!CLASSIFY) { bBeta = 0.f; bGamma = 0.f; }
            tinted = false;
            sp = 0;
            gBase = 0; gBits = 0x80000000u;  // root: one pending child, imask 0 -> node index 0
            tBase = 0; tBits = 0;
            if (COUNT) { nv = 0; np = 0; shadowEnd = 0; }
          }
        }
        if (CHUNKED) {
          if (poolPos >= poolCount) {
            if (moreChunks) {
              // claim the next 32 now: the queue load is in flight while the warp goes back to traversing
              uint32_t nb = 0;
              if (lane == 0) nb = atomicAdd(job.cursor, 32u);
              nb = __shfl_sync(FULL, nb, 0);
              poolCount = nb < jobCount ? min(32u, jobCount - nb) : 0u;
              poolPos = 0;
              if ((uint32_t)lane < poolCount) poolId = job.queue ? MOX_LD_STREAM(job.queue + nb + lane) : nb + lane;
              if (nb + 32u >= jobCount) moreChunks = false;
            }
            if (poolPos >= poolCount && !moreChunks) exhausted = true;
          }
        } else if (base + __popc(idle) >= jobCount) exhausted = true;
      }
    }
    if (!__any_sync(FULL, MOX_LANE_BUSY())) {
      if (exhausted) break;
      continue;
    }
    // ---------------- traverse until too few lanes are busy
    while (true) {
      const bool isTri = tBits != 0u;
      const bool isNode = !isTri && (gBits & 0xff000000u) != 0u;
#ifndef MOX_VOTE_TRI_WEIGHT
#define MOX_VOTE_TRI_WEIGHT 3  // a primitive step costs less than half a node step: vote by cost, not by head count (round 1: 1: 1115, 2: 1191, 3: 1206, 4: 1214, 6: 1207, 32: 1069 Mrays/s; with the cheaper node step of round 2 and two primitives per step: 2: 1693-1696, 3: 1691-1692, 4: 1671)
#endif
#ifndef MOX_VOTE_POPC
      // one ballot + one warp reduction (REDUX) of the weighted vote instead of two ballots + two POPCs — POPC shares the
      // slow conversion pipe with the 48 I2F of a node step (ncu: 15-19 % of the stall samples are mio_throttle):
      // extend 71.6 -> 71.0 ms, shadow 92.2 -> 91.7 ms per step (MOX_VOTE_POPC restores the counted form)
      const unsigned bm = __ballot_sync(FULL, isNode | isTri);
      const int score = __reduce_add_sync(FULL, isNode ? 1 : (isTri ? -MOX_VOTE_TRI_WEIGHT : 0));
      if (bm == 0u || (!exhausted && __popc(bm) < job.fetchThreshold)) break;
      if (score >= 0) {
#else
      const unsigned nm = __ballot_sync(FULL, isNode), tm = __ballot_sync(FULL, isTri);
      const int nNode = __popc(nm), nTri = __popc(tm);  // disjoint masks: busy lanes = nNode + nTri (POPC shares the slow conversion pipe)
      if ((nm | tm) == 0u || (!exhausted && nNode + nTri < job.fetchThreshold)) break;
      if (nNode >= MOX_VOTE_TRI_WEIGHT * nTri) {
#endif
        if (isNode) {
          // ---- pop the front-most pending child of G
          const uint32_t bit = highestBit(gBits & 0xff000000u);
          const uint32_t imaskG = gBits & 0xffu;
          gBits &= ~(1u << bit);
          const uint32_t slot = (bit - 24u) ^ octinv;
          const uint32_t nodeIdx = gBase + __popc(imaskG & ((1u << slot) - 1u));
          if (gBits & 0xff000000u) stack[sp++] = make_uint2(gBase, gBits);
          // ---- fetch and test the node
#ifdef MOX_TOP_SMEM
          float4 n0, n1, n2, n3, n4;
          if (nodeIdx < (uint32_t)MOX_TOP_SMEM) {
            const float4* t = sTop + nodeIdx * 5u;
            n0 = t[0]; n1 = t[1]; n2 = t[2]; n3 = t[3]; n4 = t[4];
          } else {
            const BvhNode8* nd = s.nodes8 + nodeIdx;
            n0 = __ldg(&nd->n0); n1 = __ldg(&nd->n1); n2 = __ldg(&nd->n2); n3 = __ldg(&nd->n3); n4 = __ldg(&nd->n4);
          }
#else
          const BvhNode8* nd = s.nodes8 + nodeIdx;
          const float4 n0 = __ldg(&nd->n0), n1 = __ldg(&nd->n1), n2 = __ldg(&nd->n2), n3 = __ldg(&nd->n3), n4 = __ldg(&nd->n4);
#endif
          if (COUNT) nv++;
          const uint32_t ew = __float_as_uint(n0.w);
          const float iax = __uint_as_float((ew & 0xffu) << 23) * idir.x;
          const float iay = __uint_as_float(((ew >> 8) & 0xffu) << 23) * idir.y;
          const float iaz = __uint_as_float(((ew >> 16) & 0xffu) << 23) * idir.z;
          const float oax = (n0.x - o.x) * idir.x, oay = (n0.y - o.y) * idir.y, oaz = (n0.z - o.z) * idir.z;
#ifdef MOX_ONE_FMA
          // ... or one FMA per node step (0 * finite + 1; neither nvcc nor ptxas may fold it) instead of a constant
          // load and an ALU instruction to rebuild the value ptxas does not keep in a register across the loop
          const uint32_t one = __float_as_uint(fmaf(0.f, idir.x, 1.0f));
#endif
          // conservative plane terms (planeTerms above); the far side is additionally widened by 1e-5 relative
          constexpr unsigned MIX = MOX_BYTE_MIX;
          float snx, onx, sfx, ofx, sny, ony, sfy, ofy, snz, onz, sfz, ofz;
          planeTerms<(MIX & 1u) != 0, (MIX & 8u) != 0>(iax, oax, snx, onx, sfx, ofx);
          planeTerms<(MIX & 2u) != 0, (MIX & 16u) != 0>(iay, oay, sny, ony, sfy, ofy);
          planeTerms<(MIX & 4u) != 0, (MIX & 32u) != 0>(iaz, oaz, snz, onz, sfz, ofz);
#if !defined(MOX_NODE_META) && MOX_HIT_SIGN == 3
          float hitf = 16777215.f;   // inner children in bits 16..23 until the mask is an integer again
          uint32_t hitmask;
#elif !defined(MOX_NODE_META) && MOX_HIT_SIGN
          uint32_t hitmask = 0xff00ffffu;
#else
          uint32_t hitmask = 0;
#endif
#pragma unroll
          for (int half = 0; half < 2; ++half) {
#ifdef MOX_NODE_META
            const uint32_t meta4 = __float_as_uint(half ? n1.w : n1.z);
            // four children at once: inner children (position field 24..31 = 0b11xxx) get their slot
            // XOR-ed with the ray's octant mask, which orders them front to back
            const uint32_t inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
            const uint32_t pos4 = (meta4 ^ ((inner4 >> 4) * octinv)) & 0x1f1f1f1fu;
            const uint32_t cnt4 = (meta4 >> 5) & 0x07070707u;
#endif
            const uint32_t qlx = __float_as_uint(half ? n2.y : n2.x), qly = __float_as_uint(half ? n2.w : n2.z);
            const uint32_t qlz = __float_as_uint(half ? n3.y : n3.x), qhx = __float_as_uint(half ? n3.w : n3.z);
            const uint32_t qhy = __float_as_uint(half ? n4.y : n4.x), qhz = __float_as_uint(half ? n4.w : n4.z);
            const uint32_t nx = idir.x < 0.f ? qhx : qlx, fx = idir.x < 0.f ? qlx : qhx;
            const uint32_t ny = idir.y < 0.f ? qhy : qly, fy = idir.y < 0.f ? qly : qhy;
            const uint32_t nz = idir.z < 0.f ? qhz : qlz, fz = idir.z < 0.f ? qlz : qhz;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float tnx = fmaf(byteToFloat<(MIX & 1u) != 0>(nx, i, one), snx, onx), tfx = fmaf(byteToFloat<(MIX & 8u) != 0>(fx, i, one), sfx, ofx);
              const float tny = fmaf(byteToFloat<(MIX & 2u) != 0>(ny, i, one), sny, ony), tfy = fmaf(byteToFloat<(MIX & 16u) != 0>(fy, i, one), sfy, ofy);
              const float tnz = fmaf(byteToFloat<(MIX & 4u) != 0>(nz, i, one), snz, onz), tfz = fmaf(byteToFloat<(MIX & 32u) != 0>(fz, i, one), sfz, ofz);
              const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
#if !defined(MOX_NODE_META) && MOX_HIT_SIGN
              // hit <=> tf * 1.00001 - tn >= 0, read off the sign bit of one FMA (the FMA pipe has room, the ALU pipe
              // that runs FSETP / SEL does not): every miss clears the child's constant from an all-hit mask
              const float dd = fmaf(fminf(fminf(tfx, tfy), fminf(tfz, tBest)), 1.00001f, -tn);
              const uint32_t K = (1u << (24 + 4 * half + i)) | (3u << (2 * (4 * half + i)));
#if MOX_HIT_SIGN == 1
              asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hitmask) : "r"(__float_as_uint(dd) >> 31), "r"(0u - K));
#elif MOX_HIT_SIGN == 3
              float miss;
              asm("mul.ftz.sat.f32 %0, %1, %2;" : "=f"(miss) : "f"(dd), "f"(-3.0e38f));
              hitf = fmaf(miss, -(float)((1u << (16 + 4 * half + i)) | (3u << (2 * (4 * half + i)))), hitf);
#else
              hitmask &= ~((uint32_t)(__float_as_int(dd) >> 31) & K);
#endif
#else
              const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tBest)) * 1.00001f;
#endif
#ifdef MOX_NODE_META
              // branch-free: an empty slot has count bits 0 (and an inverted box)
              const uint32_t bitsI = tn <= tf ? ((cnt4 >> (8 * i)) & 0xffu) : 0u;
              hitmask |= bitsI << ((pos4 >> (8 * i)) & 0xffu);
#elif !MOX_HIT_SIGN
              // fixed positions per slot: one select of a compile-time constant per child (gpu_types.h)
              hitmask |= tn <= tf ? ((1u << (24 + 4 * half + i)) | (3u << (2 * (4 * half + i)))) : 0u;
#endif
            }
          }
          gBase = __float_as_uint(n1.x);
          tBase = __float_as_uint(n1.y);
#ifdef MOX_NODE_META
          gBits = (hitmask & 0xff000000u) | (ew >> 24);
          tBits = hitmask & 0x00ffffffu;
#else
          // keep what exists (inner children: imask, primitives: V), then bring the inner byte into front-to-back
          // order: bit 24 + s moves to 24 + (s ^ octinv) — three conditional swaps (nibbles, pairs, neighbours) written
          // as shifts by 0 or 4 / 2 / 1, so there is no branch and no select: (x << 0 & m) | (x >> 0 & ~m) = x
#if MOX_HIT_SIGN == 3
          hitmask = __float2uint_rz(hitf) & __float_as_uint(n1.z);   // n1.z = imask << 16 | V in this build
          tBits = hitmask & 0x0000ffffu;
          hitmask <<= 8;
#else
          hitmask &= __float_as_uint(n1.z);
          tBits = hitmask & 0x0000ffffu;
#endif
          {
            uint32_t x = hitmask;
#ifdef MOX_SWAP_PREDICATED
            // the octant mask is made of the signs of 1/d (see the refill), the predicates the near/far selects of this
            // node step hold anyway: three predicated instructions per swap
            if (!(idir.x < 0.f)) x = bitSelect(x << 4, x >> 4, 0xf0000000u);
            if (!(idir.y < 0.f)) x = bitSelect(x << 2, x >> 2, 0xcc000000u);
            if (!(idir.z < 0.f)) x = bitSelect(x << 1, x >> 1, 0xaa000000u);
#else
            const uint32_t s4 = octinv & 4u, s2 = octinv & 2u, s1 = octinv & 1u;
            x = bitSelect(x << s4, x >> s4, 0xf0000000u);
            x = bitSelect(x << s2, x >> s2, 0xcc000000u);
            x = bitSelect(x << s1, x >> s1, 0xaa000000u);
#endif
            gBits = (x & 0xff000000u) | __float_as_uint(n1.w);   // n1.w = V << 8 | imask
          }
#endif
        }
      } else {
#ifndef MOX_PRIMS_PER_STEP
#define MOX_PRIMS_PER_STEP 2
#endif
        // ---- up to MOX_PRIMS_PER_STEP primitives of the current group.  Leaf children hold up to two primitives and a
        // ray often hits the boxes of several, so a lane that is in the primitive phase usually has more than one bit
        // pending: testing two per step halves the vote / loop overhead per primitive, and (MOX_PRIM_PREFETCH) both
        // records are fetched before the first test, so the second fetch hides behind the first test.
        // Measured (Mrays/s): 1 per step 1 656, 2 per step 1 691 (vote weight 3) / 1 696 (weight 2).
        auto recordOf = [&](uint32_t k) -> const float4* {
#ifdef MOX_NODE_META
          return s.packed8 + (size_t)(tBase + k) * MOX_PACKED_F4;
#else
          // offset from the node's first primitive = rank of bit k among the primitive bits that exist (V sits in
          // bits 8..23 of the node-group word for as long as primitives of that node are pending)
          return s.packed8 + (size_t)(tBase + __popc((gBits >> 8) & ~(0xffffffffu << k))) * MOX_PACKED_F4;
#endif
        };
        // one primitive test; returns true when a shadow ray was blocked (the lane has nothing pending any more)
        auto testPrimitive = [&](const float4 r0, const float4 r1, const float4 r2) -> bool {
          bool blockedNow = false;
          if (COUNT) np++;
          const uint32_t idbits = __float_as_uint(r0.w);
          const uint32_t type = (idbits >> 30) | __float_as_uint(r2.w) | (__float_as_uint(r1.w) >> 8);
          const int id = (int)(idbits & 0x3fffffffu);
          float t = 0.f, be = 0.f, ga = 0.f;
          bool hit;
          // (t, id) rule: an equal t wins only against a hit with a higher id (CLASSIFY: the id is the low 28 bits)
#define MOX_TIE_WINS() (!ANYHIT && (CLASSIFY ? (bPrim >= 0 && id < (int)((uint32_t)bPrim & MOX_HIT_ID_MASK)) : id < bPrim))
          if (type == PT_TRI) {
            hit = (WT ? triTestWt(wr, o, tmin, mk3(r0), mk3(r1), mk3(r2), t, be, ga) : triTest(o, d, tmin, mk3(r0), mk3(r1), mk3(r2), t, be, ga)) &&
                  (t < tBest || (t == tBest && MOX_TIE_WINS()));
          } else if (type == PT_SPHERE) {
            hit = sphereTest(make_float4(r1.x, r1.y, r1.z, r2.x), o, d, tmin, tBest, MOX_TIE_WINS(), t);
          } else {
            const Analytic* an = s.analytic + __float_as_int(r0.x);
            Analytic q;
            q.a = __ldg(&an->a); q.b = __ldg(&an->b); q.c = __ldg(&an->c); q.d = __ldg(&an->d);
            hit = quadTest(q, o, d, tmin, t, be, ga) && (t < tBest || (t == tBest && MOX_TIE_WINS()));
          }
#undef MOX_TIE_WINS
          if (hit) {
            if (ANYHIT) {
              // shadow class from the record itself (k_pack): only tinting glass still needs its material
              const uint32_t cls = __float_as_uint(r1.w) & 3u;
              if (cls == MOX_SHADOW_BLOCKS) {
                // blocked: zero the contribution without reading it, drop all pending work (the lane is idle from here)
                MOX_ST_STREAM(job.shC + rayId, make_float4(0.f, 0.f, 0.f, 0.f));
                tinted = false;
                tBits = 0u; gBits = 0u; sp = 0;
                blockedNow = true;
                if (COUNT) shadowEnd = 1u;
              } else if (cls == MOX_SHADOW_TINTS) {
                const GpuMaterial* m = s.mats + (__ldg(&s.prims[id].typeMat) >> 2);
                const float3 col = mk3(__ldg(&m->dis.color.x), __ldg(&m->dis.color.y), __ldg(&m->dis.color.z));
                float3 a = col;   // = (1, 1, 1) * col
                if (tinted) {
                  const uint2 p0 = stack[MOX_WIDE_STACK - 2], p1 = stack[MOX_WIDE_STACK - 1];
                  a = mk3(__uint_as_float(p0.x) * col.x, __uint_as_float(p0.y) * col.y, __uint_as_float(p1.x) * col.z);
                }
                stack[MOX_WIDE_STACK - 2] = make_uint2(__float_as_uint(a.x), __float_as_uint(a.y));
                stack[MOX_WIDE_STACK - 1] = make_uint2(__float_as_uint(a.z), 0u);
                tinted = true;
              }
            } else {
              tBest = t;
              if (CLASSIFY) bPrim = (int)((uint32_t)id | ((__float_as_uint(r1.w) >> MOX_CLASS_SHIFT) << MOX_HIT_ID_BITS));
              else { bPrim = id; bBeta = be; bGamma = ga; }
            }
          }
          return blockedNow;
        };
#if MOX_PRIMS_PER_STEP == 2 && defined(MOX_PRIM_PREFETCH)
        if (isTri) {
          const uint32_t k1 = highestBit(tBits);
          const float4* recA = recordOf(k1);
          tBits &= ~(1u << k1);
          const bool two = tBits != 0u;
          const uint32_t k2 = two ? highestBit(tBits) : k1;
          const float4* recB = recordOf(k2);
          tBits &= ~(1u << k2);
          // All words are fetched before a type is known (see below)
          const float4 a0 = __ldg(recA), a1 = __ldg(recA + 1), a2 = __ldg(recA + 2);
          const float4 b0 = __ldg(recB), b1 = __ldg(recB + 1), b2 = __ldg(recB + 2);
          const bool blockedA = testPrimitive(a0, a1, a2);
          if (two && !blockedA) testPrimitive(b0, b1, b2);
        }
#else
#pragma unroll
        for (int rep = 0; rep < MOX_PRIMS_PER_STEP; ++rep)
        if (rep == 0 ? isTri : tBits != 0u) {
          const uint32_t k = highestBit(tBits);
          const float4* rec = recordOf(k);
          tBits &= ~(1u << k);
          // All three words are fetched before the type is known: the tag is folded from words 0 and 2 and the
          // (zero) high bits of word 1, so the loads stay together ahead of the branch — waiting for word 0
          // first would put a second memory latency into every triangle test.
          const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1), r2 = __ldg(rec + 2);
          testPrimitive(r0, r1, r2);
        }
#endif
      }
      // ---- out of work in the current groups: resume a postponed node group, or finish the ray.  Only a lane that
      // was busy when this iteration started can have run dry in it.
      if ((isTri || isNode) && !MOX_LANE_BUSY()) {
        if (sp > 0) {
          const uint2 g = stack[--sp];
          gBase = g.x; gBits = g.y;
        } else {
          if (ANYHIT) {
            // unoccluded: the contribution stays as the shade kernel wrote it; blocked: already zeroed at the hit;
            // only a ray tinted by glass needs the read-modify-write (the load stalls the whole warp)
            if (tinted) {
              const uint2 p0 = stack[MOX_WIDE_STACK - 2], p1 = stack[MOX_WIDE_STACK - 1];
              float4 c = MOX_LD_STREAM(job.shC + rayId);
              MOX_ST_STREAM(job.shC + rayId, make_float4(c.x * __uint_as_float(p0.x), c.y * __uint_as_float(p0.y), c.z * __uint_as_float(p1.x), c.w));
              tinted = false;
              if (COUNT) shadowEnd = 2u;
            }
          } else if (CLASSIFY) {
            MOX_ST_STREAM(job.hits2 + rayId, make_float2(tBest, __int_as_float(bPrim)));   // -1 = MOX_HIT_MISS
          } else {
            MOX_ST_STREAM(job.hits + rayId, make_float4(tBest, __int_as_float(bPrim), bBeta, bGamma));
          }
          if (COUNT) {
            atomicAdd((unsigned long long*)(job.counters + (ANYHIT ? 16 : 10)), (unsigned long long)nv);
            atomicAdd((unsigned long long*)(job.counters + (ANYHIT ? 18 : 12)), (unsigned long long)np);
            if (ANYHIT && shadowEnd) atomicAdd(job.counters + (shadowEnd == 1u ? 20 : 21), 1u);   // C_SH_BLOCKED / C_SH_TINTED
          }
        }
      }
    }
  }
#undef MOX_LANE_BUSY
}
