"""The bit arithmetic of the fixed-slot wide-BVH node (gpu_types.h::BvhNode8, traverse_wide.cuh), restated with Python
integers and checked exhaustively where the domain is small.  The GPU kernels are covered by the parity tests (the
8-wide, the binary and the brute-force traversals must agree bit for bit); these tests pin the identities the node
step relies on, so that a change of the layout fails here first, on the CPU:
  * the three branch-free conditional swaps move hit bit 24 + s to 24 + (s ^ octinv) for every mask and octant;
  * the rank of a primitive bit among the bits that exist is its offset from primBase;
  * subtracting a missed child's constant from the all-hit mask (the IMAD form) equals OR-ing the hit children's
    constants (the select form), and masking with the node's valid word leaves exactly the children that exist;
  * the byte -> float conversion by PRMT (m = 1 + b * 2^-15) folded into the plane FMA is the affine map it replaces,
    within the rounding bound the kernel adds to its conservative margin."""
import itertools
import struct

import numpy as np
import pytest

M32 = 0xFFFFFFFF


def bit_select(a, b, m):
    """lop3 0xE4: (a & m) | (b & ~m)"""
    return ((a & m) | (b & ~m)) & M32


def octant_order(hitmask, octinv):
    """traverse_wide.cuh: the inner byte of the hit mask brought into front-to-back order."""
    s4, s2, s1 = octinv & 4, octinv & 2, octinv & 1
    x = hitmask
    x = bit_select((x << s4) & M32, x >> s4, 0xF0000000)
    x = bit_select((x << s2) & M32, x >> s2, 0xCC000000)
    x = bit_select((x << s1) & M32, x >> s1, 0xAA000000)
    return x & 0xFF000000


def child_constant(slot):
    return (1 << (24 + slot)) | (3 << (2 * slot))


@pytest.mark.parametrize("octinv", range(8))
def test_conditional_swaps_are_the_xor_permutation(octinv):
    rng = np.random.default_rng(octinv)
    for inner in range(256):
        low = int(rng.integers(0, 1 << 24))               # primitive bits and whatever else sits below: must not leak
        want = 0
        for s in range(8):
            if inner >> s & 1:
                want |= 1 << (24 + (s ^ octinv))
        assert octant_order((inner << 24) | low, octinv) == want


def test_front_most_child_is_the_highest_bit():
    # slots are assigned by centroid octant; a ray with octant mask o visits slot (7 ^ o) ^ ... first: popping the highest
    # bit of the permuted byte and XOR-ing back gives the slot, as the kernel does
    for octinv, inner in itertools.product(range(8), range(1, 256)):
        g = octant_order(inner << 24, octinv)
        bit = g.bit_length() - 1
        slot = (bit - 24) ^ octinv
        assert inner >> slot & 1
        assert all((s ^ octinv) <= (slot ^ octinv) for s in range(8) if inner >> s & 1)


def test_primitive_rank_is_the_offset_from_prim_base():
    rng = np.random.default_rng(7)
    for _ in range(2000):
        counts = rng.integers(0, 3, 8)                    # 0 = inner or empty slot, 1..2 primitives of a leaf child
        valid, offsets, off = 0, {}, 0
        for s, c in enumerate(counts):                    # bvh_wide.cu: primitives stored in slot order
            for k in range(int(c)):
                valid |= 1 << (2 * s + k)
                offsets[2 * s + k] = off
                off += 1
        g_bits = (int(rng.integers(0, 256)) << 24) | (valid << 8) | int(rng.integers(0, 256))   # hit byte | V | imask
        for k, o in offsets.items():
            rank = bin((g_bits >> 8) & ~(M32 << k) & M32).count("1")
            assert rank == o


def test_imad_form_equals_select_form_and_valid_mask():
    rng = np.random.default_rng(11)
    assert sum(child_constant(s) for s in range(8)) == 0xFF00FFFF
    for _ in range(2000):
        hit = rng.integers(0, 2, 8)
        selected = 0
        imad = 0xFF00FFFF
        for s in range(8):
            if hit[s]:
                selected |= child_constant(s)
            else:
                imad = (imad + (1 * ((0 - child_constant(s)) & M32))) & M32   # mad.lo.u32: miss bit * -K + mask
        assert imad == selected
        # the node's valid word keeps inner children (imask) and existing primitives (V) only
        kinds = rng.integers(0, 4, 8)                     # 0 empty, 1 inner, 2 leaf with 1, 3 leaf with 2 primitives
        imask = sum(1 << s for s in range(8) if kinds[s] == 1)
        v = sum(((1 << (int(kinds[s]) - 1)) - 1) << (2 * s) for s in range(8) if kinds[s] >= 2)
        got = selected & ((imask << 24) | v)
        for s in range(8):
            assert (got >> (24 + s) & 1) == int(bool(hit[s]) and kinds[s] == 1)
            assert (got >> (2 * s) & 3) == ((1 << (int(kinds[s]) - 1)) - 1 if hit[s] and kinds[s] >= 2 else 0)
        # float accumulator of MOX_HIT_SIGN=3: the same sum with the inner bits at 16..23 stays below 2^24 (exact in fp32)
        acc = np.float32(16777215.0)
        for s in range(8):
            if not hit[s]:
                acc = np.float32(acc - np.float32((1 << (16 + s)) | (3 << (2 * s))))
        assert int(acc) == ((selected >> 24) << 16) | (selected & 0xFFFF)


def f32(x):
    return struct.unpack("<f", struct.pack("<f", x))[0]


def test_prmt_byte_conversion_is_the_affine_map_within_the_stated_bound():
    rng = np.random.default_rng(3)
    one = 0x3F800000
    for _ in range(4000):
        b = int(rng.integers(0, 256))
        m = struct.unpack("<f", struct.pack("<I", one | (b << 8)))[0]      # prmt: byte b into mantissa bits 8..15
        assert m == 1.0 + b * 2.0 ** -15
        ia = f32(float(rng.normal()) * 2.0 ** int(rng.integers(-12, 6)))   # grid scale * 1/d
        oa = f32(float(rng.normal()) * 2.0 ** int(rng.integers(-6, 12)))   # (node origin - ray origin) * 1/d
        k = f32(ia * 32768.0)
        exact = b * ia + oa
        e = f32(2.0 ** -22 * abs(k) + 2.0 ** -21 * abs(oa))
        near = f32(m * k + f32(f32(oa - e) - k))            # one FMA on the GPU: compare with its exact value
        far = f32(m * k + f32(f32(oa + e) - k))
        slack = 2.0 ** -23 * (abs(near) + abs(far)) + 1e-30  # the FMA's own rounding, covered by the 1e-5 far-side widening
        assert near <= exact + slack and far >= exact - slack
