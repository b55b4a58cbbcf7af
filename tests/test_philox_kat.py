"""CPU: known-answer test of the Philox4x32-10 restatement the parity tests feed the oracle with
(tests/gpu_parity_lib.py philox4x32_10_np).  Vectors: Random123 (D. E. Shaw Research) examples/kat_vectors, lines
"philox4x32 10 ..."; the generator is the one of Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11).
The device's philox4x32_10 (agarcl_b200/csrc/device_math.cuh) is pinned against this restatement by the RNG_PHILOX
parity run of tests/test_gpu_parity.py (every spawn point of 2000 env-steps of 8 instances must match)."""
import numpy as np

from gpu_parity_lib import philox4x32_10_np, philox_uniform_np

KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox4x32_10_known_answers():
    for ctr, key, want in KAT:
        got = tuple(int(x) for x in philox4x32_10_np(ctr, key))
        assert got == want, (ctr, key, [hex(g) for g in got])


def test_uniform_mapping_and_counter_layout():
    # draw k = word (k & 3) of block k >> 2; 24 bits -> [0, 1); the instance index is counter word 2, the seed the key
    seed, inst = 0x299f31d0a4093822, 0x13198a2e
    k = np.arange(0x243f6a88 * 4, 0x243f6a88 * 4 + 4, dtype=np.uint64)
    u = philox_uniform_np(seed, inst, k)
    blk = philox4x32_10_np((0x243f6a88, 0, inst, 0), (seed & 0xffffffff, seed >> 32))
    for j in range(4):
        assert u[j] == np.float32(int(blk[j]) >> 8) * np.float32(2.0 ** -24)
    d = philox_uniform_np(7, 3, np.arange(1 << 16))
    assert d.dtype == np.float32 and (d >= 0).all() and (d < 1).all() and abs(float(d.mean()) - 0.5) < 0.01
    assert not np.array_equal(d, philox_uniform_np(7, 4, np.arange(1 << 16)))  # streams of two instances differ
