"""CPU model of the one piece of device arithmetic in the b16 visited set that is not integer-exact by construction:
`b16_of` (csrc/hnsw_device.cuh) reduces x = (pid >> 15) + scramble(tag) modulo the bucket count `nb` with an fp32 reciprocal
(q = uint(float(x) * (1.0f / nb)), rem = x - q * nb, one +-nb fix-up).  The set is exact only if that equals x mod nb for every x
the kernel can form (x < 2^17 + 2^22, exact in fp32) and every bucket count the host can choose (<= 2040) — checked here
exhaustively with numpy's IEEE fp32 — and if (home bucket, 15-bit tag) is then an injective function of the PointId while
ceil(n / 32768) <= nb (the host-side admission rule in Index::select_visited_tier)."""
import numpy as np
import pytest


def home_bucket_device(x, nb):
    """The device code path, in IEEE fp32."""
    inv = np.float32(1.0) / np.float32(nb)
    q = (x.astype(np.float32) * inv).astype(np.uint32)  # float -> uint conversion truncates, like the cvt.rzi in the kernel
    rem = x.astype(np.int64) - q.astype(np.int64) * nb
    rem = np.where(rem < 0, rem + nb, rem)
    rem = np.where(rem >= nb, rem - nb, rem)
    return rem


@pytest.mark.parametrize("nb", [1, 2, 3, 7, 15, 16, 31, 254, 255, 256, 510, 1000, 1016, 1023, 1024, 1535, 2039, 2040])
def test_fp32_reciprocal_reduction_equals_integer_modulo(nb):
    x = np.arange(0, (1 << 17) + (1 << 22), dtype=np.uint32)
    assert (home_bucket_device(x, nb) == x.astype(np.int64) % nb).all()


@pytest.mark.parametrize("n,nb", [(1_000_000, 1016), (5_000_000, 1016), (10_000_000, 306), (33_292_288, 1016), (1_250_000, 39), (70_000, 3)])
def test_home_bucket_and_tag_identify_the_point(n, nb):
    assert (n + 32767) // 32768 <= nb  # the host only selects the b16 flavour under this condition
    pid = np.arange(n, dtype=np.uint32)
    tag = pid & np.uint32(0x7FFF)
    x = (pid >> np.uint32(15)) + ((tag * np.uint32(0x9E3779B1)) >> np.uint32(10))
    home = home_bucket_device(x, nb).astype(np.uint64)
    packed = home * np.uint64(32768) + tag.astype(np.uint64)
    assert np.unique(packed).size == n
