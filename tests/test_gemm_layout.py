"""The operand staging maps of uavnet_gemm (csrc/tc_gemm.cuh: Stager<TRANS>::init, store_tile) restated in Python and
checked on the CPU: every element of a tile is written exactly once, at the byte offset of the canonical K-major
no-swizzle UMMA layout, and the store instructions are free of shared-memory bank conflicts.  (The kernel itself is
checked against float64 products on the GPU, tests/test_gpu_gemm.py; this pins the index arithmetic it relies on.)"""
import numpy as np
import pytest

NTHR, KC = 256, 32


def canonical_offset(r, k):
    """K-major, no swizzle: 8 x 16-byte core matrices, LBO 128 (next 4 k), SBO 1024 (next 8 rows)"""
    return (r >> 3) * 1024 + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4


def rowmajor_items(R):
    """thread tid, item j -> (row, first k, byte offset) for a row-major source: one float4 = 4 consecutive k"""
    out = []
    for tid in range(NTHR):
        l, w = tid & 31, tid >> 5
        r_in, q, g = l & 7, ((w & 1) << 2) | (l >> 3), w >> 1
        n_items = max(0, (R * 8 - tid + NTHR - 1) // NTHR)
        for j in range(n_items):
            out.append((tid, j, g * 8 + r_in + 32 * j, q * 4, g * 1024 + q * 128 + r_in * 16 + j * 4096))
    return out


def transposed_items(R):
    """thread tid, item j -> (first row, k, per-component byte offsets in store order) for a [k, rows] source"""
    out = []
    for tid in range(NTHR):
        l, w = tid & 31, tid >> 5
        t, octet = l >> 3, w >> 1
        kk = ((w & 1) << 4) | (t << 2) | (l & 3)
        rlo = ((l >> 2) & 1) << 2
        base = octet * 1024 + (kk >> 2) * 128 + (kk & 3) * 4
        n_items = max(0, (R * 8 - tid + NTHR - 1) // NTHR)
        for j in range(n_items):
            comps = [((i + t) & 3) for i in range(4)]                      # store i writes component (i + rot) & 3
            offs = [base + (rlo + c) * 16 + j * 4096 for c in comps]
            out.append((tid, j, octet * 8 + rlo + 32 * j, kk, comps, offs))
    return out


@pytest.mark.parametrize("R", [16, 48, 112, 128, 208, 224, 256])
def test_rowmajor_staging_covers_the_tile_in_canonical_layout(R):
    seen = np.zeros((R, KC), dtype=np.int32)
    for tid, j, row, k, off in rowmajor_items(R):
        assert 0 <= row < R and off == canonical_offset(row, k)
        seen[row, k:k + 4] += 1
    assert (seen == 1).all()
    # every quarter-warp of a store instruction writes one whole core matrix (128 contiguous bytes): no bank conflicts
    items = {(tid, j): off for tid, j, _, _, off in rowmajor_items(R)}
    for warp in range(NTHR // 32):
        for j in range(R * 8 // NTHR + 1):
            for quarter in range(4):
                offs = sorted(items[(warp * 32 + quarter * 8 + i, j)] for i in range(8) if (warp * 32 + quarter * 8 + i, j) in items)
                if offs:
                    assert len(offs) == 8 and offs == list(range(offs[0], offs[0] + 128, 16)) and offs[0] % 128 == 0


@pytest.mark.parametrize("R", [16, 48, 112, 128, 208, 224, 256])
def test_transposed_staging_covers_the_tile_and_is_bank_conflict_free(R):
    seen = np.zeros((R, KC), dtype=np.int32)
    for tid, j, row0, k, comps, offs in transposed_items(R):
        for c, off in zip(comps, offs):
            assert 0 <= row0 + c < R and off == canonical_offset(row0 + c, k)
            seen[row0 + c, k] += 1
    assert (seen == 1).all()
    # store instruction i of item j: the 32 lanes of a warp hit 32 different banks
    table = {(tid, j): offs for tid, j, _, _, _, offs in transposed_items(R)}
    for warp in range(NTHR // 32):
        for j in range(R * 8 // NTHR + 1):
            lanes = [table.get((warp * 32 + lane, j)) for lane in range(32)]
            if lanes[0] is None:
                continue
            assert all(x is not None for x in lanes)                       # R * 8 is a multiple of 128: whole warps
            for i in range(4):
                banks = {(x[i] // 4) % 32 for x in lanes}
                assert len(banks) == 32
    # a warp item reads 16 k-rows x 32 contiguous bytes of the source (one fully used sector per k)
    by_warp = {}
    for tid, j, row0, k, _, _ in transposed_items(R):
        by_warp.setdefault((tid >> 5, j), []).append((k, row0))
    for rows in by_warp.values():
        ks = sorted({k for k, _ in rows})
        assert len(ks) == 16
        for k in ks:
            r = sorted(r0 for kk, r0 in rows if kk == k)
            assert len(r) == 2 and r[1] == r[0] + 4 and r[0] % 8 == 0


def test_descriptor_fields_match_the_cutlass_encoding():
    """smem_desc / smem_desc_sw128 / instr_desc of tc_gemm.cuh against the bit positions of cute::UMMA::SmemDescriptor
    and InstrDescriptor (start address >> 4 at 0, LBO >> 4 at 16, SBO >> 4 at 32, version 1 at 46, layout type at 61;
    D format at 4, A / B format at 7 / 10, N >> 3 at 17, M >> 4 at 24)."""
    def smem_desc(addr, lbo, sbo, layout=0):
        return ((addr & 0x3FFFF) >> 4) | ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | 1 << 46 | layout << 61
    d = smem_desc(0x12340, 128, 1024)
    assert d & 0x3FFF == 0x1234 and (d >> 16) & 0x3FFF == 8 and (d >> 32) & 0x3FFF == 64 and (d >> 46) & 3 == 1 and d >> 61 == 0
    s = smem_desc(0x400, 16, 1024, layout=2)
    assert (s >> 16) & 0x3FFF == 1 and (s >> 32) & 0x3FFF == 64 and s >> 61 == 2
    idesc = (1 << 4) | (2 << 7) | (2 << 10) | ((208 >> 3) << 17) | ((128 >> 4) << 24)
    assert (idesc >> 4) & 3 == 1 and (idesc >> 7) & 7 == 2 and (idesc >> 10) & 7 == 2 and (idesc >> 15) & 3 == 0
    assert ((idesc >> 17) & 0x3F) * 8 == 208 and ((idesc >> 24) & 0x1F) * 16 == 128
