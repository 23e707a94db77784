"""CPU model of the index math of the staged epilogues (conv_igemm, conv_wgrad, conv_halo) on the unvalidated
branch: every output element must be written exactly once, from the staging slot that holds it.
Run: python scripts/r02/sim_epilogue.py"""
import itertools
import numpy as np

K_STG_ROW = 256 + 16


def sim_igemm(block_n, out_fp32, cout, rows_valid):
    """One warp (32 accumulator rows), one tile; returns dict global_byte_offset -> (row, col_byte) source."""
    esize = 4 if out_fp32 else 2
    chunk_cols = min(block_n, 256 // esize)
    j_per_chunk = chunk_cols // 32
    written = {}
    stage = {}                                      # (lane_row, byte) -> (row, col)
    row_off = {lane: -1 for lane in range(32)}
    for j in range(block_n // 32):
        col0 = j * 32
        for lane in range(32):
            if rows_valid[lane]:
                off = lane * 1000003 * cout + col0          # pix * cout + c0 with a fake, unique pixel id per row
                if j % j_per_chunk == 0:
                    row_off[lane] = off
                base = lane * K_STG_ROW + (j % j_per_chunk) * 32 * esize
                for g in range(4):
                    for i in range(8):
                        col = col0 + g * 8 + i
                        byte = base + (g * 32 if out_fp32 else g * 16) + i * esize
                        for bb in range(esize):
                            stage[byte + bb] = (lane, col, bb)
            elif j % j_per_chunk == 0:
                row_off[lane] = -1
        if (j + 1) % j_per_chunk == 0:
            lanes_per_row = (chunk_cols * esize) >> 4
            rows_per_pass = 32 // lanes_per_row
            for r0 in range(0, 32, rows_per_pass):
                for lane in range(32):
                    sub = lane % lanes_per_row
                    rr = r0 + lane // lanes_per_row
                    o_el = row_off[rr]
                    if o_el >= 0:
                        for bb in range(16):
                            src = stage[rr * K_STG_ROW + sub * 16 + bb]
                            dst = o_el * esize + sub * 16 + bb
                            assert dst not in written, "double write"
                            written[dst] = src
    # check: each valid row wrote exactly block_n columns, each byte from the matching (row, col)
    for lane in range(32):
        if not rows_valid[lane]:
            continue
        for col in range(block_n):
            for bb in range(esize):
                dst = (lane * 1000003 * cout + col) * esize + bb
                assert written.get(dst) == (lane, col, bb), (lane, col, bb, written.get(dst))
    assert len(written) == sum(rows_valid) * block_n * esize
    return True


def sim_wgrad(block_c):
    written = {}
    stage = {}
    for j in range(block_c // 32):
        for lane in range(32):
            base = lane * K_STG_ROW + (j & 1) * 128
            for g in range(8):
                for i in range(4):
                    stage[base + g * 16 + i * 4] = (lane, j * 32 + g * 4 + i)
        last = (j + 1 == block_c // 32)
        if (j & 1) or last:
            cols = 64 if (j & 1) else 32
            lanes_per_row = cols // 4
            rows_per_pass = 32 // lanes_per_row
            jc0 = j - 1 if (j & 1) else j
            for r0 in range(0, 32, rows_per_pass):
                for lane in range(32):
                    sub = lane % lanes_per_row
                    rr = r0 + lane // lanes_per_row
                    for i in range(4):
                        src = stage[rr * K_STG_ROW + sub * 16 + i * 4]
                        dst = (rr, jc0 * 32 + sub * 4 + i)
                        assert dst not in written
                        written[dst] = src
    for rr in range(32):
        for c in range(block_c):
            assert written[(rr, c)] == (rr, c), (rr, c, written.get((rr, c)))
    assert len(written) == 32 * block_c
    return True


def sim_halo_windows(pitch=24, rows=18):
    """Every (tap, sub, row r) window address stays inside the box and maps to pixel (y + ty, x + tx + 8 sub)."""
    for tap, sub in itertools.product(range(9), range(2)):
        ty, tx = divmod(tap, 3)
        start = (ty * pitch + tx + 8 * sub) * 128
        for r in range(128):
            y, x = divmod(r, 8)
            addr = start + y * (pitch * 128) + x * 128          # start + (r/8)*SBO + (r%8)*128
            row = addr // 128
            assert row == (y + ty) * pitch + (x + tx + 8 * sub) and row < pitch * rows
            assert x + tx + 8 * sub < pitch and y + ty < rows
    smem = 2 * pitch * rows * 128 + 4 * 128 * 128 + 256 + 4 * 32 * K_STG_ROW + 1024
    assert smem <= 227 * 1024, smem
    return smem


if __name__ == "__main__":
    rng = np.random.RandomState(0)
    for block_n, fp32 in itertools.product((32, 64, 128, 256), (False, True)):
        for trial in range(3):
            valid = [True] * 32 if trial == 0 else list(rng.rand(32) > 0.3)
            sim_igemm(block_n, fp32, cout=max(block_n, 64), rows_valid=valid)
    print("conv_igemm staged epilogue: every element written once from its own slot")
    for bc in (64, 128, 192, 256):
        sim_wgrad(bc)
    print("conv_wgrad staged reduction: every element reduced once from its own slot")
    print("conv_halo windows in bounds; shared memory", sim_halo_windows(), "bytes")
