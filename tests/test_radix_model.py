"""CPU model of the fused per-node kernel's radix sort (erasor_b200/csrc/updater_kernels.cu, k_node_fused): 9-bit digits, only the
passes the key width needs, per-segment histograms (counted by RED while the previous pass scatters), one exclusive row scan per
digit, and a stable per-segment placement behind  digit base + row prefix.  The model restates that data flow in numpy and checks
it against a stable argsort -- the property voxelize_preserving_labels needs (voxel members stay in cloud order) -- for segment
lengths beyond the 8-row minimum and key spaces from 1 bit to 31 bits; it also restates the scratch carve-up so that a change of
the layout that would overrun the buffer shows up without a GPU."""
import numpy as np
import pytest

RB, RD = 9, 512
RS_ROWS_MIN, RS_MAX_SEGS = 8, 8192
PART_CHUNK, K_FUSED_MAX_GRID = 4096, 1024


def rs_seg_len(n):
    rows = (n + 32 * RS_MAX_SEGS - 1) // (32 * RS_MAX_SEGS)
    rows = (rows + RS_ROWS_MIN - 1) // RS_ROWS_MIN * RS_ROWS_MIN
    return 32 * max(rows, RS_ROWS_MIN)


def model_sort(keys, seg):
    """(sorted keys, sorted cloud indices) by the kernel's passes."""
    n = len(keys)
    nseg = (n + seg - 1) // seg
    lim = int(keys.max()) + 1 if n else 1
    bits = 1
    while (1 << bits) < lim:
        bits += 1
    npass = (bits + RB - 1) // RB
    k, v = keys.astype(np.uint32).copy(), np.arange(n, dtype=np.uint32)
    # pass 0's histogram is counted while the keys are computed: cnt[digit][segment of the SOURCE position]
    cnt = np.zeros((RD, nseg), dtype=np.int64)
    np.add.at(cnt, (k & (RD - 1), np.arange(n) // seg), 1)
    for p in range(npass):
        shift = RB * p
        row_prefix = np.cumsum(cnt, axis=1) - cnt                    # (a) one warp per digit: exclusive scan over the segments
        tot = cnt.sum(axis=1)
        base = np.cumsum(tot) - tot                                  # exclusive scan of the digit totals (every scatter warp redoes it)
        ko, vo = np.empty_like(k), np.empty_like(v)
        cnt_next = np.zeros((RD, nseg), dtype=np.int64)
        for s in range(nseg):                                        # (b) one warp per segment, rows of 32 in order
            nxt = base + row_prefix[:, s]                            # next free destination per digit
            b0, b1 = s * seg, min(n, (s + 1) * seg)
            for r0 in range(b0, b1, 32):
                d = (k[r0:min(r0 + 32, b1)] >> shift) & (RD - 1)
                for lane, dig in enumerate(d):                       # match.any ranks = lane order inside a run of equal digits
                    o = nxt[dig]
                    nxt[dig] += 1
                    ko[o], vo[o] = k[r0 + lane], v[r0 + lane]
                    if p + 1 < npass:
                        cnt_next[(int(k[r0 + lane]) >> (shift + RB)) & (RD - 1), o // seg] += 1
        k, v, cnt = ko, vo, cnt_next
    return k, v, npass


@pytest.mark.parametrize("n,key_bits,seg", [(1, 1, 256), (300, 5, 256), (5000, 9, 256), (5000, 10, 256), (7000, 27, 256), (6000, 31, 256),
                                            (9000, 18, 512), (4097, 27, 1024)])
def test_model_sort_is_the_stable_sort(n, key_bits, seg):
    rng = np.random.default_rng(n + key_bits)
    keys = rng.integers(0, 1 << key_bits, n, dtype=np.int64).astype(np.uint32)
    keys[rng.integers(0, n, n // 3)] = keys[0]                        # long runs of one voxel, spread over the segments
    if key_bits > 1:
        keys[-1] = (1 << key_bits) - 1                                # the key width is really needed
    k, v, npass = model_sort(keys, seg)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(v, order.astype(np.uint32)) and np.array_equal(k, keys[order])
    assert npass == max(1, -(-max(1, int(keys.max()).bit_length()) // RB))


def test_segment_length_and_scratch_carve_up():
    for n in (0, 1, 255, 256, 110_000, 2_097_152, 2_097_153, 50_000_000, 4_000_000_000):
        seg = rs_seg_len(n)
        nseg = (n + seg - 1) // seg
        assert seg % 256 == 0 and seg >= 256 and nseg <= RS_MAX_SEGS
        # vox_plan(): key / idx ping-pong, two histogram matrices of RD x (nseg + 1), digit totals, head-chunk counters, voxel starts / keys, partials
        used = 4 * n + 2 * RD * (nseg + 1) + RD + ((n + PART_CHUNK - 1) // PART_CHUNK + 1) + 2 * (n + 2) + 6 * K_FUSED_MAX_GRID
        # voxelize_tmp_bytes() in words
        have = 4 * n + 2 * RD * ((n + seg - 1) // seg + 1) + RD + ((n + PART_CHUNK - 1) // PART_CHUNK + 1) + 2 * (n + 2) + 64 + 6 * K_FUSED_MAX_GRID
        assert used <= have
