"""Host-side emulation of index formulas used inside the CUDA kernels (no GPU needed), in the spirit of
tests/test_fft_indexing.py: each check restates a formula from the cited kernel source in numpy and proves the
property the kernel relies on (bijectivity, exact cover, balanced barrier traffic).  They guard the algebra, not the
CUDA text: when a formula in the kernel changes, the restatement here has to change with it."""
import numpy as np
import pytest


# ---- gemm_epilogue.cuh: epi_store_staged (swizzled smem staging, cooperative coalesced stores) ----------------
@pytest.mark.parametrize("out_bf16", [False, True])
def test_epilogue_staging_swizzle_is_a_bijection_and_stores_are_coalesced(out_bf16):
    row_bytes = 64 if out_bf16 else 128                     # 32 columns per chunk
    nseg = row_bytes // 16                                  # 16-byte segments per row
    buf = -np.ones((128 * row_bytes // 16,), dtype=np.int64)  # segment slot -> (row, logical segment)
    for r in range(128):                                    # writer: thread r owns row r
        sw = ((r >> 1) & 3) if out_bf16 else (r & 7)
        for j in range(nseg):
            slot = r * nseg + (j ^ sw)
            assert buf[slot] == -1                          # no two writes collide
            buf[slot] = r * 100 + j
    assert (buf >= 0).all()
    seen = set()
    for et in range(128):                                   # reader: cooperative pattern
        q = et & (nseg - 1)
        for i in range(nseg):
            row = i * (128 // nseg) + (et >> (2 if out_bf16 else 3))
            sw = ((row >> 1) & 3) if out_bf16 else (row & 7)
            v = buf[row * nseg + (q ^ sw)]
            assert v == row * 100 + q                       # reads back logical segment q of that row
            seen.add((row, q))
    assert len(seen) == 128 * nseg                          # every (row, segment) stored exactly once
    # coalescing: the nseg lanes with consecutive `et` cover one row contiguously
    for et0 in range(0, 128, nseg):
        rows = {(et >> (2 if out_bf16 else 3)) for et in range(et0, et0 + nseg)}
        assert len(rows) == 1


# ---- gemm2_sm100.cuh / gemm_epilogue.cuh: two epilogue groups interleave 64-column units ------------------------
@pytest.mark.parametrize("bn", [128, 192, 256])
def test_pair_kernel_epilogue_groups_cover_every_column_once(bn):
    cover = np.zeros(bn, dtype=int)
    per_group = []
    for grp in range(2):
        units = list(range(grp, bn // 64, 2))               # epi_drain_tile(cc0 = grp, cc_step = 2)
        per_group.append(len(units))
        for cc in units:
            cover[cc * 64:(cc + 1) * 64] += 1
    assert (cover == 1).all()
    assert max(per_group) - min(per_group) <= 1             # 2+1 units at BN=192, even otherwise


def test_single_cta_epilogue_groups_split_the_tile_in_halves():
    bn, groups = 128, 2                                     # GemmEpi<128, 6>
    cols = [list(range(g * (bn // groups), (g + 1) * (bn // groups))) for g in range(groups)]
    assert sorted(sum(cols, [])) == list(range(bn))
    # warps 2-5 and 6-9 both cover the four TMEM lane quarters (warp % 4)
    for first in (2, 6):
        assert sorted(w % 4 for w in range(first, first + 4)) == [0, 1, 2, 3]


# ---- gemm_sm100.cuh: weight tiles requested before the PDL wait --------------------------------------------------
@pytest.mark.parametrize("num_kb,stages", [(16, 6), (32, 6), (2, 6), (16, 3), (1, 4)])
def test_early_weight_tiles_plus_main_loop_issue_every_k_block_once(num_kb, stages):
    early_b = min(stages, num_kb)
    b_loads = [kb for kb in range(early_b)]                 # before pdl_wait: stage kb, expect_tx armed there
    expect = [kb for kb in range(early_b)]
    a_loads = []
    for kb in range(num_kb):                                # producer loop
        if kb >= early_b:
            expect.append(kb)
            b_loads.append(kb)
        a_loads.append(kb)
    assert b_loads == list(range(num_kb)) and a_loads == list(range(num_kb)) and expect == list(range(num_kb))
    assert all(kb % stages == kb for kb in range(early_b))  # early tiles land in the stage the loop will use


# ---- attention2_sm100.cuh: P written to TMEM as the A operand of the TS-form MMA ---------------------------------
def test_attention_p_tiles_fill_the_tmem_columns_the_mma_reads():
    # softmax thread: chunk c (8 probabilities) -> 4 packed bf16x2 words; chunks (2q, 2q+1) stored by one
    # tcgen05.st.x8 at column offset q*8.  MMA k-step k (16 keys) reads columns [k*8, k*8+8).
    col_of_key = {}
    for c in range(16):
        for w in range(4):
            col = (c >> 1) * 8 + (c & 1) * 4 + w
            for half in range(2):
                key = c * 8 + 2 * w + half
                assert key not in col_of_key
                col_of_key[key] = (col, half)                # low half = even key (pack_bf16x2(lo, hi))
    assert sorted(col_of_key) == list(range(128))
    for k in range(8):
        keys = [key for key, (col, _) in col_of_key.items() if k * 8 <= col < k * 8 + 8]
        assert sorted(keys) == list(range(16 * k, 16 * k + 16))
    for key, (col, half) in col_of_key.items():
        assert col == key // 2 and half == key % 2           # K-major: two consecutive keys per 32-bit column
    # TMEM map: S0 S1 O0 O1 P0 P1 = exactly the 512 columns of an SM
    regions = [(0, 128), (128, 128), (256, 64), (320, 64), (384, 64), (448, 64)]
    assert sum(n for _, n in regions) == 512
    assert all(regions[i][0] + regions[i][1] == regions[i + 1][0] for i in range(5))


# ---- attention2_sm100.cuh: hand-off barriers between the two softmax groups --------------------------------------
@pytest.mark.parametrize("num_kv", [1, 2, 3, 8, 9])
@pytest.mark.parametrize("mode", [1, 2])
def test_attention_handoff_barriers_are_balanced_and_order_the_groups(num_kv, mode):
    """bar 2: group 0 syncs, group 1 arrives; bar 3: group 1 syncs, group 0 arrives.  A sync may only complete when
    a matching arrive has been posted; no arrive may be left unconsumed at exit."""
    pending = {2: 0, 3: 0}
    pending[2] += 1                                          # group 1's initial arrive: group 0 goes first
    pc = {0: 0, 1: 0}                                        # next tile of each group
    waiting = {0: True, 1: True}                             # at the sync in front of the exponential loop
    order = []
    for _ in range(4 * num_kv + 4):
        progressed = False
        for g, bar, other_bar in ((0, 2, 3), (1, 3, 2)):
            if pc[g] < num_kv and waiting[g] and pending[bar] > 0:
                pending[bar] -= 1
                order.append((g, pc[g]))
                # exponential loop; the release is at half time (mode 2) or at the end (mode 1) — same counts
                if g == 0 or pc[g] + 1 < num_kv:
                    pending[other_bar] += 1
                pc[g] += 1
                progressed = True
        if not progressed:
            break
    assert pc == {0: num_kv, 1: num_kv}, "a group starved (deadlock)"
    assert pending == {2: 0, 3: 0}, "an arrive was left unconsumed at exit"
    assert order == [(g, j) for j in range(num_kv) for g in (0, 1)]   # strict alternation 0,1,0,1,...


def _tile_walk(cluster_id, num_clusters, n_tiles, total_tiles):
    """Python restatement of gemm2_sm100.cuh tile_walk()."""
    m_tiles = total_tiles // n_tiles
    G = num_clusters // m_tiles if m_tiles > 0 else 0
    sticky = total_tiles > num_clusters and G >= 1 and -(-n_tiles // G) <= -(-total_tiles // num_clusters)
    if sticky:
        m, j = divmod(cluster_id, G)
        first, stride = m * n_tiles + j, G
        count = -(-(n_tiles - j) // G) if (m < m_tiles and j < n_tiles) else 0
    else:
        first, stride = cluster_id, num_clusters
        count = -(-(total_tiles - cluster_id) // num_clusters) if cluster_id < total_tiles else 0
    return first, stride, count, sticky


def test_pair_gemm_tile_walk_covers_every_tile_once_and_row_sticky_keeps_rows():
    """The persistent CTA-pair GEMM's tile walk: every tile exactly once for any (clusters, n_tiles, m_tiles); in the
    row-sticky mode (QKV at batch 1: 8 x 16 tiles on 74 clusters) all tiles of a cluster share the row block and
    no cluster gets more tiles than with the round-robin walk."""
    seen_sticky = False
    for clusters in (1, 2, 7, 64, 74):
        for n_tiles in (1, 2, 4, 12, 16, 24):
            for m_tiles in (1, 2, 8, 9, 75, 469):
                total = n_tiles * m_tiles
                launched = min(clusters, total)
                tiles, per = [], []
                for c in range(launched):
                    first, stride, count, sticky = _tile_walk(c, launched, n_tiles, total)
                    mine = [first + i * stride for i in range(count)]
                    tiles += mine
                    per.append(len(mine))
                    if sticky:
                        seen_sticky = True
                        assert len({t // n_tiles for t in mine}) <= 1
                assert sorted(tiles) == list(range(total)), (clusters, n_tiles, m_tiles)
                assert max(per) <= -(-total // launched)
    assert seen_sticky
    first, stride, count, sticky = _tile_walk(10, 74, 16, 128)
    assert sticky and (first, stride, count) == (1 * 16 + 1, 9, 2)
