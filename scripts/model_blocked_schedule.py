"""Design model (CPU, no GPU): the blocked line-systolic schedule proposed in DESIGN.md section 7.1.

A lane resolves a BX x BY x BZ block of grid rows per warp step instead of one row.  Lane l of a task owns the line
group J = 32 * jb + l (BY adjacent lines of BZ adjacent planes) and walks its column groups front to back, skewed by
one step per lane: at step s it resolves column group I = s - l.  The script
  (1) checks by brute force that every dependency of a 7-point lower-triangular solve (offsets 1, nx, nx*ny) is
      resolved either earlier inside the same block (internal level a + b + c) or in a block that the producing
      lane / warp finished at an earlier step (same lane: s - 1; previous lane: s - 1; previous plane group: the warp
      one step ahead), for ragged grid sizes too;
  (2) prints the dependent warp steps on the critical path of the BASELINE grids for 1x1x1 (today) and blocked variants.

usage: python scripts/model_blocked_schedule.py
"""
import itertools


def schedule(nx, ny, nz, bx, by, bz):
    """step at which every row (i, j, k) is resolved, and the rank of the row inside its block"""
    gx, gy, gz = -(-nx // bx), -(-ny // by), -(-nz // bz)
    nyb = -(-gy // 32)                      # line-group blocks of 32 lanes
    step, rank = {}, {}
    for i, j, k in itertools.product(range(nx), range(ny), range(nz)):
        I, J, K = i // bx, j // by, k // bz
        lane, jb = J % 32, J // 32
        # task (jb, K) starts 33 steps after (jb - 1, K) (lane 31 of the previous block is 31 steps in + 1 hand-off)
        # and 1 step after (jb, K - 1)
        start = 33 * jb + K
        step[(i, j, k)] = start + I + lane
        rank[(i, j, k)] = (i % bx) + (j % by) + (k % bz)   # internal wavefront level inside the block
    return step, rank, (33 * (nyb - 1) + (gz - 1) + gx + 31)


def check(nx, ny, nz, bx, by, bz):
    step, rank, _ = schedule(nx, ny, nz, bx, by, bz)
    for (i, j, k), s in step.items():
        for di, dj, dk in ((1, 0, 0), (0, 1, 0), (0, 0, 1)):
            d = (i - di, j - dj, k - dk)
            if min(d) < 0:
                continue
            sd = step[d]
            same_block = (d[0] // bx, d[1] // by, d[2] // bz) == (i // bx, j // by, k // bz)
            if same_block:
                assert rank[d] < rank[(i, j, k)], ("internal order", (i, j, k), d)
            else:
                assert sd < s, ("cross-block dependency not ready", (nx, ny, nz, bx, by, bz), (i, j, k), d, sd, s)
    return True


def main():
    for dims in ((5, 7, 3), (9, 70, 4), (16, 33, 5), (12, 65, 6)):
        for blk in ((1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2)):
            check(*dims, *blk)
    print("dependency order verified for ragged grids and blocks 1x1x1, 2x1x1, 2x2x1, 2x2x2")
    print("%-26s %10s %10s %10s %10s" % ("grid", "1x1x1", "2x1x1", "2x2x1", "2x2x2"))
    for name, dims in (("poisson3d(128)", (128, 128, 128)), ("convdiff3d(160)", (160, 160, 160)),
                       ("poisson2d(256)", (256, 256, 1))):
        cols = []
        for blk in ((1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2)):
            gx, gy, gz = (-(-d // b) for d, b in zip(dims, blk))
            nyb = -(-gy // 32)
            cols.append(33 * (nyb - 1) + (gz - 1) + gx + 31)
        print("%-26s %10d %10d %10d %10d   dependent warp steps per solve" % ((name,) + tuple(cols)))


if __name__ == "__main__":
    main()
