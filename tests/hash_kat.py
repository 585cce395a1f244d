"""Known-answer vectors for the hash-grid INDEX arithmetic, computed with Python integers / numpy float32 scalars that
follow the reference line by line — independent of oracle/ngp_oracle.c and of the CUDA kernels:

  scale       = base_res * exp(level * log_b) - 1                       modules/hash_encoder.py:73-76 (f32)
  resolution  = uint32(ceil(scale)) + 1                                 :78-80
  pos         = xyz * scale + 0.5; pos_grid = uint32(floor(pos)); pos -= pos_grid       :108-110
  corner idx  : bit d of idx selects pos_grid[d] (+1), weight (1 - pos[d]) or pos[d]    :116-126
  under_hash  = x + y*res + z*res^2      (uint32 wrap)                  :53-60
  fast_hash   = x*1 ^ y*2654435761 ^ z*805459861   (uint32 wrap)        :43-51
  index       = hash % map_size;  element = offset*F + index*F + f      :62-71, :106, :134-139

The table used by the KAT stores its own entry index (feature 0 = index, feature 1 = -index; exactly representable
in fp32 below 2^24, the stock table has 5,710,032 entries), so the encoder output at a point equals
sum_c w_c * index_c and any wrong prime, stride, modulo, offset or corner order shows up as an O(index) error.
"""
import numpy as np

PRIMES = (1, 2654435761, 805459861)
M32 = 0xFFFFFFFF


def fast_hash(p):
    r = 0
    for i in range(3):
        r ^= (int(p[i]) * PRIMES[i]) & M32
    return r & M32


def under_hash(p, res):
    r, stride = 0, 1
    for i in range(3):
        r = (r + ((int(p[i]) * stride) & M32)) & M32
        stride = (stride * res) & M32
    return r


def corner_table(lay, xyz, level):
    """[(entry index incl. level offset, weight as np.float32)] * 8 for one point and one level."""
    f = np.float32
    scale = f(lay.scales[level])
    res = int(lay.resolutions[level])
    pos = [f(f(f(v) * scale) + f(0.5)) for v in xyz]
    grid = [int(np.floor(p)) for p in pos]
    frac = [f(p - f(g)) for p, g in zip(pos, grid)]
    out = []
    for idx in range(8):
        w = f(1.0)
        pl = [0, 0, 0]
        for d in range(3):
            if (idx & (1 << d)) == 0:
                pl[d] = grid[d]
                w = f(w * f(f(1.0) - frac[d]))
            else:
                pl[d] = grid[d] + 1
                w = f(w * frac[d])
        h = under_hash(pl, res) if level < lay.begin_fast_hash_level else fast_hash(pl)
        out.append((lay.offsets[level] + h % lay.map_sizes[level], w))
    return out


def index_table(lay):
    """fp32 table [entries, 2]: feature 0 = the entry's own index, feature 1 = its negative."""
    idx = np.arange(lay.total_entries, dtype=np.float32)
    return np.stack([idx, -idx], 1)


def expected(lay, pts):
    """[n, L*2] float64 reference output for the index table (weights are the f32 products of the reference; the sum
    is accumulated in float64 so only the kernels' own fp32 summation order remains as slack)."""
    out = np.zeros((len(pts), lay.levels * 2))
    for i, p in enumerate(pts):
        for l in range(lay.levels):
            s = sum(float(w) * float(e) for e, w in corner_table(lay, p, l))
            out[i, 2 * l] = s
            out[i, 2 * l + 1] = -s
    return out


# hand-picked points: interior, near the box faces (grid coordinate == resolution - 1 -> the +1 corner wraps in dense
# levels), exact cell boundaries at level 0, plus fixed pseudo-random ones
POINTS = np.array([[0.5, 0.5, 0.5], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [0.999999, 0.25, 0.75], [1.0 / 15.0, 2.0 / 15.0, 0.2],
                   [0.123456, 0.654321, 0.314159], [0.9, 0.05, 0.6180339], [0.3333333, 0.6666667, 0.1]], np.float32)


def hand_computed_vectors():
    """A few (point, level) -> corner-0 index values worked out by hand from the formulas above (not by any code path
    under test): stock layout, max_res 1024.
      level 15: scale 1023, x=(0.5,0.5,0.5) -> pos 512.0 -> grid (512,512,512):
                2654435761 = 0x9E3779B1, 805459861 = 0x30025795; multiplying by 512 is a 9-bit shift:
                512 ^ (0x9E3779B1 << 9 mod 2^32) ^ (0x30025795 << 9 mod 2^32)
                = 0x00000200 ^ 0x6EF36200 ^ 0x04AF2A00 = 0x6A5C4A00;  % 2^19 = 0x44A00 = 281088
      level 0 : scale 15, res 16, x=(0.5,0.5,0.5) -> pos 8.0 -> grid (8,8,8): 8 + 8*16 + 8*256 = 2184"""
    return [((0.5, 0.5, 0.5), 15, 0, 5185744 + 281088), ((0.5, 0.5, 0.5), 0, 0, 2184)]
