"""ASCII maps of the two gridworlds plus helpers to build the benchmark variants.

The two default layouts are the game boards of the reference, kept as plain ASCII rows
(social_dilemmas/constants.py:7-23 HARVEST_MAP, :25-50 CLEANUP_MAP); they are data, and a
drop-in has to ship the same boards.  Alphabet: '@' wall, 'P' agent spawn point, 'A' apple
spawn point (Harvest), 'B' apple spawn point (Cleanup), 'H' waste (start), 'R' river (waste can
appear), 'S' stream, ' ' empty.
"""

# 16 x 38: 155 apple points 'A', 20 spawn points 'P' (social_dilemmas/constants.py:7-23)
HARVEST_MAP = [
    '@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@',
    '@ P   P      A    P AAAAA    P  A P  @',
    '@  P     A P AA    P    AAA    A  A  @',
    '@     A AAA  AAA    A    A AA AAAA   @',
    '@ A  AAA A    A  A AAA  A  A   A A   @',
    '@AAA  A A    A  AAA A  AAA        A P@',
    '@ A A  AAA  AAA  A A    A AA   AA AA @',
    '@  A A  AAA    A A  AAA    AAA  A    @',
    '@   AAA  A      AAA  A    AAAA       @',
    '@ P  A       A  A AAA    A  A      P @',
    '@A  AAA  A  A  AAA A    AAAA     P   @',
    '@    A A   AAA  A A      A AA   A  P @',
    '@     AAA   A A  AAA      AA   AAA P @',
    '@ A    A     AAA  A  P          A    @',
    '@       P     A         P  P P     P @',
    '@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@',
]

# 25 x 18: 103 apple points 'B', 56 'H' + 63 'R' waste points, 12 stream 'S', 10 'P' (social_dilemmas/constants.py:25-50)
CLEANUP_MAP = [
    '@@@@@@@@@@@@@@@@@@',
    '@RRRRRR     BBBBB@',
    '@HHHHHH      BBBB@',
    '@RRRRRR     BBBBB@',
    '@RRRRR  P    BBBB@',
    '@RRRRR    P BBBBB@',
    '@HHHHH       BBBB@',
    '@RRRRR      BBBBB@',
    '@HHHHHHSSSSSSBBBB@',
    '@HHHHHHSSSSSSBBBB@',
    '@RRRRR   P P BBBB@',
    '@HHHHH   P  BBBBB@',
    '@RRRRRR    P BBBB@',
    '@HHHHHH P   BBBBB@',
    '@RRRRR       BBBB@',
    '@HHHH    P  BBBBB@',
    '@RRRRR       BBBB@',
    '@HHHHH  P P BBBBB@',
    '@RRRRR       BBBB@',
    '@HHHH       BBBBB@',
    '@RRRRR       BBBB@',
    '@HHHHH      BBBBB@',
    '@RRRRR       BBBB@',
    '@HHHH       BBBBB@',
    '@@@@@@@@@@@@@@@@@@',
]


def tile_map(ascii_map, reps_rows=2, reps_cols=2):
    """Tile a wall-enclosed map (BASELINE.json config 4: CLEANUP_MAP tiled 2x2 -> 50x36).

    Interior wall rows/columns are kept, so every tile stays enclosed and agents never index
    outside the grid (agent.py:111 reads grid[new_row, new_col] unchecked).
    """
    rows = []
    for _ in range(reps_rows):
        for line in ascii_map:
            rows.append(line * reps_cols)
    return rows


def validate_map(ascii_map):
    """The step path assumes a rectangular, wall-enclosed board."""
    if not ascii_map or any(len(r) != len(ascii_map[0]) for r in ascii_map):
        raise ValueError("ascii_map must be a non-empty rectangular list of strings")
    h, w = len(ascii_map), len(ascii_map[0])
    if h > 255 or w > 255:
        raise ValueError("maps larger than 255x255 are not supported (positions are bytes on device)")
    return h, w
