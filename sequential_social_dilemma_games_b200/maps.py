"""ASCII maps of the two gridworlds plus helpers to build the benchmark variants.

The two default layouts are the game boards of the reference
(social_dilemmas/constants.py:7-23 HARVEST_MAP, :25-50 CLEANUP_MAP); they are data, and a
drop-in has to ship the same boards.  Alphabet: '@' wall, 'P' agent spawn point, 'A' apple
spawn point (Harvest), 'B' apple spawn point (Cleanup), 'H' waste (start), 'R' river (waste can
appear), 'S' stream, ' ' empty.
"""

import re

def _decode(board):
    """Boards are stored run-length coded, rows separated by '/', '_' standing for an empty cell
    ('12@' = twelve walls).  The decoded layouts are pinned by tests/test_host_logic.py (shape,
    point counts of SURVEY.md section 0 item 3, SHA-256 of the rows)."""
    rows = []
    for code in board.split('/'):
        rows.append(''.join((' ' if ch == '_' else ch) * int(n or 1) for n, ch in re.findall(r'(\d*)(\D)', code)))
    return rows


# 16 x 38: 155 apple points 'A', 20 spawn points 'P' (constants.py:7-23)
HARVEST_MAP = _decode(
    '38@/@_P3_P6_A4_P_5A4_P2_A_P2_@/@2_P5_A_P_2A4_P4_3A4_A2_A2_@/@5_A_3A2_3A4_A4_A_2A_4A3_@/'
    '@_A2_3A_A4_A2_A_3A2_A2_A3_A_A3_@/@3A2_A_A4_A2_3A_A2_3A8_A_P@/@_A_A2_3A2_3A2_A_A4_A_2A3_2A_2A_@/'
    '@2_A_A2_3A4_A_A2_3A4_3A2_A4_@/@3_3A2_A6_3A2_A4_4A7_@/@_P2_A7_A2_A_3A4_A2_A6_P_@/'
    '@A2_3A2_A2_A2_3A_A4_4A5_P3_@/@4_A_A3_3A2_A_A6_A_2A3_A2_P_@/@5_3A3_A_A2_3A6_2A3_3A_P_@/'
    '@_A4_A5_3A2_A2_P10_A4_@/@7_P5_A9_P2_P_P5_P_@/38@')

# 25 x 18: 103 apple points 'B', 56 'H' + 63 'R' waste points, 12 stream 'S', 10 'P' (constants.py:25-50)
CLEANUP_MAP = _decode(
    '18@/@6R5_5B@/@6H6_4B@/@6R5_5B@/@5R2_P4_4B@/@5R4_P_5B@/@5H7_4B@/@5R6_5B@/@6H6S4B@/@6H6S4B@/'
    '@5R3_P_P_4B@/@5H3_P2_5B@/@6R4_P_4B@/@6H_P3_5B@/@5R7_4B@/@4H4_P2_5B@/@5R7_4B@/@5H2_P_P_5B@/'
    '@5R7_4B@/@4H7_5B@/@5R7_4B@/@5H6_5B@/@5R7_4B@/@4H7_5B@/18@')


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
