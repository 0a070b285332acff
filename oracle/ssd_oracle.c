/* TEST INFRASTRUCTURE ONLY -- see ssd_oracle.h.
 *
 * Literal, sequential restatement of the reference's step path.  Every function cites the
 * reference lines it follows (paths relative to the reference root).  The structure follows
 * the reference (ordered dicts, np.unique order, snapshot-vs-live lookups), not the GPU
 * kernels: the two implementations are deliberately independent.
 */
#include "ssd_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_THREADS 256

struct OrcEnv {
    int kind, H, W, N, r, V, beam_len;
    uint8_t* base_map;
    uint8_t color[128][3];
    double harvest_prob[4];
    double* apple_prob;
    double* waste_prob;
    int area;
    int n_apple, n_waste, n_spawn;
    int16_t* apple_pts; /* (row, col), row-major scan order: harvest.py:22-26 / cleanup.py:49-54 */
    int16_t* waste_pts; /* 'H' or 'R' cells of base_map, row-major: cleanup.py:59-60 */
    int16_t* spawn_pts;
};

typedef struct { int r, c; } Cell;

/* ------------------------------------------------------------------ Philox4x32-10 */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { STREAM_MOVE = 0, STREAM_SPAWN, STREAM_WASTE, STREAM_RPOINT, STREAM_RROT, STREAM_RSPAWN };

typedef struct {
    const OrcTape* tape; /* NULL => Philox */
    int b;               /* env index inside the batch (tape rows) */
    uint32_t key[2], env_id, t;
    int k;               /* sequential rank of the next uniform draw */
    int spawn_stream;
    int err;
} Draws;

static uint32_t px_word(const Draws* d, uint32_t stream, uint32_t i) {
    uint32_t ctr[4] = {d->env_id, d->t, stream, i >> 2}, out[4];
    orc_philox4x32_10(ctr, d->key, out);
    return out[i & 3];
}

/* np.random.rand(1)[0]: harvest.py:101, cleanup.py:139,150.  Returns (u < p). */
static int draw_less(const OrcEnv* e, Draws* d, double p) {
    int k = d->k++;
    if (d->tape) {
        if (k >= d->tape->u_stride) { d->err = -3; return 0; }
        return d->tape->uniforms[(size_t)d->b * d->tape->u_stride + k] < p;
    }
    (void)e;
    uint32_t a = px_word(d, d->spawn_stream, 2u * k), b = px_word(d, d->spawn_stream, 2u * k + 1);
    uint64_t u53 = ((uint64_t)(a >> 5) << 26) | (b >> 6);
    return (double)u53 / 9007199254740992.0 < p; /* exact: u53 < 2^53 */
}

/* ------------------------------------------------------------------ small helpers */
static const int ACT_VEC[5][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}, {0, 0}}; /* map_env.py:11-15 */
static const int ORI_VEC[4][2] = {{0, -1}, {1, 0}, {0, 1}, {-1, 0}};        /* UP RIGHT DOWN LEFT, map_env.py:19-22 */

static int is_wall(const OrcEnv* e, int r, int c) { return e->base_map[r * e->W + c] == '@'; }
static int cell_eq(Cell a, Cell b) { return a.r == b.r && a.c == b.c; }

/* `x in self.agent_pos` (map_env.py:251-253): any agent standing on the cell */
static int occupied(const Cell* pos, int N, Cell c) {
    for (int a = 0; a < N; ++a) if (cell_eq(pos[a], c)) return 1;
    return 0;
}
/* {tuple(pos): id for agents in order}[cell] -- the LAST agent on the cell wins */
static int by_pos(const Cell* pos, int N, Cell c) {
    int o = -1;
    for (int a = 0; a < N; ++a) if (cell_eq(pos[a], c)) o = a;
    return o;
}

/* ------------------------------------------------------------------ update_moves, map_env.py:357-543 */
static void update_moves(const OrcEnv* e, Cell* pos, uint8_t* ori, const int8_t* act,
                         const uint8_t* order, Draws* d) {
    const int N = e->N;
    int n_mov = 0, mover[ORC_MAX_AGENTS], is_mover[ORC_MAX_AGENTS] = {0};
    Cell moves[ORC_MAX_AGENTS]; /* agent_moves values */

    for (int k = 0; k < N; ++k) { /* map_env.py:379-392, dict order */
        int a = order ? order[k] : k;
        int action = act[a];
        if (action < 0) continue;
        if (action <= 4) {
            int v0 = ACT_VEC[action][0], v1 = ACT_VEC[action][1], r0, r1;
            switch (ori[a]) { /* rotate_action map_env.py:701-716 */
                case 0: r0 = v0; r1 = v1; break;        /* UP */
                case 3: r0 = v1; r1 = -v0; break;       /* LEFT: rotate_left */
                case 1: r0 = -v1; r1 = v0; break;       /* RIGHT: rotate_right */
                default: r0 = -v0; r1 = -v1; break;     /* DOWN */
            }
            Cell t = {pos[a].r + r0, pos[a].c + r1};
            if (is_wall(e, t.r, t.c)) t = pos[a]; /* agent.py:105-113 */
            moves[a] = t; is_mover[a] = 1; mover[n_mov++] = a;
        } else if (action == 5) {
            ori[a] = (uint8_t)((ori[a] + 1) & 3); /* TURN_CLOCKWISE map_env.py:729-737 */
        } else if (action == 6) {
            ori[a] = (uint8_t)((ori[a] + 3) & 3); /* TURN_COUNTERCLOCKWISE map_env.py:720-728 */
        }
    }
    if (n_mov == 0) return; /* map_env.py:415 */

    /* np.random.shuffle(shuffle_list) map_env.py:421-423 */
    int shuf[ORC_MAX_AGENTS];
    if (d->tape) {
        const uint8_t* mo = d->tape->move_order + (size_t)d->b * N;
        unsigned seen = 0;
        for (int i = 0; i < n_mov; ++i) {
            int a = mo[i];
            if (a >= N || !is_mover[a] || (seen >> a & 1)) { d->err = -2; return; }
            seen |= 1u << a; shuf[i] = a;
        }
    } else {
        for (int i = 0; i < n_mov; ++i) shuf[i] = mover[i];
        uint32_t w = 0;
        for (int i = n_mov - 1; i >= 1; --i) {
            uint32_t j = (uint32_t)(((uint64_t)px_word(d, STREAM_MOVE, w++) * (uint32_t)(i + 1)) >> 32);
            int tmp = shuf[i]; shuf[i] = shuf[j]; shuf[j] = tmp;
        }
    }

    /* np.unique(move_slots, axis=0): distinct ORIGINAL targets in lexicographic order, map_env.py:424-426 */
    Cell orig[ORC_MAX_AGENTS], uniq[ORC_MAX_AGENTS];
    int n_uniq = 0;
    for (int a = 0; a < N; ++a) if (is_mover[a]) orig[a] = moves[a];
    for (int i = 0; i < n_mov; ++i) {
        Cell c = orig[shuf[i]];
        int found = 0;
        for (int u = 0; u < n_uniq; ++u) if (cell_eq(uniq[u], c)) found = 1;
        if (!found) uniq[n_uniq++] = c;
    }
    for (int i = 1; i < n_uniq; ++i) { /* insertion sort, (row, col) ascending */
        Cell c = uniq[i]; int j = i - 1;
        while (j >= 0 && (uniq[j].r > c.r || (uniq[j].r == c.r && uniq[j].c > c.c))) { uniq[j + 1] = uniq[j]; --j; }
        uniq[j + 1] = c;
    }

    for (int u = 0; u < n_uniq; ++u) { /* map_env.py:435-491 */
        Cell move = uniq[u];
        int cont[ORC_MAX_AGENTS], n_cont = 0;
        for (int i = 0; i < n_mov; ++i) if (cell_eq(orig[shuf[i]], move)) cont[n_cont++] = shuf[i];
        if (n_cont <= 1) continue;
        int cell_free = 1;
        for (int i = 0; i < n_cont; ++i) {
            int a = cont[i];
            if (occupied(pos, N, move)) {            /* :449 */
                int o = by_pos(pos, N, move);        /* :452 (rebuilt after every update => live) */
                Cell curr_pos = pos[a], cpos = pos[o];
                Cell cmove = is_mover[o] ? moves[o] : cpos; /* :456 */
                if (a == o) cell_free = 0;                                      /* (1) :460 */
                else if (!is_mover[o] || cell_eq(cpos, cmove)) cell_free = 0;   /* (2) :466 */
                else if (cell_eq(moves[o], curr_pos) && cell_eq(move, pos[o])) cell_free = 0; /* (3) :472 */
            }
        }
        if (cell_free) pos[cont[0]] = move;          /* :480-483: first in shuffled order wins */
        for (int i = 0; i < n_cont; ++i) moves[cont[i]] = pos[cont[i]]; /* :486-491 */
    }

    /* remaining moves, map_env.py:494-543; agent_moves is an ordered dict in action order */
    int alive[ORC_MAX_AGENTS] = {0}, n_alive = n_mov;
    for (int i = 0; i < n_mov; ++i) alive[mover[i]] = 1;
    while (n_alive > 0) {
        Cell snap_pos[ORC_MAX_AGENTS];
        int in_copy[ORC_MAX_AGENTS], deleted[ORC_MAX_AGENTS] = {0};
        int n0 = n_alive;
        for (int a = 0; a < N; ++a) { snap_pos[a] = pos[a]; in_copy[a] = alive[a]; }
        for (int i = 0; i < n_mov; ++i) {
            int a = mover[i];
            if (!in_copy[a] || deleted[a]) continue;
            Cell move = moves[a];
            if (occupied(pos, N, move)) {                       /* :503 live positions */
                int o = by_pos(snap_pos, N, move);              /* :506 pass-start snapshot */
                if (o < 0) { d->err = -4; return; }             /* reference would KeyError */
                Cell curr_pos = pos[a], cpos = pos[o];
                Cell cmove = alive[o] ? moves[o] : cpos;        /* :509 live agent_moves */
                if (a == o) { alive[a] = 0; deleted[a] = 1; --n_alive; }                    /* (1) */
                else if (!in_copy[o] || cell_eq(cpos, cmove)) { alive[a] = 0; deleted[a] = 1; --n_alive; } /* (2) */
                else if (cell_eq(moves[o], curr_pos) && cell_eq(move, pos[o])) {            /* (3) */
                    alive[o] = 0; alive[a] = 0; deleted[a] = deleted[o] = 1; n_alive -= 2;
                }
            } else {
                pos[a] = move; alive[a] = 0; deleted[a] = 1; --n_alive; /* :532-535 */
            }
        }
        if (n_alive == n0) { /* :540-543 */
            for (int i = 0; i < n_mov; ++i) if (alive[mover[i]]) pos[mover[i]] = moves[mover[i]];
            break;
        }
    }
}

/* ------------------------------------------------------------------ beams, map_env.py:545-649 */
typedef struct { int n; int16_t cell[ORC_MAX_AGENTS * 3 * 32]; uint8_t ch[ORC_MAX_AGENTS * 3 * 32]; } BeamList;

static void fire_beam(const OrcEnv* e, uint8_t* grid, const Cell* pos, const uint8_t* ori, int a,
                      int clean, int32_t* rew, BeamList* beams, int64_t* st) {
    const int N = e->N, H = e->H, W = e->W;
    const uint8_t ch = clean ? 'C' : 'F';
    int d0 = ORI_VEC[ori[a]][0], d1 = ORI_VEC[ori[a]][1];
    int rs0 = -d1, rs1 = d0; /* rotate_right(d) map_env.py:607,715 */
    Cell start[3] = {{pos[a].r, pos[a].c},
                     {pos[a].r + rs0 - d0, pos[a].c + rs1 - d1},
                     {pos[a].r - rs0 - d0, pos[a].c - rs1 - d1}}; /* :608-609 */
    int upd[3 * 32], n_upd = 0;
    for (int s = 0; s < 3; ++s) {
        Cell nc = {start[s].r + d0, start[s].c + d1};
        for (int i = 0; i < e->beam_len; ++i) {
            if (!(nc.r >= 0 && nc.r < H && nc.c >= 0 && nc.c < W) || grid[nc.r * W + nc.c] == '@') break; /* :615-616, :645 */
            int idx = nc.r * W + nc.c;
            if (occupied(pos, N, nc)) { /* :621-629 */
                int o = by_pos(pos, N, nc);
                if (!clean) { rew[o] -= 50; st[4]++; } /* agent.py:166-168 / 212-214 */
                beams->cell[beams->n] = (int16_t)idx; beams->ch[beams->n++] = ch;
                if (clean && grid[idx] == 'H') upd[n_upd++] = idx;
                break;
            }
            if (clean && grid[idx] == 'H') upd[n_upd++] = idx; /* :632-634 */
            beams->cell[beams->n] = (int16_t)idx; beams->ch[beams->n++] = ch; /* :636 */
            if (clean && grid[idx] == 'H') break; /* blocking_cells :639 */
            nc.r += d0; nc.c += d1;
        }
    }
    for (int i = 0; i < n_upd; ++i) { grid[upd[i]] = 'R'; st[5]++; } /* update_map :551-558 */
}

/* ------------------------------------------------------------------ spawning */
/* HarvestEnv.spawn_apples harvest.py:75-104 (3x3 window: j*j + k*k <= 2) */
static void harvest_spawn(const OrcEnv* e, uint8_t* grid, const Cell* pos, Draws* d, int64_t* st) {
    const int H = e->H, W = e->W;
    int16_t newp[e->n_apple + 1]; /* VLA: no allocator traffic on the threaded baseline */
    int n_new = 0;
    for (int i = 0; i < e->n_apple; ++i) {
        Cell p = {e->apple_pts[2 * i], e->apple_pts[2 * i + 1]};
        if (occupied(pos, e->N, p) || grid[p.r * W + p.c] == 'A') continue;
        int n = 0;
        for (int j = -2; j <= 2; ++j)
            for (int k = -2; k <= 2; ++k)
                if (j * j + k * k <= 2 && p.r + j >= 0 && p.r + j < H && p.c + k >= 0 && p.c + k < W &&
                    grid[(p.r + j) * W + p.c + k] == 'A') ++n;
        if (draw_less(e, d, e->harvest_prob[n < 3 ? n : 3])) newp[n_new++] = (int16_t)(p.r * W + p.c);
    }
    for (int i = 0; i < n_new; ++i) { grid[newp[i]] = 'A'; st[6]++; } /* harvest.py:72-73 */
}

/* CleanupEnv.custom_map_update cleanup.py:113-179 */
static void cleanup_spawn(const OrcEnv* e, uint8_t* grid, const Cell* pos, Draws* d, int64_t* st) {
    const int W = e->W;
    int h = 0;
    for (int i = 0; i < e->H * e->W; ++i) h += grid[i] == 'H'; /* compute_permitted_area :173-179 */
    if (h > e->area) h = e->area;
    double apple_p = e->apple_prob[h], waste_p = e->waste_prob[h]; /* compute_probabilities :156-171 */
    int16_t newp[e->n_apple + 2];
    int n_new = 0, waste_cell = -1;
    for (int i = 0; i < e->n_apple; ++i) { /* :135-141 */
        Cell p = {e->apple_pts[2 * i], e->apple_pts[2 * i + 1]};
        if (occupied(pos, e->N, p) || grid[p.r * W + p.c] == 'A') continue;
        if (draw_less(e, d, apple_p)) newp[n_new++] = (int16_t)(p.r * W + p.c);
    }
    if (waste_p != 0.0 && e->n_waste > 0) { /* not np.isclose(p, 0): p is 0 or 0.5 here  :144 */
        if (d->tape) {
            const uint16_t* wo = d->tape->waste_order + (size_t)d->b * e->n_waste;
            for (int i = 0; i < e->n_waste; ++i) { /* :146-153 */
                int idx = wo[i];
                if (grid[idx] != 'H' && draw_less(e, d, waste_p)) { waste_cell = idx; break; }
            }
        } else {
            /* random.shuffle replacement: canonical points ordered by (32-bit key, index) */
            int n = e->n_waste;
            uint64_t keyed[n];
            for (int i = 0; i < n; ++i) keyed[i] = ((uint64_t)px_word(d, STREAM_WASTE, (uint32_t)i) << 32) | (uint32_t)i;
            for (int i = 1; i < n; ++i) { uint64_t kx = keyed[i]; int j = i - 1; while (j >= 0 && keyed[j] > kx) { keyed[j + 1] = keyed[j]; --j; } keyed[j + 1] = kx; }
            for (int i = 0; i < n; ++i) {
                int w = (int)(keyed[i] & 0xffffffffu);
                int idx = e->waste_pts[2 * w] * W + e->waste_pts[2 * w + 1];
                if (grid[idx] != 'H' && draw_less(e, d, waste_p)) { waste_cell = idx; break; }
            }
        }
    }
    for (int i = 0; i < n_new; ++i) { grid[newp[i]] = 'A'; st[6]++; } /* :116 update_map */
    if (waste_cell >= 0) { grid[waste_cell] = 'H'; st[7]++; }
}

/* ------------------------------------------------------------------ rendering */
static uint8_t agent_char(int i) { /* str(int(agent_id[-1]) + 1) stored into a <U1 array, map_env.py:290,297 */
    int v = i % 10 + 1;
    return (uint8_t)(v == 10 ? '1' : '0' + v);
}

/* get_map_with_agents :280-302 -> return_view utility_funcs.py:59-114 -> map_to_colors :316-339
 * -> rotate_view :669-689 */
static void render(const OrcEnv* e, const uint8_t* grid, const Cell* pos, const uint8_t* ori,
                   const BeamList* beams, int rotate, uint8_t* obs) {
    const int H = e->H, W = e->W, N = e->N, r = e->r, V = e->V;
    uint8_t ov[H * W];
    memcpy(ov, grid, (size_t)H * W);
    for (int a = 0; a < N; ++a)
        if (pos[a].r >= 0 && pos[a].r < H && pos[a].c >= 0 && pos[a].c < W) ov[pos[a].r * W + pos[a].c] = agent_char(a);
    if (beams) for (int i = 0; i < beams->n; ++i) ov[beams->cell[i]] = beams->ch[i];
    for (int a = 0; a < N; ++a) {
        uint8_t* o = obs + (size_t)a * V * V * 3;
        int k = rotate ? ((4 - ori[a]) & 3) : 0; /* UP 0, LEFT 1, DOWN 2, RIGHT 3 */
        for (int i = 0; i < V; ++i)
            for (int j = 0; j < V; ++j) {
                int vi, vj; /* out[i][j] = view[vi][vj], np.rot90 */
                switch (k) {
                    case 0: vi = i; vj = j; break;
                    case 1: vi = j; vj = V - 1 - i; break;
                    case 2: vi = V - 1 - i; vj = V - 1 - j; break;
                    default: vi = V - 1 - j; vj = i; break;
                }
                int mr = pos[a].r - r + vi, mc = pos[a].c - r + vj;
                uint8_t ch = (mr >= 0 && mr < H && mc >= 0 && mc < W) ? ov[mr * W + mc] : (uint8_t)'0';
                const uint8_t* rgb = e->color[ch & 127];
                o[(i * V + j) * 3 + 0] = rgb[0]; o[(i * V + j) * 3 + 1] = rgb[1]; o[(i * V + j) * 3 + 2] = rgb[2];
            }
    }
}

/* ------------------------------------------------------------------ one env step, map_env.py:152-212 */
static int step_env(const OrcEnv* e, uint8_t* grid, int16_t* pos16, uint8_t* ori, const int8_t* act,
                    const uint8_t* order, Draws* d, uint8_t* obs, int32_t* rew, int64_t* st) {
    const int N = e->N, W = e->W;
    Cell pos[ORC_MAX_AGENTS];
    BeamList beams; beams.n = 0;
    for (int a = 0; a < N; ++a) { pos[a].r = pos16[2 * a]; pos[a].c = pos16[2 * a + 1]; rew[a] = 0; }
    for (int a = 0; a < N; ++a) {
        int mx = e->kind == ORC_KIND_CLEANUP ? 8 : (e->kind == ORC_KIND_HARVEST ? 7 : 6);
        if (act[a] > mx) return -1; /* action_map KeyError agent.py:162-164 / 201-203 */
    }
    update_moves(e, pos, ori, act, order, d); /* :176 */
    if (d->err) return d->err;
    for (int a = 0; a < N; ++a) { /* consume :178-181, agent.py:177-183 */
        int idx = pos[a].r * W + pos[a].c;
        if (grid[idx] == 'A') { rew[a] += 1; grid[idx] = ' '; st[2]++; }
    }
    for (int k = 0; k < N; ++k) { /* update_custom_moves :545-552 */
        int a = order ? order[k] : k;
        if (act[a] == 7) { rew[a] -= 1; st[3]++; fire_beam(e, grid, pos, ori, a, 0, rew, &beams, st); } /* harvest.py:62-67, cleanup.py:97-101 */
        else if (act[a] == 8) fire_beam(e, grid, pos, ori, a, 1, rew, &beams, st);                       /* cleanup.py:102-110 */
    }
    d->k = 0; d->spawn_stream = STREAM_SPAWN;
    if (e->kind == ORC_KIND_HARVEST) harvest_spawn(e, grid, pos, d, st);      /* :187 */
    else if (e->kind == ORC_KIND_CLEANUP) cleanup_spawn(e, grid, pos, d, st);
    if (d->err) return d->err;
    if (obs) render(e, grid, pos, ori, &beams, 1, obs); /* :189-199 */
    for (int a = 0; a < N; ++a) { pos16[2 * a] = (int16_t)pos[a].r; pos16[2 * a + 1] = (int16_t)pos[a].c; st[1] += rew[a]; }
    st[0]++;
    return 0;
}

/* ------------------------------------------------------------------ pthread parallel-for over envs */
typedef struct {
    void (*fn)(void* ctx, int b0, int b1, int tid);
    void* ctx; int b0, b1, tid;
} PfJob;
static void* pf_thread(void* p) { PfJob* j = (PfJob*)p; j->fn(j->ctx, j->b0, j->b1, j->tid); return 0; }
static void parallel_for(int B, int n_threads, void (*fn)(void*, int, int, int), void* ctx) {
    if (n_threads > ORC_MAX_THREADS) n_threads = ORC_MAX_THREADS;
    if (n_threads > B) n_threads = B;
    if (n_threads <= 1) { fn(ctx, 0, B, 0); return; }
    pthread_t th[ORC_MAX_THREADS]; PfJob jobs[ORC_MAX_THREADS];
    for (int i = 0; i < n_threads; ++i) {
        jobs[i].fn = fn; jobs[i].ctx = ctx; jobs[i].tid = i;
        jobs[i].b0 = (int)((int64_t)B * i / n_threads); jobs[i].b1 = (int)((int64_t)B * (i + 1) / n_threads);
        pthread_create(&th[i], 0, pf_thread, &jobs[i]);
    }
    for (int i = 0; i < n_threads; ++i) pthread_join(th[i], 0);
}

typedef struct {
    const OrcEnv* e; uint8_t* grid; int16_t* pos; uint8_t* ori; const int8_t* actions;
    const uint8_t* action_order; const OrcTape* tape; uint64_t seed, env_id0; uint32_t t;
    uint8_t* obs; int32_t* reward; int32_t* n_draws; int rotate;
    int err[ORC_MAX_THREADS]; int64_t st[ORC_MAX_THREADS][ORC_NUM_STATS];
} Job;

static void step_range(void* p, int b0, int b1, int tid) {
    Job* j = (Job*)p; const OrcEnv* e = j->e;
    const int N = e->N, HW = e->H * e->W;
    const size_t obs_sz = (size_t)N * e->V * e->V * 3;
    for (int b = b0; b < b1; ++b) {
        Draws d; memset(&d, 0, sizeof d);
        d.tape = j->tape; d.b = b;
        d.key[0] = (uint32_t)j->seed; d.key[1] = (uint32_t)(j->seed >> 32);
        d.env_id = (uint32_t)(j->env_id0 + (uint64_t)b); d.t = j->t;
        int rc = step_env(e, j->grid + (size_t)b * HW, j->pos + (size_t)b * N * 2, j->ori + (size_t)b * N,
                          j->actions + (size_t)b * N, j->action_order ? j->action_order + (size_t)b * N : 0, &d,
                          j->obs ? j->obs + (size_t)b * obs_sz : 0, j->reward + (size_t)b * N, j->st[tid]);
        if (j->n_draws) j->n_draws[b] = d.k;
        if (rc && !j->err[tid]) j->err[tid] = rc;
    }
}

int orc_step(const OrcEnv* e, int B, uint8_t* grid, int16_t* pos, uint8_t* ori, const int8_t* actions,
             const uint8_t* action_order, const OrcTape* tape, uint64_t seed, uint64_t env_id0,
             uint32_t t, uint8_t* obs, int32_t* reward, int32_t* n_draws, int64_t* stats, int n_threads) {
    Job* j = (Job*)calloc(1, sizeof(Job));
    j->e = e; j->grid = grid; j->pos = pos; j->ori = ori; j->actions = actions; j->action_order = action_order;
    j->tape = tape; j->seed = seed; j->env_id0 = env_id0; j->t = t; j->obs = obs; j->reward = reward; j->n_draws = n_draws;
    parallel_for(B, n_threads, step_range, j);
    int err = 0;
    for (int i = 0; i < ORC_MAX_THREADS; ++i) {
        if (j->err[i] && !err) err = j->err[i];
        if (stats) for (int k = 0; k < ORC_NUM_STATS; ++k) stats[k] += j->st[i][k];
    }
    free(j);
    return err;
}

/* ------------------------------------------------------------------ reset, map_env.py:214-249 */
static int reset_env(const OrcEnv* e, uint8_t* grid, int16_t* pos16, uint8_t* ori, Draws* d, uint8_t* obs) {
    const int N = e->N, HW = e->H * e->W, S = e->n_spawn;
    Cell pos[ORC_MAX_AGENTS];
    int64_t st[ORC_NUM_STATS] = {0};
    /* setup_agents harvest.py:46-55 / cleanup.py:118-130 */
    for (int a = 0; a < N; ++a) {
        /* spawn_point map_env.py:651-662: shuffle, then the LAST free entry of the shuffled list;
         * shuffle replacement orders the canonical list by (key, index) => max (key, index) among free */
        int best = -1; uint64_t best_key = 0;
        for (int s = 0; s < S; ++s) {
            Cell p = {e->spawn_pts[2 * s], e->spawn_pts[2 * s + 1]};
            if (occupied(pos, a, p)) continue;
            uint64_t kx = ((uint64_t)px_word(d, STREAM_RPOINT, (uint32_t)(a * S + s)) << 32) | (uint32_t)s;
            if (best < 0 || kx > best_key) { best = s; best_key = kx; }
        }
        if (best < 0) return -5; /* AssertionError map_env.py:661 */
        pos[a].r = e->spawn_pts[2 * best]; pos[a].c = e->spawn_pts[2 * best + 1];
        /* spawn_rotation map_env.py:664-667: randint(4) indexes ['LEFT','RIGHT','UP','DOWN'] */
        static const uint8_t ROT[4] = {3, 1, 0, 2};
        ori[a] = ROT[px_word(d, STREAM_RROT, (uint32_t)a) & 3];
    }
    /* reset_map :560-564, build_walls :691-694, custom_reset harvest.py:57-60 / cleanup.py:84-92 */
    for (int i = 0; i < HW; ++i) {
        uint8_t b = e->base_map[i], c = ' ';
        if (b == '@') c = '@';
        else if (e->kind == ORC_KIND_HARVEST && b == 'A') c = 'A';
        else if (e->kind == ORC_KIND_CLEANUP && (b == 'H' || b == 'R' || b == 'S')) c = b;
        grid[i] = c;
    }
    d->k = 0; d->spawn_stream = STREAM_RSPAWN; /* custom_map_update :230 */
    if (e->kind == ORC_KIND_HARVEST) harvest_spawn(e, grid, pos, d, st);
    else if (e->kind == ORC_KIND_CLEANUP) cleanup_spawn(e, grid, pos, d, st);
    if (obs) render(e, grid, pos, ori, 0, 0, obs); /* :232-248, no rotate_view */
    for (int a = 0; a < N; ++a) { pos16[2 * a] = (int16_t)pos[a].r; pos16[2 * a + 1] = (int16_t)pos[a].c; }
    return d->err;
}

static void reset_range(void* p, int b0, int b1, int tid) {
    Job* j = (Job*)p; const OrcEnv* e = j->e;
    const int N = e->N, HW = e->H * e->W;
    const size_t obs_sz = (size_t)N * e->V * e->V * 3;
    for (int b = b0; b < b1; ++b) {
        Draws d; memset(&d, 0, sizeof d);
        d.b = b; d.key[0] = (uint32_t)j->seed; d.key[1] = (uint32_t)(j->seed >> 32);
        d.env_id = (uint32_t)(j->env_id0 + (uint64_t)b); d.t = j->t;
        int rc = reset_env(e, j->grid + (size_t)b * HW, j->pos + (size_t)b * N * 2, j->ori + (size_t)b * N, &d,
                           j->obs ? j->obs + (size_t)b * obs_sz : 0);
        if (rc && !j->err[tid]) j->err[tid] = rc;
    }
}

int orc_reset(const OrcEnv* e, int B, uint8_t* grid, int16_t* pos, uint8_t* ori, uint64_t seed,
              uint64_t env_id0, uint32_t t, uint8_t* obs, int n_threads) {
    Job* j = (Job*)calloc(1, sizeof(Job));
    j->e = e; j->grid = grid; j->pos = pos; j->ori = ori; j->seed = seed; j->env_id0 = env_id0; j->t = t; j->obs = obs;
    parallel_for(B, n_threads, reset_range, j);
    int err = 0;
    for (int i = 0; i < ORC_MAX_THREADS; ++i) if (j->err[i] && !err) err = j->err[i];
    free(j);
    return err;
}

static void render_range(void* p, int b0, int b1, int tid) {
    Job* j = (Job*)p; const OrcEnv* e = j->e; (void)tid;
    const int N = e->N, HW = e->H * e->W;
    const size_t obs_sz = (size_t)N * e->V * e->V * 3;
    for (int b = b0; b < b1; ++b) {
        Cell c[ORC_MAX_AGENTS];
        for (int a = 0; a < N; ++a) { c[a].r = j->pos[((size_t)b * N + a) * 2]; c[a].c = j->pos[((size_t)b * N + a) * 2 + 1]; }
        render(e, j->grid + (size_t)b * HW, c, j->ori + (size_t)b * N, 0, j->rotate, j->obs + (size_t)b * obs_sz);
    }
}

int orc_render(const OrcEnv* e, int B, const uint8_t* grid, const int16_t* pos, const uint8_t* ori,
               int rotate, uint8_t* obs, int n_threads) {
    Job* j = (Job*)calloc(1, sizeof(Job));
    j->e = e; j->grid = (uint8_t*)grid; j->pos = (int16_t*)pos; j->ori = (uint8_t*)ori; j->rotate = rotate; j->obs = obs;
    parallel_for(B, n_threads, render_range, j);
    free(j);
    return 0;
}

/* ------------------------------------------------------------------ construction */
OrcEnv* orc_create(int kind, int H, int W, int N, int view_radius, int beam_len, const uint8_t* base_map,
                   const uint8_t* color_lut, const double* harvest_prob, const double* apple_prob,
                   const double* waste_prob, int area, const int16_t* spawn_points, int n_spawn) {
    if (N < 1 || N > ORC_MAX_AGENTS || H < 1 || W < 1 || beam_len > 32) return 0;
    OrcEnv* e = (OrcEnv*)calloc(1, sizeof(OrcEnv));
    e->kind = kind; e->H = H; e->W = W; e->N = N; e->r = view_radius; e->V = 2 * view_radius + 1;
    e->beam_len = beam_len; e->area = area; e->n_spawn = n_spawn;
    e->base_map = (uint8_t*)malloc((size_t)H * W);
    memcpy(e->base_map, base_map, (size_t)H * W);
    memcpy(e->color, color_lut, 128 * 3);
    if (harvest_prob) memcpy(e->harvest_prob, harvest_prob, sizeof e->harvest_prob);
    if (kind == ORC_KIND_CLEANUP) {
        e->apple_prob = (double*)malloc(sizeof(double) * (area + 1));
        e->waste_prob = (double*)malloc(sizeof(double) * (area + 1));
        memcpy(e->apple_prob, apple_prob, sizeof(double) * (area + 1));
        memcpy(e->waste_prob, waste_prob, sizeof(double) * (area + 1));
    }
    e->apple_pts = (int16_t*)malloc(sizeof(int16_t) * 2 * H * W);
    e->waste_pts = (int16_t*)malloc(sizeof(int16_t) * 2 * H * W);
    e->spawn_pts = (int16_t*)malloc(sizeof(int16_t) * 2 * (n_spawn + 1));
    memcpy(e->spawn_pts, spawn_points, sizeof(int16_t) * 2 * n_spawn);
    uint8_t apple_ch = kind == ORC_KIND_HARVEST ? 'A' : (kind == ORC_KIND_CLEANUP ? 'B' : 0);
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            uint8_t b = base_map[r * W + c];
            if (apple_ch && b == apple_ch) { e->apple_pts[2 * e->n_apple] = (int16_t)r; e->apple_pts[2 * e->n_apple + 1] = (int16_t)c; e->n_apple++; }
            if (kind == ORC_KIND_CLEANUP && (b == 'H' || b == 'R')) { e->waste_pts[2 * e->n_waste] = (int16_t)r; e->waste_pts[2 * e->n_waste + 1] = (int16_t)c; e->n_waste++; }
        }
    return e;
}

void orc_destroy(OrcEnv* e) {
    if (!e) return;
    free(e->base_map); free(e->apple_prob); free(e->waste_prob);
    free(e->apple_pts); free(e->waste_pts); free(e->spawn_pts); free(e);
}
int orc_num_apple_points(const OrcEnv* e) { return e->n_apple; }
int orc_num_waste_points(const OrcEnv* e) { return e->n_waste; }
int orc_max_draws(const OrcEnv* e) { return e->n_apple + e->n_waste; }
