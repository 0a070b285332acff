// Device functions of the phases of MapEnv.step, shared by the general kernel (ssd_step_general.cu) and the
// specialised full-step kernel (ssd_step_fast.cu):
//
//   A   moves (map_env.py:357-543), consume (:178-181), beams (:545-649): ONE LANE PER AGENT, a group of
//       G = 8 or 16 lanes per env.  Conflict-free moves are resolved with shuffles; an env with any
//       contested / occupied target runs the literal emulation of update_moves, its group of lanes working
//       together (moves_coop).  A firing agent's three rays walk on 3 lanes.
//   B   custom_map_update (harvest.py:69-104, cleanup.py:113-179): the whole warp per env, ballot/popc
//       prefix ranks give every eligible cell its sequential draw index
//   C   get_map_with_agents + return_view + map_to_colors + rotate_view (map_env.py:189-199): one lane per
//       VIEW ROW; the warp packs 32 rows to their exact byte offsets in a private staging buffer
//
// Reference citations are relative to the reference root (social_dilemmas/envs/...).
#pragma once
#include <cstdio>
#include <cstdlib>

#include "ssd_device.cuh"
#include "ssd_internal.h"

namespace ssd {

__device__ __forceinline__ int tile_idx(const StepArgs& a, uint32_t key) {
    return static_cast<int>(key >> 8) * a.Ws + static_cast<int>(key & 255);
}

__device__ __forceinline__ bool is_apple(uint8_t c) { return (c & kCodeMask) == CB(C_APPLE); }

// Harvest keeps min(3, #apple neighbours) in the low two bits of every ' ' / 'A' cell (ssd_internal.h).
// Recompute it for the cell at `q` from the codes around it; called for the 8 neighbours of every
// cell that gained or lost an apple.
__device__ __forceinline__ void recount(uint8_t* q, int Ws) {
    const uint8_t c = *q, code = c & kCodeMask;
    if (code == CB(C_EMPTY) || code == CB(C_APPLE)) {
        int n = is_apple(q[-Ws - 1]) + is_apple(q[-Ws]) + is_apple(q[-Ws + 1]) + is_apple(q[-1]) + is_apple(q[1]) +
                is_apple(q[Ws - 1]) + is_apple(q[Ws]) + is_apple(q[Ws + 1]);
        *q = static_cast<uint8_t>((c & 0xFC) | (n < 3 ? n : 3));
    }
}
// All lanes call this with the warp-wide ballot of lanes whose cell base[my_off] just changed its apple
// state (`base` is warp-uniform); lanes 0..7 refresh the eight neighbours of one event at a time.
__device__ __forceinline__ void recount_events(uint32_t events, int my_off, uint8_t* base, int Ws) {
    const int lane = threadIdx.x & 31;
    const int t = lane & 7;
    const int nb = (t < 3 ? -Ws - 1 + t : (t == 3 ? -1 : (t == 4 ? 1 : Ws - 6 + t)));  // -Ws-1,-Ws,-Ws+1,-1,+1,Ws-1,Ws,Ws+1
    while (events) {
        const int src = __ffs(events) - 1;
        events &= events - 1;
        const int off = __shfl_sync(0xffffffffu, my_off, src);
        if (lane < 8) recount(base + off + nb, Ws);
        __syncwarp();
    }
}

// ---- orchard bitmaps (StepArgs::orch): the specialised Harvest kernel keeps, per env, one bit per apple point for "holds no
// apple" (emp) and "holds none and SPAWN_PROB[neighbour count] != 0" (need), so that spawn_apples only looks at candidates.
struct OrchTables {
    const uint32_t* pt_mask;  // shared memory: bit c % 32 of word c / 32 -- tile cell c is an apple point
    const uint16_t* pt_pre;   //                apple points in the words before
    int nW, stride, nz;       // words per bitmap, words per env, harvest_nz
};
// apple-point index of tile cell `cell`, or -1
__device__ __forceinline__ int cell_point(const OrchTables& T, int cell) {
    const uint32_t m = T.pt_mask[cell >> 5], b = cell & 31;
    return (m >> b) & 1u ? static_cast<int>(T.pt_pre[cell >> 5]) + __popc(m & ((1u << b) - 1u)) : -1;
}
// recount() for a cell of env `bm` (its bitmaps), plus the `need` bit when the cell is an empty apple point
__device__ __forceinline__ void recount_bm(uint8_t* q, int Ws, int cell, uint32_t* bm, const OrchTables& T) {
    const uint8_t c = *q, code = c & kCodeMask;
    if (code == CB(C_EMPTY) || code == CB(C_APPLE)) {
        int n = is_apple(q[-Ws - 1]) + is_apple(q[-Ws]) + is_apple(q[-Ws + 1]) + is_apple(q[-1]) + is_apple(q[1]) +
                is_apple(q[Ws - 1]) + is_apple(q[Ws]) + is_apple(q[Ws + 1]);
        n = n < 3 ? n : 3;
        *q = static_cast<uint8_t>((c & 0xFC) | n);
        if (code == CB(C_EMPTY)) {
            const int p = cell_point(T, cell);
            if (p >= 0) {
                const uint32_t bit = 1u << (p & 31);
                if ((T.nz >> n) & 1) atomicOr(&bm[T.nW + (p >> 5)], bit); else atomicAnd(&bm[T.nW + (p >> 5)], ~bit);
            }
        }
    }
}
// recount_events() with bitmaps: my_slot / my_cell = env slot and in-tile cell of this lane's event
__device__ __forceinline__ void recount_events_bm(uint32_t events, int my_slot, int my_cell, uint8_t* tiles0 /* cell 0 of slot 0 */, int tile_pitch,
                                                  int Ws, uint32_t* orch, const OrchTables& T) {
    const int lane = threadIdx.x & 31;
    const int t = lane & 7;
    const int nb = (t < 3 ? -Ws - 1 + t : (t == 3 ? -1 : (t == 4 ? 1 : Ws - 6 + t)));  // -Ws-1,-Ws,-Ws+1,-1,+1,Ws-1,Ws,Ws+1
    while (events) {
        const int src = __ffs(events) - 1;
        events &= events - 1;
        const int slot = __shfl_sync(0xffffffffu, my_slot, src), cell = __shfl_sync(0xffffffffu, my_cell, src);
        if (lane < 8) recount_bm(tiles0 + slot * tile_pitch + cell + nb, Ws, cell + nb, orch + slot * T.stride, T);
        __syncwarp();
    }
}

struct Counters {  // per-lane event counts, reduced per warp at the end of the kernel
    int steps, eaten, fires, hits, cleaned, apples, waste;
};

// ====================================================================== phase A: moves
__device__ __forceinline__ int by_pos(const uint16_t* p, int N, uint32_t key) {
    int o = -1;  // dict built in agent order: the LAST agent on a cell wins (map_env.py:397)
#pragma unroll 1
    for (int a = 0; a < N; ++a) o = (p[a] == key) ? a : o;
    return o;
}

// Literal emulation of the conflict resolution of update_moves (map_env.py:394-543) for one env, run by ALL lanes of the env's
// group together (lane = agent): positions, targets and the frozen targets stay in registers, `x in self.agent_pos` and
// `agent_by_pos[x]` are one ballot each, and the control flow -- contested cells in key order, then the fix-point passes in
// action order, exactly the reference's loops -- is uniform across the group.  (A single lane walking shared-memory arrays took
// up to 40 000 cycles for this, and a step kernel is as slow as its slowest warp.)  `gm` is the group's lane mask; every lane of
// the group must call this, converged.
template <bool TAPE, class ES>
__device__ __noinline__ void moves_coop(const StepArgs& a, ES& S, MoveScratch& M, uint32_t movers, int local_env, const PhiloxKey& pk,
                                        uint32_t& pos, uint32_t tgt, bool valid, int al, int G, uint32_t gm) {
    const int N = a.N;
    const int gshift = (threadIdx.x & 31) & ~(G - 1);
    const bool mover = valid && ((movers >> al) & 1u);
    // np.random.shuffle of the movers (action order), map_env.py:421-423: lane 0 of the group draws, everybody reads its rank
    if (al == 0) {
        uint8_t* shuf = M.shuf;
        int n_mov = 0;
#pragma unroll 1
        for (int k = 0; k < N; ++k) {
            const int ag = S.order[k];
            if (movers >> ag & 1) shuf[n_mov++] = static_cast<uint8_t>(ag);
        }
        if (TAPE) {
            const uint8_t* mo = a.tape_move + static_cast<size_t>(local_env) * N;
#pragma unroll 1
            for (int i = 0; i < n_mov; ++i) shuf[i] = mo[i] < N ? mo[i] : static_cast<uint8_t>(N - 1);  // malformed tapes must not fault
        } else {
            uint4 blk = make_uint4(0, 0, 0, 0);
            uint32_t w = 0;
#pragma unroll 1
            for (int i = n_mov - 1; i >= 1; --i, ++w) {
                if ((w & 3) == 0) blk = philox4x32_10(pk.env, pk.t, STREAM_MOVE, w >> 2, pk.k0, pk.k1);
                const uint32_t j = __umulhi(pick_word(blk, w), static_cast<uint32_t>(i + 1));
                const uint8_t tmp = shuf[i]; shuf[i] = shuf[j]; shuf[j] = tmp;
            }
        }
#pragma unroll 1
        for (int i = 0; i < n_mov; ++i) M.orig[shuf[i]] = static_cast<uint16_t>(i);  // rank of every mover in the shuffled order
    }
    __syncwarp(gm);
    const uint32_t rank = mover ? M.orig[al] : 0xFFu;
    const uint32_t orig = mover ? tgt : 0xFFFFFFFFu;  // search_list: targets frozen before the contested pass (:426)
    auto occupied_now = [&](uint32_t cell) { return __ballot_sync(gm, valid && pos == cell) != 0u; };          // x in self.agent_pos
    auto by_pos_of = [&](uint32_t key, uint32_t cell) {                                                        // the LAST agent on a cell wins (:397)
        const uint32_t m = (__ballot_sync(gm, valid && key == cell) & gm) >> gshift;
        return m ? 31 - __clz(m) : -1;
    };

    // contested cells in lexicographic (row, col) order == ascending key (np.unique axis=0, :424)
    int prev = -1;
    while (true) {
        const uint32_t cand = (mover && static_cast<int>(orig) > prev) ? orig : 0x10000u;
        const uint32_t cell = __reduce_min_sync(gm, cand);
        if (cell == 0x10000u) break;
        prev = static_cast<int>(cell);
        const bool in_cell = mover && orig == cell;
        const uint32_t same = __ballot_sync(gm, in_cell);
        if (__popc(same) < 2) continue;
        bool bad = false;  // conditions (1)-(3) of :449-478 for this agent; nothing they read changes inside the reference's loop
        if (occupied_now(cell)) {
            const int o = by_pos_of(pos, cell);
            const uint32_t o_tgt = __shfl_sync(gm, tgt, o, G);
            const bool o_moves = (movers >> o) & 1u;
            // cpos == cell: o stands on the cell
            bad = in_cell && (al == o || !o_moves || o_tgt == cell || o_tgt == pos);
        }
        const bool cell_free = __ballot_sync(gm, bad) == 0u;
        const int winner = static_cast<int>(__reduce_min_sync(gm, in_cell ? (rank << 8 | static_cast<uint32_t>(al)) : 0xFFFFu) & 255u);  // first in shuffled order (:481)
        if (cell_free && al == winner) pos = cell;  // :480-483
        if (in_cell) tgt = pos;                      // :486-491 (the winner included: everybody is a "stay" now)
    }

    // remaining moves: fix-point loop, map_env.py:494-543
    uint32_t alive = movers;
    while (alive) {
        const uint32_t snap = pos;       // agent_by_pos snapshot (:495)
        const uint32_t in_copy = alive;  // moves_copy (:498)
        uint32_t deleted = 0;
#pragma unroll 1
        for (int k = 0; k < N; ++k) {
            const int ag = S.order[k];
            if (!(in_copy >> ag & 1) || (deleted >> ag & 1)) continue;
            const uint32_t mv = __shfl_sync(gm, tgt, ag, G), ag_pos = __shfl_sync(gm, pos, ag, G);
            if (occupied_now(mv)) {                     // :503 live positions
                const int o = by_pos_of(snap, mv);      // :506 snapshot
                if (o < 0) continue;                    // reference would KeyError; unreachable
                const uint32_t cpos = __shfl_sync(gm, pos, o, G), o_tgt = __shfl_sync(gm, tgt, o, G);
                const uint32_t cmove = (alive >> o & 1) ? o_tgt : cpos;  // :509 live agent_moves
                if (ag == o) { alive &= ~(1u << ag); deleted |= 1u << ag; }                                    // (1)
                else if (!(in_copy >> o & 1) || cpos == cmove) { alive &= ~(1u << ag); deleted |= 1u << ag; }  // (2)
                else if (o_tgt == ag_pos && mv == cpos) {                                                      // (3)
                    alive &= ~((1u << ag) | (1u << o)); deleted |= (1u << ag) | (1u << o);
                }
            } else {
                if (al == ag) pos = mv;                 // :532-535
                alive &= ~(1u << ag); deleted |= 1u << ag;
            }
        }
        if (alive == in_copy) {  // nobody could move freely: move them all (:540-543)
            if (valid && (alive >> al & 1)) pos = tgt;
            break;
        }
    }
}

struct AgentLane {
    uint32_t key;  // row << 8 | col
    int ori, act, rew;
};

// update_moves for one group of G lanes (= one env).  All 32 lanes of the warp call this.
template <bool TAPE, class ES>
__device__ __forceinline__ void moves_group(const StepArgs& a, ES& S, MoveScratch& M, const uint8_t* g, AgentLane& me, bool valid,
                                            int al, int G, int local_env, const PhiloxKey& pk) {
    const int act = me.act;
    bool mover = false;
    uint32_t tgt = me.key;
    if (valid && act >= 0) {  // map_env.py:379-392
        if (act <= 4) {
            const int v0 = (act == 0) ? -1 : (act == 1) ? 1 : 0;  // ACTIONS map_env.py:11-15
            const int v1 = (act == 2) ? -1 : (act == 3) ? 1 : 0;
            const int o = me.ori;
            int r0, r1;  // rotate_action map_env.py:701-716
            if (o == 0) { r0 = v0; r1 = v1; } else if (o == 3) { r0 = v1; r1 = -v0; }
            else if (o == 1) { r0 = -v1; r1 = v0; } else { r0 = -v0; r1 = -v1; }
            const uint32_t nkey = static_cast<uint32_t>((static_cast<int>(me.key >> 8) + r0) << 8 | (static_cast<int>(me.key & 255) + r1));
            tgt = (g[tile_idx(a, nkey)] == CB(C_WALL)) ? me.key : nkey;  // agent.py:105-113 you can't walk through walls
            mover = true;
        } else if (act == 5) {
            me.ori = (me.ori + 1) & 3;  // TURN_CLOCKWISE map_env.py:729-737
        } else if (act == 6) {
            me.ori = (me.ori + 3) & 3;  // TURN_COUNTERCLOCKWISE map_env.py:720-728
        }
    }
    // Fast path: all targets distinct and no target currently occupied by ANOTHER agent => the
    // contested pass is empty and the first fix-point pass moves everybody (a STAY hits rule (1)
    // and keeps its place).  Anything else runs the literal emulation.
    const uint32_t pos_x = valid ? me.key : 0xFFFF0000u | al;  // never equal to a real cell
    const uint32_t tgt_x = mover ? tgt : 0xFFFE0000u | al;
    bool conflict = false;
    for (int d = 1; d < G; ++d) {
        const int src = (al + d) & (G - 1);
        const uint32_t pos_y = __shfl_sync(0xffffffffu, pos_x, src, G);
        const uint32_t tgt_y = __shfl_sync(0xffffffffu, tgt_x, src, G);
        conflict |= mover && (pos_y == tgt || tgt_y == tgt);
    }
    const int gshift = (threadIdx.x & 31) & ~(G - 1);
    const uint32_t gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << gshift;
    const uint32_t conf_all = __ballot_sync(0xffffffffu, conflict);
    const uint32_t conf = conf_all & gmask;
    const uint32_t movers = (__ballot_sync(0xffffffffu, mover) & gmask) >> gshift;
    if (conf == 0 && mover) me.key = tgt;
    if (conf != 0 && !SSD_SKIP(a.debug, 16))  // all lanes of a group with a conflict: the literal emulation, run by the group together
        moves_coop<TAPE, ES>(a, S, M, movers, local_env, pk, me.key, tgt, valid, al, G, gmask);
    __syncwarp();
}

// One ray of a beam (map_env.py:566-649); the three rays of a firing agent walk on lanes 0..2 of
// its group.  The map is wall-enclosed (checked by ssd_create), so the reference's bounds test
// (:615) can never fire before the wall test (:616).  Agent cells carry kFlag, so the position
// table is only searched when a ray actually runs into somebody.  Returns the painted cell count.
// ray_probe has no side effects: it returns the number of painted cells, the cell a CLEAN beam turns into river (`upd`, -1: none)
// and the agent an 'F' beam hits (`hit`, -1: none; a ray stops at the first agent, so there is at most one).
template <class ES>
__device__ __forceinline__ int ray_probe(const StepArgs& a, const ES& S, const uint8_t* g, uint32_t key, int ori, int s,
                                         bool clean, int& upd, int& hit) {
    const int d0 = (ori == 1) - (ori == 3), d1 = (ori == 2) - (ori == 0);  // ORIENTATIONS map_env.py:19-22
    int r = static_cast<int>(key >> 8) + d0, c = static_cast<int>(key & 255) + d1;  // :608-613
    if (s == 1) { r += -d1 - d0; c += d0 - d1; }  // start + rotate_right(d) - d   (:607-609)
    if (s == 2) { r -= -d1 + d0; c -= d0 + d1; }  // start - rotate_right(d) - d
    const int dp = d0 * a.Ws + d1;
    int p = r * a.Ws + c;
    int n = 0;
    for (int i = 0; i < a.beam_len; ++i) {
        const uint8_t raw = g[p], cell = raw & 0x7F;
        if (cell == CB(C_WALL)) break;                          // :616
        const bool isH = clean && cell == CB(C_WASTE);
        if (raw & kFlag) {                                      // :621-629 agents absorb beams
            if (!clean) hit = by_pos(S.pos, a.N, static_cast<uint32_t>(r << 8 | c));  // agent.py:166-168, 212-214
            ++n;                                                // :624
            if (isH) upd = p;                                   // :625-628
            break;
        }
        if (isH) upd = p;                                       // :632-634
        ++n;                                                    // :636
        if (isH) break;                                         // blocking_cells :639
        r += d0; c += d1; p += dp;
    }
    return n;
}
template <class ES, bool ATOMIC = false>
__device__ __forceinline__ int ray_walk(const StepArgs& a, ES& S, uint8_t* g, uint32_t key, int ori, int s,
                                        bool clean, int& upd, int& hits) {
    int hit = -1;
    const int n = ray_probe(a, S, g, key, ori, s, clean, upd, hit);
    if (hit >= 0) {
        if constexpr (ATOMIC) atomicAdd(&S.rew[hit], -50); else S.rew[hit] -= 50;
        ++hits;
    }
    return n;
}

// ====================================================================== phase B: spawning (one warp per env)
// Agent cells are flagged with bit 7 while the spawn pass runs ("[row, col] not in self.agent_pos",
// harvest.py:90, cleanup.py:138); consume already turned every apple under an agent into ' '.
template <bool TAPE>
__device__ __forceinline__ void harvest_spawn(const StepArgs& a, uint8_t* g, const uint16_t* s_apple, uint32_t* list,
                                              int local_env, const PhiloxKey& pk, int lane, Counters& cnt) {
    const int n_apple = a.n_apple;
    constexpr uint8_t A = CB(C_APPLE);
    // One scan in row-major apple-point order (harvest.py:87-101).  The apple table is padded to a
    // multiple of 32 with a harmless interior cell, so every lane loads unconditionally.  `base`
    // counts eligible points: the k-th eligible point consumes the k-th np.random.rand.  The number
    // of apples in the 3x3 window (harvest.py:92-100) is cached in the low bits of the cell, so the
    // scan is one byte load per point.  Only points with SPAWN_PROB[n] != 0 can spawn; they are
    // compacted into `list` as cell | n << 16 | draw index << 18 and drawn for afterwards.
    int base = 0, n_need = 0;
#pragma unroll 1
    for (int i0 = 0; i0 < n_apple; i0 += 32) {
        const int i = i0 + lane;
        const uint32_t cell = s_apple[i];
        const uint8_t c = g[cell];
        const bool el = (i < n_apple) & ((c & 0xFC) != A) & (c < kFlag);  // not an apple, no agent on it (harvest.py:90)
        const uint32_t m = __ballot_sync(0xffffffffu, el);
        const int n = c & 3;
        const bool need = el & ((a.harvest_nz >> n) & 1);
        const uint32_t m2 = __ballot_sync(0xffffffffu, need);
        if (need) list[n_need + __popc(m2 & lanemask_lt())] = cell | static_cast<uint32_t>(n) << 16 |
                                                               static_cast<uint32_t>(base + __popc(m & lanemask_lt())) << 18;
        base += __popc(m);
        n_need += __popc(m2);
    }
    if (TAPE && a.n_draws_out != nullptr && lane == 0) a.n_draws_out[local_env] = base;
    __syncwarp();  // every count was read from the pre-spawn grid; writes happen after the scan (harvest.py:72-73)
#pragma unroll 1
    for (int j0 = 0; j0 < n_need; j0 += 32) {
        const int j = j0 + lane;
        bool spawn = false;
        uint32_t en = 0;
        if (j < n_need) {
            en = list[j];
            const int n = (en >> 16) & 3;
            const uint32_t k = en >> 18;
            if (TAPE) spawn = a.tape_u[static_cast<size_t>(local_env) * a.u_stride + k] < a.harvest_p[n];
            else spawn = philox_u53(pk, a.spawn_stream, k) < a.harvest_thr[n];  // u < p  <=>  u53 < ceil(p * 2^53)
            if (spawn) { g[en & 0xffffu] = static_cast<uint8_t>(A | (g[en & 0xffffu] & 3)); ++cnt.apples; }  // keep the CURRENT cached count
        }
        const uint32_t ms = __ballot_sync(0xffffffffu, spawn);
        if (ms) {  // refresh the cached counts around the new apples
            __syncwarp();
            recount_events(ms, static_cast<int>(en & 0xffffu), g, a.Ws);
        }
    }
}

// spawn_apples for all four envs of a warp (specialised kernel).  The scans run env by env, two groups of
// 32 apple points per trip so that their loads overlap; the candidates of ALL envs go into one list
// (cell | n << 16 | env slot << 18 | draw index << 20), so the Philox draws of the whole warp are one
// or two passes instead of one per env.  A list that could overflow is drained between two scans --
// never inside one: every count must be read from the pre-spawn grid of its env.
template <bool TAPE>
__device__ __forceinline__ void harvest_drain(const StepArgs& a, uint8_t* tiles, int tile_pitch, const uint32_t* list, int n_list,
                                              int we, PhiloxKey pk, int lane, Counters& cnt) {
    constexpr uint8_t A = CB(C_APPLE);
    __syncwarp();
#pragma unroll 1
    for (int j0 = 0; j0 < n_list; j0 += 32) {
        const int j = j0 + lane;
        bool spawn = false;
        int off = 0;
        if (j < n_list) {
            const uint32_t en = list[j];
            const int n = (en >> 16) & 3, slot = (en >> 18) & 3;
            const uint32_t k = en >> 20;
            off = a.pad_bytes + slot * tile_pitch + static_cast<int>(en & 0xffffu);
            if (TAPE) {
                spawn = a.tape_u[static_cast<size_t>(we + slot) * a.u_stride + k] < a.harvest_p[n];
            } else {
                pk.env = static_cast<uint32_t>(a.env_id0 + static_cast<uint64_t>(we + slot));
                spawn = philox_u53(pk, a.spawn_stream, k) < a.harvest_thr[n];  // u < p  <=>  u53 < ceil(p * 2^53)
            }
            if (spawn) { tiles[off] = static_cast<uint8_t>(A | (tiles[off] & 3)); ++cnt.apples; }  // keep the CURRENT cached count: an earlier batch may have refreshed it
        }
        const uint32_t ms = __ballot_sync(0xffffffffu, spawn);
        if (ms) {  // refresh the cached counts around the new apples
            __syncwarp();
            recount_events(ms, off, tiles, a.Ws);
        }
    }
    __syncwarp();
}

template <bool TAPE, int EPW>
__device__ __forceinline__ void harvest_spawn_warp(const StepArgs& a, uint8_t* tiles, int tile_pitch, const uint16_t* __restrict__ s_apple,
                                                   uint32_t* __restrict__ list, int cap, int we, const PhiloxKey& pk, int lane, Counters& cnt) {
    const int n_apple = a.n_apple;
    constexpr uint8_t A = CB(C_APPLE);
    const uint32_t lt = lanemask_lt();
    int n_list = 0;
#pragma unroll 1
    for (int q = 0; q < EPW; ++q) {
        if (n_list + n_apple > cap) { harvest_drain<TAPE>(a, tiles, tile_pitch, list, n_list, we, pk, lane, cnt); n_list = 0; }
        const uint8_t* __restrict__ g = tiles + a.pad_bytes + q * tile_pitch;
        int base = 0;
#pragma unroll 1
        for (int i0 = 0; i0 < n_apple; i0 += 64) {  // the table is padded to a multiple of 64 points
            const int i = i0 + lane;
            const uint32_t cell0 = s_apple[i], cell1 = s_apple[i + 32];
            const uint8_t c0 = g[cell0], c1 = g[cell1];
            const bool el0 = (i < n_apple) & ((c0 & 0xFC) != A) & (c0 < kFlag);  // not an apple, no agent on it (harvest.py:90)
            const bool el1 = (i + 32 < n_apple) & ((c1 & 0xFC) != A) & (c1 < kFlag);
            const uint32_t m0 = __ballot_sync(0xffffffffu, el0), m1 = __ballot_sync(0xffffffffu, el1);
            const int n0 = c0 & 3, n1 = c1 & 3;  // cached count of apples in the 3x3 window (harvest.py:92-100)
            const bool need0 = el0 & ((a.harvest_nz >> n0) & 1), need1 = el1 & ((a.harvest_nz >> n1) & 1);
            const uint32_t w0 = __ballot_sync(0xffffffffu, need0), w1 = __ballot_sync(0xffffffffu, need1);
            const int base1 = base + __popc(m0), nl1 = n_list + __popc(w0);
            if (need0) list[n_list + __popc(w0 & lt)] = cell0 | static_cast<uint32_t>(n0) << 16 | static_cast<uint32_t>(q) << 18 |
                                                          static_cast<uint32_t>(base + __popc(m0 & lt)) << 20;
            if (need1) list[nl1 + __popc(w1 & lt)] = cell1 | static_cast<uint32_t>(n1) << 16 | static_cast<uint32_t>(q) << 18 |
                                                      static_cast<uint32_t>(base1 + __popc(m1 & lt)) << 20;
            base = base1 + __popc(m1);
            n_list = nl1 + __popc(w1);
        }
        if (TAPE && a.n_draws_out != nullptr && lane == 0) a.n_draws_out[we + q] = base;
    }
    harvest_drain<TAPE>(a, tiles, tile_pitch, list, n_list, we, pk, lane, cnt);
}

// spawn_apples (harvest.py:75-104) for all envs of a warp from the orchard bitmaps.  Lane L < EPW * nW owns word g = L % nW
// of env slot q = L / nW: e = points that hold no apple and no agent (the agents' bits are masked out by the caller) -- the
// eligible points, whose count before a point is the index of its np.random.rand draw -- and c = the candidates among them
// (SPAWN_PROB != 0).  Only candidates are looked at: one list entry each (point | n << 12 | slot << 14 | draw index << 16),
// then the same drain as above with the bitmaps kept current.  All reads of an env happen before its first write.
template <bool TAPE, int EPW>
__device__ __forceinline__ void harvest_spawn_bm(const StepArgs& a, uint8_t* tiles, int tile_pitch, const uint16_t* __restrict__ s_apple,
                                                 uint32_t* orch, const OrchTables& T, uint32_t* __restrict__ list, int cap, int we,
                                                 PhiloxKey pk, int lane, Counters& cnt) {
    constexpr uint8_t A = CB(C_APPLE);
    const int nW = T.nW;
    const uint32_t lt = lanemask_lt();
    const int q = lane / nW, g = lane - q * nW;
    const bool mine = q < EPW;
    uint8_t* const tiles0 = tiles + a.pad_bytes;
    uint32_t e = 0, c = 0;
    if (mine) { const uint32_t* bm = orch + q * T.stride; e = bm[g]; c = bm[nW + g] & e; }
    const int cnt_e = __popc(e);
    int base = cnt_e;  // draws before this word's points: exclusive scan over the nW lanes of the env
    for (int d = 1; d < nW; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, base, d);
        if (g >= d) base += v;
    }
    if (TAPE && a.n_draws_out != nullptr && mine && g == nW - 1) a.n_draws_out[we + q] = base;
    base -= cnt_e;
    // all candidates of the warp in one list when they fit; env by env otherwise (one env's candidates always fit)
    const int total = __reduce_add_sync(0xffffffffu, __popc(c));
    const int n_batch = total <= cap ? 1 : EPW;
#pragma unroll 1
    for (int b = 0; b < n_batch; ++b) {
        uint32_t cc = (n_batch == 1 || q == b) ? c : 0u;
        int n_list = 0;
#pragma unroll 1
        while (true) {
            const uint32_t m = __ballot_sync(0xffffffffu, cc != 0);
            if (m == 0) break;
            if (cc != 0) {
                const int bit = __ffs(cc) - 1;
                cc &= cc - 1;
                const int p = 32 * g + bit;
                const uint32_t n = tiles0[q * tile_pitch + s_apple[p]] & 3u;  // cached count of apples in the 3x3 window (harvest.py:92-100)
                const uint32_t k = static_cast<uint32_t>(base + __popc(e & ((1u << bit) - 1u)));
                list[n_list + __popc(m & lt)] = static_cast<uint32_t>(p) | n << 12 | static_cast<uint32_t>(q) << 14 | k << 16;
            }
            n_list += __popc(m);
        }
        __syncwarp();
#pragma unroll 1
        for (int j0 = 0; j0 < n_list; j0 += 32) {
            const int j = j0 + lane;
            bool spawn = false;
            int slot = 0, cell = 0;
            if (j < n_list) {
                const uint32_t en = list[j];
                const int p = en & 0xfffu, n = (en >> 12) & 3;
                const uint32_t k = en >> 16;
                slot = (en >> 14) & 3;
                cell = s_apple[p];
                if (TAPE) {
                    spawn = a.tape_u[static_cast<size_t>(we + slot) * a.u_stride + k] < a.harvest_p[n];
                } else {
                    pk.env = static_cast<uint32_t>(a.env_id0 + static_cast<uint64_t>(we + slot));
                    spawn = philox_u53(pk, a.spawn_stream, k) < a.harvest_thr[n];  // u < p  <=>  u53 < ceil(p * 2^53)
                }
                if (spawn) {
                    uint8_t* t = tiles0 + slot * tile_pitch + cell;
                    *t = static_cast<uint8_t>(A | (*t & 3));  // keep the CURRENT cached count: an earlier batch may have refreshed it
                    uint32_t* bm = orch + slot * T.stride;
                    const uint32_t bit = 1u << (p & 31);
                    atomicAnd(&bm[p >> 5], ~bit);
                    atomicAnd(&bm[nW + (p >> 5)], ~bit);
                    ++cnt.apples;
                }
            }
            const uint32_t ms = __ballot_sync(0xffffffffu, spawn);
            if (ms) {  // refresh the cached counts (and `need` bits) around the new apples
                __syncwarp();
                recount_events_bm(ms, slot, cell, tiles0, tile_pitch, a.Ws, orch, T);
            }
        }
        __syncwarp();
    }
}

// Number of 'H' cells of one env tile (all lanes get the sum).
__device__ __forceinline__ int count_waste(const StepArgs& a, const uint8_t* g, int lane) {
    int nh = 0;
    for (int i = lane * 16; i < a.env_bytes; i += 512) {
        const uint4 v = *reinterpret_cast<const uint4*>(g + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t x = (w[q] & 0x7F7F7F7Fu) ^ (0x01010101u * CB(C_WASTE));  // byte == 0  <=>  cell is 'H'
            const uint32_t nz = ((x + 0x7F7F7F7Fu) | x) & 0x80808080u;      // bit 7 set  <=>  byte != 0
            nh += 4 - __popc(nz);
        }
    }
    return __reduce_add_sync(0xffffffffu, nh);
}

// hcount: the env's running number of 'H' cells (specialised kernel: kept in HBM next to the grid, StepArgs::orch word 0, and
// adjusted by every clean / waste spawn), or nullptr to count them here.
template <bool TAPE>
__device__ __forceinline__ void cleanup_spawn(const StepArgs& a, uint8_t* g, const uint16_t* s_apple, uint32_t* keys,
                                              int local_env, const PhiloxKey& pk, int lane, Counters& cnt, int* hcount = nullptr) {
    const int n_apple = a.n_apple, n_waste = a.n_waste;
    // compute_permitted_area / compute_probabilities, cleanup.py:156-179: the number of 'H' cells of the whole grid
    int h = hcount != nullptr ? *hcount : count_waste(a, g, lane);
    h = h < a.area ? h : a.area;
    const double apple_p = a.apple_p[h], waste_p = a.waste_p[h];
    const uint64_t apple_thr = a.apple_thr[h], waste_thr = a.waste_thr[h];

    // Nothing can spawn above the depletion threshold (both probabilities are 0, cleanup.py:160-162): without a tape
    // to account draws for, the scans below have no effect.  Under random play this is the common case.
    if (!TAPE && apple_thr == 0 && (waste_thr == 0 || n_waste == 0)) return;

    int base = 0;
    if (TAPE) {
        for (int i0 = 0; i0 < n_apple; i0 += 32) {  // apple pass, cleanup.py:135-141 (a draw per eligible point)
            const int i = i0 + lane;
            bool el = false;
            int idx = 0;
            if (i < n_apple) { idx = s_apple[i]; const uint8_t c = g[idx]; el = (c != CB(C_APPLE)) && !(c & kFlag); }
            const uint32_t m = __ballot_sync(0xffffffffu, el);
            if (el) {
                const int rank = base + __popc(m & lanemask_lt());
                if (a.tape_u[static_cast<size_t>(local_env) * a.u_stride + rank] < apple_p) { g[idx] = CB(C_APPLE); ++cnt.apples; }
            }
            base += __popc(m);
        }
    } else {
        // Two groups of 32 points per trip (the table is padded to 64).  Their <= 64 draws are the words of <= 33 Philox
        // blocks: lane L evaluates block (base >> 1) + L ONCE and parks it in `keys` (free until the waste pass), every
        // eligible lane picks its half -- half the Philox evaluations of one per group.
        uint4* const blocks = reinterpret_cast<uint4*>(keys);
        const uint32_t lt = lanemask_lt();
#pragma unroll 1
        for (int i0 = 0; i0 < n_apple; i0 += 64) {
            const int i = i0 + lane;
            const int idx0 = s_apple[i], idx1 = s_apple[i + 32];
            const uint8_t c0 = g[idx0], c1 = g[idx1];
            const bool el0 = (i < n_apple) & (c0 != CB(C_APPLE)) & (c0 < kFlag);  // no apple, no agent (cleanup.py:138)
            const bool el1 = (i + 32 < n_apple) & (c1 != CB(C_APPLE)) & (c1 < kFlag);
            const uint32_t m0 = __ballot_sync(0xffffffffu, el0), m1 = __ballot_sync(0xffffffffu, el1);
            const int cnt0 = __popc(m0), total = cnt0 + __popc(m1);
            if (apple_thr != 0 && total != 0) {
                const uint32_t first = static_cast<uint32_t>(base) >> 1;
                blocks[lane] = philox4x32_10(pk.env, pk.t, a.spawn_stream, first + lane, pk.k0, pk.k1);
                __syncwarp();
                const uint32_t r0 = base + __popc(m0 & lt), r1 = base + cnt0 + __popc(m1 & lt);
                auto draw = [&](uint32_t r) -> uint64_t {  // k-th 53-bit uniform = words (2k, 2k+1) of the stream
                    if ((r >> 1) - first < 32u) {
                        const uint2 w = reinterpret_cast<const uint2*>(blocks)[r - 2 * first];
                        return (static_cast<uint64_t>(w.x >> 5) << 26) | (w.y >> 6);
                    }
                    return philox_u53(pk, a.spawn_stream, r);  // draw 64 of a trip that started on an odd index
                };
                if (el0 && draw(r0) < apple_thr) { g[idx0] = CB(C_APPLE); ++cnt.apples; }  // apple cells are never read again
                if (el1 && draw(r1) < apple_thr) { g[idx1] = CB(C_APPLE); ++cnt.apples; }
                __syncwarp();
            }
            base += total;
        }
    }

    if (TAPE && a.n_draws_out != nullptr && lane == 0) a.n_draws_out[local_env] = base;  // apple pass; the waste pass adds its own
    if (waste_p != 0.0 && n_waste > 0) {  // `not np.isclose(p, 0)`: p is 0 or wasteSpawnProbability (cleanup.py:144)
        if (TAPE) {
            const uint16_t* wo = a.tape_waste + static_cast<size_t>(local_env) * n_waste;  // order after random.shuffle (:145)
            for (int i0 = 0; i0 < n_waste; i0 += 32) {
                const int i = i0 + lane;
                bool el = false;
                int idx = 0;
                if (i < n_waste) {  // tape cells are row*W+col of the reference's map
                    const int cell = wo[i];
                    idx = (cell / a.W) * a.Ws + cell % a.W;
                    el = (g[idx] & 0x7F) != CB(C_WASTE);  // :149
                }
                const uint32_t m = __ballot_sync(0xffffffffu, el);
                bool ok = false;
                if (el) ok = a.tape_u[static_cast<size_t>(local_env) * a.u_stride + base + __popc(m & lanemask_lt())] < waste_p;
                const uint32_t s = __ballot_sync(0xffffffffu, ok);
                if (s) {  // first success spawns and breaks (:151-153); waste may appear under an agent
                    if (lane == __ffs(s) - 1) { g[idx] = CB(C_WASTE) | (g[idx] & kFlag); ++cnt.waste; if (hcount != nullptr) ++*hcount; }
                    base += __popc(m & ((2u << (__ffs(s) - 1)) - 1u));  // draws up to and including the winner
                    break;
                }
                base += __popc(m);
            }
            if (a.n_draws_out != nullptr && lane == 0) a.n_draws_out[local_env] = base;
        } else {
            // random.shuffle replacement: canonical waste points ordered by (32-bit key, index).
            __syncwarp();
            for (int i0 = 0; i0 < n_waste; i0 += 128) {  // one Philox block = keys of 4 consecutive points
                const int i = i0 + lane * 4;
                if (i < n_waste) {
                    const uint4 k4 = philox4x32_10(pk.env, pk.t, STREAM_WASTE, i >> 2, pk.k0, pk.k1);
                    keys[i] = k4.x;
                    if (i + 1 < n_waste) keys[i + 1] = k4.y;
                    if (i + 2 < n_waste) keys[i + 2] = k4.z;
                    if (i + 3 < n_waste) keys[i + 3] = k4.w;
                }
            }
            // eligibility (not 'H', cleanup.py:149) of the points lane, lane + 32, ... as a bit mask per lane
            uint32_t el_bits = 0;
            int el_cnt = 0;
            for (int i = lane, k = 0; i < n_waste; i += 32, ++k) {
                const bool el = (g[a.waste_cell[i]] & 0x7F) != CB(C_WASTE);
                el_bits |= static_cast<uint32_t>(el) << (k & 31);
                el_cnt += el;
            }
            const int n_el = __reduce_add_sync(0xffffffffu, el_cnt);
            __syncwarp();
            // the k-th scanned non-'H' cell draws uniform #(base + k); the first success wins.  32 candidates k at a time.
            int kstar = n_el;
            for (int k0 = 0; k0 < n_el; k0 += 32) {
                const bool ok = k0 + lane < n_el && philox_u53(pk, a.spawn_stream, base + k0 + lane) < waste_thr;
                const uint32_t sm = __ballot_sync(0xffffffffu, ok);
                if (sm) { kstar = k0 + __ffs(sm) - 1; break; }
            }
            if (kstar < n_el) {
                uint64_t prev = 0;  // select the (kstar+1)-th smallest (key, index) among the eligible cells
                bool first = true;
                for (int it = 0; it <= kstar; ++it) {
                    uint64_t best = ~0ull;
                    if (n_waste <= 1024) {
                        for (uint32_t bits = el_bits; bits; bits &= bits - 1) {
                            const int i = lane + 32 * (__ffs(bits) - 1);
                            const uint64_t kx = static_cast<uint64_t>(keys[i]) << 32 | static_cast<uint32_t>(i);
                            if ((first || kx > prev) && kx < best) best = kx;
                        }
                    } else {  // more than 32 points per lane: the mask wrapped, test the cells again
                        for (int i = lane; i < n_waste; i += 32) {
                            if ((g[a.waste_cell[i]] & 0x7F) == CB(C_WASTE)) continue;
                            const uint64_t kx = static_cast<uint64_t>(keys[i]) << 32 | static_cast<uint32_t>(i);
                            if ((first || kx > prev) && kx < best) best = kx;
                        }
                    }
                    prev = warp_min_u64(best);
                    first = false;
                }
                if (lane == 0) {
                    const int idx = a.waste_cell[static_cast<uint32_t>(prev)];
                    g[idx] = CB(C_WASTE) | (g[idx] & kFlag);
                    ++cnt.waste;
                    if (hcount != nullptr) ++*hcount;
                }
            }
        }
    }
}

// ====================================================================== phase C: rendering
// Per-agent window geometry (np.rot90 index algebra of rotate_view map_env.py:669-689 folded with
// return_view utility_funcs.py:59-114): view pixel (i, j) reads the warp-tile byte a0 + i*si + j*sj.
// Tiles are framed by >= r*Ws + r zero bytes and rows end in r zero bytes, so no pixel needs a
// bounds test: everything outside the map reads as C_PAD.
__device__ __forceinline__ uint2 view_param(const StepArgs& a, const EnvScratch& S, int tile_off, int ag) {
    const int pr = S.pos[ag] >> 8, pc = S.pos[ag] & 255, r = a.r, Ws = a.Ws;
    const int k = a.rotate ? ((4 - S.ori[ag]) & 3) : 0;  // UP 0, LEFT 1, DOWN 2, RIGHT 3
    int a0, si, sj;
    if (k == 0)      { a0 = (pr - r) * Ws + pc - r; si = Ws;  sj = 1; }    // cell (pr-r+i, pc-r+j)
    else if (k == 2) { a0 = (pr + r) * Ws + pc + r; si = -Ws; sj = -1; }   // cell (pr+r-i, pc+r-j)
    else if (k == 1) { a0 = (pr - r) * Ws + pc + r; si = -1;  sj = Ws; }   // cell (pr-r+j, pc+r-i)
    else             { a0 = (pr + r) * Ws + pc - r; si = 1;   sj = -Ws; }  // cell (pr+r-j, pc-r+i)
    return make_uint2(static_cast<uint32_t>(a0 + tile_off), (static_cast<uint32_t>(si) & 0xffffu) | static_cast<uint32_t>(sj) << 16);
}

__device__ __forceinline__ uint32_t cell_color(const uint32_t* s_color, uint8_t cell) {  // table indexed by the grid byte
    return s_color[cell];
}

// One lane renders one row of one agent's view (V pixels = 3V bytes).  The warp owns `total_rows`
// consecutive rows of the obs tensor starting at `dst` (4-byte aligned); 32 rows = 96V bytes are
// packed to their exact byte offsets in the warp's staging buffer (3V is odd, so consecutive rows
// start at byte phases 0,1,2,3: each lane owns the 32-bit words whose FIRST byte lies in its row
// and takes the first pixel of the next row from the neighbouring lane).  The staging buffer is
// shifted by (dst & 15) so that shared and global addresses are congruent mod 16 and the body of
// every chunk leaves with 16-byte stores.
template <int VT>
__device__ __forceinline__ void render_rows(const uint2* s_view, const uint8_t* tiles, const uint32_t* s_color,
                                            uint32_t* stage, uint8_t* dst, int total_rows) {
    constexpr int RB = 3 * VT;           // bytes per view row
    constexpr int NP = (RB + 3 + 3) / 4; // words covering the row plus the next row's first pixel
    const int lane = threadIdx.x & 31;
    const int mis = static_cast<int>(reinterpret_cast<uintptr_t>(dst) & 15);  // multiple of 4
    uint32_t* st = stage + (mis >> 2);
    for (int base = 0; base < total_rows; base += 32) {
        const int R = base + lane;
        uint32_t X[VT + 2];
#pragma unroll
        for (int j = 0; j < VT + 2; ++j) X[j] = 0;
        if (R < total_rows) {
            const int ga = R / VT, i = R - ga * VT;  // rows are ordered (env, agent, i)
            const uint2 vp = s_view[ga];
            const int si = static_cast<int16_t>(vp.y & 0xffffu), sj = static_cast<int32_t>(vp.y) >> 16;
            const uint8_t* g = tiles + static_cast<int32_t>(vp.x) + i * si;
#pragma unroll
            for (int j = 0; j < VT; ++j) X[j] = cell_color(s_color, g[j * sj]);
        }
        X[VT] = __shfl_down_sync(0xffffffffu, X[0], 1);
        uint32_t P[NP + 1];
#pragma unroll
        for (int w = 0; w < NP; ++w) {
            const int p = (4 * w) / 3, ph = (4 * w) % 3;
            P[w] = __byte_perm(X[p], X[p + 1], ph == 0 ? 0x4210u : (ph == 1 ? 0x5421u : 0x6542u));
        }
        P[NP] = 0;
        if (R < total_rows) {
            const uint32_t o = static_cast<uint32_t>(lane) * RB;  // the chunk starts word aligned
            const uint32_t d = (4 - (o & 3)) & 3;
            const uint32_t w0 = (o + 3) >> 2, w1 = (o + RB - 1) >> 2;
            const int M = w1 - w0 + 1;
#pragma unroll
            for (int m = 0; m < NP; ++m)
                if (m < M) st[w0 + m] = __funnelshift_r(P[m], P[m + 1], 8 * d);
        }
        __syncwarp();
        const int nbytes = min(32, total_rows - base) * RB;  // multiple of 4
        uint8_t* out = dst + static_cast<size_t>(base) * RB - mis;  // 16-byte aligned
        const uint8_t* sb = reinterpret_cast<const uint8_t*>(stage);
#pragma unroll
        for (int it = 0; it < (32 * RB + 16 + 511) / 512; ++it) {
            const int off = it * 512 + lane * 16;  // slot [off, off+16) of the shifted chunk [mis, mis+nbytes)
            if (off >= mis && off + 16 <= mis + nbytes) {
                *reinterpret_cast<uint4*>(out + off) = *reinterpret_cast<const uint4*>(sb + off);
            } else if (off + 16 > mis && off < mis + nbytes) {
#pragma unroll
                for (int w = 0; w < 16; w += 4)
                    if (off + w >= mis && off + w < mis + nbytes)
                        *reinterpret_cast<uint32_t*>(out + off + w) = *reinterpret_cast<const uint32_t*>(sb + off + w);
            }
        }
        __syncwarp();
    }
}

// Any view size / partially valid warps: one lane per pixel, byte stores straight to HBM.
__device__ __forceinline__ void render_generic(const StepArgs& a, const EnvScratch* envs, const uint2* s_view,
                                               const uint8_t* tiles, const uint32_t* s_color, uint8_t* dst, int n_envs) {
    const int V = a.V, N = a.N;
    const int total = n_envs * N * V * V;
    for (int p = threadIdx.x & 31; p < total; p += 32) {
        const int j = p % V, i = (p / V) % V, ga = p / (V * V);
        if (!envs[ga / N].active) continue;
        const uint2 vp = s_view[ga];
        const int si = static_cast<int16_t>(vp.y & 0xffffu), sj = static_cast<int32_t>(vp.y) >> 16;
        const uint32_t c = cell_color(s_color, tiles[static_cast<int32_t>(vp.x) + i * si + j * sj]);
        dst[3 * static_cast<size_t>(p)] = c & 255; dst[3 * static_cast<size_t>(p) + 1] = (c >> 8) & 255; dst[3 * static_cast<size_t>(p) + 2] = (c >> 16) & 255;
    }
}

__device__ __forceinline__ uint8_t agent_cell(int i) {  // str(int(agent_id[-1]) + 1) in a <U1 array (map_env.py:290,297)
    const int v = i % 10;  // '1'..'9', and agent-9 / agent-19 alias '1' ('10' truncated)
    return CB(static_cast<uint8_t>(C_AGENT + (v == 9 ? 0 : v)));
}


}  // namespace ssd
