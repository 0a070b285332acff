// Fused MapEnv.step kernel for sm_100a.  Every WARP steps its own 32/G environments end to end
// (no CTA barrier after the table load), so warps in different phases overlap freely:
//
//   load   the warp's grid tiles by TMA bulk copies into zero-framed shared-memory tiles
//          (ssd_internal.h), agent table and actions straight into registers (one lane per agent)
//   A      moves (map_env.py:357-543), consume (:178-181), beams (:545-649): ONE LANE PER AGENT, a
//          group of G = 8 (or 16) lanes per env.  Conflict-free moves are resolved with shuffles; an
//          env with any contested / occupied target falls back to the literal sequential emulation
//          of update_moves on the group's first lane.  A firing agent's three rays walk on 3 lanes.
//   B      custom_map_update (harvest.py:69-104, cleanup.py:113-179): the whole warp per env,
//          ballot/popc prefix ranks give every eligible cell its sequential draw index
//   store  grid, agent table, rewards back to HBM
//   C      get_map_with_agents + return_view + map_to_colors + rotate_view (map_env.py:189-199):
//          one lane per VIEW ROW; the warp packs 32 rows to their exact byte offsets in a private
//          staging buffer and writes them with coalesced 16-byte stores
//
// Reference citations are relative to the reference root (social_dilemmas/envs/...).
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

struct Counters {  // per-lane event counts, reduced per warp at the end of the kernel
    int steps, eaten, fires, hits, cleaned, apples, waste;
};

// ====================================================================== phase A: moves
__device__ __forceinline__ bool occupied(const uint16_t* p, int N, uint32_t key) {
    bool f = false;  // `x in self.agent_pos` (map_env.py:251-253)
    for (int a = 0; a < N; ++a) f |= (p[a] == key);
    return f;
}
__device__ __forceinline__ int by_pos(const uint16_t* p, int N, uint32_t key) {
    int o = -1;  // dict built in agent order: the LAST agent on a cell wins (map_env.py:397)
    for (int a = 0; a < N; ++a) o = (p[a] == key) ? a : o;
    return o;
}

// Literal emulation of the conflict resolution of update_moves (map_env.py:394-543) for one env,
// run by a single lane.  S.pos / S.tgt hold positions and wall-clipped targets of the movers.
template <bool TAPE, class ES>
__device__ __noinline__ void moves_slow(const StepArgs& a, ES& S, MoveScratch& M, uint32_t movers, int local_env,
                                         const PhiloxKey& pk) {
    const int N = a.N;
    // mover list in action order (agent_moves is an insertion-ordered dict, map_env.py:400-412)
    uint8_t* shuf = M.shuf;
    int n_mov = 0;
    for (int k = 0; k < N; ++k) {
        const int ag = S.order[k];
        if (movers >> ag & 1) shuf[n_mov++] = static_cast<uint8_t>(ag);
    }
    // np.random.shuffle(shuffle_list), map_env.py:421-423
    if (TAPE) {
        const uint8_t* mo = a.tape_move + static_cast<size_t>(local_env) * N;
        for (int i = 0; i < n_mov; ++i) shuf[i] = mo[i] < N ? mo[i] : static_cast<uint8_t>(N - 1);  // malformed tapes must not fault
    } else {
        uint4 blk = make_uint4(0, 0, 0, 0);
        uint32_t w = 0;
        for (int i = n_mov - 1; i >= 1; --i, ++w) {
            if ((w & 3) == 0) blk = philox4x32_10(pk.env, pk.t, STREAM_MOVE, w >> 2, pk.k0, pk.k1);
            const uint32_t j = __umulhi(pick_word(blk, w), static_cast<uint32_t>(i + 1));
            const uint8_t tmp = shuf[i]; shuf[i] = shuf[j]; shuf[j] = tmp;
        }
    }
    for (int ag = 0; ag < N; ++ag) M.orig[ag] = (movers >> ag & 1) ? M.tgt[ag] : 0xFFFFu;

    // contested cells in lexicographic (row, col) order == ascending key (np.unique axis=0, :424)
    int prev = -1;
    while (true) {
        int cell = 0x10000, cnt = 0;
        for (int ag = 0; ag < N; ++ag) {
            const int o = M.orig[ag];
            if (o != 0xFFFF && o > prev) {
                if (o < cell) { cell = o; cnt = 1; } else if (o == cell) ++cnt;
            }
        }
        if (cell == 0x10000) break;
        prev = cell;
        if (cnt < 2) continue;
        bool cell_free = true;
        int winner = -1;
        for (int i = 0; i < n_mov; ++i) {  // conflicting agents in shuffled order (:441-442)
            const int ag = shuf[i];
            if (M.orig[ag] != cell) continue;
            if (winner < 0) winner = ag;  // agent_to_slot[index]: first occurrence (:481)
            if (occupied(S.pos, N, cell)) {                       // :449
                const int o = by_pos(S.pos, N, cell);             // :452 (rebuilt after every update)
                const uint32_t cpos = S.pos[o];
                const bool o_moves = movers >> o & 1;
                const uint32_t cmove = o_moves ? M.tgt[o] : cpos; // :456
                if (ag == o) cell_free = false;                                   // (1) :460
                else if (!o_moves || cpos == cmove) cell_free = false;            // (2) :466
                else if (M.tgt[o] == S.pos[ag] && cell == (int)S.pos[o]) cell_free = false;  // (3) :472
            }
        }
        if (cell_free) S.pos[winner] = static_cast<uint16_t>(cell);  // :480-483
        for (int i = 0; i < n_mov; ++i) {                            // :486-491
            const int ag = shuf[i];
            if (M.orig[ag] == cell) M.tgt[ag] = S.pos[ag];
        }
    }

    // remaining moves: fix-point loop, map_env.py:494-543
    uint32_t alive = movers;
    while (alive) {
        for (int ag = 0; ag < N; ++ag) M.snap[ag] = S.pos[ag];  // agent_by_pos snapshot (:495)
        const uint32_t in_copy = alive;                         // moves_copy (:498)
        uint32_t deleted = 0;
        for (int k = 0; k < N; ++k) {
            const int ag = S.order[k];
            if (!(in_copy >> ag & 1) || (deleted >> ag & 1)) continue;
            const uint32_t mv = M.tgt[ag];
            if (occupied(S.pos, N, mv)) {                   // :503 live positions
                const int o = by_pos(M.snap, N, mv);        // :506 snapshot
                if (o < 0) continue;                        // reference would KeyError; unreachable
                const uint32_t cpos = S.pos[o];
                const uint32_t cmove = (alive >> o & 1) ? M.tgt[o] : cpos;  // :509 live agent_moves
                if (ag == o) { alive &= ~(1u << ag); deleted |= 1u << ag; }                                  // (1)
                else if (!(in_copy >> o & 1) || cpos == cmove) { alive &= ~(1u << ag); deleted |= 1u << ag; }  // (2)
                else if (M.tgt[o] == S.pos[ag] && mv == S.pos[o]) {                                           // (3)
                    alive &= ~((1u << ag) | (1u << o)); deleted |= (1u << ag) | (1u << o);
                }
            } else {
                S.pos[ag] = static_cast<uint16_t>(mv);      // :532-535
                alive &= ~(1u << ag); deleted |= 1u << ag;
            }
        }
        if (alive == in_copy) {  // nobody could move freely: move them all (:540-543)
            for (int ag = 0; ag < N; ++ag) if (alive >> ag & 1) S.pos[ag] = M.tgt[ag];
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
    if (conf_all != 0) {  // warp-uniform branch: groups without a conflict just keep the barriers company
        if (conf != 0 && valid) { S.pos[al] = static_cast<uint16_t>(me.key); M.tgt[al] = static_cast<uint16_t>(tgt); }
        __syncwarp();
        if (conf != 0 && al == 0 && !SSD_SKIP(a.debug, 16)) moves_slow<TAPE, ES>(a, S, M, movers, local_env, pk);
        __syncwarp();
        if (conf != 0 && valid) me.key = S.pos[al];
    }
}

// One ray of a beam (map_env.py:566-649); the three rays of a firing agent walk on lanes 0..2 of
// its group.  The map is wall-enclosed (checked by ssd_create), so the reference's bounds test
// (:615) can never fire before the wall test (:616).  Agent cells carry kFlag, so the position
// table is only searched when a ray actually runs into somebody.  Returns the painted cell count.
template <class ES, bool ATOMIC = false>
__device__ __forceinline__ int ray_walk(const StepArgs& a, ES& S, uint8_t* g, uint32_t key, int ori, int s,
                                        bool clean, int& upd, int& hits) {
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
            if (!clean) {  // agent.py:166-168, 212-214
                const int v = by_pos(S.pos, a.N, static_cast<uint32_t>(r << 8 | c));
                if constexpr (ATOMIC) atomicAdd(&S.rew[v], -50); else S.rew[v] -= 50;
                ++hits;
            }
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

template <bool TAPE>
__device__ __forceinline__ void cleanup_spawn(const StepArgs& a, uint8_t* g, const uint16_t* s_apple, uint32_t* keys,
                                              int local_env, const PhiloxKey& pk, int lane, Counters& cnt) {
    const int n_apple = a.n_apple, n_waste = a.n_waste;
    // compute_permitted_area / compute_probabilities, cleanup.py:156-179: count 'H' over the whole grid
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
    int h = __reduce_add_sync(0xffffffffu, nh);
    h = h < a.area ? h : a.area;
    const double apple_p = a.apple_p[h], waste_p = a.waste_p[h];
    const uint64_t apple_thr = a.apple_thr[h], waste_thr = a.waste_thr[h];

    int base = 0;
    for (int i0 = 0; i0 < n_apple; i0 += 32) {  // apple pass, cleanup.py:135-141 (a draw per eligible point)
        const int i = i0 + lane;
        bool el = false;
        int idx = 0;
        if (i < n_apple) { idx = s_apple[i]; const uint8_t c = g[idx]; el = (c != CB(C_APPLE)) && !(c & kFlag); }
        const uint32_t m = __ballot_sync(0xffffffffu, el);
        if (el) {
            const int rank = base + __popc(m & lanemask_lt());
            bool spawn;
            if (TAPE) spawn = a.tape_u[static_cast<size_t>(local_env) * a.u_stride + rank] < apple_p;
            else spawn = apple_thr != 0 && philox_u53(pk, a.spawn_stream, rank) < apple_thr;
            if (spawn) { g[idx] = CB(C_APPLE); ++cnt.apples; }  // apple cells are never read again in this pass
        }
        base += __popc(m);
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
                    if (lane == __ffs(s) - 1) { g[idx] = CB(C_WASTE) | (g[idx] & kFlag); ++cnt.waste; }
                    base += __popc(m & ((2u << (__ffs(s) - 1)) - 1u));  // draws up to and including the winner
                    break;
                }
                base += __popc(m);
            }
            if (a.n_draws_out != nullptr && lane == 0) a.n_draws_out[local_env] = base;
        } else {
            // random.shuffle replacement: canonical waste points ordered by (32-bit key, index).
            int n_el = 0;
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
            for (int i = lane; i < n_waste; i += 32) n_el += (g[a.waste_cell[i]] & 0x7F) != CB(C_WASTE);
            n_el = __reduce_add_sync(0xffffffffu, n_el);
            __syncwarp();
            // the k-th scanned non-'H' cell draws uniform #(base + k); the first success wins
            int kstar = 0;
            while (kstar < n_el && !(philox_u53(pk, a.spawn_stream, base + kstar) < waste_thr)) ++kstar;
            if (kstar < n_el) {
                uint64_t prev = 0;  // select the (kstar+1)-th smallest (key, index) among the eligible cells
                bool first = true;
                for (int it = 0; it <= kstar; ++it) {
                    uint64_t best = ~0ull;
                    for (int i = lane; i < n_waste; i += 32) {
                        if ((g[a.waste_cell[i]] & 0x7F) == CB(C_WASTE)) continue;
                        const uint64_t kx = static_cast<uint64_t>(keys[i]) << 32 | static_cast<uint32_t>(i);
                        if ((first || kx > prev) && kx < best) best = kx;
                    }
                    prev = warp_min_u64(best);
                    first = false;
                }
                if (lane == 0) {
                    const int idx = a.waste_cell[static_cast<uint32_t>(prev)];
                    g[idx] = CB(C_WASTE) | (g[idx] & kFlag);
                    ++cnt.waste;
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

// ====================================================================== the fused kernel
template <int KIND, bool TAPE, int VT>
__global__ void __launch_bounds__(kMaxThreads) ssd_step_kernel(const __grid_constant__ StepArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_color[kLutEntries];
    __shared__ int s_cta_stats[SSD_NUM_STATS];
    __shared__ int s_done;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int N = a.N, G = a.G, EPW = 32 / G;
    const int phases = a.phases;
    uint16_t* s_apple = reinterpret_cast<uint16_t*>(smem + a.L.apple);

    // ---- CTA-shared tables, then the only CTA barrier of the kernel
    for (int i = tid; i < kLutEntries; i += nthr) s_color[i] = a.color[i];
    if (tid < SSD_NUM_STATS) s_cta_stats[tid] = 0;
    if (tid == 0) s_done = 0;
    if (phases & SSD_PHASE_SPAWN)
        for (int i = tid; i < ((a.n_apple + 63) & ~63); i += nthr) s_apple[i] = i < a.n_apple ? a.apple_cell[i] : static_cast<uint16_t>(a.Ws + 1);
    __syncthreads();

    // ---- this warp's envs
    uint8_t* wbase = smem + a.L.warp0 + warp * a.L.warp_stride;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(wbase + a.L.w_mbar);
    uint8_t* tiles = wbase + a.L.w_tiles;
    EnvScratch* envs = reinterpret_cast<EnvScratch*>(wbase + a.L.w_env);
    const int tile_pitch = a.env_bytes + a.pad_bytes;
    const int we = a.env_begin + (blockIdx.x * nwarps + warp) * EPW;  // first local env of this warp
    const int nvalid = max(0, min(EPW, a.env_end - we));
    Counters cnt = {0, 0, 0, 0, 0, 0, 0};

    if (nvalid > 0) {
        // ---- load: one TMA bulk copy per env tile; zero frames while they are in flight
        if (lane == 0) { mbar_init(mbar, 1); mbar_expect_tx(mbar, static_cast<uint32_t>(EPW) * a.env_bytes); }
        __syncwarp();
        if (lane < EPW)
            bulk_g2s(tiles + a.pad_bytes + lane * tile_pitch, a.grid + static_cast<size_t>(we + lane) * a.env_bytes, a.env_bytes, mbar);
        {
            const uint4 z = make_uint4(0, 0, 0, 0);
            for (int q = 0; q <= EPW; ++q)
                for (int i = lane * 16; i < a.pad_bytes; i += 512) *reinterpret_cast<uint4*>(tiles + q * tile_pitch + i) = z;
        }
        mbar_wait(mbar, 0);  // tiles landed
        __syncwarp();

        // ---- phase A: one lane per agent, G lanes per env
        const int al = lane & (G - 1), gbase = lane & ~(G - 1), j = lane / G;  // j: env slot of this lane's group
        EnvScratch& S = envs[j];
        uint8_t* g = tiles + a.pad_bytes + j * tile_pitch;
        const int e = we + j;
        const bool active = j < nvalid && (a.mask == nullptr || a.mask[e] != 0);
        const bool valid = al < N;
        const size_t gi = static_cast<size_t>(e) * N + (valid ? al : 0);
        PhiloxKey pk;
        pk.k0 = a.key0; pk.k1 = a.key1; pk.t = a.t;
        pk.env = static_cast<uint32_t>(a.env_id0 + static_cast<uint64_t>(e));
        AgentLane me;
        me.key = 0x0101; me.ori = 0; me.act = -1; me.rew = 0;
        if (valid) {
            const uint32_t w = a.agents[gi];
            me.key = (w & 255) << 8 | ((w >> 8) & 255);
            me.ori = (w >> 16) & 3;
            if (active && a.actions) me.act = a.actions[gi];
            if (active && a.rew_accumulate && a.rew) me.rew = a.rew[gi];
            S.order[al] = (active && a.order) ? a.order[gi] : static_cast<uint8_t>(al);
            S.rew[al] = 0;
            S.firech[al] = 0;
        }
        if (al == 0) S.active = active;
        __syncwarp();
        if (phases & SSD_PHASE_MOVES) {
            moves_group<TAPE>(a, S, reinterpret_cast<MoveScratch*>(wbase + a.L.w_union)[j], g, me, valid && active, al, G, e, pk);
            cnt.steps += (active && al == 0);
        }
        if (valid) { S.pos[al] = static_cast<uint16_t>(me.key); S.ori[al] = static_cast<uint8_t>(me.ori); }
        __syncwarp();
        const int my_idx = tile_idx(a, me.key);
        if (phases & SSD_PHASE_CONSUME) {  // map_env.py:178-181, agent.py:177-183 / 216-222
            const uint8_t under = g[my_idx];
            const bool on_apple = valid && active && is_apple(under);
            // agents sharing a cell: the first one in agent order eats
            const uint32_t same = __match_any_sync(0xffffffffu, on_apple ? (me.key | static_cast<uint32_t>(gbase) << 16) : (0x80000000u | lane));
            __syncwarp();
            const bool ate = on_apple && (__ffs(same) - 1) == lane;
            if (ate) { g[my_idx] = CB(C_EMPTY) | (under & 3); me.rew += 1; ++cnt.eaten; }
            __syncwarp();
            if (KIND == SSD_KIND_HARVEST) recount_events(__ballot_sync(0xffffffffu, ate), static_cast<int>(g - tiles) + my_idx, tiles, a.Ws);
        }
        if ((phases & (SSD_PHASE_BEAMS | SSD_PHASE_SPAWN)) && valid && active) g[my_idx] |= kFlag;  // "an agent stands here"
        __syncwarp();
        if ((phases & SSD_PHASE_BEAMS) && KIND != SSD_KIND_PLAIN) {  // update_custom_moves map_env.py:545-552
            for (int k = 0; k < N; ++k) {  // action-dict order
                const int ag = S.order[k];
                const int act_k = __shfl_sync(0xffffffffu, me.act, ag, G);
                const uint32_t key_k = __shfl_sync(0xffffffffu, me.key, ag, G);
                const int ori_k = __shfl_sync(0xffffffffu, me.ori, ag, G);
                const bool fire = active && (act_k == 7 || (KIND == SSD_KIND_CLEANUP && act_k == 8));
                if (!__any_sync(0xffffffffu, fire)) continue;
                const bool clean = act_k == 8;
                int upd = -1, hits = 0, n = 0;
                if (fire && al < 3) n = ray_walk(a, S, g, key_k, ori_k, al, clean, upd, hits);
                if (fire && al == ag && !clean) { me.rew -= 1; ++cnt.fires; }  // fire_beam agent.py:170-172
                __syncwarp();
                if (fire && al < 3) {
                    S.raylen[k * 3 + al] = static_cast<uint8_t>(n);
                    if (al == 0) S.firech[k] = clean ? CB(C_CLEAN) : CB(C_FIRE);
                    if (upd >= 0) { g[upd] = CB(C_RIVER) | (g[upd] & kFlag); ++cnt.cleaned; }  // update_map :551-558, before the next agent fires
                    cnt.hits += hits;
                }
                __syncwarp();
            }
        }
        if (valid && active) {
            me.rew += S.rew[al];  // -50 per hit taken
            if (phases & (SSD_PHASE_MOVES | SSD_PHASE_CONSUME | SSD_PHASE_BEAMS)) {
                a.agents[gi] = (me.key >> 8) | (me.key & 255) << 8 | static_cast<uint32_t>(me.ori) << 16;
                if (a.rew) a.rew[gi] = me.rew;
            }
        }
        // beams recorded by an earlier phase call of this step (phase-split mode only)
        if (a.use_beam_buf) {
            const bool load = (phases & SSD_PHASE_RENDER) && !(phases & SSD_PHASE_BEAMS);
            const bool store = (phases & SSD_PHASE_BEAMS) && !(phases & SSD_PHASE_RENDER);
            for (int q = 0; q < EPW; ++q) {
                if (!envs[q].active) continue;
                for (int i = lane; i < 64; i += 32) {
                    uint8_t* p = i < 48 ? &envs[q].raylen[i] : &envs[q].firech[i - 48];
                    uint8_t* gp = a.beam_buf + static_cast<size_t>(we + q) * 64 + i;
                    if (load) *p = *gp;
                    if (store) *gp = *p;
                }
            }
        }
        __syncwarp();

        // ---- phase B: the whole warp per env
        if ((phases & SSD_PHASE_SPAWN) && KIND != SSD_KIND_PLAIN) {
            for (int q = 0; q < EPW; ++q) {
                if (!envs[q].active) continue;
                pk.env = static_cast<uint32_t>(a.env_id0 + static_cast<uint64_t>(we + q));
                void* scratch = wbase + a.L.w_union;
                if (KIND == SSD_KIND_HARVEST)
                    harvest_spawn<TAPE>(a, tiles + a.pad_bytes + q * tile_pitch, s_apple, static_cast<uint32_t*>(scratch), we + q, pk, lane, cnt);
                else
                    cleanup_spawn<TAPE>(a, tiles + a.pad_bytes + q * tile_pitch, s_apple, static_cast<uint32_t*>(scratch), we + q, pk, lane, cnt);
                __syncwarp();
            }
        }

        // ---- store: grid rows back to HBM (kFlag stripped on the way out)
        if (phases & (SSD_PHASE_CONSUME | SSD_PHASE_BEAMS | SSD_PHASE_SPAWN)) {
            for (int q = 0; q < EPW; ++q) {
                if (!envs[q].active) continue;
                uint4* gdst = reinterpret_cast<uint4*>(a.grid + static_cast<size_t>(we + q) * a.env_bytes);
                const uint4* gsrc = reinterpret_cast<const uint4*>(tiles + a.pad_bytes + q * tile_pitch);
                for (int i = lane; i < a.env_bytes / 16; i += 32) {
                    uint4 v = gsrc[i];
                    v.x &= 0x7F7F7F7Fu; v.y &= 0x7F7F7F7Fu; v.z &= 0x7F7F7F7Fu; v.w &= 0x7F7F7F7Fu;
                    gdst[i] = v;
                }
            }
        }

        // ---- phase C: overlay + render + coalesced stores
        if ((phases & SSD_PHASE_RENDER) && a.obs != nullptr) {
            __syncwarp();  // the write-back above has read the tiles
            // get_map_with_agents map_env.py:280-302: agents in agent order (the last one on a cell wins),
            // then beams in firing order (a later beam overwrites an earlier one)
            {
                const uint32_t key = valid ? S.pos[al] : 0x0101u;
                const uint32_t same = __match_any_sync(0xffffffffu, valid ? (key | static_cast<uint32_t>(gbase) << 16) : (0x80000000u | lane));
                if (valid && (31 - __clz(same)) == lane) g[tile_idx(a, key)] = agent_cell(al);
                __syncwarp();
                if (KIND != SSD_KIND_PLAIN) {
                    for (int k = 0; k < N; ++k) {
                        const uint32_t ch = S.firech[k];
                        if (!__any_sync(0xffffffffu, ch != 0)) continue;
                        if (ch != 0 && al < 3) {
                            const int ag = S.order[k];
                            const int ori = S.ori[ag];
                            const int d0 = (ori == 1) - (ori == 3), d1 = (ori == 2) - (ori == 0);
                            int r = static_cast<int>(S.pos[ag] >> 8) + d0, c = static_cast<int>(S.pos[ag] & 255) + d1;
                            if (al == 1) { r += -d1 - d0; c += d0 - d1; }
                            if (al == 2) { r -= -d1 + d0; c -= d0 + d1; }
                            const int n = S.raylen[k * 3 + al], dp = d0 * a.Ws + d1;
                            int p = r * a.Ws + c;
                            for (int i = 0; i < n; ++i) { g[p] = static_cast<uint8_t>(ch); p += dp; }
                        }
                        __syncwarp();
                    }
                }
            }
            uint2* s_view = reinterpret_cast<uint2*>(wbase + a.L.w_union);
            for (int i = lane; i < EPW * N; i += 32) s_view[i] = view_param(a, envs[i / N], a.pad_bytes + (i / N) * tile_pitch, i % N);
            __syncwarp();
            uint8_t* dst = a.obs + static_cast<size_t>(we) * a.obs_env;
            const bool all_active = (nvalid == EPW) && (a.mask == nullptr);
            if constexpr (VT > 0) {
                if (all_active) render_rows<VT>(s_view, tiles, s_color, reinterpret_cast<uint32_t*>(wbase + a.L.w_union + a.L.u_stage), dst, EPW * N * VT);
                else render_generic(a, envs, s_view, tiles, s_color, dst, nvalid);
            } else {
                render_generic(a, envs, s_view, tiles, s_color, dst, nvalid);
            }
        }
    }

    // ---- stats: warp -> CTA -> one set of global atomics per CTA (issued by the last warp to finish)
    if (a.stats != nullptr) {
        const int v[7] = {cnt.steps, cnt.eaten, cnt.fires, cnt.hits, cnt.cleaned, cnt.apples, cnt.waste};
        const int slot[7] = {0, 2, 3, 4, 5, 6, 7};
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const int tot = __reduce_add_sync(0xffffffffu, v[i]);
            if (lane == 0 && tot) atomicAdd(&s_cta_stats[slot[i]], tot);
        }
        __syncwarp();
        int last = 0;
        if (lane == 0) { __threadfence_block(); last = (atomicAdd(&s_done, 1) == nwarps - 1); }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last && lane < SSD_NUM_STATS) {
            const int tot = *reinterpret_cast<volatile int*>(&s_cta_stats[lane]);
            if (tot) atomicAdd(&a.stats[lane], static_cast<unsigned long long>(tot));
        }
    }
}

// ====================================================================== the fast path of a full step
// Same algorithm as ssd_step_kernel, specialised for what a production step is: all phases, every env
// stepped (no mask), actions in agent order, N <= 8 (8 lanes per env, 4 envs per warp), a packed
// row renderer for this view size and only whole warps (the launcher sends any tail envs through the
// general kernel).  What the specialisation buys: no per-env / per-phase flag tests, fire flags from
// one ballot, view geometry straight from the agent registers, and observation rows that leave
// shared memory as TMA bulk stores instead of LDS.128 / STG.128 pairs.

// Packed rows with TMA copy-out.  As render_rows, but every chunk of 32 rows (32 * 3V bytes, a multiple
// of 16) is written by ONE cp.async.bulk from the staging buffer.  The warp's slab starts at `dst`
// = 16-byte aligned base + mis; the buffer holds the aligned image [base + c*CH, base + (c+1)*CH) of
// chunk c: the mis/4 words that spill over the end of a chunk are kept in registers by lane 31
// (they are its last words) and stored at the head of the next chunk's image.  Only the first
// 16 - mis and the last mis bytes of the slab are written with plain 4-byte stores.
__host__ __device__ constexpr int row_words(int o, int RB) { return ((o + RB - 1) >> 2) - ((o + 3) >> 2) + 1; }
__host__ __device__ constexpr int row_words_min(int RB) {
    int m = row_words(0, RB);
    for (int r = 1; r < 4; ++r) m = row_words(r, RB) < m ? row_words(r, RB) : m;
    return m;
}

template <int VT>
__device__ __forceinline__ void render_rows_tma(const uint2* s_view, const uint8_t* tiles, const uint32_t* s_color,
                                                uint32_t* stage, uint8_t* dst, int total_rows, int debug) {
    constexpr int RB = 3 * VT;            // bytes per view row
    constexpr int CH = 32 * RB;           // bytes per chunk
    constexpr int NP = (RB + 3 + 3) / 4;  // words covering the row plus the next row's first pixel
    constexpr int O31 = 31 * RB, M31 = ((O31 + RB - 1) >> 2) - ((O31 + 3) >> 2) + 1;  // lane 31 owns the chunk's last words
    constexpr int MMIN = row_words_min(RB);  // every lane owns at least this many words of a chunk
    static_assert(M31 >= 3 && M31 <= NP, "carry words must all live in lane 31");
    const int lane = threadIdx.x & 31;
    const int mis = static_cast<int>(reinterpret_cast<uintptr_t>(dst) & 15);  // multiple of 4
    const int mis4 = mis >> 2;
    const uint32_t o = static_cast<uint32_t>(lane) * RB;
    const uint32_t d8 = 8u * ((4u - (o & 3u)) & 3u);
    const uint32_t w0 = (o + 3) >> 2, w1 = (o + RB - 1) >> 2;
    const bool extra = static_cast<int>(w1 - w0) + 1 > MMIN;  // this lane owns MMIN + 1 words of every chunk
    uint32_t* st = stage + mis4 + w0;
    const int n_chunks = (total_rows + 31) >> 5;  // >= 2: a warp renders at least 4 * VT rows
    const int end_last = mis + (total_rows - (n_chunks - 1) * 32) * RB;  // valid image bytes of the last chunk (multiple of 4)
    const int hi_last = min(CH, end_last & ~15);
    uint8_t* const img0 = dst - mis;  // 16-byte aligned image of chunk 0
    const uint32_t stage32 = smem_u32(stage);
    // where lane 31 parks the words that spill over a chunk: the head of the next image, or a dummy slot
    uint32_t* const k1 = stage + (mis4 >= 1 ? mis4 - 1 : CH / 4 + 5);
    uint32_t* const k2 = stage + (mis4 >= 2 ? mis4 - 2 : CH / 4 + 6);
    uint32_t* const k3 = stage + (mis4 >= 3 ? mis4 - 3 : CH / 4 + 7);
    const int lo0 = mis != 0 ? 16 : 0;  // the first 16 - mis bytes of the slab leave with plain stores
    uint8_t* gp = img0 + lo0;           // destination, source and size of the next bulk store
    uint32_t sp = stage32 + lo0, nb = CH - lo0;
    uint32_t c0 = 0, c1 = 0, c2 = 0;
#pragma unroll 1
    for (int c = 0; c < n_chunks; ++c) {
        const int R = min(c * 32 + lane, total_rows - 1);  // lanes past the end redo the last row; their words are never copied out
        const int ga = R / VT, i = R - ga * VT;            // rows are ordered (env, agent, i)
        uint32_t X[VT + 2];
        {
            const uint2 vp = s_view[ga];
            const int si = static_cast<int16_t>(vp.y & 0xffffu), sj = static_cast<int32_t>(vp.y) >> 16;
            const uint8_t* g = tiles + static_cast<int32_t>(vp.x) + i * si;
#pragma unroll
            for (int j = 0; j < VT; ++j) X[j] = cell_color(s_color, g[j * sj]);
        }
        X[VT] = __shfl_down_sync(0xffffffffu, X[0], 1);
        X[VT + 1] = 0;
        uint32_t P[NP + 1];
#pragma unroll
        for (int w = 0; w < NP; ++w) {
            const int p = (4 * w) / 3, ph = (4 * w) % 3;
            P[w] = __byte_perm(X[p], X[p + 1], ph == 0 ? 0x4210u : (ph == 1 ? 0x5421u : 0x6542u));
        }
        P[NP] = 0;
        uint32_t Q[MMIN + 1];
#pragma unroll
        for (int m = 0; m <= MMIN; ++m) Q[m] = __funnelshift_r(P[m], P[m + 1], d8);
        bulk_wait_read();  // the previous chunk's bulk store (issued by lane 0) must have finished READING the buffer
        __syncwarp();
#pragma unroll
        for (int m = 0; m < MMIN; ++m) st[m] = Q[m];
        if (extra) st[MMIN] = Q[MMIN];
        if (lane == 31) { *k1 = c2; *k2 = c1; *k3 = c0; }  // head of this image = spill of the previous chunk (unused for c == 0)
        c0 = Q[M31 - 3]; c1 = Q[M31 - 2]; c2 = Q[M31 - 1];
        fence_async_smem();
        __syncwarp();
        if (c == n_chunks - 1) nb = hi_last;  // n_chunks >= 2: the last chunk starts at the head of the buffer
        if (lane == 0 && !SSD_SKIP(debug, 1)) {
            bulk_s2g_u32(gp, sp, nb);
            bulk_commit();
        }
        if (c == 0 && lane >= mis4 && lane < 4 && mis != 0)  // first bytes of the slab
            *reinterpret_cast<uint32_t*>(img0 + 4 * lane) = stage[lane];
        gp += nb; sp = stage32; nb = CH;
    }
    {   // last bytes of the slab
        const int off = hi_last + 4 * lane;
        if (off < end_last) *reinterpret_cast<uint32_t*>(img0 + static_cast<size_t>(n_chunks - 1) * CH + off) = stage[off >> 2];
    }
    bulk_wait_read();  // shared memory must outlive the last bulk read
}

template <int KIND, bool TAPE, int VT, int G>
__global__ void __launch_bounds__(kMaxThreads) ssd_step_fast_kernel(const __grid_constant__ StepArgs a) {
    constexpr int EPW = 32 / G;                                       // envs per warp: 4 (N <= 8) or 2 (N <= 16)
    constexpr uint32_t kSlotLsb = G == 8 ? 0x01010101u : 0x00010001u;  // bit 0 of every env's lane group
    using FastScratch = FastScratchT<G>;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_color[kLutEntries];
    __shared__ int s_cta_stats[SSD_NUM_STATS];
    __shared__ int s_done;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int N = a.N;
    uint16_t* s_apple = reinterpret_cast<uint16_t*>(smem + a.Lf.apple);
    pdl_launch_dependents();  // a chained next step may start launching (it waits per task, see below)

    for (int i = tid; i < kLutEntries; i += nthr) s_color[i] = a.color[i];
    if (tid < SSD_NUM_STATS) s_cta_stats[tid] = 0;
    if (tid == 0) s_done = 0;
    if (KIND != SSD_KIND_PLAIN)
#pragma unroll 1
        for (int i = tid; i < ((a.n_apple + 63) & ~63); i += nthr) s_apple[i] = i < a.n_apple ? a.apple_cell[i] : static_cast<uint16_t>(a.Ws + 1);
    __syncthreads();

    uint8_t* wbase = smem + a.Lf.warp0 + warp * a.Lf.warp_stride;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(wbase + a.Lf.w_mbar);
    uint8_t* tiles = wbase + a.Lf.w_tiles;
    FastScratch* envs = reinterpret_cast<FastScratch*>(wbase + a.Lf.w_env);
    const int tile_pitch = a.env_bytes + a.pad_bytes;
    const int we = a.env_begin + (blockIdx.x * nwarps + warp) * EPW;  // the launcher only sends whole warps
    Counters cnt = {0, 0, 0, 0, 0, 0, 0};

    if (we < a.env_end) {
        // ---- load: one TMA bulk copy per env tile; zero the frames while they are in flight
        if (lane == 0) {
            mbar_init(mbar, 1);
            mbar_expect_tx(mbar, static_cast<uint32_t>(EPW) * a.env_bytes);
            if (a.dep_wait) {  // chained step: the previous step's kernel may still be running; wait for OUR four envs only
                while (ld_acquire_u32(a.done + we / EPW) != a.epoch - 1) __nanosleep(64);
                fence_async_all();  // its ordinary stores -> our TMA loads
            }
        }
        __syncwarp();
        if (lane < EPW)
            bulk_g2s(tiles + a.pad_bytes + lane * tile_pitch, a.grid + static_cast<size_t>(we + lane) * a.env_bytes, a.env_bytes, mbar);
        {
            const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int q = 0; q <= EPW; ++q)
#pragma unroll 1
                for (int i = lane * 16; i < a.pad_bytes; i += 512) *reinterpret_cast<uint4*>(tiles + q * tile_pitch + i) = z;
        }
        // agent words and actions travel while the tiles do
        const int al = lane & (G - 1), gbase = lane & ~(G - 1), j = lane / G;
        FastScratch& S = envs[j];
        uint8_t* g = tiles + a.pad_bytes + j * tile_pitch;
        const int e = we + j;
        const bool valid = al < N;
        const size_t gi = static_cast<size_t>(e) * N + (valid ? al : 0);
        PhiloxKey pk;
        pk.k0 = a.key0; pk.k1 = a.key1; pk.t = a.t;
        pk.env = static_cast<uint32_t>(a.env_id0 + static_cast<uint64_t>(e));
        AgentLane me;
        me.key = 0x0101; me.ori = 0; me.act = -1; me.rew = 0;
        if (valid) {
            const uint32_t w = __ldcg(a.agents + gi);  // L2: a chained predecessor may just have written it
            me.act = a.actions[gi];
            me.key = (w & 255) << 8 | ((w >> 8) & 255);
            me.ori = (w >> 16) & 3;
            S.order[al] = static_cast<uint8_t>(al);
            S.rew[al] = 0;
        }
        mbar_wait(mbar, 0);  // tiles landed
        __syncwarp();

        // ---- phase A: one lane per agent
        moves_group<TAPE>(a, S, reinterpret_cast<MoveScratch*>(wbase + a.Lf.w_union)[j], g, me, valid, al, G, e, pk);
        cnt.steps += (al == 0);
        if (valid) S.pos[al] = static_cast<uint16_t>(me.key);
        const int my_idx = tile_idx(a, me.key);
        {   // consume, map_env.py:178-181: of agents sharing a cell (appendix A.2 quirk) the first in agent order eats
            const uint8_t under = g[my_idx];
            const bool on_apple = valid && is_apple(under);
            const uint32_t same = __match_any_sync(0xffffffffu, on_apple ? (me.key | static_cast<uint32_t>(gbase) << 16) : (0x80000000u | lane));
            const bool ate = on_apple && (__ffs(same) - 1) == lane;
            if (ate) { g[my_idx] = CB(C_EMPTY) | (under & 3); me.rew += 1; ++cnt.eaten; }
            __syncwarp();
            if (KIND == SSD_KIND_HARVEST) recount_events(__ballot_sync(0xffffffffu, ate), static_cast<int>(g - tiles) + my_idx, tiles, a.Ws);
        }
        __syncwarp();
        if (KIND != SSD_KIND_PLAIN && valid) g[my_idx] |= kFlag;  // "an agent stands here"
        __syncwarp();
        uint32_t fmask = 0;  // bit (8 * env slot + agent): that agent fires
        uint32_t* const fire_list = reinterpret_cast<uint32_t*>(wbase + a.Lf.w_union);
        const int ray_f = lane / 3, ray_s = lane - 3 * ray_f;  // Harvest: ray lane -> (firing agent of the round, ray)
        uint32_t fire_ent = 0;
        if (KIND == SSD_KIND_HARVEST && !SSD_SKIP(a.debug, 8)) {
            // update_custom_moves map_env.py:545-552.  Harvest has only 'F' beams: they change no cell, so the
            // firing order is irrelevant and the rays of ALL firing agents of the warp walk at once, 3 lanes each.
            const bool fire_me = me.act == 7;
            fmask = __ballot_sync(0xffffffffu, fire_me);
            if (fmask) {
                fire_ent = me.key | static_cast<uint32_t>(me.ori) << 16 | static_cast<uint32_t>(j) << 18 | static_cast<uint32_t>(al) << 21;
                if (fire_me) { fire_list[__popc(fmask & lanemask_lt())] = fire_ent; me.rew -= 1; ++cnt.fires; }  // fire_beam agent.py:170-172
                __syncwarp();
                const int nf = __popc(fmask);
#pragma unroll 1
                for (int f0 = 0; f0 < nf; f0 += 10) {
                    if (lane < 30 && f0 + ray_f < nf) {
                        const uint32_t en = fire_list[f0 + ray_f];
                        const int slot = (en >> 18) & 3, ag = en >> 21;
                        int upd = -1, hits = 0;
                        const int n = ray_walk<FastScratch, true>(a, envs[slot], tiles + a.pad_bytes + slot * tile_pitch, en & 0xffffu,
                                                                  (en >> 16) & 3, ray_s, false, upd, hits);
                        envs[slot].raylen[ag * 3 + ray_s] = static_cast<uint8_t>(n);
                        cnt.hits += hits;
                    }
                }
                __syncwarp();
            }
        }
        if (KIND == SSD_KIND_CLEANUP && !SSD_SKIP(a.debug, 8)) {  // firing order matters: a CLEAN beam turns 'H' into 'R' for the next one
            fmask = __ballot_sync(0xffffffffu, me.act == 7 || me.act == 8);
            for (int k = 0; k < N; ++k) {
                if (!((fmask >> k) & kSlotLsb)) continue;  // nobody in this warp fires in slot k
                const bool fire = (fmask >> (gbase + k)) & 1u;
                const int act_k = __shfl_sync(0xffffffffu, me.act, k, G);
                const uint32_t key_k = __shfl_sync(0xffffffffu, me.key, k, G);
                const int ori_k = __shfl_sync(0xffffffffu, me.ori, k, G);
                const bool clean = act_k == 8;
                int upd = -1, hits = 0, n = 0;
                if (fire && al < 3) n = ray_walk(a, S, g, key_k, ori_k, al, clean, upd, hits);
                if (fire && al == k && !clean) { me.rew -= 1; ++cnt.fires; }  // fire_beam agent.py:170-172
                __syncwarp();
                if (fire && al < 3) {
                    S.raylen[k * 3 + al] = static_cast<uint8_t>(n);
                    if (upd >= 0) { g[upd] = CB(C_RIVER) | (g[upd] & kFlag); ++cnt.cleaned; }  // update_map :551-558, before the next agent fires
                    cnt.hits += hits;
                }
                __syncwarp();
            }
        }
        if (valid) {
            me.rew += S.rew[al];  // -50 per hit taken
            a.agents[gi] = (me.key >> 8) | (me.key & 255) << 8 | static_cast<uint32_t>(me.ori) << 16;
            a.rew[gi] = me.rew;
        }

        // ---- phase B: the whole warp per env
        if (KIND != SSD_KIND_PLAIN && !SSD_SKIP(a.debug, 4)) {
            void* scratch = wbase + a.Lf.w_union;
            if (KIND == SSD_KIND_HARVEST) {
                harvest_spawn_warp<TAPE, EPW>(a, tiles, tile_pitch, s_apple, static_cast<uint32_t*>(scratch), a.Lf.u_words, we, pk, lane, cnt);
            } else {
#pragma unroll 1
                for (int q = 0; q < EPW; ++q) {
                    pk.env = static_cast<uint32_t>(a.env_id0 + static_cast<uint64_t>(we + q));
                    cleanup_spawn<TAPE>(a, tiles + a.pad_bytes + q * tile_pitch, s_apple, static_cast<uint32_t*>(scratch), we + q, pk, lane, cnt);
                    __syncwarp();
                }
            }
            if (valid) g[my_idx] &= 0x7F;  // agents sharing a cell all write the same byte
            __syncwarp();
        }

        // ---- store: grid tiles back to HBM
        if (!SSD_SKIP(a.debug, 64)) {
            const int n16 = a.env_bytes >> 4;
#pragma unroll
            for (int q = 0; q < EPW; ++q) {
                uint4* gdst = reinterpret_cast<uint4*>(a.grid + static_cast<size_t>(we + q) * a.env_bytes);
                const uint4* gsrc = reinterpret_cast<const uint4*>(tiles + a.pad_bytes + q * tile_pitch);
#pragma unroll 1
                for (int i = lane; i < n16; i += 32) gdst[i] = gsrc[i];
            }
        }
        __syncwarp();  // the write-back above has read the tiles

        // ---- phase C: overlay (get_map_with_agents map_env.py:280-302), view geometry, packed rows
        {
            const uint32_t same = __match_any_sync(0xffffffffu, valid ? (me.key | static_cast<uint32_t>(gbase) << 16) : (0x80000000u | lane));
            if (valid && (31 - __clz(same)) == lane) g[my_idx] = agent_cell(al);  // the last agent on a cell wins
            __syncwarp();
            if (KIND == SSD_KIND_HARVEST && fmask) {  // all beams are 'F': the painting order is irrelevant
                if ((fmask >> lane) & 1u) fire_list[__popc(fmask & lanemask_lt())] = fire_ent;  // the union was reused by the spawn pass
                __syncwarp();
                const int nf = __popc(fmask);
#pragma unroll 1
                for (int f0 = 0; f0 < nf; f0 += 10) {
                    if (lane < 30 && f0 + ray_f < nf) {
                        const uint32_t en = fire_list[f0 + ray_f];
                        const int slot = (en >> 18) & 3, ag = en >> 21, ori = (en >> 16) & 3;
                        const int d0 = (ori == 1) - (ori == 3), d1 = (ori == 2) - (ori == 0);
                        int r = static_cast<int>((en >> 8) & 255) + d0, c = static_cast<int>(en & 255) + d1;
                        if (ray_s == 1) { r += -d1 - d0; c += d0 - d1; }
                        if (ray_s == 2) { r -= -d1 + d0; c -= d0 + d1; }
                        const int n = envs[slot].raylen[ag * 3 + ray_s], dp = d0 * a.Ws + d1;
                        uint8_t* p = tiles + a.pad_bytes + slot * tile_pitch + r * a.Ws + c;
#pragma unroll 1
                        for (int i = 0; i < n; ++i) { *p = CB(C_FIRE); p += dp; }
                    }
                }
                __syncwarp();
            }
            if (KIND == SSD_KIND_CLEANUP) {  // beams in firing order: a later beam overwrites an earlier one
                for (int k = 0; k < N; ++k) {
                    if (!((fmask >> k) & kSlotLsb)) continue;
                    const uint32_t key_k = __shfl_sync(0xffffffffu, me.key, k, G);
                    const int ori_k = __shfl_sync(0xffffffffu, me.ori, k, G);
                    const int act_k = __shfl_sync(0xffffffffu, me.act, k, G);
                    if (((fmask >> (gbase + k)) & 1u) && al < 3) {
                        const int d0 = (ori_k == 1) - (ori_k == 3), d1 = (ori_k == 2) - (ori_k == 0);
                        int r = static_cast<int>(key_k >> 8) + d0, c = static_cast<int>(key_k & 255) + d1;
                        if (al == 1) { r += -d1 - d0; c += d0 - d1; }
                        if (al == 2) { r -= -d1 + d0; c -= d0 + d1; }
                        const int n = S.raylen[k * 3 + al], dp = d0 * a.Ws + d1;
                        const uint8_t ch = act_k == 8 ? CB(C_CLEAN) : CB(C_FIRE);
                        int p = r * a.Ws + c;
#pragma unroll 1
                        for (int i = 0; i < n; ++i) { g[p] = ch; p += dp; }
                    }
                    __syncwarp();
                }
            }
            uint2* s_view = reinterpret_cast<uint2*>(wbase + a.Lf.w_union);
            if (valid) {  // rot90 folded into strides, see view_param
                const int pr = me.key >> 8, pc = me.key & 255, r = a.r, Ws = a.Ws;
                const int k = (4 - me.ori) & 3;
                int a0, si, sj;
                if (k == 0)      { a0 = (pr - r) * Ws + pc - r; si = Ws;  sj = 1; }
                else if (k == 2) { a0 = (pr + r) * Ws + pc + r; si = -Ws; sj = -1; }
                else if (k == 1) { a0 = (pr - r) * Ws + pc + r; si = -1;  sj = Ws; }
                else             { a0 = (pr + r) * Ws + pc - r; si = 1;   sj = -Ws; }
                s_view[j * N + al] = make_uint2(static_cast<uint32_t>(a0 + a.pad_bytes + j * tile_pitch),
                                                (static_cast<uint32_t>(si) & 0xffffu) | static_cast<uint32_t>(sj) << 16);
            }
            __syncwarp();
            if (!SSD_SKIP(a.debug, 2))
            render_rows_tma<VT>(s_view, tiles, s_color, reinterpret_cast<uint32_t*>(wbase + a.Lf.w_union + a.Lf.u_stage),
                                a.obs + static_cast<size_t>(we) * a.obs_env, EPW * N * VT, a.debug);
        }
        if (a.publish) {  // everything this task wrote (grid, agent words, rewards, observation rows) is visible before the word is
            __syncwarp();
            if (lane == 0) {
                bulk_wait_all();
                __threadfence();
                st_release_u32(a.done + we / EPW, a.epoch);
            }
        }
    }

    // ---- stats: warp -> CTA -> one set of global atomics per CTA (issued by the last warp to finish)
    if (a.stats != nullptr && !SSD_SKIP(a.debug, 32)) {
        // per-warp totals are small (<= 32 agents): two packed reductions carry all seven counters
        //   r0: steps | eaten << 8 | fires << 16 | hits << 24 (hits <= 3 per shooter)     r1: cleaned | waste << 8 | apples << 12
        const uint32_t r0 = __reduce_add_sync(0xffffffffu, static_cast<uint32_t>(cnt.steps | cnt.eaten << 8 | cnt.fires << 16 | cnt.hits << 24));
        const uint32_t r1 = __reduce_add_sync(0xffffffffu, static_cast<uint32_t>(cnt.cleaned | cnt.waste << 8 | cnt.apples << 12));
        if (lane < 7) {
            const uint32_t v = lane < 4 ? (r0 >> (8 * lane)) & 255u : (lane == 4 ? r1 & 255u : (lane == 5 ? r1 >> 12 : (r1 >> 8) & 15u));
            if (v) atomicAdd(&s_cta_stats[lane == 0 ? 0 : lane + 1], static_cast<int>(v));
        }
        __syncwarp();
        int last = 0;
        if (lane == 0) { __threadfence_block(); last = (atomicAdd(&s_done, 1) == nwarps - 1); }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last && lane < SSD_NUM_STATS) {
            const int tot = *reinterpret_cast<volatile int*>(&s_cta_stats[lane]);
            if (tot) atomicAdd(&a.stats[lane], static_cast<unsigned long long>(tot));
        }
    }
}

// ====================================================================== reset: setup_agents + reset_map
// map_env.py:214-229: spawn_point (:651-662) with the shuffle replaced by a (key, index) order --
// the reference takes the LAST free entry of the shuffled list = the free entry with the largest
// (key, index); spawn_rotation (:664-667) indexes ['LEFT','RIGHT','UP','DOWN'] with randint(4).
__global__ void __launch_bounds__(128) ssd_reset_kernel(const ResetArgs a) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.env_end) return;
    if (a.mask != nullptr && a.mask[e] == 0) return;
    const uint32_t env = static_cast<uint32_t>(a.env_id0 + static_cast<uint64_t>(e));
    uint16_t taken[kMaxAgents];
    for (int ag = 0; ag < a.N; ++ag) {
        uint64_t best = 0;
        int best_s = -1;
        uint4 blk = make_uint4(0, 0, 0, 0);
        int blk_id = -1;
        for (int s = 0; s < a.n_spawn; ++s) {
            const uint32_t wi = ag * a.n_spawn + s;
            if (static_cast<int>(wi >> 2) != blk_id) { blk_id = wi >> 2; blk = philox4x32_10(env, a.t, STREAM_RPOINT, blk_id, a.key0, a.key1); }
            const uint16_t key = a.spawn_key[s];
            bool free_cell = true;
            for (int q = 0; q < ag; ++q) free_cell &= (taken[q] != key);
            const uint64_t kx = static_cast<uint64_t>(pick_word(blk, wi)) << 32 | static_cast<uint32_t>(s);
            if (free_cell && (best_s < 0 || kx > best)) { best = kx; best_s = s; }
        }
        const uint16_t key = a.spawn_key[best_s < 0 ? 0 : best_s];
        taken[ag] = key;
        const uint32_t rot = philox_word(PhiloxKey{a.key0, a.key1, env, a.t}, STREAM_RROT, ag) & 3;
        const uint32_t ori = (rot == 0) ? 3u : (rot == 1) ? 1u : (rot == 2) ? 0u : 2u;  // LEFT, RIGHT, UP, DOWN
        a.agents[static_cast<size_t>(e) * a.N + ag] = (key >> 8) | (key & 255) << 8 | ori << 16;
    }
    // reset_map + build_walls + custom_reset (map_env.py:560-564, harvest.py:57-60, cleanup.py:84-92)
    const uint4* src = reinterpret_cast<const uint4*>(a.init_grid);
    uint4* dst = reinterpret_cast<uint4*>(a.grid + static_cast<size_t>(e) * a.env_bytes);
    for (int i = 0; i < a.env_bytes / 16; ++i) dst[i] = src[i];
}

// ====================================================================== state pack / unpack, selftest
__device__ __forceinline__ uint8_t dev_ascii_to_cell(uint8_t ch) {
    switch (ch) {
        case '0': return CB(C_PAD);
        case ' ': return CB(C_EMPTY);
        case '@': return CB(C_WALL);
        case 'A': return CB(C_APPLE);
        case 'H': return CB(C_WASTE);
        case 'R': return CB(C_RIVER);
        case 'S': return CB(C_STREAM);
        case 'F': return CB(C_FIRE);
        case 'C': return CB(C_CLEAN);
        default: return (ch >= '1' && ch <= '9') ? CB(static_cast<uint8_t>(C_AGENT + ch - '1')) : CB(C_OTHER);
    }
}
__device__ __forceinline__ uint8_t dev_cell_to_ascii(uint8_t cell) {
    const uint8_t code = (cell & 0x7F) >> 2;
    switch (code) {
        case C_PAD: return '0';
        case C_EMPTY: return ' ';
        case C_WALL: return '@';
        case C_APPLE: return 'A';
        case C_WASTE: return 'H';
        case C_RIVER: return 'R';
        case C_STREAM: return 'S';
        case C_FIRE: return 'F';
        case C_CLEAN: return 'C';
        default: return (code >= C_AGENT && code < C_AGENT + 9) ? static_cast<uint8_t>('1' + code - C_AGENT) : static_cast<uint8_t>('?');
    }
}
__global__ void pack_state_kernel(int kind, int B, int N, int H, int W, int Ws, int env_bytes, const uint8_t* grid_in, const int16_t* pos_in,
                                  const uint8_t* ori_in, uint8_t* grid, uint32_t* agents) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < static_cast<size_t>(B) * env_bytes) {
        const size_t b = i / env_bytes, q = i % env_bytes;
        const int r = static_cast<int>(q / Ws), c = static_cast<int>(q % Ws);
        uint8_t cell = 0;
        if (r < H && c < W) {
            const uint8_t* gi = grid_in + b * H * W;
            const uint8_t ch = gi[r * W + c];
            cell = dev_ascii_to_cell(ch);
            if (kind == SSD_KIND_HARVEST && (ch == ' ' || ch == 'A')) {  // cached neighbourhood count (ssd_internal.h)
                int n = 0;
                for (int dr = -1; dr <= 1; ++dr)
                    for (int dc = -1; dc <= 1; ++dc) {
                        const int rr = r + dr, cc = c + dc;
                        n += (dr || dc) && rr >= 0 && rr < H && cc >= 0 && cc < W && gi[rr * W + cc] == 'A';
                    }
                cell |= static_cast<uint8_t>(n < 3 ? n : 3);
            }
        }
        grid[i] = cell;
    }
    if (i < static_cast<size_t>(B) * N)
        agents[i] = (pos_in[2 * i] & 255) | (pos_in[2 * i + 1] & 255) << 8 | static_cast<uint32_t>(ori_in[i] & 3) << 16;
}
__global__ void unpack_state_kernel(int B, int N, int H, int W, int Ws, int env_bytes, const uint8_t* grid, const uint32_t* agents,
                                    uint8_t* grid_out, int16_t* pos_out, uint8_t* ori_out) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t HW = static_cast<size_t>(H) * W;
    if (grid_out != nullptr && i < static_cast<size_t>(B) * HW) {
        const size_t b = i / HW, q = i % HW, r = q / W, c = q % W;
        grid_out[i] = dev_cell_to_ascii(grid[b * env_bytes + r * Ws + c]);
    }
    if (i < static_cast<size_t>(B) * N) {
        const uint32_t w = agents[i];
        if (pos_out != nullptr) { pos_out[2 * i] = w & 255; pos_out[2 * i + 1] = (w >> 8) & 255; }
        if (ori_out != nullptr) ori_out[i] = (w >> 16) & 3;
    }
}
// Full-map frames: map_to_colors(get_map_with_agents()) (map_env.py:280-339) for every env, uint8 [B][H][W][3].
// One thread per output pixel triple; the agents of the env are painted in agent order (the last one on a cell wins).
// Beams are not part of the persistent state (map_env.py:169 clears them every step) and are not drawn.
__global__ void render_map_kernel(int B, int N, int H, int W, int Ws, int env_bytes, const uint8_t* grid, const uint32_t* agents,
                                  const uint32_t* color, uint8_t* out) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t HW = static_cast<size_t>(H) * W;
    if (i >= static_cast<size_t>(B) * HW) return;
    const size_t b = i / HW, q = i % HW;
    const uint32_t r = static_cast<uint32_t>(q / W), c = static_cast<uint32_t>(q % W);
    uint8_t cell = grid[b * env_bytes + r * Ws + c] & 0x7F;
    for (int ag = 0; ag < N; ++ag) {
        const uint32_t w = agents[b * N + ag];
        if ((w & 255u) == r && ((w >> 8) & 255u) == c) cell = agent_cell(ag);
    }
    const uint32_t rgb = color[cell];
    out[3 * i] = rgb & 255; out[3 * i + 1] = (rgb >> 8) & 255; out[3 * i + 2] = (rgb >> 16) & 255;
}

__global__ void philox_selftest_kernel(const uint32_t* ck, uint32_t* out) {
    const uint4 v = philox4x32_10(ck[0], ck[1], ck[2], ck[3], ck[4], ck[5]);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}

// ====================================================================== launchers
constexpr int kMaxDevices = 64;
template <int KIND, bool TAPE>
static cudaError_t launch_v(const StepArgs& a, int threads, cudaStream_t stream, bool fast_rows) {
    const int envs_per_cta = (threads / 32) * (32 / a.G);
    const int ctas = (a.env_end - a.env_begin + envs_per_cta - 1) / envs_per_cta;
    if (ctas <= 0) return cudaSuccess;
    const int vt = fast_rows ? a.V : 0;
#define SSD_LAUNCH(VT_)                                                                                         \
    do {                                                                                                        \
        auto kern = ssd_step_kernel<KIND, TAPE, VT_>;                                                           \
        static uint32_t smem_set[kMaxDevices] = {};  /* the attribute is per device */                          \
        int dev_ = 0;                                                                                           \
        cudaGetDevice(&dev_);                                                                                   \
        dev_ = dev_ < kMaxDevices ? dev_ : kMaxDevices - 1;                                                     \
        if (a.L.total > smem_set[dev_] || dev_ == kMaxDevices - 1) {                                            \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, a.L.total); \
            if (e != cudaSuccess) return e;                                                                     \
            smem_set[dev_] = a.L.total;                                                                         \
        }                                                                                                       \
        kern<<<ctas, threads, a.L.total, stream>>>(a);                                                          \
        return cudaGetLastError();                                                                              \
    } while (0)
    switch (vt) {
        case 11: SSD_LAUNCH(11);
        case 15: SSD_LAUNCH(15);
        case 21: SSD_LAUNCH(21);
        default: SSD_LAUNCH(0);
    }
#undef SSD_LAUNCH
}

template <int KIND, bool TAPE, int G>
static cudaError_t launch_fast(const StepArgs& a, int threads, cudaStream_t stream) {
    const int envs_per_cta = (threads / 32) * (32 / G);
    const int ctas = (a.env_end - a.env_begin + envs_per_cta - 1) / envs_per_cta;
    if (ctas <= 0) return cudaSuccess;
#define SSD_LAUNCH_FAST(VT_)                                                                                    \
    do {                                                                                                        \
        auto kern = ssd_step_fast_kernel<KIND, TAPE, VT_, G>;                                                      \
        static uint32_t smem_set[kMaxDevices] = {};  /* the attribute is per device */                          \
        int dev_ = 0;                                                                                           \
        cudaGetDevice(&dev_);                                                                                   \
        dev_ = dev_ < kMaxDevices ? dev_ : kMaxDevices - 1;                                                     \
        if (a.Lf.total > smem_set[dev_] || dev_ == kMaxDevices - 1) {                                           \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, a.Lf.total); \
            if (e != cudaSuccess) return e;                                                                     \
            smem_set[dev_] = a.Lf.total;                                                                        \
        }                                                                                                       \
        cudaLaunchConfig_t lc = {};                                                                             \
        lc.gridDim = dim3(ctas); lc.blockDim = dim3(threads); lc.dynamicSmemBytes = a.Lf.total; lc.stream = stream; \
        cudaLaunchAttribute at[1];                                                                              \
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                          \
        at[0].val.programmaticStreamSerializationAllowed = 1;                                                   \
        lc.attrs = at; lc.numAttrs = a.dep_wait ? 1 : 0;                                                        \
        return cudaLaunchKernelEx(&lc, kern, a);                                                                \
    } while (0)
    switch (a.V) {
        case 11: SSD_LAUNCH_FAST(11);
        case 15: SSD_LAUNCH_FAST(15);
        default: SSD_LAUNCH_FAST(21);
    }
#undef SSD_LAUNCH_FAST
}

static cudaError_t launch_general(const StepArgs& a, int threads, cudaStream_t stream, bool fast_rows) {
    const bool tape = a.tape_u != nullptr || a.tape_move != nullptr;
    switch (a.kind) {
        case SSD_KIND_HARVEST:
            return tape ? launch_v<SSD_KIND_HARVEST, true>(a, threads, stream, fast_rows)
                        : launch_v<SSD_KIND_HARVEST, false>(a, threads, stream, fast_rows);
        case SSD_KIND_CLEANUP:
            return tape ? launch_v<SSD_KIND_CLEANUP, true>(a, threads, stream, fast_rows)
                        : launch_v<SSD_KIND_CLEANUP, false>(a, threads, stream, fast_rows);
        default:
            return tape ? launch_v<SSD_KIND_PLAIN, true>(a, threads, stream, fast_rows)
                        : launch_v<SSD_KIND_PLAIN, false>(a, threads, stream, fast_rows);
    }
}

cudaError_t launch_step(const StepArgs& a, int threads, cudaStream_t stream, ChainState* chain) {
    // the packed row renderers need every warp's slab of 32/G envs to start 4-byte aligned
    const bool fast_rows = (((32 / a.G) * a.obs_env) % 4 == 0) && (reinterpret_cast<uintptr_t>(a.obs) % 4 == 0);
    const bool tape = a.tape_u != nullptr || a.tape_move != nullptr;
    // a production step: everything the specialised kernel assumes (see ssd_step_fast_kernel)
    static const bool no_fast = getenv("SSD_NO_FAST") != nullptr;
    const bool full = !no_fast && a.phases == SSD_PHASE_ALL && a.mask == nullptr && a.order == nullptr && !a.use_beam_buf &&
                      !a.rew_accumulate && a.obs != nullptr && a.rew != nullptr && a.actions != nullptr && fast_rows &&
                      (a.V == 11 || a.V == 15 || a.V == 21) && a.env_begin % (32 / a.G) == 0;
    if (!full) {
        if (chain) chain->valid = false;
        return launch_general(a, threads, stream, fast_rows);
    }
    StepArgs f = a;
    const int epw = 32 / a.G;
    f.env_end = a.env_begin + (a.env_end - a.env_begin) / epw * epw;  // whole warps
    const bool has_tail = f.env_end != a.env_end;
    // Chaining pays when a step is a few waves of CTAs (it hides launch, ramp and the half-empty last wave); a single
    // wave has no tail to hide and very long grids amortise it anyway, while the completion words cost a little
    // (profiles/r01h_sweep.md): chain between 1.5 and 12 waves.
    bool chain_here = false;
    if (chain && chain->enabled && chain->done != nullptr) {
        static int slots = 0;  // resident CTAs of the whole GPU at this CTA shape (8 per SM on B200)
        if (slots == 0) {
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            slots = sms * 8;
        }
        const int ctas = (f.env_end - f.env_begin + (threads / 32) * epw - 1) / ((threads / 32) * epw);
        chain_here = 2 * ctas >= 3 * slots && ctas <= 12 * slots;
        static const bool chain_always = getenv("SSD_CHAIN_ALWAYS") != nullptr;  // experiments
        if (chain_always) chain_here = true;
    }
    if (chain_here) {
        // Chained steps (SSD_OPT_CHAIN_STEPS): this launch may overlap the previous step's kernel when that was the
        // same kind of launch on the same stream; either way it publishes per-task completion words for the next one.
        f.done = chain->done;
        f.epoch = ++chain->epoch;
        f.publish = 1;
        f.dep_wait = chain->valid && chain->stream == stream && chain->env_begin == f.env_begin && chain->env_end == f.env_end;
        chain->valid = !has_tail;
        chain->stream = stream; chain->env_begin = f.env_begin; chain->env_end = f.env_end;
    } else if (chain) {
        chain->valid = false;
    }
    cudaError_t e = cudaSuccess;
#define SSD_FAST(KIND_)                                                                                                     \
    e = a.G == 8 ? (tape ? launch_fast<KIND_, true, 8>(f, threads, stream) : launch_fast<KIND_, false, 8>(f, threads, stream)) \
                 : (tape ? launch_fast<KIND_, true, 16>(f, threads, stream) : launch_fast<KIND_, false, 16>(f, threads, stream))
    switch (a.kind) {
        case SSD_KIND_HARVEST: SSD_FAST(SSD_KIND_HARVEST); break;
        case SSD_KIND_CLEANUP: SSD_FAST(SSD_KIND_CLEANUP); break;
        default: SSD_FAST(SSD_KIND_PLAIN); break;
    }
#undef SSD_FAST
    if (e != cudaSuccess || !has_tail) return e;
    StepArgs tail = a;  // the envs that do not fill a warp
    tail.env_begin = f.env_end;
    return launch_general(tail, threads, stream, fast_rows);
}

cudaError_t launch_reset(const ResetArgs& a, cudaStream_t stream) {
    if (a.env_end <= 0) return cudaSuccess;
    ssd_reset_kernel<<<(a.env_end + 127) / 128, 128, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_pack_state(int kind, int B, int N, int H, int W, int Ws, int env_bytes, const uint8_t* grid_in, const int16_t* pos_in,
                              const uint8_t* ori_in, uint8_t* grid, uint32_t* agents, cudaStream_t stream) {
    const size_t n = static_cast<size_t>(B) * env_bytes;
    pack_state_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(kind, B, N, H, W, Ws, env_bytes, grid_in, pos_in, ori_in, grid, agents);
    return cudaGetLastError();
}
cudaError_t launch_unpack_state(int B, int N, int H, int W, int Ws, int env_bytes, const uint8_t* grid, const uint32_t* agents,
                                uint8_t* grid_out, int16_t* pos_out, uint8_t* ori_out, cudaStream_t stream) {
    const size_t hw = static_cast<size_t>(H) * W;
    const size_t n = static_cast<size_t>(B) * (hw > static_cast<size_t>(N) ? hw : N);
    unpack_state_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(B, N, H, W, Ws, env_bytes, grid, agents, grid_out, pos_out, ori_out);
    return cudaGetLastError();
}
cudaError_t launch_render_map(int B, int N, int H, int W, int Ws, int env_bytes, const uint8_t* grid, const uint32_t* agents,
                              const uint32_t* color, uint8_t* out, cudaStream_t stream) {
    const size_t n = static_cast<size_t>(B) * H * W;
    render_map_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(B, N, H, W, Ws, env_bytes, grid, agents, color, out);
    return cudaGetLastError();
}
cudaError_t launch_philox_selftest(const uint32_t* ctr_key, uint32_t* out, cudaStream_t stream) {
    philox_selftest_kernel<<<1, 1, 0, stream>>>(ctr_key, out);
    return cudaGetLastError();
}

}  // namespace ssd
