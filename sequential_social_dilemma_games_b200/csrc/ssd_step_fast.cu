// The specialised full-step kernel for sm_100a and the dispatcher launch_step (which kernel steps what, and when
// consecutive steps are chained with programmatic dependent launch).
// Reference citations are relative to the reference root (social_dilemmas/envs/...).
#include "ssd_phases.cuh"

namespace ssd {

// ====================================================================== the fast path of a full step
// Same algorithm as ssd_step_kernel, specialised for what a production step is: all phases, every env
// stepped (no mask), N <= 8 (8 lanes per env, 4 envs per warp) or N <= 16 (16 lanes, 2 envs), a packed
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
    // lane 31 parks the mis4 words that spill over a chunk at the head of the next image (mis4 is warp-uniform: a slab that
    // starts 16-byte aligned spills nothing and stores nothing)
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
        if (lane == 31) {  // head of this image = spill of the previous chunk (unused for c == 0)
            if (mis4 >= 1) stage[mis4 - 1] = c2;
            if (mis4 >= 2) stage[mis4 - 2] = c1;
            if (mis4 >= 3) stage[mis4 - 3] = c0;
        }
        c0 = Q[M31 - 3]; c1 = Q[M31 - 2]; c2 = Q[M31 - 1];
        fence_async_smem();
        __syncwarp();
        if (c == n_chunks - 1) nb = hi_last;  // n_chunks >= 2: the last chunk starts at the head of the buffer
        if (elect_one() && !SSD_SKIP(debug, 1)) {  // the elected lane is lane 0 of the converged warp: it owns all bulk groups
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

// WIDE = false: at most 64 registers, eight CTAs of 128 threads per SM -- the shape for batches that fill the GPU, where the
// kernel lives on latency hiding.  WIDE = true: 128 registers, four CTAs per SM: without the register squeeze a warp gets
// through its task about a third faster, which is what counts when the whole batch is less than half a wave of CTAs.
template <int KIND, bool TAPE, int VT, int G, bool ORCH, bool WIDE>
__global__ void __launch_bounds__(kMaxThreads, WIDE ? 2 : 4) ssd_step_fast_kernel(const __grid_constant__ StepArgs a) {
    static_assert(!ORCH || KIND == SSD_KIND_HARVEST, "orchard bitmaps are a Harvest structure");
    constexpr int EPW = 32 / G;                                       // envs per warp: 4 (N <= 8) or 2 (N <= 16)
    using FastScratch = FastScratchT<G>;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(16) uint32_t s_color[kLutEntries];
    __shared__ __align__(8) uint64_t s_tab_bar;
    __shared__ int s_cta_stats[SSD_NUM_STATS];
    __shared__ int s_done;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int N = a.N;
    uint16_t* s_apple = reinterpret_cast<uint16_t*>(smem + a.Lf.apple);
    pdl_launch_dependents();  // a chained next step may start launching (it waits per task, see below)
    SSD_TICK_DECL;

    // CTA-shared tables (colours, apple points, cell -> point): static data, fetched by four bulk copies that travel while the
    // warps set up and issue their tile loads; nobody waits for them before the tiles have been asked for.
    static_assert(kLutEntries * 4 % 16 == 0, "the colour table is one bulk copy");
    if (tid == 0) {
        const uint32_t b_apple = KIND != SSD_KIND_PLAIN ? static_cast<uint32_t>((a.n_apple + 63) & ~63) * 2u : 0u;
        const uint32_t ncw = static_cast<uint32_t>(a.env_bytes + 31) / 32u;
        const uint32_t b_mask = ORCH ? ((ncw * 4u + 15u) & ~15u) : 0u, b_pre = ORCH ? ((ncw * 2u + 15u) & ~15u) : 0u;
        mbar_init(&s_tab_bar, 1);
        mbar_expect_tx(&s_tab_bar, kLutEntries * 4u + b_apple + b_mask + b_pre);
        bulk_g2s(s_color, a.color, kLutEntries * 4u, &s_tab_bar);
        if (b_apple) bulk_g2s(s_apple, a.apple_cell, b_apple, &s_tab_bar);
        if (ORCH) {
            bulk_g2s(smem + a.Lf.pt_mask, a.pt_mask, b_mask, &s_tab_bar);
            bulk_g2s(smem + a.Lf.pt_pre, a.pt_pre, b_pre, &s_tab_bar);
        }
    }
    if (tid < SSD_NUM_STATS) s_cta_stats[tid] = 0;
    if (tid == 0) s_done = 0;
    OrchTables T;
    T.pt_mask = reinterpret_cast<const uint32_t*>(smem + a.Lf.pt_mask); T.pt_pre = reinterpret_cast<const uint16_t*>(smem + a.Lf.pt_pre);
    T.nW = a.nW; T.stride = a.orch_stride; T.nz = a.harvest_nz;
    __syncthreads();  // the table barrier is initialised

    uint8_t* wbase = smem + a.Lf.warp0 + warp * a.Lf.warp_stride;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(wbase + a.Lf.w_mbar);
    uint8_t* tiles = wbase + a.Lf.w_tiles;
    FastScratch* envs = reinterpret_cast<FastScratch*>(wbase + a.Lf.w_env);
    uint32_t* const orch_s = reinterpret_cast<uint32_t*>(wbase + a.Lf.w_union + a.Lf.u_orch);  // [EPW][orch_stride], phases A and B
    const int tile_pitch = a.env_bytes + a.pad_bytes;
    const int we = a.env_begin + (blockIdx.x * nwarps + warp) * EPW;  // the launcher only sends whole warps
    Counters cnt = {0, 0, 0, 0, 0, 0, 0};

    // Every fast launch carries the programmatic-serialization attribute: the CTAs become resident while the previous
    // kernel of the stream drains and do everything above (tables, barrier) early.  Unless this step is chained to it
    // task by task (dep_wait), nothing that kernel wrote is touched before it has completed.
    SSD_TICK_INIT;
    if (a.pdl_wait) pdl_wait();
    SSD_TICK(0);  // waiting for the previous kernel

    if (we < a.env_end) {
        // agent words and actions first: their latency hides behind the set-up of the tile loads.  (A chained step may read
        // the agent words only after it has seen its predecessor's completion word; its actions exist by contract.)
        const int al = lane & (G - 1), gbase = lane & ~(G - 1), j = lane / G;
        const int e = we + j;
        const bool valid = al < N;
        const size_t gi = static_cast<size_t>(e) * N + (valid ? al : 0);
        uint32_t w_agent = 0;
        int act_in = -1;
        if (valid) {
            act_in = a.actions[gi];
            if (!a.dep_wait) w_agent = __ldcg(a.agents + gi);
        }
        FastScratch& S = envs[j];
        uint8_t* g = tiles + a.pad_bytes + j * tile_pitch;
        PhiloxKey pk;
        pk.k0 = a.key0; pk.k1 = a.key1; pk.t = a.t;
        pk.env = static_cast<uint32_t>(a.env_id0 + static_cast<uint64_t>(e));
        AgentLane me;
        me.key = 0x0101; me.ori = 0; me.act = -1; me.rew = 0;
        // A scripted rollout (ssd_rollout: every step's actions exist up front) runs all its steps in this one launch: the warp
        // keeps stepping ITS envs -- they never interact with anybody else's -- so there is no launch, no grid-wide drain and no
        // convoy of warps in the same phase between steps.  The tiles still make the round trip through HBM every step (phase C
        // paints the overlay into them), ordered by the warp's own bulk-copy groups.
        const int n_steps = WIDE ? a.n_steps : 1;  // only the wide kernel carries the step loop (launch_fast)
#pragma unroll 1
      for (int s = 0; s < n_steps; ++s) {
        const size_t so = static_cast<size_t>(s) * a.step_stride;  // [B][N] elements per step of actions / rewards
        if (s > 0) {
            pk.t = a.t + static_cast<uint32_t>(s);
            if (valid) act_in = a.actions[so + gi];
            fence_async_smem();  // the overlay this warp painted into the tiles (generic proxy) before the copy engine overwrites them
        }
        // ---- load: one TMA bulk copy per env tile; zero the frames while they are in flight
        if (lane == 0) {
            if (s == 0) mbar_init(mbar, 1);
            mbar_expect_tx(mbar, static_cast<uint32_t>(EPW) * (a.env_bytes + (ORCH ? 4u * a.orch_stride : 0u)));
            if (a.dep_wait) {  // chained step: the previous step's kernel may still be running; wait for OUR four envs only
                while (ld_acquire_u32(a.done + we / EPW) != a.epoch - 1) __nanosleep(64);
            }
        }
        __syncwarp();
        if (elect_one()) {  // one lane issues all tile loads: uniform operands, no per-lane replay
            if (a.dep_wait) fence_async_all();  // the predecessor's ordinary stores (acquired above) -> our TMA loads
            if (s > 0) bulk_wait_all();         // the previous step's write-back of these tiles has landed
#pragma unroll
            for (int q = 0; q < EPW; ++q)
                bulk_g2s(tiles + a.pad_bytes + q * tile_pitch, a.grid + static_cast<size_t>(we + q) * a.env_bytes, a.env_bytes, mbar);
            if (ORCH) bulk_g2s(orch_s, a.orch + static_cast<size_t>(we) * a.orch_stride, static_cast<uint32_t>(EPW) * 4u * a.orch_stride, mbar);
        }
        if (s == 0) {  // the frames are never written again: rays stop at the walls, windows only read
            const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int q = 0; q <= EPW; ++q)
#pragma unroll 1
                for (int i = lane * 16; i < a.pad_bytes; i += 512) *reinterpret_cast<uint4*>(tiles + q * tile_pitch + i) = z;
        }
        if (valid) {
            if (s == 0) {
                if (a.dep_wait) w_agent = __ldcg(a.agents + gi);  // L2: a chained predecessor may just have written it
                me.key = (w_agent & 255) << 8 | ((w_agent >> 8) & 255);
                me.ori = (w_agent >> 16) & 3;
            }
            me.act = (w_agent >> 24) & 1u ? -2 : act_in;  // -2: parked on a wall cell by ssd_set_state, never acts
            me.rew = 0;
            S.order[al] = a.order != nullptr ? a.order[gi] : static_cast<uint8_t>(al);  // action-dict order (NULL: agent order)
            S.rew[al] = 0;
        }
        SSD_TICK(1);  // issue loads, zero frames
        mbar_wait(&s_tab_bar, 0);  // tables landed (long ago, except in the first tasks of a wave)
        mbar_wait(mbar, static_cast<uint32_t>(s) & 1u);  // tiles landed
        __syncwarp();
        SSD_TICK(2);  // TMA wait

        // ---- phase A: one lane per agent
        moves_group<TAPE>(a, S, reinterpret_cast<MoveScratch*>(wbase + a.Lf.w_union)[j], g, me, valid, al, G, e, pk);
        cnt.steps += (al == 0);
        SSD_TICK(3);  // moves
        if (valid) S.pos[al] = static_cast<uint16_t>(me.key);
        const int my_idx = tile_idx(a, me.key);
        uint32_t* const my_bm = orch_s + j * a.orch_stride;          // ORCH: the bitmaps of this lane's env ...
        const int my_pt = (ORCH && valid) ? cell_point(T, my_idx) : -1;  // ... and the apple point this agent stands on, if any
        {   // consume, map_env.py:178-181: of agents sharing a cell (appendix A.2 quirk) the first in agent order eats
            const uint8_t under = g[my_idx];
            const bool on_apple = valid && is_apple(under);
            const uint32_t same = __match_any_sync(0xffffffffu, on_apple ? (me.key | static_cast<uint32_t>(gbase) << 16) : (0x80000000u | lane));
            const bool ate = on_apple && (__ffs(same) - 1) == lane;
            if (ate) {
                g[my_idx] = CB(C_EMPTY) | (under & 3); me.rew += 1; ++cnt.eaten;
                if (ORCH) {  // the point holds no apple now; it is a spawn candidate if its cached count says so
                    atomicOr(&my_bm[my_pt >> 5], 1u << (my_pt & 31));
                    if ((a.harvest_nz >> (under & 3)) & 1) atomicOr(&my_bm[a.nW + (my_pt >> 5)], 1u << (my_pt & 31));
                }
            }
            __syncwarp();
            if (KIND == SSD_KIND_HARVEST) {
                if (ORCH) recount_events_bm(__ballot_sync(0xffffffffu, ate), j, my_idx, tiles + a.pad_bytes, tile_pitch, a.Ws, orch_s, T);
                else recount_events(__ballot_sync(0xffffffffu, ate), static_cast<int>(g - tiles) + my_idx, tiles, a.Ws);
            }
            // "[row, col] not in self.agent_pos" (harvest.py:90): a point under an agent is not eligible.  After consume it never
            // holds an apple, so its `emp` bit is set: clear it for the spawn pass, set it again afterwards.
            if (ORCH && my_pt >= 0) atomicAnd(&my_bm[my_pt >> 5], ~(1u << (my_pt & 31)));
        }
        __syncwarp();
        if (KIND != SSD_KIND_PLAIN && valid) g[my_idx] |= kFlag;  // "an agent stands here"
        __syncwarp();
        SSD_TICK(4);  // consume, flags
        uint32_t fmask = 0;  // bit (8 * env slot + agent): that agent fires
        uint32_t* const fire_list = reinterpret_cast<uint32_t*>(wbase + a.Lf.w_union);
        const int ray_f = lane / 3, ray_s = lane - 3 * ray_f;  // ray lane -> (firing agent of the round, ray)
        uint32_t fire_ent = 0;
        // Cleanup with an explicit action-dict order takes the literal loop further down; everything else walks the rays of ALL
        // firing agents of the warp at once, three lanes per agent, ten agents per round.
        const bool literal_beams = KIND == SSD_KIND_CLEANUP && a.order != nullptr;
        if (KIND == SSD_KIND_CLEANUP && al == 0 && s == 0) S.hcount = static_cast<int32_t>(__ldcg(a.orch + static_cast<size_t>(e) * a.orch_stride));
        if (KIND != SSD_KIND_PLAIN && !literal_beams && !SSD_SKIP(a.debug, 8)) {
            // update_custom_moves map_env.py:545-552.  'F' beams change no cell, so their order is irrelevant (Harvest has no
            // others).  A CLEAN beam turns the 'H' cell that stops it into 'R' BEFORE the next agent fires (:551-558), so a later
            // beam depends on an earlier one exactly when both stop at the same 'H' cell: the rays of a round are probed without
            // side effects, and only if two of them would clean the same cell is that round replayed agent by agent.  Rounds
            // follow the firing order (env slot, agent), and a round is committed before the next one is probed.
            const bool clean_me = KIND == SSD_KIND_CLEANUP && me.act == 8;
            const bool fire_me = me.act == 7 || clean_me;
            fmask = __ballot_sync(0xffffffffu, fire_me);
            if (fmask) {
                fire_ent = me.key | static_cast<uint32_t>(me.ori) << 16 | static_cast<uint32_t>(j) << 18 | static_cast<uint32_t>(al) << 21 |
                           static_cast<uint32_t>(clean_me) << 25;
                if (fire_me) {
                    fire_list[__popc(fmask & lanemask_lt())] = fire_ent;
                    if (!clean_me) { me.rew -= 1; ++cnt.fires; }  // fire_beam agent.py:170-172 (a CLEAN beam is free, :205-207)
                }
                __syncwarp();
                const int nf = __popc(fmask);
                auto commit = [&](uint32_t en, int ray, int n, int upd, int hit) {
                    const int slot = (en >> 18) & 3, ag = (en >> 21) & 15;
                    envs[slot].raylen[ag * 3 + ray] = static_cast<uint8_t>(n);
                    if (hit >= 0) { atomicAdd(&envs[slot].rew[hit], -50); ++cnt.hits; }  // agent.hit('F') agent.py:166-168
                    if (KIND == SSD_KIND_CLEANUP && upd >= 0) {                          // update_map :551-558
                        uint8_t* t = tiles + a.pad_bytes + slot * tile_pitch + upd;
                        *t = CB(C_RIVER) | (*t & kFlag);
                        atomicSub(&envs[slot].hcount, 1);
                        ++cnt.cleaned;
                    }
                };
#pragma unroll 1
                for (int f0 = 0; f0 < nf; f0 += 10) {
                    const bool act_lane = lane < 30 && f0 + ray_f < nf;
                    uint32_t en = 0;
                    int n = 0, upd = -1, hit = -1;
                    if (act_lane) {
                        en = fire_list[f0 + ray_f];
                        const int slot = (en >> 18) & 3;
                        n = ray_probe(a, envs[slot], tiles + a.pad_bytes + slot * tile_pitch, en & 0xffffu, (en >> 16) & 3, ray_s, (en >> 25) & 1, upd, hit);
                    }
                    bool replay = false;
                    if (KIND == SSD_KIND_CLEANUP) {
                        const uint32_t same = __match_any_sync(0xffffffffu, upd >= 0 ? (static_cast<uint32_t>(upd) | ((en >> 18) & 3) << 16) : (0x80000000u | lane));
                        replay = __any_sync(0xffffffffu, upd >= 0 && (same & (same - 1)) != 0);
                    }
                    if (!replay) {
                        if (act_lane) commit(en, ray_s, n, upd, hit);
                    } else {
                        const int f1 = min(f0 + 10, nf);
#pragma unroll 1
                        for (int f = f0; f < f1; ++f) {  // one agent at a time: its three rays see what the agents before it cleaned
                            n = 0; upd = -1; hit = -1;
                            if (lane < 3) {
                                en = fire_list[f];
                                const int slot = (en >> 18) & 3;
                                n = ray_probe(a, envs[slot], tiles + a.pad_bytes + slot * tile_pitch, en & 0xffffu, (en >> 16) & 3, lane, (en >> 25) & 1, upd, hit);
                            }
                            __syncwarp();  // all three rays are walked before the agent's updates are applied (:551-552)
                            if (lane < 3) commit(en, lane, n, upd, hit);
                            __syncwarp();
                        }
                    }
                    __syncwarp();
                }
            }
        }
        if (literal_beams && !SSD_SKIP(a.debug, 8)) {  // explicit action order: the k-th entry of the action dict is agent S.order[k]
            fmask = __ballot_sync(0xffffffffu, me.act == 7 || me.act == 8);
            for (int k = 0; k < N; ++k) {
                const int ag = S.order[k];
                const bool fire = (fmask >> (gbase + ag)) & 1u;
                if (!__any_sync(0xffffffffu, fire)) continue;  // nobody in this warp fires in slot k
                const int act_k = __shfl_sync(0xffffffffu, me.act, ag, G);
                const uint32_t key_k = __shfl_sync(0xffffffffu, me.key, ag, G);
                const int ori_k = __shfl_sync(0xffffffffu, me.ori, ag, G);
                const bool clean = act_k == 8;
                int upd = -1, hits = 0, n = 0;
                if (fire && al < 3) n = ray_walk(a, S, g, key_k, ori_k, al, clean, upd, hits);
                if (fire && al == ag && !clean) { me.rew -= 1; ++cnt.fires; }  // fire_beam agent.py:170-172
                __syncwarp();
                if (fire && al < 3) {
                    S.raylen[k * 3 + al] = static_cast<uint8_t>(n);
                    if (upd >= 0) { g[upd] = CB(C_RIVER) | (g[upd] & kFlag); atomicSub(&S.hcount, 1); ++cnt.cleaned; }  // update_map :551-558, before the next agent fires
                    cnt.hits += hits;
                }
                __syncwarp();
            }
        }
        if (valid) me.rew += S.rew[al];  // -50 per hit taken
        SSD_TICK(5);  // beams

        // ---- phase B: the whole warp per env
        if (KIND != SSD_KIND_PLAIN && !SSD_SKIP(a.debug, 4)) {
            void* scratch = wbase + a.Lf.w_union;
            if (KIND == SSD_KIND_HARVEST) {
                if (ORCH) {
                    harvest_spawn_bm<TAPE, EPW>(a, tiles, tile_pitch, s_apple, orch_s, T, static_cast<uint32_t*>(scratch), a.Lf.u_words, we, pk, lane, cnt);
                    if (my_pt >= 0) atomicOr(&my_bm[my_pt >> 5], 1u << (my_pt & 31));  // the points under agents are empty points again
                } else {
                    harvest_spawn_warp<TAPE, EPW>(a, tiles, tile_pitch, s_apple, static_cast<uint32_t*>(scratch), a.Lf.u_words, we, pk, lane, cnt);
                }
            } else {
                PhiloxKey pq = pk;
#pragma unroll 1
                for (int q = 0; q < EPW; ++q) {
                    pq.env = static_cast<uint32_t>(a.env_id0 + static_cast<uint64_t>(we + q));
                    cleanup_spawn<TAPE>(a, tiles + a.pad_bytes + q * tile_pitch, s_apple, static_cast<uint32_t*>(scratch), we + q, pq, lane, cnt, &envs[q].hcount);
                    __syncwarp();
                }
            }
            if (valid) g[my_idx] &= 0x7F;  // agents sharing a cell all write the same byte
            __syncwarp();
        }

        SSD_TICK(6);  // spawn
        // ---- store: grid tiles back to HBM, one TMA bulk store per tile.  The overlay below repaints the tiles, so the
        // warp waits until the copy engine has READ them (not until the writes have landed).
        if (!SSD_SKIP(a.debug, 64)) {
            fence_async_smem();  // the phases above wrote the tiles with ordinary stores
            __syncwarp();
            if (elect_one()) {
#pragma unroll
                for (int q = 0; q < EPW; ++q)
                    bulk_s2g(a.grid + static_cast<size_t>(we + q) * a.env_bytes, tiles + a.pad_bytes + q * tile_pitch, a.env_bytes);
                if (ORCH) bulk_s2g(a.orch + static_cast<size_t>(we) * a.orch_stride, orch_s, static_cast<uint32_t>(EPW) * 4u * a.orch_stride);
                bulk_commit();
            }
            bulk_wait_read();
        }
        __syncwarp();  // the write-back above has read the tiles
        SSD_TICK(7);  // grid write-back (until the copy engine has read the tiles)

        // ---- phase C: overlay (get_map_with_agents map_env.py:280-302), view geometry, packed rows
        {
            const uint32_t same = __match_any_sync(0xffffffffu, valid ? (me.key | static_cast<uint32_t>(gbase) << 16) : (0x80000000u | lane));
            if (valid && (31 - __clz(same)) == lane && me.act != -2) g[my_idx] = agent_cell(al);  // the last agent on a cell wins
            __syncwarp();
            if (KIND != SSD_KIND_PLAIN && !literal_beams && fmask) {
                // Beams are painted in firing order and a later one overwrites an earlier one (map_env.py:298-301).  All rays paint
                // at once; that is the same picture unless an 'F' and a 'C' beam cross, which the lanes find out by reading their
                // cells back -- then everything is painted again, agent by agent.
                if ((fmask >> lane) & 1u) fire_list[__popc(fmask & lanemask_lt())] = fire_ent;  // the union was reused by the spawn pass
                __syncwarp();
                const int nf = __popc(fmask);
                auto paint = [&](uint32_t en, int ray, bool check) -> bool {
                    const int slot = (en >> 18) & 3, ag = (en >> 21) & 15, ori = (en >> 16) & 3;
                    const int d0 = (ori == 1) - (ori == 3), d1 = (ori == 2) - (ori == 0);
                    int r = static_cast<int>((en >> 8) & 255) + d0, c = static_cast<int>(en & 255) + d1;
                    if (ray == 1) { r += -d1 - d0; c += d0 - d1; }
                    if (ray == 2) { r -= -d1 + d0; c -= d0 + d1; }
                    const int n = envs[slot].raylen[ag * 3 + ray], dp = d0 * a.Ws + d1;
                    const uint8_t ch = (KIND == SSD_KIND_CLEANUP && ((en >> 25) & 1)) ? CB(C_CLEAN) : CB(C_FIRE);
                    uint8_t* p = tiles + a.pad_bytes + slot * tile_pitch + r * a.Ws + c;
                    bool lost = false;
#pragma unroll 1
                    for (int i = 0; i < n; ++i) {
                        if (check) lost |= (*p != ch); else *p = ch;
                        p += dp;
                    }
                    return lost;
                };
#pragma unroll 1
                for (int f0 = 0; f0 < nf; f0 += 10)
                    if (lane < 30 && f0 + ray_f < nf) paint(fire_list[f0 + ray_f], ray_s, false);
                __syncwarp();
                // (only where some env of the warp has both kinds of beam: bit 25 of an entry marks a CLEAN beam)
                bool mixed = false;
                if (KIND == SSD_KIND_CLEANUP) {
                    const uint32_t cmask = __ballot_sync(0xffffffffu, ((fmask >> lane) & 1u) && ((fire_ent >> 25) & 1u));
                    const uint32_t grp = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << gbase;
                    mixed = __any_sync(0xffffffffu, (cmask & grp) != 0u && (fmask & ~cmask & grp) != 0u);
                }
                if (KIND == SSD_KIND_CLEANUP && mixed) {
                    bool lost = false;
#pragma unroll 1
                    for (int f0 = 0; f0 < nf; f0 += 10)
                        if (lane < 30 && f0 + ray_f < nf) lost |= paint(fire_list[f0 + ray_f], ray_s, true);
                    if (__any_sync(0xffffffffu, lost)) {
#pragma unroll 1
                        for (int f = 0; f < nf; ++f) {
                            if (lane < 3) paint(fire_list[f], lane, false);
                            __syncwarp();
                        }
                    }
                }
                __syncwarp();
            }
            if (literal_beams) {  // beams in firing order: a later beam overwrites an earlier one
                for (int k = 0; k < N; ++k) {
                    const int ag = S.order[k];
                    const bool fired = (fmask >> (gbase + ag)) & 1u;
                    if (!__any_sync(0xffffffffu, fired)) continue;
                    const uint32_t key_k = __shfl_sync(0xffffffffu, me.key, ag, G);
                    const int ori_k = __shfl_sync(0xffffffffu, me.ori, ag, G);
                    const int act_k = __shfl_sync(0xffffffffu, me.act, ag, G);
                    if (fired && al < 3) {
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
            SSD_TICK(8);  // overlay, view parameters
            if (!SSD_SKIP(a.debug, 2))
            render_rows_tma<VT>(s_view, tiles, s_color, reinterpret_cast<uint32_t*>(wbase + a.Lf.w_union + a.Lf.u_stage),
                                a.obs + static_cast<size_t>(s % a.ring_slots) * a.obs_slot_stride + static_cast<size_t>(we) * a.obs_env, EPW * N * VT, a.debug);
        }
        SSD_TICK(9);  // rows
        if (valid) {  // agent words and rewards last: no ordinary global store is in flight when the row loop fences
            if (me.act != -2 && s == n_steps - 1) a.agents[gi] = (me.key >> 8) | (me.key & 255) << 8 | static_cast<uint32_t>(me.ori) << 16;
            a.rew[so + gi] = me.rew;
        }
        if (KIND == SSD_KIND_CLEANUP && al == 0 && s == n_steps - 1) a.orch[static_cast<size_t>(e) * a.orch_stride] = static_cast<uint32_t>(S.hcount);
        if (a.stats != nullptr && !SSD_SKIP(a.debug, 32)) {
            // per-warp totals of ONE step are small (<= 32 agents): two packed reductions carry all seven counters
            //   r0: steps | eaten << 8 | fires << 16 | hits << 24 (hits <= 3 per shooter)     r1: cleaned | waste << 8 | apples << 12
            const uint32_t r0 = __reduce_add_sync(0xffffffffu, static_cast<uint32_t>(cnt.steps | cnt.eaten << 8 | cnt.fires << 16 | cnt.hits << 24));
            const uint32_t r1 = __reduce_add_sync(0xffffffffu, static_cast<uint32_t>(cnt.cleaned | cnt.waste << 8 | cnt.apples << 12));
            if (lane < 7) {
                const uint32_t v = lane < 4 ? (r0 >> (8 * lane)) & 255u : (lane == 4 ? r1 & 255u : (lane == 5 ? r1 >> 12 : (r1 >> 8) & 15u));
                if (v) atomicAdd(&s_cta_stats[lane == 0 ? 0 : lane + 1], static_cast<int>(v));
            }
            cnt = Counters{0, 0, 0, 0, 0, 0, 0};
        }
      }  // steps of a scripted rollout
        if (a.publish) {  // everything this task wrote (grid, agent words, rewards, observation rows) is visible before the word is
            bulk_wait_all();  // all lanes: only the lane that committed the bulk stores actually waits
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                st_release_u32(a.done + we / EPW, a.epoch);
            }
        }
    }

    SSD_TICK(10);  // agent words, rewards, publish
    // ---- stats: warp -> CTA (once per step, above) -> one set of global atomics per CTA (issued by the last warp to finish)
    if (a.stats != nullptr && !SSD_SKIP(a.debug, 32)) {
        __syncwarp();
        int last = 0;
        if (lane == 0) { __threadfence_block(); last = (atomicAdd(&s_done, 1) == nwarps - 1); }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last && lane < SSD_NUM_STATS) {
            const int tot = *reinterpret_cast<volatile int*>(&s_cta_stats[lane]);
            if (tot) atomicAdd(&a.stats[lane], static_cast<unsigned long long>(tot));
        }
    }
    SSD_TICK_FLUSH;
}

// ====================================================================== launchers
template <int KIND, bool TAPE, int G, bool ORCH, bool WIDE>
static cudaError_t launch_fast_w(const StepArgs& a, int threads, cudaStream_t stream) {
    const int envs_per_cta = (threads / 32) * (32 / G);
    const int ctas = (a.env_end - a.env_begin + envs_per_cta - 1) / envs_per_cta;
    if (ctas <= 0) return cudaSuccess;
#define SSD_LAUNCH_FAST(VT_)                                                                                    \
    do {                                                                                                        \
        auto kern = ssd_step_fast_kernel<KIND, TAPE, VT_, G, ORCH, WIDE>;                                                      \
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
        lc.attrs = at; lc.numAttrs = (a.dep_wait || a.pdl_wait) ? 1 : 0;                                                        \
        return cudaLaunchKernelEx(&lc, kern, a);                                                                \
    } while (0)
    switch (a.V) {
        case 11: SSD_LAUNCH_FAST(11);
        case 15: SSD_LAUNCH_FAST(15);
        default: SSD_LAUNCH_FAST(21);
    }
#undef SSD_LAUNCH_FAST
}


// the packed row renderers need every warp's slab of 32/G envs to start 4-byte aligned
static bool packed_rows_ok(const StepArgs& a) {
    return (((32 / a.G) * a.obs_env) % 4 == 0) && (reinterpret_cast<uintptr_t>(a.obs) % 4 == 0) && (a.obs_slot_stride % 4 == 0);
}
// a production step: everything the specialised kernel assumes (see ssd_step_fast_kernel)
static bool production_step(const StepArgs& a, const ChainState* chain) {
    static const bool no_fast = knob("SSD_NO_FAST") != nullptr;
    return !no_fast && !(chain && chain->general_only) && a.phases == SSD_PHASE_ALL && a.mask == nullptr && a.rows == nullptr && !a.use_beam_buf &&
           !a.rew_accumulate && a.obs != nullptr && a.rew != nullptr && a.actions != nullptr && packed_rows_ok(a) &&
           (a.V == 11 || a.V == 15 || a.V == 21) && a.env_begin % (32 / a.G) == 0;
}
// ... and through its wide variant, the one with the step loop: the whole grid resident at four CTAs per SM (launch_fast)
bool specialised_for_all(const StepArgs& a, const ChainState* chain, int threads) {
    const int envs_per_cta = (threads / 32) * (32 / a.G);
    const int ctas = (a.env_end - a.env_begin + envs_per_cta - 1) / envs_per_cta;
    const int cta_slots = chain && chain->cta_slots > 0 ? chain->cta_slots : 148 * 8;
    static const bool no_wide = knob("SSD_NO_WIDE") != nullptr;
    return production_step(a, chain) && (a.env_end - a.env_begin) % (32 / a.G) == 0 && a.tape_u == nullptr && a.tape_move == nullptr &&
           threads <= 128 && 2 * ctas <= cta_slots && !no_wide;
}

// The wide-register kernel when the whole grid is resident at four CTAs per SM anyway (and never for tape replays: parity runs
// need no second set of kernels).
template <int KIND, bool TAPE, int G, bool ORCH = false>
static cudaError_t launch_fast(const StepArgs& a, int threads, cudaStream_t stream, int cta_slots) {
    const int envs_per_cta = (threads / 32) * (32 / G);
    const int ctas = (a.env_end - a.env_begin + envs_per_cta - 1) / envs_per_cta;
    static const bool no_wide = knob("SSD_NO_WIDE") != nullptr;
    if constexpr (!TAPE) {
        if (threads <= 128 && 2 * ctas <= cta_slots && !no_wide) return launch_fast_w<KIND, TAPE, G, ORCH, true>(a, threads, stream);
    }
    return launch_fast_w<KIND, TAPE, G, ORCH, false>(a, threads, stream);
}

cudaError_t launch_step(const StepArgs& a, int threads, cudaStream_t stream, ChainState* chain) {
    const bool fast_rows = packed_rows_ok(a);
    const bool tape = a.tape_u != nullptr || a.tape_move != nullptr;
    const bool full = production_step(a, chain);
    if (a.n_steps != 1 && !specialised_for_all(a, chain, threads)) return cudaErrorInvalidValue;  // ssd_rollout checks before it asks
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
    if (chain && chain->enabled && chain->done != nullptr && a.n_steps == 1) {
        const int slots = chain->cta_slots > 0 ? chain->cta_slots : 148 * 8;  // resident CTAs of this handle's GPU (8 per SM on B200)
        const int ctas = (f.env_end - f.env_begin + (threads / 32) * epw - 1) / ((threads / 32) * epw);
        chain_here = 2 * ctas >= 3 * slots && ctas <= 12 * slots;
        static const bool chain_always = knob("SSD_CHAIN_ALWAYS") != nullptr;  // experiments
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
    static const bool no_pdl = knob("SSD_NO_PDL") != nullptr;  // experiments
    f.pdl_wait = !f.dep_wait && !no_pdl;
    cudaError_t e = cudaSuccess;
    const int cta_slots = chain && chain->cta_slots > 0 ? chain->cta_slots : 148 * 8;
#define SSD_FAST(KIND_)                                                                                                     \
    e = a.G == 8 ? (tape ? launch_fast<KIND_, true, 8>(f, threads, stream, cta_slots) : launch_fast<KIND_, false, 8>(f, threads, stream, cta_slots)) \
                 : (tape ? launch_fast<KIND_, true, 16>(f, threads, stream, cta_slots) : launch_fast<KIND_, false, 16>(f, threads, stream, cta_slots))
    switch (a.kind) {
        case SSD_KIND_HARVEST:
            if (a.use_orch) {
                e = a.G == 8 ? (tape ? launch_fast<SSD_KIND_HARVEST, true, 8, true>(f, threads, stream, cta_slots) : launch_fast<SSD_KIND_HARVEST, false, 8, true>(f, threads, stream, cta_slots))
                             : (tape ? launch_fast<SSD_KIND_HARVEST, true, 16, true>(f, threads, stream, cta_slots) : launch_fast<SSD_KIND_HARVEST, false, 16, true>(f, threads, stream, cta_slots));
            } else {
                SSD_FAST(SSD_KIND_HARVEST);
            }
            break;
        case SSD_KIND_CLEANUP: SSD_FAST(SSD_KIND_CLEANUP); break;
        default: SSD_FAST(SSD_KIND_PLAIN); break;
    }
#undef SSD_FAST
    if (e != cudaSuccess || !has_tail) return e;
    StepArgs tail = a;  // the envs that do not fill a warp
    tail.env_begin = f.env_end;
    return launch_general(tail, threads, stream, fast_rows);
}


}  // namespace ssd
