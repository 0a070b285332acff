// The general MapEnv.step kernel for sm_100a: every WARP steps its own 32/G environments end to end (no CTA
// barrier after the table load), with runtime flags for everything a caller may ask for -- a subset of the
// phases (the per-env adapters run MOVES|CONSUME, a Python hook, then SPAWN|RENDER), a reset mask, an explicit
// action order, beams crossing phase calls through HBM, any view size.  Production steps take the specialised
// kernel in ssd_step_fast.cu.
// Reference citations are relative to the reference root (social_dilemmas/envs/...).
#include "ssd_phases.cuh"

namespace ssd {

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
    if (phases & (SSD_PHASE_SPAWN | SSD_PHASE_CONSUME))
        for (int i = tid; i < ((a.n_apple + 63) & ~63); i += nthr) s_apple[i] = i < a.n_apple ? a.apple_cell[i] : static_cast<uint16_t>(a.Ws + 1);
    __syncthreads();

    // ---- this warp's envs
    uint8_t* wbase = smem + a.L.warp0 + warp * a.L.warp_stride;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(wbase + a.L.w_mbar);
    uint8_t* tiles = wbase + a.L.w_tiles;
    EnvScratch* envs = reinterpret_cast<EnvScratch*>(wbase + a.L.w_env);
    const int tile_pitch = a.env_bytes + a.pad_bytes;
    // a.rows (ssd_reset_rows): one warp per listed env; it loads the group of EPW envs around it and steps that one only
    int we = a.env_begin + (blockIdx.x * nwarps + warp) * EPW;  // first local env of this warp
    int row = -1;
    if (a.rows != nullptr) {
        const int wi = blockIdx.x * nwarps + warp;
        row = wi < a.n_rows ? a.rows[wi] : -1;
        we = (row >= 0 && row < a.env_end) ? row / EPW * EPW : a.env_end;
    }
    const int nvalid = max(0, min(EPW, a.env_end - we));
    Counters cnt = {0, 0, 0, 0, 0, 0, 0};

    if (nvalid > 0) {
        // ---- load: one TMA bulk copy per env tile; zero frames while they are in flight
        if (lane == 0) { mbar_init(mbar, 1); mbar_expect_tx(mbar, static_cast<uint32_t>(EPW) * a.env_bytes); }
        __syncwarp();
        if (elect_one())  // one lane issues all tile loads: uniform operands, no per-lane replay
            for (int q = 0; q < EPW; ++q)
                bulk_g2s(tiles + a.pad_bytes + q * tile_pitch, a.grid + static_cast<size_t>(we + q) * a.env_bytes, a.env_bytes, mbar);
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
        const bool active = j < nvalid && (a.mask == nullptr || a.mask[e] != 0) && (a.rows == nullptr || e == row);
        const bool valid = al < N;
        const size_t gi = static_cast<size_t>(e) * N + (valid ? al : 0);
        PhiloxKey pk;
        pk.k0 = a.key0; pk.k1 = a.key1; pk.t = a.t;
        pk.env = static_cast<uint32_t>(a.env_id0 + static_cast<uint64_t>(e));
        AgentLane me;
        me.key = 0x0101; me.ori = 0; me.act = -1; me.rew = 0;
        bool parked = false;  // uploaded onto a wall cell by ssd_set_state: never acts, never painted
        if (valid) {
            const uint32_t w = a.agents[gi];
            me.key = (w & 255) << 8 | ((w >> 8) & 255);
            me.ori = (w >> 16) & 3;
            parked = (w >> 24) & 1u;
            if (active && a.actions && !parked) me.act = a.actions[gi];
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
                a.agents[gi] = (me.key >> 8) | (me.key & 255) << 8 | static_cast<uint32_t>(me.ori) << 16 | static_cast<uint32_t>(parked) << 24;
                if (a.rew) a.rew[gi] = me.rew;
            }
        }
        // beams recorded by an earlier phase call of this step (phase-split mode only)
        if (a.use_beam_buf) {
            // a call that moves the agents without a beam phase starts a NEW step whose beams (if any) come from the caller's own
            // custom_action hook: forget what an earlier step's device beam phase recorded, or a later RENDER call paints it
            const bool fresh = (phases & SSD_PHASE_MOVES) && !(phases & SSD_PHASE_BEAMS);
            const bool load = !fresh && (phases & SSD_PHASE_RENDER) && !(phases & SSD_PHASE_BEAMS);
            const bool store = (phases & SSD_PHASE_BEAMS) && !(phases & SSD_PHASE_RENDER);
            for (int q = 0; q < EPW; ++q) {
                if (!envs[q].active) continue;
                for (int i = lane; i < 64; i += 32) {
                    uint8_t* p = i < 48 ? &envs[q].raylen[i] : &envs[q].firech[i - 48];
                    uint8_t* gp = a.beam_buf + static_cast<size_t>(we + q) * 64 + i;
                    if (fresh) { *gp = 0; *p = 0; }
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

        // ---- Harvest: the orchard bitmaps of the specialised kernel (StepArgs::orch), rebuilt from the tiles this call changed
        if (KIND == SSD_KIND_HARVEST && a.nW > 0 && (phases & (SSD_PHASE_CONSUME | SSD_PHASE_SPAWN))) {
            for (int q = 0; q < EPW; ++q) {
                if (!envs[q].active) continue;
                const uint8_t* tq = tiles + a.pad_bytes + q * tile_pitch;
                uint32_t* bm = a.orch + static_cast<size_t>(we + q) * a.orch_stride;
                for (int w = 0; w < a.nW; ++w) {
                    const int i = 32 * w + lane;
                    const uint8_t c = i < a.n_apple ? tq[s_apple[i]] : 0;
                    const bool emp = i < a.n_apple && (c & kCodeMask) == CB(C_EMPTY);
                    const uint32_t m_e = __ballot_sync(0xffffffffu, emp), m_n = __ballot_sync(0xffffffffu, emp && ((a.harvest_nz >> (c & 3)) & 1));
                    if (lane == 0) { bm[w] = m_e; bm[a.nW + w] = m_n; }
                }
            }
        }

        if (KIND == SSD_KIND_CLEANUP && (phases & (SSD_PHASE_BEAMS | SSD_PHASE_SPAWN))) {  // the running 'H' count of the specialised kernel
            for (int q = 0; q < EPW; ++q) {
                if (!envs[q].active) continue;
                const int nh = count_waste(a, tiles + a.pad_bytes + q * tile_pitch, lane);
                if (lane == 0) a.orch[static_cast<size_t>(we + q) * a.orch_stride] = static_cast<uint32_t>(nh);
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
                if (valid && (31 - __clz(same)) == lane && !parked) g[tile_idx(a, key)] = agent_cell(al);
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
            const bool all_active = (nvalid == EPW) && (a.mask == nullptr) && (a.rows == nullptr);
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

// ====================================================================== launchers
template <int KIND, bool TAPE>
static cudaError_t launch_v(const StepArgs& a, int threads, cudaStream_t stream, bool fast_rows) {
    const int envs_per_cta = (threads / 32) * (32 / a.G);
    const int ctas = a.rows != nullptr ? (a.n_rows + threads / 32 - 1) / (threads / 32)  // one warp per listed env
                                       : (a.env_end - a.env_begin + envs_per_cta - 1) / envs_per_cta;
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


cudaError_t launch_general(const StepArgs& a, int threads, cudaStream_t stream, bool fast_rows) {
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


}  // namespace ssd
