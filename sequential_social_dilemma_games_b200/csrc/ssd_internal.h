// Internal definitions shared by the kernels (ssd_step.cu) and the C-ABI host layer (ssd_capi.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/ssd_b200.h"

namespace ssd {

constexpr int kMaxAgents = SSD_MAX_AGENTS;
constexpr int kMaxThreads = 256;

// Philox stream ids (oracle/philox_ref.py restates them for the tests).
enum : uint32_t {
    STREAM_MOVE = 0,   // move-priority Fisher-Yates words        map_env.py:422
    STREAM_SPAWN = 1,  // uniform draws of custom_map_update      harvest.py:101, cleanup.py:139,150
    STREAM_WASTE = 2,  // sort keys of the waste-point order      cleanup.py:145
    STREAM_RPOINT = 3, // sort keys of the spawn-point order      map_env.py:656
    STREAM_RROT = 4,   // spawn rotation                          map_env.py:666
    STREAM_RSPAWN = 5  // uniform draws of reset()'s spawn pass   map_env.py:230
};

// Grid layout.  HBM: [B_pad][H][Ws] bytes, Ws = round_up(W + r, 16): every row carries >= r zero
// bytes after its W cells.  Shared memory: one tile of (H + 2r) rows x Ws per env -- r zero rows,
// the H rows bulk-copied from HBM, r zero rows -- so every cell an egocentric (2r+1)^2 window can
// touch exists and reads as 0 (black, the '0' padding of utility_funcs.py:94-114) outside the map.
// In-tile index of map cell (row, col): (row + r) * Ws + col.

// Per-environment scratch in shared memory.
struct EnvScratch {
    uint16_t pos[kMaxAgents];   // row << 8 | col (live positions)
    uint16_t tgt[kMaxAgents];   // agent_moves values (map_env.py:400)
    uint16_t orig[kMaxAgents];  // search_list: targets frozen before the contested pass (:426)
    uint16_t snap[kMaxAgents];  // agent_by_pos snapshot of one fix-point pass (:495)
    int32_t rew[kMaxAgents];    // -50 per 'F' hit taken in this step (agent.py:166-168)
    uint8_t ori[kMaxAgents];
    uint8_t shuf[kMaxAgents];   // movers after np.random.shuffle (:422)
    uint8_t order[kMaxAgents];  // action-dict iteration order
    uint8_t firech[kMaxAgents]; // [k] beam char of the k-th entry of the action order (0 = did not fire)
    uint8_t raylen[3 * kMaxAgents];  // [k*3+s] painted cells of ray s (beam_pos, map_env.py:648)
    int32_t active;             // 0: env masked out of this launch
    int32_t pad[3];
};
static_assert(sizeof(EnvScratch) % 16 == 0, "EnvScratch must stay 16-byte sized");

// Byte offsets of the dynamic shared-memory carve-up of one CTA (all 16-byte aligned).
struct SmemLayout {
    uint32_t mbar, grid, color, apple, env, list, view, stage, stats, total;
    uint32_t list_stride;  // bytes of spawn scratch per warp
    uint32_t stage_stride; // bytes of render staging per warp (32 view rows)
};

struct StepArgs {
    // ---- static game description
    int kind, H, W, N, r, V, beam_len;
    int Ws;               // grid row stride (bytes)
    int env_bytes;        // H * Ws: one env's grid in HBM
    int pad_bytes;        // r * Ws: zero rows above / below the map in the shared-memory tile
    int tile_stride;      // (H + 2r) * Ws: one env's tile in shared memory
    int n_apple, n_waste, area;
    int obs_env;          // N*V*V*3 bytes
    // ---- launch description
    int E;                // envs per CTA
    int G;                // lanes per env in phase A: 8 (N <= 8) or 16
    int env_begin;        // first local env of this launch (multiple of E)
    int env_end;          // one past the last valid local env
    int phases, rotate;
    uint32_t spawn_stream;
    int use_beam_buf;     // beams cross phase calls through HBM
    int rew_accumulate;
    uint32_t key0, key1, t;
    uint64_t env_id0;     // global id of local env 0
    SmemLayout L;
    // ---- static tables (device)
    const uint16_t* apple_cell; // [n_apple] in-tile cell ids, row-major (harvest.py:22-26, cleanup.py:53-54)
    const uint16_t* waste_cell; // [n_waste] in-tile cell ids, row-major (cleanup.py:59-60)
    const uint32_t* color;      // [128] 0x00BBGGRR by ASCII code
    const uint64_t* harvest_thr; const double* harvest_p;  // [4]
    const uint64_t* apple_thr;   const double* apple_p;    // [area+1]
    const uint64_t* waste_thr;   const double* waste_p;    // [area+1]
    // ---- state (device)
    uint8_t* grid;        // [B_pad][H][Ws]
    uint32_t* agents;     // [B_pad][N] row | col<<8 | ori<<16
    uint8_t* beam_buf;    // [B_pad][64] raylen + firech between phase-split calls
    // ---- I/O (device)
    const int8_t* actions; const uint8_t* order; const uint8_t* mask;
    const uint8_t* tape_move; const double* tape_u; int u_stride; const uint16_t* tape_waste;
    uint8_t* obs; int32_t* rew;
    unsigned long long* stats;
};

struct ResetArgs {
    int N, n_spawn, env_bytes, env_end;
    uint32_t key0, key1, t;
    uint64_t env_id0;
    const uint16_t* spawn_key;  // [n_spawn] row<<8|col, canonical order
    const uint8_t* init_grid;   // [env_bytes]
    const uint8_t* mask;
    uint8_t* grid; uint32_t* agents;
};

// Launchers implemented in ssd_step.cu.
cudaError_t launch_step(const StepArgs& a, int threads, cudaStream_t stream);
cudaError_t launch_reset(const ResetArgs& a, cudaStream_t stream);
cudaError_t launch_pack_state(int B, int N, int H, int W, int Ws, const uint8_t* grid_in, const int16_t* pos_in,
                              const uint8_t* ori_in, uint8_t* grid, uint32_t* agents, cudaStream_t stream);
cudaError_t launch_unpack_state(int B, int N, int H, int W, int Ws, const uint8_t* grid, const uint32_t* agents,
                                uint8_t* grid_out, int16_t* pos_out, uint8_t* ori_out, cudaStream_t stream);
cudaError_t launch_philox_selftest(const uint32_t* ctr_key, uint32_t* out, cudaStream_t stream);

}  // namespace ssd
