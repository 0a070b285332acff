// Internal definitions shared by the kernels (ssd_step.cu) and the C-ABI host layer (ssd_capi.cu).
#pragma once
#include <cstdint>
#include <cstdlib>
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

// ---------------------------------------------------------------------------------------------
// Device cell encoding.  A grid byte on the device is (cell code) << 2 | nb:
//   bits 2..6  cell code (below)
//   bits 0..1  nb, Harvest only, ' ' and 'A' cells only: min(3, number of 'A' cells among the 8
//              neighbours) -- the neighbourhood count of spawn_apples (harvest.py:92-100) cached in
//              the grid and refreshed around every cell that gains or loses an apple; 0 elsewhere
//   bit 7      transient "an agent stands here" flag between the consume and spawn phases
// The byte (without bit 7) indexes the colour table directly.  ssd_set_state / ssd_get_state
// translate from / to the reference's ASCII characters (map_env.py:24-41, cleanup.py:15-18).
// ---------------------------------------------------------------------------------------------
enum : uint8_t {
    C_PAD = 0,     // '0'  outside the map (utility_funcs.py:94-114)
    C_EMPTY = 1,   // ' '
    C_WALL = 2,    // '@'
    C_APPLE = 3,   // 'A'
    C_WASTE = 4,   // 'H'
    C_RIVER = 5,   // 'R'
    C_STREAM = 6,  // 'S'
    C_FIRE = 7,    // 'F'  overlay only
    C_CLEAN = 8,   // 'C'  overlay only
    C_AGENT = 9,   // '1'..'9' -> 9..17, overlay only
    C_OTHER = 18,  // any other character handed to ssd_set_state
    kNumCodes = 19,
    kLutEntries = 4 * kNumCodes  // colour table indexed by the grid byte
};
__host__ __device__ constexpr uint8_t CB(uint8_t code) { return static_cast<uint8_t>(code * 4); }
// Phase-skip knobs for timing experiments (profiles/skip_sweep.py).  A production build compiles them out:
// SSD_SKIP(a, bit) is the constant false unless the library is built with -DSSD_PROFILING_KNOBS.
#ifdef SSD_PROFILING_KNOBS
#define SSD_SKIP(dbg, bit) (((dbg) & (bit)) != 0)
#else
#define SSD_SKIP(dbg, bit) false
#endif
// Environment-variable knobs of the tuning scripts under profiles/ (SSD_THREADS, SSD_EXTRA_SMEM, SSD_DEBUG_SKIP, SSD_NO_FAST,
// SSD_CHAIN_ALWAYS, SSD_NO_PDL, SSD_EPW).  A production build never reads the environment: knob() is the constant nullptr unless
// the library is built with -DSSD_PROFILING_KNOBS.
#ifdef SSD_PROFILING_KNOBS
inline const char* knob(const char* name) { return getenv(name); }
#else
inline const char* knob(const char*) { return nullptr; }
#endif

// Per-phase clock accounting of the specialised kernel (profiles/phase_clocks.py): SSD_TICK(i) adds the cycles since the
// previous tick of this warp to prof[i] and counts the tick in prof[16 + i].  Compiled out of production builds.
#ifdef SSD_PROFILING_KNOBS
#define SSD_TICK(i)                                                                                              \
    do {                                                                                                         \
        if (a.prof != nullptr && (threadIdx.x & 31) == 0) {                                                      \
            const long long now_ = clock64();                                                                    \
            atomicAdd(s_prof + (i), static_cast<unsigned long long>(now_ - tick_));                              \
            atomicAdd(s_prof + 16 + (i), 1ull);                                                                  \
            atomicMax(s_prof + 32 + (i), static_cast<unsigned long long>(now_ - tick_));                         \
            tick_ = clock64();                                                                                   \
        }                                                                                                        \
    } while (0)
#define SSD_TICK_DECL __shared__ unsigned long long s_prof[48]; if (threadIdx.x < 48) s_prof[threadIdx.x] = 0
#define SSD_TICK_INIT long long tick_ = clock64()
#define SSD_TICK_FLUSH                                                                                            \
    do {                                                                                                         \
        __syncthreads();                                                                                         \
        if (a.prof != nullptr && threadIdx.x < 32 && s_prof[threadIdx.x]) atomicAdd(a.prof + threadIdx.x, s_prof[threadIdx.x]); \
        if (a.prof != nullptr && threadIdx.x >= 32 && threadIdx.x < 48) atomicMax(a.prof + threadIdx.x, s_prof[threadIdx.x]); \
    } while (0)
#else
#define SSD_TICK(i) do { } while (0)
#define SSD_TICK_DECL do { } while (0)
#define SSD_TICK_INIT do { } while (0)
#define SSD_TICK_FLUSH do { } while (0)
#endif

constexpr uint8_t kFlag = 0x80;
constexpr uint8_t kCodeMask = 0x7C;  // cell code without the neighbour count and the agent flag

// Grid layout.  HBM: [B_pad][env_bytes], row stride Ws = W + r (>= r zero bytes after the W cells of
// every row), env_bytes = round_up(H * Ws, 16).  Shared memory, per warp: the tiles of its envs
// separated (and framed) by pad_bytes >= r * Ws + r zero bytes, so every cell an egocentric
// (2r+1)^2 window can touch exists and reads as C_PAD.  In-tile index of cell (row, col): row * Ws + col.

// Per-environment scratch in shared memory that lives for the whole step.
struct EnvScratch {
    uint16_t pos[kMaxAgents];   // row << 8 | col (live positions)
    int16_t rew[kMaxAgents];    // -50 per 'F' hit taken in this step (agent.py:166-168)
    uint8_t ori[kMaxAgents];
    uint8_t order[kMaxAgents];  // action-dict iteration order
    uint8_t firech[kMaxAgents]; // [k] beam cell byte of the k-th entry of the action order (0 = did not fire)
    uint8_t raylen[3 * kMaxAgents];  // [k*3+s] painted cells of ray s (beam_pos, map_env.py:648)
    int32_t active;             // 0: env masked out of this launch
    int32_t pad[3];
};
static_assert(sizeof(EnvScratch) % 16 == 0, "EnvScratch must stay 16-byte sized");

// The same for the specialised full-step kernel (G = 8 or 16 lanes per env, actions in agent order).
template <int G>
struct FastScratchT {
    uint16_t pos[G];
    int32_t rew[G];   // 32-bit: rays of different shooters add their -50 with shared-memory atomics
    uint8_t raylen[3 * G];
    uint8_t order[G];
    int32_t hcount;   // Cleanup: running number of 'H' cells of the env (StepArgs::orch word 0)
    int32_t pad[3];
};
static_assert(sizeof(FastScratchT<8>) == 96 && sizeof(FastScratchT<16>) == 176, "FastScratchT is 10 bytes per lane + 16");

// Scratch of the literal update_moves emulation (moves_coop); lives in the per-warp phase union.
struct MoveScratch {
    uint16_t orig[kMaxAgents];  // rank of every mover in the shuffled order
    uint8_t shuf[kMaxAgents];   // movers after np.random.shuffle (map_env.py:422)
};
static_assert(sizeof(MoveScratch) % 16 == 0, "MoveScratch must stay 16-byte sized");

// Dynamic shared memory: [apple table][warp 0 region][warp 1 region]...; offsets inside a warp region.
// The phases of a step never overlap inside a warp, so their scratch shares one union:
//   moves:  MoveScratch per env        spawn: need-list / waste keys        render: view params + staging
struct SmemLayout {
    uint32_t apple;                                  // CTA-shared table
    uint32_t pt_mask, pt_pre;                        // CTA-shared cell -> apple-point tables (orchard bitmaps, specialised Harvest kernel)
    uint32_t u_orch;                                 // orchard bitmaps of the warp's envs: the tail of the union, live in phases A and B
    uint32_t warp0, warp_stride;                     // first warp region, bytes per warp
    uint32_t w_mbar, w_tiles, w_env, w_union;        // offsets inside a warp region
    uint32_t u_stage;                                // staging buffer inside the union (after the view params)
    uint32_t u_words;                                // 32-bit words in the union
    uint32_t total;
};

struct StepArgs {
    // ---- static game description
    int kind, H, W, N, r, V, beam_len;
    int Ws;               // grid row stride (bytes)
    int env_bytes;        // round_up(H * Ws, 16): one env's grid in HBM
    int pad_bytes;        // zero bytes between / around the tiles in shared memory
    int n_apple, n_waste, area;
    int harvest_nz;       // bit n: SPAWN_PROB[n] != 0 (harvest.py:13)
    int nW;               // Harvest: 32-bit words per orchard bitmap = ceil(n_apple / 32)
    int orch_stride;      // 32-bit words of orchard bitmaps per env in HBM and in shared memory (2 * nW rounded up to 16 bytes)
    int use_orch;         // the specialised kernel scans the bitmaps instead of the apple points (all the warp's words fit one pass)
    int debug;            // SSD_DEBUG_SKIP bits; only read by builds with -DSSD_PROFILING_KNOBS (profiles/skip_sweep.py)
    int obs_env;          // N*V*V*3 bytes
    // ---- launch description
    int G;                // lanes per env in phase A: 8 (N <= 8) or 16
    int env_begin;        // first local env of this launch (multiple of the CTA tile)
    int env_end;          // one past the last valid local env
    int phases, rotate;
    uint32_t spawn_stream;
    int use_beam_buf;     // beams cross phase calls through HBM
    int rew_accumulate;
    uint32_t key0, key1, t;
    uint64_t env_id0;     // global id of local env 0
    SmemLayout L;         // general kernel
    SmemLayout Lf;        // specialised full-step kernel
    // ---- static tables (device)
    const uint16_t* apple_cell; // [n_apple] in-tile cell ids, row-major (harvest.py:22-26, cleanup.py:53-54)
    const uint16_t* waste_cell; // [n_waste] in-tile cell ids, row-major (cleanup.py:59-60)
    const uint32_t* color;      // [32] 0x00BBGGRR by cell code
    const uint64_t* harvest_thr; const double* harvest_p;  // [4]
    const uint64_t* apple_thr;   const double* apple_p;    // [area+1]
    const uint64_t* waste_thr;   const double* waste_p;    // [area+1]
    // ---- state (device)
    uint8_t* grid;        // [B_pad][env_bytes]
    uint32_t* agents;     // [B_pad][N] row | col<<8 | ori<<16 | parked<<24 (parked: uploaded onto a '@' cell; never acts, never painted)
    uint8_t* beam_buf;    // [B_pad][64] raylen + firech between phase-split calls
    // Harvest orchard bitmaps, [B_pad][orch_stride] words: bit p of word p/32 of `emp` (words 0..nW-1) = apple point p holds no
    // apple; of `need` (words nW..2nW-1) = it holds none AND SPAWN_PROB[cached neighbour count] != 0, i.e. it is a candidate of
    // spawn_apples (harvest.py:87-101).  Kept exact by every kernel that changes a grid.
    // Cleanup: [B_pad][4] words, word 0 = number of 'H' cells of the env's grid (compute_permitted_area, cleanup.py:173-179).
    uint32_t* orch;
    const uint32_t* pt_mask;    // [ceil(env_bytes/32)] bit c%32 of word c/32: tile cell c is an apple point
    const uint16_t* pt_pre;     // [ceil(env_bytes/32)] number of apple points in the words before
    // ---- I/O (device)
    const int8_t* actions; const uint8_t* order; const uint8_t* mask;
    const int32_t* rows; int n_rows;  // general kernel: step only these envs, one warp per listed row (ssd_reset_rows)
    const uint8_t* tape_move; const double* tape_u; int u_stride; const uint16_t* tape_waste; int32_t* n_draws_out;
    uint8_t* obs; int32_t* rew;
    unsigned long long* stats;
    // ---- chained steps (SSD_OPT_CHAIN_STEPS): done[task] = epoch of the last full step that finished the
    // task's envs (task = 4 consecutive envs = one warp of the specialised kernel)
    uint32_t* done;
    uint32_t epoch;
    int dep_wait;         // wait for done[task] == epoch - 1 instead of relying on stream order
    int publish;          // write done[task] = epoch when the task's results are visible
    int pdl_wait;         // launched early behind the previous kernel of the stream: griddepcontrol.wait after the prologue
    // ---- scripted rollouts (ssd_rollout): n_steps > 1 steps in ONE launch of the specialised kernel; step s reads actions + s *
    // step_stride, writes rewards + s * step_stride and observations into slot s % ring_slots (obs + slot * obs_slot_stride)
    int n_steps, ring_slots;
    size_t step_stride, obs_slot_stride;
    unsigned long long* prof;  // SSD_PROFILING_KNOBS builds with SSD_PROF set: [48] per-phase cycle sums, tick counts, maxima
};

// Host-side bookkeeping of the step chain (one per handle).
struct ChainState {
    bool enabled = false, valid = false;
    bool general_only = false;  // SSD_OPT_GENERAL_KERNEL
    cudaStream_t stream = nullptr;
    int env_begin = 0, env_end = 0;
    uint32_t epoch = 0;
    uint32_t* done = nullptr;
    int cta_slots = 0;          // resident CTAs of the handle's GPU at the step kernel's CTA shape (set by ssd_create)
};

struct ResetArgs {
    int N, n_spawn, env_bytes, env_end;
    uint32_t key0, key1, t;
    uint64_t env_id0;
    const uint16_t* spawn_key;  // [n_spawn] row<<8|col, canonical order
    const uint8_t* init_grid;   // [env_bytes]
    const uint8_t* mask;
    const int32_t* rows; int n_rows;  // reset only these envs (ssd_reset_rows); NULL: all envs below env_end
    uint8_t* grid; uint32_t* agents;
    uint32_t* orch; int orch_stride;  // Harvest orchard bitmaps: all zero after reset_map (every point holds an apple) ...
    uint32_t orch_word0;              // ... Cleanup: word 0 = number of 'H' cells of the reset grid
};

// Launchers.  launch_step (ssd_step_fast.cu) decides which kernel steps what; launch_general (ssd_step_general.cu) is
// the kernel with all the runtime flags; the rest live in ssd_aux.cu.
constexpr int kMaxDevices = 64;
cudaError_t launch_general(const StepArgs& a, int threads, cudaStream_t stream, bool fast_rows);
cudaError_t launch_step(const StepArgs& a, int threads, cudaStream_t stream, ChainState* chain = nullptr);
// true when launch_step would send ALL envs of `a` through the wide variant of the specialised kernel (so a.n_steps > 1 is allowed)
bool specialised_for_all(const StepArgs& a, const ChainState* chain, int threads);
cudaError_t launch_reset(const ResetArgs& a, cudaStream_t stream);
cudaError_t launch_pack_state(int kind, int B, int N, int H, int W, int Ws, int env_bytes, const uint8_t* grid_in, const int16_t* pos_in,
                              const uint8_t* ori_in, uint8_t* grid, uint32_t* agents, cudaStream_t stream);
// Harvest: orchard bitmaps of every env recomputed from the grid (after ssd_set_state); one warp per env
cudaError_t launch_build_orch(int kind, int B, int n_apple, int nW, int orch_stride, int harvest_nz, int env_bytes, const uint16_t* apple_cell,
                              const uint8_t* grid, uint32_t* orch, cudaStream_t stream);
// counts the agents whose (row, col) lies outside the map into *bad (device int, zeroed by the caller)
cudaError_t launch_check_positions(int B, int N, int H, int W, const int16_t* pos_in, int* bad, cudaStream_t stream);
cudaError_t launch_unpack_state(int B, int N, int H, int W, int Ws, int env_bytes, const uint8_t* grid, const uint32_t* agents,
                                uint8_t* grid_out, int16_t* pos_out, uint8_t* ori_out, cudaStream_t stream);
cudaError_t launch_render_map(int B, int N, int H, int W, int Ws, int env_bytes, const uint8_t* grid, const uint32_t* agents,
                              const uint32_t* color, uint8_t* out, cudaStream_t stream);
cudaError_t launch_philox_selftest(const uint32_t* ctr_key, uint32_t* out, cudaStream_t stream);

// ASCII <-> device cell byte (host helpers shared with ssd_capi.cu)
inline uint8_t ascii_to_cell(uint8_t ch) {
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

// records the thread-local message behind ssd_last_error() and returns `code` (ssd_capi.cu)
int set_error(int code, const char* msg);

}  // namespace ssd
