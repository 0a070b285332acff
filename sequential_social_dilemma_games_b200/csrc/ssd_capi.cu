// C-ABI host layer of libssd_b200 (include/ssd_b200.h): handle management, static tables,
// shared-memory carve-up, launch plumbing.  No torch types, no exceptions across the boundary.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "ssd_internal.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) return fail(SSD_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

uint32_t up16(uint32_t x) { return (x + 15u) & ~15u; }

// u < p for u = k / 2^53  <=>  k < ceil(p * 2^53); p * 2^53 is an exact scaling of a double.
uint64_t threshold53(double p) {
    if (!(p > 0.0)) return 0;
    if (p >= 1.0) return 1ull << 53;
    return static_cast<uint64_t>(std::ceil(std::ldexp(p, 53)));
}

}  // namespace

namespace ssd {
int set_error(int code, const char* msg) { return fail(code, "%s", msg); }
}  // namespace ssd

struct SsdEnv {
    SsdConfig cfg{};
    int B = 0, B_pad = 0, E = 0, threads = 128;
    int HW = 0, Ws = 0, env_bytes = 0, pad_bytes = 0, V = 0, obs_env = 0, n_apple = 0, n_waste = 0, n_spawn = 0;
    uint64_t seed = 0;
    int harvest_nz = 0;
    int distinct_spawn = 0;
    uint32_t t = 0;
    int64_t launches = 0;
    ssd::SmemLayout L{}, Lf{};
    ssd::ChainState chain{};
    // device allocations
    std::vector<void*> allocs;
    uint16_t* d_apple = nullptr;
    uint16_t* d_waste = nullptr; uint16_t* d_spawn = nullptr; uint32_t* d_color = nullptr;
    uint64_t* d_hthr = nullptr; double* d_hp = nullptr;
    uint64_t* d_athr = nullptr; double* d_ap = nullptr; uint64_t* d_wthr = nullptr; double* d_wp = nullptr;
    uint8_t* d_init_grid = nullptr;
    uint8_t* d_grid = nullptr; uint32_t* d_agents = nullptr; uint8_t* d_beam_buf = nullptr;
    uint32_t* d_orch = nullptr; uint32_t* d_pt_mask = nullptr; uint16_t* d_pt_pre = nullptr;
    int nW = 0, orch_stride = 0, use_orch = 0;
    uint32_t init_waste = 0;  // 'H' cells of the reset grid (Cleanup)
    unsigned long long* d_stats = nullptr;
    int* d_bad = nullptr;  // ssd_set_state: number of agent positions outside the map
    unsigned long long* d_prof = nullptr;  // profiling builds only (SSD_PROF)
    int32_t* d_rows = nullptr; int rows_cap = 0;  // ssd_reset_rows: device copy of a host row list
    // ssd_step_host plumbing
    bool host_ready = false;
    cudaEvent_t host_ev = nullptr;
    cudaStream_t hs[2] = {nullptr, nullptr};
    int8_t* d_act_stage = nullptr; uint8_t* d_obs_stage = nullptr; int32_t* d_rew_stage = nullptr;
    // staging for host-pointer set/get state
    uint8_t* d_io_grid = nullptr; int16_t* d_io_pos = nullptr; uint8_t* d_io_ori = nullptr;

    template <typename T>
    int alloc(T** p, size_t n) {
        void* q = nullptr;
        if (cudaMalloc(&q, n * sizeof(T) > 0 ? n * sizeof(T) : 16) != cudaSuccess) return -1;
        allocs.push_back(q);
        *p = static_cast<T*>(q);
        return 0;
    }
    template <typename T>
    int upload(T** p, const std::vector<T>& v) {
        if (alloc(p, v.size())) return -1;
        if (!v.empty() && cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
        return 0;
    }
};

namespace {

bool is_device_ptr(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

ssd::SmemLayout make_layout(const SsdEnv& h, int threads, bool fast = false) {
    ssd::SmemLayout L{};
    const uint32_t G = h.cfg.num_agents <= 8 ? 8 : 16, epw = 32 / G;
    uint32_t off = 0;
    L.apple = off; off += up16(((h.n_apple + 63) & ~63) * 2);  // padded to two whole warps
    const uint32_t orch_bytes = (fast && h.use_orch) ? epw * h.orch_stride * 4u : 0u;
    if (orch_bytes) {  // cell -> apple-point tables of the orchard bitmaps
        const uint32_t ncw = (h.env_bytes + 31) / 32;
        L.pt_mask = off; off += up16(ncw * 4);
        L.pt_pre = off; off += up16(ncw * 2);
    }
    L.warp0 = off;
    uint32_t w = 0;
    L.w_mbar = w; w += 16;
    L.w_tiles = w; w += epw * (h.env_bytes + h.pad_bytes) + h.pad_bytes;
    L.w_env = w; w += epw * (fast ? 10u * G + 16u : sizeof(ssd::EnvScratch));  // FastScratchT<G> is 10 bytes per lane + 16
    L.w_union = w;
    L.u_stage = up16(epw * h.cfg.num_agents * 8);                          // view params first
    const uint32_t u_render = L.u_stage + up16(32u * 3u * h.V) + 32;       // + staging of 32 view rows, spill and dummy words
    const uint32_t u_spawn = up16(std::max(std::max(h.n_apple * 4, h.n_waste * 4), h.cfg.kind == SSD_KIND_CLEANUP ? 512 : 0)); // need-list / waste keys / 32 Philox blocks
    const uint32_t u_moves = epw * sizeof(ssd::MoveScratch);
    // the orchard bitmaps sit behind whatever phases A and B keep in the union (they are stored back before phase C reuses it)
    const uint32_t u_ab = up16(std::max(u_spawn, std::max(u_moves, 32u * 4u)));  // need-list / move scratch / fire list
    L.u_orch = u_ab;
    const uint32_t union_bytes = std::max(u_render, u_ab + orch_bytes);
    L.u_words = (orch_bytes ? u_ab : union_bytes) / 4;  // capacity of the candidate list
    w += union_bytes;
    if (const char* x = ssd::knob("SSD_EXTRA_SMEM")) w += up16(static_cast<uint32_t>(atoi(x)));  // occupancy experiments
    L.warp_stride = w;
    L.total = off + (threads / 32) * w;
    return L;
}

void fill_args(SsdEnv* h, ssd::StepArgs& a) {
    memset(&a, 0, sizeof a);
    const SsdConfig& c = h->cfg;
    a.kind = c.kind; a.H = c.height; a.W = c.width; a.N = c.num_agents; a.r = c.view_radius; a.V = h->V;
    a.beam_len = c.beam_len; a.Ws = h->Ws; a.env_bytes = h->env_bytes; a.pad_bytes = h->pad_bytes;
    a.n_apple = h->n_apple; a.n_waste = h->n_waste; a.area = c.potential_waste_area;
    a.harvest_nz = h->harvest_nz;
    a.nW = h->nW; a.orch_stride = h->orch_stride; a.use_orch = h->use_orch;
    a.orch = h->d_orch; a.pt_mask = h->d_pt_mask; a.pt_pre = h->d_pt_pre;
    { static const int dbg = ssd::knob("SSD_DEBUG_SKIP") ? atoi(ssd::knob("SSD_DEBUG_SKIP")) : 0; a.debug = dbg; }
    a.obs_env = h->obs_env;
    a.G = h->cfg.num_agents <= 8 ? 8 : 16; a.env_begin = 0; a.env_end = h->B;
    a.phases = SSD_PHASE_ALL; a.rotate = 1; a.spawn_stream = ssd::STREAM_SPAWN;
    a.n_steps = 1; a.ring_slots = 1;
    a.key0 = static_cast<uint32_t>(h->seed); a.key1 = static_cast<uint32_t>(h->seed >> 32); a.t = h->t;
    a.env_id0 = c.env_id_offset;
    a.L = h->L; a.Lf = h->Lf;
    a.apple_cell = h->d_apple; a.waste_cell = h->d_waste;
    a.color = h->d_color; a.harvest_thr = h->d_hthr; a.harvest_p = h->d_hp;
    a.apple_thr = h->d_athr; a.apple_p = h->d_ap; a.waste_thr = h->d_wthr; a.waste_p = h->d_wp;
    a.grid = h->d_grid; a.agents = h->d_agents; a.beam_buf = h->d_beam_buf;
    a.stats = h->d_stats;
    a.prof = h->d_prof;
}

int check_handle(ssd_handle h) { return h ? 0 : fail(SSD_ERR_INVALID, "null handle"); }

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

const char* ssd_last_error(void) { return g_err; }
int ssd_abi_version(void) { return SSD_ABI_VERSION; }

int ssd_create(const SsdConfig* cfg, ssd_handle* out) {
    if (!cfg || !out) return fail(SSD_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->abi_version != SSD_ABI_VERSION) return fail(SSD_ERR_INVALID, "ABI version mismatch: caller %d, library %d", cfg->abi_version, SSD_ABI_VERSION);
    const int H = cfg->height, W = cfg->width, N = cfg->num_agents;
    if (cfg->kind < 0 || cfg->kind > SSD_KIND_PLAIN) return fail(SSD_ERR_INVALID, "unknown kind %d", cfg->kind);
    if (H < 3 || W < 3 || H > 255 || W > 255) return fail(SSD_ERR_INVALID, "map shape %dx%d outside 3..255", H, W);
    if (N < 1 || N > SSD_MAX_AGENTS) return fail(SSD_ERR_INVALID, "num_agents %d outside 1..%d", N, SSD_MAX_AGENTS);
    if (cfg->view_radius < 0 || cfg->view_radius > 31) return fail(SSD_ERR_INVALID, "view_radius %d outside 0..31", cfg->view_radius);
    if (cfg->beam_len < 0 || cfg->beam_len > 32) return fail(SSD_ERR_INVALID, "beam_len %d outside 0..32", cfg->beam_len);
    if (cfg->num_envs < 1) return fail(SSD_ERR_INVALID, "num_envs must be positive");
    if (!cfg->base_map || !cfg->color_lut) return fail(SSD_ERR_INVALID, "base_map and color_lut are required");
    if (cfg->kind == SSD_KIND_HARVEST && !cfg->harvest_spawn_prob) return fail(SSD_ERR_INVALID, "harvest_spawn_prob is required");
    if (cfg->kind == SSD_KIND_CLEANUP && (!cfg->cleanup_apple_prob || !cfg->cleanup_waste_prob || cfg->potential_waste_area < 0))
        return fail(SSD_ERR_INVALID, "cleanup probability tables are required");
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            const uint8_t ch = cfg->base_map[r * W + c];
            if (ch >= 128) return fail(SSD_ERR_INVALID, "map characters must be 7-bit ASCII");
            if ((r == 0 || c == 0 || r == H - 1 || c == W - 1) && ch != '@')
                return fail(SSD_ERR_INVALID, "the map must be enclosed by '@' walls (cell %d,%d): the reference indexes grid[new_row, new_col] unchecked (agent.py:111)", r, c);
        }
    int distinct_spawn = 0;
    {
        std::vector<int> seen;
        for (int s = 0; s < cfg->num_spawn_points; ++s) {
            const int r = cfg->spawn_points[2 * s], c = cfg->spawn_points[2 * s + 1];
            if (r < 0 || r >= H || c < 0 || c >= W) return fail(SSD_ERR_INVALID, "spawn point %d outside the map", s);
            const int key = r * W + c;
            bool dup = false;
            for (int k : seen) dup |= (k == key);
            if (!dup) seen.push_back(key);
        }
        distinct_spawn = static_cast<int>(seen.size());  // checked by ssd_reset: agents may also be placed with ssd_set_state
    }
    CUDA_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) return fail(SSD_ERR_UNSUPPORTED, "libssd_b200 is built for sm_100a only; device %d is sm_%d%d", cfg->device, prop.major, prop.minor);

    SsdEnv* h = new (std::nothrow) SsdEnv();
    if (!h) return fail(SSD_ERR_INVALID, "out of host memory");
    h->cfg = *cfg;
    h->distinct_spawn = distinct_spawn;
    h->B = cfg->num_envs; h->HW = H * W;
    // Row stride of the grid: at least r zero bytes after the W cells of every row (the right-hand padding of return_view), and
    // such that the 2r+1 rows of one egocentric window start in 2r+1 different shared-memory banks whatever the window's
    // alignment -- the lanes of the row renderer read one window row each (Harvest: 38 + 7 = 45 already is; Cleanup: 25 -> 27).
    h->Ws = W + cfg->view_radius;
    {
        const int V = 2 * cfg->view_radius + 1;
        auto conflict_free = [V](int ws) {
            if (V > 32) return true;
            for (int a0 = 0; a0 < 4; ++a0) {
                uint32_t seen = 0;
                for (int i = 0; i < V; ++i) {
                    const uint32_t bank = 1u << (((a0 + ws * i) >> 2) & 31);
                    if (seen & bank) return false;
                    seen |= bank;
                }
            }
            return true;
        };
        for (int extra = 0; extra < 16 && !conflict_free(h->Ws); ++extra) ++h->Ws;
        if (!conflict_free(h->Ws)) h->Ws = W + cfg->view_radius;
    }
    h->env_bytes = static_cast<int>(up16(H * h->Ws));
    h->pad_bytes = static_cast<int>(up16(cfg->view_radius * h->Ws + cfg->view_radius));
    h->V = 2 * cfg->view_radius + 1; h->obs_env = N * h->V * h->V * 3;
    if (h->env_bytes > 65535) { delete h; return fail(SSD_ERR_UNSUPPORTED, "map tile of %d bytes exceeds the 16-bit cell index range", h->env_bytes); }

    // static tables
    std::vector<uint32_t> color(ssd::kLutEntries, 0);
    std::vector<uint16_t> apple, waste, spawn;
    std::vector<uint8_t> init_grid(h->env_bytes, 0);
    const uint8_t apple_ch = cfg->kind == SSD_KIND_HARVEST ? 'A' : (cfg->kind == SSD_KIND_CLEANUP ? 'B' : 0);
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            const uint8_t ch = cfg->base_map[r * W + c];
            const uint16_t tile_cell = static_cast<uint16_t>(r * h->Ws + c);
            uint8_t g = ' ';  // reset_map + build_walls + custom_reset (map_env.py:560-564, harvest.py:57-60, cleanup.py:84-92)
            if (ch == '@') g = '@';
            else if (cfg->kind == SSD_KIND_HARVEST && ch == 'A') g = 'A';
            else if (cfg->kind == SSD_KIND_CLEANUP && (ch == 'H' || ch == 'R' || ch == 'S')) g = ch;
            init_grid[r * h->Ws + c] = ssd::ascii_to_cell(g);
            if (apple_ch && ch == apple_ch) apple.push_back(tile_cell);
            if (cfg->kind == SSD_KIND_CLEANUP && (ch == 'H' || ch == 'R')) waste.push_back(tile_cell);
        }
    if (cfg->kind == SSD_KIND_HARVEST)  // cached neighbourhood counts (ssd_internal.h: device cell encoding)
        for (int r = 1; r + 1 < H; ++r)
            for (int c = 1; c + 1 < W; ++c) {
                uint8_t& g = init_grid[r * h->Ws + c];
                if (g != ssd::CB(ssd::C_EMPTY) && g != ssd::CB(ssd::C_APPLE)) continue;
                int n = 0;
                for (int dr = -1; dr <= 1; ++dr)
                    for (int dc = -1; dc <= 1; ++dc)
                        n += (dr || dc) && (init_grid[(r + dr) * h->Ws + c + dc] & ssd::kCodeMask) == ssd::CB(ssd::C_APPLE);
                g |= static_cast<uint8_t>(n < 3 ? n : 3);
            }
    for (int s = 0; s < cfg->num_spawn_points; ++s)
        spawn.push_back(static_cast<uint16_t>(cfg->spawn_points[2 * s] << 8 | cfg->spawn_points[2 * s + 1]));
    {   // colour table by cell code: every code takes the colour of its ASCII character
        const char* chars = "0 @AHRSFC123456789";
        for (int code = 0; chars[code]; ++code) {
            const int i = static_cast<uint8_t>(chars[code]);
            for (int nb = 0; nb < 4; ++nb)  // the table is indexed by the grid byte: code << 2 | neighbour count
                color[4 * code + nb] = cfg->color_lut[3 * i] | cfg->color_lut[3 * i + 1] << 8 | cfg->color_lut[3 * i + 2] << 16;
        }
    }
    h->n_apple = static_cast<int>(apple.size()); h->n_waste = static_cast<int>(waste.size()); h->n_spawn = static_cast<int>(spawn.size());
    if (h->n_apple >= 0x1000) { delete h; return fail(SSD_ERR_UNSUPPORTED, "too many apple points"); }
    if (cfg->kind == SSD_KIND_CLEANUP && cfg->potential_waste_area < h->n_waste) {
        delete h;
        return fail(SSD_ERR_INVALID, "potential_waste_area %d smaller than the number of 'H'/'R' cells %d", cfg->potential_waste_area, h->n_waste);
    }
    std::vector<uint64_t> hthr(4, 0), athr, wthr;
    std::vector<double> hp(4, 0.0), ap, wp;
    if (cfg->harvest_spawn_prob)
        for (int i = 0; i < 4; ++i) { hp[i] = cfg->harvest_spawn_prob[i]; hthr[i] = threshold53(hp[i]); h->harvest_nz |= (hthr[i] != 0) << i; }
    const int area = cfg->kind == SSD_KIND_CLEANUP ? cfg->potential_waste_area : 0;
    for (int i = 0; i <= area; ++i) {
        const double pa = cfg->kind == SSD_KIND_CLEANUP ? cfg->cleanup_apple_prob[i] : 0.0;
        const double pw = cfg->kind == SSD_KIND_CLEANUP ? cfg->cleanup_waste_prob[i] : 0.0;
        ap.push_back(pa); athr.push_back(threshold53(pa)); wp.push_back(pw); wthr.push_back(threshold53(pw));
    }

    // Harvest orchard bitmaps (ssd_internal.h): two bitmaps over the apple points per env; the specialised kernel scans them
    // when all the words of a warp's envs fit one pass of 32 lanes
    std::vector<uint32_t> pt_mask((h->env_bytes + 31) / 32, 0);
    std::vector<uint16_t> pt_pre((h->env_bytes + 31) / 32, 0);
    if (cfg->kind == SSD_KIND_HARVEST && h->n_apple > 0) {
        h->nW = (h->n_apple + 31) / 32;
        h->orch_stride = (2 * h->nW + 3) & ~3;
        h->use_orch = (N <= 8 ? 4 : 2) * h->nW <= 32 && !ssd::knob("SSD_NO_ORCH");
        for (uint16_t c : apple) pt_mask[c >> 5] |= 1u << (c & 31);
        for (size_t w2 = 1; w2 < pt_mask.size(); ++w2) pt_pre[w2] = static_cast<uint16_t>(pt_pre[w2 - 1] + __builtin_popcount(pt_mask[w2 - 1]));
    }
    if (cfg->kind == SSD_KIND_CLEANUP) {  // word 0 of an env's record: its number of 'H' cells
        h->orch_stride = 4;
        for (uint8_t c : init_grid) h->init_waste += c == ssd::CB(ssd::C_WASTE);
    }

    // CTA shape: every warp owns 32/G envs; SSD_THREADS (64, 128 or 256) overrides the CTA size for tuning.
    const int smem_max = static_cast<int>(prop.sharedMemPerBlockOptin);
    auto env_int = [](const char* name, int dflt) { const char* v = ssd::knob(name); return v && *v ? atoi(v) : dflt; };
    int threads = env_int("SSD_THREADS", 128);
    if (threads != 32 && threads != 64 && threads != 128 && threads != 256) { delete h; return fail(SSD_ERR_INVALID, "SSD_THREADS must be 32, 64, 128 or 256"); }
    if (cfg->envs_per_cta != 0) {  // explicit request: must be a whole number of warps
        const int epw = N <= 8 ? 4 : 2;
        if (cfg->envs_per_cta % epw != 0 || cfg->envs_per_cta / epw > 8) { delete h; return fail(SSD_ERR_INVALID, "envs_per_cta must be a multiple of %d and at most %d", epw, 8 * epw); }
        threads = cfg->envs_per_cta / epw * 32;
    }
    h->threads = threads;
    h->L = make_layout(*h, threads);
    while (static_cast<int>(h->L.total) > smem_max && h->threads > 32) { h->threads >>= 1; h->L = make_layout(*h, h->threads); }
    if (static_cast<int>(h->L.total) > smem_max) {
        const unsigned need = h->L.total;
        delete h;
        return fail(SSD_ERR_UNSUPPORTED, "one warp's tile set needs %u bytes of shared memory (limit %d)", need, smem_max);
    }
    h->Lf = make_layout(*h, h->threads, true);
    const int E = (h->threads / 32) * (N <= 8 ? 4 : 2);
    h->E = E;
    h->B_pad = (h->B + E - 1) / E * E;

    // the specialised kernel fetches the tables with bulk copies of whole 16-byte units: pad them (ssd_step_fast.cu)
    while (apple.size() % 64) apple.push_back(static_cast<uint16_t>(h->Ws + 1));  // a harmless interior cell
    while (pt_mask.size() % 4) pt_mask.push_back(0);
    while (pt_pre.size() % 8) pt_pre.push_back(0);
    int bad = 0;
    bad |= h->upload(&h->d_apple, apple);
    bad |= h->upload(&h->d_waste, waste); bad |= h->upload(&h->d_spawn, spawn); bad |= h->upload(&h->d_color, color);
    bad |= h->upload(&h->d_hthr, hthr); bad |= h->upload(&h->d_hp, hp); bad |= h->upload(&h->d_athr, athr);
    bad |= h->upload(&h->d_ap, ap); bad |= h->upload(&h->d_wthr, wthr); bad |= h->upload(&h->d_wp, wp);
    bad |= h->upload(&h->d_init_grid, init_grid);
    bad |= h->alloc(&h->d_grid, static_cast<size_t>(h->B_pad) * h->env_bytes);
    bad |= h->alloc(&h->d_agents, static_cast<size_t>(h->B_pad) * N);
    bad |= h->alloc(&h->d_beam_buf, static_cast<size_t>(h->B_pad) * 64);
    bad |= h->alloc(&h->d_orch, static_cast<size_t>(h->B_pad) * std::max(h->orch_stride, 4));
    bad |= h->upload(&h->d_pt_mask, pt_mask); bad |= h->upload(&h->d_pt_pre, pt_pre);
    bad |= h->alloc(&h->d_stats, static_cast<size_t>(SSD_NUM_STATS));
    bad |= h->alloc(&h->chain.done, static_cast<size_t>(h->B_pad / 2 + 1));  // one word per task (4 or 2 envs)
    bad |= h->alloc(&h->d_bad, 1);
    if (ssd::knob("SSD_PROF")) { bad |= h->alloc(&h->d_prof, 48); if (!bad) cudaMemset(h->d_prof, 0, 48 * sizeof(unsigned long long)); }
    h->chain.cta_slots = prop.multiProcessorCount * 8;  // resident CTAs at the specialised kernel's shape (8 per SM)
    if (bad) { const char* m = cudaGetErrorString(cudaGetLastError()); ssd_destroy(h); return fail(SSD_ERR_CUDA, "device allocation failed: %s", m); }
    // initial state: post-reset_map grid, agents parked on the first spawn point (or cell 1,1)
    {
        std::vector<uint8_t> g(static_cast<size_t>(h->B_pad) * h->env_bytes);
        for (int b = 0; b < h->B_pad; ++b) memcpy(g.data() + static_cast<size_t>(b) * h->env_bytes, init_grid.data(), h->env_bytes);
        const uint32_t park = spawn.empty() ? (1u | 1u << 8) : ((spawn[0] >> 8) | (spawn[0] & 255u) << 8);
        std::vector<uint32_t> ag(static_cast<size_t>(h->B_pad) * N, park);
        cudaError_t e1 = cudaMemcpy(h->d_grid, g.data(), g.size(), cudaMemcpyHostToDevice);
        cudaError_t e2 = cudaMemcpy(h->d_agents, ag.data(), ag.size() * 4, cudaMemcpyHostToDevice);
        cudaError_t e3 = cudaMemset(h->d_stats, 0, SSD_NUM_STATS * sizeof(unsigned long long));
        cudaError_t e4 = cudaMemset(h->d_beam_buf, 0, static_cast<size_t>(h->B_pad) * 64);
        if (e4 == cudaSuccess) e4 = cudaMemset(h->chain.done, 0, (static_cast<size_t>(h->B_pad) / 2 + 1) * sizeof(uint32_t));
        // every apple point holds an apple after reset_map (harvest.py:57-60): both orchard bitmaps are empty
        if (e4 == cudaSuccess) e4 = cudaMemset(h->d_orch, 0, static_cast<size_t>(h->B_pad) * std::max(h->orch_stride, 4) * sizeof(uint32_t));
        if (e4 == cudaSuccess && cfg->kind == SSD_KIND_CLEANUP) {
            std::vector<uint32_t> o(static_cast<size_t>(h->B_pad) * 4, 0u);
            for (int b = 0; b < h->B_pad; ++b) o[static_cast<size_t>(b) * 4] = h->init_waste;
            e4 = cudaMemcpy(h->d_orch, o.data(), o.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
        }
        if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess) {
            ssd_destroy(h);
            return fail(SSD_ERR_CUDA, "state initialisation failed");
        }
    }
    *out = h;
    return SSD_OK;
}

int ssd_destroy(ssd_handle h) {
    if (!h) return SSD_OK;
    cudaSetDevice(h->cfg.device);
    if (h->d_prof) {  // profiling builds: per-phase cycles of the specialised kernel, per warp task
        unsigned long long p[48];
        if (cudaMemcpy(p, h->d_prof, sizeof p, cudaMemcpyDeviceToHost) == cudaSuccess) {
            fprintf(stderr, "SSD_PROF B=%d:", h->B);
            for (int i = 0; i < 16; ++i) if (p[16 + i]) fprintf(stderr, " t%d=%.0f", i, static_cast<double>(p[i]) / static_cast<double>(p[16 + i]));
            fprintf(stderr, "\nSSD_PROF max:");
            for (int i = 0; i < 16; ++i) if (p[16 + i]) fprintf(stderr, " t%d=%llu", i, p[32 + i]);
            fprintf(stderr, "\n");
        }
    }
    for (int i = 0; i < 2; ++i) if (h->hs[i]) cudaStreamDestroy(h->hs[i]);
    if (h->host_ev) cudaEventDestroy(h->host_ev);
    for (void* p : h->allocs) cudaFree(p);
    delete h;
    return SSD_OK;
}

int ssd_num_apple_points(ssd_handle h) { return h ? h->n_apple : -1; }
int ssd_num_waste_points(ssd_handle h) { return h ? h->n_waste : -1; }
int64_t ssd_obs_bytes_per_env(ssd_handle h) { return h ? h->obs_env : -1; }
int ssd_envs_per_cta(ssd_handle h) { return h ? h->E : -1; }
int64_t ssd_launch_count(ssd_handle h) { return h ? h->launches : -1; }

int64_t ssd_algorithmic_bytes_per_env_step(ssd_handle h) {
    if (!h) return -1;
    // grid read+write, agent state 8 B read+write, actions, rewards, observations.
    // (The reference's persistent waste order carries no information under either replay mode and
    // is not device state, so its 4*n_waste term is not counted.)
    const int64_t N = h->cfg.num_agents;
    return 2ll * h->HW + 16 * N + N + 4 * N + h->obs_env;
}

int ssd_seed(ssd_handle h, uint64_t seed, uint32_t t) {
    if (check_handle(h)) return SSD_ERR_INVALID;
    h->seed = seed; h->t = t;
    return SSD_OK;
}
int ssd_get_counter(ssd_handle h, uint32_t* t) {
    if (check_handle(h) || !t) return SSD_ERR_INVALID;
    *t = h->t;
    return SSD_OK;
}

int ssd_set_state(ssd_handle h, const uint8_t* grid, const int16_t* pos, const uint8_t* ori, void* stream) {
    if (check_handle(h)) return SSD_ERR_INVALID;
    h->chain.valid = false;
    if (!grid || !pos || !ori) return fail(SSD_ERR_INVALID, "grid, pos and ori are required");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int N = h->cfg.num_agents;
    const size_t ng = static_cast<size_t>(h->B) * h->HW, np = static_cast<size_t>(h->B) * N;
    if (!is_device_ptr(grid) || !is_device_ptr(pos) || !is_device_ptr(ori)) {
        if (!h->d_io_grid) {
            if (h->alloc(&h->d_io_grid, ng) || h->alloc(&h->d_io_pos, np * 2) || h->alloc(&h->d_io_ori, np))
                return fail(SSD_ERR_CUDA, "staging allocation failed");
        }
        CUDA_TRY(cudaMemcpyAsync(h->d_io_grid, grid, ng, cudaMemcpyDefault, st));
        CUDA_TRY(cudaMemcpyAsync(h->d_io_pos, pos, np * 2 * sizeof(int16_t), cudaMemcpyDefault, st));
        CUDA_TRY(cudaMemcpyAsync(h->d_io_ori, ori, np, cudaMemcpyDefault, st));
        grid = h->d_io_grid; pos = h->d_io_pos; ori = h->d_io_ori;
    }
    {   // positions index the shared-memory tiles unchecked inside the kernels: refuse anything outside the map up front
        int bad = 0;
        CUDA_TRY(cudaMemsetAsync(h->d_bad, 0, sizeof(int), st));
        CUDA_TRY(ssd::launch_check_positions(h->B, N, h->cfg.height, h->cfg.width, pos, h->d_bad, st));
        CUDA_TRY(cudaMemcpyAsync(&bad, h->d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        h->launches++;
        if (bad) return fail(SSD_ERR_INVALID, "%d agent position(s) outside the %dx%d map; the state was not changed", bad, h->cfg.height, h->cfg.width);
    }
    CUDA_TRY(ssd::launch_pack_state(h->cfg.kind, h->B, N, h->cfg.height, h->cfg.width, h->Ws, h->env_bytes, grid, pos, ori, h->d_grid, h->d_agents, st));
    h->launches++;
    if (h->orch_stride > 0) {
        CUDA_TRY(ssd::launch_build_orch(h->cfg.kind, h->B, h->n_apple, h->nW, h->orch_stride, h->harvest_nz, h->env_bytes, h->d_apple, h->d_grid, h->d_orch, st));
        h->launches++;
    }
    return SSD_OK;
}

int ssd_get_state(ssd_handle h, uint8_t* grid, int16_t* pos, uint8_t* ori, void* stream) {
    if (check_handle(h)) return SSD_ERR_INVALID;
    h->chain.valid = false;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int N = h->cfg.num_agents;
    const size_t ng = static_cast<size_t>(h->B) * h->HW, np = static_cast<size_t>(h->B) * N;
    const bool dev = (!grid || is_device_ptr(grid)) && (!pos || is_device_ptr(pos)) && (!ori || is_device_ptr(ori));
    if (dev) {
        CUDA_TRY(ssd::launch_unpack_state(h->B, N, h->cfg.height, h->cfg.width, h->Ws, h->env_bytes, h->d_grid, h->d_agents, grid, pos, ori, st));
        h->launches++;
        return SSD_OK;
    }
    if (!h->d_io_grid) {
        if (h->alloc(&h->d_io_grid, ng) || h->alloc(&h->d_io_pos, np * 2) || h->alloc(&h->d_io_ori, np))
            return fail(SSD_ERR_CUDA, "staging allocation failed");
    }
    CUDA_TRY(ssd::launch_unpack_state(h->B, N, h->cfg.height, h->cfg.width, h->Ws, h->env_bytes, h->d_grid, h->d_agents, h->d_io_grid, h->d_io_pos, h->d_io_ori, st));
    h->launches++;
    if (grid) CUDA_TRY(cudaMemcpyAsync(grid, h->d_io_grid, ng, cudaMemcpyDefault, st));
    if (pos) CUDA_TRY(cudaMemcpyAsync(pos, h->d_io_pos, np * 2 * sizeof(int16_t), cudaMemcpyDefault, st));
    if (ori) CUDA_TRY(cudaMemcpyAsync(ori, h->d_io_ori, np, cudaMemcpyDefault, st));
    CUDA_TRY(cudaStreamSynchronize(st));  // host destinations are complete on return
    return SSD_OK;
}

static int reset_impl(ssd_handle h, const uint8_t* mask, const int32_t* rows, int n_rows, uint8_t* obs_out, void* stream) {
    h->chain.valid = false;
    if (h->n_spawn == 0) return fail(SSD_ERR_INVALID, "the map has no 'P' spawn points");
    if (h->distinct_spawn < h->cfg.num_agents)  // map_env.py:661
        return fail(SSD_ERR_INVALID, "There are not enough spawn points! Check your map? (%d distinct, %d agents; map_env.py:661)", h->distinct_spawn, h->cfg.num_agents);
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ssd::ResetArgs r{};
    r.N = h->cfg.num_agents; r.n_spawn = h->n_spawn; r.env_bytes = h->env_bytes; r.env_end = h->B;
    r.key0 = static_cast<uint32_t>(h->seed); r.key1 = static_cast<uint32_t>(h->seed >> 32); r.t = h->t;
    r.env_id0 = h->cfg.env_id_offset;
    r.spawn_key = h->d_spawn; r.init_grid = h->d_init_grid; r.mask = mask; r.rows = rows; r.n_rows = n_rows; r.grid = h->d_grid; r.agents = h->d_agents;
    r.orch = h->orch_stride > 0 ? h->d_orch : nullptr; r.orch_stride = h->orch_stride; r.orch_word0 = h->init_waste;
    CUDA_TRY(ssd::launch_reset(r, st));
    ssd::StepArgs a;
    fill_args(h, a);
    a.phases = SSD_PHASE_SPAWN | (obs_out ? SSD_PHASE_RENDER : 0);  // custom_map_update + un-rotated render (map_env.py:230-248)
    a.rotate = 0; a.spawn_stream = ssd::STREAM_RSPAWN; a.mask = mask; a.rows = rows; a.n_rows = n_rows; a.obs = obs_out;
    CUDA_TRY(ssd::launch_step(a, h->threads, st));
    h->launches += 2;
    return SSD_OK;
}

int ssd_reset(ssd_handle h, const uint8_t* mask, uint8_t* obs_out, void* stream) {
    if (check_handle(h)) return SSD_ERR_INVALID;
    return reset_impl(h, mask, nullptr, 0, obs_out, stream);
}

int ssd_reset_rows(ssd_handle h, const int32_t* rows, int n_rows, uint8_t* obs_out, void* stream) {
    if (check_handle(h)) return SSD_ERR_INVALID;
    if (n_rows < 0 || (n_rows > 0 && !rows)) return fail(SSD_ERR_INVALID, "rows / n_rows");
    if (n_rows == 0) return SSD_OK;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (!is_device_ptr(rows)) {  // a host list: range-check it here and stage it on the device
        for (int i = 0; i < n_rows; ++i)
            if (rows[i] < 0 || rows[i] >= h->B) return fail(SSD_ERR_INVALID, "row %d outside 0..%d", rows[i], h->B - 1);
        if (n_rows > h->rows_cap) {
            int cap = n_rows < 256 ? 256 : n_rows;
            if (h->alloc(&h->d_rows, static_cast<size_t>(cap))) return fail(SSD_ERR_CUDA, "staging allocation failed");
            h->rows_cap = cap;
        }
        CUDA_TRY(cudaMemcpyAsync(h->d_rows, rows, static_cast<size_t>(n_rows) * sizeof(int32_t), cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
        rows = h->d_rows;
    }
    return reset_impl(h, nullptr, rows, n_rows, obs_out, stream);
}

int ssd_step_phases(ssd_handle h, int phases, const int8_t* actions, const uint8_t* action_order, const SsdTape* tape,
                    uint8_t* obs_out, int32_t* reward_out, void* stream) {
    if (check_handle(h)) return SSD_ERR_INVALID;
    if ((phases & ~SSD_PHASE_ALL) || phases == 0) return fail(SSD_ERR_INVALID, "bad phase mask %d", phases);
    if ((phases & (SSD_PHASE_MOVES | SSD_PHASE_BEAMS)) && !actions) return fail(SSD_ERR_INVALID, "actions are required");
    if (tape && h->cfg.kind == SSD_KIND_CLEANUP && h->n_waste > 0 && !tape->waste_order && (phases & SSD_PHASE_SPAWN))
        return fail(SSD_ERR_INVALID, "tape.waste_order is required for Cleanup");
    if (tape && (!tape->move_order || !tape->uniforms)) return fail(SSD_ERR_INVALID, "tape.move_order and tape.uniforms are required");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    ssd::StepArgs a;
    fill_args(h, a);
    a.phases = phases;
    a.use_beam_buf = phases != SSD_PHASE_ALL;
    a.rew_accumulate = phases != SSD_PHASE_ALL;
    a.actions = actions; a.order = action_order; a.obs = obs_out; a.rew = reward_out;
    if (tape) { a.tape_move = tape->move_order; a.tape_u = tape->uniforms; a.u_stride = tape->u_stride; a.tape_waste = tape->waste_order; a.n_draws_out = tape->n_draws_out; }
    CUDA_TRY(ssd::launch_step(a, h->threads, static_cast<cudaStream_t>(stream), &h->chain));
    h->launches++;
    if (phases & SSD_PHASE_SPAWN) h->t++;
    return SSD_OK;
}

int ssd_step(ssd_handle h, const int8_t* actions, const uint8_t* action_order, const SsdTape* tape, uint8_t* obs_out,
             int32_t* reward_out, void* stream) {
    return ssd_step_phases(h, SSD_PHASE_ALL, actions, action_order, tape, obs_out, reward_out, stream);
}

int ssd_rollout(ssd_handle h, int num_steps, const int8_t* actions, uint8_t* obs_ring, int ring_slots, int32_t* reward_out, void* stream) {
    if (check_handle(h)) return SSD_ERR_INVALID;
    if (num_steps < 0 || ring_slots < 1) return fail(SSD_ERR_INVALID, "num_steps must be >= 0 and ring_slots >= 1");
    if (!actions || !obs_ring || !reward_out) return fail(SSD_ERR_INVALID, "actions, obs_ring and reward_out are required");
    const size_t bn = static_cast<size_t>(h->B) * h->cfg.num_agents, obs_bytes = static_cast<size_t>(h->B) * h->obs_env;
    if (num_steps == 0) return SSD_OK;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    {   // A batch of less than half a wave of CTAs, all of it for the specialised kernel: ONE launch in which every warp runs all
        // the steps of its own envs.  Larger batches gain nothing over chained launches (they are throughput-bound).
        ssd::StepArgs a;
        fill_args(h, a);
        a.actions = actions; a.obs = obs_ring; a.rew = reward_out;
        a.n_steps = num_steps; a.ring_slots = ring_slots; a.step_stride = bn; a.obs_slot_stride = obs_bytes;
        if (num_steps > 1 && ssd::specialised_for_all(a, &h->chain, h->threads)) {
            h->chain.valid = false;
            CUDA_TRY(ssd::launch_step(a, h->threads, static_cast<cudaStream_t>(stream), &h->chain));
            h->chain.valid = false;
            h->launches++;
            h->t += static_cast<uint32_t>(num_steps);
            return SSD_OK;
        }
    }
    // otherwise one launch per step; every step's inputs exist before the first launch: exactly the precondition of chained steps
    const bool was_enabled = h->chain.enabled;
    h->chain.enabled = true;
    int rc = SSD_OK;
    for (int s = 0; s < num_steps && rc == SSD_OK; ++s)
        rc = ssd_step_phases(h, SSD_PHASE_ALL, actions + s * bn, nullptr, nullptr, obs_ring + (s % ring_slots) * obs_bytes,
                             reward_out + s * bn, stream);
    h->chain.enabled = was_enabled;
    h->chain.valid = false;
    return rc;
}

int ssd_get_beams(ssd_handle h, uint8_t* out, void* stream) {
    if (check_handle(h) || !out) return SSD_ERR_INVALID;
    h->chain.valid = false;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    std::vector<uint8_t> tmp(static_cast<size_t>(h->B) * 64);
    CUDA_TRY(cudaMemcpyAsync(tmp.data(), h->d_beam_buf, tmp.size(), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int b = 0; b < h->B; ++b)  // device cell bytes -> the reference's beam characters
        for (int k = 48; k < 64; ++k) {
            uint8_t& c = tmp[static_cast<size_t>(b) * 64 + k];
            c = c == ssd::CB(ssd::C_FIRE) ? 'F' : (c == ssd::CB(ssd::C_CLEAN) ? 'C' : 0);
        }
    CUDA_TRY(cudaMemcpy(out, tmp.data(), tmp.size(), cudaMemcpyDefault));
    return SSD_OK;
}

int ssd_render(ssd_handle h, int rotate, uint8_t* obs_out, void* stream) {
    if (check_handle(h)) return SSD_ERR_INVALID;
    h->chain.valid = false;
    if (!obs_out) return fail(SSD_ERR_INVALID, "obs_out is required");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    ssd::StepArgs a;
    fill_args(h, a);
    a.phases = SSD_PHASE_RENDER; a.rotate = rotate ? 1 : 0; a.obs = obs_out;
    CUDA_TRY(ssd::launch_step(a, h->threads, static_cast<cudaStream_t>(stream)));
    h->launches++;
    return SSD_OK;
}

int ssd_render_map(ssd_handle h, uint8_t* rgb_out, void* stream) {
    if (check_handle(h)) return SSD_ERR_INVALID;
    if (!rgb_out) return fail(SSD_ERR_INVALID, "rgb_out is required");
    h->chain.valid = false;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(ssd::launch_render_map(h->B, h->cfg.num_agents, h->cfg.height, h->cfg.width, h->Ws, h->env_bytes, h->d_grid, h->d_agents,
                                    h->d_color, rgb_out, static_cast<cudaStream_t>(stream)));
    h->launches++;
    return SSD_OK;
}

int ssd_step_host(ssd_handle h, const int8_t* actions_host, uint8_t* obs_host, int32_t* reward_host, void* stream) {
    if (check_handle(h)) return SSD_ERR_INVALID;
    h->chain.valid = false;
    if (!actions_host || !reward_host) return fail(SSD_ERR_INVALID, "actions_host and reward_host are required");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const int N = h->cfg.num_agents, B = h->B, E = h->E;
    if (!h->host_ready) {
        if (!h->hs[0]) CUDA_TRY(cudaStreamCreateWithFlags(&h->hs[0], cudaStreamNonBlocking));
        if (!h->hs[1]) CUDA_TRY(cudaStreamCreateWithFlags(&h->hs[1], cudaStreamNonBlocking));
        if (!h->host_ev) CUDA_TRY(cudaEventCreateWithFlags(&h->host_ev, cudaEventDisableTiming));
        if ((!h->d_act_stage && h->alloc(&h->d_act_stage, static_cast<size_t>(h->B_pad) * N)) ||
            (!h->d_rew_stage && h->alloc(&h->d_rew_stage, static_cast<size_t>(h->B_pad) * N)) ||
            (!h->d_obs_stage && h->alloc(&h->d_obs_stage, static_cast<size_t>(h->B_pad) * h->obs_env)))
            return fail(SSD_ERR_CUDA, "staging allocation failed");
        h->host_ready = true;
    }
    // The chunks run on two internal non-blocking streams: make them wait for whatever the caller already queued on
    // `stream` (an asynchronous ssd_reset / ssd_set_state / ssd_step writes the state these kernels read).
    CUDA_TRY(cudaEventRecord(h->host_ev, static_cast<cudaStream_t>(stream)));
    CUDA_TRY(cudaStreamWaitEvent(h->hs[0], h->host_ev, 0));
    CUDA_TRY(cudaStreamWaitEvent(h->hs[1], h->host_ev, 0));
    // chunks of whole CTAs, ~8 per step, alternating between two streams so that the D2H copy of
    // chunk i overlaps the kernel of chunk i+1
    int chunk = ((B + 7) / 8 + E - 1) / E * E;
    if (chunk < E) chunk = E;
    int k = 0;
    for (int b0 = 0; b0 < B; b0 += chunk, ++k) {
        const int b1 = b0 + chunk < B ? b0 + chunk : B;
        cudaStream_t st = h->hs[k & 1];
        const size_t na = static_cast<size_t>(b1 - b0) * N;
        CUDA_TRY(cudaMemcpyAsync(h->d_act_stage + static_cast<size_t>(b0) * N, actions_host + static_cast<size_t>(b0) * N, na, cudaMemcpyHostToDevice, st));
        ssd::StepArgs a;
        fill_args(h, a);
        a.env_begin = b0; a.env_end = b1;
        a.actions = h->d_act_stage; a.obs = obs_host ? h->d_obs_stage : nullptr; a.rew = h->d_rew_stage;
        CUDA_TRY(ssd::launch_step(a, h->threads, st));
        h->launches++;
        CUDA_TRY(cudaMemcpyAsync(reward_host + static_cast<size_t>(b0) * N, h->d_rew_stage + static_cast<size_t>(b0) * N, na * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        if (obs_host)
            CUDA_TRY(cudaMemcpyAsync(obs_host + static_cast<size_t>(b0) * h->obs_env, h->d_obs_stage + static_cast<size_t>(b0) * h->obs_env,
                                     static_cast<size_t>(b1 - b0) * h->obs_env, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(h->hs[0]));
    CUDA_TRY(cudaStreamSynchronize(h->hs[1]));
    h->t++;
    return SSD_OK;
}

int ssd_set_option(ssd_handle h, int option, int64_t value) {
    if (check_handle(h)) return SSD_ERR_INVALID;
    switch (option) {
        case SSD_OPT_CHAIN_STEPS:
            h->chain.enabled = value != 0;
            h->chain.valid = false;
            return SSD_OK;
        case SSD_OPT_GENERAL_KERNEL:
            h->chain.general_only = value != 0;
            h->chain.valid = false;
            return SSD_OK;
        default:
            return fail(SSD_ERR_INVALID, "unknown option %d", option);
    }
}

int ssd_stats(ssd_handle h, int64_t* out_host, void* stream) {
    if (check_handle(h) || !out_host) return SSD_ERR_INVALID;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long tmp[SSD_NUM_STATS];
    CUDA_TRY(cudaMemcpyAsync(tmp, h->d_stats, sizeof tmp, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int i = 0; i < SSD_NUM_STATS; ++i) out_host[i] = static_cast<int64_t>(tmp[i]);
    out_host[1] = out_host[2] - out_host[3] - 50 * out_host[4];  // reward_sum = apples - fires - 50 * hits
    return SSD_OK;
}

int ssd_philox_selftest(int device, const uint32_t ctr[4], const uint32_t key[2], uint32_t* out_host) {
    if (!ctr || !key || !out_host) return fail(SSD_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(device));
    uint32_t host[6] = {ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]};
    uint32_t* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 10 * sizeof(uint32_t)));
    cudaError_t e = cudaMemcpy(d, host, sizeof host, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = ssd::launch_philox_selftest(d, d + 6, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(out_host, d + 6, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(SSD_ERR_CUDA, "philox selftest: %s", cudaGetErrorString(e));
    return SSD_OK;
}

}  // extern "C"
#pragma GCC visibility pop
