// Small kernels around the step path: reset, ASCII <-> device cell bytes for state upload / download,
// full-map frames, the Philox self-test.
// Reference citations are relative to the reference root (social_dilemmas/envs/...).
#include "ssd_phases.cuh"

namespace ssd {

// ====================================================================== reset: setup_agents + reset_map
// map_env.py:214-229: spawn_point (:651-662) with the shuffle replaced by a (key, index) order --
// the reference takes the LAST free entry of the shuffled list = the free entry with the largest
// (key, index); spawn_rotation (:664-667) indexes ['LEFT','RIGHT','UP','DOWN'] with randint(4).
__global__ void __launch_bounds__(128) ssd_reset_kernel(const ResetArgs a) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (a.rows != nullptr) {  // ssd_reset_rows: one thread per listed env
        if (e >= a.n_rows) return;
        e = a.rows[e];
        if (e < 0) return;
    }
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
    if (a.orch != nullptr)  // every apple point holds an apple again: empty orchard bitmaps
        for (int i = 0; i < a.orch_stride; ++i) a.orch[static_cast<size_t>(e) * a.orch_stride + i] = i == 0 ? a.orch_word0 : 0u;
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
    if (i < static_cast<size_t>(B) * N) {  // positions were range-checked by check_positions_kernel
        const int r = pos_in[2 * i], c = pos_in[2 * i + 1];
        // An agent uploaded onto a wall cell (the adapters park a stand-in there when an env has no agents) is marked
        // "parked": the step kernels never let it act or paint it, so no ray ever starts outside the wall enclosure.
        const uint32_t parked = grid_in[(i / N) * H * W + r * W + c] == '@';
        agents[i] = (r & 255) | (c & 255) << 8 | static_cast<uint32_t>(ori_in[i] & 3) << 16 | parked << 24;
    }
}
// Orchard bitmaps of one env from its grid (ssd_internal.h, StepArgs::orch): one warp per env.
__global__ void build_orch_kernel(int kind, int B, int n_apple, int nW, int orch_stride, int harvest_nz, int env_bytes, const uint16_t* apple_cell,
                                  const uint8_t* grid, uint32_t* orch) {
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (e >= B) return;
    const uint8_t* g = grid + static_cast<size_t>(e) * env_bytes;
    if (kind == SSD_KIND_CLEANUP) {  // word 0: the number of 'H' cells
        int nh = 0;
        for (int i = lane; i < env_bytes; i += 32) nh += (g[i] & 0x7F) == CB(C_WASTE);
        nh = __reduce_add_sync(0xffffffffu, nh);
        if (lane == 0) orch[static_cast<size_t>(e) * orch_stride] = static_cast<uint32_t>(nh);
        return;
    }
    for (int w = 0; w < nW; ++w) {
        const int i = 32 * w + lane;
        const uint8_t c = i < n_apple ? g[apple_cell[i]] : 0;
        const bool emp = i < n_apple && (c & kCodeMask) == CB(C_EMPTY);
        const uint32_t me = __ballot_sync(0xffffffffu, emp), mn = __ballot_sync(0xffffffffu, emp && ((harvest_nz >> (c & 3)) & 1));
        if (lane == 0) { orch[static_cast<size_t>(e) * orch_stride + w] = me; orch[static_cast<size_t>(e) * orch_stride + nW + w] = mn; }
    }
}
__global__ void check_positions_kernel(int B, int N, int H, int W, const int16_t* pos_in, int* bad) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<size_t>(B) * N) return;
    const int r = pos_in[2 * i], c = pos_in[2 * i + 1];
    if (r < 0 || r >= H || c < 0 || c >= W) atomicAdd(bad, 1);
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
        if ((w & 255u) == r && ((w >> 8) & 255u) == c && !((w >> 24) & 1u)) cell = agent_cell(ag);
    }
    const uint32_t rgb = color[cell];
    out[3 * i] = rgb & 255; out[3 * i + 1] = (rgb >> 8) & 255; out[3 * i + 2] = (rgb >> 16) & 255;
}

__global__ void philox_selftest_kernel(const uint32_t* ck, uint32_t* out) {
    const uint4 v = philox4x32_10(ck[0], ck[1], ck[2], ck[3], ck[4], ck[5]);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}

// ====================================================================== launchers
cudaError_t launch_reset(const ResetArgs& a, cudaStream_t stream) {
    const int n = a.rows != nullptr ? a.n_rows : a.env_end;
    if (n <= 0) return cudaSuccess;
    ssd_reset_kernel<<<(n + 127) / 128, 128, 0, stream>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_build_orch(int kind, int B, int n_apple, int nW, int orch_stride, int harvest_nz, int env_bytes, const uint16_t* apple_cell,
                              const uint8_t* grid, uint32_t* orch, cudaStream_t stream) {
    if (B <= 0 || orch_stride <= 0) return cudaSuccess;
    build_orch_kernel<<<(B + 3) / 4, 128, 0, stream>>>(kind, B, n_apple, nW, orch_stride, harvest_nz, env_bytes, apple_cell, grid, orch);
    return cudaGetLastError();
}
cudaError_t launch_check_positions(int B, int N, int H, int W, const int16_t* pos_in, int* bad, cudaStream_t stream) {
    const size_t n = static_cast<size_t>(B) * N;
    check_positions_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(B, N, H, W, pos_in, bad);
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
