// Standalone timer of the fused policy trunk (no Python): ms per call for M agents.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../include -o policy_timing policy_timing.cu
#include "../../sequential_social_dilemma_games_b200/csrc/ssd_policy.cu"

namespace ssd { int set_error(int code, const char* msg) { fprintf(stderr, "error %d: %s\n", code, msg); return code; } }

int main(int argc, char** argv) {
    const long long M = argc > 1 ? atoll(argv[1]) : 327680;
    std::vector<float> cw(162), cb(6), w1(1014 * 32), b1(32), w2(1024), b2(32);
    unsigned s = 1;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0f - 0.5f; };
    for (auto& x : cw) x = rnd() * 0.4f; for (auto& x : cb) x = rnd() * 0.1f; for (auto& x : w1) x = rnd() * 0.1f;
    for (auto& x : b1) x = rnd() * 0.1f; for (auto& x : w2) x = rnd() * 0.3f; for (auto& x : b2) x = rnd() * 0.1f;
    ssd_policy_t p;
    if (ssd_policy_create(7, 0, cw.data(), cb.data(), w1.data(), b1.data(), w2.data(), b2.data(), &p)) return 1;
    uint8_t* obs; float* out;
    cudaMalloc(&obs, M * 675 + 256); cudaMalloc(&out, M * 32 * 4);
    std::vector<uint8_t> h(M * 675); for (auto& x : h) { s = s * 1664525u + 1013904223u; x = s >> 24; }
    cudaMemcpy(obs, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) ssd_policy_features(p, obs, M, out, nullptr);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) ssd_policy_features(p, obs, M, out, nullptr);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("M=%lld  %.4f ms per call  (%s)\n", M, ms / 20, cudaGetErrorString(cudaGetLastError()));
#ifdef SSD_POLICY_TIMING
    unsigned long long c[3][8];
    cudaMemcpyFromSymbol(c, g_policy_cycles, sizeof c);
    const char* names[3][8] = {{"wait obs", "wait row_free", "build row", "publish previous row", "group end", "issue tmem st", "wait::st", "syncwarp"},
                               {"wait row_full", "wait rows+d1_free (o)", "issue e+d, commit", "wait c_full/w1", "issue o", "", "last fc2", ""},
                               {"wait d1_full", "tmem ld", "convert", "wait c_free", "store C", "group tail", "", ""}};
    const char* roles[3] = {"producer thread 0", "MMA thread", "drain thread 160 (warp 5)"};
    for (int r = 0; r < 3; ++r) {
        unsigned long long tot = 0; for (int i = 0; i < 8; ++i) tot += c[r][i];
        printf("%s: %llu cycles\n", roles[r], tot);
        for (int i = 0; i < 8; ++i) if (names[r][i][0]) printf("    %-20s %10llu  %5.1f %%\n", names[r][i], c[r][i], 100.0 * c[r][i] / tot);
    }
    {
        long long tl[3][32][4];
        cudaMemcpyFromSymbol(tl, g_policy_tl, sizeof tl);
        const long long base = tl[1][0][0];
        printf("timeline of CTA 0 (cycles relative to MMA step 40 block 1)\n");
        printf("idx | producer(row idx): built  free  stored  published(prev) | MMA(step idx): b1  commit  b2  b2done | drain(tile idx): step_seen  ld_done  c_full  iter_end\n");
        for (int i = 0; i < 20; ++i) {
            printf("%3d |", 40 + i);
            for (int r = 0; r < 3; ++r) { for (int e = 0; e < 4; ++e) printf(" %7lld", tl[r][i][e] - base); printf(" |"); }
            printf("\n");
        }
    }
#endif
    return 0;
}
