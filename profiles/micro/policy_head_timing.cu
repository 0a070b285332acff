// Standalone timer of the fused LSTM + heads kernel; with -DSSD_POLICY_TIMING also cycles per phase (thread 0 of CTA 0).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../include [-DSSD_POLICY_TIMING] -o policy_head_timing policy_head_timing.cu
#include "../../sequential_social_dilemma_games_b200/csrc/ssd_policy_head.cu"

namespace ssd { int set_error(int code, const char* msg) { fprintf(stderr, "error %d: %s\n", code, msg); return code; } }

int main(int argc, char** argv) {
    const long long M = argc > 1 ? atoll(argv[1]) : 327680;
    const int A = 8;
    unsigned s = 1;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0f - 0.5f; };
    std::vector<float> lw(32 * 512), lu(128 * 512), lb(512), gw(128 * A), gb(A), vw(128), vb(1);
    for (auto& x : lw) x = rnd() * 0.3f; for (auto& x : lu) x = rnd() * 0.2f; for (auto& x : lb) x = rnd() * 0.1f;
    for (auto& x : gw) x = rnd() * 0.3f; for (auto& x : gb) x = rnd() * 0.1f; for (auto& x : vw) x = rnd() * 0.3f; vb[0] = 0.1f;
    SsdPolicy pol; pol.device = 0; cudaDeviceGetAttribute(&pol.sms, cudaDevAttrMultiProcessorCount, 0);
    if (ssd_policy_set_head(&pol, 128, A, lw.data(), lu.data(), lb.data(), gw.data(), gb.data(), vw.data(), vb.data())) return 1;
    float *feat, *h, *c, *h2, *c2, *lg, *vl; int8_t* ac;
    const long long Mp = (M + 127) / 128 * 128;
    cudaMalloc(&feat, M * 32 * 4); cudaMalloc(&h, Mp * 512); cudaMalloc(&c, Mp * 512); cudaMalloc(&h2, Mp * 512); cudaMalloc(&c2, Mp * 512);
    cudaMalloc(&lg, M * A * 4); cudaMalloc(&vl, M * 4); cudaMalloc(&ac, M);
    cudaMemset(feat, 0, M * 128); cudaMemset(h, 0, Mp * 512); cudaMemset(c, 0, Mp * 512);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) ssd_policy_lstm_heads(&pol, feat, h, c, h2, c2, lg, vl, ac, M, 1, i, nullptr);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) ssd_policy_lstm_heads(&pol, feat, h, c, h2, c2, lg, vl, ac, M, 1, i, nullptr);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("M=%lld  %.4f ms per call  (%s)\n", M, ms / 20, cudaGetErrorString(cudaGetLastError()));
#ifdef SSD_POLICY_TIMING
    unsigned long long cy[8];
    cudaMemcpyFromSymbol(cy, g_head_cycles, sizeof cy);
    const char* names[5] = {"build A (HBM -> smem)", "gate MMAs", "cell update", "head MMAs", "logits / sample"};
    unsigned long long tot = 0; for (int i = 0; i < 5; ++i) tot += cy[i];
    for (int i = 0; i < 5; ++i) printf("  %-24s %10llu cycles  %5.1f %%\n", names[i], cy[i], 100.0 * cy[i] / tot);
#endif
    return 0;
}
