/* c_abi_demo.c - the C ABI of libhexb.so used from plain C (no Python, no torch): 11x11 self-play against the random opponent
 * with the fused random agent, N_STEPS env steps for N_GAMES games, then the episode statistics.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/c_abi_demo.c -o examples/c_abi_demo \
 *       -L hex_gym_env_b200 -lhexb -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN/../hex_gym_env_b200'
 *   ./examples/c_abi_demo [games] [steps] [seed]
 */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>

#include "hexb.h"

#define CHECK(x)                                                                      \
    do {                                                                              \
        int rc_ = (x);                                                                \
        if (rc_ != HEXB_OK) {                                                         \
            fprintf(stderr, "%s failed: %s (cuda %d)\n", #x, hexb_strerror(rc_), hexb_last_cuda_error()); \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

int main(int argc, char **argv) {
    const long long games = argc > 1 ? atoll(argv[1]) : 65536;
    const int steps = argc > 2 ? atoi(argv[2]) : 200;
    const unsigned long long seed = argc > 3 ? strtoull(argv[3], NULL, 10) : 0;
    hexb_config cfg = {0};
    cfg.board_size = 11;
    cfg.variant = HEXB_VARIANT_B;
    cfg.num_games = games;
    cfg.seed = seed;
    cfg.agent_mode = HEXB_AGENT_RANDOM;
    cfg.auto_reset = 1;
    const size_t nbytes = hexb_state_bytes(&cfg);
    if (!nbytes) { fprintf(stderr, "bad config\n"); return 1; }
    void *state = NULL, *obs = NULL, *mask = NULL, *reward = NULL, *done = NULL, *stats = NULL;
    const size_t cells = (size_t)games * 121;
    if (cudaMalloc(&state, nbytes) || cudaMalloc(&obs, cells) || cudaMalloc(&mask, cells) || cudaMalloc(&reward, games * 4) ||
        cudaMalloc(&done, games) || cudaMalloc(&stats, 64)) {
        fprintf(stderr, "cudaMalloc failed (no GPU?)\n");
        return 2;
    }
    hexb_env *env = NULL;
    CHECK(hexb_create(&cfg, state, nbytes, NULL, &env));
    CHECK(hexb_reset(env, NULL, NULL, (int8_t *)obs, (uint8_t *)mask, NULL));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, NULL);
    for (int t = 0; t < steps; ++t)
        CHECK(hexb_step(env, NULL, NULL, (int8_t *)obs, (uint8_t *)mask, (float *)reward, (uint8_t *)done, NULL, NULL, NULL));
    cudaEventRecord(e1, NULL);
    CHECK(hexb_stats(env, (int64_t *)stats, NULL));
    long long h[8];
    if (cudaMemcpy(h, stats, 64, cudaMemcpyDeviceToHost) != cudaSuccess) { fprintf(stderr, "copy failed\n"); return 3; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("stats %lld %lld %lld %lld %lld %lld %lld %lld\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    printf("%lld games x %d steps in %.3f ms: %.3e env-steps/s (ABI version %d.%d)\n", games, steps, ms,
           (double)games * steps / (ms * 1e-3), hexb_version() >> 16, hexb_version() & 0xffff);
    CHECK(hexb_destroy(env));
    cudaFree(state); cudaFree(obs); cudaFree(mask); cudaFree(reward); cudaFree(done); cudaFree(stats);
    return 0;
}
