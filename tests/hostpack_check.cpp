// Host-only check of csrc/hexb_hostpack.cpp (the expansion of the 2-bit transport of hexb_step_host): every SIMD level and
// ragged word ranges against a plain per-cell loop. Built and run by tests/test_hostpack_cpu.py; test infrastructure.
#include "../hex_gym_env_b200/csrc/hexb_hostpack.cpp"
#include <stdio.h>
int main() {
    int bad = 0;
    for (int variant = 0; variant < 2; ++variant) {
        const long long cells = 7 * 361 * 333 + 5, words = (cells + 15) / 16;
        uint32_t *packed; int8_t *obs, *ref_o; uint8_t *mask, *ref_m;
        if (posix_memalign((void **)&packed, 64, words * 4 + 64) || posix_memalign((void **)&obs, 64, cells + 128) ||
            posix_memalign((void **)&mask, 64, cells + 128) || posix_memalign((void **)&ref_o, 64, cells + 128) ||
            posix_memalign((void **)&ref_m, 64, cells + 128)) return 2;
        unsigned long long s = 88172645463325252ull + variant;
        for (long long i = 0; i < words; ++i) {
            uint32_t w = 0;
            for (int k = 0; k < 16; ++k) {
                s ^= s << 13; s ^= s >> 7; s ^= s << 17;
                uint32_t c = (s >> 20) % 3;
                if (variant && c == 2) c = 3;   // variant B codes: 0 empty, 1 opponent, 3 own (-1); variant A: 0, 1, 2
                w |= c << (2 * k);
            }
            packed[i] = w;
        }
        for (long long c = 0; c < cells; ++c) {   // the definition, cell by cell
            const uint32_t code = (packed[c / 16] >> (2 * (c % 16))) & 3u;
            ref_o[c] = (int8_t)(variant ? (code == 3u ? -1 : (int)code) : (int)code);
            ref_m[c] = (uint8_t)(variant ? (code == 0u) : (code == 2u));
        }
        const long long ranges[][2] = {{0, words}, {1, words - 1}, {3, 1001}, {4, 64}, {5, 1}, {words - 7, 7}, {0, 3}};
        for (int off = 0; off < 2; ++off)         // aligned and unaligned destinations
            for (auto &r : ranges) {
                memset(obs, 7, cells + 128); memset(mask, 7, cells + 128);
                hexb_hostpack_expand(packed, r[0], r[1], cells, variant, obs + off, mask + off);
                long long c0 = 16 * r[0], c1 = 16 * (r[0] + r[1]);
                if (c1 > cells) c1 = cells;
                int ok = memcmp(obs + off + c0, ref_o + c0, c1 - c0) == 0 && memcmp(mask + off + c0, ref_m + c0, c1 - c0) == 0;
                for (long long c = 0; c < cells + 64 && ok; ++c)   // nothing outside the range is touched
                    if ((c < c0 || c >= c1) && (obs[off + c] != 7 || mask[off + c] != 7)) ok = 0;
                if (!ok) { printf("MISMATCH variant %d off %d range %lld+%lld\n", variant, off, r[0], r[1]); bad = 1; }
            }
        // the streaming form hexb_step_host uses: the pool gets the whole range, the words are published piece by piece
        for (int rep = 0; rep < 3; ++rep) {
            memset(obs, 7, cells + 128); memset(mask, 7, cells + 128);
            const long long first = rep == 0 ? 0 : 242 * rep, n = words - first;
            hexb_hostpack_begin(packed, first, n, cells, variant, obs, mask);
            const long long cuts[] = {first + 1, first + 17, first + n / 16, first + n / 3, first + n / 3, first + n - 1, first + n};
            for (long long c : cuts) {
                for (volatile int spin = 0; spin < 20000 * rep; ++spin) {}   // let the pool run ahead of the arrivals
                hexb_hostpack_publish(c);
            }
            hexb_hostpack_finish(0);
            const long long c0 = 16 * first;
            int ok = memcmp(obs + c0, ref_o + c0, cells - c0) == 0 && memcmp(mask + c0, ref_m + c0, cells - c0) == 0;
            for (long long c = 0; c < c0 && ok; ++c) ok = obs[c] == 7 && mask[c] == 7;
            for (long long c = cells; c < cells + 64 && ok; ++c) ok = obs[c] == 7 && mask[c] == 7;
            if (!ok) { printf("MISMATCH streaming variant %d rep %d\n", variant, rep); bad = 1; }
        }
        {   // an aborted job must return (threads drop the blocks they wait for) and leave the pool usable
            hexb_hostpack_begin(packed, 0, words, cells, variant, obs, mask);
            hexb_hostpack_publish(5000);
            hexb_hostpack_finish(1);
            memset(obs, 7, cells + 128); memset(mask, 7, cells + 128);
            hexb_hostpack_expand(packed, 0, words, cells, variant, obs, mask);
            if (memcmp(obs, ref_o, cells) || memcmp(mask, ref_m, cells)) { printf("MISMATCH after abort variant %d\n", variant); bad = 1; }
        }
        free(packed); free(obs); free(mask); free(ref_o); free(ref_m);
    }
    printf(bad ? "FAILED\n" : "hostpack ok (%d threads)\n", hexb_hostpack_threads());
    return bad;
}
