// hexb_hostpack.cpp - host half of hexb_step_host_packed (include/hexb.h): expands the 2-bit observation transport into the
// API's int8 obs[G,N,N] and uint8 mask[G,N*N] host arrays on a small pool of host threads.
//
// Why it exists: with HOST buffers the step is bound by the device->host path (obs + mask = 2*N*N bytes per game and step,
// 259 MB per 1 Mi-game 11x11 step at ~53 GB/s). Two bits per cell carry the same information (the mask is `cell empty`), so
// 8x fewer bytes cross PCIe and host cores write the API's arrays instead of the DMA engine. Whether that is faster depends on
// the host (profiles/r2*_e2e_packed.json has the A/B at 1 and 8 GPUs).
//
// The packed words of a step arrive in pieces (csrc/hexb_kernels.cu: host_step_enqueue / host_step_finish); the pool below works on
// them as they arrive: hexb_hostpack_begin / _publish / _finish.
//
// Plain C++ (g++), no CUDA: 16 cells per packed word, cell i of the flat [G*N*N] order in bits 2*(i%16).. of word i/16,
// code = obs byte & 3 (variant B: -1/0/+1 -> 3/0/1; variant A: BLACK 0, WHITE 1, EMPTY 2).
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

struct Job {
    const uint32_t *packed;
    long long first, count, cells;
    int variant;
    int8_t *obs;
    uint8_t *mask;
};

// ---- one worker's share: words [w0, w1) of the job
void expand_scalar(const Job &j, long long w0, long long w1) {
    // byte-wise LUT: 4 cells per packed byte; built once by whichever thread gets here first (a function-local static is
    // initialised exactly once under the language's own guard, so no worker ever reads a table another one is still writing)
    struct Lut {
        uint32_t obs[2][256], msk[2][256];
        Lut() {
            for (int v = 0; v < 2; ++v)
                for (int b = 0; b < 256; ++b) {
                    uint32_t o = 0, m = 0;
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t c = (b >> (2 * k)) & 3u;
                        const uint32_t ob = v == 1 ? (c == 3u ? 0xffu : c) : c;          // variant B: code 3 is -1
                        const uint32_t mk = v == 1 ? (c == 0u) : (c == 2u);              // legal == empty
                        o |= ob << (8 * k);
                        m |= mk << (8 * k);
                    }
                    obs[v][b] = o;
                    msk[v][b] = m;
                }
        }
    };
    static const Lut lut;
    const uint32_t *lo = lut.obs[j.variant ? 1 : 0], *lm = lut.msk[j.variant ? 1 : 0];
    for (long long w = w0; w < w1; ++w) {
        const uint32_t x = j.packed[w];
        const long long c0 = 16 * w;
        if (c0 + 16 <= j.cells) {
            uint32_t *po = reinterpret_cast<uint32_t *>(j.obs + c0), *pm = reinterpret_cast<uint32_t *>(j.mask + c0);
            for (int k = 0; k < 4; ++k) {
                const uint32_t b = (x >> (8 * k)) & 0xffu, o = lo[b], m = lm[b];
                memcpy(po + k, &o, 4);
                memcpy(pm + k, &m, 4);
            }
        } else {
            for (int k = 0; k < 16 && c0 + k < j.cells; ++k) {
                const uint32_t c = (x >> (2 * k)) & 3u;
                j.obs[c0 + k] = (int8_t)(j.variant ? (c == 3u ? -1 : (int)c) : (int)c);
                j.mask[c0 + k] = (uint8_t)(j.variant ? (c == 0u) : (c == 2u));
            }
        }
    }
}

#if defined(__x86_64__)
// 32 cells (two packed words) per iteration: bytes are spread with PSHUFB, the 2-bit field of each byte position is isolated
// with 16-bit shifts, and two 16-entry PSHUFB tables map code -> obs byte and code -> mask byte. Results leave the core with
// non-temporal stores when the destination is 32-byte aligned (nothing reads them back here).
__attribute__((target("avx2"))) void expand_avx2(const Job &j, long long w0, long long w1) {
    const __m256i spread = _mm256_setr_epi8(0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7);
    const __m256i three = _mm256_set1_epi8(3);
    const __m256i sel0 = _mm256_set1_epi32(0x000000ff), sel1 = _mm256_set1_epi32(0x0000ff00), sel2 = _mm256_set1_epi32(0x00ff0000);
    const __m256i tab_obs = j.variant ? _mm256_setr_epi8(0, 1, 2, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 2, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)
                                      : _mm256_setr_epi8(0, 1, 2, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 2, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i tab_msk = j.variant ? _mm256_setr_epi8(1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)
                                      : _mm256_setr_epi8(0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    long long w = w0;
    if ((w & 1) && w < w1) { expand_scalar(j, w, w + 1); ++w; }   // keep the 32-cell groups 32-byte aligned relative to cell 0
    const bool nt = ((((uintptr_t)j.obs) | ((uintptr_t)j.mask)) & 31) == 0;
    for (; w + 2 <= w1 && 16 * (w + 2) <= j.cells; w += 2) {
        uint64_t x;
        memcpy(&x, j.packed + w, 8);
        const __m128i lo = _mm_cvtsi64_si128((long long)x);
        // both 128-bit lanes get the 8 packed bytes; lane 0 uses bytes 0-3, lane 1 bytes 4-7
        const __m256i v = _mm256_shuffle_epi8(_mm256_broadcastsi128_si256(lo), spread);
        const __m256i t0 = _mm256_and_si256(v, three);
        const __m256i t1 = _mm256_and_si256(_mm256_srli_epi16(v, 2), three);
        const __m256i t2 = _mm256_and_si256(_mm256_srli_epi16(v, 4), three);
        const __m256i t3 = _mm256_and_si256(_mm256_srli_epi16(v, 6), three);
        // byte position p of every 4-byte group takes t_p
        const __m256i code = _mm256_or_si256(_mm256_or_si256(_mm256_and_si256(t0, sel0), _mm256_and_si256(t1, sel1)),
                                             _mm256_or_si256(_mm256_and_si256(t2, sel2), _mm256_andnot_si256(_mm256_or_si256(_mm256_or_si256(sel0, sel1), sel2), t3)));
        const __m256i o = _mm256_shuffle_epi8(tab_obs, code), m = _mm256_shuffle_epi8(tab_msk, code);
        const long long c0 = 16 * w;
        if (nt) {
            _mm256_stream_si256(reinterpret_cast<__m256i *>(j.obs + c0), o);
            _mm256_stream_si256(reinterpret_cast<__m256i *>(j.mask + c0), m);
        } else {
            _mm256_storeu_si256(reinterpret_cast<__m256i *>(j.obs + c0), o);
            _mm256_storeu_si256(reinterpret_cast<__m256i *>(j.mask + c0), m);
        }
    }
    if (nt) _mm_sfence();
    if (w < w1) expand_scalar(j, w, w1);
}

// The same in 512-bit registers: 64 cells (four packed words) per iteration, and each result is ONE full cache line written by one
// non-temporal store (two 32-byte halves of a line go through a write-combining buffer that can be flushed half full).
__attribute__((target("avx512f,avx512bw"))) void expand_avx512(const Job &j, long long w0, long long w1) {
    long long w = w0;
    while ((w & 3) && w < w1) { expand_scalar(j, w, w + 1); ++w; }   // 64-cell groups are 64-byte aligned relative to cell 0
    const bool nt = ((((uintptr_t)j.obs) | ((uintptr_t)j.mask)) & 63) == 0;
    if (nt) {
        // lane q of the 512-bit register takes packed bytes 4q .. 4q+3, each four times
        const __m512i spread = _mm512_set_epi8(15, 15, 15, 15, 14, 14, 14, 14, 13, 13, 13, 13, 12, 12, 12, 12, 11, 11, 11, 11, 10, 10, 10, 10, 9, 9, 9, 9,
                                               8, 8, 8, 8, 7, 7, 7, 7, 6, 6, 6, 6, 5, 5, 5, 5, 4, 4, 4, 4, 3, 3, 3, 3, 2, 2, 2, 2, 1, 1, 1, 1, 0, 0, 0, 0);
        const __m512i three = _mm512_set1_epi8(3);
        const __m512i sel0 = _mm512_set1_epi32(0x000000ff), sel1 = _mm512_set1_epi32(0x0000ff00), sel2 = _mm512_set1_epi32(0x00ff0000),
                      sel3 = _mm512_set1_epi32((int)0xff000000u);
        const __m128i t_obs = j.variant ? _mm_setr_epi8(0, 1, 2, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0) : _mm_setr_epi8(0, 1, 2, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
        const __m128i t_msk = j.variant ? _mm_setr_epi8(1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0) : _mm_setr_epi8(0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
        const __m512i tab_obs = _mm512_broadcast_i32x4(t_obs), tab_msk = _mm512_broadcast_i32x4(t_msk);
        for (; w + 4 <= w1 && 16 * (w + 4) <= j.cells; w += 4) {
            const __m128i x = _mm_loadu_si128(reinterpret_cast<const __m128i *>(j.packed + w));
            const __m512i v = _mm512_shuffle_epi8(_mm512_broadcast_i32x4(x), spread);
            const __m512i t0 = _mm512_and_si512(v, three);
            const __m512i t1 = _mm512_and_si512(_mm512_srli_epi16(v, 2), three);
            const __m512i t2 = _mm512_and_si512(_mm512_srli_epi16(v, 4), three);
            const __m512i t3 = _mm512_and_si512(_mm512_srli_epi16(v, 6), three);
            // byte position p of every 4-byte group takes t_p
            const __m512i code = _mm512_or_si512(_mm512_or_si512(_mm512_and_si512(t0, sel0), _mm512_and_si512(t1, sel1)),
                                                 _mm512_or_si512(_mm512_and_si512(t2, sel2), _mm512_and_si512(t3, sel3)));
            const long long c0 = 16 * w;
            _mm512_stream_si512(reinterpret_cast<__m512i *>(j.obs + c0), _mm512_shuffle_epi8(tab_obs, code));
            _mm512_stream_si512(reinterpret_cast<__m512i *>(j.mask + c0), _mm512_shuffle_epi8(tab_msk, code));
        }
        _mm_sfence();
    }
    if (w < w1) expand_avx2(j, w, w1);
}
#endif

void expand_range(const Job &j, long long w0, long long w1) {
#if defined(__x86_64__)
    static const int level = [] {   // HEXB_HOST_SIMD = 0 (scalar) / 2 (AVX2) / 512 caps what the CPU offers (experiments)
        const char *e = getenv("HEXB_HOST_SIMD");
        const int cap = e ? atoi(e) : 512;
        if (cap >= 512 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw")) return 512;
        if (cap >= 2 && __builtin_cpu_supports("avx2")) return 2;
        return 0;
    }();
    if (level == 512) { expand_avx512(j, w0, w1); return; }
    if (level == 2) { expand_avx2(j, w0, w1); return; }
#endif
    expand_scalar(j, w0, w1);
}

// ---- a persistent pool. One job = the packed words of one host-buffer step, cut into blocks that the threads take from a shared
// counter (so a thread that is late - it slept, or its core was taken - simply takes fewer blocks), and the words of the job become
// AVAILABLE piece by piece while the threads are already at work: the submitting thread waits for each piece's device->host copy
// and publishes how far the words have arrived; a thread whose block has not arrived yet spins for it. There is one join, at the
// end of the step, instead of one per piece.
struct Pool {
    pthread_mutex_t mu;
    pthread_cond_t cv_work, cv_done;
    int nthreads, started;
    unsigned long long generation;
    int remaining;
    Job job;
    long long blk, first_block, nblocks;   // block size in words; absolute index of the job's first block; number of blocks
    long long next;                        // next block to hand out (atomic)
    long long avail;                       // absolute word index up to which the packed words have arrived (atomic)
    int abort;                             // the submitter gave up (a failed copy): threads drop the blocks they wait for
    pthread_t th[64];
};
Pool g_pool = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, 0, 0, 0, 0, {}, 0, 0, 0, 0, 0, 0, {}};
pthread_mutex_t g_submit = PTHREAD_MUTEX_INITIALIZER;   // one job at a time (several handles / host threads may call in)
const long long kBlockWords = 4096;   // 65,536 cells = 128 KB of obs + mask per block; a multiple of 4 (the AVX-512 loop's alignment)

// The pieces of one host-buffer step reach the pool a few hundred microseconds apart, and a thread that sleeps on a condition
// variable takes tens of microseconds to run again (more inside a VM). So between jobs the workers (and the submitting thread,
// waiting for them) first SPIN on the shared counters for a bounded time and only then sleep.
double now_us() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e6 * (double)ts.tv_sec + 1e-3 * (double)ts.tv_nsec;
}
inline void cpu_relax() {
#if defined(__x86_64__)
    _mm_pause();
#endif
}
// HEXB_HOST_SPIN_US: how long a pool thread spins for the next job before it sleeps (default 400; 0 = always sleep, e.g. when a
// host policy needs the cores between steps)
double spin_us() {
    static const double v = [] {
        const char *e = getenv("HEXB_HOST_SPIN_US");
        const double x = e ? atof(e) : 400.0;
        return x < 0.0 ? 0.0 : (x > 1e5 ? 1e5 : x);
    }();
    return v;
}

// one thread's part of the running job: blocks from the shared counter, each as soon as its words have arrived
void run_blocks() {
    const Job j = g_pool.job;
    const long long blk = g_pool.blk, b0 = g_pool.first_block, nb = g_pool.nblocks, end = j.first + j.count;
    for (;;) {
        const long long b = __atomic_fetch_add(&g_pool.next, 1, __ATOMIC_RELAXED);
        if (b >= nb) break;
        long long lo = (b0 + b) * blk, hi = lo + blk;
        if (lo < j.first) lo = j.first;
        if (hi > end) hi = end;
        while (__atomic_load_n(&g_pool.avail, __ATOMIC_ACQUIRE) < hi) {
            if (__atomic_load_n(&g_pool.abort, __ATOMIC_RELAXED)) return;
            for (int i = 0; i < 16; ++i) cpu_relax();
        }
        if (hi > lo) expand_range(j, lo, hi);
    }
}

void *worker(void *) {
    unsigned long long seen = 0;
    for (;;) {
        for (const double t_end = now_us() + spin_us(); __atomic_load_n(&g_pool.generation, __ATOMIC_ACQUIRE) == seen && now_us() < t_end;)
            for (int i = 0; i < 64; ++i) cpu_relax();
        pthread_mutex_lock(&g_pool.mu);
        while (g_pool.generation == seen) pthread_cond_wait(&g_pool.cv_work, &g_pool.mu);
        seen = g_pool.generation;
        pthread_mutex_unlock(&g_pool.mu);
        run_blocks();
        pthread_mutex_lock(&g_pool.mu);
        const int left = __atomic_sub_fetch(&g_pool.remaining, 1, __ATOMIC_ACQ_REL);
        if (left == 0) pthread_cond_signal(&g_pool.cv_done);
        pthread_mutex_unlock(&g_pool.mu);
    }
    return nullptr;
}

int pool_threads() {
    int n = 0;
    const char *e = getenv("HEXB_HOST_THREADS");
    if (e) n = atoi(e);
    if (n <= 0) {
        cpu_set_t set;
        CPU_ZERO(&set);
        n = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : 1;
    }
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    return n;
}

void after_fork_in_child() {   // threads do not survive fork(): the child starts its own pool on first use
    g_pool.started = 0;
    g_pool.nthreads = 0;
    g_pool.generation = 0;
    g_pool.remaining = 0;
    pthread_mutex_init(&g_pool.mu, nullptr);
    pthread_cond_init(&g_pool.cv_work, nullptr);
    pthread_cond_init(&g_pool.cv_done, nullptr);
    pthread_mutex_init(&g_submit, nullptr);
}

void pool_start() {   // called with g_submit held
    if (g_pool.started) return;
    static int atfork_set = 0;
    if (!atfork_set) {
        pthread_atfork(nullptr, nullptr, after_fork_in_child);
        atfork_set = 1;
    }
    g_pool.nthreads = pool_threads();
    // the calling thread works too (after its last publish); the others are spawned once and live for the process
    for (int k = 1; k < g_pool.nthreads; ++k) {
        if (pthread_create(&g_pool.th[k], nullptr, worker, nullptr) != 0) {
            g_pool.nthreads = k;
            break;
        }
        pthread_detach(g_pool.th[k]);
    }
    g_pool.started = 1;
}

}  // namespace

// Size of the pool (as it is, or as it will be: the threads are only created by the first expansion)
extern "C" __attribute__((visibility("hidden"))) int hexb_hostpack_threads(void) {
    pthread_mutex_lock(&g_submit);
    const int n = g_pool.started ? g_pool.nthreads : pool_threads();
    pthread_mutex_unlock(&g_submit);
    return n;
}

// A job in three calls, all from ONE thread: _begin hands the whole range [first_word, first_word + n_words) to the pool (nothing
// of it is available yet) and keeps the pool locked for this caller; _publish(w) says that the packed words below absolute index w
// have arrived; _finish makes the caller work along, waits for the pool and unlocks it. abort != 0: the words will not arrive
// (a failed copy) - threads drop what they wait for; the outputs are then incomplete and the caller reports the error.
extern "C" __attribute__((visibility("hidden"))) void hexb_hostpack_begin(const uint32_t *packed, long long first_word, long long n_words,
                                                                           long long n_cells, int variant, int8_t *obs, uint8_t *mask) {
    pthread_mutex_lock(&g_submit);
    pool_start();
    pthread_mutex_lock(&g_pool.mu);
    g_pool.job = Job{packed, first_word, n_words, n_cells, variant, obs, mask};
    g_pool.blk = kBlockWords;
    g_pool.first_block = first_word / kBlockWords;
    g_pool.nblocks = n_words > 0 ? (first_word + n_words + kBlockWords - 1) / kBlockWords - g_pool.first_block : 0;
    __atomic_store_n(&g_pool.next, 0, __ATOMIC_RELAXED);
    __atomic_store_n(&g_pool.avail, first_word, __ATOMIC_RELAXED);
    __atomic_store_n(&g_pool.abort, 0, __ATOMIC_RELAXED);
    __atomic_store_n(&g_pool.remaining, g_pool.nthreads - 1, __ATOMIC_RELEASE);
    if (g_pool.nthreads > 1) {
        __atomic_store_n(&g_pool.generation, g_pool.generation + 1, __ATOMIC_RELEASE);   // spinning workers see this without the lock
        pthread_cond_broadcast(&g_pool.cv_work);
    }
    pthread_mutex_unlock(&g_pool.mu);
}

extern "C" __attribute__((visibility("hidden"))) void hexb_hostpack_publish(long long words_arrived_abs) {
    __atomic_store_n(&g_pool.avail, words_arrived_abs, __ATOMIC_RELEASE);
}

extern "C" __attribute__((visibility("hidden"))) void hexb_hostpack_finish(int abort) {
    if (abort) __atomic_store_n(&g_pool.abort, 1, __ATOMIC_RELEASE);
    run_blocks();
    if (g_pool.nthreads > 1) {
        for (const double t_end = now_us() + spin_us(); __atomic_load_n(&g_pool.remaining, __ATOMIC_ACQUIRE) != 0 && now_us() < t_end;)
            for (int i = 0; i < 64; ++i) cpu_relax();
        pthread_mutex_lock(&g_pool.mu);
        while (g_pool.remaining != 0) pthread_cond_wait(&g_pool.cv_done, &g_pool.mu);
        pthread_mutex_unlock(&g_pool.mu);
    }
    pthread_mutex_unlock(&g_submit);
}

// Expand packed words [first_word, first_word + n_words) into obs / mask (whole arrays' base pointers; n_cells = G*N*N), all of
// them already in host memory.
extern "C" __attribute__((visibility("hidden"))) void hexb_hostpack_expand(const uint32_t *packed, long long first_word, long long n_words,
                                                                            long long n_cells, int variant, int8_t *obs, uint8_t *mask) {
    hexb_hostpack_begin(packed, first_word, n_words, n_cells, variant, obs, mask);
    hexb_hostpack_publish(first_word + n_words);
    hexb_hostpack_finish(0);
}
