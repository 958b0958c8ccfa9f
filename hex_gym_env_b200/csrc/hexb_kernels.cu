// hexb_kernels.cu - sm_100a kernels and the C ABI (include/hexb.h) of the batched Hex simulator.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC (hex_gym_env_b200/_native.py: build).
//
// Kernel inventory:
//   K1/K2/K3  hexb_step_kernel<N>  one warp per chunk of 32 games, HEXB_WARPS_PER_CTA (default 1) warps per CTA;
//             MODE_RESET / MODE_STEP / MODE_PLY / MODE_HALF, and T steps per launch for hexb_rollout
//   K4        hexb_sample_kernel   standalone k-th-empty-cell sampler
//   K5        hexb_encode_kernel   standalone observation + mask encoder (either view)
//   K6        hexb_export_kernel / hexb_import_kernel   reference-layout dump / preset boards
//   K7        statistics: warp __reduce_add_sync + one atomic per warp and counter (striped); hexb_stats sums the stripes
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/hexb.h"
#include "hexb_views.cuh"

using namespace hexb;

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (spin > (1u << 26)) __trap();  // never hang the GPU: a lost copy becomes a CUDA error
    }
}
// global -> shared bulk copy (TMA 1-D), completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// L2 eviction policies for the chunk copies: a fixed part of the state (the first keep_chunks chunks) is marked evict_last and
// stays in the 126 MB L2 from one step to the next, the rest is marked evict_first so that it does not push that part out.
// (Marking ALL of a state larger than L2 evict_last just recreates LRU thrashing: every line is evicted before its reuse.)
__device__ __forceinline__ uint64_t l2_policy(bool keep) {
    uint64_t p;
    if (keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void *dst, const void *src_smem, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(smem_u32(src_smem)),
                 "r"(bytes), "l"(policy)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ step kernel
// One warp = one chunk of 32 games; a CTA is HEXB_WARPS_PER_CTA independent warps (default 1; no __syncthreads anywhere). Per warp:
//   lane 0 starts ONE bulk asynchronous copy (cp.async.bulk = the 1-D TMA path, completion on the warp's own mbarrier)
//   of the chunk (32 games' label bytes + record words, one contiguous block) into shared memory; meanwhile every lane
//   fetches its game's meta / stream-position words and runs the Philox rounds of the step's two draws; then the
//   thread-per-game plies, the rare finished-game rows, the elementwise obs/mask encode with 16-byte coalesced stores, the
//   warp-per-game relabel sweeps, and one bulk copy of the chunk back to global memory.
#ifndef HEXB_WARPS_PER_CTA
#define HEXB_WARPS_PER_CTA 1   // measured (r1h): 1 warp per CTA 90.0 us, 2: 94.0 us, 4: 94.7 us per 1 Mi-game step (finer-grained tail)
#endif
constexpr int kWarpsPerCta = HEXB_WARPS_PER_CTA;   // Gpad is a multiple of kTile = 128 games, so 1, 2 and 4 all divide it
constexpr int kCtaThreads = kWarpsPerCta * kWarp;
// resident CTAs per SM the register allocation should allow (the hardware holds at most 32 CTAs per SM, i.e. 32 warps with one
// warp per CTA; the step kernel uses 56 registers, so registers are not the limit), for large boards whatever the shared-memory
// footprint of the chunk permits
constexpr int min_ctas(int n) {
    const int smem = kWarpsPerCta * 32 * (n * n + 4 * ((n * n + 31) / 32 + 2)) + 64;
    const int by_smem = 220 * 1024 / smem;
    const int want = 10 * 4 / kWarpsPerCta;   // 40 warps per SM (48 registers): measured equal to 48 warps, and no spills
    const int cap = 8 * 4 / kWarpsPerCta;
    return n <= 12 ? (want > 32 ? 32 : want) : (by_smem < 1 ? 1 : (by_smem > cap ? cap : by_smem));
}

template <int N>
struct SmemLayout {
    static constexpr int CHUNK = Geo<N>::CHUNK_STATE;    // labels + records of 32 games
    static constexpr int BAR = kWarpsPerCta * CHUNK;   // multiple of 16
    static constexpr int BYTES = BAR + kWarpsPerCta * 8;
};

__device__ __forceinline__ void st_hint(uint4 *p, const uint4 &v, uint64_t policy) {
    asm volatile("st.global.cs.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy)
                 : "memory");
}
template <int N, int VARIANT, bool HINT>
__device__ __forceinline__ void encode_loop(const uint4 *src, uint4 *po, uint4 *pm, int lane, uint32_t one, uint64_t pol) {
#pragma unroll 4
    for (int i = lane; i < Chunk<N>::VECS; i += kWarp) {
        const uint4 x = src[i];
        Vec4 in = {x.x, x.y, x.z, x.w}, o, m;
        encode_vec_v<VARIANT>(in, one, o, m);
        if (HINT) {
            st_hint(po + i, make_uint4(o.x, o.y, o.z, o.w), pol);
            st_hint(pm + i, make_uint4(m.x, m.y, m.z, m.w), pol);
        } else {
            __stcs(po + i, make_uint4(o.x, o.y, o.z, o.w));
            __stcs(pm + i, make_uint4(m.x, m.y, m.z, m.w));
        }
    }
}
template <int N>
__device__ __forceinline__ void encode_chunk(const uint8_t *chunk, const Params &P, long long g0, int t, int lane) {
    constexpr int C = Geo<N>::C;
    const long long out0 = (g0 + (long long)t * P.G) * C;   // byte offset of the chunk in obs / mask (row t of [T,G,C])
    const long long limit = ((long long)t + 1) * P.G * C;   // bytes of that row that exist in the caller's buffers
    const uint4 *src = reinterpret_cast<const uint4 *>(chunk);
    uint8_t *obs = reinterpret_cast<uint8_t *>(P.obs);
    uint8_t *msk = P.mask;
    const bool vec_ok = ((((uintptr_t)obs) | ((uintptr_t)msk) | (uintptr_t)out0) & 15) == 0;
    if (vec_ok && obs && msk && out0 + Chunk<N>::BYTES <= limit) {
        // the common case (warp-uniform): whole chunk inside the buffers, both outputs wanted, 16-byte aligned
        uint4 *po = reinterpret_cast<uint4 *>(obs + out0), *pm = reinterpret_cast<uint4 *>(msk + out0);
        // Optional (HEXB_L2_OUT_HINT=1): the streaming stores also carry an explicit L2 evict_first policy. Measured both ways
        // with 20 MiB of state kept in L2: 95.5 -> 92.4 us per 1 Mi-game step in one process layout (tools/graph_probe.py) but
        // 92.3 -> 96.3 us in bench.py on another box, so it stays off by default.
        const uint64_t pol = l2_policy(false);
        if (P.variant == VARIANT_B) {
            if (P.out_hint) encode_loop<N, VARIANT_B, true>(src, po, pm, lane, P.one, pol);
            else encode_loop<N, VARIANT_B, false>(src, po, pm, lane, P.one, pol);
        } else {
            if (P.out_hint) encode_loop<N, VARIANT_A, true>(src, po, pm, lane, P.one, pol);
            else encode_loop<N, VARIANT_A, false>(src, po, pm, lane, P.one, pol);
        }
        return;
    }
    for (int i = lane; i < Chunk<N>::VECS; i += kWarp) {
        const uint4 x = src[i];
        Vec4 in = {x.x, x.y, x.z, x.w}, o, m;
        encode_vec<N>(in, P.variant, o, m);
        const long long off = out0 + 16ll * i;
        if (vec_ok && off + 16 <= limit) {
            if (obs) __stcs(reinterpret_cast<uint4 *>(obs + off), make_uint4(o.x, o.y, o.z, o.w));
            if (msk) __stcs(reinterpret_cast<uint4 *>(msk + off), make_uint4(m.x, m.y, m.z, m.w));
        } else {
            if (obs) store_tail(obs, off, limit, o);
            if (msk) store_tail(msk, off, limit, m);
        }
    }
}

// K7: lane 0 adds the warp's packed, reduced statistics (see the packing at the call sites) to the warp's stripe
__device__ __forceinline__ void add_stats(const Params &P, long long wglobal, int lane, uint32_t sa, uint32_t sb) {
    if (lane != 0) return;
    unsigned long long *stripe = reinterpret_cast<unsigned long long *>(P.stats) + 8 * (wglobal & (kStatStripes - 1));
    if (sa) {
        if (sa & 63u) atomicAdd(stripe + 0, (unsigned long long)(sa & 63u));
        if ((sa >> 6) & 63u) atomicAdd(stripe + 1, (unsigned long long)((sa >> 6) & 63u));
        if ((sa >> 12) & 63u) atomicAdd(stripe + 2, (unsigned long long)((sa >> 12) & 63u));
        if ((sa >> 18) & 63u) atomicAdd(stripe + 3, (unsigned long long)((sa >> 18) & 63u));
        if ((sa >> 24) & 63u) atomicAdd(stripe + 5, (unsigned long long)((sa >> 24) & 63u));
        atomicAdd(stripe + 4, (unsigned long long)(sb & 0x3fffu));
    }
    if ((sb >> 14) & 63u) atomicAdd(stripe + 6, (unsigned long long)((sb >> 14) & 63u));
    if (sb >> 20) atomicAdd(stripe + 7, (unsigned long long)(sb >> 20));
}

// KIND_STEP: the instantiation the timed path launches (one env step, nothing else compiled in, step index folded to 0);
// KIND_ROLLOUT: P.steps env steps per launch on the resident chunk (hexb_rollout);
// KIND_OTHER: reset / raw ply / half step, selected at run time by P.mode.
enum : int { KIND_OTHER = 0, KIND_STEP = 1, KIND_ROLLOUT = 2 };
template <int N, int KIND, bool BATCHED>
__global__ void __launch_bounds__(kCtaThreads, min_ctas(N)) hexb_step_kernel(const Params P) {
    constexpr bool STEP_ONLY = KIND != KIND_OTHER;
    extern __shared__ __align__(128) uint8_t smem[];
    using SL = SmemLayout<N>;
    constexpr int C = Geo<N>::C;
    constexpr uint32_t FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint8_t *chunk = smem + wid * SL::CHUNK;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + SL::BAR) + wid;
    const long long wglobal = (long long)blockIdx.x * kWarpsPerCta + wid;
    const long long g0 = wglobal * kWarp;   // first game of the chunk
    const long long g = g0 + lane;          // this lane's game
    uint8_t *gl = P.state + wglobal * SL::CHUNK;

    // ---- programmatic dependent launch: the step kernels are launched with the programmatic-stream-serialization attribute, so
    //      a launch that follows another kernel on the stream may be set up while that kernel drains; it waits HERE, before its
    //      first global access, until the previous grid has completed and its writes are visible (a no-op without the attribute).
    //      This hides ~2 us of launch gap per step for steps issued one by one (1 Mi games 11x11: 93.6 -> 91.4 us; a CUDA graph
    //      has no such gap). Letting the next grid in EARLY (griddepcontrol.launch_dependents at the top) was measured and
    //      rejected: its waiting CTAs take slots from this grid (65,536 games of 7x7: 9.1 -> 12.7 us per step).
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // ---- chunk in (asynchronous: labels and records in ONE bulk copy); meanwhile the two draws of the (first) step, which
    //      only need the game's meta and stream-position words (two plain loads of lines the bulk copy is fetching anyway)
    const bool use_hint = P.keep_chunks > 0;               // warp-uniform (kernel-uniform)
    const bool keep = wglobal < P.keep_chunks;
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, SL::CHUNK);
        if (use_hint) bulk_g2s_hint(chunk, gl, SL::CHUNK, bar, l2_policy(keep));
        else bulk_g2s(chunk, gl, SL::CHUNK, bar);
    }
    uint32_t *recw = reinterpret_cast<uint32_t *>(chunk + Geo<N>::CHUNK_LAB) + lane;   // this lane's record word 0 (shared memory)
    uint32_t *lab32 = reinterpret_cast<uint32_t *>(chunk);
    double u_agent = 0.0, u_opp = 0.0;
    if (STEP_ONLY && g < P.G) {
        const uint32_t *grec = reinterpret_cast<const uint32_t *>(gl + Geo<N>::CHUNK_LAB) + lane;
        pre_draws(P, grec[Geo<N>::W * kRecStride], grec[(Geo<N>::W + 1) * kRecStride],
                  (unsigned long long)(P.game_offset + g), u_agent, u_opp);
    }
    __syncwarp();  // the barrier's initialisation is visible to the other lanes
    mbar_wait(bar, 0);
    Rec<N> rec;
    load_rec<N>(recw, rec);
    uint8_t *L = chunk + lane * C;

    // One env step per iteration. hexb_step launches with steps == 1; hexb_rollout runs T steps with the chunk staying in
    // shared memory and the records in registers: the state crosses HBM once per launch instead of once per step.
    const int steps = KIND == KIND_ROLLOUT ? P.steps : 1;
    for (int tt = 0; tt < steps; ++tt) {
        const int t = KIND == KIND_ROLLOUT ? tt : 0;   // a compile-time 0 on the single-step path
        // ---- thread-per-game phase
        uint32_t prmA = 0, prmB = 0, flg = 0;
        if (STEP_ONLY) {
            if (t > 0 && g < P.G) pre_draws(P, rec.meta, rec.draws, (unsigned long long)(P.game_offset + g), u_agent, u_opp);
            Loc loc;
            game_step<N>(L, P, g, t, rec, u_agent, u_opp, loc, prmA, prmB, flg);
            // K7: episode statistics - the eight per-game increments are packed into two words (fields wide enough for the
            // sum over 32 lanes), reduced with two redux.sync, and lane 0 adds the non-zero counters to this warp's stripe
            const uint32_t pa = (uint32_t)loc.st[0] | ((uint32_t)loc.st[1] << 6) | ((uint32_t)loc.st[2] << 12) |
                                ((uint32_t)loc.st[3] << 18) | ((uint32_t)loc.st[5] << 24);
            const uint32_t pb = (uint32_t)loc.st[4] | ((uint32_t)loc.st[6] << 14) | ((uint32_t)loc.st[7] << 20);
            add_stats(P, wglobal, lane, __reduce_add_sync(FULL, pa), __reduce_add_sync(FULL, pb));
        } else if (P.mode == MODE_RESET) {
            game_reset<N>(P, g, rec, flg);
        } else if (P.mode == MODE_HALF) {
            Loc loc;
            game_half<N>(L, P, g, rec, loc, prmA, prmB, flg);
            const uint32_t pa = (uint32_t)loc.st[0] | ((uint32_t)loc.st[1] << 6) | ((uint32_t)loc.st[2] << 12) |
                                ((uint32_t)loc.st[3] << 18) | ((uint32_t)loc.st[5] << 24);
            const uint32_t pb = (uint32_t)loc.st[4] | ((uint32_t)loc.st[6] << 14) | ((uint32_t)loc.st[7] << 20);
            add_stats(P, wglobal, lane, __reduce_add_sync(FULL, pa), __reduce_add_sync(FULL, pb));
        } else {
            game_ply<N>(L, P, g, rec, prmA, flg);
        }
        if ((KIND != KIND_ROLLOUT || tt == steps - 1) && g < P.G) store_rec<N>(recw, rec);
        __syncwarp();  // every game's new stones (and record) are in shared memory

        // ---- warp-per-game row jobs, part 1 (rare): games that finished - terminal observation, clear, opening stone
        uint32_t pending = __ballot_sync(FULL, (flg & (F_RESET | F_TERM)) != 0u);
        while (pending) {
            const int r = __ffs(pending) - 1;
            pending &= pending - 1;
            const uint32_t rf = __shfl_sync(FULL, flg, r);
            row_job_lane<N>(chunk, r, rf, P, g0 + r + (long long)t * P.G, lane, [] { __syncwarp(); });
            __syncwarp();
        }

        // ---- observation + mask. They depend on emptiness and owner bits only, not on the labels, so they are issued
        //      BEFORE the relabel sweeps: the output stores drain to HBM while the warp works through its relabel rows.
        if ((STEP_ONLY || (P.mode != MODE_PLY && P.mode != MODE_HALF)) && (P.obs || P.mask)) {
            encode_chunk<N>(chunk, P, g0, t, lane);
            uint32_t views = __ballot_sync(FULL, (flg & F_VIEW_OPP) != 0u);  // only without auto-reset: finished by the agent's own ply
            if (views) __syncwarp();
            while (views) {
                const int r = __ffs(views) - 1;
                views &= views - 1;
                view_row_lane<N>(chunk, r, P, g0 + r + (long long)t * P.G, lane);
            }
        }

        // ---- row jobs, part 2: relabel sweeps (regions[regions == label] = new label, both plies of the step at once)
        if (!BATCHED) {
            // one row per pass, one word per lane: the faster form when the launch is several waves deep (HBM-bound regime)
            const bool need = (flg & (F_RELABEL | F_RESET)) == F_RELABEL;
            RelabelReq q = {0u, 0u, 0u, 0u, 0u};
            if (need) prep_request(prmA, prmB, q);   // each game's own lane decodes its requests once
            const uint32_t all_rows = __ballot_sync(FULL, need);
            const uint32_t extra = __ballot_sync(FULL, need && q.nx != 0u);   // rows with more than one (old -> new) pair
            uint32_t pending = all_rows & ~extra;                             // the common case first: one pair, no inner loop
            while (pending) {
                const int r = __ffs(pending) - 1;
                pending &= pending - 1;
                const uint32_t so = __shfl_sync(FULL, q.so0, r), sn = __shfl_sync(FULL, q.sn0, r);
                relabel_row_lane2<N, false>(lab32, row_desc<N>(r), lane, so, sn, 0u, 0u, 0, P.one);
                __syncwarp();
            }
            pending = extra;
            while (pending) {
                const int r = __ffs(pending) - 1;
                pending &= pending - 1;
                const uint32_t so = __shfl_sync(FULL, q.so0, r), sn = __shfl_sync(FULL, q.sn0, r);
                const uint32_t xo = __shfl_sync(FULL, q.xo, r), xn = __shfl_sync(FULL, q.xn, r), nx = __shfl_sync(FULL, q.nx, r);
                relabel_row_lane2<N, true>(lab32, row_desc<N>(r), lane, so, sn, xo, xn, (int)nx, P.one);
                __syncwarp();
            }
        } else {
            // RPS rows per pass (Sweep<N>): fewer, wider passes shorten a warp's dependent chain - the faster form when the
            // launch is at most about one wave of warps and the step time is a single warp's latency (small batches, rollouts)
            using SW = Sweep<N>;
            const bool need = (flg & (F_RELABEL | F_RESET)) == F_RELABEL;
            uint32_t olds_l = 0, news_l = 0, n_l = 0;
            if (need) canon_request(prmA, prmB, olds_l, news_l, n_l);
            const uint32_t pend_all = __ballot_sync(FULL, need);
            const int sg = lane / SW::LPR, sl = lane % SW::LPR;
            constexpr int CLASSES = (Chunk<N>::ALIGNED_ROWS || SW::RPS == 1) ? 1 : 2;
#pragma unroll
            for (int cls = 0; cls < CLASSES; ++cls) {
                uint32_t pp = CLASSES == 1 ? pend_all : (pend_all & (cls ? 0xaaaaaaaau : 0x55555555u));
                while (pp) {
                    const int row = pick_row<N>(pp, sg);
                    const int srcl = row & 31;
                    const uint32_t o = __shfl_sync(FULL, olds_l, srcl), nw = __shfl_sync(FULL, news_l, srcl);
                    const uint32_t n = __shfl_sync(FULL, n_l, srcl);
                    relabel_rows_lane<N>(lab32, row, sl, o, nw, (int)n, P.one);
                    __syncwarp();
                }
            }
        }
    }

    // ---- chunk out
    fence_async_smem();  // generic-proxy writes to shared memory -> visible to the async proxy
    __syncwarp();
    if (lane == 0) {
        if (use_hint) bulk_s2g_hint(gl, chunk, SL::CHUNK, l2_policy(keep));
        else bulk_s2g(gl, chunk, SL::CHUNK);
        bulk_wait_read();
    }
}

// ------------------------------------------------------------------------------------------------ small kernels (hexb_views.cuh)
__global__ void hexb_encode_kernel(View V, int view, int8_t *obs, uint8_t *mask) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < V.G * V.N * V.N) encode_at(V, view, i, obs, mask);
}
template <int N>
__global__ void hexb_sample_kernel(View V, int view, const double *u, int32_t *out) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < V.G) sample_at<N>(V, view, g, u, out);
}
__global__ void hexb_export_kernel(View V, double *board, double *regions, double *counter, int8_t *cur, uint8_t *done,
                                   int8_t *winner, int8_t *agent, uint32_t *draws) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < V.G * 2 * (V.N + 2) * (V.N + 2)) export_at(V, i, board, regions, counter, cur, done, winner, agent, draws);
}
template <int N>
__global__ void hexb_import_kernel(Params P, const int8_t *board_true, const int8_t *to_move, const uint8_t *import_mask) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < P.G) import_game<N>(P, g, board_true, to_move, import_mask);
}

__global__ void hexb_stats_kernel(const long long *src, int64_t *dst) {
    if (threadIdx.x < 8) {
        long long v = 0;
        for (int s = 0; s < kStatStripes; ++s) v += src[8 * s + threadIdx.x];
        dst[threadIdx.x] = (int64_t)v;
    }
}

// K8: masked categorical sampling, one warp per game. Cells are dealt to lanes in blocks of 32 (cell = 32*k + lane), so a
// block's inclusive scan over lanes continues the running CDF in cell order.
__global__ void __launch_bounds__(128) hexb_masked_sample_kernel(const float *__restrict__ logits, const uint8_t *__restrict__ mask,
                                                                 const double *__restrict__ u, long long G, int C, int32_t *actions,
                                                                 float *logp, float *entropy) {
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr int KMAX = (HEXB_MAX_BOARD * HEXB_MAX_BOARD + 31) / 32;
    const int lane = threadIdx.x & 31;
    const long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= G) return;
    const float *lg = logits + g * C;
    const uint8_t *mk = mask + g * C;
    const int K = (C + 31) >> 5;
    float l[KMAX];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        const int c = 32 * k + lane;
        l[k] = (k < K && c < C && mk[c]) ? lg[c] : -INFINITY;
        mx = fmaxf(mx, l[k]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
    if (mx == -INFINITY) {  // no legal cell
        if (lane == 0) {
            if (actions) actions[g] = -1;
            if (logp) logp[g] = 0.f;
            if (entropy) entropy[g] = 0.f;
        }
        return;
    }
    float e[KMAX], z = 0.f, sel = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        e[k] = (l[k] == -INFINITY) ? 0.f : __expf(l[k] - mx);
        z += e[k];
        sel += e[k] * ((l[k] == -INFINITY) ? 0.f : (l[k] - mx));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        z += __shfl_xor_sync(FULL, z, o);
        sel += __shfl_xor_sync(FULL, sel, o);
    }
    const float logz = __logf(z);
    const float target = (float)(u[g] * (double)z);
    float run = 0.f;
    int pick = -1, last_legal = -1;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
            float s = e[k];  // inclusive scan over lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float t = __shfl_up_sync(FULL, s, o);
                if (lane >= o) s += t;
            }
            const float cum = run + s;
            const uint32_t hit = __ballot_sync(FULL, e[k] > 0.f && cum > target);
            const uint32_t legal = __ballot_sync(FULL, e[k] > 0.f);
            if (pick < 0 && hit) pick = 32 * k + (__ffs(hit) - 1);
            if (legal) last_legal = 32 * k + (31 - __clz(legal));
            run += __shfl_sync(FULL, s, 31);
        }
    }
    if (pick < 0) pick = last_legal;  // u*z rounded up to the total: the last legal cell
    const int pk = pick >> 5, pl = pick & 31;
    float lp = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
        if (k == pk) lp = l[k];
    lp = __shfl_sync(FULL, lp, pl) - mx - logz;
    if (lane == 0) {
        if (actions) actions[g] = pick;
        if (logp) logp[g] = lp;
        if (entropy) entropy[g] = logz - sel / z;
    }
}

// ------------------------------------------------------------------------------------------------ host side
struct hexb_env {
    hexb_config cfg;
    Params base;  // state pointers + config, I/O pointers null
};

static thread_local int g_last_cuda = 0;
static int cuda_fail(cudaError_t e) {
    g_last_cuda = (int)e;
    return HEXB_ERR_CUDA;
}
#define CK(call)                                  \
    do {                                          \
        cudaError_t e_ = (call);                  \
        if (e_ != cudaSuccess) return cuda_fail(e_); \
    } while (0)

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static bool cfg_ok(const hexb_config *c) {
    if (!c) return false;
    if (c->board_size < HEXB_MIN_BOARD || c->board_size > HEXB_MAX_BOARD) return false;
    if (c->variant != HEXB_VARIANT_A && c->variant != HEXB_VARIANT_B) return false;
    if (c->num_games < 1 || c->game_offset < 0) return false;
    if (c->agent_mode < 0 || c->agent_mode > 2) return false;
    if (c->variant == HEXB_VARIANT_A && c->agent_mode != HEXB_AGENT_BLACK) return false;
    if (c->pool_size < 0 || (c->manual_opponent && c->raw)) return false;
    return true;
}

struct Layout {
    long long Gpad;
    size_t stats_off, total;   // the chunks (labels + records of 32 games each) start at offset 0; the statistics follow
};
static Layout layout_of(const hexb_config *c) {
    Layout L;
    const long long C = (long long)c->board_size * c->board_size;
    L.Gpad = (c->num_games + kTile - 1) / kTile * kTile;
    L.stats_off = align256((size_t)(L.Gpad / 32 * chunk_state_bytes((int)C)));
    L.total = L.stats_off + align256((size_t)kStatStripes * 8 * 8);
    return L;
}

extern "C" {

int32_t hexb_version(void) { return (1 << 16) | 2; }   // 1.2: hexb_get_config, state without the transposed bitboard

const char *hexb_strerror(int32_t code) {
    switch (code) {
        case HEXB_OK: return "ok";
        case HEXB_ERR_ARG: return "bad argument or unsupported configuration";
        case HEXB_ERR_CUDA: return "CUDA call failed";
        case HEXB_ERR_STATE: return "state buffer too small or misaligned";
        case HEXB_ERR_NOGPU: return "no usable CUDA device";
        default: return "unknown error";
    }
}

int32_t hexb_last_cuda_error(void) { return g_last_cuda; }

size_t hexb_state_bytes(const hexb_config *cfg) { return cfg_ok(cfg) ? layout_of(cfg).total : 0; }

int32_t hexb_create(const hexb_config *cfg, void *state, size_t state_bytes, void *stream, hexb_env **out) {
    if (!cfg_ok(cfg) || !out) return HEXB_ERR_ARG;
    const Layout L = layout_of(cfg);
    if (!state || state_bytes < L.total || ((uintptr_t)state & 255)) return HEXB_ERR_STATE;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) return HEXB_ERR_NOGPU;
    CK(cudaSetDevice(cfg->device));
    CK(cudaMemsetAsync(state, 0, L.total, (cudaStream_t)stream));
    hexb_env *e = (hexb_env *)calloc(1, sizeof(hexb_env));
    if (!e) return HEXB_ERR_ARG;
    e->cfg = *cfg;
    Params &P = e->base;
    P.state = (uint8_t *)state;
    P.stats = (long long *)((uint8_t *)state + L.stats_off);
    P.G = cfg->num_games;
    P.Gpad = L.Gpad;
    P.game_offset = cfg->game_offset;
    P.seed = cfg->seed;
    P.variant = cfg->variant;
    P.auto_reset = cfg->auto_reset;
    P.eval_state = cfg->eval_state;
    P.opponent_first = cfg->opponent_first;
    P.agent_mode = cfg->agent_mode;
    P.raw = cfg->raw;
    P.manual_opponent = cfg->manual_opponent;
    P.pool_size = cfg->pool_size;
    P.one = 1u;
    {   // L2 policy of the chunk copies (see l2_policy): when the packed state is too large to live in L2 anyway, keep a fixed
        // 20 MiB of it there across steps and stream the rest. Measured on 11x11: 1 Mi games 99.4 -> 93.0 us per step,
        // 768 Ki 76.2 -> 70.6, 2 Mi 191.2 -> 182.7; 16-24 MiB is the optimum (32: 94.1, 48: 97.2, everything: 99.6 us - the L2
        // holds only so many evict_last lines before they thrash among themselves); below ~64 MiB of state it does not pay.
        // HEXB_L2_KEEP_MB overrides the amount (0 = never).
        const long long cb = chunk_state_bytes(cfg->board_size * cfg->board_size);
        const long long state = L.Gpad / 32 * cb;
        long long keep_bytes = state > (64ll << 20) ? (20ll << 20) : 0;
        const char *kmb = getenv("HEXB_L2_KEEP_MB");
        if (kmb) keep_bytes = atoll(kmb) * (1ll << 20);
        P.keep_chunks = keep_bytes > 0 ? keep_bytes / cb : 0;
        const char *oh = getenv("HEXB_L2_OUT_HINT");
        P.out_hint = (P.keep_chunks > 0 && oh && atoi(oh) == 1) ? 1 : 0;
    }
    *out = e;
    return HEXB_OK;
}

int32_t hexb_destroy(hexb_env *env) {
    if (!env) return HEXB_ERR_ARG;
    free(env);
    return HEXB_OK;
}

int32_t hexb_get_config(const hexb_env *env, hexb_config *out) {
    if (!env || !out) return HEXB_ERR_ARG;
    *out = env->cfg;
    return HEXB_OK;
}

}  // extern "C"

// number of chunk-warps one wave of the step kernel holds on this device (SMs x resident CTAs per SM x warps per CTA)
static long long wave_warps(int device) {
    static long long cached[64] = {0};
    if (device < 0 || device >= 64) return 148ll * 32;
    if (!cached[device]) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0) sms = 148;
        cached[device] = (long long)sms * 32;
    }
    return cached[device];
}

template <int N>
static int launch_tile(const hexb_env *e, const Params &P, cudaStream_t s) {
    constexpr int smem = SmemLayout<N>::BYTES;
    static bool attr_done_dev[64] = {false};   // function attributes are per device
    const int dev = e->cfg.device;
    bool never = false;
    bool &attr_done = (dev >= 0 && dev < 64) ? attr_done_dev[dev] : never;
    if (!attr_done) {
#define HEXB_ATTR(K, B)                                                                                                        \
    CK(cudaFuncSetAttribute(hexb_step_kernel<N, K, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                    \
    CK(cudaFuncSetAttribute(hexb_step_kernel<N, K, B>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        HEXB_ATTR(KIND_STEP, false) HEXB_ATTR(KIND_STEP, true) HEXB_ATTR(KIND_ROLLOUT, false) HEXB_ATTR(KIND_ROLLOUT, true)
        HEXB_ATTR(KIND_OTHER, false)
#undef HEXB_ATTR
        attr_done = true;
    }
    const unsigned grid = (unsigned)(P.Gpad / kCtaThreads);
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid);
    lc.blockDim = dim3(kCtaThreads);
    lc.dynamicSmemBytes = smem;
    lc.stream = s;
    cudaLaunchAttribute la[1];
    la[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    la[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool pdl = !(getenv("HEXB_PDL") && atoi(getenv("HEXB_PDL")) == 0);   // HEXB_PDL=0: plain stream-ordered launches
    lc.attrs = la;
    lc.numAttrs = pdl ? 1 : 0;
    // at most about one wave of warps: the step time is one warp's latency -> the batched relabel sweep; deeper launches are
    // HBM-bound and run the one-row-per-pass sweep (measured on 1 Mi games 11x11: 99.5 us vs 105.0 us; on 4,096 games 6x6: 6.7 vs 6.0 us)
    const bool small = P.Gpad / kWarp <= wave_warps(e->cfg.device);
    const Params &Q = P;
    if (P.mode == MODE_STEP && P.steps == 1) {
        if (small) CK(cudaLaunchKernelEx(&lc, hexb_step_kernel<N, KIND_STEP, true>, Q));
        else CK(cudaLaunchKernelEx(&lc, hexb_step_kernel<N, KIND_STEP, false>, Q));
    } else if (P.mode == MODE_STEP) {
        if (small) CK(cudaLaunchKernelEx(&lc, hexb_step_kernel<N, KIND_ROLLOUT, true>, Q));
        else CK(cudaLaunchKernelEx(&lc, hexb_step_kernel<N, KIND_ROLLOUT, false>, Q));
    } else {
        CK(cudaLaunchKernelEx(&lc, hexb_step_kernel<N, KIND_OTHER, false>, Q));
    }
    CK(cudaGetLastError());
    return HEXB_OK;
}

#define HEXB_FOR_N(X) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18) X(19)

static int dispatch_tile(const hexb_env *e, const Params &P, cudaStream_t s) {
    switch (e->cfg.board_size) {
#define X(n) \
    case n: return launch_tile<n>(e, P, s);
        HEXB_FOR_N(X)
#undef X
    }
    return HEXB_ERR_ARG;
}

static View view_of(const hexb_env *e) {
    View V;
    V.state = e->base.state;
    V.G = e->base.G;
    V.Gpad = e->base.Gpad;
    V.N = e->cfg.board_size;
    V.variant = e->cfg.variant;
    V.raw = e->cfg.raw;
    return V;
}

extern "C" {

int32_t hexb_set_info_buffers(hexb_env *env, int32_t *last_move_opponent, int8_t *winner) {
    if (!env) return HEXB_ERR_ARG;
    env->base.info_opp = last_move_opponent;
    env->base.info_winner = winner;
    return HEXB_OK;
}

int32_t hexb_set_opponent_buffers(hexb_env *env, int32_t *opp_index, uint8_t *to_move) {
    if (!env) return HEXB_ERR_ARG;
    env->base.opp_index = opp_index;
    env->base.to_move = to_move;
    return HEXB_OK;
}

int32_t hexb_half_step(hexb_env *env, int32_t side, const int32_t *actions, float *reward, uint8_t *done, int8_t *term_obs,
                       void *stream) {
    // actions == null is allowed for side 1 only: the built-in random opponent moves (positions imported with the opponent to move)
    if (!env || env->cfg.raw || (side != 0 && side != 1) || (!actions && side != 1)) return HEXB_ERR_ARG;
    if (!env->cfg.manual_opponent) return HEXB_ERR_ARG;   // fused handles step with hexb_step; their resets already open
    CK(cudaSetDevice(env->cfg.device));
    Params P = env->base;
    P.mode = MODE_HALF;
    P.half_side = side;
    P.actions = actions;
    P.reward = reward;
    P.done = done;
    P.term_obs = term_obs;
    return dispatch_tile(env, P, (cudaStream_t)stream);
}

int32_t hexb_reset(hexb_env *env, const uint8_t *reset_mask, const double *open_u, int8_t *obs, uint8_t *mask, void *stream) {
    if (!env) return HEXB_ERR_ARG;
    CK(cudaSetDevice(env->cfg.device));
    Params P = env->base;
    P.mode = MODE_RESET;
    P.reset_mask = reset_mask;
    P.open_u = open_u;
    P.obs = obs;
    P.mask = mask;
    return dispatch_tile(env, P, (cudaStream_t)stream);
}

int32_t hexb_step(hexb_env *env, const int32_t *actions, const double *opp_u, int8_t *obs, uint8_t *mask, float *reward,
                  uint8_t *done, int8_t *term_obs, int32_t *actions_out, void *stream) {
    if (!env || env->cfg.raw || env->cfg.manual_opponent) return HEXB_ERR_ARG;
    CK(cudaSetDevice(env->cfg.device));
    Params P = env->base;
    P.mode = MODE_STEP;
    P.steps = 1;
    P.actions = actions;
    P.opp_u = opp_u;
    P.obs = obs;
    P.mask = mask;
    P.reward = reward;
    P.done = done;
    P.term_obs = term_obs;
    P.actions_out = actions_out;
    return dispatch_tile(env, P, (cudaStream_t)stream);
}

int32_t hexb_rollout(hexb_env *env, int32_t num_steps, int8_t *obs, uint8_t *mask, float *reward, uint8_t *done, int8_t *term_obs,
                     int32_t *actions_out, void *stream) {
    if (!env || env->cfg.raw || env->cfg.manual_opponent || num_steps < 1 || num_steps > 65536) return HEXB_ERR_ARG;
    CK(cudaSetDevice(env->cfg.device));
    Params P = env->base;
    P.mode = MODE_STEP;
    P.steps = num_steps;
    P.obs = obs;
    P.mask = mask;
    P.reward = reward;
    P.done = done;
    P.term_obs = term_obs;
    P.actions_out = actions_out;
    return dispatch_tile(env, P, (cudaStream_t)stream);
}

size_t hexb_host_workspace_bytes(const hexb_config *cfg) {
    if (!cfg_ok(cfg)) return 0;
    const size_t G = (size_t)cfg->num_games, C = (size_t)cfg->board_size * cfg->board_size;
    return align256(G * 4) + 2 * align256(G * C) + align256(G * 4) + align256(G);
}

int32_t hexb_step_host(hexb_env *env, void *workspace, const int32_t *actions_host, int8_t *obs_host, uint8_t *mask_host,
                       float *reward_host, uint8_t *done_host, void *stream) {
    if (!env || !workspace || env->cfg.raw) return HEXB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t G = (size_t)env->cfg.num_games, C = (size_t)env->cfg.board_size * env->cfg.board_size;
    uint8_t *w = (uint8_t *)workspace;
    int32_t *d_act = (int32_t *)w;            w += align256(G * 4);
    int8_t *d_obs = (int8_t *)w;              w += align256(G * C);
    uint8_t *d_mask = w;                      w += align256(G * C);
    float *d_rew = (float *)w;                w += align256(G * 4);
    uint8_t *d_done = w;
    CK(cudaSetDevice(env->cfg.device));
    if (actions_host) CK(cudaMemcpyAsync(d_act, actions_host, G * 4, cudaMemcpyHostToDevice, s));
    const int rc = hexb_step(env, actions_host ? d_act : nullptr, nullptr, obs_host ? d_obs : nullptr, mask_host ? d_mask : nullptr,
                             reward_host ? d_rew : nullptr, done_host ? d_done : nullptr, nullptr, nullptr, stream);
    if (rc != HEXB_OK) return rc;
    if (obs_host) CK(cudaMemcpyAsync(obs_host, d_obs, G * C, cudaMemcpyDeviceToHost, s));
    if (mask_host) CK(cudaMemcpyAsync(mask_host, d_mask, G * C, cudaMemcpyDeviceToHost, s));
    if (reward_host) CK(cudaMemcpyAsync(reward_host, d_rew, G * 4, cudaMemcpyDeviceToHost, s));
    if (done_host) CK(cudaMemcpyAsync(done_host, d_done, G, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return HEXB_OK;
}

int32_t hexb_ply(hexb_env *env, const int32_t *actions, int8_t *ret, void *stream) {
    if (!env || !actions || !env->cfg.raw) return HEXB_ERR_ARG;
    CK(cudaSetDevice(env->cfg.device));
    Params P = env->base;
    P.mode = MODE_PLY;
    P.actions = actions;
    P.ret = ret;
    return dispatch_tile(env, P, (cudaStream_t)stream);
}

int32_t hexb_encode(hexb_env *env, int32_t view, int8_t *obs, uint8_t *mask, void *stream) {
    if (!env || (view != 0 && view != 1)) return HEXB_ERR_ARG;
    CK(cudaSetDevice(env->cfg.device));
    const View V = view_of(env);
    const long long n = V.G * V.N * V.N;
    hexb_encode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(V, view, obs, mask);
    CK(cudaGetLastError());
    return HEXB_OK;
}

int32_t hexb_sample_actions(hexb_env *env, int32_t view, const double *u, int32_t *actions_out, void *stream) {
    if (!env || !u || !actions_out || (view != 0 && view != 1)) return HEXB_ERR_ARG;
    CK(cudaSetDevice(env->cfg.device));
    const View V = view_of(env);
    const unsigned grid = (unsigned)((V.G + 127) / 128);
    switch (V.N) {
#define X(n) \
    case n: hexb_sample_kernel<n><<<grid, 128, 0, (cudaStream_t)stream>>>(V, view, u, actions_out); break;
        HEXB_FOR_N(X)
#undef X
    }
    CK(cudaGetLastError());
    return HEXB_OK;
}

int32_t hexb_export_state(hexb_env *env, double *board, double *regions, double *region_counter, int8_t *cur, uint8_t *done,
                          int8_t *winner, int8_t *agent, uint32_t *draws, void *stream) {
    if (!env) return HEXB_ERR_ARG;
    CK(cudaSetDevice(env->cfg.device));
    const View V = view_of(env);
    const long long n = V.G * 2 * (V.N + 2) * (V.N + 2);
    hexb_export_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(V, board, regions, region_counter, cur, done, winner,
                                                                                       agent, draws);
    CK(cudaGetLastError());
    return HEXB_OK;
}

int32_t hexb_import_boards(hexb_env *env, const int8_t *board_true, const int8_t *to_move, const uint8_t *import_mask, void *stream) {
    if (!env || !board_true) return HEXB_ERR_ARG;
    CK(cudaSetDevice(env->cfg.device));
    const Params P = env->base;
    const unsigned grid = (unsigned)((P.G + 127) / 128);
    switch (env->cfg.board_size) {
#define X(n) \
    case n: hexb_import_kernel<n><<<grid, 128, 0, (cudaStream_t)stream>>>(P, board_true, to_move, import_mask); break;
        HEXB_FOR_N(X)
#undef X
    }
    CK(cudaGetLastError());
    return HEXB_OK;
}

int32_t hexb_stats(hexb_env *env, int64_t *out8, void *stream) {
    if (!env || !out8) return HEXB_ERR_ARG;
    CK(cudaSetDevice(env->cfg.device));
    hexb_stats_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(env->base.stats, out8);
    CK(cudaGetLastError());
    return HEXB_OK;
}

int32_t hexb_masked_sample(const float *logits, const uint8_t *mask, const double *u, int64_t num_games, int32_t num_cells,
                           int32_t *actions, float *logp, float *entropy, int32_t device, void *stream) {
    if (!logits || !mask || !u || num_games < 1 || num_cells < 1 || num_cells > HEXB_MAX_BOARD * HEXB_MAX_BOARD) return HEXB_ERR_ARG;
    CK(cudaSetDevice(device));
    const unsigned grid = (unsigned)((num_games * 32 + 127) / 128);
    hexb_masked_sample_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(logits, mask, u, num_games, num_cells, actions, logp, entropy);
    CK(cudaGetLastError());
    return HEXB_OK;
}

}  // extern "C"
