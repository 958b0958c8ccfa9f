// hexb_phases.cuh - the phases of the fused step kernel, one function per phase.
//
// A CTA owns a TILE of kTile consecutive games. Its label bytes (kTile*C contiguous bytes, the same
// layout as the obs / mask outputs) sit in shared memory for the whole step:
//
//   phase_agent      thread-per-game   agent ply: validity, stone, neighbour labels, merge request, win
//   pass_relabel     cooperative       byte-SIMD relabel of the tile's label words (request #1)
//   phase_opponent   thread-per-game   random opponent ply (k-th empty cell, column-major), win, reward/done,
//                                      episode accounting, auto-reset bookkeeping (+ opening stone)
//   pass_encode      cooperative       relabel request #2 fused with obs + mask encoding and the coalesced
//                                      global stores; rare byte-wise path for straddling words, resets,
//                                      terminal observations and the opponent's (transposed) view
//   phase_clear      thread-per-game   games that reset: zero their label bytes, drop the opening stone
//
// Between phases the kernel puts a __syncthreads(); the host emulator (tests/emu) simply runs each
// phase for all threads before the next one.
#pragma once
#include "hexb_core.cuh"

namespace hexb {

template <int N>
struct Tile {
    uint8_t *lab;     // [kTile*C] label bytes, 16-byte aligned
    uint32_t *prm1;   // [kTile] relabel request of the first ply
    uint32_t *prm2;   // [kTile] relabel request of the second ply
    uint32_t *flg;    // [kTile] reset / view / terminal flags for pass_encode
    long long g0;     // first game of the tile
};

// per-thread values that live from phase_agent to phase_opponent
struct Loc {
    uint32_t f;     // bit 0 inactive (g >= G or never reset), bit 1 was done before this step
    float reward;
    int action;
    int st[8];      // episode statistics increments (see hexb.h: hexb_stats)
};
constexpr uint32_t L_INACTIVE = 1u, L_WASDONE = 2u;

template <int N>
HEXB_HD void load_rec(const Params &P, long long g, Rec<N> &r) {
    constexpr int W = Geo<N>::W;
    const uint32_t *b = P.rec + g;
#pragma unroll
    for (int w = 0; w < W; ++w) r.occ_rm[w] = b[(long long)w * P.Gpad];
#pragma unroll
    for (int w = 0; w < W; ++w) r.occ_cm[w] = b[(long long)(W + w) * P.Gpad];
    r.meta = b[(long long)(2 * W) * P.Gpad];
    r.draws = b[(long long)(2 * W + 1) * P.Gpad];
    r.aux = b[(long long)(2 * W + 2) * P.Gpad];
}
template <int N>
HEXB_HD void store_rec(const Params &P, long long g, const Rec<N> &r) {
    constexpr int W = Geo<N>::W;
    uint32_t *b = P.rec + g;
#pragma unroll
    for (int w = 0; w < W; ++w) b[(long long)w * P.Gpad] = r.occ_rm[w];
#pragma unroll
    for (int w = 0; w < W; ++w) b[(long long)(W + w) * P.Gpad] = r.occ_cm[w];
    b[(long long)(2 * W) * P.Gpad] = r.meta;
    b[(long long)(2 * W + 1) * P.Gpad] = r.draws;
    b[(long long)(2 * W + 2) * P.Gpad] = r.aux;
}

// reward HexEnv.step would hand out again for an already finished variant-A game (HexGame.py:250,267-279)
HEXB_HD float stale_reward_A(uint32_t meta) {
    if (meta & M_INVALID) return -100.f;
    const uint32_t w = (meta & M_WIN_MASK) >> M_WIN_SHIFT;
    return w == 1u ? 1.f : (w == 2u ? -1.f : 0.f);
}

// ---------------------------------------------------------------------------------------------- agent ply
// SelfPlayEnv.step -> HexEnv.step (SelfplayWrapper.py:174-176, HexSingleGame.py:233-263) or variant-A
// HexEnv.step (HexGame.py:244-253). With actions == null the agent is BaseRandomPolicy / random_policy
// itself and takes one draw from the game's stream first, exactly like the reference loop
//   a = BaseRandomPolicy().choose_action(obs); env.step(a).
template <int N>
HEXB_HD void phase_agent(const Tile<N> &T, const Params &P, int t, Rec<N> &rec, Loc &loc) {
    constexpr int C = Geo<N>::C;
    loc.f = 0; loc.reward = 0.f; loc.action = -1;
#pragma unroll
    for (int i = 0; i < 8; ++i) loc.st[i] = 0;
    T.prm1[t] = 0; T.prm2[t] = 0; T.flg[t] = 0;
    const long long g = T.g0 + t;
    if (g >= P.G || !(rec.meta & M_LIVE)) { loc.f = L_INACTIVE; return; }
    if (rec.meta & M_DONE) {
        loc.f = L_WASDONE;
        loc.reward = P.variant == VARIANT_A ? stale_reward_A(rec.meta) : 0.f;
        return;
    }
    const unsigned long long gid = (unsigned long long)(P.game_offset + g);
    int a;
    if (P.actions) a = P.actions[g];
    else {
        const int n = count_empty<N>(rec.occ_rm);
        a = select_kth_zero<N>(rec.occ_rm, choice_of(draw01(P.seed, gid, rec.draws++), n));
    }
    loc.action = a;
    loc.st[6] = 1;
    const bool valid = (unsigned)a < (unsigned)C && !test_bit<N>(rec.occ_rm, a);
    if (!valid) {  // fast_move returns 3, state untouched; the env ends the episode (HexGame.py:252-253, HexSingleGame.py:240-241)
        rec.meta |= M_DONE | M_INVALID | M_AGENT_ENDED;
        loc.reward = P.variant == VARIANT_A ? -100.f : 0.f;
        return;
    }
    uint32_t prm;
    const bool won = place_stone<N>(T.lab + t * C, rec, 0, a, prm);
    T.prm1[t] = prm;
    rec.aux++;
    loc.st[7]++;
    rec.meta ^= M_TOMOVE;
    if (won) {
        rec.meta |= M_DONE | (1u << M_WIN_SHIFT) | M_AGENT_ENDED;
        loc.reward = 1.f;
    } else if (P.variant == VARIANT_B && count_empty<N>(rec.occ_rm) == 0) {
        rec.meta |= M_DONE | M_AGENT_ENDED;  // HexSingleGame.py:117-119 (cannot happen from an empty start)
    }
}

// ---------------------------------------------------------------------------------------------- raw ply (hexb_ply)
// Batched HexGame.make_move: variant A in true coordinates (HexGame.py:85-111, no done guard); variant B
// with the action in the mover's perspective, as HexEnv.step feeds it (HexSingleGame.py:239, 98-106).
template <int N>
HEXB_HD void phase_ply(const Tile<N> &T, const Params &P, int t, Rec<N> &rec) {
    constexpr int C = Geo<N>::C;
    T.prm1[t] = 0; T.prm2[t] = 0; T.flg[t] = 0;
    const long long g = T.g0 + t;
    if (g >= P.G || !(rec.meta & M_LIVE)) return;
    const int a = P.actions[g];
    const int p = (rec.meta & M_TOMOVE) ? 1 : 0;
    int r = 3;
    if ((unsigned)a < (unsigned)C) {
        const int cell = (P.variant == VARIANT_B && p) ? transpose_cell<N>(a) : a;
        if (!test_bit<N>(rec.occ_rm, cell)) {
            uint32_t prm;
            const bool won = place_stone<N>(T.lab + t * C, rec, p, cell, prm);
            T.prm1[t] = prm;
            rec.aux++;
            rec.meta ^= M_TOMOVE;
            r = -1;
            if (won) {
                rec.meta = (rec.meta & ~M_WIN_MASK) | M_DONE | ((uint32_t)(p + 1) << M_WIN_SHIFT);
                r = p;
            } else if (P.variant == VARIANT_B && count_empty<N>(rec.occ_rm) == 0) {
                rec.meta |= M_DONE;
            }
        }
    }
    if (r == 3 && P.variant == VARIANT_B) rec.meta |= M_DONE | M_INVALID;  // HexSingleGame.py:240-241
    if (P.ret) P.ret[g] = (int8_t)r;
}

// ---------------------------------------------------------------------------------------------- relabel pass
// regions[regions == label] = new_region_label (HexGame.py:141-142, HexSingleGame.py:152-153) for every game
// of the tile at once: thread `tid` owns words tid, tid+kTile, ... of the tile's label bytes. A word whose four
// cells straddle two games is handled byte by byte. Warps skip words whose games requested nothing.
template <int N>
HEXB_HD void pass_relabel(const Tile<N> &T, const uint32_t *prm, int tid) {
    constexpr int C = Geo<N>::C;
    uint32_t *lab32 = reinterpret_cast<uint32_t *>(T.lab);
    for (int j = tid; j < Geo<N>::TILE_WORDS; j += kTile) {
        const int byte0 = 4 * j;
        const int gA = byte0 / C, c0 = byte0 - gA * C;
        const bool whole = c0 + 4 <= C;
        const uint32_t pa = prm[gA];
        const uint32_t pb = whole ? 0u : prm[gA + 1];
        if (!((pa | pb) & P_NEED)) continue;  // a warp whose 32 words (about one game) asked for nothing skips as a whole
        if (whole) {
            if (pa & P_NEED) {
                const uint32_t x = lab32[j];
                const uint32_t x2 = relabel_word(x, pa);
                if (x2 != x) lab32[j] = x2;
            }
        } else {
            for (int k = 0; k < 4; ++k) {
                const uint32_t pk = (c0 + k < C) ? pa : pb;
                const uint32_t b = T.lab[byte0 + k];
                const uint32_t b2 = relabel_byte(b, pk);
                if (b2 != b) T.lab[byte0 + k] = (uint8_t)b2;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- opponent ply
// SelfPlayEnv.continue_game (SelfplayWrapper.py:146-172) / HexEnv.opponent_move (HexGame.py:332-349) with the
// random policy, then reward / done, DummyVecEnv-style auto-reset and the episode counters.
template <int N>
HEXB_HD void phase_opponent(const Tile<N> &T, const Params &P, int t, Rec<N> &rec, Loc &loc) {
    constexpr int C = Geo<N>::C;
    const long long g = T.g0 + t;
    if (g >= P.G) return;
    if (loc.f & L_INACTIVE) {
        if (P.reward) P.reward[g] = 0.f;
        if (P.done) P.done[g] = 1;
        if (P.actions_out) P.actions_out[g] = -1;
        return;
    }
    const unsigned long long gid = (unsigned long long)(P.game_offset + g);
    const bool was_done = (loc.f & L_WASDONE) != 0u;
    if (!was_done && !(rec.meta & M_DONE)) {
        double u;
        if (P.opp_u) u = P.opp_u[2 * g];
        else {
            if (P.variant == VARIANT_B) rec.draws++;  // rv = random.uniform(0,1), unused (SelfplayWrapper.py:159)
            u = draw01(P.seed, gid, rec.draws++);
        }
        const int n = count_empty<N>(rec.occ_cm);
        const int i = select_kth_zero<N>(rec.occ_cm, choice_of(u, n));
        const int x = i / N, y = i - x * N;
        uint32_t prm;
        const bool won = place_stone<N>(T.lab + t * C, rec, 1, y * N + x, prm);
        T.prm2[t] = prm;
        rec.aux++;
        loc.st[7]++;
        rec.meta ^= M_TOMOVE;
        if (won) {
            rec.meta |= M_DONE | (2u << M_WIN_SHIFT);
            loc.reward = -1.f;
        } else if (P.variant == VARIANT_B && n == 1) {
            rec.meta |= M_DONE;
        }
    }
    const bool is_done = (rec.meta & M_DONE) != 0u;
    uint32_t flg = 0;
    if (is_done && !was_done) {  // episode accounting (true colours)
        const uint32_t w = (rec.meta & M_WIN_MASK) >> M_WIN_SHIFT;
        const bool tr = (rec.meta & M_TRANSPOSED) != 0u;
        loc.st[0] = 1;
        loc.st[1] = (w == 1u && !tr) || (w == 2u && tr);
        loc.st[2] = (w == 2u && !tr) || (w == 1u && tr);
        loc.st[3] = (w == 1u);
        loc.st[4] = (int)(rec.aux & 0xffffu);
        loc.st[5] = (rec.meta & M_INVALID) != 0u;
        if (P.term_obs) flg |= F_TERM | (((rec.meta & M_AGENT_ENDED) && P.variant == VARIANT_B) ? F_TERM_OPP : 0u);
    }
    if (P.reward) P.reward[g] = loc.reward;
    if (P.done) P.done[g] = is_done ? 1 : 0;
    if (P.actions_out) P.actions_out[g] = loc.action;
    if (is_done) {
        if (P.auto_reset) {
            reset_game<N>(rec, P, gid, P.opp_u ? &P.opp_u[2 * g + 1] : nullptr, flg);
            if (flg & F_OPEN) loc.st[7]++;  // the opponent's opening stone is a ply of this step
        }
        else if ((rec.meta & M_AGENT_ENDED) && P.variant == VARIANT_B) flg |= F_VIEW_OPP;
    }
    T.flg[t] = flg;
}

// ---------------------------------------------------------------------------------------------- reset (hexb_reset)
template <int N>
HEXB_HD void phase_reset(const Tile<N> &T, const Params &P, int t, Rec<N> &rec) {
    T.prm1[t] = 0; T.prm2[t] = 0;
    uint32_t flg = 0;
    const long long g = T.g0 + t;
    if (g < P.G) {
        if (!P.reset_mask || P.reset_mask[g]) {
            reset_game<N>(rec, P, (unsigned long long)(P.game_offset + g), P.open_u ? &P.open_u[g] : nullptr, flg);
        } else if ((rec.meta & M_DONE) && (rec.meta & M_AGENT_ENDED) && P.variant == VARIANT_B && !P.raw) {
            flg |= F_VIEW_OPP;
        }
    }
    T.flg[t] = flg;
}

// ---------------------------------------------------------------------------------------------- encode pass
HEXB_HD void store_out_word(uint8_t *base, long long off, long long limit, uint32_t w) {
    if (off + 4 <= limit) *reinterpret_cast<uint32_t *>(base + off) = w;
    else
        for (int k = 0; k < 4; ++k)
            if (off + k < limit) base[off + k] = (uint8_t)(w >> (8 * k));
}

// get_action_mask (HexGame.py:203-204) / legal_actions (HexSingleGame.py:205-206) and the observation the env
// returns (the live simulator.board: HexGame.py:294, HexSingleGame.py:262 after invert_board :265-271), for all
// games of the tile, 4 cells per thread per iteration, written straight to global memory in [G,C] order.
template <int N>
HEXB_HD void pass_encode(const Tile<N> &T, const Params &P, int tid) {
    constexpr int C = Geo<N>::C;
    uint32_t *lab32 = reinterpret_cast<uint32_t *>(T.lab);
    const long long out0 = T.g0 * C;          // byte offset of the tile in obs / mask / term_obs
    const long long limit = P.G * C;          // bytes that exist in the caller's buffers
    for (int j = tid; j < Geo<N>::TILE_WORDS; j += kTile) {
        const int byte0 = 4 * j;
        const int gA = byte0 / C, c0 = byte0 - gA * C;
        const bool whole = c0 + 4 <= C;
        const uint32_t fa = T.flg[gA];
        uint32_t obsw, mskw;
        if (whole && fa == 0u) {  // common case: one game, no reset, agent's view
            const uint32_t pa = T.prm2[gA];
            uint32_t x = lab32[j];
            if (pa & P_NEED) {
                const uint32_t x2 = relabel_word(x, pa);
                if (x2 != x) lab32[j] = x2;
                x = x2;
            }
            encode_word(x, P.variant, obsw, mskw);
        } else {
            obsw = 0; mskw = 0;
            for (int k = 0; k < 4; ++k) {
                int gg = gA, cc = c0 + k;
                if (cc >= C) { gg += 1; cc -= C; }
                const uint32_t pk = T.prm2[gg], fk = T.flg[gg];
                const uint8_t *Lg = T.lab + gg * C;
                const uint32_t b = Lg[cc];
                const uint32_t b2 = relabel_byte(b, pk);
                if (b2 != b) T.lab[byte0 + k] = (uint8_t)b2;
                const int ct = transpose_cell<N>(cc);
                uint32_t mk, ob;
                if (fk & F_TERM) {  // info["terminal_observation"]: what step() itself returned
                    uint32_t tmk;
                    const uint32_t tb = (fk & F_TERM_OPP) ? encode_byte(relabel_byte(Lg[ct], pk), P.variant, true, tmk)
                                                          : encode_byte(b2, P.variant, false, tmk);
                    const long long o = out0 + byte0 + k;
                    if (o < limit) P.term_obs[o] = (int8_t)tb;
                }
                if (fk & F_RESET) {
                    const uint32_t bo = ((fk & F_OPEN) && (uint32_t)cc == (fk >> 16)) ? ((fk >> 8) & 0xffu) : 0u;
                    ob = encode_byte(bo, P.variant, false, mk);
                } else if (fk & F_VIEW_OPP) {
                    ob = encode_byte(relabel_byte(Lg[ct], pk), P.variant, true, mk);
                } else {
                    ob = encode_byte(b2, P.variant, false, mk);
                }
                obsw |= ob << (8 * k);
                mskw |= mk << (8 * k);
            }
        }
        if (P.obs) store_out_word(reinterpret_cast<uint8_t *>(P.obs), out0 + byte0, limit, obsw);
        if (P.mask) store_out_word(P.mask, out0 + byte0, limit, mskw);
    }
}

// ---------------------------------------------------------------------------------------------- clear
// Games that reset get an empty board (HexSingleGame.py:208-231 / HexGame.py:206-220) plus the opponent's
// opening stone when it moves first. Head and tail of the game's byte range are not word aligned, so they
// are cleared byte-wise (neighbouring games own the other bytes of those words).
template <int N>
HEXB_HD void phase_clear(const Tile<N> &T, int t) {
    constexpr int C = Geo<N>::C;
    const uint32_t f = T.flg[t];
    if (!(f & F_RESET)) return;
    int b = t * C;
    const int e = b + C;
    while ((b & 3) && b < e) T.lab[b++] = 0;
    uint32_t *w = reinterpret_cast<uint32_t *>(T.lab + b);
    const int nw = (e - b) >> 2;
    for (int i = 0; i < nw; ++i) w[i] = 0u;
    b += 4 * nw;
    while (b < e) T.lab[b++] = 0;
    if (f & F_OPEN) T.lab[t * C + (int)(f >> 16)] = (uint8_t)((f >> 8) & 0xffu);
}

}  // namespace hexb
