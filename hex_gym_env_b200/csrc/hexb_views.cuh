// hexb_views.cuh - state views that are not on the timed path: standalone observation/mask encoder (K5),
// k-th-empty sampler (K4), reference-layout export and preset-board import (K6). Written per element so the
// kernels in hexb_kernels.cu are one-line wrappers and the host emulator (tests/emu) can loop over the same code.
#pragma once
#include "hexb_phases.cuh"

namespace hexb {

struct View {
    const uint8_t *state;
    long long G, Gpad;
    int N, variant, raw;
    int obs_f32;   // dtype of the obs buffer handed to encode_at (hexb_config.obs_dtype)
};
HEXB_HD const uint32_t *view_rec(const View &V, long long g) {
    return reinterpret_cast<const uint32_t *>(V.state + rec_offset(g, V.N * V.N));
}
HEXB_HD const uint8_t *view_labels(const View &V, long long g) { return V.state + labels_offset(g, V.N * V.N); }
HEXB_HD uint32_t view_meta(const View &V, long long g) {
    const int W = (V.N * V.N + 31) / 32;
    return view_rec(V, g)[W * kRecStride];
}

// K5: obs + mask of the current state. view 0: the agent's (stored orientation, or the opponent's for an episode the
// agent's own ply ended - what step() returned); view 1: the side to move's (variant-B one-ply HexEnv).
HEXB_HD void encode_at(const View &V, int view, long long i, int8_t *obs, uint8_t *mask) {
    const int C = V.N * V.N;
    const long long g = i / C;
    const int c = (int)(i - g * C);
    const uint32_t meta = view_meta(V, g);
    bool opp = false;
    if (V.variant == VARIANT_B) {
        if (view == 1 || V.raw) opp = (meta & M_TOMOVE) != 0u;
        else opp = (meta & M_DONE) && (meta & M_AGENT_ENDED);
    } else if (view == 1 && !V.raw) {
        opp = (meta & M_TOMOVE) != 0u;  // what HexEnv.opponent_move shows its policy: invert_board (HexGame.py:333-334)
    }
    const int y = c / V.N, x = c - y * V.N;
    const uint32_t b = view_labels(V, g)[opp ? x * V.N + y : c];
    uint32_t mk;
    const uint32_t ob = encode_byte(b, V.variant, opp, mk);
    if (obs) store_obs(obs, V.obs_f32, i, ob);
    if (mask) mask[i] = (uint8_t)mk;
}

// K4: int(u * n_empty)-th empty cell in row-major order of the chosen view.
template <int N>
HEXB_HD void sample_at(const View &V, int view, long long g, const double *u, int32_t *out) {
    constexpr int W = Geo<N>::W;
    const uint32_t *rw = view_rec(V, g);
    const uint32_t meta = rw[W * kRecStride];
    const bool opp = (V.variant == VARIANT_B ? (view == 1 || V.raw) : (view == 1 && !V.raw)) && (meta & M_TOMOVE);
    uint32_t occ[W];
#pragma unroll
    for (int w = 0; w < W; ++w) occ[w] = rw[w * kRecStride];
    const int n = count_empty<N>(occ);
    if (n <= 0) { out[g] = -1; return; }
    const int k = choice_of(u[g], n);
    if (!opp) { out[g] = select_kth_zero<N>(occ, k); return; }
    int x;
    const int cell = select_kth_zero_colmajor<N>(occ, k, x);   // the opponent's view is the transpose of the stored board
    out[g] = x * N + (cell - x) / N;
}

// K6: reference-layout dump. Element i = (game, plane, padded cell); board / scalars ride on the first elements.
HEXB_HD void export_at(const View &V, long long i, double *board, double *regions, double *counter, int8_t *cur, uint8_t *done,
                       int8_t *winner, int8_t *agent, uint32_t *draws) {
    const int N = V.N, C = N * N, Pd = N + 2, P2 = Pd * Pd;
    const int W = (C + 31) / 32;
    const long long g = i / (2 * P2);
    const int r = (int)(i - g * 2 * P2);
    const int pl = r / P2, pc = r - pl * P2;
    const int Y = pc / Pd, X = pc - Y * Pd;
    const uint32_t meta = view_meta(V, g);
    const int tr = (meta & M_TRANSPOSED) ? 1 : 0;
    const int sp = pl ^ tr;  // stored player that is true colour `pl`
    const uint8_t *L = view_labels(V, g);
    if (regions) {
        const uint32_t far = (meta & (sp ? M_FAR_C1 : M_FAR_R1)) ? 1u : 2u;
        uint32_t v = 0;
        if (Y >= 1 && Y <= N && X >= 1 && X <= N) {
            const int y = Y - 1, x = X - 1;
            const uint32_t b = L[tr ? x * N + y : y * N + x];
            v = (b != 0u && (int)(b >> 7) == sp) ? (b & 0x7fu) : 0u;
        } else if (pl == 0) {  // BLACK plane: rows 0 and N+1 (HexGame.py:43,45 / HexSingleGame.py:47,49)
            v = (Y == 0) ? 1u : (Y == N + 1 ? far : 0u);
        } else {               // WHITE plane: cols 0 and N+1 (HexGame.py:42,44 / HexSingleGame.py:46,48)
            v = (X == 0) ? 1u : (X == N + 1 ? far : 0u);
        }
        regions[i] = (double)v;
    }
    if (pl == 0 && pc < C && board) {
        const int c = pc, y = c / N, x = c - y * N;
        bool opp = false;
        if (V.variant == VARIANT_B) opp = V.raw ? ((meta & M_TOMOVE) != 0u) : ((meta & M_DONE) && (meta & M_AGENT_ENDED));
        const uint32_t b = L[opp ? x * N + y : c];
        uint32_t mk;
        const uint32_t ob = encode_byte(b, V.variant, opp, mk);
        board[g * C + c] = (double)(int8_t)ob;
    }
    if (r == 0) {
        if (counter) {
            const uint32_t cr = (meta >> M_CTR_R_SHIFT) & 0xffu, cc = (meta >> M_CTR_C_SHIFT) & 0xffu;
            counter[2 * g + 0] = (double)(tr ? cc : cr);
            counter[2 * g + 1] = (double)(tr ? cr : cc);
        }
        if (cur) cur[g] = (int8_t)(((meta & M_TOMOVE) ? 1 : 0) ^ tr);
        if (done) done[g] = (meta & M_DONE) ? 1 : 0;
        if (winner) {
            const uint32_t w = (meta & M_WIN_MASK) >> M_WIN_SHIFT;
            winner[g] = (int8_t)(w == 0u ? -1 : (int)((w - 1u) ^ (uint32_t)tr));
        }
        if (agent) agent[g] = (int8_t)tr;
        if (draws) draws[g] = view_rec(V, g)[(W + 1) * kRecStride];
    }
}

// K6b: HexGame.__init__ with a preset board: raster-order flood_fill rebuild (HexGame.py:53-61, HexSingleGame.py:57-65).
template <int N>
// On an env handle (not raw) the game keeps its agent colour and its position in the random stream: the board is given in TRUE
// coordinates and colours, stones are visited in TRUE raster order (the labels depend on it) and mapped into the stored
// (agent's) orientation. `import_mask` (nullable) selects the games to overwrite.
HEXB_HD void import_game(const Params &P, long long g, const int8_t *board_true, const int8_t *to_move, const uint8_t *import_mask) {
    constexpr int C = Geo<N>::C;
    if (import_mask && !import_mask[g]) return;
    uint8_t *L = P.state + labels_offset(g, C);
    uint32_t *recw = reinterpret_cast<uint32_t *>(P.state + rec_offset(g, C));
    Rec<N> rec;
    load_rec<N>(recw, rec);
    const bool env_game = !P.raw && (rec.meta & M_COLOUR_SET);
    const uint32_t keep = env_game ? (rec.meta & (M_TRANSPOSED | M_COLOUR_SET)) : M_COLOUR_SET;
    const int tr = (keep & M_TRANSPOSED) ? 1 : 0;
    if (!env_game) rec.draws = 0;
#pragma unroll
    for (int w = 0; w < Geo<N>::W; ++w) rec.occ_rm[w] = 0u;
    rec.meta = keep | M_LIVE | (3u << M_CTR_R_SHIFT) | (3u << M_CTR_C_SHIFT);
    for (int c = 0; c < C; ++c) L[c] = 0;
    for (int c = 0; c < C; ++c) {
        const int v = board_true[g * C + c];
        if (v != 0 && v != 1) continue;
        uint32_t prm;
        place_stone<N>(L, rec, v ^ tr, tr ? transpose_cell<N>(c) : c, prm);
        if (prm & P_NEED)
            for (int k = 0; k < C; ++k) L[k] = (uint8_t)relabel_byte(L[k], prm);
    }
    if (((to_move && to_move[g]) ? 1 : 0) ^ tr) rec.meta |= M_TOMOVE;   // stored "C to move" = the side that is not the agent
    store_rec<N>(recw, rec);
}

// K6c: HexGame.__init__ with connected_stones given (HexGame.py:46-51, HexSingleGame.py:50-55): the label planes are adopted as
// they are and region_counter = max(plane) + 1 per colour - what HexEnv.reset does from its second call on with the planes it
// cached at the first one (HexGame.py:214-220, HexSingleGame.py:226-231), and with user-supplied `regions=`. After a merge the
// highest label can be lower than the number of labels ever handed out, so these counters differ from a raster-order rebuild's.
// planes u8[G,2,N+2,N+2] in the reference's layout (true colours and coordinates, borders included: the far border holds 1 once
// that colour has connected). Board size at run time: this is a set-up path.
HEXB_HD void import_labels_game(const Params &P, int N, long long g, const int8_t *board_true, const uint8_t *planes, const int8_t *to_move,
                                const uint8_t *import_mask) {
    const int C = N * N, W = (C + 31) / 32, Pd = N + 2, P2 = Pd * Pd;
    if (import_mask && !import_mask[g]) return;
    uint8_t *L = P.state + labels_offset(g, C);
    uint32_t *recw = reinterpret_cast<uint32_t *>(P.state + rec_offset(g, C));
    const uint32_t old_meta = recw[W * kRecStride];
    const bool env_game = !P.raw && (old_meta & M_COLOUR_SET);
    const uint32_t keep = env_game ? (old_meta & (M_TRANSPOSED | M_COLOUR_SET)) : M_COLOUR_SET;
    const int tr = (keep & M_TRANSPOSED) ? 1 : 0;
    const uint8_t *pl = planes + g * 2 * P2;
    for (int w = 0; w < W; ++w) recw[w * kRecStride] = 0u;
    for (int c = 0; c < C; ++c) {
        const int y = c / N, x = c - y * N;
        const int sc = tr ? x * N + y : c;              // stored cell
        const int v = board_true[g * C + c];
        uint32_t b = 0;
        if (v == 0 || v == 1) {
            b = (pl[v * P2 + (y + 1) * Pd + (x + 1)] & 0x7fu) | ((uint32_t)(v ^ tr) << 7);
            recw[(sc >> 5) * kRecStride] |= 1u << (sc & 31);
        }
        L[sc] = (uint8_t)b;
    }
    uint32_t meta = keep | M_LIVE;
    for (int col = 0; col < 2; ++col) {
        uint32_t m = 0;
        for (int i = 0; i < P2; ++i) m = pl[col * P2 + i] > m ? pl[col * P2 + i] : m;
        const uint32_t ctr = m + 1u > 255u ? 255u : m + 1u;       // region_counter = max(plane) + 1
        const int sp = col ^ tr;                                   // stored player holding true colour `col`
        meta |= ctr << (sp ? M_CTR_C_SHIFT : M_CTR_R_SHIFT);
        const uint32_t far = col == 0 ? pl[(N + 1) * Pd + 1] : pl[P2 + 1 * Pd + (N + 1)];
        if (far == 1u) meta |= sp ? M_FAR_C1 : M_FAR_R1;
    }
    if (((to_move && to_move[g]) ? 1 : 0) ^ tr) meta |= M_TOMOVE;
    recw[W * kRecStride] = meta;
    if (!env_game) recw[(W + 1) * kRecStride] = 0u;
}

}  // namespace hexb
