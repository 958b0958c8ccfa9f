/* TEST INFRASTRUCTURE ONLY - CPU restatement of the reference Hex simulator.
 *
 * Restates, in plain C, the algorithm of MBPrdctns/hex_gym_env's hot path. Every function
 * cites the reference lines it follows (paths relative to /root/reference):
 *   variant A  = minihex/HexGame.py          (board 0/1/2 in true coordinates, env with built-in opponent)
 *   variant B  = minihex/HexSingleGame.py    (board -1/+1/0 in the mover's perspective, one ply per step)
 *   self-play  = minihex/SelfplayWrapper.py  (agent ply + random-opponent reply)
 *   random_policy = minihex/__init__.py:8-12
 *
 * It keeps the reference's own data model (a board plus two zero-padded (N+2)x(N+2) region-label
 * planes, quick-find relabel over the whole padded plane) so that it can be compared cell by cell
 * with the real reference. Parity is PINNED: tests/test_oracle_golden.py checks this file against
 * tests/golden/ *.npz, which oracle/gen_golden.py produced by running the unmodified reference.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load the library built from this file. The product (hex_gym_env_b200) never does.
 *
 * Random numbers: the reference calls CPython's global random.random(); here every game owns a
 * counter-based stream (Philox4x32-10 keyed by seed and global game index, oracle/philox.py) that
 * is consumed in exactly the reference's call order.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXN 19
#define MAXC (MAXN * MAXN)
#define MAXP ((MAXN + 2) * (MAXN + 2))

#define BLACK 0
#define WHITE 1
#define NONE (-1) /* Python None */
#define INVALID 3 /* fast_move's "return 3" */

/* ------------------------------------------------------------------ Philox4x32-10 + CPython double */
static inline void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void hexref_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof(c));
}

double hexref_draw(uint64_t seed, uint64_t game, uint32_t idx) {
    uint32_t ctr[4] = {idx, (uint32_t)game, (uint32_t)(game >> 32), 0u};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    hexref_philox4x32_10(ctr, key, o);
    /* CPython random_random(): a = genrand>>5, b = genrand>>6, (a*67108864.0+b)*(1.0/9007199254740992.0) */
    return ((double)(o[0] >> 5) * 67108864.0 + (double)(o[1] >> 6)) * (1.0 / 9007199254740992.0);
}

typedef struct {
    uint64_t seed, game;
    uint32_t idx;
    const double *inject; /* when non-NULL: replay these doubles instead (tests) */
    int inject_pos;
} rng_t;

static double rng_random(rng_t *r) {
    if (r->inject) return r->inject[r->inject_pos++];
    return hexref_draw(r->seed, r->game, r->idx++);
}
/* random.uniform(0,1) == 0 + (1-0)*random(): one draw */
static double rng_uniform01(rng_t *r) { return rng_random(r); }

/* ------------------------------------------------------------------ game core (both variants) */
typedef struct {
    int N, variant; /* variant 0 = A (HexGame.py), 1 = B (HexSingleGame.py) */
    int board[MAXC];
    int regions[2][MAXP];
    int counter[2];
    int cur;   /* simulator.current_player_num */
    int done;  /* simulator.done */
    int winner; /* simulator.winner, NONE if unset */
    int empty_fields;
} game_t;

static inline int empty_code(const game_t *g) { return g->variant == 0 ? 2 : 0; }

/* HexGame.flood_fill  A: HexGame.py:124-142   B: HexSingleGame.py:135-153 */
static void flood_fill(game_t *g, int py, int px) {
    const int P = g->N + 2;
    int *reg = g->regions[g->cur];
    const int y = py + 1, x = px + 1;
    int nb[9], n = 0;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            int v = reg[(y + dy) * P + (x + dx)];
            if ((dy == -1 && dx == -1) || (dy == 1 && dx == 1)) v = 0; /* neighborhood[0,0] = neighborhood[2,2] = 0 */
            nb[n++] = v;
        }
    /* sorted(set(...)) minus the leading 0 */
    int adj[9], na = 0;
    for (int i = 0; i < 9; ++i) {
        if (nb[i] == 0) continue;
        int seen = 0;
        for (int j = 0; j < na; ++j) seen |= (adj[j] == nb[i]);
        if (!seen) adj[na++] = nb[i];
    }
    for (int i = 1; i < na; ++i) { /* insertion sort, na <= 7 */
        int v = adj[i], j = i - 1;
        while (j >= 0 && adj[j] > v) { adj[j + 1] = adj[j]; --j; }
        adj[j + 1] = v;
    }
    if (na == 0) {
        reg[y * P + x] = g->counter[g->cur];
        g->counter[g->cur] += 1;
    } else {
        const int m = adj[0];
        reg[y * P + x] = m;
        for (int i = 1; i < na; ++i)
            for (int c = 0; c < P * P; ++c) /* regions[regions == label] = new_region_label: whole padded plane */
                if (reg[c] == adj[i]) reg[c] = m;
    }
}

/* HexGame.__init__  A: HexGame.py:21-68   B: HexSingleGame.py:26-71  (board given, connected_stones=None) */
static void game_init(game_t *g, int N, int variant, int cur, const int *board /* NULL = empty */) {
    const int P = N + 2, C = N * N;
    g->N = N; g->variant = variant;
    const int E = empty_code(g);
    g->empty_fields = 0;
    for (int c = 0; c < C; ++c) {
        g->board[c] = board ? board[c] : E;
        g->empty_fields += (g->board[c] == E);
    }
    memset(g->regions, 0, sizeof(g->regions));
    for (int i = 0; i < P; ++i) g->regions[WHITE][i * P + 0] = 1;
    for (int i = 0; i < P; ++i) g->regions[BLACK][0 * P + i] = 1;
    for (int i = 0; i < P; ++i) g->regions[WHITE][i * P + (N + 1)] = 2;
    for (int i = 0; i < P; ++i) g->regions[BLACK][(N + 1) * P + i] = 2;
    g->counter[BLACK] = 3; /* max(plane) + 1 */
    g->counter[WHITE] = 3;
    if (board) {
        const int black_code = variant == 0 ? 0 : -1, white_code = 1;
        for (int y = 0; y < N; ++y)
            for (int x = 0; x < N; ++x) {
                int v = board[y * N + x];
                if (v == black_code) { g->cur = BLACK; flood_fill(g, y, x); }
                else if (v == white_code) { g->cur = WHITE; flood_fill(g, y, x); }
            }
    }
    g->cur = cur;
    g->done = 0;
    g->winner = NONE;
}

/* HexGame.__init__ with connected_stones given  A: HexGame.py:46-51,63-68   B: HexSingleGame.py:50-55,67-71: the planes are
 * adopted as they are, region_counter = max(plane) + 1 per colour, no flood_fill runs. This is what HexEnv.reset does from its
 * second call on with the planes it cached at the first (HexGame.py:214-220 / HexSingleGame.py:226-231). */
static void game_init_adopt(game_t *g, int N, int variant, int cur, const int *board, const uint8_t *planes /* [2][(N+2)^2] */) {
    const int P = N + 2, C = N * N;
    g->N = N; g->variant = variant;
    const int E = empty_code(g);
    g->empty_fields = 0;
    for (int c = 0; c < C; ++c) {
        g->board[c] = board[c];
        g->empty_fields += (g->board[c] == E);
    }
    memset(g->regions, 0, sizeof(g->regions));
    for (int p = 0; p < 2; ++p) {
        int m = 0;
        for (int i = 0; i < P * P; ++i) {
            g->regions[p][i] = planes[p * P * P + i];
            if (g->regions[p][i] > m) m = g->regions[p][i];
        }
        g->counter[p] = m + 1;
    }
    g->cur = cur;
    g->done = 0;
    g->winner = NONE;
}

/* HexGame.fast_move  A: HexGame.py:85-111   B: HexSingleGame.py:88-122.
 * Returns NONE, BLACK, WHITE or INVALID. An out-of-range action (IndexError / negative wrap in the
 * reference, i.e. undefined) is treated as INVALID. */
static int fast_move(game_t *g, int action) {
    const int N = g->N, C = N * N;
    if (action < 0 || action >= C) return INVALID;
    const int y = action / N, x = action - N * y; /* action_to_coordinate */
    if (g->board[y * N + x] != empty_code(g)) return INVALID; /* is_valid_move */
    if (g->variant == 0) {
        g->board[y * N + x] = g->cur;
        g->empty_fields -= x; /* sic, HexGame.py:96 */
        flood_fill(g, y, x);
    } else {
        g->board[y * N + x] = -1; /* mover always writes its own code into its own perspective */
        g->empty_fields -= 1;
        if (g->cur == WHITE) flood_fill(g, x, y); /* HexSingleGame.py:103-104 "switch" */
        else flood_fill(g, y, x);
    }
    int winner = NONE;
    const int P = N + 2;
    if (g->regions[g->cur][P * P - 1] == 1) { /* regions[-1, -1] == 1 */
        g->done = 1;
        winner = g->cur;
        g->winner = winner;
    } else if (g->variant == 1 && g->empty_fields <= 0) { /* HexSingleGame.py:117-119 */
        g->done = 1;
    }
    g->cur = (g->cur + 1) % 2;
    return winner;
}

/* HexEnv.invert_board  A: HexGame.py:297-303 (transpose, 0<->1)   B: HexSingleGame.py:265-271 (transpose, -1<->+1) */
static void invert_board(game_t *g) {
    const int N = g->N;
    int t[MAXC];
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; ++x) {
            int v = g->board[x * N + y];
            if (g->variant == 0) v = (v == 0) ? 1 : (v == 1) ? 0 : v;
            else v = -v;
            t[y * N + x] = v;
        }
    memcpy(g->board, t, sizeof(int) * N * N);
}

/* k-th cell equal to the empty code in row-major order of the CURRENT board array:
 *   BaseRandomPolicy.choose_action SelfplayWrapper.py:17-22 (== 0)   random_policy minihex/__init__.py:8-12 (== 2) */
static int random_choice(const game_t *g, double u) {
    const int C = g->N * g->N, E = empty_code(g);
    int n = 0;
    for (int c = 0; c < C; ++c) n += (g->board[c] == E);
    if (n == 0) return -1; /* reference: IndexError */
    int choice = (int)(u * (double)n);
    for (int c = 0; c < C; ++c)
        if (g->board[c] == E && choice-- == 0) return c;
    return -1;
}

static inline int transpose_action(int a, int N) { return (a % N) * N + a / N; }

/* ------------------------------------------------------------------ environments */
typedef struct {
    game_t g;
    rng_t rng;
    int kind;           /* 0 = variant-A HexEnv (HexGame.py:145-371), 1 = variant-B SelfPlayEnv, 2 = raw HexGame (A), 3 = raw HexGame (B) */
    int agent;          /* A: self.player (BLACK only)   B: self.agent_player_num (-1 = None) */
    int start_player;   /* A: ctor current_player_num (WHITE => opponent opens, HexGame.py:224-230) */
    int env_cur;        /* B: HexEnv.current_player_num (HexSingleGame.py:209,259) */
    int env_winner;     /* self.winner at env level */
    int eval_state;     /* SelfplayWrapper.py:92 */
    int eval_episode;   /* SelfplayWrapper.py:66,94-95: episodes started since set_eval */
    int64_t st[8];      /* episodes, black wins, white wins, agent wins, plies of finished episodes, invalid ends, env steps, plies */
    int plies;          /* plies in the running episode */
    int manual;         /* the opponent's moves come from the caller (an OpponentPolicy, SelfplayWrapper.py:26-35): resets do not open */
    int pool_size;      /* len(self.opponent_models) */
    int opp_index;      /* opponent chosen by setup_opponents: -1 = best_model, k = opponent_models[k] */
    double opp_eps;     /* variant A, caller-driven opponent: HexEnv.eps of opponent_predict (HexGame.py:354-359); < 0 = plain caller moves */
    int last_opp;       /* the opponent's latest move as HexEnv reports it (A: true cell, HexGame.py:341-348; B: its own view) */
    int info_opp, info_winner; /* info["last_move_opponent"], env.winner at the end of the latest step() (before an auto-reset) */
} env_t;

/* --- variant A */
/* HexEnv.opponent_move HexGame.py:332-349 with opponent_policy = minihex.random_policy, self.player == BLACK */
static void A_opponent_move(env_t *e, double u) {
    invert_board(&e->g);
    int a = random_choice(&e->g, u);
    invert_board(&e->g);
    a = transpose_action(a, e->g.N);
    e->env_winner = fast_move(&e->g, a);
    e->last_opp = a;
    e->plies++; e->st[7]++;
}

/* HexEnv.reset HexGame.py:206-242 */
static void A_reset(env_t *e, const double *open_u) {
    game_init(&e->g, e->g.N, 0, e->start_player, NULL);
    e->plies = 0;
    if (e->agent != e->start_player && !e->manual) A_opponent_move(e, open_u ? *open_u : rng_random(&e->rng));
}

/* HexEnv.step HexGame.py:244-295 (self.player == BLACK) */
static float A_step(env_t *e, int action, const double *opp_u) {
    if (!e->g.done) {
        e->env_winner = fast_move(&e->g, action);
        if (e->env_winner == INVALID) e->g.done = 1;
        else { e->plies++; e->st[7]++; }
    }
    if (!e->g.done) A_opponent_move(e, opp_u ? *opp_u : rng_random(&e->rng));
    if (e->env_winner == e->agent) return 1.f;
    if (e->env_winner == (e->agent + 1) % 2) return -1.f;
    if (e->env_winner == INVALID) return -100.f;
    return 0.f;
}

/* --- variant B */
/* HexEnv.step HexSingleGame.py:233-263: one ply, 2-vector reward, unconditional invert */
static void B_base_step(env_t *e, int action, int reward[2]) {
    e->env_winner = fast_move(&e->g, action);
    if (e->env_winner == INVALID) e->g.done = 1;
    else { e->plies++; e->st[7]++; }
    int r = 0;
    if (e->env_winner == e->env_cur) r = 1;
    else if (e->env_winner == (e->env_cur + 1) % 2) r = -1;
    reward[0] = -r; reward[1] = -r;
    reward[e->env_cur] = r;
    e->env_cur = (e->env_cur + 1) % 2;
    invert_board(&e->g);
}

/* SelfPlayEnv.continue_game SelfplayWrapper.py:146-172 with BaseRandomPolicy */
static void B_continue_game(env_t *e, const double *u_in, int reward[2]) {
    double u;
    if (u_in) u = *u_in;
    else {
        (void)rng_uniform01(&e->rng); /* rv = random.uniform(0,1), unused (:159) */
        u = rng_random(&e->rng);      /* BaseRandomPolicy.choose_action (:20) */
    }
    int a = random_choice(&e->g, u);
    e->last_opp = a;
    B_base_step(e, a, reward);
}

/* SelfPlayEnv.setup_opponents SelfplayWrapper.py:91-104 (with BaseRandomPolicy entries only the draws matter; with a caller-driven
 * opponent opp_index names the entry that plays). Evaluation (:92-96): episode k since set_eval meets opponent_models[k] while
 * k <= len - 1, later episodes keep the opponent they have; nothing is drawn. */
static void B_setup_opponents(env_t *e) {
    if (e->eval_state) {
        if (e->eval_episode <= e->pool_size - 1) {
            e->opp_index = e->eval_episode;
            e->eval_episode += 1;
        }
        return;
    }
    double rv = rng_uniform01(&e->rng);
    e->opp_index = -1;                       /* self.opponent_model = self.best_model */
    if (!(rv < 0.8)) {
        double ui = rng_random(&e->rng);     /* i = int(random.random() * len(self.opponent_models)) */
        if (e->pool_size > 0) e->opp_index = (int)(ui * (double)e->pool_size);
    }
}

/* SelfPlayEnv.reset SelfplayWrapper.py:69-89 + HexEnv.reset HexSingleGame.py:208-231 */
static void B_reset(env_t *e, const double *open_u) {
    e->env_cur = BLACK;
    game_init(&e->g, e->g.N, 1, BLACK, NULL);
    e->plies = 0;
    if (e->agent < 0) e->agent = (int)(rng_random(&e->rng) * 2.0); /* random.randint(0,1), once per env */
    if (!open_u) B_setup_opponents(e);
    if (e->env_cur != e->agent && !e->manual) {
        int reward[2];
        B_continue_game(e, open_u, reward);
    }
}

/* SelfPlayEnv.step SelfplayWrapper.py:174-199 */
static float B_step(env_t *e, int action, const double *opp_u) {
    int reward[2];
    B_base_step(e, action, reward);
    if (!e->g.done) B_continue_game(e, opp_u, reward);
    return (float)reward[e->agent];
}

/* ------------------------------------------------------------------ batch driver (ctypes entry points) */
typedef struct {
    int kind, N;
    int64_t G;
    env_t *envs;
} batch_t;

void *hexref_batch_create(int kind, int N, int64_t G, uint64_t seed, int64_t game_offset, int agent_mode /*0,1, 2=random(B)*/,
                          int opponent_first /*A*/, int eval_state /*B*/) {
    if (N < 2 || N > MAXN || G < 1) return NULL;
    batch_t *b = (batch_t *)calloc(1, sizeof(batch_t));
    b->kind = kind; b->N = N; b->G = G;
    b->envs = (env_t *)calloc((size_t)G, sizeof(env_t));
    for (int64_t i = 0; i < G; ++i) {
        env_t *e = &b->envs[i];
        e->kind = kind;
        e->g.N = N;
        e->rng.seed = seed; e->rng.game = (uint64_t)(game_offset + i); e->rng.idx = 0; e->rng.inject = NULL;
        e->eval_state = eval_state;
        e->opp_eps = -1.0;
        e->env_winner = NONE;
        if (kind == 0) { e->agent = BLACK; e->start_player = opponent_first ? WHITE : BLACK; }
        else if (kind == 1) { e->agent = agent_mode == 2 ? -1 : agent_mode; }
        if (kind >= 2) game_init(&e->g, N, kind - 2, BLACK, NULL);
    }
    return b;
}

void hexref_batch_destroy(void *h) {
    batch_t *b = (batch_t *)h;
    if (!b) return;
    free(b->envs);
    free(b);
}

static void emit_obs_mask(const env_t *e, int8_t *obs, uint8_t *mask) {
    const int C = e->g.N * e->g.N, E = empty_code(&e->g);
    for (int c = 0; c < C; ++c) {
        if (obs) obs[c] = (int8_t)e->g.board[c];              /* the live simulator.board */
        if (mask) mask[c] = (uint8_t)(e->g.board[c] == E);    /* get_action_mask HexGame.py:203-204 / legal_actions HexSingleGame.py:205-206 */
    }
}

static void env_reset(env_t *e, const double *open_u) {
    if (e->kind == 0) A_reset(e, open_u); else B_reset(e, open_u);
}

void hexref_batch_reset(void *h, const uint8_t *reset_mask, const double *open_u, int8_t *obs, uint8_t *mask) {
    batch_t *b = (batch_t *)h;
    const int C = b->N * b->N;
    for (int64_t i = 0; i < b->G; ++i) {
        env_t *e = &b->envs[i];
        if (!reset_mask || reset_mask[i]) {
            int64_t plies = e->st[7]; /* st[7] counts plies played inside step() (incl. auto-reset openings) only */
            env_reset(e, open_u ? &open_u[i] : NULL);
            e->st[7] = plies;
        }
        emit_obs_mask(e, obs ? obs + i * C : NULL, mask ? mask + i * C : NULL);
    }
}

static void account_episode(env_t *e) {
    e->st[0]++;
    if (e->g.winner == BLACK) e->st[1]++;
    if (e->g.winner == WHITE) e->st[2]++;
    if (e->g.winner == e->agent) e->st[3]++;
    e->st[4] += e->plies;
    if (e->env_winner == INVALID) e->st[5]++;
}

/* One env step for every game, driven the way the reference loop drives a single env:
 *   mask = env.legal_actions(); a = BaseRandomPolicy().choose_action(obs)   (only when actions == NULL; one draw)
 *   obs, r, done = env.step(a);  if done and auto_reset: obs = env.reset()  (DummyVecEnv semantics)
 * opp_u: optional [G,2] doubles replacing the stream: [g,0] the opponent's reply, [g,1] its opening move after an auto-reset. */
typedef struct {
    batch_t *b;
    int64_t lo, hi;
    const int32_t *actions; const double *opp_u; int auto_reset;
    int8_t *obs; uint8_t *mask; float *reward; uint8_t *done; int8_t *term_obs; int32_t *actions_out;
} step_job_t;

static void *step_range(void *arg) {
    step_job_t *j = (step_job_t *)arg;
    batch_t *b = j->b;
    const int C = b->N * b->N;
    for (int64_t i = j->lo; i < j->hi; ++i) {
        env_t *e = &b->envs[i];
        int was_done = e->g.done;
        int a;
        if (j->actions) a = j->actions[i];
        else a = was_done ? 0 : random_choice(&e->g, rng_random(&e->rng));
        if (j->actions_out) j->actions_out[i] = was_done ? -1 : a; /* nothing is played in a finished game */
        const double *ou = j->opp_u ? &j->opp_u[2 * i] : NULL;
        float r;
        e->last_opp = -1;
        if (was_done && e->kind == 1) r = 0.f; /* stepping a finished variant-B game: defined as a no-op here (the reference has no guard) */
        else r = e->kind == 0 ? A_step(e, a, ou) : B_step(e, a, ou);
        if (!was_done) e->st[6]++;
        if (j->reward) j->reward[i] = r;
        if (j->done) j->done[i] = (uint8_t)e->g.done;
        e->info_opp = e->last_opp;
        e->info_winner = e->env_winner;
        if (e->g.done && !was_done) account_episode(e);
        if (e->g.done && j->term_obs) emit_obs_mask(e, j->term_obs + i * C, NULL);
        if (e->g.done && j->auto_reset) env_reset(e, ou ? ou + 1 : NULL);
        emit_obs_mask(e, j->obs ? j->obs + i * C : NULL, j->mask ? j->mask + i * C : NULL);
    }
    return NULL;
}

static int g_threads = 1;
void hexref_set_threads(int n) { g_threads = n < 1 ? 1 : (n > 1024 ? 1024 : n); }

void hexref_batch_step(void *h, const int32_t *actions, const double *opp_u, int auto_reset, int8_t *obs, uint8_t *mask,
                       float *reward, uint8_t *done, int8_t *term_obs, int32_t *actions_out) {
    batch_t *b = (batch_t *)h;
    int nt = g_threads;
    if ((int64_t)nt > b->G) nt = (int)b->G;
    step_job_t jobs[1024];
    pthread_t tids[1024];
    for (int t = 0; t < nt; ++t) {
        step_job_t j = {b, b->G * t / nt, b->G * (t + 1) / nt, actions, opp_u, auto_reset, obs, mask, reward, done, term_obs, actions_out};
        jobs[t] = j;
    }
    if (nt == 1) { step_range(&jobs[0]); return; }
    for (int t = 0; t < nt; ++t) pthread_create(&tids[t], NULL, step_range, &jobs[t]);
    for (int t = 0; t < nt; ++t) pthread_join(tids[t], NULL);
}

/* Caller-driven opponent: the env with an OpponentPolicy whose actions arrive from outside. */
void hexref_batch_set_manual(void *h, int pool_size) {
    batch_t *b = (batch_t *)h;
    for (int64_t i = 0; i < b->G; ++i) { b->envs[i].manual = 1; b->envs[i].pool_size = pool_size; b->envs[i].opp_index = -1; b->envs[i].opp_eps = -1.0; }
}

/* HexEnv(opponent_policy="opponent_predict", eps=...) HexGame.py:165-167,180: the opponent's half steps mix random_policy in */
void hexref_batch_set_opponent_eps(void *h, double eps) {
    batch_t *b = (batch_t *)h;
    for (int64_t i = 0; i < b->G; ++i) b->envs[i].opp_eps = eps;
}

/* SelfPlayEnv.set_eval SelfplayWrapper.py:117-120: eval_episode = 0, eval_state = the argument; the running episode goes on */
void hexref_batch_set_eval(void *h, int eval_state) {
    batch_t *b = (batch_t *)h;
    for (int64_t i = 0; i < b->G; ++i) { b->envs[i].eval_episode = 0; b->envs[i].eval_state = eval_state ? 1 : 0; }
}

static int agent_to_move(const env_t *e) { return e->kind == 0 ? (e->g.cur == e->agent) : (e->env_cur == e->agent); }

/* One ply of `side` (0 agent, 1 opponent) in every unfinished game whose turn it is; actions in the mover's own perspective.
 *   variant B: HexEnv.step (HexSingleGame.py:233-263), for the opponent preceded by continue_game's unused draw (SelfplayWrapper.py:159)
 *   variant A: HexEnv.step's agent part (HexGame.py:250-253) / opponent_move with the action transposed back (:341-346)
 * An illegal OPPONENT move ends the episode with reward 0 in both variants (variant A's reference would hand out -100 and play on;
 * unreachable with a masked policy, not replicated). */
void hexref_batch_half_step(void *h, int side, const int32_t *actions, int auto_reset, float *reward, uint8_t *done, uint8_t *to_move,
                            int32_t *opp_index, int8_t *term_obs) {
    batch_t *b = (batch_t *)h;
    const int C = b->N * b->N;
    for (int64_t i = 0; i < b->G; ++i) {
        env_t *e = &b->envs[i];
        const int was_done = e->g.done;
        float r = 0.f;
        if (!was_done && ((side == 0) == (agent_to_move(e) != 0))) {
            if (side == 0) e->st[6]++;
            if (side == 1 && e->kind == 1) (void)rng_uniform01(&e->rng);   /* continue_game's unused draw (:159) */
            int a;
            int from_caller = actions != NULL;
            if (from_caller && side == 1 && e->kind == 0 && e->opp_eps >= 0.0) {
                /* HexEnv.opponent_predict HexGame.py:354-359: rv = random.uniform(0,1); rv < eps -> random_policy(state), else the model */
                double rv = rng_uniform01(&e->rng);
                if (rv < e->opp_eps) from_caller = 0;
            }
            if (from_caller) a = actions[i];
            else if (e->kind == 1) a = random_choice(&e->g, rng_random(&e->rng));   /* BaseRandomPolicy on the mover's view */
            else {                                                                   /* random_policy on the inverted board */
                invert_board(&e->g);
                a = random_choice(&e->g, rng_random(&e->rng));
                invert_board(&e->g);
            }
            if (e->kind == 1) {
                int rw[2];
                B_base_step(e, a, rw);
                r = (float)rw[e->agent];
            } else {
                if (side == 1) a = (a >= 0 && a < C) ? transpose_action(a, b->N) : -1;
                e->env_winner = fast_move(&e->g, a);
                if (e->env_winner == INVALID) e->g.done = 1;
                else { e->plies++; e->st[7]++; }
                if (e->env_winner == e->agent) r = 1.f;
                else if (e->env_winner == (e->agent + 1) % 2) r = -1.f;
                else if (e->env_winner == INVALID && side == 0) r = -100.f;
            }
        } else if (was_done && e->kind == 0 && side == 0) {
            if (e->env_winner == e->agent) r = 1.f;
            else if (e->env_winner == (e->agent + 1) % 2) r = -1.f;
            else if (e->env_winner == INVALID) r = -100.f;
        }
        if (reward) reward[i] = r;
        if (done) done[i] = (uint8_t)e->g.done;
        if (e->g.done && !was_done) {
            account_episode(e);
            if (term_obs) emit_obs_mask(e, term_obs + i * C, NULL);
            if (auto_reset) env_reset(e, NULL);
        }
        if (to_move) to_move[i] = e->g.done ? 2 : (agent_to_move(e) ? 0 : 1);
        if (opp_index) opp_index[i] = e->opp_index;
    }
}

void hexref_batch_info(void *h, int32_t *last_move_opponent, int8_t *winner) {
    batch_t *b = (batch_t *)h;
    for (int64_t i = 0; i < b->G; ++i) {
        if (last_move_opponent) last_move_opponent[i] = b->envs[i].info_opp;
        if (winner) winner[i] = (int8_t)b->envs[i].info_winner;
    }
}

void hexref_batch_opp_state(void *h, uint8_t *to_move, int32_t *opp_index) {
    batch_t *b = (batch_t *)h;
    for (int64_t i = 0; i < b->G; ++i) {
        const env_t *e = &b->envs[i];
        if (to_move) to_move[i] = e->g.done ? 2 : (agent_to_move(e) ? 0 : 1);
        if (opp_index) opp_index[i] = e->opp_index;
    }
}

/* Current board (the live simulator.board, i.e. the side-to-move view in variant B) + mask, without stepping. */
void hexref_batch_observe(void *h, int8_t *obs, uint8_t *mask) {
    batch_t *b = (batch_t *)h;
    const int C = b->N * b->N;
    for (int64_t i = 0; i < b->G; ++i) emit_obs_mask(&b->envs[i], obs ? obs + i * C : NULL, mask ? mask + i * C : NULL);
}

/* The board and mask the OPPONENT's policy is shown when it is to move (what hexb_encode(view 1) emits): variant B = the live board
 * (already the mover's view); variant A = the board between the two invert_board calls of HexEnv.opponent_move (HexGame.py:333-339:
 * transposed, colours swapped) and get_action_mask of that array (:358). Games with the agent to move: the live board. */
void hexref_batch_observe_opponent(void *h, int8_t *obs, uint8_t *mask) {
    batch_t *b = (batch_t *)h;
    const int C = b->N * b->N;
    for (int64_t i = 0; i < b->G; ++i) {
        env_t *e = &b->envs[i];
        const int inv = e->kind == 0 && !e->g.done && !agent_to_move(e);
        if (inv) invert_board(&e->g);
        emit_obs_mask(e, obs ? obs + i * C : NULL, mask ? mask + i * C : NULL);
        if (inv) invert_board(&e->g);
    }
}

/* Batched HexGame.make_move on raw games (kind 2 = A, kind 3 = B), or on the simulators inside envs. */
void hexref_batch_ply(void *h, const int32_t *actions, int8_t *ret) {
    batch_t *b = (batch_t *)h;
    for (int64_t i = 0; i < b->G; ++i) {
        int w = fast_move(&b->envs[i].g, actions[i]);
        if (ret) ret[i] = (int8_t)w;
    }
}

/* Reference-layout dump: board f64[G,N,N] (the live simulator.board), regions f64[G,2,N+2,N+2], region_counter f64[G,2]. */
void hexref_batch_export(void *h, double *board, double *regions, double *counter, int8_t *cur, uint8_t *done, int8_t *winner,
                         int8_t *agent, uint32_t *draws) {
    batch_t *b = (batch_t *)h;
    const int N = b->N, C = N * N, P2 = (N + 2) * (N + 2);
    for (int64_t i = 0; i < b->G; ++i) {
        const env_t *e = &b->envs[i];
        if (board) for (int c = 0; c < C; ++c) board[i * C + c] = (double)e->g.board[c];
        if (regions)
            for (int p = 0; p < 2; ++p)
                for (int c = 0; c < P2; ++c) regions[(i * 2 + p) * P2 + c] = (double)e->g.regions[p][c];
        if (counter) { counter[2 * i] = e->g.counter[0]; counter[2 * i + 1] = e->g.counter[1]; }
        if (cur) cur[i] = (int8_t)e->g.cur;
        if (done) done[i] = (uint8_t)e->g.done;
        if (winner) winner[i] = (int8_t)e->g.winner;
        if (agent) agent[i] = (int8_t)e->agent;
        if (draws) draws[i] = e->rng.idx;
    }
}

void hexref_batch_stats(void *h, int64_t out[8]) {
    batch_t *b = (batch_t *)h;
    for (int k = 0; k < 8; ++k) out[k] = 0;
    for (int64_t i = 0; i < b->G; ++i)
        for (int k = 0; k < 8; ++k) out[k] += b->envs[i].st[k];
}

/* HexEnv.reset with sample_board=True (HexSingleGame.py:217-222) for the masked envs: a new simulator on the given board
 * (true coordinates; B codes -1 BLACK / +1 WHITE / 0 empty, A codes 0/1/2), BLACK to move. The agent colour and the random
 * stream carry on; the opponent's catch-up move (SelfPlayEnv.reset -> continue_game) is a separate half step. */
void hexref_batch_env_set_board(void *h, const int8_t *boards, const uint8_t *mask) {
    batch_t *b = (batch_t *)h;
    const int C = b->N * b->N;
    for (int64_t i = 0; i < b->G; ++i) {
        if (mask && !mask[i]) continue;
        env_t *e = &b->envs[i];
        int tmp[MAXC], stones = 0;
        const int E = e->kind == 1 ? 0 : 2;
        for (int c = 0; c < C; ++c) { tmp[c] = boards[i * C + c]; stones += (tmp[c] != E); }
        game_init(&e->g, b->N, e->kind == 1 ? 1 : 0, BLACK, tmp);
        e->env_cur = BLACK;
        e->plies = stones;
        e->env_winner = NONE;
    }
}

/* Raw game from a preset board and its label planes (HexGame.__init__ with connected_stones given). */
void hexref_batch_set_board_labels(void *h, const int8_t *boards, const uint8_t *planes /*[G,2,N+2,N+2]*/, int cur) {
    batch_t *b = (batch_t *)h;
    const int C = b->N * b->N, P2 = (b->N + 2) * (b->N + 2);
    for (int64_t i = 0; i < b->G; ++i) {
        int tmp[MAXC];
        for (int c = 0; c < C; ++c) tmp[c] = boards[i * C + c];
        game_init_adopt(&b->envs[i].g, b->N, b->envs[i].g.variant, cur, tmp, planes + i * 2 * P2);
    }
}

/* Raw game from a preset board (HexGame.__init__ with connected_stones=None: raster-order flood_fill rebuild). */
void hexref_batch_set_board(void *h, const int8_t *boards /*[G,N,N] in the variant's encoding*/, int cur) {
    batch_t *b = (batch_t *)h;
    const int C = b->N * b->N;
    for (int64_t i = 0; i < b->G; ++i) {
        int tmp[MAXC];
        for (int c = 0; c < C; ++c) tmp[c] = boards[i * C + c];
        game_init(&b->envs[i].g, b->N, b->envs[i].g.variant, cur, tmp);
    }
}
