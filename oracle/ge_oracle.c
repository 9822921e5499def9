/*
 * ge_oracle.c — Oracle B: scalar CPU restatement of the referee/phase step.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product path (game_engine_b200/csrc) never links or
 * calls it and has no CPU fallback.
 *
 * What it restates (paths under /root/reference/):
 *   - BotBehaviorNode      agent/game_agent_v2.py:468-617  (+ prompt/bot_behavior_system_prompt.txt)
 *   - PhaseNode            agent/game_agent_v2.py:987-1241 (+ prompt/PhaseNode_system_prompt.txt)
 *   - RefereeNode          agent/game_agent_v2.py:619-803  (+ prompt/referee_system_prompt_{1,2}.txt)
 *   - _execute_update_player_state  agent/tools/backend_tools.py:204-225 (field writes)
 *   - phase graphs         games/werewolf-(mafia).yaml:166-666, games/two-truths-and-a-lie.yaml:145-403,
 *                          game_draft/werewolf-(mafia).yaml:146-430 (third table)
 * with every decision the reference leaves to its LLM frozen by /root/repo/SPEC.md.
 *
 * PARITY STATUS: the reference has no executable tests, golden vectors or fixtures for this path
 * (SURVEY.md section 4), and its decisions come from a remote LLM, so game-rule parity is UNPINNED by
 * the reference's own tests.  What IS pinned: this file must equal, step for step, the reference's
 * real node code driven by a rule-following stub LLM (oracle/ref_harness, fixtures in tests/golden/),
 * and the Philox4x32-10 known-answer vectors of Random123.
 *
 * Style: deliberately naive — state is unpacked into one array entry per player and every rule is a
 * loop over players.  Nothing here is shared with the CUDA implementation.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GE_STATS_LEN 560
#define MAXP 32

enum { FAM_WEREWOLF = 1, FAM_TTL = 2 };
enum { KIND_UI = 0, KIND_TIMER = 1, KIND_ACTION = 2, KIND_TERMINAL = 3 };
enum { ACT_NONE = 0, ACT_PICK_PLAYER = 1, ACT_PICK_OPTION = 2, ACT_MARK = 3 };
enum { EX_NONE = 0, EX_VOTE_KILL = 1, EX_PROTECT = 2, EX_INVESTIGATE_RESOLVE = 3, EX_DAY_VOTE = 4,
       EX_T_STATEMENTS = 16, EX_T_LIE = 17, EX_T_VOTES = 18 };
enum { EN_NONE = 0, EN_ASSIGN_ROLES = 1, EN_NIGHT_RESET = 2,
       EN_T_ROUND_START = 16, EN_T_REVEAL = 17, EN_T_SCORE = 18, EN_T_FINAL = 19 };
enum { BR_ALWAYS = 0, BR_COUNT_EQ0 = 1, BR_COUNT_GE = 2, BR_PREV_IN = 3, BR_ALL_VAL_GE = 4, BR_TIE_PENDING = 5 };

/* ---- table blob accessors (format: game_engine_b200/table.py) ---- */
typedef struct {
    const uint8_t *blob;
    int family, n_phases, P, n_preds, n_wolves, rounds, max_revotes;
    uint32_t init_masks;
} tab_t;

static uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static void wr32(uint8_t *p, uint32_t v) { p[0] = v; p[1] = v >> 8; p[2] = v >> 16; p[3] = v >> 24; }
static void wr16(uint8_t *p, uint16_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }

static int wolfy_counts_bad(const void *tv);
static int tab_open(tab_t *t, const uint8_t *blob, size_t n) {
    if (n < 32 || memcmp(blob, "GETB", 4) != 0 || rd16(blob + 4) != 1) return -1;
    t->blob = blob;
    t->family = blob[6]; t->n_phases = blob[7]; t->P = blob[8]; t->n_preds = blob[9];
    t->n_wolves = blob[10]; t->rounds = blob[11]; t->max_revotes = blob[12];
    t->init_masks = rd32(blob + 16);
    if (n < (size_t)(32 + 48 * t->n_phases + 8 * t->n_preds)) return -1;
    if (t->P < 2 || t->P > MAXP || t->n_phases < 1 || t->n_phases > 32) return -1;
    if (t->family != FAM_WEREWOLF && t->family != FAM_TTL) return -1;
    /* well-formedness (SPEC.md section 7), the same rules as the CUDA library's ge_table_create */
    for (int i = 0; i < t->n_phases; ++i) {
        const uint8_t *ph = blob + 32 + 48 * i;
        const int kind = ph[1], aop = ph[2], exo = ph[5], eno = ph[6], nbr = ph[7], wolfy = t->family == FAM_WEREWOLF;
        if (kind > KIND_TERMINAL || nbr > 4 || (kind != KIND_TERMINAL && nbr == 0)) return -1;
        if (exo != EX_NONE && kind != KIND_ACTION) return -1;
        if (wolfy ? exo > EX_DAY_VOTE : (exo != EX_NONE && (exo < EX_T_STATEMENTS || exo > EX_T_VOTES))) return -1;
        if (wolfy ? eno > EN_NIGHT_RESET : (eno != EN_NONE && (eno < EN_T_ROUND_START || eno > EN_T_FINAL))) return -1;
        if (kind == KIND_ACTION) {
            if (aop < ACT_PICK_PLAYER || aop > ACT_MARK || ph[8] >= t->n_preds) return -1;
            if (aop == ACT_PICK_PLAYER && ph[3] >= t->n_preds) return -1;
            if (aop == ACT_PICK_OPTION && ph[3] == 0) return -1;
            if (wolfy && exo != EX_NONE && aop != ACT_PICK_PLAYER) return -1;
        }
        for (int b = 0; b < nbr; ++b) {
            const uint8_t *br = ph + 16 + 8 * b;
            if (br[0] > BR_TIE_PENDING || br[1] >= t->n_phases) return -1;
            if ((br[0] == BR_COUNT_EQ0 || br[0] == BR_COUNT_GE) && br[3] >= t->n_preds) return -1;
            if (br[0] == BR_COUNT_GE && rd32(br + 4) >= (uint32_t)t->n_preds) return -1;
            if (br[0] == BR_ALL_VAL_GE && br[3] > 2) return -1;
            if (wolfy ? (br[0] == BR_ALL_VAL_GE || br[2] > 2) : br[0] == BR_TIE_PENDING) return -1;
        }
    }
    /* comparison fields: count at byte 13, four {value field, op, constant} triples at 20; the rest of the header zero */
    const int n_cmp = blob[13], wolfy_t = t->family == FAM_WEREWOLF;
    if (n_cmp > (wolfy_t ? 2 : 4) || blob[14] || blob[15]) return -1;
    for (int k = 0; k < 4; ++k) {
        const uint8_t *c = blob + 20 + 3 * k;
        if (k >= n_cmp) { if (c[0] | c[1] | c[2]) return -1; continue; }
        if (c[1] > 5 || c[0] > (wolfy_t ? 0 : 2)) return -1;
    }
    /* predicates may only name mask fields the table defines (SPEC.md section 2); the last record cannot be continued */
    unsigned defined = wolfy_t ? 0x9FFFu : 0x801Fu;
    for (int k = 0; k < n_cmp; ++k) defined |= 1u << ((wolfy_t ? 13 : 11) + k);
    for (int i = 0; i < t->n_preds; ++i) {
        const uint8_t *pr = blob + 32 + 48 * t->n_phases + 8 * i;
        const unsigned used = rd16(pr) | rd16(pr + 2) | rd16(pr + 4) | rd16(pr + 6);
        if (used & ~defined) return -1;
        if ((rd16(pr) & 0x8000) && i + 1 >= t->n_preds) return -1;
    }
    if (wolfy_counts_bad(t)) return -1;
    return 0;
}
static int wolfy_counts_bad(const void *tv) {
    const tab_t *t = (const tab_t *)tv;
    return t->family == FAM_WEREWOLF && (t->n_wolves < 1 || t->n_wolves + 2 > t->P);
}
static const uint8_t *tab_phase(const tab_t *t, int i) { return t->blob + 32 + 48 * i; }
static const uint8_t *tab_branch(const tab_t *t, int i, int b) { return tab_phase(t, i) + 16 + 8 * b; }
static const uint8_t *tab_pred(const tab_t *t, int i) { return t->blob + 32 + 48 * t->n_phases + 8 * i; }

static size_t rec_size(const tab_t *t) {
    if (t->family == FAM_WEREWOLF) return 48 + (size_t)((t->P + 7) / 8) * 8;
    return (size_t)((8 + 4 * t->P + 7) / 8) * 8;
}

/* ---- Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11; Random123 philox.h constants) ---- */
void ge_cpu_philox(const uint32_t key[2], const uint32_t ctr[4], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* SPEC section 3: word for 0-based player p0 */
static uint32_t draw(uint64_t seed, uint64_t sid, uint32_t step, uint32_t purpose, int p0) {
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t ctr[4] = { (uint32_t)sid, (uint32_t)(sid >> 32), step, (purpose << 16) | (uint32_t)(p0 >> 2) };
    uint32_t out[4];
    ge_cpu_philox(key, ctr, out);
    return out[p0 & 3];
}
static uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

/* ---- unpacked session ---- */
typedef struct {
    int phase, prev, step;
    /* werewolf */
    int winner, kill_target, protect_target, revote;
    uint8_t alive[MAXP], can_vote[MAXP], eligible[MAXP], submitted[MAXP], revealed[MAXP], investigated[MAXP];
    uint8_t wolf[MAXP], secret[MAXP], role[MAXP], target[MAXP];
    /* ttl */
    int speaker, lie_index;
    uint8_t score[MAXP], rounds_done[MAXP], vote[MAXP];
    uint8_t is_speaker[MAXP], stmts[MAXP], lie_revealed[MAXP], has_voted[MAXP];
} sess_t;

static void bits_get(uint32_t m, uint8_t *a, int P) { for (int p = 0; p < P; ++p) a[p] = (m >> p) & 1; }
static uint32_t bits_put(const uint8_t *a, int P) { uint32_t m = 0; for (int p = 0; p < P; ++p) if (a[p]) m |= 1u << p; return m; }

static void unpack(const tab_t *t, const uint8_t *r, sess_t *s) {
    int P = t->P;
    memset(s, 0, sizeof *s);
    s->phase = r[0]; s->prev = r[1]; s->step = rd16(r + 2);
    if (t->family == FAM_WEREWOLF) {
        s->winner = r[4]; s->kill_target = r[5]; s->protect_target = r[6]; s->revote = r[7];
        bits_get(rd32(r + 8), s->alive, P);      bits_get(rd32(r + 12), s->can_vote, P);
        bits_get(rd32(r + 16), s->eligible, P);  bits_get(rd32(r + 20), s->submitted, P);
        bits_get(rd32(r + 24), s->revealed, P);  bits_get(rd32(r + 28), s->investigated, P);
        bits_get(rd32(r + 32), s->wolf, P);      bits_get(rd32(r + 36), s->secret, P);
        uint32_t lo = rd32(r + 40), hi = rd32(r + 44);
        for (int p = 0; p < P; ++p) { s->role[p] = ((lo >> p) & 1) | (((hi >> p) & 1) << 1); s->target[p] = r[48 + p]; }
    } else {
        s->speaker = r[4]; s->lie_index = r[5]; s->winner = r[6];
        for (int p = 0; p < P; ++p) {
            const uint8_t *q = r + 8 + 4 * p;
            s->score[p] = q[0]; s->rounds_done[p] = q[1]; s->vote[p] = q[2];
            s->is_speaker[p] = q[3] & 1; s->stmts[p] = (q[3] >> 1) & 1; s->lie_revealed[p] = (q[3] >> 2) & 1;
            s->can_vote[p] = (q[3] >> 3) & 1; s->has_voted[p] = (q[3] >> 4) & 1;
        }
    }
}

static void pack(const tab_t *t, const sess_t *s, uint8_t *r) {
    int P = t->P;
    memset(r, 0, rec_size(t));
    r[0] = (uint8_t)s->phase; r[1] = (uint8_t)s->prev; wr16(r + 2, (uint16_t)s->step);
    if (t->family == FAM_WEREWOLF) {
        r[4] = (uint8_t)s->winner; r[5] = (uint8_t)s->kill_target; r[6] = (uint8_t)s->protect_target; r[7] = (uint8_t)s->revote;
        wr32(r + 8, bits_put(s->alive, P));      wr32(r + 12, bits_put(s->can_vote, P));
        wr32(r + 16, bits_put(s->eligible, P));  wr32(r + 20, bits_put(s->submitted, P));
        wr32(r + 24, bits_put(s->revealed, P));  wr32(r + 28, bits_put(s->investigated, P));
        wr32(r + 32, bits_put(s->wolf, P));      wr32(r + 36, bits_put(s->secret, P));
        uint32_t lo = 0, hi = 0;
        for (int p = 0; p < P; ++p) { lo |= (uint32_t)(s->role[p] & 1) << p; hi |= (uint32_t)((s->role[p] >> 1) & 1) << p; r[48 + p] = s->target[p]; }
        wr32(r + 40, lo); wr32(r + 44, hi);
    } else {
        r[4] = (uint8_t)s->speaker; r[5] = (uint8_t)s->lie_index; r[6] = (uint8_t)s->winner;
        for (int p = 0; p < P; ++p) {
            uint8_t *q = r + 8 + 4 * p;
            q[0] = s->score[p]; q[1] = s->rounds_done[p]; q[2] = s->vote[p];
            q[3] = (uint8_t)(s->is_speaker[p] | (s->stmts[p] << 1) | (s->lie_revealed[p] << 2) | (s->can_vote[p] << 3) | (s->has_voted[p] << 4));
        }
    }
}

/* comparison field k of the table: "player p's value field <op> constant" (numeric conditions of the DSL,
 * prompt/dsl_phases_generation_prompt.txt:106-128); header bytes 13 = count, 20 + 3k = {value field, op, constant} */
static int cmp_field(const tab_t *t, const sess_t *s, int k, int p) {
    if (k < 0 || k >= t->blob[13]) return 0;
    const uint8_t *c = t->blob + 20 + 3 * k;
    int v;
    if (t->family == FAM_WEREWOLF) v = s->target[p];
    else v = c[0] == 0 ? s->score[p] : c[0] == 1 ? s->rounds_done[p] : s->vote[p];
    switch (c[1]) {
    case 0: return v == c[2]; case 1: return v != c[2]; case 2: return v < c[2];
    case 3: return v <= c[2]; case 4: return v > c[2];  default: return v >= c[2];
    }
}

/* value of mask field f for player p (SPEC section 2) */
static int field_of(const tab_t *t, const sess_t *s, int f, int p) {
    if (f == 15) return 1;
    if (t->family == FAM_WEREWOLF ? (f == 13 || f == 14) : (f >= 11 && f <= 14)) return cmp_field(t, s, f - (t->family == FAM_WEREWOLF ? 13 : 11), p);
    if (t->family == FAM_WEREWOLF) {
        switch (f) {
        case 0: return s->alive[p];     case 1: return s->can_vote[p];  case 2: return s->eligible[p];
        case 3: return s->submitted[p]; case 4: return s->revealed[p];  case 5: return s->investigated[p];
        case 6: return s->wolf[p];      case 7: return s->secret[p];
        case 8: case 9: case 10: case 11: {          /* a role only counts once roles are assigned */
            int assigned = 0;
            for (int q = 0; q < t->P; ++q) assigned |= s->secret[q];
            return assigned && s->role[p] == f - 8;
        }
        case 12: {
            int assigned = 0;
            for (int q = 0; q < t->P; ++q) assigned |= s->secret[q];
            return assigned;
        }
        default: return 0;
        }
    }
    switch (f) {
    case 0: return s->is_speaker[p]; case 1: return s->stmts[p]; case 2: return s->lie_revealed[p];
    case 3: return s->can_vote[p];   case 4: return s->has_voted[p];
    default: return 0;
    }
}

/* one predicate record: a DNF of two clauses (bit 15 of pos0 is the "continued" flag, not a field) */
static int pred_record_holds(const tab_t *t, const sess_t *s, const uint8_t *q, int p) {
    for (int c = 0; c < 2; ++c) {
        uint16_t pos = rd16(q + 4 * c), neg = rd16(q + 4 * c + 2);
        if (c == 0) pos &= 0x7FFF;
        int ok = 1;
        for (int f = 0; f < 16 && ok; ++f) {
            if (!(((pos | neg) >> f) & 1)) continue;
            const int v = field_of(t, s, f, p);
            if (((pos >> f) & 1) && !v) ok = 0;
            if (((neg >> f) & 1) && v) ok = 0;
        }
        if (ok) return 1;
    }
    return 0;
}
/* a predicate of the table: the record, ORed with the following ones while they are marked continued */
static int pred_holds(const tab_t *t, const sess_t *s, int pred, int p) {
    for (;; ++pred) {
        const uint8_t *q = tab_pred(t, pred);
        if (pred_record_holds(t, s, q, p)) return 1;
        if (!(rd16(q) & 0x8000)) return 0;
    }
}
static int pred_count(const tab_t *t, const sess_t *s, int pred) {
    int n = 0;
    for (int p = 0; p < t->P; ++p) n += pred_holds(t, s, pred, p);
    return n;
}

static void die(sess_t *s, int id) {       /* id is 1-based */
    s->alive[id - 1] = 0; s->can_vote[id - 1] = 0; s->eligible[id - 1] = 0;
}

/* plurality over choices of actors; ties -> lowest id; *tied = more than one candidate on top */
static int plurality(int P, const uint8_t *is_actor, const uint8_t *choice, int *tied) {
    int cnt[MAXP + 1] = {0};
    for (int p = 0; p < P; ++p) if (is_actor[p] && choice[p]) cnt[choice[p]]++;
    int best = 0, best_n = 0, n_top = 0;
    for (int c = 1; c <= P; ++c) {
        if (cnt[c] > best_n) { best = c; best_n = cnt[c]; n_top = 1; }
        else if (cnt[c] == best_n && best_n > 0) n_top++;
    }
    if (tied) *tied = n_top > 1;
    return best;
}

/* one step of one session; returns 1 if the step counted. visits may be NULL.
 * hmask / hchoice: human seats (SPEC D3h) and their inputs for this step (0xFF = has not acted); hmask 0 = all bots.
 * Restates: human player excluded from bot actions (prompt/bot_behavior_system_prompt.txt:3,58-61), PhaseNode staying
 * — and still appending history — until the target players have acted (agent/game_agent_v2.py:1144-1170,1206-1215). */
static int step_session(const tab_t *t, sess_t *s, uint64_t seed, uint64_t sid, uint64_t *visits, uint32_t hmask, const uint8_t *hchoice) {
    const int P = t->P, X = s->phase;
    const uint8_t *ph = tab_phase(t, X);
    if (ph[1] == KIND_TERMINAL) return 0;
    if (s->step == 0) {                       /* D11 */
        s->step = 1;
        if (visits) visits[X]++;
        return 1;
    }
    const int step0 = s->step;
    uint8_t actor[MAXP] = {0}, choice[MAXP] = {0};
    int first_actor = -1;

    /* --- bots act (BotBehaviorNode) --- */
    if (ph[1] == KIND_ACTION) {
        uint8_t legal_ok[MAXP] = {0};      /* legal set on the state at the start of the step */
        for (int p = 0; p < P; ++p) actor[p] = (uint8_t)pred_holds(t, s, ph[8], p);
        if (ph[2] == ACT_PICK_PLAYER)
            for (int q = 0; q < P; ++q) legal_ok[q] = (uint8_t)pred_holds(t, s, ph[3], q);
        /* people first: is every acting human seat answered with something it may choose? */
        int human_val[MAXP];
        int waiting = 0;
        for (int p = 0; p < P; ++p) {
            human_val[p] = -1;
            if (!actor[p] || !((hmask >> p) & 1u)) continue;
            const int c = hchoice[p];
            if (ph[2] == ACT_PICK_PLAYER) {
                int n = 0, ok = 0;
                for (int q = 0; q < P; ++q) {
                    if ((ph[4] & 1) && q == p) continue;
                    if (legal_ok[q]) { ++n; if (c == q + 1) ok = 1; }
                }
                human_val[p] = n == 0 ? 0 : ok ? c : -1;
            } else if (ph[2] == ACT_PICK_OPTION) {
                human_val[p] = (c >= 1 && c <= ph[3]) ? c : -1;
            } else {
                human_val[p] = c != 0xFF ? 1 : -1;
            }
            if (human_val[p] < 0) waiting = 1;
        }
        if (waiting) {                        /* the step stays: history grows, nothing else changes */
            s->prev = X; s->step = (step0 + 1) & 0xFFFF;
            if (visits) visits[X]++;
            return 1;
        }
        for (int p = 0; p < P; ++p) {
            if (!actor[p]) continue;
            if (first_actor < 0) first_actor = p;
            if ((hmask >> p) & 1u) { choice[p] = (uint8_t)human_val[p]; continue; }
            uint32_t r = draw(seed, sid, (uint32_t)step0, 0, p);
            if (ph[2] == ACT_PICK_PLAYER) {
                int legal[MAXP], n = 0;
                for (int q = 0; q < P; ++q) {
                    if ((ph[4] & 1) && q == p) continue;
                    if (legal_ok[q]) legal[n++] = q;
                }
                choice[p] = n ? (uint8_t)(1 + legal[mulhi32(r, (uint32_t)n)]) : 0;
            } else if (ph[2] == ACT_PICK_OPTION) {
                choice[p] = (uint8_t)(1 + mulhi32(r, ph[3]));
            } else if (ph[2] == ACT_MARK) {
                choice[p] = 1;
            }
        }
    }

    /* --- branch select on the state before effects (PhaseNode) --- */
    int nbr = ph[7], taken = nbr - 1;
    for (int b = 0; b < nbr; ++b) {
        const uint8_t *br = tab_branch(t, X, b);
        uint32_t arg = rd32(br + 4);
        int ok = 0;
        switch (br[0]) {
        case BR_ALWAYS: ok = 1; break;
        case BR_COUNT_EQ0: ok = pred_count(t, s, br[3]) == 0; break;
        case BR_COUNT_GE: ok = pred_count(t, s, br[3]) >= pred_count(t, s, (int)arg); break;
        case BR_PREV_IN: ok = (arg >> s->prev) & 1; break;
        case BR_ALL_VAL_GE: {
            ok = 1;
            for (int p = 0; p < P; ++p) {
                int v = br[3] == 0 ? s->score[p] : br[3] == 1 ? s->rounds_done[p] : s->vote[p];
                if (v < (int)arg) ok = 0;
            }
        } break;
        case BR_TIE_PENDING: ok = (s->revote & 0x80) != 0; break;
        }
        if (ok) { taken = b; break; }
    }
    const uint8_t *br = tab_branch(t, X, taken);
    const int Y = br[1], tag = br[2];

    /* --- exit effect of X (RefereeNode, last_phase = X) --- */
    switch (ph[5]) {
    case EX_VOTE_KILL:
    case EX_PROTECT:
    case EX_INVESTIGATE_RESOLVE:
        for (int p = 0; p < P; ++p) if (actor[p]) { s->target[p] = choice[p]; s->submitted[p] = 1; }
        if (ph[5] == EX_VOTE_KILL) s->kill_target = plurality(P, actor, choice, NULL);
        if (ph[5] == EX_PROTECT) s->protect_target = first_actor >= 0 ? choice[first_actor] : 0;
        if (ph[5] == EX_INVESTIGATE_RESOLVE) {
            for (int p = 0; p < P; ++p) if (actor[p] && choice[p]) s->investigated[choice[p] - 1] = 1;
            if (s->kill_target && s->kill_target != s->protect_target) die(s, s->kill_target);
            s->kill_target = 0; s->protect_target = 0;
        }
        break;
    case EX_DAY_VOTE: {
        int tied = 0;
        for (int p = 0; p < P; ++p) if (actor[p]) s->target[p] = choice[p];
        int x = plurality(P, actor, choice, &tied);
        if (t->max_revotes > 0 && tied && (s->revote & 0x7f) < t->max_revotes) {
            s->revote = ((s->revote & 0x7f) + 1) | 0x80;
        } else {
            s->revote &= 0x7f;
            if (x) { die(s, x); s->revealed[x - 1] = 1; }
        }
    } break;
    case EX_T_STATEMENTS: for (int p = 0; p < P; ++p) if (actor[p]) s->stmts[p] = 1; break;
    case EX_T_LIE: if (first_actor >= 0) s->lie_index = choice[first_actor]; break;
    case EX_T_VOTES: for (int p = 0; p < P; ++p) if (actor[p]) { s->vote[p] = choice[p]; s->has_voted[p] = 1; } break;
    default: break;
    }

    /* --- entry effect of Y (RefereeNode, next phase = Y) --- */
    switch (tab_phase(t, Y)[6]) {
    case EN_ASSIGN_ROLES: {
        uint32_t key[MAXP];
        for (int p = 0; p < P; ++p) key[p] = draw(seed, sid, (uint32_t)step0, 1, p);
        for (int p = 0; p < P; ++p) {
            int rank = 0;
            for (int q = 0; q < P; ++q) if (key[q] < key[p] || (key[q] == key[p] && q < p)) rank++;
            int role = rank < t->n_wolves ? 1 : rank == t->n_wolves ? 2 : rank == t->n_wolves + 1 ? 3 : 0;
            s->role[p] = (uint8_t)role; s->wolf[p] = role == 1; s->secret[p] = role != 0; s->eligible[p] = role != 0;
        }
    } break;
    case EN_NIGHT_RESET:
        for (int p = 0; p < P; ++p) { s->submitted[p] = 0; s->target[p] = 0; }
        s->kill_target = 0; s->protect_target = 0; s->revote = 0;
        break;
    case EN_T_ROUND_START: {
        int sp = 0;
        for (int p = P - 1; p >= 0; --p) if (s->rounds_done[p] < t->rounds) sp = p + 1;
        s->speaker = sp; s->lie_index = 0;
        for (int p = 0; p < P; ++p) {
            s->is_speaker[p] = (p + 1 == sp); s->can_vote[p] = (p + 1 != sp);
            s->has_voted[p] = 0; s->stmts[p] = 0; s->lie_revealed[p] = 0; s->vote[p] = 0;
        }
    } break;
    case EN_T_REVEAL: for (int p = 0; p < P; ++p) if (s->is_speaker[p]) s->lie_revealed[p] = 1; break;
    case EN_T_SCORE: {
        int fooled = 0;
        for (int p = 0; p < P; ++p) {
            if (!(s->has_voted[p] && s->can_vote[p] && !s->is_speaker[p])) continue;
            if (s->vote[p] == s->lie_index) s->score[p]++; else fooled++;
        }
        if (s->speaker) { s->score[s->speaker - 1] = (uint8_t)(s->score[s->speaker - 1] + fooled); s->rounds_done[s->speaker - 1]++; }
    } break;
    case EN_T_FINAL: {
        int best = 0;
        for (int p = 1; p < P; ++p) if (s->score[p] > s->score[best]) best = p;
        s->winner = best + 1;
    } break;
    default: break;
    }
    if (tag) s->winner = tag;

    s->prev = X; s->phase = Y; s->step = (step0 + 1) & 0xFFFF;      /* 16-bit counter (SPEC section 7) */
    if (visits) visits[Y]++;
    return 1;
}

/* ------------------------------------------------------------------ exported C entry points */
int ge_cpu_table_check(const uint8_t *blob, size_t n) { tab_t t; return tab_open(&t, blob, n); }

size_t ge_cpu_record_size(const uint8_t *blob, size_t n) { tab_t t; return tab_open(&t, blob, n) ? 0 : rec_size(&t); }

int ge_cpu_init(const uint8_t *blob, size_t nb, uint8_t *records, uint64_t n_sessions) {
    tab_t t;
    if (tab_open(&t, blob, nb)) return -1;
    const size_t S = rec_size(&t);
    sess_t s;
    memset(&s, 0, sizeof s);
    for (int p = 0; p < t.P; ++p) {
        if (t.family == FAM_WEREWOLF) {
            s.alive[p] = (t.init_masks >> 0) & 1;    s.can_vote[p] = (t.init_masks >> 1) & 1;
            s.eligible[p] = (t.init_masks >> 2) & 1; s.submitted[p] = (t.init_masks >> 3) & 1;
            s.revealed[p] = (t.init_masks >> 4) & 1; s.secret[p] = (t.init_masks >> 7) & 1;
        } else {
            s.is_speaker[p] = (t.init_masks >> 0) & 1; s.stmts[p] = (t.init_masks >> 1) & 1;
            s.lie_revealed[p] = (t.init_masks >> 2) & 1; s.can_vote[p] = (t.init_masks >> 3) & 1;
            s.has_voted[p] = (t.init_masks >> 4) & 1;
        }
    }
    uint8_t first[256];
    pack(&t, &s, first);
    for (uint64_t i = 0; i < n_sessions; ++i) memcpy(records + i * S, first, S);
    return 0;
}

/* stats (u64[GE_STATS_LEN], may be NULL): [0] += counted steps, [260+i] += visits of phase index i */
/* hmask[i] / hchoice[i * hstride + p]: human seats of session i and their inputs for the FIRST of the n_steps steps
 * (consumed by it, like the library's ge_batch_set_human_choices); both NULL = all bots. */
int ge_cpu_step_h(const uint8_t *blob, size_t nb, uint8_t *records, uint64_t n_sessions, uint64_t first_sid,
                  uint64_t seed, int n_steps, uint64_t *stats, int n_threads,
                  const uint32_t *hmask, const uint8_t *hchoice, size_t hstride) {
    tab_t t;
    if (tab_open(&t, blob, nb)) return -1;
    const size_t S = rec_size(&t);
    uint8_t none[MAXP];
    memset(none, 0xFF, sizeof none);
    uint64_t counted = 0, visits[32] = {0};
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel
    {
        uint64_t my_counted = 0, my_visits[32] = {0};
#pragma omp for schedule(static)
        for (int64_t i = 0; i < (int64_t)n_sessions; ++i) {
            sess_t s;
            if (tab_phase(&t, records[(size_t)i * S])[1] == KIND_TERMINAL) continue;   /* frozen (SPEC D16) */
            unpack(&t, records + (size_t)i * S, &s);
            for (int k = 0; k < n_steps; ++k)
                my_counted += (uint64_t)step_session(&t, &s, seed, first_sid + (uint64_t)i, my_visits, hmask ? hmask[i] : 0u,
                                                     (hchoice && k == 0) ? hchoice + (size_t)i * hstride : none);
            pack(&t, &s, records + (size_t)i * S);
        }
#pragma omp critical
        { counted += my_counted; for (int j = 0; j < 32; ++j) visits[j] += my_visits[j]; }
    }
#else
    (void)n_threads;
    for (uint64_t i = 0; i < n_sessions; ++i) {
        sess_t s;
        if (tab_phase(&t, records[i * S])[1] == KIND_TERMINAL) continue;
        unpack(&t, records + i * S, &s);
        for (int k = 0; k < n_steps; ++k)
            counted += (uint64_t)step_session(&t, &s, seed, first_sid + i, visits, hmask ? hmask[i] : 0u,
                                              (hchoice && k == 0) ? hchoice + (size_t)i * hstride : none);
        pack(&t, &s, records + i * S);
    }
#endif
    if (stats) { stats[0] += counted; for (int j = 0; j < 32; ++j) stats[260 + j] += visits[j]; }
    return 0;
}

int ge_cpu_step(const uint8_t *blob, size_t nb, uint8_t *records, uint64_t n_sessions, uint64_t first_sid,
                uint64_t seed, int n_steps, uint64_t *stats, int n_threads) {
    return ge_cpu_step_h(blob, nb, records, n_sessions, first_sid, seed, n_steps, stats, n_threads, NULL, NULL, 0);
}

/* final-state statistics: overwrites stats[1..259] and stats[292..547] (SPEC section 6) */
int ge_cpu_stats_final(const uint8_t *blob, size_t nb, const uint8_t *records, uint64_t n_sessions, uint64_t *stats) {
    tab_t t;
    if (tab_open(&t, blob, nb)) return -1;
    const size_t S = rec_size(&t);
    for (int j = 1; j < 260; ++j) stats[j] = 0;
    for (int j = 292; j < 548; ++j) stats[j] = 0;
    for (uint64_t i = 0; i < n_sessions; ++i) {
        sess_t s;
        unpack(&t, records + i * S, &s);
        int terminal = tab_phase(&t, s.phase)[1] == KIND_TERMINAL;
        if (t.family == FAM_WEREWOLF) {
            stats[1 + (s.winner <= 2 ? s.winner : 0)]++;
            if (terminal) {
                int alive = 0;
                for (int p = 0; p < t.P; ++p) alive += s.alive[p];
                stats[292 + alive]++;
            }
        } else {
            stats[1 + (terminal ? 1 : 0)]++;
            if (terminal) for (int p = 0; p < t.P; ++p) stats[292 + s.score[p]]++;
        }
        if (terminal) stats[4 + (s.step < 255 ? s.step : 255)]++;
    }
    return 0;
}

/* choices the actors of the CURRENT phase of one record would make on its next step (0 = not an actor
 * or no legal choice); used by tests of the host adapter.  choices has n_players entries. */
int ge_cpu_peek_choices(const uint8_t *blob, size_t nb, const uint8_t *record, uint64_t sid, uint64_t seed, uint8_t *choices) {
    tab_t t;
    if (tab_open(&t, blob, nb)) return -1;
    sess_t s;
    unpack(&t, record, &s);
    memset(choices, 0, (size_t)t.P);
    const uint8_t *ph = tab_phase(&t, s.phase);
    if (ph[1] != KIND_ACTION || s.step == 0) return 0;
    for (int p = 0; p < t.P; ++p) {
        if (!pred_holds(&t, &s, ph[8], p)) continue;
        uint32_t r = draw(seed, sid, (uint32_t)s.step, 0, p);
        if (ph[2] == ACT_PICK_PLAYER) {
            int legal[MAXP], n = 0;
            for (int q = 0; q < t.P; ++q) { if ((ph[4] & 1) && q == p) continue; if (pred_holds(&t, &s, ph[3], q)) legal[n++] = q; }
            choices[p] = n ? (uint8_t)(1 + legal[mulhi32(r, (uint32_t)n)]) : 0;
        } else if (ph[2] == ACT_PICK_OPTION) choices[p] = (uint8_t)(1 + mulhi32(r, ph[3]));
        else if (ph[2] == ACT_MARK) choices[p] = 1;
    }
    return 0;
}

/* audience masks: out[i * n_preds + j] = lane mask of predicate j (8 bytes each: pos0, neg0, pos1, neg1 as u16) */
int ge_cpu_eval_preds(const uint8_t *blob, size_t nb, const uint8_t *records, uint64_t n_sessions, const uint8_t *preds,
                      int n_preds, uint32_t *out) {
    tab_t t;
    if (tab_open(&t, blob, nb)) return -1;
    const size_t S = rec_size(&t);
    for (uint64_t i = 0; i < n_sessions; ++i) {
        sess_t s;
        unpack(&t, records + i * S, &s);
        for (int j = 0; j < n_preds; ++j) {
            const uint8_t *q = preds + 8 * j;
            uint32_t m = 0;
            for (int p = 0; p < t.P; ++p)          /* each record on its own: the caller ORs continued runs */
                if (pred_record_holds(&t, &s, q, p)) m |= 1u << p;
            out[i * (uint64_t)n_preds + j] = m;
        }
    }
    return 0;
}

/* Record well-formedness (SPEC.md section 7b): the rules the CUDA library applies to every imported record.
 * ok[i] = 1 when record i may be stepped.  Returns the number of malformed records, or -1 for a bad table. */
long ge_cpu_validate_records(const uint8_t *blob, size_t nb, const uint8_t *records, uint64_t n_sessions, uint8_t *ok) {
    tab_t t;
    if (tab_open(&t, blob, nb)) return -1;
    const size_t S = rec_size(&t);
    long bad = 0;
    for (uint64_t i = 0; i < n_sessions; ++i) {
        const uint8_t *r = records + i * S;
        int good = 1;
        const unsigned step = rd16(r + 2);
        if (r[0] >= t.n_phases || r[1] >= t.n_phases) good = 0;             /* the table is indexed with both */
        if (step == 0 && (r[0] != 0 || r[1] != 0)) good = 0;                /* a session that has not started is in phase 0 */
        if (t.family == FAM_WEREWOLF) {
            if (r[4] > 2 || r[5] > t.P || r[6] > t.P) good = 0;             /* winner tag, kill / protect ids */
            if ((r[7] & 0x7F) > t.max_revotes || (t.max_revotes == 0 && r[7] != 0)) good = 0;
            for (int f = 0; f < 10; ++f)                                    /* no mask bit above the player count */
                for (int p = t.P; p < 32; ++p)
                    if ((rd32(r + 8 + 4 * f) >> p) & 1u) good = 0;
            for (size_t p = 0; 48 + p < S; ++p)                             /* targets are player ids; padding is zero */
                if (r[48 + p] > ((int)p < t.P ? t.P : 0)) good = 0;
        } else {
            if (r[4] > t.P || r[6] > t.P || r[7] != 0) good = 0;            /* speaker, winner, reserved byte */
            for (size_t p = 0; 8 + 4 * p < S; ++p) {
                if ((int)p < t.P) { if (r[11 + 4 * p] & 0xE0) good = 0; }   /* five flag bits */
                else if (r[8 + 4 * p] || r[9 + 4 * p] || r[10 + 4 * p] || r[11 + 4 * p]) good = 0;
            }
        }
        if (ok) ok[i] = (uint8_t)good;
        bad += !good;
    }
    return bad;
}

int ge_cpu_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
