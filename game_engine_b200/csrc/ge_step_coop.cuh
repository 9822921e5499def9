// ge_step_coop.cuh — "lane per player" mapping of the referee/phase step (the mapping
// BASELINE.json's north_star describes): a group of L lanes owns one session, lane j of the group is
// player j+1, and a warp steps 32/L sessions at once.  Plurality votes are
// __match_any_sync + __popc, alive/actor sets are __ballot_sync masks, tie-breaks are a max-reduce
// over (count << 8 | 255 - id).  Session-level scalars (phase, masks) are held redundantly by every
// lane of the group; each lane additionally owns its player's bytes.
//
// All warp collectives are executed by the full warp in warp-uniform control flow (groups whose
// session is terminal contribute neutral values), so sub-warp groups never diverge around a
// collective.  Rules: SPEC.md.  Same session store layout as the thread-per-session mapping.
#pragma once
#include "ge_common.cuh"

namespace ge {

template <int L>
__device__ __forceinline__ uint32_t group_max(uint32_t v) {
#pragma unroll
    for (int d = L / 2; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
    return v;
}
template <int L>
__device__ __forceinline__ uint32_t group_or(uint32_t v) {
#pragma unroll
    for (int d = L / 2; d > 0; d >>= 1) v |= __shfl_xor_sync(0xFFFFFFFFu, v, d);
    return v;
}
// ballot restricted to the caller's group, shifted so that bit j = lane j of the group
template <int L>
__device__ __forceinline__ uint32_t group_ballot(bool pred, int gshift) {
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, pred);
    return L == 32 ? b : ((b >> gshift) & ((1u << L) - 1u));
}

// =============================================================================== werewolf family
struct WScalars {
    uint32_t h0, h1, alive, can_vote, eligible, submitted, revealed, investigated, wolf, secret, role_lo, role_hi;
    uint32_t cmp0, cmp1;        // comparison fields 13 / 14 of the table, balloted from the lanes' own target bytes
};
__device__ __forceinline__ uint32_t wc_field(const WScalars& s, int f, uint32_t ALL) {
    switch (f) {
    case 0: return s.alive;      case 1: return s.can_vote;  case 2: return s.eligible;
    case 3: return s.submitted;  case 4: return s.revealed;  case 5: return s.investigated;
    case 6: return s.wolf;       case 7: return s.secret;
    case 8: return s.secret ? ~(s.role_lo | s.role_hi) & ALL : 0u;      // role 0 only once roles are assigned
    case 12: return s.secret ? ALL : 0u;
    case 9: return s.role_lo & ~s.role_hi;
    case 10: return ~s.role_lo & s.role_hi;
    case 11: return s.role_lo & s.role_hi;
    case 13: return s.cmp0;
    case 14: return s.cmp1;
    case 15: return ALL;
    default: return 0u;
    }
}
__device__ __forceinline__ uint32_t wc_pred(const DevTable& T, const WScalars& s, int pi, uint32_t ALL) {
    uint32_t out = 0;
    for (;; ++pi) {                                            // a continued predicate is a run of records, ORed
        const ge_pred_t pr = T.pred[pi];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t pos = (c ? pr.pos1 : pr.pos0) & 0x7FFFu, neg = c ? pr.neg1 : pr.neg0, m = ALL;
            while (pos) { const int f = __ffs(pos) - 1; pos &= pos - 1; m &= wc_field(s, f, ALL); }
            while (neg) { const int f = __ffs(neg) - 1; neg &= neg - 1; m &= ~wc_field(s, f, ALL); }
            out |= m;
        }
        if (!(pr.pos0 & GE_PRED_CONTINUED)) break;
    }
    return out;
}

// P8: record bucket (8/16/24/32); L lanes per session = 8, 16 or 32.
template <int P8>
__global__ void __launch_bounds__(128)
k_step_w_coop(const __grid_constant__ DevTable T, const __grid_constant__ StepArgs A) {
    uint8_t* __restrict__ tiles = A.tiles;
    const uint64_t n_sessions = A.n_sessions, n_tiles = A.n_tiles, first_sid = A.first_sid, seed = A.seed;
    unsigned long long* __restrict__ stats = A.stats;
    const int n_steps = A.n_steps;
    constexpr int S = 48 + P8;
    constexpr int L = P8 <= 8 ? 8 : P8 <= 16 ? 16 : 32;
    constexpr int G = 32 / L;                  // sessions per warp
    constexpr int BITS = P8;
    __shared__ uint32_t s_visits[32];
    if (threadIdx.x < 32) s_visits[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int grp = lane / L, p = lane % L, gshift = grp * L;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const int P = T.h.n_players;
    const uint32_t ALL = all_mask(P);
    const uint32_t me = 1u << p;
    const bool is_player = p < P;
    const bool owns_byte = p < P8;
    const uint64_t n_units = n_tiles * L;      // a unit = G sessions = one warp iteration

    for (uint64_t u = warp0; u < n_units; u += nwarps) {
        const uint64_t tile = u / L;
        const uint32_t sl = (uint32_t)(u % L) * G + grp;      // session lane inside the tile
        uint8_t* base = tiles + tile * (uint64_t)(32 * S);
        const uint64_t sess = tile * 32 + sl;
        WScalars s;
        {
            const uint4 c0 = ld128(base + sl * 16);
            s.h0 = c0.x; s.h1 = c0.y; s.alive = c0.z; s.can_vote = c0.w;
        }
        bool live = sess < n_sessions && T.phase[s.h0 & 0xFF].kind != KIND_TERMINAL;
        uint32_t my_tgt = 0;
        if (live) {
            const uint4 c1 = ld128(base + 512 + sl * 16);
            const uint4 c2 = ld128(base + 1024 + sl * 16);
            s.eligible = c1.x; s.submitted = c1.y; s.revealed = c1.z; s.investigated = c1.w;
            s.wolf = c2.x; s.secret = c2.y; s.role_lo = c2.z; s.role_hi = c2.w;
            if (owns_byte) my_tgt = base[tile_off<S>(48 + p, sl)];
        } else {
            s.eligible = s.submitted = s.revealed = s.investigated = s.wolf = s.secret = s.role_lo = s.role_hi = 0;
        }
        const uint64_t sid = first_sid + sess;
        bool d0 = false, d1 = false, d2 = false, dp = false;

        for (int it = 0; it < n_steps; ++it) {
            const int X = s.h0 & 0xFF;
            const uint32_t step0 = s.h0 >> 16;
            const ge_phase_t& ph = T.phase[X];
            if (live && ph.kind == KIND_TERMINAL) live = false;
            const bool first_visit = live && step0 == 0;
            const bool run = live && step0 != 0;
            int np = -1;

            uint32_t winner = s.h1 & 0xFF, kill = (s.h1 >> 8) & 0xFF, protect = (s.h1 >> 16) & 0xFF, revote = s.h1 >> 24;
            const uint32_t prev = (s.h0 >> 8) & 0xFF;
            // comparison fields on the state at the start of the step (warp-uniform point: full-warp ballots)
            {
                const ge_cmp_t c0 = T.h.cmp[0], c1 = T.h.cmp[1];
                s.cmp0 = group_ballot<L>(T.h.n_cmp > 0 && is_player && cmp_holds(c0.op, my_tgt, c0.constant), gshift);
                s.cmp1 = group_ballot<L>(T.h.n_cmp > 1 && is_player && cmp_holds(c1.op, my_tgt, c1.constant), gshift);
            }

            // ---- PhaseNode: branch selection (replicated scalar work)
            int Y = X; uint32_t tag = 0;
            if (run) {
                int taken = ph.n_branches - 1;
                for (int b = 0; b < ph.n_branches; ++b) {
                    const ge_branch_t br = ph.br[b];
                    bool ok;
                    switch (br.op) {
                    case BR_ALWAYS: ok = true; break;
                    case BR_COUNT_EQ0: ok = wc_pred(T, s, br.a, ALL) == 0; break;
                    case BR_COUNT_GE: ok = __popc(wc_pred(T, s, br.a, ALL)) >= __popc(wc_pred(T, s, (int)br.arg, ALL)); break;
                    case BR_PREV_IN: ok = (br.arg >> prev) & 1u; break;
                    case BR_TIE_PENDING: ok = (revote & 0x80u) != 0; break;
                    default: ok = false; break;
                    }
                    if (ok) { taken = b; break; }
                }
                Y = ph.br[taken].next; tag = ph.br[taken].tag;
            }

            // ---- BotBehaviorNode: one lane = one bot
            const bool acting = run && ph.kind == KIND_ACTION;
            const bool recording = acting && ph.exit_op >= EX_VOTE_KILL && ph.exit_op <= EX_DAY_VOTE;
            uint32_t actors = 0, choice = 0;
            bool i_act = false;
            if (__any_sync(0xFFFFFFFFu, acting)) {
                if (acting) {
                    actors = wc_pred(T, s, ph.actor_pred, ALL);
                    i_act = is_player && (actors & me);
                    const uint4 r4 = philox4x32_10((uint32_t)sid, (uint32_t)(sid >> 32), step0, (uint32_t)(p >> 2), k0, k1);
                    const uint32_t r = word_of(r4, p & 3);
                    if (i_act) {
                        if (ph.action_op == ACT_PICK_PLAYER) {
                            uint32_t legal = wc_pred(T, s, ph.action_arg, ALL);
                            if (ph.action_flags & 1) legal &= ~me;
                            const uint32_t n = __popc(legal);
                            choice = n ? 1u + (uint32_t)kth_set_bit<BITS>(legal, __umulhi(r, n)) : 0u;
                        } else if (ph.action_op == ACT_PICK_OPTION) {
                            choice = 1u + __umulhi(r, (uint32_t)ph.action_arg);
                        } else {
                            choice = 1u;
                        }
                    }
                }
                // ---- tally by match_any: lanes of one group that chose the same player find each other
                const bool voted = i_act && choice != 0 && ph.action_op == ACT_PICK_PLAYER;
                const uint32_t mkey = voted ? ((uint32_t)grp << 8) | choice : 0x10000u | (uint32_t)lane;
                const uint32_t same = __match_any_sync(0xFFFFFFFFu, mkey);
                const uint32_t cnt = voted ? (uint32_t)__popc(same) : 0u;
                const uint32_t score = voted ? (cnt << 8) | (255u - choice) : 0u;
                const uint32_t best = group_max<L>(score);
                const uint32_t top_cnt = best >> 8;
                const uint32_t top_id = best ? 255u - (best & 0xFFu) : 0u;
                const uint32_t top_lanes = group_ballot<L>(voted && cnt == top_cnt, gshift);
                const bool tied = top_cnt != 0 && (uint32_t)__popc(top_lanes) > top_cnt;
                const uint32_t chosen = group_or<L>(voted ? 1u << (choice - 1) : 0u);
                const int la = actors ? __ffs(actors) - 1 : 0;
                const uint32_t first_choice = __shfl_sync(0xFFFFFFFFu, choice, gshift + la);

                if (recording) {
                    if (i_act) my_tgt = choice;
                    dp = true;
                    switch (ph.exit_op) {
                    case EX_VOTE_KILL: s.submitted |= actors; d1 = true; kill = top_id; break;
                    case EX_PROTECT: s.submitted |= actors; d1 = true; protect = actors ? first_choice : 0u; break;
                    case EX_INVESTIGATE_RESOLVE:
                        s.submitted |= actors; s.investigated |= chosen; d1 = true;
                        if (kill != 0 && kill != protect) { const uint32_t bit = ~(1u << (kill - 1)); s.alive &= bit; s.can_vote &= bit; s.eligible &= bit; }
                        kill = 0; protect = 0;
                        break;
                    case EX_DAY_VOTE:
                        if (T.h.max_revotes > 0 && tied && (revote & 0x7Fu) < T.h.max_revotes) {
                            revote = ((revote & 0x7Fu) + 1u) | 0x80u;
                        } else {
                            revote &= 0x7Fu;
                            if (top_id) {
                                const uint32_t bit = ~(1u << (top_id - 1));
                                s.alive &= bit; s.can_vote &= bit; s.eligible &= bit; s.revealed |= ~bit; d1 = true;
                            }
                        }
                        break;
                    default: break;
                    }
                }
            }

            // ---- entry effects of Y
            const int en = run ? T.phase[Y].entry_op : EN_NONE;
            if (__any_sync(0xFFFFFFFFu, en == EN_ASSIGN_ROLES)) {
                const uint4 r4 = philox4x32_10((uint32_t)sid, (uint32_t)(sid >> 32), step0, (1u << 16) | (uint32_t)(p >> 2), k0, k1);
                const uint32_t key = word_of(r4, p & 3);
                int rank = 0;
                for (int q = 0; q < L; ++q) {
                    const uint32_t kq = __shfl_sync(0xFFFFFFFFu, key, gshift + q);
                    if (q < P && (kq < key || (kq == key && q < p))) rank++;
                }
                const int W = T.h.n_wolves;
                const int role = !is_player ? 0 : rank < W ? 1 : rank == W ? 2 : rank == W + 1 ? 3 : 0;
                const uint32_t lo = group_ballot<L>(role & 1, gshift), hi = group_ballot<L>((role >> 1) & 1, gshift);
                if (en == EN_ASSIGN_ROLES) {
                    s.role_lo = lo; s.role_hi = hi; s.wolf = lo & ~hi; s.secret = lo | hi; s.eligible = lo | hi;
                    d1 = true; d2 = true;
                }
            }
            if (en == EN_NIGHT_RESET) {
                s.submitted = 0; my_tgt = 0; kill = 0; protect = 0; revote = 0;
                d1 = true; dp = true;
            }
            if (run) {
                if (tag) winner = tag;
                s.h1 = winner | (kill << 8) | (protect << 16) | (revote << 24);
                s.h0 = (uint32_t)Y | ((uint32_t)X << 8) | ((step0 + 1u) << 16);
                np = Y; d0 = true;
            } else if (first_visit) {
                s.h0 = (s.h0 & 0xFFFFu) | (1u << 16);
                np = X; d0 = true;
            }
            // one lane per session reports the visit
            count_visit(s_visits, p == 0 ? np : -1, lane);
        }

        if (p == 0) {
            if (d0) st128(base + sl * 16, make_uint4(s.h0, s.h1, s.alive, s.can_vote));
            if (d1) st128(base + 512 + sl * 16, make_uint4(s.eligible, s.submitted, s.revealed, s.investigated));
            if (d2) st128(base + 1024 + sl * 16, make_uint4(s.wolf, s.secret, s.role_lo, s.role_hi));
        }
        if (dp && owns_byte) base[tile_off<S>(48 + p, sl)] = (uint8_t)my_tgt;
    }
    __syncthreads();
    flush_visits(s_visits, stats);
    publish_presence(A, 0xFFFFFFFFu);      // this kernel loads every column; make the next launch do the same
}

// =================================================================================== TTL family
// PB: player bucket of the device record (4/8/16/32) = lanes per session.
template <int PB>
__global__ void __launch_bounds__(128)
k_step_t_coop(const __grid_constant__ DevTable T, const __grid_constant__ StepArgs A) {
    uint8_t* __restrict__ tiles = A.tiles;
    const uint64_t n_sessions = A.n_sessions, n_tiles = A.n_tiles, first_sid = A.first_sid, seed = A.seed;
    unsigned long long* __restrict__ stats = A.stats;
    const int n_steps = A.n_steps;
    constexpr int S = 8 + 4 * PB;
    constexpr int L = PB;
    constexpr int G = 32 / L;
    __shared__ uint32_t s_visits[32];
    if (threadIdx.x < 32) s_visits[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int grp = lane / L, p = lane % L, gshift = grp * L;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const int P = T.h.n_players;
    const uint32_t ALL = all_mask(P);
    const uint32_t me = 1u << p;
    const bool is_player = p < P;
    const uint64_t n_units = n_tiles * L;

    for (uint64_t u = warp0; u < n_units; u += nwarps) {
        const uint64_t tile = u / L;
        const uint32_t sl = (uint32_t)(u % L) * G + grp;
        uint8_t* base = tiles + tile * (uint64_t)(32 * S);
        const uint64_t sess = tile * 32 + sl;
        const uint2 hd = ld64(base + tile_off<S>(0, sl));
        uint32_t h0 = hd.x, h1 = hd.y;
        uint32_t pw = *reinterpret_cast<const uint32_t*>(base + tile_off<S>(8 + 4 * p, sl));
        bool live = sess < n_sessions && T.phase[h0 & 0xFF].kind != KIND_TERMINAL;
        const uint64_t sid = first_sid + sess;
        bool dirty = false;

        for (int it = 0; it < n_steps; ++it) {
            const int X = h0 & 0xFF;
            const uint32_t step0 = h0 >> 16;
            const ge_phase_t& ph = T.phase[X];
            if (live && ph.kind == KIND_TERMINAL) live = false;
            const bool first_visit = live && step0 == 0;
            const bool run = live && step0 != 0;
            int np = -1;
            uint32_t speaker = h1 & 0xFF, lie = (h1 >> 8) & 0xFF, winner = (h1 >> 16) & 0xFF;
            uint32_t fl = pw >> 24;
            // lane masks by ballot
            uint32_t m0 = group_ballot<L>(fl & TF_SPEAKER, gshift), m1 = group_ballot<L>(fl & TF_STMTS, gshift);
            uint32_t m2 = group_ballot<L>(fl & TF_REVEALED, gshift), m3 = group_ballot<L>(fl & TF_CANVOTE, gshift);
            uint32_t m4 = group_ballot<L>(fl & TF_VOTED, gshift);
            // comparison fields (numeric conditions): every lane tests its own value byte, one ballot per field
            uint32_t cm[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const ge_cmp_t c = T.h.cmp[k];
                cm[k] = group_ballot<L>(k < T.h.n_cmp && is_player && cmp_holds(c.op, (pw >> (8 * c.value_field)) & 0xFFu, c.constant), gshift);
            }
            auto field = [&](int f) -> uint32_t {
                return f == 15 ? ALL : f == 0 ? m0 : f == 1 ? m1 : f == 2 ? m2 : f == 3 ? m3 : f == 4 ? m4
                     : f == 11 ? cm[0] : f == 12 ? cm[1] : f == 13 ? cm[2] : f == 14 ? cm[3] : 0u;
            };
            auto pred = [&](int pi) -> uint32_t {
                uint32_t out = 0;
                for (;; ++pi) {                                // a continued predicate is a run of records, ORed
                    const ge_pred_t pr = T.pred[pi];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t pos = (c ? pr.pos1 : pr.pos0) & 0x7FFFu, neg = c ? pr.neg1 : pr.neg0, mm = ALL;
                        while (pos) { const int f = __ffs(pos) - 1; pos &= pos - 1; mm &= field(f); }
                        while (neg) { const int f = __ffs(neg) - 1; neg &= neg - 1; mm &= ~field(f); }
                        out |= mm;
                    }
                    if (!(pr.pos0 & GE_PRED_CONTINUED)) break;
                }
                return out;
            };

            // branch selection; ALL_VAL_GE is a ballot over the lanes' own bytes
            int Y = X; uint32_t tag = 0;
            {
                const int nb = run ? ph.n_branches : 0;
                const int nb_max = (int)group_max<32>((uint32_t)nb);      // warp-uniform trip count
                int taken = -1;
                for (int b = 0; b < nb_max; ++b) {
                    const bool mine = b < nb;
                    const ge_branch_t br = ph.br[mine ? b : 0];
                    const bool below = mine && br.op == BR_ALL_VAL_GE && is_player && ((pw >> (8 * (br.a & 3))) & 0xFFu) < br.arg;
                    const uint32_t any_below = group_ballot<L>(below, gshift);
                    bool ok = false;
                    if (mine) {
                        switch (br.op) {
                        case BR_ALWAYS: ok = true; break;
                        case BR_COUNT_EQ0: ok = pred(br.a) == 0; break;
                        case BR_COUNT_GE: ok = __popc(pred(br.a)) >= __popc(pred((int)br.arg)); break;
                        case BR_PREV_IN: ok = (br.arg >> ((h0 >> 8) & 0xFF)) & 1u; break;
                        case BR_ALL_VAL_GE: ok = any_below == 0; break;
                        default: ok = false; break;
                        }
                    }
                    if (ok && taken < 0) taken = b;
                }
                if (run) {
                    if (taken < 0) taken = ph.n_branches - 1;
                    Y = ph.br[taken].next; tag = ph.br[taken].tag;
                }
            }

            const bool acting = run && ph.kind == KIND_ACTION;
            uint32_t actors = 0, choice = 0;
            bool i_act = false;
            if (acting) {
                actors = pred(ph.actor_pred);
                i_act = is_player && (actors & me);
                const uint4 r4 = philox4x32_10((uint32_t)sid, (uint32_t)(sid >> 32), step0, (uint32_t)(p >> 2), k0, k1);
                const uint32_t r = word_of(r4, p & 3);
                if (i_act) {
                    if (ph.action_op == ACT_PICK_PLAYER) {
                        uint32_t legal = pred(ph.action_arg);
                        if (ph.action_flags & 1) legal &= ~me;
                        const uint32_t n = __popc(legal);
                        choice = n ? 1u + (uint32_t)kth_set_bit<PB>(legal, __umulhi(r, n)) : 0u;
                    } else {
                        choice = ph.action_op == ACT_PICK_OPTION ? 1u + __umulhi(r, (uint32_t)ph.action_arg) : 1u;
                    }
                }
            }
            const int la = actors ? __ffs(actors) - 1 : 0;
            const uint32_t first_choice = __shfl_sync(0xFFFFFFFFu, choice, gshift + la);
            if (acting) {
                switch (ph.exit_op) {
                case EX_T_STATEMENTS: m1 |= actors; break;
                case EX_T_LIE: if (actors) lie = first_choice; break;
                case EX_T_VOTES: m4 |= actors; if (i_act) pw = (pw & 0xFF00FFFFu) | (choice << 16); break;
                default: break;
                }
            }

            const int en = run ? T.phase[Y].entry_op : EN_NONE;
            // collectives needed by the entry ops, executed by the whole warp
            const uint32_t pending = group_ballot<L>(is_player && ((pw >> 8) & 0xFFu) < T.h.rounds, gshift);
            const bool voter = ((m4 & m3 & ~m0) & me) != 0;
            const uint32_t wrong = group_ballot<L>(voter && ((pw >> 16) & 0xFFu) != lie, gshift);
            const uint32_t my_score = is_player ? (pw & 0xFFu) : 0u;
            const uint32_t best = group_max<L>(is_player ? (my_score << 8) | (255u - (uint32_t)p) : 0u);
            if (en == EN_T_ROUND_START) {
                speaker = pending ? (uint32_t)__ffs(pending) : 0u;
                lie = 0;
                m0 = speaker ? 1u << (speaker - 1) : 0u;
                m3 = ALL & ~m0; m1 = 0; m2 = 0; m4 = 0;
                pw &= 0xFF00FFFFu;
            } else if (en == EN_T_REVEAL) {
                m2 |= m0;
            } else if (en == EN_T_SCORE) {
                uint32_t sc = pw & 0xFFu, rd = (pw >> 8) & 0xFFu;
                if (voter && !((wrong >> p) & 1u)) sc = (sc + 1u) & 0xFFu;
                if ((uint32_t)(p + 1) == speaker) { sc = (sc + (uint32_t)__popc(wrong)) & 0xFFu; rd = (rd + 1u) & 0xFFu; }
                pw = (pw & 0xFFFF0000u) | sc | (rd << 8);
            } else if (en == EN_T_FINAL) {
                winner = 256u - (best & 0xFFu);      // = (255 - (best & 0xFF)) + 1
            }
            if (run) {
                if (tag) winner = tag;
                fl = ((m0 >> p) & 1u) | (((m1 >> p) & 1u) << 1) | (((m2 >> p) & 1u) << 2) | (((m3 >> p) & 1u) << 3) | (((m4 >> p) & 1u) << 4);
                pw = (pw & 0x00FFFFFFu) | (fl << 24);
                h1 = speaker | (lie << 8) | (winner << 16);
                h0 = (uint32_t)Y | ((uint32_t)X << 8) | ((step0 + 1u) << 16);
                np = Y; dirty = true;
            } else if (first_visit) {
                h0 = (h0 & 0xFFFFu) | (1u << 16);
                np = X; dirty = true;
            }
            count_visit(s_visits, p == 0 ? np : -1, lane);
        }
        if (dirty) {
            if (p == 0) st64(base + tile_off<S>(0, sl), make_uint2(h0, h1));
            *reinterpret_cast<uint32_t*>(base + tile_off<S>(8 + 4 * p, sl)) = pw;
        }
    }
    __syncthreads();
    flush_visits(s_visits, stats);
    publish_presence(A, 0xFFFFFFFFu);      // this kernel loads every column; make the next launch do the same
}

}  // namespace ge
