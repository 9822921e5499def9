// ge_step_tps.cuh — "thread per session" mapping of the referee/phase step.
//
// One thread owns one session; a warp owns one 32-session tile, so every column of the session store
// is one coalesced 128-bit access per lane.  All per-player state is 32-bit lane masks (bit p-1 =
// player p); votes are tallied in BIT-SLICED counters (plane b holds bit b of every candidate's vote
// count), so a plurality with lowest-id tie-break is a few AND/XORs instead of a per-candidate loop.
// Loops over players are fully unrolled so the per-player bytes live in named registers (no local
// memory).  The rules are SPEC.md; the restated reference code is BotBehaviorNode / PhaseNode /
// RefereeNode (reference agent/game_agent_v2.py:468-617, 987-1241, 619-803).
#pragma once
#include <type_traits>
#include "ge_common.cuh"

namespace ge {

enum { DIRTY_C0 = 1, DIRTY_C1 = 2, DIRTY_C2 = 4, DIRTY_PL = 8,
       DIRTY_RV = 16 };     // role_revealed / investigated changed (they sit in their own column of the 16-player packed store)

// =============================================================================== werewolf family
template <int P8>
struct WState {
    uint32_t h0, h1;                                   // phase|prev<<8|step<<16 ; winner|kill<<8|protect<<16|revote<<24
    uint32_t alive, can_vote;                          // column 0
    uint32_t eligible, submitted, revealed, investigated;   // column 1
    uint32_t wolf, secret, role_lo, role_hi;           // column 2
    uint32_t tw[P8 / 4];                               // selected_target_id bytes
    // Packed store: the ten masks as they sit in HBM (the words of the dense wire record, SPEC 5b).
    //   up to 8 players : pk0 = alive | can_vote << 8 | eligible << 16 | submitted << 24,
    //                     pk1 = revealed | investigated << 8 | wolf << 16 | secret << 24, pk2 = role_lo | role_hi << 8
    //   up to 16 players: pk0 = alive | can_vote << 16, pk1 = eligible | submitted << 16, pk2 = revealed | investigated << 16,
    //                     pk3 = wolf | secret << 16, pk4 = role_lo | role_hi << 16
    // A step unpacks them INSIDE its per-phase body and packs them again at its end, so that with a build-time table every
    // phase extracts only the fields it reads and re-inserts only the ones it writes (pack(unpack(x)) folds to x).
    uint32_t pk0, pk1, pk2, pk3, pk4;
    __device__ __forceinline__ void unpack() {
        if constexpr (P8 == 8) {
            alive = pk0 & 0xFFu; can_vote = (pk0 >> 8) & 0xFFu; eligible = (pk0 >> 16) & 0xFFu; submitted = pk0 >> 24;
            revealed = pk1 & 0xFFu; investigated = (pk1 >> 8) & 0xFFu; wolf = (pk1 >> 16) & 0xFFu; secret = pk1 >> 24;
            role_lo = pk2 & 0xFFu; role_hi = (pk2 >> 8) & 0xFFu;
        } else {
            alive = pk0 & 0xFFFFu; can_vote = pk0 >> 16; eligible = pk1 & 0xFFFFu; submitted = pk1 >> 16;
            revealed = pk2 & 0xFFFFu; investigated = pk2 >> 16; wolf = pk3 & 0xFFFFu; secret = pk3 >> 16;
            role_lo = pk4 & 0xFFFFu; role_hi = pk4 >> 16;
        }
    }
    __device__ __forceinline__ void repack() {
        if constexpr (P8 == 8) {
            pk0 = (alive & 0xFFu) | ((can_vote & 0xFFu) << 8) | ((eligible & 0xFFu) << 16) | (submitted << 24);
            pk1 = (revealed & 0xFFu) | ((investigated & 0xFFu) << 8) | ((wolf & 0xFFu) << 16) | (secret << 24);
            pk2 = (role_lo & 0xFFu) | ((role_hi & 0xFFu) << 8);
        } else {
            pk0 = (alive & 0xFFFFu) | (can_vote << 16); pk1 = (eligible & 0xFFFFu) | (submitted << 16);
            pk2 = (revealed & 0xFFFFu) | (investigated << 16); pk3 = (wolf & 0xFFFFu) | (secret << 16);
            pk4 = (role_lo & 0xFFFFu) | (role_hi << 16);
        }
    }
};


constexpr int TPS_THREADS = 128;
// resident CTAs per SM the werewolf kernels are compiled for (register budget = 65536 / (128 x this)); the 24/32-player
// figure can be overridden at build time for occupancy experiments (-DGE_P32_CTAS=5)
#ifndef GE_P32_CTAS
#define GE_P32_CTAS 5
#endif
#ifndef GE_P8_CTAS
#define GE_P8_CTAS 8
#endif
#ifndef GE_P16_CTAS
#define GE_P16_CTAS 6
#endif
#define GE_W_CTAS(P8) ((P8) <= 8 ? GE_P8_CTAS : (P8) <= 16 ? GE_P16_CTAS : GE_P32_CTAS)
// The specialised 8-player step kernel over the packed store needs 48 registers without a spill (the masks stay packed until a
// phase body wants them), so it is compiled for 10 resident CTAs per SM: with the default 8 streams x 3 CTAs per SM more of the
// ring's launches are co-resident (+2.9 %, tools/ab_occ.sh; 11 spills; the ring kernel and full-grid launches prefer 8).
#ifndef GE_P8_PACKED_SPEC_CTAS
#define GE_P8_PACKED_SPEC_CTAS 10
#endif
template <int P8, class Spec, bool PK>
constexpr int w_step_ctas() { return (P8 == 8 && PK && !std::is_void<Spec>::value) ? GE_P8_PACKED_SPEC_CTAS : GE_W_CTAS(P8); }

// Per-thread table of the 12 predicate field masks (SPEC.md section 2) in shared memory, laid out
// [field][thread] (conflict-free).  Predicates index it dynamically; this replaced a switch over the field
// id that cost a quarter of an action step's instructions.
struct FieldTable {
    uint32_t (*f)[TPS_THREADS];
    int tid;
    uint8_t* lut;       // this thread's column of the legal-position table (P8 > 16 only): entry j at lut[j * TPS_THREADS]
    template <int P8>
    __device__ __forceinline__ void fill(const WState<P8>& s, uint32_t ALL) const {
        f[0][tid] = s.alive;      f[1][tid] = s.can_vote;  f[2][tid] = s.eligible;  f[3][tid] = s.submitted;
        f[4][tid] = s.revealed;   f[5][tid] = s.investigated;  f[6][tid] = s.wolf;  f[7][tid] = s.secret;
        f[8][tid] = s.secret ? ~(s.role_lo | s.role_hi) & ALL : 0u;
        f[12][tid] = s.secret ? ALL : 0u;
        f[9][tid] = s.role_lo & ~s.role_hi;
        f[10][tid] = ~s.role_lo & s.role_hi;
        f[11][tid] = s.role_lo & s.role_hi;
        f[15][tid] = ALL;
    }
    // comparison fields of the table (numeric conditions on selected_target_id): mask fields 13 and 14
    template <int P8>
    __device__ __forceinline__ void fill_cmp(const DevTable& T, const WState<P8>& s) const {
        if (T.h.n_cmp > 0) f[13][tid] = cmp_mask_bytes(s.tw, T.h.n_players, T.h.cmp[0]);
        if (T.h.n_cmp > 1) f[14][tid] = cmp_mask_bytes(s.tw, T.h.n_players, T.h.cmp[1]);
    }
    __device__ __forceinline__ uint32_t pred(const DevTable& T, int pi, uint32_t ALL) const {
        uint32_t out = 0;
        for (;; ++pi) {                                        // a continued predicate is a run of records, ORed
            const ge_pred_t pr = T.pred[pi];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t pos = (c ? pr.pos1 : pr.pos0) & 0x7FFFu, neg = c ? pr.neg1 : pr.neg0;
                if (neg & 0x8000u) continue;                   // "& ~ALL": unused clause
                uint32_t m = ALL;
                while (pos) { const int k = __ffs(pos) - 1; pos &= pos - 1; m &= f[k][tid]; }
                while (neg) { const int k = __ffs(neg) - 1; neg &= neg - 1; m &= ~f[k][tid]; }
                out |= m;
            }
            if (!(pr.pos0 & GE_PRED_CONTINUED)) break;
        }
        return out;
    }
};

// bit-sliced vote counters
template <int NPL>
struct Tally {
    uint32_t pl[NPL];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int b = 0; b < NPL; ++b) pl[b] = 0;
    }
    __device__ __forceinline__ void add(uint32_t onehot) {
        uint32_t carry = onehot;
#pragma unroll
        for (int b = 0; b < NPL; ++b) { const uint32_t t = pl[b] & carry; pl[b] ^= carry; carry = t; }
    }
    // adds a 3-bit bit-sliced number (one count 0..7 per candidate): full adders on the low planes (one LOP3 per
    // sum and per carry), half adders above
    __device__ __forceinline__ void add3(uint32_t b0, uint32_t b1, uint32_t b2) {
        uint32_t carry;
        { const uint32_t t = pl[0] & b0; pl[0] ^= b0; carry = t; }
        if (NPL > 1) { const uint32_t a = pl[1]; pl[1] = a ^ b1 ^ carry; carry = (a & b1) | (a & carry) | (b1 & carry); }
        if (NPL > 2) { const uint32_t a = pl[2]; pl[2] = a ^ b2 ^ carry; carry = (a & b2) | (a & carry) | (b2 & carry); }
#pragma unroll
        for (int b = 3; b < NPL; ++b) { const uint32_t t = pl[b] & carry; pl[b] ^= carry; carry = t; }
    }
    // seven one-hot votes -> their per-candidate sum as three bit planes (carry-save adder tree, 8 LOP3), added in
    __device__ __forceinline__ void add7(const uint32_t (&g)[7]) {
        const uint32_t s1 = g[0] ^ g[1] ^ g[2], c1 = (g[0] & g[1]) | (g[0] & g[2]) | (g[1] & g[2]);
        const uint32_t s2 = g[3] ^ g[4] ^ g[5], c2 = (g[3] & g[4]) | (g[3] & g[5]) | (g[4] & g[5]);
        const uint32_t b0 = s1 ^ s2 ^ g[6], c3 = (s1 & s2) | (s1 & g[6]) | (s2 & g[6]);
        add3(b0, c1 ^ c2 ^ c3, (c1 & c2) | (c1 & c3) | (c2 & c3));
    }
    // returns the mask of candidates that share the highest (non-zero) count
    __device__ __forceinline__ uint32_t top() const {
        uint32_t cand = 0;
#pragma unroll
        for (int b = 0; b < NPL; ++b) cand |= pl[b];
#pragma unroll
        for (int b = NPL - 1; b >= 0; --b) { const uint32_t t = cand & pl[b]; cand = t ? t : cand; }
        return cand;
    }
};

template <int P8>
__device__ __forceinline__ void w_die(WState<P8>& s, int id) {
    const uint32_t bit = ~(1u << (id - 1));
    s.alive &= bit; s.can_vote &= bit; s.eligible &= bit;
}

template <int NW>
__device__ __forceinline__ void set_byte(uint32_t (&w)[NW], int p, uint32_t v) {
    const int wi = p >> 2, sh = (p & 3) * 8;
#pragma unroll
    for (int i = 0; i < NW; ++i)
        if (i == wi) w[i] = (w[i] & ~(0xFFu << sh)) | (v << sh);
}

// Where a night action's target byte goes when the per-player columns were NOT loaded for this launch (the host
// proves per phase that nothing reads them, DevTable::need bit 2): straight to its final place in the tile, one
// byte store per actor, instead of a read-modify-write of the whole player columns through registers.
struct PlSink {
    uint8_t* p16;       // this lane's 16 bytes of the first full player column (column 3)
    uint8_t* p8;        // this lane's 8 bytes of the trailing half column (P8 % 16 == 8)
    bool direct;
};
template <int P8>
__device__ __forceinline__ void pl_store(const PlSink& K, int p, uint32_t v) {
    constexpr int NT16 = P8 / 16;
    uint8_t* a = ((P8 % 16) != 0 && p >= 16 * NT16) ? K.p8 + (p - 16 * NT16) : K.p16 + (p >> 4) * 512 + (p & 15);
    *a = (uint8_t)v;
}

// ---- table views -----------------------------------------------------------------------------------
// The step body below is written once against a "view" of the phase it runs in.  RtView reads the
// table passed at run time (any game).  CtView<Spec, X> reads a table generated at BUILD time from the
// shipped game files (ge_spec_gen.cuh): every accessor is a constant expression, so for each phase the
// compiler folds the op dispatch, the predicates and the branch list and drops the dead code — the
// table is still the only statement of the game, it is just consumed by the compiler instead of by an
// interpreter loop.
template <int P8>
__device__ __forceinline__ uint32_t w_field_of(const WState<P8>& s, int f, uint32_t ALL) {
    switch (f) {
    case 0: return s.alive;      case 1: return s.can_vote;  case 2: return s.eligible;
    case 3: return s.submitted;  case 4: return s.revealed;  case 5: return s.investigated;
    case 6: return s.wolf;       case 7: return s.secret;
    case 8: return s.secret ? ~(s.role_lo | s.role_hi) & ALL : 0u;      // role 0 only once roles are assigned
    case 12: return s.secret ? ALL : 0u;
    case 9: return s.role_lo & ~s.role_hi;
    case 10: return ~s.role_lo & s.role_hi;
    case 11: return s.role_lo & s.role_hi;
    case 15: return ALL;
    default: return 0u;
    }
}

struct RtView {
    const DevTable& T;
    const ge_phase_t& ph;
    static constexpr bool is_const = false;
    __device__ __forceinline__ RtView(const DevTable& t, int X) : T(t), ph(t.phase[X]) {}
    __device__ __forceinline__ int kind() const { return ph.kind; }
    __device__ __forceinline__ int action_op() const { return ph.action_op; }
    __device__ __forceinline__ int action_arg() const { return ph.action_arg; }
    __device__ __forceinline__ int action_flags() const { return ph.action_flags; }
    __device__ __forceinline__ int exit_op() const { return ph.exit_op; }
    __device__ __forceinline__ int actor_pred() const { return ph.actor_pred; }
    __device__ __forceinline__ int n_branches() const { return ph.n_branches; }
    __device__ __forceinline__ ge_branch_t branch(int b) const { return ph.br[b]; }
    __device__ __forceinline__ int entry_op_after(int taken) const { return T.phase[ph.br[taken].next].entry_op; }
    __device__ __forceinline__ int n_players() const { return T.h.n_players; }
    __device__ __forceinline__ int n_wolves() const { return T.h.n_wolves; }
    __device__ __forceinline__ int max_revotes() const { return T.h.max_revotes; }
    __device__ __forceinline__ int rounds() const { return T.h.rounds; }
    __device__ __forceinline__ ge_pred_t pred_rec(int pi) const { return T.pred[pi]; }
    __device__ __forceinline__ bool wants_fields() const { return ph.kind == KIND_ACTION || ph.n_branches > 1; }
    template <int P8>
    __device__ __forceinline__ uint32_t pred(const FieldTable& F, const WState<P8>&, int pi, uint32_t ALL) const { return F.pred(T, pi, ALL); }
};

template <class Spec, int X>
struct CtView {
    static constexpr bool is_const = true;
    __device__ __forceinline__ static constexpr int kind() { return Spec::phase(X).kind; }
    __device__ __forceinline__ static constexpr int action_op() { return Spec::phase(X).action_op; }
    __device__ __forceinline__ static constexpr int action_arg() { return Spec::phase(X).action_arg; }
    __device__ __forceinline__ static constexpr int action_flags() { return Spec::phase(X).action_flags; }
    __device__ __forceinline__ static constexpr int exit_op() { return Spec::phase(X).exit_op; }
    __device__ __forceinline__ static constexpr int actor_pred() { return Spec::phase(X).actor_pred; }
    __device__ __forceinline__ static constexpr int n_branches() { return Spec::phase(X).n_branches; }
    __device__ __forceinline__ static constexpr ge_branch_t branch(int b) { return Spec::phase(X).br[b]; }
    __device__ __forceinline__ static int entry_op_after(int taken) {
        int en = EN_NONE;
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (b < Spec::phase(X).n_branches && b == taken) en = Spec::phase(Spec::phase(X).br[b].next).entry_op;
        return en;
    }
    __device__ __forceinline__ static constexpr int n_players() { return Spec::n_players; }
    __device__ __forceinline__ static constexpr int n_wolves() { return Spec::n_wolves; }
    __device__ __forceinline__ static constexpr int max_revotes() { return Spec::max_revotes; }
    __device__ __forceinline__ static constexpr int rounds() { return Spec::rounds; }
    __device__ __forceinline__ static constexpr ge_pred_t pred_rec(int pi) { return Spec::pred(pi); }
    __device__ __forceinline__ static constexpr bool wants_fields() { return false; }       // predicates fold to register ops
    template <int P8>
    __device__ __forceinline__ static uint32_t pred(const FieldTable&, const WState<P8>& s, int pi, uint32_t ALL) {
        const ge_pred_t pr = Spec::pred(pi);          // pi is a constant after inlining: the loops below fold
        uint32_t out = 0;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint32_t pos = c ? pr.pos1 : pr.pos0, neg = c ? pr.neg1 : pr.neg0;
            if (neg & 0x8000u) continue;
            uint32_t m = ALL;
#pragma unroll
            for (int f = 0; f < 16; ++f) {
                if ((pos >> f) & 1u) m &= w_field_of(s, f, ALL);
                if ((neg >> f) & 1u) m &= ~w_field_of(s, f, ALL);
            }
            out |= m;
        }
        return out;
    }
};

// One step of one session in phase X (already known not to be terminal and step0 != 0).  Returns the
// phase index entered; `dirty` collects which column groups changed.
template <int P8, class V>
__device__ __forceinline__ int w_step_body(const V v, const int X, WState<P8>& s, const FieldTable& F, const PlSink& K, uint32_t sid_lo, uint32_t sid_hi,
                                           const StepArgs& A, uint32_t& dirty, const HumanIn& H) {
    // planes of the bit-sliced vote counters: enough for P votes; a build-time table whose phase is the wolves'
    // vote needs only enough for n_wolves votes
    constexpr int NPL_FULL = P8 <= 8 ? 4 : P8 <= 16 ? 5 : 6;
    constexpr int NPL = [] {
        if constexpr (V::is_const) {
            if (V::exit_op() == EX_VOTE_KILL) { int w = V::n_wolves(), b = 1; while ((1 << b) <= w) ++b; return b < NPL_FULL ? b : NPL_FULL; }
        }
        return NPL_FULL;
    }();
    const int P = v.n_players();
    const uint32_t ALL = all_mask(P);
    const uint32_t step0 = s.h0 >> 16;
    uint32_t winner = s.h1 & 0xFF, kill = (s.h1 >> 8) & 0xFF, protect = (s.h1 >> 16) & 0xFF, revote = s.h1 >> 24;
    const uint32_t prev = (s.h0 >> 8) & 0xFF;
    const bool acting = v.kind() == KIND_ACTION;
    if (v.wants_fields()) { F.fill(s, ALL); if constexpr (!V::is_const) F.fill_cmp(v.T, s); }

    // ---- PhaseNode: ordered branch evaluation on the state before this step's effects
    const int nb = v.n_branches();
    int taken = nb - 1;
    if (nb > 1) {
        bool done = false;
#pragma unroll(V::is_const ? 4 : 1)
        for (int b = 0; b < 4; ++b) {
            if (b < nb && !done) {
                const ge_branch_t br = v.branch(b);
                bool ok;
                switch (br.op) {
                case BR_ALWAYS: ok = true; break;
                case BR_COUNT_EQ0: ok = v.template pred<P8>(F, s, br.a, ALL) == 0; break;
                case BR_COUNT_GE: ok = __popc(v.template pred<P8>(F, s, br.a, ALL)) >= __popc(v.template pred<P8>(F, s, (int)br.arg, ALL)); break;
                case BR_PREV_IN: ok = (br.arg >> prev) & 1u; break;
                case BR_TIE_PENDING: ok = (revote & 0x80u) != 0; break;
                default: ok = false; break;
                }
                if (ok) { taken = b; done = true; }
            }
        }
    }
    int Y = 0; uint32_t tag = 0;
    if constexpr (V::is_const) {
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (b < nb && b == taken) { Y = v.branch(b).next; tag = v.branch(b).tag; }
    } else {
        Y = v.branch(taken).next; tag = v.branch(taken).tag;
    }

    // ---- BotBehaviorNode: actors are visited in rank order (the i-th actor of every session in the
    // same warp iteration, so lanes stay converged); the Philox block is recomputed only when it changes.
    if (acting) {
        const uint32_t actors = v.template pred<P8>(F, s, v.actor_pred(), ALL);
        const int aop = v.action_op();
        const uint32_t legal0 = aop == ACT_PICK_PLAYER ? v.template pred<P8>(F, s, v.action_arg(), ALL) : 0u;
        const uint32_t excl = (v.action_flags() & 1) ? 0xFFFFFFFFu : 0u;
        const int exo = v.exit_op();
        const bool record = exo >= EX_VOTE_KILL && exo <= EX_DAY_VOTE;
        const bool tallying = exo == EX_VOTE_KILL || exo == EX_DAY_VOTE;
        Tally<NPL> tally; tally.clear();
        uint32_t chosen = 0, first_choice = 0;
        uint32_t nib = 0;                       // P8 <= 8: nibble-packed vote counters (one candidate per nibble)
        bool nib_used = false;
        // ---- human seats among the actors (SPEC D3h): the step waits — history grows, nothing else changes — until
        // every one of them has a valid input; bots draw on the step that completes the phase
        const uint32_t hact = actors & H.mask;
        if (hact) {
            uint32_t rem_h = hact;
            bool waiting = false;
            while (rem_h) {
                const int p = __ffs(rem_h) - 1;
                rem_h &= rem_h - 1;
                if (human_choice_of(H, p, aop, v.action_arg(), legal0 & ~(excl & (1u << p))) < 0) waiting = true;
            }
            if (waiting) {
                s.h0 = (uint32_t)X | ((uint32_t)X << 8) | ((step0 + 1u) << 16);
                return X;
            }
        }
        if (aop == ACT_PICK_PLAYER && !K.direct && hact == 0 && (uint32_t)__popc(actors) * 3u > (uint32_t)P8) {
            // ---- many actors (day vote): one statically unrolled pass over the players.  Philox words, target
            // bytes and ranks are static; the pick is a lookup in a nibble LUT of the legal players' positions.
            const uint32_t n0 = __popc(legal0);
            uint64_t lut = 0;
            if (P8 <= 16) {
                int j = 0;
#pragma unroll
                for (int q = 0; q < P8; ++q)
                    if ((legal0 >> q) & 1u) { lut |= (uint64_t)q << (4 * j); ++j; }
            }
            uint4 R[P8 / 4];
#pragma unroll
            for (int b = 0; b < P8 / 4; ++b)
                R[b] = ((actors >> (4 * b)) & 0xFu) ? philox4x32_10(sid_lo, sid_hi, step0, (uint32_t)b, A.rk) : make_uint4(0, 0, 0, 0);
            if (P8 > 16) {
                // more than 16 players: the positions of the legal players go to a per-thread byte column in shared
                // memory (entry j = position of the j-th legal player), a pick is one LDS
                uint32_t j = 0;
#pragma unroll
                for (int q = 0; q < P8; ++q)
                    if ((legal0 >> q) & 1u) { F.lut[j * TPS_THREADS] = (uint8_t)q; ++j; }
            }
            uint32_t rank = 0;
            bool have_first = false;
            nib_used = P8 <= 8;
            uint32_t g[7];                      // P8 > 8: one-hot votes of seven consecutive players (carry-save tally)
#pragma unroll
            for (int p = 0; p < P8; ++p) {
                const uint32_t self_in = (legal0 >> p) & 1u;
                uint32_t vote = 0;
                if ((actors >> p) & 1u) {
                    const uint32_t r = word_of(R[p >> 2], p & 3);
                    const uint32_t skip = excl & self_in;                 // 1 when this actor must skip itself
                    const uint32_t n = n0 - skip;
                    const uint32_t k = __umulhi(r, n);
                    const uint32_t j = k + ((skip && k >= rank) ? 1u : 0u);
                    int idx;
                    if (P8 <= 16) idx = (int)((lut >> (4 * j)) & 0xFu);
                    else idx = n ? (int)F.lut[j * TPS_THREADS] : 0;
                    const uint32_t choice = n ? (uint32_t)idx + 1u : 0u;
                    if (n) {
                        chosen |= 1u << idx;
                        if (tallying) { if (P8 <= 8) nib += 1u << (4 * idx); else vote = 1u << idx; }
                    }
                    if (!have_first) { first_choice = choice; have_first = true; }
                    if (record) s.tw[p >> 2] = __byte_perm(s.tw[p >> 2], choice, 0x3210u + ((4u - (p & 3)) << (4 * (p & 3))));   // byte p&3 <- choice (one PRMT)
                }
                rank += self_in;
                if (P8 > 8) {
                    g[p % 7] = vote;
                    if (p % 7 == 6 || p == P8 - 1) {
#pragma unroll
                        for (int i = (p % 7) + 1; i < 7; ++i) g[i] = 0;
                        tally.add7(g);
                    }
                }
            }
        } else {
        uint32_t rem = actors;
        int cur_blk = -1;
        uint4 r4 = make_uint4(0, 0, 0, 0);
        while (rem) {
            const int p = __ffs(rem) - 1;
            const bool is_first = rem == actors;
            rem &= rem - 1;
            const int blk = p >> 2;
            uint32_t choice;
            if ((hact >> p) & 1u) {                   // a person's input instead of a draw (validated above)
                choice = (uint32_t)human_choice_of(H, p, aop, v.action_arg(), legal0 & ~(excl & (1u << p)));
                if (aop == ACT_PICK_PLAYER && choice) { chosen |= 1u << (choice - 1u); if (tallying) tally.add(1u << (choice - 1u)); }
                if (is_first) first_choice = choice;
                if (record) { if (K.direct) pl_store<P8>(K, p, choice); else set_byte(s.tw, p, choice); }
                continue;
            }
            if (blk != cur_blk) { r4 = philox4x32_10(sid_lo, sid_hi, step0, (uint32_t)blk, A.rk); cur_blk = blk; }
            const uint32_t r = word_of(r4, p & 3);
            if (aop == ACT_PICK_PLAYER) {
                const uint32_t legal = legal0 & ~(excl & (1u << p));
                const uint32_t n = __popc(legal);
                const int idx = kth_set_bit<P8>(legal, __umulhi(r, n));
                choice = n ? (uint32_t)idx + 1u : 0u;
                if (n) { chosen |= 1u << idx; if (tallying) tally.add(1u << idx); }
            } else if (aop == ACT_PICK_OPTION) {
                choice = 1u + __umulhi(r, (uint32_t)v.action_arg());
            } else {
                choice = 1u;
            }
            if (is_first) first_choice = choice;
            if (record) { if (K.direct) pl_store<P8>(K, p, choice); else set_byte(s.tw, p, choice); }
        }
        }
        // plurality: candidates sharing the highest non-zero count (lowest id wins ties)
        uint32_t top = 0;
        if (tallying) {
            if (nib_used) {
                uint32_t best = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t c = (nib >> (4 * q)) & 0xFu;
                    if (c > best) { best = c; top = 1u << q; } else if (c == best && c) top |= 1u << q;
                }
            } else {
                top = tally.top();
            }
        }
        if (record && !K.direct) dirty |= DIRTY_PL;
        // ---- RefereeNode, effects of the phase just left
        switch (exo) {
        case EX_VOTE_KILL:
            s.submitted |= actors; dirty |= DIRTY_C1;
            kill = top ? (uint32_t)__ffs(top) : 0u;
            break;
        case EX_PROTECT:
            s.submitted |= actors; dirty |= DIRTY_C1;
            protect = first_choice;
            break;
        case EX_INVESTIGATE_RESOLVE:
            s.submitted |= actors; s.investigated |= chosen; dirty |= DIRTY_C1 | DIRTY_RV;
            if (kill != 0 && kill != protect) w_die(s, (int)kill);
            kill = 0; protect = 0;
            break;
        case EX_DAY_VOTE: {
            const uint32_t x = top ? (uint32_t)__ffs(top) : 0u;
            const bool tied = __popc(top) > 1;
            if (v.max_revotes() > 0 && tied && (revote & 0x7Fu) < (uint32_t)v.max_revotes()) {
                revote = ((revote & 0x7Fu) + 1u) | 0x80u;
            } else {
                revote &= 0x7Fu;
                if (x) { w_die(s, (int)x); s.revealed |= 1u << (x - 1); dirty |= DIRTY_C1 | DIRTY_RV; }
            }
        } break;
        default: break;
        }
    }

    // ---- RefereeNode, effects of entering Y
    const int en = v.entry_op_after(taken);
    if (en == EN_ASSIGN_ROLES) {
        uint32_t key[P8];
#pragma unroll
        for (int b = 0; b < P8 / 4; ++b) {
            if (4 * b < P) {
                const uint4 q4 = philox4x32_10(sid_lo, sid_hi, step0, (1u << 16) | (uint32_t)b, A.rk);
                key[4 * b] = q4.x; key[4 * b + 1] = q4.y; key[4 * b + 2] = q4.z; key[4 * b + 3] = q4.w;
            } else {
                key[4 * b] = key[4 * b + 1] = key[4 * b + 2] = key[4 * b + 3] = 0;
            }
        }
        // rank < W+2 is all that matters: pick the W+2 smallest (key, id) pairs in order
        uint32_t rem = ALL, wolf = 0, lo = 0, hi = 0;
        const int W = v.n_wolves();
        for (int round = 0; round < W + 2; ++round) {
            uint32_t best = 0; int bi = -1;
#pragma unroll
            for (int p = 0; p < P8; ++p) {
                if (((rem >> p) & 1u) && (bi < 0 || key[p] < best)) { best = key[p]; bi = p; }
            }
            if (bi < 0) break;
            rem &= ~(1u << bi);
            const uint32_t bit = 1u << bi;
            if (round < W) { wolf |= bit; lo |= bit; }            // role index 1
            else if (round == W) { hi |= bit; }                    // role index 2 (Doctor)
            else { lo |= bit; hi |= bit; }                         // role index 3 (Detective)
        }
        s.wolf = wolf; s.role_lo = lo; s.role_hi = hi;
        s.secret = lo | hi; s.eligible = lo | hi;
        dirty |= DIRTY_C1 | DIRTY_C2;
    } else if (en == EN_NIGHT_RESET) {
        s.submitted = 0;
#pragma unroll
        for (int b = 0; b < P8 / 4; ++b) s.tw[b] = 0;
        kill = 0; protect = 0; revote = 0;
        dirty |= DIRTY_C1 | DIRTY_PL;
    }
    if (tag) winner = tag;

    s.h1 = winner | (kill << 8) | (protect << 16) | (revote << 24);
    s.h0 = (uint32_t)Y | ((uint32_t)X << 8) | ((step0 + 1u) << 16);
    return Y;
}

// Generic entry: interpret the run-time table.  Returns the phase entered or -1 for a terminal session.
template <int P8, bool PK = false>
__device__ __forceinline__ int w_step(const DevTable& T, WState<P8>& s, const FieldTable& F, const PlSink& K, uint32_t sid_lo, uint32_t sid_hi,
                                      const StepArgs& A, uint32_t& dirty, const HumanIn& H) {
    const int X = s.h0 & 0xFF;
    if (T.phase[X].kind == KIND_TERMINAL) return -1;
    dirty |= DIRTY_C0;
    if ((s.h0 >> 16) == 0) { s.h0 = (s.h0 & 0xFFFFu) | (1u << 16); return X; }    // SPEC D11
    if constexpr (PK) s.unpack();
    const int y = w_step_body<P8>(RtView(T, X), X, s, F, K, sid_lo, sid_hi, A, dirty, H);
    if constexpr (PK) s.repack();
    return y;
}

// Specialised entry: one compiled body per phase of a build-time table.
template <int P8, class Spec, int X, bool PK>
__device__ __forceinline__ int w_step_spec_case(WState<P8>& s, const FieldTable& F, const PlSink& K, uint32_t sid_lo, uint32_t sid_hi,
                                                const StepArgs& A, uint32_t& dirty, const HumanIn& H) {
    if constexpr (X >= Spec::n_phases) {
        return -1;
    } else {
        if (Spec::phase(X).kind == KIND_TERMINAL) return -1;
        dirty |= DIRTY_C0;
        if ((s.h0 >> 16) == 0) { s.h0 = (s.h0 & 0xFFFFu) | (1u << 16); return X; }
        if constexpr (PK) s.unpack();
        const int y = w_step_body<P8>(CtView<Spec, X>{}, X, s, F, K, sid_lo, sid_hi, A, dirty, H);
        if constexpr (PK) s.repack();
        return y;
    }
}

// The body of phase X0.  The caller (w_tps_tiles) makes X0 uniform over the lanes that get here — the sessions of a batch
// move in lockstep, so a tile's live sessions are nearly always in ONE phase — which turns the dispatch into one indexed
// branch instead of a chain of up to n_phases compare-and-branch pairs per tile (4-7 % of the step's instructions, ncu).
template <int P8, class Spec, bool PK>
__device__ __forceinline__ int w_step_spec(const int X0, WState<P8>& s, const FieldTable& F, const PlSink& K, uint32_t sid_lo, uint32_t sid_hi,
                                           const StepArgs& A, uint32_t& dirty, const HumanIn& H) {
#define GE_W_CASE(X) case X: return w_step_spec_case<P8, Spec, X, PK>(s, F, K, sid_lo, sid_hi, A, dirty, H);
    switch (X0) {
        GE_W_CASE(0) GE_W_CASE(1) GE_W_CASE(2) GE_W_CASE(3) GE_W_CASE(4) GE_W_CASE(5) GE_W_CASE(6) GE_W_CASE(7)
        GE_W_CASE(8) GE_W_CASE(9) GE_W_CASE(10) GE_W_CASE(11) GE_W_CASE(12) GE_W_CASE(13) GE_W_CASE(14) GE_W_CASE(15)
        GE_W_CASE(16) GE_W_CASE(17) GE_W_CASE(18) GE_W_CASE(19) GE_W_CASE(20) GE_W_CASE(21) GE_W_CASE(22) GE_W_CASE(23)
        GE_W_CASE(24) GE_W_CASE(25) GE_W_CASE(26) GE_W_CASE(27) GE_W_CASE(28) GE_W_CASE(29) GE_W_CASE(30) GE_W_CASE(31)
    default: return -1;
    }
#undef GE_W_CASE
}
// Per-lane dispatch as a balanced tree of compares over the phase index: ceil(log2(n_phases)) compare-and-branch pairs
// instead of up to n_phases (the chain below); lanes of a tile that sit in different phases simply diverge.
template <int P8, class Spec, bool PK, int LO = 0, int HI = Spec::n_phases>
__device__ __forceinline__ int w_step_spec_tree(WState<P8>& s, const FieldTable& F, const PlSink& K, uint32_t sid_lo, uint32_t sid_hi,
                                                const StepArgs& A, uint32_t& dirty, const HumanIn& H) {
    if constexpr (HI - LO <= 1) {
        return w_step_spec_case<P8, Spec, LO, PK>(s, F, K, sid_lo, sid_hi, A, dirty, H);
    } else {
        constexpr int MID = (LO + HI) / 2;
        if ((int)(s.h0 & 0xFF) < MID) return w_step_spec_tree<P8, Spec, PK, LO, MID>(s, F, K, sid_lo, sid_hi, A, dirty, H);
        return w_step_spec_tree<P8, Spec, PK, MID, HI>(s, F, K, sid_lo, sid_hi, A, dirty, H);
    }
}
// A/B twin (-DGE_CHAIN_DISPATCH): the per-lane chain of compare-and-branch pairs over the phases
template <int P8, class Spec, bool PK, int X = 0>
__device__ __forceinline__ int w_step_spec_chain(WState<P8>& s, const FieldTable& F, const PlSink& K, uint32_t sid_lo, uint32_t sid_hi,
                                                 const StepArgs& A, uint32_t& dirty, const HumanIn& H) {
    if constexpr (X >= Spec::n_phases) {
        return -1;
    } else {
        if ((int)(s.h0 & 0xFF) == X) return w_step_spec_case<P8, Spec, X, PK>(s, F, K, sid_lo, sid_hi, A, dirty, H);
        return w_step_spec_chain<P8, Spec, PK, X + 1>(s, F, K, sid_lo, sid_hi, A, dirty, H);
    }
}

// A step of a session whose phase needs nothing but column 0 (UI / timer phases with no effects; the host
// proves this per phase in DevTable::need).  c = {h0, h1, is_alive, can_vote}.  Same SPEC as w_step.
// PK: the packed store (below): c.z = is_alive | can_vote << 8 (or << 16) | ... instead of two 32-bit masks.
template <int P8 = 32, bool PK = false>
__device__ __forceinline__ int w_step_light(const DevTable& T, uint4& c) {
    const int X = c.x & 0xFF;
    const uint32_t f_alive = !PK ? c.z : P8 == 8 ? (c.z & 0xFFu) : (c.z & 0xFFFFu);
    const uint32_t f_vote = !PK ? c.w : P8 == 8 ? ((c.z >> 8) & 0xFFu) : (c.z >> 16);
    const uint32_t step0 = c.x >> 16;
    const ge_phase_t& ph = T.phase[X];
    if (ph.kind == KIND_TERMINAL) return -1;
    if (step0 == 0) { c.x = (c.x & 0xFFFFu) | (1u << 16); return X; }
    const uint32_t prev = (c.x >> 8) & 0xFF;
    int taken = ph.n_branches - 1;
    if (ph.n_branches > 1) {
        const uint32_t ALL = all_mask(T.h.n_players);
        auto lp = [&](int pi) -> uint32_t {                    // predicates of need==0 phases only read fields 0, 1, 15
            uint32_t out = 0;
            for (;; ++pi) {
                const ge_pred_t pr = T.pred[pi];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const uint32_t pos = k ? pr.pos1 : pr.pos0, neg = k ? pr.neg1 : pr.neg0;
                    if (neg & 0x8000u) continue;
                    uint32_t m = ALL;
                    if (pos & 1u) m &= f_alive;
                    if (pos & 2u) m &= f_vote;
                    if (neg & 1u) m &= ~f_alive;
                    if (neg & 2u) m &= ~f_vote;
                    out |= m;
                }
                if (!(pr.pos0 & GE_PRED_CONTINUED)) break;
            }
            return out;
        };
        for (int b = 0; b < ph.n_branches; ++b) {
            const ge_branch_t br = ph.br[b];
            bool ok;
            switch (br.op) {
            case BR_ALWAYS: ok = true; break;
            case BR_COUNT_EQ0: ok = lp(br.a) == 0; break;
            case BR_COUNT_GE: ok = __popc(lp(br.a)) >= __popc(lp((int)br.arg)); break;
            case BR_PREV_IN: ok = (br.arg >> prev) & 1u; break;
            case BR_TIE_PENDING: ok = ((c.y >> 24) & 0x80u) != 0; break;
            default: ok = false; break;
            }
            if (ok) { taken = b; break; }
        }
    }
    const int Y = ph.br[taken].next;
    const uint32_t tag = ph.br[taken].tag;
    if (tag) c.y = (c.y & ~0xFFu) | tag;
    c.x = (uint32_t)Y | ((uint32_t)X << 8) | ((step0 + 1u) << 16);
    return Y;
}

// Column groups a launch must load, from the phases present in the batch (see StepArgs::presence).
__device__ __forceinline__ uint32_t need_of(const DevTable& T, uint32_t present) {
    uint32_t need = 0;
    while (present) { const int i = __ffs(present) - 1; present &= present - 1; need |= T.need[i]; }
    return need;
}

// Shared-memory working set of one CTA of the werewolf thread-per-session kernels: per-thread scratch (predicate
// field table, legal-position table) and one set of block counters per batch the launch steps.
struct BlockCounters {
    uint32_t visits[32];
    uint32_t present, live, mixed;
    // the batch's launch-time inputs, fetched ONCE per block for all batches of the launch (in a ring launch the
    // dependent global loads presence -> need and n_active would otherwise sit in front of every batch's first tile)
    uint32_t present_in;
    unsigned long long n_act, sid0;
};
// all threads; the caller synchronises afterwards.  `slot(k)` gives the k-th batch's arguments.
template <class SlotFn>
__device__ __forceinline__ void counters_init(BlockCounters* c, int n, SlotFn slot) {
    for (int i = threadIdx.x; i < n * (int)(sizeof(BlockCounters) / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(c)[i] = 0;
    // Programmatic dependent launch (GE_OPT_PDL): a step launch that was allowed to start before the previous kernel of
    // its stream has drained waits HERE — after its block prologue, before it reads anything a predecessor wrote (launch
    // inputs, records).  A no-op for ordinary launches.
    grid_dependency_wait();
    __syncthreads();
    if ((int)threadIdx.x < n) {
        const SlotArgs& A = slot((int)threadIdx.x);
        c[threadIdx.x].present_in = A.presence_override ? A.presence_override : A.presence[A.launch_idx % 3];
        c[threadIdx.x].n_act = *A.n_active;            // slots beyond it hold only terminal sessions
        c[threadIdx.x].sid0 = first_sid_of(A);
    }
}
template <int P8, int NSLOT>
struct WSmem {
    BlockCounters c[NSLOT];
    uint32_t fields[16][TPS_THREADS];
    uint8_t lut[P8 > 16 ? P8 : 1][TPS_THREADS];
};
// staging of the bulk-copy variant of the light path: per warp two stages of four 512-byte columns, one mbarrier each
struct LightBulk {
    alignas(16) uint8_t buf[TPS_THREADS / 32][2][4][512];
    alignas(8) uint64_t bar[TPS_THREADS / 32][2];
};

// This CTA's share of ONE batch: every non-terminal session of the batch's active prefix advances by C.n_steps
// steps.  C carries what a ring of batches shares (steps per launch, Philox round keys), A the batch, `bc` the block
// counters of this batch (cleared by the caller).  No block-wide barrier inside: in a ring launch the warps of a CTA
// drift from batch to batch independently.
// Spec = void: interpret the run-time table T.  Spec = a generated ge::spec struct: the same table known at
// build time (the host only selects this instantiation when the blobs are byte-identical).
// HUM: the batch has human seats (SPEC D3h); only the run-time-table instantiations k_step_*_tps_h carry that path, so
// the all-bot kernels are exactly what they were.
// TILED: per-tile column needs (batches with phase regrouping); its own instantiations (k_step_w_tps_tiled), so the
// lockstep kernels do not carry the extra words and branches (measured: -2.3 % on the headline when they did).
// PK: the PACKED session store (werewolf tables up to 16 players; ge_capi.cu GE_OPT_STORE_PACKED): a record is the dense
// wire record (SPEC 5b) in 16-byte columns.  Up to 8 players, 32 bytes: D0 = header + the eight masks is_alive ...
// has_secret_role as bytes, D1 = role_lo, role_hi bytes + the eight target bytes — instead of 3.5 columns whose mask words are
// 3/4 zeros.  Up to 16 players, 48 bytes: D0 = header + is_alive, can_vote, eligible, submitted as u16, D1 = the other six
// masks, D2 = the target bytes — instead of 4 columns.  D0 moves on every step; the others only when DevTable::need says so.
template <int P8, class Spec, bool HUM = false, bool TILED = false, bool PK = false>
__device__ __forceinline__ void w_tps_tiles(const DevTable& T, const StepArgs& C, const SlotArgs& A, BlockCounters& bc,
                                            uint32_t (*s_fields)[TPS_THREADS], uint8_t* s_lut_col, LightBulk* lb = nullptr) {
    static_assert(!PK || P8 == 8 || P8 == 16, "the packed store covers tables up to 16 players");
    constexpr int S = PK ? (P8 == 8 ? 32 : 48) : 48 + P8;
    constexpr int NT16 = P8 / 16;          // full 16-byte target columns
    constexpr bool THALF = (P8 % 16) != 0; // trailing 8-byte column
    uint32_t (&s_visits)[32] = bc.visits;
    const int lane = threadIdx.x & 31;
    const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    // which column groups can any session of this batch need?  (bit0 C1, bit1 C2, bit2 player bytes, bit3 session id)
    const uint32_t present_in = bc.present_in;
    const uint32_t need_batch = C.n_steps > 1 ? 31u : need_of(T, present_in);
    const uint64_t n_act = bc.n_act;
    const uint64_t sid0 = bc.sid0;
    const uint32_t n_tiles_act = (uint32_t)((n_act + 31) >> 5);
    const FieldTable F{s_fields, (int)threadIdx.x, s_lut_col};
    uint32_t present_out = 0, live_cnt = 0, mixed = 0;
    VisitAcc visits;

    // Slots [0, n_act): tiles below full_tiles are entirely inside the prefix, the next one has `rem` slots in it.
    // Tile addresses advance by pointer increments (one 64-bit add per tile instead of a wide multiply per column).
    const uint32_t full_tiles = (uint32_t)(n_act >> 5), rem = (uint32_t)n_act & 31u;
    const uint64_t tile_stride = (uint64_t)nwarps * (32 * S);
    uint8_t* base = A.tiles + (uint64_t)warp0 * (32 * S) + lane * 16;      // column c of this lane: base + c * 512
    if (need_batch == 0) {
        // ---- light path: every present phase touches column 0 only.  Four tiles in flight per warp.
        // Sessions move in lockstep, so usually ONE non-terminal phase x0 is present: its single successor is
        // resolved once here and the common lane only rewrites the header word.
        const uint32_t live_present = present_in & T.nonterm;
        int x0 = -1; uint32_t y0 = 0, tag0 = 0;
        if (__popc(live_present) == 1) {
            const int x = __ffs(live_present) - 1;
            if (T.phase[x].n_branches == 1 && T.phase[x].br[0].op == BR_ALWAYS) { x0 = x; y0 = T.phase[x].br[0].next; tag0 = T.phase[x].br[0].tag; }
        }
        const bool y0_live = x0 >= 0 && ((T.nonterm >> y0) & 1u);
        const uint32_t hdr0 = y0 | ((uint32_t)x0 << 8);
        // A/B variant (STEP_LIGHT_BULK): the four columns of a group arrive by cp.async.bulk into shared memory (one
        // elected lane issues four 512-byte copies, an mbarrier counts the bytes), double-buffered one group ahead,
        // instead of four warp-wide LDG.128.  Measured, not kept as the default: DESIGN section 6.
        const bool bulk = lb != nullptr && (C.flags & STEP_LIGHT_BULK);
        const int wib = threadIdx.x >> 5;
        auto issue_group = [&](uint32_t tile0, const uint8_t* gbase, int stage) {      // lane 0 only
            uint32_t n = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) n += tile0 + j * nwarps < n_tiles_act;
            mbar_expect_tx(&lb->bar[wib][stage], 512u * n);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (tile0 + j * nwarps < n_tiles_act) bulk_g2s(lb->buf[wib][stage][j], gbase - lane * 16 + j * tile_stride, 512u, &lb->bar[wib][stage]);
        };
        uint32_t group = 0;
        if (bulk) {
            if (lane == 0) { mbar_init(&lb->bar[wib][0], 1); mbar_init(&lb->bar[wib][1], 1); mbar_init_fence(); }
            __syncwarp();
            if (lane == 0 && warp0 < n_tiles_act) issue_group(warp0, base, 0);
        }
        for (uint32_t tile = warp0; tile < n_tiles_act; tile += 4 * nwarps, base += 4 * tile_stride, ++group) {
            uint4 c[4];
            if (bulk) {
                const int stage = group & 1;
                if (lane == 0 && tile + 4 * nwarps < n_tiles_act) issue_group(tile + 4 * nwarps, base + 4 * tile_stride, stage ^ 1);
                mbar_wait(&lb->bar[wib][stage], (group >> 1) & 1u);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    c[j] = tile + j * nwarps < n_tiles_act ? *reinterpret_cast<const uint4*>(&lb->buf[wib][stage][j][lane * 16]) : make_uint4(0, 0, 0, 0);
                __syncwarp();                                   // every lane has its copy before the stage is refilled
            } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                c[j] = tile + j * nwarps < n_tiles_act ? ld128(base + j * tile_stride) : make_uint4(0, 0, 0, 0);
            }

#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t t = tile + j * nwarps;
                if (t < n_tiles_act) {                        // warp-uniform
                    const bool in_range = t < full_tiles || (uint32_t)lane < rem;
                    const bool fast = in_range && (int)(c[j].x & 0xFF) == x0 && (c[j].x >> 16) != 0;
                    const uint32_t fastm = __ballot_sync(0xFFFFFFFFu, fast);
                    uint32_t lm;
                    if (fastm == 0xFFFFFFFFu) {               // whole tile in the common phase: nothing else to look at
                        c[j].x = hdr0 | ((c[j].x & 0xFFFF0000u) + 0x10000u);
                        if (tag0) c[j].y = (c[j].y & ~0xFFu) | tag0;
                        st128(base + j * tile_stride, c[j]);
                        if ((int)y0 != visits.phase) { visits.flush(s_visits, lane); visits.phase = (int)y0; }
                        visits.count += 32;
                        present_out |= 1u << y0;
                        lm = y0_live ? 0xFFFFFFFFu : 0u;
                        if (TILED && A.tile_present_out != nullptr && lane == 0) A.tile_present_out[t] = 1u << y0;
                    } else {
                        int np = -1;
                        if (fast) {
                            c[j].x = hdr0 | ((c[j].x & 0xFFFF0000u) + 0x10000u);
                            if (tag0) c[j].y = (c[j].y & ~0xFFu) | tag0;
                            np = (int)y0;
                        } else if (in_range) {
                            np = w_step_light<P8, PK>(T, c[j]);
                        }
                        if (np >= 0) st128(base + j * tile_stride, c[j]);
                        mixed += visits.add(s_visits, np, lane) > 1;
                        if (in_range) present_out |= 1u << (c[j].x & 31);
                        lm = __ballot_sync(0xFFFFFFFFu, in_range && ((T.nonterm >> (c[j].x & 31)) & 1u));
                        if (TILED && A.tile_present_out != nullptr) {
                            const uint32_t bits = __reduce_or_sync(0xFFFFFFFFu, in_range ? 1u << (c[j].x & 31) : 0u);
                            if (lane == 0) A.tile_present_out[t] = bits;
                        }
                    }
                    if (lane == 0) A.live_mask[t] = lm;
                    live_cnt += __popc(lm);
                }
            }
        }
    } else {
        // Per-tile column needs (batches with phase regrouping): between regroups a tile holds one or two phases while
        // the batch as a whole holds many, so the union `need` would make every tile move every column.  The words are
        // read two tiles ahead (one for this tile's loads, one for the next tile's prefetch).
        const uint32_t* tp = (TILED && C.n_steps == 1) ? A.tile_present : nullptr;
        uint32_t w_cur = 0, w_nxt = 0;
        if (tp) {
            if (warp0 < n_tiles_act) w_cur = tp[warp0];
            if (warp0 + nwarps < n_tiles_act) w_nxt = tp[warp0 + nwarps];
        }
        for (uint32_t tile = warp0; tile < n_tiles_act; tile += nwarps, base += tile_stride) {
            uint32_t need = need_batch, need_next = need_batch;
            if (tp) {
                uint32_t w_nn = 0;
                if (tile + 2 * nwarps < n_tiles_act) w_nn = tp[tile + 2 * nwarps];
                need = need_of(T, w_cur) & need_batch;
                need_next = need_of(T, w_nxt) & need_batch;
                w_cur = w_nxt; w_nxt = w_nn;
            }
            const bool in_range = tile < full_tiles || (uint32_t)lane < rem;
            uint32_t org = tile * 32u + lane;
            if (A.origin != nullptr && (need & 8) && in_range) org = A.origin[org];
            WState<P8> s;
            uint8_t* base8 = nullptr;
            // all loads are issued up front (no dependent second round trip)
            if constexpr (PK) {
                // up to 8 players: D0 = header + 8 mask bytes, D1 = role bytes + target bytes (need bit 4);
                // up to 16: D0 = header + alive, can_vote, eligible, submitted, D1 = the other six masks (need bit 4),
                // D2 = the target bytes (need bit 2; night actions store theirs directly, PlSink)
                const uint4 d0 = ld128(base);
                uint4 d1 = make_uint4(0, 0, 0, 0), d2 = make_uint4(0, 0, 0, 0);
                if (need & 16) d1 = ld128(base + 512);
                if (P8 == 16 && (need & 4)) d2 = ld128(base + 1024);
                if (tile + nwarps < n_tiles_act) {                       // next tile -> L1 (see below)
                    prefetch_l1(base + tile_stride);
                    if (need_next & 16) prefetch_l1(base + tile_stride + 512);
                    if (P8 == 16 && (need_next & 4)) prefetch_l1(base + tile_stride + 1024);
                }
                s.h0 = d0.x; s.h1 = d0.y;
                s.pk0 = d0.z;                                            // unpacked inside the step (WState::unpack)
                if constexpr (P8 == 8) {
                    s.pk1 = d0.w; s.pk2 = d1.x; s.pk3 = 0; s.pk4 = 0;
                    s.tw[0] = d1.y; s.tw[1] = d1.z;
                } else {
                    s.pk1 = d0.w; s.pk2 = d1.x; s.pk3 = d1.y; s.pk4 = d1.z;
                    s.tw[0] = d2.x; s.tw[1] = d2.y; s.tw[2] = d2.z; s.tw[3] = d2.w;
                }
            } else {
            const uint4 c0 = ld128(base);
            uint4 c1 = make_uint4(0, 0, 0, 0), c2 = make_uint4(0, 0, 0, 0);
            if (need & 1) c1 = ld128(base + 512);
            if (need & 2) c2 = ld128(base + 1024);
#pragma unroll
            for (int c = 0; c < NT16; ++c) {
                uint4 t = make_uint4(0, 0, 0, 0);
                if (need & 4) t = ld128(base + (3 + c) * 512);
                s.tw[4 * c] = t.x; s.tw[4 * c + 1] = t.y; s.tw[4 * c + 2] = t.z; s.tw[4 * c + 3] = t.w;
            }
            base8 = base - lane * 8 + (3 + NT16) * 512;       // trailing 8-byte column of this lane
            if (THALF) {
                uint2 t = make_uint2(0, 0);
                if (need & 4) t = ld64(base8);
                s.tw[4 * NT16] = t.x; s.tw[4 * NT16 + 1] = t.y;
            }
            // The next tile of this warp: bring its columns to L1 while this one computes, so the warp does not sit out a
            // full DRAM round trip at the top of every iteration (+7 % at 8 players; two tiles ahead, or prefetching
            // in the light path, measured no better).
            if (tile + nwarps < n_tiles_act) {
                const uint8_t* nb = base + tile_stride;
                prefetch_l1(nb);
                if (need_next & 1) prefetch_l1(nb + 512);
                if (need_next & 2) prefetch_l1(nb + 1024);
                if (need_next & 4) {
#pragma unroll
                    for (int c = 0; c < NT16; ++c) prefetch_l1(nb + (3 + c) * 512);
                    if (THALF) prefetch_l1(base8 + tile_stride);
                }
            }
            s.h0 = c0.x; s.h1 = c0.y; s.alive = c0.z; s.can_vote = c0.w;
            s.eligible = c1.x; s.submitted = c1.y; s.revealed = c1.z; s.investigated = c1.w;
            s.wolf = c2.x; s.secret = c2.y; s.role_lo = c2.z; s.role_hi = c2.w;
            }
            // (packed store: the target bytes always travel through registers — D1 is loaded whenever a phase records them)
            const PlSink K{base + (PK ? 2 : 3) * 512, base8, !(PK && P8 == 8) && (need & 4u) == 0};
            bool live = in_range;
            const uint64_t sid = sid0 + org;
            uint32_t dirty = 0;
            HumanIn H{0u, nullptr};
            if constexpr (HUM) {
                if (A.human_mask != nullptr && (need & 8) && in_range) {    // action phases only (they need the session id too)
                    H.mask = A.human_mask[org];
                    H.choice = A.human_choice + (uint64_t)org * A.human_stride;
                }
            }
            for (int it = 0; it < C.n_steps; ++it) {
                int np = -1;
                if constexpr (std::is_void<Spec>::value) {
                    if (live) np = w_step<P8, PK>(T, s, F, K, (uint32_t)sid, (uint32_t)(sid >> 32), C, dirty, H);
                } else {
                    // phase dispatch: a per-lane compare tree (default); A/B twins: the linear chain (-DGE_CHAIN_DISPATCH)
                    // and the warp-uniform switch loop (-DGE_LOOP_DISPATCH) — measured in DESIGN section 6
#if defined(GE_CHAIN_DISPATCH)
                    if (live) np = w_step_spec_chain<P8, Spec, PK>(s, F, K, (uint32_t)sid, (uint32_t)(sid >> 32), C, dirty, H);
#elif !defined(GE_LOOP_DISPATCH)
                    if (live) np = w_step_spec_tree<P8, Spec, PK>(s, F, K, (uint32_t)sid, (uint32_t)(sid >> 32), C, dirty, H);
#else
                    // one iteration per distinct phase among the tile's live sessions (nearly always one): the phase is
                    // uniform over the lanes that enter the switch
                    uint32_t todo = __ballot_sync(0xFFFFFFFFu, live);
                    while (todo) {
                        const int X0 = __shfl_sync(0xFFFFFFFFu, (int)(s.h0 & 0xFFu), __ffs(todo) - 1);
                        const bool mine = ((todo >> lane) & 1u) && (int)(s.h0 & 0xFFu) == X0;
                        if (mine) np = w_step_spec<P8, Spec, PK>(X0, s, F, K, (uint32_t)sid, (uint32_t)(sid >> 32), C, dirty, H);
                        todo &= ~__ballot_sync(0xFFFFFFFFu, mine);
                    }
#endif
                }
                if (live && np < 0) live = false;
                mixed += visits.add(s_visits, np, lane) > 1;
            }
            if (in_range) present_out |= 1u << (s.h0 & 31);
            const uint32_t lm = __ballot_sync(0xFFFFFFFFu, in_range && ((T.nonterm >> (s.h0 & 31)) & 1u));
            if (lane == 0) A.live_mask[tile] = lm;
            if (TILED && A.tile_present_out != nullptr) {          // what the next launch may skip for this tile
                const uint32_t bits = __reduce_or_sync(0xFFFFFFFFu, in_range ? 1u << (s.h0 & 31) : 0u);
                if (lane == 0) A.tile_present_out[tile] = bits;
            }
            live_cnt += __popc(lm);
            if constexpr (PK) {
                if (dirty & (DIRTY_C0 | DIRTY_C1 | DIRTY_C2)) st128(base, make_uint4(s.h0, s.h1, s.pk0, s.pk1));
                if constexpr (P8 == 8) {
                    if (dirty & (DIRTY_C2 | DIRTY_PL)) st128(base + 512, make_uint4(s.pk2, s.tw[0], s.tw[1], 0u));
                } else {
                    if (dirty & (DIRTY_C2 | DIRTY_RV)) st128(base + 512, make_uint4(s.pk2, s.pk3, s.pk4, 0u));
                    if (dirty & DIRTY_PL) st128(base + 1024, make_uint4(s.tw[0], s.tw[1], s.tw[2], s.tw[3]));
                }
            } else {
            if (dirty & DIRTY_C0) st128(base, make_uint4(s.h0, s.h1, s.alive, s.can_vote));
            if (dirty & DIRTY_C1) st128(base + 512, make_uint4(s.eligible, s.submitted, s.revealed, s.investigated));
            if (dirty & DIRTY_C2) st128(base + 1024, make_uint4(s.wolf, s.secret, s.role_lo, s.role_hi));
            if (dirty & DIRTY_PL) {
#pragma unroll
                for (int c = 0; c < NT16; ++c)
                    st128(base + (3 + c) * 512, make_uint4(s.tw[4 * c], s.tw[4 * c + 1], s.tw[4 * c + 2], s.tw[4 * c + 3]));
                if (THALF) st64(base8, make_uint2(s.tw[4 * NT16], s.tw[4 * NT16 + 1]));
            }
            }
        }
    }
    visits.flush(s_visits, lane);
    present_out = __reduce_or_sync(0xFFFFFFFFu, present_out);
    if (lane == 0) {
        if (present_out) atomicOr(&bc.present, present_out);
        if (live_cnt) atomicAdd(&bc.live, live_cnt);
        if (mixed) atomicAdd(&bc.mixed, mixed);
    }
}

// Block epilogue of one batch, after a barrier that follows w_tps_tiles: a handful of global atomics per block.
__device__ __forceinline__ void w_tps_publish(const SlotArgs& A, const BlockCounters& bc) {
    flush_visits(bc.visits, A.stats);
    publish_presence(A, bc.present);
    if (A.count_live && threadIdx.x == 0 && bc.live) atomicAdd(A.live_count, (unsigned long long)bc.live);
    if (A.count_live && A.rg != nullptr) {             // inputs of the phase regrouping check (k_regroup_plan)
        if (threadIdx.x < 32 && bc.visits[threadIdx.x]) atomicAdd(&A.rg[threadIdx.x], bc.visits[threadIdx.x]);
        if (threadIdx.x == 0 && bc.mixed) atomicAdd(&A.rg[32], bc.mixed);
    }
}

template <int P8, class Spec = void, bool PK = false>
__global__ void __launch_bounds__(TPS_THREADS, (w_step_ctas<P8, Spec, PK>()))
k_step_w_tps(const __grid_constant__ DevTable T, const __grid_constant__ StepArgs A) {
    __shared__ WSmem<P8, 1> sm;
    extern __shared__ __align__(16) uint8_t dyn_smem[];       // sizeof(LightBulk) when the launch asked for STEP_LIGHT_BULK
    LightBulk* lb = (A.flags & STEP_LIGHT_BULK) ? reinterpret_cast<LightBulk*>(dyn_smem) : nullptr;
    counters_init(sm.c, 1, [&](int) -> const SlotArgs& { return A; });
    __syncthreads();
    w_tps_tiles<P8, Spec, false, false, PK>(T, A, A, sm.c[0], sm.fields, &sm.lut[0][P8 > 16 ? threadIdx.x : 0], lb);
    grid_launch_dependents();      // this CTA's tiles are done: the stream's next launch may start filling the machine
    __syncthreads();
    w_tps_publish(A, sm.c[0]);
}

// the step kernel with per-tile column needs (batches whose sessions de-synchronise: phase regrouping on)
template <int P8, class Spec = void>
__global__ void __launch_bounds__(TPS_THREADS, GE_W_CTAS(P8))
k_step_w_tps_tiled(const __grid_constant__ DevTable T, const __grid_constant__ StepArgs A) {
    __shared__ WSmem<P8, 1> sm;
    counters_init(sm.c, 1, [&](int) -> const SlotArgs& { return A; });
    __syncthreads();
    w_tps_tiles<P8, Spec, false, true>(T, A, A, sm.c[0], sm.fields, &sm.lut[0][P8 > 16 ? threadIdx.x : 0]);
    grid_launch_dependents();      // this CTA's tiles are done: the stream's next launch may start filling the machine
    __syncthreads();
    w_tps_publish(A, sm.c[0]);
}

// the run-time-table kernel with the human-seat path (batches that have people at the table)
template <int P8>
__global__ void __launch_bounds__(TPS_THREADS, GE_W_CTAS(P8))
k_step_w_tps_h(const __grid_constant__ DevTable T, const __grid_constant__ StepArgs A) {
    __shared__ WSmem<P8, 1> sm;
    counters_init(sm.c, 1, [&](int) -> const SlotArgs& { return A; });
    __syncthreads();
    w_tps_tiles<P8, void, true>(T, A, A, sm.c[0], sm.fields, &sm.lut[0][P8 > 16 ? threadIdx.x : 0]);
    grid_launch_dependents();      // this CTA's tiles are done: the stream's next launch may start filling the machine
    __syncthreads();
    w_tps_publish(A, sm.c[0]);
}

// The ring launch: one step of EVERY batch of a ring (same table, same seed) in one launch.  A batch of 2^20 8-player
// sessions is ~10 us of work, so launched alone a step kernel spends its life in ramp-up and tail (0.43 of the
// roofline on one stream).  Here each CTA walks the batches in order (starting at a rotated one): one ramp and one
// tail per RING pass, full-occupancy grid, one stream.  Measured: 4.4e10 steps/s against 3.0e10
// for separate launches on one stream and 6.3e10 for separate launches on eight streams, which remains the default:
// the ring kernel itself issues at 64 % of the scheduler peak (ncu, profiles/r02_ring_*), but the compaction and
// re-initialisation launches that sit between ring passes on the single stream are no longer hidden.
template <int P8, class Spec = void, bool PK = false>
__global__ void __launch_bounds__(TPS_THREADS, GE_W_CTAS(P8))
k_ring_w_tps(const __grid_constant__ DevTable T, const __grid_constant__ StepArgs C, const __grid_constant__ RingArgs R) {
    __shared__ WSmem<P8, GE_RING_MAX> sm;
    counters_init(sm.c, R.n, [&](int k) -> const SlotArgs& { return R.slot[k]; });
    __syncthreads();
    int i = ring_first_slot(R);
    for (int k = 0; k < R.n; ++k) {                    // no barrier between batches: warps drift freely
        w_tps_tiles<P8, Spec, false, false, PK>(T, C, R.slot[i], sm.c[i], sm.fields, &sm.lut[0][P8 > 16 ? threadIdx.x : 0]);
        i = i + 1 == R.n ? 0 : i + 1;
    }
    grid_launch_dependents();
    __syncthreads();
    for (int k = 0; k < R.n; ++k) w_tps_publish(R.slot[k], sm.c[k]);
}

// =================================================================================== TTL family
// Device record for player bucket PB: 8-byte header + PB player words (score|rounds<<8|vote<<16|flags<<24).
template <int PB>
struct TState {
    uint32_t h0, h1;              // phase|prev<<8|step<<16 ; speaker|lie<<8|winner<<16
    uint32_t pw[PB];
};

enum { TF_SPEAKER = 1, TF_STMTS = 2, TF_REVEALED = 4, TF_CANVOTE = 8, TF_VOTED = 16 };

// The per-player flag bytes of four players gathered in one word (byte j = player 4i+j).  A mask field is
// then one shift + AND away ("byte-sliced"), and converting it to / from a lane mask (bit p = player p) is one
// multiply: gather  ((w >> f) & 0x01010101) * 0x01020408 >> 24,  spread  (m * 0x00204081) & 0x01010101.
template <int PB>
struct TFlags {
    static constexpr int NW = PB / 4;
    uint32_t w[NW];
    __device__ __forceinline__ void load(const uint32_t (&pw)[PB]) {
#pragma unroll
        for (int i = 0; i < NW; ++i)
            w[i] = __byte_perm(__byte_perm(pw[4 * i], pw[4 * i + 1], 0x0073), __byte_perm(pw[4 * i + 2], pw[4 * i + 3], 0x0073), 0x5410);
    }
    __device__ __forceinline__ void store(uint32_t (&pw)[PB]) const {
#pragma unroll
        for (int p = 0; p < PB; ++p) pw[p] = __byte_perm(pw[p], w[p >> 2], ((4u + (p & 3)) << 12) | 0x210u);
    }
    __device__ __forceinline__ uint32_t get(int f) const {
        uint32_t m = 0;
#pragma unroll
        for (int i = 0; i < NW; ++i) m |= ((((w[i] >> f) & 0x01010101u) * 0x01020408u) >> 24) << (4 * i);
        return m;
    }
    __device__ __forceinline__ void set(int f, uint32_t m) {
#pragma unroll
        for (int i = 0; i < NW; ++i)
            w[i] = (w[i] & ~(0x01010101u << f)) | (((((m >> (4 * i)) & 0xFu) * 0x00204081u) & 0x01010101u) << f);
    }
};

template <class V, class FieldFn>
__device__ __forceinline__ uint32_t t_pred(const V& v, FieldFn field, int pi, uint32_t ALL) {
    uint32_t out = 0;
#pragma unroll 1
    for (;; ++pi) {                                        // a continued predicate is a run of records, ORed
        const ge_pred_t pr = v.pred_rec(pi);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint32_t pos = c ? pr.pos1 : pr.pos0, neg = c ? pr.neg1 : pr.neg0;
            if (neg & 0x8000u) continue;                   // "& ~ALL": unused clause
            uint32_t m = ALL;
#pragma unroll
            for (int f = 0; f < 15; ++f) {                 // 0-4 flag fields, 11-14 comparison fields (the rest: never set)
                if (f >= 5 && f < 11) continue;
                if ((pos >> f) & 1u) m &= field(f);
                if ((neg >> f) & 1u) m &= ~field(f);
            }
            out |= m;
        }
        if (V::is_const || !(pr.pos0 & GE_PRED_CONTINUED)) break;      // build-time tables have no continued records
    }
    return out;
}

// One step of one TTL session in phase X (not terminal, step0 != 0), against a table view (RtView: the run-time
// table; CtView<Spec, X>: a build-time table, every accessor a constant).  Returns the phase entered.
template <int PB, class V>
__device__ __forceinline__ int t_step_body(const V v, const int X, TState<PB>& s, uint32_t sid_lo, uint32_t sid_hi,
                                           const StepArgs& A, uint32_t& dirty, const HumanIn& H) {
    const int P = v.n_players();
    const uint32_t ALL = all_mask(P);
    const uint32_t step0 = s.h0 >> 16;
    uint32_t speaker = s.h1 & 0xFF, lie = (s.h1 >> 8) & 0xFF, winner = (s.h1 >> 16) & 0xFF;
    TFlags<PB> F;
    F.load(s.pw);
    bool flags_dirty = false;
    // comparison fields of a run-time table (numeric conditions): "value field <op> constant" per player, computed
    // once per step; build-time tables have none (specgen asserts it)
    uint32_t cm[4] = {0u, 0u, 0u, 0u};
    if constexpr (!V::is_const) {
#pragma unroll 1
        for (int k = 0; k < v.T.h.n_cmp; ++k) {
            const ge_cmp_t c = v.T.h.cmp[k];
            uint32_t m = 0;
#pragma unroll
            for (int p = 0; p < PB; ++p)
                if (p < P && cmp_holds(c.op, (s.pw[p] >> (8 * c.value_field)) & 0xFFu, c.constant)) m |= 1u << p;
            cm[k] = m;
        }
    }
    auto field = [&](int f) -> uint32_t { return f < 5 ? F.get(f) : (f >= 11 && f < 15) ? cm[f - 11] : 0u; };

    const int nb = v.n_branches();
    int taken = nb - 1;
    if (nb > 1) {
        bool done = false;
#pragma unroll(V::is_const ? 4 : 1)
        for (int b = 0; b < 4; ++b) {
            if (b < nb && !done) {
                const ge_branch_t br = v.branch(b);
                bool ok;
                switch (br.op) {
                case BR_ALWAYS: ok = true; break;
                case BR_COUNT_EQ0: ok = t_pred(v, field, br.a, ALL) == 0; break;
                case BR_COUNT_GE: ok = __popc(t_pred(v, field, br.a, ALL)) >= __popc(t_pred(v, field, (int)br.arg, ALL)); break;
                case BR_PREV_IN: ok = (br.arg >> ((s.h0 >> 8) & 0xFF)) & 1u; break;
                case BR_ALL_VAL_GE: {
                    ok = true;
#pragma unroll
                    for (int p = 0; p < PB; ++p)
                        if (p < P && ((s.pw[p] >> (8 * br.a)) & 0xFFu) < br.arg) ok = false;
                } break;
                default: ok = false; break;
                }
                if (ok) { taken = b; done = true; }
            }
        }
    }
    int Y = 0; uint32_t tag = 0;
    if constexpr (V::is_const) {
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (b < nb && b == taken) { Y = v.branch(b).next; tag = v.branch(b).tag; }
    } else {
        Y = v.branch(taken).next; tag = v.branch(taken).tag;
    }

    if (v.kind() == KIND_ACTION) {
        const uint32_t actors = t_pred(v, field, v.actor_pred(), ALL);
        const int aop = v.action_op(), exo = v.exit_op();
        const uint32_t legal0 = aop == ACT_PICK_PLAYER ? t_pred(v, field, v.action_arg(), ALL) : 0u;
        uint32_t first_choice = 0; bool have_first = false;
        const uint32_t hact = actors & H.mask;           // human seats among the actors (SPEC D3h)
        if (hact) {
            uint32_t rem_h = hact;
            bool waiting = false;
            while (rem_h) {
                const int p = __ffs(rem_h) - 1;
                rem_h &= rem_h - 1;
                if (human_choice_of(H, p, aop, v.action_arg(), (v.action_flags() & 1) ? legal0 & ~(1u << p) : legal0) < 0) waiting = true;
            }
            if (waiting) {
                s.h0 = (uint32_t)X | ((uint32_t)X << 8) | ((step0 + 1u) << 16);
                return X;
            }
        }
#pragma unroll
        for (int b = 0; b < PB / 4; ++b) {
            const uint32_t ab = (actors >> (4 * b)) & 0xFu;
            if (ab) {
                uint4 r4 = make_uint4(0, 0, 0, 0);
                if (aop != ACT_MARK) r4 = philox4x32_10(sid_lo, sid_hi, step0, (uint32_t)b, A.rk);     // MARK draws nothing
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int p = 4 * b + j;
                    if ((ab >> j) & 1u) {
                        uint32_t choice;
                        if ((hact >> p) & 1u) {
                            choice = (uint32_t)human_choice_of(H, p, aop, v.action_arg(), (v.action_flags() & 1) ? legal0 & ~(1u << p) : legal0);
                        } else if (aop == ACT_PICK_PLAYER) {
                            const uint32_t legal = (v.action_flags() & 1) ? legal0 & ~(1u << p) : legal0;
                            const uint32_t n = __popc(legal);
                            choice = n ? 1u + (uint32_t)kth_set_bit<PB>(legal, __umulhi(word_of(r4, j), n)) : 0u;
                        } else {
                            choice = aop == ACT_PICK_OPTION ? 1u + __umulhi(word_of(r4, j), (uint32_t)v.action_arg()) : 1u;
                        }
                        if (!have_first) { first_choice = choice; have_first = true; }
                        if (exo == EX_T_VOTES) s.pw[p] = (s.pw[p] & 0xFF00FFFFu) | (choice << 16);
                    }
                }
            }
        }
        switch (exo) {
        case EX_T_STATEMENTS: F.set(1, F.get(1) | actors); flags_dirty = true; break;
        case EX_T_LIE: if (have_first) lie = first_choice; break;
        case EX_T_VOTES: F.set(4, F.get(4) | actors); flags_dirty = true; break;
        default: break;
        }
    }

    const int en = v.entry_op_after(taken);
    if (en == EN_T_ROUND_START) {
        speaker = 0;
#pragma unroll
        for (int p = PB - 1; p >= 0; --p)
            if (p < P && ((s.pw[p] >> 8) & 0xFFu) < (uint32_t)v.rounds()) speaker = p + 1;
        lie = 0;
        const uint32_t sp = speaker ? 1u << (speaker - 1) : 0u;
#pragma unroll
        for (int i = 0; i < PB / 4; ++i) F.w[i] = 0;
        F.set(0, sp);
        F.set(3, ALL & ~sp);
        flags_dirty = true;
#pragma unroll
        for (int p = 0; p < PB; ++p) s.pw[p] &= 0xFF00FFFFu;
    } else if (en == EN_T_REVEAL) {
        F.set(2, F.get(2) | F.get(0));
        flags_dirty = true;
    } else if (en == EN_T_SCORE) {
        const uint32_t voters = F.get(4) & F.get(3) & ~F.get(0);
        uint32_t fooled = 0;
#pragma unroll
        for (int p = 0; p < PB; ++p) {
            if ((voters >> p) & 1u) {
                if (((s.pw[p] >> 16) & 0xFFu) == lie) s.pw[p] = (s.pw[p] & ~0xFFu) | ((s.pw[p] + 1u) & 0xFFu);
                else fooled++;
            }
        }
#pragma unroll
        for (int p = 0; p < PB; ++p) {
            if ((uint32_t)(p + 1) == speaker) {
                const uint32_t sc = ((s.pw[p] & 0xFFu) + fooled) & 0xFFu;
                const uint32_t rd = (((s.pw[p] >> 8) & 0xFFu) + 1u) & 0xFFu;
                s.pw[p] = (s.pw[p] & 0xFFFF0000u) | sc | (rd << 8);
            }
        }
        dirty |= DIRTY_PL;
    } else if (en == EN_T_FINAL) {
        uint32_t best = s.pw[0] & 0xFFu; winner = 1;
#pragma unroll
        for (int p = 1; p < PB; ++p)
            if (p < P && (s.pw[p] & 0xFFu) > best) { best = s.pw[p] & 0xFFu; winner = p + 1; }
    }
    if (tag) winner = tag;
    if (v.kind() == KIND_ACTION && v.exit_op() == EX_T_VOTES) dirty |= DIRTY_PL;
    if (flags_dirty) { F.store(s.pw); dirty |= DIRTY_PL; }
    s.h1 = speaker | (lie << 8) | (winner << 16);
    s.h0 = (uint32_t)Y | ((uint32_t)X << 8) | ((step0 + 1u) << 16);
    return Y;
}

template <int PB>
__device__ __forceinline__ int t_step(const DevTable& T, TState<PB>& s, uint32_t sid_lo, uint32_t sid_hi,
                                      const StepArgs& A, uint32_t& dirty, const HumanIn& H) {
    const int X = s.h0 & 0xFF;
    if (T.phase[X].kind == KIND_TERMINAL) return -1;
    dirty |= DIRTY_C0;
    if ((s.h0 >> 16) == 0) { s.h0 = (s.h0 & 0xFFFFu) | (1u << 16); return X; }    // SPEC D11
    return t_step_body<PB>(RtView(T, X), X, s, sid_lo, sid_hi, A, dirty, H);
}

template <int PB, class Spec, int X = 0>
__device__ __forceinline__ int t_step_spec(TState<PB>& s, uint32_t sid_lo, uint32_t sid_hi, const StepArgs& A, uint32_t& dirty, const HumanIn& H) {
    if constexpr (X >= Spec::n_phases) {
        return -1;
    } else {
        if ((int)(s.h0 & 0xFF) == X) {
            if (Spec::phase(X).kind == KIND_TERMINAL) return -1;
            dirty |= DIRTY_C0;
            if ((s.h0 >> 16) == 0) { s.h0 = (s.h0 & 0xFFFFu) | (1u << 16); return X; }
            return t_step_body<PB>(CtView<Spec, X>{}, X, s, sid_lo, sid_hi, A, dirty, H);
        }
        return t_step_spec<PB, Spec, X + 1>(s, sid_lo, sid_hi, A, dirty, H);
    }
}

// Column 0 (16 bytes) holds the header and the first two player words; the host proves per phase whether a
// step needs the player words at all (DevTable::need bit 2): launches whose sessions are all in header-only
// phases move one column in and out instead of the whole record.
// This CTA's share of one TTL batch (see w_tps_tiles: no block-wide barrier inside).
template <int PB, class Spec, bool HUM = false>
__device__ __forceinline__ void t_tps_tiles(const DevTable& T, const StepArgs& C, const SlotArgs& A, BlockCounters& bc) {
    constexpr int S = 8 + 4 * PB;           // device record (PB even => S % 8 == 0)
    constexpr int NW = S / 4;
    constexpr int N16 = S / 16;
    constexpr bool HALF = (S % 16) != 0;
    uint32_t (&s_visits)[32] = bc.visits;
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t present_in = bc.present_in;
    const uint32_t need = C.n_steps > 1 ? 15u : need_of(T, present_in);
    const bool full = (need & 4u) != 0;
    const bool use_origin = A.origin != nullptr && ((need & 8u) || (HUM && A.human_mask != nullptr));
    const uint64_t n_act = bc.n_act;
    const uint64_t sid0 = bc.sid0;
    const uint64_t n_tiles_act = (n_act + 31) >> 5;
    uint32_t present_out = 0;
    VisitAcc visits;
    if (!full) {
        // ---- light path: every phase present only rewrites the header (no actors, no predicate or value test, no
        // entry effect behind its branches: DevTable::need).  Four tiles in flight per warp; a tile whose 32 sessions
        // are all in the one live phase present takes a branch-free route, like the werewolf kernel's light path.
        const uint32_t live_present = present_in & T.nonterm;
        int x0 = -1; uint32_t y0 = 0, tag0 = 0;
        if (__popc(live_present) == 1) {
            const int x = __ffs(live_present) - 1;
            if (T.phase[x].n_branches == 1 && T.phase[x].br[0].op == BR_ALWAYS) { x0 = x; y0 = T.phase[x].br[0].next; tag0 = T.phase[x].br[0].tag; }
        }
        const bool y0_live = x0 >= 0 && ((T.nonterm >> y0) & 1u);
        const uint32_t hdr0 = y0 | ((uint32_t)x0 << 8);
        const uint64_t full_tiles = n_act >> 5;
        const uint32_t rem = (uint32_t)n_act & 31u;
        const uint64_t tile_stride = nwarps * (uint64_t)(32 * S);
        uint8_t* base = A.tiles + warp0 * (uint64_t)(32 * S) + lane * 16;
        auto light = [&](uint4& c) -> int {             // one header-only step from the run-time table; -1 = terminal
            const int X = c.x & 0xFF;
            const ge_phase_t& ph = T.phase[X];
            if (ph.kind == KIND_TERMINAL) return -1;
            if ((c.x >> 16) == 0) { c.x = (c.x & 0xFFFFu) | (1u << 16); return X; }
            int taken = ph.n_branches - 1;
            for (int b = 0; b < ph.n_branches; ++b) {
                const ge_branch_t br = ph.br[b];
                const bool ok = br.op == BR_ALWAYS || (br.op == BR_PREV_IN && ((br.arg >> ((c.x >> 8) & 0xFF)) & 1u));
                if (ok) { taken = b; break; }
            }
            const uint32_t tag = ph.br[taken].tag;
            if (tag) c.y = (c.y & ~0xFF0000u) | (tag << 16);
            c.x = (uint32_t)ph.br[taken].next | ((uint32_t)X << 8) | ((c.x & 0xFFFF0000u) + 0x10000u);
            return ph.br[taken].next;
        };
        for (uint64_t tile = warp0; tile < n_tiles_act; tile += 4 * nwarps, base += 4 * tile_stride) {
            uint4 c[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                c[j] = tile + j * nwarps < n_tiles_act ? ld128(base + j * tile_stride) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint64_t t = tile + j * nwarps;
                if (t < n_tiles_act) {                        // warp-uniform
                    const bool in_range = t < full_tiles || (uint32_t)lane < rem;
                    const bool fast = in_range && (int)(c[j].x & 0xFF) == x0 && (c[j].x >> 16) != 0;
                    uint32_t lm;
                    if (__ballot_sync(0xFFFFFFFFu, fast) == 0xFFFFFFFFu) {
                        c[j].x = hdr0 | ((c[j].x & 0xFFFF0000u) + 0x10000u);
                        if (tag0) c[j].y = (c[j].y & ~0xFF0000u) | (tag0 << 16);
                        st128(base + j * tile_stride, c[j]);
                        if ((int)y0 != visits.phase) { visits.flush(s_visits, lane); visits.phase = (int)y0; }
                        visits.count += 32;
                        present_out |= 1u << y0;
                        lm = y0_live ? 0xFFFFFFFFu : 0u;
                    } else {
                        int np = -1;
                        if (in_range) np = light(c[j]);
                        if (np >= 0) st128(base + j * tile_stride, c[j]);
                        visits.add(s_visits, np, lane);
                        if (in_range) present_out |= 1u << (c[j].x & 31);
                        lm = __ballot_sync(0xFFFFFFFFu, in_range && ((T.nonterm >> (c[j].x & 31)) & 1u));
                    }
                    if (lane == 0) { A.live_mask[t] = lm; if (A.count_live && lm) atomicAdd(A.live_count, (unsigned long long)__popc(lm)); }
                }
            }
        }
    } else
    for (uint64_t tile = warp0; tile < n_tiles_act; tile += nwarps) {
        uint8_t* base = A.tiles + tile * (uint64_t)(32 * S);
        const uint64_t sess = tile * 32 + lane;
        const bool in_range = sess < n_act;
        uint64_t org = sess;
        if (use_origin && in_range) org = A.origin[sess];
        uint32_t w[NW];
        {
            const uint4 v = ld128(base + lane * 16);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        }
#pragma unroll
        for (int c = 1; c < N16; ++c) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (full) v = ld128(base + c * 512 + lane * 16);
            w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
        }
        if (HALF) {
            uint2 v = make_uint2(0, 0);
            if (full) v = ld64(base + N16 * 512 + lane * 8);
            w[4 * N16] = v.x; w[4 * N16 + 1] = v.y;
        }
        TState<PB> s;
        s.h0 = w[0]; s.h1 = w[1];
#pragma unroll
        for (int p = 0; p < PB; ++p) s.pw[p] = w[2 + p];
        bool live = in_range;
        const uint64_t sid = sid0 + org;
        uint32_t dirty = 0;
        HumanIn H{0u, nullptr};
        if constexpr (HUM) {
            if (A.human_mask != nullptr && in_range) {
                H.mask = A.human_mask[org];
                H.choice = A.human_choice + org * A.human_stride;
            }
        }
        for (int it = 0; it < C.n_steps; ++it) {
            int np = -1;
            if (live) {
                if constexpr (std::is_void<Spec>::value)
                    np = t_step<PB>(T, s, (uint32_t)sid, (uint32_t)(sid >> 32), C, dirty, H);
                else
                    np = t_step_spec<PB, Spec>(s, (uint32_t)sid, (uint32_t)(sid >> 32), C, dirty, H);
                if (np < 0) live = false;
            }
            visits.add(s_visits, np, lane);
        }
        if (in_range) present_out |= 1u << (s.h0 & 31);
        {
            const uint32_t lm = __ballot_sync(0xFFFFFFFFu, in_range && T.phase[s.h0 & 31].kind != KIND_TERMINAL);
            if (lane == 0) { A.live_mask[tile] = lm; if (A.count_live && lm) atomicAdd(A.live_count, (unsigned long long)__popc(lm)); }
        }
        if (dirty) {
            w[0] = s.h0; w[1] = s.h1;
#pragma unroll
            for (int p = 0; p < PB; ++p) w[2 + p] = s.pw[p];
            st128(base + lane * 16, make_uint4(w[0], w[1], w[2], w[3]));
            if (dirty & DIRTY_PL) {
#pragma unroll
                for (int c = 1; c < N16; ++c) st128(base + c * 512 + lane * 16, make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]));
                if (HALF) st64(base + N16 * 512 + lane * 8, make_uint2(w[4 * N16], w[4 * N16 + 1]));
            }
        }
    }
    visits.flush(s_visits, lane);
    present_out = __reduce_or_sync(0xFFFFFFFFu, present_out);
    if (lane == 0 && present_out) atomicOr(&bc.present, present_out);
}

__device__ __forceinline__ void t_tps_publish(const SlotArgs& A, const BlockCounters& bc) {
    flush_visits(bc.visits, A.stats);
    publish_presence(A, bc.present);
}

template <int PB, class Spec = void>
__global__ void __launch_bounds__(128)
k_step_t_tps(const __grid_constant__ DevTable T, const __grid_constant__ StepArgs A) {
    __shared__ BlockCounters bc[1];
    counters_init(bc, 1, [&](int) -> const SlotArgs& { return A; });
    __syncthreads();
    t_tps_tiles<PB, Spec>(T, A, A, bc[0]);
    grid_launch_dependents();
    __syncthreads();
    t_tps_publish(A, bc[0]);
}

template <int PB>
__global__ void __launch_bounds__(128)
k_step_t_tps_h(const __grid_constant__ DevTable T, const __grid_constant__ StepArgs A) {
    __shared__ BlockCounters bc[1];
    counters_init(bc, 1, [&](int) -> const SlotArgs& { return A; });
    __syncthreads();
    t_tps_tiles<PB, void, true>(T, A, A, bc[0]);
    grid_launch_dependents();
    __syncthreads();
    t_tps_publish(A, bc[0]);
}

// ring launch of the TTL family (see k_ring_w_tps)
template <int PB, class Spec = void>
__global__ void __launch_bounds__(128)
k_ring_t_tps(const __grid_constant__ DevTable T, const __grid_constant__ StepArgs C, const __grid_constant__ RingArgs R) {
    __shared__ BlockCounters bc[GE_RING_MAX];
    counters_init(bc, R.n, [&](int k) -> const SlotArgs& { return R.slot[k]; });
    __syncthreads();
    int i = ring_first_slot(R);
    for (int k = 0; k < R.n; ++k) {
        t_tps_tiles<PB, Spec>(T, C, R.slot[i], bc[i]);
        i = i + 1 == R.n ? 0 : i + 1;
    }
    grid_launch_dependents();
    __syncthreads();
    for (int k = 0; k < R.n; ++k) t_tps_publish(R.slot[k], bc[k]);
}

}  // namespace ge
