// ge_common.cuh — device-side building blocks shared by the two step-kernel mappings.
//
// Data layout in HBM ("session store"): sessions are grouped in tiles of 32.  A canonical record of
// S bytes (SPEC.md section 5, S % 8 == 0) is split into N16 = S/16 columns of 16 bytes plus, when
// S % 16 == 8, one trailing column of 8 bytes.  Column c of a tile is stored contiguously for the 32
// sessions of the tile:   tile_base + c*512 + lane*16   (trailing column: tile_base + N16*512 + lane*8).
// A warp that owns one tile therefore reads/writes every column with one fully coalesced 128-bit
// (or 64-bit) access per lane: 512 contiguous bytes per instruction, no padding bytes moved.
// Werewolf tables up to 8 / 16 players keep the PACKED record instead (the 32 / 48 bytes of the dense wire format, two /
// three 16-byte columns, same tile addressing; ge_step_tps.cuh, ge_capi.cu ensure_store).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/game_engine_b200.h"

namespace ge {

enum { FAM_WEREWOLF = 1, FAM_TTL = 2 };
enum { KIND_UI = 0, KIND_TIMER = 1, KIND_ACTION = 2, KIND_TERMINAL = 3 };
enum { ACT_NONE = 0, ACT_PICK_PLAYER = 1, ACT_PICK_OPTION = 2, ACT_MARK = 3 };
enum { EX_NONE = 0, EX_VOTE_KILL = 1, EX_PROTECT = 2, EX_INVESTIGATE_RESOLVE = 3, EX_DAY_VOTE = 4,
       EX_T_STATEMENTS = 16, EX_T_LIE = 17, EX_T_VOTES = 18 };
enum { EN_NONE = 0, EN_ASSIGN_ROLES = 1, EN_NIGHT_RESET = 2,
       EN_T_ROUND_START = 16, EN_T_REVEAL = 17, EN_T_SCORE = 18, EN_T_FINAL = 19 };
enum { BR_ALWAYS = 0, BR_COUNT_EQ0 = 1, BR_COUNT_GE = 2, BR_PREV_IN = 3, BR_ALL_VAL_GE = 4, BR_TIE_PENDING = 5 };

// stats word offsets (SPEC.md section 6)
enum { ST_COUNTED = 0, ST_WINNER = 1, ST_LENGTH = 4, ST_VISITS = 260, ST_TAIL = 292 };

// The compiled table travels to the kernels BY VALUE as a __grid_constant__ parameter: it lands in
// the constant bank (uniform, cached) and there is no module-global symbol to race on between handles.
struct DevTable {
    ge_table_header_t h;
    ge_phase_t phase[GE_MAX_PHASES];
    ge_pred_t pred[GE_MAX_PREDS];
    // per phase: which column groups a step that starts in this phase must READ besides column 0
    // (bit0 = dynamic masks column, bit1 = role/team column, bit2 = per-player bytes, bit3 = session id; bit4 = column D1
    // of the PACKED store: role bytes + target bytes); host-computed.
    uint8_t need[GE_MAX_PHASES];
    uint32_t nonterm;          // bit i: phase index i is not terminal (host-computed)
};
static_assert(sizeof(ge_phase_t) == 48 && sizeof(ge_pred_t) == 8 && sizeof(ge_table_header_t) == 32, "table ABI");
static_assert(sizeof(DevTable) <= 4000, "table must fit the kernel parameter space");

// Per-launch arguments of every step kernel: the part that belongs to ONE batch (SlotArgs) and the part a ring of
// batches stepped by one launch shares (StepArgs adds it: seed, steps per launch, Philox round keys).
struct SlotArgs {
    uint8_t* tiles;
    uint64_t n_sessions, n_tiles, first_sid;
    unsigned long long* stats;
    // Phase-presence masks (3 rotating words): launch k READS word k%3 (bit i = some session was in
    // phase index i after launch k-1), ORs the phases its sessions end in into word (k+1)%3 and clears
    // word (k+2)%3 for launch k+1.  presence_override != 0 replaces the read (first launch after a
    // reset or an import, when the device words are not valid).
    uint32_t* presence;
    uint32_t launch_idx, presence_override;
    // Active-prefix compaction (ge_capi.cu): only slots [0, *n_active) can hold live sessions; `origin`
    // maps a slot to the session's original index (NULL = identity; session id = first_sid + origin);
    // the step kernel publishes which lanes of each tile are still live after the step in live_mask.
    const unsigned long long* n_active;
    const uint32_t* origin;
    uint32_t* live_mask;
    unsigned long long* live_count;   // when count_live != 0 the launch adds its number of live sessions here
    uint32_t count_live;
    // Phase regrouping (ge_capi.cu, k_regroup_*): on a counted launch (count_live != 0) the werewolf
    // thread-per-session kernel also adds, per phase index, the sessions that entered it to rg[0..31] and the
    // number of tiles whose live sessions sit in more than one phase to rg[32].  NULL = not collected.
    uint32_t* rg;
    // Per-TILE phase presence (tables whose sessions de-synchronise, i.e. batches with phase regrouping): word t = the
    // phases the 32 sessions of tile t were in after the previous launch.  tile_present_out (NULL = not kept) is written
    // by every launch; tile_present (NULL = not valid: first launch after a reset / import) lets a tile load only the
    // columns ITS phases need instead of the union over the whole batch.
    const uint32_t* tile_present;
    uint32_t* tile_present_out;
    // Human seats (SPEC.md section 1, D3h): human_mask[i] = seats of session i (original index) played by people,
    // human_choice[i * stride + p] = the input of seat p+1 for THIS step (0xFF = has not acted).  NULL = all bots.
    const uint32_t* human_mask;
    const uint8_t* human_choice;
    uint32_t human_stride;
    // Auto-reset (ge_capi.cu, k_autoreset_*): n_active[8] counts the device-side re-initialisations; the session
    // in slot i then has id first_sid + n_active[8] * sid_stride + origin[i].  0 = off.
    uint64_t sid_stride;
};
enum { STEP_LIGHT_BULK = 1 };     // StepArgs::flags
struct StepArgs : SlotArgs {
    uint64_t seed;
    int n_steps;
    uint32_t flags;               // STEP_LIGHT_BULK: header-only launches fetch their tiles with cp.async.bulk + mbarrier (A/B knob)
    // Philox round keys (k0 + r*W0, k1 + r*W1 for r = 0..9), expanded once on the host: as kernel parameters they
    // are constant-bank operands of the round's XOR, so the key schedule costs no instructions.
    uint32_t rk[20];
};
// A ring of independent batches of the SAME table and seed, stepped once each by ONE launch (k_ring_*): the launch
// walks the slots in order; nothing orders one slot against another (their sessions are independent).
constexpr int GE_RING_MAX = 16;
struct RingArgs {
    int n;
    // CTA b starts at slot (b + b / rot_div) % n (rot_div = SM count): the CTAs resident on one SM then work on
    // DIFFERENT batches at any moment, so a batch in a bandwidth-bound phase (header-only steps) shares the SM with
    // one in an issue-bound phase (votes) — the overlap that separate launches on separate streams used to give.
    uint32_t rot_div;
    SlotArgs slot[GE_RING_MAX];
};
__device__ __forceinline__ int ring_first_slot(const RingArgs& R) {
    return (int)((blockIdx.x + blockIdx.x / R.rot_div) % (uint32_t)R.n);
}

__device__ __forceinline__ uint64_t first_sid_of(const SlotArgs& A) {
    return A.sid_stride ? A.first_sid + A.n_active[8] * A.sid_stride : A.first_sid;
}

__device__ __forceinline__ void publish_presence(const SlotArgs& A, uint32_t block_present) {
    // called by all threads after a __syncthreads() that orders the block's shared accumulation
    if (threadIdx.x == 0) {
        if (block_present) atomicOr(&A.presence[(A.launch_idx + 1) % 3], block_present);
        if (blockIdx.x == 0) A.presence[(A.launch_idx + 2) % 3] = 0;
    }
}

// ---- Philox4x32-10 (Random123 constants; SPEC.md section 3) -------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
// same function with the round keys pre-expanded (StepArgs::rk)
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t (&rk)[20]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ rk[2 * r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ rk[2 * r + 1];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint32_t word_of(const uint4& v, int j) {
    return j == 0 ? v.x : j == 1 ? v.y : j == 2 ? v.z : v.w;
}

// index (from bit 0) of the k-th (0-based) set bit of m; BITS = number of low bits that can be set
template <int BITS>
__device__ __forceinline__ int kth_set_bit(uint32_t m, uint32_t k) {
    int pos = 0;
    if (BITS > 16) { const uint32_t c = __popc(m & 0xFFFFu); if (k >= c) { k -= c; pos += 16; m >>= 16; } }
    if (BITS > 8)  { const uint32_t c = __popc(m & 0xFFu);   if (k >= c) { k -= c; pos += 8;  m >>= 8; } }
    if (BITS > 4)  { const uint32_t c = __popc(m & 0xFu);    if (k >= c) { k -= c; pos += 4;  m >>= 4; } }
    { const uint32_t c = __popc(m & 0x3u); if (k >= c) { k -= c; pos += 2; m >>= 2; } }
    { const uint32_t c = m & 1u; if (k >= c) pos += 1; }
    return pos;
}

// What a step needs to know about the people at the table: which seats are human and where their inputs are.
struct HumanIn {
    uint32_t mask;
    const uint8_t* choice;      // this session's row of SlotArgs::human_choice (valid when mask != 0)
};
enum { HUMAN_NONE = 0xFF };
// The input of human seat p for an action (op, arg), or -1 when the step must wait for it (missing or not valid).
// legal = the seat's legal target set for PICK_PLAYER (already without itself when the action excludes it).
__device__ __forceinline__ int human_choice_of(const HumanIn& H, int p, int op, int arg, uint32_t legal) {
    const uint32_t c = H.choice[p];
    if (op == ACT_PICK_PLAYER) {
        if (legal == 0) return 0;                                  // nobody to pick: same as a bot, never waits
        return (c >= 1 && c <= 32 && ((legal >> (c - 1)) & 1u)) ? (int)c : -1;
    }
    if (op == ACT_PICK_OPTION) return (c >= 1 && c <= (uint32_t)arg) ? (int)c : -1;
    return c != HUMAN_NONE ? 1 : -1;                               // MARK
}

// comparison fields (numeric conditions of the DSL; include/game_engine_b200.h ge_cmp_t)
__device__ __forceinline__ bool cmp_holds(int op, uint32_t v, uint32_t c) {
    return op == 0 ? v == c : op == 1 ? v != c : op == 2 ? v < c : op == 3 ? v <= c : op == 4 ? v > c : v >= c;
}
// lane mask of "byte p of the packed words <op> constant" over the first P players (werewolf: selected_target_id)
template <int NW>
__device__ __forceinline__ uint32_t cmp_mask_bytes(const uint32_t (&w)[NW], int P, const ge_cmp_t& c) {
    uint32_t m = 0;
#pragma unroll
    for (int p = 0; p < 4 * NW; ++p)
        if (p < P && cmp_holds(c.op, (w[p >> 2] >> (8 * (p & 3))) & 0xFFu, c.constant)) m |= 1u << p;
    return m;
}

__device__ __forceinline__ uint32_t all_mask(int P) { return P >= 32 ? 0xFFFFFFFFu : ((1u << P) - 1u); }

// byte offset inside a tile of record byte `o` of session-lane `sl` (S = record bytes)
template <int S>
__device__ __host__ __forceinline__ constexpr uint32_t tile_off(uint32_t o, uint32_t sl) {
    return (o / 16u < (uint32_t)(S / 16)) ? (o / 16u) * 512u + sl * 16u + (o % 16u)
                                          : (uint32_t)(S / 16) * 512u + sl * 8u + (o - 16u * (uint32_t)(S / 16));
}

// ---- bulk asynchronous copies (cp.async.bulk, the non-tensor TMA path) with mbarrier completion -------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!ok);
}

// programmatic dependent launch (sm_90+): see counters_init in ge_step_tps.cuh
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

// L1 prefetch of the 128-byte line that holds p (one warp instruction covers a 512-byte column of a tile)
__device__ __forceinline__ void prefetch_l1(const uint8_t* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint4 ld128(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ uint2 ld64(const uint8_t* p) { return *reinterpret_cast<const uint2*>(p); }
__device__ __forceinline__ void st128(uint8_t* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }
__device__ __forceinline__ void st64(uint8_t* p, const uint2& v) { *reinterpret_cast<uint2*>(p) = v; }

// Warp-uniform visit accumulator.  Live sessions of a batch move in lockstep, so nearly every tile a warp
// processes enters the same phase: the (phase, count) pair is kept in registers and only spilled to the
// block's shared counters when the phase changes (and once at the end).
struct VisitAcc {
    int phase = -1;
    uint32_t count = 0;
    // returns the number of distinct phases entered by the tile's sessions
    __device__ __forceinline__ int add(uint32_t* s_visits, int np, int lane) {
        uint32_t todo = __ballot_sync(0xFFFFFFFFu, np >= 0);
        int distinct = 0;
        while (todo) {                                   // one iteration per distinct phase in the tile (usually 1)
            const int v = __shfl_sync(0xFFFFFFFFu, np, __ffs(todo) - 1);
            const uint32_t same = __ballot_sync(0xFFFFFFFFu, np == v);
            if (v != phase) { flush(s_visits, lane); phase = v; }
            count += __popc(same);
            todo &= ~same;
            ++distinct;
        }
        return distinct;
    }
    __device__ __forceinline__ void flush(uint32_t* s_visits, int lane) {
        if (lane == 0 && count) atomicAdd(&s_visits[phase], count);
        count = 0;
    }
};

// block-level visit counters: warp-aggregated shared atomics, flushed once per block
__device__ __forceinline__ void count_visit(uint32_t* s_visits, int new_phase, int lane) {
    const unsigned m = __match_any_sync(0xFFFFFFFFu, new_phase);
    if (new_phase >= 0 && lane == __ffs(m) - 1) atomicAdd(&s_visits[new_phase], (uint32_t)__popc(m));
}
__device__ __forceinline__ void flush_visits(const uint32_t* s_visits, unsigned long long* stats) {
    // called by all threads after __syncthreads(); warp 0 writes
    if (threadIdx.x < 32) {
        const uint32_t v = s_visits[threadIdx.x];
        if (v) atomicAdd(&stats[ST_VISITS + threadIdx.x], (unsigned long long)v);
        uint32_t tot = v;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) tot += __shfl_xor_sync(0xFFFFFFFFu, tot, d);
        if (threadIdx.x == 0 && tot) atomicAdd(&stats[ST_COUNTED], (unsigned long long)tot);
    }
}

}  // namespace ge
