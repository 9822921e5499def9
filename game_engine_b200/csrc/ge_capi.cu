// ge_capi.cu — C ABI of the batched referee/phase-step simulator (include/game_engine_b200.h).
//
// Host-side runtime in C++: table validation, session-store allocation, launch geometry and the
// glue kernels (init / import / export / statistics).  The step kernels live in ge_step_tps.cuh
// (thread per session) and ge_step_coop.cuh (lane per player) and are instantiated in their own translation
// units (ge_k_generic.cu, ge_k_spec.cu; registry: ge_kernels.h).  No CPU fallback exists: every compute entry
// point needs a CUDA device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "ge_common.cuh"
#include "ge_step_tps.cuh"      // LightBulk, TPS_THREADS (the kernels themselves are instantiated in ge_k_*.cu)
#include "ge_spec_gen.cuh"
#include "ge_kernels.h"
#include "ge_glue.cuh"

using namespace ge;

// ------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(GE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

#define SYNC(b)                                                                                    \
    do {                                                                                           \
        const int rc_ = sync_and_check(b);                                                         \
        if (rc_ != GE_OK) return rc_;                                                              \
    } while (0)

extern "C" const char* ge_last_error(void) { return g_err.c_str(); }
extern "C" const char* ge_version(void) { return "game_engine_b200 0.1.0 (sm_100a)"; }

// ------------------------------------------------------------------------------------ handles
struct ge_table {
    DevTable dev;
    int family, P, bucket;       // bucket: werewolf P8 (8/16/24/32), TTL PB (4/8/16/32)
    size_t rec_canon, rec_dev;   // canonical / device record bytes
    uint16_t io_read[GE_MAX_PHASES], io_write[GE_MAX_PHASES];   // bytes a step that starts in phase i must read / write (ge_table_phase_io)
    uint16_t io_read_pk[GE_MAX_PHASES], io_write_pk[GE_MAX_PHASES];   // the same for the packed store (0 = table not packable)
    uint32_t init_words[40];     // initial device record
    KernelSet ks;                // the run-time-table kernels of this (family, bucket)  (ge_kernels.h)
    SpecKernels spec;            // build-time specialised twins for this exact table (all NULL when it is not a shipped one)
};

struct ge_batch {
    ge_table* tab;
    int device, sm_count;
    uint64_t n, n_tiles, first_sid, seed;
    uint8_t* d_tiles;
    size_t tiles_bytes;
    unsigned long long* d_stats;      // accumulator: counted/visits (step kernels) + harvested histograms
    unsigned long long* d_stats_out;  // snapshot returned by ge_stats / ge_stats_device_ptr
    uint8_t* d_stage;
    size_t stage_bytes;
    cudaStream_t stream;          // stream in use (own_stream unless ge_batch_set_stream bound another)
    cudaStream_t own_stream;
    // active-prefix compaction (see k_compact_*): slot -> original index, per-tile live masks, scratch
    uint32_t* d_origin;
    uint32_t* d_live_mask;
    uint32_t* d_prefix;           // per-tile prefix inside a scan block
    uint32_t* d_blk;              // per-scan-block prefix
    int scan_blocks, dead_shift;
    unsigned long long* d_cstate; // [0] n_active  [1] n_live  [2] pairs to swap  [3] live already in front  [4] old tiles
    int compact_every, since_compact;
    bool host_fused;              // ge_run_host[_async]: apply n_steps in one fused launch
    unsigned long long* h_hint;   // pinned, device-mapped {n_active, epoch}: stored by the compaction / regroup kernels
    unsigned long long* d_hint;   // the same words as the device sees them
    // Learned check schedule: h_sched[c] = host epoch in which compaction check number c of an epoch last RAN,
    // h_sched[64 + c] = in which it last FIRED (stored by k_compact_scan through the mapping).  Games of one table have the
    // same length distribution in every epoch, so a check that did not fire the last time it was observed is not launched
    // again until the periodic refresh epoch (want_check).  checks_seen = due checks since the last (re)initialisation.
    uint32_t* h_sched;
    uint32_t* d_sched;
    uint32_t checks_seen;
    bool sched_valid;             // the batch's sessions started their games together at the last (re)initialisation (not imported)
    unsigned long long epoch;     // bumped by every (re)initialisation; stale hints are ignored
    bool compacted;               // origin may differ from identity
    bool origin_iota;             // !compacted, and d_origin already holds the identity (written by k_reinit): no k_iota needed
    // phase regrouping (see k_regroup_*): counting sort of the active prefix by phase, through a scratch store
    uint32_t* d_rg;               // RG_WORDS control words (histogram, trigger, bases, cursors)
    uint8_t* d_rg_tiles;          // scratch session store (same size as d_tiles)
    uint32_t* d_rg_origin;
    uint32_t* d_tile_present;     // per-tile phase presence (SlotArgs::tile_present), kept while regrouping is on
    bool tile_valid;              // the words describe the records as they are now (false after a reset / import / re-order)
    int regroup_every, rg_mixed_shift;
    uint64_t sid_stride;          // auto-reset: session ids of device epoch e start at first_sid + e * sid_stride (0 = off)
    uint32_t* d_presence;         // 3 rotating phase-presence words (StepArgs::presence)
    uint32_t launch_idx;          // index of the next step launch
    uint32_t next_override;       // presence override for the next launch (0 = read the device word)
    int kernel;
    step_fn fn[4];               // by kernel id (COOP, TPS, TPS_GENERIC)
    ring_fn rfn[4];              // ring-launch twins (thread-per-session kernels only)
    int grid[4], occ[4];         // persistent grid size / occupancy limit (CTAs per SM) per kernel id
    int ctas_per_sm;             // ge_batch_set_grid's request (0 = the occupancy limit)
    uint64_t launches;
    uint32_t* d_hmask;            // human seats per session (NULL = all bots) and their inputs for the next step
    uint8_t* d_hchoice;
    bool hchoice_set;             // inputs are pending: the next step launch consumes them
    uint32_t step_flags;          // StepArgs::flags (ge_batch_set_option)
    int wire;                     // host-buffer record format (GE_WIRE_*); rec_wire = its record size
    size_t rec_wire;
    uint32_t* h_err;              // import validation: [0] rejected records, [1] max(~index): pinned host words that k_import
    uint32_t* d_err;              //   bumps THROUGH THE MAPPING (d_err) only when it finds a bad record — no memset, no copy
    cudaEvent_t fence;            // orders a caller-supplied stream against the batch's own (ge_step / ge_run_fused / ge_stats_refresh)
    // Store format (GE_OPT_STORE_PACKED): werewolf tables up to 8 players can keep their records PACKED — the 32 bytes of
    // the dense wire record in two 16-byte columns instead of 56 bytes in 3.5 — for the all-bot thread-per-session
    // kernels.  want_packed = the option; packed = how d_tiles is laid out right now; rec_store = its record bytes.
    // Everything that does not speak the packed layout (lane-per-player kernels, human seats, regrouping, audience
    // masks) converts the store back first (ensure_store).
    bool want_packed, packed;
    size_t rec_store;
    bool pdl;                     // GE_OPT_PDL: step launches carry the programmatic-stream-serialization attribute
};

static int restore_order(ge_batch* b, bool keep_records);
static size_t human_stride(const ge_table* t);
// Synchronise the batch's stream, then report what the last import found (the verdict travels back asynchronously,
// so the asynchronous host-buffer call can stay asynchronous): a batch that was fed malformed records says so at
// its next synchronising call.  The offending records were replaced by initial records on the device.
static int sync_and_check(ge_batch* b);
static int lanes_per_session(const ge_table* t);
extern "C" int ge_batch_set_regroup(ge_batch* b, int every_n_steps, int min_mixed_shift);
static int ensure_stage(ge_batch* b, size_t bytes);
static int ensure_store(ge_batch* b, bool packed, cudaStream_t st);
extern "C" size_t ge_table_wire_size(const ge_table* t, int wire);

static int sync_and_check(ge_batch* b) {
    CU(cudaStreamSynchronize(b->stream));
    const uint32_t bad = ((volatile uint32_t*)b->h_err)[0];
    if (bad) {
        const uint32_t first = ~((volatile uint32_t*)b->h_err)[1];
        b->h_err[0] = 0; b->h_err[1] = 0;
        return fail(GE_ERR_ARG, "import: " + std::to_string(bad) + " record(s) failed validation (first at index " + std::to_string(first) +
                                "): phase / player ids outside the table, mask bits above the player count, or step 0 outside phase 0; "
                                "they were replaced by initial records");
    }
    return GE_OK;
}

// ------------------------------------------------------------------------------------ table
static int validate_and_build(const uint8_t* blob, size_t n, ge_table* t) {
    if (!blob || n < sizeof(ge_table_header_t)) return fail(GE_ERR_ARG, "table blob too small");
    ge_table_header_t h;
    memcpy(&h, blob, sizeof h);
    if (memcmp(h.magic, "GETB", 4) != 0 || h.version != 1) return fail(GE_ERR_ARG, "bad table magic/version");
    if (h.n_phases < 1 || h.n_phases > GE_MAX_PHASES || h.n_preds > GE_MAX_PREDS) return fail(GE_ERR_ARG, "table counts out of range");
    if (h.n_players < 2 || h.n_players > 32) return fail(GE_ERR_ARG, "n_players must be 2..32");
    const size_t need = sizeof h + (size_t)h.n_phases * sizeof(ge_phase_t) + (size_t)h.n_preds * sizeof(ge_pred_t);
    if (n < need) return fail(GE_ERR_ARG, "table blob truncated");
    memset(&t->dev, 0, sizeof t->dev);
    t->dev.h = h;
    memcpy(t->dev.phase, blob + sizeof h, (size_t)h.n_phases * sizeof(ge_phase_t));
    memcpy(t->dev.pred, blob + sizeof h + (size_t)h.n_phases * sizeof(ge_phase_t), (size_t)h.n_preds * sizeof(ge_pred_t));
    // comparison fields (numeric conditions): werewolf up to 2 on selected_target_id, TTL up to 4 on its value fields
    if (h.n_cmp > (h.family == FAM_WEREWOLF ? 2 : 4) || h.reserved[0] || h.reserved[1]) return fail(GE_ERR_ARG, "bad comparison-field count");
    for (int k = 0; k < 4; ++k) {
        const ge_cmp_t& c = h.cmp[k];
        if (k >= h.n_cmp) { if (c.value_field | c.op | c.constant) return fail(GE_ERR_ARG, "unused comparison field is not zero"); continue; }
        if (c.op > 5 || c.value_field > (h.family == FAM_WEREWOLF ? 0 : 2)) return fail(GE_ERR_ARG, "bad comparison field");
    }
    // predicates may only name mask fields the table defines (SPEC.md section 2): werewolf 0-12, TTL 0-4, the table's
    // comparison fields, and 15; bit 15 of pos0 chains the record to the next one
    uint16_t defined = h.family == FAM_WEREWOLF ? 0x9FFFu : 0x801Fu;
    for (int k = 0; k < h.n_cmp; ++k) defined |= (uint16_t)(1u << ((h.family == FAM_WEREWOLF ? 13 : 11) + k));
    for (int i = 0; i < h.n_preds; ++i) {
        const ge_pred_t& p = t->dev.pred[i];
        if ((p.pos0 | p.neg0 | p.pos1 | p.neg1) & ~defined) return fail(GE_ERR_ARG, "predicate names an undefined mask field");
        if ((p.pos0 & GE_PRED_CONTINUED) && i + 1 >= h.n_preds) return fail(GE_ERR_ARG, "the last predicate record is marked continued");
    }
    for (int i = 0; i < h.n_phases; ++i) {
        const ge_phase_t& ph = t->dev.phase[i];
        if (ph.kind > KIND_TERMINAL || ph.n_branches > 4) return fail(GE_ERR_ARG, "bad phase record");
        if (ph.kind != KIND_TERMINAL && ph.n_branches == 0) return fail(GE_ERR_ARG, "non-terminal phase without next_phase");
        // well-formedness (SPEC.md section 7): effects belong to the table's family and to action phases
        const bool wolfy = h.family == FAM_WEREWOLF;
        if (ph.exit_op != EX_NONE && ph.kind != KIND_ACTION) return fail(GE_ERR_ARG, "exit effect on a phase without actors");
        if (wolfy ? ph.exit_op > EX_DAY_VOTE : (ph.exit_op != EX_NONE && (ph.exit_op < EX_T_STATEMENTS || ph.exit_op > EX_T_VOTES)))
            return fail(GE_ERR_ARG, "exit effect of another rule family");
        if (wolfy ? ph.entry_op > EN_NIGHT_RESET : (ph.entry_op != EN_NONE && (ph.entry_op < EN_T_ROUND_START || ph.entry_op > EN_T_FINAL)))
            return fail(GE_ERR_ARG, "entry effect of another rule family");
        if (ph.kind == KIND_ACTION && (ph.action_op < ACT_PICK_PLAYER || ph.action_op > ACT_MARK)) return fail(GE_ERR_ARG, "action phase without an action op");
        if (wolfy && ph.exit_op != EX_NONE && ph.action_op != ACT_PICK_PLAYER) return fail(GE_ERR_ARG, "werewolf exit effects record a chosen player: the action must be PICK_PLAYER");
        if (ph.kind == KIND_ACTION) {
            if (ph.actor_pred >= h.n_preds) return fail(GE_ERR_ARG, "actor predicate out of range");
            if (ph.action_op == ACT_PICK_PLAYER && ph.action_arg >= h.n_preds) return fail(GE_ERR_ARG, "legal-set predicate out of range");
            if (ph.action_op == ACT_PICK_OPTION && ph.action_arg == 0) return fail(GE_ERR_ARG, "PICK_OPTION needs options > 0");
        }
        for (int b = 0; b < ph.n_branches; ++b) {
            const ge_branch_t& br = ph.br[b];
            if (br.next >= h.n_phases) return fail(GE_ERR_ARG, "branch target out of range");
            if ((br.op == BR_COUNT_EQ0 || br.op == BR_COUNT_GE) && br.a >= h.n_preds) return fail(GE_ERR_ARG, "branch predicate out of range");
            if (br.op == BR_COUNT_GE && br.arg >= h.n_preds) return fail(GE_ERR_ARG, "branch predicate out of range");
            if (br.op == BR_ALL_VAL_GE && br.a > 2) return fail(GE_ERR_ARG, "value field out of range");
            if (br.op > BR_TIE_PENDING) return fail(GE_ERR_ARG, "unknown branch op");
            if (h.family == FAM_WEREWOLF ? (br.op == BR_ALL_VAL_GE || br.tag > 2) : br.op == BR_TIE_PENDING)
                return fail(GE_ERR_ARG, "branch op / winner tag of another rule family");
        }
    }
    // Column groups a step that STARTS in phase i must read besides column 0 (bit0 = dynamic masks C1:
    // fields 2..5; bit1 = role/team C2: fields 6..11; bit2 = per-player bytes).  A group is needed when a
    // predicate reads one of its fields or an effect read-modify-writes it; effects that overwrite a
    // whole group (ASSIGN_ROLES -> C2, NIGHT_RESET -> player bytes) need no read.
    auto pred_need = [&](int pi) -> uint8_t {
        uint8_t out = 0;
        for (; pi >= 0 && pi < h.n_preds; ++pi) {             // the whole run of a continued predicate
            const ge_pred_t& p = t->dev.pred[pi];
            const uint16_t lits[4] = {p.pos0, p.neg0, p.pos1, p.neg1};
            for (int c = 0; c < 2; ++c) {
                if (lits[2 * c + 1] & 0x8000u) continue;      // clause marked empty (& ~ALL)
                const uint16_t used = lits[2 * c] | lits[2 * c + 1];
                if (used & 0x003Cu) out |= 1;
                if (used & 0x1FC0u) out |= 2;
                if (used & 0x6000u) out |= 4;                 // comparison fields read the per-player bytes
            }
            if (!(p.pos0 & GE_PRED_CONTINUED)) break;
        }
        return out;
    };
    auto entry_need = [&](int en) -> uint8_t { return (en == EN_ASSIGN_ROLES || en == EN_NIGHT_RESET) ? 1 : 0; };
    for (int i = 0; i < h.n_phases; ++i) {
        const ge_phase_t& ph = t->dev.phase[i];
        uint8_t need = 0;
        if (ph.kind == KIND_TERMINAL) { t->dev.need[i] = 0; continue; }
        if (h.family != FAM_WEREWOLF) {
            // TTL: bit2 = the player words (everything but column 0's header), bit3 = session id.  Header-only
            // phases: no actors, no predicate / value test in the branches, no entry effect on the way out.
            if (ph.kind == KIND_ACTION) need |= 4 | (ph.action_op != ACT_MARK ? 8 : 0);
            for (int b = 0; b < ph.n_branches; ++b) {
                const ge_branch_t& br = ph.br[b];
                if (br.op == BR_COUNT_EQ0 || br.op == BR_COUNT_GE || br.op == BR_ALL_VAL_GE) need |= 4;
                if (t->dev.phase[br.next].entry_op != EN_NONE) need |= 4;
            }
            t->dev.need[i] = need;
            continue;
        }
        if (ph.kind == KIND_ACTION) {
            need |= 8;                                        // bots draw from the session's Philox stream (needs its id)
            need |= pred_need(ph.actor_pred);
            if (ph.action_op == ACT_PICK_PLAYER) need |= pred_need(ph.action_arg);
            // recording exits read-modify-write the submitted mask (C1).  Only the day vote goes through the player
            // bytes in registers; night actions store their few target bytes directly (ge_step_tps.cuh: PlSink)
            if (ph.exit_op >= EX_VOTE_KILL && ph.exit_op <= EX_DAY_VOTE) need |= 1;
            // (up to 8 players the bytes are one 8-byte column: through registers is as cheap, measured)
            if (ph.exit_op == EX_DAY_VOTE || (h.n_players <= 8 && ph.exit_op >= EX_VOTE_KILL)) need |= 4;
        }
        for (int b = 0; b < ph.n_branches; ++b) {
            const ge_branch_t& br = ph.br[b];
            if (br.op == BR_COUNT_EQ0 || br.op == BR_COUNT_GE) need |= pred_need(br.a);
            if (br.op == BR_COUNT_GE) need |= pred_need((int)br.arg);
            need |= entry_need(t->dev.phase[br.next].entry_op);
            if (t->dev.phase[br.next].entry_op == EN_ASSIGN_ROLES) need |= 8;
        }
        // bit 4: column D1 of the PACKED store (role_lo / role_hi bytes + target bytes; ge_step_tps.cuh) must be read:
        // a predicate names a role field (8-11), anything touches the target bytes, or an entry effect rewrites part of
        // the column (ASSIGN_ROLES the role bytes, NIGHT_RESET the target bytes)
        {
            auto pred_roles = [&](int pi) -> bool {
                for (; pi >= 0 && pi < h.n_preds; ++pi) {
                    const ge_pred_t& p = t->dev.pred[pi];
                    const uint16_t lits[4] = {p.pos0, p.neg0, p.pos1, p.neg1};
                    for (int c = 0; c < 2; ++c)
                        if (!(lits[2 * c + 1] & 0x8000u) && ((lits[2 * c] | lits[2 * c + 1]) & (h.n_players > 8 ? 0x1FF0u : 0x0F00u))) return true;
                    if (!(p.pos0 & GE_PRED_CONTINUED)) break;
                }
                return false;
            };
            // (up to 16 players the packed store has three columns: D1 = role_revealed, investigated, team, secret and role
            // masks — fields 4..12 —, read when a predicate names one or an effect changes part of it (INVESTIGATE_RESOLVE,
            // DAY_VOTE's reveal, ASSIGN_ROLES); the target bytes are D2, governed by bit 2 as in the canonical store)
            const bool wide = h.n_players > 8;
            bool d1 = !wide && (need & 4) != 0;
            if (ph.kind == KIND_ACTION) {
                d1 |= pred_roles(ph.actor_pred);
                if (ph.action_op == ACT_PICK_PLAYER) d1 |= pred_roles(ph.action_arg);
                if (!wide && ph.exit_op >= EX_VOTE_KILL && ph.exit_op <= EX_DAY_VOTE) d1 = true;      // recorded targets
                if (wide && (ph.exit_op == EX_INVESTIGATE_RESOLVE || ph.exit_op == EX_DAY_VOTE)) d1 = true;
            }
            for (int b = 0; b < ph.n_branches; ++b) {
                const ge_branch_t& br = ph.br[b];
                if (br.op == BR_COUNT_EQ0 || br.op == BR_COUNT_GE) d1 |= pred_roles(br.a);
                if (br.op == BR_COUNT_GE) d1 |= pred_roles((int)br.arg);
                const int en = t->dev.phase[br.next].entry_op;
                if (en == EN_ASSIGN_ROLES || (!wide && en == EN_NIGHT_RESET)) d1 = true;
            }
            if (d1) need |= 16;
        }
        t->dev.need[i] = need;
    }
    // NECESSARY bytes per session of a step that starts in phase i (ge_table_phase_io): the columns `need` proves it
    // must read, and the columns its effects can change.  Column 0 (header, is_alive, can_vote) moves both ways on
    // every step.  This is what an ideal implementation of the same column layout moves; the §8(d) "algorithmic"
    // figure 2·S assumes every byte of the record moves on every step.
    for (int i = 0; i < h.n_phases; ++i) {
        const ge_phase_t& ph = t->dev.phase[i];
        t->io_read_pk[i] = t->io_write_pk[i] = 0;
        if (ph.kind == KIND_TERMINAL) { t->io_read[i] = t->io_write[i] = 0; continue; }
        const uint8_t need = t->dev.need[i];
        bool en_assign = false, en_reset = false, en_any = false;
        for (int b = 0; b < ph.n_branches; ++b) {
            const int en = t->dev.phase[ph.br[b].next].entry_op;
            en_assign |= en == EN_ASSIGN_ROLES; en_reset |= en == EN_NIGHT_RESET; en_any |= en != EN_NONE;
        }
        if (h.family == FAM_WEREWOLF) {
            const int P8 = ((h.n_players + 7) / 8) * 8;
            int rd = 16 + ((need & 1) ? 16 : 0) + ((need & 2) ? 16 : 0) + ((need & 4) ? P8 : 0);
            int wr = 16;
            const bool records = ph.kind == KIND_ACTION && ph.exit_op >= EX_VOTE_KILL && ph.exit_op <= EX_DAY_VOTE;
            if (records || en_assign || en_reset) wr += 16;                        // submitted / revealed / eligible live in C1
            if (en_assign) wr += 16;                                               // roles and team: C2
            if (en_reset || (records && (ph.exit_op == EX_DAY_VOTE || h.n_players <= 8))) wr += P8;
            else if (records) wr += ph.exit_op == EX_VOTE_KILL ? h.n_wolves : 1;   // one target byte per actor, stored directly
            t->io_read[i] = (uint16_t)rd; t->io_write[i] = (uint16_t)wr;
            // packed store (up to 8 players): D0 both ways on every step, D1 when `need` bit 4 says so / when the role or
            // target bytes change
            if (P8 == 8) {
                t->io_read_pk[i] = (uint16_t)(16 + ((need & 16) ? 16 : 0));
                t->io_write_pk[i] = (uint16_t)(16 + ((records || en_assign || en_reset) ? 16 : 0));
            } else if (P8 == 16) {
                const bool d1w = en_assign || (ph.kind == KIND_ACTION && (ph.exit_op == EX_INVESTIGATE_RESOLVE || ph.exit_op == EX_DAY_VOTE));
                int wr_pk = 16 + (d1w ? 16 : 0);
                if (en_reset || (records && ph.exit_op == EX_DAY_VOTE)) wr_pk += 16;
                else if (records) wr_pk += ph.exit_op == EX_VOTE_KILL ? h.n_wolves : 1;
                t->io_read_pk[i] = (uint16_t)(16 + ((need & 16) ? 16 : 0) + ((need & 4) ? 16 : 0));
                t->io_write_pk[i] = (uint16_t)wr_pk;
            }
        } else {
            const int bucket = h.n_players <= 4 ? 4 : h.n_players <= 8 ? 8 : h.n_players <= 16 ? 16 : 32;
            const int S = 8 + 4 * bucket;
            const bool touches = (ph.kind == KIND_ACTION && ph.exit_op != EX_NONE) || en_any;
            t->io_read[i] = (uint16_t)((need & 4) ? S : 16);
            t->io_write[i] = (uint16_t)(touches ? S : 16);
        }
    }
    if (t->dev.phase[0].id != 0) return fail(GE_ERR_ARG, "phase index 0 must be DSL phase 0");
    t->dev.nonterm = 0;
    for (int i = 0; i < h.n_phases; ++i)
        if (t->dev.phase[i].kind != KIND_TERMINAL) t->dev.nonterm |= 1u << i;
    t->family = h.family;
    t->P = h.n_players;
    memset(t->init_words, 0, sizeof t->init_words);
    const uint32_t ALL = h.n_players >= 32 ? 0xFFFFFFFFu : ((1u << h.n_players) - 1u);
    if (h.family == FAM_WEREWOLF) {
        if (h.n_wolves < 1 || h.n_wolves + 2 > h.n_players) return fail(GE_ERR_ARG, "bad wolf count");
        t->bucket = ((h.n_players + 7) / 8) * 8;
        t->rec_canon = t->rec_dev = 48 + (size_t)t->bucket;
        // word index = byte offset / 4 (SPEC.md section 5); mask field f lives at 8 + 4*slot
        static const int slot_of_field[8] = {0, 1, 2, 3, 4, 5, 6, 7};
        for (int f = 0; f < 8; ++f)
            if ((h.init_masks >> f) & 1u) t->init_words[2 + slot_of_field[f]] = ALL;
    } else if (h.family == FAM_TTL) {
        t->bucket = h.n_players <= 4 ? 4 : h.n_players <= 8 ? 8 : h.n_players <= 16 ? 16 : 32;
        t->rec_canon = (size_t)((8 + 4 * h.n_players + 7) / 8) * 8;
        t->rec_dev = 8 + 4 * (size_t)t->bucket;
        uint32_t fl = 0;
        for (int f = 0; f < 5; ++f)
            if ((h.init_masks >> f) & 1u) fl |= 1u << f;
        for (int p = 0; p < h.n_players; ++p) t->init_words[2 + p] = fl << 24;
    } else {
        return fail(GE_ERR_UNSUPPORTED, "unknown rule family");
    }
    return GE_OK;
}

// build-time specialised kernels (ge_k_spec.cu, one translation unit per shipped table), matched by byte-identical blobs
#define GE_DECL_SPEC(S, FAM, BUCKET) void ge_spec_kernels_##S(SpecKernels* e);
GE_SPEC_LIST(GE_DECL_SPEC)
#undef GE_DECL_SPEC
typedef void (*spec_getter)(SpecKernels*);
#define GE_SPEC_GETTER(S, FAM, BUCKET) ge_spec_kernels_##S,
static const spec_getter g_spec_getters[] = { GE_SPEC_LIST(GE_SPEC_GETTER) nullptr };
#undef GE_SPEC_GETTER

static bool kernel_set_of(int family, int bucket, KernelSet* out) {
#define GE_PICK_KSET(F, B) if (family == F && bucket == B) { ge_kernel_set_##F##_##B(out); return true; }
    GE_KERNEL_SETS(GE_PICK_KSET)
#undef GE_PICK_KSET
    return false;
}

extern "C" int ge_table_create(const uint8_t* blob, size_t n, ge_table** out) {
    if (!out) return fail(GE_ERR_ARG, "out is NULL");
    ge_table* t = new (std::nothrow) ge_table;
    if (!t) return fail(GE_ERR_NOMEM, "out of host memory");
    const int rc = validate_and_build(blob, n, t);
    if (rc != GE_OK) { delete t; return rc; }
    t->spec = SpecKernels{};
    if (!kernel_set_of(t->family, t->bucket, &t->ks)) { delete t; return fail(GE_ERR_UNSUPPORTED, "no kernels for this table's player bucket"); }
    const size_t used = sizeof(ge_table_header_t) + (size_t)t->dev.h.n_phases * sizeof(ge_phase_t) + (size_t)t->dev.h.n_preds * sizeof(ge_pred_t);
    for (const spec_getter* g = g_spec_getters; *g; ++g) {
        SpecKernels e;
        (*g)(&e);
        if (e.len == used && memcmp(e.blob, blob, used) == 0) t->spec = e;
    }
    *out = t;
    return GE_OK;
}
extern "C" void ge_table_destroy(ge_table* t) { delete t; }
extern "C" size_t ge_table_record_size(const ge_table* t) { return t ? t->rec_canon : 0; }
extern "C" int ge_table_n_players(const ge_table* t) { return t ? t->P : 0; }
extern "C" int ge_table_phase_io(const ge_table* t, int phase_index, uint32_t* read_bytes, uint32_t* write_bytes) {
    if (!t || phase_index < 0 || phase_index >= t->dev.h.n_phases) return fail(GE_ERR_ARG, "bad arguments to ge_table_phase_io");
    if (read_bytes) *read_bytes = t->io_read[phase_index];
    if (write_bytes) *write_bytes = t->io_write[phase_index];
    return GE_OK;
}
extern "C" int ge_table_phase_io_packed(const ge_table* t, int phase_index, uint32_t* read_bytes, uint32_t* write_bytes) {
    if (!t || phase_index < 0 || phase_index >= t->dev.h.n_phases) return fail(GE_ERR_ARG, "bad arguments to ge_table_phase_io_packed");
    if (!(t->family == FAM_WEREWOLF && t->bucket <= 16)) return fail(GE_ERR_UNSUPPORTED, "the packed store covers werewolf-family tables up to 16 players");
    if (read_bytes) *read_bytes = t->io_read_pk[phase_index];
    if (write_bytes) *write_bytes = t->io_write_pk[phase_index];
    return GE_OK;
}
// dense wire records exist for werewolf tables up to 16 players (SPEC.md section 5b); everything else travels canonical
static bool has_dense(const ge_table* t) { return t->family == FAM_WEREWOLF && t->bucket <= 16; }
extern "C" size_t ge_table_wire_size(const ge_table* t, int wire) {
    if (!t || (wire != GE_WIRE_CANONICAL && wire != GE_WIRE_DENSE)) return 0;
    return (wire == GE_WIRE_DENSE && has_dense(t)) ? (t->bucket == 8 ? 32 : 48) : t->rec_canon;
}

// ------------------------------------------------------------------------------------ dispatch
static step_fn pick_fn(const ge_table* t, int kernel) { return kernel == GE_KERNEL_COOP ? t->ks.coop : t->ks.tps; }
static step_fn pick_tiled_fn(const ge_table* t, int kernel) {
    if (t->family != FAM_WEREWOLF) return nullptr;
    return (kernel == GE_KERNEL_TPS && t->spec.tiled) ? t->spec.tiled : t->ks.tiled;
}
static step_fn pick_human_fn(const ge_table* t) { return t->ks.human; }
static ring_fn pick_ring_fn(const ge_table* t) { return t->ks.ring; }

static int lanes_per_session(const ge_table* t) {
    if (t->family == FAM_WEREWOLF) return t->bucket <= 8 ? 8 : t->bucket <= 16 ? 16 : 32;
    return t->bucket;
}

static int glue_grid(const ge_batch* b, uint64_t items, int block) {
    uint64_t g = (items + block - 1) / block;
    const uint64_t cap = (uint64_t)b->sm_count * 16;
    if (g > cap) g = cap;
    return g < 1 ? 1 : (int)g;
}

// ------------------------------------------------------------------------------------ batch
// the initial record in the batch's CURRENT store format (canonical words, or the packed words of SPEC 5b)
static InitRec init_rec(const ge_batch* b) {
    InitRec rec;
    memset(&rec, 0, sizeof rec);
    const uint32_t* w = b->tab->init_words;
    if (!b->packed) { memcpy(rec.w, w, sizeof rec.w); return rec; }
    rec.w[0] = w[0]; rec.w[1] = w[1];
    if (b->tab->bucket == 8) {
        rec.w[2] = (w[2] & 0xFFu) | ((w[3] & 0xFFu) << 8) | ((w[4] & 0xFFu) << 16) | ((w[5] & 0xFFu) << 24);
        rec.w[3] = (w[6] & 0xFFu) | ((w[7] & 0xFFu) << 8) | ((w[8] & 0xFFu) << 16) | ((w[9] & 0xFFu) << 24);
        rec.w[4] = (w[10] & 0xFFu) | ((w[11] & 0xFFu) << 8);
        rec.w[5] = w[12]; rec.w[6] = w[13]; rec.w[7] = 0;
    } else {
        for (int k = 0; k < 5; ++k) rec.w[2 + k] = (w[2 + 2 * k] & 0xFFFFu) | ((w[3 + 2 * k] & 0xFFFFu) << 16);
        rec.w[7] = 0;
        for (int k = 0; k < 4; ++k) rec.w[8 + k] = w[12 + k];
    }
    return rec;
}
static InitRec init_rec_canon(const ge_batch* b) {
    InitRec rec;
    memcpy(rec.w, b->tab->init_words, sizeof rec.w);
    return rec;
}
static bool packable(const ge_table* t) { return t->family == FAM_WEREWOLF && t->bucket <= 16 && t->ks.tps_pk != nullptr; }
static size_t packed_record(const ge_table* t) { return t->bucket == 8 ? 32 : 48; }
// where is_alive sits in word 2 of the stored record (k_stats)
static uint32_t alive_mask_of(const ge_batch* b) { return !b->packed ? 0xFFFFFFFFu : b->tab->bucket == 8 ? 0xFFu : 0xFFFFu; }
// the layout the batch's next step launch wants
static bool store_should_be_packed(const ge_batch* b) {
    return b->want_packed && packable(b->tab) && b->kernel != GE_KERNEL_COOP && !b->d_hmask && b->regroup_every == 0;
}

static int init_sessions(ge_batch* b, uint64_t first_session_id, uint64_t seed) {
    b->first_sid = first_session_id;
    b->seed = seed;
    const InitRec rec = init_rec(b);
    k_init<<<glue_grid(b, b->n_tiles * 32, 256), 256, 0, b->stream>>>(b->d_tiles, b->n_tiles, (uint32_t)b->rec_store, rec);
    CU(cudaGetLastError());
    b->launches++;
    CU(cudaMemsetAsync(b->d_presence, 0, 3 * sizeof(uint32_t), b->stream));
    b->next_override = 1u;        // every session is in phase index 0
    b->epoch++;
    k_cstate_reset<<<1, 32, 0, b->stream>>>(b->d_cstate, b->n, b->epoch, 0);
    CU(cudaGetLastError());
    b->compacted = false;
    b->origin_iota = false;
    b->since_compact = 0;
    b->checks_seen = 0;
    b->sched_valid = true;
    b->tile_valid = false;
    return GE_OK;
}

// Statistics are cumulative over the life of the handle: before the sessions are overwritten their
// final-state histograms are folded into the accumulator ("harvest"), so win rates cover every session
// the batch ever simulated.
extern "C" int ge_batch_reset(ge_batch* b, uint64_t first_session_id, uint64_t seed) {
    if (!b) return fail(GE_ERR_ARG, "batch is NULL");
    CU(cudaSetDevice(b->device));
    // harvest + initial records + identity slot order + bookkeeping in one launch (k_reinit)
    b->first_sid = first_session_id;
    b->seed = seed;
    b->epoch++;
    k_reinit<<<glue_grid(b, b->n_tiles * 32, 256), 256, 0, b->stream>>>(b->tab->dev, b->d_tiles, (uint32_t)b->rec_store, b->n, b->n_tiles, init_rec(b),
                                                                         b->d_origin, b->d_stats, b->d_cstate, b->d_presence, b->epoch, alive_mask_of(b));
    CU(cudaGetLastError());
    b->launches++;
    b->next_override = 1u;        // every session is in phase index 0
    b->compacted = false;
    b->origin_iota = true;
    b->since_compact = 0;
    b->checks_seen = 0;
    b->sched_valid = true;
    b->tile_valid = false;
    return GE_OK;
}

extern "C" int ge_batch_clear_stats(ge_batch* b) {
    if (!b) return fail(GE_ERR_ARG, "batch is NULL");
    CU(cudaSetDevice(b->device));
    CU(cudaMemsetAsync(b->d_stats, 0, GE_STATS_LEN * sizeof(unsigned long long), b->stream));
    return GE_OK;
}

extern "C" int ge_batch_create(ge_table* t, int device, uint64_t n_sessions, uint64_t first_session_id, uint64_t seed,
                               ge_batch** out) {
    if (!t || !out || n_sessions == 0) return fail(GE_ERR_ARG, "bad arguments to ge_batch_create");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(GE_ERR_ARG, "no such CUDA device");
    CU(cudaSetDevice(device));
    ge_batch* b = new (std::nothrow) ge_batch;
    if (!b) return fail(GE_ERR_NOMEM, "out of host memory");
    memset(b, 0, sizeof *b);
    b->tab = t; b->device = device; b->n = n_sessions; b->n_tiles = (n_sessions + 31) / 32;
    b->tiles_bytes = (size_t)b->n_tiles * 32 * t->rec_dev;      // allocated for the canonical columns; the packed store uses a prefix
    b->rec_store = t->rec_dev;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete b; return fail(GE_ERR_CUDA, cudaGetErrorString(e)); }
    b->sm_count = prop.multiProcessorCount;
    e = cudaMalloc(&b->d_tiles, b->tiles_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&b->d_stats, GE_STATS_LEN * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_stats_out, GE_STATS_LEN * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_presence, 3 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_origin, b->n_tiles * 32 * sizeof(uint32_t));
    // padded so that the scan kernel's 1024 x (multiple of 4) ranges never leave the allocation
    const size_t mask_words = ((b->n_tiles + 4095) / 4096 + 1) * 4096;
    if (e == cudaSuccess) e = cudaMalloc(&b->d_live_mask, mask_words * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(b->d_live_mask, 0, mask_words * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_prefix, mask_words * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_cstate, 16 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaHostAlloc(&b->h_hint, 2 * sizeof(unsigned long long) + 128 * sizeof(uint32_t), cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&b->d_hint, b->h_hint, 0);
    if (e == cudaSuccess) {
        b->h_sched = reinterpret_cast<uint32_t*>(b->h_hint + 2);
        b->d_sched = reinterpret_cast<uint32_t*>(b->d_hint + 2);
        memset(b->h_sched, 0, 128 * sizeof(uint32_t));
    }
    if (e == cudaSuccess) e = cudaHostAlloc(&b->h_err, 2 * sizeof(uint32_t), cudaHostAllocMapped);
    if (e == cudaSuccess) { b->h_err[0] = 0; b->h_err[1] = 0; }
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&b->d_err, b->h_err, 0);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->fence, cudaEventDisableTiming);
    if (e == cudaSuccess) { b->h_hint[0] = n_sessions; b->h_hint[1] = 0; }
    b->scan_blocks = (int)((b->n_tiles + CS_TILES - 1) / CS_TILES);
    if (e == cudaSuccess) e = cudaMalloc(&b->d_blk, ((size_t)b->scan_blocks + 2) * sizeof(uint32_t));
    b->dead_shift = 2;                                   // compact when >= 1/4 of the active prefix is dead
    // Default cadence of the compaction check: every 5 steps for the werewolf family (measured best at 8 players:
    // +7 % over every 8; neutral at 16 / 32), off for the TTL family, whose games all have the same length (nothing
    // to compact until every game ends at once: the check only costs launches, −4 %).  One finishing block scans
    // the block totals, hence the size limit.
    b->compact_every = (b->scan_blocks <= 1024 && t->family == FAM_WEREWOLF) ? 5 : 0;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&b->own_stream, cudaStreamNonBlocking);
    b->stream = b->own_stream;
    if (e != cudaSuccess) {
        cudaFree(b->d_tiles); cudaFree(b->d_stats); cudaFree(b->d_stats_out); cudaFree(b->d_presence);
        cudaFree(b->d_origin); cudaFree(b->d_live_mask); cudaFree(b->d_prefix); cudaFree(b->d_blk); cudaFree(b->d_cstate); cudaFreeHost(b->h_hint);
        cudaFreeHost(b->h_err); if (b->fence) cudaEventDestroy(b->fence);
        delete b;
        return fail(e == cudaErrorMemoryAllocation ? GE_ERR_NOMEM : GE_ERR_CUDA, std::string("ge_batch_create: ") + cudaGetErrorString(e));
    }
    for (int k = GE_KERNEL_COOP; k <= GE_KERNEL_TPS_GENERIC; ++k) {
        b->fn[k] = k == GE_KERNEL_TPS_GENERIC ? pick_fn(t, GE_KERNEL_TPS) : (k == GE_KERNEL_TPS && t->spec.tps) ? t->spec.tps : pick_fn(t, k);
        b->rfn[k] = k == GE_KERNEL_COOP ? nullptr : (k == GE_KERNEL_TPS && t->spec.ring) ? t->spec.ring : pick_ring_fn(t);
        if (!b->fn[k]) { ge_batch_destroy(b); return fail(GE_ERR_UNSUPPORTED, "no kernel for this table"); }
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)b->fn[k], 128, 0);
        if (e != cudaSuccess || per_sm < 1) per_sm = 4;
        // (measured on the canonical-store kernel; its packed-store twin is launched with the same grids — the specialised
        // 8-player one fits 10 CTAs per SM, which only matters when SMALL grids of several streams share the machine)
        b->occ[k] = per_sm;
        const uint64_t warps = k == GE_KERNEL_COOP ? b->n_tiles * (uint64_t)lanes_per_session(t) : b->n_tiles;
        uint64_t g = (warps + 3) / 4;
        const uint64_t cap = (uint64_t)b->sm_count * per_sm;      // persistent grid: whole multiples of the SM count
        if (g > cap) g = cap;
        b->grid[k] = g < 1 ? 1 : (int)g;
    }
    b->kernel = GE_KERNEL_TPS;
    b->wire = GE_WIRE_CANONICAL; b->rec_wire = t->rec_canon;
    // tables with a tie -> re-vote loop de-synchronise their sessions: regroup them by phase (k_regroup_*)
    bool desync = false;
    for (int i = 0; i < t->dev.h.n_phases; ++i)
        for (int k = 0; k < t->dev.phase[i].n_branches; ++k)
            if (t->dev.phase[i].br[k].op == BR_TIE_PENDING) desync = true;
    // tables up to 8 players keep their records packed (32 bytes in two columns; +10 % on the headline workload, DESIGN
    // section 6) unless something the packed layout does not serve is switched on (GE_OPT_STORE_PACKED 0 = canonical)
    b->pdl = true;                // programmatic dependent launches: never slower, +16 % on a single stream
    b->want_packed = packable(t);
    b->packed = b->want_packed && !(desync && t->family == FAM_WEREWOLF && b->n <= (1ull << 31));
    b->rec_store = b->packed ? packed_record(t) : t->rec_dev;
    int rc = ge_batch_clear_stats(b);
    if (rc == GE_OK) rc = init_sessions(b, first_session_id, seed);
    if (rc == GE_OK && desync && t->family == FAM_WEREWOLF && b->n <= (1ull << 31)) rc = ge_batch_set_regroup(b, 5, 3);
    if (rc != GE_OK) { ge_batch_destroy(b); return rc; }
    *out = b;
    return GE_OK;
}

extern "C" void ge_batch_destroy(ge_batch* b) {
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    if (b->own_stream) { cudaStreamSynchronize(b->own_stream); cudaStreamDestroy(b->own_stream); }
    cudaFree(b->d_tiles); cudaFree(b->d_stats); cudaFree(b->d_stats_out); cudaFree(b->d_stage); cudaFree(b->d_presence);
    cudaFree(b->d_origin); cudaFree(b->d_live_mask); cudaFree(b->d_prefix); cudaFree(b->d_blk); cudaFree(b->d_cstate); cudaFreeHost(b->h_hint);
    cudaFree(b->d_rg); cudaFree(b->d_rg_tiles); cudaFree(b->d_rg_origin); cudaFree(b->d_tile_present);
    cudaFreeHost(b->h_err); if (b->fence) cudaEventDestroy(b->fence);
    cudaFree(b->d_hmask); cudaFree(b->d_hchoice);
    delete b;
}

extern "C" int ge_batch_set_stream(ge_batch* b, void* cuda_stream) {
    if (!b) return fail(GE_ERR_ARG, "batch is NULL");
    CU(cudaSetDevice(b->device));
    CU(cudaStreamSynchronize(b->stream));
    b->stream = cuda_stream ? (cudaStream_t)cuda_stream : b->own_stream;
    return GE_OK;
}

extern "C" int ge_batch_set_wire(ge_batch* b, int wire) {
    if (!b || (wire != GE_WIRE_CANONICAL && wire != GE_WIRE_DENSE)) return fail(GE_ERR_ARG, "bad wire format");
    b->wire = wire;
    b->rec_wire = ge_table_wire_size(b->tab, wire);
    return GE_OK;
}
extern "C" size_t ge_batch_wire_size(const ge_batch* b) { return b ? b->rec_wire : 0; }

// Human seats (SPEC.md section 1, D3h).  host_masks[i] = seats of session i played by people (bit p-1 = player p);
// NULL returns the batch to all bots.  Synchronous copy.
extern "C" int ge_batch_set_human_seats(ge_batch* b, const uint32_t* host_masks) {
    if (!b) return fail(GE_ERR_ARG, "batch is NULL");
    CU(cudaSetDevice(b->device));
    CU(cudaStreamSynchronize(b->stream));
    if (!host_masks) {
        cudaFree(b->d_hmask); cudaFree(b->d_hchoice);
        b->d_hmask = nullptr; b->d_hchoice = nullptr; b->hchoice_set = false;
        return GE_OK;
    }
    if (b->kernel == GE_KERNEL_COOP) return fail(GE_ERR_UNSUPPORTED, "human seats are served by the thread-per-session kernels");
    const uint32_t hi = b->tab->P >= 32 ? 0u : ~((1u << b->tab->P) - 1u);
    for (uint64_t i = 0; i < b->n; ++i)
        if (host_masks[i] & hi) return fail(GE_ERR_ARG, "human seat above the player count");
    if (!b->d_hmask) {
        cudaError_t e = cudaMalloc(&b->d_hmask, b->n * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&b->d_hchoice, b->n * human_stride(b->tab));
        if (e == cudaSuccess) e = cudaMemset(b->d_hchoice, HUMAN_NONE, b->n * human_stride(b->tab));
        if (e != cudaSuccess) {
            cudaFree(b->d_hmask); cudaFree(b->d_hchoice);
            b->d_hmask = nullptr; b->d_hchoice = nullptr;
            return fail(e == cudaErrorMemoryAllocation ? GE_ERR_NOMEM : GE_ERR_CUDA, std::string("ge_batch_set_human_seats: ") + cudaGetErrorString(e));
        }
    }
    CU(cudaMemcpy(b->d_hmask, host_masks, b->n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return GE_OK;
}

// Inputs of the human seats for the NEXT step launch: host_choices[i * stride + p] = what seat p+1 of session i chose
// (a player id for PICK_PLAYER, 1..n for PICK_OPTION, anything for MARK; 0xFF = has not acted), stride =
// ge_table_human_stride.  The next step launch consumes them; a session whose acting human seats are not all
// answered stays in its phase (its history still grows by one entry).  Asynchronous on the batch's stream.
extern "C" int ge_batch_set_human_choices(ge_batch* b, const uint8_t* host_choices) {
    if (!b || !host_choices) return fail(GE_ERR_ARG, "bad arguments to ge_batch_set_human_choices");
    if (!b->d_hmask) return fail(GE_ERR_ARG, "ge_batch_set_human_choices: the batch has no human seats (ge_batch_set_human_seats)");
    CU(cudaSetDevice(b->device));
    CU(cudaMemcpyAsync(b->d_hchoice, host_choices, b->n * human_stride(b->tab), cudaMemcpyHostToDevice, b->stream));
    b->hchoice_set = true;
    return GE_OK;
}
extern "C" size_t ge_table_human_stride(const ge_table* t) { return t ? human_stride(t) : 0; }

extern "C" int ge_batch_set_option(ge_batch* b, int option, int value) {
    if (!b) return fail(GE_ERR_ARG, "batch is NULL");
    if (option == GE_OPT_LIGHT_BULK) {
        b->step_flags = value ? (b->step_flags | STEP_LIGHT_BULK) : (b->step_flags & ~(uint32_t)STEP_LIGHT_BULK);
        return GE_OK;
    }
    if (option == GE_OPT_PDL) { b->pdl = value != 0; return GE_OK; }
    if (option == GE_OPT_STORE_PACKED) {
        if (value && !packable(b->tab)) return fail(GE_ERR_UNSUPPORTED, "the packed store covers werewolf-family tables up to 16 players");
        b->want_packed = value != 0;
        CU(cudaSetDevice(b->device));
        return ensure_store(b, store_should_be_packed(b), b->stream);
    }
    return fail(GE_ERR_ARG, "unknown option");
}

extern "C" int ge_batch_set_kernel(ge_batch* b, int kernel) {
    if (!b || kernel < GE_KERNEL_AUTO || kernel > GE_KERNEL_TPS_GENERIC) return fail(GE_ERR_ARG, "bad kernel id");
    kernel = kernel == GE_KERNEL_AUTO ? GE_KERNEL_TPS : kernel;
    if (kernel == GE_KERNEL_COOP && b->d_hmask) return fail(GE_ERR_UNSUPPORTED, "human seats are served by the thread-per-session kernels");
    if (kernel == GE_KERNEL_COOP && b->sid_stride != 0) return fail(GE_ERR_UNSUPPORTED, "auto-reset rides on the thread-per-session kernels");
    if (kernel == GE_KERNEL_COOP && b->kernel != GE_KERNEL_COOP) {
        // the lane-per-player kernels walk every slot in session order: undo any compaction first
        CU(cudaSetDevice(b->device));
        const int rc = restore_order(b, true);
        if (rc) return rc;
    }
    b->kernel = kernel;
    return GE_OK;
}

extern "C" int ge_batch_set_compaction(ge_batch* b, int every_n_steps, int min_dead_shift) {
    if (!b || every_n_steps < 0 || min_dead_shift < 0 || min_dead_shift > 16) return fail(GE_ERR_ARG, "bad arguments to ge_batch_set_compaction");
    if (every_n_steps > 0 && b->scan_blocks > 1024) return fail(GE_ERR_UNSUPPORTED, "compaction supports batches up to 2^25 sessions");
    b->compact_every = every_n_steps;
    b->dead_shift = min_dead_shift;
    // the learned schedule belongs to a cadence and a threshold: start over (a check still in flight may log a stale entry,
    // which the periodic refresh corrects)
    memset(b->h_sched, 0, 128 * sizeof(uint32_t));
    b->checks_seen = 0;
    b->since_compact = 0;
    return GE_OK;
}

// Persistent grid of the step launches = SM count x ctas_per_sm (clamped to the kernel's occupancy limit and to
// the work available); 0 restores the occupancy limit.
extern "C" int ge_batch_set_grid(ge_batch* b, int ctas_per_sm) {
    if (!b || ctas_per_sm < 0) return fail(GE_ERR_ARG, "bad arguments to ge_batch_set_grid");
    b->ctas_per_sm = ctas_per_sm;
    for (int k = GE_KERNEL_COOP; k <= GE_KERNEL_TPS_GENERIC; ++k) {
        int per_sm = b->occ[k];
        if (ctas_per_sm > 0 && ctas_per_sm < per_sm) per_sm = ctas_per_sm;
        const uint64_t warps = k == GE_KERNEL_COOP ? b->n_tiles * (uint64_t)lanes_per_session(b->tab) : b->n_tiles;
        uint64_t g = (warps + 3) / 4;
        const uint64_t cap = (uint64_t)b->sm_count * per_sm;
        if (g > cap) g = cap;
        b->grid[k] = g < 1 ? 1 : (int)g;
    }
    return GE_OK;
}

extern "C" int ge_batch_set_regroup(ge_batch* b, int every_n_steps, int min_mixed_shift) {
    if (!b || every_n_steps < 0 || min_mixed_shift < 0 || min_mixed_shift > 16) return fail(GE_ERR_ARG, "bad arguments to ge_batch_set_regroup");
    if (every_n_steps == 0) { b->regroup_every = 0; return GE_OK; }
    if (b->tab->family != FAM_WEREWOLF) return fail(GE_ERR_UNSUPPORTED, "phase regrouping covers the werewolf family (TTL sessions never de-synchronise)");
    if (b->n > (1ull << 31)) return fail(GE_ERR_UNSUPPORTED, "phase regrouping supports batches up to 2^31 sessions");
    CU(cudaSetDevice(b->device));
    if (!b->d_rg) {
        cudaError_t e = cudaMalloc(&b->d_rg, RG_WORDS * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMemset(b->d_rg, 0, RG_WORDS * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&b->d_rg_tiles, b->tiles_bytes);
        if (e == cudaSuccess) e = cudaMalloc(&b->d_rg_origin, b->n_tiles * 32 * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&b->d_tile_present, (b->n_tiles + 1) * sizeof(uint32_t));
        b->tile_valid = false;
        if (e != cudaSuccess) {
            cudaFree(b->d_rg); cudaFree(b->d_rg_tiles); cudaFree(b->d_rg_origin); cudaFree(b->d_tile_present);
            b->d_rg = nullptr; b->d_rg_tiles = nullptr; b->d_rg_origin = nullptr; b->d_tile_present = nullptr;
            return fail(e == cudaErrorMemoryAllocation ? GE_ERR_NOMEM : GE_ERR_CUDA, std::string("ge_batch_set_regroup: ") + cudaGetErrorString(e));
        }
    }
    b->regroup_every = every_n_steps;
    b->rg_mixed_shift = min_mixed_shift;
    b->since_compact = 0;
    return GE_OK;
}

extern "C" int ge_batch_set_autoreset(ge_batch* b, uint64_t sid_stride) {
    if (!b) return fail(GE_ERR_ARG, "batch is NULL");
    if (sid_stride != 0 && sid_stride < b->n) return fail(GE_ERR_ARG, "sid_stride must be 0 (off) or >= n_sessions (ids of different epochs must not overlap)");
    if (sid_stride != 0 && b->kernel == GE_KERNEL_COOP) return fail(GE_ERR_UNSUPPORTED, "auto-reset rides on the thread-per-session kernels (the lane-per-player kernels do not follow device epochs)");
    b->sid_stride = sid_stride;
    if (sid_stride != 0 && b->compact_every == 0 && b->regroup_every == 0 && b->scan_blocks <= 1024)
        b->compact_every = 8;                 // the check that notices "every game is over" rides on compaction
    return GE_OK;
}

extern "C" int ge_batch_epochs(ge_batch* b, uint64_t* out) {
    if (!b || !out) return fail(GE_ERR_ARG, "bad arguments to ge_batch_epochs");
    CU(cudaSetDevice(b->device));
    CU(cudaMemcpyAsync(out, b->d_cstate + 8, sizeof(uint64_t), cudaMemcpyDeviceToHost, b->stream));
    SYNC(b);
    return GE_OK;
}

extern "C" int ge_batch_set_host_fused(ge_batch* b, int on) {
    if (!b) return fail(GE_ERR_ARG, "batch is NULL");
    b->host_fused = on != 0;
    return GE_OK;
}

extern "C" int ge_batch_active_hint(ge_batch* b, uint64_t* out) {
    if (!b || !out) return fail(GE_ERR_ARG, "bad arguments to ge_batch_active_hint");
    // the pair is written by two ordered copies; a hint from an older epoch (before the last reset) is ignored
    const unsigned long long ep = ((volatile unsigned long long*)b->h_hint)[1];
    const unsigned long long v = ((volatile unsigned long long*)b->h_hint)[0];
    *out = (ep == b->epoch && ((volatile unsigned long long*)b->h_hint)[1] == ep) ? v : b->n;
    return GE_OK;
}

extern "C" int ge_batch_active(ge_batch* b, uint64_t* out) {
    if (!b || !out) return fail(GE_ERR_ARG, "bad arguments to ge_batch_active");
    CU(cudaSetDevice(b->device));
    CU(cudaMemcpyAsync(out, b->d_cstate, sizeof(uint64_t), cudaMemcpyDeviceToHost, b->stream));
    SYNC(b);
    return GE_OK;
}
extern "C" int ge_batch_get_kernel(const ge_batch* b) { return b ? b->kernel : GE_ERR_ARG; }

// scan -> rank -> swap -> commit on the live masks of the step that just ran, for a list of batches of one table in
// ONE launch pair (all asynchronous, no host sync, no copies: the kernels store the host's progress hint themselves)
// (check_idx >= 0: the check's number inside the batch's epoch, logged for the learned schedule; -1 = not logged)
static int enqueue_compaction(ge_batch** list, int n, cudaStream_t st, int check_idx = -1) {
    CompactArgs ca;
    memset(&ca, 0, sizeof ca);
    ca.n = n;
    ca.S = (uint32_t)list[0]->rec_store;
    ca.dead_shift = (uint32_t)list[0]->dead_shift;
    int max_scan = 1;
    uint64_t max_tiles = 1;
    for (int i = 0; i < n; ++i) {
        ge_batch* b = list[i];
        if (!b->compacted) {
            if (!b->origin_iota) {
                k_iota<<<glue_grid(b, b->n_tiles * 32, 256), 256, 0, st>>>(b->d_origin, b->n_tiles * 32);
                b->launches++;
            }
            b->compacted = true;
            b->origin_iota = false;
        }
        ca.s[i] = CompactSlot{b->d_tiles, b->d_origin, b->d_live_mask, b->d_prefix, b->d_blk, b->d_cstate, b->d_hint,
                              check_idx >= 0 ? b->d_sched : nullptr, (uint32_t)(check_idx < 0 ? 0 : check_idx)};
        if (b->scan_blocks > max_scan) max_scan = b->scan_blocks;
        if (b->n_tiles > max_tiles) max_tiles = b->n_tiles;
        b->since_compact = 0;
    }
    list[0]->launches += 2;                           // one launch pair for the whole list
    uint64_t sg = (max_tiles * 16 + 127) / 128;
    const uint64_t cap = (uint64_t)list[0]->sm_count * 8 / (n > 4 ? 4 : n);       // the list shares the machine
    if (sg > cap) sg = cap;
    if (list[0]->pdl) {                               // the check rides the same dependent-launch chain as the steps around it
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.stream = st; cfg.attrs = at; cfg.numAttrs = 1;
        cfg.gridDim = dim3((unsigned)max_scan, (unsigned)n); cfg.blockDim = dim3(1024);
        CU(cudaLaunchKernelEx(&cfg, k_compact_scan, ca));
        cfg.gridDim = dim3((unsigned)(sg < 1 ? 1 : sg), (unsigned)n); cfg.blockDim = dim3(128);
        CU(cudaLaunchKernelEx(&cfg, k_compact_swap, ca));
    } else {
        k_compact_scan<<<dim3(max_scan, n), 1024, 0, st>>>(ca);
        k_compact_swap<<<dim3((unsigned)(sg < 1 ? 1 : sg), n), 128, 0, st>>>(ca);
    }
    CU(cudaGetLastError());
    return GE_OK;
}

// Should compaction check number c of the batch's current epoch be launched?  Yes while nothing is known about it, on every
// 8th epoch (refresh: the distribution may drift, a threshold may be crossed one check later), and when it FIRED the last
// time it ran.  A check that would not fire changes nothing, so skipping it does not move the checks that do.
static bool want_check(const ge_batch* b, int c) {
    if (c < 0 || c >= 64) return true;
    const uint32_t seen = ((volatile uint32_t*)b->h_sched)[c], fired = ((volatile uint32_t*)b->h_sched)[64 + c];
    if (seen == 0 || (b->epoch & 7ull) == 0) return true;
    return fired == seen;
}

// plan -> scatter -> copy back on the histogram of the counted step that just ran (all asynchronous, no host sync)
static int enqueue_regroup(ge_batch* b, cudaStream_t st) {
    const uint32_t S = (uint32_t)b->rec_store;
    k_regroup_plan<<<1, 32, 0, st>>>(b->d_cstate, b->d_rg, b->tab->dev.nonterm, (uint32_t)b->rg_mixed_shift, (uint32_t)b->dead_shift);
    uint64_t g = (b->n + 1023) / 1024;
    if (g > (uint64_t)b->sm_count * 2) g = (uint64_t)b->sm_count * 2;
    k_regroup_scatter<<<(int)g, 1024, 0, st>>>(b->d_tiles, S, b->d_origin, b->d_rg, b->tab->dev.nonterm, b->d_rg_tiles, b->d_rg_origin);
    k_regroup_copyback<<<glue_grid(b, b->n, 256), 256, 0, st>>>(b->d_tiles, S, b->d_origin, b->d_rg, b->d_rg_tiles, b->d_rg_origin, b->d_cstate, b->d_hint, b->d_tile_present);
    CU(cudaGetLastError());
    b->launches += 3;
    b->since_compact = 0;
    return GE_OK;
}

// bytes per session of the human-input rows: one byte per seat, rounded up to 8
static size_t human_stride(const ge_table* t) { return (size_t)((t->P + 7) / 8) * 8; }

// the inputs of the human seats are for ONE step: once a launch has consumed them they are cleared to "has not acted"
static int consume_human_inputs(ge_batch* b, cudaStream_t st) {
    if (b->hchoice_set) {
        CU(cudaMemsetAsync(b->d_hchoice, HUMAN_NONE, b->n * human_stride(b->tab), st));
        b->hchoice_set = false;
    }
    return GE_OK;
}

static void fill_common(const ge_batch* b, StepArgs& a, int steps_per_launch) {
    a.seed = b->seed; a.n_steps = steps_per_launch; a.flags = b->step_flags;
    for (int r = 0; r < 10; ++r) {
        a.rk[2 * r] = (uint32_t)b->seed + (uint32_t)r * 0x9E3779B9u;
        a.rk[2 * r + 1] = (uint32_t)(b->seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
}

// the per-batch arguments of the batch's NEXT step launch (advances its launch index)
static void fill_slot(ge_batch* b, SlotArgs& a, bool count_live, bool regroup) {
    a.tiles = b->d_tiles; a.n_sessions = b->n; a.n_tiles = b->n_tiles; a.first_sid = b->first_sid;
    a.stats = b->d_stats; a.presence = b->d_presence;
    a.n_active = b->d_cstate; a.live_mask = b->d_live_mask; a.live_count = b->d_cstate + 5;
    a.sid_stride = b->sid_stride;
    a.rg = regroup ? b->d_rg : nullptr;
    // per-tile column needs ride on regrouping (tables whose sessions de-synchronise); the device-side auto-reset
    // rewrites every record behind the host's back, so the words are not trusted while it is on
    a.tile_present_out = (regroup && b->sid_stride == 0) ? b->d_tile_present : nullptr;
    a.tile_present = (a.tile_present_out && b->tile_valid) ? b->d_tile_present : nullptr;
    b->tile_valid = a.tile_present_out != nullptr;          // this launch writes the word of every tile it steps
    a.count_live = count_live ? 1u : 0u;
    a.launch_idx = b->launch_idx++;
    a.presence_override = b->next_override;
    b->next_override = 0;
    a.origin = b->compacted ? b->d_origin : nullptr;
    a.human_mask = b->d_hmask; a.human_choice = b->d_hchoice; a.human_stride = (uint32_t)human_stride(b->tab);
}

static int launch_steps(ge_batch* b, int n_launches, int steps_per_launch, cudaStream_t st) {
    if (b->d_hmask && (b->kernel == GE_KERNEL_COOP || steps_per_launch > 1))
        return fail(GE_ERR_UNSUPPORTED, "human seats are served by single-step launches of the thread-per-session kernels");
    // a batch with people at the table runs the run-time-table kernel that carries the human-seat path
    const bool regroup_on = b->regroup_every > 0 && b->kernel != GE_KERNEL_COOP && steps_per_launch == 1;
    const bool tiled = regroup_on && b->sid_stride == 0 && !b->d_hmask && pick_tiled_fn(b->tab, b->kernel) != nullptr;
    {
        const int rc = ensure_store(b, store_should_be_packed(b), st);
        if (rc != GE_OK) return rc;
    }
    const step_fn fn = b->packed ? ((b->kernel == GE_KERNEL_TPS && b->tab->spec.tps_pk) ? b->tab->spec.tps_pk : b->tab->ks.tps_pk)
                     : b->d_hmask ? pick_human_fn(b->tab) : tiled ? pick_tiled_fn(b->tab, b->kernel) : b->fn[b->kernel];
    StepArgs a;
    fill_common(b, a, steps_per_launch);
    // phase regrouping replaces the swap compaction (it also moves finished games behind the live ones)
    const bool regroup = b->regroup_every > 0 && b->kernel != GE_KERNEL_COOP && steps_per_launch == 1;
    for (int i = 0; i < n_launches; ++i) {
        if (regroup && !b->compacted) {              // the origin map must exist before the first regrouping
            if (!b->origin_iota) {
                k_iota<<<glue_grid(b, b->n_tiles * 32, 256), 256, 0, st>>>(b->d_origin, b->n_tiles * 32);
                b->launches++;
            }
            b->compacted = true;
            b->origin_iota = false;
        }
        const bool regroup_after = regroup && b->since_compact + 1 >= b->regroup_every;
        const bool check_due = !regroup && b->compact_every > 0 && b->kernel != GE_KERNEL_COOP && b->since_compact + 1 >= b->compact_every;
        // learned schedule: checks that did not fire the last time they ran are skipped (want_check); only batches whose
        // epochs the host starts itself (no device-side auto-reset) and that are stepped one step per launch are scheduled
        const bool scheduled = check_due && b->sched_valid && b->sid_stride == 0 && !b->d_hmask && steps_per_launch == 1;
        const int check_idx = scheduled ? (int)b->checks_seen++ : -1;
        const bool compact_after = check_due && (!scheduled || want_check(b, check_idx));
        if (check_due && !compact_after) b->since_compact = -1;       // skipped: the cadence restarts with this launch
        fill_slot(b, a, compact_after || regroup_after, regroup);
        if (!tiled) { a.tile_present = nullptr; a.tile_present_out = nullptr; b->tile_valid = false; }
        // the bulk-copy variant of the light path stages its tiles in dynamic shared memory (werewolf single-batch kernels)
        const bool bulk = (a.flags & STEP_LIGHT_BULK) && b->tab->family == FAM_WEREWOLF && !b->d_hmask && b->kernel != GE_KERNEL_COOP;
        if (!bulk) a.flags &= ~(uint32_t)STEP_LIGHT_BULK;
        if (b->pdl && b->kernel != GE_KERNEL_COOP) {        // (the lane-per-player kernels have no device-side wait)
            // the launch may begin while the previous kernel of the stream drains; it waits (griddepcontrol.wait) before
            // it reads anything that kernel wrote
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.gridDim = dim3((unsigned)b->grid[b->kernel]); cfg.blockDim = dim3(128);
            cfg.dynamicSmemBytes = bulk ? sizeof(LightBulk) : 0; cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            CU(cudaLaunchKernelEx(&cfg, fn, b->tab->dev, a));
        } else {
            fn<<<b->grid[b->kernel], 128, bulk ? sizeof(LightBulk) : 0, st>>>(b->tab->dev, a);
        }
        b->launches++;
        b->since_compact++;
        if (b->hchoice_set) { const int rc = consume_human_inputs(b, st); if (rc != GE_OK) return rc; }
        if (regroup_after) {
            const int rc = enqueue_regroup(b, st);
            if (rc != GE_OK) return rc;
        } else if (compact_after) {
            const int rc = enqueue_compaction(&b, 1, st, check_idx);
            if (rc != GE_OK) return rc;
        }
        if ((regroup_after || compact_after) && b->sid_stride != 0) {      // every game over? start the next epoch on the device
            const InitRec rec = init_rec(b);
            k_autoreset_apply<<<glue_grid(b, b->n_tiles * 32, 256), 256, 0, st>>>(
                b->tab->dev, b->d_tiles, (uint32_t)b->rec_store, b->n, b->n_tiles, rec, b->d_origin, b->d_stats, b->d_cstate, alive_mask_of(b));
            k_autoreset_commit<<<1, 32, 0, st>>>(b->d_cstate, b->d_presence, regroup ? b->d_rg : nullptr, b->launch_idx, b->n);
            CU(cudaGetLastError());
            b->launches += 2;
        }
    }
    CU(cudaGetLastError());
    return GE_OK;
}

// One step of every batch of a ring in ONE launch (k_ring_*), n_rounds times, all on the FIRST batch's stream.
// The batches must share table, device, seed, kernel (thread-per-session) and stream, and use neither phase
// regrouping nor auto-reset; compaction checks that fall due are batched into one launch pair per round.
extern "C" int ge_step_ring(ge_batch** batches, int n_batches, int n_rounds) {
    if (!batches || n_batches < 1 || n_batches > GE_RING_MAX || n_rounds < 0) return fail(GE_ERR_ARG, "bad arguments to ge_step_ring (1..16 batches)");
    ge_batch* b0 = batches[0];
    if (!b0) return fail(GE_ERR_ARG, "NULL batch in ge_step_ring");
    for (int i = 0; i < n_batches; ++i) {
        const ge_batch* b = batches[i];
        if (!b) return fail(GE_ERR_ARG, "NULL batch in ge_step_ring");
        if (b->tab != b0->tab || b->device != b0->device || b->seed != b0->seed || b->kernel != b0->kernel || b->stream != b0->stream)
            return fail(GE_ERR_ARG, "ge_step_ring: the batches must share table, device, seed, kernel and stream (ge_batch_set_stream)");
        if (b->kernel == GE_KERNEL_COOP || b->regroup_every > 0 || b->sid_stride != 0 || b->d_hmask)
            return fail(GE_ERR_UNSUPPORTED, "ge_step_ring covers the thread-per-session kernels without phase regrouping / auto-reset / human seats");
        for (int j = 0; j < i; ++j)
            if (batches[j] == b) return fail(GE_ERR_ARG, "ge_step_ring: a batch appears twice");
    }
    CU(cudaSetDevice(b0->device));
    for (int i = 0; i < n_batches; ++i) {
        if (batches[i]->want_packed != b0->want_packed) return fail(GE_ERR_ARG, "ge_step_ring: the batches must share the store format (GE_OPT_STORE_PACKED)");
        const int rc = ensure_store(batches[i], store_should_be_packed(batches[i]), b0->stream);
        if (rc != GE_OK) return rc;
    }
    const ring_fn fn = b0->packed ? ((b0->kernel == GE_KERNEL_TPS && b0->tab->spec.ring_pk) ? b0->tab->spec.ring_pk : b0->tab->ks.ring_pk)
                                  : b0->rfn[b0->kernel];
    if (!fn) return fail(GE_ERR_UNSUPPORTED, "no ring kernel for this table");
    StepArgs c;
    memset(&c, 0, sizeof c);
    fill_common(b0, c, 1);
    // every CTA walks all the slots, so the grid is sized for the largest batch at the kernel's occupancy limit
    uint64_t warps = 0;
    for (int i = 0; i < n_batches; ++i) if (batches[i]->n_tiles > warps) warps = batches[i]->n_tiles;
    uint64_t g = (warps + 3) / 4;
    // (ge_batch_set_grid on the first batch asks for a smaller grid, so that ring launches of OTHER rings on other streams
    // can be resident at the same time)
    int per_sm = b0->occ[b0->kernel];
    if (b0->ctas_per_sm > 0 && b0->ctas_per_sm < per_sm) per_sm = b0->ctas_per_sm;
    const uint64_t cap = (uint64_t)b0->sm_count * per_sm;
    if (g > cap) g = cap;
    for (int r = 0; r < n_rounds; ++r) {
        RingArgs ra;
        ra.n = n_batches;
        ra.rot_div = (uint32_t)b0->sm_count;
        ge_batch* due[GE_RING_MAX];
        int n_due = 0;
        // the ring is checked as a whole, on the first batch's cadence: one launch pair for all its batches
        const bool ring_due = b0->compact_every > 0 && b0->since_compact + 1 >= b0->compact_every;
        for (int i = 0; i < n_batches; ++i) {
            ge_batch* b = batches[i];
            const bool compact_after = ring_due && b->compact_every > 0;
            fill_slot(b, ra.slot[i], compact_after, false);
            b->since_compact++;
            if (compact_after) due[n_due++] = b;
        }
        if (b0->pdl) {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.gridDim = dim3((unsigned)(g < 1 ? 1 : g)); cfg.blockDim = dim3(128); cfg.stream = b0->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            CU(cudaLaunchKernelEx(&cfg, fn, b0->tab->dev, c, ra));
        } else {
            fn<<<(unsigned)(g < 1 ? 1 : g), 128, 0, b0->stream>>>(b0->tab->dev, c, ra);
        }
        b0->launches++;                               // ONE launch for the whole ring
        if (n_due) {
            const int rc = enqueue_compaction(due, n_due, b0->stream);
            if (rc != GE_OK) return rc;
        }
    }
    CU(cudaGetLastError());
    return GE_OK;
}

// A caller-supplied stream is ordered against the batch's own on both sides: it first waits for everything the batch
// has enqueued (initialisation, imports, compaction bookkeeping), and the batch's stream then waits for the work
// enqueued on it, so later exports / steps of the batch see the result.  No host synchronisation.
static int fence_in(ge_batch* b, cudaStream_t other) {
    if (other == b->stream) return GE_OK;
    CU(cudaEventRecord(b->fence, b->stream));
    CU(cudaStreamWaitEvent(other, b->fence, 0));
    return GE_OK;
}
static int fence_out(ge_batch* b, cudaStream_t other) {
    if (other == b->stream) return GE_OK;
    CU(cudaEventRecord(b->fence, other));
    CU(cudaStreamWaitEvent(b->stream, b->fence, 0));
    return GE_OK;
}

extern "C" int ge_step(ge_batch* b, int n_steps, void* cuda_stream) {
    if (!b || n_steps < 0) return fail(GE_ERR_ARG, "bad arguments to ge_step");
    CU(cudaSetDevice(b->device));
    const cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : b->stream;
    int rc = fence_in(b, st);
    if (rc == GE_OK) rc = launch_steps(b, n_steps, 1, st);
    if (rc == GE_OK) rc = fence_out(b, st);
    return rc;
}

// ---- host-side launch workers of ge_step_many --------------------------------------------------------------------
// A step launch of a 2^20-session batch is ~8 us of device time in the co-resident state, so a ring of 8 batches wants
// ~190 000 launches per second (steps + compaction checks): as much as ONE host thread can enqueue (4-7 us per launch
// with this parameter block), and on a slower host the device starves (measured: 4.99e10 instead of 6.2e10 steps/s
// with enqueue time > device time).  The batches of a ring are independent and sit on their own streams, so their
// launch sequences can be enqueued by different host threads: a small persistent pool, worker w takes batches
// w, w + T, ... (a batch always goes to the same worker, so its own launches stay in order).  GE_STEP_THREADS sets
// T (default 4, 1 = the calling thread only).
namespace {
struct StepPool {
    std::mutex m;
    std::condition_variable cv_go, cv_done;
    std::vector<std::thread> workers;
    ge_batch** batches = nullptr;
    int n = 0, rounds = 0, T = 1, pending = 0;
    unsigned long long gen = 0;
    bool stop = false;
    int rc[16];
    std::string err[16];

    void share(int w) {
        int cur_dev = -1, r_ = GE_OK;
        for (int r = 0; r < rounds && r_ == GE_OK; ++r)
            for (int i = w; i < n && r_ == GE_OK; i += T) {
                ge_batch* b = batches[i];
                if (b->device != cur_dev) {
                    if (cudaSetDevice(b->device) != cudaSuccess) { r_ = fail(GE_ERR_CUDA, "cudaSetDevice failed in a launch worker"); break; }
                    cur_dev = b->device;
                }
                r_ = launch_steps(b, 1, 1, b->stream);
            }
        rc[w] = r_;
        if (r_ != GE_OK) err[w] = g_err;
    }
    void loop(int w) {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv_go.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
            }
            share(w);
            std::lock_guard<std::mutex> lk(m);
            if (--pending == 0) cv_done.notify_one();
        }
    }
    explicit StepPool(int t) : T(t) {
        for (int w = 1; w < T; ++w) workers.emplace_back([this, w] { loop(w); });
    }
    ~StepPool() {
        { std::lock_guard<std::mutex> lk(m); stop = true; }
        cv_go.notify_all();
        for (auto& t : workers) t.join();
    }
    int run(ge_batch** b, int n_, int rounds_) {
        {
            std::lock_guard<std::mutex> lk(m);
            batches = b; n = n_; rounds = rounds_; pending = T - 1; ++gen;
        }
        cv_go.notify_all();
        share(0);
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return pending == 0; });
        for (int w = 0; w < T; ++w)
            if (rc[w] != GE_OK) return fail(rc[w], err[w]);
        return GE_OK;
    }
};
std::mutex g_pool_mutex;
StepPool* g_pool = nullptr;
int step_threads() {
    static const int t = [] {
        const char* e = getenv("GE_STEP_THREADS");
        int v = e ? atoi(e) : 4;
        const unsigned hw = std::thread::hardware_concurrency();
        if (hw && (unsigned)v > hw) v = (int)hw;
        return v < 1 ? 1 : v > 16 ? 16 : v;
    }();
    return t;
}
}  // namespace

extern "C" int ge_step_many(ge_batch** batches, int n_batches, int n_rounds) {
    if (!batches || n_batches < 0 || n_rounds < 0) return fail(GE_ERR_ARG, "bad arguments to ge_step_many");
    for (int i = 0; i < n_batches; ++i)
        if (!batches[i]) return fail(GE_ERR_ARG, "NULL batch in ge_step_many");
    // several launch workers pay off for a ring of distinct batches and more than a handful of launches
    bool distinct = true;
    for (int i = 0; i < n_batches && distinct; ++i)
        for (int j = 0; j < i; ++j)
            if (batches[j] == batches[i]) { distinct = false; break; }
    const int T = step_threads();
    if (T > 1 && distinct && n_batches >= 4 && (long long)n_batches * n_rounds >= 16) {
        std::lock_guard<std::mutex> lk(g_pool_mutex);            // one ge_step_many at a time uses the pool
        if (!g_pool) g_pool = new (std::nothrow) StepPool(T);
        if (g_pool) return g_pool->run(batches, n_batches, n_rounds);
    }
    int cur_dev = -1;
    for (int r = 0; r < n_rounds; ++r)
        for (int i = 0; i < n_batches; ++i) {
            ge_batch* b = batches[i];
            if (b->device != cur_dev) { CU(cudaSetDevice(b->device)); cur_dev = b->device; }
            const int rc = launch_steps(b, 1, 1, b->stream);
            if (rc != GE_OK) return rc;
        }
    return GE_OK;
}

extern "C" int ge_run_fused(ge_batch* b, int n_steps, void* cuda_stream) {
    if (!b || n_steps < 0) return fail(GE_ERR_ARG, "bad arguments to ge_run_fused");
    if (n_steps == 0) return GE_OK;
    CU(cudaSetDevice(b->device));
    const cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : b->stream;
    int rc = fence_in(b, st);
    if (rc == GE_OK) rc = launch_steps(b, 1, n_steps, st);
    if (rc == GE_OK) rc = fence_out(b, st);
    return rc;
}

extern "C" int ge_sync(ge_batch* b) {
    if (!b) return fail(GE_ERR_ARG, "batch is NULL");
    CU(cudaSetDevice(b->device));
    SYNC(b);
    return GE_OK;
}

static int ensure_stage(ge_batch* b, size_t bytes) {
    if (b->stage_bytes >= bytes) return GE_OK;
    if (b->d_stage) { CU(cudaStreamSynchronize(b->stream)); CU(cudaFree(b->d_stage)); b->d_stage = nullptr; b->stage_bytes = 0; }
    cudaError_t e = cudaMalloc(&b->d_stage, bytes);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? GE_ERR_NOMEM : GE_ERR_CUDA, cudaGetErrorString(e));
    b->stage_bytes = bytes;
    return GE_OK;
}

// Bring the session store to the wanted layout (canonical columns <-> packed, ge_batch::packed).  Slot by slot, so the
// slot order (compaction's origin map, the active prefix) is untouched; through the staging buffer, stream-ordered.
static int ensure_store(ge_batch* b, bool packed, cudaStream_t st) {
    if (b->packed == packed) return GE_OK;
    const size_t S_dst = packed ? packed_record(b->tab) : b->tab->rec_dev;
    const size_t bytes = (size_t)b->n_tiles * 32 * S_dst;
    if (st != b->stream) CU(cudaStreamSynchronize(b->stream));      // the staging buffer belongs to the batch's own stream
    int rc = ensure_stage(b, bytes);
    if (rc) return rc;
    if (b->tab->bucket == 8) k_repack<8><<<glue_grid(b, b->n_tiles * 32, 256), 256, 0, st>>>(b->d_tiles, b->d_stage, b->n_tiles * 32, packed ? 1 : 0);
    else k_repack<16><<<glue_grid(b, b->n_tiles * 32, 256), 256, 0, st>>>(b->d_tiles, b->d_stage, b->n_tiles * 32, packed ? 1 : 0);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(b->d_tiles, b->d_stage, bytes, cudaMemcpyDeviceToDevice, st));
    if (st != b->stream) CU(cudaStreamSynchronize(st));
    b->launches++;
    b->packed = packed;
    b->rec_store = S_dst;
    b->tile_valid = false;
    return GE_OK;
}

static int export_async(ge_batch* b, uint64_t first, uint64_t count, void* host_buf, int wire) {
    const bool dense = wire == GE_WIRE_DENSE && has_dense(b->tab);
    const size_t S = dense ? ge_table_wire_size(b->tab, GE_WIRE_DENSE) : b->tab->rec_canon;
    int rc = ensure_stage(b, count * S);
    if (rc) return rc;
    if (dense || b->packed) {
        const uint32_t* org = b->compacted ? b->d_origin : nullptr;
        const int g = glue_grid(b, org ? b->n : count, 256);
        const bool p8 = b->tab->bucket == 8;
        if (b->packed && dense && p8) k_export_w<8, true, true><<<g, 256, 0, b->stream>>>(b->d_tiles, org, b->n, first, count, b->d_stage);
        else if (b->packed && dense) k_export_w<16, true, true><<<g, 256, 0, b->stream>>>(b->d_tiles, org, b->n, first, count, b->d_stage);
        else if (b->packed && p8) k_export_w<8, false, true><<<g, 256, 0, b->stream>>>(b->d_tiles, org, b->n, first, count, b->d_stage);
        else if (b->packed) k_export_w<16, false, true><<<g, 256, 0, b->stream>>>(b->d_tiles, org, b->n, first, count, b->d_stage);
        else if (p8) k_export_w<8, true, false><<<g, 256, 0, b->stream>>>(b->d_tiles, org, b->n, first, count, b->d_stage);
        else k_export_w<16, true, false><<<g, 256, 0, b->stream>>>(b->d_tiles, org, b->n, first, count, b->d_stage);
    } else if (b->compacted)
        k_export_perm<<<glue_grid(b, b->n, 256), 256, 0, b->stream>>>(b->d_tiles, (uint32_t)b->rec_store, (uint32_t)S, b->d_origin, b->n, first, count, b->d_stage);
    else
        k_export<<<glue_grid(b, count, 256), 256, 0, b->stream>>>(b->d_tiles, (uint32_t)b->rec_store, (uint32_t)S, first, count, b->d_stage);
    CU(cudaGetLastError());
    b->launches++;
    CU(cudaMemcpyAsync(host_buf, b->d_stage, count * S, cudaMemcpyDeviceToHost, b->stream));
    return GE_OK;
}

// Undo the slot permutation of compaction: records go back to slot == session index, every slot active.
// keep_records = false: the caller is about to overwrite every record, only the bookkeeping is reset.
// The device epoch of auto-reset (cstate[8]) survives: the resident sessions' ids depend on it.
static int restore_order(ge_batch* b, bool keep_records) {
    if (b->compacted && keep_records) {
        const size_t S = b->tab->rec_canon;
        int rc = ensure_stage(b, b->n * S);
        if (rc) return rc;
        const InitRec rec = init_rec_canon(b);
        if (b->packed && b->tab->bucket == 8) {
            k_export_w<8, false, true><<<glue_grid(b, b->n, 256), 256, 0, b->stream>>>(b->d_tiles, b->d_origin, b->n, 0, b->n, b->d_stage);
            k_import_w<8, false, true><<<glue_grid(b, b->n, 256), 256, 0, b->stream>>>(b->tab->dev, rec, b->d_tiles, 0, b->n, b->d_stage, nullptr, ImportReset{nullptr, nullptr, 0, 0});
        } else if (b->packed) {
            k_export_w<16, false, true><<<glue_grid(b, b->n, 256), 256, 0, b->stream>>>(b->d_tiles, b->d_origin, b->n, 0, b->n, b->d_stage);
            k_import_w<16, false, true><<<glue_grid(b, b->n, 256), 256, 0, b->stream>>>(b->tab->dev, rec, b->d_tiles, 0, b->n, b->d_stage, nullptr, ImportReset{nullptr, nullptr, 0, 0});
        } else {
        k_export_perm<<<glue_grid(b, b->n, 256), 256, 0, b->stream>>>(b->d_tiles, (uint32_t)b->rec_store, (uint32_t)S, b->d_origin, b->n, 0, b->n, b->d_stage);
        k_import<<<glue_grid(b, b->n, 256), 256, 0, b->stream>>>(b->tab->dev, rec, b->d_tiles, (uint32_t)b->rec_store, (uint32_t)S, 0, b->n, b->d_stage, nullptr, ImportReset{nullptr, nullptr, 0, 0});
        }
        CU(cudaGetLastError());
        b->launches += 2;
    }
    b->compacted = false;
    b->origin_iota = false;
    b->sched_valid = false;
    b->tile_valid = false;
    b->epoch++;
    k_cstate_reset<<<1, 32, 0, b->stream>>>(b->d_cstate, b->n, b->epoch, 1);
    CU(cudaGetLastError());
    b->since_compact = 0;
    CU(cudaMemsetAsync(b->d_presence, 0, 3 * sizeof(uint32_t), b->stream));
    b->next_override = 0xFFFFFFFFu;
    return GE_OK;
}

// Every import path comes through here: the records are range-checked ON THE DEVICE by k_import (record_invalid,
// ge_glue.cuh); a bad record bumps two words of mapped pinned host memory (system-scope atomics, only on the rare
// malformed record), which sync_and_check reads after the next synchronisation and clears.
static int import_async(ge_batch* b, uint64_t first, uint64_t count, const void* host_buf, int wire) {
    const bool dense = wire == GE_WIRE_DENSE && has_dense(b->tab);
    const size_t S = dense ? ge_table_wire_size(b->tab, GE_WIRE_DENSE) : b->tab->rec_canon;
    // imported sessions may be live anywhere: slot order and the active prefix go back to "everything".  A whole-batch
    // import overwrites every record, so only the bookkeeping is reset — by the import kernel itself.
    const bool whole = first == 0 && count == b->n;
    ImportReset R{b->d_presence, nullptr, b->n, 0};
    b->tile_valid = false;
    if (whole) {
        b->compacted = false;
        b->origin_iota = false;
        b->sched_valid = false;
        b->epoch++;
        b->since_compact = 0;
        R.cstate = b->d_cstate; R.epoch = b->epoch;
    } else {
        const int rc0 = restore_order(b, true);
        if (rc0) return rc0;
    }
    int rc = ensure_stage(b, count * S);
    if (rc) return rc;
    const InitRec rec = init_rec_canon(b);
    CU(cudaMemcpyAsync(b->d_stage, host_buf, count * S, cudaMemcpyHostToDevice, b->stream));
    const int g = glue_grid(b, count, 256);
    const bool p8 = b->tab->bucket == 8;
    if (b->packed && dense && p8) k_import_w<8, true, true><<<g, 256, 0, b->stream>>>(b->tab->dev, rec, b->d_tiles, first, count, b->d_stage, b->d_err, R);
    else if (b->packed && dense) k_import_w<16, true, true><<<g, 256, 0, b->stream>>>(b->tab->dev, rec, b->d_tiles, first, count, b->d_stage, b->d_err, R);
    else if (b->packed && p8) k_import_w<8, false, true><<<g, 256, 0, b->stream>>>(b->tab->dev, rec, b->d_tiles, first, count, b->d_stage, b->d_err, R);
    else if (b->packed) k_import_w<16, false, true><<<g, 256, 0, b->stream>>>(b->tab->dev, rec, b->d_tiles, first, count, b->d_stage, b->d_err, R);
    else if (dense && b->tab->bucket == 8) k_import_w<8, true, false><<<g, 256, 0, b->stream>>>(b->tab->dev, rec, b->d_tiles, first, count, b->d_stage, b->d_err, R);
    else if (dense) k_import_w<16, true, false><<<g, 256, 0, b->stream>>>(b->tab->dev, rec, b->d_tiles, first, count, b->d_stage, b->d_err, R);
    else k_import<<<g, 256, 0, b->stream>>>(b->tab->dev, rec, b->d_tiles, (uint32_t)b->rec_store, (uint32_t)S, first, count, b->d_stage, b->d_err, R);
    CU(cudaGetLastError());
    b->launches++;
    b->next_override = 0xFFFFFFFFu;   // imported sessions can be in any phase (the import kernel cleared the presence words)
    return GE_OK;
}

extern "C" int ge_export_state(ge_batch* b, uint64_t first, uint64_t count, void* host_buf) {
    if (!b || !host_buf || first + count > b->n) return fail(GE_ERR_ARG, "bad arguments to ge_export_state");
    if (count == 0) return GE_OK;
    CU(cudaSetDevice(b->device));
    int rc = export_async(b, first, count, host_buf, b->wire);
    if (rc) return rc;
    SYNC(b);
    return GE_OK;
}

extern "C" int ge_import_state(ge_batch* b, uint64_t first, uint64_t count, const void* host_buf) {
    if (!b || !host_buf || first + count > b->n) return fail(GE_ERR_ARG, "bad arguments to ge_import_state");
    if (count == 0) return GE_OK;
    CU(cudaSetDevice(b->device));
    int rc = import_async(b, first, count, host_buf, b->wire);
    if (rc) return rc;
    SYNC(b);
    return GE_OK;
}

// Step log: canonical records of a window of sessions after each of n_steps single-step launches (record 0 =
// the state before the first of them).  The whole batch is stepped; only the window is exported.
extern "C" int ge_trace(ge_batch* b, uint64_t first, uint64_t count, int n_steps, void* host_records) {
    if (!b || !host_records || n_steps < 0 || count == 0 || first + count > b->n) return fail(GE_ERR_ARG, "bad arguments to ge_trace");
    CU(cudaSetDevice(b->device));
    const size_t frame = count * b->tab->rec_canon;
    uint8_t* out = static_cast<uint8_t*>(host_records);
    int rc = export_async(b, first, count, out, GE_WIRE_CANONICAL);
    for (int k = 1; rc == GE_OK && k <= n_steps; ++k) {
        rc = launch_steps(b, 1, 1, b->stream);
        if (rc == GE_OK) rc = export_async(b, first, count, out + (size_t)k * frame, GE_WIRE_CANONICAL);
    }
    if (rc != GE_OK) return rc;
    SYNC(b);
    return GE_OK;
}

extern "C" int ge_eval_preds(ge_batch* b, const ge_pred_t* preds, int n_preds, uint64_t first, uint64_t count, uint32_t* host_masks) {
    if (!b || !preds || !host_masks || n_preds < 1 || n_preds > 32 || first + count > b->n) return fail(GE_ERR_ARG, "bad arguments to ge_eval_preds");
    if (count == 0) return GE_OK;
    CU(cudaSetDevice(b->device));
    const size_t bytes = count * (size_t)n_preds * sizeof(uint32_t);
    int rc = ensure_stage(b, bytes);
    if (rc) return rc;
    PredList pl;
    memset(&pl, 0, sizeof pl);
    memcpy(pl.p, preds, (size_t)n_preds * sizeof(ge_pred_t));
    pl.n = n_preds;
    k_eval_preds<<<glue_grid(b, b->n, 256), 256, 0, b->stream>>>(b->tab->dev, pl, b->d_tiles, (uint32_t)b->rec_store,
                                                               b->compacted ? b->d_origin : nullptr, b->n, first, count,
                                                               reinterpret_cast<uint32_t*>(b->d_stage), b->packed ? b->tab->bucket : 0);
    CU(cudaGetLastError());
    b->launches++;
    CU(cudaMemcpyAsync(host_masks, b->d_stage, bytes, cudaMemcpyDeviceToHost, b->stream));
    SYNC(b);
    return GE_OK;
}

extern "C" int ge_stats_refresh(ge_batch* b, void* cuda_stream) {
    if (!b) return fail(GE_ERR_ARG, "batch is NULL");
    CU(cudaSetDevice(b->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : b->stream;
    int rc = fence_in(b, st);
    if (rc != GE_OK) return rc;
    // snapshot = accumulator + histograms of the sessions currently resident
    CU(cudaMemcpyAsync(b->d_stats_out, b->d_stats, GE_STATS_LEN * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    k_stats<<<glue_grid(b, b->n, 256), 256, 0, st>>>(b->tab->dev, b->d_tiles, (uint32_t)b->rec_store, b->n, b->d_stats_out, alive_mask_of(b));
    CU(cudaGetLastError());
    b->launches++;
    return fence_out(b, st);
}

extern "C" int ge_stats(ge_batch* b, uint64_t* host_hist, size_t n) {
    if (!b || !host_hist || n < GE_STATS_LEN) return fail(GE_ERR_ARG, "bad arguments to ge_stats");
    int rc = ge_stats_refresh(b, nullptr);
    if (rc) return rc;
    CU(cudaMemcpyAsync(host_hist, b->d_stats_out, GE_STATS_LEN * sizeof(uint64_t), cudaMemcpyDeviceToHost, b->stream));
    SYNC(b);
    return GE_OK;
}

extern "C" void* ge_stats_device_ptr(ge_batch* b) { return b ? (void*)b->d_stats_out : nullptr; }

extern "C" int ge_counted_steps(ge_batch* b, uint64_t* out) {
    if (!b || !out) return fail(GE_ERR_ARG, "bad arguments to ge_counted_steps");
    CU(cudaSetDevice(b->device));
    CU(cudaMemcpyAsync(out, b->d_stats, sizeof(uint64_t), cudaMemcpyDeviceToHost, b->stream));
    SYNC(b);
    return GE_OK;
}

// Asynchronous variant for pipelined measurement: the counted-steps word as of this point of the batch's stream is
// copied to *pinned_out (page-locked host memory) without synchronising; valid after the next ge_sync.
extern "C" int ge_counted_steps_async(ge_batch* b, uint64_t* pinned_out) {
    if (!b || !pinned_out) return fail(GE_ERR_ARG, "bad arguments to ge_counted_steps_async");
    CU(cudaSetDevice(b->device));
    CU(cudaMemcpyAsync(pinned_out, b->d_stats, sizeof(uint64_t), cudaMemcpyDeviceToHost, b->stream));
    return GE_OK;
}

extern "C" int ge_run_host_async(ge_batch* b, const void* records_in, void* records_out, int n_steps, uint64_t* host_stats) {
    if (!b || n_steps < 0) return fail(GE_ERR_ARG, "bad arguments to ge_run_host_async");
    CU(cudaSetDevice(b->device));
    int rc;
    if (records_in && (rc = import_async(b, 0, b->n, records_in, b->wire)) != GE_OK) return rc;
    if (b->host_fused && n_steps > 1) rc = launch_steps(b, 1, n_steps, b->stream);
    else rc = launch_steps(b, n_steps, 1, b->stream);
    if (rc != GE_OK) return rc;
    if (records_out && (rc = export_async(b, 0, b->n, records_out, b->wire)) != GE_OK) return rc;
    if (host_stats) {
        if ((rc = ge_stats_refresh(b, nullptr)) != GE_OK) return rc;
        CU(cudaMemcpyAsync(host_stats, b->d_stats_out, GE_STATS_LEN * sizeof(uint64_t), cudaMemcpyDeviceToHost, b->stream));
    }
    return GE_OK;
}

extern "C" int ge_run_host(ge_batch* b, const void* records_in, void* records_out, int n_steps, uint64_t* host_stats) {
    const int rc = ge_run_host_async(b, records_in, records_out, n_steps, host_stats);
    if (rc != GE_OK) return rc;
    SYNC(b);
    return GE_OK;
}

extern "C" int ge_host_alloc(void** p, size_t bytes) {
    if (!p) return fail(GE_ERR_ARG, "p is NULL");
    CU(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
    return GE_OK;
}
extern "C" void ge_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" void* ge_state_device_ptr(ge_batch* b) { return b ? (void*)b->d_tiles : nullptr; }
extern "C" size_t ge_state_device_bytes(const ge_batch* b) { return b ? (size_t)b->n_tiles * 32 * b->rec_store : 0; }
extern "C" uint64_t ge_launch_count(const ge_batch* b) { return b ? b->launches : 0; }
