/*
 * game_engine_b200.h — C ABI of the batched referee/phase-step simulator (B200, sm_100a).
 *
 * This is the drop-in boundary for the ONE data-parallel path of liruihan000/game_engine: the
 * deterministic core of a graph run BotBehaviorNode -> PhaseNode -> RefereeNode.  Every entry point
 * names the reference interface it replaces (paths under the reference repo).  Plain pointers and
 * sizes only; bind with ctypes / cffi / cgo.
 *
 * All functions return 0 on success or a negative GE_ERR_* code; ge_last_error() gives the message
 * of the last failure on the calling thread.  The library is re-entrant across handles; one handle
 * must not be used from two threads at once.  There is no CPU fallback: without a CUDA device every
 * compute call fails with GE_ERR_CUDA.
 */
#ifndef GAME_ENGINE_B200_H
#define GAME_ENGINE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GE_STATS_LEN 560          /* u64 words, layout in SPEC.md section 6 */
#define GE_MAX_PHASES 32
#define GE_MAX_PREDS 32

#define GE_OK 0
#define GE_ERR_ARG (-1)           /* bad argument / malformed table blob */
#define GE_ERR_CUDA (-2)          /* CUDA runtime error (message has the cudaError string) */
#define GE_ERR_UNSUPPORTED (-3)   /* table uses a family / player count the kernels do not cover */
#define GE_ERR_NOMEM (-4)

/* record formats at the host boundary (ge_batch_set_wire) */
#define GE_WIRE_CANONICAL 0       /* SPEC.md section 5: S = 48 + P8 (werewolf), roundup8(8 + 4P) (TTL) */
#define GE_WIRE_DENSE 1           /* SPEC.md section 5b: werewolf tables up to 16 players, masks as u8 / u16: 32 / 48 bytes */

/* kernel mappings (ge_batch_set_kernel) */
#define GE_KERNEL_AUTO 0
#define GE_KERNEL_COOP 1          /* one lane per player, warp ballots / match_any (32/P sessions per warp) */
#define GE_KERNEL_TPS 2           /* one thread per session, bit-sliced tallies; uses the build-time specialised
                                     instantiation when the table is byte-identical to a shipped game's */
#define GE_KERNEL_TPS_GENERIC 3   /* one thread per session, always the run-time table interpreter */

/* ---- binary transition table (produced by game_engine_b200/compiler.py) -------------------------
 * Replaces: the per-step LLM reading of dsl['phases'] (reference agent/game_agent_v2.py:1022-1103,
 * accessor agent/tools/utils.py:19-31, loader utils.py:557-581). */
typedef struct {
    uint8_t op, next, tag, a;
    uint32_t arg;
} ge_branch_t;

typedef struct {
    uint8_t id, kind, action_op, action_arg, action_flags, exit_op, entry_op, n_branches;
    uint8_t actor_pred, pad[7];
    ge_branch_t br[4];
} ge_phase_t;

/* DNF of two clauses over the mask fields of SPEC.md section 2; clause = AND(pos fields) & ~OR(neg fields).  Bit 15 of
 * pos0 = "continued": the predicate is this record OR the next one (a DNF of any length is a run of records). */
typedef struct {
    uint16_t pos0, neg0, pos1, neg1;
} ge_pred_t;
#define GE_PRED_CONTINUED 0x8000u

/* Comparison field: "player's value field <op> constant" as a derived mask field (numeric conditions of the DSL).
 * Comparison k is mask field 13 + k (werewolf, k < 2; value field 0 = selected_target_id) or 11 + k (TTL, k < 4; value
 * fields 0 total_score, 1 rounds_as_speaker, 2 vote_choice).  ops: 0 ==, 1 !=, 2 <, 3 <=, 4 >, 5 >=. */
typedef struct {
    uint8_t value_field, op, constant;
} ge_cmp_t;

typedef struct {
    char magic[4];                /* "GETB" */
    uint16_t version;             /* 1 */
    uint8_t family, n_phases, n_players, n_preds, n_wolves, rounds, max_revotes, n_cmp, reserved[2];
    uint32_t init_masks;
    ge_cmp_t cmp[4];
} ge_table_header_t;

typedef struct ge_table ge_table;
typedef struct ge_batch ge_batch;

/* Parse and validate a table blob.  Replaces load_dsl_by_gamename + yaml.safe_load per thread
 * (reference agent/tools/utils.py:557-581). */
int ge_table_create(const uint8_t *blob, size_t n, ge_table **out);
void ge_table_destroy(ge_table *t);
/* canonical packed record size S in bytes (SPEC.md section 5) */
size_t ge_table_record_size(const ge_table *t);
int ge_table_n_players(const ge_table *t);
/* Bytes per session that a step STARTING in phase `phase_index` must read and may write, given the column layout of
 * the session store (column 0 = header + is_alive + can_vote always moves; the other column groups only when a
 * predicate or effect of the phase touches them).  Weighted by the visit histogram of the statistics this gives the
 * necessary DRAM traffic per session-phase-step, the honest lower bound next to the 2*S "algorithmic" figure. */
int ge_table_phase_io(const ge_table *t, int phase_index, uint32_t *read_bytes, uint32_t *write_bytes);
/* The same for the packed session store (GE_OPT_STORE_PACKED; werewolf-family tables up to 8 players, else
 * GE_ERR_UNSUPPORTED): column D0 (16 bytes) both ways on every step, column D1 (16 bytes) only when the phase reads or
 * changes the role bytes or the target bytes. */
int ge_table_phase_io_packed(const ge_table *t, int phase_index, uint32_t *read_bytes, uint32_t *write_bytes);

/* Allocate n_sessions sessions on `device` and initialise them from the DSL template.
 * Session i has id first_session_id + i.  Replaces initialize_player_states_from_dsl
 * (reference agent/tools/utils.py:584-653) and AgentState defaults (game_agent_v2.py:97-117). */
int ge_batch_create(ge_table *t, int device, uint64_t n_sessions, uint64_t first_session_id, uint64_t seed,
                    ge_batch **out);
/* Re-initialise all sessions with new ids / seed.  Statistics are cumulative over the life of the
 * handle: the final-state histograms of the sessions being replaced are folded into the accumulator
 * first, so win rates cover every session the batch ever simulated.  ge_batch_clear_stats zeroes it. */
int ge_batch_reset(ge_batch *b, uint64_t first_session_id, uint64_t seed);
int ge_batch_clear_stats(ge_batch *b);
void ge_batch_destroy(ge_batch *b);
/* Bind all work of this batch to an external CUDA stream (e.g. the caller's torch stream); NULL
 * returns to the batch's own stream.  The stream is borrowed, not owned. */
int ge_batch_set_stream(ge_batch *b, void *cuda_stream);
int ge_batch_set_kernel(ge_batch *b, int kernel);
/* Active-prefix compaction (thread-per-session kernels): every `every_n_steps` launches a device-side check
 * runs and, when at least 1/2^min_dead_shift of the active prefix holds finished games, the live sessions
 * are swapped in front of them (on the device, no host sync) so later launches walk only the live prefix.
 * Session ids, export order and statistics are unaffected.  every_n_steps = 0 turns it off; default (5, 2) for the
 * werewolf family and off for the TTL family (fixed-length games: nothing to compact).
 * A batch that is re-initialised epoch after epoch (ge_batch_reset) learns which of an epoch's checks fire — the check kernel
 * logs it through mapped host memory — and stops launching the ones that do not (every 8th epoch observes all of them again);
 * results never depend on when compaction happens.  Calling this function forgets the learned schedule.
 * ge_batch_active returns the current prefix length (synchronises); ge_batch_active_hint the value of the most recent
 * completed check without synchronising. */
int ge_batch_set_compaction(ge_batch *b, int every_n_steps, int min_dead_shift);
int ge_batch_active(ge_batch *b, uint64_t *out);
/* Launch geometry: the step kernels run a persistent grid of (SM count x ctas_per_sm) CTAs of 128 threads, by
 * default as many as fit (the kernel's occupancy limit, 8 at 8 players).  A smaller grid leaves room for the
 * step launches of OTHER batches on other streams to be resident at the same time, so that one launch's ramp
 * and tail overlap another's steady state (ring of 8 batches on 8 streams: 3 CTAs per SM is ~12% faster than 8).
 * 0 restores the default. */
int ge_batch_set_grid(ge_batch *b, int ctas_per_sm);
/* Phase regrouping (thread-per-session werewolf kernels): every `every_n_steps` launches a device-side check runs
 * and, when at least 1/2^min_mixed_shift of the tiles hold sessions in more than one phase (or the compaction
 * threshold of dead sessions is reached), the active prefix is counting-sorted by phase, finished games last, so
 * that a warp's 32 sessions share a phase again.  Needed by tables whose sessions de-synchronise (loops of
 * data-dependent length such as tie -> re-vote); it replaces the swap compaction while on and is enabled
 * automatically (5, 3) for tables that use the TIE_PENDING branch.  Costs a second session store.  Results,
 * session ids, export order and statistics are unaffected.  every_n_steps = 0 turns it off. */
int ge_batch_set_regroup(ge_batch *b, int every_n_steps, int min_mixed_shift);
/* Auto-reset (continuous simulation): with sid_stride != 0, whenever the compaction / regrouping check finds that
 * every game of the batch is over, the batch starts its next EPOCH on the device — final-state histograms folded
 * into the statistics, all slots re-initialised, session i gets id first_session_id + epoch * sid_stride + i — with
 * no host round trip and no launches wasted on finished games.  Needs compaction or regrouping on (the thread-per-
 * session kernels; switched on (8, 2) if neither is); sid_stride must be >= n_sessions.  ge_batch_reset returns to epoch 0.  ge_batch_epochs
 * reports the number of device-side re-initialisations so far (synchronises). */
int ge_batch_set_autoreset(ge_batch *b, uint64_t sid_stride);
int ge_batch_epochs(ge_batch *b, uint64_t *out);
/* Non-blocking variant: the value an asynchronous copy brought to the host after the most recent completed
 * compaction of the current epoch (an upper bound that only decreases; n_sessions until the first one).
 * 0 means every game of the batch is over — the cue to re-initialise it without waiting for a step cap. */
int ge_batch_active_hint(ge_batch *b, uint64_t *out);
int ge_batch_get_kernel(const ge_batch *b);
/* Tuning options of a batch.  GE_OPT_LIGHT_BULK (werewolf family, specialised / interpreter single-batch kernels):
 * launches whose sessions are all in header-only phases fetch their 512-byte columns with cp.async.bulk into shared
 * memory, completion counted by an mbarrier, double-buffered one group of four tiles ahead, instead of warp-wide
 * 128-bit loads.  Off by default (measured: DESIGN section 6).
 * GE_OPT_STORE_PACKED (werewolf-family tables up to 8 players): the session store in HBM keeps a record as the 32 bytes of
 * the dense wire format (two 16-byte columns: header + eight mask bytes | role bytes + target bytes) instead of the
 * canonical 56 bytes in 3.5 columns whose mask words are three-quarters zeros; the all-bot thread-per-session kernels
 * run on it directly.  Records at the ABI are unchanged (canonical or dense wire, as ge_batch_set_wire says); callers
 * the packed layout does not serve (lane-per-player kernels, human seats, phase regrouping) convert the
 * store back transparently.  ON by default for the tables it covers (value 0 = keep the canonical columns);
 * GE_ERR_UNSUPPORTED when switched on for other tables. */
#define GE_OPT_LIGHT_BULK 1
#define GE_OPT_STORE_PACKED 2
/* GE_OPT_PDL: step launches (ge_step, ge_step_many, ge_step_ring) are programmatic dependent launches: a launch may start
 * filling the machine while the previous kernel of its stream drains and waits on the device (griddepcontrol.wait) before it
 * reads that kernel's results.  On by default (measured: DESIGN section 6: +0.5 % with 8 streams, +16 % on one); 0 = plain
 * stream-ordered launches. */
#define GE_OPT_PDL 3
int ge_batch_set_option(ge_batch *b, int option, int value);

/* Record format of the host-buffer calls of this batch (ge_export_state, ge_import_state, ge_run_host[_async]):
 * canonical (default) or dense.  The dense format carries the same fields in 32 bytes (<= 8 players) or 48 bytes
 * (<= 16 players) instead of 56 / 64 — the conversion runs inside the import / export kernels, so 1.75x / 1.33x fewer
 * bytes cross PCIe per session.  Tables it does not cover (more than 16 players, TTL) keep the canonical record:
 * ge_table_wire_size / ge_batch_wire_size give the size in force.  ge_trace always writes canonical records. */
size_t ge_table_wire_size(const ge_table *t, int wire);
int ge_batch_set_wire(ge_batch *b, int wire);
size_t ge_batch_wire_size(const ge_batch *b);

/* Human seats (SPEC.md section 1, D3h).  Replaces: the reference excluding the human player from the bots' actions
 * (agent/prompt/bot_behavior_system_prompt.txt:3,58-61), logging the person's action in the router
 * (agent/tools/utils.py:310-358, called agent/game_agent_v2.py:324-332) and PhaseNode staying at the phase — while
 * still appending to phase_history — until every target player has acted (game_agent_v2.py:1144-1170,1206-1215).
 * host_masks[i]: seats of session i played by people (bit p-1 = player p); NULL = all bots again.
 * host_choices[i * ge_table_human_stride(t) + p]: the input of seat p+1 for the NEXT step launch (a player id for
 * PICK_PLAYER, 1..n for PICK_OPTION, anything for MARK; 0xFF = has not acted); the launch consumes them.  A session
 * whose acting human seats are not all answered with a valid input STAYS: step + 1, prev = phase, nothing else
 * changes, bots do not draw (they act on the step that completes the phase).  Thread-per-session kernels,
 * single-step launches (ge_step, ge_step_many, ge_run_host with the fused mode off); such a batch runs the
 * run-time-table kernel. */
int ge_batch_set_human_seats(ge_batch *b, const uint32_t *host_masks);
int ge_batch_set_human_choices(ge_batch *b, const uint8_t *host_choices);
size_t ge_table_human_stride(const ge_table *t);

/* Apply n_steps session-phase-steps to every non-terminal session: n_steps launches of the step
 * kernel on `cuda_stream` (NULL = the batch's own stream), each reading and writing the state once.
 * Asynchronous.  A caller-supplied stream is fenced against the batch's own stream with events on both sides (it
 * waits for the batch's pending work; the batch's later work waits for it).  Replaces one graph run BotBehaviorNode -> PhaseNode -> RefereeNode per step
 * (reference agent/game_agent_v2.py:468-617, 987-1241, 619-803) including the tool applications
 * _execute_update_player_actions / _execute_update_player_state (agent/tools/backend_tools.py:285-344,
 * 204-225). */
int ge_step(ge_batch *b, int n_steps, void *cuda_stream);
/* Round-robin over several batches: n_rounds times, one ge_step(b, 1) for every batch in order, each on
 * its own (bound) stream.  One call instead of n_rounds * n_batches keeps the host out of the way when the
 * launches are short. */
int ge_step_many(ge_batch **batches, int n_batches, int n_rounds);
/* The same round-robin as ONE launch per round: a ring kernel whose CTAs walk the batches in order, on the first
 * batch's stream, with the compaction checks that fall due batched into one launch pair.  A 2^20-session batch is only
 * ~10 us of work, so separate launches live in their ramp-up and tail unless many streams overlap them; the ring
 * launch pays one ramp and one tail per round.  Up to 16 batches that share table, device, seed, kernel (thread per
 * session) and stream (ge_batch_set_stream), without phase regrouping or auto-reset.  Results are identical to
 * ge_step_many's. */
int ge_step_ring(ge_batch **batches, int n_batches, int n_rounds);
/* Same semantics with the state kept in registers for up to n_steps steps (one launch). */
int ge_run_fused(ge_batch *b, int n_steps, void *cuda_stream);
int ge_sync(ge_batch *b);

/* Canonical records (SPEC.md section 5), count * S bytes, sessions [first, first+count).
 * Replaces reading / writing AgentState.player_states, current_phase_id, phase_history length
 * (reference agent/game_agent_v2.py:97-117).  Synchronous. */
int ge_export_state(ge_batch *b, uint64_t first, uint64_t count, void *host_buf);
int ge_import_state(ge_batch *b, uint64_t first, uint64_t count, const void *host_buf);
/* Every import path (ge_import_state, ge_run_host, ge_run_host_async) range-checks its records ON THE DEVICE
 * (SPEC.md section 7b: phase / prev inside the table, step 0 only in phase 0, winner / player ids <= P, re-vote
 * counter <= max_revotes, no mask or flag bits above the player count, padding zero).  A record that fails is
 * replaced by the table's initial record (always safe to step) and the call — or, for the asynchronous call, the
 * next synchronising call on the batch (ge_sync, ge_export_state, ge_stats, ...) — returns GE_ERR_ARG naming the
 * first offending index. */

/* Step log for replay: applies n_steps single-step launches to the whole batch and writes the canonical records
 * of sessions [first, first+count) after each of them: host_records holds (n_steps + 1) frames of count * S
 * bytes, frame 0 = the state before the first step.  The host adapter turns a session's frames into the
 * reference's on-wire AgentState JSON, one object per graph run (reference src/lib/canvas/types.ts:338-360,
 * agent/game_agent_v2.py:97-117).  Synchronous. */
int ge_trace(ge_batch *b, uint64_t first, uint64_t count, int n_steps, void *host_records);

/* End-to-end call with HOST buffers: records_in (NULL = keep current device state) -> device,
 * n_steps steps, records_out (NULL = skip) and stats (NULL = skip) back to the host.  Synchronous.
 * Pinned buffers (ge_host_alloc) make the copies asynchronous DMA. */
int ge_run_host(ge_batch *b, const void *records_in, void *records_out, int n_steps, uint64_t *host_stats);
/* The same work enqueued on the batch's stream without the final synchronisation (host buffers must be pinned
 * and stay valid until ge_sync).  Several batches driven this way overlap H2D, compute and D2H. */
int ge_run_host_async(ge_batch *b, const void *records_in, void *records_out, int n_steps, uint64_t *host_stats);
/* on != 0: the host-buffer calls apply their n_steps in ONE fused launch (state in registers across the steps,
 * as ge_run_fused) instead of n_steps launches — the right mode for run-to-completion calls on small batches. */
int ge_batch_set_host_fused(ge_batch *b, int on);
int ge_host_alloc(void **p, size_t bytes);
void ge_host_free(void *p);

/* Audience masks: evaluates n_preds (<= 32) DNF predicates over the mask fields of SPEC.md section 2 for the
 * sessions [first, first+count); host_masks[i * n_preds + j] = lane mask (bit p-1 = player p) of predicate j in
 * session first+i.  Replaces the LLM reading declaration.audience_groups[*].selection_criteria
 * (reference games/werewolf-(mafia).yaml:138-165) when ActionExecutor fills the UI tools' audience_ids
 * (agent/game_agent_v2.py:1243-1568, src/lib/canvas/types.ts:14-17).  Synchronous. */
int ge_eval_preds(ge_batch *b, const ge_pred_t *preds, int n_preds, uint64_t first, uint64_t count, uint32_t *host_masks);

/* Statistics (SPEC.md section 6).  ge_stats recomputes the final-state histograms, then copies
 * GE_STATS_LEN words (n must be >= GE_STATS_LEN).  ge_stats_device_ptr returns the device buffer
 * (valid after ge_stats_refresh) for an NCCL / torch.distributed all-reduce.  No reference analogue. */
int ge_stats_refresh(ge_batch *b, void *cuda_stream);
int ge_stats(ge_batch *b, uint64_t *host_hist, size_t n);
void *ge_stats_device_ptr(ge_batch *b);
int ge_counted_steps(ge_batch *b, uint64_t *out);
/* Non-blocking: copies the counter as of this point of the batch's stream to page-locked host memory (ge_host_alloc);
 * the value is valid after the next ge_sync. */
int ge_counted_steps_async(ge_batch *b, uint64_t *pinned_out);

/* device pointer / size of the tiled session store (for profiling and tests) */
void *ge_state_device_ptr(ge_batch *b);
size_t ge_state_device_bytes(const ge_batch *b);
/* number of kernel launches (step + glue kernels) issued by this batch since creation */
uint64_t ge_launch_count(const ge_batch *b);

const char *ge_last_error(void);
const char *ge_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GAME_ENGINE_B200_H */
