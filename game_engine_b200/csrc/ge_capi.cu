// ge_capi.cu — C ABI of the batched referee/phase-step simulator (include/game_engine_b200.h).
//
// Host-side runtime in C++: table validation, session-store allocation, launch geometry and the
// glue kernels (init / import / export / statistics).  The step kernels live in ge_step_tps.cuh
// (thread per session) and ge_step_coop.cuh (lane per player).  No CPU fallback exists: every
// compute entry point needs a CUDA device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "ge_common.cuh"
#include "ge_step_tps.cuh"
#include "ge_step_coop.cuh"

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

extern "C" const char* ge_last_error(void) { return g_err.c_str(); }
extern "C" const char* ge_version(void) { return "game_engine_b200 0.1.0 (sm_100a)"; }

// ------------------------------------------------------------------------------------ handles
struct ge_table {
    DevTable dev;
    int family, P, bucket;       // bucket: werewolf P8 (8/16/24/32), TTL PB (4/8/16/32)
    size_t rec_canon, rec_dev;   // canonical / device record bytes
    uint32_t init_words[40];     // initial device record
};

typedef void (*step_fn)(const DevTable, const StepArgs);

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
    uint32_t* d_presence;         // 3 rotating phase-presence words (StepArgs::presence)
    uint32_t launch_idx;          // index of the next step launch
    uint32_t next_override;       // presence override for the next launch (0 = read the device word)
    int kernel;
    step_fn fn[3];               // by kernel id (COOP, TPS)
    int grid[3];
    uint64_t launches;
};

// ------------------------------------------------------------------------------------ glue kernels
struct InitRec { uint32_t w[40]; };

__device__ __forceinline__ uint32_t rt_tile_off(uint32_t o, uint32_t sl, uint32_t n16) {
    return (o / 16u < n16) ? (o / 16u) * 512u + sl * 16u + (o % 16u) : n16 * 512u + sl * 8u + (o - 16u * n16);
}

__global__ void k_init(uint8_t* tiles, uint64_t n_tiles, uint32_t S, const __grid_constant__ InitRec rec) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = n_tiles * 32;
    const uint32_t n16 = S / 16;
    for (uint64_t i = t; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint8_t* base = tiles + (i >> 5) * (uint64_t)(32 * S);
        const uint32_t sl = (uint32_t)(i & 31);
        for (uint32_t k = 0; k < S / 8; ++k)
            *reinterpret_cast<uint2*>(base + rt_tile_off(8 * k, sl, n16)) = make_uint2(rec.w[2 * k], rec.w[2 * k + 1]);
    }
}

// tiles -> canonical AoS records (count sessions starting at `first`)
__global__ void k_export(const uint8_t* tiles, uint32_t S_dev, uint32_t S_canon, uint64_t first, uint64_t count, uint8_t* out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n16 = S_dev / 16;
    for (uint64_t j = t; j < count; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = first + j;
        const uint8_t* base = tiles + (i >> 5) * (uint64_t)(32 * S_dev);
        const uint32_t sl = (uint32_t)(i & 31);
        for (uint32_t k = 0; k < S_canon / 8; ++k)
            *reinterpret_cast<uint2*>(out + j * S_canon + 8 * k) = *reinterpret_cast<const uint2*>(base + rt_tile_off(8 * k, sl, n16));
    }
}

__global__ void k_import(uint8_t* tiles, uint32_t S_dev, uint32_t S_canon, uint64_t first, uint64_t count, const uint8_t* in) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n16 = S_dev / 16;
    for (uint64_t j = t; j < count; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = first + j;
        uint8_t* base = tiles + (i >> 5) * (uint64_t)(32 * S_dev);
        const uint32_t sl = (uint32_t)(i & 31);
        for (uint32_t k = 0; k < S_dev / 8; ++k) {
            uint2 v = make_uint2(0, 0);
            if (k < S_canon / 8) v = *reinterpret_cast<const uint2*>(in + j * S_canon + 8 * k);
            *reinterpret_cast<uint2*>(base + rt_tile_off(8 * k, sl, n16)) = v;
        }
    }
}

// final-state histograms (SPEC.md section 6): winner, length, survivors / scores
__global__ void __launch_bounds__(256)
k_stats(const __grid_constant__ DevTable T, const uint8_t* tiles, uint32_t S_dev, uint64_t n, unsigned long long* stats) {
    __shared__ uint32_t sh[3 + 256 + 256];
    for (int i = threadIdx.x; i < 515; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint32_t n16 = S_dev / 16;
    const int P = T.h.n_players;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t* base = tiles + (i >> 5) * (uint64_t)(32 * S_dev);
        const uint32_t sl = (uint32_t)(i & 31);
        const uint4 c0 = *reinterpret_cast<const uint4*>(base + sl * 16);
        const bool terminal = T.phase[c0.x & 31].kind == KIND_TERMINAL;
        const uint32_t step = c0.x >> 16;
        if (T.h.family == FAM_WEREWOLF) {
            const uint32_t w = c0.y & 0xFF;
            atomicAdd(&sh[w <= 2 ? w : 0], 1u);
            if (terminal) atomicAdd(&sh[3 + 256 + __popc(c0.z)], 1u);
        } else {
            atomicAdd(&sh[terminal ? 1 : 0], 1u);
            if (terminal)
                for (int p = 0; p < P; ++p) {
                    const uint32_t pw = *reinterpret_cast<const uint32_t*>(base + rt_tile_off(8 + 4 * p, sl, n16));
                    atomicAdd(&sh[3 + 256 + (pw & 0xFF)], 1u);
                }
        }
        if (terminal) atomicAdd(&sh[3 + (step < 255 ? step : 255)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 515; i += blockDim.x) {
        const uint32_t v = sh[i];
        if (!v) continue;
        const int dst = i < 3 ? ST_WINNER + i : i < 259 ? ST_LENGTH + (i - 3) : ST_TAIL + (i - 259);
        atomicAdd(&stats[dst], (unsigned long long)v);
    }
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
    for (int i = 0; i < h.n_phases; ++i) {
        const ge_phase_t& ph = t->dev.phase[i];
        if (ph.kind > KIND_TERMINAL || ph.n_branches > 4) return fail(GE_ERR_ARG, "bad phase record");
        if (ph.kind != KIND_TERMINAL && ph.n_branches == 0) return fail(GE_ERR_ARG, "non-terminal phase without next_phase");
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
        }
    }
    // Column groups a step that STARTS in phase i must read besides column 0 (bit0 = dynamic masks C1:
    // fields 2..5; bit1 = role/team C2: fields 6..11; bit2 = per-player bytes).  A group is needed when a
    // predicate reads one of its fields or an effect read-modify-writes it; effects that overwrite a
    // whole group (ASSIGN_ROLES -> C2, NIGHT_RESET -> player bytes) need no read.
    auto pred_need = [&](int pi) -> uint8_t {
        if (pi < 0 || pi >= h.n_preds) return 0;
        const ge_pred_t& p = t->dev.pred[pi];
        uint8_t out = 0;
        const uint16_t lits[4] = {p.pos0, p.neg0, p.pos1, p.neg1};
        for (int c = 0; c < 2; ++c) {
            if (lits[2 * c + 1] & 0x8000u) continue;          // clause marked empty (& ~ALL)
            const uint16_t used = lits[2 * c] | lits[2 * c + 1];
            if (used & 0x003Cu) out |= 1;
            if (used & 0x0FC0u) out |= 2;
        }
        return out;
    };
    auto entry_need = [&](int en) -> uint8_t { return (en == EN_ASSIGN_ROLES || en == EN_NIGHT_RESET) ? 1 : 0; };
    for (int i = 0; i < h.n_phases; ++i) {
        const ge_phase_t& ph = t->dev.phase[i];
        uint8_t need = 0;
        if (h.family != FAM_WEREWOLF) { t->dev.need[i] = 7; continue; }
        if (ph.kind == KIND_TERMINAL) { t->dev.need[i] = 0; continue; }
        if (ph.kind == KIND_ACTION) {
            need |= pred_need(ph.actor_pred);
            if (ph.action_op == ACT_PICK_PLAYER) need |= pred_need(ph.action_arg);
            if (ph.exit_op >= EX_VOTE_KILL && ph.exit_op <= EX_DAY_VOTE) need |= 1 | 4;
        }
        for (int b = 0; b < ph.n_branches; ++b) {
            const ge_branch_t& br = ph.br[b];
            if (br.op == BR_COUNT_EQ0 || br.op == BR_COUNT_GE) need |= pred_need(br.a);
            if (br.op == BR_COUNT_GE) need |= pred_need((int)br.arg);
            need |= entry_need(t->dev.phase[br.next].entry_op);
        }
        t->dev.need[i] = need;
    }
    if (t->dev.phase[0].id != 0) return fail(GE_ERR_ARG, "phase index 0 must be DSL phase 0");
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

extern "C" int ge_table_create(const uint8_t* blob, size_t n, ge_table** out) {
    if (!out) return fail(GE_ERR_ARG, "out is NULL");
    ge_table* t = new (std::nothrow) ge_table;
    if (!t) return fail(GE_ERR_NOMEM, "out of host memory");
    const int rc = validate_and_build(blob, n, t);
    if (rc != GE_OK) { delete t; return rc; }
    *out = t;
    return GE_OK;
}
extern "C" void ge_table_destroy(ge_table* t) { delete t; }
extern "C" size_t ge_table_record_size(const ge_table* t) { return t ? t->rec_canon : 0; }
extern "C" int ge_table_n_players(const ge_table* t) { return t ? t->P : 0; }

// ------------------------------------------------------------------------------------ dispatch
static step_fn pick_fn(const ge_table* t, int kernel) {
    if (t->family == FAM_WEREWOLF) {
        switch (t->bucket) {
        case 8: return kernel == GE_KERNEL_COOP ? (step_fn)k_step_w_coop<8> : (step_fn)k_step_w_tps<8>;
        case 16: return kernel == GE_KERNEL_COOP ? (step_fn)k_step_w_coop<16> : (step_fn)k_step_w_tps<16>;
        case 24: return kernel == GE_KERNEL_COOP ? (step_fn)k_step_w_coop<24> : (step_fn)k_step_w_tps<24>;
        case 32: return kernel == GE_KERNEL_COOP ? (step_fn)k_step_w_coop<32> : (step_fn)k_step_w_tps<32>;
        }
    } else {
        switch (t->bucket) {
        case 4: return kernel == GE_KERNEL_COOP ? (step_fn)k_step_t_coop<4> : (step_fn)k_step_t_tps<4>;
        case 8: return kernel == GE_KERNEL_COOP ? (step_fn)k_step_t_coop<8> : (step_fn)k_step_t_tps<8>;
        case 16: return kernel == GE_KERNEL_COOP ? (step_fn)k_step_t_coop<16> : (step_fn)k_step_t_tps<16>;
        case 32: return kernel == GE_KERNEL_COOP ? (step_fn)k_step_t_coop<32> : (step_fn)k_step_t_tps<32>;
        }
    }
    return nullptr;
}

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
static int init_sessions(ge_batch* b, uint64_t first_session_id, uint64_t seed) {
    b->first_sid = first_session_id;
    b->seed = seed;
    InitRec rec;
    memcpy(rec.w, b->tab->init_words, sizeof rec.w);
    k_init<<<glue_grid(b, b->n_tiles * 32, 256), 256, 0, b->stream>>>(b->d_tiles, b->n_tiles, (uint32_t)b->tab->rec_dev, rec);
    CU(cudaGetLastError());
    b->launches++;
    CU(cudaMemsetAsync(b->d_presence, 0, 3 * sizeof(uint32_t), b->stream));
    b->next_override = 1u;        // every session is in phase index 0
    return GE_OK;
}

// Statistics are cumulative over the life of the handle: before the sessions are overwritten their
// final-state histograms are folded into the accumulator ("harvest"), so win rates cover every session
// the batch ever simulated.
extern "C" int ge_batch_reset(ge_batch* b, uint64_t first_session_id, uint64_t seed) {
    if (!b) return fail(GE_ERR_ARG, "batch is NULL");
    CU(cudaSetDevice(b->device));
    k_stats<<<glue_grid(b, b->n, 256), 256, 0, b->stream>>>(b->tab->dev, b->d_tiles, (uint32_t)b->tab->rec_dev, b->n, b->d_stats);
    CU(cudaGetLastError());
    b->launches++;
    return init_sessions(b, first_session_id, seed);
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
    b->tiles_bytes = (size_t)b->n_tiles * 32 * t->rec_dev;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete b; return fail(GE_ERR_CUDA, cudaGetErrorString(e)); }
    b->sm_count = prop.multiProcessorCount;
    e = cudaMalloc(&b->d_tiles, b->tiles_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&b->d_stats, GE_STATS_LEN * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_stats_out, GE_STATS_LEN * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_presence, 3 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&b->own_stream, cudaStreamNonBlocking);
    b->stream = b->own_stream;
    if (e != cudaSuccess) {
        cudaFree(b->d_tiles); cudaFree(b->d_stats); cudaFree(b->d_stats_out); cudaFree(b->d_presence); delete b;
        return fail(e == cudaErrorMemoryAllocation ? GE_ERR_NOMEM : GE_ERR_CUDA, std::string("ge_batch_create: ") + cudaGetErrorString(e));
    }
    for (int k = GE_KERNEL_COOP; k <= GE_KERNEL_TPS; ++k) {
        b->fn[k] = pick_fn(t, k);
        if (!b->fn[k]) { ge_batch_destroy(b); return fail(GE_ERR_UNSUPPORTED, "no kernel for this table"); }
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)b->fn[k], 128, 0);
        if (e != cudaSuccess || per_sm < 1) per_sm = 4;
        const uint64_t warps = k == GE_KERNEL_COOP ? b->n_tiles * (uint64_t)lanes_per_session(t) : b->n_tiles;
        uint64_t g = (warps + 3) / 4;
        const uint64_t cap = (uint64_t)b->sm_count * per_sm;      // persistent grid: whole multiples of the SM count
        if (g > cap) g = cap;
        b->grid[k] = g < 1 ? 1 : (int)g;
    }
    b->kernel = GE_KERNEL_TPS;
    int rc = ge_batch_clear_stats(b);
    if (rc == GE_OK) rc = init_sessions(b, first_session_id, seed);
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
    delete b;
}

extern "C" int ge_batch_set_stream(ge_batch* b, void* cuda_stream) {
    if (!b) return fail(GE_ERR_ARG, "batch is NULL");
    CU(cudaSetDevice(b->device));
    CU(cudaStreamSynchronize(b->stream));
    b->stream = cuda_stream ? (cudaStream_t)cuda_stream : b->own_stream;
    return GE_OK;
}

extern "C" int ge_batch_set_kernel(ge_batch* b, int kernel) {
    if (!b || kernel < GE_KERNEL_AUTO || kernel > GE_KERNEL_TPS) return fail(GE_ERR_ARG, "bad kernel id");
    b->kernel = kernel == GE_KERNEL_AUTO ? GE_KERNEL_TPS : kernel;
    return GE_OK;
}
extern "C" int ge_batch_get_kernel(const ge_batch* b) { return b ? b->kernel : GE_ERR_ARG; }

static int launch_steps(ge_batch* b, int n_launches, int steps_per_launch, cudaStream_t st) {
    const step_fn fn = b->fn[b->kernel];
    StepArgs a;
    a.tiles = b->d_tiles; a.n_sessions = b->n; a.n_tiles = b->n_tiles; a.first_sid = b->first_sid; a.seed = b->seed;
    a.stats = b->d_stats; a.presence = b->d_presence; a.n_steps = steps_per_launch;
    for (int i = 0; i < n_launches; ++i) {
        a.launch_idx = b->launch_idx++;
        a.presence_override = b->next_override;
        b->next_override = 0;
        fn<<<b->grid[b->kernel], 128, 0, st>>>(b->tab->dev, a);
        b->launches++;
    }
    CU(cudaGetLastError());
    return GE_OK;
}

extern "C" int ge_step(ge_batch* b, int n_steps, void* cuda_stream) {
    if (!b || n_steps < 0) return fail(GE_ERR_ARG, "bad arguments to ge_step");
    CU(cudaSetDevice(b->device));
    return launch_steps(b, n_steps, 1, cuda_stream ? (cudaStream_t)cuda_stream : b->stream);
}

extern "C" int ge_run_fused(ge_batch* b, int n_steps, void* cuda_stream) {
    if (!b || n_steps < 0) return fail(GE_ERR_ARG, "bad arguments to ge_run_fused");
    if (n_steps == 0) return GE_OK;
    CU(cudaSetDevice(b->device));
    return launch_steps(b, 1, n_steps, cuda_stream ? (cudaStream_t)cuda_stream : b->stream);
}

extern "C" int ge_sync(ge_batch* b) {
    if (!b) return fail(GE_ERR_ARG, "batch is NULL");
    CU(cudaSetDevice(b->device));
    CU(cudaStreamSynchronize(b->stream));
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

static int export_async(ge_batch* b, uint64_t first, uint64_t count, void* host_buf) {
    const size_t S = b->tab->rec_canon;
    int rc = ensure_stage(b, count * S);
    if (rc) return rc;
    k_export<<<glue_grid(b, count, 256), 256, 0, b->stream>>>(b->d_tiles, (uint32_t)b->tab->rec_dev, (uint32_t)S, first, count, b->d_stage);
    CU(cudaGetLastError());
    b->launches++;
    CU(cudaMemcpyAsync(host_buf, b->d_stage, count * S, cudaMemcpyDeviceToHost, b->stream));
    return GE_OK;
}

static int import_async(ge_batch* b, uint64_t first, uint64_t count, const void* host_buf) {
    const size_t S = b->tab->rec_canon;
    int rc = ensure_stage(b, count * S);
    if (rc) return rc;
    CU(cudaMemcpyAsync(b->d_stage, host_buf, count * S, cudaMemcpyHostToDevice, b->stream));
    k_import<<<glue_grid(b, count, 256), 256, 0, b->stream>>>(b->d_tiles, (uint32_t)b->tab->rec_dev, (uint32_t)S, first, count, b->d_stage);
    CU(cudaGetLastError());
    b->launches++;
    CU(cudaMemsetAsync(b->d_presence, 0, 3 * sizeof(uint32_t), b->stream));
    b->next_override = 0xFFFFFFFFu;   // imported sessions can be in any phase
    return GE_OK;
}

extern "C" int ge_export_state(ge_batch* b, uint64_t first, uint64_t count, void* host_buf) {
    if (!b || !host_buf || first + count > b->n) return fail(GE_ERR_ARG, "bad arguments to ge_export_state");
    if (count == 0) return GE_OK;
    CU(cudaSetDevice(b->device));
    int rc = export_async(b, first, count, host_buf);
    if (rc) return rc;
    CU(cudaStreamSynchronize(b->stream));
    return GE_OK;
}

extern "C" int ge_import_state(ge_batch* b, uint64_t first, uint64_t count, const void* host_buf) {
    if (!b || !host_buf || first + count > b->n) return fail(GE_ERR_ARG, "bad arguments to ge_import_state");
    if (count == 0) return GE_OK;
    // reject records whose phase index is outside the table: the kernels index the table with it
    const size_t S = b->tab->rec_canon;
    const uint8_t* p = static_cast<const uint8_t*>(host_buf);
    for (uint64_t i = 0; i < count; ++i)
        if (p[i * S] >= b->tab->dev.h.n_phases || p[i * S + 1] >= b->tab->dev.h.n_phases)
            return fail(GE_ERR_ARG, "record has a phase index outside the table");
    CU(cudaSetDevice(b->device));
    int rc = import_async(b, first, count, host_buf);
    if (rc) return rc;
    CU(cudaStreamSynchronize(b->stream));
    return GE_OK;
}

extern "C" int ge_stats_refresh(ge_batch* b, void* cuda_stream) {
    if (!b) return fail(GE_ERR_ARG, "batch is NULL");
    CU(cudaSetDevice(b->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : b->stream;
    // snapshot = accumulator + histograms of the sessions currently resident
    CU(cudaMemcpyAsync(b->d_stats_out, b->d_stats, GE_STATS_LEN * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    k_stats<<<glue_grid(b, b->n, 256), 256, 0, st>>>(b->tab->dev, b->d_tiles, (uint32_t)b->tab->rec_dev, b->n, b->d_stats_out);
    CU(cudaGetLastError());
    b->launches++;
    return GE_OK;
}

extern "C" int ge_stats(ge_batch* b, uint64_t* host_hist, size_t n) {
    if (!b || !host_hist || n < GE_STATS_LEN) return fail(GE_ERR_ARG, "bad arguments to ge_stats");
    int rc = ge_stats_refresh(b, nullptr);
    if (rc) return rc;
    CU(cudaMemcpyAsync(host_hist, b->d_stats_out, GE_STATS_LEN * sizeof(uint64_t), cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return GE_OK;
}

extern "C" void* ge_stats_device_ptr(ge_batch* b) { return b ? (void*)b->d_stats_out : nullptr; }

extern "C" int ge_counted_steps(ge_batch* b, uint64_t* out) {
    if (!b || !out) return fail(GE_ERR_ARG, "bad arguments to ge_counted_steps");
    CU(cudaSetDevice(b->device));
    CU(cudaMemcpyAsync(out, b->d_stats, sizeof(uint64_t), cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return GE_OK;
}

extern "C" int ge_run_host(ge_batch* b, const void* records_in, void* records_out, int n_steps, uint64_t* host_stats) {
    if (!b || n_steps < 0) return fail(GE_ERR_ARG, "bad arguments to ge_run_host");
    CU(cudaSetDevice(b->device));
    int rc;
    if (records_in && (rc = import_async(b, 0, b->n, records_in)) != GE_OK) return rc;
    if ((rc = launch_steps(b, n_steps, 1, b->stream)) != GE_OK) return rc;
    if (records_out && (rc = export_async(b, 0, b->n, records_out)) != GE_OK) return rc;
    if (host_stats) {
        if ((rc = ge_stats_refresh(b, nullptr)) != GE_OK) return rc;
        CU(cudaMemcpyAsync(host_stats, b->d_stats_out, GE_STATS_LEN * sizeof(uint64_t), cudaMemcpyDeviceToHost, b->stream));
    }
    CU(cudaStreamSynchronize(b->stream));
    return GE_OK;
}

extern "C" int ge_host_alloc(void** p, size_t bytes) {
    if (!p) return fail(GE_ERR_ARG, "p is NULL");
    CU(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
    return GE_OK;
}
extern "C" void ge_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" void* ge_state_device_ptr(ge_batch* b) { return b ? (void*)b->d_tiles : nullptr; }
extern "C" size_t ge_state_device_bytes(const ge_batch* b) { return b ? b->tiles_bytes : 0; }
extern "C" uint64_t ge_launch_count(const ge_batch* b) { return b ? b->launches : 0; }
