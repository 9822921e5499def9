// ge_glue.cuh — the kernels around the step kernels: initial records, canonical AoS <-> tiled store, final-state
// statistics, audience masks, active-prefix compaction, phase regrouping, device-side auto-reset, and the
// measurement spin kernel.  All of them are thread-per-session (or per tile / per swapped pair) and HBM-bound;
// the host runtime that launches them is ge_capi.cu.
#pragma once
#include "ge_common.cuh"

using namespace ge;

// ------------------------------------------------------------------------------------ glue kernels
struct InitRec { uint32_t w[40]; };

__device__ __forceinline__ uint32_t rt_tile_off(uint32_t o, uint32_t sl, uint32_t n16) {
    return (o / 16u < n16) ? (o / 16u) * 512u + sl * 16u + (o % 16u) : n16 * 512u + sl * 8u + (o - 16u * n16);
}

__global__ void k_set_u64(unsigned long long* p, unsigned long long v) { *p = v; }

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

// Bookkeeping an import resets, done by the import kernel's first thread instead of separate stream operations (an
// end-to-end call is a handful of short operations per sub-batch; each one saved shows): the phase-presence words
// (the next step launch starts from an override) and, for a whole-batch import, the compaction state (every slot
// active again, slot order = session order; the device epoch of auto-reset, word 8, survives).
struct ImportReset {
    uint32_t* presence;               // NULL = leave alone
    unsigned long long* cstate;       // NULL = leave alone
    unsigned long long n, epoch;
};
__device__ __forceinline__ void import_reset(const ImportReset& R) {
    if (R.presence) { R.presence[0] = 0; R.presence[1] = 0; R.presence[2] = 0; }
    if (R.cstate)
        for (int i = 0; i < 16; ++i)
            if (i != 8) R.cstate[i] = i == 0 ? R.n : i == 7 ? R.epoch : 0ull;
}

// Record validation (SPEC.md section 7b), shared by every import path: the step kernels index the table with the
// phase bytes and shift by player ids, so a record must be range-checked before it reaches them.  0 = well-formed.
__device__ __forceinline__ int record_invalid(const DevTable& T, const uint8_t* r, uint32_t S_canon) {
    const int P = T.h.n_players;
    const uint32_t hi = P >= 32 ? 0u : ~((1u << P) - 1u);
    const uint32_t step = r[2] | ((uint32_t)r[3] << 8);
    if (r[0] >= T.h.n_phases || r[1] >= T.h.n_phases) return 1;
    if (step == 0 && (r[0] != 0 || r[1] != 0)) return 2;
    if (T.h.family == FAM_WEREWOLF) {
        if (r[4] > 2 || r[5] > P || r[6] > P) return 3;
        if ((r[7] & 0x7Fu) > T.h.max_revotes || (T.h.max_revotes == 0 && r[7] != 0)) return 4;
        for (int f = 0; f < 10; ++f) {
            const uint32_t m = r[8 + 4 * f] | ((uint32_t)r[9 + 4 * f] << 8) | ((uint32_t)r[10 + 4 * f] << 16) | ((uint32_t)r[11 + 4 * f] << 24);
            if (m & hi) return 5;
        }
        for (uint32_t p = 0; 48 + p < S_canon; ++p)
            if (r[48 + p] > ((int)p < P ? P : 0)) return 6;
    } else {
        if (r[4] > P || r[6] > P || r[7] != 0) return 3;
        for (uint32_t p = 0; 8 + 4 * p < S_canon; ++p) {
            if ((int)p < P) { if (r[11 + 4 * p] & 0xE0u) return 5; }
            else if (r[8 + 4 * p] | r[9 + 4 * p] | r[10 + 4 * p] | r[11 + 4 * p]) return 6;
        }
    }
    return 0;
}

// canonical AoS records -> tiles.  A record that fails validation is replaced by the table's initial record (always
// safe to step) and reported: err[0] counts such records, err[1] = max(~index) (so ~err[1] is the first one).
__global__ void k_import(const __grid_constant__ DevTable T, const __grid_constant__ InitRec init, uint8_t* tiles, uint32_t S_dev, uint32_t S_canon,
                         uint64_t first, uint64_t count, const uint8_t* in, uint32_t* err, ImportReset R) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) import_reset(R);
    const uint32_t n16 = S_dev / 16;
    for (uint64_t j = t; j < count; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = first + j;
        uint8_t* base = tiles + (i >> 5) * (uint64_t)(32 * S_dev);
        const uint32_t sl = (uint32_t)(i & 31);
        const uint8_t* r = in + j * S_canon;
        const bool bad = err != nullptr && record_invalid(T, r, S_canon) != 0;
        if (bad) { atomicAdd_system(&err[0], 1u); atomicMax_system(&err[1], ~(uint32_t)(j > 0xFFFFFFFEull ? 0xFFFFFFFEull : j)); }
        for (uint32_t k = 0; k < S_dev / 8; ++k) {
            uint2 v = make_uint2(0, 0);
            if (bad) v = make_uint2(init.w[2 * k], init.w[2 * k + 1]);
            else if (k < S_canon / 8) v = *reinterpret_cast<const uint2*>(r + 8 * k);
            *reinterpret_cast<uint2*>(base + rt_tile_off(8 * k, sl, n16)) = v;
        }
    }
}

// ------------------------------------------------------------------------------------ dense wire format
// Werewolf tables with P <= 16 players: the ten lane masks of the canonical record carry P8 bits each but occupy
// 32, so at the host boundary (PCIe) a record can travel in 32 bytes (P8 = 8) or 48 bytes (P8 = 16) instead of
// 56 / 64 (SPEC.md section 5b).  Words (little-endian u32):
//   P8 = 8 : 0-1 header (as canonical) | 2 alive,can_vote,eligible,submitted (u8 each) | 3 revealed,investigated,wolf,secret
//            | 4 role_lo,role_hi,0,0 | 5-6 selected_target_id[0..7] | 7 zero
//   P8 = 16: 0-1 header | 2 alive,can_vote (u16 each) | 3 eligible,submitted | 4 revealed,investigated | 5 wolf,secret
//            | 6 role_lo,role_hi | 7 zero | 8-11 selected_target_id[0..15]
// The conversion happens in the import / export kernels.  The session store in HBM keeps the canonical columns, or —
// GE_OPT_STORE_PACKED — exactly these 32 / 48 bytes in two / three 16-byte columns (the packed store).
template <int P8> struct DenseW { static constexpr int WORDS = P8 == 8 ? 8 : 12; static constexpr int CW = 12 + P8 / 4; };

template <int P8>
__device__ __forceinline__ void dense_pack(const uint32_t (&w)[12 + P8 / 4], uint32_t (&d)[DenseW<P8>::WORDS]) {
    d[0] = w[0]; d[1] = w[1];
    if (P8 == 8) {
        d[2] = (w[2] & 0xFFu) | ((w[3] & 0xFFu) << 8) | ((w[4] & 0xFFu) << 16) | ((w[5] & 0xFFu) << 24);
        d[3] = (w[6] & 0xFFu) | ((w[7] & 0xFFu) << 8) | ((w[8] & 0xFFu) << 16) | ((w[9] & 0xFFu) << 24);
        d[4] = (w[10] & 0xFFu) | ((w[11] & 0xFFu) << 8);
        d[5] = w[12]; d[6] = w[13]; d[7] = 0;
    } else {
#pragma unroll
        for (int k = 0; k < 5; ++k) d[2 + k] = (w[2 + 2 * k] & 0xFFFFu) | ((w[3 + 2 * k] & 0xFFFFu) << 16);
        d[7] = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) d[8 + k] = w[12 + k];
    }
}
// returns non-zero when the padding of the dense record is not zero (then the record is malformed)
template <int P8>
__device__ __forceinline__ uint32_t dense_unpack(const uint32_t (&d)[DenseW<P8>::WORDS], uint32_t (&w)[12 + P8 / 4]) {
    w[0] = d[0]; w[1] = d[1];
    if (P8 == 8) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { w[2 + k] = (d[2] >> (8 * k)) & 0xFFu; w[6 + k] = (d[3] >> (8 * k)) & 0xFFu; }
        w[10] = d[4] & 0xFFu; w[11] = (d[4] >> 8) & 0xFFu;
        w[12] = d[5]; w[13] = d[6];
        return (d[4] >> 16) | d[7];
    } else {
#pragma unroll
        for (int k = 0; k < 5; ++k) { w[2 + 2 * k] = d[2 + k] & 0xFFFFu; w[3 + 2 * k] = d[2 + k] >> 16; }
#pragma unroll
        for (int k = 0; k < 4; ++k) w[12 + k] = d[8 + k];
        return d[7];
    }
}

// record_invalid on the canonical WORDS of a werewolf record (same rules; the two are held to each other by the tests)
template <int P8>
__device__ __forceinline__ int words_invalid_w(const DevTable& T, const uint32_t (&w)[12 + P8 / 4]) {
    const int P = T.h.n_players;
    const uint32_t hi = P >= 32 ? 0u : ~((1u << P) - 1u);
    const uint32_t ph = w[0] & 0xFFu, pv = (w[0] >> 8) & 0xFFu, step = w[0] >> 16;
    if (ph >= T.h.n_phases || pv >= T.h.n_phases) return 1;
    if (step == 0 && (ph != 0 || pv != 0)) return 2;
    if ((w[1] & 0xFFu) > 2u || ((w[1] >> 8) & 0xFFu) > (uint32_t)P || ((w[1] >> 16) & 0xFFu) > (uint32_t)P) return 3;
    const uint32_t rv = w[1] >> 24;
    if ((rv & 0x7Fu) > T.h.max_revotes || (T.h.max_revotes == 0 && rv != 0)) return 4;
    uint32_t any = 0;
#pragma unroll
    for (int f = 0; f < 10; ++f) any |= w[2 + f];
    if (any & hi) return 5;
#pragma unroll
    for (int p = 0; p < P8; ++p)
        if (((w[12 + (p >> 2)] >> (8 * (p & 3))) & 0xFFu) > (uint32_t)(p < P ? P : 0)) return 6;
    return 0;
}

// canonical words of one slot of the tiled store (werewolf) and back.  PK: the slot lives in the PACKED store (two
// 16-byte columns holding the dense wire record, ge_step_tps.cuh), otherwise in the canonical columns.
template <int P8, bool PK = false>
__device__ __forceinline__ void tile_load_words(const uint8_t* tiles, uint64_t slot, uint32_t (&w)[12 + P8 / 4]) {
    const uint32_t sl = (uint32_t)(slot & 31);
    if constexpr (PK) {
        static_assert(P8 == 8 || P8 == 16, "packed store: up to 16 players");
        constexpr int DW = DenseW<P8>::WORDS;
        const uint8_t* base = tiles + (slot >> 5) * (uint64_t)(32 * 4 * DW);
        uint32_t d[DW];
#pragma unroll
        for (int c = 0; c < DW / 4; ++c) {
            const uint4 v = *reinterpret_cast<const uint4*>(base + c * 512 + sl * 16);
            d[4 * c] = v.x; d[4 * c + 1] = v.y; d[4 * c + 2] = v.z; d[4 * c + 3] = v.w;
        }
        dense_unpack<P8>(d, w);
    } else {
        constexpr int S = 48 + P8;
        const uint8_t* base = tiles + (slot >> 5) * (uint64_t)(32 * S);
#pragma unroll
        for (int c = 0; c < S / 16; ++c) {
            const uint4 v = *reinterpret_cast<const uint4*>(base + c * 512 + sl * 16);
            w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
        }
        if (S % 16) {
            const uint2 v = *reinterpret_cast<const uint2*>(base + (S / 16) * 512 + sl * 8);
            w[4 * (S / 16)] = v.x; w[4 * (S / 16) + 1] = v.y;
        }
    }
}
template <int P8, bool PK = false>
__device__ __forceinline__ void tile_store_words(uint8_t* tiles, uint64_t slot, const uint32_t (&w)[12 + P8 / 4]) {
    const uint32_t sl = (uint32_t)(slot & 31);
    if constexpr (PK) {
        constexpr int DW = DenseW<P8>::WORDS;
        uint8_t* base = tiles + (slot >> 5) * (uint64_t)(32 * 4 * DW);
        uint32_t d[DW];
        dense_pack<P8>(w, d);
#pragma unroll
        for (int c = 0; c < DW / 4; ++c)
            *reinterpret_cast<uint4*>(base + c * 512 + sl * 16) = make_uint4(d[4 * c], d[4 * c + 1], d[4 * c + 2], d[4 * c + 3]);
    } else {
        constexpr int S = 48 + P8;
        uint8_t* base = tiles + (slot >> 5) * (uint64_t)(32 * S);
#pragma unroll
        for (int c = 0; c < S / 16; ++c)
            *reinterpret_cast<uint4*>(base + c * 512 + sl * 16) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        if (S % 16)
            *reinterpret_cast<uint2*>(base + (S / 16) * 512 + sl * 8) = make_uint2(w[4 * (S / 16)], w[4 * (S / 16) + 1]);
    }
}

// AoS records of one wire format -> canonical words and back (WD: dense wire record, else the canonical record).
// wire_load returns non-zero when the padding of a dense record is not zero.
template <int P8, bool WD>
__device__ __forceinline__ uint32_t wire_load(const uint8_t* in, uint64_t j, uint32_t (&w)[12 + P8 / 4]) {
    if constexpr (WD) {
        constexpr int DW = DenseW<P8>::WORDS;
        uint32_t d[DW];
        const uint4* src = reinterpret_cast<const uint4*>(in + j * (4 * DW));
#pragma unroll
        for (int k = 0; k < DW / 4; ++k) { const uint4 v = src[k]; d[4 * k] = v.x; d[4 * k + 1] = v.y; d[4 * k + 2] = v.z; d[4 * k + 3] = v.w; }
        return dense_unpack<P8>(d, w);
    } else {
        constexpr int CW = 12 + P8 / 4;                      // canonical records are 8-byte aligned (S % 8 == 0)
        const uint2* src = reinterpret_cast<const uint2*>(in + j * (4 * CW));
#pragma unroll
        for (int k = 0; k < CW / 2; ++k) { const uint2 v = src[k]; w[2 * k] = v.x; w[2 * k + 1] = v.y; }
        return 0u;
    }
}
template <int P8, bool WD>
__device__ __forceinline__ void wire_store(uint8_t* out, uint64_t j, const uint32_t (&w)[12 + P8 / 4]) {
    if constexpr (WD) {
        constexpr int DW = DenseW<P8>::WORDS;
        uint32_t d[DW];
        dense_pack<P8>(w, d);
        uint4* dst = reinterpret_cast<uint4*>(out + j * (4 * DW));
#pragma unroll
        for (int k = 0; k < DW / 4; ++k) dst[k] = make_uint4(d[4 * k], d[4 * k + 1], d[4 * k + 2], d[4 * k + 3]);
    } else {
        constexpr int CW = 12 + P8 / 4;
        uint2* dst = reinterpret_cast<uint2*>(out + j * (4 * CW));
#pragma unroll
        for (int k = 0; k < CW / 2; ++k) dst[k] = make_uint2(w[2 * k], w[2 * k + 1]);
    }
}

// AoS records -> tiles for the werewolf family, any wire format (WD) and any store format (PK), validated like k_import
// (err == NULL: trusted records, no validation).  One thread per session, wide accesses on both sides.  `init` holds
// the CANONICAL words of the initial record.
template <int P8, bool WD, bool PK>
__global__ void __launch_bounds__(256)
k_import_w(const __grid_constant__ DevTable T, const __grid_constant__ InitRec init, uint8_t* tiles, uint64_t first, uint64_t count,
           const uint8_t* in, uint32_t* err, ImportReset R) {
    if (blockIdx.x == 0 && threadIdx.x == 0) import_reset(R);
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t w[12 + P8 / 4];
        const uint32_t pad = wire_load<P8, WD>(in, j, w);
        if (err != nullptr && (pad != 0 || words_invalid_w<P8>(T, w) != 0)) {
            atomicAdd_system(&err[0], 1u);
            atomicMax_system(&err[1], ~(uint32_t)(j > 0xFFFFFFFEull ? 0xFFFFFFFEull : j));
#pragma unroll
            for (int k = 0; k < 12 + P8 / 4; ++k) w[k] = init.w[k];
        }
        tile_store_words<P8, PK>(tiles, first + j, w);
    }
}

// tiles -> AoS records; origin != NULL: slots are permuted (compaction), records leave in original order
template <int P8, bool WD, bool PK>
__global__ void __launch_bounds__(256)
k_export_w(const uint8_t* tiles, const uint32_t* __restrict__ origin, uint64_t n, uint64_t first, uint64_t count, uint8_t* out) {
    const uint64_t lo = origin ? 0 : first, hi = origin ? n : first + count;
    for (uint64_t slot = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; slot < hi; slot += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t o = origin ? origin[slot] : slot;
        if (o < first || o >= first + count) continue;
        uint32_t w[12 + P8 / 4];
        tile_load_words<P8, PK>(tiles, slot, w);
        wire_store<P8, WD>(out, o - first, w);
    }
}

// canonical columns <-> packed store, slot by slot (slot order, and with it the origin map of compaction, is kept)
template <int P8>
__global__ void __launch_bounds__(256)
k_repack(const uint8_t* src, uint8_t* dst, uint64_t slots, int to_packed) {
    for (uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; slot < slots; slot += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t w[12 + P8 / 4];
        if (to_packed) { tile_load_words<P8, false>(src, slot, w); tile_store_words<P8, true>(dst, slot, w); }
        else { tile_load_words<P8, true>(src, slot, w); tile_store_words<P8, false>(dst, slot, w); }
    }
}

// final-state histograms (SPEC.md section 6): winner, length, survivors / scores.  sh = 515 shared counters.
// alive_mask: where is_alive sits in word 2 (canonical: the whole word; packed store: its low byte / half)
__device__ __forceinline__ void stats_one(const DevTable& T, const uint8_t* tiles, uint32_t S_dev, uint64_t i, uint32_t* sh, uint32_t alive_mask) {
    const uint32_t n16 = S_dev / 16;
    const int P = T.h.n_players;
    const uint8_t* base = tiles + (i >> 5) * (uint64_t)(32 * S_dev);
    const uint32_t sl = (uint32_t)(i & 31);
    const uint4 c0 = *reinterpret_cast<const uint4*>(base + sl * 16);
    const bool terminal = T.phase[c0.x & 31].kind == KIND_TERMINAL;
    const uint32_t step = c0.x >> 16;
    if (T.h.family == FAM_WEREWOLF) {
        const uint32_t w = c0.y & 0xFF;
        atomicAdd(&sh[w <= 2 ? w : 0], 1u);
        if (terminal) atomicAdd(&sh[3 + 256 + __popc(c0.z & alive_mask)], 1u);
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
__device__ __forceinline__ void stats_flush(const uint32_t* sh, unsigned long long* stats) {
    for (int i = threadIdx.x; i < 515; i += blockDim.x) {
        const uint32_t v = sh[i];
        if (!v) continue;
        const int dst = i < 3 ? ST_WINNER + i : i < 259 ? ST_LENGTH + (i - 3) : ST_TAIL + (i - 259);
        atomicAdd(&stats[dst], (unsigned long long)v);
    }
}

__global__ void __launch_bounds__(256)
k_stats(const __grid_constant__ DevTable T, const uint8_t* tiles, uint32_t S_dev, uint64_t n, unsigned long long* stats, uint32_t alive_mask) {
    __shared__ uint32_t sh[3 + 256 + 256];
    for (int i = threadIdx.x; i < 515; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        stats_one(T, tiles, S_dev, i, sh, alive_mask);
    __syncthreads();
    stats_flush(sh, stats);
}

// Auto-reset (ge_batch_set_autoreset): when the check that follows a counted step finds no live session left
// (n_active == 0), the batch starts over ON THE DEVICE with fresh session ids, no host round trip:
// k_autoreset_apply folds the finished sessions' histograms into the accumulator and rewrites every slot with the
// initial record (slot order back to identity); k_autoreset_commit then republishes n_active, bumps the device
// epoch (session id = first_sid + epoch * sid_stride + index) and announces "phase 0" to the next launch.
// Both are no-ops (one uniform load) while games are still running.
__global__ void __launch_bounds__(256)
k_autoreset_apply(const __grid_constant__ DevTable T, uint8_t* tiles, uint32_t S_dev, uint64_t n, uint64_t n_tiles,
                  const __grid_constant__ InitRec rec, uint32_t* origin, unsigned long long* stats, const unsigned long long* cstate, uint32_t alive_mask) {
    __shared__ uint32_t sh[3 + 256 + 256];
    if (cstate[0] != 0) return;
    for (int i = threadIdx.x; i < 515; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint32_t n16 = S_dev / 16;
    const uint64_t total = n_tiles * 32;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        if (i < n) stats_one(T, tiles, S_dev, i, sh, alive_mask);        // read the finished game first ...
        uint8_t* base = tiles + (i >> 5) * (uint64_t)(32 * S_dev);
        const uint32_t sl = (uint32_t)(i & 31);
        for (uint32_t k = 0; k < S_dev / 8; ++k)                   // ... then overwrite the same slot
            *reinterpret_cast<uint2*>(base + rt_tile_off(8 * k, sl, n16)) = make_uint2(rec.w[2 * k], rec.w[2 * k + 1]);
        origin[i] = (uint32_t)i;
    }
    __syncthreads();
    stats_flush(sh, stats);
}

// ge_batch_reset in ONE launch: fold the resident sessions' final-state histograms into the accumulator (k_stats), rewrite
// every slot with the initial record (k_init), put the slot order back to identity (k_iota) and reset the compaction state and
// the phase-presence words (k_cstate_reset + a memset).  A batch of a steady-state ring starts over once per game cycle; as
// five stream operations that was 3.4 % of the (serialised) GPU time of the headline workload.
__global__ void __launch_bounds__(256)
k_reinit(const __grid_constant__ DevTable T, uint8_t* tiles, uint32_t S_dev, uint64_t n, uint64_t n_tiles, const __grid_constant__ InitRec rec,
         uint32_t* origin, unsigned long long* stats, unsigned long long* cstate, uint32_t* presence, unsigned long long epoch,
         uint32_t alive_mask) {
    __shared__ uint32_t sh[3 + 256 + 256];
    for (int i = threadIdx.x; i < 515; i += blockDim.x) sh[i] = 0;
    if (blockIdx.x == 0 && threadIdx.x < 16) cstate[threadIdx.x] = threadIdx.x == 0 ? n : threadIdx.x == 7 ? epoch : 0ull;
    if (blockIdx.x == 0 && threadIdx.x < 3) presence[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t n16 = S_dev / 16;
    const uint64_t total = n_tiles * 32;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        if (i < n) stats_one(T, tiles, S_dev, i, sh, alive_mask);  // read the finished game first ...
        uint8_t* base = tiles + (i >> 5) * (uint64_t)(32 * S_dev);
        const uint32_t sl = (uint32_t)(i & 31);
        for (uint32_t k = 0; k < S_dev / 8; ++k)                   // ... then overwrite the same slot
            *reinterpret_cast<uint2*>(base + rt_tile_off(8 * k, sl, n16)) = make_uint2(rec.w[2 * k], rec.w[2 * k + 1]);
        origin[i] = (uint32_t)i;
    }
    __syncthreads();
    stats_flush(sh, stats);
}

__global__ void k_autoreset_commit(unsigned long long* cstate, uint32_t* presence, uint32_t* rg, uint32_t next_launch_idx, unsigned long long n) {
    if (threadIdx.x != 0 || cstate[0] != 0) return;
    cstate[0] = n; cstate[5] = 0; cstate[8] += 1;
    presence[next_launch_idx % 3] = 1u;                            // every session is in phase index 0
    if (rg) for (int i = 0; i <= 33; ++i) rg[i] = 0;
}

// ------------------------------------------------------------------------------------ compaction
// Active-prefix compaction.  All live sessions of a batch advance in lockstep, so finished games leave
// holes that still cost issue slots.  Periodically the live sessions at the back are swapped, in place,
// with terminal sessions at the front; step kernels then walk only slots [0, n_active).  Inputs are the
// per-tile live masks and the live count the step kernel publishes — no pass over the records.  A swap
// costs about one step of traffic, so it only runs when a quarter of the prefix is dead (dead_shift = 2).
// Deterministic: ranks come from an exclusive scan, not from atomics.
// cstate: [0] n_active [1] n_live [2] pairs [3] live already in front [4] tiles of the old prefix
//         [5] live sessions after the last counted step (written by the step kernel) [6] scan ticket
constexpr int CS_TILES = 1024;          // tiles per scan block

__global__ void k_iota(uint32_t* origin, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) origin[i] = (uint32_t)i;
}

__global__ void k_cstate_reset(unsigned long long* cstate, unsigned long long n, unsigned long long epoch, int keep_device_epoch) {
    // [7] = host epoch tag, [8] = number of device-side re-initialisations (auto-reset) since the last host one
    // (kept when only the slot order is restored: the resident sessions' ids depend on it)
    if (threadIdx.x < 16 && !(keep_device_epoch && threadIdx.x == 8))
        cstate[threadIdx.x] = threadIdx.x == 0 ? n : threadIdx.x == 7 ? epoch : 0ull;
}

__device__ __forceinline__ uint32_t block_excl_scan_1024(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += u; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_warp[lane];
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, wi, d); if (lane >= d) wi += u; }
        s_warp[lane] = wi - w;
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    return s_warp[warp] + incl - v;
}

// The compaction kernels take a LIST of batches (blockIdx.y selects one): a ring stepped by one launch is also
// checked by one launch pair, and a single batch is a list of one.  `hint` is the batch's pinned host word pair
// {n_active, host epoch} (mapped into the device address space): the swap kernel stores the new values there
// directly, so the host's non-blocking progress hint costs no copy operations on the stream.
struct CompactSlot {
    uint8_t* tiles;
    uint32_t* origin;
    uint32_t* live_mask;
    uint32_t* loc;
    uint32_t* blk;
    unsigned long long* cstate;
    unsigned long long* hint;
    // Check log (mapped host memory, NULL = not kept): the scan kernel stores the host epoch tag in sched[check_idx] whenever
    // check number check_idx of an epoch RAN and in sched[64 + check_idx] when it also FIRED, so that the host can stop
    // launching the checks that never fire (ge_capi.cu: want_check).
    uint32_t* sched;
    uint32_t check_idx;
};
struct CompactArgs {
    int n;
    uint32_t S, dead_shift;
    CompactSlot s[GE_RING_MAX];
};

// grid = (ceil(max tiles / 1024), batches) blocks of 1024 threads.  loc[t] = live sessions in tiles of the same block
// before t, blk[b] = live sessions in blocks before b (filled in by the last block to finish).
__global__ void __launch_bounds__(1024)
k_compact_scan(const __grid_constant__ CompactArgs CA) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_total;
    __shared__ uint32_t s_last;
    // programmatic dependent launch (GE_OPT_PDL): wait for the counted step launch, then let the swap kernel start filling in
    grid_dependency_wait();
    grid_launch_dependents();
    const CompactSlot& Q = CA.s[blockIdx.y];
    const uint32_t* __restrict__ live_mask = Q.live_mask;
    uint32_t* loc = Q.loc;
    uint32_t* blk = Q.blk;
    unsigned long long* cstate = Q.cstate;
    const uint32_t dead_shift = CA.dead_shift;
    const uint64_t n_act = cstate[0];
    const uint64_t live_now = cstate[5];
    const uint64_t nt = (n_act + 31) >> 5;
    const uint32_t nblk = (uint32_t)((nt + CS_TILES - 1) / CS_TILES);
    // every block takes the same decision from the same two words (nobody writes them in this kernel)
    if (live_now > n_act || ((n_act - live_now) << dead_shift) < n_act) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            cstate[1] = n_act; cstate[2] = 0;
            if (Q.sched && Q.check_idx < 64) ((volatile uint32_t*)Q.sched)[Q.check_idx] = (uint32_t)cstate[7];
        }
        return;
    }
    if (blockIdx.x >= nblk) return;
    const uint64_t t = (uint64_t)blockIdx.x * CS_TILES + threadIdx.x;
    const uint32_t cnt = t < nt ? (uint32_t)__popc(live_mask[t]) : 0u;
    const uint32_t excl = block_excl_scan_1024(cnt, s_warp, &s_total);
    if (t < nt) loc[t] = excl;
    if (threadIdx.x == 0) {
        blk[blockIdx.x] = s_total;
        __threadfence();
        s_last = atomicAdd(&cstate[6], 1ull) == (unsigned long long)(nblk - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last block: exclusive scan of the block totals (nblk <= 1024), then the bounds of the swap
    const uint32_t bt = threadIdx.x < nblk ? ((volatile uint32_t*)blk)[threadIdx.x] : 0u;
    const uint32_t be = block_excl_scan_1024(bt, s_warp, &s_total);
    if (threadIdx.x < nblk) blk[threadIdx.x] = be;
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint64_t n_live = s_total;
        uint64_t pairs = 0, in_front = n_live;
        if (n_live < n_act) {
            const uint64_t tb = n_live >> 5;
            const uint32_t lb = (uint32_t)(n_live & 31);
            in_front = ((volatile uint32_t*)blk)[tb / CS_TILES] + ((volatile uint32_t*)loc)[tb] + __popc(live_mask[tb] & ((1u << lb) - 1u));
            pairs = n_live - in_front;
        }
        cstate[0] = n_live; cstate[1] = n_live; cstate[2] = pairs; cstate[3] = in_front; cstate[4] = nt; cstate[6] = 0;
        if (Q.sched && Q.check_idx < 64) {
            ((volatile uint32_t*)Q.sched)[64 + Q.check_idx] = (uint32_t)cstate[7];
            ((volatile uint32_t*)Q.sched)[Q.check_idx] = (uint32_t)cstate[7];
        }
    }
}

// Pair i swaps the i-th terminal session of the front region [0, n_live) with the i-th live session of the
// back region [n_live, old prefix).  Both are located by binary search over the two-level prefix
// P(t) = blk[t / 1024] + loc[t] (terminals before tile t = 32t - P(t)).  One thread per pair.
__global__ void k_compact_swap(const __grid_constant__ CompactArgs CA) {
    grid_dependency_wait();            // (GE_OPT_PDL) the scan's results; the batch's next step launch may start behind us
    grid_launch_dependents();
    const CompactSlot& Q = CA.s[blockIdx.y];
    uint8_t* tiles = Q.tiles;
    const uint32_t S = CA.S;
    uint32_t* origin = Q.origin;
    const uint32_t* __restrict__ live_mask = Q.live_mask;
    const uint32_t* __restrict__ loc = Q.loc;
    const uint32_t* __restrict__ blk = Q.blk;
    unsigned long long* cstate = Q.cstate;
    const uint64_t n_live = cstate[1], pairs = cstate[2], in_front = cstate[3], nt = cstate[4];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        cstate[5] = 0;                                           // the next counted step starts from zero
        if (Q.hint) {                                            // progress hint for the host: value first, then its epoch tag
            ((volatile unsigned long long*)Q.hint)[0] = cstate[0];
            __threadfence_system();
            ((volatile unsigned long long*)Q.hint)[1] = cstate[7];
        }
    }
    if (pairs == 0) return;
    const uint32_t n16 = S / 16;
    const bool half = (S % 16) != 0;
    // Compaction covers batches up to 2^25 sessions (ge_batch_set_compaction), so slots, tiles and counts fit 32 bits.  Both
    // searches are two-level like the prefix itself: first the scan block (blk[], at most a few dozen entries that stay in
    // L1), then the tile inside it with the block's base hoisted — one load and 32-bit arithmetic per probe instead of two
    // loads and 64-bit arithmetic (the swap was 38 % of a step launch's warp-instructions per firing check, ncu).
    const uint32_t n_live32 = (uint32_t)n_live, pairs32 = (uint32_t)pairs, in_front32 = (uint32_t)in_front, nt32 = (uint32_t)nt;
    const uint32_t tb = n_live32 >> 5, lb = n_live32 & 31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < pairs32; i += gridDim.x * blockDim.x) {
        // front: largest tile t in [0, tb] with D(t) = 32 t - P(t) <= i  (terminal sessions before tile t)
        uint32_t lo = 0, hi = tb / CS_TILES;
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (mid * (32u * CS_TILES) - blk[mid] <= i) lo = mid; else hi = mid - 1;
        }
        uint32_t base = blk[lo];
        hi = min(tb, lo * CS_TILES + (CS_TILES - 1));
        lo = lo * CS_TILES;
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (32u * mid - base - loc[mid] <= i) lo = mid; else hi = mid - 1;
        }
        const uint32_t valid_a = lo == tb ? ((1u << lb) - 1u) : 0xFFFFFFFFu;
        const uint32_t a = 32u * lo + (uint32_t)kth_set_bit<32>(~live_mask[lo] & valid_a, i - (32u * lo - base - loc[lo]));
        // back: largest tile t in [tb, nt - 1] with P(t) <= r  (live sessions before tile t)
        const uint32_t r = in_front32 + i;
        lo = tb / CS_TILES; hi = (nt32 - 1) / CS_TILES;
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (blk[mid] <= r) lo = mid; else hi = mid - 1;
        }
        base = blk[lo];
        hi = min(nt32 - 1, lo * CS_TILES + (CS_TILES - 1));
        lo = max(tb, lo * CS_TILES);
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (base + loc[mid] <= r) lo = mid; else hi = mid - 1;
        }
        const uint32_t b = 32u * lo + (uint32_t)kth_set_bit<32>(live_mask[lo], r - (base + loc[lo]));
        uint8_t* ba = tiles + (uint64_t)(a >> 5) * (32ull * S);
        uint8_t* bb = tiles + (uint64_t)(b >> 5) * (32ull * S);
        for (uint32_t c = 0; c < n16; ++c) {
            uint4* pa = reinterpret_cast<uint4*>(ba + c * 512u + (a & 31) * 16u);
            uint4* pb = reinterpret_cast<uint4*>(bb + c * 512u + (b & 31) * 16u);
            const uint4 t = *pa; *pa = *pb; *pb = t;
        }
        if (half) {
            uint2* pa = reinterpret_cast<uint2*>(ba + n16 * 512u + (a & 31) * 8u);
            uint2* pb = reinterpret_cast<uint2*>(bb + n16 * 512u + (b & 31) * 8u);
            const uint2 t = *pa; *pa = *pb; *pb = t;
        }
        const uint32_t t = origin[a]; origin[a] = origin[b]; origin[b] = t;
    }
}

// tiles -> canonical records in ORIGINAL session order when slots have been permuted by compaction
__global__ void k_export_perm(const uint8_t* tiles, uint32_t S_dev, uint32_t S_canon, const uint32_t* __restrict__ origin,
                              uint64_t n, uint64_t first, uint64_t count, uint8_t* out) {
    const uint32_t n16 = S_dev / 16;
    for (uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t o = origin[slot];
        if (o < first || o >= first + count) continue;
        const uint8_t* base = tiles + (slot >> 5) * (uint64_t)(32 * S_dev);
        const uint32_t sl = (uint32_t)(slot & 31);
        for (uint32_t k = 0; k < S_canon / 8; ++k)
            *reinterpret_cast<uint2*>(out + (o - first) * S_canon + 8 * k) = *reinterpret_cast<const uint2*>(base + rt_tile_off(8 * k, sl, n16));
    }
}

// ------------------------------------------------------------------------------------ phase regrouping
// Sessions of one batch normally advance in lockstep, so a warp's 32 sessions are in the same phase and the
// warp-uniform phase switch of the step kernel costs one body.  Tables with loops of data-dependent length
// (the tie -> re-vote loop of werewolf-revote) de-synchronise them: every warp then runs every phase body
// present in its tile, i.e. each launch pays for the most expensive phase.  Regrouping is a counting sort of
// the active prefix by phase index (terminal sessions last, which also makes it a compaction): afterwards at
// most one tile per phase is mixed.  It is decided and done on the device (no host synchronisation):
//   step kernel (counted launch): rg[0..31] += sessions that entered phase i, rg[32] += mixed tiles
//   k_regroup_plan: trigger when >= 1/2^mixed_shift of the tiles are mixed or >= 1/2^dead_shift of the prefix
//                   is dead; exclusive scan of the live phases' counts -> first slot of every phase
//   k_regroup_scatter: blocks of 1024 slots; ranks inside a block from warp match + shared counters, the
//                   block's range inside each phase from one global atomic per phase; record -> scratch
//   k_regroup_copyback: scratch -> session store (the prefix only)
// Slot order inside a phase is not deterministic, and does not have to be: session ids come from the origin
// map, exports are in original order and the statistics are sums.
enum { RG_HIST = 0, RG_MIXED = 32, RG_OLD = 33, RG_BASE = 64, RG_CURSOR = 100, RG_WORDS = 160, RG_TERM = 32 };

__global__ void k_regroup_plan(unsigned long long* cstate, uint32_t* rg, uint32_t nonterm, uint32_t mixed_shift, uint32_t dead_shift) {
    const int lane = threadIdx.x;                      // one warp
    const uint64_t n_act = cstate[0];
    const uint32_t cnt = ((nonterm >> lane) & 1u) ? rg[RG_HIST + lane] : 0u;
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += u; }
    const uint64_t n_live = __shfl_sync(0xFFFFFFFFu, incl, 31);
    const uint64_t mixed = rg[RG_MIXED];
    const uint64_t tiles = (n_act + 31) >> 5;
    const bool trig = n_act > 0 && n_live <= n_act &&
                      (((mixed << mixed_shift) >= tiles && mixed > 0) || (((n_act - n_live) << dead_shift) >= n_act && n_live < n_act));
    __syncwarp();
    rg[RG_BASE + lane] = incl - cnt;
    rg[RG_CURSOR + lane] = 0;
    rg[RG_HIST + lane] = 0;
    if (lane == 0) {
        rg[RG_BASE + RG_TERM] = (uint32_t)n_live;
        rg[RG_CURSOR + RG_TERM] = 0;
        rg[RG_MIXED] = 0;
        rg[RG_OLD] = trig ? (uint32_t)n_act : 0u;
        if (trig) cstate[0] = n_live;
        cstate[5] = 0;
    }
}

__global__ void __launch_bounds__(1024)
k_regroup_scatter(const uint8_t* __restrict__ tiles, uint32_t S, const uint32_t* __restrict__ origin, uint32_t* rg, uint32_t nonterm,
                  uint8_t* __restrict__ out_tiles, uint32_t* __restrict__ out_origin) {
    __shared__ uint32_t s_cnt[RG_TERM + 1], s_base[RG_TERM + 1];
    const uint32_t old = rg[RG_OLD];
    if (old == 0) return;
    const uint32_t n16 = S / 16;
    const bool half = (S % 16) != 0;
    const int lane = threadIdx.x & 31;
    for (uint32_t c0 = blockIdx.x * 1024u; c0 < old; c0 += gridDim.x * 1024u) {
        if (threadIdx.x <= RG_TERM) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t slot = c0 + threadIdx.x;
        const bool valid = slot < old;
        const uint8_t* src = tiles + (uint64_t)(slot >> 5) * (32ull * S);
        int key = -1;
        if (valid) {
            const uint32_t ph = *reinterpret_cast<const uint32_t*>(src + (slot & 31) * 16u) & 31u;
            key = ((nonterm >> ph) & 1u) ? (int)ph : RG_TERM;
        }
        const uint32_t same = __match_any_sync(0xFFFFFFFFu, key);
        const int leader = __ffs(same) - 1;
        uint32_t wbase = 0;
        if (valid && lane == leader) wbase = atomicAdd(&s_cnt[key], (uint32_t)__popc(same));
        wbase = __shfl_sync(0xFFFFFFFFu, wbase, leader);
        const uint32_t rank = __popc(same & ((1u << lane) - 1u));
        __syncthreads();
        if (threadIdx.x <= RG_TERM) {
            const uint32_t n = s_cnt[threadIdx.x];
            s_base[threadIdx.x] = n ? rg[RG_BASE + threadIdx.x] + atomicAdd(&rg[RG_CURSOR + threadIdx.x], n) : 0u;
        }
        __syncthreads();
        if (valid) {
            const uint32_t dst = s_base[key] + wbase + rank;
            uint8_t* db = out_tiles + (uint64_t)(dst >> 5) * (32ull * S);
            for (uint32_t c = 0; c < n16; ++c)
                *reinterpret_cast<uint4*>(db + c * 512u + (dst & 31) * 16u) = *reinterpret_cast<const uint4*>(src + c * 512u + (slot & 31) * 16u);
            if (half)
                *reinterpret_cast<uint2*>(db + n16 * 512u + (dst & 31) * 8u) = *reinterpret_cast<const uint2*>(src + n16 * 512u + (slot & 31) * 8u);
            out_origin[dst] = origin[slot];
        }
        __syncthreads();
    }
}

__global__ void k_regroup_copyback(uint8_t* __restrict__ tiles, uint32_t S, uint32_t* __restrict__ origin, const uint32_t* __restrict__ rg,
                                   const uint8_t* __restrict__ in_tiles, const uint32_t* __restrict__ in_origin,
                                   const unsigned long long* cstate, unsigned long long* hint, uint32_t* tile_present) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && hint) {           // progress hint for the host (see k_compact_swap)
        ((volatile unsigned long long*)hint)[0] = cstate[0];
        __threadfence_system();
        ((volatile unsigned long long*)hint)[1] = cstate[7];
    }
    const uint32_t old = rg[RG_OLD];
    const uint32_t n16 = S / 16;
    const bool half = (S % 16) != 0;
    // whole tiles (a warp = the 32 slots of one tile), so that the per-tile phase-presence word of every tile the sort
    // touched can be rebuilt on the way: the step kernel's per-tile column needs stay valid across a regroup
    const uint32_t end = (old + 31u) & ~31u;
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < end; slot += gridDim.x * blockDim.x) {
        const uint64_t off = (uint64_t)(slot >> 5) * (32ull * S);
        uint32_t ph;
        if (slot < old) {
            uint4 h = make_uint4(0, 0, 0, 0);
            for (uint32_t c = 0; c < n16; ++c) {
                const uint4 v = *reinterpret_cast<const uint4*>(in_tiles + off + c * 512u + (slot & 31) * 16u);
                if (c == 0) h = v;
                *reinterpret_cast<uint4*>(tiles + off + c * 512u + (slot & 31) * 16u) = v;
            }
            if (half)
                *reinterpret_cast<uint2*>(tiles + off + n16 * 512u + (slot & 31) * 8u) = *reinterpret_cast<const uint2*>(in_tiles + off + n16 * 512u + (slot & 31) * 8u);
            origin[slot] = in_origin[slot];
            ph = h.x & 31u;
        } else {                                                  // beyond the sorted prefix: the record stays, its phase still counts
            ph = *reinterpret_cast<const uint32_t*>(tiles + off + (slot & 31) * 16u) & 31u;
        }
        if (tile_present) {
            const uint32_t bits = __reduce_or_sync(0xFFFFFFFFu, 1u << ph);
            if ((threadIdx.x & 31) == 0) tile_present[slot >> 5] = bits;
        }
    }
}

// ------------------------------------------------------------------------------------ audience masks
// Evaluates up to 32 DNF predicates (same encoding as the table's) for every session of a window: the lane
// masks of the DSL's audience_groups (reference games/werewolf-(mafia).yaml:138-165; consumed by the UI tools'
// audience_ids, src/lib/canvas/types.ts:14-17).  Output row = session (original order), column = predicate.
struct PredList { ge_pred_t p[32]; int n; };

// packed_bucket: 0 = canonical columns, 8 / 16 = the packed store of that bucket (werewolf family)
__global__ void k_eval_preds(const __grid_constant__ DevTable T, const __grid_constant__ PredList PL, const uint8_t* tiles,
                             uint32_t S_dev, const uint32_t* __restrict__ origin, uint64_t n, uint64_t first, uint64_t count,
                             uint32_t* out, int packed_bucket) {
    const uint32_t n16 = S_dev / 16;
    const int P = T.h.n_players;
    const uint32_t ALL = P >= 32 ? 0xFFFFFFFFu : ((1u << P) - 1u);
    for (uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t o = origin ? origin[slot] : slot;
        if (o < first || o >= first + count) continue;
        const uint8_t* base = tiles + (slot >> 5) * (uint64_t)(32 * S_dev);
        const uint32_t sl = (uint32_t)(slot & 31);
        uint32_t F[16];
#pragma unroll
        for (int f = 0; f < 16; ++f) F[f] = 0;
        F[15] = ALL;
        if (T.h.family == FAM_WEREWOLF && packed_bucket != 0) {
            uint32_t w[16];
            if (packed_bucket == 8) {
                uint32_t w8[14];
                tile_load_words<8, true>(tiles, slot, w8);
#pragma unroll
                for (int k = 0; k < 14; ++k) w[k] = w8[k];
                w[14] = 0; w[15] = 0;
            } else {
                tile_load_words<16, true>(tiles, slot, w);
            }
#pragma unroll
            for (int f = 0; f < 8; ++f) F[f] = w[2 + f];
            const uint32_t lo = w[10], hi = w[11];
            F[8] = F[7] ? ~(lo | hi) & ALL : 0u; F[9] = lo & ~hi; F[10] = ~lo & hi; F[11] = lo & hi;
            F[12] = F[7] ? ALL : 0u;
            for (int k = 0; k < T.h.n_cmp && k < 2; ++k)
                for (int p = 0; p < P; ++p)
                    if (cmp_holds(T.h.cmp[k].op, (w[12 + (p >> 2)] >> (8 * (p & 3))) & 0xFFu, T.h.cmp[k].constant)) F[13 + k] |= 1u << p;
        } else if (T.h.family == FAM_WEREWOLF) {
            const uint4 c0 = *reinterpret_cast<const uint4*>(base + sl * 16);
            const uint4 c1 = *reinterpret_cast<const uint4*>(base + 512 + sl * 16);
            const uint4 c2 = *reinterpret_cast<const uint4*>(base + 1024 + sl * 16);
            F[0] = c0.z; F[1] = c0.w; F[2] = c1.x; F[3] = c1.y; F[4] = c1.z; F[5] = c1.w; F[6] = c2.x; F[7] = c2.y;
            F[8] = c2.y ? ~(c2.z | c2.w) & ALL : 0u; F[9] = c2.z & ~c2.w; F[10] = ~c2.z & c2.w; F[11] = c2.z & c2.w;
            F[12] = c2.y ? ALL : 0u;
            for (int k = 0; k < T.h.n_cmp && k < 2; ++k)            // comparison fields: selected_target_id <op> constant
                for (int p = 0; p < P; ++p)
                    if (cmp_holds(T.h.cmp[k].op, base[rt_tile_off(48 + p, sl, n16)], T.h.cmp[k].constant)) F[13 + k] |= 1u << p;
        } else {
            for (int p = 0; p < P; ++p) {
                const uint32_t pw = *reinterpret_cast<const uint32_t*>(base + rt_tile_off(8 + 4 * p, sl, n16));
                const uint32_t fl = pw >> 24;
#pragma unroll
                for (int f = 0; f < 5; ++f) F[f] |= ((fl >> f) & 1u) << p;
                for (int k = 0; k < T.h.n_cmp && k < 4; ++k)
                    if (cmp_holds(T.h.cmp[k].op, (pw >> (8 * T.h.cmp[k].value_field)) & 0xFFu, T.h.cmp[k].constant)) F[11 + k] |= 1u << p;
            }
        }
        for (int j = 0; j < PL.n; ++j) {
            const ge_pred_t pr = PL.p[j];
            uint32_t res = 0;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint32_t pos = (c ? pr.pos1 : pr.pos0) & 0x7FFFu, neg = c ? pr.neg1 : pr.neg0;      // bit 15 of pos0: "continued" (the caller ORs)
                uint32_t m = ALL;
#pragma unroll
                for (int f = 0; f < 16; ++f) {
                    if ((pos >> f) & 1u) m &= F[f];
                    if ((neg >> f) & 1u) m &= ~F[f];
                }
                res |= m;
            }
            out[(o - first) * PL.n + j] = res;
        }
    }
}
