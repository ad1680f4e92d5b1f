// Warp-specialised tcgen05 GEMM skeleton with pluggable A-producer and epilogue (sm_100a).
//
//   D[128 x N] (TMEM, fp32)  +=  A[128 x 64-chunk] (smem, bf16 hi/lo)  *  W[N x 64-chunk]^T (smem, bf16 hi/lo)
//
// Roles (one CTA per SM, persistent over 128-row tiles; PW = 4 or 8 producer warps):
//   warps 0..PW-1 A producers: build the 128x64 A chunk IN SHARED MEMORY (plain fp32 rows, or the
//                            PointConv gather+aggregate, or the cost-volume gather+add+act), split
//                            fp32 -> bf16 hi/lo, write it in the canonical SWIZZLE_128B layout
//   warp  PW   MMA issuer  : warp-uniform loop, the tcgen05 instructions under elect.sync: 3 tcgen05.mma per K-step
//                            (hi*hi, hi*lo, lo*hi).  (Producer thread 0 also issues the 1-D bulk TMA of the pre-packed
//                            weight chunk for the stage it is about to fill: no separate loader warp.)
//   warps PW+1..PW+4 epilogue: tcgen05.ld the accumulator, scale/shift/activation, store
//                            (kMergedIssuer producers: warp PW is issuer AND epilogue of TMEM lane quarter 0, the
//                            epilogue warps are PW..PW+3 - 12 warps, 168 registers per thread)
// Pipelines: smem stages (full_a/full_b/empty mbarriers) and two TMEM accumulator buffers
// (tmem_full/tmem_empty) so the epilogue of tile i overlaps the main loop of tile i+1.
#pragma once
#include <cuda.h>           // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include "tc_common.cuh"

// A/B switch (kdpc_tc_set_async): 0 = synchronous register-staged producers everywhere
extern "C" int kdpc_tc_async_enabled(void);
// debug: device buffer of 200 x 16 int64 receiving CTA 0's per-iteration clock64 stamps (kdpc_tc_set_trace); NULL = off
extern "C" void *kdpc_tc_trace_buffer(void);

namespace kdpc {
namespace tc {

constexpr int MAX_STAGES = 4;
// threads of a kernel instance: PW producer warps + MMA warp + 4 epilogue warps.  Producers that declare
// kMergedIssuer run the MMA issuer INSIDE the first epilogue warp (12 warps instead of 13): the register file is
// split per SM sub-partition, so 13 warps (4 on one sub-partition) cap every thread at 128 registers while 12 warps
// (3 per sub-partition) allow 168 - what a producer that keeps a whole neighbourhood in registers needs.
template <class P, class = void> struct merged_issuer { static constexpr bool value = false; };
template <class P> struct merged_issuer<P, decltype((void)P::kMergedIssuer)> { static constexpr bool value = P::kMergedIssuer; };
// Producers that declare kOwnsLoop implement run_tile(tile, c_begin, c_end, ptid, raw_base, raw_full, acquire, release)
// instead of fill(); with kIssuers > 0 they own the raw staging area and its mbarriers (kIssuers arrivals each)
template <class P, class = void> struct owns_loop { static constexpr bool value = false; };
template <class P> struct owns_loop<P, decltype((void)P::kOwnsLoop)> { static constexpr bool value = P::kOwnsLoop; };
// Asynchronous producers that declare kIssuerWarps > 0 get that many extra warps which do nothing but issue the
// cp.async gathers (blocking on the load/store unit as long as they like), kWarps warps only convert: issue and
// conversion overlap instead of alternating in the same warps.  raw_full[] then counts the issuer threads
// (kIssuers = 32 * kIssuerWarps) and raw_empty[] hands a staging slot back (one arrival per converting warp).
template <class P, class = void> struct issuer_warps { static constexpr int value = 0; };
template <class P> struct issuer_warps<P, decltype((void)P::kIssuerWarps)> { static constexpr int value = P::kIssuerWarps; };
// Producers that declare kEpilogueWarps = 8 get two epilogue warps per TMEM lane quarter (the second four receive
// quarter + 4 and work on their own column range of the accumulator): for tiles whose epilogue, not the MMA, paces the pipeline.
template <class P, class = void> struct epilogue_warps { static constexpr int value = 4; };
template <class P> struct epilogue_warps<P, decltype((void)P::kEpilogueWarps)> { static constexpr int value = P::kEpilogueWarps; };
// P::kRelaxedWaits = true: the converting warps' wait for the gathered rows and the epilogue warps' wait for a finished
// accumulator back off with nanosleep between polls.  In the paired cost volume HALF of all executed instructions were
// bare try_wait polls (ncu source page: 35.7 M + 29.6 M of 131 M, issue slots 73 % busy) - on the schedulers that also
// have to issue the gathers those warps are waiting for.
template <class P, class = void> struct relaxed_waits { static constexpr bool value = false; };
template <class P> struct relaxed_waits<P, decltype((void)P::kRelaxedWaits)> { static constexpr bool value = P::kRelaxedWaits; };
template <bool RELAXED, unsigned NS>
__device__ __forceinline__ void mbar_wait_poll(uint64_t *bar, uint32_t parity) {
    if constexpr (!RELAXED) {
        mbar_wait(bar, parity);
    } else {
        for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
            __nanosleep(NS);
            if (spins > (1u << 22)) __trap();
        }
    }
}
template <class Producer> constexpr int num_threads() {
    return (Producer::kWarps + issuer_warps<Producer>::value + (merged_issuer<Producer>::value ? 0 : 1) + epilogue_warps<Producer>::value) * 32;
}

struct GemmShape {
    long long m;          // rows
    int n;                // logical output columns
    int n_pad;            // UMMA N: multiple of 16, 16..256
    int acc_stride;       // TMEM columns per accumulator buffer: round_up(n_pad, 32)
    int tmem_cols;        // power of two >= nacc * acc_stride
    int nacc_log2;        // accumulator buffers in TMEM: 4 (log2 = 2) when 4 * acc_stride <= 512 columns, else 2
    int num_chunks;       // pipeline iterations per tile: K chunks of 64 (x neighbour passes of the PointConv producer)
    int wchunks;          // chunks of the packed weight: iteration c uses weight chunk c % wchunks
    int k_total;          // packed K (multiple of 16): K-steps beyond it are skipped
    int stages;
    int splits;           // split-K: each tile's K chunks are spread over `splits` work items (partial sums to a workspace)
    int chunks_per_split;
    int nsplit;           // 1: the `splits` work items of a tile are COLUMN blocks of n_pad outputs each (whole K, final
                          // results, no workspace): split index j computes output columns j * n_pad .. (plan_split_n)
    int w_n_pad;          // rows of the packed weight (= n_pad unless nsplit)
    int raw_bytes;        // bytes of one raw staging buffer of an asynchronous producer (0: none)
    int raw_stages;       // lookahead + 1
    long long num_tiles;
    const unsigned char *wpacked;   // [chunk][hi|lo][n_pad][128 B]
    long long *trace;     // debug (tools/trace_*.py): per-iteration clock64 stamps of CTA 0; nullptr = off (one uniform
                          // predicate per stamp site, nothing is written)
};

constexpr int MAX_RAW_STAGES = 4;
constexpr int SMEM_BUDGET = 200 * 1024;

static inline int b_stage_bytes(int n_pad) { return 2 * n_pad * 128; }

static inline size_t smem_bytes(int n_pad, int stages, int raw_total = 0) {
    return (size_t)stages * (A_STAGE_BYTES + b_stage_bytes(n_pad)) + (size_t)raw_total + 1024;     // + alignment slack
}

static inline int pick_stages(int n_pad, int raw_total = 0) {
    int s = (int)((SMEM_BUDGET - 1024 - raw_total) / (A_STAGE_BYTES + b_stage_bytes(n_pad)));
    return s > MAX_STAGES ? MAX_STAGES : s;
}

// raw_bytes / raw_stages: staging of an asynchronous producer (Producer::kRawBytes, kLookahead + 1); the shape is
// unusable (stages < 2) when the operand stages no longer fit next to it: callers fall back to a synchronous producer.
static inline GemmShape make_shape(long long m, int n, int k_packed, const void *wpacked, int raw_bytes = 0,
                                   int raw_stages = 0) {
    GemmShape g;
    g.m = m;
    g.n = n;
    g.n_pad = (n + 15) / 16 * 16;
    g.acc_stride = (g.n_pad + 31) / 32 * 32;
    int c = 32;
    g.nacc_log2 = 4 * g.acc_stride <= 512 ? 2 : 1;               // more buffers: the MMA runs ahead of a slower epilogue
    while (c < (g.acc_stride << g.nacc_log2)) c <<= 1;
    g.tmem_cols = c;
    g.k_total = (k_packed + 15) / 16 * 16;
    g.num_chunks = (k_packed + CHUNK_K - 1) / CHUNK_K;
    g.wchunks = g.num_chunks;
    g.splits = 1;
    g.chunks_per_split = g.num_chunks;
    g.nsplit = 0;
    g.w_n_pad = g.n_pad;
    g.raw_bytes = raw_bytes;
    g.raw_stages = raw_stages;
    g.stages = pick_stages(g.n_pad, raw_bytes * raw_stages);
    g.num_tiles = (m + TILE_M - 1) / TILE_M;
    g.wpacked = reinterpret_cast<const unsigned char *>(wpacked);
    g.trace = reinterpret_cast<long long *>(kdpc_tc_trace_buffer());
    return g;
}

// Split-K for small-M / large-K layers (a handful of 128-row tiles would otherwise leave most SMs idle while each
// CTA walks >100 K-chunks): partial accumulators go to a workspace [splits][M][n_pad] and splitk_reduce_kernel adds
// them in split order (deterministic) and applies the epilogue.  (Fusing that reduction into the GEMM - the last work item of
// a tile to finish sums the partials - was measured 10 % slower end to end: one warp per tile quarter then walks
// splits x 32 rows of L2 reads on the kernel's critical path.)
static inline void plan_split_k(GemmShape &g) {
    g.splits = 1;
    g.chunks_per_split = g.num_chunks;
    const int sms = device_sms();
    if (g.num_tiles * 2 > sms || g.num_chunks < 4) return;
    int want = (int)(sms / g.num_tiles);
    if (want > g.num_chunks / 2) want = g.num_chunks / 2;       // >= 2 chunks per work item
    if (want > 16) want = 16;
    if (want < 2) return;
    g.chunks_per_split = (g.num_chunks + want - 1) / want;
    g.splits = (g.num_chunks + g.chunks_per_split - 1) / g.chunks_per_split;
}
static inline size_t split_k_ws_bytes(const GemmShape &g) {
    return g.splits > 1 && !g.nsplit ? (size_t)g.splits * (size_t)g.m * g.n_pad * sizeof(float) : 0;
}
// Split-N for small-M plain layers with >= 64 outputs (the 64 .. 4096-point levels: 4 .. 64 row tiles): a work item
// is a 128-row x 64- (32- for 64-wide layers, or 128-) column block over the WHOLE K.  Against split-K: final results straight from the
// epilogue (no partial sums, no reduce launch), weight stages of 16 KB instead of 64, four pipeline stages; the rows
// are re-read from L2 once per column block.  (4096 x 256 -> 256: 25-29 us as split-K + reduce.)
static inline bool plan_split_n(GemmShape &g) {
    const int sms = device_sms();
    // (long K stays with split-K: its shorter accumulation chains keep the 1e-5 error budget of the K = 3120 layers)
    if (g.num_tiles * 2 > sms || g.n_pad < 64 || (g.n_pad & 63) != 0 || g.num_chunks > 8) return false;
    if (g.n_pad == 64 && g.num_chunks < 2) return false;         // (a single-chunk 64-wide layer: nothing to gain)
    int nb = g.n_pad == 64 ? 32 : 64;
    if (g.num_tiles * (g.n_pad / 64) > sms && g.n_pad > 128 && (g.n_pad & 127) == 0) nb = 128;
    g.w_n_pad = g.n_pad;
    g.splits = g.n_pad / nb;
    g.nsplit = 1;
    g.chunks_per_split = g.num_chunks;
    g.n_pad = nb;
    g.acc_stride = nb;
    g.nacc_log2 = 2;
    g.tmem_cols = 4 * nb;
    g.stages = pick_stages(nb, g.raw_bytes * g.raw_stages);
    return true;
}

// Producer concept:
//   struct P { struct Args {...};
//              static constexpr int kWarps, kGroups;   // producer warps, independent groups (kWarps/kGroups warps fill one stage)
//              static constexpr bool kAsync;           // true: prime/issue/convert with kLookahead, kIssuers, raw staging (see the loop)
//              static __device__ void prologue(const Args&, int tid, int nthreads);      // all threads, before the role split
//              __device__ P(const Args&, const GemmShape&);
//              __device__ void begin_tile(long long tile, int r);                       // r = producer thread 0..127 = tile row
//              __device__ void fill(int chunk, unsigned char *a_hi, unsigned char *a_lo, int r); };
// Epilogue concept:
//   struct E { struct Args {...};
//              __device__ void tile(const Args&, const GemmShape&, long long tile, int split, uint32_t tmem_acc, int quarter, int lane); };
// work item -> (tile, split): work items are < 2^31 (checked by the launchers' shapes) and almost always splits == 1; a
// 64-bit division per tile in the single-warp MMA / epilogue loops cost ~1000 cycles of dependent integer code
__device__ __forceinline__ void work_item(long long w, int splits, long long &tile, int &split) {
    if (splits == 1) { tile = w; split = 0; return; }
    const unsigned t = (unsigned)w / (unsigned)splits;
    tile = t;
    split = (int)((unsigned)w - t * (unsigned)splits);
}

template <class Producer, class Epilogue>
__global__ void __launch_bounds__(num_threads<Producer>(), 1)
tc_gemm_kernel(const GemmShape g, const __grid_constant__ typename Producer::Args pa, const typename Epilogue::Args ea) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_a[MAX_STAGES], full_b[MAX_STAGES], empty[MAX_STAGES];
    __shared__ __align__(8) uint64_t tmem_full[4], tmem_empty[4];
    __shared__ __align__(8) uint64_t raw_full[MAX_RAW_STAGES], raw_empty[MAX_RAW_STAGES];
    __shared__ uint32_t tmem_base_smem;

    constexpr int PW = Producer::kWarps;                                     // 4 or 8: epilogue warps PW+1..PW+4 have (warp & 3) = 1,2,3,0
    constexpr bool MG = merged_issuer<Producer>::value;                      // warp PW = issuer AND epilogue of quarter 0
    constexpr int IW = issuer_warps<Producer>::value;                        // dedicated gather-issue warps PW..PW+IW-1
    constexpr int MW = PW + IW;                                              // the MMA issuer warp (epilogue warps follow)
    constexpr int EW = epilogue_warps<Producer>::value;                      // 4, or 8 (two per TMEM lane quarter)
    static_assert(EW == 4 || (EW == 8 && !MG), "epilogue warps: 4, or 8 without a merged issuer");
    static_assert(IW == 0 || (Producer::kAsync && !MG && ((MW + 1) & 3) == 1), "issuer warps: async producers, epilogue quarters 1,2,3,0");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);       // 1024-byte aligned tiles
    const int bbytes = 2 * g.n_pad * 128;
    unsigned char *a_base = smem;
    unsigned char *b_base = smem + (size_t)g.stages * A_STAGE_BYTES;
    unsigned char *raw_base = b_base + (size_t)g.stages * bbytes;            // asynchronous producers only

    if (tid == 0) {
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(&full_a[s], PW / Producer::kGroups);   // one arrive per producer warp of the filling group
            mbar_init(&full_b[s], 1);        // expect_tx arrive of the W loader
            mbar_init(&empty[s], 1);         // tcgen05.commit
        }
        if constexpr (Producer::kIssuers > 0)
            for (int r = 0; r < MAX_RAW_STAGES; ++r) {
                mbar_init(&raw_full[r], Producer::kIssuers);
                mbar_init(&raw_empty[r], PW);                     // (issuer-warp producers) one arrival per converting warp
            }   // arrive.expect_tx per issuing thread
        for (int a = 0; a < 4; ++a) {
            mbar_init(&tmem_full[a], 1);     // tcgen05.commit
            mbar_init(&tmem_empty[a], EW);   // one arrive per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == MW) tmem_alloc(&tmem_base_smem, (uint32_t)g.tmem_cols);
    // everything above is private to the CTA; from here on global memory written by the previous kernel is read
    asm volatile("griddepcontrol.wait;" ::: "memory");
    Producer::prologue(pa, tid, (int)blockDim.x);                            // optional CTA-wide staging (before the role split)
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_smem;
    const uint32_t nacc_mask = (1u << g.nacc_log2) - 1u;

    if (warp < PW) {
        // ================= A producers =================
        // PG independent groups of GW warps; group g fills pipeline iterations it = g (mod PG), so PG
        // gather round-trips are in flight per SM.
        constexpr int PG = Producer::kGroups, GW = PW / PG;
        const int grp = warp / GW;
        const int ptid = tid - grp * GW * 32;
        Producer prod(pa, g);
        if constexpr (Producer::kAsync) {
            // Asynchronous producer: the global -> shared gathers of iteration it + LA are issued (bulk copies that
            // complete on raw_full[]) while iteration it is converted fp32 -> bf16 hi/lo into the operand stage, so
            // the gather latency never sits on the critical path and no register holds data in flight.
            // (all bookkeeping is incremental 32-bit arithmetic: 64-bit divisions cost more than the conversion itself)
            const int RAW = g.raw_stages, LA = RAW - 1;                     // 2 or 3 raw buffers: lookahead 1 or 2
            // (split-N plans: the cursors walk WORK ITEMS = (tile, column block), tl() is an item's row tile)
            const int nsp = g.nsplit ? g.splits : 1;
            auto tl = [&](int item) { return nsp == 1 ? item : item / nsp; };
            const int tile_step = (int)gridDim.x, ntiles = (int)g.num_tiles * nsp, nchunks = g.num_chunks;
            int total = ((ntiles - (int)blockIdx.x + tile_step - 1) / tile_step) * nchunks;       // iterations of this CTA
            // cursors: `ah` = the iteration being issued (i + LA), `nx` = the one after it (index prefetch), `cu` = converted
            int ah_tile = (int)blockIdx.x, ah_chunk = 0, nx_tile = ah_tile, nx_chunk = 0, cu_tile = ah_tile, cu_chunk = 0;
            auto step = [&](int &t, int &c) { if (++c == nchunks) { c = 0; t += tile_step; } };
            step(nx_tile, nx_chunk);
            int ah_slot = 0, issued = 0;
            if constexpr (IW == 0) prod.prime(tl(ah_tile), ptid);
            auto issue_next = [&]() {
                prod.issue(tl(ah_tile), ah_chunk, nx_tile < ntiles ? tl(nx_tile) : -1, raw_base + (size_t)ah_slot * g.raw_bytes,
                           &raw_full[ah_slot], ptid);
                step(ah_tile, ah_chunk);
                step(nx_tile, nx_chunk);
                ah_slot = ah_slot + 1 == RAW ? 0 : ah_slot + 1;
                ++issued;
            };
            if constexpr (IW == 0) {
                for (int la = 0; la < LA; ++la)
                    if (issued < total) issue_next();
            }
            int s = 0, cu_slot = 0;
            uint32_t ph = 0, raw_ph = 0;
            const bool b_resident = g.wchunks == 1 && nsp == 1;   // one weight chunk: each stage's copy is loaded once
            auto stamp = [&](int i, int ev) {                     // debug trace of CTA 0 (kdpc_tc_set_trace)
                if (g.trace != nullptr && blockIdx.x == 0 && tid == 0 && i < 200) g.trace[i * 16 + ev] = clock64();
            };
            for (int i = 0; i < total; ++i) {
                stamp(i, 0);
                // every producer finished converting i-1: its raw slot may be refilled.  (Reads and cp.async writes are
                // both generic-proxy accesses: the barrier orders them, no proxy fence is needed here.)
                if constexpr (IW == 0) {
                    asm volatile("bar.sync 1, %0;" ::"n"(PW * 32) : "memory");
                    stamp(i, 1);
                    if (issued < total) issue_next();
                }
                stamp(i, 2);
                mbar_wait(&empty[s], ph ^ 1);
                stamp(i, 3);
                if (ptid == 0 && (!b_resident || i < g.stages)) {
                    mbar_expect_tx(&full_b[s], (uint32_t)bbytes);
                    if (nsp > 1) {                                // rows (item % splits) * n_pad .. of the chunk: its hi and lo pieces
                        const int nb = cu_tile - tl(cu_tile) * nsp;
                        const unsigned char *src = g.wpacked + (size_t)(cu_chunk % g.wchunks) * (2 * g.w_n_pad * 128) + (size_t)nb * g.n_pad * 128;
                        tma_load_1d(b_base + (size_t)s * bbytes, src, (uint32_t)bbytes / 2, &full_b[s]);
                        tma_load_1d(b_base + (size_t)s * bbytes + bbytes / 2, src + (size_t)g.w_n_pad * 128, (uint32_t)bbytes / 2, &full_b[s]);
                    } else {
                        tma_load_1d(b_base + (size_t)s * bbytes, g.wpacked + (size_t)(b_resident ? 0 : cu_chunk % g.wchunks) * bbytes, (uint32_t)bbytes, &full_b[s]);
                    }
                }
                mbar_wait_poll<relaxed_waits<Producer>::value, 32>(&raw_full[cu_slot], raw_ph);
                stamp(i, 4);
                unsigned char *a_hi = a_base + (size_t)s * A_STAGE_BYTES;
                prod.convert(tl(cu_tile), cu_chunk, raw_base + (size_t)cu_slot * g.raw_bytes, a_hi, a_hi + A_PART_BYTES, ptid);
                stamp(i, 5);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&full_a[s]);
                    if constexpr (IW > 0) mbar_arrive(&raw_empty[cu_slot]);   // the issuers may refill this staging slot
                }
                stamp(i, 6);
                step(cu_tile, cu_chunk);
                if (++s == g.stages) { s = 0; ph ^= 1; }
                if (++cu_slot == RAW) { cu_slot = 0; raw_ph ^= 1; }
            }
        } else {
        uint32_t it = 0;
        int ol_s = 0, ol_wc = 0;                                  // kOwnsLoop cursors: operand stage, its phase, weight chunk
        uint32_t ol_ph = 0;
        const long long work = g.num_tiles * g.splits;
        for (long long w = blockIdx.x; w < work; w += gridDim.x) {
            long long tile;
            int split_;
            work_item(w, g.splits, tile, split_);
            const int c_begin = g.nsplit ? 0 : split_ * g.chunks_per_split;
            const int c_end = min(g.num_chunks, c_begin + g.chunks_per_split);
            if constexpr (owns_loop<Producer>::value) {
                // the producer drives the K loop of its tile itself (tight inner loops with its state in registers):
                // acquire(c) waits for a free operand stage, starts the weight TMA and returns the A tile; release()
                // publishes it to the MMA issuer
                // (stage / phase / weight-chunk cursors are carried incrementally: a runtime % and / per chunk cost
                // more instructions than the hi/lo split of the chunk)
                auto stamp = [&](int ev) {                        // debug trace of CTA 0 (kdpc_tc_set_trace)
                    if (g.trace != nullptr && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == PW - 1) && it < 200)
                        g.trace[it * 16 + (warp == 0 ? 0 : 4) + ev] = clock64();
                };
                auto acquire = [&](int /*c*/) -> unsigned char * {
                    stamp(0);
                    mbar_wait(&empty[ol_s], ol_ph ^ 1u);
                    stamp(1);
                    if (ptid == 0) {
                        mbar_expect_tx(&full_b[ol_s], (uint32_t)bbytes);
                        tma_load_1d(b_base + (size_t)ol_s * bbytes, g.wpacked + (size_t)ol_wc * bbytes, (uint32_t)bbytes, &full_b[ol_s]);
                    }
                    return a_base + (size_t)ol_s * A_STAGE_BYTES;
                };
                auto release = [&]() {
                    stamp(2);
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_a[ol_s]);
                    stamp(3);
                    ++it;
                    if (++ol_s == g.stages) { ol_s = 0; ol_ph ^= 1u; }
                    if (++ol_wc == g.wchunks) ol_wc = 0;
                };
                ol_wc = c_begin == 0 ? 0 : c_begin % g.wchunks;
                prod.run_tile(tile, c_begin, c_end, ptid, raw_base, raw_full, acquire, release);
            } else {
            bool began = false;
            for (int c = c_begin; c < c_end; ++c, ++it) {
                if (PG > 1 && (int)(it % PG) != grp) continue;
                if (!began) { prod.begin_tile(tile, ptid); began = true; }
                const int s = it % g.stages;
                const uint32_t ph = (it / g.stages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                if (ptid == 0) {                                  // weight chunk for this stage (bulk TMA, async)
                    mbar_expect_tx(&full_b[s], (uint32_t)bbytes);
                    if (g.nsplit) {                               // rows split_ * n_pad .. of the chunk: its hi and lo pieces
                        const unsigned char *src = g.wpacked + (size_t)(c % g.wchunks) * (2 * g.w_n_pad * 128) + (size_t)split_ * g.n_pad * 128;
                        tma_load_1d(b_base + (size_t)s * bbytes, src, (uint32_t)bbytes / 2, &full_b[s]);
                        tma_load_1d(b_base + (size_t)s * bbytes + bbytes / 2, src + (size_t)g.w_n_pad * 128, (uint32_t)bbytes / 2, &full_b[s]);
                    } else {
                        tma_load_1d(b_base + (size_t)s * bbytes, g.wpacked + (size_t)(c % g.wchunks) * bbytes, (uint32_t)bbytes, &full_b[s]);
                    }
                }
                unsigned char *a_hi = a_base + (size_t)s * A_STAGE_BYTES;
                prod.fill(c, a_hi, a_hi + A_PART_BYTES, ptid);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_a[s]);
            }
            }
        }
        }
    } else if (warp < MW) {
        // ================= gather issuers (kIssuerWarps > 0) =================
        if constexpr (IW > 0) {
            const int itid = tid - PW * 32;
            Producer prod(pa, g);
            const int RAW = g.raw_stages;
            const int nsp = g.nsplit ? g.splits : 1;              // (work items = (tile, column block), as in the producers)
            auto tl = [&](int item) { return nsp == 1 ? item : item / nsp; };
            const int tile_step = (int)gridDim.x, ntiles = (int)g.num_tiles * nsp, nchunks = g.num_chunks;
            const int total = ((ntiles - (int)blockIdx.x + tile_step - 1) / tile_step) * nchunks;
            int ah_tile = (int)blockIdx.x, ah_chunk = 0, nx_tile = ah_tile, nx_chunk = 0;
            auto step = [&](int &t, int &c) { if (++c == nchunks) { c = 0; t += tile_step; } };
            step(nx_tile, nx_chunk);
            prod.prime(tl(ah_tile), itid);
            int slot = 0;
            uint32_t eph = 0;                                     // parity of the slot's previous hand-back
            for (int i = 0; i < total; ++i) {
                if (i >= RAW) mbar_wait(&raw_empty[slot], eph);   // the converters are done with this slot's last use
                prod.issue(tl(ah_tile), ah_chunk, nx_tile < ntiles ? tl(nx_tile) : -1, raw_base + (size_t)slot * g.raw_bytes,
                           &raw_full[slot], itid);
                step(ah_tile, ah_chunk);
                step(nx_tile, nx_chunk);
                if (++slot == RAW) { slot = 0; if (i >= RAW) eph ^= 1u; }
            }
        }
    } else if (warp == MW) {
        // ================= MMA issuer (+ epilogue of TMEM lane quarter 0 when merged) =================
        const uint32_t idesc = make_idesc_bf16(TILE_M, g.n_pad);
        const uint32_t a_base_u32 = smem_u32(a_base), b_base_u32 = smem_u32(b_base);
        uint32_t it = 0, tcount = 0, mph = 0;
        int ms = 0;                                               // operand stage / phase of iteration `it`
        const long long work = g.num_tiles * g.splits;
        Epilogue epi;
        long long owed = -1;                                      // merged: work item whose epilogue this warp still owes
        auto run_epilogue = [&](long long w, uint32_t tc) {
            long long tile;
            int split_;
            work_item(w, g.splits, tile, split_);
            const uint32_t acc = tc & nacc_mask;
            mbar_wait(&tmem_full[acc], (tc >> g.nacc_log2) & 1);
            fence_after_sync();
            epi.tile(ea, g, tile, split_, tmem_base + acc * (uint32_t)g.acc_stride, 0, lane);
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        };
        for (long long w = blockIdx.x; w < work; w += gridDim.x, ++tcount) {
            if (g.trace != nullptr && blockIdx.x == 0 && lane == 0 && it < 200) g.trace[it * 16 + 7] = clock64();
            long long tile_;
            int split_;
            work_item(w, g.splits, tile_, split_);
            const int c_begin = g.nsplit ? 0 : split_ * g.chunks_per_split;
            const int c_end = min(g.num_chunks, c_begin + g.chunks_per_split);
            const uint32_t acc = tcount & nacc_mask;
            mbar_wait(&tmem_empty[acc], ((tcount >> g.nacc_log2) & 1) ^ 1);
            if (g.trace != nullptr && blockIdx.x == 0 && lane == 0 && it < 200) g.trace[it * 16 + 15] = clock64();
            fence_after_sync();
            const uint32_t d_addr = tmem_base + acc * (uint32_t)g.acc_stride;
            int wc = c_begin == 0 ? 0 : c_begin % g.wchunks;      // weight chunk of iteration c
            for (int c = c_begin; c < c_end; ++c, ++it) {
                const int s = ms;
                const uint32_t ph = mph;
                if (++ms == g.stages) { ms = 0; mph ^= 1u; }
                const bool trm = g.trace != nullptr && blockIdx.x == 0 && lane == 0 && it < 200;
                if (trm) g.trace[it * 16 + 8] = clock64();
                mbar_wait(&full_a[s], ph);
                if (trm) g.trace[it * 16 + 9] = clock64();
                if (!(Producer::kAsync && g.wchunks == 1 && !g.nsplit && it >= (uint32_t)g.stages)) mbar_wait(&full_b[s], ph);   // (resident weights)
                if (trm) g.trace[it * 16 + 10] = clock64();
                fence_after_sync();
                {
                    // Everything here is warp-uniform and computed by ALL lanes, only the tcgen05 instructions sit
                    // under the elect predicate: descriptors then live in uniform registers.  (Issuing from inside an
                    // `if (lane == 0)` block made ptxas wrap every MMA in an ELECT + 7x R2UR waterfall loop: ~155
                    // cycles of issue per MMA, twice the MMA's own 64 cycles.)
                    const uint32_t a_hi = a_base_u32 + (uint32_t)s * A_STAGE_BYTES;
                    const uint32_t a_lo = a_hi + A_PART_BYTES;
                    const uint32_t b_hi = b_base_u32 + (uint32_t)s * (uint32_t)bbytes;
                    const uint32_t b_lo = b_hi + (uint32_t)g.n_pad * 128u;
                    const int ksteps = min(CHUNK_K / UMMA_K, (g.k_total - wc * CHUNK_K) / UMMA_K);
                    const bool leader = elect_one();
                    for (int kk = 0; kk < ksteps; ++kk) {
                        const uint32_t off = (uint32_t)kk * (UMMA_K * 2);
                        const uint64_t dah = make_smem_desc_sw128(a_hi + off), dal = make_smem_desc_sw128(a_lo + off);
                        const uint64_t dbh = make_smem_desc_sw128(b_hi + off), dbl = make_smem_desc_sw128(b_lo + off);
                        if (leader) {
                            umma_bf16(d_addr, dah, dbh, idesc, (c != c_begin) || (kk != 0));
                            umma_bf16(d_addr, dah, dbl, idesc, 1);
                            umma_bf16(d_addr, dal, dbh, idesc, 1);
                        }
                    }
                    if (leader) {
                        umma_commit(&empty[s]);                   // frees the smem stage when the MMAs retire
                        if (c == c_end - 1) umma_commit(&tmem_full[acc]);
                        if (trm) g.trace[it * 16 + 11] = clock64();
                    }
                    if (++wc == g.wchunks) wc = 0;
                }
                __syncwarp();
                if constexpr (MG) {
                    // the previous tile's accumulator completes while the producers fill this tile's second chunk:
                    // drain this warp's quarter of it now (the MMAs just issued keep the tensor pipe busy meanwhile)
                    if (c == c_begin && owed >= 0) { run_epilogue(owed, tcount - 1); owed = -1; }
                }
            }
            if constexpr (MG) owed = w;
        }
        if constexpr (MG)
            if (owed >= 0) run_epilogue(owed, tcount - 1);
    } else {
        // ================= epilogue =================
        const int quarter = warp & 3;                             // TMEM lane quarter this warp may read
        Epilogue epi;
        uint32_t tcount = 0;
        const long long work = g.num_tiles * g.splits;
        for (long long w = blockIdx.x; w < work; w += gridDim.x, ++tcount) {
            long long tile;
            int split_;
            work_item(w, g.splits, tile, split_);
            const uint32_t acc = tcount & nacc_mask;
            const bool tre = g.trace != nullptr && blockIdx.x == 0 && warp == MW + 1 && lane == 0 && tcount < 200;
            if (tre) g.trace[tcount * 16 + 12] = clock64();
            mbar_wait_poll<relaxed_waits<Producer>::value, 64>(&tmem_full[acc], (tcount >> g.nacc_log2) & 1);
            if (tre) g.trace[tcount * 16 + 13] = clock64();
            fence_after_sync();
            const uint32_t t_acc = tmem_base + acc * (uint32_t)g.acc_stride + ((uint32_t)(quarter * 32) << 16);
            epi.tile(ea, g, tile, split_, t_acc, EW > 4 ? quarter + 4 * ((warp - MW - 1) >> 2) : quarter, lane);
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (tre) g.trace[tcount * 16 + 14] = clock64();
        }
    }

    fence_before_sync();
    __syncthreads();
    if (warp == MW) {
        fence_after_sync();
        tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
    }
}

// Launch with PROGRAMMATIC DEPENDENT LAUNCH (kdpc_tc_set_pdl(1); default off - measured neutral): the kernel's CTAs may be scheduled while the
// previous kernel of the stream is still draining, run their set-up (mbarrier init, TMEM allocation, tensor-map
// prefetch, role split) and wait at `griddepcontrol.wait` - placed in tc_gemm_kernel right before the first access to
// global memory that a predecessor may have written - until that kernel has completed and flushed.  Stream capture
// records the attribute as a programmatic edge, so the ~150 launches of a graphed forward lose most of their
// launch-to-launch gaps.  Safe behind any predecessor (it need not trigger anything: its completion releases the wait).
extern "C" int kdpc_tc_pdl_enabled(void);
template <class Kern, class... Args>
static inline void launch_tc(Kern kern, unsigned grid, unsigned threads, size_t smem, cudaStream_t st, const Args &...args) {
    if (!kdpc_tc_pdl_enabled()) {
        kern<<<grid, threads, smem, st>>>(args...);
        return;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, args...);
}

// ------------------------------------------------------------------------------------------
// Epilogue: y = act(acc * scale[n] + shift[n]) stored row-major fp32 [M, ldo].
// (Linear bias, and eval-mode BatchNorm folded on the host: scale = gamma/sqrt(var+eps),
//  shift = (bias - mean) * scale + beta.)
struct StoreEpilogue {
    struct Args {
        const float *scale;   // [n] or nullptr (= 1)
        const float *shift;   // [n] or nullptr (= 0)
        float slope;          // LeakyReLU slope; 1.0 = no activation
        float lo, hi;         // clamp range (applied after the activation); lo > hi = none
        const float *residual;   // optional [M, ldo] added last (flow = flow_local + up_flow)
        float *out;
        int ldo;
        float *partial;          // split-K workspace [splits][M][n_pad] (g.splits > 1): raw sums, epilogue applied by the reducer
        // optional row order (never with split-K): tile row p = b * rows_per_cloud + i is output row
        // b * rows_per_cloud + row_order[b * order_stride + i] (the producer processed that row at position p)
        const int *row_order = nullptr;
        int order_stride = 0;
        int rows_per_cloud = 0;
    };
    // tcgen05.ld hands each lane ONE ROW of the 32 x 32 block (v[j] = column j).  Storing that directly makes every
    // store instruction touch 32 different rows, 16 bytes each: measured 4x slower than the whole rest of the kernel
    // (partial-sector writes, 32 transactions per instruction).  So the block is transposed through a small
    // per-warp shared-memory tile and stored with each instruction covering 4 rows x 128 contiguous bytes; the
    // affine / activation / clamp / residual are applied after the transpose, where a lane owns 4 fixed columns.
    static constexpr int TP = 36;                             // floats per staged row (32 + 4): conflict-free 16-byte accesses
    __device__ __forceinline__ void tile(const Args &e, const GemmShape &g, long long tile, int split, uint32_t t_acc,
                                         int quarter, int lane) const {
        if (g.nsplit) {                                           // column block `split`: the same epilogue on shifted pointers
            const int n0 = split * g.n_pad;
            Args e2 = e;
            e2.out += n0;
            if (e2.scale) e2.scale += n0;
            if (e2.shift) e2.shift += n0;
            if (e2.residual) e2.residual += n0;
            tile_impl(e2, g, g.n - n0 < g.n_pad ? g.n - n0 : g.n_pad, false, tile, 0, t_acc, quarter, lane);
            return;
        }
        tile_impl(e, g, g.n, g.splits > 1, tile, split, t_acc, quarter, lane);
    }
    __device__ __forceinline__ void tile_impl(const Args &e, const GemmShape &g, const int gn, const bool partial, long long tile,
                                              int split, uint32_t t_acc, int quarter, int lane) const {
        __shared__ __align__(16) float stage[4][32 * TP];
        float *st = stage[quarter];
        const long long row0 = tile * TILE_M + quarter * 32;
        float *obase = partial ? e.partial + (size_t)split * g.m * g.n_pad : e.out;
        const int ld = partial ? g.n_pad : e.ldo;
        const int ncols = partial ? g.n_pad : gn;
        const bool vec = (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(obase) & 15) == 0 &&
                         (e.residual == nullptr || partial || (reinterpret_cast<uintptr_t>(e.residual) & 15) == 0);
        const int rsub = lane >> 3, c4 = (lane & 7) * 4;     // this lane's row within a group of 4, and its 4 columns
        // Fast path (every streaming layer of the model): whole 32-row block inside M, plain row order, 16-byte aligned
        // rows, no residual.  Branch-free and fully unrolled: the generic path below carries a 64-bit division, eight
        // predicated scalar scale/shift loads and ~10 branches per row group and cost ~2700 cycles per 32 x 32 block
        // (tools/trace_linear.py: the epilogue, not HBM, paced the streaming layers at 26-34 % of the copy peak).
        if (!partial && (e.row_order == nullptr || e.rows_per_cloud >= 32) && e.residual == nullptr && vec && (gn & 3) == 0 &&
            row0 + 32 <= g.m && ((reinterpret_cast<uintptr_t>(e.scale) | reinterpret_cast<uintptr_t>(e.shift)) & 15) == 0) {
            const bool clampd = e.lo <= e.hi;
            // destination of this lane's 8 rows (i * 4 + rsub): plain order, or through the row order (the fused PointConv
            // processes its rows in Morton order; one division per block, the block may straddle one cloud boundary)
            float *orow[8];
            if (e.row_order == nullptr) {
#pragma unroll
                for (int i = 0; i < 8; ++i) orow[i] = e.out + (row0 + rsub + i * 4) * (long long)ld + c4;
            } else {
                const long long b0 = (long long)((unsigned)row0 / (unsigned)e.rows_per_cloud);   // rows < 2^31 (launcher shapes)
                const int p0 = (int)(row0 - b0 * e.rows_per_cloud) + rsub;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    int pp = p0 + i * 4;
                    long long b = b0;
                    if (pp >= e.rows_per_cloud) { pp -= e.rows_per_cloud; ++b; }
                    const long long row = b * e.rows_per_cloud + __ldg(e.row_order + b * e.order_stride + pp);
                    orow[i] = e.out + row * (long long)ld + c4;
                }
            }
            for (int c0 = 0; c0 < g.n_pad; c0 += 32) {
                const int col = c0 + c4;
                const bool live = col < gn;                      // (n % 4 == 0: a lane's 4 columns are all in or all out)
                float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live && e.scale) sc = __ldg(reinterpret_cast<const float4 *>(e.scale + col));   // in flight during the TMEM load
                if (live && e.shift) sh = __ldg(reinterpret_cast<const float4 *>(e.shift + col));
                float v[32];
                tmem_ld_32x32(t_acc + (uint32_t)c0, v);           // warp-collective: no divergence around it
                if (c0 >= gn) continue;
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4 *>(st + lane * TP + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                __syncwarp();
                if (live) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 t = *reinterpret_cast<const float4 *>(st + (i * 4 + rsub) * TP + c4);
                        float4 y;
                        y.x = t.x * sc.x + sh.x; y.y = t.y * sc.y + sh.y; y.z = t.z * sc.z + sh.z; y.w = t.w * sc.w + sh.w;
                        y.x = y.x > 0.f ? y.x : y.x * e.slope; y.y = y.y > 0.f ? y.y : y.y * e.slope;
                        y.z = y.z > 0.f ? y.z : y.z * e.slope; y.w = y.w > 0.f ? y.w : y.w * e.slope;
                        if (clampd) {
                            y.x = fminf(fmaxf(y.x, e.lo), e.hi); y.y = fminf(fmaxf(y.y, e.lo), e.hi);
                            y.z = fminf(fmaxf(y.z, e.lo), e.hi); y.w = fminf(fmaxf(y.w, e.lo), e.hi);
                        }
                        *reinterpret_cast<float4 *>(orow[i] + c0) = y;
                    }
                }
            }
            return;
        }
        for (int c0 = 0; c0 < g.n_pad; c0 += 32) {
            float v[32];
            tmem_ld_32x32(t_acc + (uint32_t)c0, v);               // warp-collective: no divergence around it
            if (c0 >= ncols) continue;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4 *>(st + lane * TP + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            __syncwarp();
            const int col = c0 + c4;
            float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
            if (!partial) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (col + j < gn) {
                        if (e.scale) sc[j] = __ldg(e.scale + col + j);
                        if (e.shift) sh[j] = __ldg(e.shift + col + j);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = i * 4 + rsub;
                const long long prow = row0 + r;
                const float4 t = *reinterpret_cast<const float4 *>(st + r * TP + c4);
                float y[4] = {t.x, t.y, t.z, t.w};
                if (prow < g.m && col < ncols) {
                    long long row = prow;
                    if (e.row_order != nullptr) {
                        const long long b = prow / e.rows_per_cloud;
                        row = b * e.rows_per_cloud + __ldg(e.row_order + b * e.order_stride + (prow - b * e.rows_per_cloud));
                    }
                    float *o = obase + row * ld + col;
                    if (!partial) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float z = y[j] * sc[j] + sh[j];
                            z = z > 0.f ? z : z * e.slope;
                            if (e.lo <= e.hi) z = fminf(fmaxf(z, e.lo), e.hi);
                            y[j] = z;
                        }
                        if (e.residual) {
                            const float *rp = e.residual + row * ld + col;
                            if (vec && col + 4 <= ncols) {
                                const float4 rr = __ldg(reinterpret_cast<const float4 *>(rp));
                                y[0] += rr.x; y[1] += rr.y; y[2] += rr.z; y[3] += rr.w;
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (col + j < ncols) y[j] += __ldg(rp + j);
                            }
                        }
                    }
                    if (vec && col + 4 <= ncols) {
                        *reinterpret_cast<float4 *>(o) = make_float4(y[0], y[1], y[2], y[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (col + j < ncols) o[j] = y[j];
                    }
                }
            }
        }
    }
};

// out[row, col] = epilogue(sum_s partial[s][row][col]), splits added in index order
static __global__ void __launch_bounds__(256)
splitk_reduce_kernel(long long m, int n, int n_pad, int splits, const StoreEpilogue::Args e) {
    const int ng = (n + 3) >> 2;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * ng) return;
    const long long row = t / ng;
    const int c0 = (int)(t - row * ng) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
        const float4 v = *reinterpret_cast<const float4 *>(e.partial + ((size_t)s * m + row) * n_pad + c0);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const float a[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int col = c0 + j;
        if (col < n) {
            float y = a[j];
            if (e.scale) y *= __ldg(e.scale + col);
            if (e.shift) y += __ldg(e.shift + col);
            y = y > 0.f ? y : y * e.slope;
            if (e.lo <= e.hi) y = fminf(fmaxf(y, e.lo), e.hi);
            if (e.residual) y += __ldg(e.residual + row * e.ldo + col);
            e.out[row * e.ldo + col] = y;
        }
    }
}

static inline int launch_splitk_reduce(const GemmShape &g, const StoreEpilogue::Args &e, cudaStream_t st) {
    const long long total = g.m * ((g.n + 3) / 4);
    splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g.m, g.n, g.n_pad, g.splits, e);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Producer: plain fp32 rows x[M, ldx] (K contiguous).
struct PlainProducer {
    static constexpr int kWarps = 8, kGroups = 2;        // two groups of 4 warps alternate K-chunks
    static constexpr bool kAsync = false;
    static constexpr int kIssuers = 0, kLookahead = 0;
    struct Args {
        const float *x;
        int ldx;
        int k;
    };
    static __device__ __forceinline__ void prologue(const Args &, int, int) {}
    const Args &a;
    const GemmShape &g;
    const float *rowp;
    bool valid;
    __device__ PlainProducer(const Args &a_, const GemmShape &g_) : a(a_), g(g_), rowp(nullptr), valid(false) {}
    __device__ __forceinline__ void begin_tile(long long tile, int r) {
        const long long row = tile * TILE_M + r;
        valid = row < g.m;
        rowp = a.x + (valid ? row : 0) * a.ldx;
    }
    __device__ __forceinline__ void fill(int chunk, unsigned char *a_hi, unsigned char *a_lo, int r) {
        const int k0 = chunk * CHUNK_K;
        const bool fast = valid && (a.ldx & 3) == 0 && k0 + CHUNK_K <= a.k;
        float4 ld[16];
        if (fast) {
            const float4 *p4 = reinterpret_cast<const float4 *>(rowp + k0);
#pragma unroll
            for (int i = 0; i < 16; ++i) ld[i] = __ldg(p4 + i);
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float t[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = k0 + i * 4 + j;
                    t[j] = (valid && k < a.k) ? __ldg(rowp + k) : 0.f;
                }
                ld[i] = make_float4(t[0], t[1], t[2], t[3]);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float v[8] = {ld[2 * u].x, ld[2 * u].y, ld[2 * u].z, ld[2 * u].w,
                                ld[2 * u + 1].x, ld[2 * u + 1].y, ld[2 * u + 1].z, ld[2 * u + 1].w};
            uint4 hi, lo;
            split8(v, hi, lo);
            const uint32_t off = sw128_offset(r, u);
            *reinterpret_cast<uint4 *>(a_hi + off) = hi;
            *reinterpret_cast<uint4 *>(a_lo + off) = lo;
        }
    }
};


// ------------------------------------------------------------------------------------------
// Asynchronous producer for plain fp32 rows (K % 8 == 0, ldx % 4 == 0): the 128 x 64 chunk is fetched with
// 16-byte cp.async one or two pipeline iterations ahead (no registers hold data in flight, so the HBM latency
// is off the critical path), then converted fp32 -> bf16 hi/lo from shared memory.  Thread (row, half) owns 32
// of the 64 columns.
struct PlainAsyncProducer {
    static constexpr int kWarps = 8, kGroups = 1;
    static constexpr bool kAsync = true;
    static constexpr int kIssuerWarps = 0;                   // dedicated issue warps: measured neutral for these streaming rows (4 tried)
    static constexpr int kIssueThreads = kIssuerWarps > 0 ? 32 * kIssuerWarps : 256;
    // kBulkRows: every row's chunk slice (<= 256 contiguous bytes) travels as ONE 1-D bulk TMA copy (cp.async.bulk,
    // completion by mbarrier transaction bytes) issued by thread r of the first 128 producer threads: 4 warp instructions
    // per chunk instead of 64 LDGSTS requests (tools/trace_linear.py: ~1100 of a chunk's ~3500 producer cycles were
    // the load/store unit working through those requests).  Needs 16-byte aligned rows (ldx % 4 == 0, checked by the
    // launcher).  raw_full[] then counts ONE arrival (the expect_tx of producer thread 0).
    // MEASURED AND OFF (round 2): 128 per-row bulk copies take ~2300 cycles to issue (the TMA unit accepts roughly
    // one request per 17 cycles whichever warp sends it) against ~1100 for the 64 LDGSTS requests: 25.2 us instead of
    // 23.2 us at M = 65536, K = N = 128.  One request per chunk needs a 2-D tensor map (UTMALDG), not per-row copies.
    static constexpr bool kBulkRows = false;
    static constexpr int kIssuers = kBulkRows ? 1 : kIssueThreads, kLookahead = 2;
    static constexpr int ROW_PITCH = 272;                    // 256 B payload + 16 B: conflict-free 16-byte reads by row
    static constexpr int kRawBytes = TILE_M * ROW_PITCH;
    using Args = PlainProducer::Args;
    static __device__ __forceinline__ void prologue(const Args &, int, int) {}
    const Args &a;
    const GemmShape &g;
    __device__ PlainAsyncProducer(const Args &a_, const GemmShape &g_) : a(a_), g(g_) {}
    __device__ __forceinline__ void prime(int, int) {}
    // 16 consecutive lanes copy the (up to) 256 contiguous bytes of ONE row's chunk slice, thread t serves rows
    // (t >> 4) + RSTEP j: a warp-level LDGSTS touches 4 lines.  (One thread per row-half copying its pieces one after the
    // other made every request touch 32 lines - the load/store unit then needs ~28 cycles per request, see costvol_tc.cu.)
    __device__ __forceinline__ void issue(int tile, int chunk, int /*next_tile*/, unsigned char *raw, uint64_t *bar, int ptid) {
        const int k0 = chunk * CHUNK_K;
        const int pieces = max(0, min(CHUNK_K, a.k - k0)) >> 2;
        const long long row0 = (long long)tile * TILE_M;
        if constexpr (kBulkRows) {
            const uint32_t bytes = (uint32_t)pieces * 16u;
            if (ptid == 0) mbar_expect_tx(bar, bytes * TILE_M);
            if (ptid < TILE_M && bytes != 0) {
                long long row = row0 + ptid;
                if (row >= g.m) row = g.m - 1;               // padded rows repeat the last row; never stored
                tma_load_1d(raw + ptid * ROW_PITCH, a.x + row * a.ldx + k0, bytes, bar);
            }
            return;
        }
        const int q = ptid & 15, rb = ptid >> 4;
        if (q < pieces) {
            constexpr int RSTEP = kIssueThreads / 16;        // rows covered by one pass of the issuing threads
#pragma unroll
            for (int j = 0; j < TILE_M / RSTEP; ++j) {
                const int r = rb + RSTEP * j;
                long long row = row0 + r;
                if (row >= g.m) row = g.m - 1;               // padded rows repeat the last row; never stored
                cp_async_16(smem_u32(raw + r * ROW_PITCH + q * 16), a.x + row * a.ldx + k0 + q * 4);
            }
        }
        cp_async_mbar_arrive(bar);
    }
    __device__ __forceinline__ void convert(int /*tile*/, int chunk, const unsigned char *raw, unsigned char *a_hi,
                                            unsigned char *a_lo, int ptid) {
        const int r = ptid & 127, half = ptid >> 7;
        const float4 *rr = reinterpret_cast<const float4 *>(raw + r * ROW_PITCH);
        const int units = min(8, (a.k - chunk * CHUNK_K) >> 3);          // k % 8 == 0
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
            const int u = half * 4 + uu;
            uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
            if (u < units) {
                const float4 g0 = rr[2 * u], g1 = rr[2 * u + 1];
                const float v[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                split8(v, hi, lo);
            }
            if (u <= units) {                                // unit == units: the zero half of an odd last 16-wide K-step
                const uint32_t off = sw128_offset(r, u);
                *reinterpret_cast<uint4 *>(a_hi + off) = hi;
                *reinterpret_cast<uint4 *>(a_lo + off) = lo;
            }
        }
    }
};


// ------------------------------------------------------------------------------------------
// Plain fp32 rows fetched by 2-D TENSOR-MAP TMA (cp.async.bulk.tensor.2d -> SASS UTMALDG): ONE elected thread issues two
// box copies per 128 x 64 chunk (columns c0 .. c0+31 and c0+32 .. c0+63: a SWIZZLE_128B box is at most 128 bytes wide),
// instead of 64 warp-level LDGSTS requests (~1100 cycles of load/store-unit time per chunk, tools/trace_linear.py) or
// 128 per-row bulk copies (~2300 cycles of TMA issue).  Rows beyond M and columns beyond K are zero-filled by the TMA
// unit; the boxes land densely (128 B per row, 16-byte units XOR-swizzled by row & 7 - the same pattern the operand
// tiles use), so the converting threads read them conflict-free without padding.
struct PlainTmaProducer {
    static constexpr int kWarps = 8, kGroups = 1;
    static constexpr bool kAsync = true;
    static constexpr int kIssuerWarps = 0;
    static constexpr int kIssuers = 1, kLookahead = 2;       // raw_full[]: the expect_tx arrival of producer thread 0
    static constexpr int BOX_COLS = 32, BOX_BYTES = TILE_M * 128;
    static constexpr int kRawBytes = 2 * BOX_BYTES;
    struct alignas(64) Args {
        CUtensorMap tmap;        // [M, K] fp32, row pitch ldx * 4 bytes, box 32 x 128, SWIZZLE_128B, zero fill
        int k;
    };
    static __device__ __forceinline__ void prologue(const Args &a, int tid, int) {
        if (tid == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.tmap)) : "memory");
    }
    const Args &a;
    const GemmShape &g;
    __device__ PlainTmaProducer(const Args &a_, const GemmShape &g_) : a(a_), g(g_) {}
    __device__ __forceinline__ void prime(int, int) {}
    static __device__ __forceinline__ void tma_box(void *smem_dst, const CUtensorMap *tm, int col, int row, uint64_t *bar) {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(col), "r"(row) : "memory");
    }
    __device__ __forceinline__ void issue(int tile, int chunk, int /*next_tile*/, unsigned char *raw, uint64_t *bar, int ptid) {
        if (ptid != 0) return;
        const int c0 = chunk * CHUNK_K;
        const int boxes = (a.k - c0 > BOX_COLS) ? 2 : 1;
        mbar_expect_tx(bar, (uint32_t)(boxes * BOX_BYTES));
        tma_box(raw, &a.tmap, c0, tile * TILE_M, bar);
        if (boxes == 2) tma_box(raw + BOX_BYTES, &a.tmap, c0 + BOX_COLS, tile * TILE_M, bar);
    }
    __device__ __forceinline__ void convert(int /*tile*/, int chunk, const unsigned char *raw, unsigned char *a_hi,
                                            unsigned char *a_lo, int ptid) {
        const int r = ptid & 127, half = ptid >> 7;
        const int units = min(8, (a.k - chunk * CHUNK_K) >> 3);          // k % 8 == 0
        const unsigned char *row = raw + r * 128;
        const int sw = r & 7;
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
            const int u = half * 4 + uu;
            uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
            if (u < units) {
                const unsigned char *box = row + (u >> 2) * BOX_BYTES;   // unit u = floats 8u .. 8u+7 = pieces 2u, 2u+1 of its box
                const int p0 = (2 * u) & 7;
                const float4 g0 = *reinterpret_cast<const float4 *>(box + ((p0 ^ sw) << 4));
                const float4 g1 = *reinterpret_cast<const float4 *>(box + (((p0 + 1) ^ sw) << 4));
                const float v[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                split8(v, hi, lo);
            }
            if (u <= units) {                                // unit == units: the zero half of an odd last 16-wide K-step
                const uint32_t off = sw128_offset(r, u);
                *reinterpret_cast<uint4 *>(a_hi + off) = hi;
                *reinterpret_cast<uint4 *>(a_lo + off) = lo;
            }
        }
    }
};

}  // namespace tc
}  // namespace kdpc
