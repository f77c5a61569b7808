// select.cuh -- CUB-free cluster radix-select + in-smem bitonic sort ("top-k, sorted").
//
// One thread-block CLUSTER of kClusterSize CTAs serves one segment (an (image, level) score map for
// a3, an (image, pos|neg) candidate set for the a8 samplers).  Each CTA owns a contiguous slice of the
// segment, turns every element into a 64-bit composite
//        comp = (key32 << 32) | ~index         (larger comp == better; ties -> lower index first)
// caches key32 in shared memory on the first pass, and the cluster then runs an MSB-first radix
// select over the composite with 8-bit digits:
//   * per-CTA histogram  hist[256][32]  is replicated per lane (bank == lane -> conflict-free, and no
//     same-address serialisation inside a warp even when all scores share an exponent),
//   * per-CTA totals are exchanged through DISTRIBUTED SHARED MEMORY (cluster.map_shared_rank),
//   * the pass loop stops as soon as the threshold bin holds exactly the number still needed
//     (normally after the 4 key digits; the index digits only run when a tie straddles K).
// Every CTA then collects ITS winners into its own shared memory (two sweeps over the cached keys: count, block scan,
// write), sorts them (shuffle-based bitonic sort, ~K/8 entries per CTA), and after one cluster barrier pulls the other
// CTAs' sorted lists over DSMEM: a composite's final rank is its position in its own list plus the lower bounds in
// the other seven, so the owner calls Sink::emit(seg, rank, comp) itself -- sort, ranking and emit are spread over the
// cluster.  Segments with N <= kSelSoloMax and K <= kSelSoloMaxK are served by the leader alone ("solo": the other
// CTAs exit at once, no cluster barrier); segments with N <= kSelDirectMax skip the radix passes (one sort).
// Pass 0 of Srcs with raw_ptr()/from_raw() pulls the whole slice in with cp.async first.  -DMD_SEL_TIMING builds print
// the cycles of every phase of the first cluster's leader.
//
// Semantics == oracle o_topk / o_sample (oracle/CONVENTIONS.md #4, #13).  Reference idiom being
// replaced: ops.TopK(sorted=True) at centerpoint/det3d_ms/models/bbox_heads/center_head.py:435 and
// pointpillars/src/pointpillars.py:764.
#pragma once
#include <cooperative_groups.h>
#ifdef MD_SEL_TIMING
#include <cstdio>
#endif

#include "common.cuh"
#include "launch.cuh"

namespace md {
namespace cg = cooperative_groups;

#ifndef MD_SEL_CLUSTER
#define MD_SEL_CLUSTER 8
#endif
constexpr int kClusterSize = MD_SEL_CLUSTER;    // > 8 is a non-portable cluster size (opt-in below); measured with 16:
                                                // 90 us vs 66 us for the 40 proposal segments (fewer clusters resident)
constexpr int kSelThreads = 512;
constexpr int kSelMaxK = 2048;            // sorted output limit per segment
constexpr int kSelDirectMax = 1024;       // segments up to this length skip the radix passes: the leader sorts them all
                                          // (a single-CTA bitonic sort of 4096 costs ~50 us: measured, so keep this small)
constexpr int kSelSoloMax = 16384;       // segments up to this length are served by one CTA of the cluster ("solo") ...
constexpr int kSelSoloMaxK = 512;        // ... when K is small: sorting and emitting 2000 winners is worth spreading over 8 CTAs
constexpr int kSelMaxIndexBits = 22;      // segment length < 4 Mi elements
constexpr int kSelMaxCacheElems = 44 * 1024; // key cache per CTA (dynamic smem)

struct alignas(16) SelShared {
    union {
        uint32_t hist[256 * 32];            // [bin][lane]  (select passes)
        struct {
            unsigned long long all[kSelMaxK];   // every CTA: the cluster's selected composites, list after list (ranking);
                                                // leader of a short segment: the whole segment
            unsigned long long mine[kSelMaxK];  // this CTA's selected composites, sorted (read remotely through DSMEM)
        } lists;
    };
    uint32_t local[2][256];    // this CTA's per-bin totals (read remotely through DSMEM); double-buffered per pass
    uint32_t tot[256];         // cluster totals
    uint32_t mine_count;       // entries of lists.mine (read remotely)
    uint32_t cnt[kClusterSize];
    uint32_t digit, above, eq, total, found;
};
static_assert(sizeof(SelShared) % 16 == 0, "the key cache behind SelShared is filled with 16-byte cp.async");

// Optional Src extension ("raw prefetch"): when the keys of a slice are a pure function of one contiguous 32-bit word
// per element, the Src exposes
//   __device__ const uint32_t *raw_ptr(const Ctx&, int m) const;                       address of element m's word
//   __device__ bool from_raw(const Ctx&, int m, uint32_t raw, uint32_t &key) const;    == load()
// and pass 0 pulls the whole slice into the key cache with cp.async (every load of the slice in flight at once)
// instead of UNR loads per thread per round trip.
template <class Src, class = void> struct SrcHasRaw { static constexpr bool value = false; };
template <class Src> struct SrcHasRaw<Src, decltype((void)&Src::raw_ptr)> { static constexpr bool value = true; };

// Optional Sink extension: `Pre prefetch(seg, comp)` starts the global loads emit needs (they depend on the composite
// only) before the ranking, `emit_pre(seg, rank, comp, pre)` finishes the job.
template <class Sink, class = void> struct SinkHasPre { static constexpr bool value = false; };
template <class Sink> struct SinkHasPre<Sink, decltype((void)&Sink::prefetch)> { static constexpr bool value = true; };

MD_DEVINL void sel_cp_async(void *smem, const void *gmem, bool wide)
{
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    if (wide) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gmem) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(sa), "l"(gmem) : "memory");
}

// Bitonic sort (descending) of n2 (power of two, >= 32, <= kSelMaxK) 64-bit values in shared memory by one CTA of
// kSelThreads.  Element e lives in lane e & 31, so every compare-exchange distance below 32 is a warp shuffle on
// register copies: only the distances >= 32 go through shared memory with a block barrier (21 instead of 66 barriers
// for 2048 values).
MD_DEVINL void bitonic_sort_desc(unsigned long long *v, int n2)
{
    constexpr int R = kSelMaxK / kSelThreads;           // values per thread
    const int tid = threadIdx.x, lane = tid & 31;
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j >= 32; j >>= 1) {
            for (int i = tid; i < n2; i += kSelThreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = v[i], b = v[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { v[i] = b; v[ixj] = a; }
                }
            }
            __syncthreads();
        }
        if (k >= 64 || k == 32) {
            // all remaining distances (< 32) of this k, plus -- for k == 32 -- the whole k = 2..32 prefix, in registers
            // (m * kSelThreads < n2 is uniform over the CTA: short lists skip the register slots they do not fill)
            unsigned long long r[R];
#pragma unroll
            for (int m = 0; m < R; m++) r[m] = (tid + m * kSelThreads < n2) ? v[tid + m * kSelThreads] : 0ull;
            for (int kk = (k == 32 ? 2 : k); kk <= k; kk <<= 1) {
                for (int j = min(kk >> 1, 16); j > 0; j >>= 1) {
#pragma unroll
                    for (int m = 0; m < R; m++) {
                        if (m * kSelThreads >= n2) break;
                        const int i = tid + m * kSelThreads;
                        const unsigned long long o = __shfl_xor_sync(0xffffffffu, r[m], j);
                        const bool desc = (i & kk) == 0, lower = (lane & j) == 0;
                        // the lower index keeps the larger value when descending
                        const bool take_max = desc == lower;
                        r[m] = take_max ? (r[m] > o ? r[m] : o) : (r[m] < o ? r[m] : o);
                    }
                }
            }
#pragma unroll
            for (int m = 0; m < R; m++) if (tid + m * kSelThreads < n2) v[tid + m * kSelThreads] = r[m];
            __syncthreads();
        }
    }
}

// Src concept:
//   __device__ int segment_of(int cluster, int it) const;          it-th segment served by this cluster, -1 = no more
//                                                                    (a cluster works through its segments in turn; give
//                                                                    the clusters similar totals and the longest first)
//   struct Ctx;  __device__ Ctx prepare(int seg) const;            per-segment constants (hoisted out of the loops)
//   __device__ bool active(const Ctx&) const;                       false = this launch leaves the segment untouched
//   __device__ int  length(const Ctx&) const;                       elements in the segment (memory order)
//   __device__ int  want(const Ctx&) const;                         K requested (<= kSelMaxK)
//   __device__ bool load(const Ctx&, int m, uint32_t &key) const;   false = not a candidate
//   __device__ uint32_t index_of(const Ctx&, int m) const;          logical index of memory position m (only evaluated
//                                                                    for tie digits and for the selected elements)
// Sink concept:
//   __device__ void emit(int seg, int rank, unsigned long long comp) const;
//   __device__ void pad(int seg, int rank) const;                 ranks >= #selected, < want(seg)
//   __device__ void finish(int seg, int selected, int candidates) const;   once per segment (thread 0)
template <class Src, class Sink>
__device__ void select_segment(const Src &src, const Sink &sink, const int cache_elems, const int seg,
                               SelShared &sh, uint32_t *keys, uint32_t *vbits, cg::cluster_group &cluster, const int rank)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const typename Src::Ctx ctx = src.prepare(seg);
    if (!src.active(ctx)) return;            // uniform over the cluster: nobody reaches a cluster barrier

    const int N = src.length(ctx);
    const int K = min(src.want(ctx), kSelMaxK);
    // slice (multiple of 32 so validity words never straddle CTAs)
    // solo: a segment this short is served by the leader CTA alone -- the other seven exit at once and free their SMs
    // for the next cluster, and the leader skips every cluster barrier and DSMEM exchange (for a 12,600-element level
    // those fixed costs were most of the time)
    const bool solo = N <= kSelSoloMax && (K <= kSelSoloMaxK || N <= kSelDirectMax);
    if (solo && rank != 0) return;               // uniform over the cluster; the leader touches no distributed memory
    const int nranks = solo ? 1 : kClusterSize;
    int per = (N + nranks - 1) / nranks;
    per = (per + 31) & ~31;
    const int begin = min(N, rank * per), end = min(N, begin + per);
    const int len = end - begin;
    const bool cached = len <= cache_elems;

    if (tid == 0) sh.mine_count = 0;
#ifdef MD_SEL_TIMING
    long long t_ph[20]; int n_ph = 0;
#define MD_STAMP() do { if (n_ph < 20) t_ph[n_ph++] = clock64(); } while (0)
#else
#define MD_STAMP() do {} while (0)
#endif
    MD_STAMP();

    if (N <= kSelDirectMax) {     // (always solo)
        // ---- short segment: no radix passes.  The leader gathers every candidate, sorts them all and emits the K best.
        __syncthreads();
        for (int base = 0; base < N; base += kSelThreads) {
            const int i = base + tid;
            uint32_t key = 0u;
            const bool ok = i < N && src.load(ctx, i, key);
            const uint32_t m = __ballot_sync(0xffffffffu, ok);
            if (m) {
                uint32_t pos = 0;
                const int leader = __ffs(m) - 1;
                if (lane == leader) pos = atomicAdd(&sh.mine_count, (uint32_t)__popc(m));
                pos = __shfl_sync(0xffffffffu, pos, leader);
                if (ok) sh.lists.all[pos + __popc(m & ((1u << lane) - 1u))] = ((unsigned long long)key << 32) | (uint32_t)~src.index_of(ctx, i);
            }
        }
        __syncthreads();
        const int ncand = (int)sh.mine_count;
        int n2 = 32;
        while (n2 < ncand) n2 <<= 1;
        for (int i = ncand + tid; i < n2; i += kSelThreads) sh.lists.all[i] = 0ull;
        __syncthreads();
        bitonic_sort_desc(sh.lists.all, n2);
        const int sel = min(K, ncand), want = src.want(ctx);
        for (int i = tid; i < want; i += kSelThreads) {
            if (i < sel) sink.emit(seg, i, sh.lists.all[i]);
            else sink.pad(seg, i);
        }
        if (tid == 0) sink.finish(seg, sel, ncand);
        return;
    }

    // composite = (key << 32) | ~index.  Every composite has low bits 31..22 set (index < 2^22), so those
    // bits are "already matched".  The 4 key digits never need the index; it is only computed for the tie
    // digits (passes 4..6, when a tie straddles K) and for the selected elements.
    uint32_t prefix_hi = 0u, known_hi = 0u, prefix_lo = 0xFFC00000u, known_lo = 0xFFC00000u;
    int need = K;          // how many still to take among elements matching the prefix
    int candidates = 0;
    bool done = false;
    constexpr int UNR = 4;
    constexpr bool kRaw = SrcHasRaw<Src>::value;
    if constexpr (kRaw) {
        if (cached) {
            const uint32_t *g = src.raw_ptr(ctx, begin);
            const bool wide = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
            const int nwide = wide ? (len & ~3) : 0;
            for (int i = tid * 4; i < nwide; i += kSelThreads * 4) sel_cp_async(keys + i, g + i, true);
            for (int i = nwide + tid; i < len; i += kSelThreads) sel_cp_async(keys + i, g + i, false);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            // the barrier after the histogram reset below publishes the slice to the whole CTA
        }
    }
    // digit schedule: key bits 31..0 (4 digits), then index bits 21..16, 15..8, 7..0
    for (int pass = 0; pass < 7 && !done; pass++) {
        const bool on_key = pass < 4;
        const int shift = on_key ? 24 - 8 * pass : 16 - 8 * (pass - 4);
        const uint32_t wmask = pass == 4 ? 0x3Fu : 0xFFu;
        for (int i = tid; i < 256 * 32; i += kSelThreads) sh.hist[i] = 0;
        __syncthreads();
        // ---- scan the slice ------------------------------------------------------------------------------
        const uint32_t hist_sa = smem_addr(sh.hist) + 4u * lane;
        auto tally = [&](uint32_t key, int i) {       // element i of the slice is a candidate with this key
            if ((key & known_hi) != prefix_hi) return;
            if (on_key) {
                reds_add(hist_sa + 128u * ((key >> shift) & 0xFFu), 1u);
            } else {
                const uint32_t lo = ~src.index_of(ctx, begin + i);
                if ((lo & known_lo) == prefix_lo) reds_add(hist_sa + 128u * ((lo >> shift) & wmask), 1u);
            }
        };
        bool scanned = false;
        if (cached) {
            const int n4 = (len + 3) >> 2;
            const uint32_t keys_sa = smem_addr(keys), vbits_sa = smem_addr(vbits);
            if (pass > 0) {
                // cached keys, four per thread and load: after pass 0 almost nothing matches the prefix, so the scan is
                // one LDS.128 + four compares per thread and round (it was ~50 instructions per element, one at a time)
                for (int q0 = tid; q0 < n4; q0 += UNR * kSelThreads) {
                    uint4 kk[UNR];
                    uint32_t vb[UNR];
#pragma unroll
                    for (int j = 0; j < UNR; j++) {          // all loads first: the atomics below order against them
                        const int q = q0 + j * kSelThreads;
                        vb[j] = 0u;
                        if (q < n4) { kk[j] = lds128(keys_sa + 16u * q); vb[j] = (lds32(vbits_sa + 4u * (q >> 3)) >> ((q & 7) * 4)) & 0xFu; }
                    }
#pragma unroll
                    for (int j = 0; j < UNR; j++) {
                        const int q = q0 + j * kSelThreads;
                        if (vb[j] & 1u) tally(kk[j].x, 4 * q);
                        if (vb[j] & 2u) tally(kk[j].y, 4 * q + 1);
                        if (vb[j] & 4u) tally(kk[j].z, 4 * q + 2);
                        if (vb[j] & 8u) tally(kk[j].w, 4 * q + 3);
                    }
                }
                scanned = true;
            } else if constexpr (kRaw) {
                // pass 0 over the prefetched raw words: four keys per thread, written back in place
                for (int q0 = 0; q0 < n4; q0 += kSelThreads) {       // whole warps: the nibble exchange below shuffles
                    const int q = q0 + tid;
                    uint4 kk = make_uint4(0u, 0u, 0u, 0u);
                    uint32_t vb = 0u;
                    if (q < n4) {
                        // four independent evaluations (words past the end of the slice are evaluated and masked off:
                        // a branch per word would serialise the four dependency chains)
                        const uint4 rw = lds128(keys_sa + 16u * q);
                        const int i = 4 * q;
                        const bool o0 = src.from_raw(ctx, begin + i, rw.x, kk.x);
                        const bool o1 = src.from_raw(ctx, begin + i + 1, rw.y, kk.y);
                        const bool o2 = src.from_raw(ctx, begin + i + 2, rw.z, kk.z);
                        const bool o3 = src.from_raw(ctx, begin + i + 3, rw.w, kk.w);
                        vb = (o0 ? 1u : 0u) | (o1 ? 2u : 0u) | (o2 ? 4u : 0u) | (o3 ? 8u : 0u);
                        const int left = len - i;               // >= 1
                        if (left < 4) vb &= (1u << left) - 1u;
                        sts128(keys_sa + 16u * q, kk);
                    }
                    uint32_t word = vb << ((lane & 7) * 4);          // 8 lanes share one validity word
                    word |= __shfl_xor_sync(0xffffffffu, word, 1);
                    word |= __shfl_xor_sync(0xffffffffu, word, 2);
                    word |= __shfl_xor_sync(0xffffffffu, word, 4);
                    if ((lane & 7) == 0 && q < n4) vbits[q >> 3] = word;
                    if (vb & 1u) tally(kk.x, 4 * q);
                    if (vb & 2u) tally(kk.y, 4 * q + 1);
                    if (vb & 4u) tally(kk.z, 4 * q + 2);
                    if (vb & 8u) tally(kk.w, 4 * q + 3);
                }
                scanned = true;
            }
        }
        if (!scanned) {
            // generic path: Src::load per element, UNR independent loads in flight per thread
            for (int base = 0; base < len; base += UNR * kSelThreads) {
                bool ok[UNR];
                uint32_t key[UNR];
#pragma unroll
                for (int j = 0; j < UNR; j++) {
                    const int i = base + j * kSelThreads + tid;
                    ok[j] = false; key[j] = 0u;
                    if (i < len) ok[j] = src.load(ctx, begin + i, key[j]);
                }
#pragma unroll
                for (int j = 0; j < UNR; j++) {
                    const int i = base + j * kSelThreads + tid;
                    if (pass == 0 && cached) {
                        if (i < len) keys[i] = key[j];
                        const uint32_t bal = __ballot_sync(0xffffffffu, ok[j]);
                        if (lane == 0 && i < len) vbits[i >> 5] = bal;
                    }
                    if (ok[j]) tally(key[j], i);
                }
            }
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t s = 0;
#pragma unroll 8
            for (int r = 0; r < 32; r++) s += sh.hist[tid * 32 + ((r + tid) & 31)];
            sh.local[pass & 1][tid] = s;
        }
        if (solo) {
            if (tid < 256) sh.tot[tid] = sh.local[pass & 1][tid];
        } else {
            cluster.sync();
            if (tid < 256) {
                uint32_t s = 0;
                for (int r = 0; r < kClusterSize; r++) s += cluster.map_shared_rank(sh.local[pass & 1], r)[tid];
                sh.tot[tid] = s;
            }
        }
        __syncthreads();
        // ---- find the threshold digit: above = #matching elements in higher bins --------------
        if (tid < 32) {
            // lane handles bins 255-8*lane .. 255-8*lane-7 (descending)
            uint32_t c[8], s = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) { c[j] = sh.tot[255 - (lane * 8 + j)]; s += c[j]; }
            uint32_t incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (lane == 0) { sh.total = total; sh.found = 0; }
            __syncwarp();
            const uint32_t excl = incl - s;
            const uint32_t want_n = (uint32_t)min(need, (int)min(total, 0x7FFFFFFFu));
            if (want_n > 0 && excl < want_n && want_n <= incl) {   // exactly one lane
                uint32_t run = excl;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (run < want_n && want_n <= run + c[j]) {
                        sh.digit = 255 - (lane * 8 + j);
                        sh.above = run;
                        sh.eq = c[j];
                        sh.found = 1;
                    }
                    run += c[j];
                }
            }
        }
        __syncthreads();
        if (pass == 0) candidates = (int)sh.total;
        need = min(need, (int)sh.total);
        if (need == 0 || !sh.found) {
            done = true;
        } else {
            if (on_key) { prefix_hi |= sh.digit << shift; known_hi |= 0xFFu << shift; }
            else { prefix_lo |= sh.digit << shift; known_lo |= wmask << shift; }
            need -= (int)sh.above;
            if ((int)sh.eq == need) done = true;   // everything matching the prefix is selected
        }
        MD_STAMP();
        // no second cluster barrier: the next pass writes the OTHER sh.local buffer, and the barrier of that pass
        // orders this pass's remote reads before the buffer is written again two passes later
    }
    // selected set: every candidate with (comp & known) >= prefix
    const int selected = min(K, candidates);
    const bool lo_free = known_lo == 0xFFC00000u;    // no index digit fixed: the key alone decides

    // ---- collect this CTA's selected composites into its own shared memory ---------------------------
    unsigned long long *mine = sh.lists.mine, *all = sh.lists.all;
    if (selected > 0) {
        // decide(key, i): is candidate i of the slice selected.  Its composite carries the slice position for now; the
        // logical index is filled in below
        auto decide = [&](uint32_t key, int i) -> bool {
            const uint32_t kh = key & known_hi;
            if (kh < prefix_hi) return false;
            if (kh == prefix_hi && !lo_free) return ((~src.index_of(ctx, begin + i)) & known_lo) >= prefix_lo;
            return true;
        };
        const uint32_t mine_sa = smem_addr(mine);
        MD_STAMP();
        if (cached) {
            // two sweeps over the cached keys, four per thread and load: count, block-wide exclusive scan, write.  (A
            // ballot + shared atomic per round cost ~150 instructions per warp and round -- 12k cycles for a 25,200
            // element slice -- although one element in a hundred is taken.)
            const int n4 = (len + 3) >> 2;
            const uint32_t keys_sa = smem_addr(keys), vbits_sa = smem_addr(vbits);
            uint32_t mycount = 0;
            for (int q = tid; q < n4; q += kSelThreads) {
                const uint4 kk = lds128(keys_sa + 16u * q);
                const uint32_t vb = (lds32(vbits_sa + 4u * (q >> 3)) >> ((q & 7) * 4)) & 0xFu;
                mycount += ((vb & 1u) && decide(kk.x, 4 * q)) ? 1u : 0u;
                mycount += ((vb & 2u) && decide(kk.y, 4 * q + 1)) ? 1u : 0u;
                mycount += ((vb & 4u) && decide(kk.z, 4 * q + 2)) ? 1u : 0u;
                mycount += ((vb & 8u) && decide(kk.w, 4 * q + 3)) ? 1u : 0u;
            }
            MD_STAMP();
            uint32_t incl = mycount;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) sh.tot[tid >> 5] = incl;           // sh.tot is free after the passes
            __syncthreads();
            MD_STAMP();
            uint32_t pos = incl - mycount, all_warps = 0;
#pragma unroll
            for (int w = 0; w < kSelThreads / 32; w++) {
                const uint32_t t = sh.tot[w];
                if (w < (tid >> 5)) pos += t;
                all_warps += t;
            }
            if (tid == 0) sh.mine_count = all_warps;
            if (mycount) {
                for (int q = tid; q < n4; q += kSelThreads) {
                    const uint4 kk = lds128(keys_sa + 16u * q);
                    const uint32_t vb = (lds32(vbits_sa + 4u * (q >> 3)) >> ((q & 7) * 4)) & 0xFu;
                    if ((vb & 1u) && decide(kk.x, 4 * q)) sts64(mine_sa + 8u * pos++, ((unsigned long long)kk.x << 32) | (uint32_t)(4 * q));
                    if ((vb & 2u) && decide(kk.y, 4 * q + 1)) sts64(mine_sa + 8u * pos++, ((unsigned long long)kk.y << 32) | (uint32_t)(4 * q + 1));
                    if ((vb & 4u) && decide(kk.z, 4 * q + 2)) sts64(mine_sa + 8u * pos++, ((unsigned long long)kk.z << 32) | (uint32_t)(4 * q + 2));
                    if ((vb & 8u) && decide(kk.w, 4 * q + 3)) sts64(mine_sa + 8u * pos++, ((unsigned long long)kk.w << 32) | (uint32_t)(4 * q + 3));
                }
            }
        } else {
            const uint32_t count_sa = smem_addr(&sh.mine_count);
            for (int base = 0; base < len; base += UNR * kSelThreads) {
                bool take[UNR];
                unsigned long long comp[UNR];
                uint32_t m[UNR];
#pragma unroll
                for (int j = 0; j < UNR; j++) {
                    const int i = base + j * kSelThreads + tid;
                    uint32_t key = 0;
                    take[j] = i < len && src.load(ctx, begin + i, key) && decide(key, i);
                    comp[j] = ((unsigned long long)key << 32) | (uint32_t)i;
                }
                uint32_t total = 0;
#pragma unroll
                for (int j = 0; j < UNR; j++) { m[j] = __ballot_sync(0xffffffffu, take[j]); total += __popc(m[j]); }
                if (total) {
                    uint32_t pos = 0;
                    if (lane == 0) pos = atoms_add(count_sa, total);
                    pos = __shfl_sync(0xffffffffu, pos, 0);
#pragma unroll
                    for (int j = 0; j < UNR; j++) {
                        if (take[j]) sts64(mine_sa + 8u * (pos + __popc(m[j] & ((1u << lane) - 1u))), comp[j]);
                        pos += __popc(m[j]);
                    }
                }
            }
        }
    }
    MD_STAMP();
    __syncthreads();
    MD_STAMP();
    const int cnt = (int)sh.mine_count;
    // the logical index (an integer division or two per element in most Srcs) is computed here, on the dense list with
    // every lane busy, not inside the sparse, divergent scan above
    for (int p = tid; p < cnt; p += kSelThreads) {
        const unsigned long long c = mine[p];
        mine[p] = (c & 0xFFFFFFFF00000000ull) | (uint32_t)~src.index_of(ctx, begin + (int)(uint32_t)c);
    }
    MD_STAMP();
    // ---- every CTA sorts its own list (descending); composites are unique, real ones are > 0 ------------
    {
        int n2 = 32;
        while (n2 < cnt) n2 <<= 1;
        for (int i = cnt + tid; i < n2; i += kSelThreads) mine[i] = 0ull;
        __syncthreads();
        if (cnt > 1) bitonic_sort_desc(mine, n2);
    }
    MD_STAMP();
    if (solo) {
        // the sorted list is the answer
        __syncthreads();
        for (int p = tid; p < cnt; p += kSelThreads) sink.emit(seg, p, mine[p]);
        const int want = src.want(ctx);
        for (int i = selected + tid; i < want; i += kSelThreads) sink.pad(seg, i);
        if (tid == 0) sink.finish(seg, selected, candidates);
        return;
    }
    cluster.sync();      // every CTA's sorted list and count are visible
    MD_STAMP();
    // ---- pull the other CTAs' lists into local shared memory, list after list ---------------------------
    if (tid < kClusterSize) sh.cnt[tid] = *cluster.map_shared_rank(&sh.mine_count, tid);
    __syncthreads();
    int off[kClusterSize + 1];
    off[0] = 0;
#pragma unroll
    for (int r = 0; r < kClusterSize; r++) off[r + 1] = off[r] + (int)sh.cnt[r];
    const int gathered = min(off[kClusterSize], kSelMaxK);      // == selected
#pragma unroll 4
    for (int g = tid; g < gathered; g += kSelThreads) {
        int r = 0, o = 0;
#pragma unroll
        for (int q = 1; q < kClusterSize; q++)
            if (g >= off[q]) { r = q; o = off[q]; }
        if (r != rank) all[g] = cluster.map_shared_rank(mine, r)[g - o];
    }
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");    // done with the other CTAs' memory
    __syncthreads();
    MD_STAMP();
    // ---- rank of each of my composites = its position here + how many larger ones every other list holds; the
    //      owner emits it straight to its final slot (the sort, the ranking and Sink::emit are spread over the cluster)
    // branch-free lower bounds, all lists in lockstep (the seven searches are independent: one step is seven loads in
    // flight instead of seven dependent searches one after the other)
    int max_cnt = 0;
#pragma unroll
    for (int r = 0; r < kClusterSize; r++) max_cnt = max(max_cnt, off[r + 1] - off[r]);
    int top_step = 1;
    while (top_step * 2 <= max_cnt) top_step <<= 1;
    const uint32_t all_sa = smem_addr(all);
    auto rank_of = [&](unsigned long long e, int p) {
        int pos[kClusterSize];
#pragma unroll
        for (int r = 0; r < kClusterSize; r++) pos[r] = 0;
        for (int st = top_step; st > 0; st >>= 1) {
#pragma unroll
            for (int r = 0; r < kClusterSize; r++) {
                const int probe = pos[r] + st;                       // lists hold their elements in descending order
                if (probe <= off[r + 1] - off[r] && lds64(all_sa + 8u * (off[r] + probe - 1)) > e) pos[r] = probe;
            }
        }
        int rk = p;
#pragma unroll
        for (int r = 0; r < kClusterSize; r++) rk += (r == rank) ? 0 : pos[r];
        return rk;
    };
    for (int p = tid; p < cnt; p += kSelThreads) {
        const unsigned long long e = mine[p];
        if constexpr (SinkHasPre<Sink>::value) {
            const auto pre = sink.prefetch(seg, e);
            sink.emit_pre(seg, rank_of(e, p), e, pre);
        } else {
            sink.emit(seg, rank_of(e, p), e);
        }
    }
    if (rank == 0) {
        const int want = src.want(ctx);
        for (int i = selected + tid; i < want; i += kSelThreads) sink.pad(seg, i);
        if (tid == 0) sink.finish(seg, selected, candidates);
    }
    MD_STAMP();
#ifdef MD_SEL_TIMING
    if (tid == 0 && blockIdx.x == 0) {       // phase durations of the first cluster's leader (one-off instrumented builds)
        printf("select N=%d K=%d sel=%d mine=%d:", N, K, selected, cnt);
        for (int i = 1; i < n_ph; i++) printf(" %lld", t_ph[i] - t_ph[i - 1]);
        printf(" cycles\n");
    }
#endif
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");      // nobody still reads my list
}

template <class Src, class Sink>
__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kSelThreads)
select_sorted_kernel(const Src src, const Sink sink, const int cache_elems)
{
    pdl_entry();
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    SelShared &sh = *reinterpret_cast<SelShared *>(dyn_smem);
    uint32_t *keys = reinterpret_cast<uint32_t *>(dyn_smem + sizeof(SelShared));
    uint32_t *vbits = keys + cache_elems;   // [cache_elems/32] validity bits
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / kClusterSize;
    for (int it = 0;; it++) {
        const int seg = src.segment_of(cid, it);       // uniform over the cluster
        if (seg < 0) break;
        select_segment(src, sink, cache_elems, seg, sh, keys, vbits, cluster, rank);
        // Shared memory is reused by the next segment.  Inside the CTA this barrier is enough; across the cluster the
        // closing barrier.cluster.wait of a cluster-mode segment has already seen every CTA finish its remote reads,
        // and solo segments touch no remote memory (CTAs that sat one out simply wait at the next cluster barrier).
        __syncthreads();
    }
}

// max_slice: the largest per-CTA slice any segment of this launch will have (host-known); decides
// how much dynamic shared memory is spent on the key cache.
template <class Src, class Sink>
cudaError_t launch_select_sorted(const Src &src, const Sink &sink, int nclusters, int max_segment_len,
                                 cudaStream_t stream)
{
    const int nseg = nclusters;          // one cluster per segment unless Src::segment_of hands a cluster several
    if (nseg <= 0) return cudaSuccess;
    auto kern = select_sorted_kernel<Src, Sink>;
    int per = (max_segment_len + kClusterSize - 1) / kClusterSize;
    if (per < kSelSoloMax) per = max_segment_len < kSelSoloMax ? max_segment_len : kSelSoloMax;   // solo slices are whole segments
    per = (per + 31) & ~31;
    int cache = per <= kSelMaxCacheElems ? per : 0;   // 0 -> recompute keys every pass
    cache = (cache + 31) & ~31;
    const size_t dyn = sizeof(SelShared) + (size_t)(cache + cache / 32 + 8) * sizeof(uint32_t);
    {
        // The attribute is per DEVICE, so a process-wide "already configured" flag is wrong on a second GPU and racy between
        // host threads.  Always allow the largest slice this kernel can ask for: a constant, identical from every caller.
        constexpr size_t kMaxDyn = sizeof(SelShared) + (size_t)(kSelMaxCacheElems + 32 + kSelMaxCacheElems / 32 + 8) * sizeof(uint32_t);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(dyn > kMaxDyn ? dyn : kMaxDyn));
        if (e != cudaSuccess) return e;
        if (kClusterSize > 8) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e != cudaSuccess) return e;
        }
    }
    return launch_pdl(kern, dim3(nseg * kClusterSize), dim3(kSelThreads), dyn, stream, src, sink, cache);
}

}  // namespace md
