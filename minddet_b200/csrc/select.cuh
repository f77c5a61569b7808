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
// Selected composites are appended to the leader CTA's shared memory through DSMEM atomics, the
// leader bitonic-sorts them (<= 2048) and calls Sink::emit(seg, rank, comp) in sorted order.
//
// Semantics == oracle o_topk / o_sample (oracle/CONVENTIONS.md #4, #13).  Reference idiom being
// replaced: ops.TopK(sorted=True) at centerpoint/det3d_ms/models/bbox_heads/center_head.py:435 and
// pointpillars/src/pointpillars.py:764.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace md {
namespace cg = cooperative_groups;

constexpr int kClusterSize = 8;
constexpr int kSelThreads = 512;
constexpr int kSelMaxK = 2048;            // sorted output limit per segment
constexpr int kSelDirectMax = 1024;       // segments up to this length skip the radix passes: the leader sorts them all
                                          // (a single-CTA bitonic sort of 4096 costs ~50 us: measured, so keep this small)
constexpr int kSelMaxIndexBits = 22;      // segment length < 4 Mi elements
constexpr int kSelMaxCacheElems = 44 * 1024; // key cache per CTA (dynamic smem)

struct SelShared {
    union {
        uint32_t hist[256 * 32];            // [bin][lane]  (select passes)
        unsigned long long cand[kSelMaxK];  // leader only: selected composites (after the passes) / the whole short segment
    };
    uint32_t local[2][256];    // this CTA's per-bin totals (read remotely through DSMEM); double-buffered per pass
    uint32_t tot[256];         // cluster totals
    uint32_t cand_count;       // leader only
    uint32_t digit, above, eq, total, found;
};

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
            unsigned long long r[R];
#pragma unroll
            for (int m = 0; m < R; m++) r[m] = (tid + m * kSelThreads < n2) ? v[tid + m * kSelThreads] : 0ull;
            for (int kk = (k == 32 ? 2 : k); kk <= k; kk <<= 1) {
                for (int j = min(kk >> 1, 16); j > 0; j >>= 1) {
#pragma unroll
                    for (int m = 0; m < R; m++) {
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
//   __device__ int segment_of(int launch_index) const;              launch order -> segment id (put the longest first)
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
__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kSelThreads)
select_sorted_kernel(const Src src, const Sink sink, const int cache_elems)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    SelShared &sh = *reinterpret_cast<SelShared *>(dyn_smem);
    uint32_t *keys = reinterpret_cast<uint32_t *>(dyn_smem + sizeof(SelShared));
    uint32_t *vbits = keys + cache_elems;   // [cache_elems/32] validity bits
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int seg = src.segment_of(blockIdx.x / kClusterSize);   // launch order -> segment (largest segments first)
    const int tid = threadIdx.x, lane = tid & 31;
    const typename Src::Ctx ctx = src.prepare(seg);
    if (!src.active(ctx)) return;            // uniform over the cluster: nobody reaches a cluster barrier

    const int N = src.length(ctx);
    const int K = min(src.want(ctx), kSelMaxK);
    // slice (multiple of 32 so validity words never straddle CTAs)
    int per = (N + kClusterSize - 1) / kClusterSize;
    per = (per + 31) & ~31;
    const int begin = min(N, rank * per), end = min(N, begin + per);
    const int len = end - begin;
    const bool cached = len <= cache_elems;

    if (tid == 0) sh.cand_count = 0;

    if (N <= kSelDirectMax) {
        // ---- short segment: no radix passes.  The leader gathers every candidate, sorts them all and emits the K best.
        if (rank != 0) return;                   // uniform over the cluster; nobody touches distributed shared memory
        __syncthreads();
        for (int base = 0; base < N; base += kSelThreads) {
            const int i = base + tid;
            uint32_t key = 0u;
            const bool ok = i < N && src.load(ctx, i, key);
            const uint32_t m = __ballot_sync(0xffffffffu, ok);
            if (m) {
                uint32_t pos = 0;
                const int leader = __ffs(m) - 1;
                if (lane == leader) pos = atomicAdd(&sh.cand_count, (uint32_t)__popc(m));
                pos = __shfl_sync(0xffffffffu, pos, leader);
                if (ok) sh.cand[pos + __popc(m & ((1u << lane) - 1u))] = ((unsigned long long)key << 32) | (uint32_t)~src.index_of(ctx, i);
            }
        }
        __syncthreads();
        const int ncand = (int)sh.cand_count;
        int n2 = 32;
        while (n2 < ncand) n2 <<= 1;
        for (int i = ncand + tid; i < n2; i += kSelThreads) sh.cand[i] = 0ull;
        __syncthreads();
        bitonic_sort_desc(sh.cand, n2);
        const int sel = min(K, ncand), want = src.want(ctx);
        for (int i = tid; i < want; i += kSelThreads) {
            if (i < sel) sink.emit(seg, i, sh.cand[i]);
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
    // digit schedule: key bits 31..0 (4 digits), then index bits 21..16, 15..8, 7..0
    for (int pass = 0; pass < 7 && !done; pass++) {
        const bool on_key = pass < 4;
        const int shift = on_key ? 24 - 8 * pass : 16 - 8 * (pass - 4);
        const uint32_t wmask = pass == 4 ? 0x3Fu : 0xFFu;
        for (int i = tid; i < 256 * 32; i += kSelThreads) sh.hist[i] = 0;
        __syncthreads();
        // ---- scan the slice (UNR independent loads in flight per thread) -------------------------------
        for (int base = 0; base < len; base += UNR * kSelThreads) {
            bool ok[UNR];
            uint32_t key[UNR];
#pragma unroll
            for (int j = 0; j < UNR; j++) {
                const int i = base + j * kSelThreads + tid;
                ok[j] = false; key[j] = 0u;
                if (i < len) {
                    if (pass == 0 || !cached) ok[j] = src.load(ctx, begin + i, key[j]);
                    else { ok[j] = (vbits[i >> 5] >> (i & 31)) & 1u; key[j] = keys[i]; }
                }
            }
#pragma unroll
            for (int j = 0; j < UNR; j++) {
                const int i = base + j * kSelThreads + tid;
                if (pass == 0 && cached) {
                    if (i < len) keys[i] = key[j];
                    const uint32_t bal = __ballot_sync(0xffffffffu, ok[j]);
                    if (lane == 0 && i < len) vbits[i >> 5] = bal;
                }
                if (ok[j] && (key[j] & known_hi) == prefix_hi) {
                    if (on_key) {
                        atomicAdd(&sh.hist[((key[j] >> shift) & 0xFFu) * 32 + lane], 1u);
                    } else {
                        const uint32_t lo = ~src.index_of(ctx, begin + i);
                        if ((lo & known_lo) == prefix_lo) atomicAdd(&sh.hist[((lo >> shift) & wmask) * 32 + lane], 1u);
                    }
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
        cluster.sync();
        if (tid < 256) {
            uint32_t s = 0;
            for (int r = 0; r < kClusterSize; r++) s += cluster.map_shared_rank(sh.local[pass & 1], r)[tid];
            sh.tot[tid] = s;
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
        // no second cluster barrier: the next pass writes the OTHER sh.local buffer, and the barrier of that pass
        // orders this pass's remote reads before the buffer is written again two passes later
    }
    // selected set: every candidate with (comp & known) >= prefix
    const int selected = min(K, candidates);
    const bool lo_free = known_lo == 0xFFC00000u;    // no index digit fixed: the key alone decides

    // ---- collect into the leader's shared memory (DSMEM) -------------------------------------------
    if (selected > 0) {
        unsigned long long *lead_cand = cluster.map_shared_rank(sh.cand, 0);
        uint32_t *lead_count = cluster.map_shared_rank(&sh.cand_count, 0);
        for (int base = 0; base < len; base += kSelThreads) {
            const int i = base + tid;
            bool ok = false;
            uint32_t key = 0;
            if (i < len) {
                if (cached) { ok = (vbits[i >> 5] >> (i & 31)) & 1u; key = keys[i]; }
                else ok = src.load(ctx, begin + i, key);
            }
            bool take = false;
            uint32_t lo = 0u;
            if (ok) {
                const uint32_t kh = key & known_hi;
                if (kh > prefix_hi) take = true;
                else if (kh == prefix_hi) {
                    if (lo_free) take = true;
                    else { lo = ~src.index_of(ctx, begin + i); take = (lo & known_lo) >= prefix_lo; }
                }
                if (take && lo == 0u) lo = ~src.index_of(ctx, begin + i);
            }
            const unsigned long long comp = ((unsigned long long)key << 32) | lo;
            const uint32_t m = __ballot_sync(0xffffffffu, take);
            if (m) {
                uint32_t pos = 0;
                const int leader = __ffs(m) - 1;
                if (lane == leader) pos = atomicAdd(lead_count, (uint32_t)__popc(m));
                pos = __shfl_sync(0xffffffffu, pos, leader);
                if (take) lead_cand[pos + __popc(m & ((1u << lane) - 1u))] = comp;
            }
        }
    }
    cluster.sync();
    if (rank != 0) return;

    // ---- leader: bitonic sort (descending) and emit -------------------------------------------------
    int n2 = 32;
    while (n2 < selected) n2 <<= 1;
    for (int i = selected + tid; i < n2; i += kSelThreads) sh.cand[i] = 0ull;
    __syncthreads();
    bitonic_sort_desc(sh.cand, n2);
    const int want = src.want(ctx);
    for (int i = tid; i < want; i += kSelThreads) {
        if (i < selected) sink.emit(seg, i, sh.cand[i]);
        else sink.pad(seg, i);
    }
    if (tid == 0) sink.finish(seg, selected, candidates);
}

// max_slice: the largest per-CTA slice any segment of this launch will have (host-known); decides
// how much dynamic shared memory is spent on the key cache.
template <class Src, class Sink>
cudaError_t launch_select_sorted(const Src &src, const Sink &sink, int nseg, int max_segment_len,
                                 cudaStream_t stream)
{
    if (nseg <= 0) return cudaSuccess;
    auto kern = select_sorted_kernel<Src, Sink>;
    int per = (max_segment_len + kClusterSize - 1) / kClusterSize;
    per = (per + 31) & ~31;
    int cache = per <= kSelMaxCacheElems ? per : 0;   // 0 -> recompute keys every pass
    cache = (cache + 31) & ~31;
    const size_t dyn = sizeof(SelShared) + (size_t)(cache + cache / 32 + 8) * sizeof(uint32_t);
    static size_t configured = 0;   // per instantiation
    if (dyn > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return e;
        configured = dyn;
    }
    kern<<<dim3(nseg * kClusterSize), dim3(kSelThreads), dyn, stream>>>(src, sink, cache);
    return cudaGetLastError();
}

}  // namespace md
