// The fused transmission pass (native-RNG mode): reference sim.py:622-649 + utils.py:93-128.
//
// The reference makes n_variants x n_layers x 2 passes over the edge lists per day (compute_trans_sus
// + compute_infections per variant and layer, both directions) and re-gathers per-agent values with
// fancy indexing each time.  Here one pass evaluates both directions and all variants.  Three forms,
// with the same probability chain, the same Philox key (layer, edge) and the same winner key, so they
// give identical results:
//
//   * adjacency form (edge_pass_sparse_kernel; static layers): one warp per TRANSMITTER walks its row of
//     the bidirectional adjacency -- only the edges of the few percent of agents who can transmit today;
//   * dense streaming form (edge_pass_kernel; dynamic layers or use_adjacency=False): every edge
//     (p1:int32, p2:int32, beta:f32 = 12 bytes) is streamed from HBM exactly once per day with 128-bit
//     loads.  Algorithmic bytes per day: 12*E_total + 8*N.  One persistent CTA of 512 threads per SM:
//       1. the day's "can transmit" bitmap (1 bit per agent) is staged in SHARED MEMORY (125 KB at 1M agents;
//          a global / L1 path is used when it does not fit);
//       2. each thread loads two quads of four edges (six 128-bit loads, the next tile's already in flight,
//          the one after prefetched into L2), requests all sixteen bitmap words, then tests the endpoint
//          bits -- ~80 % of the edges have no transmitting endpoint and are dropped here;
//       3. survivors are appended to a per-warp shared-memory queue (ballot + popc compaction) and popped
//          32 at a time; the two record gathers of a batch are issued when it is popped and consumed when the
//          NEXT batch is popped, so the gathers, the float32 chain and the Philox draw run on full warps and
//          the gather latency overlaps the filtering of the following quads;
//   * agent-partitioned form (edge_pass_partition_kernel): a rank walks, for every GLOBAL transmitter, the
//     edges that end in one of its own agents (covasim_b200/partition.py).
//
// Per-agent inputs are the 16-byte agent records written by prepare_transmission (cvb_device.cuh:AgentRecord;
// the per-layer factors are applied here), or, for the single-variant dense form, the reference's per-layer
// {rel_trans, rel_sus} pairs.
//
// Randomness: one Philox4x32-10 call per LIVE edge (non-zero probability in either direction), keyed
// (seed, P_EDGE, layer, day, edge index): words 0-1 give the p1->p2 uniform, words 2-3 the p2->p1 one.
//
// Output: for each target hit at least once, infect_key[target] = min over successful transmissions
// of (variant, layer, direction, edge) -- exactly the reference's winner: variants and layers are
// processed in order, infect() drops already-infected targets, and np.unique(return_index=True)
// keeps the first occurrence in [direction 0 edges..., direction 1 edges...] (people.py:465-473).
// The first hit on a target also appends it to the day's candidate list (warp-aggregated atomics).
#include <stdlib.h>
#include "cvb_internal.cuh"

namespace cvb {

struct EdgeParams {
    float beta[CVB_MAX_VARIANTS];
    float beta_layer[CVB_MAX_LAYERS], iso_factor[CVB_MAX_LAYERS], quar_factor[CVB_MAX_LAYERS];   // sim.py:640-642
    float asymp_factor, vl_early, vl_late, pad_;
    uint64_t seed;
    int64_t n;
    int64_t n_words;        // words of the transmit bitmap
    int32_t t, nv;
};

constexpr int kQueueCap = 160;            // < 32 left over + up to 128 appended per iteration

__device__ __forceinline__ void record_hit(unsigned long long* __restrict__ infect_key, int32_t* __restrict__ cand,
                                           unsigned int* __restrict__ n_cand, int target, unsigned long long key) {
    unsigned long long old = atomicMin(infect_key + target, key);
    if (old == kEmptyKey) {
        unsigned int pos = warp_append32(n_cand);
        cand[pos] = target;
    }
}

// Evaluate one candidate edge in both directions (reference utils.py:113-123) and record transmissions.  Split in two so
// that the dense pass can issue the two record gathers of a batch of candidates and consume them one batch later.
__device__ __forceinline__ void gather_records(const TransRecords& rec, int a, int b, float4& ra, float4& rb) {
    ra = __ldg(rec.rec + a);
    rb = __ldg(rec.rec + b);
}

// probability that `src` (record rs_) infects `tgt` (record rt_) over an edge of weight w on layer l; 0 if src cannot transmit
template <bool MULTI>
__device__ __forceinline__ float direction_prob(const TransRecords& rec, const EdgeParams& ep, const float4 rs_, const float4 rt_, int tgt,
                                                float w, int l, int& variant) {
    const uint32_t cs = __float_as_uint(rs_.w);
    variant = 0;
    if (!(cs & 7u) || rt_.y == 0.0f) return 0.0f;                  // source cannot transmit / target not susceptible
    variant = (int)(cs & 7u) - 1;
    const float t = record_trans(rs_.x, cs, ep.asymp_factor, ep.iso_factor[l], ep.quar_factor[l], ep.beta_layer[l], ep.vl_early, ep.vl_late);
    float imm = rt_.z;
    if (MULTI && variant > 0) imm = __ldg(rec.sus_imm + (int64_t)variant * ep.n + tgt);
    const float sus = record_sus(rt_.y, __float_as_uint(rt_.w), ep.quar_factor[l], imm);
    return edge_prob(ep.beta[variant], w, t, sus);
}

template <bool MULTI>
__device__ __forceinline__ void finish_edge(const TransRecords& rec, const EdgeParams& ep, int a, int b, float w, int l, int64_t e,
        const float4 ra, const float4 rb, unsigned long long* __restrict__ infect_key, int32_t* __restrict__ cand, unsigned int* __restrict__ n_cand) {
    int va, vb;
    const float p01 = direction_prob<MULTI>(rec, ep, ra, rb, b, w, l, va);
    const float p10 = direction_prob<MULTI>(rec, ep, rb, ra, a, w, l, vb);
    if (p01 != 0.0f || p10 != 0.0f) {
        const u32x4 r = keyed_words(ep.seed, P_EDGE, (uint32_t)l, ep.t, e, 0);
        const unsigned long long base = ((unsigned long long)l << 48) | (unsigned long long)e;
        if (p01 != 0.0f && u53(r.x, r.y) < (double)p01)
            record_hit(infect_key, cand, n_cand, b, ((unsigned long long)va << 56) | base);
        if (p10 != 0.0f && u53(r.z, r.w) < (double)p10)
            record_hit(infect_key, cand, n_cand, a, ((unsigned long long)vb << 56) | (1ull << 40) | base);
    }
}

// the dense pass gathers 8-byte per-layer pairs with one variant (MULTI = false) and 16-byte agent records otherwise
template <bool MULTI> struct RecType { typedef float4 type; };
template <> struct RecType<false> { typedef float2 type; };

// ---- the same for the per-layer {rel_trans, rel_sus} pairs (one variant; what the dense streaming pass gathers) ----
__device__ __forceinline__ void gather_records(const TransRecords& rec, int64_t n, int l, int a, int b, float2& ra, float2& rb) {
    const float2* __restrict__ ts = rec.ts8 + (int64_t)l * n;
    ra = __ldg(ts + a);
    rb = __ldg(ts + b);
}
__device__ __forceinline__ void gather_records(const TransRecords& rec, int64_t, int, int a, int b, float4& ra, float4& rb) {
    gather_records(rec, a, b, ra, rb);
}

template <bool MULTI>
__device__ __forceinline__ void finish_edge(const TransRecords& rec, const EdgeParams& ep, int a, int b, float w, int l, int64_t e,
        const float2 ra, const float2 rb, unsigned long long* __restrict__ infect_key, int32_t* __restrict__ cand, unsigned int* __restrict__ n_cand) {
    const float p01 = ra.x != 0.0f ? edge_prob(ep.beta[0], w, ra.x, rb.y) : 0.0f;
    const float p10 = rb.x != 0.0f ? edge_prob(ep.beta[0], w, rb.x, ra.y) : 0.0f;
    if (p01 != 0.0f || p10 != 0.0f) {
        const u32x4 r = keyed_words(ep.seed, P_EDGE, (uint32_t)l, ep.t, e, 0);
        const unsigned long long base = ((unsigned long long)l << 48) | (unsigned long long)e;
        if (p01 != 0.0f && u53(r.x, r.y) < (double)p01) record_hit(infect_key, cand, n_cand, b, base);
        if (p10 != 0.0f && u53(r.z, r.w) < (double)p10) record_hit(infect_key, cand, n_cand, a, (1ull << 40) | base);
    }
}

template <bool MULTI>
__device__ __forceinline__ void process_edge(const TransRecords& rec, const EdgeParams& ep, int a, int b, float w, int l, int64_t e,
        unsigned long long* __restrict__ infect_key, int32_t* __restrict__ cand, unsigned int* __restrict__ n_cand) {
    typename RecType<MULTI>::type ra, rb;
    gather_records(rec, ep.n, l, a, b, ra, rb);
    finish_edge<MULTI>(rec, ep, a, b, w, l, e, ra, rb, infect_key, cand, n_cand);
}

// One bit test per endpoint: word >> (index mod 32) with the hardware's wrap-around funnel shift (no mask instruction)
__device__ __forceinline__ unsigned endpoint_bit(unsigned word, int index) { return __funnelshift_r(word, 0u, (unsigned)index); }

struct EdgeQuad {                       // four consecutive edges of one layer, as loaded by one thread
    int4 a, b;
    float4 w;
    int cnt;                            // how many of the four exist (4 except at the ragged end of a layer)
};

__device__ __forceinline__ void load_quad(const int32_t* __restrict__ p1, const int32_t* __restrict__ p2, const float* __restrict__ beta,
                                          int64_t n_edges, unsigned q, EdgeQuad& Q) {
    const int64_t e0 = (int64_t)q * 4;
    if (e0 + 4 <= n_edges) {
        Q.a = ld_stream(reinterpret_cast<const int4*>(p1) + q);
        Q.b = ld_stream(reinterpret_cast<const int4*>(p2) + q);
        Q.w = ld_stream(reinterpret_cast<const float4*>(beta) + q);
        Q.cnt = 4;
    } else {                                                        // ragged end (or a lane past the end): scalar loads
        int* pa = reinterpret_cast<int*>(&Q.a); int* pb = reinterpret_cast<int*>(&Q.b); float* pw = reinterpret_cast<float*>(&Q.w);
        Q.cnt = e0 < n_edges ? (int)(n_edges - e0) : 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool ok = k < Q.cnt;
            pa[k] = ok ? p1[e0 + k] : 0; pb[k] = ok ? p2[e0 + k] : 0; pw[k] = ok ? beta[e0 + k] : 0.0f;
        }
    }
}

// THREADS per CTA (one persistent CTA per SM when the bitmap is staged in shared memory) and QPT quads (of four edges)
// per thread per iteration.  More quads per thread = more independent loads in flight per warp (3*QPT 128-bit streaming
// loads + 8*QPT shared-memory bit lookups issued back to back) at the price of registers, i.e. of warps per SM.
// Ask L2 for the cache line holding `p` (fire and forget): the streaming loads two tiles later then find their data in L2
// (~0.4 us) instead of HBM (~2 us under load), so the bytes a warp must keep in flight in REGISTERS shrink accordingly
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

template <bool MULTI, bool SMEM_BITS, int THREADS, int QPT, int PF>
__global__ void __launch_bounds__(THREADS) edge_pass_kernel(const __grid_constant__ LayerTable L, TransRecords rec,
        const __grid_constant__ EdgeParams ep, const unsigned int* __restrict__ inf_bits, unsigned long long* __restrict__ infect_key,
        int32_t* __restrict__ cand, unsigned int* __restrict__ n_cand) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kWarps = THREADS / 32;
    // layout: [queue entries uint4 x kWarps*kQueueCap][bitmap words].  A queue holds candidate edges of ONE layer (it is
    // flushed when the CTA finishes a layer), so an entry is just {p1, p2, beta, edge index}.
    uint4* q_edge = reinterpret_cast<uint4*>(smem_raw) + warp_id() * kQueueCap;
    const unsigned int* bits = inf_bits;
    if (SMEM_BITS) {
        unsigned int* s_bits = reinterpret_cast<unsigned int*>(smem_raw + (size_t)kWarps * kQueueCap * sizeof(uint4));
        const int64_t n4 = ep.n_words >> 2;                            // 128-bit copies, then the tail
        for (int64_t wd = threadIdx.x; wd < n4; wd += THREADS) reinterpret_cast<uint4*>(s_bits)[wd] = __ldg(reinterpret_cast<const uint4*>(inf_bits) + wd);
        for (int64_t wd = (n4 << 2) + threadIdx.x; wd < ep.n_words; wd += THREADS) s_bits[wd] = inf_bits[wd];
        bits = s_bits;
        __syncthreads();
    }
    const int lane = lane_id();
    const unsigned lt_mask = (1u << lane) - 1u;
    constexpr unsigned kTileQuads = THREADS * QPT;                  // quads per CTA per iteration (build_layer_table's tile = 4x this)
    const unsigned stride = gridDim.x * kTileQuads;

    for (int entry = 0; entry < L.n_layers; ++entry) {
        const int32_t* __restrict__ p1 = L.l[entry].p1;
        const int32_t* __restrict__ p2 = L.l[entry].p2;
        const float* __restrict__ beta = L.l[entry].beta;
        const int64_t n_edges = L.l[entry].n_edges;
        const int l = L.layer_id[entry];
        const unsigned nq = (unsigned)((n_edges + 3) >> 2);           // quads of this layer (n_edges < 2^32)
        // the CTA's first tile in this layer continues the round-robin over ALL layers' tiles, so the load stays balanced
        const unsigned rot = (unsigned)(L.tile_start[entry] % gridDim.x);
        const unsigned first = (blockIdx.x + gridDim.x - rot) % gridDim.x;
        // quad j of this thread in a tile is first_quad + j * THREADS: every 128-bit load of a warp is one contiguous 512 B
        unsigned q = first * kTileQuads + threadIdx.x;                // q - lane is warp-uniform
        int qn = 0;                                                   // warp-uniform queue length
        // deferred drain: the record gathers of a batch of 32 candidates are issued when the batch is popped and consumed
        // when the NEXT batch is popped, so their L2 / HBM latency overlaps the filtering of the following quads
        bool pend = false;                                            // warp-uniform
        uint4 pc = make_uint4(0u, 0u, 0u, 0u);
        typename RecType<MULTI>::type pra = {}, prb = {};

        auto load_tile = [&](unsigned qq, EdgeQuad (&T)[QPT]) {
#pragma unroll
            for (int j = 0; j < QPT; ++j) load_quad(p1, p2, beta, n_edges, qq + (unsigned)j * THREADS, T[j]);
            if (PF > 0) {                                             // L2 prefetch of the tile PF iterations ahead of this load
#pragma unroll
                for (int j = 0; j < QPT; ++j) {
                    const unsigned qp = qq + (unsigned)PF * stride + (unsigned)j * THREADS;
                    if (qp < nq && (lane & 7) == 0) {                 // one request per 128-byte line
                        prefetch_l2(reinterpret_cast<const int4*>(p1) + qp);
                        prefetch_l2(reinterpret_cast<const int4*>(p2) + qp);
                        prefetch_l2(reinterpret_cast<const float4*>(beta) + qp);
                    }
                }
            }
        };
        // filter QPT quads per lane into the warp's queue and drain full warps of candidates
        auto filter_tile = [&](const EdgeQuad (&T)[QPT], unsigned qq) {
            // keep an edge only if one endpoint can transmit.  All 8*QPT bitmap words are requested before any is used, so
            // the shared-memory latency is paid once per tile, not once per edge.
            unsigned wa[QPT][4], wb[QPT][4];
#pragma unroll
            for (int j = 0; j < QPT; ++j) {
                const int a[4] = {T[j].a.x, T[j].a.y, T[j].a.z, T[j].a.w}, b[4] = {T[j].b.x, T[j].b.y, T[j].b.z, T[j].b.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) { wa[j][k] = bits[a[k] >> 5]; wb[j][k] = bits[b[k] >> 5]; }
            }
#pragma unroll
            for (int j = 0; j < QPT; ++j) {
                const int a[4] = {T[j].a.x, T[j].a.y, T[j].a.z, T[j].a.w}, b[4] = {T[j].b.x, T[j].b.y, T[j].b.z, T[j].b.w};
                const float w[4] = {T[j].w.x, T[j].w.y, T[j].w.z, T[j].w.w};
                const unsigned e_base = (qq + (unsigned)j * THREADS) * 4u;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool keep = ((endpoint_bit(wa[j][k], a[k]) | endpoint_bit(wb[j][k], b[k])) & 1u) != 0 && k < T[j].cnt;
                    const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
                    if (keep) q_edge[qn + __popc(m & lt_mask)] = make_uint4((unsigned)a[k], (unsigned)b[k], __float_as_uint(w[k]), e_base + (unsigned)k);
                    qn += __popc(m);
                }
                __syncwarp();
                while (qn >= 32) {
                    qn -= 32;
                    const uint4 c = q_edge[qn + lane];
                    __syncwarp();
                    typename RecType<MULTI>::type ra, rb;
                    gather_records(rec, ep.n, l, (int)c.x, (int)c.y, ra, rb);
                    if (pend) finish_edge<MULTI>(rec, ep, (int)pc.x, (int)pc.y, __uint_as_float(pc.z), l, (int64_t)pc.w, pra, prb, infect_key, cand, n_cand);
                    pc = c; pra = ra; prb = rb; pend = true;
                }
                __syncwarp();
            }
        };

        // software pipeline, two register sets: the 3*QPT 128-bit loads of the NEXT tile are in flight while this one is
        // filtered and drained
        EdgeQuad A[QPT], B[QPT];
        unsigned qbase = q - lane;
        if (qbase < nq) load_tile(q, A);
        while (qbase < nq) {
            const unsigned q2 = q + stride, qbase2 = qbase + stride;
            if (qbase2 < nq) load_tile(q2, B);
            filter_tile(A, q);
            if (qbase2 >= nq) break;
            q = q2 + stride; qbase = qbase2 + stride;
            if (qbase < nq) load_tile(q, A);
            filter_tile(B, q2);
        }
        if (pend) finish_edge<MULTI>(rec, ep, (int)pc.x, (int)pc.y, __uint_as_float(pc.z), l, (int64_t)pc.w, pra, prb, infect_key, cand, n_cand);
        if (lane < qn) {                                              // leftovers of this layer
            const uint4 c = q_edge[lane];
            process_edge<MULTI>(rec, ep, (int)c.x, (int)c.y, __uint_as_float(c.z), l, (int64_t)c.w, infect_key, cand, n_cand);
        }
        __syncwarp();
    }
}

// ---- dense streaming form, tiles staged by the TMA engine ------------------------------------------------------------------
// Same pass as edge_pass_kernel, with the edge tiles moved HBM -> shared memory by bulk asynchronous copies (cp.async.bulk,
// completion on an mbarrier) instead of register-held 128-bit loads.  Every warp owns a private ring of NS stages of
// {p1, p2, beta} x 32 quads (1.5 KB): its lane 0 issues the three copies of a stage, all lanes wait on the stage's mbarrier, pull
// their quad into registers, and lane 0 immediately refills the stage with the warp's tile NS iterations ahead -- so NS - 1 stages
// per warp are in flight for the whole kernel, they cost no registers and no issue slots of the filtering code, and no warp ever
// waits for another (a first version with CTA-wide stages and a producer warp was gated by its slowest warp: 72 us against 56).
// MEASURED (profiles/r2/README.md, "dense pass with bulk-copy staging"): at C2 this form takes 74 us (20 warps x 2 stages; 89 us with
// 16 x 3, 101 us with 12 x 4) against 56 us for edge_pass_kernel<512 threads x 2 quads>.  The time follows the number of filtering
// warps, not the stages in flight: next to the 125 KB transmit bitmap shared memory holds at most ~20 warps' queues and stages, each
// warp filters one quad per wait instead of two, and the copies are issued by a lane of the filtering warps.  The pass is bound by
// the issue / latency chain of the filter and the drain (bit tests, ballots, Philox for the ~20 % live edges), not by bytes in
// flight, so the register-staged kernel stays the default; this one is selectable with CVB_DENSE_VARIANT=3.
// Layout of dynamic shared memory:
//   [mbarriers: CW x NS] [per-warp queues: CW x kQueueCapT uint4] [stages: CW x NS x 3 x 32 x 16 B] [transmit bitmap]
constexpr int kQueueCapT = 96;            // < 32 left over + 64 appended between two drains (a drain after every two edges of a quad)
constexpr int kMaxStages = 4;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const uint32_t a = smem_addr(bar);
    const long long t0 = clock64();
    for (unsigned spins = 0;; ++spins) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
        if ((spins & 1023u) == 1023u && clock64() - t0 > 4000000000ll) __trap();      // ~2 s: a lost completion must not hang the GPU
    }
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

template <bool MULTI, int CW>
__global__ void __launch_bounds__(CW * 32) edge_pass_tma_kernel(const __grid_constant__ LayerTable L, TransRecords rec,
        const __grid_constant__ EdgeParams ep, const unsigned int* __restrict__ inf_bits, unsigned long long* __restrict__ infect_key,
        int32_t* __restrict__ cand, unsigned int* __restrict__ n_cand, int n_stages) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int TQ = CW * 32;                                       // quads per CTA tile: one per lane
    constexpr unsigned kStageBytes = 3u * 32u * 16u;                  // one warp's share of a tile
    constexpr size_t kBarBytes = ((size_t)CW * kMaxStages * 8 + 127) / 128 * 128;
    const int warp = warp_id(), lane = lane_id();
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw) + warp * kMaxStages;
    uint4* q_edge = reinterpret_cast<uint4*>(smem_raw + kBarBytes) + warp * kQueueCapT;
    unsigned char* stage_base = smem_raw + kBarBytes + (size_t)CW * kQueueCapT * sizeof(uint4);
    unsigned char* stages = stage_base + (size_t)warp * n_stages * kStageBytes;
    unsigned int* s_bits = reinterpret_cast<unsigned int*>(stage_base + (size_t)CW * n_stages * kStageBytes);
    if (lane == 0) for (int k = 0; k < n_stages; ++k) mbar_init(full + k, 1u);
    if (threadIdx.x == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    {
        const int64_t n4 = ep.n_words >> 2;
        for (int64_t wd = threadIdx.x; wd < n4; wd += blockDim.x) reinterpret_cast<uint4*>(s_bits)[wd] = __ldg(reinterpret_cast<const uint4*>(inf_bits) + wd);
        for (int64_t wd = (n4 << 2) + threadIdx.x; wd < ep.n_words; wd += blockDim.x) s_bits[wd] = inf_bits[wd];
    }
    __syncthreads();
    const unsigned int* bits = s_bits;
    const unsigned grid = gridDim.x;
    const unsigned lt_mask = (1u << lane) - 1u;

    // the warp's tile sequence over all layers, walked twice: by the issuing side (NS ahead) and by the consuming side
    struct Cursor { int entry; unsigned tile, n_full; };
    auto open_layer = [&](Cursor& c) {                                // first full tile of layer c.entry for this CTA, or move on
        for (; c.entry < L.n_layers; ++c.entry) {
            c.n_full = (unsigned)(L.l[c.entry].n_edges / (TQ * 4));
            const unsigned rot = (unsigned)(L.tile_start[c.entry] % grid);
            c.tile = (blockIdx.x + grid - rot) % grid;
            if (c.tile < c.n_full) return;
        }
    };
    auto issue = [&](Cursor& c, unsigned it) {                        // lane 0: the copies of the tile under the cursor, then advance it
        if (c.entry >= L.n_layers) return;
        const unsigned sidx = it % (unsigned)n_stages;
        unsigned char* dst = stages + (size_t)sidx * kStageBytes;
        const size_t off = ((size_t)c.tile * TQ + (size_t)warp * 32) * 16;
        mbar_arrive_expect_tx(full + sidx, kStageBytes);
        bulk_copy_g2s(dst, reinterpret_cast<const unsigned char*>(L.l[c.entry].p1) + off, 512u, full + sidx);
        bulk_copy_g2s(dst + 512, reinterpret_cast<const unsigned char*>(L.l[c.entry].p2) + off, 512u, full + sidx);
        bulk_copy_g2s(dst + 1024, reinterpret_cast<const unsigned char*>(L.l[c.entry].beta) + off, 512u, full + sidx);
        c.tile += grid;
        if (c.tile >= c.n_full) { ++c.entry; open_layer(c); }
    };
    Cursor ahead{0, 0u, 0u};
    unsigned it_issue = 0;
    if (lane == 0) {
        open_layer(ahead);
        for (; it_issue < (unsigned)n_stages; ++it_issue) issue(ahead, it_issue);
    }

    unsigned it = 0;
    for (int entry = 0; entry < L.n_layers; ++entry) {
        const int32_t* __restrict__ p1 = L.l[entry].p1;
        const int32_t* __restrict__ p2 = L.l[entry].p2;
        const float* __restrict__ beta = L.l[entry].beta;
        const int64_t n_edges = L.l[entry].n_edges;
        const int l = L.layer_id[entry];
        const unsigned n_full = (unsigned)(n_edges / (TQ * 4));
        const unsigned rot = (unsigned)(L.tile_start[entry] % grid);
        const unsigned first = (blockIdx.x + grid - rot) % grid;
        int qn = 0;                                                   // warp-uniform queue length
        bool pend = false;                                            // deferred drain, as in edge_pass_kernel
        uint4 pc = make_uint4(0u, 0u, 0u, 0u);
        typename RecType<MULTI>::type pra = {}, prb = {};

        auto drain = [&]() {
            while (qn >= 32) {
                qn -= 32;
                const uint4 c = q_edge[qn + lane];
                __syncwarp();
                typename RecType<MULTI>::type ra, rb;
                gather_records(rec, ep.n, l, (int)c.x, (int)c.y, ra, rb);
                if (pend) finish_edge<MULTI>(rec, ep, (int)pc.x, (int)pc.y, __uint_as_float(pc.z), l, (int64_t)pc.w, pra, prb, infect_key, cand, n_cand);
                pc = c; pra = ra; prb = rb; pend = true;
            }
        };
        auto filter_quad = [&](const EdgeQuad& Q, unsigned qq) {
            const int a[4] = {Q.a.x, Q.a.y, Q.a.z, Q.a.w}, b[4] = {Q.b.x, Q.b.y, Q.b.z, Q.b.w};
            const float w[4] = {Q.w.x, Q.w.y, Q.w.z, Q.w.w};
            unsigned wa[4], wb[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { wa[k] = bits[a[k] >> 5]; wb[k] = bits[b[k] >> 5]; }
            const unsigned e_base = qq * 4u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool keep = ((endpoint_bit(wa[k], a[k]) | endpoint_bit(wb[k], b[k])) & 1u) != 0 && k < Q.cnt;
                const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
                if (keep) q_edge[qn + __popc(m & lt_mask)] = make_uint4((unsigned)a[k], (unsigned)b[k], __float_as_uint(w[k]), e_base + (unsigned)k);
                qn += __popc(m);
                if (k & 1) { __syncwarp(); drain(); __syncwarp(); }
            }
        };

        for (unsigned tile = first; tile < n_full; tile += grid, ++it) {
            const unsigned sidx = it % (unsigned)n_stages, use = it / (unsigned)n_stages;
            mbar_wait(full + sidx, use & 1u);
            const uint4* st4 = reinterpret_cast<const uint4*>(stages + (size_t)sidx * kStageBytes);
            EdgeQuad Q;
            const uint4 va = st4[lane], vb = st4[32 + lane], vw = st4[64 + lane];
            Q.a = make_int4((int)va.x, (int)va.y, (int)va.z, (int)va.w);
            Q.b = make_int4((int)vb.x, (int)vb.y, (int)vb.z, (int)vb.w);
            Q.w = make_float4(__uint_as_float(vw.x), __uint_as_float(vw.y), __uint_as_float(vw.z), __uint_as_float(vw.w));
            Q.cnt = 4;
            __syncwarp();                                             // every lane has its quad: the stage can be refilled
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(ahead, it_issue);
                ++it_issue;
            }
            filter_quad(Q, tile * (unsigned)TQ + (unsigned)(warp * 32 + lane));
        }
        // the layer's ragged last tile (fewer than TQ quads): straight from global memory, by the CTA whose turn it is
        if ((int64_t)n_full * (TQ * 4) < n_edges && n_full % grid == first) {
            const unsigned qq = n_full * (unsigned)TQ + (unsigned)(warp * 32 + lane);
            EdgeQuad Q;
            load_quad(p1, p2, beta, n_edges, qq, Q);
            filter_quad(Q, qq);
        }
        if (pend) finish_edge<MULTI>(rec, ep, (int)pc.x, (int)pc.y, __uint_as_float(pc.z), l, (int64_t)pc.w, pra, prb, infect_key, cand, n_cand);
        if (lane < qn) {                                              // leftovers of this layer
            const uint4 c = q_edge[lane];
            process_edge<MULTI>(rec, ep, (int)c.x, (int)c.y, __uint_as_float(c.z), l, (int64_t)c.w, infect_key, cand, n_cand);
        }
        __syncwarp();
    }
}

// ---- sparse form: only the edges of today's transmitters, through the bidirectional adjacency ----------
// One warp per transmitter; lanes stride over its adjacency range (all static layers, both directions:
// ~36 entries of 16 bytes per agent in a hybrid population, contiguous).  Same probability chain, same
// Philox key (layer, edge) and same winner key as the dense pass, so results are identical.
template <bool MULTI>
__global__ void __launch_bounds__(kThreads) edge_pass_sparse_kernel(TransRecords rec, const __grid_constant__ EdgeParams ep,
        const long long* __restrict__ adj_ptr, const uint4* __restrict__ adj, const int32_t* __restrict__ trans_list,
        const unsigned int* __restrict__ n_trans_ptr, unsigned long long* __restrict__ infect_key, int32_t* __restrict__ cand,
        unsigned int* __restrict__ n_cand, unsigned long long* __restrict__ work_row) {
    const unsigned int n_trans = *n_trans_ptr;
    unsigned long long visited = 0;
    const int64_t n = ep.n;
    const int lane = lane_id();
    const unsigned int warps_total = (gridDim.x * blockDim.x) >> 5;
    for (unsigned int ti = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ti < n_trans; ti += warps_total) {
        const int i = trans_list[ti];
        const long long beg = adj_ptr[i], end = adj_ptr[i + 1];
        visited += (unsigned long long)(end - beg);
        const float4 ri = __ldg(rec.rec + i);                           // the transmitter's own record
        const uint32_t ci = __float_as_uint(ri.w);
        const int vi = (int)(ci & 7u) - 1;
        const float beta_v = ep.beta[vi < 0 ? 0 : vi];
        for (long long off = beg + lane; off < end; off += 32) {
            const uint4 en = __ldg(adj + off);
            const int j = (int)en.x;
            const int l = (int)(en.z >> 1);
            const int dir = (int)(en.z & 1u);
            const float t_i = record_trans(ri.x, ci, ep.asymp_factor, ep.iso_factor[l], ep.quar_factor[l], ep.beta_layer[l], ep.vl_early, ep.vl_late);
            if (t_i == 0.0f) continue;                        // cannot transmit on this layer
            const float4 rj = __ldg(rec.rec + j);
            if (rj.y == 0.0f) continue;                       // target not susceptible
            float imm = rj.z;
            if (MULTI && vi > 0) imm = __ldg(rec.sus_imm + (int64_t)vi * n + j);
            const float s_j = record_sus(rj.y, __float_as_uint(rj.w), ep.quar_factor[l], imm);
            const float p = edge_prob(beta_v, __uint_as_float(en.w), t_i, s_j);
            if (p != 0.0f) {
                const int64_t e = (int64_t)en.y;
                const u32x4 r = keyed_words(ep.seed, P_EDGE, (uint32_t)l, ep.t, e, 0);
                const double u = dir == 0 ? u53(r.x, r.y) : u53(r.z, r.w);
                if (u < (double)p)
                    record_hit(infect_key, cand, n_cand, j, ((unsigned long long)vi << 56) | ((unsigned long long)l << 48) |
                                                            ((unsigned long long)dir << 40) | (unsigned long long)e);
            }
        }
    }
    // work accounting for the roofline: adjacency entries visited and transmitters today
    if (lane == 0 && visited) atomicAdd(work_row, visited);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(work_row + 1, (unsigned long long)n_trans);
}

// ---- sparse form v2 (fused day pipeline): transmitter ENTRIES instead of a list of agents ----------------------------------
// day_mid_kernel writes one 32-byte entry per transmitter: {agent, row length, row begin (64 bit)} {rel_trans, transmit code}, so
// a group's dependent chain is entry -> adjacency entries -> neighbour records (three loads instead of the five of the list form:
// list -> row pointers -> own record -> entries -> records).  G lanes share a transmitter and each lane keeps U adjacency entries
// in flight: all U entry loads are issued, then all U record gathers, then the arithmetic -- a row of 32 entries costs two
// memory round trips per group.  Same probability chain, Philox key and winner key as every other form.
// The kernel is launched with programmatic dependent launch, so it may be resident while day_mid_kernel is still writing the
// entries and records: they are read with L2 loads (ld.global.cg), never through the non-coherent path.
constexpr unsigned int kHitBuf = 512;
template <bool MULTI, int G, int U>
__global__ void __launch_bounds__(kThreads) edge_pass_sparse2_kernel(TransRecords rec, const __grid_constant__ EdgeParams ep,
        const uint4* __restrict__ adj, const uint4* __restrict__ ents, const unsigned int* __restrict__ n_trans_ptr,
        unsigned long long* __restrict__ infect_key, int32_t* __restrict__ cand, unsigned int* __restrict__ n_cand,
        unsigned long long* __restrict__ work_row) {
    __shared__ unsigned long long s_visited;
    __shared__ int32_t s_hit[kHitBuf];                          // this CTA's newly hit targets: ONE global atomic reserves their list slots
    __shared__ unsigned int s_n_hit, s_hit_base;
    pdl_trigger();
    if (threadIdx.x == 0) { s_visited = 0; s_n_hit = 0; }
    __syncthreads();
    pdl_wait();                                                 // the entries and records are written by day_mid_kernel
    const unsigned int n_trans = __ldcg(n_trans_ptr);           // (L2 loads: the kernel may have been resident while its inputs were written)
    unsigned long long visited = 0;
    const int64_t n = ep.n;
    const int gl = threadIdx.x & (G - 1);
    const unsigned int groups_total = (gridDim.x * blockDim.x) / G;
    for (unsigned int ti = (blockIdx.x * blockDim.x + threadIdx.x) / G; ti < n_trans; ti += groups_total) {
        const uint4 e0 = __ldcg(ents + 2 * (int64_t)ti), e1 = __ldcg(ents + 2 * (int64_t)ti + 1);
        const int len = (int)e0.y;
        const long long beg = (long long)(((unsigned long long)e0.w << 32) | (unsigned long long)e0.z);
        const float rt = __uint_as_float(e1.x);
        const uint32_t ci = e1.y;
        const int vi = (int)(ci & 7u) - 1;
        const float beta_v = ep.beta[vi < 0 ? 0 : vi];
        if (gl == 0) visited += (unsigned long long)len;
        for (int base = gl; base < len; base += G * U) {
            uint4 en[U];
            float4 rj[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int off = base + u * G;
                en[u] = off < len ? __ldg(adj + beg + off) : make_uint4(0u, 0u, 0xFFFFFFFFu, 0u);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) rj[u] = en[u].z != 0xFFFFFFFFu ? __ldcg(rec.rec + en[u].x) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (en[u].z == 0xFFFFFFFFu || rj[u].y == 0.0f) continue;      // past the row's end / target not susceptible
                const int j = (int)en[u].x;
                const int l = (int)(en[u].z >> 1);
                const int dir = (int)(en[u].z & 1u);
                const float t_i = record_trans(rt, ci, ep.asymp_factor, ep.iso_factor[l], ep.quar_factor[l], ep.beta_layer[l], ep.vl_early, ep.vl_late);
                if (t_i == 0.0f) continue;                                    // cannot transmit on this layer
                float imm = rj[u].z;
                if (MULTI && vi > 0) imm = __ldcg(rec.sus_imm + (int64_t)vi * n + j);
                const float s_j = record_sus(rj[u].y, __float_as_uint(rj[u].w), ep.quar_factor[l], imm);
                const float p = edge_prob(beta_v, __uint_as_float(en[u].w), t_i, s_j);
                if (p != 0.0f) {
                    const int64_t e = (int64_t)en[u].y;
                    const u32x4 r = keyed_words(ep.seed, P_EDGE, (uint32_t)l, ep.t, e, 0);
                    const double uu = dir == 0 ? u53(r.x, r.y) : u53(r.z, r.w);
                    if (uu < (double)p) {
                        const unsigned long long key = ((unsigned long long)vi << 56) | ((unsigned long long)l << 48) |
                                                       ((unsigned long long)dir << 40) | (unsigned long long)e;
                        if (atomicMin(infect_key + j, key) == kEmptyKey) {              // first hit on this target today
                            const unsigned int k = atomicAdd(&s_n_hit, 1u);
                            if (k < kHitBuf) s_hit[k] = j;
                            else cand[atomicAdd(n_cand, 1u)] = j;                       // (buffer full: straight to the global list)
                        }
                    }
                }
            }
        }
    }
    // work accounting for the roofline (one global atomic per CTA: ~10^4 same-address atomics would serialise in L2)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) visited += __shfl_down_sync(0xFFFFFFFFu, visited, d);
    if (lane_id() == 0 && visited) atomicAdd(&s_visited, visited);
    __syncthreads();
    const unsigned int n_hit = s_n_hit < kHitBuf ? s_n_hit : kHitBuf;
    if (threadIdx.x == 0 && n_hit) s_hit_base = atomicAdd(n_cand, n_hit);
    __syncthreads();
    for (unsigned int q = threadIdx.x; q < n_hit; q += blockDim.x) cand[s_hit_base + q] = s_hit[q];
    if (threadIdx.x == 0 && s_visited) atomicAdd(work_row, s_visited);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(work_row + 1, (unsigned long long)n_trans);
}

// ---- agent-partitioned form (one large simulation over several GPUs) ------------------------------------
// Every GPU owns a contiguous range of agents and evaluates the transmissions whose TARGET it owns.  Its
// adjacency has one row per GLOBAL agent (the possible sources) holding the edges that end in a LOCAL
// target.  What it needs to know about a remote source is one byte (transmit_code) all-gathered by the host
// once per day plus the replicated initial rel_trans; per-layer transmissibility is rebuilt from those with
// the same float32 chain as prepare_transmission, so probabilities, Philox keys (layer, edge) and winner keys
// are those of the single-GPU run, bit for bit.

// Warp-level reservation: every lane asks for `c` slots of a global list; one atomic per warp
__device__ __forceinline__ unsigned int warp_reserve(unsigned int* counter, int c) {
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane_id() >= d) incl += o;
    }
    const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    unsigned int base = 0;
    if (lane_id() == 31 && total) base = atomicAdd(counter, (unsigned int)total);
    base = __shfl_sync(0xFFFFFFFFu, base, 31);
    return base + (unsigned int)(incl - c);
}

// non-zero bytes of the gathered code array -> compact (unordered) list of global transmitter ids
__global__ void __launch_bounds__(kThreads) codes_to_list_kernel(const uint8_t* __restrict__ codes, int64_t n_slots,
        int32_t* __restrict__ list, unsigned int* __restrict__ n_list) {
    const int64_t n16 = n_slots / 16;                                 // n_slots is a multiple of 32
    const int64_t n16_pad = (n16 + 31) / 32 * 32;                     // whole warps stay in the loop (shuffles)
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n16_pad; g += (int64_t)gridDim.x * blockDim.x) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (g < n16) v = __ldg(reinterpret_cast<const uint4*>(codes) + g);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const int c = count4(v.x) + count4(v.y) + count4(v.z) + count4(v.w);
        if (!__any_sync(0xFFFFFFFFu, c != 0)) continue;
        unsigned int pos = warp_reserve(n_list, c);
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if ((w[q] >> (8 * k)) & 0xFFu) list[pos++] = (int32_t)(g * 16 + q * 4 + k);
    }
}

// set bits of a gathered bitmap -> compact (unordered) list of global ids
__global__ void __launch_bounds__(kThreads) bits_to_list_kernel(const unsigned int* __restrict__ bits, int64_t n_words,
        int32_t* __restrict__ list, unsigned int* __restrict__ n_list) {
    const int64_t n_pad = (n_words + 31) / 32 * 32;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_pad; g += (int64_t)gridDim.x * blockDim.x) {
        unsigned w = g < n_words ? __ldg(bits + g) : 0u;
        if (!__any_sync(0xFFFFFFFFu, w != 0u)) continue;
        unsigned int pos = warp_reserve(n_list, __popc(w));
        while (w) {
            const int b = __ffs(w) - 1;
            w &= w - 1;
            list[pos++] = (int32_t)(g * 32 + b);
        }
    }
}

struct PartHits {                         // one entry per successful transmission (parallel arrays, shared counter)
    int32_t* tgt; int32_t* src; unsigned long long* key; unsigned int* n; unsigned int* dropped; int64_t cap;
};

// G lanes per transmitter, U adjacency entries in flight per lane (with W ranks a transmitter's row holds ~36/W local entries, so a
// full warp per row would idle most of its lanes).  A group's dependent chain per transmitter is list index -> row header
// {row pointers, transmit code, base rel_trans} -> adjacency entries -> target records; the first two levels are software-
// pipelined across the group's transmitters (the index two ahead and the header one ahead are in flight while a row is
// processed), the last two issue all U loads before consuming any.
struct PartHeader { long long beg, end; unsigned code; float rt; };

template <bool MULTI, int G, int U>
__global__ void __launch_bounds__(kThreads) edge_pass_partition_kernel(TransRecords rec, const __grid_constant__ EdgeParams ep,
        float trans_redux, const long long* __restrict__ adj_ptr, const uint4* __restrict__ adj,
        const int32_t* __restrict__ glist, const unsigned int* __restrict__ n_glist, const uint8_t* __restrict__ codes,
        const float* __restrict__ base_trans, unsigned long long* __restrict__ infect_key, PartHits hits,
        unsigned long long* __restrict__ work_row) {
    const unsigned int n_trans = *n_glist;
    unsigned long long visited = 0;
    const int64_t n = ep.n;                                              // local agents
    const int gl = threadIdx.x & (G - 1);
    const unsigned int stride = (gridDim.x * blockDim.x) / G;
    auto header = [&](int i) {
        PartHeader h{0, 0, 0u, 0.0f};
        if (i >= 0) { h.beg = __ldg(adj_ptr + i); h.end = __ldg(adj_ptr + i + 1); h.code = __ldg(codes + i); h.rt = __ldg(base_trans + i); }
        return h;
    };
    unsigned int ti = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    int i0 = ti < n_trans ? __ldg(glist + ti) : -1;                      // global id of the source
    int i1 = ti + stride < n_trans ? __ldg(glist + ti + stride) : -1;
    PartHeader h0 = header(i0);
    for (; ti < n_trans; ti += stride) {
        const PartHeader h1 = header(i1);
        const int i2 = (unsigned long long)ti + 2ull * stride < n_trans ? __ldg(glist + ti + 2 * stride) : -1;
        const int len = (int)(h0.end - h0.beg);                          // 0: no contact on this GPU
        if (len > 0) {
            if (gl == 0) visited += (unsigned long long)len;
            const uint32_t ci = h0.code;
            const int vi = (int)(ci & 7u) - 1;
            const float rt = (ci & 128u) ? fmul(h0.rt, trans_redux) : h0.rt;
            const float beta_v = ep.beta[vi < 0 ? 0 : vi];
            for (int base = gl; base - gl < len; base += G * U) {
                uint4 en[U];
                float4 rj[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int off = base + u * G;
                    en[u] = off < len ? __ldg(adj + h0.beg + off) : make_uint4(0u, 0u, 0xFFFFFFFFu, 0u);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) rj[u] = en[u].z != 0xFFFFFFFFu ? __ldg(rec.rec + en[u].x) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (en[u].z == 0xFFFFFFFFu || rj[u].y == 0.0f) continue;      // past the row's end / target not susceptible
                    const int j = (int)en[u].x;                                   // LOCAL target
                    const int l = (int)(en[u].z >> 1);
                    const int dir = (int)(en[u].z & 1u);
                    const float t_i = record_trans(rt, ci, ep.asymp_factor, ep.iso_factor[l], ep.quar_factor[l], ep.beta_layer[l], ep.vl_early, ep.vl_late);
                    if (t_i == 0.0f) continue;
                    float imm = rj[u].z;
                    if (MULTI && vi > 0) imm = __ldg(rec.sus_imm + (int64_t)vi * n + j);
                    const float s_j = record_sus(rj[u].y, __float_as_uint(rj[u].w), ep.quar_factor[l], imm);
                    const float p = edge_prob(beta_v, __uint_as_float(en[u].w), t_i, s_j);
                    if (p != 0.0f) {
                        const int64_t e = (int64_t)en[u].y;
                        const u32x4 r = keyed_words(ep.seed, P_EDGE, (uint32_t)l, ep.t, e, 0);
                        const double uu = dir == 0 ? u53(r.x, r.y) : u53(r.z, r.w);
                        if (uu < (double)p) {
                            const unsigned long long key = ((unsigned long long)vi << 56) | ((unsigned long long)l << 48) |
                                                           ((unsigned long long)dir << 40) | (unsigned long long)e;
                            atomicMin(infect_key + j, key);
                            const unsigned int pos = warp_append32(hits.n);      // aggregated over the lanes active here (any group)
                            if ((int64_t)pos < hits.cap) { hits.tgt[pos] = j; hits.src[pos] = i0; hits.key[pos] = key; }
                            else atomicAdd(hits.dropped, 1u);
                        }
                    }
                }
            }
        }
        i0 = i1; h0 = h1; i1 = i2;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) visited += __shfl_down_sync(0xFFFFFFFFu, visited, d);
    if (lane_id() == 0 && visited) atomicAdd(work_row, visited);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(work_row + 1, (unsigned long long)n_trans);
}

int build_layer_table(cvb_sim* s, LayerTable& L, int tile_edges, uint32_t skip_mask);

// the TMA-staged form; returns -1 if it does not apply (bitmap + two stages per warp do not fit shared memory, unaligned arrays)
template <bool MULTI, int CW>
static int launch_edge_pass_tma(cvb_sim* s, uint32_t skip_mask, const EdgeParams& ep, size_t bitmap_bytes, cudaStream_t st, int max_stages) {
    constexpr size_t stage = (size_t)CW * 3u * 32u * 16u;             // one stage of every warp
    const size_t fixed = ((size_t)CW * kMaxStages * 8 + 127) / 128 * 128 + (size_t)CW * kQueueCapT * sizeof(uint4) + ((bitmap_bytes + 15) / 16) * 16;
    const size_t limit = 227 * 1024;
    if (fixed + 2 * stage > limit) return -1;
    int n_stages = (int)((limit - fixed) / stage);
    if (n_stages > max_stages) n_stages = max_stages;
    LayerTable L;
    if (build_layer_table(s, L, CW * 32 * kEdgesPerThread, skip_mask)) return 1;
    const int64_t tiles = L.tile_start[L.n_layers];
    if (tiles == 0) return 0;
    for (int q = 0; q < L.n_layers; ++q)
        if ((((uintptr_t)L.l[q].p1 | (uintptr_t)L.l[q].p2 | (uintptr_t)L.l[q].beta) & 15) != 0) return -1;
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, s->device);
    const int grid = (int)(tiles < n_sm ? tiles : n_sm);
    auto kern = edge_pass_tma_kernel<MULTI, CW>;
    static bool configured[64] = {false};
    const int dev = s->device & 63;
    if (!configured[dev]) {
        CVB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured[dev] = true;
    }
    const size_t smem = fixed + (size_t)n_stages * stage;
    kern<<<grid, CW * 32, smem, st>>>(L, s->rec, ep, s->inf_bits, s->infect_key, s->cand, s->n_cand, n_stages);
    CVB_LAUNCH_CHECK();
    return 0;
}

template <bool MULTI, bool SMEM_BITS, int THREADS, int QPT, int PF>
static int launch_edge_pass(cvb_sim* s, uint32_t skip_mask, const EdgeParams& ep, size_t bitmap_bytes, int ctas_per_sm, cudaStream_t st) {
    auto kern = edge_pass_kernel<MULTI, SMEM_BITS, THREADS, QPT, PF>;
    LayerTable L;
    if (build_layer_table(s, L, THREADS * QPT * kEdgesPerThread, skip_mask)) return 1;
    const int64_t tiles = L.tile_start[L.n_layers];
    if (tiles == 0) return 0;
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, s->device);
    const int64_t max_grid = (int64_t)n_sm * ctas_per_sm;
    const int grid = (int)(tiles < max_grid ? tiles : max_grid);
    const size_t smem = (size_t)(THREADS / 32) * kQueueCap * sizeof(uint4) + (SMEM_BITS ? bitmap_bytes : 0);
    static bool configured[64] = {false};                            // per instantiation and per device
    const int dev = s->device & 63;
    if (!configured[dev]) {
        CVB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured[dev] = true;
    }
    kern<<<grid, THREADS, smem, st>>>(L, s->rec, ep, s->inf_bits, s->infect_key, s->cand, s->n_cand);
    CVB_LAUNCH_CHECK();
    return 0;
}

}  // namespace cvb

using namespace cvb;

int cvb::list_from_bits(cvb_sim* s, const unsigned int* bits, int64_t n_words, cudaStream_t st) {
    CVB_CHECK(cudaMemsetAsync(s->n_glist, 0, sizeof(unsigned int), st));
    bits_to_list_kernel<<<grid_for(n_words, kThreads, 148 * 8), kThreads, 0, st>>>(bits, n_words, s->glist, s->n_glist);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb::build_layer_table(cvb_sim* s, LayerTable& L, int tile_edges, uint32_t skip_mask) {
    int64_t acc = 0;
    int q = 0;
    for (int l = 0; l < s->pars.n_layers; ++l) {
        if (skip_mask & (1u << l)) continue;
        L.l[q] = s->layers[l];
        L.layer_id[q] = l;
        CVB_REQUIRE(L.l[q].n_edges == 0 || L.l[q].p1, "layer %d is not bound (cvb_bind_layer)", l);
        CVB_REQUIRE(L.l[q].n_edges < (1ll << 32), "layer %d has too many edges for the 32-bit edge index of the streaming pass", l);
        L.tile_start[q] = acc;
        acc += (L.l[q].n_edges + tile_edges - 1) / tile_edges;
        ++q;
    }
    L.n_layers = q;
    for (int j = q; j < CVB_MAX_LAYERS; ++j) { L.l[j] = LayerPtrs{nullptr, nullptr, nullptr, 0}; L.layer_id[j] = 0; }
    for (int j = q; j <= CVB_MAX_LAYERS; ++j) L.tile_start[j] = acc;
    return 0;
}

extern "C" int cvb_edge_pass(cvb_sim* s, int32_t t, cvb_stream st) {
    return cvb::edge_pass_impl(s, t, (cudaStream_t)st, false);
}

// from_entries: the adjacency part reads the transmitter entries written by day_mid_kernel (fused day pipeline) instead of the
// plain transmitter list written by post_prepare_kernel
int cvb::edge_pass_impl(cvb_sim* s, int32_t t, cudaStream_t st, bool from_entries) {
    CVB_REQUIRE(s && s->pars_set, "cvb_edge_pass: handle not ready");
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_edge_pass: day %d outside [0,%d)", t, s->npts);
    CVB_REQUIRE((s->rec.rec || s->rec.ts8) && s->rec_layers >= s->pars.n_layers, "cvb_edge_pass: call cvb_prepare_transmission first");
    EdgeParams ep;
    for (int v = 0; v < CVB_MAX_VARIANTS; ++v) ep.beta[v] = s->pars.beta[v];
    ep.seed = s->seed; ep.n = s->n; ep.t = t; ep.nv = s->nv;
    for (int l = 0; l < CVB_MAX_LAYERS; ++l) { ep.beta_layer[l] = s->pars.beta_layer[l]; ep.iso_factor[l] = s->pars.iso_factor[l]; ep.quar_factor[l] = s->pars.quar_factor[l]; }
    ep.asymp_factor = s->pars.asymp_factor; ep.pad_ = 0.0f;
    ep.vl_early = viral_load_value(true, s->pars.frac_time, s->pars.load_ratio);
    ep.vl_late = viral_load_value(false, s->pars.frac_time, s->pars.load_ratio);
    ep.n_words = (s->n + 31) / 32;
    const bool multi = s->nv > 1;
    uint32_t skip_mask = 0;
    CVB_REQUIRE(s->rec.rec || !(s->partitioned || (s->adj && s->adj_layer_mask) || multi), "cvb_edge_pass: agent records missing (layers changed since cvb_prepare_transmission)");
    if (s->partitioned) {
        // agent-partitioned form: global transmitter list from the all-gathered codes, then the local adjacency rows
        CVB_REQUIRE(s->padj_ptr && s->padj && s->codes_global, "cvb_edge_pass: partitioned handle without adjacency / codes (cvb_bind_partition_adjacency, cvb_set_partition)");
        for (int l = 0; l < s->pars.n_layers; ++l)
            CVB_REQUIRE(((s->padj_layer_mask >> l) & 1u) || s->layers[l].n_edges == 0, "cvb_edge_pass: layer %d is not covered by the partitioned adjacency (dynamic layers cannot be partitioned)", l);
        CVB_CHECK(cudaMemsetAsync(s->n_glist, 0, sizeof(unsigned int), st));
        codes_to_list_kernel<<<grid_for(s->n_slots / 16, kThreads, 148 * 8), kThreads, 0, st>>>(s->codes_global, s->n_slots, s->glist, s->n_glist);
        CVB_LAUNCH_CHECK();
        PartHits hits{s->cand, s->hit_src, s->hit_key, s->n_cand, s->part_flags, s->hit_cap};
        unsigned long long* work_row = s->edge_work + (int64_t)t * 2;
        const int grid = 148 * 8;
        // lanes per transmitter from the mean local row length (entries per agent slot, both directions)
        const double mean_row = (double)s->padj_entries / (double)(s->n_global > 0 ? s->n_global : 1);
#define CVB_PART(M, G, U) edge_pass_partition_kernel<M, G, U><<<grid, kThreads, 0, st>>>(s->rec, ep, s->pars.trans_redux, s->padj_ptr, s->padj, s->glist, s->n_glist, \
                                                                                      s->codes_global, s->rel_trans_global, s->infect_key, hits, work_row)
        const int shape = s->tune[4];
        if (shape == 1 || (!shape && mean_row >= 24.0)) { if (multi) CVB_PART(true, 16, 4); else CVB_PART(false, 16, 4); }
        else if (shape == 2 || (!shape && mean_row >= 12.0)) { if (multi) CVB_PART(true, 16, 2); else CVB_PART(false, 16, 2); }
        else if (shape == 3 || (!shape && mean_row >= 6.0)) { if (multi) CVB_PART(true, 8, 2); else CVB_PART(false, 8, 2); }
        else { if (multi) CVB_PART(true, 4, 2); else CVB_PART(false, 4, 2); }
#undef CVB_PART
        CVB_LAUNCH_CHECK();
        return 0;
    }
    if (s->adj && s->adj_layer_mask) {
        // static layers: visit only the adjacency ranges of today's transmitters (a device-side count: the grid is
        // sized for a large outbreak and surplus warps exit after one load)
        skip_mask = s->adj_layer_mask;
        const int grid = 148 * 8;
        unsigned long long* work_row = s->edge_work + (int64_t)t * 2;
        if (from_entries) {
            CVB_REQUIRE(s->trans_ent, "cvb_edge_pass: transmitter entries missing");
            const uint4* adj = s->adj; const uint4* ents = s->trans_ent; const unsigned int* nt = s->n_trans;
            const int shape = s->tune[4];                               // lanes per transmitter x entries in flight per lane (cvb_tune)
            const int g2 = s->tune[5] > 0 ? 148 * s->tune[5] : grid;
#define CVB_SP2(G, U) do { if (multi) CVB_CHECK(launch_pdl(edge_pass_sparse2_kernel<true, G, U>, g2, kThreads, 0, st, s->rec, ep, adj, ents, nt, s->infect_key, s->cand, s->n_cand, work_row)); \
                           else CVB_CHECK(launch_pdl(edge_pass_sparse2_kernel<false, G, U>, g2, kThreads, 0, st, s->rec, ep, adj, ents, nt, s->infect_key, s->cand, s->n_cand, work_row)); } while (0)
            switch (shape) {
                case 1:  CVB_SP2(8, 6); break;
                case 2:  CVB_SP2(16, 2); break;
                case 3:  CVB_SP2(16, 3); break;
                case 5:  CVB_SP2(32, 2); break;
                case 6:  CVB_SP2(4, 8); break;
                case 7:  CVB_SP2(8, 4); break;
                default: CVB_SP2(16, 4); break;          // measured best on the C2 workload (profiles/r2/README.md)
            }
#undef CVB_SP2
        }
        else if (multi) edge_pass_sparse_kernel<true><<<grid, kThreads, 0, st>>>(s->rec, ep, s->adj_ptr, s->adj, s->trans_list, s->n_trans,
                                                                               s->infect_key, s->cand, s->n_cand, work_row);
        else edge_pass_sparse_kernel<false><<<grid, kThreads, 0, st>>>(s->rec, ep, s->adj_ptr, s->adj, s->trans_list, s->n_trans,
                                                                      s->infect_key, s->cand, s->n_cand, work_row);
        CVB_LAUNCH_CHECK();
        bool rest = false;
        for (int l = 0; l < s->pars.n_layers; ++l) rest |= !(skip_mask & (1u << l)) && s->layers[l].n_edges > 0;
        if (!rest) return 0;
    }
    // dense streaming pass over the remaining (dynamic) layers.  Shared-memory budget: per-warp candidate queue (16 B x 160
    // entries) + the transmit bitmap; one persistent CTA per SM while the bitmap fits, else L1/L2-cached bit tests.
    bool rest = false;
    for (int l = 0; l < s->pars.n_layers; ++l) rest |= !(skip_mask & (1u << l)) && s->layers[l].n_edges > 0;
    if (!rest) return 0;
    if (!multi) {
        for (int l = 0; l < s->pars.n_layers; ++l)
            CVB_REQUIRE((skip_mask >> l) & 1u || s->layers[l].n_edges == 0 || (s->rec.ts8 && ((s->rec.ts8_mask >> l) & 1u)),
                        "cvb_edge_pass: layer %d changed between cvb_prepare_transmission and cvb_edge_pass (its per-layer records are missing)", l);
    }
    const size_t bitmap_bytes = (size_t)ep.n_words * sizeof(unsigned int);
    const size_t limit = 227 * 1024;
    auto queue_bytes = [](int threads) { return (size_t)(threads / 32) * kQueueCap * sizeof(uint4); };
    // Tile shape chosen on the B200 (profiles/r1/README.md, "dense edge pass v4"): 512 threads x 2 quads per thread beats 1024 x 1
    // (more independent loads per warp: the pass is latency bound, not issue bound), L2 prefetch one tile ahead.
    // CVB_DENSE_VARIANT selects the other shapes for profiling.
    static int variant = -1;
    if (variant < 0) { const char* v = getenv("CVB_DENSE_VARIANT"); variant = v ? atoi(v) : 0; }
#define CVB_DENSE(THREADS, QPT, SMEM, CTAS, PF) \
    return multi ? launch_edge_pass<true, SMEM, THREADS, QPT, PF>(s, skip_mask, ep, bitmap_bytes, CTAS, st) \
                 : launch_edge_pass<false, SMEM, THREADS, QPT, PF>(s, skip_mask, ep, bitmap_bytes, CTAS, st)
    if (variant == 3) {                                               // tiles staged by bulk copies (cp.async.bulk + mbarrier): measured SLOWER, kept for profiling
        const int rc = multi ? launch_edge_pass_tma<true, 20>(s, skip_mask, ep, bitmap_bytes, st, 2) : launch_edge_pass_tma<false, 20>(s, skip_mask, ep, bitmap_bytes, st, 2);
        if (rc >= 0) return rc;
    }
    if (bitmap_bytes + queue_bytes(1024) <= limit && variant == 1) { CVB_DENSE(1024, 1, true, 1, 1); }
    if (bitmap_bytes + queue_bytes(512) <= limit) {
        if (variant == 2) { CVB_DENSE(512, 2, true, 1, 0); }
        CVB_DENSE(512, 2, true, 1, 1);
    }
    // large populations: bitmap stays in global memory (L1/L2-cached bit tests), 256-thread CTAs
    CVB_DENSE(256, 1, false, 4, 1);
#undef CVB_DENSE
}
