// The fused transmission pass (native-RNG mode): reference sim.py:622-649 + utils.py:93-128.
//
// The reference makes n_variants x n_layers x 2 passes over the edge lists per day (compute_trans_sus
// + compute_infections per variant and layer, both directions) and re-gathers per-agent values with
// fancy indexing each time.  Here every edge (p1:int32, p2:int32, beta:f32 = 12 bytes) of every layer
// is streamed from HBM exactly ONCE per day with 128-bit coalesced loads; both directions and all
// variants are evaluated from the two 8-byte {rel_trans, rel_sus} records of its endpoints (written by
// prepare_transmission; L2-resident gathers).  Algorithmic bytes per day: 12*E_total + 8*N.
//
// Randomness: one Philox4x32-10 call per LIVE edge (non-zero probability in either direction), keyed
// (seed, P_EDGE, layer, day, edge index): words 0-1 give the p1->p2 uniform, words 2-3 the p2->p1 one.
// Dead edges (>= 85 % of them even at the epidemic peak) cost no RNG work.
//
// Output: for each target hit at least once, infect_key[target] = min over successful transmissions
// of (variant, layer, direction, edge) -- exactly the reference's winner: variants and layers are
// processed in order, infect() drops already-infected targets, and np.unique(return_index=True)
// keeps the first occurrence in [direction 0 edges..., direction 1 edges...] (people.py:465-473).
// The first hit on a target also appends it to the day's candidate list (warp-aggregated atomics).
#include "cvb_internal.cuh"

namespace cvb {

struct EdgeParams {
    float beta[CVB_MAX_VARIANTS];
    uint64_t seed;
    int64_t n;
    int32_t t, nv;
};

__device__ __forceinline__ void record_hit(unsigned long long* __restrict__ infect_key, int32_t* __restrict__ cand,
                                           unsigned int* __restrict__ n_cand, int target, unsigned long long key) {
    unsigned long long old = atomicMin(infect_key + target, key);
    if (old == kEmptyKey) {
        unsigned int pos = warp_append32(n_cand);
        cand[pos] = target;
    }
}

template <bool MULTI>
__global__ void __launch_bounds__(kThreads) edge_pass_kernel(const __grid_constant__ LayerTable L, TransRecords rec,
        const __grid_constant__ EdgeParams ep, unsigned long long* __restrict__ infect_key, int32_t* __restrict__ cand,
        unsigned int* __restrict__ n_cand) {
    const int64_t n = ep.n;
    const int64_t total_tiles = L.tile_start[L.n_layers];
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int l = 0;
#pragma unroll
        for (int q = 1; q < CVB_MAX_LAYERS; ++q) l += (q < L.n_layers && tile >= L.tile_start[q]);
        const LayerPtrs& lay = L.l[l];
        const int64_t e0 = (tile - L.tile_start[l]) * kTileEdges + (int64_t)threadIdx.x * kEdgesPerThread;
        if (e0 >= lay.n_edges) continue;
        int a[4], b[4];
        float w[4];
        int cnt;
        if (e0 + 4 <= lay.n_edges) {
            int4 va = ld_stream(reinterpret_cast<const int4*>(lay.p1 + e0));
            int4 vb = ld_stream(reinterpret_cast<const int4*>(lay.p2 + e0));
            float4 vw = ld_stream(reinterpret_cast<const float4*>(lay.beta + e0));
            a[0] = va.x; a[1] = va.y; a[2] = va.z; a[3] = va.w;
            b[0] = vb.x; b[1] = vb.y; b[2] = vb.z; b[3] = vb.w;
            w[0] = vw.x; w[1] = vw.y; w[2] = vw.z; w[3] = vw.w;
            cnt = 4;
        } else {
            cnt = (int)(lay.n_edges - e0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                bool ok = k < cnt;
                a[k] = ok ? lay.p1[e0 + k] : 0; b[k] = ok ? lay.p2[e0 + k] : 0; w[k] = ok ? lay.beta[e0 + k] : 0.0f;
            }
        }
        const float2* __restrict__ ts = rec.ts + (int64_t)l * n;
        // issue all eight gathers before using any of them (memory-level parallelism, Guideline 7)
        float2 ra[4], rb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { ra[k] = __ldg(ts + a[k]); rb[k] = __ldg(ts + b[k]); }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k >= cnt) break;
            float p01 = 0.0f, p10 = 0.0f;
            int va_ = 0, vb_ = 0;
            if (ra[k].x != 0.0f) {                       // a can transmit
                float sb = rb[k].y;
                if (MULTI) {
                    va_ = rec.ivar[a[k]];
                    if (va_ > 0) sb = __ldg(rec.sus_extra + ((int64_t)l * (ep.nv - 1) + (va_ - 1)) * n + b[k]);
                }
                p01 = edge_prob(ep.beta[va_], w[k], ra[k].x, sb);
            }
            if (rb[k].x != 0.0f) {                       // b can transmit
                float sa = ra[k].y;
                if (MULTI) {
                    vb_ = rec.ivar[b[k]];
                    if (vb_ > 0) sa = __ldg(rec.sus_extra + ((int64_t)l * (ep.nv - 1) + (vb_ - 1)) * n + a[k]);
                }
                p10 = edge_prob(ep.beta[vb_], w[k], rb[k].x, sa);
            }
            if (p01 != 0.0f || p10 != 0.0f) {
                const int64_t e = e0 + k;
                u32x4 r = keyed_words(ep.seed, P_EDGE, (uint32_t)l, ep.t, e, 0);
                unsigned long long base = ((unsigned long long)l << 48) | (unsigned long long)e;
                if (p01 != 0.0f && u53(r.x, r.y) < (double)p01)
                    record_hit(infect_key, cand, n_cand, b[k], ((unsigned long long)va_ << 56) | base);
                if (p10 != 0.0f && u53(r.z, r.w) < (double)p10)
                    record_hit(infect_key, cand, n_cand, a[k], ((unsigned long long)vb_ << 56) | (1ull << 40) | base);
            }
        }
    }
}

}  // namespace cvb

using namespace cvb;

namespace cvb { int build_layer_table(cvb_sim* s, LayerTable& L); }

int cvb::build_layer_table(cvb_sim* s, LayerTable& L) {
    L.n_layers = s->pars.n_layers;
    int64_t acc = 0;
    for (int l = 0; l < CVB_MAX_LAYERS; ++l) {
        L.tile_start[l] = acc;
        if (l < L.n_layers) {
            L.l[l] = s->layers[l];
            CVB_REQUIRE(L.l[l].n_edges == 0 || L.l[l].p1, "layer %d is not bound (cvb_bind_layer)", l);
            CVB_REQUIRE(L.l[l].n_edges < (1ll << 40), "layer %d has too many edges for the 40-bit edge field", l);
            acc += (L.l[l].n_edges + kTileEdges - 1) / kTileEdges;
        } else {
            L.l[l] = LayerPtrs{nullptr, nullptr, nullptr, 0};
        }
    }
    for (int l = L.n_layers; l <= CVB_MAX_LAYERS; ++l) L.tile_start[l] = acc;
    return 0;
}

extern "C" int cvb_edge_pass(cvb_sim* s, int32_t t, cvb_stream st) {
    CVB_REQUIRE(s && s->pars_set, "cvb_edge_pass: handle not ready");
    CVB_REQUIRE(s->rec.ts && s->rec_layers >= s->pars.n_layers, "cvb_edge_pass: call cvb_prepare_transmission first");
    LayerTable L;
    if (build_layer_table(s, L)) return 1;
    int64_t total_tiles = L.tile_start[L.n_layers];
    if (total_tiles == 0) return 0;
    EdgeParams ep;
    for (int v = 0; v < CVB_MAX_VARIANTS; ++v) ep.beta[v] = s->pars.beta[v];
    ep.seed = s->seed; ep.n = s->n; ep.t = t; ep.nv = s->nv;
    // persistent-style grid: a multiple of the SM count, 8 resident CTAs of 256 threads per SM
    int grid = (int)(total_tiles < 148 * 8 ? total_tiles : 148 * 8);
    if (s->nv > 1)
        edge_pass_kernel<true><<<grid, kThreads, 0, (cudaStream_t)st>>>(L, s->rec, ep, s->infect_key, s->cand, s->n_cand);
    else
        edge_pass_kernel<false><<<grid, kThreads, 0, (cudaStream_t)st>>>(L, s->rec, ep, s->infect_key, s->cand, s->n_cand);
    CVB_LAUNCH_CHECK();
    return 0;
}
