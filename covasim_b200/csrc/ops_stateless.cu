// Stateless operators: device forms of the reference's four Numba kernels (covasim/utils.py:39-147)
// plus the ordered stream compaction they and the People index helpers (utils.py:494-506) need.
//
// All are HBM-streaming kernels: 128-bit coalesced loads of the edge arrays, L2-resident gathers of
// the per-agent values, one CTA = 256 threads x 4 edges.
#include "cvb_internal.cuh"

namespace cvb {

// ---- A2 ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) viral_load_kernel(int32_t t, const float* __restrict__ d_inf,
        const float* __restrict__ d_rec, const float* __restrict__ d_dead, float frac_time, float load_ratio,
        float high_cap, float* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = viral_load(t, d_inf[i], d_rec[i], d_dead[i], frac_time, load_ratio, high_cap);
}

// ---- A3 ---------------------------------------------------------------------------------------
// keyed uniforms for the device-side population generator (covasim_b200/population.py:make_keyed_pop)
__global__ void __launch_bounds__(kThreads) keyed_uniform_kernel(uint64_t seed, uint32_t purpose, uint32_t sub, int32_t day, int64_t index0,
        int64_t n, uint32_t slot, double* __restrict__ out) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        out[k] = keyed_uniform(seed, purpose, sub, day, index0 + k, slot);
}

__global__ void __launch_bounds__(kThreads) trans_sus_kernel(const float* __restrict__ rel_trans, const float* __restrict__ rel_sus,
        const uint8_t* __restrict__ inf, const uint8_t* __restrict__ sus, float beta_layer, const float* __restrict__ vload,
        const uint8_t* __restrict__ symp, const uint8_t* __restrict__ iso, const uint8_t* __restrict__ quar,
        float asymp_factor, float iso_factor, float quar_factor, const float* __restrict__ imm,
        float* __restrict__ out_trans, float* __restrict__ out_sus, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        bool q = quar[i] != 0;
        out_trans[i] = rel_trans_layer(rel_trans[i], inf[i] != 0, symp[i] != 0, iso[i] != 0, q, asymp_factor, iso_factor,
                                       quar_factor, beta_layer, vload[i]);
        out_sus[i] = rel_sus_layer(rel_sus[i], sus[i] != 0, q, quar_factor, imm[i]);
    }
}

// ---- edge loading -------------------------------------------------------------------------------
struct Edge4 { int a[4]; int b[4]; float w[4]; int n; };

__device__ __forceinline__ Edge4 load_edge4(const int32_t* __restrict__ p1, const int32_t* __restrict__ p2,
                                            const float* __restrict__ beta, int64_t e0, int64_t n_edges, bool vec_ok) {
    Edge4 r;
    if (e0 + 4 <= n_edges && vec_ok) {
        int4 a = ld_stream(reinterpret_cast<const int4*>(p1 + e0));
        int4 b = ld_stream(reinterpret_cast<const int4*>(p2 + e0));
        float4 w = ld_stream(reinterpret_cast<const float4*>(beta + e0));
        r.a[0] = a.x; r.a[1] = a.y; r.a[2] = a.z; r.a[3] = a.w;
        r.b[0] = b.x; r.b[1] = b.y; r.b[2] = b.z; r.b[3] = b.w;
        r.w[0] = w.x; r.w[1] = w.y; r.w[2] = w.z; r.w[3] = w.w;
        r.n = 4;
    } else {
        r.n = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (e0 + k < n_edges) { r.a[k] = p1[e0 + k]; r.b[k] = p2[e0 + k]; r.w[k] = beta[e0 + k]; r.n = k + 1; }
            else { r.a[k] = 0; r.b[k] = 0; r.w[k] = 0.0f; }
        }
    }
    return r;
}

// Probabilities of the (up to) four edges of a thread, both directions (reference utils.py:115-118)
__device__ __forceinline__ void edge4_probs(const Edge4& e, float beta, const float* __restrict__ rel_trans,
                                            const float* __restrict__ rel_sus, float p01[4], float p10[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        p01[k] = 0.0f; p10[k] = 0.0f;
        if (k < e.n) {
            float ta = __ldg(rel_trans + e.a[k]), tb = __ldg(rel_trans + e.b[k]);
            // the reference drops zero-transmissibility sources before forming the product (utils.py:113-114)
            if (ta != 0.0f) p01[k] = edge_prob(beta, e.w[k], ta, __ldg(rel_sus + e.b[k]));
            if (tb != 0.0f) p10[k] = edge_prob(beta, e.w[k], tb, __ldg(rel_sus + e.a[k]));
        }
    }
}

// ---- A4 (replay form), pass 1: how many draws does each tile consume per direction ---------------
__global__ void __launch_bounds__(kThreads) infections_count_kernel(float beta, const int32_t* __restrict__ p1,
        const int32_t* __restrict__ p2, const float* __restrict__ lbeta, int64_t n_edges, const float* __restrict__ rel_trans,
        const float* __restrict__ rel_sus, unsigned int* __restrict__ tile_cnt, int64_t n_tiles, bool vec_ok) {
    __shared__ int s_cnt[2];
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        int64_t e0 = tile * kTileEdges + (int64_t)threadIdx.x * kEdgesPerThread;
        Edge4 e = load_edge4(p1, p2, lbeta, e0, n_edges, vec_ok);
        float p01[4], p10[4];
        edge4_probs(e, beta, rel_trans, rel_sus, p01, p10);
        int c0 = 0, c1 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { c0 += (p01[k] != 0.0f); c1 += (p10[k] != 0.0f); }
        int w0 = __reduce_add_sync(0xFFFFFFFFu, c0), w1 = __reduce_add_sync(0xFFFFFFFFu, c1);
        if (lane_id() == 0) { if (w0) atomicAdd(&s_cnt[0], w0); if (w1) atomicAdd(&s_cnt[1], w1); }
        __syncthreads();
        if (threadIdx.x == 0) { tile_cnt[tile] = s_cnt[0]; tile_cnt[n_tiles + tile] = s_cnt[1]; }
        __syncthreads();
    }
}

// ---- A4 pass 2: consume the uniforms in the reference's order, record hits ----------------------
// tile_off = exclusive scan of tile_cnt over [direction 0 tiles..., direction 1 tiles...]: the rank of a
// surviving edge-direction in the reference's draw order is tile_off + (rank inside the tile).
__global__ void __launch_bounds__(kThreads) infections_draw_kernel(float beta, const int32_t* __restrict__ p1,
        const int32_t* __restrict__ p2, const float* __restrict__ lbeta, int64_t n_edges, const float* __restrict__ rel_trans,
        const float* __restrict__ rel_sus, const unsigned int* __restrict__ tile_off, const double* __restrict__ uniforms,
        uint8_t* __restrict__ hit_mask, unsigned int* __restrict__ hit_cnt, int64_t n_tiles, bool vec_ok) {
    __shared__ int warp_sums[kThreads / 32 + 1];
    __shared__ int s_cnt[2];
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
        int64_t e0 = tile * kTileEdges + (int64_t)threadIdx.x * kEdgesPerThread;
        Edge4 e = load_edge4(p1, p2, lbeta, e0, n_edges, vec_ok);
        float p01[4], p10[4];
        edge4_probs(e, beta, rel_trans, rel_sus, p01, p10);
        int c0 = 0, c1 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { c0 += (p01[k] != 0.0f); c1 += (p10[k] != 0.0f); }
        int tot;
        int r0 = block_exclusive_scan(c0, warp_sums, tot);
        int r1 = block_exclusive_scan(c1, warp_sums, tot);
        int64_t base0 = (int64_t)tile_off[tile] + r0, base1 = (int64_t)tile_off[n_tiles + tile] + r1;
        unsigned m = 0;
        int h0 = 0, h1 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (p01[k] != 0.0f) { if (uniforms[base0++] < (double)p01[k]) { m |= 1u << k; ++h0; } }
            if (p10[k] != 0.0f) { if (uniforms[base1++] < (double)p10[k]) { m |= 16u << k; ++h1; } }
        }
        hit_mask[tile * kThreads + threadIdx.x] = (uint8_t)m;
        int w0 = __reduce_add_sync(0xFFFFFFFFu, h0), w1 = __reduce_add_sync(0xFFFFFFFFu, h1);
        if (lane_id() == 0) { if (w0) atomicAdd(&s_cnt[0], w0); if (w1) atomicAdd(&s_cnt[1], w1); }
        __syncthreads();
        if (threadIdx.x == 0) { hit_cnt[tile] = s_cnt[0]; hit_cnt[n_tiles + tile] = s_cnt[1]; }
        __syncthreads();
    }
}

// ---- A4 pass 3: ordered (source, target) lists, direction 0 first (reference utils.py:124-127) ----
__global__ void __launch_bounds__(kThreads) infections_write_kernel(const int32_t* __restrict__ p1, const int32_t* __restrict__ p2,
        int64_t n_edges, const uint8_t* __restrict__ hit_mask, const unsigned int* __restrict__ hit_cnt_raw_next,
        const unsigned int* __restrict__ hit_off, int32_t* __restrict__ out_src, int32_t* __restrict__ out_tgt, int64_t n_tiles) {
    __shared__ int warp_sums[kThreads / 32 + 1];
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        unsigned m = hit_mask[tile * kThreads + threadIdx.x];
        int any = __syncthreads_or((int)m);
        if (!any) continue;
        int tot;
        int r0 = block_exclusive_scan(__popc(m & 15u), warp_sums, tot);
        int r1 = block_exclusive_scan(__popc(m >> 4), warp_sums, tot);
        int64_t o0 = (int64_t)hit_off[tile] + r0, o1 = (int64_t)hit_off[n_tiles + tile] + r1;
        int64_t e0 = tile * kTileEdges + (int64_t)threadIdx.x * kEdgesPerThread;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (m & (1u << k)) { out_src[o0] = p1[e0 + k]; out_tgt[o0] = p2[e0 + k]; ++o0; }
            if (m & (16u << k)) { out_src[o1] = p2[e0 + k]; out_tgt[o1] = p1[e0 + k]; ++o1; }
        }
    }
    (void)hit_cnt_raw_next;
}

// ---- ordered compaction of byte flags -> ascending indices (utils.py:494-506 true()) ---------------
__global__ void __launch_bounds__(kThreads) flag_count_kernel(const uint8_t* __restrict__ flags, int64_t n,
                                                              unsigned int* __restrict__ tile_cnt, int64_t n_tiles) {
    __shared__ int s_cnt;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        int64_t i0 = tile * kTileEdges + (int64_t)threadIdx.x * 4;
        int c = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) c += (i0 + k < n && flags[i0 + k] != 0);
        int w = __reduce_add_sync(0xFFFFFFFFu, c);
        if (lane_id() == 0 && w) atomicAdd(&s_cnt, w);
        __syncthreads();
        if (threadIdx.x == 0) tile_cnt[tile] = s_cnt;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads) flag_write_kernel(const uint8_t* __restrict__ flags, int64_t n,
        const unsigned int* __restrict__ tile_off, int32_t* __restrict__ out, int64_t n_tiles) {
    __shared__ int warp_sums[kThreads / 32 + 1];
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int64_t i0 = tile * kTileEdges + (int64_t)threadIdx.x * 4;
        unsigned m = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) if (i0 + k < n && flags[i0 + k] != 0) m |= 1u << k;
        int tot;
        int r = block_exclusive_scan(__popc(m), warp_sums, tot);
        int64_t o = (int64_t)tile_off[tile] + r;
#pragma unroll
        for (int k = 0; k < 4; ++k) if (m & (1u << k)) out[o++] = (int32_t)(i0 + k);
    }
}

// ---- A13: contacts of a set of agents -----------------------------------------------------------
__global__ void mark_members_kernel(const int64_t* __restrict__ inds, int64_t n_inds, uint8_t* __restrict__ member, int64_t n) {
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_inds; j += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = inds[j];
        if (i >= 0 && i < n) member[i] = 1;
    }
}

__global__ void __launch_bounds__(kThreads) find_contacts_kernel(const int32_t* __restrict__ p1, const int32_t* __restrict__ p2,
        int64_t n_edges, const uint8_t* __restrict__ member, uint8_t* __restrict__ found, bool vec_ok) {
    int64_t n_tiles = (n_edges + kTileEdges - 1) / kTileEdges;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int64_t e0 = tile * kTileEdges + (int64_t)threadIdx.x * kEdgesPerThread;
        int a[4], b[4], cnt = 0;
        if (e0 + 4 <= n_edges && vec_ok) {
            int4 va = ld_stream(reinterpret_cast<const int4*>(p1 + e0)), vb = ld_stream(reinterpret_cast<const int4*>(p2 + e0));
            a[0] = va.x; a[1] = va.y; a[2] = va.z; a[3] = va.w; b[0] = vb.x; b[1] = vb.y; b[2] = vb.z; b[3] = vb.w; cnt = 4;
        } else {
            for (int k = 0; k < 4; ++k) if (e0 + k < n_edges) { a[k] = p1[e0 + k]; b[k] = p2[e0 + k]; cnt = k + 1; }
        }
        for (int k = 0; k < cnt; ++k) {
            if (__ldg(member + a[k])) found[b[k]] = 1;
            if (__ldg(member + b[k])) found[a[k]] = 1;
        }
    }
}

static int compact_flags(cvb_sim* s, const uint8_t* flags, int64_t n, int32_t* out, int64_t* host_n_out, cudaStream_t st) {
    int64_t n_tiles = (n + kTileEdges - 1) / kTileEdges;
    if (ensure_u32(&s->tile_cnt, &s->tile_cnt_cap, n_tiles + 1)) return 1;
    flag_count_kernel<<<grid_for(n_tiles, 1), kThreads, 0, st>>>(flags, n, s->tile_cnt, n_tiles);
    CVB_LAUNCH_CHECK();
    if (exclusive_scan_u32(s->tile_cnt, n_tiles, s->dev_scalars, st)) return 1;
    flag_write_kernel<<<grid_for(n_tiles, 1), kThreads, 0, st>>>(flags, n, s->tile_cnt, out, n_tiles);
    CVB_LAUNCH_CHECK();
    CVB_CHECK(cudaMemcpyAsync(s->host_scalars, s->dev_scalars, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CVB_CHECK(cudaStreamSynchronize(st));
    *host_n_out = (int64_t)s->host_scalars[0];
    return 0;
}

static inline bool aligned16(const void* a, const void* b, const void* c) {
    return ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)c)) & 15) == 0;
}

}  // namespace cvb

using namespace cvb;

extern "C" {

int cvb_keyed_uniform(uint64_t seed, uint32_t purpose, uint32_t sub, int32_t day, int64_t index0, int64_t n, uint32_t slot,
                      double* out, cvb_stream st) {
    CVB_REQUIRE(n >= 0 && (n == 0 || out), "cvb_keyed_uniform: bad argument");
    if (n == 0) return 0;
    keyed_uniform_kernel<<<grid_for(n, kThreads, 148 * 16), kThreads, 0, (cudaStream_t)st>>>(seed, purpose, sub, day, index0, n, slot, out);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_compute_viral_load(int32_t t, const float* date_inf, const float* date_rec, const float* date_dead,
                           float frac_time, float load_ratio, float high_cap, float* out, int64_t n, cvb_stream st) {
    CVB_REQUIRE(n >= 0 && (n == 0 || (date_inf && date_rec && date_dead && out)), "cvb_compute_viral_load: bad argument");
    if (n == 0) return 0;
    viral_load_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)st>>>(t, date_inf, date_rec, date_dead, frac_time, load_ratio, high_cap, out, n);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_compute_trans_sus(const float* rel_trans, const float* rel_sus, const uint8_t* inf, const uint8_t* sus,
                          float beta_layer, const float* viral_load, const uint8_t* symp, const uint8_t* iso,
                          const uint8_t* quar, float asymp_factor, float iso_factor, float quar_factor,
                          const float* immunity_factors, float* out_trans, float* out_sus, int64_t n, cvb_stream st) {
    CVB_REQUIRE(n >= 0, "cvb_compute_trans_sus: negative n");
    if (n == 0) return 0;
    CVB_REQUIRE(rel_trans && rel_sus && inf && sus && viral_load && symp && iso && quar && immunity_factors && out_trans && out_sus,
                "cvb_compute_trans_sus: NULL array");
    trans_sus_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)st>>>(rel_trans, rel_sus, inf, sus, beta_layer, viral_load, symp, iso, quar,
                                                                    asymp_factor, iso_factor, quar_factor, immunity_factors, out_trans, out_sus, n);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_infections_count(cvb_sim* s, float beta, const int32_t* p1, const int32_t* p2, const float* layer_betas,
                         int64_t n_edges, const float* rel_trans, const float* rel_sus, int64_t* host_n_draws, cvb_stream st_) {
    cudaStream_t st = (cudaStream_t)st_;
    CVB_REQUIRE(s && host_n_draws, "cvb_infections_count: NULL argument");
    host_n_draws[0] = host_n_draws[1] = 0;
    if (n_edges == 0) return 0;
    CVB_REQUIRE(p1 && p2 && layer_betas && rel_trans && rel_sus, "cvb_infections_count: NULL array");
    int64_t n_tiles = (n_edges + kTileEdges - 1) / kTileEdges;
    if (ensure_u32(&s->tile_cnt, &s->tile_cnt_cap, 4 * n_tiles + 4)) return 1;
    infections_count_kernel<<<grid_for(n_tiles, 1), kThreads, 0, st>>>(beta, p1, p2, layer_betas, n_edges, rel_trans, rel_sus,
                                                                       s->tile_cnt, n_tiles, aligned16(p1, p2, layer_betas));
    CVB_LAUNCH_CHECK();
    // keep the raw per-direction totals: scan direction 0 and direction 1 tiles as one sequence, and
    // read the boundary value (offset of the first direction-1 tile) to split the total
    if (exclusive_scan_u32(s->tile_cnt, 2 * n_tiles, s->dev_scalars, st)) return 1;
    CVB_CHECK(cudaMemcpyAsync(s->host_scalars, s->dev_scalars, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CVB_CHECK(cudaMemcpyAsync(s->host_scalars + 1, s->tile_cnt + n_tiles, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    CVB_CHECK(cudaStreamSynchronize(st));
    unsigned int first_dir1 = *(unsigned int*)(s->host_scalars + 1);
    host_n_draws[0] = (int64_t)first_dir1;
    host_n_draws[1] = (int64_t)s->host_scalars[0] - (int64_t)first_dir1;
    return 0;
}

int cvb_infections_draw(cvb_sim* s, float beta, const int32_t* p1, const int32_t* p2, const float* layer_betas,
                        int64_t n_edges, const float* rel_trans, const float* rel_sus, const double* uniforms,
                        int32_t* out_src, int32_t* out_tgt, int64_t* host_n_out, cvb_stream st_) {
    cudaStream_t st = (cudaStream_t)st_;
    CVB_REQUIRE(s && host_n_out, "cvb_infections_draw: NULL argument");
    *host_n_out = 0;
    if (n_edges == 0) return 0;
    CVB_REQUIRE(p1 && p2 && layer_betas && rel_trans && rel_sus, "cvb_infections_draw: NULL array");
    int64_t n_tiles = (n_edges + kTileEdges - 1) / kTileEdges;
    CVB_REQUIRE(s->tile_cnt && s->tile_cnt_cap >= 4 * n_tiles + 4, "cvb_infections_draw: call cvb_infections_count on the same layer first");
    if (ensure_u8(&s->hit_mask, &s->hit_mask_cap, n_tiles * kThreads)) return 1;
    unsigned int* hit_cnt = s->tile_cnt + 2 * n_tiles;
    infections_draw_kernel<<<grid_for(n_tiles, 1), kThreads, 0, st>>>(beta, p1, p2, layer_betas, n_edges, rel_trans, rel_sus, s->tile_cnt,
                                                                      uniforms, s->hit_mask, hit_cnt, n_tiles, aligned16(p1, p2, layer_betas));
    CVB_LAUNCH_CHECK();
    if (exclusive_scan_u32(hit_cnt, 2 * n_tiles, s->dev_scalars, st)) return 1;
    infections_write_kernel<<<grid_for(n_tiles, 1), kThreads, 0, st>>>(p1, p2, n_edges, s->hit_mask, nullptr, hit_cnt, out_src, out_tgt, n_tiles);
    CVB_LAUNCH_CHECK();
    CVB_CHECK(cudaMemcpyAsync(s->host_scalars, s->dev_scalars, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CVB_CHECK(cudaStreamSynchronize(st));
    *host_n_out = (int64_t)s->host_scalars[0];
    return 0;
}

int cvb_true_indices(cvb_sim* s, const uint8_t* flags, int64_t n, int32_t* out, int64_t* host_n_out, cvb_stream st) {
    CVB_REQUIRE(s && host_n_out, "cvb_true_indices: NULL argument");
    *host_n_out = 0;
    if (n == 0) return 0;
    CVB_REQUIRE(flags && out, "cvb_true_indices: NULL array");
    return compact_flags(s, flags, n, out, host_n_out, (cudaStream_t)st);
}

int cvb_find_contacts(cvb_sim* s, const int32_t* p1, const int32_t* p2, int64_t n_edges, const int64_t* inds,
                      int64_t n_inds, int32_t* out, int64_t* host_n_out, cvb_stream st_) {
    cudaStream_t st = (cudaStream_t)st_;
    CVB_REQUIRE(s && host_n_out, "cvb_find_contacts: NULL argument");
    *host_n_out = 0;
    if (n_edges == 0 || n_inds == 0) return 0;
    CVB_REQUIRE(p1 && p2 && inds && out, "cvb_find_contacts: NULL array");
    if (ensure_u8(&s->flag_tmp, &s->flag_tmp_cap, 2 * s->n)) return 1;
    uint8_t* member = s->flag_tmp;
    uint8_t* found = s->flag_tmp + s->n;
    CVB_CHECK(cudaMemsetAsync(s->flag_tmp, 0, (size_t)(2 * s->n), st));
    mark_members_kernel<<<grid_for(n_inds), kThreads, 0, st>>>(inds, n_inds, member, s->n);
    CVB_LAUNCH_CHECK();
    int64_t n_tiles = (n_edges + kTileEdges - 1) / kTileEdges;
    find_contacts_kernel<<<grid_for(n_tiles, 1), kThreads, 0, st>>>(p1, p2, n_edges, member, found, aligned16(p1, p2, p1));
    CVB_LAUNCH_CHECK();
    return compact_flags(s, found, s->n, out, host_n_out, st);
}

}  // extern "C"
