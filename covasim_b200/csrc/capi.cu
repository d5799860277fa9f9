// Handle management, binding and small shared utilities of libcovasim_b200.so.
#include <stdarg.h>
#include <string.h>
#include <new>
#include "cvb_internal.cuh"

namespace cvb {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

template <typename T> static int ensure(T** p, int64_t* cap, int64_t need) {
    if (need <= *cap) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    int64_t newcap = need + need / 4 + 1024;
    CVB_CHECK(cudaMalloc((void**)p, (size_t)newcap * sizeof(T)));
    *cap = newcap;
    return 0;
}
int ensure_u32(unsigned int** p, int64_t* cap, int64_t need) { return ensure(p, cap, need); }
int ensure_u8(uint8_t** p, int64_t* cap, int64_t need) { return ensure(p, cap, need); }
int ensure_f64(double** p, int64_t* cap, int64_t need) { return ensure(p, cap, need); }

// ---- exclusive scan of uint32 counts, in place (single CTA; inputs are per-tile counts, <= ~1e5) ----
__global__ void __launch_bounds__(1024) scan_u32_kernel(unsigned int* data, int64_t n, unsigned long long* total_out) {
    __shared__ int warp_sums[33];
    unsigned long long running = 0;
    for (int64_t base = 0; base < n; base += blockDim.x) {
        int64_t i = base + threadIdx.x;
        int v = i < n ? (int)data[i] : 0;
        int total;
        int excl = block_exclusive_scan(v, warp_sums, total);
        if (i < n) data[i] = (unsigned int)(running + (unsigned long long)excl);
        running += (unsigned long long)total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = running;
}

int exclusive_scan_u32(unsigned int* data, int64_t n, unsigned long long* total_out, cudaStream_t st) {
    scan_u32_kernel<<<1, 1024, 0, st>>>(data, n, total_out);
    CVB_LAUNCH_CHECK();
    return 0;
}

// All-gather over peer memory (agent-partitioned runs on one NVSwitch node): every rank stores its chunk straight into the exchange
// buffer of every rank (its own included) -- W coalesced 16-byte stores per item, the remote ones travel over NVLink -- so the day's
// exchange is this one kernel plus a signal barrier instead of a library collective (covasim_b200/partition.py:PeerExchange)
constexpr int kMaxPeers = 16;
struct PeerPtrs { unsigned char* p[kMaxPeers]; };
template <typename V>
__global__ void __launch_bounds__(kThreads) peer_push_kernel(const V* __restrict__ src, int64_t items, const __grid_constant__ PeerPtrs P, int world) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < items; k += (int64_t)gridDim.x * blockDim.x) {
        const V v = __ldg(src + k);
        for (int p = 0; p < world; ++p) reinterpret_cast<V*>(P.p[p])[k] = v;
    }
}

// Compact restore of the People arena (Sim.restore of a compact snapshot): most per-agent arrays of a saved state are one value
// almost everywhere (unset dates, cleared flags), so the snapshot keeps {segment, fill value} + the exceptions and only the
// genuinely dense arrays travel over PCIe.  Segments are 32-bit words of the arena; blockIdx.y = segment.
__global__ void __launch_bounds__(kThreads) fill_segments_kernel(uint32_t* __restrict__ words, const long long* __restrict__ seg) {
    const long long begin = seg[3 * blockIdx.y], count = seg[3 * blockIdx.y + 1];
    const uint32_t v = (uint32_t)seg[3 * blockIdx.y + 2];
    uint32_t* p = words + begin;
    const long long n4 = ((uintptr_t)p & 15) == 0 ? count / 4 : 0;
    const uint4 v4 = make_uint4(v, v, v, v);
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n4; k += (long long)gridDim.x * blockDim.x) reinterpret_cast<uint4*>(p)[k] = v4;
    for (long long k = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; k < count; k += (long long)gridDim.x * blockDim.x) p[k] = v;
}
__global__ void __launch_bounds__(kThreads) scatter_words_kernel(uint32_t* __restrict__ words, const long long* __restrict__ idx,
                                                                 const long long* __restrict__ val, int64_t n) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) words[idx[k]] = (uint32_t)val[k];
}

__global__ void fill_f32_kernel(float* p, int64_t n, float v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace cvb

using namespace cvb;

extern "C" {

const char* cvb_last_error(void) { return g_err; }
int32_t cvb_abi_version(void) { return CVB_ABI_VERSION; }
int64_t cvb_launch_count(void) { return (int64_t)g_launches; }

/* sizes of the by-value structs, so a binding can verify its own layout */
int cvb_struct_sizes(int64_t* out5) {
    CVB_REQUIRE(out5, "cvb_struct_sizes: NULL output");
    out5[0] = (int64_t)sizeof(cvb_pars); out5[1] = (int64_t)sizeof(cvb_dist); out5[2] = (int64_t)sizeof(cvb_test_prob_pars);
    out5[3] = (int64_t)sizeof(cvb_trace_pars); out5[4] = (int64_t)sizeof(cvb_vaccinate_pars);
    return 0;
}

int cvb_create(cvb_sim** out, int64_t n_agents, int32_t n_variants, int32_t npts, uint64_t seed) {
    CVB_REQUIRE(out != nullptr, "cvb_create: out is NULL");
    CVB_REQUIRE(n_agents > 0, "cvb_create: n_agents must be positive");
    CVB_REQUIRE(n_variants >= 1 && n_variants <= CVB_MAX_VARIANTS, "cvb_create: n_variants must be in [1,%d]", CVB_MAX_VARIANTS);
    CVB_REQUIRE(npts >= 1, "cvb_create: npts must be positive");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("cvb_create: no CUDA device available (%s); covasim_b200 has no CPU fallback", cudaGetErrorString(e));
        return 100;
    }
    cvb_sim* s = new (std::nothrow) cvb_sim();
    CVB_REQUIRE(s != nullptr, "cvb_create: out of host memory");
    memset(s, 0, sizeof(*s));
    s->n = n_agents;
    s->nv = n_variants;
    s->npts = npts;
    s->seed = seed;
    cudaGetDevice(&s->device);
    s->quar_horizon = 1;
    int rc = 0;
    do {
        if ((rc = check_cuda(cudaMalloc((void**)&s->cand, (size_t)n_agents * sizeof(int32_t)), "cand"))) break;
        if ((rc = check_cuda(cudaMalloc((void**)&s->infect_key, (size_t)n_agents * sizeof(unsigned long long)), "infect_key"))) break;
        if ((rc = check_cuda(cudaMemset(s->infect_key, 0xFF, (size_t)n_agents * sizeof(unsigned long long)), "infect_key"))) break;
        if ((rc = check_cuda(cudaMalloc((void**)&s->n_cand, 64), "n_cand"))) break;
        if ((rc = check_cuda(cudaMemset(s->n_cand, 0, 64), "n_cand"))) break;
        if ((rc = check_cuda(cudaMalloc((void**)&s->beds, (size_t)npts * 2 * sizeof(unsigned long long)), "beds"))) break;
        if ((rc = check_cuda(cudaMemset(s->beds, 0, (size_t)npts * 2 * sizeof(unsigned long long)), "beds"))) break;
        if ((rc = check_cuda(cudaMalloc((void**)&s->edge_work, (size_t)npts * 2 * sizeof(unsigned long long)), "edge_work"))) break;
        if ((rc = check_cuda(cudaMemset(s->edge_work, 0, (size_t)npts * 2 * sizeof(unsigned long long)), "edge_work"))) break;
        if ((rc = check_cuda(cudaMalloc((void**)&s->quar_ring, (size_t)n_agents * sizeof(float)), "quar_ring"))) break;
        fill_f32_kernel<<<grid_for(n_agents), kThreads>>>(s->quar_ring, n_agents, -1.0f);
        if ((rc = check_cuda(cudaGetLastError(), "fill quar_ring"))) break;
        int64_t words = (n_agents + 31) / 32;
        if ((rc = check_cuda(cudaMalloc((void**)&s->case_bits, (size_t)words * sizeof(unsigned int)), "case_bits"))) break;
        if ((rc = check_cuda(cudaMemset(s->case_bits, 0, (size_t)words * sizeof(unsigned int)), "case_bits"))) break;
        if ((rc = check_cuda(cudaMalloc((void**)&s->inf_bits, (size_t)(words + 4) * sizeof(unsigned int)), "inf_bits"))) break;
        if ((rc = check_cuda(cudaMemset(s->inf_bits, 0, (size_t)(words + 4) * sizeof(unsigned int)), "inf_bits"))) break;
        if ((rc = check_cuda(cudaMalloc((void**)&s->trans_list, (size_t)n_agents * sizeof(int32_t)), "trans_list"))) break;
        if ((rc = check_cuda(cudaMalloc((void**)&s->case_list, (size_t)n_agents * sizeof(int32_t)), "case_list"))) break;
        if ((rc = check_cuda(cudaMalloc((void**)&s->n_trans, 64), "n_trans"))) break;
        if ((rc = check_cuda(cudaMemset(s->n_trans, 0, 64), "n_trans"))) break;
        if ((rc = check_cuda(cudaMalloc((void**)&s->n_case_list, 64), "n_case_list"))) break;
        if ((rc = check_cuda(cudaMemset(s->n_case_list, 0, 64), "n_case_list"))) break;
        if ((rc = check_cuda(cudaMalloc((void**)&s->n_cases, 64), "n_cases"))) break;
        if ((rc = check_cuda(cudaMemset(s->n_cases, 0, 64), "n_cases"))) break;
        if ((rc = check_cuda(cudaMalloc((void**)&s->dev_scalars, 64 * sizeof(unsigned long long)), "dev_scalars"))) break;
        if ((rc = check_cuda(cudaMemset(s->dev_scalars, 0, 64 * sizeof(unsigned long long)), "dev_scalars"))) break;
        if ((rc = check_cuda(cudaMallocHost((void**)&s->host_scalars, 64 * sizeof(unsigned long long)), "host_scalars"))) break;
        if ((rc = check_cuda(cudaDeviceSynchronize(), "cvb_create sync"))) break;
    } while (0);
    if (rc) { cvb_destroy(s); return rc; }
    *out = s;
    return 0;
}

int cvb_destroy(cvb_sim* s) {
    if (!s) return 0;
    cudaFree(s->cand); cudaFree(s->n_cand); cudaFree(s->infect_key); if (!s->beds_external) cudaFree(s->beds); cudaFree(s->edge_work); cudaFree(s->quar_ring); cudaFree(s->case_bits); cudaFree(s->inf_bits);
    cudaFree(s->trans_list); cudaFree(s->case_list); cudaFree(s->n_trans); cudaFree(s->n_case_list);
    cudaFree(s->n_cases); cudaFree(s->dev_scalars); cudaFree(s->rec_store); cudaFree(s->ts8_store);
    cudaFree(s->nab_kin); cudaFree(s->tile_cnt); cudaFree(s->hit_mask); cudaFree(s->flag_tmp); cudaFree(s->partial);
    cudaFree(s->glist); cudaFree(s->n_glist); cudaFree(s->hit_src); cudaFree(s->hit_key); cudaFree(s->part_flags);
    cudaFree(s->state); cudaFree(s->trans_ent); cudaFree(s->case_ent); cudaFree(s->stock_base);
    if (s->plan) { for (int k = 0; k < 4; ++k) free(s->plan->vacc_days[k]); }
    delete s->plan;
    cvb_timing_enable(s, 0);
    if (s->host_scalars) cudaFreeHost(s->host_scalars);
    delete s;
    return 0;
}

int cvb_reset(cvb_sim* s, cvb_stream st_) {
    CVB_REQUIRE(s, "cvb_reset: NULL handle");
    cudaStream_t st = (cudaStream_t)st_;
    s->state_valid = 0;
    CVB_CHECK(cudaMemsetAsync(s->infect_key, 0xFF, (size_t)s->n * sizeof(unsigned long long), st));
    CVB_CHECK(cudaMemsetAsync(s->n_cand, 0, 64, st));
    CVB_CHECK(cudaMemsetAsync(s->n_cases, 0, 64, st));
    CVB_CHECK(cudaMemsetAsync(s->beds, 0, (size_t)s->npts * 2 * sizeof(unsigned long long), st));
    CVB_CHECK(cudaMemsetAsync(s->edge_work, 0, (size_t)s->npts * 2 * sizeof(unsigned long long), st));
    if (s->part_flags) CVB_CHECK(cudaMemsetAsync(s->part_flags, 0, 64, st));
    fill_f32_kernel<<<grid_for((int64_t)s->quar_horizon * s->n), kThreads, 0, st>>>(s->quar_ring, (int64_t)s->quar_horizon * s->n, -1.0f);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_get_edge_work(cvb_sim* s, int64_t* host_out /* [npts][2] */) {
    CVB_REQUIRE(s && host_out, "cvb_get_edge_work: NULL argument");
    CVB_CHECK(cudaMemcpy(host_out, s->edge_work, (size_t)s->npts * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return 0;
}

int cvb_set_seed(cvb_sim* s, uint64_t seed) {
    CVB_REQUIRE(s, "cvb_set_seed: NULL handle");
    s->seed = seed;
    return 0;
}

int cvb_set_pars(cvb_sim* s, const cvb_pars* p) {
    CVB_REQUIRE(s && p, "cvb_set_pars: NULL argument");
    CVB_REQUIRE(p->n_variants == s->nv, "cvb_set_pars: n_variants %d does not match the handle's %d", p->n_variants, s->nv);
    CVB_REQUIRE(p->n_layers >= 0 && p->n_layers <= CVB_MAX_LAYERS, "cvb_set_pars: n_layers must be in [0,%d]", CVB_MAX_LAYERS);
    CVB_REQUIRE(p->n_vaccines >= 0 && p->n_vaccines <= CVB_MAX_VACCINES, "cvb_set_pars: n_vaccines must be in [0,%d]", CVB_MAX_VACCINES);
    s->pars = *p;
    s->pars_set = true;
    return 0;
}

int cvb_set_nab_kin(cvb_sim* s, const double* host_kin, int64_t n) {
    CVB_REQUIRE(s && host_kin && n > 0, "cvb_set_nab_kin: bad argument");
    if (s->nab_kin) cudaFree(s->nab_kin);
    s->nab_kin = nullptr;
    CVB_CHECK(cudaMalloc((void**)&s->nab_kin, (size_t)n * sizeof(double)));
    CVB_CHECK(cudaMemcpy(s->nab_kin, host_kin, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    s->nab_kin_len = n;
    return 0;
}

int cvb_bind_field(cvb_sim* s, int32_t field, void* ptr) {
    CVB_REQUIRE(s, "cvb_bind_field: NULL handle");
    CVB_REQUIRE(field >= 0 && field < CVB_N_FIELDS, "cvb_bind_field: field id %d out of range", field);
    CVB_REQUIRE(ptr != nullptr, "cvb_bind_field: NULL pointer for field %d", field);
    s->people.f[field] = ptr;
    return 0;
}

int cvb_bind_layer(cvb_sim* s, int32_t layer, int32_t* p1, int32_t* p2, float* beta, int64_t n_edges) {
    CVB_REQUIRE(s, "cvb_bind_layer: NULL handle");
    CVB_REQUIRE(layer >= 0 && layer < CVB_MAX_LAYERS, "cvb_bind_layer: layer %d out of range", layer);
    CVB_REQUIRE(n_edges >= 0, "cvb_bind_layer: negative edge count");
    CVB_REQUIRE(n_edges == 0 || (p1 && p2 && beta), "cvb_bind_layer: NULL edge array");
    CVB_REQUIRE((((uintptr_t)p1 | (uintptr_t)p2 | (uintptr_t)beta) & 15) == 0,
                "cvb_bind_layer: edge arrays must be 16-byte aligned (128-bit vector loads)");
    s->layers[layer].p1 = p1;
    s->layers[layer].p2 = p2;
    s->layers[layer].beta = beta;
    s->layers[layer].n_edges = n_edges;
    return 0;
}

int cvb_bind_adjacency(cvb_sim* s, const int64_t* adj_ptr, const void* adj, int64_t n_entries, uint32_t layer_mask) {
    CVB_REQUIRE(s, "cvb_bind_adjacency: NULL handle");
    if (layer_mask == 0 || n_entries == 0) {             // unbind: every layer goes through the dense passes
        s->adj_ptr = nullptr; s->adj = nullptr; s->adj_entries = 0; s->adj_layer_mask = 0;
        return 0;
    }
    CVB_REQUIRE(adj_ptr && adj, "cvb_bind_adjacency: NULL array");
    CVB_REQUIRE(((uintptr_t)adj & 15) == 0, "cvb_bind_adjacency: entries must be 16-byte aligned");
    s->adj_ptr = (const long long*)adj_ptr; s->adj = (const uint4*)adj; s->adj_entries = n_entries; s->adj_layer_mask = layer_mask;
    return 0;
}

int cvb_restore_compact(void* arena, int64_t arena_bytes, const int64_t* dev_table, int32_t n_seg, int64_t n_exc, cvb_stream st) {
    CVB_REQUIRE(arena && dev_table && n_seg >= 0 && n_exc >= 0 && ((uintptr_t)arena & 3) == 0 && arena_bytes % 4 == 0, "cvb_restore_compact: bad argument");
    (void)arena_bytes;                                                 // (the table was built for this arena by the caller: sizes are checked there)
    if (n_seg > 0) {
        dim3 grid(148, (unsigned)n_seg);
        cvb::fill_segments_kernel<<<grid, cvb::kThreads, 0, (cudaStream_t)st>>>((uint32_t*)arena, (const long long*)dev_table);
        CVB_LAUNCH_CHECK();
    }
    if (n_exc > 0) {
        const long long* idx = (const long long*)dev_table + 3 * (int64_t)n_seg;
        cvb::scatter_words_kernel<<<cvb::grid_for(n_exc, cvb::kThreads, 148 * 4), cvb::kThreads, 0, (cudaStream_t)st>>>((uint32_t*)arena, idx, idx + n_exc, n_exc);
        CVB_LAUNCH_CHECK();
    }
    return 0;
}

int cvb_peer_push(const void* src, int64_t n_bytes, const uint64_t* host_peer_ptrs, int32_t world, int64_t dst_offset_bytes, cvb_stream st) {
    CVB_REQUIRE(src && host_peer_ptrs && world >= 1 && world <= cvb::kMaxPeers && n_bytes >= 0 && n_bytes % 4 == 0 && dst_offset_bytes % 4 == 0,
                "cvb_peer_push: bad argument (at most %d peers, sizes in multiples of 4 bytes)", cvb::kMaxPeers);
    if (n_bytes == 0) return 0;
    cvb::PeerPtrs P;
    bool vec = ((uintptr_t)src & 15) == 0 && n_bytes % 16 == 0 && dst_offset_bytes % 16 == 0;
    for (int p = 0; p < world; ++p) {
        CVB_REQUIRE(host_peer_ptrs[p], "cvb_peer_push: NULL peer buffer");
        P.p[p] = (unsigned char*)(uintptr_t)host_peer_ptrs[p] + dst_offset_bytes;
        vec = vec && (host_peer_ptrs[p] & 15) == 0;
    }
    const int64_t items = vec ? n_bytes / 16 : n_bytes / 4;
    const int grid = cvb::grid_for(items, cvb::kThreads, 148 * 4);
    if (vec) cvb::peer_push_kernel<uint4><<<grid, cvb::kThreads, 0, (cudaStream_t)st>>>((const uint4*)src, items, P, world);
    else cvb::peer_push_kernel<uint32_t><<<grid, cvb::kThreads, 0, (cudaStream_t)st>>>((const uint32_t*)src, items, P, world);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_set_partition(cvb_sim* s, int64_t id0, int64_t n_global, int64_t chunk, int32_t world, const float* rel_trans_global,
                      uint8_t* codes_local, const uint8_t* codes_global, uint32_t* case_bits_local, const uint32_t* case_bits_global,
                      int64_t hit_capacity) {
    CVB_REQUIRE(s, "cvb_set_partition: NULL handle");
    CVB_REQUIRE(world >= 1 && chunk > 0 && chunk % 32 == 0, "cvb_set_partition: chunk must be a positive multiple of 32");
    CVB_REQUIRE(id0 >= 0 && id0 % chunk == 0 && id0 + s->n <= n_global && s->n <= chunk, "cvb_set_partition: agents [%lld, %lld) do not fit chunk %lld of %lld",
                (long long)id0, (long long)(id0 + s->n), (long long)chunk, (long long)n_global);
    CVB_REQUIRE(n_global <= (int64_t)world * chunk && (int64_t)world * chunk < (1ll << 31), "cvb_set_partition: world * chunk must cover n_global and fit int32");
    CVB_REQUIRE(s->nv <= 7, "cvb_set_partition: at most 7 variants fit the 1-byte transmit code");
    CVB_REQUIRE(rel_trans_global && codes_local && codes_global && case_bits_local && case_bits_global, "cvb_set_partition: NULL array");
    CVB_REQUIRE((((uintptr_t)codes_local | (uintptr_t)codes_global | (uintptr_t)case_bits_local | (uintptr_t)case_bits_global) & 15) == 0,
                "cvb_set_partition: arrays must be 16-byte aligned");
    s->id0 = id0; s->n_global = n_global; s->chunk = chunk; s->n_slots = (int64_t)world * chunk;
    s->rel_trans_global = rel_trans_global;
    s->codes_local = codes_local; s->codes_global = codes_global;
    s->case_bits_local = case_bits_local; s->case_bits_global = case_bits_global;
    if (hit_capacity <= 0) hit_capacity = s->n > 65536 ? s->n : 65536;   // default: one successful transmission per local agent per day
    if (hit_capacity < 1024) hit_capacity = 1024;
    cudaFree(s->cand); cudaFree(s->hit_src); cudaFree(s->hit_key); cudaFree(s->glist); cudaFree(s->n_glist); cudaFree(s->part_flags);
    s->cand = nullptr; s->hit_src = nullptr; s->hit_key = nullptr; s->glist = nullptr; s->n_glist = nullptr; s->part_flags = nullptr;
    // cand doubles as the unique-target list of cvb_infect_list (up to n_agents entries), so it never shrinks below n_agents
    CVB_CHECK(cudaMalloc((void**)&s->cand, (size_t)(hit_capacity > s->n ? hit_capacity : s->n) * sizeof(int32_t)));
    CVB_CHECK(cudaMalloc((void**)&s->hit_src, (size_t)hit_capacity * sizeof(int32_t)));
    CVB_CHECK(cudaMalloc((void**)&s->hit_key, (size_t)hit_capacity * sizeof(unsigned long long)));
    s->hit_cap = hit_capacity;
    s->glist_cap = s->n_slots;
    CVB_CHECK(cudaMalloc((void**)&s->glist, (size_t)s->glist_cap * sizeof(int32_t)));
    CVB_CHECK(cudaMalloc((void**)&s->n_glist, 64));
    CVB_CHECK(cudaMemset(s->n_glist, 0, 64));
    CVB_CHECK(cudaMalloc((void**)&s->part_flags, 64));
    CVB_CHECK(cudaMemset(s->part_flags, 0, 64));
    s->partitioned = 1;
    return 0;
}

int cvb_set_exchange_buffers(cvb_sim* s, const uint8_t* codes_global, const uint32_t* case_bits_global) {
    CVB_REQUIRE(s && s->partitioned, "cvb_set_exchange_buffers: call cvb_set_partition first");
    CVB_REQUIRE((((uintptr_t)codes_global | (uintptr_t)case_bits_global) & 15) == 0, "cvb_set_exchange_buffers: arrays must be 16-byte aligned");
    if (codes_global) s->codes_global = codes_global;
    if (case_bits_global) s->case_bits_global = case_bits_global;
    return 0;
}

int cvb_bind_partition_adjacency(cvb_sim* s, const int64_t* adj_ptr, const void* adj, int64_t n_entries, uint32_t layer_mask) {
    CVB_REQUIRE(s && s->partitioned, "cvb_bind_partition_adjacency: call cvb_set_partition first");
    CVB_REQUIRE(adj_ptr && layer_mask, "cvb_bind_partition_adjacency: NULL row pointers / empty layer mask");
    CVB_REQUIRE(n_entries == 0 || adj, "cvb_bind_partition_adjacency: NULL entries");
    CVB_REQUIRE(((uintptr_t)adj & 15) == 0, "cvb_bind_partition_adjacency: entries must be 16-byte aligned");
    s->padj_ptr = (const long long*)adj_ptr; s->padj = (const uint4*)adj; s->padj_entries = n_entries; s->padj_layer_mask = layer_mask;
    return 0;
}

int cvb_partition_status(cvb_sim* s, int64_t* host_out2) {
    CVB_REQUIRE(s && s->partitioned && host_out2, "cvb_partition_status: not a partitioned handle");
    unsigned int f[2] = {0, 0};
    CVB_CHECK(cudaMemcpy(f, s->part_flags, sizeof(f), cudaMemcpyDeviceToHost));
    host_out2[0] = f[0]; host_out2[1] = f[1];
    return 0;
}

int cvb_bind_results(cvb_sim* s, int64_t* counters, int64_t* vcounters, double* sums) {
    CVB_REQUIRE(s && counters && vcounters && sums, "cvb_bind_results: NULL argument");
    s->res.counters = (unsigned long long*)counters;
    s->res.vcounters = (unsigned long long*)vcounters;
    s->res.sums = sums;
    return 0;
}

int cvb_bind_beds(cvb_sim* s, int64_t* beds) {
    CVB_REQUIRE(s && beds, "cvb_bind_beds: NULL argument");
    if (!s->beds_external) cudaFree(s->beds);
    s->beds = (unsigned long long*)beds;
    s->beds_external = 1;
    return 0;
}

int cvb_bind_log(cvb_sim* s, int32_t* source, int32_t* target, int32_t* date, int8_t* layer, int8_t* variant,
                 int64_t cap, int64_t* count) {
    CVB_REQUIRE(s && source && target && date && layer && variant && count && cap > 0, "cvb_bind_log: bad argument");
    s->log.source = source; s->log.target = target; s->log.date = date; s->log.layer = layer; s->log.variant = variant;
    s->log.cap = cap; s->log.count = (unsigned long long*)count;
    return 0;
}

int cvb_clone_scratch(cvb_sim* dst, const cvb_sim* src, cvb_stream st_) {
    cudaStream_t st = (cudaStream_t)st_;
    CVB_REQUIRE(dst && src, "cvb_clone_scratch: NULL handle");
    CVB_REQUIRE(dst->n == src->n && dst->npts == src->npts && dst->nv == src->nv, "cvb_clone_scratch: the handles have different shapes");
    if (src->quar_horizon > dst->quar_horizon && cvb_set_quar_horizon(dst, src->quar_horizon)) return 1;
    CVB_REQUIRE(dst->quar_horizon == src->quar_horizon, "cvb_clone_scratch: the destination ring is larger than the source's");
    CVB_CHECK(cudaMemcpyAsync(dst->quar_ring, src->quar_ring, (size_t)src->quar_horizon * src->n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CVB_CHECK(cudaMemcpyAsync(dst->beds, src->beds, (size_t)src->npts * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    CVB_CHECK(cudaMemcpyAsync(dst->edge_work, src->edge_work, (size_t)src->npts * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    dst->last_t = src->last_t;
    dst->state_valid = 0;
    return 0;
}

int cvb_set_quar_horizon(cvb_sim* s, int32_t horizon) {
    CVB_REQUIRE(s && horizon >= 1 && horizon <= 64, "cvb_set_quar_horizon: horizon must be in [1,64]");
    if (horizon <= s->quar_horizon) return 0;
    // The ring is indexed by day % horizon.  Requests already pending for the days [today, today + old horizon) move to the slots
    // the same days have in the larger ring (people.py:620-640: a custom intervention may ask for a start date beyond the
    // current horizon in the middle of a run); "today" is the last day an entry point was called for.
    float* ring = nullptr;
    CVB_CHECK(cudaMalloc((void**)&ring, (size_t)horizon * s->n * sizeof(float)));
    fill_f32_kernel<<<grid_for(horizon * s->n), kThreads>>>(ring, horizon * s->n, -1.0f);
    CVB_LAUNCH_CHECK();
    CVB_CHECK(cudaDeviceSynchronize());
    for (int32_t d = s->last_t; d < s->last_t + s->quar_horizon; ++d)
        CVB_CHECK(cudaMemcpy(ring + (int64_t)(d % horizon) * s->n, s->quar_ring + (int64_t)(d % s->quar_horizon) * s->n,
                             (size_t)s->n * sizeof(float), cudaMemcpyDeviceToDevice));
    cudaFree(s->quar_ring);
    s->quar_ring = ring;
    s->quar_horizon = horizon;
    s->state_valid = 0;
    return 0;
}

}  // extern "C"
