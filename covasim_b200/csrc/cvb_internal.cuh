// Internal definitions shared by the .cu files of libcovasim_b200.so: the handle, the by-value
// kernel argument blocks, launch/error helpers and the block-level reduce / scan primitives.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "covasim_b200.h"
#include "cvb_device.cuh"

namespace cvb {

constexpr int kThreads = 256;                 // CTA size of every streaming kernel
constexpr int kEdgesPerThread = 4;            // 128-bit loads of p1 / p2 / beta
constexpr int kTileEdges = kThreads * kEdgesPerThread;
constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

// Device pointers of the per-agent arrays, passed to kernels by value (66 pointers = 528 B)
struct PeoplePtrs {
    void* f[CVB_N_FIELDS];
    template <typename T> __host__ __device__ __forceinline__ T* get(int id) const { return (T*)f[id]; }
};
#define PF(P, name)  ((P).template get<float>(CVB_F_##name))
#define PB(P, name)  ((P).template get<uint8_t>(CVB_F_##name))
#define PI(P, name)  ((P).template get<int32_t>(CVB_F_##name))

struct LayerPtrs {
    int32_t* p1;
    int32_t* p2;
    float* beta;
    int64_t n_edges;
};

struct LayerTable {                            // all layers, by value, for the fused edge pass
    LayerPtrs l[CVB_MAX_LAYERS];
    int64_t tile_start[CVB_MAX_LAYERS + 1];    // prefix sum of ceil(n_edges / tile size) over the table's entries
    int32_t layer_id[CVB_MAX_LAYERS];          // the layer each entry is (entries may skip layers)
    int32_t n_layers;                          // number of entries
};

struct ResultPtrs {
    unsigned long long* counters;              // [npts][CVB_N_COUNTERS]
    unsigned long long* vcounters;             // [npts][nv][CVB_N_VCOUNTERS]
    double* sums;                              // [npts][4]
};

// Built-in interventions registered with the handle so that cvb_run_days can run whole days without the host
struct DayPlan {
    int32_t has_test, test_start, test_end;         // test_prob (end < 0: none)
    cvb_test_prob_pars test;
    int32_t has_trace, trace_start, trace_end;      // contact_tracing
    cvb_trace_pars trace;
    uint32_t regen_mask;                            // dynamic layers regenerated every day (Layer.update, frac = 1)
    // vaccinate_prob interventions (interventions.py:1257-1662): per day bit 0 = first doses are offered, bit 1 = second doses fall due
    int32_t n_vacc;
    cvb_vaccinate_pars vacc[4];
    uint8_t* vacc_days[4];                          // host copies, [npts]
    int32_t* vacc_doses[4]; int32_t* vacc_due[4];   // the intervention's device arrays (cvb_vaccinate_prob)
};

struct FusedTiming;

struct LogPtrs {
    int32_t* source; int32_t* target; int32_t* date; int8_t* layer; int8_t* variant;
    int64_t cap; unsigned long long* count;
};

// Transmission records written by prepare_transmission and gathered by the edge pass: one 16-byte AgentRecord per agent
// (cvb_device.cuh); susceptibility against variants > 0 comes straight from the People array sus_imm
struct TransRecords {
    float4* rec;             // [N] AgentRecord {t, s, imm0, code}
    const float* sus_imm;    // People.sus_imm [n_variants][N]
    float2* ts8;             // [n_layers][N] {rel_trans, rel_sus} per layer, written only for the layers the dense streaming
                             // pass will read (dynamic layers / no adjacency) when n_variants == 1; NULL otherwise
    uint32_t ts8_mask;       // the layers ts8 holds today
};

}  // namespace cvb

struct cvb_sim {
    int64_t n;
    int32_t nv, npts, device;
    uint64_t seed;
    cvb_pars pars;
    bool pars_set;
    cvb::PeoplePtrs people;
    cvb::LayerPtrs layers[CVB_MAX_LAYERS];
    cvb::ResultPtrs res;
    cvb::LogPtrs log;
    // scratch owned by the library
    cvb::TransRecords rec; int32_t rec_layers;
    float2* ts8_store; int32_t ts8_layers;          // backing store of rec.ts8
    float4* rec_store;                              // backing store of rec.rec
    int32_t* cand; unsigned int* n_cand;            // today's newly infected candidates
    unsigned long long* infect_key;                 // [N] winning transmission key per target (kEmptyKey = none)
    unsigned long long* beds;                       // [npts][2] severe / critical after update_states_pre
    int32_t beds_external;                          // bound by the caller (cvb_bind_beds): agent-partitioned runs all-reduce the day's row
    unsigned long long* edge_work;                  // [npts][2] adjacency entries visited / transmitters, per day (sparse edge pass)
    double* nab_kin; int64_t nab_kin_len;           // NAb kinetics table (immunity.py:298)
    float* quar_ring; int32_t quar_horizon;         // [quar_horizon][N] pending quarantine end days, -1 = none
    int32_t last_t;                                 // the last day update_states_pre ran for ("today" when the ring is re-sized)
    unsigned int* case_bits; unsigned int* n_cases; // contact tracing: bitmap of today's cases
    unsigned int* inf_bits;                         // [ceil(N/32)] agents that can transmit today (written by prepare_transmission)
    int32_t* trans_list; unsigned int* n_trans;     // the same set as a compact (unordered) list
    int32_t* case_list; unsigned int* n_case_list;  // contact tracing: today's cases as a compact list
    // bidirectional adjacency (CSR over agents) of the static layers, bound by the host (cvb_bind_adjacency)
    const long long* adj_ptr; const uint4* adj; int64_t adj_entries; uint32_t adj_layer_mask;
    // agent partition of one large simulation over several GPUs (cvb_set_partition): this handle owns the agents
    // [id0, id0 + n) of n_global.  Per-agent Philox keys and logged ids use GLOBAL ids, so a partitioned run is
    // bit-identical to the single-GPU run of the same simulation.
    int64_t id0, n_global, chunk; int32_t partitioned;
    const float* rel_trans_global;                  // [n_global] transmissibility at initialisation (replicated)
    uint8_t* codes_local; const uint8_t* codes_global;   // [chunk] written by prepare_transmission / [world*chunk] all-gathered by the host
    unsigned int* case_bits_local; const unsigned int* case_bits_global;   // [chunk/32] / [world*chunk/32]
    int64_t n_slots;                                // world * chunk: number of global code slots
    const long long* padj_ptr; const uint4* padj; int64_t padj_entries; uint32_t padj_layer_mask;   // rows: GLOBAL source slot; entries: LOCAL target
    int32_t* glist; unsigned int* n_glist; int64_t glist_cap;   // compact list of global transmitters / cases
    int32_t* hit_src; unsigned long long* hit_key; int64_t hit_cap;   // per successful transmission (parallel to cand)
    unsigned int* part_flags;                       // [0] hits dropped (capacity), [1] agents whose rel_trans the code cannot express
    // scan / compaction workspace
    unsigned int* tile_cnt; int64_t tile_cnt_cap;   // per-tile counts (and their exclusive scan, in place)
    uint8_t* hit_mask; int64_t hit_mask_cap;
    uint8_t* flag_tmp; int64_t flag_tmp_cap;
    double* partial; int64_t partial_cap;           // per-block partial sums
    unsigned long long* dev_scalars;                // small device scratch for counts returned to the host
    unsigned long long* host_scalars;               // pinned mirror
    // ---- fused day pipeline (day_fused.cu) ----------------------------------------------------------------------------
    // state[i]: packed copy of the 16 bool states + a few "is there anything to read" bits, kept in step with the public
    // People arrays by the fused kernels; rebuilt from the arrays (pack_state_kernel) whenever anything else may have written them
    uint32_t* state; int32_t state_valid;
    uint4* trans_ent;                               // [N][2] today's transmitters {agent, row length, row begin, rel_trans, code}
    uint4* case_ent;                                // [N] today's traced cases {agent, row length, row begin}
    unsigned long long* stock_base;                 // absolute stock counts of the packed words after a pack (counter-row layout)
    int32_t block_packed;                           // the state words were rebuilt at the start of the current block of days
    int32_t begin_grid;                             // grid of the last day_begin_kernel launch (= number of per-CTA partial sums)
    int32_t tune[8];                                // launch-shape overrides (cvb_tune): 0/1 day_begin CTA size / chunk, 2/3 day_mid, 4/5 edge pass lanes / unroll
    cvb::DayPlan* plan;                             // built-in interventions the C day loop runs itself (cvb_plan_*)
    cvb::FusedTiming* timing;                       // per-kernel CUDA-event timing of the day loop, when enabled
};

namespace cvb {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int ensure_u32(unsigned int** p, int64_t* cap, int64_t need);
int ensure_u8(uint8_t** p, int64_t* cap, int64_t need);
int ensure_f64(double** p, int64_t* cap, int64_t need);
int exclusive_scan_u32(unsigned int* data, int64_t n, unsigned long long* total_out, cudaStream_t st);
// agent-partitioned runs: set bits of a gathered bitmap -> s->glist / s->n_glist (edge_pass.cu)
int list_from_bits(cvb_sim* s, const unsigned int* bits, int64_t n_words, cudaStream_t st);
int edge_pass_impl(cvb_sim* s, int32_t t, cudaStream_t st, bool from_entries);       // edge_pass.cu
int launch_trace_sparse2(cvb_sim* s, int32_t t, const cvb_trace_pars* tr, cudaStream_t st);   // interventions.cu
int launch_infect_winners(cvb_sim* s, int32_t t, bool with_state, cudaStream_t st);   // infect.cu
int launch_trace_partition(cvb_sim* s, int32_t t, const cvb_trace_pars* tr, cudaStream_t st);   // interventions.cu
int launch_vaccinate_fused(cvb_sim* s, int32_t t, const cvb_vaccinate_pars* vp, int32_t* iv_doses, int32_t* due_day, cudaStream_t st);   // interventions.cu
// every entry point that writes People flags outside the fused pipeline calls this: the packed state must be rebuilt
inline void state_touched(cvb_sim* s) { s->state_valid = 0; }

#define CVB_CHECK(call) do { int rc_ = cvb::check_cuda((call), #call); if (rc_) return rc_; } while (0)
#define CVB_REQUIRE(cond, ...) do { if (!(cond)) { cvb::set_error(__VA_ARGS__); return 1; } } while (0)
extern unsigned long long g_launches;        // kernels launched by this library (every launch site uses CVB_LAUNCH_CHECK)
#define CVB_LAUNCH_CHECK() do { __atomic_fetch_add(&cvb::g_launches, 1ull, __ATOMIC_RELAXED); CVB_CHECK(cudaGetLastError()); } while (0)     // (atomic: cvb_run_days_multi launches from several host threads)

inline int grid_for(int64_t n, int per_block = kThreads, int64_t cap = 148 * 16) {
    int64_t g = (n + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

#if defined(__CUDACC__)
// ---- programmatic dependent launch (the fused day pipeline is a chain of five dependent kernels per day) -----------------------
// A kernel launched with launch_pdl may be scheduled while its predecessor in the stream is still draining: its CTAs run their
// prologue (parameters, shared-memory setup) and then block in pdl_wait() until the predecessor has completed and its writes are
// visible.  pdl_trigger() tells the scheduler that the successor may be brought in; every kernel of the chain calls it first.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---- warp / block primitives ---------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// Sum `v` over the CTA and add the result to *dst with ONE global atomic (Guideline 12)
__device__ __forceinline__ void block_add(unsigned long long* dst, int v, int* smem_slot) {
    int w = __reduce_add_sync(0xFFFFFFFFu, v);
    if (lane_id() == 0 && w) atomicAdd(smem_slot, w);
}

// Exclusive prefix sum of one int per thread across a CTA of kThreads threads; returns the thread's
// offset and, through `total`, the CTA sum.  `warp_sums` is shared int[kThreads/32].
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int& total) {
    int lane = lane_id(), w = warp_id();
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    int nw = blockDim.x >> 5;
    if (w == 0) {
        int s = lane < nw ? warp_sums[lane] : 0;
        int si = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int o = __shfl_up_sync(0xFFFFFFFFu, si, d);
            if (lane >= d) si += o;
        }
        if (lane < nw) warp_sums[lane] = si - s;          // exclusive warp offsets
        if (lane == nw - 1) warp_sums[nw] = si;           // total
    }
    __syncthreads();
    total = warp_sums[nw];
    int out = warp_sums[w] + incl - v;
    __syncthreads();
    return out;
}

// Warp-aggregated "append one item": returns this lane's slot in a global list (Guideline 12)
__device__ __forceinline__ unsigned long long warp_append(unsigned long long* counter) {
    unsigned mask = __activemask();
    int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane_id() == leader) base = atomicAdd(counter, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane_id()) - 1u));
}
__device__ __forceinline__ unsigned int warp_append32(unsigned int* counter) {
    unsigned mask = __activemask();
    int leader = __ffs(mask) - 1;
    unsigned int base = 0;
    if (lane_id() == leader) base = atomicAdd(counter, (unsigned int)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane_id()) - 1u));
}

// ---- four agents per thread: vector loads of the structure-of-arrays fields -------------------------
constexpr int kAPT = 4;                                    // agents per thread

__device__ __forceinline__ void load4(const float* __restrict__ p, int64_t i0, int64_t n, bool vec, float fill, float o[4]) {
    if (vec && i0 + 4 <= n) {
        float4 v = *reinterpret_cast<const float4*>(p + i0);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = (i0 + k < n) ? p[i0 + k] : fill;
    }
}
__device__ __forceinline__ void load4(const int32_t* __restrict__ p, int64_t i0, int64_t n, bool vec, int32_t o[4]) {
    if (vec && i0 + 4 <= n) {
        int4 v = *reinterpret_cast<const int4*>(p + i0);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = (i0 + k < n) ? p[i0 + k] : 0;
    }
}
// four bool bytes as one 32-bit word (byte k = agent i0+k)
__device__ __forceinline__ uint32_t load4b(const uint8_t* __restrict__ p, int64_t i0, int64_t n, bool vec) {
    if (vec && i0 + 4 <= n) return *reinterpret_cast<const uint32_t*>(p + i0);
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (i0 + k < n) w |= (uint32_t)p[i0 + k] << (8 * k);
    return w;
}
__device__ __forceinline__ bool flag(uint32_t w, int k) { return ((w >> (8 * k)) & 0xFFu) != 0; }
__device__ __forceinline__ int count4(uint32_t w) { return __popc(__vcmpne4(w, 0u) & 0x01010101u); }

// Per-thread event counters -> CTA totals in shared memory.  Most counters are flows of rare events (a death, a new
// diagnosis): a warp whose 32 lanes all hold 0 skips the reduction after one vote.
template <int NK>
__device__ __forceinline__ void reduce_counters(const int (&c)[NK], int* s_cnt) {
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        if (!__any_sync(0xFFFFFFFFu, c[k] != 0)) continue;
        int w = __reduce_add_sync(0xFFFFFFFFu, c[k]);
        if (lane_id() == 0) atomicAdd(&s_cnt[k], w);
    }
}


// ---- the packed state word ----------------------------------------------------------------------------------------------------
// bits 0-15: the bool states in the order of defaults.states (== CVB_F_susceptible + bit)
enum : uint32_t {
    SB_SUS = 1u << 0, SB_NAIVE = 1u << 1, SB_EXP = 1u << 2, SB_INF = 1u << 3, SB_SYMP = 1u << 4, SB_SEV = 1u << 5, SB_CRIT = 1u << 6,
    SB_TESTED = 1u << 7, SB_DIAG = 1u << 8, SB_REC = 1u << 9, SB_KDEAD = 1u << 10, SB_DEAD = 1u << 11, SB_KCONTACT = 1u << 12,
    SB_QUAR = 1u << 13, SB_ISO = 1u << 14, SB_VACC = 1u << 15,
    SB_HAS_NAB = 1u << 16,      // peak_nab != 0 (update_nab / check_immunity have work to do)
    SB_IMM_NZ = 1u << 17,       // some sus_imm / symp_imm / sev_imm entry of the agent may be non-zero
    SB_QPEND = 1u << 18,        // a quarantine request is waiting in the pending ring
    SB_DPEND = 1u << 19,        // date_diagnosed is set and the agent is not (yet) diagnosed
    SB_RS_VALID = 1u << 20,     // the stored agent record has the simple form {0, rel_sus or 0, 0, quarantine bit} ...
    SB_RS_SUS = 1u << 21,       // ... written with this susceptible flag
    SB_RS_QUAR = 1u << 22,      // ... and this quarantined flag
    SB_IBV = 1u << 23,          // infectious_by_variant[EBV - 1] is set
};
constexpr int kEbvShift = 24;   // bits 24-27: variant + 1 of the set exposed_by_variant row (0: none)
constexpr int kRvShift = 28;    // bits 28-31: recovered_variant + 1 while t >= date_recovered (the natural-immunity source), else 0
__device__ __forceinline__ int sb_ebv(uint32_t s) { return (int)((s >> kEbvShift) & 15u); }
__device__ __forceinline__ int sb_rv(uint32_t s) { return (int)((s >> kRvShift) & 15u); }
constexpr uint32_t kEbvMask = 15u << kEbvShift, kRvMask = 15u << kRvShift;

// ---- stock counters as running totals ---------------------------------------------------------------------------------------
// The unfused path counts every state flag of every agent at the end of every day (sim.py:652-664).  Here the stock columns of the
// day's counter row start as a copy of the previous day's (added by day_begin_kernel) and every kernel that changes a state word
// adds the difference: a handful of shared-memory atomics for the few agents whose state changed, instead of 13 counts over all.
constexpr uint32_t kStockMask = SB_SUS | SB_EXP | SB_INF | SB_SYMP | SB_SEV | SB_CRIT | SB_DIAG | SB_REC | SB_KDEAD | SB_DEAD | SB_QUAR | SB_ISO | SB_VACC;
constexpr int kStockSlots = 16 + 1 + 2 * CVB_MAX_VARIANTS;      // 16 state bits, alive, exposed / infectious by variant

__device__ __forceinline__ void stock_delta(uint32_t o, uint32_t s, int* __restrict__ s_delta) {
    const uint32_t ch = (o ^ s) & (kStockMask | kEbvMask | SB_IBV);
    if (!ch) return;
    uint32_t fb = ch & kStockMask;
    while (fb) {
        const int b = __ffs(fb) - 1;
        fb &= fb - 1;
        atomicAdd(&s_delta[b], ((s >> b) & 1u) ? 1 : -1);
    }
    if (ch & SB_DEAD) atomicAdd(&s_delta[16], (s & SB_DEAD) ? -1 : 1);
    if (ch & (kEbvMask | SB_IBV)) {
        const int eo = sb_ebv(o), en = sb_ebv(s);
        if (eo) { atomicAdd(&s_delta[17 + 2 * (eo - 1)], -1); if (o & SB_IBV) atomicAdd(&s_delta[18 + 2 * (eo - 1)], -1); }
        if (en) { atomicAdd(&s_delta[17 + 2 * (en - 1)], 1); if (s & SB_IBV) atomicAdd(&s_delta[18 + 2 * (en - 1)], 1); }
    }
}
// where stock slot k lives in a day's counter rows (NULL: not a result stock)
__device__ __forceinline__ unsigned long long* stock_slot(unsigned long long* row, unsigned long long* vrow, int nv, int k) {
    // state bit -> stock counter (defaults.result_stocks order: susceptible, exposed, infectious, symptomatic, severe, critical,
    // recovered, dead, diagnosed, known_dead, quarantined, isolated, vaccinated)
    // {0, -, 1, 2, 3, 4, 5, -, 8, 6, 9, 7, -, 10, 11, 12} as sixteen nibbles (15 = not a stock), bit 0 in the lowest
    if (k < 16) {
        const int st = (int)((0xCBAF7968F54321F0ull >> (4 * k)) & 15ull);
        return st != 15 ? row + CVB_C_n_susceptible + st : nullptr;
    }
    if (k == 16) return row + CVB_C_n_alive_agents;
    const int q = k - 17, var = q >> 1;
    return var < nv ? vrow + (int64_t)var * CVB_N_VCOUNTERS + ((q & 1) ? CVB_VC_n_infectious_by_variant : CVB_VC_n_exposed_by_variant) : nullptr;
}
__device__ __forceinline__ void flush_stock_delta(const int* s_delta, unsigned long long* row, unsigned long long* vrow, int nv) {
    if (threadIdx.x < kStockSlots && s_delta[threadIdx.x]) {
        unsigned long long* p = stock_slot(row, vrow, nv, threadIdx.x);
        if (p) atomicAdd(p, (unsigned long long)(long long)s_delta[threadIdx.x]);      // two's complement: negative differences wrap correctly
    }
}

// 128-bit streaming loads that do not allocate in L1 (edge arrays are read exactly once per pass)
__device__ __forceinline__ int4 ld_stream(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
#endif

}  // namespace cvb
