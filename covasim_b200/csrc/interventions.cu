// Built-in interventions as device passes (native-RNG mode) and dynamic-layer regeneration:
//   test_prob + People.test            reference interventions.py:857-981, people.py:589-617
//   test_num                           reference interventions.py:718-854 (weights + exponential-clock keys; tests for a list)
//   contact_tracing                    reference interventions.py:984-1145, base.py:1808-1846, utils.py:131-147
//   vaccinate_prob + vaccinate         reference interventions.py:1257-1662, immunity.py:138-202
//   Layer.update (frac = 1)            reference base.py:1849-1876
// All Bernoulli draws are keyed per agent (or per agent x layer for tracing), so results are
// independent of thread order; see oracle/cvoracle.py for the CPU restatement they are tested against.
#include <string.h>
#include "cvb_internal.cuh"

namespace cvb {

// ================================================================================================
// test_prob
// ================================================================================================
__global__ void __launch_bounds__(kThreads, 3) test_prob_kernel(PeoplePtrs P, const __grid_constant__ cvb_test_prob_pars tp, uint64_t seed,
        int64_t n, int64_t id0, int32_t t, bool vec, unsigned long long* __restrict__ counters, const double* __restrict__ prob_override,
        const double* __restrict__ tape /* verification: the three uniforms of every agent given instead of keyed, [n][3]; or NULL */) {
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int c = 0;
    const float tf = (float)t;
    const float qnan = nanf32();
    const uint8_t* symptomatic = PB(P, symptomatic); const uint8_t* diagnosed = PB(P, diagnosed); const uint8_t* quarantined = PB(P, quarantined);
    const uint8_t* infectious = PB(P, infectious); uint8_t* tested = PB(P, tested);
    const float* d_quar = PF(P, date_quarantined); const float* d_end_quar = PF(P, date_end_quarantine);
    float* d_tested = PF(P, date_tested); float* d_diag = PF(P, date_diagnosed); float* d_pos = PF(P, date_pos_test);
    const int policy = tp.quar_policy;
    const int64_t n_groups = (n + kAPT - 1) / kAPT;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_groups; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = g * kAPT;
        const uint32_t w_symp = load4b(symptomatic, i0, n, vec), w_diag = load4b(diagnosed, i0, n, vec), w_inf = load4b(infectious, i0, n, vec);
        uint32_t w_quar = 0;
        float dq[4] = {qnan, qnan, qnan, qnan}, deq[4] = {qnan, qnan, qnan, qnan}, ddiag[4];
        if (policy == 0 || policy == 2) load4(d_quar, i0, n, vec, qnan, dq);
        if (policy == 1 || policy == 2) load4(d_end_quar, i0, n, vec, qnan, deq);
        if (policy == 3) w_quar = load4b(quarantined, i0, n, vec);
        load4(d_diag, i0, n, vec, qnan, ddiag);
#pragma unroll
        for (int k = 0; k < kAPT; ++k) {
            const int64_t i = i0 + k;
            if (i >= n) break;
            if (flag(w_diag, k)) continue;                             // diagnosed people do not test (interventions.py:973)
            const bool symp = flag(w_symp, k);
            bool qt;                                                   // interventions.py:691-715 get_quar_inds
            switch (policy) {
                case 0:  qt = dq[k] == tf - 1.0f; break;
                case 1:  qt = deq[k] == tf + 1.0f; break;
                case 2:  qt = (dq[k] == tf - 1.0f) || (deq[k] == tf + 1.0f); break;
                default: qt = flag(w_quar, k); break;
            }
            double prob = qt ? (symp ? tp.symp_quar_prob : tp.asymp_quar_prob) : (symp ? tp.symp_prob : tp.asymp_prob);
            if (prob_override) {                                       // subtarget (interventions.py:971-973): explicit probabilities win
                const double ov = prob_override[i];
                if (ov == ov) prob = ov;
            }
            if (!(prob > 0.0)) continue;
            if (!((tape ? tape[i * 3] : keyed_uniform(seed, P_TEST, (uint32_t)tp.index, t, i + id0, 0)) < prob)) continue;
            // People.test (people.py:589-617)
            ++c;
            tested[i] = 1;
            d_tested[i] = tf;
            if (!flag(w_inf, k)) continue;
            if (!((tape ? tape[i * 3 + 1] : keyed_uniform(seed, P_TEST_SENS, (uint32_t)tp.index, t, i + id0, 0)) < tp.sensitivity)) continue;
            if (!is_nan(ddiag[k])) continue;
            if (!((tape ? tape[i * 3 + 2] : keyed_uniform(seed, P_TEST_LOSS, (uint32_t)tp.index, t, i + id0, 0)) < 1.0 - tp.loss_prob)) continue;
            d_diag[i] = (float)(t + tp.test_delay);
            d_pos[i] = tf;
        }
    }
    int w = __reduce_add_sync(0xFFFFFFFFu, c);
    if (lane_id() == 0 && w) atomicAdd(&s_cnt, w);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(counters + (int64_t)t * CVB_N_COUNTERS + CVB_C_new_tests, (unsigned long long)s_cnt);
}

// ================================================================================================
// test_num (reference interventions.py:718-854): a fixed number of tests per day, handed out by weight
// ================================================================================================
// Weighted sampling without replacement of n_tests agents is the n_tests smallest of the keys  -log(1 - u_i) / w_i  (exponential
// clocks; Efraimidis & Spirakis 2006) with u_i the agent's keyed uniform: this kernel writes the weights (the reference's
// test_probs) and the keys, the host picks the n smallest (torch.topk) and cvb_test_list administers the tests.
__global__ void __launch_bounds__(kThreads) test_num_keys_kernel(PeoplePtrs P, const __grid_constant__ cvb_test_num_pars tp, uint64_t seed, int64_t n,
        int64_t id0, int32_t t, double* __restrict__ weight, double* __restrict__ key) {
    const float tf = (float)t;
    const uint8_t* symptomatic = PB(P, symptomatic); const uint8_t* diagnosed = PB(P, diagnosed); const uint8_t* quarantined = PB(P, quarantined);
    const float* d_quar = PF(P, date_quarantined); const float* d_end_quar = PF(P, date_end_quarantine);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double w = 1.0;
        if (symptomatic[i]) w = dmul(w, tp.symp_test);
        bool qt;                                                       // interventions.py:691-715 get_quar_inds
        switch (tp.quar_policy) {
            case 0:  qt = d_quar[i] == tf - 1.0f; break;
            case 1:  qt = d_end_quar[i] == tf + 1.0f; break;
            case 2:  qt = (d_quar[i] == tf - 1.0f) || (d_end_quar[i] == tf + 1.0f); break;
            case 3:  qt = quarantined[i] != 0; break;
            default: qt = false; break;                                // 4: the caller applies its own quarantine-testing set
        }
        if (qt) w = dmul(w, tp.quar_test);
        if (diagnosed[i]) w = 0.0;                                     // diagnosed people do not test
        weight[i] = w;
        const double u = keyed_uniform(seed, P_TEST, (uint32_t)tp.index, t, i + id0, 0);
        key[i] = w > 0.0 ? -log(1.0 - u) / w : __longlong_as_double(0x7ff0000000000000ll);
    }
}

// People.test (people.py:589-617) for an explicit list of distinct agents
__global__ void __launch_bounds__(kThreads) test_list_kernel(PeoplePtrs P, const int32_t* __restrict__ inds, int64_t n_inds, uint64_t seed, int64_t n,
        int64_t id0, int32_t t, double sensitivity, double loss_prob, int32_t test_delay, uint32_t index) {
    const float tf = (float)t;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_inds; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = inds[j];
        if (i < 0 || i >= n) continue;
        PB(P, tested)[i] = 1;
        PF(P, date_tested)[i] = tf;
        if (!PB(P, infectious)[i]) continue;
        if (!(keyed_uniform(seed, P_TEST_SENS, index, t, i + id0, 0) < sensitivity)) continue;
        if (!is_nan(PF(P, date_diagnosed)[i])) continue;
        if (!(keyed_uniform(seed, P_TEST_LOSS, index, t, i + id0, 0) < 1.0 - loss_prob)) continue;
        PF(P, date_diagnosed)[i] = (float)(t + test_delay);
        PF(P, date_pos_test)[i] = tf;
    }
}

// ================================================================================================
// contact_tracing
// ================================================================================================
// Today's cases as a bitmap: each thread tests 4 agents (one 128-bit load), 8 lanes assemble a 32-bit word
__global__ void __launch_bounds__(kThreads) trace_select_kernel(PeoplePtrs P, int64_t n, int32_t t, int presumptive, bool vec,
        unsigned int* __restrict__ case_bits, int32_t* __restrict__ case_list, unsigned int* __restrict__ n_case_list) {
    const float tf = (float)t;
    const float* d_diag = PF(P, date_diagnosed); const float* d_tested = PF(P, date_tested); const uint8_t* exposed = PB(P, exposed);
    const int64_t n_groups = (n + 3) / 4;
    const int64_t n_groups_pad = (n_groups + 31) / 32 * 32;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_groups_pad; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = g * 4;
        float d[4];
        unsigned nib = 0;
        if (!presumptive) {
            if (vec && i0 + 4 <= n) { float4 v = *reinterpret_cast<const float4*>(d_diag + i0); d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
            else { for (int k = 0; k < 4; ++k) d[k] = (i0 + k < n) ? d_diag[i0 + k] : -1.0f; }
#pragma unroll
            for (int k = 0; k < 4; ++k) nib |= (unsigned)(i0 + k < n && d[k] == tf) << k;
        } else {
            for (int k = 0; k < 4; ++k) nib |= (unsigned)(i0 + k < n && d_tested[i0 + k] == tf && exposed[i0 + k] != 0) << k;
        }
        unsigned word = nib << (4 * (lane_id() & 7));
        word |= __shfl_xor_sync(0xFFFFFFFFu, word, 1);
        word |= __shfl_xor_sync(0xFFFFFFFFu, word, 2);
        word |= __shfl_xor_sync(0xFFFFFFFFu, word, 4);
        const int64_t widx = (i0 - (int64_t)(lane_id() & 7) * 4) / 32;
        if ((lane_id() & 7) == 0 && widx * 32 < n) case_bits[widx] = word;
        if (nib) {                                                  // and as a compact list for the adjacency form
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (nib & (1u << k)) case_list[atomicAdd(n_case_list, 1u)] = (int32_t)(i0 + k);
        }
    }
}

// an explicit case list (contact_tracing with a capacity: the host picked who is traced) -> case bitmap + case list
__global__ void __launch_bounds__(kThreads) set_cases_kernel(const int32_t* __restrict__ inds, int64_t n_inds, int64_t n,
        unsigned int* __restrict__ case_bits, int32_t* __restrict__ case_list, unsigned int* __restrict__ n_case_list) {
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_inds; j += (int64_t)gridDim.x * blockDim.x) {
        const int i = inds[j];
        if (i < 0 || i >= n) continue;
        atomicOr(case_bits + (i >> 5), 1u << (i & 31));
        case_list[j] = i;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_case_list = (unsigned int)n_inds;
}

struct TraceTable {                      // the traced layers of one contact_tracing intervention, by value
    const int32_t* p1[CVB_MAX_LAYERS];
    const int32_t* p2[CVB_MAX_LAYERS];
    int* quar_slot[CVB_MAX_LAYERS];      // ring slot of the day the contact is notified (t + trace_time)
    double trace_prob[CVB_MAX_LAYERS];
    int64_t n_edges[CVB_MAX_LAYERS];
    int64_t tile_start[CVB_MAX_LAYERS + 1];
    float notify_day[CVB_MAX_LAYERS];
    int32_t layer_id[CVB_MAX_LAYERS];
    int32_t entry_of_layer[CVB_MAX_LAYERS];   // inverse of layer_id (-1: layer not traced)
    int32_t n_entries;
    float end_day;                       // t + quar_period (start + (quar_period - trace_time), interventions.py:1144)
    uint64_t seed;
    int64_t n, n_words;
    int64_t id0;                         // global id of local agent 0 (agent-partitioned runs): Philox keys use global ids
    int32_t t, index;
    const double* tape;                  // verification (cvb_contact_tracing_taped): the uniform of (layer, contact) is tape[layer * n + contact]
};

__device__ __forceinline__ void trace_notify(const PeoplePtrs& P, const TraceTable& T, int q, int c) {
    const uint32_t sub = ((uint32_t)T.index << 8) | (uint32_t)T.layer_id[q];
    const double u = T.tape ? T.tape[(int64_t)T.layer_id[q] * T.n + c] : keyed_uniform(T.seed, P_TRACE, sub, T.t, (int64_t)c + T.id0, 0);
    if (!(u < T.trace_prob[q])) return;                                 // binomial_filter, interventions.py:1116
    if (PB(P, dead)[c]) return;                                        // interventions.py:1139-1141
    PB(P, known_contact)[c] = 1;
    // date_known_contact = fmin(old, notify_day): for non-negative floats and NaN the unsigned bit
    // patterns order the same way (NaN = 0x7fc00000 is the largest), so atomicMin does fmin
    atomicMin((unsigned int*)PF(P, date_known_contact) + c, (unsigned int)__float_as_int(T.notify_day[q]));
    atomicMax(T.quar_slot[q] + c, __float_as_int(T.end_day));          // people.py:620-640 schedule_quarantine
}

// Fused day pipeline: the cases come as entries {agent, row length, row begin} written by day_begin_kernel, the contact's "dead" flag
// is read from the packed state word, and the word's known_contact / pending-request bits are set with the People arrays
__device__ __forceinline__ void trace_notify2(const PeoplePtrs& P, uint32_t* __restrict__ S, const TraceTable& T, int q, int c) {
    const uint32_t sub = ((uint32_t)T.index << 8) | (uint32_t)T.layer_id[q];
    const double u = T.tape ? T.tape[(int64_t)T.layer_id[q] * T.n + c] : keyed_uniform(T.seed, P_TRACE, sub, T.t, (int64_t)c + T.id0, 0);
    if (!(u < T.trace_prob[q])) return;                                 // binomial_filter, interventions.py:1116
    if (S[c] & (1u << 11)) return;                                     // dead (interventions.py:1139-1141)
    PB(P, known_contact)[c] = 1;
    atomicOr(S + c, (1u << 12) | (1u << 18));                          // known_contact, request pending
    atomicMin((unsigned int*)PF(P, date_known_contact) + c, (unsigned int)__float_as_int(T.notify_day[q]));
    atomicMax(T.quar_slot[q] + c, __float_as_int(T.end_day));          // people.py:620-640 schedule_quarantine
}

// Adjacency form: G lanes per case walk the case's own edges (static layers, both directions).  A group's chain per case is list index ->
// row pointers -> entries; the index two ahead and the row pointers one ahead are loaded while a row is processed.  G follows the mean
// row length: 32 for a whole population's rows (~36 entries), down to 4 for the rows of an 8-way agent partition (~4.5 local entries).
template <int G>
__global__ void __launch_bounds__(kThreads) trace_sparse_kernel(PeoplePtrs P, const __grid_constant__ TraceTable T,
        const long long* __restrict__ adj_ptr, const uint4* __restrict__ adj, const int32_t* __restrict__ case_list,
        const unsigned int* __restrict__ n_case_ptr, uint32_t layer_mask, uint32_t* __restrict__ S /* fused day: packed state words, or NULL */) {
    const unsigned int n_cases = *n_case_ptr;
    const int gl = threadIdx.x & (G - 1);
    const unsigned int stride = (gridDim.x * blockDim.x) / G;
    unsigned int ci = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    int i0 = ci < n_cases ? __ldg(case_list + ci) : -1;
    int i1 = ci + stride < n_cases ? __ldg(case_list + ci + stride) : -1;
    long long beg = 0, end = 0;
    if (i0 >= 0) { beg = __ldg(adj_ptr + i0); end = __ldg(adj_ptr + i0 + 1); }
    for (; ci < n_cases; ci += stride) {
        long long beg1 = 0, end1 = 0;
        if (i1 >= 0) { beg1 = __ldg(adj_ptr + i1); end1 = __ldg(adj_ptr + i1 + 1); }
        const int i2 = (unsigned long long)ci + 2ull * stride < n_cases ? __ldg(case_list + ci + 2 * stride) : -1;
        for (long long off = beg + gl; off < end; off += G) {
            const uint4 en = __ldg(adj + off);
            const int l = (int)(en.z >> 1);
            if (!((layer_mask >> l) & 1u)) continue;
            const int q = T.entry_of_layer[l];
            if (q < 0) continue;                                       // layer not traced
            if (S) trace_notify2(P, S, T, q, (int)en.x);
            else trace_notify(P, T, q, (int)en.x);
        }
        i0 = i1; i1 = i2; beg = beg1; end = end1;
    }
}

static int launch_trace_sparse(const PeoplePtrs& P, const TraceTable& T, const long long* adj_ptr, const uint4* adj, const int32_t* case_list,
                               const unsigned int* n_case, uint32_t layer_mask, uint32_t* S, double mean_row, cudaStream_t st) {
    const int grid = 148 * 8;
    if (mean_row >= 24.0) trace_sparse_kernel<32><<<grid, kThreads, 0, st>>>(P, T, adj_ptr, adj, case_list, n_case, layer_mask, S);
    else if (mean_row >= 12.0) trace_sparse_kernel<16><<<grid, kThreads, 0, st>>>(P, T, adj_ptr, adj, case_list, n_case, layer_mask, S);
    else if (mean_row >= 6.0) trace_sparse_kernel<8><<<grid, kThreads, 0, st>>>(P, T, adj_ptr, adj, case_list, n_case, layer_mask, S);
    else trace_sparse_kernel<4><<<grid, kThreads, 0, st>>>(P, T, adj_ptr, adj, case_list, n_case, layer_mask, S);
    CVB_LAUNCH_CHECK();
    return 0;
}

__global__ void __launch_bounds__(kThreads) trace_sparse2_kernel(PeoplePtrs P, uint32_t* __restrict__ S, const __grid_constant__ TraceTable T,
        const uint4* __restrict__ adj, const uint4* __restrict__ case_ent, const unsigned int* __restrict__ n_case_ptr, uint32_t layer_mask) {
    pdl_trigger();
    pdl_wait();                                                        // the cases are written by day_begin_kernel
    const unsigned int n_cases = __ldcg(n_case_ptr);                   // (L2 loads: this kernel may have been resident while they were written)
    const int lane = lane_id();
    const unsigned int warps_total = (gridDim.x * blockDim.x) >> 5;
    for (unsigned int ci = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ci < n_cases; ci += warps_total) {
        const uint4 ce = __ldcg(case_ent + ci);
        const int len = (int)ce.y;
        const long long beg = (long long)(((unsigned long long)ce.w << 32) | (unsigned long long)ce.z);
        for (int off = lane; off < len; off += 32) {
            const uint4 en = __ldg(adj + beg + off);
            const int l = (int)(en.z >> 1);
            if (!((layer_mask >> l) & 1u)) continue;
            const int q = T.entry_of_layer[l];
            if (q < 0) continue;                                       // layer not traced
            trace_notify2(P, S, T, q, (int)en.x);
        }
    }
}

// One streaming pass over (p1, p2) of every traced layer; the case bitmap is staged in shared memory when it fits
template <bool SMEM_BITS, int THREADS>
__global__ void __launch_bounds__(THREADS) trace_edges_kernel(PeoplePtrs P, const __grid_constant__ TraceTable T,
                                                              const unsigned int* __restrict__ case_bits) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kTile = THREADS * kEdgesPerThread;
    const unsigned int* bits = case_bits;
    if (SMEM_BITS) {
        unsigned int* s_bits = reinterpret_cast<unsigned int*>(smem_raw);
        unsigned any = 0;
        for (int64_t wd = threadIdx.x; wd < T.n_words; wd += THREADS) { const unsigned v = case_bits[wd]; s_bits[wd] = v; any |= v; }
        if (!__syncthreads_or((int)(any != 0))) return;              // nobody was diagnosed today
        bits = s_bits;
    }
    const int64_t total_tiles = T.tile_start[T.n_entries];
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int q = 0;
#pragma unroll
        for (int j = 1; j < CVB_MAX_LAYERS; ++j) q += (j < T.n_entries && tile >= T.tile_start[j]);
        const int64_t ne = T.n_edges[q];
        const int64_t e0 = (tile - T.tile_start[q]) * kTile + (int64_t)threadIdx.x * kEdgesPerThread;
        if (e0 >= ne) continue;
        const int32_t* __restrict__ p1 = T.p1[q];
        const int32_t* __restrict__ p2 = T.p2[q];
        int a[4], b[4], cnt;
        if (e0 + 4 <= ne) {
            const int4 va = ld_stream(reinterpret_cast<const int4*>(p1 + e0)), vb = ld_stream(reinterpret_cast<const int4*>(p2 + e0));
            a[0] = va.x; a[1] = va.y; a[2] = va.z; a[3] = va.w; b[0] = vb.x; b[1] = vb.y; b[2] = vb.z; b[3] = vb.w; cnt = 4;
        } else {
            cnt = (int)(ne - e0);
#pragma unroll
            for (int k = 0; k < 4; ++k) { a[k] = k < cnt ? p1[e0 + k] : 0; b[k] = k < cnt ? p2[e0 + k] : 0; }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k >= cnt) break;
            const unsigned wa = bits[a[k] >> 5], wb = bits[b[k] >> 5];
            if ((wa >> (a[k] & 31)) & 1u) trace_notify(P, T, q, b[k]);
            if ((wb >> (b[k] & 31)) & 1u) trace_notify(P, T, q, a[k]);
        }
    }
}

// ================================================================================================
// vaccinate_prob
// ================================================================================================
__global__ void __launch_bounds__(kThreads) vaccinate_kernel(PeoplePtrs P, const __grid_constant__ cvb_vaccinate_pars vp, uint64_t seed,
        int64_t n, int64_t id0, int32_t t, int32_t* __restrict__ iv_doses, int32_t* __restrict__ due_day, unsigned long long* __restrict__ counters,
        const double* __restrict__ prob_override, uint32_t* __restrict__ S /* fused day: packed state words, or NULL */,
        unsigned long long* __restrict__ vcounters_row, int32_t nv, const double* __restrict__ tape = nullptr /* verification: the NAb sample of agent i */) {
    __shared__ int s_cnt[2];
    __shared__ int s_delta[kStockSlots];
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x < kStockSlots) s_delta[threadIdx.x] = 0;
    __syncthreads();
    int c_doses = 0, c_new = 0;
    uint8_t* vaccinated = PB(P, vaccinated); const uint8_t* dead = PB(P, dead);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        bool picked = false;
        const bool vacc = vaccinated[i] != 0;
        if (vp.first_dose_today) {                                     // interventions.py:1631-1653 select_people
            const bool eligible = vp.booster ? vacc : !vacc;
            double prob = eligible ? vp.prob : 0.0;
            if (prob_override) {                                       // subtarget (interventions.py:1644-1647): explicit probabilities win
                const double ov = prob_override[i];
                if (ov == ov) prob = ov;
            }
            if (prob > 0.0 && keyed_uniform(seed, P_VACC, (uint32_t)vp.index, t, i + id0, 0) < prob) {
                picked = true;
                if (vp.interval >= 0 && t + vp.interval < vp.n_days) due_day[i] = t + vp.interval;
            }
        }
        if (due_day[i] == t) picked = true;                            // second dose (interventions.py:1655-1660)
        if (!picked) continue;
        // BaseVaccination.vaccinate (interventions.py:1428-1482)
        if (dead[i]) continue;
        const int d = iv_doses[i];
        if (!(d < vp.max_doses)) continue;
        iv_doses[i] = d + 1;
        ++c_doses;
        c_new += !vacc;
        vaccinated[i] = 1;
        PI(P, vaccine_source)[i] = vp.vaccine_index;
        PI(P, doses)[i] += 1;
        PF(P, date_vaccinated)[i] = (float)t;
        // update_peak_nab with the vaccine's parameters (immunity.py:138-202, symp=None)
        if (PF(P, nab)[i] > 0.0f) {
            PF(P, peak_nab)[i] = vp.nab_boost_is_f64 ? (float)dmul((double)PF(P, peak_nab)[i], vp.nab_boost_f64) : fmul(PF(P, peak_nab)[i], vp.nab_boost);
        } else {
            double x = tape ? tape[i] : dist_from_normal(vp.nab_init, keyed_normal(seed, P_NAB_VACC, (uint32_t)vp.index, t, i + id0, 0));
            PF(P, peak_nab)[i] = (float)pow(2.0, x);
        }
        PI(P, t_nab_event)[i] = t;
        if (S) {                                                       // vaccinated, with antibodies from now on
            const uint32_t so = S[i], sn = so | SB_VACC | SB_HAS_NAB;
            if (sn != so) { S[i] = sn; stock_delta(so, sn, s_delta); }
        }
    }
    int w0 = __reduce_add_sync(0xFFFFFFFFu, c_doses), w1 = __reduce_add_sync(0xFFFFFFFFu, c_new);
    if (lane_id() == 0) { if (w0) atomicAdd(&s_cnt[0], w0); if (w1) atomicAdd(&s_cnt[1], w1); }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long* row = counters + (int64_t)t * CVB_N_COUNTERS;
        if (s_cnt[0]) atomicAdd(row + CVB_C_new_doses, (unsigned long long)s_cnt[0]);
        if (s_cnt[1]) atomicAdd(row + CVB_C_new_vaccinated, (unsigned long long)s_cnt[1]);
    }
    if (S) flush_stock_delta(s_delta, counters + (int64_t)t * CVB_N_COUNTERS, vcounters_row, nv);
}

// ================================================================================================
// dynamic layer regeneration (frac = 1): every edge gets two fresh uniformly random endpoints
// ================================================================================================
__global__ void __launch_bounds__(kThreads) layer_regen_kernel(int32_t* __restrict__ p1, int32_t* __restrict__ p2, float* __restrict__ beta,
        int64_t n_edges, int64_t n, uint64_t seed, int32_t layer, int32_t t) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
        u32x4 r = keyed_words(seed, P_DYNLAYER, (uint32_t)layer, t, e, 0);
        int64_t a = (int64_t)dmul(u53(r.x, r.y), (double)n), b = (int64_t)dmul(u53(r.z, r.w), (double)n);
        p1[e] = (int32_t)(a < n - 1 ? a : n - 1);
        p2[e] = (int32_t)(b < n - 1 ? b : n - 1);
        beta[e] = 1.0f;
    }
}

// the same for a chosen subset of the edges (frac < 1): edge e keeps the key it has in a full regeneration
__global__ void __launch_bounds__(kThreads) layer_regen_list_kernel(int32_t* __restrict__ p1, int32_t* __restrict__ p2, float* __restrict__ beta,
        const int64_t* __restrict__ inds, int64_t n_inds, int64_t n_edges, int64_t n, uint64_t seed, int32_t layer, int32_t t) {
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_inds; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = inds[j];
        if (e < 0 || e >= n_edges) continue;
        u32x4 r = keyed_words(seed, P_DYNLAYER, (uint32_t)layer, t, e, 0);
        int64_t a = (int64_t)dmul(u53(r.x, r.y), (double)n), b = (int64_t)dmul(u53(r.z, r.w), (double)n);
        p1[e] = (int32_t)(a < n - 1 ? a : n - 1);
        p2[e] = (int32_t)(b < n - 1 ? b : n - 1);
        beta[e] = 1.0f;
    }
}

}  // namespace cvb

using namespace cvb;

extern "C" {

int cvb_layer_regenerate_list(cvb_sim* s, int32_t layer, int32_t t, const int64_t* inds, int64_t n_inds, cvb_stream st) {
    CVB_REQUIRE(s && layer >= 0 && layer < CVB_MAX_LAYERS && (n_inds == 0 || inds), "cvb_layer_regenerate_list: bad argument");
    cvb::LayerPtrs& L = s->layers[layer];
    if (L.n_edges == 0 || n_inds == 0) return 0;
    layer_regen_list_kernel<<<grid_for(n_inds), kThreads, 0, (cudaStream_t)st>>>(L.p1, L.p2, L.beta, inds, n_inds, L.n_edges, s->n, s->seed, layer, t);
    CVB_LAUNCH_CHECK();
    return 0;
}

static int test_prob_impl(cvb_sim* s, int32_t t, const cvb_test_prob_pars* tp, const double* prob_override, const double* tape, cudaStream_t st) {
    if (s) cvb::state_touched(s);
    CVB_REQUIRE(s && tp && s->res.counters, "cvb_test_prob: handle not ready");
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_test_prob: day %d outside [0,%d)", t, s->npts);
    uintptr_t al = 0;
    for (int f = 0; f < CVB_N_FIELDS; ++f) al |= (uintptr_t)s->people.f[f];
    test_prob_kernel<<<grid_for((s->n + kAPT - 1) / kAPT, kThreads, 148 * 8), kThreads, 0, st>>>(
        s->people, *tp, s->seed, s->n, s->partitioned ? s->id0 : 0, t, (al & 15) == 0 && s->n % 4 == 0, s->res.counters, prob_override, tape);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_test_prob(cvb_sim* s, int32_t t, const cvb_test_prob_pars* tp, const double* prob_override, cvb_stream st) {
    return test_prob_impl(s, t, tp, prob_override, nullptr, (cudaStream_t)st);
}

int cvb_test_prob_taped(cvb_sim* s, int32_t t, const cvb_test_prob_pars* tp, const double* prob_override, const double* tape, cvb_stream st) {
    CVB_REQUIRE(tape, "cvb_test_prob_taped: NULL tape");
    return test_prob_impl(s, t, tp, prob_override, tape, (cudaStream_t)st);
}

// today's cases -> bitmap (+ compact list): phase one of contact tracing
static int trace_select(cvb_sim* s, int32_t t, const cvb_trace_pars* tr, cudaStream_t st) {
    uintptr_t al = (uintptr_t)s->people.f[CVB_F_date_diagnosed];
    CVB_CHECK(cudaMemsetAsync(s->n_case_list, 0, sizeof(unsigned int), st));
    trace_select_kernel<<<grid_for((s->n + 3) / 4, kThreads, 148 * 8), kThreads, 0, st>>>(s->people, s->n, t, tr->presumptive,
        (al & 15) == 0, s->partitioned ? s->case_bits_local : s->case_bits, s->case_list, s->n_case_list);
    CVB_LAUNCH_CHECK();
    return 0;
}

// every traced layer gets a table entry (probability, notification day, ring slot); only the layers that are NOT
// covered by an adjacency (`adj_mask`) contribute tiles to the dense streaming pass
static int build_trace_table(cvb_sim* s, int32_t t, const cvb_trace_pars* tr, uint32_t adj_mask, int tile_edges, TraceTable& T,
                             int64_t& acc, bool& any_sparse) {
    memset(&T, 0, sizeof(T));
    T.seed = s->seed; T.n = s->n; T.n_words = (s->n + 31) / 32; T.t = t; T.index = tr->index;
    T.id0 = s->partitioned ? s->id0 : 0;
    T.end_day = (float)(t + tr->quar_period);
    acc = 0;
    any_sparse = false;
    int q = 0;
    for (int l = 0; l < CVB_MAX_LAYERS; ++l) T.entry_of_layer[l] = -1;
    for (int l = 0; l < s->pars.n_layers; ++l) {
        const bool sparse = (adj_mask >> l) & 1u;
        if (!(tr->trace_prob[l] > 0.0) || (s->layers[l].n_edges == 0 && !sparse)) continue;
        CVB_REQUIRE(tr->trace_time[l] >= 0 && tr->trace_time[l] < s->quar_horizon,
                    "cvb_contact_tracing: trace_time %d needs cvb_set_quar_horizon(%d)", tr->trace_time[l], tr->trace_time[l] + 1);
        const int notify = t + tr->trace_time[l];
        any_sparse |= sparse;
        T.p1[q] = s->layers[l].p1; T.p2[q] = s->layers[l].p2; T.n_edges[q] = sparse ? 0 : s->layers[l].n_edges;
        T.quar_slot[q] = (int*)(s->quar_ring + (int64_t)(notify % s->quar_horizon) * s->n);
        T.trace_prob[q] = tr->trace_prob[l];
        T.notify_day[q] = (float)notify;
        T.layer_id[q] = l;
        T.entry_of_layer[l] = q;
        T.tile_start[q] = acc;
        acc += (T.n_edges[q] + tile_edges - 1) / tile_edges;
        ++q;
    }
    T.n_entries = q;
    for (int j = q; j <= CVB_MAX_LAYERS; ++j) T.tile_start[j] = acc;
    return 0;
}

}  // extern "C"

// fused day: one registered vaccinate_prob intervention on a day it acts (first doses and / or second doses due), state words kept in step
int cvb::launch_vaccinate_fused(cvb_sim* s, int32_t t, const cvb_vaccinate_pars* vp, int32_t* iv_doses, int32_t* due_day, cudaStream_t st) {
    vaccinate_kernel<<<grid_for(s->n), kThreads, 0, st>>>(s->people, *vp, s->seed, s->n, s->partitioned ? s->id0 : 0, t, iv_doses, due_day, s->res.counters, nullptr,
                                                          s->state, s->res.vcounters + (int64_t)t * s->nv * CVB_N_VCOUNTERS, s->nv);
    CVB_LAUNCH_CHECK();
    return 0;
}

// agent-partitioned fused day: the LOCAL contacts of every GLOBAL case (the all-gathered bitmap), with the state words kept in step
int cvb::launch_trace_partition(cvb_sim* s, int32_t t, const cvb_trace_pars* tr, cudaStream_t st) {
    CVB_REQUIRE(s->padj_ptr && s->case_bits_global, "fused day: partitioned adjacency / case bitmap not bound");
    TraceTable T;
    int64_t acc;
    bool any_sparse;
    if (build_trace_table(s, t, tr, s->padj_layer_mask, kTileEdges, T, acc, any_sparse)) return 1;
    CVB_REQUIRE(acc == 0, "fused day: a traced layer is not covered by the partitioned adjacency");
    if (!any_sparse) return 0;
    if (cvb::list_from_bits(s, s->case_bits_global, s->n_slots / 32, st)) return 1;
    return launch_trace_sparse(s->people, T, s->padj_ptr, s->padj, s->glist, s->n_glist, s->padj_layer_mask, s->state,
                               (double)s->padj_entries / (double)(s->n_global > 0 ? s->n_global : 1), st);
}

int cvb::launch_trace_sparse2(cvb_sim* s, int32_t t, const cvb_trace_pars* tr, cudaStream_t st) {
    TraceTable T;
    int64_t acc;
    bool any_sparse;
    if (build_trace_table(s, t, tr, s->adj_layer_mask, kTileEdges, T, acc, any_sparse)) return 1;
    CVB_REQUIRE(acc == 0, "fused day: a traced layer is not covered by the adjacency");
    if (!any_sparse) return 0;
    {
        const uint4* adj = s->adj; const uint4* ce = s->case_ent; const unsigned int* nc = s->n_case_list;
        CVB_CHECK(launch_pdl(trace_sparse2_kernel, 148 * 2, kThreads, 0, st, s->people, s->state, T, adj, ce, nc, s->adj_layer_mask));
    }
    CVB_LAUNCH_CHECK();
    return 0;
}

extern "C" {

int cvb_trace_select_cases(cvb_sim* s, int32_t t, const cvb_trace_pars* tr, cvb_stream st) {
    CVB_REQUIRE(s && tr && s->pars_set, "cvb_trace_select_cases: handle not ready");
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_trace_select_cases: day %d outside [0,%d)", t, s->npts);
    return trace_select(s, t, tr, (cudaStream_t)st);
}

int cvb_trace_notify_contacts(cvb_sim* s, int32_t t, const cvb_trace_pars* tr, cvb_stream st_) {
    if (s) cvb::state_touched(s);
    cudaStream_t st = (cudaStream_t)st_;
    CVB_REQUIRE(s && tr && s->pars_set, "cvb_trace_notify_contacts: handle not ready");
    CVB_REQUIRE(s->partitioned, "cvb_trace_notify_contacts: only for agent-partitioned handles (use cvb_contact_tracing)");
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_trace_notify_contacts: day %d outside [0,%d)", t, s->npts);
    CVB_REQUIRE(s->padj_ptr && s->case_bits_global, "cvb_trace_notify_contacts: partitioned adjacency / case bitmap not bound");
    TraceTable T;
    int64_t acc;
    bool any_sparse;
    if (build_trace_table(s, t, tr, s->padj_layer_mask, kTileEdges, T, acc, any_sparse)) return 1;
    CVB_REQUIRE(acc == 0, "cvb_trace_notify_contacts: a traced layer is not covered by the partitioned adjacency");
    if (!any_sparse) return 0;
    if (cvb::list_from_bits(s, s->case_bits_global, s->n_slots / 32, st)) return 1;
    return launch_trace_sparse(s->people, T, s->padj_ptr, s->padj, s->glist, s->n_glist, s->padj_layer_mask, nullptr,
                               (double)s->padj_entries / (double)(s->n_global > 0 ? s->n_global : 1), st);
}

int cvb_test_num_keys(cvb_sim* s, int32_t t, const cvb_test_num_pars* tp, double* weight, double* key, cvb_stream st) {
    CVB_REQUIRE(s && tp && weight && key, "cvb_test_num_keys: bad argument");
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_test_num_keys: day %d outside [0,%d)", t, s->npts);
    test_num_keys_kernel<<<grid_for(s->n, kThreads, 148 * 8), kThreads, 0, (cudaStream_t)st>>>(s->people, *tp, s->seed, s->n,
        s->partitioned ? s->id0 : 0, t, weight, key);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_test_list(cvb_sim* s, int32_t t, const int32_t* inds, int64_t n_inds, double sensitivity, double loss_prob, int32_t test_delay,
                  int32_t index, cvb_stream st) {
    if (s) cvb::state_touched(s);
    CVB_REQUIRE(s && (n_inds == 0 || inds), "cvb_test_list: bad argument");
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_test_list: day %d outside [0,%d)", t, s->npts);
    if (n_inds == 0) return 0;
    test_list_kernel<<<grid_for(n_inds), kThreads, 0, (cudaStream_t)st>>>(s->people, inds, n_inds, s->seed, s->n, s->partitioned ? s->id0 : 0, t,
                                                                          sensitivity, loss_prob, test_delay, (uint32_t)index);
    CVB_LAUNCH_CHECK();
    return 0;
}

// notify the contacts of the current case set (bitmap + list) of a non-partitioned handle: phase two of contact tracing
static int trace_notify_local(cvb_sim* s, int32_t t, const cvb_trace_pars* tr, cudaStream_t st, const double* tape = nullptr) {
    const uint32_t adj_mask = (s->adj && s->adj_layer_mask) ? s->adj_layer_mask : 0u;
    const size_t bitmap_bytes = (size_t)((s->n + 31) / 32) * sizeof(unsigned int);
    const bool smem_bits = bitmap_bytes <= 200 * 1024;
    const int threads = smem_bits ? 1024 : 256;
    TraceTable T;
    int64_t acc;
    bool any_sparse;
    if (build_trace_table(s, t, tr, adj_mask, threads * kEdgesPerThread, T, acc, any_sparse)) return 1;
    T.tape = tape;
    if (any_sparse) {
        if (launch_trace_sparse(s->people, T, s->adj_ptr, s->adj, s->case_list, s->n_case_list, adj_mask, nullptr, 36.0, st)) return 1;
    }
    if (acc == 0) return 0;
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, s->device);
    if (smem_bits) {
        static bool configured[64] = {false};                        // the attribute is per device
        if (!configured[s->device & 63]) {
            CVB_CHECK(cudaFuncSetAttribute(trace_edges_kernel<true, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            configured[s->device & 63] = true;
        }
        const int grid = (int)(acc < n_sm ? acc : n_sm);
        trace_edges_kernel<true, 1024><<<grid, 1024, bitmap_bytes, st>>>(s->people, T, s->case_bits);
    } else {
        const int grid = (int)(acc < (int64_t)n_sm * 8 ? acc : (int64_t)n_sm * 8);
        trace_edges_kernel<false, 256><<<grid, 256, 0, st>>>(s->people, T, s->case_bits);
    }
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_contact_tracing(cvb_sim* s, int32_t t, const cvb_trace_pars* tr, cvb_stream st_) {
    if (s) cvb::state_touched(s);
    cudaStream_t st = (cudaStream_t)st_;
    CVB_REQUIRE(s && tr && s->pars_set, "cvb_contact_tracing: handle not ready");
    CVB_REQUIRE(!s->partitioned, "cvb_contact_tracing: agent-partitioned handles trace in two phases (cvb_trace_select_cases, all-gather, cvb_trace_notify_contacts)");
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_contact_tracing: day %d outside [0,%d)", t, s->npts);
    if (trace_select(s, t, tr, st)) return 1;
    return trace_notify_local(s, t, tr, st);
}

int cvb_contact_tracing_taped(cvb_sim* s, int32_t t, const cvb_trace_pars* tr, const double* tape, cvb_stream st_) {
    if (s) cvb::state_touched(s);
    cudaStream_t st = (cudaStream_t)st_;
    CVB_REQUIRE(s && tr && tape && s->pars_set && !s->partitioned, "cvb_contact_tracing_taped: bad argument");
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_contact_tracing_taped: day %d outside [0,%d)", t, s->npts);
    if (trace_select(s, t, tr, st)) return 1;
    return trace_notify_local(s, t, tr, st, tape);
}

int cvb_pending_quarantine(cvb_sim* s, int32_t start_day, float* out_end_day, cvb_stream st) {
    CVB_REQUIRE(s && out_end_day && start_day >= 0, "cvb_pending_quarantine: bad argument");
    const int slot = start_day % s->quar_horizon;
    CVB_CHECK(cudaMemcpyAsync(out_end_day, s->quar_ring + (int64_t)slot * s->n, (size_t)s->n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)st));
    return 0;
}

int cvb_set_pending_quarantine(cvb_sim* s, int32_t start_day, const float* end_day, cvb_stream st) {
    if (s) cvb::state_touched(s);
    CVB_REQUIRE(s && end_day && start_day >= 0, "cvb_set_pending_quarantine: bad argument");
    const int slot = start_day % s->quar_horizon;
    CVB_CHECK(cudaMemcpyAsync(s->quar_ring + (int64_t)slot * s->n, end_day, (size_t)s->n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)st));
    return 0;
}

int cvb_contact_tracing_list(cvb_sim* s, int32_t t, const cvb_trace_pars* tr, const int32_t* case_inds, int64_t n_cases, cvb_stream st_) {
    if (s) cvb::state_touched(s);
    cudaStream_t st = (cudaStream_t)st_;
    CVB_REQUIRE(s && tr && s->pars_set && (n_cases == 0 || case_inds), "cvb_contact_tracing_list: bad argument");
    CVB_REQUIRE(!s->partitioned, "cvb_contact_tracing_list: not for agent-partitioned handles");
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_contact_tracing_list: day %d outside [0,%d)", t, s->npts);
    CVB_REQUIRE(n_cases <= s->n, "cvb_contact_tracing_list: more cases than agents");
    if (n_cases == 0) return 0;
    CVB_CHECK(cudaMemsetAsync(s->case_bits, 0, (size_t)((s->n + 31) / 32) * sizeof(unsigned int), st));
    set_cases_kernel<<<grid_for(n_cases), kThreads, 0, st>>>(case_inds, n_cases, s->n, s->case_bits, s->case_list, s->n_case_list);
    CVB_LAUNCH_CHECK();
    return trace_notify_local(s, t, tr, st);
}

int cvb_vaccinate_prob(cvb_sim* s, int32_t t, const cvb_vaccinate_pars* vp, int32_t* iv_doses, int32_t* due_day, const double* prob_override,
                       cvb_stream st) {
    if (s) cvb::state_touched(s);
    CVB_REQUIRE(s && vp && iv_doses && due_day && s->res.counters, "cvb_vaccinate_prob: bad argument");
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_vaccinate_prob: day %d outside [0,%d)", t, s->npts);
    CVB_REQUIRE(vp->vaccine_index >= 0 && vp->vaccine_index < CVB_MAX_VACCINES, "cvb_vaccinate_prob: vaccine index out of range");
    vaccinate_kernel<<<grid_for(s->n), kThreads, 0, (cudaStream_t)st>>>(s->people, *vp, s->seed, s->n, s->partitioned ? s->id0 : 0, t, iv_doses, due_day, s->res.counters, prob_override,
                                                                      nullptr, nullptr, s->nv);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_vaccinate_taped(cvb_sim* s, int32_t t, const cvb_vaccinate_pars* vp, int32_t* iv_doses, int32_t* due_day, const double* prob_override,
                        const double* tape, cvb_stream st) {
    if (s) cvb::state_touched(s);
    CVB_REQUIRE(s && vp && iv_doses && due_day && tape && s->res.counters, "cvb_vaccinate_taped: bad argument");
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_vaccinate_taped: day %d outside [0,%d)", t, s->npts);
    vaccinate_kernel<<<grid_for(s->n), kThreads, 0, (cudaStream_t)st>>>(s->people, *vp, s->seed, s->n, s->partitioned ? s->id0 : 0, t, iv_doses, due_day, s->res.counters, prob_override,
                                                                      nullptr, nullptr, s->nv, tape);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_layer_regenerate(cvb_sim* s, int32_t layer, int32_t t, cvb_stream st) {
    CVB_REQUIRE(s && layer >= 0 && layer < CVB_MAX_LAYERS, "cvb_layer_regenerate: bad argument");
    cvb::LayerPtrs& L = s->layers[layer];
    if (L.n_edges == 0) return 0;
    layer_regen_kernel<<<grid_for(L.n_edges), kThreads, 0, (cudaStream_t)st>>>(L.p1, L.p2, L.beta, L.n_edges, s->n, s->seed, layer, t);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_step_day(cvb_sim* s, int32_t t, cvb_stream st) {
    int rc;
    if ((rc = cvb_update_states_pre(s, t, st))) return rc;
    if ((rc = cvb_post_and_prepare(s, t, st))) return rc;
    if ((rc = cvb_edge_pass(s, t, st))) return rc;
    if ((rc = cvb_infect_winners(s, t, st))) return rc;
    return cvb_update_nab_count(s, t, st);
}

}  // extern "C"
