// Built-in interventions as device passes (native-RNG mode) and dynamic-layer regeneration:
//   test_prob + People.test            reference interventions.py:857-981, people.py:589-617
//   contact_tracing                    reference interventions.py:984-1145, base.py:1808-1846, utils.py:131-147
//   vaccinate_prob + vaccinate         reference interventions.py:1257-1662, immunity.py:138-202
//   Layer.update (frac = 1)            reference base.py:1849-1876
// All Bernoulli draws are keyed per agent (or per agent x layer for tracing), so results are
// independent of thread order; see oracle/cvoracle.py for the CPU restatement they are tested against.
#include "cvb_internal.cuh"

namespace cvb {

// ================================================================================================
// test_prob
// ================================================================================================
__global__ void __launch_bounds__(kThreads) test_prob_kernel(PeoplePtrs P, const __grid_constant__ cvb_test_prob_pars tp, uint64_t seed,
        int64_t n, int32_t t, unsigned long long* __restrict__ counters) {
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int c = 0;
    const float tf = (float)t;
    const uint8_t* symptomatic = PB(P, symptomatic); const uint8_t* diagnosed = PB(P, diagnosed); const uint8_t* quarantined = PB(P, quarantined);
    const uint8_t* infectious = PB(P, infectious); uint8_t* tested = PB(P, tested);
    const float* d_quar = PF(P, date_quarantined); const float* d_end_quar = PF(P, date_end_quarantine);
    float* d_tested = PF(P, date_tested); float* d_diag = PF(P, date_diagnosed); float* d_pos = PF(P, date_pos_test);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double prob = 0.0;
        if (!diagnosed[i]) {                                           // interventions.py:973
            const bool symp = symptomatic[i] != 0;
            bool qt;                                                   // interventions.py:691-715 get_quar_inds
            switch (tp.quar_policy) {
                case 0:  qt = d_quar[i] == tf - 1.0f; break;
                case 1:  qt = d_end_quar[i] == tf + 1.0f; break;
                case 2:  qt = (d_quar[i] == tf - 1.0f) || (d_end_quar[i] == tf + 1.0f); break;
                default: qt = quarantined[i] != 0; break;
            }
            prob = qt ? (symp ? tp.symp_quar_prob : tp.asymp_quar_prob) : (symp ? tp.symp_prob : tp.asymp_prob);
        }
        if (!(prob > 0.0)) continue;
        if (!(keyed_uniform(seed, P_TEST, (uint32_t)tp.index, t, i, 0) < prob)) continue;
        // People.test (people.py:589-617)
        ++c;
        tested[i] = 1;
        d_tested[i] = tf;
        if (!infectious[i]) continue;
        if (!(keyed_uniform(seed, P_TEST_SENS, (uint32_t)tp.index, t, i, 0) < tp.sensitivity)) continue;
        if (!is_nan(d_diag[i])) continue;
        if (!(keyed_uniform(seed, P_TEST_LOSS, (uint32_t)tp.index, t, i, 0) < 1.0 - tp.loss_prob)) continue;
        d_diag[i] = (float)(t + tp.test_delay);
        d_pos[i] = tf;
    }
    int w = __reduce_add_sync(0xFFFFFFFFu, c);
    if (lane_id() == 0 && w) atomicAdd(&s_cnt, w);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(counters + (int64_t)t * CVB_N_COUNTERS + CVB_C_new_tests, (unsigned long long)s_cnt);
}

// ================================================================================================
// contact_tracing
// ================================================================================================
__global__ void __launch_bounds__(kThreads) trace_select_kernel(PeoplePtrs P, int64_t n, int32_t t, int presumptive,
        unsigned int* __restrict__ case_bits, unsigned int* __restrict__ n_cases) {
    // one warp covers 32 consecutive agents = one bitmap word, written without atomics
    const float tf = (float)t;
    const float* d_diag = PF(P, date_diagnosed); const float* d_tested = PF(P, date_tested); const uint8_t* exposed = PB(P, exposed);
    const int64_t n_words = (n + 31) / 32;
    int found = 0;
    for (int64_t wd = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32; wd < n_words; wd += ((int64_t)gridDim.x * blockDim.x) / 32) {
        int64_t i = wd * 32 + lane_id();
        bool is_case = false;
        if (i < n) is_case = presumptive ? (d_tested[i] == tf && exposed[i] != 0) : (d_diag[i] == tf);
        unsigned bits = __ballot_sync(0xFFFFFFFFu, is_case);
        if (lane_id() == 0) { case_bits[wd] = bits; found += __popc(bits); }
    }
    if (found) atomicAdd(n_cases, (unsigned int)found);
}

struct TraceArgs {
    double trace_prob;
    uint64_t seed;
    int64_t n, n_edges;
    int32_t t, layer, sub, notify_day;   // sub = (intervention index << 8) | layer
    float end_day;
};

__device__ __forceinline__ void trace_notify(PeoplePtrs& P, const TraceArgs& ta, int c, int* __restrict__ quar_slot) {
    if (!(keyed_uniform(ta.seed, P_TRACE, (uint32_t)ta.sub, ta.t, c, 0) < ta.trace_prob)) return;
    if (PB(P, dead)[c]) return;                                        // interventions.py:1139-1141
    PB(P, known_contact)[c] = 1;
    // date_known_contact = fmin(old, notify_day): for non-negative floats and NaN the unsigned bit
    // patterns order the same way (NaN = 0x7fc00000 is the largest), so atomicMin does fmin
    atomicMin((unsigned int*)PF(P, date_known_contact) + c, (unsigned int)__float_as_int((float)ta.notify_day));
    atomicMax(quar_slot + c, __float_as_int(ta.end_day));              // people.py:620-640 schedule_quarantine
}

__global__ void __launch_bounds__(kThreads) trace_edges_kernel(PeoplePtrs P, const __grid_constant__ TraceArgs ta,
        const int32_t* __restrict__ p1, const int32_t* __restrict__ p2, const unsigned int* __restrict__ case_bits,
        const unsigned int* __restrict__ n_cases, int* __restrict__ quar_slot) {
    if (*n_cases == 0) return;                                         // nobody to trace today
    const int64_t n_tiles = (ta.n_edges + kTileEdges - 1) / kTileEdges;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t e0 = tile * kTileEdges + (int64_t)threadIdx.x * kEdgesPerThread;
        if (e0 >= ta.n_edges) continue;
        int a[4], b[4], cnt;
        if (e0 + 4 <= ta.n_edges) {
            int4 va = ld_stream(reinterpret_cast<const int4*>(p1 + e0)), vb = ld_stream(reinterpret_cast<const int4*>(p2 + e0));
            a[0] = va.x; a[1] = va.y; a[2] = va.z; a[3] = va.w; b[0] = vb.x; b[1] = vb.y; b[2] = vb.z; b[3] = vb.w; cnt = 4;
        } else {
            cnt = (int)(ta.n_edges - e0);
            for (int k = 0; k < 4; ++k) { a[k] = k < cnt ? p1[e0 + k] : 0; b[k] = k < cnt ? p2[e0 + k] : 0; }
        }
        unsigned wa[4], wb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { wa[k] = __ldg(case_bits + (a[k] >> 5)); wb[k] = __ldg(case_bits + (b[k] >> 5)); }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k >= cnt) break;
            if ((wa[k] >> (a[k] & 31)) & 1u) trace_notify(P, ta, b[k], quar_slot);
            if ((wb[k] >> (b[k] & 31)) & 1u) trace_notify(P, ta, a[k], quar_slot);
        }
    }
}

__global__ void reset_cases_kernel(unsigned int* n_cases) { *n_cases = 0; }

// ================================================================================================
// vaccinate_prob
// ================================================================================================
__global__ void __launch_bounds__(kThreads) vaccinate_kernel(PeoplePtrs P, const __grid_constant__ cvb_vaccinate_pars vp, uint64_t seed,
        int64_t n, int32_t t, int32_t* __restrict__ iv_doses, int32_t* __restrict__ due_day, unsigned long long* __restrict__ counters) {
    __shared__ int s_cnt[2];
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    int c_doses = 0, c_new = 0;
    uint8_t* vaccinated = PB(P, vaccinated); const uint8_t* dead = PB(P, dead);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        bool picked = false;
        const bool vacc = vaccinated[i] != 0;
        if (vp.first_dose_today) {                                     // interventions.py:1631-1653 select_people
            const bool eligible = vp.booster ? vacc : !vacc;
            if (eligible && vp.prob > 0.0 && keyed_uniform(seed, P_VACC, (uint32_t)vp.index, t, i, 0) < vp.prob) {
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
            PF(P, peak_nab)[i] = fmul(PF(P, peak_nab)[i], vp.nab_boost);
        } else {
            double x = dist_from_normal(vp.nab_init, keyed_normal(seed, P_NAB_VACC, (uint32_t)vp.index, t, i, 0));
            PF(P, peak_nab)[i] = (float)pow(2.0, x);
        }
        PI(P, t_nab_event)[i] = t;
    }
    int w0 = __reduce_add_sync(0xFFFFFFFFu, c_doses), w1 = __reduce_add_sync(0xFFFFFFFFu, c_new);
    if (lane_id() == 0) { if (w0) atomicAdd(&s_cnt[0], w0); if (w1) atomicAdd(&s_cnt[1], w1); }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long* row = counters + (int64_t)t * CVB_N_COUNTERS;
        if (s_cnt[0]) atomicAdd(row + CVB_C_new_doses, (unsigned long long)s_cnt[0]);
        if (s_cnt[1]) atomicAdd(row + CVB_C_new_vaccinated, (unsigned long long)s_cnt[1]);
    }
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

}  // namespace cvb

using namespace cvb;

extern "C" {

int cvb_test_prob(cvb_sim* s, int32_t t, const cvb_test_prob_pars* tp, cvb_stream st) {
    CVB_REQUIRE(s && tp && s->res.counters, "cvb_test_prob: handle not ready");
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_test_prob: day %d outside [0,%d)", t, s->npts);
    test_prob_kernel<<<grid_for(s->n), kThreads, 0, (cudaStream_t)st>>>(s->people, *tp, s->seed, s->n, t, s->res.counters);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_contact_tracing(cvb_sim* s, int32_t t, const cvb_trace_pars* tr, cvb_stream st_) {
    cudaStream_t st = (cudaStream_t)st_;
    CVB_REQUIRE(s && tr && s->pars_set, "cvb_contact_tracing: handle not ready");
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_contact_tracing: day %d outside [0,%d)", t, s->npts);
    reset_cases_kernel<<<1, 1, 0, st>>>(s->n_cases);
    CVB_LAUNCH_CHECK();
    trace_select_kernel<<<grid_for(s->n), kThreads, 0, st>>>(s->people, s->n, t, tr->presumptive, s->case_bits, s->n_cases);
    CVB_LAUNCH_CHECK();
    for (int l = 0; l < s->pars.n_layers; ++l) {
        if (!(tr->trace_prob[l] > 0.0) || s->layers[l].n_edges == 0) continue;
        CVB_REQUIRE(tr->trace_time[l] >= 0 && tr->trace_time[l] < s->quar_horizon,
                    "cvb_contact_tracing: trace_time %d needs cvb_set_quar_horizon(%d)", tr->trace_time[l], tr->trace_time[l] + 1);
        TraceArgs ta;
        ta.trace_prob = tr->trace_prob[l];
        ta.seed = s->seed; ta.n = s->n; ta.n_edges = s->layers[l].n_edges;
        ta.t = t; ta.layer = l; ta.sub = (tr->index << 8) | l;
        ta.notify_day = t + tr->trace_time[l];
        ta.end_day = (float)(t + tr->quar_period);                    // start + (quar_period - trace_time), interventions.py:1144
        int slot = ta.notify_day % s->quar_horizon;
        int64_t n_tiles = (ta.n_edges + kTileEdges - 1) / kTileEdges;
        int grid = (int)(n_tiles < 148 * 8 ? n_tiles : 148 * 8);
        trace_edges_kernel<<<grid, kThreads, 0, st>>>(s->people, ta, s->layers[l].p1, s->layers[l].p2, s->case_bits, s->n_cases,
                                                      (int*)(s->quar_ring + (int64_t)slot * s->n));
        CVB_LAUNCH_CHECK();
    }
    return 0;
}

int cvb_vaccinate_prob(cvb_sim* s, int32_t t, const cvb_vaccinate_pars* vp, int32_t* iv_doses, int32_t* due_day, cvb_stream st) {
    CVB_REQUIRE(s && vp && iv_doses && due_day && s->res.counters, "cvb_vaccinate_prob: bad argument");
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_vaccinate_prob: day %d outside [0,%d)", t, s->npts);
    CVB_REQUIRE(vp->vaccine_index >= 0 && vp->vaccine_index < CVB_MAX_VACCINES, "cvb_vaccinate_prob: vaccine index out of range");
    vaccinate_kernel<<<grid_for(s->n), kThreads, 0, (cudaStream_t)st>>>(s->people, *vp, s->seed, s->n, t, iv_doses, due_day, s->res.counters);
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
    if ((rc = cvb_update_states_post(s, t, st))) return rc;
    if ((rc = cvb_prepare_transmission(s, t, st))) return rc;
    if ((rc = cvb_edge_pass(s, t, st))) return rc;
    if ((rc = cvb_infect_winners(s, t, st))) return rc;
    return cvb_update_nab_count(s, t, st);
}

}  // extern "C"
