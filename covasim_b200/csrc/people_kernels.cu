// Per-agent state-transition kernels over the structure-of-arrays People state:
//   update_states_pre  (+ check_immunity)      reference people.py:164-186, immunity.py:303-350
//   update_states_post                          reference people.py:189-196, 315-366
//   prepare_transmission (viral load + per-layer rel_trans / rel_sus records)   sim.py:602-643
//   update_nab + stock counts + population means                                 immunity.py:205-213, sim.py:652-674
// One thread per agent, coalesced SoA accesses, flows reduced per CTA and added to the per-day
// counter row with one atomic per counter per CTA.  HBM-bound: at most one read and one write of the
// touched arrays per day.
#include "cvb_internal.cuh"

namespace cvb {

// ================================================================================================
// update_states_pre
// ================================================================================================
enum { PRE_INFECTIOUS = 0, PRE_SYMPTOMATIC, PRE_SEVERE, PRE_CRITICAL, PRE_RECOVERIES, PRE_DEATHS, PRE_KNOWN_DEATHS,
       PRE_BED_SEVERE, PRE_BED_CRITICAL, PRE_NK };

__global__ void __launch_bounds__(kThreads) states_pre_kernel(PeoplePtrs P, const __grid_constant__ cvb_pars pars, int64_t n, int32_t t,
        unsigned long long* __restrict__ counters, unsigned long long* __restrict__ vcounters, unsigned long long* __restrict__ beds) {
    __shared__ int s_cnt[PRE_NK + CVB_MAX_VARIANTS];
    const int nv = pars.n_variants;
    if (threadIdx.x < PRE_NK + CVB_MAX_VARIANTS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    int c[PRE_NK];
#pragma unroll
    for (int k = 0; k < PRE_NK; ++k) c[k] = 0;
    int cv[CVB_MAX_VARIANTS];
#pragma unroll
    for (int k = 0; k < CVB_MAX_VARIANTS; ++k) cv[k] = 0;

    uint8_t* exposed = PB(P, exposed); uint8_t* infectious = PB(P, infectious); uint8_t* symptomatic = PB(P, symptomatic);
    uint8_t* severe = PB(P, severe); uint8_t* critical = PB(P, critical); uint8_t* recovered = PB(P, recovered);
    uint8_t* dead = PB(P, dead); uint8_t* diagnosed = PB(P, diagnosed); uint8_t* susceptible = PB(P, susceptible);
    uint8_t* isolated = PB(P, isolated); uint8_t* known_dead = PB(P, known_dead); uint8_t* known_contact = PB(P, known_contact);
    uint8_t* quarantined = PB(P, quarantined); uint8_t* vaccinated = PB(P, vaccinated);
    uint8_t* exp_by_var = PB(P, exposed_by_variant); uint8_t* inf_by_var = PB(P, infectious_by_variant);
    float* exp_var = PF(P, exposed_variant); float* inf_var = PF(P, infectious_variant); float* rec_var = PF(P, recovered_variant);
    const float* d_inf = PF(P, date_infectious); const float* d_symp = PF(P, date_symptomatic); const float* d_sev = PF(P, date_severe);
    const float* d_crit = PF(P, date_critical); const float* d_rec = PF(P, date_recovered); const float* d_dead = PF(P, date_dead);
    const float* d_end_iso = PF(P, date_end_isolation);
    const float* nab = PF(P, nab); const int32_t* vsrc = PI(P, vaccine_source);
    float* sus_imm = PF(P, sus_imm); float* symp_imm = PF(P, symp_imm); float* sev_imm = PF(P, sev_imm);

    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool was_exposed = exposed[i] != 0;          // is_exp is taken once, before any transition (people.py:169)
        float date_rec_i = d_rec[i];
        bool sev_now = severe[i] != 0, crit_now = critical[i] != 0;
        if (was_exposed) {
            // infectious (people.py:222-232)
            if (!infectious[i] && due(d_inf[i], t)) {
                infectious[i] = 1;
                float ev = exp_var[i];
                inf_var[i] = ev;
                int v = (int)ev;
                if (v >= 0 && v < nv) {
                    inf_by_var[(int64_t)v * n + i] = 1;
#pragma unroll
                    for (int k = 0; k < CVB_MAX_VARIANTS; ++k) cv[k] += (k == v);      // static indices keep cv[] in registers
                }
                ++c[PRE_INFECTIOUS];
            }
            // symptomatic / severe / critical (people.py:235-253)
            if (!symptomatic[i] && due(d_symp[i], t)) { symptomatic[i] = 1; ++c[PRE_SYMPTOMATIC]; }
            if (!sev_now && due(d_sev[i], t)) { severe[i] = 1; sev_now = true; ++c[PRE_SEVERE]; }
            if (!crit_now && due(d_crit[i], t)) { critical[i] = 1; crit_now = true; ++c[PRE_CRITICAL]; }
            // recovery (people.py:256-291)
            if (!recovered[i] && due(date_rec_i, t)) {
                exposed[i] = 0; infectious[i] = 0; symptomatic[i] = 0; severe[i] = 0; critical[i] = 0;
                sev_now = false; crit_now = false;
                recovered[i] = 1;
                rec_var[i] = exp_var[i];
                inf_var[i] = nanf32();
                exp_var[i] = nanf32();
                for (int v = 0; v < nv; ++v) { exp_by_var[(int64_t)v * n + i] = 0; inf_by_var[(int64_t)v * n + i] = 0; }
                if (pars.use_waning) { susceptible[i] = 1; diagnosed[i] = 0; }
                ++c[PRE_RECOVERIES];
            }
        }
        // leave isolation (people.py:368-374) -- every agent
        if (isolated[i] && due(d_end_iso[i], t)) isolated[i] = 0;
        if (was_exposed) {
            // death (people.py:294-312)
            if (!dead[i] && due(d_dead[i], t)) {
                dead[i] = 1;
                if (diagnosed[i]) { known_dead[i] = 1; ++c[PRE_KNOWN_DEATHS]; }
                susceptible[i] = 0; exposed[i] = 0; infectious[i] = 0; symptomatic[i] = 0; severe[i] = 0; critical[i] = 0;
                sev_now = false; crit_now = false;
                known_contact[i] = 0; quarantined[i] = 0; recovered[i] = 0;
                inf_var[i] = nanf32(); exp_var[i] = nanf32(); rec_var[i] = nanf32();
                ++c[PRE_DEATHS];
            }
        }
        c[PRE_BED_SEVERE] += sev_now;
        c[PRE_BED_CRITICAL] += crit_now;

        // check_immunity (immunity.py:303-350): float64 arithmetic, rounded once to float32
        if (pars.use_waning) {
            float nab_i = nab[i];
            bool was_inf = due(date_rec_i, t);
            int rv = was_inf ? (int)rec_var[i] : -1;
            bool vacc = pars.has_vaccine_pars && vaccinated[i] != 0;
            int vs = vacc ? vsrc[i] : 0;
            for (int v = 0; v < nv; ++v) {
                double natural = (rv >= 0 && rv < nv) ? (double)pars.immunity[v][rv] : 0.0;
                double vaccine = (vacc && vs >= 0 && vs < CVB_MAX_VACCINES) ? pars.vaccine_imm[vs][v] : 0.0;
                double enab = dmul((double)nab_i, fmax(natural, vaccine));
                float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
                if (enab != 0.0) {       // 0**beta == 0 -> protection exactly 0
                    s0 = calc_ve(enab, pars.exp_alpha_inf, pars.beta_inf);
                    s1 = calc_ve(enab, pars.exp_alpha_symp_inf, pars.beta_symp_inf);
                    s2 = calc_ve(enab, pars.exp_alpha_sev_symp, pars.beta_sev_symp);
                }
                sus_imm[(int64_t)v * n + i] = s0;
                symp_imm[(int64_t)v * n + i] = s1;
                sev_imm[(int64_t)v * n + i] = s2;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < PRE_NK; ++k) {
        int w = __reduce_add_sync(0xFFFFFFFFu, c[k]);
        if (lane_id() == 0 && w) atomicAdd(&s_cnt[k], w);
    }
#pragma unroll
    for (int k = 0; k < CVB_MAX_VARIANTS; ++k) {
        int w = __reduce_add_sync(0xFFFFFFFFu, cv[k]);
        if (lane_id() == 0 && w) atomicAdd(&s_cnt[PRE_NK + k], w);
    }
    __syncthreads();
    if (threadIdx.x < PRE_NK + CVB_MAX_VARIANTS) {
        int v = s_cnt[threadIdx.x];
        if (v) {
            unsigned long long* row = counters + (int64_t)t * CVB_N_COUNTERS;
            switch (threadIdx.x) {
                case PRE_INFECTIOUS:   atomicAdd(row + CVB_C_new_infectious, (unsigned long long)v); break;
                case PRE_SYMPTOMATIC:  atomicAdd(row + CVB_C_new_symptomatic, (unsigned long long)v); break;
                case PRE_SEVERE:       atomicAdd(row + CVB_C_new_severe, (unsigned long long)v); break;
                case PRE_CRITICAL:     atomicAdd(row + CVB_C_new_critical, (unsigned long long)v); break;
                case PRE_RECOVERIES:   atomicAdd(row + CVB_C_new_recoveries, (unsigned long long)v); break;
                case PRE_DEATHS:       atomicAdd(row + CVB_C_new_deaths, (unsigned long long)v); break;
                case PRE_KNOWN_DEATHS: atomicAdd(row + CVB_C_new_known_deaths, (unsigned long long)v); break;
                case PRE_BED_SEVERE:   atomicAdd(beds + (int64_t)t * 2 + 0, (unsigned long long)v); break;
                case PRE_BED_CRITICAL: atomicAdd(beds + (int64_t)t * 2 + 1, (unsigned long long)v); break;
                default: {
                    int var = threadIdx.x - PRE_NK;
                    if (var < nv) atomicAdd(vcounters + ((int64_t)t * nv + var) * CVB_N_VCOUNTERS + CVB_VC_new_infectious_by_variant,
                                            (unsigned long long)v);
                }
            }
        }
    }
}

// ================================================================================================
// update_states_post
// ================================================================================================
enum { POST_DIAGNOSES = 0, POST_QUARANTINED, POST_ISOLATED, POST_NK };

__global__ void __launch_bounds__(kThreads) states_post_kernel(PeoplePtrs P, int64_t n, int32_t t, float* __restrict__ quar_slot,
                                                               unsigned long long* __restrict__ counters) {
    __shared__ int s_cnt[POST_NK];
    if (threadIdx.x < POST_NK) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    int c[POST_NK] = {0, 0, 0};
    uint8_t* diagnosed = PB(P, diagnosed); uint8_t* quarantined = PB(P, quarantined); uint8_t* isolated = PB(P, isolated);
    const uint8_t* dead = PB(P, dead); const uint8_t* recovered = PB(P, recovered);
    float* d_pos = PF(P, date_pos_test); const float* d_diag = PF(P, date_diagnosed); float* d_quar = PF(P, date_quarantined);
    float* d_end_quar = PF(P, date_end_quarantine); float* d_end_iso = PF(P, date_end_isolation); const float* d_rec = PF(P, date_recovered);
    const float tf = (float)t;

    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        // check_diagnosed (people.py:315-332)
        bool diag = diagnosed[i] != 0;
        float ddiag = d_diag[i];
        if (!diag) {
            if (due(d_pos[i], t)) { d_pos[i] = nanf32(); ++c[POST_DIAGNOSES]; }
            if (due(ddiag, t)) { diagnosed[i] = 1; diag = true; }
        }
        // check_quar (people.py:335-358): the pending request holds the max end day asked for today
        bool quar = quarantined[i] != 0;
        float pend = quar_slot[i];
        float end_q = 0.0f;
        bool end_loaded = false;
        if (pend >= 0.0f) {
            quar_slot[i] = -1.0f;
            if (quar) {
                end_q = d_end_quar[i]; end_loaded = true;
                if (pend > end_q) { end_q = pend; d_end_quar[i] = end_q; }           // Python max(old, requested)
            } else if (!(dead[i] || recovered[i] || diag || isolated[i])) {
                quarantined[i] = 1; quar = true;
                d_quar[i] = tf;
                end_q = pend; end_loaded = true;
                d_end_quar[i] = end_q;
                ++c[POST_QUARANTINED];
            }
        }
        if (quar) {
            if (!end_loaded) end_q = d_end_quar[i];
            if (ddiag == tf) { end_q = tf; d_end_quar[i] = tf; }
            if (due(end_q, t)) quarantined[i] = 0;
        }
        // check_enter_iso (people.py:361-366)
        if (ddiag == tf) {
            isolated[i] = 1;
            d_end_iso[i] = d_rec[i];
            ++c[POST_ISOLATED];
        }
    }
#pragma unroll
    for (int k = 0; k < POST_NK; ++k) {
        int w = __reduce_add_sync(0xFFFFFFFFu, c[k]);
        if (lane_id() == 0 && w) atomicAdd(&s_cnt[k], w);
    }
    __syncthreads();
    if (threadIdx.x < POST_NK && s_cnt[threadIdx.x]) {
        unsigned long long* row = counters + (int64_t)t * CVB_N_COUNTERS;
        const int ids[POST_NK] = {CVB_C_new_diagnoses, CVB_C_new_quarantined, CVB_C_new_isolated};
        atomicAdd(row + ids[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
    }
}

__global__ void schedule_quar_kernel(const int32_t* __restrict__ inds, int64_t n_inds, int* __restrict__ slot, float end_day, int64_t n) {
    int bits = __float_as_int(end_day);
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_inds; j += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = inds[j];
        if (i >= 0 && i < n) atomicMax(slot + i, bits);     // non-negative floats order like their bit patterns; -1.0f (empty) is negative
    }
}

// ================================================================================================
// prepare_transmission: viral load + {rel_trans, rel_sus} per layer (sim.py:602-643, utils.py:39-90)
// ================================================================================================
__global__ void __launch_bounds__(kThreads) prepare_kernel(PeoplePtrs P, const __grid_constant__ cvb_pars pars, int64_t n, int32_t t,
        TransRecords rec, unsigned int* __restrict__ n_cand) {
    const int nv = pars.n_variants, nl = pars.n_layers;
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_cand = 0;       // today's candidate list starts empty
    const float* rel_trans = PF(P, rel_trans); const float* rel_sus = PF(P, rel_sus);
    const uint8_t* infectious = PB(P, infectious); const uint8_t* susceptible = PB(P, susceptible);
    const uint8_t* symptomatic = PB(P, symptomatic); const uint8_t* isolated = PB(P, isolated); const uint8_t* quarantined = PB(P, quarantined);
    const float* inf_var = PF(P, infectious_variant);
    const float* d_inf = PF(P, date_infectious); const float* d_rec = PF(P, date_recovered); const float* d_dead = PF(P, date_dead);
    const float* sus_imm = PF(P, sus_imm);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        bool inf = infectious[i] != 0, sus = susceptible[i] != 0;
        int var = 0;
        float rt = 0.0f, vl = 0.0f;
        if (inf) {
            float fv = inf_var[i];
            var = (int)fv;
            if (!(var >= 0 && var < nv)) { inf = false; var = 0; }     // infectious_variant == v never true (sim.py:629)
        }
        bool symp = false, iso = false;
        bool quar = (inf || sus) ? quarantined[i] != 0 : false;
        if (inf) {
            rt = rel_trans[i];
            vl = viral_load(t, d_inf[i], d_rec[i], d_dead[i], pars.frac_time, pars.load_ratio, pars.high_cap);
            symp = symptomatic[i] != 0;
            iso = isolated[i] != 0;
        }
        float rs = sus ? rel_sus[i] : 0.0f;
        if (nv > 1) rec.ivar[i] = (uint8_t)var;
        for (int l = 0; l < nl; ++l) {
            float2 o;
            o.x = inf ? rel_trans_layer(rt, true, symp, iso, quar, pars.asymp_factor, pars.iso_factor[l], pars.quar_factor[l],
                                        pars.beta_layer[l], vl) : 0.0f;
            o.y = sus ? rel_sus_layer(rs, true, quar, pars.quar_factor[l], sus_imm[i]) : 0.0f;
            rec.ts[(int64_t)l * n + i] = o;
            for (int v = 1; v < nv; ++v)
                rec.sus_extra[((int64_t)l * (nv - 1) + (v - 1)) * n + i] =
                    sus ? rel_sus_layer(rs, true, quar, pars.quar_factor[l], sus_imm[(int64_t)v * n + i]) : 0.0f;
        }
    }
}

// ================================================================================================
// update_nab + stock counts + population means
// ================================================================================================
constexpr int kNStocks = 13;
__global__ void __launch_bounds__(kThreads) nab_count_kernel(PeoplePtrs P, const __grid_constant__ cvb_pars pars, int64_t n, int32_t t,
        const double* __restrict__ nab_kin, int64_t nab_kin_len, unsigned long long* __restrict__ counters,
        unsigned long long* __restrict__ vcounters, double* __restrict__ partial) {
    __shared__ int s_cnt[kNStocks + 1 + 2 * CVB_MAX_VARIANTS];
    __shared__ double s_sum[3][kThreads / 32];
    const int nv = pars.n_variants;
    const int NK = kNStocks + 1 + 2 * CVB_MAX_VARIANTS;
    if (threadIdx.x < NK) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    int c[kNStocks + 1];
#pragma unroll
    for (int k = 0; k < kNStocks + 1; ++k) c[k] = 0;
    int cv[2 * CVB_MAX_VARIANTS];
#pragma unroll
    for (int k = 0; k < 2 * CVB_MAX_VARIANTS; ++k) cv[k] = 0;
    double sum_nab = 0.0, sum_sus = 0.0, sum_symp = 0.0;

    const uint8_t* st[kNStocks] = {PB(P, susceptible), PB(P, exposed), PB(P, infectious), PB(P, symptomatic), PB(P, severe), PB(P, critical),
                                   PB(P, recovered), PB(P, dead), PB(P, diagnosed), PB(P, known_dead), PB(P, quarantined), PB(P, isolated),
                                   PB(P, vaccinated)};
    const uint8_t* exp_by_var = PB(P, exposed_by_variant); const uint8_t* inf_by_var = PB(P, infectious_by_variant);
    float* nab = PF(P, nab); const float* peak = PF(P, peak_nab); const int32_t* t_event = PI(P, t_nab_event);
    const float* sus_imm = PF(P, sus_imm); const float* symp_imm = PF(P, symp_imm);

    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        bool is_dead = false;
#pragma unroll
        for (int k = 0; k < kNStocks; ++k) { bool b = st[k][i] != 0; c[k] += b; if (k == 7) is_dead = b; }
        c[kNStocks] += !is_dead;
#pragma unroll
        for (int v = 0; v < CVB_MAX_VARIANTS; ++v)
            if (v < nv) { cv[2 * v] += exp_by_var[(int64_t)v * n + i] != 0; cv[2 * v + 1] += inf_by_var[(int64_t)v * n + i] != 0; }
        float nab_i = nab[i];
        if (pars.use_waning) {
            float pk = peak[i];
            if (pk != 0.0f) {                                    // has_nabs = true(peak_nab)  (sim.py:666-669)
                int64_t dt = (int64_t)t - (int64_t)t_event[i];
                if (dt < 0) dt += nab_kin_len;                   // NumPy negative index wraps
                double kin = (dt >= 0 && dt < nab_kin_len) ? nab_kin[dt] : 0.0;
                nab_i = nab_step(nab_i, pk, kin);
                nab[i] = nab_i;
            }
        }
        if (!is_dead) sum_nab += (double)nab_i;
        for (int v = 0; v < nv; ++v) { sum_sus += (double)sus_imm[(int64_t)v * n + i]; sum_symp += (double)symp_imm[(int64_t)v * n + i]; }
    }
#pragma unroll
    for (int k = 0; k < kNStocks + 1; ++k) {
        int w = __reduce_add_sync(0xFFFFFFFFu, c[k]);
        if (lane_id() == 0 && w) atomicAdd(&s_cnt[k], w);
    }
#pragma unroll
    for (int k = 0; k < 2 * CVB_MAX_VARIANTS; ++k) {
        int w = __reduce_add_sync(0xFFFFFFFFu, cv[k]);
        if (lane_id() == 0 && w) atomicAdd(&s_cnt[kNStocks + 1 + k], w);
    }
    // deterministic float64 sums: fixed-shape shuffle tree per warp, then per CTA, then one partial per CTA
    double sums[3] = {sum_nab, sum_sus, sum_symp};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        double v = sums[q];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
        if (lane_id() == 0) s_sum[q][warp_id()] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) v += s_sum[threadIdx.x][w];
        partial[(int64_t)blockIdx.x * 3 + threadIdx.x] = v;
    }
    if (threadIdx.x < NK && s_cnt[threadIdx.x]) {
        int k = threadIdx.x;
        unsigned long long v = (unsigned long long)s_cnt[k];
        unsigned long long* row = counters + (int64_t)t * CVB_N_COUNTERS;
        if (k < kNStocks) atomicAdd(row + CVB_C_n_susceptible + k, v);
        else if (k == kNStocks) atomicAdd(row + CVB_C_n_alive_agents, v);
        else {
            int q = k - kNStocks - 1, var = q >> 1;
            if (var < nv) atomicAdd(vcounters + ((int64_t)t * nv + var) * CVB_N_VCOUNTERS +
                                    ((q & 1) ? CVB_VC_n_infectious_by_variant : CVB_VC_n_exposed_by_variant), v);
        }
    }
}

__global__ void finish_sums_kernel(const double* __restrict__ partial, int n_blocks, double* __restrict__ sums_row) {
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int b = 0; b < n_blocks; ++b) v += partial[(int64_t)b * 3 + threadIdx.x];
        sums_row[threadIdx.x] = v;
    }
}

}  // namespace cvb

using namespace cvb;

static int require_ready(cvb_sim* s, const char* who, bool need_results = true) {
    CVB_REQUIRE(s, "%s: NULL handle", who);
    CVB_REQUIRE(s->pars_set, "%s: cvb_set_pars has not been called", who);
    for (int f = 0; f < CVB_N_FIELDS; ++f) CVB_REQUIRE(s->people.f[f] != nullptr, "%s: people field %d is not bound", who, f);
    if (need_results) CVB_REQUIRE(s->res.counters && s->res.vcounters && s->res.sums, "%s: result tables are not bound", who);
    return 0;
}

extern "C" {

int cvb_update_states_pre(cvb_sim* s, int32_t t, cvb_stream st) {
    if (require_ready(s, "cvb_update_states_pre")) return 1;
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_update_states_pre: day %d outside [0,%d)", t, s->npts);
    states_pre_kernel<<<grid_for(s->n), kThreads, 0, (cudaStream_t)st>>>(s->people, s->pars, s->n, t, s->res.counters, s->res.vcounters, s->beds);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_schedule_quarantine(cvb_sim* s, const int32_t* inds, int64_t n, int32_t start_day, float end_day, cvb_stream st) {
    CVB_REQUIRE(s, "cvb_schedule_quarantine: NULL handle");
    if (n == 0) return 0;
    CVB_REQUIRE(inds, "cvb_schedule_quarantine: NULL index array");
    CVB_REQUIRE(end_day >= 0.0f, "cvb_schedule_quarantine: end day must be non-negative");
    int slot = ((start_day % s->quar_horizon) + s->quar_horizon) % s->quar_horizon;
    schedule_quar_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)st>>>(inds, n, (int*)(s->quar_ring + (int64_t)slot * s->n), end_day, s->n);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_update_states_post(cvb_sim* s, int32_t t, cvb_stream st) {
    if (require_ready(s, "cvb_update_states_post")) return 1;
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_update_states_post: day %d outside [0,%d)", t, s->npts);
    int slot = t % s->quar_horizon;
    states_post_kernel<<<grid_for(s->n), kThreads, 0, (cudaStream_t)st>>>(s->people, s->n, t, s->quar_ring + (int64_t)slot * s->n, s->res.counters);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_prepare_transmission(cvb_sim* s, int32_t t, cvb_stream st) {
    if (require_ready(s, "cvb_prepare_transmission", false)) return 1;
    int nl = s->pars.n_layers;
    CVB_REQUIRE(nl >= 1, "cvb_prepare_transmission: no contact layers");
    if (s->rec_layers < nl) {
        cudaFree(s->rec.ts); cudaFree(s->rec.sus_extra); cudaFree(s->rec.ivar);
        s->rec.ts = nullptr; s->rec.sus_extra = nullptr; s->rec.ivar = nullptr; s->rec_layers = 0;
        CVB_CHECK(cudaMalloc((void**)&s->rec.ts, (size_t)nl * s->n * sizeof(float2)));
        if (s->nv > 1) {
            CVB_CHECK(cudaMalloc((void**)&s->rec.sus_extra, (size_t)nl * (s->nv - 1) * s->n * sizeof(float)));
            CVB_CHECK(cudaMalloc((void**)&s->rec.ivar, (size_t)s->n));
        }
        s->rec_layers = nl;
    }
    prepare_kernel<<<grid_for(s->n), kThreads, 0, (cudaStream_t)st>>>(s->people, s->pars, s->n, t, s->rec, s->n_cand);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_update_nab_count(cvb_sim* s, int32_t t, cvb_stream st) {
    if (require_ready(s, "cvb_update_nab_count")) return 1;
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_update_nab_count: day %d outside [0,%d)", t, s->npts);
    CVB_REQUIRE(!s->pars.use_waning || s->nab_kin, "cvb_update_nab_count: NAb kinetics table not set (cvb_set_nab_kin)");
    int grid = grid_for(s->n);
    if (ensure_f64(&s->partial, &s->partial_cap, (int64_t)grid * 3)) return 1;
    nab_count_kernel<<<grid, kThreads, 0, (cudaStream_t)st>>>(s->people, s->pars, s->n, t, s->nab_kin, s->nab_kin_len, s->res.counters,
                                                             s->res.vcounters, s->partial);
    CVB_LAUNCH_CHECK();
    finish_sums_kernel<<<1, 32, 0, (cudaStream_t)st>>>(s->partial, grid, s->res.sums + (int64_t)t * 4);
    CVB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
