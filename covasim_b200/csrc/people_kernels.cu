// Per-agent state-transition kernels over the structure-of-arrays People state:
//   update_states_pre  (+ check_immunity)      reference people.py:164-186, immunity.py:303-350
//   update_states_post                          reference people.py:189-196, 315-366
//   prepare_transmission (viral load + one 16-byte agent record per agent + transmit bitmap / list / codes)   sim.py:602-643
//   update_nab + stock counts + population means                                 immunity.py:205-213, sim.py:652-674
//
// Every array a kernel needs is read at most once and written only where it changes (ncu: DRAM traffic = algorithmic bytes);
// what limits them today is latency, not HBM bandwidth (profiles/r1/README.md).
// Each thread owns FOUR consecutive agents and loads every field it may need up front with 32-bit
// (4 x bool) and 128-bit (4 x float32 / int32) coalesced loads, all independent, so ~20 loads per thread
// are in flight at once (the first version chained conditional byte loads and was latency-bound at 5-16 %
// of DRAM throughput: profiles/r1).  State changes are sparse, so stores stay scalar and conditional.
// Flow / stock counters are reduced per warp (__reduce_add_sync), per CTA (shared atomics) and added to
// the day's counter row with one global atomic per counter per CTA.
#include "cvb_internal.cuh"

namespace cvb {

// ================================================================================================
// update_states_pre
// ================================================================================================
enum { PRE_INFECTIOUS = 0, PRE_SYMPTOMATIC, PRE_SEVERE, PRE_CRITICAL, PRE_RECOVERIES, PRE_DEATHS, PRE_KNOWN_DEATHS,
       PRE_BED_SEVERE, PRE_BED_CRITICAL, PRE_NK };

// check_immunity for one queued agent x variant (immunity.py:303-350): float64 arithmetic, rounded once to float32.
// entry = {agent, nab bits, packed (variant | recovered variant + 1 << 4 | vaccine source << 8 | vaccinated << 12), -}
__device__ __forceinline__ void immunity_eval(const uint4 en, int64_t n, const cvb_pars& pars, float* __restrict__ sus_imm,
                                              float* __restrict__ symp_imm, float* __restrict__ sev_imm) {
    const int64_t i = (int64_t)en.x;
    const float nab = __uint_as_float(en.y);
    const int v = (int)(en.z & 15u), rvi = (int)((en.z >> 4) & 15u) - 1, vsi = (int)((en.z >> 8) & 15u);
    const bool vacc = (en.z >> 12) & 1u;
    const double natural = rvi >= 0 ? (double)pars.immunity[v][rvi] : 0.0;
    const double vaccine = vacc ? pars.vaccine_imm[vsi][v] : 0.0;
    const double enab = dmul((double)nab, fmax(natural, vaccine));
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
    if (enab > 0.0)          // 0**beta == 0 -> protection exactly 0
        calc_ve3(enab, pars.exp_alpha_inf, pars.beta_inf, pars.exp_alpha_symp_inf, pars.beta_symp_inf,
                 pars.exp_alpha_sev_symp, pars.beta_sev_symp, s0, s1, s2);
    sus_imm[(int64_t)v * n + i] = s0;
    symp_imm[(int64_t)v * n + i] = s1;
    sev_imm[(int64_t)v * n + i] = s2;
}

constexpr int kImmQueueCap = 64;            // < 32 left over + up to 32 appended per (agent slot, variant) step

__global__ void __launch_bounds__(kThreads) states_pre_kernel(PeoplePtrs P, const __grid_constant__ cvb_pars pars, int64_t n, int32_t t, bool vec,
        unsigned long long* __restrict__ counters, unsigned long long* __restrict__ vcounters, unsigned long long* __restrict__ beds) {
    __shared__ int s_cnt[PRE_NK + CVB_MAX_VARIANTS];
    __shared__ uint4 s_queue[(kThreads / 32) * kImmQueueCap];
    uint4* q_imm = s_queue + warp_id() * kImmQueueCap;
    int qn = 0;                                                     // warp-uniform queue length
    const unsigned lt_mask = (1u << lane_id()) - 1u;
    const int nv = pars.n_variants;
    const bool waning = pars.use_waning != 0;
    const bool vaxpars = pars.has_vaccine_pars != 0;
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
    uint8_t* quarantined = PB(P, quarantined); const uint8_t* vaccinated = PB(P, vaccinated);
    uint8_t* exp_by_var = PB(P, exposed_by_variant); uint8_t* inf_by_var = PB(P, infectious_by_variant);
    float* exp_var = PF(P, exposed_variant); float* inf_var = PF(P, infectious_variant); float* rec_var = PF(P, recovered_variant);
    const float* d_inf = PF(P, date_infectious); const float* d_symp = PF(P, date_symptomatic); const float* d_sev = PF(P, date_severe);
    const float* d_crit = PF(P, date_critical); const float* d_rec = PF(P, date_recovered); const float* d_dead = PF(P, date_dead);
    const float* d_end_iso = PF(P, date_end_isolation);
    const float* nab = PF(P, nab); const int32_t* vsrc = PI(P, vaccine_source);
    float* sus_imm = PF(P, sus_imm); float* symp_imm = PF(P, symp_imm); float* sev_imm = PF(P, sev_imm);
    const float qnan = nanf32();

    // the loop is warp-uniform (whole warps stay in it: the immunity queue is filled with ballots); agents past n are inert
    const int64_t n_groups = (n + kAPT - 1) / kAPT;
    const int64_t n_groups_pad = (n_groups + 31) / 32 * 32;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_groups_pad; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = g * kAPT;
        // ---- all loads up front (independent, coalesced) ----
        const uint32_t w_exp = load4b(exposed, i0, n, vec), w_inf = load4b(infectious, i0, n, vec), w_symp = load4b(symptomatic, i0, n, vec);
        const uint32_t w_sev = load4b(severe, i0, n, vec), w_crit = load4b(critical, i0, n, vec), w_rec = load4b(recovered, i0, n, vec);
        const uint32_t w_dead = load4b(dead, i0, n, vec), w_diag = load4b(diagnosed, i0, n, vec), w_iso = load4b(isolated, i0, n, vec);
        float di[4], ds[4], dv[4], dc[4], dr[4], dd[4], dei[4];
        load4(d_rec, i0, n, vec, qnan, dr);
        load4(d_end_iso, i0, n, vec, qnan, dei);
        float nb[4] = {0.0f, 0.0f, 0.0f, 0.0f}, rv[4] = {qnan, qnan, qnan, qnan};
        uint32_t w_vacc = 0;
        int32_t vs[4] = {0, 0, 0, 0};
        if (waning) {
            load4(nab, i0, n, vec, 0.0f, nb);
            load4(rec_var, i0, n, vec, qnan, rv);
            if (vaxpars) { w_vacc = load4b(vaccinated, i0, n, vec); load4(vsrc, i0, n, vec, vs); }
        }
        float ev[4] = {qnan, qnan, qnan, qnan};
        if (w_exp) {                                               // dates only matter for exposed agents
            load4(d_inf, i0, n, vec, qnan, di); load4(d_symp, i0, n, vec, qnan, ds); load4(d_sev, i0, n, vec, qnan, dv);
            load4(d_crit, i0, n, vec, qnan, dc); load4(d_dead, i0, n, vec, qnan, dd); load4(exp_var, i0, n, vec, qnan, ev);
        }
        float rec_var_now[4];
#pragma unroll
        for (int k = 0; k < kAPT; ++k) {
            const int64_t i = i0 + k;
            rec_var_now[k] = rv[k];
            if (i >= n) continue;
            const bool was_exposed = flag(w_exp, k);               // is_exp is taken once, before any transition (people.py:169)
            bool sev_now = flag(w_sev, k), crit_now = flag(w_crit, k);
            bool diag_now = flag(w_diag, k);
            if (was_exposed) {
                // infectious (people.py:222-232)
                if (!flag(w_inf, k) && due(di[k], t)) {
                    infectious[i] = 1;
                    inf_var[i] = ev[k];
                    const int v = (int)ev[k];
                    if (v >= 0 && v < nv) {
                        inf_by_var[(int64_t)v * n + i] = 1;
#pragma unroll
                        for (int q = 0; q < CVB_MAX_VARIANTS; ++q) cv[q] += (q == v);
                    }
                    ++c[PRE_INFECTIOUS];
                }
                // symptomatic / severe / critical (people.py:235-253)
                if (!flag(w_symp, k) && due(ds[k], t)) { symptomatic[i] = 1; ++c[PRE_SYMPTOMATIC]; }
                if (!sev_now && due(dv[k], t)) { severe[i] = 1; sev_now = true; ++c[PRE_SEVERE]; }
                if (!crit_now && due(dc[k], t)) { critical[i] = 1; crit_now = true; ++c[PRE_CRITICAL]; }
                // recovery (people.py:256-291)
                if (!flag(w_rec, k) && due(dr[k], t)) {
                    exposed[i] = 0; infectious[i] = 0; symptomatic[i] = 0; severe[i] = 0; critical[i] = 0;
                    sev_now = false; crit_now = false;
                    recovered[i] = 1;
                    rec_var[i] = ev[k]; rec_var_now[k] = ev[k];
                    inf_var[i] = qnan;
                    exp_var[i] = qnan;
                    for (int v = 0; v < nv; ++v) { exp_by_var[(int64_t)v * n + i] = 0; inf_by_var[(int64_t)v * n + i] = 0; }
                    if (waning) { susceptible[i] = 1; diagnosed[i] = 0; diag_now = false; }
                    ++c[PRE_RECOVERIES];
                }
            }
            // leave isolation (people.py:368-374) -- every agent
            if (flag(w_iso, k) && due(dei[k], t)) isolated[i] = 0;
            if (was_exposed) {
                // death (people.py:294-312)
                if (!flag(w_dead, k) && due(dd[k], t)) {
                    dead[i] = 1;
                    if (diag_now) { known_dead[i] = 1; ++c[PRE_KNOWN_DEATHS]; }
                    susceptible[i] = 0; exposed[i] = 0; infectious[i] = 0; symptomatic[i] = 0; severe[i] = 0; critical[i] = 0;
                    sev_now = false; crit_now = false;
                    known_contact[i] = 0; quarantined[i] = 0; recovered[i] = 0;
                    inf_var[i] = qnan; exp_var[i] = qnan; rec_var[i] = qnan; rec_var_now[k] = qnan;
                    ++c[PRE_DEATHS];
                }
            }
            c[PRE_BED_SEVERE] += sev_now;
            c[PRE_BED_CRITICAL] += crit_now;
        }

        // check_immunity (immunity.py:303-350).  The float64 log / exp work is needed only for agents with neutralising
        // antibodies and a non-zero immunity factor.  Evaluating it in the per-agent loop above would make every warp that
        // holds ONE such agent among its 128 run the ~300-instruction path (four times); instead those agents are appended
        // to the warp's shared-memory queue (ballot compaction) and evaluated 32 at a time by full warps.
        if (waning) {
#pragma unroll
            for (int k = 0; k < kAPT; ++k) {
                const int64_t i = i0 + k;
                const bool valid = i < n;
                const bool was_inf = due(dr[k], t);
                const int rvi = was_inf ? (int)rec_var_now[k] : -1;
                const bool rv_ok = rvi >= 0 && rvi < nv;
                const int vsi = vs[k];
                const bool vacc = vaxpars && flag(w_vacc, k) && vsi >= 0 && vsi < CVB_MAX_VACCINES;
                const bool heavy = valid && nb[k] > 0.0f && (rv_ok || vacc);
                const unsigned packed = ((unsigned)(rv_ok ? rvi + 1 : 0) << 4) | ((unsigned)(vacc ? vsi : 0) << 8) | ((unsigned)vacc << 12);
                for (int v = 0; v < nv; ++v) {
                    if (valid && !heavy) {                         // no NAbs or no immunity source: protection exactly 0
                        sus_imm[(int64_t)v * n + i] = 0.0f;
                        symp_imm[(int64_t)v * n + i] = 0.0f;
                        sev_imm[(int64_t)v * n + i] = 0.0f;
                    }
                    const unsigned m = __ballot_sync(0xFFFFFFFFu, heavy);
                    if (heavy) q_imm[qn + __popc(m & lt_mask)] = make_uint4((unsigned)i, __float_as_uint(nb[k]), packed | (unsigned)v, 0u);
                    qn += __popc(m);
                    __syncwarp();
                    if (qn >= 32) {
                        qn -= 32;
                        const uint4 en = q_imm[qn + lane_id()];
                        __syncwarp();
                        immunity_eval(en, n, pars, sus_imm, symp_imm, sev_imm);
                    }
                }
            }
        }
    }
    if (waning && lane_id() < qn) immunity_eval(q_imm[lane_id()], n, pars, sus_imm, symp_imm, sev_imm);
    reduce_counters(c, s_cnt);
    reduce_counters(cv, s_cnt + PRE_NK);
    __syncthreads();
    if (threadIdx.x < PRE_NK + CVB_MAX_VARIANTS) {
        const int v = s_cnt[threadIdx.x];
        if (v) {
            unsigned long long* row = counters + (int64_t)t * CVB_N_COUNTERS;
            const unsigned long long uv = (unsigned long long)v;
            switch (threadIdx.x) {
                case PRE_INFECTIOUS:   atomicAdd(row + CVB_C_new_infectious, uv); break;
                case PRE_SYMPTOMATIC:  atomicAdd(row + CVB_C_new_symptomatic, uv); break;
                case PRE_SEVERE:       atomicAdd(row + CVB_C_new_severe, uv); break;
                case PRE_CRITICAL:     atomicAdd(row + CVB_C_new_critical, uv); break;
                case PRE_RECOVERIES:   atomicAdd(row + CVB_C_new_recoveries, uv); break;
                case PRE_DEATHS:       atomicAdd(row + CVB_C_new_deaths, uv); break;
                case PRE_KNOWN_DEATHS: atomicAdd(row + CVB_C_new_known_deaths, uv); break;
                case PRE_BED_SEVERE:   atomicAdd(beds + (int64_t)t * 2 + 0, uv); break;
                case PRE_BED_CRITICAL: atomicAdd(beds + (int64_t)t * 2 + 1, uv); break;
                default: {
                    const int var = threadIdx.x - PRE_NK;
                    if (var < nv) atomicAdd(vcounters + ((int64_t)t * nv + var) * CVB_N_VCOUNTERS + CVB_VC_new_infectious_by_variant, uv);
                }
            }
        }
    }
}

// ================================================================================================
// update_states_post and/or prepare_transmission (one kernel: they read the same flags back to back)
// ================================================================================================
enum { POST_DIAGNOSES = 0, POST_QUARANTINED, POST_ISOLATED, POST_NK };

template <bool DO_POST, bool DO_PREP>
__global__ void __launch_bounds__(kThreads) post_prepare_kernel(PeoplePtrs P, const __grid_constant__ cvb_pars pars, int64_t n, int32_t t, bool vec,
        float* __restrict__ quar_slot, unsigned long long* __restrict__ counters, TransRecords rec, unsigned int* __restrict__ inf_bits,
        int32_t* __restrict__ trans_list, unsigned int* __restrict__ n_trans, unsigned int* __restrict__ n_cand,
        uint8_t* __restrict__ codes, const float* __restrict__ base_trans, unsigned int* __restrict__ part_flags) {
    __shared__ int s_cnt[POST_NK];
    if (threadIdx.x < POST_NK) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    int c[POST_NK] = {0, 0, 0};
    const int nv = pars.n_variants;
    if (DO_PREP && blockIdx.x == 0 && threadIdx.x == 0) *n_cand = 0;      // today's candidate list starts empty
    uint8_t* diagnosed = PB(P, diagnosed); uint8_t* quarantined = PB(P, quarantined); uint8_t* isolated = PB(P, isolated);
    const uint8_t* dead = PB(P, dead); const uint8_t* recovered = PB(P, recovered);
    float* d_pos = PF(P, date_pos_test); const float* d_diag = PF(P, date_diagnosed); float* d_quar = PF(P, date_quarantined);
    float* d_end_quar = PF(P, date_end_quarantine); float* d_end_iso = PF(P, date_end_isolation); const float* d_rec = PF(P, date_recovered);
    const float* rel_trans = PF(P, rel_trans); const float* rel_sus = PF(P, rel_sus);
    const uint8_t* infectious = PB(P, infectious); const uint8_t* susceptible = PB(P, susceptible); const uint8_t* symptomatic = PB(P, symptomatic);
    const float* inf_var = PF(P, infectious_variant); const float* d_inf = PF(P, date_infectious); const float* d_dead = PF(P, date_dead);
    const float* sus_imm = PF(P, sus_imm);
    const float tf = (float)t;
    const float qnan = nanf32();

    // the loop is warp-aligned: a warp covers 128 consecutive agents = four words of the infectious bitmap
    const int64_t n_groups = (n + kAPT - 1) / kAPT;
    const int64_t n_groups_pad = (n_groups + 31) / 32 * 32;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_groups_pad; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = g * kAPT;
        uint32_t w_diag = 0, w_quar = load4b(quarantined, i0, n, vec), w_iso = load4b(isolated, i0, n, vec);
        uint32_t w_dead = 0, w_rec = 0;
        float dpos[4], ddiag[4], pend[4], endq[4], drec[4];
        load4(d_rec, i0, n, vec, qnan, drec);
        if (DO_POST) {
            w_diag = load4b(diagnosed, i0, n, vec); w_dead = load4b(dead, i0, n, vec); w_rec = load4b(recovered, i0, n, vec);
            load4(d_pos, i0, n, vec, qnan, dpos); load4(d_diag, i0, n, vec, qnan, ddiag);
            load4(quar_slot, i0, n, vec, -1.0f, pend); load4(d_end_quar, i0, n, vec, qnan, endq);
        }
        uint32_t w_inf = 0, w_sus = 0, w_symp = 0;
        float rt4[4], rs4[4], iv4[4], dinf[4], ddead[4], imm0[4];
        if (DO_PREP) {
            w_inf = load4b(infectious, i0, n, vec); w_sus = load4b(susceptible, i0, n, vec); w_symp = load4b(symptomatic, i0, n, vec);
            load4(rel_sus, i0, n, vec, 0.0f, rs4); load4(sus_imm, i0, n, vec, 0.0f, imm0);
            if (w_inf) {
                load4(rel_trans, i0, n, vec, 0.0f, rt4); load4(inf_var, i0, n, vec, qnan, iv4);
                load4(d_inf, i0, n, vec, qnan, dinf); load4(d_dead, i0, n, vec, qnan, ddead);
            }
        }
        unsigned inf_nibble = 0;
        uint32_t code_word = 0;                                     // partitioned form: the four agents' transmit codes
        float4 out4[4];
#pragma unroll
        for (int k = 0; k < kAPT; ++k) out4[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
        for (int k = 0; k < kAPT; ++k) {
            const int64_t i = i0 + k;
            if (i >= n) break;
            bool quar = flag(w_quar, k), iso = flag(w_iso, k);
            if (DO_POST) {
                // check_diagnosed (people.py:315-332)
                bool diag = flag(w_diag, k);
                if (!diag) {
                    if (due(dpos[k], t)) { d_pos[i] = qnan; ++c[POST_DIAGNOSES]; }
                    if (due(ddiag[k], t)) { diagnosed[i] = 1; diag = true; }
                }
                // check_quar (people.py:335-358): the pending slot holds the max end day requested for today
                float end_q = endq[k];
                if (pend[k] >= 0.0f) {
                    quar_slot[i] = -1.0f;
                    if (quar) {
                        if (pend[k] > end_q) { end_q = pend[k]; d_end_quar[i] = end_q; }          // Python max(old, requested)
                    } else if (!(flag(w_dead, k) || flag(w_rec, k) || diag || iso)) {
                        quarantined[i] = 1; quar = true;
                        d_quar[i] = tf;
                        end_q = pend[k];
                        d_end_quar[i] = end_q;
                        ++c[POST_QUARANTINED];
                    }
                }
                if (quar) {
                    if (ddiag[k] == tf) { end_q = tf; d_end_quar[i] = tf; }
                    if (due(end_q, t)) { quarantined[i] = 0; quar = false; }
                }
                // check_enter_iso (people.py:361-366)
                if (ddiag[k] == tf) {
                    isolated[i] = 1; iso = true;
                    d_end_iso[i] = drec[k];
                    ++c[POST_ISOLATED];
                }
            }
            if (DO_PREP) {
                // prepare_transmission (sim.py:602-643): ONE 16-byte record per agent; the edge pass applies the per-layer
                // factors to the edges it evaluates (cvb_device.cuh:AgentRecord)
                bool inf = flag(w_inf, k);
                const bool sus = flag(w_sus, k);
                int var = 0;
                if (inf) {
                    var = (int)iv4[k];
                    if (!(var >= 0 && var < nv)) { inf = false; var = 0; }      // infectious_variant == v is never true (sim.py:629)
                }
                const bool symp = flag(w_symp, k);
                uint32_t code = quar ? 32u : 0u;                              // the quarantine bit matters for targets too
                float rt = 0.0f;
                if (inf && rt4[k] != 0.0f) {                                   // can transmit (a zero rel_trans never does)
                    rt = rt4[k];
                    const bool early = viral_load_early(t, dinf[k], drec[k], ddead[k], pars.frac_time, pars.high_cap);
                    bool redux = false;
                    if (codes) {                                               // partitioned: other GPUs rebuild rt from its initial value
                        const float base = base_trans[i];
                        redux = rt != base;
                        if (redux && rt != fmul(base, pars.trans_redux)) atomicAdd(part_flags + 1, 1u);
                    }
                    code = transmit_code(var, symp, iso, quar, early, redux);
                    inf_nibble |= 1u << k;
                    if (codes) code_word |= code << (8 * k);
                    else trans_list[warp_append32(n_trans)] = (int32_t)i;    // compact (unordered) list of today's transmitters
                }
                // With one variant, immunity is applied here for agents who are not quarantined (their susceptibility is the
                // same on every layer), so the edge pass does the float64 product only for quarantined targets.
                float s_rec = sus ? rs4[k] : 0.0f, imm_rec = imm0[k];
                if (nv == 1 && !quar) { s_rec = record_sus(s_rec, 0u, 1.0f, imm_rec); imm_rec = 0.0f; }
                out4[k] = make_float4(rt, s_rec, imm_rec, __uint_as_float(code));
                if (rec.ts8) {
                    // the layers the dense streaming pass reads get the reference's per-layer pair as well: that pass
                    // evaluates ~20 % of ALL edges, so the layer factors are cheaper applied once per agent here
                    const float vl = rt != 0.0f ? viral_load_value((code & 64u) != 0, pars.frac_time, pars.load_ratio) : 0.0f;
                    for (int l = 0; l < pars.n_layers; ++l) {
                        if (!((rec.ts8_mask >> l) & 1u)) continue;
                        float2 o;
                        o.x = rt != 0.0f ? rel_trans_layer(rt, true, symp, iso, quar, pars.asymp_factor, pars.iso_factor[l], pars.quar_factor[l],
                                                           pars.beta_layer[l], vl) : 0.0f;
                        o.y = sus ? rel_sus_layer(rs4[k], true, quar, pars.quar_factor[l], imm0[k]) : 0.0f;
                        rec.ts8[(int64_t)l * n + i] = o;
                    }
                }
            }
        }
        if (DO_PREP && rec.rec) {                                  // (not needed when every layer goes through the dense pass)
            if (vec && i0 + 4 <= n) {
#pragma unroll
                for (int k = 0; k < kAPT; ++k) rec.rec[i0 + k] = out4[k];
            } else {
                for (int k = 0; k < kAPT && i0 + k < n; ++k) rec.rec[i0 + k] = out4[k];
            }
        }
        if (DO_PREP && codes) {
            if (vec && i0 + 4 <= n) *reinterpret_cast<uint32_t*>(codes + i0) = code_word;
            else for (int k = 0; k < kAPT && i0 + k < n; ++k) codes[i0 + k] = (uint8_t)(code_word >> (8 * k));
        }
        if (DO_PREP) {
            // bitmap of agents that can transmit in at least one layer: lane L holds bits 4(L&7).. of word (L>>3)
            unsigned word = inf_nibble << (4 * (lane_id() & 7));
            word |= __shfl_xor_sync(0xFFFFFFFFu, word, 1);
            word |= __shfl_xor_sync(0xFFFFFFFFu, word, 2);
            word |= __shfl_xor_sync(0xFFFFFFFFu, word, 4);
            const int64_t widx = (i0 - (int64_t)(lane_id() & 7) * kAPT) / 32;
            if ((lane_id() & 7) == 0 && widx * 32 < n) inf_bits[widx] = word;
        }
    }
    if (DO_POST) {
        reduce_counters(c, s_cnt);
        __syncthreads();
        if (threadIdx.x < POST_NK && s_cnt[threadIdx.x]) {
            unsigned long long* row = counters + (int64_t)t * CVB_N_COUNTERS;
            const int ids[POST_NK] = {CVB_C_new_diagnoses, CVB_C_new_quarantined, CVB_C_new_isolated};
            atomicAdd(row + ids[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
        }
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
// update_nab + stock counts + population means
// ================================================================================================
constexpr int kNStocks = 13;
__global__ void __launch_bounds__(kThreads, 3) nab_count_kernel(PeoplePtrs P, const __grid_constant__ cvb_pars pars, int64_t n, int32_t t, bool vec,
        const double* __restrict__ nab_kin, int64_t nab_kin_len, unsigned long long* __restrict__ counters,
        unsigned long long* __restrict__ vcounters, double* __restrict__ partial, unsigned int* __restrict__ ticket,
        double* __restrict__ sums_row) {
    __shared__ int s_cnt[kNStocks + 1 + 2 * CVB_MAX_VARIANTS];
    __shared__ bool s_last;
    __shared__ double s_sum[3][kThreads / 32];
    const int nv = pars.n_variants;
    const bool waning = pars.use_waning != 0;
    constexpr int NK = kNStocks + 1 + 2 * CVB_MAX_VARIANTS;
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

    const int64_t n_groups = (n + kAPT - 1) / kAPT;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_groups; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = g * kAPT;
        uint32_t w[kNStocks];
#pragma unroll
        for (int k = 0; k < kNStocks; ++k) w[k] = load4b(st[k], i0, n, vec);
        float nb[4], pk[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        int32_t te[4] = {0, 0, 0, 0};
        load4(nab, i0, n, vec, 0.0f, nb);
        if (waning) { load4(peak, i0, n, vec, 0.0f, pk); load4(t_event, i0, n, vec, te); }
#pragma unroll
        for (int k = 0; k < kNStocks; ++k) c[k] += count4(w[k]);
        const int valid = (int)((n - i0) < 4 ? (n - i0) : 4);
        c[kNStocks] += valid - count4(w[7]);                      // alive = not dead
#pragma unroll
        for (int v = 0; v < CVB_MAX_VARIANTS; ++v)
            if (v < nv) {
                cv[2 * v] += count4(load4b(exp_by_var + (int64_t)v * n, i0, n, vec));
                cv[2 * v + 1] += count4(load4b(inf_by_var + (int64_t)v * n, i0, n, vec));
                float a4[4], b4[4];
                load4(sus_imm + (int64_t)v * n, i0, n, vec, 0.0f, a4);
                load4(symp_imm + (int64_t)v * n, i0, n, vec, 0.0f, b4);
#pragma unroll
                for (int k = 0; k < 4; ++k) { sum_sus += (double)a4[k]; sum_symp += (double)b4[k]; }
            }
#pragma unroll
        for (int k = 0; k < kAPT; ++k) {
            const int64_t i = i0 + k;
            if (i >= n) break;
            float nab_i = nb[k];
            if (waning && pk[k] != 0.0f) {                           // has_nabs = true(peak_nab)  (sim.py:666-669)
                int64_t dt = (int64_t)t - (int64_t)te[k];
                if (dt < 0) dt += nab_kin_len;                       // NumPy negative index wraps
                const double kin = (dt >= 0 && dt < nab_kin_len) ? nab_kin[dt] : 0.0;
                nab_i = nab_step(nab_i, pk[k], kin);
                nab[i] = nab_i;
            }
            if (!flag(w[7], k)) sum_nab += (double)nab_i;
        }
    }
    reduce_counters(c, s_cnt);
    reduce_counters(cv, s_cnt + kNStocks + 1);
    // deterministic float64 sums: fixed-shape shuffle tree per warp, fixed order across warps, one partial per CTA
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
        for (int wq = 0; wq < kThreads / 32; ++wq) v += s_sum[threadIdx.x][wq];
        partial[(int64_t)blockIdx.x * 3 + threadIdx.x] = v;
    }
    if (threadIdx.x < NK && s_cnt[threadIdx.x]) {
        const int k = threadIdx.x;
        const unsigned long long v = (unsigned long long)s_cnt[k];
        unsigned long long* row = counters + (int64_t)t * CVB_N_COUNTERS;
        if (k < kNStocks) atomicAdd(row + CVB_C_n_susceptible + k, v);
        else if (k == kNStocks) atomicAdd(row + CVB_C_n_alive_agents, v);
        else {
            const int q = k - kNStocks - 1, var = q >> 1;
            if (var < nv) atomicAdd(vcounters + ((int64_t)t * nv + var) * CVB_N_VCOUNTERS +
                                    ((q & 1) ? CVB_VC_n_infectious_by_variant : CVB_VC_n_exposed_by_variant), v);
        }
    }
    // The last CTA to finish adds up the per-CTA partial sums (no second launch): one warp per sum, lanes stride over the
    // partials in a fixed pattern, then a fixed shuffle tree -- the result does not depend on which CTA came last.
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x < 96) {
        __threadfence();
        const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
        double v = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) v += __ldcg(partial + (int64_t)b * 3 + q);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
        if (lane == 0) sums_row[q] = v;
        if (threadIdx.x == 0) *ticket = 0;                          // ready for the next launch
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

// 128-bit loads need every bound array 16-byte aligned and (for the [n_variants, N] arrays) N % 4 == 0
static bool vector_ok(const cvb_sim* s) {
    if (s->n % 4 != 0) return false;
    uintptr_t all = 0;
    for (int f = 0; f < CVB_N_FIELDS; ++f) all |= (uintptr_t)s->people.f[f];
    return (all & 15) == 0;
}

static int grid_agents(int64_t n) { return grid_for((n + kAPT - 1) / kAPT, kThreads, 148 * 8); }

static int ensure_records(cvb_sim* s) {
    CVB_REQUIRE(s->pars.n_layers >= 1, "prepare_transmission: no contact layers");
    if (!s->rec_store) CVB_CHECK(cudaMalloc((void**)&s->rec_store, (size_t)s->n * sizeof(float4)));
    s->rec.sus_imm = (const float*)s->people.f[CVB_F_sus_imm];
    // layers the dense streaming pass will read today: not covered by an adjacency, not partitioned, non-empty
    uint32_t dense = 0;
    if (!s->partitioned && s->nv == 1)
        for (int l = 0; l < s->pars.n_layers; ++l)
            if (s->layers[l].n_edges > 0 && !((s->adj && ((s->adj_layer_mask >> l) & 1u)))) dense |= 1u << l;
    if (dense && (!s->ts8_store || s->ts8_layers < s->pars.n_layers)) {
        cudaFree(s->ts8_store);
        s->ts8_store = nullptr;
        CVB_CHECK(cudaMalloc((void**)&s->ts8_store, (size_t)s->pars.n_layers * s->n * sizeof(float2)));
        s->ts8_layers = s->pars.n_layers;
    }
    s->rec.ts8 = dense ? s->ts8_store : nullptr;
    s->rec.ts8_mask = dense;
    // the 16-byte agent records feed the adjacency / partitioned passes and the multi-variant dense pass; when every non-empty
    // layer is read through ts8 nobody needs them
    uint32_t nonempty = 0;
    for (int l = 0; l < s->pars.n_layers; ++l) if (s->layers[l].n_edges > 0) nonempty |= 1u << l;
    s->rec.rec = (dense && dense == nonempty) ? nullptr : s->rec_store;
    s->rec_layers = s->pars.n_layers;
    return 0;
}

template <bool DO_POST, bool DO_PREP>
static int launch_post_prepare(cvb_sim* s, int32_t t, cudaStream_t st) {
    const int slot = t % s->quar_horizon;
    if (DO_PREP) CVB_CHECK(cudaMemsetAsync(s->n_trans, 0, sizeof(unsigned int), st));
    post_prepare_kernel<DO_POST, DO_PREP><<<grid_agents(s->n), kThreads, 0, st>>>(s->people, s->pars, s->n, t, vector_ok(s),
        s->quar_ring + (int64_t)slot * s->n, s->res.counters, s->rec, s->inf_bits, s->trans_list, s->n_trans, s->n_cand,
        s->partitioned ? s->codes_local : nullptr, s->partitioned ? s->rel_trans_global + s->id0 : nullptr, s->part_flags);
    CVB_LAUNCH_CHECK();
    return 0;
}

extern "C" {

int cvb_update_states_pre(cvb_sim* s, int32_t t, cvb_stream st) {
    if (s) cvb::state_touched(s);
    if (require_ready(s, "cvb_update_states_pre")) return 1;
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_update_states_pre: day %d outside [0,%d)", t, s->npts);
    s->last_t = t;
    states_pre_kernel<<<grid_agents(s->n), kThreads, 0, (cudaStream_t)st>>>(s->people, s->pars, s->n, t, vector_ok(s), s->res.counters,
                                                                          s->res.vcounters, s->beds);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_schedule_quarantine(cvb_sim* s, const int32_t* inds, int64_t n, int32_t start_day, float end_day, cvb_stream st) {
    if (s) cvb::state_touched(s);
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
    if (s) cvb::state_touched(s);
    if (require_ready(s, "cvb_update_states_post")) return 1;
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_update_states_post: day %d outside [0,%d)", t, s->npts);
    return launch_post_prepare<true, false>(s, t, (cudaStream_t)st);
}

int cvb_prepare_transmission(cvb_sim* s, int32_t t, cvb_stream st) {
    if (require_ready(s, "cvb_prepare_transmission", false)) return 1;
    if (ensure_records(s)) return 1;
    return launch_post_prepare<false, true>(s, t, (cudaStream_t)st);
}

int cvb_post_and_prepare(cvb_sim* s, int32_t t, cvb_stream st) {
    if (s) cvb::state_touched(s);
    if (require_ready(s, "cvb_post_and_prepare")) return 1;
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_post_and_prepare: day %d outside [0,%d)", t, s->npts);
    if (ensure_records(s)) return 1;
    return launch_post_prepare<true, true>(s, t, (cudaStream_t)st);
}

int cvb_update_nab_count(cvb_sim* s, int32_t t, cvb_stream st) {
    if (s) cvb::state_touched(s);
    if (require_ready(s, "cvb_update_nab_count")) return 1;
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_update_nab_count: day %d outside [0,%d)", t, s->npts);
    CVB_REQUIRE(!s->pars.use_waning || s->nab_kin, "cvb_update_nab_count: NAb kinetics table not set (cvb_set_nab_kin)");
    int grid = grid_for((s->n + kAPT - 1) / kAPT, kThreads, 148 * 4);
    if (ensure_f64(&s->partial, &s->partial_cap, (int64_t)grid * 3)) return 1;
    nab_count_kernel<<<grid, kThreads, 0, (cudaStream_t)st>>>(s->people, s->pars, s->n, t, vector_ok(s), s->nab_kin, s->nab_kin_len,
                                                             s->res.counters, s->res.vcounters, s->partial,
                                                             reinterpret_cast<unsigned int*>(s->dev_scalars + 8), s->res.sums + (int64_t)t * 4);
    CVB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
