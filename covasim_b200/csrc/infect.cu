// People.infect (reference people.py:435-586) + update_peak_nab (immunity.py:138-202), native-RNG form.
//
// Sixteen lanes per newly infected agent.  Every random draw is a pure function of
// (seed, P_INFECT, day, agent, slot), so the result does not depend on the order in which the edge
// pass discovered the targets.  Slots: 0 exp2inf, 1 symptomatic?, 2 asym2rec|inf2sym, 3 severe?,
// 4 mild2rec|sym2sev, 5 critical?, 6 sev2rec|sev2crit, 7 dies?, 8 crit2rec|crit2die, 9 nab_init.
#include "cvb_internal.cuh"

namespace cvb {

int build_layer_table(cvb_sim* s, LayerTable& L, int tile_edges, uint32_t skip_mask);

enum { INF_INFECTIONS = 0, INF_REINFECTIONS, INF_NK };

struct InfectArgs {
    uint64_t seed;
    int64_t n;
    int32_t t;
    int32_t count_flows;     // 0 for seed infections at initialisation (reference sim.py:505-532: flows are discarded)
    int32_t list_layer_code; // layer code logged for list-mode keys (CVB_LAYER_SEED / CVB_LAYER_IMPORT)
    int32_t hosp_max, icu_max; // 0 / 1 as given by the caller, -1 = evaluate from today's severe / critical counts (sim.py:579-580)
    int64_t id0;             // global id of local agent 0 (agent-partitioned runs; 0 otherwise): Philox keys and the log use global ids
    const int32_t* hit_src;  // partitioned edge pass: cand[] lists every successful transmission, with its source and key; the
    const unsigned long long* hit_key;   // entry whose key equals infect_key[target] is the winner (NULL: cand[] lists unique targets)
    int64_t hit_cap;
    const long long* adj_ptr; const uint4* adj; uint32_t adj_mask;   // fused pipeline: the source of a transmission over a layer the
                                             // adjacency covers is looked up in the TARGET's row (the raw edge lists are not read)
    const unsigned long long* log_base;      // winners mode: the log length before this kernel; candidate j's entry goes to *log_base + j
                                             // (no atomics; a candidate that is not infected leaves a tombstone, target = -1)
    const double* tape;                      // verification (cvb_infect_list_taped): draws given by the caller, [list position][16 slots]
    const unsigned long long* beds_direct;   // fused pipeline: {n_severe, n_critical} of today's counter row (running totals); NULL: beds[t]
    unsigned long long* vcounters_row;       // today's by-variant counter row (stock differences)
};

#ifndef CVB_INFECT_MINB
#define CVB_INFECT_MINB 1
#endif
__global__ void __launch_bounds__(kThreads, CVB_INFECT_MINB) infect_kernel(PeoplePtrs P, const __grid_constant__ cvb_pars pars, const __grid_constant__ LayerTable L,
        const __grid_constant__ InfectArgs ia, const int32_t* __restrict__ cand, const unsigned int* __restrict__ n_cand_ptr,
        unsigned long long* __restrict__ infect_key, const unsigned long long* __restrict__ beds, ResultPtrs res, LogPtrs log,
        uint32_t* __restrict__ S /* packed state words of the fused day pipeline (day_fused.cu), or NULL */) {
    __shared__ int s_cnt[INF_NK + 3 * CVB_MAX_VARIANTS];
    __shared__ int s_delta[kStockSlots];
    pdl_trigger();
    pdl_wait();                                                // the candidates come from the edge pass before this kernel
    // the grid is sized for a large outbreak: CTAs whose first pair of candidates does not exist have nothing to do at all
    if ((unsigned long long)blockIdx.x * (blockDim.x >> 5) * 2 >= (unsigned long long)__ldcg(n_cand_ptr)) return;
    const int NK = INF_NK + 3 * CVB_MAX_VARIANTS;
    if (threadIdx.x < NK) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x < kStockSlots) s_delta[threadIdx.x] = 0;
    __syncthreads();
    int c[INF_NK] = {0, 0};
    int cv[3 * CVB_MAX_VARIANTS];
#pragma unroll
    for (int k = 0; k < 3 * CVB_MAX_VARIANTS; ++k) cv[k] = 0;

    const int64_t n = ia.n;
    const int32_t t = ia.t;
    unsigned int n_cand = __ldcg(n_cand_ptr);                  // (L2 loads: the kernel may have been resident while the edge pass wrote them)
    if (ia.hit_key && (int64_t)n_cand > ia.hit_cap) n_cand = (unsigned int)ia.hit_cap;
    const unsigned long long* bd = ia.beds_direct ? ia.beds_direct : beds + (int64_t)t * 2;
    const bool hosp_max = ia.hosp_max >= 0 ? ia.hosp_max != 0 : (pars.n_beds_hosp >= 0 && (long long)__ldcg(bd) > pars.n_beds_hosp);
    const bool icu_max = ia.icu_max >= 0 ? ia.icu_max != 0 : (pars.n_beds_icu >= 0 && (long long)__ldcg(bd + 1) > pars.n_beds_icu);
    const float tf = (float)t;

    // Sixteen lanes per newly infected agent: lane s of the half-warp computes the Philox draw of slot s (its uniform, and
    // for the duration / NAb slots the normal deviate pushed through BOTH distributions the slot may feed), so the ~10
    // float64 Box-Muller + exp chains of one agent run side by side instead of one after the other; lane 0 of the
    // half-warp then walks the prognosis tree with the finished values.  (One thread per agent spent its time on
    // instruction-cache misses and dependent float64 latency: profiles/r1.)
    const int lane = lane_id(), half = lane >> 4, slot = lane & 15, base = half << 4;
    const unsigned int warps_total = (gridDim.x * blockDim.x) >> 5;
    for (unsigned int jw = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 2; jw < n_cand; jw += warps_total * 2) {
        const unsigned int j = jw + half;
        const bool valid = j < n_cand;
        const int64_t i = valid ? __ldcg(cand + j) : 0;
        const int64_t gi = i + ia.id0;                             // global id: Philox index and logged target
        // the leader's per-agent inputs are requested HERE, as soon as the agent is known, so that they travel while the winning key,
        // the adjacency row and the draws are worked out (they used to start a second chain of memory round trips after all that)
        const bool lead = valid && slot == 0;
        bool is_sus = false;
        float in_peak = 0.0f, in_nab = 0.0f, in_rel_trans = 0.0f, in_drec = 0.0f, in_symp_prob = 0.0f, in_sev_prob = 0.0f, in_crit_prob = 0.0f, in_death_prob = 0.0f;
        int in_nbreak = 0, in_ninf = 0;
        uint32_t so = 0u;
        unsigned long long log_base_v = 0ull;
        if (lead) {
            is_sus = PB(P, susceptible)[i] != 0;
            in_peak = PF(P, peak_nab)[i]; in_nab = PF(P, nab)[i]; in_rel_trans = PF(P, rel_trans)[i];
            in_nbreak = PI(P, n_breakthroughs)[i]; in_ninf = PI(P, n_infections)[i];
            in_drec = PF(P, date_recovered)[i];
            in_symp_prob = PF(P, symp_prob)[i]; in_sev_prob = PF(P, severe_prob)[i]; in_crit_prob = PF(P, crit_prob)[i]; in_death_prob = PF(P, death_prob)[i];
            if (S) so = S[i];
            if (ia.log_base) log_base_v = __ldcg(ia.log_base);
        }
        // who infected i: the winning key names (layer, edge, direction); i's adjacency row holds that edge with the other
        // direction bit and the source as the neighbour.  The 16 lanes scan the row while the draws below are computed.
        int32_t src_adj = -1;
        if (ia.adj_ptr && valid) {
            const unsigned long long k0 = __ldcg(infect_key + i);
            const unsigned lf = (unsigned)(k0 >> 48) & 0xFFu;
            if (k0 != kEmptyKey && lf != 0xFFu && ((ia.adj_mask >> lf) & 1u)) {
                const unsigned e32 = (unsigned)(k0 & 0xFFFFFFFFull), want = (lf << 1) | (1u - (unsigned)((k0 >> 40) & 1ull));
                const long long beg = ia.adj_ptr[i], end = ia.adj_ptr[i + 1];
                for (long long off = beg + slot; off < end; off += 16) {
                    const uint4 en = __ldg(ia.adj + off);
                    if (en.y == e32 && en.z == want) src_adj = (int32_t)en.x;
                }
            }
        }
#pragma unroll
        for (int d = 8; d > 0; d >>= 1) src_adj = max(src_adj, __shfl_xor_sync(0xFFFFFFFFu, src_adj, d));
        const u32x4 words = keyed_words(ia.seed, P_INFECT, 0, t, gi, (uint32_t)slot);
        double u_mine = u53(words.x, words.y);
        double d0 = 0.0, d1 = 0.0;
        if (ia.tape) {
            // taped draws: the value every prognosis step consumed in a recorded run of the reference (the uniform of a Bernoulli
            // step, the finished duration / NAb sample of the others), indexed by the agent's position in the list
            const unsigned long long kt = valid ? __ldcg(infect_key + i) : kEmptyKey;
            const double v = kt != kEmptyKey ? ia.tape[(int64_t)(kt & 0xFFFFFFFFFFull) * 16 + slot] : 0.0;
            u_mine = v; d0 = v; d1 = v;
        } else if (slot <= 9 && !(slot & 1 && slot != 9)) {                // slots 0, 2, 4, 6, 8 (durations) and 9 (initial NAb level)
            // ONE copy of the Box-Muller / exp code for every slot (the kernel is launched cold every day with a handful
            // of warps: its time is instruction fetch, so code size matters more than the selects below)
            const double z = normal_from_words(words);
            const int ka = slot == 0 ? CVB_DUR_exp2inf : slot == 2 ? CVB_DUR_asym2rec : slot == 4 ? CVB_DUR_mild2rec : slot == 6 ? CVB_DUR_sev2rec : CVB_DUR_crit2rec;
            const int kb = slot == 2 ? CVB_DUR_inf2sym : slot == 4 ? CVB_DUR_sym2sev : slot == 6 ? CVB_DUR_sev2crit : CVB_DUR_crit2die;
            const cvb_dist da = slot == 9 ? pars.nab_init : pars.dur[ka];
            cvb_dist db = pars.dur[kb];
            if (slot == 0 || slot == 9) db.kind = CVB_DIST_ZERO;
#pragma unroll 1
            for (int alt = 0; alt < 2; ++alt) {
                const double d = dist_from_normal(alt ? db : da, z);
                if (alt) d1 = d; else d0 = d;
            }
        }
        const double x_exp2inf = __shfl_sync(0xFFFFFFFFu, d0, base + 0);
        const double u_symp = __shfl_sync(0xFFFFFFFFu, u_mine, base + 1);
        const double x_asym2rec = __shfl_sync(0xFFFFFFFFu, d0, base + 2), x_inf2sym = __shfl_sync(0xFFFFFFFFu, d1, base + 2);
        const double u_sev = __shfl_sync(0xFFFFFFFFu, u_mine, base + 3);
        const double x_mild2rec = __shfl_sync(0xFFFFFFFFu, d0, base + 4), x_sym2sev = __shfl_sync(0xFFFFFFFFu, d1, base + 4);
        const double u_crit = __shfl_sync(0xFFFFFFFFu, u_mine, base + 5);
        const double x_sev2rec = __shfl_sync(0xFFFFFFFFu, d0, base + 6), x_sev2crit = __shfl_sync(0xFFFFFFFFu, d1, base + 6);
        const double u_death = __shfl_sync(0xFFFFFFFFu, u_mine, base + 7);
        const double x_crit2rec = __shfl_sync(0xFFFFFFFFu, d0, base + 8), x_crit2die = __shfl_sync(0xFFFFFFFFu, d1, base + 8);
        const double x_nab = __shfl_sync(0xFFFFFFFFu, d0, base + 9);
        if (slot != 0 || !valid) continue;                         // the half-warp's leader applies the infection
        const unsigned long long dense_pos = ia.log_base ? log_base_v + j : 0ull;
        if (ia.log_base && (int64_t)dense_pos < log.cap) log.target[dense_pos] = -1;       // tombstone unless the infection happens
        unsigned long long key;
        if (ia.hit_key) {                                          // one entry per hit: only the winning one proceeds
            key = ia.hit_key[j];
            if (infect_key[i] != key) continue;
            infect_key[i] = kEmptyKey;
        } else {
            key = infect_key[i];
            infect_key[i] = kEmptyKey;
            if (key == kEmptyKey) continue;
        }
        const int v = (int)(key >> 56) & 0x7F;
        const int lfield = (int)(key >> 48) & 0xFF;
        const int dir = (int)(key >> 40) & 1;
        const int64_t e = (int64_t)(key & 0xFFFFFFFFFFull);
        int32_t source = -1;
        int layer_code = ia.list_layer_code;
        if (lfield != 0xFF) {
            if (ia.hit_src) source = ia.hit_src[j];
            else if (ia.adj_ptr && ((ia.adj_mask >> lfield) & 1u)) source = src_adj;
            else source = dir == 0 ? L.l[lfield].p1[e] : L.l[lfield].p2[e];
            layer_code = lfield;
        }
        // (the variant is known only now: the two protection values are the one late gather)
        const float in_symp_imm = PF(P, symp_imm)[(int64_t)v * n + i], in_sev_imm = PF(P, sev_imm)[(int64_t)v * n + i];
        if (!is_sus) continue;                                     // people.py:470-473

        // breakthrough infections (people.py:486-491, 501)
        if (in_peak != 0.0f) {
            if (in_nbreak == 0) PF(P, rel_trans)[i] = fmul(in_rel_trans, pars.trans_redux);
            PI(P, n_breakthroughs)[i] = in_nbreak + 1;
        }
        // flags (people.py:494-503)
        PB(P, susceptible)[i] = 0; PB(P, naive)[i] = 0; PB(P, recovered)[i] = 0; PB(P, diagnosed)[i] = 0; PB(P, exposed)[i] = 1;
        PI(P, n_infections)[i] = in_ninf + 1;
        PF(P, exposed_variant)[i] = (float)v;
        PB(P, exposed_by_variant)[(int64_t)v * n + i] = 1;
        ++c[INF_INFECTIONS];
        c[INF_REINFECTIONS] += !is_nan(in_drec);
        // infection log (people.py:508-511)
        {
            const unsigned long long pos = ia.log_base ? dense_pos : warp_append(log.count);
            if ((int64_t)pos < log.cap) {
                log.source[pos] = source; log.target[pos] = (int32_t)gi; log.date[pos] = t;
                log.layer[pos] = (int8_t)layer_code; log.variant[pos] = (int8_t)v;
            }
        }
        // exposed -> infectious (people.py:513-520)
        const float e2i = (float)x_exp2inf;
        PF(P, dur_exp2inf)[i] = e2i;
        PF(P, date_exposed)[i] = tf;
        const float d_inf = fadd(e2i, tf);
        PF(P, date_infectious)[i] = d_inf;
        float d_symp = nanf32(), d_sev = nanf32(), d_crit = nanf32(), d_rec = nanf32();
        PF(P, date_diagnosed)[i] = nanf32();
        float dur_disease;
        int symp_class;     // 0 asymptomatic, 1 mild, 2 severe (incl. critical) -- for NAb scaling
        bool is_symp_f = false, is_sev_f = false;

        // prognosis tree (people.py:522-580)
        const float p_symp = prog_prob_imm(pars.rel_symp[v], in_symp_prob, in_symp_imm);
        if (!(u_symp < (double)p_symp)) {
            const double d = x_asym2rec;
            d_rec = (float)dadd((double)d_inf, d);
            dur_disease = (float)dadd((double)e2i, d);
            symp_class = 0;
        } else {
            is_symp_f = true;
            const float i2s = (float)x_inf2sym;
            PF(P, dur_inf2sym)[i] = i2s;
            d_symp = fadd(d_inf, i2s);
            const float p_sev = prog_prob_imm(pars.rel_severe[v], in_sev_prob, in_sev_imm);
            if (!(u_sev < (double)p_sev)) {
                const double d = x_mild2rec;
                d_rec = (float)dadd((double)d_symp, d);
                dur_disease = (float)dadd((double)fadd(e2i, i2s), d);
                symp_class = 1;
            } else {
                is_sev_f = true;
                symp_class = 2;
                const float s2s = (float)x_sym2sev;
                PF(P, dur_sym2sev)[i] = s2s;
                d_sev = fadd(d_symp, s2s);
                const float p_crit = prog_prob_fac(pars.rel_crit[v], in_crit_prob, hosp_max ? pars.no_hosp_factor : 1.0f);
                if (!(u_crit < (double)p_crit)) {
                    const double d = x_sev2rec;
                    d_rec = (float)dadd((double)d_sev, d);
                    dur_disease = (float)dadd((double)fadd(fadd(e2i, i2s), s2s), d);
                } else {
                    const float s2c = (float)x_sev2crit;
                    PF(P, dur_sev2crit)[i] = s2c;
                    d_crit = fadd(d_sev, s2c);
                    const float p_death = prog_prob_fac(pars.rel_death[v], in_death_prob, icu_max ? pars.no_icu_factor : 1.0f);
                    const float pre = fadd(fadd(fadd(e2i, i2s), s2s), s2c);
                    if (!(u_death < (double)p_death)) {
                        const double d = x_crit2rec;
                        d_rec = (float)dadd((double)d_crit, d);
                        dur_disease = (float)dadd((double)pre, d);
                    } else {
                        const double d = x_crit2die;
                        PF(P, date_dead)[i] = (float)dadd((double)d_crit, d);
                        dur_disease = (float)dadd((double)pre, d);
                        d_rec = nanf32();
                    }
                }
            }
        }
        PF(P, date_symptomatic)[i] = d_symp;
        PF(P, date_severe)[i] = d_sev;
        PF(P, date_critical)[i] = d_crit;
        PF(P, date_recovered)[i] = d_rec;
        PF(P, dur_disease)[i] = dur_disease;
#pragma unroll
        for (int k = 0; k < CVB_MAX_VARIANTS; ++k) {
            cv[3 * k + 0] += (k == v);
            cv[3 * k + 1] += (k == v) && is_symp_f;
            cv[3 * k + 2] += (k == v) && is_sev_f;
        }

        // NAbs (immunity.py:138-202 with symp != None)
        if (pars.use_waning) {
            if (in_nab > 0.0f) {
                PF(P, peak_nab)[i] = fmul(in_peak, pars.nab_boost);
            } else {
                double x = x_nab;
                double level = exp2(x);                   // 2**x (NumPy pow in the reference; peak_nab is compared at 1e-6)
                double scale = symp_class == 0 ? pars.rel_imm_asymp : (symp_class == 1 ? pars.rel_imm_mild : pars.rel_imm_severe);
                PF(P, peak_nab)[i] = (float)dmul(dmul(level, scale), pars.nab_norm);
            }
            PI(P, t_nab_event)[i] = t;
        }
        if (S) {
            // the packed state word follows (bit layout: day_fused.cu): not susceptible / naive / recovered / diagnosed any more, no
            // diagnosis date, exposed to variant v, no natural-immunity source until recovery; antibodies from now on if waning
            uint32_t sk = so & ~(SB_SUS | SB_NAIVE | SB_REC | SB_DIAG | SB_DPEND | kEbvMask | kRvMask);
            sk |= SB_EXP | ((uint32_t)(v + 1) << kEbvShift);
            if (pars.use_waning) sk |= SB_HAS_NAB;
            S[i] = sk;
            stock_delta(so, sk, s_delta);
        }
    }

    if (ia.log_base && blockIdx.x == 0 && threadIdx.x == 0) *log.count = __ldcg(ia.log_base) + n_cand;     // (only this thread writes it)
    reduce_counters(c, s_cnt);
    reduce_counters(cv, s_cnt + INF_NK);
    __syncthreads();
    if (S) flush_stock_delta(s_delta, res.counters + (int64_t)t * CVB_N_COUNTERS, ia.vcounters_row, pars.n_variants);
    if (ia.count_flows && threadIdx.x < NK && s_cnt[threadIdx.x]) {
        const int k = threadIdx.x;
        const unsigned long long val = (unsigned long long)s_cnt[k];
        unsigned long long* row = res.counters + (int64_t)t * CVB_N_COUNTERS;
        if (k == INF_INFECTIONS) atomicAdd(row + CVB_C_new_infections, val);
        else if (k == INF_REINFECTIONS) atomicAdd(row + CVB_C_new_reinfections, val);
        else {
            const int q = k - INF_NK, var = q / 3, which = q % 3;
            if (var < pars.n_variants) {
                const int id = which == 0 ? CVB_VC_new_infections_by_variant : (which == 1 ? CVB_VC_new_symptomatic_by_variant : CVB_VC_new_severe_by_variant);
                atomicAdd(res.vcounters + ((int64_t)t * pars.n_variants + var) * CVB_N_VCOUNTERS + id, val);
            }
        }
    }
}

// List mode: claim each listed agent once (first occurrence wins, like np.unique(return_index=True))
__global__ void claim_list_kernel(const int32_t* __restrict__ inds, int64_t n_inds, int32_t variant, int64_t n,
        unsigned long long* __restrict__ infect_key, int32_t* __restrict__ cand, unsigned int* __restrict__ n_cand) {
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_inds; j += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = inds[j];
        if (i < 0 || i >= n) continue;
        unsigned long long key = ((unsigned long long)variant << 56) | (0xFFull << 48) | (unsigned long long)j;
        unsigned long long old = atomicMin(infect_key + i, key);
        if (old == kEmptyKey) { unsigned int pos = warp_append32(n_cand); cand[pos] = (int32_t)i; }
    }
}

__global__ void reset_u32_kernel(unsigned int* p) { *p = 0; }

static int launch_infect(cvb_sim* s, int32_t t, int32_t count_flows, int32_t list_layer_code, int32_t hosp_max, int32_t icu_max,
                         int64_t max_items, cudaStream_t st, bool hits = false, bool with_state = false, const double* tape = nullptr) {
    CVB_REQUIRE(s->log.count, "infect: infection log is not bound (cvb_bind_log)");
    LayerTable L;
    if (build_layer_table(s, L, kTileEdges, 0)) return 1;          // no layer skipped: entry index == layer id
    InfectArgs ia;
    ia.seed = s->seed; ia.n = s->n; ia.t = t; ia.count_flows = count_flows; ia.list_layer_code = list_layer_code;
    ia.hosp_max = hosp_max; ia.icu_max = icu_max;
    ia.id0 = s->partitioned ? s->id0 : 0;
    ia.hit_src = hits ? s->hit_src : nullptr; ia.hit_key = hits ? s->hit_key : nullptr; ia.hit_cap = s->hit_cap;
    const bool use_adj = with_state && s->adj && s->adj_layer_mask;
    ia.adj_ptr = use_adj ? s->adj_ptr : nullptr; ia.adj = use_adj ? s->adj : nullptr; ia.adj_mask = use_adj ? s->adj_layer_mask : 0u;
    ia.tape = tape;
    ia.log_base = nullptr;
    if (!hits) {
        // the candidates are distinct agents (winners of the edge pass, or a claimed list), so candidate j logs at (length before) + j --
        // the length is snapshotted by a stream-ordered copy here and by day_mid_kernel in the fused day
        unsigned long long* snap = s->dev_scalars + 40;
        if (!with_state) CVB_CHECK(cudaMemcpyAsync(snap, s->log.count, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
        ia.log_base = snap;
    }
    ia.beds_direct = with_state ? s->res.counters + (int64_t)t * CVB_N_COUNTERS + CVB_C_n_severe : nullptr;      // n_severe, n_critical are adjacent
    ia.vcounters_row = s->res.vcounters + (int64_t)t * s->nv * CVB_N_VCOUNTERS;
    int grid = grid_for(max_items * 16, kThreads, 148 * (s->tune[6] > 0 ? s->tune[6] : 4 * (CVB_INFECT_MINB > 2 ? 2 : 1)));       // sixteen lanes per agent
    {
        const int32_t* cand = s->cand; const unsigned int* nc = s->n_cand; const unsigned long long* beds = s->beds;
        uint32_t* state = with_state ? s->state : nullptr;
        if (with_state) CVB_CHECK(launch_pdl(infect_kernel, grid, kThreads, 0, st, s->people, s->pars, L, ia, cand, nc, s->infect_key, beds, s->res, s->log, state));
        else infect_kernel<<<grid, kThreads, 0, st>>>(s->people, s->pars, L, ia, cand, nc, s->infect_key, beds, s->res, s->log, state);
    }
    CVB_LAUNCH_CHECK();
    return 0;
}

int launch_infect_winners(cvb_sim* s, int32_t t, bool with_state, cudaStream_t st) {
    int64_t guess = s->n / 64 + 1024;
    return launch_infect(s, t, 1, CVB_LAYER_SEED, -1, -1, guess, st, s->partitioned != 0, with_state);
}

}  // namespace cvb

using namespace cvb;

extern "C" {

int cvb_infect_winners(cvb_sim* s, int32_t t, cvb_stream st) {
    CVB_REQUIRE(s && s->pars_set && s->res.counters, "cvb_infect_winners: handle not ready");
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_infect_winners: day %d outside [0,%d)", t, s->npts);
    cvb::state_touched(s);
    // the number of candidates is only known on the device: size the grid for a large outbreak and let
    // surplus CTAs exit after one load of n_cand
    return launch_infect_winners(s, t, false, (cudaStream_t)st);
}

static int infect_list_impl(cvb_sim* s, const int32_t* inds, int64_t n, int32_t variant, int32_t layer_code, int32_t t,
                            int32_t count_flows, int32_t hosp_max, int32_t icu_max, const double* tape, cudaStream_t st);

int cvb_infect_list(cvb_sim* s, const int32_t* inds, int64_t n, int32_t variant, int32_t layer_code, int32_t t,
                    int32_t count_flows, int32_t hosp_max, int32_t icu_max, cvb_stream st) {
    return infect_list_impl(s, inds, n, variant, layer_code, t, count_flows, hosp_max, icu_max, nullptr, (cudaStream_t)st);
}

int cvb_infect_list_taped(cvb_sim* s, const int32_t* inds, int64_t n, int32_t variant, int32_t layer_code, int32_t t,
                          int32_t count_flows, int32_t hosp_max, int32_t icu_max, const double* tape, cvb_stream st) {
    CVB_REQUIRE(tape, "cvb_infect_list_taped: NULL tape");
    return infect_list_impl(s, inds, n, variant, layer_code, t, count_flows, hosp_max, icu_max, tape, (cudaStream_t)st);
}

}  // extern "C"

static int infect_list_impl(cvb_sim* s, const int32_t* inds, int64_t n, int32_t variant, int32_t layer_code, int32_t t,
                            int32_t count_flows, int32_t hosp_max, int32_t icu_max, const double* tape, cudaStream_t st) {
    CVB_REQUIRE(s && s->pars_set && s->res.counters, "cvb_infect_list: handle not ready");
    CVB_REQUIRE(variant >= 0 && variant < s->nv, "cvb_infect_list: variant %d out of range", variant);
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_infect_list: day %d outside [0,%d)", t, s->npts);
    if (n == 0) return 0;
    cvb::state_touched(s);
    CVB_REQUIRE(inds, "cvb_infect_list: NULL index array");
    reset_u32_kernel<<<1, 1, 0, st>>>(s->n_cand);
    CVB_LAUNCH_CHECK();
    claim_list_kernel<<<grid_for(n), kThreads, 0, st>>>(inds, n, variant, s->n, s->infect_key, s->cand, s->n_cand);
    CVB_LAUNCH_CHECK();
    if (launch_infect(s, t, count_flows, layer_code, hosp_max, icu_max, n, st, false, false, tape)) return 1;
    reset_u32_kernel<<<1, 1, 0, st>>>(s->n_cand);
    CVB_LAUNCH_CHECK();
    return 0;
}
