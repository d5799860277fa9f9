// The fused day pipeline: one simulated day (reference sim.py:558-685) as FIVE launches, driven by a C loop (cvb_run_days) that
// runs whole blocks of days without returning to the host:
//
//   day_begin_kernel   update_nab + stock counts of day t-1 (sim.py:652-674), update_states_pre + check_immunity of day t
//                      (people.py:164-186, immunity.py:303-350), test_prob + People.test (interventions.py:857-981, people.py:589-617)
//                      and contact_tracing.select_cases (interventions.py:1066-1085) -- everything that only needs the agent's OWN state
//   trace_sparse2      contact_tracing.identify / notify_contacts over the cases' adjacency rows (interventions.cu)
//   day_mid_kernel     update_states_post (people.py:189-196, 315-366) + viral load + the 16-byte agent records + today's transmitter
//                      entries (sim.py:602-643)
//   edge_sparse2 / edge_pass_kernel    transmission (edge_pass.cu)
//   infect_kernel      People.infect for the winners (infect.cu)
//
// What makes the per-agent passes cheap is a library-owned PACKED STATE WORD per agent (cvb_sim::state): the 16 bool states plus a
// few "is there anything to read" bits.  A pass reads 4 bytes per agent and touches the float32 date / NAb / immunity arrays only for
// the agents whose bits say they matter (exposed, isolated, has antibodies, has a diagnosis pending, has a quarantine request pending),
// and the agent record is rewritten only when it changes.  The public People arrays stay the truth for everybody else: every change
// is written to them as before; the word is a cache, rebuilt by pack_state_kernel whenever another entry point (or Python) may have
// written the arrays (cvb_sim::state_valid), and cvb_state_check recomputes it from the arrays to verify the bookkeeping (tests).
//
// Results are identical to the unfused kernels (people_kernels.cu, interventions.cu): same arithmetic (cvb_device.cuh), same Philox
// keys, same order of the per-agent steps; tests/test_gpu_fused.py checks both paths against the oracle.
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>
#include <string>
#include <thread>
#include "cvb_internal.cuh"

namespace cvb {

// Rebuild the word of every agent from the public arrays; t_done = the last completed day.  viol counts agents the word cannot
// express (more than one by-variant row set, a by-variant row that disagrees with exposed_variant, antibodies without a peak)
__global__ void __launch_bounds__(kThreads) pack_state_kernel(PeoplePtrs P, uint32_t* __restrict__ S, const uint32_t* __restrict__ S_old,
        int64_t n, int32_t nv, int32_t t_done, const float* __restrict__ ring, int32_t horizon, unsigned int* __restrict__ viol,
        unsigned int* __restrict__ mismatch, int32_t* __restrict__ mismatch_at) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t s = 0;
#pragma unroll
        for (int b = 0; b < 16; ++b) if (P.template get<uint8_t>(CVB_F_susceptible + b)[i]) s |= 1u << b;
        const float peak = PF(P, peak_nab)[i];
        if (peak != 0.0f) s |= SB_HAS_NAB;
        else if (PF(P, nab)[i] != 0.0f) atomicAdd(viol, 1u);
        bool nz = false;
        int ebv = 0, ibv = 0;
        for (int v = 0; v < nv; ++v) {
            nz |= PF(P, sus_imm)[(int64_t)v * n + i] != 0.0f || PF(P, symp_imm)[(int64_t)v * n + i] != 0.0f || PF(P, sev_imm)[(int64_t)v * n + i] != 0.0f;
            if (PB(P, exposed_by_variant)[(int64_t)v * n + i]) { if (ebv) atomicAdd(viol, 1u); ebv = v + 1; }
            if (PB(P, infectious_by_variant)[(int64_t)v * n + i]) { if (ibv) atomicAdd(viol, 1u); ibv = v + 1; }
        }
        if (ibv && ibv != ebv) atomicAdd(viol, 1u);
        if (s & SB_EXP) { const float ev = PF(P, exposed_variant)[i]; if (is_nan(ev) || (int)ev + 1 != ebv) atomicAdd(viol, 1u); }
        if (s & SB_INF) { const float iv = PF(P, infectious_variant)[i]; if (is_nan(iv) || (int)iv + 1 != ebv || !ibv) atomicAdd(viol, 1u); }
        if (nz) s |= SB_IMM_NZ;
        if (ibv) s |= SB_IBV;
        s |= (uint32_t)ebv << kEbvShift;
        for (int h = 0; h < horizon; ++h) if (ring[(int64_t)h * n + i] >= 0.0f) s |= SB_QPEND;
        if (!(s & SB_DIAG) && !is_nan(PF(P, date_diagnosed)[i])) s |= SB_DPEND;
        if (due(PF(P, date_recovered)[i], t_done)) {
            const float rvf = PF(P, recovered_variant)[i];
            const int rv = is_nan(rvf) ? -1 : (int)rvf;
            if (rv >= 0 && rv < nv) s |= (uint32_t)(rv + 1) << kRvShift;
        }
        if (S_old) {                                                 // verification mode: compare, do not store
            const uint32_t ignore = SB_RS_VALID | SB_RS_SUS | SB_RS_QUAR;
            uint32_t o = S_old[i];
            // QPEND may stay set after its request was served when the ring has several slots; IMM_NZ may stay set for a day
            uint32_t soft = SB_QPEND;
            if (((o ^ s) & ~ignore & ~soft) || ((s & soft) & ~o)) {
                const unsigned int k = atomicAdd(mismatch, 1u);
                if (k < 8) { mismatch_at[3 * k] = (int32_t)i; mismatch_at[3 * k + 1] = (int32_t)o; mismatch_at[3 * k + 2] = (int32_t)s; }
            }
        } else {
            S[i] = s;
        }
    }
}

// absolute stock counts of the packed words (the base of the first fused day): one launch after pack_state_kernel
__global__ void __launch_bounds__(kThreads) count_state_kernel(const uint32_t* __restrict__ S, int64_t n, int32_t nv,
                                                               unsigned long long* __restrict__ row, unsigned long long* __restrict__ vrow) {
    __shared__ int s_cnt[kStockSlots];
    if (threadIdx.x < kStockSlots) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t n_pad = (n + 31) / 32 * 32;
    int c[17];
#pragma unroll
    for (int k = 0; k < 17; ++k) c[k] = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_pad; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t s = i < n ? S[i] : SB_DEAD;
#pragma unroll
        for (int b = 0; b < 16; ++b) c[b] += __popc(__ballot_sync(0xFFFFFFFFu, i < n && ((s >> b) & 1u)));
        c[16] += __popc(__ballot_sync(0xFFFFFFFFu, !(s & SB_DEAD)));
        const int e = i < n ? sb_ebv(s) : 0;
        if (e) { atomicAdd(&s_cnt[17 + 2 * (e - 1)], 1); if (s & SB_IBV) atomicAdd(&s_cnt[18 + 2 * (e - 1)], 1); }
    }
    if (lane_id() == 0) {
#pragma unroll
        for (int k = 0; k < 17; ++k) if (c[k]) atomicAdd(&s_cnt[k], c[k]);
    }
    __syncthreads();
    flush_stock_delta(s_cnt, row, vrow, nv);
}

// ================================================================================================================================
// day_begin_kernel: one agent per thread, one contiguous chunk of agents per CTA
// ================================================================================================================================
enum { F_INFECTIOUS = 0, F_SYMPTOMATIC, F_SEVERE, F_CRITICAL, F_RECOVERIES, F_DEATHS, F_KNOWN_DEATHS, F_TESTS, F_NK };
constexpr int kImmQueueCap2 = 64;

struct DayBeginArgs {
    int64_t n, id0, chunk;              // chunk: agents per CTA (a multiple of 32)
    int32_t t;                          // the day whose update_states_pre / interventions run; the END part closes day t - 1
    int32_t nv, waning, vaxpars;
    const double* nab_kin; int64_t nab_kin_len;
    unsigned long long* counters; unsigned long long* vcounters;
    const unsigned long long* base_row; const unsigned long long* base_vrow;    // yesterday's stock counts (PRE: copied into today's row)
    double* partial;                                           // [gridDim.x][3] per-CTA float64 sums, added up by the next kernel
    unsigned int* n_trans; unsigned int* n_case;
    // test_prob
    cvb_test_prob_pars tp; int32_t test_plain;                 // test_plain: quarantine state does not change the probability
    uint64_t seed;
    // contact tracing: cases -> entries (or, agent-partitioned: -> the local case bitmap that the host all-gathers)
    const long long* adj_ptr; uint4* case_ent;
    unsigned int* case_bits;
};

// check_immunity for one queued agent x variant (immunity.py:303-350): float64, rounded once to float32.
// entry = {agent, nab bits, variant | natural-immunity source + 1 << 4 | vaccine source << 8 | vaccinated << 12, -}
__device__ __noinline__ float2 immunity_eval2(const uint4 en, int64_t n, const cvb_pars& pars, float* __restrict__ sus_imm,
                                              float* __restrict__ symp_imm, float* __restrict__ sev_imm) {
    const int64_t i = (int64_t)en.x;
    const float nab = __uint_as_float(en.y);
    const int v = (int)(en.z & 15u), rvi = (int)((en.z >> 4) & 15u) - 1, vsi = (int)((en.z >> 8) & 15u);
    const bool vacc = (en.z >> 12) & 1u;
    const double natural = rvi >= 0 ? (double)pars.immunity[v][rvi] : 0.0;
    const double vaccine = vacc ? pars.vaccine_imm[vsi][v] : 0.0;
    const double enab = dmul((double)nab, fmax(natural, vaccine));
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
    if (enab > 0.0)
        calc_ve3(enab, pars.exp_alpha_inf, pars.beta_inf, pars.exp_alpha_symp_inf, pars.beta_symp_inf, pars.exp_alpha_sev_symp, pars.beta_sev_symp, s0, s1, s2);
    sus_imm[(int64_t)v * n + i] = s0;
    symp_imm[(int64_t)v * n + i] = s1;
    sev_imm[(int64_t)v * n + i] = s2;
    return make_float2(s0, s1);
}

#ifndef CVB_BEGIN_MINB
#define CVB_BEGIN_MINB 4      // 64 registers: occupancy beats the few spilled values (measured: 49 -> 42 us at C2, profiles/r2/README.md)
#endif
#ifndef CVB_MID_MINB
#define CVB_MID_MINB 4
#endif
template <bool END, bool PRE, bool TEST, bool TSEL>
__global__ void __launch_bounds__(kThreads, CVB_BEGIN_MINB) day_begin_kernel(PeoplePtrs P, uint32_t* __restrict__ S, const __grid_constant__ cvb_pars pars,
                                                               const __grid_constant__ DayBeginArgs A) {
    __shared__ int s_flow[F_NK + CVB_MAX_VARIANTS];
    __shared__ int s_delta[kStockSlots];
    __shared__ uint4 s_queue[(kThreads / 32) * kImmQueueCap2];
    __shared__ double s_sum[3][kThreads / 32];
    const int lane = lane_id();
    const unsigned lt_mask = (1u << lane) - 1u;
    uint4* q_imm = s_queue + warp_id() * kImmQueueCap2;
    int qn = 0;
    const int64_t n = A.n;
    const int32_t t = A.t;
    const int nv = A.nv;
    const bool waning = A.waning != 0;
    const float tf = (float)t;
    const float qnan = nanf32();
    pdl_trigger();
    for (int k = threadIdx.x; k < F_NK + CVB_MAX_VARIANTS; k += blockDim.x) s_flow[k] = 0;
    for (int k = threadIdx.x; k < kStockSlots; k += blockDim.x) s_delta[k] = 0;
    pdl_wait();                                                 // everything below reads what the previous kernel wrote
    if (PRE && blockIdx.x == 0) {
        if (threadIdx.x == 0) *A.n_trans = 0;                  // today's transmitter list starts empty (filled by day_mid_kernel)
        if (threadIdx.x < kStockSlots) {                        // today's stock counts start from yesterday's
            unsigned long long* dst = stock_slot(A.counters + (int64_t)t * CVB_N_COUNTERS, A.vcounters + (int64_t)t * nv * CVB_N_VCOUNTERS, nv, threadIdx.x);
            const unsigned long long* src = stock_slot(const_cast<unsigned long long*>(A.base_row), const_cast<unsigned long long*>(A.base_vrow), nv, threadIdx.x);
            if (dst) atomicAdd(dst, *src);
        }
    }
    __syncthreads();
    int c_tests = 0;
    double sum_nab = 0.0, sum_sus = 0.0, sum_symp = 0.0;

    uint8_t* exposed = PB(P, exposed); uint8_t* infectious = PB(P, infectious); uint8_t* symptomatic = PB(P, symptomatic);
    uint8_t* severe = PB(P, severe); uint8_t* critical = PB(P, critical); uint8_t* recovered = PB(P, recovered);
    uint8_t* dead = PB(P, dead); uint8_t* diagnosed = PB(P, diagnosed); uint8_t* susceptible = PB(P, susceptible);
    uint8_t* isolated = PB(P, isolated); uint8_t* known_dead = PB(P, known_dead); uint8_t* known_contact = PB(P, known_contact);
    uint8_t* quarantined = PB(P, quarantined); uint8_t* tested = PB(P, tested);
    uint8_t* exp_by_var = PB(P, exposed_by_variant); uint8_t* inf_by_var = PB(P, infectious_by_variant);
    float* exp_var = PF(P, exposed_variant); float* inf_var = PF(P, infectious_variant); float* rec_var = PF(P, recovered_variant);
    float* nab = PF(P, nab);
    float* sus_imm = PF(P, sus_imm); float* symp_imm = PF(P, symp_imm); float* sev_imm = PF(P, sev_imm);
    float* d_tested = PF(P, date_tested); float* d_diag = PF(P, date_diagnosed); float* d_pos = PF(P, date_pos_test);

    const int64_t lo = (int64_t)blockIdx.x * A.chunk;
    const int64_t hi = lo + A.chunk < n ? lo + A.chunk : n;
    const int64_t hi_pad = lo + (hi - lo + 31) / 32 * 32;                    // whole warps stay in the loop (ballots)
    for (int64_t i = lo + threadIdx.x; i < hi_pad; i += blockDim.x) {
        const bool in = i < hi;
        const uint32_t s0 = in ? S[i] : 0u;
        uint32_t s = s0;

        // ---- every load this agent needs follows from its state word: issued together, before any store ----
        const bool has_nab = waning && (s0 & SB_HAS_NAB);
        float nb = 0.0f, pk = 0.0f;
        int32_t te = 0, vsi = 0;
        if (has_nab) {
            nb = nab[i];
            if (END) { pk = PF(P, peak_nab)[i]; te = PI(P, t_nab_event)[i]; }
            if (PRE && A.vaxpars && (s0 & SB_VACC)) vsi = PI(P, vaccine_source)[i];
        }
        float di = qnan, ds = qnan, dv = qnan, dc = qnan, dr = qnan, dd = qnan, dei = qnan, dq = qnan, deq = qnan;
        if (PRE && (s0 & SB_EXP)) {
            di = PF(P, date_infectious)[i]; ds = PF(P, date_symptomatic)[i]; dv = PF(P, date_severe)[i];
            dc = PF(P, date_critical)[i]; dr = PF(P, date_recovered)[i]; dd = PF(P, date_dead)[i];
        }
        if (PRE && (s0 & SB_ISO)) dei = PF(P, date_end_isolation)[i];
        if (TEST && !A.test_plain && in) {
            if (A.tp.quar_policy == 0 || A.tp.quar_policy == 2) dq = PF(P, date_quarantined)[i];
            if (A.tp.quar_policy == 1 || A.tp.quar_policy == 2) deq = PF(P, date_end_quarantine)[i];
        }

        // ---- close day t-1: update_nab (immunity.py:205-213) and the sum of NAbs over the living (sim.py:666-672) ----
        if (END && has_nab) {                                                // has_nabs = true(peak_nab)  (sim.py:666-669)
            int64_t dt = (int64_t)(t - 1) - (int64_t)te;
            if (dt < 0) dt += A.nab_kin_len;                                 // NumPy negative index wraps
            const double kin = (dt >= 0 && dt < A.nab_kin_len) ? A.nab_kin[dt] : 0.0;
            nb = nab_step(nb, pk, kin);
            nab[i] = nb;
            if (!(s0 & SB_DEAD)) sum_nab += (double)nb;
        }

        if (PRE) {
            // ---- update_states_pre (people.py:164-186); is_exp is taken once, before any transition (people.py:169) ----
            const int ev = sb_ebv(s0) - 1;
            const float evf = (float)ev;
            if (s0 & SB_EXP) {
                if (!(s & SB_INF) && due(di, t)) {                           // people.py:222-232
                    infectious[i] = 1; inf_var[i] = evf; s |= SB_INF;
                    if (ev >= 0 && ev < nv) { inf_by_var[(int64_t)ev * n + i] = 1; s |= SB_IBV; atomicAdd(&s_flow[F_NK + ev], 1); }
                    atomicAdd(&s_flow[F_INFECTIOUS], 1);
                }
                if (!(s & SB_SYMP) && due(ds, t)) { symptomatic[i] = 1; s |= SB_SYMP; atomicAdd(&s_flow[F_SYMPTOMATIC], 1); }     // people.py:235-253
                if (!(s & SB_SEV) && due(dv, t)) { severe[i] = 1; s |= SB_SEV; atomicAdd(&s_flow[F_SEVERE], 1); }
                if (!(s & SB_CRIT) && due(dc, t)) { critical[i] = 1; s |= SB_CRIT; atomicAdd(&s_flow[F_CRITICAL], 1); }
                if (!(s & SB_REC) && due(dr, t)) {                           // people.py:256-291
                    exposed[i] = 0; infectious[i] = 0; symptomatic[i] = 0; severe[i] = 0; critical[i] = 0;
                    recovered[i] = 1;
                    rec_var[i] = evf; inf_var[i] = qnan; exp_var[i] = qnan;
                    for (int v = 0; v < nv; ++v) { exp_by_var[(int64_t)v * n + i] = 0; inf_by_var[(int64_t)v * n + i] = 0; }
                    s &= ~(SB_EXP | SB_INF | SB_SYMP | SB_SEV | SB_CRIT | SB_IBV | kEbvMask | kRvMask);
                    s |= SB_REC;
                    if (ev >= 0 && ev < nv) s |= (uint32_t)(ev + 1) << kRvShift;
                    if (waning) {
                        susceptible[i] = 1; diagnosed[i] = 0;
                        if (s & SB_DIAG) s |= SB_DPEND;                      // the (old) date_diagnosed is still set
                        s |= SB_SUS; s &= ~SB_DIAG;
                    }
                    atomicAdd(&s_flow[F_RECOVERIES], 1);
                }
            }
            if ((s0 & SB_ISO) && due(dei, t)) { isolated[i] = 0; s &= ~SB_ISO; }     // people.py:368-374
            if ((s0 & SB_EXP) && !(s & SB_DEAD) && due(dd, t)) {             // people.py:294-312
                dead[i] = 1;
                if (s & SB_DIAG) { known_dead[i] = 1; s |= SB_KDEAD; atomicAdd(&s_flow[F_KNOWN_DEATHS], 1); }
                susceptible[i] = 0; exposed[i] = 0; infectious[i] = 0; symptomatic[i] = 0; severe[i] = 0; critical[i] = 0;
                known_contact[i] = 0; quarantined[i] = 0; recovered[i] = 0;
                inf_var[i] = qnan; exp_var[i] = qnan; rec_var[i] = qnan;
                s &= ~(SB_SUS | SB_EXP | SB_INF | SB_SYMP | SB_SEV | SB_CRIT | SB_KCONTACT | SB_QUAR | SB_REC | kRvMask);
                s |= SB_DEAD;
                atomicAdd(&s_flow[F_DEATHS], 1);
            }

            // ---- check_immunity (immunity.py:303-350): agents with antibodies and an immunity source are queued per warp and
            //      evaluated 32 at a time; everybody else has protection exactly 0 (written only if something non-zero is stored) ----
            if (waning) {
                const int rv1 = sb_rv(s);
                const bool vacc = A.vaxpars && (s & SB_VACC) && vsi >= 0 && vsi < CVB_MAX_VACCINES;
                const bool heavy = (s & SB_HAS_NAB) && nb > 0.0f && (rv1 != 0 || vacc);
                if (!heavy && (s & SB_IMM_NZ))
                    for (int v = 0; v < nv; ++v) { sus_imm[(int64_t)v * n + i] = 0.0f; symp_imm[(int64_t)v * n + i] = 0.0f; sev_imm[(int64_t)v * n + i] = 0.0f; }
                if (heavy) s |= SB_IMM_NZ; else s &= ~SB_IMM_NZ;
                const unsigned m = __ballot_sync(0xFFFFFFFFu, heavy);
                if (m) {
                    const unsigned packed = ((unsigned)rv1 << 4) | ((unsigned)(vacc ? vsi : 0) << 8) | ((unsigned)vacc << 12);
                    for (int v = 0; v < nv; ++v) {
                        if (heavy) q_imm[qn + __popc(m & lt_mask)] = make_uint4((unsigned)i, __float_as_uint(nb), packed | (unsigned)v, 0u);
                        qn += __popc(m);
                        __syncwarp();
                        if (qn >= 32) {
                            qn -= 32;
                            const uint4 en = q_imm[qn + lane];
                            __syncwarp();
                            const float2 r = immunity_eval2(en, n, pars, sus_imm, symp_imm, sev_imm);
                            sum_sus += (double)r.x; sum_symp += (double)r.y;
                        }
                    }
                }
            }
        }

        // ---- test_prob + People.test (interventions.py:921-981, people.py:589-617) and today's cases (interventions.py:1066-1085) ----
        if ((TEST || TSEL) && in) {
            float ddiag_new = qnan;
            bool ddiag_known = false;
            if (TEST && !(s & SB_DIAG)) {                                    // diagnosed people do not test (interventions.py:973)
                const bool symp = (s & SB_SYMP) != 0;
                bool qt = false;                                             // interventions.py:691-715 get_quar_inds
                if (!A.test_plain) {
                    switch (A.tp.quar_policy) {
                        case 0:  qt = dq == tf - 1.0f; break;
                        case 1:  qt = deq == tf + 1.0f; break;
                        case 2:  qt = (dq == tf - 1.0f) || (deq == tf + 1.0f); break;
                        default: qt = (s & SB_QUAR) != 0; break;
                    }
                }
                const double prob = qt ? (symp ? A.tp.symp_quar_prob : A.tp.asymp_quar_prob) : (symp ? A.tp.symp_prob : A.tp.asymp_prob);
                if (prob > 0.0 && keyed_uniform(A.seed, P_TEST, (uint32_t)A.tp.index, t, i + A.id0, 0) < prob) {
                    ++c_tests;
                    tested[i] = 1; s |= SB_TESTED;
                    d_tested[i] = tf;
                    if ((s & SB_INF) && keyed_uniform(A.seed, P_TEST_SENS, (uint32_t)A.tp.index, t, i + A.id0, 0) < A.tp.sensitivity) {
                        const float old = d_diag[i];
                        ddiag_known = true; ddiag_new = old;
                        if (is_nan(old) && keyed_uniform(A.seed, P_TEST_LOSS, (uint32_t)A.tp.index, t, i + A.id0, 0) < 1.0 - A.tp.loss_prob) {
                            ddiag_new = (float)(t + A.tp.test_delay);
                            d_diag[i] = ddiag_new;
                            d_pos[i] = tf;
                            s |= SB_DPEND;
                        }
                    }
                }
            }
            if (TSEL && (s & SB_DPEND)) {                                    // a case: date_diagnosed == t
                const float dg = ddiag_known ? ddiag_new : d_diag[i];
                if (dg == tf) {
                    if (A.case_bits) atomicOr(A.case_bits + (i >> 5), 1u << (i & 31));
                    else {
                        const long long beg = A.adj_ptr[i], end = A.adj_ptr[i + 1];
                        A.case_ent[warp_append32(A.n_case)] = make_uint4((unsigned)i, (unsigned)(end - beg), (unsigned)(unsigned long long)beg, (unsigned)((unsigned long long)beg >> 32));
                    }
                }
            }
        }
        if (s != s0) { S[i] = s; stock_delta(s0, s, s_delta); }
    }
    if (PRE && waning && lane < qn) {
        const float2 r = immunity_eval2(q_imm[lane], n, pars, sus_imm, symp_imm, sev_imm);
        sum_sus += (double)r.x; sum_symp += (double)r.y;
    }

    // ---- counters: flows and stock differences of day t; float64 sums ----
    if (TEST) {
        const int w = __reduce_add_sync(0xFFFFFFFFu, c_tests);
        if (lane == 0 && w) atomicAdd(&s_flow[F_TESTS], w);
    }
    double sums[3] = {sum_nab, sum_sus, sum_symp};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        double v = sums[q];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
        if (lane == 0) s_sum[q][warp_id()] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int wq = 0; wq < (int)(blockDim.x >> 5); ++wq) v += s_sum[threadIdx.x][wq];
        A.partial[(int64_t)blockIdx.x * 3 + threadIdx.x] = v;
    }
    if ((PRE || TEST) && threadIdx.x < F_NK + CVB_MAX_VARIANTS && s_flow[threadIdx.x]) {
        unsigned long long* row = A.counters + (int64_t)t * CVB_N_COUNTERS;
        const unsigned long long uv = (unsigned long long)s_flow[threadIdx.x];
        switch (threadIdx.x) {
            case F_INFECTIOUS:   atomicAdd(row + CVB_C_new_infectious, uv); break;
            case F_SYMPTOMATIC:  atomicAdd(row + CVB_C_new_symptomatic, uv); break;
            case F_SEVERE:       atomicAdd(row + CVB_C_new_severe, uv); break;
            case F_CRITICAL:     atomicAdd(row + CVB_C_new_critical, uv); break;
            case F_RECOVERIES:   atomicAdd(row + CVB_C_new_recoveries, uv); break;
            case F_DEATHS:       atomicAdd(row + CVB_C_new_deaths, uv); break;
            case F_KNOWN_DEATHS: atomicAdd(row + CVB_C_new_known_deaths, uv); break;
            case F_TESTS:        atomicAdd(row + CVB_C_new_tests, uv); break;
            default: {
                const int var = threadIdx.x - F_NK;
                if (var < nv) atomicAdd(A.vcounters + ((int64_t)t * nv + var) * CVB_N_VCOUNTERS + CVB_VC_new_infectious_by_variant, uv);
            }
        }
    }
    if (PRE || TEST) flush_stock_delta(s_delta, A.counters + (int64_t)t * CVB_N_COUNTERS, A.vcounters + (int64_t)t * nv * CVB_N_VCOUNTERS, nv);
}

// The per-CTA partial sums of the day_begin_kernel that has just finished, added up in a fixed order by ONE warp-triple of the next
// kernel in the stream (deterministic float64 sums; no ticket / fence / extra barrier in the big kernel)
__device__ __forceinline__ void sum_partials(const double* __restrict__ partial, int n_part, double* __restrict__ sums, int32_t t_end, int32_t t_pre) {
    if (threadIdx.x < 96) {
        const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
        double v = 0.0;
        for (int b = lane; b < n_part; b += 32) v += __ldcg(partial + (int64_t)b * 3 + q);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
        if (lane == 0) {
            if (q == 0) { if (t_end >= 0) sums[(int64_t)t_end * 4 + 0] = v; }
            else if (t_pre >= 0) sums[(int64_t)t_pre * 4 + q] = v;
        }
    }
}
__global__ void sum_partials_kernel(const double* __restrict__ partial, int n_part, double* __restrict__ sums, int32_t t_end, int32_t t_pre) {
    pdl_wait();
    sum_partials(partial, n_part, sums, t_end, t_pre);
}

// ================================================================================================================================
// day_mid_kernel: update_states_post + prepare_transmission; one agent per thread, one contiguous chunk of agents per CTA
// ================================================================================================================================
enum { M_DIAGNOSES = 0, M_QUARANTINED, M_ISOLATED, M_NK };
constexpr int kMidChunk = 1024;                     // agents per CTA: bounds the shared-memory transmitter buffer (32 B per entry)

struct DayMidArgs {
    int64_t n;
    int32_t t, nv, horizon, dense;                  // dense: some layer is streamed densely today (records for everyone + ts8 + bitmap)
    int32_t chunk;                                  // agents per CTA (<= kMidChunk)
    float* quar_slot;
    unsigned long long* counters; unsigned long long* vcounters;
    TransRecords rec;
    unsigned int* inf_bits;
    const long long* adj_ptr;                       // NULL: no adjacency (every layer dense)
    uint4* trans_ent; unsigned int* n_trans;
    unsigned int* n_cand; unsigned int* n_case;
    int32_t* trans_list;                            // the plain list, in the same order as the entries
    const unsigned long long* log_count; unsigned long long* log_base;
    uint8_t* codes; const float* base_trans; unsigned int* part_flags;     // agent-partitioned: the 1-byte transmit codes the host all-gathers
    const double* partial; int32_t n_part, t_end;   // day_begin_kernel's per-CTA sums (t_end = t - 1 if it closed a day, else -1)
    double* sums;
};

__global__ void __launch_bounds__(kThreads, CVB_MID_MINB) day_mid_kernel(PeoplePtrs P, uint32_t* __restrict__ S, const __grid_constant__ cvb_pars pars,
                                                             const __grid_constant__ DayMidArgs A) {
    __shared__ int s_cnt[M_NK];
    __shared__ int s_delta[kStockSlots];
    __shared__ uint4 s_ent[2 * kMidChunk];                                   // this CTA's transmitter entries, flushed with ONE global atomic
    __shared__ unsigned int s_n_ent, s_base;
    pdl_trigger();
    if (threadIdx.x < M_NK) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x < kStockSlots) s_delta[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_n_ent = 0;
    pdl_wait();                                                 // everything below reads what the previous kernels wrote
    if (blockIdx.x == 0 && threadIdx.x == 0) { *A.n_cand = 0; *A.n_case = 0; *A.log_base = *A.log_count; }   // today's candidates start empty; the case list was
                                                                                                          // consumed; infect_kernel logs at (this length) + j
    if (blockIdx.x == gridDim.x - 1) sum_partials(A.partial, A.n_part, A.sums, A.t_end, A.t);     // day_begin_kernel's float64 sums
    __syncthreads();
    const int64_t n = A.n;
    const int32_t t = A.t;
    const int nv = A.nv;
    const float tf = (float)t;
    const float qnan = nanf32();
    const int lane = lane_id();
    uint8_t* diagnosed = PB(P, diagnosed); uint8_t* quarantined = PB(P, quarantined); uint8_t* isolated = PB(P, isolated);
    float* d_pos = PF(P, date_pos_test); const float* d_diag = PF(P, date_diagnosed); float* d_quar = PF(P, date_quarantined);
    float* d_end_quar = PF(P, date_end_quarantine); float* d_end_iso = PF(P, date_end_isolation); const float* d_rec = PF(P, date_recovered);
    const float* rel_trans = PF(P, rel_trans); const float* rel_sus = PF(P, rel_sus);
    const float* d_inf = PF(P, date_infectious); const float* d_dead = PF(P, date_dead);
    const float* sus_imm = PF(P, sus_imm);

    const int64_t lo = (int64_t)blockIdx.x * A.chunk;
    const int64_t hi = lo + A.chunk < n ? lo + A.chunk : n;
    const int64_t hi_pad = lo + (hi - lo + 31) / 32 * 32;
    for (int64_t i = lo + threadIdx.x; i < hi_pad; i += blockDim.x) {
        const bool in = i < hi;
        const uint32_t s0 = in ? S[i] : 0u;
        uint32_t s = s0;
        bool can_trans = false;
        // nothing can change for this agent today and its stored record is right: the common case costs one 4-byte load
        const bool settled = !A.dense && (s0 & SB_RS_VALID) && !(s0 & (SB_INF | SB_IMM_NZ | SB_QPEND | SB_DPEND | SB_QUAR)) &&
                             (((s0 & SB_SUS) != 0) == ((s0 & SB_RS_SUS) != 0)) && !(s0 & SB_RS_QUAR);
        if (in && !settled) {
            // ---- every load follows from the state word: issued together, before any store ----
            float ddiag = qnan, dpos = qnan, pend = -1.0f, end_q = qnan, drec = qnan;
            if (s0 & SB_DPEND) { ddiag = d_diag[i]; dpos = d_pos[i]; }
            if (s0 & SB_QPEND) pend = A.quar_slot[i];
            if (s0 & SB_QUAR) end_q = d_end_quar[i];
            if (s0 & (SB_DPEND | SB_INF)) drec = d_rec[i];
            const float rs = rel_sus[i];
            const float imm0 = (s0 & SB_IMM_NZ) ? sus_imm[i] : 0.0f;
            float rtv = 0.0f, dinf = qnan, ddead = qnan;
            long long beg = 0, end = 0;
            if (s0 & SB_INF) {
                rtv = rel_trans[i]; dinf = d_inf[i]; ddead = d_dead[i];
                if (A.adj_ptr) { beg = A.adj_ptr[i]; end = A.adj_ptr[i + 1]; }
            }
            // ---- update_states_post (people.py:189-196) ----
            if ((s & SB_DPEND) && !(s & SB_DIAG)) {                          // check_diagnosed (people.py:315-332)
                if (due(dpos, t)) { d_pos[i] = qnan; atomicAdd(&s_cnt[M_DIAGNOSES], 1); }
                if (due(ddiag, t)) { diagnosed[i] = 1; s |= SB_DIAG; s &= ~SB_DPEND; }
            }
            bool quar = (s & SB_QUAR) != 0;
            if (pend >= 0.0f) {                                              // check_quar (people.py:335-358)
                A.quar_slot[i] = -1.0f;
                if (quar) {
                    if (pend > end_q) { end_q = pend; d_end_quar[i] = end_q; }            // Python max(old, requested)
                } else if (!(s & (SB_DEAD | SB_REC | SB_DIAG | SB_ISO))) {
                    quarantined[i] = 1; quar = true;
                    d_quar[i] = tf;
                    end_q = pend;
                    d_end_quar[i] = end_q;
                    atomicAdd(&s_cnt[M_QUARANTINED], 1);
                }
            }
            if ((s & SB_QPEND) && A.horizon == 1) s &= ~SB_QPEND;            // (several slots: the bit stays, the slot is re-read)
            if (quar) {
                if (ddiag == tf) { end_q = tf; d_end_quar[i] = tf; }
                if (due(end_q, t)) { quarantined[i] = 0; quar = false; }
            }
            if (quar) s |= SB_QUAR; else s &= ~SB_QUAR;
            if (ddiag == tf) {                                               // check_enter_iso (people.py:361-366)
                isolated[i] = 1; s |= SB_ISO;
                d_end_iso[i] = drec;
                atomicAdd(&s_cnt[M_ISOLATED], 1);
            }
            // ---- prepare_transmission (sim.py:602-643): ONE 16-byte record per agent (cvb_device.cuh:AgentRecord) ----
            const bool iso = (s & SB_ISO) != 0;
            bool inf = (s & SB_INF) != 0;
            const bool sus = (s & SB_SUS) != 0;
            int var = sb_ebv(s) - 1;
            if (inf && !(var >= 0 && var < nv)) { inf = false; var = 0; }
            const bool simple = !inf && !(s & SB_IMM_NZ) && !A.dense;
            if (simple) {
                const uint32_t want = SB_RS_VALID | (sus ? SB_RS_SUS : 0u) | (quar ? SB_RS_QUAR : 0u);
                if ((s & (SB_RS_VALID | SB_RS_SUS | SB_RS_QUAR)) != want) {
                    if (A.codes) A.codes[i] = 0;
                    A.rec.rec[i] = make_float4(0.0f, sus ? rs : 0.0f, 0.0f, __uint_as_float(quar ? 32u : 0u));
                    s = (s & ~(SB_RS_VALID | SB_RS_SUS | SB_RS_QUAR)) | want;
                }
            } else {
                s &= ~(SB_RS_VALID | SB_RS_SUS | SB_RS_QUAR);
                const bool symp = (s & SB_SYMP) != 0;
                uint32_t code = quar ? 32u : 0u;                              // the quarantine bit matters for targets too
                float rt = 0.0f;
                if (inf && rtv != 0.0f) {                                     // can transmit (a zero rel_trans never does)
                    rt = rtv;
                    const bool early = viral_load_early(t, dinf, drec, ddead, pars.frac_time, pars.high_cap);
                    bool redux = false;
                    if (A.codes) {                                            // other GPUs rebuild rt from its initial value
                        const float base = A.base_trans[i];
                        redux = rt != base;
                        if (redux && rt != fmul(base, pars.trans_redux)) atomicAdd(A.part_flags + 1, 1u);
                    }
                    code = transmit_code(var, symp, iso, quar, early, redux);
                    can_trans = true;
                    if (!A.codes) {
                        const unsigned int pos = atomicAdd(&s_n_ent, 1u);     // shared memory: this CTA's entries
                        s_ent[2 * pos] = make_uint4((unsigned)i, (unsigned)(end - beg), (unsigned)(unsigned long long)beg, (unsigned)((unsigned long long)beg >> 32));
                        s_ent[2 * pos + 1] = make_uint4(__float_as_uint(rt), code, 0u, 0u);
                    }
                }
                if (A.codes) A.codes[i] = can_trans ? (uint8_t)code : (uint8_t)0;
                float s_rec = sus ? rs : 0.0f, imm_rec = imm0;
                if (nv == 1 && !quar) { s_rec = record_sus(s_rec, 0u, 1.0f, imm_rec); imm_rec = 0.0f; }
                if (A.rec.rec) A.rec.rec[i] = make_float4(rt, s_rec, imm_rec, __uint_as_float(code));
                if (A.rec.ts8) {
                    const float vl = rt != 0.0f ? viral_load_value((code & 64u) != 0, pars.frac_time, pars.load_ratio) : 0.0f;
                    for (int l = 0; l < pars.n_layers; ++l) {
                        if (!((A.rec.ts8_mask >> l) & 1u)) continue;
                        float2 o;
                        o.x = rt != 0.0f ? rel_trans_layer(rt, true, symp, iso, quar, pars.asymp_factor, pars.iso_factor[l], pars.quar_factor[l],
                                                           pars.beta_layer[l], vl) : 0.0f;
                        o.y = sus ? rel_sus_layer(rs, true, quar, pars.quar_factor[l], imm0) : 0.0f;
                        A.rec.ts8[(int64_t)l * n + i] = o;
                    }
                }
            }
            if (s != s0) { S[i] = s; stock_delta(s0, s, s_delta); }
        }
        if (A.dense) {                                                       // transmit bitmap for the dense streaming pass: a warp = one word
            const unsigned word = __ballot_sync(0xFFFFFFFFu, can_trans);
            if (lane == 0 && i < n) A.inf_bits[i >> 5] = word;
        }
    }
    __syncthreads();
    // flush this CTA's transmitter entries: one global atomic, coalesced copies; the list stays grouped by agent range, so
    // neighbouring groups of the edge pass walk neighbouring adjacency rows
    if (threadIdx.x == 0 && s_n_ent) s_base = atomicAdd(A.n_trans, s_n_ent);
    __syncthreads();
    const unsigned int cnt = s_n_ent;
    for (unsigned int k = threadIdx.x; k < 2 * cnt; k += blockDim.x) A.trans_ent[2 * (int64_t)s_base + k] = s_ent[k];
    for (unsigned int k = threadIdx.x; k < cnt; k += blockDim.x) A.trans_list[s_base + k] = (int32_t)s_ent[2 * k].x;
    if (threadIdx.x < M_NK && s_cnt[threadIdx.x]) {
        unsigned long long* row = A.counters + (int64_t)t * CVB_N_COUNTERS;
        const int ids[M_NK] = {CVB_C_new_diagnoses, CVB_C_new_quarantined, CVB_C_new_isolated};
        atomicAdd(row + ids[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
    }
    flush_stock_delta(s_delta, A.counters + (int64_t)t * CVB_N_COUNTERS, A.vcounters + (int64_t)t * nv * CVB_N_VCOUNTERS, nv);
}

}  // namespace cvb

using namespace cvb;

// ---- launch helpers -----------------------------------------------------------------------------------------------------------
static int fused_ready(cvb_sim* s, const char* who) {
    CVB_REQUIRE(s && s->pars_set, "%s: handle not ready", who);
    for (int f = 0; f < CVB_N_FIELDS; ++f) CVB_REQUIRE(s->people.f[f] != nullptr, "%s: people field %d is not bound", who, f);
    CVB_REQUIRE(s->res.counters && s->res.vcounters && s->res.sums, "%s: result tables are not bound", who);
    CVB_REQUIRE(s->log.count, "%s: infection log is not bound", who);
    CVB_REQUIRE(!s->partitioned || (s->pars.n_beds_hosp < 0 && s->pars.n_beds_icu < 0), "%s: bed limits of an agent-partitioned run need the per-step path (global counts)", who);
    CVB_REQUIRE(s->n % 4 == 0, "%s: the fused pipeline needs a population size that is a multiple of 4", who);
    uintptr_t all = 0;
    for (int f = 0; f < CVB_N_FIELDS; ++f) all |= (uintptr_t)s->people.f[f];
    CVB_REQUIRE((all & 15) == 0, "%s: People arrays must be 16-byte aligned", who);
    CVB_REQUIRE(!s->pars.use_waning || s->nab_kin, "%s: NAb kinetics table not set (cvb_set_nab_kin)", who);
    return 0;
}

constexpr size_t kStockBaseBytes = (CVB_N_COUNTERS + CVB_MAX_VARIANTS * CVB_N_VCOUNTERS) * sizeof(unsigned long long);

static int ensure_fused_buffers(cvb_sim* s) {
    if (!s->state) {
        CVB_CHECK(cudaMalloc((void**)&s->state, (size_t)s->n * sizeof(uint32_t)));
        s->state_valid = 0;
    }
    if (!s->trans_ent) CVB_CHECK(cudaMalloc((void**)&s->trans_ent, (size_t)s->n * 2 * sizeof(uint4)));
    if (!s->case_ent) CVB_CHECK(cudaMalloc((void**)&s->case_ent, (size_t)s->n * sizeof(uint4)));
    if (!s->stock_base) CVB_CHECK(cudaMalloc((void**)&s->stock_base, kStockBaseBytes));
    return 0;
}

constexpr int64_t kBeginChunk = 1024;      // agents per CTA of day_begin_kernel

// layout of the scratch rows that hold the absolute stock counts after a pack (the base of the first fused day)
static unsigned long long* base_row(cvb_sim* s) { return s->stock_base; }
static unsigned long long* base_vrow(cvb_sim* s) { return s->stock_base + CVB_N_COUNTERS; }

// pack (verify = false) or verify (true: returns the number of agents whose word differs from the arrays in host_out[1])
static int pack_or_check(cvb_sim* s, int32_t t_done, bool verify, int64_t* host_out, cudaStream_t st) {
    unsigned int* scal = reinterpret_cast<unsigned int*>(s->dev_scalars + 16);       // [0] violations, [1] mismatches, then 8 x 3 int32
    CVB_CHECK(cudaMemsetAsync(scal, 0, 2 * sizeof(unsigned int) + 24 * sizeof(int32_t), st));
    pack_state_kernel<<<grid_for(s->n, kThreads, 148 * 8), kThreads, 0, st>>>(s->people, s->state, verify ? s->state : nullptr, s->n, s->nv, t_done,
        s->quar_ring, s->quar_horizon, scal, scal + 1, reinterpret_cast<int32_t*>(scal + 2));
    CVB_LAUNCH_CHECK();
    if (!verify) {
        CVB_CHECK(cudaMemsetAsync(s->stock_base, 0, kStockBaseBytes, st));
        count_state_kernel<<<grid_for(s->n, kThreads, 148 * 4), kThreads, 0, st>>>(s->state, s->n, s->nv, base_row(s), base_vrow(s));
        CVB_LAUNCH_CHECK();
    }
    if (host_out) {
        unsigned int h[26];
        CVB_CHECK(cudaMemcpyAsync(h, scal, sizeof(h), cudaMemcpyDeviceToHost, st));
        CVB_CHECK(cudaStreamSynchronize(st));
        host_out[0] = h[0]; host_out[1] = h[1];
        for (int k = 0; k < 24; ++k) host_out[2 + k] = (int32_t)h[2 + k];
    }
    return 0;
}

template <bool END, bool PRE>
static int launch_day_begin(cvb_sim* s, int32_t t, bool test, bool tsel, bool base_from_pack, cudaStream_t st) {
    DayBeginArgs A;
    memset(&A, 0, sizeof(A));
    A.n = s->n; A.id0 = s->partitioned ? s->id0 : 0; A.chunk = kBeginChunk; A.t = t; A.nv = s->nv; A.waning = s->pars.use_waning; A.vaxpars = s->pars.has_vaccine_pars;
    A.nab_kin = s->nab_kin; A.nab_kin_len = s->nab_kin_len;
    A.counters = s->res.counters; A.vcounters = s->res.vcounters;
    if (PRE) {
        CVB_REQUIRE(base_from_pack || t > 0, "cvb_run_days: no stock counts to start day 0 from");
        A.base_row = base_from_pack ? base_row(s) : s->res.counters + (int64_t)(t - 1) * CVB_N_COUNTERS;
        A.base_vrow = base_from_pack ? base_vrow(s) : s->res.vcounters + (int64_t)(t - 1) * s->nv * CVB_N_VCOUNTERS;
    }
    const int threads = s->tune[0] > 0 ? s->tune[0] : kThreads;
    A.chunk = s->tune[1] > 0 ? s->tune[1] : kBeginChunk;
    const int grid = (int)((s->n + A.chunk - 1) / A.chunk);
    if (ensure_f64(&s->partial, &s->partial_cap, (int64_t)grid * 3)) return 1;
    A.partial = s->partial;
    s->begin_grid = grid;
    A.n_trans = s->n_trans; A.n_case = s->n_case_list;
    A.seed = s->seed;
    if (test) {
        A.tp = s->plan->test;
        A.test_plain = (A.tp.symp_quar_prob == A.tp.symp_prob && A.tp.asymp_quar_prob == A.tp.asymp_prob) ? 1 : 0;
    }
    A.adj_ptr = s->adj_ptr; A.case_ent = s->case_ent;
    A.case_bits = s->partitioned ? s->case_bits_local : nullptr;
#define CVB_DB(T1, T2) CVB_CHECK(launch_pdl(day_begin_kernel<END, PRE, T1, T2>, grid, threads, 0, st, s->people, s->state, s->pars, A))
    if constexpr (PRE) {
        if (test && tsel) CVB_DB(true, true);
        else if (test) CVB_DB(true, false);
        else if (tsel) CVB_DB(false, true);
        else CVB_DB(false, false);
    } else {
        CVB_DB(false, false);
    }
#undef CVB_DB
    CVB_LAUNCH_CHECK();
    return 0;
}

static int ensure_records_fused(cvb_sim* s, bool& dense_any) {
    // same bookkeeping as people_kernels.cu:ensure_records (which layers the dense pass reads today)
    CVB_REQUIRE(s->pars.n_layers >= 1, "cvb_run_days: no contact layers");
    if (!s->rec_store) CVB_CHECK(cudaMalloc((void**)&s->rec_store, (size_t)s->n * sizeof(float4)));
    s->rec.sus_imm = (const float*)s->people.f[CVB_F_sus_imm];
    uint32_t dense = 0, nonempty = 0;
    for (int l = 0; l < s->pars.n_layers; ++l) {
        if (s->layers[l].n_edges > 0) nonempty |= 1u << l;
        if (s->layers[l].n_edges > 0 && !(s->adj && ((s->adj_layer_mask >> l) & 1u))) dense |= 1u << l;
    }
    if (s->partitioned) dense = 0;                          // every layer is covered by the partitioned adjacency
    dense_any = dense != 0;
    const uint32_t ts8_layers = s->nv == 1 ? dense : 0u;
    if (ts8_layers && (!s->ts8_store || s->ts8_layers < s->pars.n_layers)) {
        cudaFree(s->ts8_store);
        s->ts8_store = nullptr;
        CVB_CHECK(cudaMalloc((void**)&s->ts8_store, (size_t)s->pars.n_layers * s->n * sizeof(float2)));
        s->ts8_layers = s->pars.n_layers;
    }
    s->rec.ts8 = ts8_layers ? s->ts8_store : nullptr;
    s->rec.ts8_mask = ts8_layers;
    s->rec.rec = (ts8_layers && ts8_layers == nonempty) ? nullptr : s->rec_store;
    s->rec_layers = s->pars.n_layers;
    return 0;
}

static int launch_day_mid(cvb_sim* s, int32_t t, bool closes_previous, cudaStream_t st) {
    bool dense_any = false;
    if (ensure_records_fused(s, dense_any)) return 1;
    DayMidArgs A;
    memset(&A, 0, sizeof(A));
    A.n = s->n; A.t = t; A.nv = s->nv; A.horizon = s->quar_horizon; A.dense = dense_any ? 1 : 0;
    A.quar_slot = s->quar_ring + (int64_t)(t % s->quar_horizon) * s->n;
    A.counters = s->res.counters; A.vcounters = s->res.vcounters; A.rec = s->rec; A.inf_bits = s->inf_bits;
    A.adj_ptr = (s->adj && s->adj_layer_mask) ? s->adj_ptr : nullptr;
    A.trans_ent = s->trans_ent; A.n_trans = s->n_trans; A.n_cand = s->n_cand; A.n_case = s->n_case_list; A.trans_list = s->trans_list;
    A.log_count = s->log.count; A.log_base = s->dev_scalars + 40;
    if (s->partitioned) { A.codes = s->codes_local; A.base_trans = s->rel_trans_global + s->id0; A.part_flags = s->part_flags; }
    A.partial = s->partial; A.n_part = s->begin_grid; A.t_end = closes_previous ? t - 1 : -1; A.sums = s->res.sums;
    A.chunk = s->tune[3] > 0 && s->tune[3] <= kMidChunk ? s->tune[3] : kMidChunk;
    const int threads = s->tune[2] > 0 ? s->tune[2] : kThreads;
    CVB_CHECK(launch_pdl(day_mid_kernel, (int)((s->n + A.chunk - 1) / A.chunk), threads, 0, st, s->people, s->state, s->pars, A));
    CVB_LAUNCH_CHECK();
    return 0;
}

// optional per-kernel timing of the day loop (bench.py's roofline table): CUDA events around every launch, summed when read
namespace cvb {
struct FusedTiming {
    std::vector<cudaEvent_t> ev[CVB_N_TIMED];        // (begin, end) pairs
    double ms[CVB_N_TIMED];
    long long launches[CVB_N_TIMED];
};
}
struct TimedScope {
    cvb::FusedTiming* tm; int kind; cudaStream_t st; cudaEvent_t e1;
    TimedScope(cvb_sim* s, int kind_, cudaStream_t st_) : tm(s->timing), kind(kind_), st(st_), e1(nullptr) {
        if (!tm) return;
        cudaEvent_t e0;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        tm->ev[kind].push_back(e0); tm->ev[kind].push_back(e1);
    }
    ~TimedScope() { if (tm) cudaEventRecord(e1, st); }
};

extern "C" {

int cvb_timing_enable(cvb_sim* s, int32_t on) {
    CVB_REQUIRE(s, "cvb_timing_enable: NULL handle");
    if (on && !s->timing) {
        s->timing = new (std::nothrow) cvb::FusedTiming();
        CVB_REQUIRE(s->timing, "cvb_timing_enable: out of host memory");
        for (int k = 0; k < CVB_N_TIMED; ++k) { s->timing->ms[k] = 0.0; s->timing->launches[k] = 0; }
    } else if (!on && s->timing) {
        for (int k = 0; k < CVB_N_TIMED; ++k) for (cudaEvent_t e : s->timing->ev[k]) cudaEventDestroy(e);
        delete s->timing;
        s->timing = nullptr;
    }
    return 0;
}

int cvb_timing_read(cvb_sim* s, double* host_ms, int64_t* host_launches) {
    CVB_REQUIRE(s && host_ms && host_launches, "cvb_timing_read: NULL argument");
    CVB_REQUIRE(s->timing, "cvb_timing_read: timing is not enabled (cvb_timing_enable)");
    CVB_CHECK(cudaDeviceSynchronize());
    cvb::FusedTiming* tm = s->timing;
    for (int k = 0; k < CVB_N_TIMED; ++k) {
        for (size_t j = 0; j + 1 < tm->ev[k].size(); j += 2) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, tm->ev[k][j], tm->ev[k][j + 1]) == cudaSuccess) { tm->ms[k] += ms; tm->launches[k] += 1; }
            cudaEventDestroy(tm->ev[k][j]); cudaEventDestroy(tm->ev[k][j + 1]);
        }
        tm->ev[k].clear();
        host_ms[k] = tm->ms[k]; host_launches[k] = tm->launches[k];
        tm->ms[k] = 0.0; tm->launches[k] = 0;
    }
    return 0;
}

int cvb_plan_clear(cvb_sim* s) {
    CVB_REQUIRE(s, "cvb_plan_clear: NULL handle");
    if (!s->plan) { s->plan = new (std::nothrow) DayPlan(); CVB_REQUIRE(s->plan, "cvb_plan_clear: out of host memory"); memset(s->plan, 0, sizeof(DayPlan)); }
    for (int k = 0; k < 4; ++k) free(s->plan->vacc_days[k]);
    memset(s->plan, 0, sizeof(DayPlan));
    return 0;
}

int cvb_plan_vaccinate(cvb_sim* s, const cvb_vaccinate_pars* vp, const uint8_t* host_day_flags, int32_t* iv_doses, int32_t* due_day) {
    CVB_REQUIRE(s && vp && host_day_flags && iv_doses && due_day, "cvb_plan_vaccinate: NULL argument");
    if (!s->plan && cvb_plan_clear(s)) return 1;
    CVB_REQUIRE(s->plan->n_vacc < 4, "cvb_plan_vaccinate: the day plan holds at most four vaccination interventions");
    CVB_REQUIRE(vp->vaccine_index >= 0 && vp->vaccine_index < CVB_MAX_VACCINES, "cvb_plan_vaccinate: vaccine index out of range");
    const int k = s->plan->n_vacc++;
    s->plan->vacc[k] = *vp;
    s->plan->vacc_days[k] = (uint8_t*)malloc((size_t)s->npts);
    CVB_REQUIRE(s->plan->vacc_days[k], "cvb_plan_vaccinate: out of host memory");
    memcpy(s->plan->vacc_days[k], host_day_flags, (size_t)s->npts);
    s->plan->vacc_doses[k] = iv_doses; s->plan->vacc_due[k] = due_day;
    return 0;
}

int cvb_plan_test_prob(cvb_sim* s, const cvb_test_prob_pars* tp, int32_t start_day, int32_t end_day) {
    CVB_REQUIRE(s && tp, "cvb_plan_test_prob: NULL argument");
    if (!s->plan && cvb_plan_clear(s)) return 1;
    CVB_REQUIRE(!s->plan->has_test, "cvb_plan_test_prob: the day plan holds one test_prob intervention");
    CVB_REQUIRE(!s->plan->has_trace, "cvb_plan_test_prob: testing must be registered before tracing (the order interventions are applied in)");
    s->plan->has_test = 1; s->plan->test = *tp; s->plan->test_start = start_day; s->plan->test_end = end_day;
    return 0;
}

int cvb_plan_contact_tracing(cvb_sim* s, const cvb_trace_pars* tr, int32_t start_day, int32_t end_day) {
    CVB_REQUIRE(s && tr, "cvb_plan_contact_tracing: NULL argument");
    if (!s->plan && cvb_plan_clear(s)) return 1;
    CVB_REQUIRE(!s->plan->has_trace, "cvb_plan_contact_tracing: the day plan holds one contact_tracing intervention");
    CVB_REQUIRE(!tr->presumptive, "cvb_plan_contact_tracing: presumptive tracing is not part of the fused day (use cvb_contact_tracing)");
    s->plan->has_trace = 1; s->plan->trace = *tr; s->plan->trace_start = start_day; s->plan->trace_end = end_day;
    return 0;
}

int cvb_plan_dynamic_layers(cvb_sim* s, uint32_t layer_mask) {
    CVB_REQUIRE(s, "cvb_plan_dynamic_layers: NULL handle");
    if (!s->plan && cvb_plan_clear(s)) return 1;
    s->plan->regen_mask = layer_mask;
    return 0;
}

int cvb_tune(cvb_sim* s, int32_t what, int32_t value) {
    CVB_REQUIRE(s && what >= 0 && what < 8, "cvb_tune: bad argument");
    if (what == 0 || what == 2) CVB_REQUIRE(value == 0 || (value >= 128 && value <= kThreads && value % 32 == 0), "cvb_tune: CTA size must be a multiple of 32 in [128, %d]", kThreads);
    if (what == 1) CVB_REQUIRE(value == 0 || (value >= 32 && value % 32 == 0), "cvb_tune: chunk must be a multiple of 32");
    if (what == 3) CVB_REQUIRE(value == 0 || (value >= 32 && value % 32 == 0 && value <= kMidChunk), "cvb_tune: chunk must be a multiple of 32 up to %d", kMidChunk);
    s->tune[what] = value;
    return 0;
}

int cvb_state_invalidate(cvb_sim* s) {
    CVB_REQUIRE(s, "cvb_state_invalidate: NULL handle");
    s->state_valid = 0;
    return 0;
}

int cvb_state_check(cvb_sim* s, int32_t t_done, int64_t* host_out26, cvb_stream st) {
    CVB_REQUIRE(s && host_out26, "cvb_state_check: NULL argument");
    CVB_REQUIRE(s->state && s->state_valid, "cvb_state_check: the packed state is not valid (nothing to check)");
    return pack_or_check(s, t_done, true, host_out26, (cudaStream_t)st);
}

// the registered vaccinate_prob interventions that act on day t (list order; after testing and tracing, before update_states_post)
static int run_vaccinations(cvb_sim* s, int32_t t, cudaStream_t st) {
    DayPlan& plan = *s->plan;
    for (int k = 0; k < plan.n_vacc; ++k) {
        const uint8_t f = plan.vacc_days[k][t];
        if (!f) continue;
        cvb_vaccinate_pars vp = plan.vacc[k];
        vp.first_dose_today = f & 1; vp.second_dose_today = (f >> 1) & 1;
        TimedScope ts(s, CVB_TIMED_vaccinate, st);
        if (launch_vaccinate_fused(s, t, &vp, plan.vacc_doses[k], plan.vacc_due[k], st)) return 1;
    }
    return 0;
}

// cvb_run_days in three parts, so that several handles can be advanced in lockstep from one host thread (cvb_run_days_multi)
static int run_days_begin(cvb_sim* s, int32_t t0, int32_t t1, cudaStream_t st, bool& packed) {
    if (fused_ready(s, "cvb_run_days")) return 1;
    CVB_REQUIRE(t0 >= 0 && t0 < t1 && t1 <= s->npts, "cvb_run_days: days [%d,%d) outside [0,%d)", t0, t1, s->npts);
    if (!s->plan && cvb_plan_clear(s)) return 1;
    const DayPlan& plan = *s->plan;
    if (ensure_fused_buffers(s)) return 1;
    if (plan.has_trace && !s->partitioned) {
        CVB_REQUIRE(s->adj && s->adj_layer_mask, "cvb_run_days: contact tracing in the fused day needs the adjacency (cvb_bind_adjacency)");
        for (int l = 0; l < s->pars.n_layers; ++l)
            CVB_REQUIRE(!(plan.trace.trace_prob[l] > 0.0) || s->layers[l].n_edges == 0 || ((s->adj_layer_mask >> l) & 1u),
                        "cvb_run_days: traced layer %d is not covered by the adjacency", l);
    }
    packed = false;
    if (!s->state_valid || t0 == 0) {
        int64_t out[26];
        packed = true;
        if (pack_or_check(s, t0 - 1, false, out, st)) return 1;
        CVB_REQUIRE(out[0] == 0, "cvb_run_days: %lld agents are in a state the packed word cannot express (several by-variant rows set, or NAbs without a peak)", (long long)out[0]);
        s->state_valid = 1;
    }
    CVB_CHECK(cudaMemsetAsync(s->n_trans, 0, sizeof(unsigned int), st));
    CVB_CHECK(cudaMemsetAsync(s->n_case_list, 0, sizeof(unsigned int), st));
    CVB_CHECK(cudaMemsetAsync(s->n_cand, 0, sizeof(unsigned int), st));
    return 0;
}

static int run_one_day(cvb_sim* s, int32_t t, int32_t t0, bool packed, cudaStream_t st) {
    const DayPlan& plan = *s->plan;
    s->last_t = t;
    for (int l = 0; l < s->pars.n_layers; ++l)
        if ((plan.regen_mask >> l) & 1u) { TimedScope ts(s, CVB_TIMED_regen, st); if (cvb_layer_regenerate(s, l, t, (cvb_stream)st)) return 1; }
    const bool test = plan.has_test && t >= plan.test_start && (plan.test_end < 0 || t <= plan.test_end);
    const bool trace = plan.has_trace && t >= plan.trace_start && (plan.trace_end < 0 || t <= plan.trace_end);
    {
        TimedScope ts(s, CVB_TIMED_day_begin, st);
        int rc = t > t0 ? launch_day_begin<true, true>(s, t, test, trace, false, st) : launch_day_begin<false, true>(s, t, test, trace, packed, st);
        if (rc) return rc;
    }
    if (trace) { TimedScope ts(s, CVB_TIMED_trace, st); if (launch_trace_sparse2(s, t, &plan.trace, st)) return 1; }
    if (run_vaccinations(s, t, st)) return 1;
    { TimedScope ts(s, CVB_TIMED_day_mid, st); if (launch_day_mid(s, t, t > t0, st)) return 1; }
    { TimedScope ts(s, CVB_TIMED_edge_pass, st); if (edge_pass_impl(s, t, st, true)) return 1; }
    { TimedScope ts(s, CVB_TIMED_infect, st); if (launch_infect_winners(s, t, true, st)) return 1; }
    return 0;
}

static int run_days_end(cvb_sim* s, int32_t t1, cudaStream_t st) {
    s->state_valid = 1;                                     // (the edge pass / layer regeneration do not touch People flags)
    TimedScope ts(s, CVB_TIMED_day_end, st);
    if (launch_day_begin<true, false>(s, t1, false, false, false, st)) return 1;      // closes day t1 - 1 (its argument is the day AFTER the one it closes)
    sum_partials_kernel<<<1, 96, 0, st>>>(s->partial, s->begin_grid, s->res.sums, t1 - 1, -1);
    CVB_LAUNCH_CHECK();
    return 0;
}

// One day of an AGENT-PARTITIONED simulation through the fused kernels, in the phases between which the host exchanges data
// (partition.py): 0 = day_begin (first_of_block: the state words are rebuilt if needed; the local case bitmap is cleared first),
// -> all-gather of the case bitmap on tracing days -> 1 = notify the local contacts of every global case -> 2 = day_mid (writes the
// 1-byte transmit codes) -> all-gather of the codes -> 3 = edge pass + infect; 4 = close day t - 1 at the end of a block.
int cvb_fused_phase(cvb_sim* s, int32_t t, int32_t phase, int32_t first_of_block, cvb_stream st_) {
    cudaStream_t st = (cudaStream_t)st_;
    CVB_REQUIRE(s && s->partitioned, "cvb_fused_phase: for agent-partitioned handles (use cvb_run_days)");
    CVB_REQUIRE(phase >= 0 && phase <= 4, "cvb_fused_phase: phase %d out of range", phase);
    if (phase == 4) return run_days_end(s, t, st);
    CVB_REQUIRE(t >= 0 && t < s->npts, "cvb_fused_phase: day %d outside [0,%d)", t, s->npts);
    const DayPlan& plan = *s->plan;
    const bool test = plan.has_test && t >= plan.test_start && (plan.test_end < 0 || t <= plan.test_end);
    const bool trace = plan.has_trace && t >= plan.trace_start && (plan.trace_end < 0 || t <= plan.trace_end);
    switch (phase) {
        case 0: {
            bool packed = false;
            if (first_of_block) {
                if (run_days_begin(s, t, t + 1, st, packed)) return 1;
                s->block_packed = packed;
            }
            s->last_t = t;
            if (trace) CVB_CHECK(cudaMemsetAsync(s->case_bits_local, 0, (size_t)(s->chunk / 32) * sizeof(unsigned int), st));
            TimedScope ts(s, CVB_TIMED_day_begin, st);
            return first_of_block ? launch_day_begin<false, true>(s, t, test, trace, s->block_packed != 0, st)
                                  : launch_day_begin<true, true>(s, t, test, trace, false, st);
        }
        case 1: { TimedScope ts(s, CVB_TIMED_trace, st); return trace ? launch_trace_partition(s, t, &plan.trace, st) : 0; }
        case 2: {
            if (run_vaccinations(s, t, st)) return 1;
            TimedScope ts(s, CVB_TIMED_day_mid, st);
            return launch_day_mid(s, t, !first_of_block, st);
        }
        default: {
            { TimedScope ts(s, CVB_TIMED_edge_pass, st); if (edge_pass_impl(s, t, st, true)) return 1; }
            TimedScope ts(s, CVB_TIMED_infect, st);
            return launch_infect_winners(s, t, true, st);
        }
    }
}

int cvb_run_days(cvb_sim* s, int32_t t0, int32_t t1, cvb_stream st_) {
    cudaStream_t st = (cudaStream_t)st_;
    CVB_REQUIRE(s && !s->partitioned, "cvb_run_days: agent-partitioned handles advance phase by phase (cvb_fused_phase)");
    bool packed = false;
    if (run_days_begin(s, t0, t1, st, packed)) return 1;
    for (int32_t t = t0; t < t1; ++t)
        if (run_one_day(s, t, t0, packed, st)) return 1;
    return run_days_end(s, t1, st);
}

// members [a, b) through days [t0, t1): days outermost, members innermost, so that one host thread keeps every one of its members' streams fed
static int run_days_slice(cvb_sim** handles, cvb_stream* streams, int a, int b, int32_t t0, int32_t t1) {
    std::vector<char> packed((size_t)(b - a), 0);
    int dev = -1;
    auto on_device = [&](int m) { if (handles[m]->device != dev) { dev = handles[m]->device; return cudaSetDevice(dev); } return cudaSuccess; };
    for (int m = a; m < b; ++m) {
        CVB_CHECK(on_device(m));
        bool p = false;
        if (run_days_begin(handles[m], t0, t1, (cudaStream_t)streams[m], p)) return 1;
        packed[m - a] = p;
    }
    for (int32_t t = t0; t < t1; ++t)
        for (int m = a; m < b; ++m) {
            CVB_CHECK(on_device(m));
            if (run_one_day(handles[m], t, t0, packed[m - a] != 0, (cudaStream_t)streams[m])) return 1;
        }
    for (int m = a; m < b; ++m) {
        CVB_CHECK(on_device(m));
        if (run_days_end(handles[m], t1, (cudaStream_t)streams[m])) return 1;
    }
    return 0;
}

int cvb_run_days_multi(cvb_sim** handles, int32_t n_handles, int32_t t0, int32_t t1, cvb_stream* streams) {
    CVB_REQUIRE(handles && streams && n_handles > 0, "cvb_run_days_multi: bad argument");
    CVB_REQUIRE(n_handles <= 65536, "cvb_run_days_multi: too many handles");
    for (int m = 0; m < n_handles; ++m) CVB_REQUIRE(handles[m], "cvb_run_days_multi: NULL handle %d", m);
    // Small members are bound by the HOST's launch rate (five launches of 3-6 us kernels per member-day, ~2.4 us each from one thread), so the
    // members are split over a few host threads, each feeding its own members' streams (CVB_MULTI_THREADS overrides the number)
    int n_threads = n_handles >= 32 ? 4 : (n_handles >= 8 ? 2 : 1);
    if (const char* e = getenv("CVB_MULTI_THREADS")) { const int v = atoi(e); if (v >= 1) n_threads = v; }
    if (n_threads > n_handles) n_threads = n_handles;
    if (n_threads > 64) n_threads = 64;
    if (n_threads == 1) return run_days_slice(handles, streams, 0, n_handles, t0, t1);
    std::vector<int> rc((size_t)n_threads, 0);
    std::vector<std::string> err((size_t)n_threads);
    std::vector<std::thread> workers;
    for (int k = 0; k < n_threads; ++k) {
        const int a = (int)((int64_t)n_handles * k / n_threads), b = (int)((int64_t)n_handles * (k + 1) / n_threads);
        workers.emplace_back([&, k, a, b]() {
            rc[k] = run_days_slice(handles, streams, a, b, t0, t1);
            if (rc[k]) err[k] = cvb_last_error();               // (the message buffer is per thread)
        });
    }
    for (auto& w : workers) w.join();
    for (int k = 0; k < n_threads; ++k)
        if (rc[k]) { cvb::set_error("%s", err[k].c_str()); return rc[k]; }
    return 0;
}

}  // extern "C"
