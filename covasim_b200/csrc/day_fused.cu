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
#include <string.h>
#include <new>
#include <vector>
#include "cvb_internal.cuh"

namespace cvb {

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
__host__ __device__ __forceinline__ int sb_ebv(uint32_t s) { return (int)((s >> kEbvShift) & 15u); }
__host__ __device__ __forceinline__ int sb_rv(uint32_t s) { return (int)((s >> kRvShift) & 15u); }
constexpr uint32_t kEbvMask = 15u << kEbvShift, kRvMask = 15u << kRvShift;

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

// ================================================================================================================================
// day_begin_kernel
// ================================================================================================================================
enum { F_INFECTIOUS = 0, F_SYMPTOMATIC, F_SEVERE, F_CRITICAL, F_RECOVERIES, F_DEATHS, F_KNOWN_DEATHS, F_BED_SEVERE, F_BED_CRITICAL, F_TESTS, F_NK };
constexpr int kStockSlots = 16 + 1 + 2 * CVB_MAX_VARIANTS;      // 16 state bits, alive, exposed / infectious by variant
constexpr int kImmQueueCap2 = 64;

struct DayBeginArgs {
    int64_t n, id0;
    int32_t t;                          // the day whose update_states_pre / interventions run; the END part closes day t - 1
    int32_t nv, waning, vaxpars;
    const double* nab_kin; int64_t nab_kin_len;
    unsigned long long* counters; unsigned long long* vcounters; unsigned long long* beds;
    double* partial; unsigned int* ticket; double* sums;       // sums = table base [npts][4]
    unsigned int* n_trans; unsigned int* n_case;
    // test_prob
    cvb_test_prob_pars tp; int32_t test_plain;                 // test_plain: quarantine state does not change the probability
    uint64_t seed;
    // contact tracing: cases -> entries
    const long long* adj_ptr; uint4* case_ent;
};

// check_immunity for one queued agent x variant (immunity.py:303-350): float64, rounded once to float32.
// entry = {agent, nab bits, variant | natural-immunity source + 1 << 4 | vaccine source << 8 | vaccinated << 12, -}
__device__ __forceinline__ void immunity_eval2(const uint4 en, int64_t n, const cvb_pars& pars, float* __restrict__ sus_imm,
                                               float* __restrict__ symp_imm, float* __restrict__ sev_imm, double& sum_sus, double& sum_symp) {
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
    sum_sus += (double)s0;
    sum_symp += (double)s1;
}

__device__ __forceinline__ float4 ld4f(const float* p, int64_t i0) { return *reinterpret_cast<const float4*>(p + i0); }
__device__ __forceinline__ void unpack4(const float4 v, float o[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }

template <bool END, bool PRE, bool TEST, bool TSEL>
__global__ void __launch_bounds__(kThreads, 2) day_begin_kernel(PeoplePtrs P, uint32_t* __restrict__ S, const __grid_constant__ cvb_pars pars,
                                                               const __grid_constant__ DayBeginArgs A) {
    __shared__ int s_flow[F_NK + CVB_MAX_VARIANTS];
    __shared__ int s_stock[kStockSlots];
    __shared__ uint4 s_queue[(kThreads / 32) * kImmQueueCap2];
    __shared__ double s_sum[3][kThreads / 32];
    __shared__ bool s_last;
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
    for (int k = threadIdx.x; k < F_NK + CVB_MAX_VARIANTS; k += blockDim.x) s_flow[k] = 0;
    for (int k = threadIdx.x; k < kStockSlots; k += blockDim.x) s_stock[k] = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0 && PRE) *A.n_trans = 0;       // today's transmitter list starts empty (filled by day_mid_kernel)
    __syncthreads();
    int c[F_NK];
#pragma unroll
    for (int k = 0; k < F_NK; ++k) c[k] = 0;
    int cvn[CVB_MAX_VARIANTS];
#pragma unroll
    for (int k = 0; k < CVB_MAX_VARIANTS; ++k) cvn[k] = 0;
    int stock[17];                                                        // warp-uniform: state-bit counts + alive
#pragma unroll
    for (int k = 0; k < 17; ++k) stock[k] = 0;
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

    // warp-aligned loop: whole warps stay in it (ballots); n is a multiple of 4 (checked by the launcher)
    const int64_t n_groups = n >> 2;
    const int64_t n_groups_pad = (n_groups + 31) / 32 * 32;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_groups_pad; g += (int64_t)gridDim.x * blockDim.x) {
        const bool in = g < n_groups;
        const int64_t i0 = g << 2;
        uint4 sv = make_uint4(0u, 0u, 0u, 0u);
        if (in) sv = *reinterpret_cast<const uint4*>(S + i0);
        uint32_t s[4] = {sv.x, sv.y, sv.z, sv.w};
        const uint32_t any = sv.x | sv.y | sv.z | sv.w;

        // ---- loads this group needs, issued together ----
        float nb[4] = {0.0f, 0.0f, 0.0f, 0.0f}, pk[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        int32_t te[4] = {0, 0, 0, 0}, vs[4] = {0, 0, 0, 0};
        const bool g_nab = waning && (any & SB_HAS_NAB);
        if (g_nab) {
            unpack4(ld4f(nab, i0), nb);
            if (END) {
                unpack4(ld4f(PF(P, peak_nab), i0), pk);
                const int4 tv = *reinterpret_cast<const int4*>(PI(P, t_nab_event) + i0);
                te[0] = tv.x; te[1] = tv.y; te[2] = tv.z; te[3] = tv.w;
            }
            if (PRE && A.vaxpars && (any & SB_VACC)) {
                const int4 vv = *reinterpret_cast<const int4*>(PI(P, vaccine_source) + i0);
                vs[0] = vv.x; vs[1] = vv.y; vs[2] = vv.z; vs[3] = vv.w;
            }
        }
        float di[4], ds[4], dv[4], dc[4], dr[4], dd[4], dei[4];
        const bool g_exp = PRE && (any & SB_EXP);
        if (g_exp) {
            unpack4(ld4f(PF(P, date_infectious), i0), di); unpack4(ld4f(PF(P, date_symptomatic), i0), ds);
            unpack4(ld4f(PF(P, date_severe), i0), dv); unpack4(ld4f(PF(P, date_critical), i0), dc);
            unpack4(ld4f(PF(P, date_recovered), i0), dr); unpack4(ld4f(PF(P, date_dead), i0), dd);
        }
        const bool g_iso = PRE && (any & SB_ISO);
        if (g_iso) unpack4(ld4f(PF(P, date_end_isolation), i0), dei);
        float dq[4] = {qnan, qnan, qnan, qnan}, deq[4] = {qnan, qnan, qnan, qnan};
        if (TEST && !A.test_plain && in) {
            if (A.tp.quar_policy == 0 || A.tp.quar_policy == 2) unpack4(ld4f(PF(P, date_quarantined), i0), dq);
            if (A.tp.quar_policy == 1 || A.tp.quar_policy == 2) unpack4(ld4f(PF(P, date_end_quarantine), i0), deq);
        }

        // ---- close day t-1: stock counts (sim.py:652-664), update_nab (immunity.py:205-213), sum of NAbs over the living ----
        if (END) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t sk = s[k];
#pragma unroll
                for (int b = 0; b < 16; ++b) {
                    if (b == 1 || b == 7 || b == 12) continue;              // naive, tested, known_contact are not result stocks
                    stock[b] += __popc(__ballot_sync(0xFFFFFFFFu, (sk >> b) & 1u));
                }
                stock[16] += __popc(__ballot_sync(0xFFFFFFFFu, in && !(sk & SB_DEAD)));
                const unsigned m_e = __ballot_sync(0xFFFFFFFFu, (sk & kEbvMask) != 0u);
                if (m_e) {                                                   // by-variant stocks (exposed / infectious rows)
                    const int ev = sb_ebv(sk) - 1;
                    for (int v = 0; v < nv; ++v) {
                        const int ce = __popc(__ballot_sync(0xFFFFFFFFu, ev == v));
                        const int ci = __popc(__ballot_sync(0xFFFFFFFFu, ev == v && (sk & SB_IBV)));
                        if (lane == 0) { if (ce) atomicAdd(&s_stock[17 + 2 * v], ce); if (ci) atomicAdd(&s_stock[17 + 2 * v + 1], ci); }
                    }
                }
            }
            if (g_nab) {
                bool upd = false;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (s[k] & SB_HAS_NAB) {                                 // has_nabs = true(peak_nab)  (sim.py:666-669)
                        int64_t dt = (int64_t)(t - 1) - (int64_t)te[k];
                        if (dt < 0) dt += A.nab_kin_len;                     // NumPy negative index wraps
                        const double kin = (dt >= 0 && dt < A.nab_kin_len) ? A.nab_kin[dt] : 0.0;
                        nb[k] = nab_step(nb[k], pk[k], kin);
                        upd = true;
                    }
                    if (!(s[k] & SB_DEAD)) sum_nab += (double)nb[k];
                }
                if (upd) *reinterpret_cast<float4*>(nab + i0) = make_float4(nb[0], nb[1], nb[2], nb[3]);
            }
        }

        if (PRE) {
            // ---- update_states_pre (people.py:164-186) ----
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int64_t i = i0 + k;
                uint32_t sk = s[k];
                const bool was_exposed = (sk & SB_EXP) != 0;                 // is_exp is taken once, before any transition (people.py:169)
                const int ev = sb_ebv(sk) - 1;
                const float evf = (float)ev;
                if (was_exposed) {
                    if (!(sk & SB_INF) && due(di[k], t)) {                   // people.py:222-232
                        infectious[i] = 1; inf_var[i] = evf; sk |= SB_INF;
                        if (ev >= 0 && ev < nv) {
                            inf_by_var[(int64_t)ev * n + i] = 1; sk |= SB_IBV;
#pragma unroll
                            for (int q = 0; q < CVB_MAX_VARIANTS; ++q) cvn[q] += (q == ev);
                        }
                        ++c[F_INFECTIOUS];
                    }
                    if (!(sk & SB_SYMP) && due(ds[k], t)) { symptomatic[i] = 1; sk |= SB_SYMP; ++c[F_SYMPTOMATIC]; }     // people.py:235-253
                    if (!(sk & SB_SEV) && due(dv[k], t)) { severe[i] = 1; sk |= SB_SEV; ++c[F_SEVERE]; }
                    if (!(sk & SB_CRIT) && due(dc[k], t)) { critical[i] = 1; sk |= SB_CRIT; ++c[F_CRITICAL]; }
                    if (!(sk & SB_REC) && due(dr[k], t)) {                   // people.py:256-291
                        exposed[i] = 0; infectious[i] = 0; symptomatic[i] = 0; severe[i] = 0; critical[i] = 0;
                        recovered[i] = 1;
                        rec_var[i] = evf; inf_var[i] = qnan; exp_var[i] = qnan;
                        for (int v = 0; v < nv; ++v) { exp_by_var[(int64_t)v * n + i] = 0; inf_by_var[(int64_t)v * n + i] = 0; }
                        sk &= ~(SB_EXP | SB_INF | SB_SYMP | SB_SEV | SB_CRIT | SB_IBV | kEbvMask | kRvMask);
                        sk |= SB_REC;
                        if (ev >= 0 && ev < nv) sk |= (uint32_t)(ev + 1) << kRvShift;
                        if (waning) {
                            susceptible[i] = 1; diagnosed[i] = 0;
                            if (sk & SB_DIAG) sk |= SB_DPEND;                // the (old) date_diagnosed is still set
                            sk |= SB_SUS; sk &= ~SB_DIAG;
                        }
                        ++c[F_RECOVERIES];
                    }
                }
                if ((s[k] & SB_ISO) && due(dei[k], t)) { isolated[i] = 0; sk &= ~SB_ISO; }     // people.py:368-374
                if (was_exposed) {
                    if (!(sk & SB_DEAD) && due(dd[k], t)) {                  // people.py:294-312
                        dead[i] = 1;
                        if (sk & SB_DIAG) { known_dead[i] = 1; sk |= SB_KDEAD; ++c[F_KNOWN_DEATHS]; }
                        susceptible[i] = 0; exposed[i] = 0; infectious[i] = 0; symptomatic[i] = 0; severe[i] = 0; critical[i] = 0;
                        known_contact[i] = 0; quarantined[i] = 0; recovered[i] = 0;
                        inf_var[i] = qnan; exp_var[i] = qnan; rec_var[i] = qnan;
                        sk &= ~(SB_SUS | SB_EXP | SB_INF | SB_SYMP | SB_SEV | SB_CRIT | SB_KCONTACT | SB_QUAR | SB_REC | kRvMask);
                        sk |= SB_DEAD;
                        ++c[F_DEATHS];
                    }
                }
                c[F_BED_SEVERE] += (sk & SB_SEV) != 0;
                c[F_BED_CRITICAL] += (sk & SB_CRIT) != 0;
                s[k] = sk;
            }

            // ---- check_immunity (immunity.py:303-350): agents with antibodies and an immunity source are queued per warp and
            //      evaluated 32 at a time; everybody else has protection exactly 0 (written only if something non-zero is stored) ----
            if (waning) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int64_t i = i0 + k;
                    const uint32_t sk = s[k];
                    const int rv1 = sb_rv(sk);
                    const int vsi = vs[k];
                    const bool vacc = A.vaxpars && (sk & SB_VACC) && vsi >= 0 && vsi < CVB_MAX_VACCINES;
                    const bool heavy = (sk & SB_HAS_NAB) && nb[k] > 0.0f && (rv1 != 0 || vacc);
                    const bool clear = !heavy && (sk & SB_IMM_NZ);
                    const unsigned packed = ((unsigned)rv1 << 4) | ((unsigned)(vacc ? vsi : 0) << 8) | ((unsigned)vacc << 12);
                    for (int v = 0; v < nv; ++v) {
                        if (clear) {
                            sus_imm[(int64_t)v * n + i] = 0.0f; symp_imm[(int64_t)v * n + i] = 0.0f; sev_imm[(int64_t)v * n + i] = 0.0f;
                        }
                        const unsigned m = __ballot_sync(0xFFFFFFFFu, heavy);
                        if (m) {
                            if (heavy) q_imm[qn + __popc(m & lt_mask)] = make_uint4((unsigned)i, __float_as_uint(nb[k]), packed | (unsigned)v, 0u);
                            qn += __popc(m);
                            __syncwarp();
                            if (qn >= 32) {
                                qn -= 32;
                                const uint4 en = q_imm[qn + lane];
                                __syncwarp();
                                immunity_eval2(en, n, pars, sus_imm, symp_imm, sev_imm, sum_sus, sum_symp);
                            }
                        }
                    }
                    if (heavy) s[k] = sk | SB_IMM_NZ; else s[k] = sk & ~SB_IMM_NZ;
                }
            }
        }

        // ---- test_prob + People.test (interventions.py:921-981, people.py:589-617) and today's cases (interventions.py:1066-1085) ----
        if (TEST || TSEL) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int64_t i = i0 + k;
                uint32_t sk = s[k];
                if (!in) continue;
                float ddiag_new = qnan;
                bool ddiag_known = false;
                if (TEST && !(sk & SB_DIAG)) {                               // diagnosed people do not test (interventions.py:973)
                    const bool symp = (sk & SB_SYMP) != 0;
                    bool qt = false;                                         // interventions.py:691-715 get_quar_inds
                    if (!A.test_plain) {
                        switch (A.tp.quar_policy) {
                            case 0:  qt = dq[k] == tf - 1.0f; break;
                            case 1:  qt = deq[k] == tf + 1.0f; break;
                            case 2:  qt = (dq[k] == tf - 1.0f) || (deq[k] == tf + 1.0f); break;
                            default: qt = (sk & SB_QUAR) != 0; break;
                        }
                    }
                    const double prob = qt ? (symp ? A.tp.symp_quar_prob : A.tp.asymp_quar_prob) : (symp ? A.tp.symp_prob : A.tp.asymp_prob);
                    if (prob > 0.0 && keyed_uniform(A.seed, P_TEST, (uint32_t)A.tp.index, t, i + A.id0, 0) < prob) {
                        ++c[F_TESTS];
                        tested[i] = 1; sk |= SB_TESTED;
                        d_tested[i] = tf;
                        if ((sk & SB_INF) && keyed_uniform(A.seed, P_TEST_SENS, (uint32_t)A.tp.index, t, i + A.id0, 0) < A.tp.sensitivity) {
                            const float old = d_diag[i];
                            ddiag_known = true; ddiag_new = old;
                            if (is_nan(old) && keyed_uniform(A.seed, P_TEST_LOSS, (uint32_t)A.tp.index, t, i + A.id0, 0) < 1.0 - A.tp.loss_prob) {
                                ddiag_new = (float)(t + A.tp.test_delay);
                                d_diag[i] = ddiag_new;
                                d_pos[i] = tf;
                                sk |= SB_DPEND;
                            }
                        }
                    }
                }
                if (TSEL && (sk & SB_DPEND)) {                               // a case: date_diagnosed == t
                    const float dg = ddiag_known ? ddiag_new : d_diag[i];
                    if (dg == tf) {
                        const long long beg = A.adj_ptr[i], end = A.adj_ptr[i + 1];
                        A.case_ent[warp_append32(A.n_case)] = make_uint4((unsigned)i, (unsigned)(end - beg), (unsigned)(unsigned long long)beg, (unsigned)((unsigned long long)beg >> 32));
                    }
                }
                s[k] = sk;
            }
        }
        if (in && (PRE || TEST) && ((s[0] ^ sv.x) | (s[1] ^ sv.y) | (s[2] ^ sv.z) | (s[3] ^ sv.w)))
            *reinterpret_cast<uint4*>(S + i0) = make_uint4(s[0], s[1], s[2], s[3]);
    }
    if (PRE && waning && lane < qn) immunity_eval2(q_imm[lane], n, pars, sus_imm, symp_imm, sev_imm, sum_sus, sum_symp);

    // ---- counters: flows of day t, stocks of day t - 1 ----
    if (PRE || TEST) {
        reduce_counters(c, s_flow);
        reduce_counters(cvn, s_flow + F_NK);
    }
    if (END && lane == 0) {
#pragma unroll
        for (int b = 0; b < 17; ++b) if (stock[b]) atomicAdd(&s_stock[b], stock[b]);
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
        for (int wq = 0; wq < kThreads / 32; ++wq) v += s_sum[threadIdx.x][wq];
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
            case F_BED_SEVERE:   atomicAdd(A.beds + (int64_t)t * 2 + 0, uv); break;
            case F_BED_CRITICAL: atomicAdd(A.beds + (int64_t)t * 2 + 1, uv); break;
            case F_TESTS:        atomicAdd(row + CVB_C_new_tests, uv); break;
            default: {
                const int var = threadIdx.x - F_NK;
                if (var < nv) atomicAdd(A.vcounters + ((int64_t)t * nv + var) * CVB_N_VCOUNTERS + CVB_VC_new_infectious_by_variant, uv);
            }
        }
    }
    if (END && threadIdx.x < kStockSlots && s_stock[threadIdx.x]) {
        const int k = threadIdx.x;
        const unsigned long long v = (unsigned long long)s_stock[k];
        unsigned long long* row = A.counters + (int64_t)(t - 1) * CVB_N_COUNTERS;
        // state bit -> stock counter (defaults.result_stocks order: susceptible, exposed, infectious, symptomatic, severe, critical,
        // recovered, dead, diagnosed, known_dead, quarantined, isolated, vaccinated)
        const int stock_of_bit[16] = {0, -1, 1, 2, 3, 4, 5, -1, 8, 6, 9, 7, -1, 10, 11, 12};
        if (k < 16) { if (stock_of_bit[k] >= 0) atomicAdd(row + CVB_C_n_susceptible + stock_of_bit[k], v); }
        else if (k == 16) atomicAdd(row + CVB_C_n_alive_agents, v);
        else {
            const int q = k - 17, var = q >> 1;
            if (var < nv) atomicAdd(A.vcounters + ((int64_t)(t - 1) * nv + var) * CVB_N_VCOUNTERS + ((q & 1) ? CVB_VC_n_infectious_by_variant : CVB_VC_n_exposed_by_variant), v);
        }
    }
    // the last CTA to finish adds up the per-CTA partial sums in a fixed order (deterministic float64 sums, no second launch)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(A.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x < 96) {
        __threadfence();
        const int q = threadIdx.x >> 5;
        double v = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) v += __ldcg(A.partial + (int64_t)b * 3 + q);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
        if (lane == 0) {
            if (q == 0) { if (END) A.sums[(int64_t)(t - 1) * 4 + 0] = v; }
            else if (PRE) A.sums[(int64_t)t * 4 + q] = v;
        }
        if (threadIdx.x == 0) *A.ticket = 0;
    }
}

// ================================================================================================================================
// day_mid_kernel: update_states_post + prepare_transmission
// ================================================================================================================================
enum { M_DIAGNOSES = 0, M_QUARANTINED, M_ISOLATED, M_NK };

struct DayMidArgs {
    int64_t n;
    int32_t t, nv, horizon, dense;                  // dense: some layer is streamed densely today (records for everyone + ts8 + bitmap)
    float* quar_slot;
    unsigned long long* counters;
    TransRecords rec;
    unsigned int* inf_bits;
    const long long* adj_ptr;                       // NULL: no adjacency (every layer dense)
    uint4* trans_ent; unsigned int* n_trans;
    unsigned int* n_cand; unsigned int* n_case;
    int32_t* trans_list;                            // the plain list (used by nobody in the fused pipeline; kept for cvb_get_edge_work parity)
};

__global__ void __launch_bounds__(kThreads, 3) day_mid_kernel(PeoplePtrs P, uint32_t* __restrict__ S, const __grid_constant__ cvb_pars pars,
                                                             const __grid_constant__ DayMidArgs A) {
    __shared__ int s_cnt[M_NK];
    if (threadIdx.x < M_NK) s_cnt[threadIdx.x] = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) { *A.n_cand = 0; *A.n_case = 0; }   // today's candidates start empty; the case list was consumed
    __syncthreads();
    int c[M_NK] = {0, 0, 0};
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

    const int64_t n_groups = n >> 2;
    const int64_t n_groups_pad = (n_groups + 31) / 32 * 32;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_groups_pad; g += (int64_t)gridDim.x * blockDim.x) {
        const bool in = g < n_groups;
        const int64_t i0 = g << 2;
        uint4 sv = make_uint4(0u, 0u, 0u, 0u);
        if (in) sv = *reinterpret_cast<const uint4*>(S + i0);
        uint32_t s[4] = {sv.x, sv.y, sv.z, sv.w};
        const uint32_t any = sv.x | sv.y | sv.z | sv.w;
        float pend[4] = {-1.0f, -1.0f, -1.0f, -1.0f};
        if (any & SB_QPEND) unpack4(ld4f(A.quar_slot, i0), pend);
        float rs4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        // rel_sus is needed unless all four stored records are certainly still right (simple form, same flags, and nothing that
        // update_states_post could change today: no pending request / diagnosis, not in quarantine)
        bool need_rs = A.dense != 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) need_rs |= in && !((s[k] & SB_RS_VALID) && !(s[k] & (SB_INF | SB_IMM_NZ | SB_QPEND | SB_DPEND | SB_QUAR)) &&
                                                      (((s[k] & SB_SUS) != 0) == ((s[k] & SB_RS_SUS) != 0)) && (((s[k] & SB_QUAR) != 0) == ((s[k] & SB_RS_QUAR) != 0)));
        if (need_rs) unpack4(ld4f(rel_sus, i0), rs4);
        unsigned inf_nibble = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = i0 + k;
            if (!in) continue;
            uint32_t sk = s[k];
            // ---- update_states_post (people.py:189-196) ----
            float ddiag = qnan;
            if (sk & SB_DPEND) {                                             // check_diagnosed (people.py:315-332)
                ddiag = d_diag[i];
                const float dpos = d_pos[i];
                if (!(sk & SB_DIAG)) {
                    if (due(dpos, t)) { d_pos[i] = qnan; ++c[M_DIAGNOSES]; }
                    if (due(ddiag, t)) { diagnosed[i] = 1; sk |= SB_DIAG; sk &= ~SB_DPEND; }
                }
            }
            bool quar = (sk & SB_QUAR) != 0;
            float end_q = qnan;
            bool end_q_known = false;
            if (pend[k] >= 0.0f) {                                           // check_quar (people.py:335-358)
                A.quar_slot[i] = -1.0f;
                if (quar) {
                    end_q = d_end_quar[i]; end_q_known = true;
                    if (pend[k] > end_q) { end_q = pend[k]; d_end_quar[i] = end_q; }      // Python max(old, requested)
                } else if (!(sk & (SB_DEAD | SB_REC | SB_DIAG | SB_ISO))) {
                    quarantined[i] = 1; quar = true;
                    d_quar[i] = tf;
                    end_q = pend[k]; end_q_known = true;
                    d_end_quar[i] = end_q;
                    ++c[M_QUARANTINED];
                }
            }
            if ((sk & SB_QPEND) && A.horizon == 1) sk &= ~SB_QPEND;          // (several slots: the bit stays, the slot is re-read)
            if (quar) {
                if (!end_q_known) end_q = d_end_quar[i];
                if (ddiag == tf) { end_q = tf; d_end_quar[i] = tf; }
                if (due(end_q, t)) { quarantined[i] = 0; quar = false; }
            }
            if (quar) sk |= SB_QUAR; else sk &= ~SB_QUAR;
            if (ddiag == tf) {                                               // check_enter_iso (people.py:361-366)
                isolated[i] = 1; sk |= SB_ISO;
                d_end_iso[i] = d_rec[i];
                ++c[M_ISOLATED];
            }
            // ---- prepare_transmission (sim.py:602-643): ONE 16-byte record per agent (cvb_device.cuh:AgentRecord) ----
            const bool iso = (sk & SB_ISO) != 0;
            bool inf = (sk & SB_INF) != 0;
            const bool sus = (sk & SB_SUS) != 0;
            int var = sb_ebv(sk) - 1;
            if (inf && !(var >= 0 && var < nv)) { inf = false; var = 0; }
            const bool simple = !inf && !(sk & SB_IMM_NZ) && !A.dense;
            if (simple) {
                const uint32_t want = SB_RS_VALID | (sus ? SB_RS_SUS : 0u) | (quar ? SB_RS_QUAR : 0u);
                if ((sk & (SB_RS_VALID | SB_RS_SUS | SB_RS_QUAR)) != want) {
                    A.rec.rec[i] = make_float4(0.0f, sus ? rs4[k] : 0.0f, 0.0f, __uint_as_float(quar ? 32u : 0u));
                    sk = (sk & ~(SB_RS_VALID | SB_RS_SUS | SB_RS_QUAR)) | want;
                }
            } else {
                sk &= ~(SB_RS_VALID | SB_RS_SUS | SB_RS_QUAR);
                const bool symp = (sk & SB_SYMP) != 0;
                uint32_t code = quar ? 32u : 0u;                              // the quarantine bit matters for targets too
                float rt = 0.0f;
                if (inf) {
                    const float rtv = rel_trans[i];
                    if (rtv != 0.0f) {                                        // can transmit (a zero rel_trans never does)
                        rt = rtv;
                        const bool early = viral_load_early(t, d_inf[i], d_rec[i], d_dead[i], pars.frac_time, pars.high_cap);
                        code = transmit_code(var, symp, iso, quar, early, false);
                        inf_nibble |= 1u << k;
                        const unsigned int pos = warp_append32(A.n_trans);
                        A.trans_list[pos] = (int32_t)i;
                        if (A.adj_ptr) {
                            const long long beg = A.adj_ptr[i], end = A.adj_ptr[i + 1];
                            A.trans_ent[2 * (int64_t)pos] = make_uint4((unsigned)i, (unsigned)(end - beg), (unsigned)(unsigned long long)beg, (unsigned)((unsigned long long)beg >> 32));
                            A.trans_ent[2 * (int64_t)pos + 1] = make_uint4(__float_as_uint(rt), code, 0u, 0u);
                        }
                    }
                }
                const float imm0 = (sk & SB_IMM_NZ) ? sus_imm[i] : 0.0f;
                float s_rec = sus ? rs4[k] : 0.0f, imm_rec = imm0;
                if (nv == 1 && !quar) { s_rec = record_sus(s_rec, 0u, 1.0f, imm_rec); imm_rec = 0.0f; }
                if (A.rec.rec) A.rec.rec[i] = make_float4(rt, s_rec, imm_rec, __uint_as_float(code));
                if (A.rec.ts8) {
                    const float vl = rt != 0.0f ? viral_load_value((code & 64u) != 0, pars.frac_time, pars.load_ratio) : 0.0f;
                    for (int l = 0; l < pars.n_layers; ++l) {
                        if (!((A.rec.ts8_mask >> l) & 1u)) continue;
                        float2 o;
                        o.x = rt != 0.0f ? rel_trans_layer(rt, true, symp, iso, quar, pars.asymp_factor, pars.iso_factor[l], pars.quar_factor[l],
                                                           pars.beta_layer[l], vl) : 0.0f;
                        o.y = sus ? rel_sus_layer(rs4[k], true, quar, pars.quar_factor[l], imm0) : 0.0f;
                        A.rec.ts8[(int64_t)l * n + i] = o;
                    }
                }
            }
            s[k] = sk;
        }
        if (in && ((s[0] ^ sv.x) | (s[1] ^ sv.y) | (s[2] ^ sv.z) | (s[3] ^ sv.w)))
            *reinterpret_cast<uint4*>(S + i0) = make_uint4(s[0], s[1], s[2], s[3]);
        if (A.dense) {                                                       // transmit bitmap for the dense streaming pass
            unsigned word = inf_nibble << (4 * (lane & 7));
            word |= __shfl_xor_sync(0xFFFFFFFFu, word, 1);
            word |= __shfl_xor_sync(0xFFFFFFFFu, word, 2);
            word |= __shfl_xor_sync(0xFFFFFFFFu, word, 4);
            const int64_t widx = (i0 - (int64_t)(lane & 7) * 4) / 32;
            if ((lane & 7) == 0 && widx * 32 < n) A.inf_bits[widx] = word;
        }
    }
    reduce_counters(c, s_cnt);
    __syncthreads();
    if (threadIdx.x < M_NK && s_cnt[threadIdx.x]) {
        unsigned long long* row = A.counters + (int64_t)t * CVB_N_COUNTERS;
        const int ids[M_NK] = {CVB_C_new_diagnoses, CVB_C_new_quarantined, CVB_C_new_isolated};
        atomicAdd(row + ids[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
    }
}

}  // namespace cvb

using namespace cvb;

// ---- launch helpers -----------------------------------------------------------------------------------------------------------
static int fused_ready(cvb_sim* s, const char* who) {
    CVB_REQUIRE(s && s->pars_set, "%s: handle not ready", who);
    for (int f = 0; f < CVB_N_FIELDS; ++f) CVB_REQUIRE(s->people.f[f] != nullptr, "%s: people field %d is not bound", who, f);
    CVB_REQUIRE(s->res.counters && s->res.vcounters && s->res.sums, "%s: result tables are not bound", who);
    CVB_REQUIRE(s->log.count, "%s: infection log is not bound", who);
    CVB_REQUIRE(!s->partitioned, "%s: agent-partitioned handles are stepped by the host (one exchange per day)", who);
    CVB_REQUIRE(s->n % 4 == 0, "%s: the fused pipeline needs a population size that is a multiple of 4", who);
    uintptr_t all = 0;
    for (int f = 0; f < CVB_N_FIELDS; ++f) all |= (uintptr_t)s->people.f[f];
    CVB_REQUIRE((all & 15) == 0, "%s: People arrays must be 16-byte aligned", who);
    CVB_REQUIRE(!s->pars.use_waning || s->nab_kin, "%s: NAb kinetics table not set (cvb_set_nab_kin)", who);
    return 0;
}

static int ensure_fused_buffers(cvb_sim* s) {
    if (!s->state) {
        CVB_CHECK(cudaMalloc((void**)&s->state, (size_t)s->n * sizeof(uint32_t)));
        s->state_valid = 0;
    }
    if (!s->trans_ent) CVB_CHECK(cudaMalloc((void**)&s->trans_ent, (size_t)s->n * 2 * sizeof(uint4)));
    if (!s->case_ent) CVB_CHECK(cudaMalloc((void**)&s->case_ent, (size_t)s->n * sizeof(uint4)));
    return 0;
}

static int grid_agents4(int64_t n) { return grid_for((n + 3) / 4, kThreads, 148 * 8); }

// pack (verify = false) or verify (true: returns the number of agents whose word differs from the arrays in host_out[1])
static int pack_or_check(cvb_sim* s, int32_t t_done, bool verify, int64_t* host_out, cudaStream_t st) {
    unsigned int* scal = reinterpret_cast<unsigned int*>(s->dev_scalars + 16);       // [0] violations, [1] mismatches, then 8 x 3 int32
    CVB_CHECK(cudaMemsetAsync(scal, 0, 2 * sizeof(unsigned int) + 24 * sizeof(int32_t), st));
    pack_state_kernel<<<grid_for(s->n, kThreads, 148 * 8), kThreads, 0, st>>>(s->people, s->state, verify ? s->state : nullptr, s->n, s->nv, t_done,
        s->quar_ring, s->quar_horizon, scal, scal + 1, reinterpret_cast<int32_t*>(scal + 2));
    CVB_LAUNCH_CHECK();
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
static int launch_day_begin(cvb_sim* s, int32_t t, bool test, bool tsel, cudaStream_t st) {
    DayBeginArgs A;
    memset(&A, 0, sizeof(A));
    A.n = s->n; A.id0 = 0; A.t = t; A.nv = s->nv; A.waning = s->pars.use_waning; A.vaxpars = s->pars.has_vaccine_pars;
    A.nab_kin = s->nab_kin; A.nab_kin_len = s->nab_kin_len;
    A.counters = s->res.counters; A.vcounters = s->res.vcounters; A.beds = s->beds;
    const int grid = grid_agents4(s->n);
    if (ensure_f64(&s->partial, &s->partial_cap, (int64_t)grid * 3)) return 1;
    A.partial = s->partial; A.ticket = reinterpret_cast<unsigned int*>(s->dev_scalars + 8); A.sums = s->res.sums;
    A.n_trans = s->n_trans; A.n_case = s->n_case_list;
    A.seed = s->seed;
    if (test) {
        A.tp = s->plan->test;
        A.test_plain = (A.tp.symp_quar_prob == A.tp.symp_prob && A.tp.asymp_quar_prob == A.tp.asymp_prob) ? 1 : 0;
    }
    A.adj_ptr = s->adj_ptr; A.case_ent = s->case_ent;
#define CVB_DB(T1, T2) day_begin_kernel<END, PRE, T1, T2><<<grid, kThreads, 0, st>>>(s->people, s->state, s->pars, A)
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

static int launch_day_mid(cvb_sim* s, int32_t t, cudaStream_t st) {
    bool dense_any = false;
    if (ensure_records_fused(s, dense_any)) return 1;
    DayMidArgs A;
    memset(&A, 0, sizeof(A));
    A.n = s->n; A.t = t; A.nv = s->nv; A.horizon = s->quar_horizon; A.dense = dense_any ? 1 : 0;
    A.quar_slot = s->quar_ring + (int64_t)(t % s->quar_horizon) * s->n;
    A.counters = s->res.counters; A.rec = s->rec; A.inf_bits = s->inf_bits;
    A.adj_ptr = (s->adj && s->adj_layer_mask) ? s->adj_ptr : nullptr;
    A.trans_ent = s->trans_ent; A.n_trans = s->n_trans; A.n_cand = s->n_cand; A.n_case = s->n_case_list; A.trans_list = s->trans_list;
    day_mid_kernel<<<grid_agents4(s->n), kThreads, 0, st>>>(s->people, s->state, s->pars, A);
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
    if (!s->plan) { s->plan = new (std::nothrow) DayPlan(); CVB_REQUIRE(s->plan, "cvb_plan_clear: out of host memory"); }
    memset(s->plan, 0, sizeof(DayPlan));
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

int cvb_run_days(cvb_sim* s, int32_t t0, int32_t t1, cvb_stream st_) {
    cudaStream_t st = (cudaStream_t)st_;
    if (fused_ready(s, "cvb_run_days")) return 1;
    CVB_REQUIRE(t0 >= 0 && t0 < t1 && t1 <= s->npts, "cvb_run_days: days [%d,%d) outside [0,%d)", t0, t1, s->npts);
    if (!s->plan && cvb_plan_clear(s)) return 1;
    const DayPlan& plan = *s->plan;
    if (ensure_fused_buffers(s)) return 1;
    if (plan.has_trace) {
        CVB_REQUIRE(s->adj && s->adj_layer_mask, "cvb_run_days: contact tracing in the fused day needs the adjacency (cvb_bind_adjacency)");
        for (int l = 0; l < s->pars.n_layers; ++l)
            CVB_REQUIRE(!(plan.trace.trace_prob[l] > 0.0) || s->layers[l].n_edges == 0 || ((s->adj_layer_mask >> l) & 1u),
                        "cvb_run_days: traced layer %d is not covered by the adjacency", l);
    }
    if (!s->state_valid) {
        int64_t out[26];
        if (pack_or_check(s, t0 - 1, false, out, st)) return 1;
        CVB_REQUIRE(out[0] == 0, "cvb_run_days: %lld agents are in a state the packed word cannot express (several by-variant rows set, or NAbs without a peak)", (long long)out[0]);
        s->state_valid = 1;
    }
    CVB_CHECK(cudaMemsetAsync(s->n_trans, 0, sizeof(unsigned int), st));
    CVB_CHECK(cudaMemsetAsync(s->n_case_list, 0, sizeof(unsigned int), st));
    CVB_CHECK(cudaMemsetAsync(s->n_cand, 0, sizeof(unsigned int), st));
    for (int32_t t = t0; t < t1; ++t) {
        for (int l = 0; l < s->pars.n_layers; ++l)
            if ((plan.regen_mask >> l) & 1u) { TimedScope ts(s, CVB_TIMED_regen, st); if (cvb_layer_regenerate(s, l, t, st_)) return 1; }
        const bool test = plan.has_test && t >= plan.test_start && (plan.test_end < 0 || t <= plan.test_end);
        const bool trace = plan.has_trace && t >= plan.trace_start && (plan.trace_end < 0 || t <= plan.trace_end);
        {
            TimedScope ts(s, CVB_TIMED_day_begin, st);
            int rc = t > t0 ? launch_day_begin<true, true>(s, t, test, trace, st) : launch_day_begin<false, true>(s, t, test, trace, st);
            if (rc) return rc;
        }
        if (trace) { TimedScope ts(s, CVB_TIMED_trace, st); if (launch_trace_sparse2(s, t, &plan.trace, st)) return 1; }
        { TimedScope ts(s, CVB_TIMED_day_mid, st); if (launch_day_mid(s, t, st)) return 1; }
        { TimedScope ts(s, CVB_TIMED_edge_pass, st); if (edge_pass_impl(s, t, st, true)) return 1; }
        { TimedScope ts(s, CVB_TIMED_infect, st); if (launch_infect_winners(s, t, true, st)) return 1; }
    }
    s->state_valid = 1;                                     // (the edge pass / layer regeneration above do not touch People flags)
    TimedScope ts(s, CVB_TIMED_day_end, st);
    return launch_day_begin<true, false>(s, t1, false, false, st);      // closes day t1 - 1 (its argument is the day AFTER the one it closes)
}

}  // extern "C"
