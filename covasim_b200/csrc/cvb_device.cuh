// Per-element arithmetic of the Covasim hot path, written once as __host__ __device__ functions.
// The CUDA kernels (people_kernels.cu, edge_pass.cu, infect.cu, ops_stateless.cu) call these per
// agent / per edge; tests/hostcheck compiles the same functions for the host so the float32/float64
// recipe can be checked against the oracle without a GPU.  (The host build is test-only: the
// product never runs it.)
//
// Precision recipe: explicit *_rn intrinsics on the device and -fmad=false /
// -ffp-contract=off at compile time, so no product is ever fused into an FMA.
#pragma once
#include <stdint.h>
#include <math.h>
#include "covasim_b200.h"

#if defined(__CUDACC__)
#define CVB_HD __host__ __device__ __forceinline__
#else
#define CVB_HD inline
#endif

namespace cvb {

// ---- float helpers -------------------------------------------------------------------------
CVB_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
CVB_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
CVB_HD float fsub(float a, float b) { return fadd(a, -b); }
CVB_HD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b; return r;
#endif
}
CVB_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
CVB_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}
CVB_HD bool is_nan(float x) { return x != x; }
CVB_HD float nanf32() {
#if defined(__CUDA_ARCH__)
    return __int_as_float(0x7fc00000);
#else
    return NAN;
#endif
}
// "date is set and has been reached": NaN compares false (reference people.py:211-219)
CVB_HD bool due(float date, int t) { return (float)t >= date; }

// ---- Philox4x32-10 (Salmon et al., SC'11) -----------------------------------------------------
enum purpose : uint32_t { P_EDGE = 1, P_INFECT = 2, P_TEST = 3, P_TEST_SENS = 4, P_TEST_LOSS = 5, P_TRACE = 6,
                          P_VACC = 7, P_NAB_VACC = 8, P_DYNLAYER = 9, P_POP = 10 };

struct u32x4 { uint32_t x, y, z, w; };

CVB_HD u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += W0;
        k1 += W1;
    }
    u32x4 o = {c0, c1, c2, c3};
    return o;
}

// key = (seed lo, seed hi ^ purpose<<24 ^ sub), counter = (index lo, index hi, day, slot); oracle/philox.py:keyed_words
CVB_HD u32x4 keyed_words(uint64_t seed, uint32_t purpose, uint32_t sub, int32_t day, int64_t index, uint32_t slot) {
    uint32_t k0 = (uint32_t)seed;
    uint32_t k1 = (uint32_t)(seed >> 32) ^ ((purpose & 0xFFu) << 24) ^ (sub & 0xFFFFFFu);
    uint64_t ix = (uint64_t)index;
    return philox4x32_10((uint32_t)ix, (uint32_t)(ix >> 32), (uint32_t)day, slot, k0, k1);
}

// 53-bit uniform in [0,1) from two words: the MT19937 random_sample recipe
CVB_HD double u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}
CVB_HD double keyed_uniform(uint64_t seed, uint32_t purpose, uint32_t sub, int32_t day, int64_t index, uint32_t slot) {
    u32x4 w = keyed_words(seed, purpose, sub, day, index, slot);
    return u53(w.x, w.y);
}
// Box-Muller, cosine branch (oracle/philox.py:keyed_normal)
CVB_HD double normal_from_words(const u32x4 w) {
    double u1 = u53(w.x, w.y), u2 = u53(w.z, w.w);
    return dmul(sqrt(dmul(-2.0, log(1.0 - u1))), cos(dmul(6.283185307179586, u2)));
}
CVB_HD double keyed_normal(uint64_t seed, uint32_t purpose, uint32_t sub, int32_t day, int64_t index, uint32_t slot) {
    return normal_from_words(keyed_words(seed, purpose, sub, day, index, slot));
}
// One draw of a cvb_dist from a standard normal z (reference utils.py:211-231)
CVB_HD double dist_from_normal(const cvb_dist& d, double z) {
    switch (d.kind) {
        case CVB_DIST_NORMAL:        return dadd(d.a, dmul(d.b, z));
        case CVB_DIST_NORMAL_POS:    return fabs(dadd(d.a, dmul(d.b, z)));
        case CVB_DIST_NORMAL_INT:    return rint(fabs(dadd(d.a, dmul(d.b, z))));
        case CVB_DIST_LOGNORMAL:     return exp(dadd(d.a, dmul(d.b, z)));
        case CVB_DIST_LOGNORMAL_INT: return rint(exp(dadd(d.a, dmul(d.b, z))));
        default:                     return 0.0;
    }
}

// ---- A2: viral load (reference utils.py:39-79) -----------------------------------------------
// split in two so that the agent-partitioned form can ship the 1-bit "early" decision instead of three dates
CVB_HD bool viral_load_early(int32_t t, float d_inf, float d_rec, float d_dead, float frac_time, float high_cap) {
    float stop = is_nan(d_dead) ? d_rec : d_dead;
    float total = fsub(stop, d_inf);
    float trans_day = fmul(frac_time, total);
    float trans_point = (trans_day > high_cap) ? fdiv(high_cap, total) : frac_time;
    // the comparison is evaluated in float64 (int32 - float32 promotes in the reference)
    return ((double)t - (double)d_inf) / (double)total < (double)trans_point;
}
CVB_HD float viral_load_value(bool early, float frac_time, float load_ratio) {
    float denom = fadd(1.0f, fmul(frac_time, fsub(load_ratio, 1.0f)));
    return early ? fdiv(load_ratio, denom) : fdiv(1.0f, denom);
}
CVB_HD float viral_load(int32_t t, float d_inf, float d_rec, float d_dead, float frac_time, float load_ratio, float high_cap) {
    return viral_load_value(viral_load_early(t, d_inf, d_rec, d_dead, frac_time, high_cap), frac_time, load_ratio);
}

// ---- transmit code: what another GPU needs to rebuild an agent's per-layer transmissibility (1 byte) ----
// bits 0-2: carried variant + 1 (0 = cannot transmit today), 3: symptomatic, 4: isolated, 5: quarantined,
// 6: early (high) viral load, 7: rel_trans has been reduced by a breakthrough infection (people.py:486-491)
CVB_HD uint8_t transmit_code(int variant, bool symp, bool iso, bool quar, bool early, bool redux) {
    return (uint8_t)(((variant + 1) & 7) | (symp ? 8 : 0) | (iso ? 16 : 0) | (quar ? 32 : 0) | (early ? 64 : 0) | (redux ? 128 : 0));
}

// ---- A3: transmissibility / susceptibility for one layer (reference utils.py:82-90) -------------
CVB_HD float rel_trans_layer(float rel_trans, bool inf, bool symp, bool iso, bool quar, float asymp_factor,
                             float iso_factor, float quar_factor, float beta_layer, float vload) {
    float f_asymp = symp ? 1.0f : asymp_factor;
    float f_iso = iso ? iso_factor : 1.0f;
    float f_quar = quar ? quar_factor : 1.0f;
    float r = fmul(rel_trans, inf ? 1.0f : 0.0f);
    r = fmul(r, f_quar);
    r = fmul(r, f_asymp);
    r = fmul(r, f_iso);
    r = fmul(r, beta_layer);
    return fmul(r, vload);
}
CVB_HD float rel_sus_layer(float rel_sus, bool sus, bool quar, float quar_factor, float imm) {
    float f_quar = quar ? quar_factor : 1.0f;
    float r = fmul(fmul(rel_sus, sus ? 1.0f : 0.0f), f_quar);
    return (float)dmul((double)r, 1.0 - (double)imm);      // float64 multiply, rounded once to float32
}

// ---- per-agent transmission record (16 bytes) and what the edge passes rebuild from it ---------------------------
// prepare_transmission writes ONE record per agent instead of one {rel_trans, rel_sus} pair per layer; the per-layer factors
// are applied by the edge pass for the edges it actually evaluates, with the same float32 chain (rel_trans_layer /
// rel_sus_layer above), so probabilities are bit-identical to the per-layer tables of the reference (utils.py:82-90).
//   t    : rel_trans if the agent can transmit today, else 0
//   s    : rel_sus if the agent is susceptible, else 0
//   imm0 : sus_imm against variant 0 (other variants are read from the People array)
//   code : transmit_code bits; bits 0-2 = 0 for agents that cannot transmit, bit 5 (quarantined) is set for everyone
struct AgentRecord { float t, s, imm0; uint32_t code; };

CVB_HD float record_trans(float t, uint32_t code, float asymp_factor, float iso_factor, float quar_factor, float beta_layer,
                          float vl_early, float vl_late) {
    return rel_trans_layer(t, true, (code & 8u) != 0, (code & 16u) != 0, (code & 32u) != 0, asymp_factor, iso_factor, quar_factor,
                           beta_layer, (code & 64u) ? vl_early : vl_late);
}
// rel_sus_layer(s, true, quar, quar_factor, imm) without the float64 round trip when imm == 0: r * (1 - 0) == r exactly
CVB_HD float record_sus(float s, uint32_t code, float quar_factor, float imm) {
    const float r = (code & 32u) ? fmul(s, quar_factor) : s;
    return imm == 0.0f ? r : (float)dmul((double)r, 1.0 - (double)imm);
}

// ---- A4: per-edge transmission probability (reference utils.py:117) ---------------------------
CVB_HD float edge_prob(float beta, float layer_beta, float trans_src, float sus_tgt) {
    return fmul(fmul(fmul(beta, layer_beta), trans_src), sus_tgt);
}

// ---- A8: NAb -> protection (reference immunity.py:216-247) ----------------------------------
CVB_HD float calc_ve(double enab, double exp_alpha, double beta) {
    double lo = dmul(exp_alpha, pow(enab, beta));
    return (float)(lo / dadd(1.0, lo));
}

// The three protection axes share the NAb level: pow(x, b) = exp(b * log(x)) with ONE log (x > 0).  Relative
// error ~1e-15, far below the float32 rounding of the stored result (the parity tests allow 1e-6 on these fields).
CVB_HD void calc_ve3(double enab, double ea0, double b0, double ea1, double b1, double ea2, double b2, float& s0, float& s1, float& s2) {
    const double lg = log(enab);
    const double l0 = dmul(ea0, exp(dmul(b0, lg))), l1 = dmul(ea1, exp(dmul(b1, lg))), l2 = dmul(ea2, exp(dmul(b2, lg)));
    s0 = (float)(l0 / dadd(1.0, l0));
    s1 = (float)(l1 / dadd(1.0, l1));
    s2 = (float)(l2 / dadd(1.0, l2));
}

// ---- A9: one day of NAb kinetics (reference immunity.py:205-213) -----------------------------
CVB_HD float nab_step(float nab, float peak, double kin) {
    float v = (float)dadd((double)nab, dmul(kin, (double)peak));
    if (v < 0.0f) v = 0.0f;
    if (v > peak) v = peak;
    return v;
}

// ---- prognosis probabilities (reference people.py:523, 538, 552, 565): float32 throughout ------
CVB_HD float prog_prob_imm(float rel, float base, float imm) { return fmul(fmul(rel, base), fsub(1.0f, imm)); }
CVB_HD float prog_prob_fac(float rel, float base, float factor) { return fmul(fmul(rel, base), factor); }

}  // namespace cvb
