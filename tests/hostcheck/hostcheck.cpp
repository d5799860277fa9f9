// TEST-ONLY host build of the per-element arithmetic in covasim_b200/csrc/cvb_device.cuh.
// Compiled with g++ (-ffp-contract=off) into tests/hostcheck/_hostcheck.so by tests/test_hostcheck.py
// so that the float32/float64 recipe and the Philox keying can be compared with the oracle on a
// machine without a GPU.  The product never loads this library: it is not a fallback.
#include "cvb_device.cuh"

extern "C" {

void hc_viral_load(int t, const float* di, const float* dr, const float* dd, float ft, float lr, float hc, float* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = cvb::viral_load(t, di[i], dr[i], dd[i], ft, lr, hc);
}

void hc_trans_sus(const float* rt, const float* rs, const unsigned char* inf, const unsigned char* sus, float beta_layer,
                  const float* vl, const unsigned char* symp, const unsigned char* iso, const unsigned char* quar, float af,
                  float isf, float qf, const float* imm, float* ot, float* os, long n) {
    for (long i = 0; i < n; ++i) {
        ot[i] = cvb::rel_trans_layer(rt[i], inf[i] != 0, symp[i] != 0, iso[i] != 0, quar[i] != 0, af, isf, qf, beta_layer, vl[i]);
        os[i] = cvb::rel_sus_layer(rs[i], sus[i] != 0, quar[i] != 0, qf, imm[i]);
    }
}

// the 16-byte agent record path (prepare_transmission writes one record per agent, the edge pass applies the layer factors):
// must give exactly the per-layer tables of hc_trans_sus
void hc_record_trans_sus(const float* rt, const float* rs, const unsigned char* inf, const unsigned char* sus, float beta_layer,
                         const unsigned char* early, const unsigned char* symp, const unsigned char* iso, const unsigned char* quar,
                         float af, float isf, float qf, const float* imm, float frac_time, float load_ratio, float* ot, float* os, long n) {
    const float vl_early = cvb::viral_load_value(true, frac_time, load_ratio), vl_late = cvb::viral_load_value(false, frac_time, load_ratio);
    for (long i = 0; i < n; ++i) {
        unsigned code = quar[i] ? 32u : 0u;
        float t = 0.0f;
        if (inf[i] && rt[i] != 0.0f) { t = rt[i]; code = cvb::transmit_code(0, symp[i] != 0, iso[i] != 0, quar[i] != 0, early[i] != 0, false); }
        const float s = sus[i] ? rs[i] : 0.0f;
        ot[i] = (code & 7u) ? cvb::record_trans(t, code, af, isf, qf, beta_layer, vl_early, vl_late) : 0.0f;
        os[i] = s != 0.0f ? cvb::record_sus(s, code, qf, imm[i]) : 0.0f;
    }
}

void hc_edge_prob(float beta, const float* lb, const float* ts, const float* st, float* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = cvb::edge_prob(beta, lb[i], ts[i], st[i]);
}

void hc_keyed_uniform2(unsigned long long seed, unsigned purpose, unsigned sub, int day, const long long* idx, unsigned slot,
                       double* u1, double* u2, long n) {
    for (long i = 0; i < n; ++i) {
        cvb::u32x4 w = cvb::keyed_words(seed, purpose, sub, day, idx[i], slot);
        u1[i] = cvb::u53(w.x, w.y);
        u2[i] = cvb::u53(w.z, w.w);
    }
}

void hc_keyed_normal(unsigned long long seed, unsigned purpose, unsigned sub, int day, const long long* idx, unsigned slot,
                     double* z, long n) {
    for (long i = 0; i < n; ++i) z[i] = cvb::keyed_normal(seed, purpose, sub, day, idx[i], slot);
}

void hc_dist(int kind, double a, double b, const double* z, double* out, long n) {
    cvb_dist d; d.kind = kind; d.pad_ = 0; d.a = a; d.b = b;
    for (long i = 0; i < n; ++i) out[i] = cvb::dist_from_normal(d, z[i]);
}

void hc_calc_ve(const double* enab, double exp_alpha, double beta, float* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = enab[i] != 0.0 ? cvb::calc_ve(enab[i], exp_alpha, beta) : 0.0f;
}

void hc_calc_ve3(const double* enab, double ea0, double b0, double ea1, double b1, double ea2, double b2, float* o0, float* o1, float* o2, long n) {
    for (long i = 0; i < n; ++i) {
        o0[i] = o1[i] = o2[i] = 0.0f;
        if (enab[i] > 0.0) cvb::calc_ve3(enab[i], ea0, b0, ea1, b1, ea2, b2, o0[i], o1[i], o2[i]);
    }
}

void hc_nab_step(const float* nab, const float* peak, const double* kin, float* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = cvb::nab_step(nab[i], peak[i], kin[i]);
}

void hc_prog_prob(float rel, const float* base, const float* imm, float factor, float* out_imm, float* out_fac, long n) {
    for (long i = 0; i < n; ++i) { out_imm[i] = cvb::prog_prob_imm(rel, base[i], imm[i]); out_fac[i] = cvb::prog_prob_fac(rel, base[i], factor); }
}

}
