/*
 * covasim_b200 -- C ABI of the B200-native Covasim hot path (libcovasim_b200.so).
 *
 * The reference (Covasim 3.1.7) has no FFI for this path: its "operator API" is the set of Python
 * functions Sim.step() calls by name (covasim/sim.py:558-685).  Each entry point below replaces one of
 * them and cites it.  Conventions:
 *   - plain pointers and sizes only; every array pointer is a DEVICE pointer unless marked "host";
 *   - the caller owns all array memory (the Python host binds torch tensors' data_ptr());
 *     the library owns only its scratch, freed by cvb_destroy();
 *   - every function returns 0 on success, non-zero on failure; cvb_last_error() gives the message;
 *   - `cvb_stream` is a cudaStream_t (NULL = the legacy default stream);
 *   - a cvb_sim handle is not thread-safe; distinct handles are independent (ensembles run one
 *     handle per member, each on its own stream);
 *   - bool arrays are one byte per agent (NumPy/torch bool layout), dates are float32 with NaN =
 *     "not set", exactly as in the reference's People (covasim/people.py:47-118).
 */
#ifndef COVASIM_B200_H
#define COVASIM_B200_H

#include <stdint.h>
#include "cvb_fields.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CVB_MAX_VARIANTS 8
#define CVB_MAX_LAYERS   8
#define CVB_MAX_VACCINES 8
#define CVB_N_DURS       9
#define CVB_ABI_VERSION  2      /* 2: round 2 -- fused day / plans / exchange entry points, cvb_vaccinate_pars grew (float64 boost) */

typedef struct cvb_sim cvb_sim;
typedef void* cvb_stream;

/* Duration / NAb distributions (reference utils.py:156-237).  For the lognormal kinds `a`,`b` are the
 * mean and sigma of the UNDERLYING normal (utils.py:223-225), computed by the host in float64. */
enum cvb_dist_kind { CVB_DIST_ZERO = 0, CVB_DIST_NORMAL = 1, CVB_DIST_NORMAL_POS = 2, CVB_DIST_NORMAL_INT = 3,
                     CVB_DIST_LOGNORMAL = 4, CVB_DIST_LOGNORMAL_INT = 5 };
typedef struct cvb_dist { int32_t kind; int32_t pad_; double a; double b; } cvb_dist;

/* Order of cvb_pars.dur[] (reference parameters.py:85-95) */
enum cvb_dur { CVB_DUR_exp2inf = 0, CVB_DUR_inf2sym, CVB_DUR_sym2sev, CVB_DUR_sev2crit, CVB_DUR_asym2rec,
               CVB_DUR_mild2rec, CVB_DUR_sev2rec, CVB_DUR_crit2rec, CVB_DUR_crit2die };

/* Layer codes stored in the infection log for infections that did not come through a contact layer */
#define CVB_LAYER_SEED   (-1)
#define CVB_LAYER_IMPORT (-2)

/* Scalars the kernels read.  Interventions may change them between days (reference sim.py:602-642
 * re-reads them every step), so the host re-sends the struct whenever it changes. */
typedef struct cvb_pars {
    int32_t n_variants, n_layers, use_waning, n_vaccines;
    int32_t quar_period, has_vaccine_pars;
    int64_t n_beds_hosp, n_beds_icu;                 /* < 0: no limit (reference sim.py:579-580) */
    float asymp_factor, frac_time, load_ratio, high_cap;
    float trans_redux, no_hosp_factor, no_icu_factor, nab_boost;
    float beta[CVB_MAX_VARIANTS];                    /* f32(beta*rel_beta*variant rel_beta), sim.py:627 */
    float rel_symp[CVB_MAX_VARIANTS], rel_severe[CVB_MAX_VARIANTS];
    float rel_crit[CVB_MAX_VARIANTS], rel_death[CVB_MAX_VARIANTS];   /* people.py:476-481 */
    float beta_layer[CVB_MAX_LAYERS], iso_factor[CVB_MAX_LAYERS], quar_factor[CVB_MAX_LAYERS];
    float immunity[CVB_MAX_VARIANTS][CVB_MAX_VARIANTS];              /* immunity.py:284-295 */
    double vaccine_imm[CVB_MAX_VACCINES][CVB_MAX_VARIANTS];          /* immunity.py:332-341 */
    double exp_alpha_inf, beta_inf, exp_alpha_symp_inf, beta_symp_inf, exp_alpha_sev_symp, beta_sev_symp; /* immunity.py:216-247 */
    double rel_imm_asymp, rel_imm_mild, rel_imm_severe, nab_norm;    /* immunity.py:184-194; nab_norm = 1+alpha_inf_diff */
    cvb_dist dur[CVB_N_DURS];
    cvb_dist nab_init;
} cvb_pars;

const char* cvb_last_error(void);
int32_t cvb_abi_version(void);
/* number of CUDA kernels this library has launched in this process (for bench.py's gpu_launches) */
int64_t cvb_launch_count(void);
/* sizeof of {cvb_pars, cvb_dist, cvb_test_prob_pars, cvb_trace_pars, cvb_vaccinate_pars} -> host int64[5] */
int cvb_struct_sizes(int64_t* out5);

/* ------------------------------------------------------------------------------------------------
 * Handle: sizes, bound arrays, parameters, scratch.
 * ---------------------------------------------------------------------------------------------- */
int cvb_create(cvb_sim** out, int64_t n_agents, int32_t n_variants, int32_t npts, uint64_t seed);
int cvb_destroy(cvb_sim* s);
int cvb_set_seed(cvb_sim* s, uint64_t seed);
/* Work done by the adjacency form of the edge pass, per day: host int64[npts][2] = {adjacency entries visited,
 * transmitters}.  bench.py derives the kernel's algorithmic bytes from it. */
int cvb_get_edge_work(cvb_sim* s, int64_t* host_out);
/* Clear the library-owned per-run scratch (pending-quarantine ring, winner keys, bed counts) so that the
 * handle can run the same simulation again from a restored People state (Sim.restore) */
int cvb_reset(cvb_sim* s, cvb_stream st);
int cvb_set_pars(cvb_sim* s, const cvb_pars* host_pars);
/* Bind one per-agent array (length n_agents, or n_variants*n_agents for the by-variant / immunity
 * fields).  Replaces People.__setitem__ on the NumPy arrays (reference base.py:1007-1027). */
int cvb_bind_field(cvb_sim* s, int32_t field, void* ptr);
/* Bind one contact layer's edge list (reference base.py:1651-1676: p1:int32[E], p2:int32[E], beta:f32[E]) */
int cvb_bind_layer(cvb_sim* s, int32_t layer, int32_t* p1, int32_t* p2, float* beta, int64_t n_edges);
/* Bidirectional adjacency of the STATIC layers (the CSR form of Base.Contacts): for agent i the entries
 * adj[adj_ptr[i] .. adj_ptr[i+1]) list every edge it belongs to as 16 bytes
 *   {uint32 neighbour, uint32 edge index within its layer, uint32 (layer << 1) | direction, float32 beta}
 * with direction 0 if i is the edge's p1 and 1 if it is its p2.  layer_mask has bit l set for every layer the
 * adjacency covers; transmission and contact tracing then visit only the edges of transmitters / cases for
 * those layers (same results as the dense passes: same Philox keys, same winner keys) and stream the rest
 * (dynamic layers) densely.  Pass layer_mask = 0 to unbind. */
int cvb_bind_adjacency(cvb_sim* s, const int64_t* adj_ptr, const void* adj, int64_t n_entries, uint32_t layer_mask);
/* ------------------------------------------------------------------------------------------------
 * Agent partition of ONE large simulation over several GPUs (BASELINE config 4).
 * Rank r owns the agents [r*chunk, r*chunk + n_agents) of n_global (chunk a multiple of 32; every rank but the
 * last owns exactly `chunk` agents) and evaluates the transmissions and contact notifications whose TARGET it
 * owns.  Per-agent Philox keys and logged ids are GLOBAL ids, so results are bit-identical to the single-GPU run.
 * Per day the HOST performs two fixed-size exchanges (ncclAllGather through torch.distributed):
 *   codes_local  uint8[chunk]     written by cvb_post_and_prepare   -> codes_global  uint8[world*chunk]   read by cvb_edge_pass
 *   case_bits_local u32[chunk/32] written by cvb_trace_select_cases -> case_bits_global u32[world*chunk/32] read by cvb_trace_notify_contacts
 * One code byte (variant, symptomatic, isolated, quarantined, early viral load, breakthrough) plus the replicated
 * initial rel_trans float32[n_global] is all a GPU needs to rebuild a remote source's per-layer transmissibility.
 * hit_capacity bounds the successful transmissions a rank can record per day (<= 0: max(n_agents, 65536)); a day that
 * exceeds it is reported by cvb_partition_status, never dropped silently.
 * ---------------------------------------------------------------------------------------------- */
int cvb_set_partition(cvb_sim* s, int64_t id0, int64_t n_global, int64_t chunk, int32_t world, const float* rel_trans_global,
                      uint8_t* codes_local, const uint8_t* codes_global, uint32_t* case_bits_local, const uint32_t* case_bits_global,
                      int64_t hit_capacity);
/* Exchange over peer memory instead of a library all-gather (ranks of one NVLink / NVSwitch node whose exchange buffers are mapped into
 * each other's address space, e.g. torch symmetric memory): store `n_bytes` of `src` at `dst_offset_bytes` of every rank's buffer
 * (host_peer_ptrs[world] device addresses as seen from THIS rank, own buffer included).  The caller separates pushes from reads with a
 * cross-rank barrier and alternates between two buffers, then points the handle at the one just filled. */
int cvb_peer_push(const void* src, int64_t n_bytes, const uint64_t* host_peer_ptrs, int32_t world, int64_t dst_offset_bytes, cvb_stream st);
int cvb_set_exchange_buffers(cvb_sim* s, const uint8_t* codes_global /* or NULL: unchanged */, const uint32_t* case_bits_global /* or NULL */);
/* Adjacency of a partitioned handle: rows adj_ptr[0 .. world*chunk] are indexed by the GLOBAL id of the source; the
 * 16-byte entries are those of cvb_bind_adjacency with `neighbour` = LOCAL index of the target */
int cvb_bind_partition_adjacency(cvb_sim* s, const int64_t* adj_ptr, const void* adj, int64_t n_entries, uint32_t layer_mask);
/* host int64[2] = {transmissions dropped because hit_capacity was too small, agents whose rel_trans the code cannot
 * express}; both must be 0 for a valid run (synchronises) */
int cvb_partition_status(cvb_sim* s, int64_t* host_out2);

/* Per-day result tables: counters int64[npts][CVB_N_COUNTERS], vcounters int64[npts][n_variants][CVB_N_VCOUNTERS],
 * sums double[npts][4] = {sum nab over alive, sum sus_imm, sum symp_imm, unused} (reference sim.py:652-674) */
int cvb_bind_results(cvb_sim* s, int64_t* counters, int64_t* vcounters, double* sums);
/* Optional: the caller's own int64[npts][2] table for the day's {severe, critical} counts after update_states_pre (sim.py:579-580: the bed
 * limits).  An agent-partitioned run binds one and sums the day's row over the ranks (one 16-byte all-reduce) before anybody is infected */
int cvb_bind_beds(cvb_sim* s, int64_t* beds);
/* Device infection log (reference people.py:508-511): parallel arrays of capacity `cap`, *count is device int64 */
int cvb_bind_log(cvb_sim* s, int32_t* source, int32_t* target, int32_t* date, int8_t* layer, int8_t* variant,
                 int64_t cap, int64_t* count);

/* ------------------------------------------------------------------------------------------------
 * Stateless operators: 1:1 with the reference's Numba kernels (covasim/utils.py:39-147).
 * ---------------------------------------------------------------------------------------------- */
/* Counter-based uniforms: out[k] = 53-bit uniform in [0,1) from Philox4x32-10 with key (seed, purpose, sub) and counter
 * (index0 + k, day, slot) -- the keyed draw every native-RNG kernel uses (csrc/cvb_device.cuh:keyed_uniform).  The device-side
 * population generator (reference population.py:143-364 make_randpop / make_*_contacts) is built from it. */
int cvb_keyed_uniform(uint64_t seed, uint32_t purpose, uint32_t sub, int32_t day, int64_t index0, int64_t n, uint32_t slot,
                      double* out, cvb_stream st);
/* utils.py:39-79 compute_viral_load */
int cvb_compute_viral_load(int32_t t, const float* date_inf, const float* date_rec, const float* date_dead,
                           float frac_time, float load_ratio, float high_cap, float* out, int64_t n, cvb_stream st);
/* utils.py:82-90 compute_trans_sus */
int cvb_compute_trans_sus(const float* rel_trans, const float* rel_sus, const uint8_t* inf, const uint8_t* sus,
                          float beta_layer, const float* viral_load, const uint8_t* symp, const uint8_t* iso,
                          const uint8_t* quar, float asymp_factor, float iso_factor, float quar_factor,
                          const float* immunity_factors, float* out_trans, float* out_sus, int64_t n, cvb_stream st);
/* utils.py:93-128 compute_infections, replay form.  The reference draws one uniform per edge-direction
 * whose float32 probability is non-zero, in edge order, direction p1->p2 first.  _count reports how many
 * draws each direction consumes (host int64[2]; synchronises the stream); _draw consumes
 * n_draws[0]+n_draws[1] uniforms (device, float64) and writes the ordered (source, target) lists
 * (capacity n_draws[0]+n_draws[1]) and their length (host int64). */
int cvb_infections_count(cvb_sim* s, float beta, const int32_t* p1, const int32_t* p2, const float* layer_betas,
                         int64_t n_edges, const float* rel_trans, const float* rel_sus, int64_t* host_n_draws,
                         cvb_stream st);
int cvb_infections_draw(cvb_sim* s, float beta, const int32_t* p1, const int32_t* p2, const float* layer_betas,
                        int64_t n_edges, const float* rel_trans, const float* rel_sus, const double* uniforms,
                        int32_t* out_src, int32_t* out_tgt, int64_t* host_n_out, cvb_stream st);
/* utils.py:131-147 find_contacts + base.py:1842-1844: sorted unique partners of `inds` (int64[n_inds]);
 * out has capacity n_agents; *host_n_out receives the count (synchronises) */
int cvb_find_contacts(cvb_sim* s, const int32_t* p1, const int32_t* p2, int64_t n_edges, const int64_t* inds,
                      int64_t n_inds, int32_t* out, int64_t* host_n_out, cvb_stream st);
/* utils.py:494-506 true(): ascending indices of non-zero bytes; *host_n_out receives the count (synchronises) */
int cvb_true_indices(cvb_sim* s, const uint8_t* flags, int64_t n, int32_t* out, int64_t* host_n_out, cvb_stream st);

/* ------------------------------------------------------------------------------------------------
 * One simulated day on the bound People / Layers (reference sim.py:558-685).
 * ---------------------------------------------------------------------------------------------- */
/* people.py:164-186 update_states_pre (+ immunity.py:303-350 check_immunity when use_waning) */
int cvb_update_states_pre(cvb_sim* s, int32_t t, cvb_stream st);
/* people.py:620-640 schedule_quarantine: request quarantine of `inds` starting on start_day and ending on
 * end_day.  The host dict People._pending_quarantine becomes a device ring of per-agent "max requested
 * end day" slots indexed by start_day % horizon (requests for one agent and start day are order-
 * independent: the reference extends to the max end and counts the agent once, people.py:339-346). */
int cvb_schedule_quarantine(cvb_sim* s, const int32_t* inds, int64_t n, int32_t start_day, float end_day, cvb_stream st);
/* how many days ahead requests may start (1 = today only); grows the ring (requests already pending keep their days) */
int cvb_set_quar_horizon(cvb_sim* s, int32_t horizon);
/* base.py:444-446 Sim.copy of a running simulation: the library-owned per-run state of `src` (pending quarantine requests, bed counts, edge-work
 * counters) copied into `dst`, a handle of the same shape whose arrays the caller has already bound */
int cvb_clone_scratch(cvb_sim* dst, const cvb_sim* src, cvb_stream st);
/* Compact restore of a saved People state (base.py:444-446 Sim.copy / the caller's own checkpoints): `arena` is the caller's device
 * buffer holding every per-agent array; dev_table (device int64) = n_seg x {first 32-bit word, word count, fill value} followed by
 * n_exc word indices and n_exc values.  Every segment is filled with its value, then the exceptions are written; arrays that are
 * dense in the saved state are copied by the caller.  Bit-identical to copying the whole buffer. */
int cvb_restore_compact(void* arena, int64_t arena_bytes, const int64_t* dev_table, int32_t n_seg, int64_t n_exc, cvb_stream st);
/* immunity.py:298 pars['nab_kin']: per-day NAb increments (host float64[n]) */
int cvb_set_nab_kin(cvb_sim* s, const double* host_kin, int64_t n);
/* people.py:189-196 update_states_post (check_diagnosed, check_quar, check_enter_iso) */
int cvb_update_states_post(cvb_sim* s, int32_t t, cvb_stream st);
/* sim.py:602-643: viral load + transmissibility / susceptibility: one 16-byte record per agent for the edge pass (which applies the
 * per-layer factors), plus the reference's per-layer {rel_trans, rel_sus} pairs for the layers the dense streaming pass reads */
int cvb_prepare_transmission(cvb_sim* s, int32_t t, cvb_stream st);
/* update_states_post and prepare_transmission as ONE pass over the agents (what Sim.step uses) */
int cvb_post_and_prepare(cvb_sim* s, int32_t t, cvb_stream st);
/* sim.py:622-649 native-RNG form: ONE pass over every layer's edges, both directions, all variants;
 * per-edge Philox4x32-10 uniforms keyed (seed, day, layer, edge); winners by atomicMin of
 * (variant, layer, direction, edge) per target == the reference's first-occurrence rule (people.py:465-467) */
int cvb_edge_pass(cvb_sim* s, int32_t t, cvb_stream st);
/* people.py:435-586 infect (+ immunity.py:138-202 update_peak_nab) for the edge pass's winners; keyed draws */
int cvb_infect_winners(cvb_sim* s, int32_t t, cvb_stream st);
/* people.py:435-586 infect for an explicit index list (seed infections sim.py:528, imports sim.py:587,
 * variant imports immunity.py:128); duplicates / non-susceptibles are dropped; keyed draws */
int cvb_infect_list(cvb_sim* s, const int32_t* inds, int64_t n, int32_t variant, int32_t layer_code, int32_t t,
                    int32_t count_flows /* 0 for seed infections at initialisation: their flows are discarded */,
                    int32_t hosp_max, int32_t icu_max /* 0 / 1 as the caller decided (people.py:435), -1: from today's severe /
                                                         critical counts vs n_beds_* (sim.py:579-580) */,
                    cvb_stream st);
/* The same with the random draws GIVEN instead of keyed: tape = device float64[n][16], row j for inds[j] (distinct agents), column =
 * prognosis step (0 exp2inf, 1 symptomatic?, 2 asym2rec | inf2sym, 3 severe?, 4 mild2rec | sym2sev, 5 critical?, 6 sev2rec | sev2crit,
 * 7 dies?, 8 crit2rec | crit2die, 9 initial NAb level): the uniform of a Bernoulli step, the finished sample of the others.  Verification
 * entry point: feeds the kernel the draws a recorded run of the reference consumed (tests/golden/infect_tape.npz) */
int cvb_infect_list_taped(cvb_sim* s, const int32_t* inds, int64_t n, int32_t variant, int32_t layer_code, int32_t t,
                          int32_t count_flows, int32_t hosp_max, int32_t icu_max, const double* tape, cvb_stream st);
/* immunity.py:205-213 update_nab + sim.py:652-674 stock counts and population means */
int cvb_update_nab_count(cvb_sim* s, int32_t t, cvb_stream st);
/* All of the above for day t in reference order, with no built-in interventions in between */
int cvb_step_day(cvb_sim* s, int32_t t, cvb_stream st);

/* ------------------------------------------------------------------------------------------------
 * Built-in interventions as device passes ("next" rows of SURVEY.md section 8(f)).
 * ---------------------------------------------------------------------------------------------- */
typedef struct cvb_test_prob_pars {        /* interventions.py:857-981 */
    double symp_prob, asymp_prob, symp_quar_prob, asymp_quar_prob, sensitivity, loss_prob;
    int32_t quar_policy;                   /* 0 start, 1 end, 2 both, 3 daily, 4 none here (days-since-start lists, functions: the caller applies them) */
    int32_t test_delay, index, pad_;
} cvb_test_prob_pars;
/* prob_override: NULL, or float64[n] of explicit per-agent probabilities (NaN = none): the `subtarget` option (interventions.py:971-973) */
int cvb_test_prob(cvb_sim* s, int32_t t, const cvb_test_prob_pars* host_pars, const double* prob_override, cvb_stream st);
/* Verification: the same with every agent's three uniforms GIVEN (device float64[n][3]: tested today?, test positive?, not lost to follow-up?)
 * -- the draws a recorded run of the reference consumed (tests/golden/test_tape.npz) */
int cvb_test_prob_taped(cvb_sim* s, int32_t t, const cvb_test_prob_pars* host_pars, const double* prob_override, const double* tape, cvb_stream st);

typedef struct cvb_test_num_pars {         /* interventions.py:718-854 */
    double symp_test, quar_test;
    int32_t quar_policy;                   /* 0 start, 1 end, 2 both, 3 daily, 4 none here (days-since-start lists, functions: the caller applies them) */
    int32_t index;
} cvb_test_num_pars;
/* test_num, device part one: per-agent testing weights (the reference's test_probs, float64[n]) and exponential-clock keys
 * -log(1 - u) / w (float64[n], +inf where w == 0); the n_tests agents with the smallest keys are a weighted sample without
 * replacement.  The caller selects them and hands them to cvb_test_list. */
int cvb_test_num_keys(cvb_sim* s, int32_t t, const cvb_test_num_pars* host_pars, double* weight, double* key, cvb_stream st);
/* people.py:589-617 People.test for a list of distinct agents (int32[n_inds]); keyed sensitivity / loss-to-follow-up draws */
int cvb_test_list(cvb_sim* s, int32_t t, const int32_t* inds, int64_t n_inds, double sensitivity, double loss_prob, int32_t test_delay,
                  int32_t index, cvb_stream st);

typedef struct cvb_trace_pars {            /* interventions.py:984-1145 */
    double trace_prob[CVB_MAX_LAYERS];
    int32_t trace_time[CVB_MAX_LAYERS];
    int32_t presumptive, quar_period, index, pad_;
} cvb_trace_pars;
/* Marks contacts of today's cases, sets known_contact/date_known_contact and queues quarantine.
 * Requests with trace_time 0 go to pend_quar_end; later ones to the per-day ring. */
int cvb_contact_tracing(cvb_sim* s, int32_t t, const cvb_trace_pars* host_pars, cvb_stream st);
/* The same for an explicit list of distinct cases (int32[n_cases]) chosen by the caller: contact_tracing with a `capacity`
 * (interventions.py:1079-1083 picks `capacity` of today's cases at random) */
int cvb_contact_tracing_list(cvb_sim* s, int32_t t, const cvb_trace_pars* host_pars, const int32_t* case_inds, int64_t n_cases, cvb_stream st);
/* Verification: cvb_contact_tracing with the uniform of every (layer, contact) GIVEN (device float64[n_layers][n], read only for contacts
 * of today's cases) -- the draws binomial_filter consumed in a recorded run of the reference (interventions.py:1109-1116) */
int cvb_contact_tracing_taped(cvb_sim* s, int32_t t, const cvb_trace_pars* host_pars, const double* tape, cvb_stream st);
/* Checkpoints and verification: the pending quarantine requests starting on start_day (people.py:620-640 _pending_quarantine[start_day]) as the latest
 * requested end day of every agent (device float32[n], -1 = none) */
int cvb_pending_quarantine(cvb_sim* s, int32_t start_day, float* out_end_day, cvb_stream st);
/* ... and the inverse, for restoring a checkpoint taken while requests were pending: the requests starting on start_day become exactly `end_day`
 * (device float32[n], -1 = none); start_day must lie within the horizon of the current day */
int cvb_set_pending_quarantine(cvb_sim* s, int32_t start_day, const float* end_day, cvb_stream st);
/* The same in two phases for agent-partitioned handles: select today's local cases into case_bits_local; (the host
 * all-gathers the bitmap); notify the LOCAL contacts of every GLOBAL case */
int cvb_trace_select_cases(cvb_sim* s, int32_t t, const cvb_trace_pars* host_pars, cvb_stream st);
int cvb_trace_notify_contacts(cvb_sim* s, int32_t t, const cvb_trace_pars* host_pars, cvb_stream st);

typedef struct cvb_vaccinate_pars {        /* interventions.py:1257-1662 */
    double prob;
    cvb_dist nab_init;
    float nab_boost;
    int32_t booster, vaccine_index, max_doses, index, first_dose_today, second_dose_today, interval, n_days;
    /* the boost as a float64 factor: the reference multiplies peak_nab (float32) by the vaccine's nab_boost with NumPy's scalar rules -- a
     * Python number acts in float32 (nab_boost above), a NumPy float64 (what `target_eff` computes, interventions.py:1393-1394) in float64 */
    double nab_boost_f64;
    int32_t nab_boost_is_f64, pad_;
} cvb_vaccinate_pars;
/* `iv_doses` int32[n] is this intervention's own dose count, `due_day` int32[n] the day an agent's second
 * dose is due (-1 none) -- the device form of second_dose_days (interventions.py:1655-1660) */
int cvb_vaccinate_prob(cvb_sim* s, int32_t t, const cvb_vaccinate_pars* host_pars, int32_t* iv_doses, int32_t* due_day,
                       const double* prob_override /* NULL or float64[n], NaN = none: `subtarget` (interventions.py:1644-1647) */, cvb_stream st);

/* Verification: the same with every agent's initial NAb sample GIVEN (device float64[n]: the value immunity.py:178 drew for agents without
 * prior antibodies in a recorded run of the reference) */
int cvb_vaccinate_taped(cvb_sim* s, int32_t t, const cvb_vaccinate_pars* host_pars, int32_t* iv_doses, int32_t* due_day, const double* prob_override,
                        const double* tape, cvb_stream st);

/* base.py:1849-1876 Layer.update with frac=1: regenerate every edge of a dynamic layer on the device */
int cvb_layer_regenerate(cvb_sim* s, int32_t layer, int32_t t, cvb_stream st);
/* ... with frac < 1: only the listed edges (device int64[n_inds], chosen by the caller: base.py:1866 cvu.choose) get new endpoints */
int cvb_layer_regenerate_list(cvb_sim* s, int32_t layer, int32_t t, const int64_t* inds, int64_t n_inds, cvb_stream st);

/* ------------------------------------------------------------------------------------------------
 * Whole blocks of days without returning to the host (sim.py:688-761 Sim.run -> sim.py:558-685 Sim.step).
 * The built-in interventions that need no host decision are registered once (the "day plan"); cvb_run_days then runs
 * days [t0, t1) as five launches per day: (stocks + update_nab of day t-1, update_states_pre + check_immunity, test_prob,
 * case selection) -> contact tracing over the cases' adjacency rows -> (update_states_post + transmission records) ->
 * transmission -> infect.  Results are identical to calling the per-step entry points above day by day.
 * A packed per-agent state word (library-owned) is kept in step with the People flag arrays by these kernels; it is rebuilt
 * from the arrays on the first cvb_run_days after any other entry point, or after cvb_state_invalidate (call it when the host
 * wrote People arrays itself).
 * ---------------------------------------------------------------------------------------------- */
int cvb_plan_clear(cvb_sim* s);
/* interventions.py:857-981 test_prob active on days [start_day, end_day] (end_day < 0: to the end); no per-agent overrides */
int cvb_plan_test_prob(cvb_sim* s, const cvb_test_prob_pars* host_pars, int32_t start_day, int32_t end_day);
/* interventions.py:984-1145 contact_tracing on days [start_day, end_day]; applied after the plan's test_prob; every traced
 * layer must be covered by the adjacency; not presumptive, no capacity */
int cvb_plan_contact_tracing(cvb_sim* s, const cvb_trace_pars* host_pars, int32_t start_day, int32_t end_day);
/* interventions.py:1257-1662 vaccinate_prob without per-agent overrides: host_day_flags uint8[npts], bit 0 = first doses are offered that day, bit 1 =
 * second doses fall due; iv_doses / due_day as for cvb_vaccinate_prob.  Up to four; applied in registration order after testing and tracing */
int cvb_plan_vaccinate(cvb_sim* s, const cvb_vaccinate_pars* host_pars, const uint8_t* host_day_flags, int32_t* iv_doses, int32_t* due_day);
/* people.py:199-206 update_contacts: the dynamic layers (bit l = layer l) regenerated at the start of every day */
int cvb_plan_dynamic_layers(cvb_sim* s, uint32_t layer_mask);
int cvb_run_days(cvb_sim* s, int32_t t0, int32_t t1, cvb_stream st);
/* The same for several handles in lockstep (ensembles of small simulations, run.py:1406-1519 multi_run): day by day, every member's
 * launches go to its own stream from ONE host thread, so the members' kernels overlap on the GPU.  host arrays of n_handles entries */
int cvb_run_days_multi(cvb_sim** handles, int32_t n_handles, int32_t t0, int32_t t1, cvb_stream* streams);
/* The fused day of an AGENT-PARTITIONED handle, phase by phase (the host exchanges data in between): 0 day_begin -> [all-gather of the case
 * bitmap on tracing days] -> 1 notify the local contacts of every global case -> 2 day_mid (writes the 1-byte transmit codes) -> [all-gather of
 * the codes] -> 3 edge pass + infect; 4 closes day t - 1 at the end of a block of days.  first_of_block: this is the first fused day since anything
 * else ran (the state words are rebuilt if needed) */
int cvb_fused_phase(cvb_sim* s, int32_t t, int32_t phase, int32_t first_of_block, cvb_stream st);
int cvb_state_invalidate(cvb_sim* s);
/* Launch-shape overrides of the fused day kernels, for tuning runs (0 = default): what 0 / 1 = CTA size / agents per CTA of the
 * first per-agent kernel, 2 / 3 = the same for the second */
int cvb_tune(cvb_sim* s, int32_t what, int32_t value);
/* Per-kernel timing of cvb_run_days (CUDA events around every launch; off by default).  cvb_timing_read synchronises, returns
 * the milliseconds and launch counts accumulated since the last read, per kernel kind, and resets them */
enum cvb_timed { CVB_TIMED_day_begin = 0, CVB_TIMED_trace, CVB_TIMED_day_mid, CVB_TIMED_edge_pass, CVB_TIMED_infect, CVB_TIMED_day_end,
                 CVB_TIMED_regen, CVB_TIMED_vaccinate, CVB_N_TIMED };
int cvb_timing_enable(cvb_sim* s, int32_t on);
int cvb_timing_read(cvb_sim* s, double* host_ms, int64_t* host_launches);
/* Verification: recompute every agent's state word from the People arrays (t_done = last completed day) and compare with the
 * stored one.  host int64[26] = {agents the word cannot express, agents that differ, then up to 8 x (agent, stored, recomputed)};
 * synchronises */
int cvb_state_check(cvb_sim* s, int32_t t_done, int64_t* host_out26, cvb_stream st);

#ifdef __cplusplus
}
#endif
#endif
