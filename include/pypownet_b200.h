/* pypownet_b200 -- C ABI of the B200-native batched power-grid step path.
 *
 * Drop-in boundary for the per-timestep hot path of pypownet (reference @ /root/reference):
 *   RunEnv.step / simulate / process_game_over      pypownet/environment.py:848-888
 *   Game.step, apply_action, _verify_illegal_action,
 *   load_entries_from_next_timestep,
 *   _compute_loadflow_cascading, process_game_over   pypownet/game.py:405-501, 503-589, 591-753, 762-885, 887-943
 *   Grid.compute_loadflow + the four pypower.api
 *   calls it makes (runpf / rundcpf)                 pypownet/grid.py:62-65, 140-264
 *   Grid.extract_flows_a, load_timestep_injections,
 *   apply_topology, export_to_observation            pypownet/grid.py:112-138, 273-311, 360-423, 496-566
 *   Observation.as_array layout                      pypownet/environment.py:451-466, 511-517, 583-595
 *
 * Conventions: plain pointers and sizes, no C++ or torch types.  Every function returns 0 on success or a
 * negative PPN_E_* code; ppn_last_error() gives the message.  Pointers named *_dev are device pointers on the
 * handle's GPU, *_host are host pointers.  The caller owns every buffer it passes; the library owns the env
 * state.  Work is enqueued on the CUDA stream passed as `stream` (a cudaStream_t cast to void*, NULL = default
 * stream); calls on one handle must be serialised by the caller.  In-step "exceptions" of the reference
 * (game.py:861-885 returns exception INSTANCES) are returned as per-env integer flags.
 */
#ifndef PYPOWNET_B200_H
#define PYPOWNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPN_OK 0
#define PPN_E_INVALID (-1)   /* bad argument / malformed case */
#define PPN_E_CUDA (-2)      /* CUDA runtime error */
#define PPN_E_STATE (-3)     /* call order (e.g. step before load_chronics/reset) */
#define PPN_E_UNSUPPORTED (-4)

/* per-env flag codes (replace the exception instances of game.py:861-885 / environment.py:14-35) */
#define PPN_FLAG_NONE 0
#define PPN_FLAG_ILLEGAL_ACTION 1        /* IllegalActionException  (step still played with the corrected action) */
#define PPN_FLAG_DIVERGING_LOADFLOW 2    /* DivergingLoadflowException, done */
#define PPN_FLAG_TOO_MANY_LOADS_CUT 3    /* TooManyConsumptionsCut, done */
#define PPN_FLAG_TOO_MANY_PRODS_CUT 4    /* TooManyProductionsCut, done */

/* Grid family: the MATPOWER-format case of Grid.__init__ (grid.py:40-93) in struct-of-arrays form.
 * 2*n_sub buses: bus s is node 0 of substation s, bus s+n_sub its artificial '666' sister (node 1). */
typedef struct ppn_case {
    int32_t n_sub, n_gen, n_load, n_line;
    double base_mva;
    const int32_t* sub_ids;       /* [n_sub]  external substation ids (observation tail only) */
    const int32_t* gen_sub;       /* [n_gen]  substation index of each generator, strictly ascending */
    const int32_t* load_sub;      /* [n_load] substation index of each load, strictly ascending */
    const int32_t* line_or_sub;   /* [n_line] */
    const int32_t* line_ex_sub;   /* [n_line] */
    const double* line_r;         /* [n_line] p.u. */
    const double* line_x;
    const double* line_b;         /* total line charging */
    const double* line_tap;       /* off-nominal ratio, 0 means 1 (makeYbus) */
    const uint8_t* line_status0;  /* [n_line] initial service status */
    const double* bus_gs;         /* [2*n_sub] MW at 1 p.u. */
    const double* bus_bs;         /* [2*n_sub] MVAr at 1 p.u. */
    const double* bus_basekv;     /* [2*n_sub] */
    const double* bus_vm0;        /* [2*n_sub] initial magnitudes, p.u. */
    const double* bus_va0;        /* [2*n_sub] initial angles, DEGREES (bus[:, VA]) */
    const double* gen_qmin;       /* [n_gen] */
    const double* gen_qmax;       /* [n_gen] */
    const double* gen_pg0;        /* [n_gen] case-file PG, QG, VG: the state before the first chronic row is loaded */
    const double* gen_qg0;
    const double* gen_vg0;
    const double* load_pd0;       /* [n_load] case-file PD, QD of the load buses */
    const double* load_qd0;
    int32_t slack_sub;            /* substation whose node-0 bus is type 3 in the case file (grid.py:74) */
    const double* thermal_limits; /* [n_line] A; the first chronic's imaps (game.py:301-304) */
} ppn_case;

/* configuration.yaml + RunEnv arguments (game.py:263-298, parameters.py:89-153) */
typedef struct ppn_config {
    int32_t dc;                          /* loadflow_mode == DC */
    double hard_overflow_coefficient;    /* 1e9 when without_overflow_cutoff */
    int32_t n_timesteps_hard_overflow_is_broken;
    double n_timesteps_consecutive_soft_overflow_breaks; /* 1e12 when without_overflow_cutoff */
    int32_t n_timesteps_soft_overflow_is_broken;
    int32_t n_timesteps_horizon_maintenance;
    int32_t max_number_prods_game_over;
    int32_t max_number_loads_game_over;
    int32_t n_timesteps_actionned_line_reactionable;
    int32_t n_timesteps_actionned_node_reactionable;
    int32_t max_number_actionned_substations;
    int32_t max_number_actionned_lines;
    int32_t max_number_actionned_total;
    int32_t hard_game_over;              /* game_over_mode == 'hard' */
    int32_t loop_mode;                   /* 0 natural, 1 random (counter-based hash, not numpy's stream), 2 fixed */
    double pf_tol;                       /* 1e-6  (grid.py:63) */
    int32_t pf_max_it;                   /* 25    (grid.py:63) */
    double reward_constant;              /* `constant` of the shipped CustomRewardSignal (14 / 30 / 118) */
    uint64_t seed;                       /* loop_mode 1 only */
    int32_t threads_per_env;             /* 0 = automatic (one warp per env up to 32 substations, else one CTA of 128 threads); 16, 32, 128 or 256 */
    int32_t pf_alg;                      /* PYPOWER's PF_ALG for AC: 0 or 2 = fast-decoupled XB (what the reference runs, grid.py:63);
                                            1 = Newton-Raphson (newtonpf, PF_MAX_IT 10, same pf_tol) -- an option the reference does not use */
} ppn_config;

/* One chronic (chronic.py:174-246): float32 tables with n_rows rows, planned tables ALREADY shifted by one row. */
typedef struct ppn_chronic {
    int32_t n_rows;
    const float* prods_p;          /* [n_rows, n_gen] */
    const float* prods_v;          /* [n_rows, n_gen]  kV, <= 0 means generator off */
    const float* loads_p;          /* [n_rows, n_load] */
    const float* loads_q;
    const float* prods_p_planned;
    const float* prods_v_planned;
    const float* loads_p_planned;
    const float* loads_q_planned;
    const float* maintenance;      /* [n_rows, n_line] integer-valued durations */
    const float* hazards;          /* [n_rows, n_line] */
    const int32_t* ids;            /* [n_rows] simu ids (unique) */
    const int32_t* datetimes;      /* [n_rows, 6] year month day hour minute second */
} ppn_chronic;

typedef struct ppn_env ppn_env;

/* state fields for ppn_get_state / ppn_set_state (device buffers, row-major [n_envs, width], see ppn_state_width) */
#define PPN_STATE_REAL 0      /* double: Vm[2S] | Va[2S] degrees | load P[L] | load Q[L] | gen Pg[G] | gen Qg[G] | gen Vg[G] */
#define PPN_STATE_TOPOLOGY 1  /* uint8 : prods node[G] | loads node[L] | lines or node[N] | lines ex node[N] | line status[N] | gen status[G] */
#define PPN_STATE_COUNTERS 2  /* int32 : reconnectable[N] | line reactionable[N] | soft-overflow count[N] | node reactionable[S] |
                                         cursor[4] = chronic index, row (-1 none yet, -2 just switched), next chronic, rng counter */

int ppn_create(const ppn_case* grid, const ppn_config* cfg, int n_envs, int device, ppn_env** out);
int ppn_load_chronics(ppn_env* env, int n_chronics, const ppn_chronic* host_tables);

/* Game.__init__ for every env (game.py:296-340): pristine grid, chronic chronic_idx[e] (NULL: 0), first row
 * played row0[e] (NULL: 0), first load-flow cascade.  obs_dev (may be NULL) receives the dynamic observation prefix,
 * flag_dev (may be NULL) PPN_FLAG_DIVERGING_LOADFLOW where that first cascade diverged -- the reference raises from
 * the constructor there (game.py:340); a batch cannot, so such envs immediately run process_game_over. */
int ppn_reset(ppn_env* env, const int32_t* chronic_idx_host, const int32_t* row0_host, double* obs_dev,
              int64_t obs_stride, int32_t* flag_dev, void* stream);

/* RunEnv.step for every env (environment.py:848-866).  act_dev uint8 [n_envs, action_length]; obs_dev double
 * [n_envs, obs_stride] or NULL (only the dynamic prefix, obs_dynamic_length values per row, is written; rows of
 * `done` envs are left untouched unless auto_reset); reward_dev double [n_envs, 5] (the shipped five-term reward);
 * done_dev uint8 [n_envs]; flag_dev int32 [n_envs]; illegal_dev uint8 [n_envs, 1+2*n_line+n_sub] or NULL:
 * has_too_much_activations | illegal reconnections | on-cooldown line switches | on-cooldown substations.
 * auto_reset != 0: envs that are done immediately run process_game_over (game.py:762-780) and their obs row
 * receives the post-reset observation (Runner.step semantics, runner.py:84-87). */
int ppn_step(ppn_env* env, const uint8_t* act_dev, double* obs_dev, int64_t obs_stride, double* reward_dev,
             uint8_t* done_dev, int32_t* flag_dev, uint8_t* illegal_dev, int auto_reset, void* stream);

/* RunEnv.simulate (environment.py:868-884, game.py:887-943) for n_candidates actions per env, no state commit.
 * act_dev [n_envs, n_candidates, action_length]; outputs have n_envs*n_candidates rows. */
int ppn_simulate(ppn_env* env, int n_candidates, const uint8_t* act_dev, double* obs_dev, int64_t obs_stride,
                 double* reward_dev, uint8_t* done_dev, int32_t* flag_dev, uint8_t* illegal_dev, void* stream);

/* RunEnv.process_game_over (environment.py:886-888) for envs with mask_dev[e] != 0 (NULL: all). */
int ppn_process_game_over(ppn_env* env, const uint8_t* mask_dev, double* obs_dev, int64_t obs_stride, void* stream);

/* Game.is_action_valid (game.py:755-760): valid_dev uint8 [n_envs]. */
int ppn_action_valid(ppn_env* env, const uint8_t* act_dev, uint8_t* valid_dev, void* stream);

/* Diagnostic, host only (no GPU): builds the sparse LDL^T tables of a grid (elimination order, fill pattern, per-level
 * update lists, packed row / column views, sparse / dense cut -- everything the kernels index with) exactly as
 * ppn_create does, and replays the kernels' factorisation and solves on the host against a dense Gaussian elimination,
 * on a random symmetric positive definite matrix over a random topology.  full = 0: one row per substation, 1: two
 * (split buses).  *max_err_out = largest deviation of the solutions; info_out (5 ints or NULL) = rows, off-diagonal
 * entries of L, levels, first dense level, rows of the dense block.  A check of the tables, not a compute path. */
int ppn_sparse_selfcheck(int n_sub, int n_line, const int32_t* line_or_sub, const int32_t* line_ex_sub, int full,
                         uint32_t seed, double* max_err_out, int32_t* info_out);

/* Optional: every following ppn_step also writes one packed row per env, reward[5] | done | flag as 7 doubles, into
 * pack_dev [n_envs][7] (device memory, caller-owned; NULL switches it off).  It is the row an env-sharded run gathers
 * over NCCL each step: written by the step kernel itself, no packing kernels between step and collective. */
int ppn_set_result_pack(ppn_env* env, double* pack_dev);

/* Diagnostic: every following ppn_step also writes, per env, trace_dev[e][0..3] = SM clock cycles the env's warp / CTA
 * spent in the call, load-flows, fast-decoupled iterations, restarts (device memory [n_envs][4] int64; NULL switches it
 * off).  A step lasts as long as its slowest env: this is the tool that shows which one and why (tools/env_trace.py). */
int ppn_set_env_trace(ppn_env* env, int64_t* trace_dev);

/* ---- env-sharded runs over the GPUs of one box (the reference is single-process: no counterpart) ----
 * The packed result rows can be written by the step kernel straight into a buffer on ANOTHER GPU (the collecting rank's)
 * over NVLink: ppn_peer_alloc creates such a buffer (zero-filled device memory + a 64-byte CUDA IPC handle to hand to the
 * other processes), ppn_peer_open maps it in a peer process (the pointer + an offset is what ppn_set_result_pack takes),
 * ppn_peer_signal stores `value` to a 64-bit counter after everything enqueued earlier on `stream` (release, system scope),
 * ppn_peer_wait blocks `stream` until flags_dev[0..n) are all >= value (acquire; the flags may live on a peer GPU),
 * ppn_peer_read enqueues a device -> host copy of a peer buffer.  pypownet_b200/sharding.py PeerGather is the host side. */
int ppn_peer_alloc(int device, uint64_t bytes, void** dev_out, uint8_t* handle_out /* [64] */);
int ppn_peer_open(int device, const uint8_t* handle /* [64] */, void** dev_out);
int ppn_peer_close(int device, void* dev_ptr);
int ppn_peer_free(int device, void* dev_ptr);
int ppn_peer_signal(int device, uint64_t* flag_dev, uint64_t value, void* stream);
int ppn_peer_wait(int device, const uint64_t* flags_dev, int n, uint64_t value, void* stream);
int ppn_peer_read(int device, void* host_dst, const void* dev_src, uint64_t bytes, void* stream);

/* Host-buffer form of ppn_step (what a host-side agent calls, RunEnv.step semantics, environment.py:848-866); returns
 * when every result is in the host buffers.  When every result buffer is page-locked (cudaMallocHost / cudaHostRegister /
 * torch pin_memory) there is ONE launch and no copy: the step kernel writes rewards / done / flags straight into host
 * memory, and the observation rows either the same way (rows above 4 KB) or through a device buffer that a small drain
 * kernel empties into the host buffer while the other envs still iterate (small rows: DESIGN.md section 5).  Otherwise the
 * batch is cut into chunks, each chunk runs  actions H2D -> step kernel -> results D2H  on its own stream so that copies
 * overlap the kernels of the other chunks; pageable buffers are staged.  act_host NULL =
 * do-nothing; obs_host NULL = no observation.  Without auto_reset the observation rows of envs that ended stay untouched
 * (the reference returns None).  Work enqueued earlier through the device-pointer calls is synchronised first. */
int ppn_step_host(ppn_env* env, const uint8_t* act_host, double* obs_host, int64_t obs_stride, double* reward_host,
                  uint8_t* done_host, int32_t* flag_host, uint8_t* illegal_host, int auto_reset);

/* ppn_step_host with float32 observation rows (obs_stride counts floats): half the bytes over PCIe -- each value is the
 * float64 one rounded once.  An extension for host-side agents that feed a float32 network; the reference's as_array is
 * float64 and ppn_step_host stays the drop-in.  Every result buffer must be page-locked. */
int ppn_step_host_f32(ppn_env* env, const uint8_t* act_host, float* obs_host, int64_t obs_stride, double* reward_host,
                      uint8_t* done_host, int32_t* flag_host, uint8_t* illegal_host, int auto_reset);

int ppn_state_width(const ppn_env* env, int field);   /* elements per env of a PPN_STATE_* field */
int ppn_get_state(ppn_env* env, int field, void* out_dev, void* stream);
int ppn_set_state(ppn_env* env, int field, const void* in_dev, void* stream);
/* static tail of Observation.as_array (environment.py:583-595), obs_length - obs_dynamic_length doubles, host */
int ppn_observation_static(ppn_env* env, double* out_host);

int ppn_n_envs(const ppn_env* env);
int ppn_action_length(const ppn_env* env);
int ppn_obs_length(const ppn_env* env);
int ppn_obs_dynamic_length(const ppn_env* env);
int ppn_device(const ppn_env* env);
/* number of kernel launches issued by this handle so far, and cumulative device counters:
 * out_host[0] load-flows, [1] fast-decoupled iterations, [2] env steps, [3] game-over resets, [4] max cascade depth,
 * [5] kernel launches, [6] shared-memory bytes per env, [7] threads per env, [8] most load-flows and [9] most
 * fast-decoupled iterations spent by one env in one call since the previous ppn_get_counters */
int ppn_get_counters(ppn_env* env, int64_t* out_host /* [10] */);
/* how deep the cascading-failure loop of game.py:503-589 went, cumulative over the env-steps played so far:
 * out_host[d] = env-steps whose cascade saw d rounds with an overflowed line, d = 0..6, out_host[7] = seven or more */
int ppn_get_cascade_histogram(ppn_env* env, int64_t* out_host /* [8] */);
const char* ppn_last_error(const ppn_env* env);
const char* ppn_build_info(void);
void ppn_destroy(ppn_env* env);

#ifdef __cplusplus
}
#endif
#endif /* PYPOWNET_B200_H */
