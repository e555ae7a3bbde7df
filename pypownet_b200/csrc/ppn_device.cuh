// Device-side data model of the batched pypownet step path (shared by ppn_kernels.cu and ppn_api.cu).
//
// One env = one substation-level grid copy.  Static grid data (PpnDevCase) and chronic tables (PpnDevChronics)
// are shared by all envs of a handle; the per-env state is three row-major arrays (PpnDevState) so that the
// threads that own an env (one warp, or one CTA) read and write it as contiguous, coalesced rows.
//
// Reference objects this replaces: the `mpc` dict of pypownet/grid.py:40-93 (bus/gen/branch tables), the
// counters of pypownet/game.py:306-327 and the chronic cursor of game.py:309-313, 333-334.
#pragma once
#include <stdint.h>

#define PPN_MODE_STEP 0       // RunEnv.step            (environment.py:848-866)
#define PPN_MODE_SIMULATE 1   // RunEnv.simulate        (environment.py:868-884), no commit
#define PPN_MODE_GAME_OVER 2  // RunEnv.process_game_over (environment.py:886-888) for masked envs
#define PPN_MODE_INIT 3       // Game.__init__          (game.py:296-340): pristine grid, first row, first cascade

#define PPN_BT_ISOLATED 0
#define PPN_BT_PQ 1
#define PPN_BT_PV 2
#define PPN_BT_REF 3

struct PpnDevCase {
    int S, G, L, N, NB, A, OBSD;
    int slack_bus;
    double base_mva;
    const int* gen_sub;       // [G]
    const int* load_sub;      // [L]
    const int* lor_sub;       // [N]
    const int* lex_sub;       // [N]
    const int* gen_of_sub;    // [S] generator index or -1
    const int* load_of_sub;   // [S] load index or -1
    const int* adj_ptr;       // [S+1] CSR over substations
    const int* adj;           // [2N]  line*2 + end (0 origin, 1 extremity)
    const int* elem_sub;      // [G+L+2N] substation of each topology element (grid.py:428-494)
    const double* line_y;     // [N][8] yff.re yff.im yft.re yft.im ytf.re ytf.im ytt.re ytt.im  (makeYbus)
    const double* line_bp;    // [N] 1/x           (B' weights, XB)
    const double* line_bdc;   // [N] 1/x/tap       (makeBdc)
    const double* bus_ysh_r;  // [NB] Gs/baseMVA
    const double* bus_ysh_i;  // [NB] Bs/baseMVA
    const double* bus_basekv; // [NB]
    const double* bus_vm0;    // [NB]
    const double* bus_va0;    // [NB] degrees
    const double* gen_qmin;   // [G]
    const double* gen_qmax;
    const double* gen_pg0;
    const double* gen_qg0;
    const double* gen_vg0;
    const double* load_pd0;   // [L]
    const double* load_qd0;
    const double* thermal;    // [N] amperes
    const uint8_t* line_status0;  // [N]
};

// All chronics of a handle in ONE table of 32-bit words: each row is the record an env reads per timestep
// (chronic.py:220-232 `TimestepEntries`), so a step is one contiguous read per env.
struct PpnDevChronics {
    const float* rows;
    int row_words;            // words per row (multiple of 4)
    int n_chronics;
    const int* row_off;       // [n_chronics] first row of chronic c in `rows`
    const int* n_rows;        // [n_chronics]
    const int* row_after_switch;  // [n_chronics] row played first after a chronic change (game.py:399, 492), -1: none
    const int* last_id_zero;  // [n_chronics] 1 when the last row's simu id is 0
    // word offsets inside a row
    int o_pp, o_pv, o_lp, o_lq, o_mt, o_hz, o_ppp, o_pvp, o_lpp, o_lqp, o_pm, o_dt;
};

struct PpnDevCfg {
    int dc;
    double hard_coef;
    int n_hard_broken;
    double n_soft_consec;
    int n_soft_broken;
    int max_prods_go, max_loads_go;
    int n_line_react, n_node_react;
    int max_sub, max_lines, max_total;
    int hard_mode, loop_mode;
    double tol;
    int max_it;
    double reward_k;
    unsigned long long seed;
    int max_reset_attempts;
};

struct PpnDevState {
    double* real;      // [B][rw]  Vm[NB] | Va[NB] deg | load P[L] | load Q[L] | gen Pg[G] | Qg[G] | Vg[G]
    uint8_t* topo;     // [B][tw]  gen node[G] | load node[L] | or node[N] | ex node[N] | status[N] | gen status[G]
    int32_t* cnt;      // [B][cw]  reconnectable[N] | line react[N] | soft count[N] | node react[S] | cursor[4]
    int rw, tw, cw;
};

struct PpnStepArgs {
    int mode;
    int n_envs;                 // envs of this launch
    int env_off;                // first env of this launch (chunked launches of the host-buffer entry point); the
                                // output pointers below are already offset to its first row
    int n_cand;                 // simulate: candidates per env (else 1)
    int auto_reset;
    const uint8_t* act;         // [n_envs*n_cand][A] or NULL (do-nothing)
    const uint8_t* mask;        // PPN_MODE_GAME_OVER: [n_envs] or NULL
    const int32_t* init_chronic;  // PPN_MODE_INIT: [n_envs] or NULL
    const int32_t* init_row0;     // PPN_MODE_INIT: [n_envs] or NULL
    double* obs;                // [rows][obs_stride] or NULL
    long long obs_stride;
    int obs_bulk;               // 1: rows leave through one TMA bulk store each (16-byte aligned rows)
    double* reward;             // [rows][5] or NULL
    uint8_t* done;              // [rows] or NULL
    int32_t* flag;              // [rows] or NULL
    uint8_t* illegal;           // [rows][1+2N+S] or NULL
    double* ws;                 // global workspace for matrices that do not fit the shared-memory budget
    long long ws_stride;        // doubles per env
    int mat_cap;                // doubles of shared memory per env for B' and B''
    unsigned long long* stats;  // [8] or NULL
};

// Shared-memory footprint of one env (bytes), excluding the matrix area.  Must match the carve-up in ppn_kernels.cu.
static inline __host__ __device__ int ppn_env_smem_fixed_bytes(int S, int G, int L, int N, int tpe) {
    const int NB = 2 * S;
    const int A = G + L + 3 * N;
    const int nw = (tpe + 31) / 32;
    const int un = 4 * NB > 5 * N ? 4 * NB : 5 * N;     // FD work arrays share storage with the branch results
    int dbl = 8 * NB + un + 4 * N + 2 * L + 4 * G + nw * 2;
    int i32 = 3 * N + S + 4 + nw * 2 + 8;
    int i16 = 2 * N + 4 * NB + G + L + 4 * N;
    int u8 = (2 * G + L + 3 * N) + 2 * NB + N + A + S + (1 + 2 * N + S) + NB;
    int bytes = dbl * 8 + i32 * 4 + ((i16 * 2 + 3) & ~3) + ((u8 + 7) & ~7);
    return (bytes + 15) & ~15;
}
