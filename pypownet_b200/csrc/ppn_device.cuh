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

// Static symbolic structure of the sparse LDL^T factorisation of B' / B'' / Bdc (all share the pattern of the line
// graph; SHIFT = 0 makes them symmetric).  Rows follow a minimum-degree elimination order of the substation graph,
// computed once per grid on the host.  Two structures per grid: U = one row per substation (no bus is split, the
// common case), F = two adjacent rows per substation (bus s and its sister s+S).  Rows of buses that take no part in a
// system (isolated, reference, PV for B'') are identity rows, so the pattern never changes with the topology.
struct PpnDevSparse {
    int n;                    // rows
    int nnz;                  // off-diagonal entries of L
    int n_lev;                // levels of the elimination tree
    int cut_lev;              // hybrid solver: levels >= cut_lev (the narrow top of the elimination tree) form a dense block
    int cut_row;              // first row of that block (rows are sorted by level)
    int cut_ent;              // first entry whose column belongs to the block (entries are column-major)
    int nt;                   // rows of the block
    int blob_words;           // 32-bit words of the table blob
    int hyb_words;            // words of its prefix the hybrid solver needs (the tables it uses come first; = blob_words
                              // when the warp schedule below is not available)
    int o_wsched;             // int, or -1: the hybrid SOLVE as a schedule for ONE warp, a row per lane, every step's
                              // entries laid out lane-major so that no index load depends on another one:
                              //   [0] forward steps (sparse levels 1..cut-1, then the rows of the top block gathering
                              //   the columns below the cut), [1] backward steps (levels cut-1..0), then per step
                              //   (first row | rows << 16), (entries per row | halfword offset << 8), then the 16-bit
                              //   entries (entry id << 7 | column or row), 0xffff = none, 32 per (step, entry slot)
    const int* blob;          // every table below in one contiguous block, so that a CTA can stage it in shared memory
    // word offsets into the blob (int tables first, then the short tables)
    int o_colptr;             // int   [n+1]  off-diagonal entries of L by column, rows ascending
    int o_lev_ptr;            // int   [n_lev+1] into lev_ent
    int o_lev_ent;            // int   targets of each level: entry id e < nnz, or nnz + column for a diagonal
    int o_trip_ptr;           // int   [nnz+n+1] update terms of each target
    int o_trip;               // int   (entry (i,k) << 16) | entry (j,k): target (i,j) -= T(i,k) L(j,k), k ascending
    int o_line_pos;           // int   U: [N] entry of the line's off-diagonal term; F: [N][2][2] by (origin node, extremity node)
    int o_rowptr;             // int   [n+1] entries of L by row (for the forward substitution), columns ascending
    int o_rowent;             // int   [nnz] entry ids
    int o_lev_rows_ptr;       // int   [n_lev+1] rows (= columns) of each level of the elimination tree
    int o_lev_rows;           // int   [n]
    int o_rpack;              // int   [nnz] row-wise entries packed as (8 * entry id << 16) | 8 * column (byte offsets)
    int o_rowpk;              // int   [n] (first row-wise entry << 8) | number of entries of the row
    int o_colpk;              // int   [n] (first column entry << 8) | number of entries of the column
    int o_rowoff;             // short [nnz] 8 * row of each column-wise entry (byte offset into the solve vector)
    int o_row_lev;            // short [n] level of each row
    int o_rowidx;             // short [nnz]
    int o_ecol;               // short [nnz] column of each entry
    int o_parent;             // short [n] elimination tree (-1 root)
    int o_bus_row;            // short [NB] row of bus b, -1 when the bus has no row (sisters in U)
};

// doubles of per-env storage of ONE factor: T[nnz], L[nnz], d[n] (doubles), entry / row offsets into the inverse
// (int nnz + n), compact index of each row (short n)
static inline __host__ __device__ int ppn_sp_factor_doubles(int n, int nnz) {
    return 2 * nnz + n + (nnz + n + 1) / 2 + (n + 3) / 4;
}

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
    PpnDevSparse sp[2];       // [0] U (un-split grid), [1] F (any bus split)
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
    int alg;          // 2 (or 0): fast-decoupled XB, what the reference runs; 1: Newton-Raphson
    int max_it_nr;    // PYPOWER's PF_MAX_IT (10)
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
    int obs_f32;                // 1: `obs` holds float rows (obs_stride counts floats): ppn_step_host_f32
    double* reward;             // [rows][5] or NULL
    uint8_t* done;              // [rows] or NULL
    int32_t* flag;              // [rows] or NULL
    uint8_t* illegal;           // [rows][1+2N+S] or NULL
    double* pack;               // [rows][7] reward[5] | done | flag as doubles (the row a sharded run all-gathers) or NULL
    double* ws;                 // global workspace for matrices that do not fit the shared-memory budget
    long long ws_stride;        // doubles per env
    long long ws_dense;         // sparse solver: offset of the factor storage inside the env's slice
    int sparse;                 // 0: dense Gauss-Jordan inverses; 1: sparse LDL^T on the static pattern, then explicit
                                // inverses; 2: sparse LDL^T and level-scheduled triangular solves every half-iteration;
                                // 3: hybrid -- sparse LDL^T for the wide bottom levels of the elimination tree, explicit
                                // dense inverse of the Schur complement of its narrow top (a few dozen rows), so a solve
                                // is a handful of wide parallel steps and no full dense matrix is ever stored
    int mat_cap;                // doubles of shared memory per env for B' and B''
    unsigned long long* stats;  // [8] or NULL
    unsigned* row_flag;         // [rows] or NULL: after a row's results are complete in (device) memory, its thread 0
    unsigned epoch;             // release-stores `epoch` there (| 0x80000000 when the row carries no observation): the
                                // signal the drain kernel of ppn_step_host waits for (ppn_api.cu)
    long long* trace;           // [rows][4] or NULL: clock cycles, load-flows, fast-decoupled iterations and restarts each
                                // launch row spent in this call (ppn_set_env_trace: where does a step's time go)
    int* split_flag;            // page-locked host word (device alias) set to 1 when an env applies a node switch:
                                // tells the host that buses may be split from now on (shared-memory plan), or NULL
};

// Shared-memory footprint of one env (bytes), excluding the matrix area.  Must match the carve-up in ppn_kernels.cu.
static inline __host__ __device__ int ppn_env_smem_fixed_bytes(int S, int G, int L, int N, int tpe) {
    const int NB = 2 * S;
    const int A = G + L + 3 * N;
    const int nw = (tpe + 31) / 32;
    const int un = ((4 * NB > 5 * N ? 4 * NB : 5 * N) + 1) & ~1;   // FD work arrays share storage with the branch results (even: 16-byte pairs behind it)
    int dbl = 8 * NB + un + 4 * N + 2 * L + 4 * G + nw * 2;
    int i32 = 3 * N + S + 4 + nw * 2 + 8;
    int i16 = 2 * N + 4 * NB + G + L + 4 * N;
    int u8 = (2 * G + L + 3 * N) + 2 * NB + N + A + S + (1 + 2 * N + S) + NB;
    int bytes = dbl * 8 + i32 * 4 + ((i16 * 2 + 3) & ~3) + ((u8 + 7) & ~7);
    return (bytes + 15) & ~15;
}
