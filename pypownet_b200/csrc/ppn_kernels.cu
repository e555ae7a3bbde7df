// Batched pypownet step path as one fused sm_100a kernel: every env (grid copy) is owned by TPE threads -- half a warp
// for IEEE-14, one warp for IEEE-30 sized grids, one CTA for IEEE-118 -- that keep the whole working set of the
// timestep in shared memory: topology, bus types, fast-decoupled B'/B'' inverses, voltages, flows, counters.  HBM
// traffic per env-step is the state row in, one chronic record in, the action in, and state row + observation +
// reward/done/flag out.
//
// Reference path (pypownet @ /root/reference):
//   Game.step / apply_action / _verify_illegal_action        pypownet/game.py:591-753, 799-885
//   load_entries_from_next_timestep / _timestep_id           pypownet/game.py:405-501, pypownet/grid.py:266-311
//   _compute_loadflow_cascading                              pypownet/game.py:503-589
//   Grid.compute_loadflow, _synchronize_bus_types            pypownet/grid.py:140-264
//   runpf(PF_ALG=2: fast-decoupled XB, 25 it, tol 1e-6) / rundcpf   PYPOWER 5.1.4 via grid.py:62-65, 226-231
//   extract_flows_a                                          pypownet/grid.py:112-138
//   process_game_over / reset_grid / simulate                pypownet/game.py:762-797, 887-943
//   export_observation / Observation.as_array                pypownet/game.py:945-978, environment.py:451-517
//   CustomRewardSignal.compute_reward                        parameters/default14/reward_signal.py:45-169
// The arithmetic is the restated PYPOWER algorithm of SURVEY.md Appendix A; linear algebra is a dense in-shared-memory
// Gauss-Jordan inverse of B' and B'' (symmetric, positive definite for connected grids without phase shifters).
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include "ppn_device.cuh"

#define PPN_FULL 0xffffffffu
#define PPN_SQRT3 1.7320508075688772
#define PPN_PI 3.14159265358979323846

#ifdef PPN_TIMING
// Phase timing of the first load-flow of output row 0 (debug builds only: tools/phase_timing.py)
__device__ long long ppn_timing_buf[64];
#define PPN_TICK(id) do { if (slot == 0 && tid == 0 && ppn_timing_buf[63] == 0) ppn_timing_buf[id] = clock64(); } while (0)
#define PPN_TICK_ACC(id, t0) do { if (slot == 0 && tid == 0 && ppn_timing_buf[63] == 0) ppn_timing_buf[id] += clock64() - (t0); } while (0)
#else
#define PPN_TICK(id)
#define PPN_TICK_ACC(id, t0)
#endif

namespace {

// ---------------------------------------------------------------------------------------------- env-wide primitives
// An env is owned by a "group": TPE <= 32 lanes of one warp (mask = the group's lanes), or a whole CTA (TPE > 32).
template <int TPE> __device__ __forceinline__ void env_sync(unsigned mask) {
    if (TPE <= 32) __syncwarp(mask); else __syncthreads();
}

template <int TPE> __device__ __forceinline__ bool env_any(bool p, unsigned mask) {
    if (TPE <= 32) return __any_sync(mask, p);
    return __syncthreads_or(p) != 0;
}

// NaN-propagating maximum over the env's threads, identical in every thread, fixed order (deterministic).
template <int TPE> __device__ __forceinline__ double env_max_nan(double v, double* red, int tid, unsigned mask) {
#pragma unroll
    for (int o = (TPE < 32 ? TPE : 32) / 2; o > 0; o >>= 1) {
        double w = __shfl_xor_sync(TPE <= 32 ? mask : PPN_FULL, v, o);
        v = (w > v || w != w) ? w : v;
    }
    if (TPE <= 32) return v;
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double r = red[0];
#pragma unroll
    for (int k = 1; k < (TPE + 31) / 32; k++) {
        double w = red[k];
        r = (w > r || w != w) ? w : r;
    }
    return r;
}

template <int TPE> __device__ __forceinline__ int env_sum_int(int v, int* red, int tid, unsigned mask) {
    v = __reduce_add_sync(TPE <= 32 ? mask : PPN_FULL, v);
    if (TPE <= 32) return v;
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    int r = 0;
#pragma unroll
    for (int k = 0; k < (TPE + 31) / 32; k++) r += red[k];
    return r;
}

template <int TPE> __device__ __forceinline__ int env_min_int(int v, int* red, int tid, unsigned mask) {
    v = __reduce_min_sync(TPE <= 32 ? mask : PPN_FULL, v);
    if (TPE <= 32) return v;
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    int r = red[0];
#pragma unroll
    for (int k = 1; k < (TPE + 31) / 32; k++) r = min(r, red[k]);
    return r;
}

template <int TPE> __device__ __forceinline__ double env_sum_double(double v, double* red, int tid, unsigned mask) {
#pragma unroll
    for (int o = (TPE < 32 ? TPE : 32) / 2; o > 0; o >>= 1) v += __shfl_xor_sync(TPE <= 32 ? mask : PPN_FULL, v, o);
    if (TPE <= 32) return v;
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double r = 0;
#pragma unroll
    for (int k = 0; k < (TPE + 31) / 32; k++) r += red[k];
    return r;
}

// ------------------------------------------------------------------------------------------------------ env context
// Grid sizes: compile-time for the three IEEE families (array offsets fold into the instructions, loops unroll),
// run-time for any other grid.
template <int S_, int G_, int L_, int N_> struct StaticDims {
    static constexpr int S = S_, G = G_, L = L_, N = N_, NB = 2 * S_, A = G_ + L_ + 3 * N_;
    static constexpr int NB_MAX = 2 * S_;   // compile-time bound on the buses of an env
    __device__ __forceinline__ void init_dims(const PpnDevCase&) {}
};
struct DynDims {
    static constexpr int NB_MAX = 0;        // unknown at compile time
    int S, G, L, N, NB, A;
    __device__ __forceinline__ void init_dims(const PpnDevCase& c) { S = c.S; G = c.G; L = c.L; N = c.N; NB = c.NB; A = c.A; }
};

// Shared-memory image of one env.  Layout (must match ppn_env_smem_fixed_bytes): doubles | int32 | int16 | bytes |
// pad to 16 | matrix area.  Every array is `base + offset(dims)`; nothing but `base` lives in a register.
template <int TPE, class D> struct Env : D {
    int tid;
    unsigned mask;   // lanes of this env's group (TPE <= 32)
    int shift;       // first lane of the group
    unsigned char* base;
    int fixed_bytes;
    static constexpr int NW = (TPE + 31) / 32;
#define PPN_DBL(name, expr) __device__ __forceinline__ double* name() const { return reinterpret_cast<double*>(base) + (expr); }
    // U: the fast-decoupled work arrays (P, Q mismatches, Ybus diagonal) share their storage with the branch results
    // (flows, amperes), which are only written once the iteration is over.
    __device__ __forceinline__ int n_union() const { return ((4 * this->NB > 5 * this->N ? 4 * this->NB : 5 * this->N) + 1) & ~1; }
    // vri: rectangular voltages as (re, im) PAIRS, bus b at [2b], [2b+1] -- one 16-byte load per neighbour in the mismatch
    // gather; the DC path uses the first NB entries as its angle vector (theta)
    PPN_DBL(vm, 0) PPN_DBL(va, this->NB) PPN_DBL(vri, 2 * this->NB) PPN_DBL(theta, 2 * this->NB)
    PPN_DBL(pin, 4 * this->NB) PPN_DBL(qin, 5 * this->NB) PPN_DBL(cs, 6 * this->NB) PPN_DBL(sn, 7 * this->NB)
    PPN_DBL(P, 8 * this->NB) PPN_DBL(Q, 9 * this->NB) PPN_DBL(ydr, 10 * this->NB) PPN_DBL(ydi, 11 * this->NB)
    PPN_DBL(pf, 8 * this->NB) PPN_DBL(qf, 8 * this->NB + this->N) PPN_DBL(pt, 8 * this->NB + 2 * this->N)
    PPN_DBL(qt, 8 * this->NB + 3 * this->N) PPN_DBL(amp, 8 * this->NB + 4 * this->N)
    // ey: off-diagonal admittance of every line-end entry as a (re, im) pair, entry k at [2k], [2k+1]
    PPN_DBL(ey, 8 * this->NB + n_union())
    PPN_DBL(lpd, 8 * this->NB + n_union() + 4 * this->N) PPN_DBL(lqd, 8 * this->NB + n_union() + 4 * this->N + this->L)
    PPN_DBL(gpg, 8 * this->NB + n_union() + 4 * this->N + 2 * this->L)
    PPN_DBL(gqg, 8 * this->NB + n_union() + 4 * this->N + 2 * this->L + this->G)
    PPN_DBL(gvg, 8 * this->NB + n_union() + 4 * this->N + 2 * this->L + 2 * this->G)
    PPN_DBL(gkv, 8 * this->NB + n_union() + 4 * this->N + 2 * this->L + 3 * this->G)
    PPN_DBL(redd, 8 * this->NB + n_union() + 4 * this->N + 2 * this->L + 4 * this->G)
#undef PPN_DBL
    __device__ __forceinline__ int n_dbl() const { return 8 * this->NB + n_union() + 4 * this->N + 2 * this->L + 4 * this->G + 2 * NW; }
#define PPN_I32(name, expr) __device__ __forceinline__ int* name() const { return reinterpret_cast<int*>(base + 8 * n_dbl()) + (expr); }
    PPN_I32(recon, 0) PPN_I32(lreact, this->N) PPN_I32(soft, 2 * this->N) PPN_I32(nreact, 3 * this->N)
    PPN_I32(cursor, 3 * this->N + this->S) PPN_I32(redi, 3 * this->N + this->S + 4) PPN_I32(misc, 3 * this->N + this->S + 4 + 2 * NW)
#undef PPN_I32
    __device__ __forceinline__ int n_i32() const { return 3 * this->N + this->S + 4 + 2 * NW + 8; }
#define PPN_I16(name, expr) __device__ __forceinline__ short* name() const { return reinterpret_cast<short*>(base + 8 * n_dbl() + 4 * n_i32()) + (expr); }
    PPN_I16(fbus, 0) PPN_I16(tbus, this->N) PPN_I16(idxp, 2 * this->N) PPN_I16(idxq, 2 * this->N + this->NB)
    PPN_I16(busp, 2 * this->N + 2 * this->NB) PPN_I16(busq, 2 * this->N + 3 * this->NB)
    PPN_I16(gbus, 2 * this->N + 4 * this->NB) PPN_I16(lbus, 2 * this->N + 4 * this->NB + this->G)
    PPN_I16(eline, 2 * this->N + 4 * this->NB + this->G + this->L) PPN_I16(eoth, 4 * this->N + 4 * this->NB + this->G + this->L)
#undef PPN_I16
    __device__ __forceinline__ int n_i16() const { return 6 * this->N + 4 * this->NB + this->G + this->L; }
#define PPN_U8(name, expr) __device__ __forceinline__ uint8_t* name() const { return base + 8 * n_dbl() + 4 * n_i32() + ((2 * n_i16() + 3) & ~3) + (expr); }
    // topo row: gnode | lnode | onode | enode | status | gstat
    PPN_U8(gnode, 0) PPN_U8(lnode, this->G) PPN_U8(onode, this->G + this->L) PPN_U8(enode, this->G + this->L + this->N)
    PPN_U8(status, this->G + this->L + 2 * this->N) PPN_U8(gstat, this->G + this->L + 3 * this->N)
    PPN_U8(btype, 2 * this->G + this->L + 3 * this->N) PPN_U8(mark, 2 * this->G + this->L + 3 * this->N + this->NB)
    PPN_U8(over, 2 * this->G + this->L + 3 * this->N + 2 * this->NB) PPN_U8(act, 2 * this->G + this->L + 4 * this->N + 2 * this->NB)
    PPN_U8(subch, 2 * this->G + this->L + 4 * this->N + 2 * this->NB + this->A)
    PPN_U8(ill, 2 * this->G + this->L + 4 * this->N + 2 * this->NB + this->A + this->S)
    PPN_U8(deg, 2 * this->G + this->L + 6 * this->N + 2 * this->NB + this->A + 2 * this->S + 1)
#undef PPN_U8
    __device__ __forceinline__ double* mat() const { return reinterpret_cast<double*>(base + fixed_bytes); }
};

// ------------------------------------------------------------------------------------------------- dense inverse
// In-place Gauss-Jordan inverse without pivoting of the n x n matrix a (row stride ld, odd => conflict-free column
// walks).  Rows are spread over the lanes of the group (TPE <= 32) or over lanes with columns over warps (CTA).
// A zero/NaN pivot yields inf/NaN, which the caller's mismatch test turns into "diverging" (the reference's splu
// raises on an exactly singular factor).
template <int TPE, int MAXR> __device__ void gj_invert(double* a, int n, int ld, int tid, unsigned mask) {
    constexpr int RW = TPE < 32 ? TPE : 32;                 // lanes that share the rows
    constexpr int W = TPE <= 32 ? 1 : TPE / 32;             // column stripes
    const int lane = TPE <= 32 ? tid : (tid & 31), w = TPE <= 32 ? 0 : (tid >> 5);
    for (int k = 0; k < n; k++) {
        const double p = 1.0 / a[k * ld + k];
        double cm[MAXR];
#pragma unroll
        for (int r = 0; r < MAXR; r++) {
            const int i = lane + RW * r;
            cm[r] = (i < n && i != k) ? a[i * ld + k] * p : 0.0;
        }
        env_sync<TPE>(mask);
#pragma unroll
        for (int r = 0; r < MAXR; r++) {
            const int i = lane + RW * r;
            if (i < n && i != k) {
                const double ci = cm[r];
                double* ai = a + i * ld;
                const double* ak = a + k * ld;
                // column k is fixed just below; four elements per trip, loads first (ai and ak never overlap: i != k)
                int j = w;
                for (; j + 3 * W < n; j += 4 * W) {
                    const double k0 = ak[j], k1 = ak[j + W], k2 = ak[j + 2 * W], k3 = ak[j + 3 * W];
                    const double x0 = ai[j], x1 = ai[j + W], x2 = ai[j + 2 * W], x3 = ai[j + 3 * W];
                    ai[j] = fma(-ci, k0, x0); ai[j + W] = fma(-ci, k1, x1);
                    ai[j + 2 * W] = fma(-ci, k2, x2); ai[j + 3 * W] = fma(-ci, k3, x3);
                }
                for (; j < n; j += W) ai[j] = fma(-ci, ak[j], ai[j]);
                if (w == k % W) ai[k] = -ci;
            }
        }
        env_sync<TPE>(mask);
        for (int j = tid; j < n; j += TPE) a[k * ld + j] = (j == k) ? p : a[k * ld + j] * p;
        env_sync<TPE>(mask);
    }
}

// Explicit shared-memory accesses (32-bit shared-window addresses): the matrices may also live in the HBM workspace,
// so their pointers are generic; these keep the common in-shared-memory path on LDS/STS.
__device__ __forceinline__ unsigned saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds64(unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts64(unsigned a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

// Gauss-Jordan inverse of a matrix in SHARED memory with every row held in the registers of one lane.  Per pivot
// step the pivot row goes through a small shared buffer (one store by its owner, broadcast loads by everyone) and the
// rank-1 update runs on registers.  `i` = row of this lane, `nsteps` is uniform over `syncmask` (max n of the
// matrices inverted side by side), `piv_s` = NR doubles per matrix.  Same operation order as gj_invert.
// NR = 16: pivot steps fully unrolled (static register indices, no select chains); two matrices fit one warp.
// Eight consecutive doubles from a 16-byte aligned shared address in ONE asm statement: every load has registers of its
// own and they issue back to back (separate volatile loads keep their program order and end up sharing one register
// pair, i.e. a load -> FMA -> load chain of ~35 cycles per element: it was most of a pivot step)
__device__ __forceinline__ void lds_8(unsigned a, double (&p)[8]) {
    asm volatile(
        "ld.shared.v2.f64 {%0, %1}, [%8];\n\t"
        "ld.shared.v2.f64 {%2, %3}, [%8+16];\n\t"
        "ld.shared.v2.f64 {%4, %5}, [%8+32];\n\t"
        "ld.shared.v2.f64 {%6, %7}, [%8+48];\n\t"
        : "=d"(p[0]), "=d"(p[1]), "=d"(p[2]), "=d"(p[3]), "=d"(p[4]), "=d"(p[5]), "=d"(p[6]), "=d"(p[7])
        : "r"(a));
}

// 1/a without the special-case branch of __drcp_rn (pivots are ordinary numbers; a zero pivot still gives inf/NaN, i.e.
// "diverging"): hardware seed + the same two refinement steps, within an ulp -- branch-free, so it interleaves with the
// independent FMAs of the elimination
__device__ __forceinline__ double rcp_plain(double a) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    double e = fma(-a, x, 1.0);
    e = fma(e, e, e);
    x = fma(x, e, x);
    e = fma(-a, x, 1.0);
    return fma(x, e, x);
}

template <int NR>
__device__ __forceinline__ void gj_rows_in_registers_impl(unsigned m_s, int n, int ld, int i, int nsteps, unsigned piv_s,
                                                          unsigned syncmask) {
    static_assert(NR % 8 == 0, "rows are read eight doubles at a time");
    double row[NR];
    const bool mine = i < n;
    piv_s = (piv_s + 15u) & ~15u;   // the pivot row buffer holds NR + 1 doubles
#pragma unroll
    for (int j = 0; j < NR; j++) row[j] = (mine && j < n) ? lds64(m_s + 8u * (unsigned)(i * ld + j)) : 0.0;
    // The reciprocal of a pivot is the other long chain of a step: every lane updates the column of the NEXT pivot first
    // and starts that reciprocal at once (only lane k+1's is used), so it overlaps the rest of the elimination; the owner
    // publishes its row already scaled and the others eliminate with their plain entry.
    double pinv = rcp_plain(row[0]);
#pragma unroll
    for (int k = 0; k < NR; k++) {
        if (k < nsteps) {
            const bool act = mine && k < n;
            if (act && i == k) {
#pragma unroll
                for (int j = 0; j < NR; j++) {
                    row[j] = (j == k) ? pinv : row[j] * pinv;
                    sts64(piv_s + 8u * j, row[j]);
                }
            }
            __syncwarp(syncmask);
            if (act && i != k) {
                const double ci = row[k];
                row[k] = 0.0;   // becomes -ci * pinv below
                constexpr int NBATCH = NR / 8;
                const int first = (k + 1 < NR) ? (k + 1) / 8 : 0;   // the batch that holds the next pivot's column goes first
#pragma unroll
                for (int bb = 0; bb < NBATCH; bb++) {
                    const int b8 = (first + bb) % NBATCH;
                    double pv[8];
                    lds_8(piv_s + 64u * b8, pv);
                    if (bb == 0 && k + 1 < NR) {
                        row[k + 1] = fma(-ci, pv[(k + 1) % 8], row[k + 1]);
                        pinv = rcp_plain(row[k + 1]);
                    }
#pragma unroll
                    for (int q = 0; q < 8; q++)
                        if (!(bb == 0 && k + 1 < NR && q == (k + 1) % 8)) row[8 * b8 + q] = fma(-ci, pv[q], row[8 * b8 + q]);
                }
            }
            __syncwarp(syncmask);
        }
    }
    if (mine) {
#pragma unroll
        for (int j = 0; j < NR; j++)
            if (j < n) sts64(m_s + 8u * (unsigned)(i * ld + j), row[j]);
    }
    __syncwarp(syncmask);
}

__device__ __noinline__ void gj16_rows_in_registers(unsigned m_s, int n, int ld, int i, int nsteps, unsigned piv_s,
                                                    unsigned syncmask) {
    gj_rows_in_registers_impl<16>(m_s, n, ld, i, nsteps, piv_s, syncmask);
}

// NR = 24: the dense top block of the hybrid solver (IEEE-118: 17 rows), one warp per matrix, a row per lane
__device__ __noinline__ void gj24_rows_in_registers(unsigned m_s, int n, int ld, int i, unsigned piv_s) {
    gj_rows_in_registers_impl<24>(m_s, n, ld, i, n, piv_s, PPN_FULL);
}

// ------------------------------------------------------------------------------------- sparse LDL^T (static pattern)
// Index tables of one structure (PpnDevSparse), read from the global blob or from its copy in shared memory.
struct SpView {
    int n, nnz, n_lev;
    bool full;   // F structure: two rows per substation
    const PpnDevSparse* d;   // the structure (table offsets)
    unsigned tb;             // shared-window address of the staged tables when tables AND factors are in shared
                             // memory (the LDS/STS code path of a CTA-per-env kernel), else 0
    bool hyb;                // hybrid factor (dense inverse of the top block) -- needs tb
    const int *colptr, *lev_ptr, *lev_ent, *trip_ptr, *trip, *line_pos, *rowptr, *rowent, *lev_rows_ptr, *lev_rows;
    const short *rowidx, *ecol, *parent, *bus_row;
};

__device__ __forceinline__ SpView sp_view(const PpnDevSparse& sp, const int* base, int S) {
    SpView v;
    v.n = sp.n; v.nnz = sp.nnz; v.n_lev = sp.n_lev; v.full = sp.n > S; v.d = &sp; v.tb = 0; v.hyb = false;
    v.colptr = base + sp.o_colptr; v.lev_ptr = base + sp.o_lev_ptr; v.lev_ent = base + sp.o_lev_ent;
    v.trip_ptr = base + sp.o_trip_ptr; v.trip = base + sp.o_trip; v.line_pos = base + sp.o_line_pos;
    v.rowptr = base + sp.o_rowptr; v.rowent = base + sp.o_rowent;
    v.lev_rows_ptr = base + sp.o_lev_rows_ptr; v.lev_rows = base + sp.o_lev_rows;
    v.rowidx = reinterpret_cast<const short*>(base + sp.o_rowidx); v.ecol = reinterpret_cast<const short*>(base + sp.o_ecol);
    v.parent = reinterpret_cast<const short*>(base + sp.o_parent); v.bus_row = reinterpret_cast<const short*>(base + sp.o_bus_row);
    return v;
}

// Storage of one factor (ppn_sp_factor_doubles): T[nnz] (entries times the pivot of their column), Lv[nnz] (unit lower
// factor; holds the lower triangle of the matrix on entry), dg[n] (diagonal of the matrix on entry, pivots after the
// factorisation, their reciprocals during the inversion), eoff[nnz] / koff[n] (offset into a column of the inverse of
// each entry's row / of each row), cp[n] (compact index of each row's bus in the system being solved, or n_act for
// rows that take no part = identity rows).
struct SpFactor {
    double* T; double* Lv; double* dg; int* eoff; int* koff; short* cp;
};

__device__ __forceinline__ SpFactor sp_carve(double* base, const SpView& sp, int which) {
    SpFactor f;
    double* p = base + which * ppn_sp_factor_doubles(sp.n, sp.nnz);
    f.T = p; f.Lv = p + sp.nnz; f.dg = p + 2 * sp.nnz;
    f.eoff = reinterpret_cast<int*>(p + 2 * sp.nnz + sp.n);
    f.koff = f.eoff + sp.nnz;
    f.cp = reinterpret_cast<short*>(p + 2 * sp.nnz + sp.n + (sp.nnz + sp.n + 1) / 2);
    return f;
}

// identity start of a factor: no couplings, unit diagonal; cp = n_act for every row
template <int TPE> __device__ __forceinline__ void sp_clear(const SpView& sp, const SpFactor& f, int n_act, int tid) {
    for (int i = tid; i < sp.nnz; i += TPE) f.Lv[i] = 0.0;
    for (int i = tid; i < sp.n; i += TPE) { f.dg[i] = 1.0; f.cp[i] = (short)n_act; }
}

// Left-looking numeric factorisation by levels of the elimination tree: every target (i,j) of a level gathers its
// update terms in a fixed order (deterministic, no atomics), then the level's columns are scaled by their pivots.
// NF = 2: B' and B'' side by side, one half of the env's threads each (same structure, shared barriers).
template <int TPE, int NF> __device__ __forceinline__ void sp_factor(const SpView& sp, const SpFactor& f1, const SpFactor& f2, int tid,
                                                                     unsigned mask) {
    constexpr int H = TPE / NF;
    const SpFactor& f = (NF == 1 || tid < H) ? f1 : f2;
    const int ht = (NF == 1 || tid < H) ? tid : tid - H;
    const int nnz = sp.nnz;
    for (int lv = 0; lv < sp.n_lev; lv++) {
        const int a = sp.lev_ptr[lv], b = sp.lev_ptr[lv + 1];
        for (int q = a + ht; q < b; q += H) {
            const int id = sp.lev_ent[q];
            double acc = id < nnz ? f.Lv[id] : f.dg[id - nnz];
            const int t1 = sp.trip_ptr[id + 1];
            for (int t = sp.trip_ptr[id]; t < t1; t++) {
                const unsigned pk = (unsigned)sp.trip[t];
                acc = fma(-f.T[pk >> 16], f.Lv[pk & 0xffffu], acc);
            }
            if (id < nnz) f.T[id] = acc; else f.dg[id - nnz] = acc;
        }
        env_sync<TPE>(mask);
        for (int q = a + ht; q < b; q += H) {
            const int id = sp.lev_ent[q];
            if (id < nnz) f.Lv[id] = f.T[id] / f.dg[sp.ecol[id]];
        }
        env_sync<TPE>(mask);
    }
}

// pivots -> reciprocal pivots (once per factorisation, before the solves)
template <int TPE> __device__ __forceinline__ void sp_recip(const SpView& sp, const SpFactor& f1, const SpFactor& f2, bool two, int tid,
                                                            unsigned mask) {
    for (int i = tid; i < sp.n; i += TPE) {
        f1.dg[i] = 1.0 / f1.dg[i];
        if (two) f2.dg[i] = 1.0 / f2.dg[i];
    }
    env_sync<TPE>(mask);
}

// L D L^T x = w in place (w: one value per row, zero on identity rows), level-scheduled: a row of the forward sweep
// only gathers rows of lower levels of the elimination tree, a row of the backward sweep only rows of higher levels,
// so the rows of one level run in parallel and every sum has a fixed order.
template <int TPE> __device__ __forceinline__ void sp_solve(const SpView& sp, const SpFactor& f, double* w, int tid, unsigned mask) {
    for (int lv = 1; lv < sp.n_lev; lv++) {   // rows of level 0 depend on nothing
        const int q1 = sp.lev_rows_ptr[lv + 1];
        for (int q = sp.lev_rows_ptr[lv] + tid; q < q1; q += TPE) {
            const int i = sp.lev_rows[q];
            double acc = w[i];
            const int t1 = sp.rowptr[i + 1];
            for (int t = sp.rowptr[i]; t < t1; t++) {
                const int en = sp.rowent[t];
                acc = fma(-f.Lv[en], w[sp.ecol[en]], acc);
            }
            w[i] = acc;
        }
        env_sync<TPE>(mask);
    }
    for (int lv = sp.n_lev - 1; lv >= 0; lv--) {
        const int q1 = sp.lev_rows_ptr[lv + 1];
        for (int q = sp.lev_rows_ptr[lv] + tid; q < q1; q += TPE) {
            const int k = sp.lev_rows[q];
            double acc = w[k] * f.dg[k];
            const int e1 = sp.colptr[k + 1];
            for (int en = sp.colptr[k]; en < e1; en++) acc = fma(-f.Lv[en], w[sp.rowidx[en]], acc);
            w[k] = acc;
        }
        env_sync<TPE>(mask);
    }
}

// ---- the same factorisation and solve for a CTA-per-env kernel whose tables and factors all sit in SHARED memory:
// 32-bit shared-window addresses and LDS/STS only, every thread owns one row of the solve (its level and the bounds of
// its row / column lists stay in registers), and only the warps that own rows meet at the per-level barrier.
__device__ __forceinline__ int lds32(unsigned a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds16(unsigned a) {
    short v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=h"(v) : "r"(a));
    return (int)v;
}

__device__ __forceinline__ unsigned lds16u(unsigned a) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return (unsigned)v;
}

struct SpsFactor { unsigned T, Lv, dg; };   // shared-window addresses of one factor's arrays

template <int TPE> __device__ __forceinline__ void sps_factor2(const PpnDevSparse& sp, unsigned tb, const SpsFactor& f1, const SpsFactor& f2,
                                                               int tid) {
    constexpr int H = TPE / 2;
    const SpsFactor f = tid < H ? f1 : f2;
    const int ht = tid < H ? tid : tid - H;
    const int nnz = sp.nnz;
    const unsigned lev_ptr = tb + 4u * sp.o_lev_ptr, lev_ent = tb + 4u * sp.o_lev_ent, trip_ptr = tb + 4u * sp.o_trip_ptr,
                   trip = tb + 4u * sp.o_trip, ecol = tb + 4u * sp.o_ecol;
    int a = lds32(lev_ptr);
    for (int lv = 0; lv < sp.n_lev; lv++) {
        const int b = lds32(lev_ptr + 4u * (lv + 1));
#pragma unroll 1
        for (int q = a + ht; q < b; q += H) {
            const int id = lds32(lev_ent + 4u * q);
            const unsigned dst = id < nnz ? f.T + 8u * id : f.dg + 8u * (id - nnz);
            double acc = lds64(id < nnz ? f.Lv + 8u * id : dst);
            const int t1 = lds32(trip_ptr + 4u * (id + 1));
#pragma unroll 1
            for (int t = lds32(trip_ptr + 4u * id); t < t1; t++) {
                const unsigned pk = (unsigned)lds32(trip + 4u * t);
                acc = fma(-lds64(f.T + 8u * (pk >> 16)), lds64(f.Lv + 8u * (pk & 0xffffu)), acc);
            }
            sts64(dst, acc);
        }
        __syncthreads();
#pragma unroll 1
        for (int q = a + ht; q < b; q += H) {
            const int id = lds32(lev_ent + 4u * q);
            if (id < nnz) sts64(f.Lv + 8u * id, lds64(f.T + 8u * id) / lds64(f.dg + 8u * lds16(ecol + 2u * id)));
        }
        __syncthreads();
        a = b;
    }
}

// ---- hybrid: sparse LDL^T below the cut, explicit dense inverse of the Schur complement of the top block ---------
// The elimination tree of a power grid is wide at the bottom and ends in a long, narrow chain; level-scheduled
// substitution crawls through that chain one dependent row at a time.  Cutting the tree where at most ~40 rows remain
// leaves a few wide sparse levels plus one small dense block whose inverse Z is formed once per factorisation, so a
// solve is (cut_lev - 1) sparse forward steps, one gather, one dense product and cut_lev sparse backward steps, each
// spread over the CTA.  Z: nt x ldz per matrix.
template <int TPE> __device__ __noinline__ void hyb_factor2(const PpnDevSparse& sp, unsigned tb, const SpsFactor& f1, const SpsFactor& f2,
                                                               double* Z1, double* Z2, int ldz, int tid) {
    constexpr int H = TPE / 2;
    const bool second = tid >= H;
    const SpsFactor f = second ? f2 : f1;
    const int ht = second ? tid - H : tid;
    const int nnz = sp.nnz, r0 = sp.cut_row, nt = sp.nt;
    const unsigned lev_ptr = tb + 4u * sp.o_lev_ptr, lev_ent = tb + 4u * sp.o_lev_ent, trip_ptr = tb + 4u * sp.o_trip_ptr,
                   trip = tb + 4u * sp.o_trip, ecol = tb + 4u * sp.o_ecol, rowidx = tb + 4u * sp.o_rowidx;
    double* Z = second ? Z2 : Z1;
    for (int i = ht; i < nt * ldz; i += H) Z[i] = 0.0;
    __syncthreads();
    int a = lds32(lev_ptr);
    for (int lv = 0; lv < sp.cut_lev; lv++) {   // the sparse levels, as sps_factor2
        const int b = lds32(lev_ptr + 4u * (lv + 1));
#pragma unroll 1
        for (int q = a + ht; q < b; q += H) {
            const int id = lds32(lev_ent + 4u * q);
            const unsigned dst = id < nnz ? f.T + 8u * id : f.dg + 8u * (id - nnz);
            double acc = lds64(id < nnz ? f.Lv + 8u * id : dst);
            const int t1 = lds32(trip_ptr + 4u * (id + 1));
#pragma unroll 1
            for (int t = lds32(trip_ptr + 4u * id); t < t1; t++) {
                const unsigned pk = (unsigned)lds32(trip + 4u * t);
                acc = fma(-lds64(f.T + 8u * (pk >> 16)), lds64(f.Lv + 8u * (pk & 0xffffu)), acc);
            }
            sts64(dst, acc);
        }
        __syncthreads();
#pragma unroll 1
        for (int q = a + ht; q < b; q += H) {
            const int id = lds32(lev_ent + 4u * q);
            if (id < nnz) sts64(f.Lv + 8u * id, lds64(f.T + 8u * id) / lds64(f.dg + 8u * lds16(ecol + 2u * id)));
        }
        __syncthreads();
        a = b;
    }
    // Schur complement of the top block: every structural entry of the block minus its update terms from the columns
    // below the cut (the terms of a target are sorted by column, entries are numbered column-major)
    const int q1 = lds32(lev_ptr + 4u * sp.n_lev);
    const unsigned cut_ent = (unsigned)sp.cut_ent;
#pragma unroll 1
    for (int q = a + ht; q < q1; q += H) {
        const int id = lds32(lev_ent + 4u * q);
        int i, j;
        double acc;
        if (id < nnz) { i = lds16(rowidx + 2u * id); j = lds16(ecol + 2u * id); acc = lds64(f.Lv + 8u * id); }
        else { i = j = id - nnz; acc = lds64(f.dg + 8u * i); }
        const int t1 = lds32(trip_ptr + 4u * (id + 1));
#pragma unroll 1
        for (int t = lds32(trip_ptr + 4u * id); t < t1; t++) {
            const unsigned pk = (unsigned)lds32(trip + 4u * t);
            if ((pk & 0xffffu) >= cut_ent) break;
            acc = fma(-lds64(f.T + 8u * (pk >> 16)), lds64(f.Lv + 8u * (pk & 0xffffu)), acc);
        }
        Z[(i - r0) * ldz + (j - r0)] = acc;
        Z[(j - r0) * ldz + (i - r0)] = acc;
    }
    // reciprocal pivots of the sparse part
    for (int k = ht; k < r0; k += H) sts64(f.dg + 8u * k, 1.0 / lds64(f.dg + 8u * k));
    __syncthreads();
}

// Gauss-Jordan inverses of the two top blocks side by side, 64 threads each (warps 0-1: Z1, warps 2-3: Z2; the other
// warps only meet the barriers).  The 64 threads form an 8 x 8 grid and thread (ti, tj) keeps the elements
// (ti + 8a, tj + 8b), a, b < 5, in REGISTERS for the whole elimination (nt <= 40).  Per pivot the owners of column k
// and of row k publish them, one barrier, every thread updates its registers from the two published vectors, one
// barrier.  `buf`: 4 * nt doubles of scratch.  No pivoting (the Schur complement of a positive definite matrix is
// positive definite); a singular block yields inf/NaN -> "diverging".
template <int TPE> __device__ __noinline__ void hyb_invert2(double* Z1, double* Z2, int nt, int ldz, double* buf, int tid) {
    if (nt <= 24) {
        // small blocks: one WARP per matrix with a row per lane in registers (pivot row through a 24-double shared buffer,
        // __syncwarp only) -- no CTA barrier per pivot; measured on IEEE-118 (17 rows): 29 k cycles with the tiled version
        if (tid < 64) gj24_rows_in_registers(saddr(tid < 32 ? Z1 : Z2), nt, ldz, tid & 31, saddr(buf) + (tid < 32 ? 0u : 8u * 26u));   // 25 doubles each after the 16-byte alignment
        __syncthreads();
        return;
    }
    constexpr int T = 5;
    const bool work = tid < 128, second = tid >= 64;
    const int ht = tid & 63, ti = ht >> 3, tj = ht & 7;
    const unsigned z = saddr(second ? Z2 : Z1);
    const unsigned col = saddr(buf) + (second ? 16u * (unsigned)nt : 0u), row = col + 8u * (unsigned)nt;
    double v[T][T];
    if (work) {
#pragma unroll
        for (int a = 0; a < T; a++)
#pragma unroll
            for (int b = 0; b < T; b++) {
                const int i = ti + 8 * a, j = tj + 8 * b;
                v[a][b] = (i < nt && j < nt) ? lds64(z + 8u * (unsigned)(i * ldz + j)) : 0.0;
            }
    }
#pragma unroll 1
    for (int k = 0; k < nt; k++) {
        if (work) {
            const int ka = k >> 3, kr = k & 7;
            if (tj == kr) {   // my column kr + 8 ka is column k
#pragma unroll
                for (int a = 0; a < T; a++) {
                    double x = v[a][0];
#pragma unroll
                    for (int b = 1; b < T; b++) x = ka == b ? v[a][b] : x;
                    if (ti + 8 * a < nt) sts64(col + 8u * (unsigned)(ti + 8 * a), x);
                }
            }
            if (ti == kr) {   // my row kr + 8 ka is row k
#pragma unroll
                for (int b = 0; b < T; b++) {
                    double x = v[0][b];
#pragma unroll
                    for (int a = 1; a < T; a++) x = ka == a ? v[a][b] : x;
                    if (tj + 8 * b < nt) sts64(row + 8u * (unsigned)(tj + 8 * b), x);
                }
            }
        }
        __syncthreads();
        if (work) {
            const double p = 1.0 / lds64(col + 8u * (unsigned)k);
            double rw[T];
#pragma unroll
            for (int b = 0; b < T; b++) rw[b] = tj + 8 * b < nt ? lds64(row + 8u * (unsigned)(tj + 8 * b)) : 0.0;
#pragma unroll
            for (int a = 0; a < T; a++) {
                const int i = ti + 8 * a;
                if (i >= nt) continue;
                if (i == k) {
#pragma unroll
                    for (int b = 0; b < T; b++) v[a][b] = (tj + 8 * b == k) ? p : v[a][b] * p;
                } else {
                    const double ci = lds64(col + 8u * (unsigned)i) * p;
#pragma unroll
                    for (int b = 0; b < T; b++) v[a][b] = (tj + 8 * b == k) ? -ci : fma(-ci, rw[b], v[a][b]);
                }
            }
        }
        __syncthreads();
    }
    if (work) {
#pragma unroll
        for (int a = 0; a < T; a++)
#pragma unroll
            for (int b = 0; b < T; b++) {
                const int i = ti + 8 * a, j = tj + 8 * b;
                if (i < nt && j < nt) sts64(z + 8u * (unsigned)(i * ldz + j), v[a][b]);
            }
    }
    __syncthreads();
}

// (B)^-1 w in place with the hybrid factor.  dg: reciprocal pivots below the cut.
// The sparse levels below the cut are tiny (IEEE-118: 25, 11, 6, 6 and 5 rows with at most 8 entries each going up, 48
// rows of level 0 coming down): ONE warp walks them, a row per lane, with nothing but __syncwarp between levels, while
// the other warps wait at a single barrier -- measured against the former all-threads version (four lanes per row, a
// CTA barrier per level: 14 barriers and ~9 k cycles per solve).  The dense top block (gather, 17 x 17 product) stays
// spread over the CTA, four lanes per row combined with two shuffles: fixed order, deterministic.
template <int TPE> __device__ __noinline__ void hyb_solve(const PpnDevSparse& sp, unsigned tb, unsigned a_Lv, unsigned a_dg, const double* Z,
                                                             int ldz, unsigned a_w, int tid) {
    const unsigned a_lev = tb + 4u * sp.o_lev_rows_ptr, a_rowpk = tb + 4u * sp.o_rowpk, a_colpk = tb + 4u * sp.o_colpk,
                   a_rpack = tb + 4u * sp.o_rpack, a_rowoff = tb + 4u * sp.o_rowoff, a_z = saddr(Z);
    const int r0 = sp.cut_row, nt = sp.nt;
    const int t = tid >> 2, part = tid & 3;
    const int wrow = (tid >> 5) << 3;   // first row of a pass that falls to this warp (eight rows per warp)
    if (tid < 32) {   // sparse forward steps (rows of level 0 depend on nothing)
        int s0 = lds32(a_lev + 4u);
        for (int lv = 1; lv < sp.cut_lev; lv++) {
            const int s1 = lds32(a_lev + 4u * (lv + 1));
#pragma unroll 1
            for (int i = s0 + tid; i < s1; i += 32) {
                const unsigned pk = (unsigned)lds32(a_rowpk + 4u * i);
                const unsigned wi = a_w + 8u * i;
                double acc = lds64(wi);
                unsigned ra = a_rpack + 4u * (pk >> 8);
                int cnt = (int)(pk & 255u);
#pragma unroll 1
                for (; cnt >= 2; cnt -= 2, ra += 8u) {   // two entries per trip, loads first
                    const unsigned p0 = (unsigned)lds32(ra), p1 = (unsigned)lds32(ra + 4u);
                    const double l0 = lds64(a_Lv + (p0 >> 16)), x0 = lds64(a_w + (p0 & 0xffffu));
                    const double l1 = lds64(a_Lv + (p1 >> 16)), x1 = lds64(a_w + (p1 & 0xffffu));
                    acc = fma(-l0, x0, acc);
                    acc = fma(-l1, x1, acc);
                }
                if (cnt) {
                    const unsigned p0 = (unsigned)lds32(ra);
                    acc = fma(-lds64(a_Lv + (p0 >> 16)), lds64(a_w + (p0 & 0xffffu)), acc);
                }
                sts64(wi, acc);
            }
            __syncwarp();
            s0 = s1;
        }
    }
    __syncthreads();
    // top block: y2 = w2 - L21 y1 (entries below the cut only)
    if (wrow < nt) {
        double y = 0.0;
        if (t < nt) {
            const unsigned pk = (unsigned)lds32(a_rowpk + 4u * (r0 + t));
            const int cnt = (int)(pk & 255u);
            const unsigned ra = a_rpack + 4u * (pk >> 8), cutoff = 8u * (unsigned)r0;
#pragma unroll 1
            for (int q = part; q < cnt; q += 4) {
                const unsigned p0 = (unsigned)lds32(ra + 4u * q);
                if ((p0 & 0xffffu) < cutoff) y = fma(-lds64(a_Lv + (p0 >> 16)), lds64(a_w + (p0 & 0xffffu)), y);
            }
        }
        y += __shfl_xor_sync(PPN_FULL, y, 1);
        y += __shfl_xor_sync(PPN_FULL, y, 2);
        if (t < nt && part == 0) sts64(a_w + 8u * (r0 + t), lds64(a_w + 8u * (r0 + t)) + y);
    }
    __syncthreads();
    // x2 = Z y2
    {
        double x = 0.0, x2 = 0.0;
        if (wrow < nt && t < nt) {
            const unsigned zr = a_z + 8u * (unsigned)(t * ldz), wr = a_w + 8u * r0;
            int j = part;
#pragma unroll 1
            for (; j + 4 < nt; j += 8) {
                const double z0 = lds64(zr + 8u * j), z1 = lds64(zr + 8u * j + 32u);
                const double w0 = lds64(wr + 8u * j), w1 = lds64(wr + 8u * j + 32u);
                x = fma(z0, w0, x);
                x2 = fma(z1, w1, x2);
            }
            if (j < nt) x = fma(lds64(zr + 8u * j), lds64(wr + 8u * j), x);
            x += x2;
        }
        if (wrow < nt) {
            x += __shfl_xor_sync(PPN_FULL, x, 1);
            x += __shfl_xor_sync(PPN_FULL, x, 2);
        }
        __syncthreads();
        if (t < nt && part == 0) sts64(a_w + 8u * (r0 + t), x);
    }
    __syncthreads();
    if (tid < 32) {   // sparse backward steps, down to level 0
        int e1 = r0;
        for (int lv = sp.cut_lev - 1; lv >= 0; lv--) {
            const int e0 = lds32(a_lev + 4u * lv);
#pragma unroll 1
            for (int i = e0 + tid; i < e1; i += 32) {
                const unsigned pk = (unsigned)lds32(a_colpk + 4u * i);
                const unsigned wi = a_w + 8u * i;
                double acc = lds64(wi) * lds64(a_dg + 8u * i);
                unsigned en = pk >> 8;
                int cnt = (int)(pk & 255u);
#pragma unroll 1
                for (; cnt >= 2; cnt -= 2, en += 2u) {
                    const int q0 = lds16(a_rowoff + 2u * en), q1 = lds16(a_rowoff + 2u * en + 2u);
                    const double l0 = lds64(a_Lv + 8u * en), l1 = lds64(a_Lv + 8u * en + 8u);
                    const double x0 = lds64(a_w + q0), x1 = lds64(a_w + q1);
                    acc = fma(-l0, x0, acc);
                    acc = fma(-l1, x1, acc);
                }
                if (cnt) acc = fma(-lds64(a_Lv + 8u * en), lds64(a_w + lds16(a_rowoff + 2u * en)), acc);
                sts64(wi, acc);
            }
            __syncwarp();
            e1 = e0;
        }
    }
    __syncthreads();
}

// The same solve as ONE-WARP PROGRAM driven by the static schedule PpnDevSparse.o_wsched: a row per lane in every step,
// the entries of a step laid out lane-major (16-bit: entry id << 7 | row or column), so that within a step no index load
// depends on another one -- header and entries are fetched while the previous step's stores drain, the only chain is
// operand load -> FMA -> store -> __syncwarp.  The dense top block (nt <= 32 rows) is one more step of the same warp:
// no CTA barrier inside the solve at all; the other warps wait at the closing barrier without using issue slots.
template <int TPE> __device__ __forceinline__ void hyb_solve_warp(const PpnDevSparse& sp, unsigned tb, unsigned a_Lv, unsigned a_dg, const double* Z,
                                                                  int ldz, unsigned a_w, int tid) {
    if (tid < 32) {
        const unsigned ws = tb + 4u * sp.o_wsched;
        const int nf = lds32(ws), nb = lds32(ws + 4u);
        const unsigned a_hdr = ws + 8u, a_ent = a_hdr + 8u * (unsigned)(nf + nb), a_z = saddr(Z);
        const int r0 = sp.cut_row, nt = sp.nt;
        int h0 = lds32(a_hdr), h1 = lds32(a_hdr + 4u);
        for (int st = 0; st < nf + nb; st++) {
            if (st == nf) {   // x2 = Z y2 between the forward and the backward steps
                double x = 0.0, x2 = 0.0;
                if (tid < nt) {
                    const unsigned zr = a_z + 8u * (unsigned)(tid * ldz), wr = a_w + 8u * r0;
                    int j = 0;
#pragma unroll 1
                    for (; j + 1 < nt; j += 2) {
                        x = fma(lds64(zr + 8u * j), lds64(wr + 8u * j), x);
                        x2 = fma(lds64(zr + 8u * j + 8u), lds64(wr + 8u * j + 8u), x2);
                    }
                    if (j < nt) x = fma(lds64(zr + 8u * j), lds64(wr + 8u * j), x);
                    x += x2;
                }
                __syncwarp();
                if (tid < nt) sts64(a_w + 8u * (r0 + tid), x);
                __syncwarp();
            }
            const int row0 = h0 & 0xffff, rows = h0 >> 16, maxcnt = h1 & 255;
            unsigned ea = a_ent + 2u * ((unsigned)(h1 >> 8) + (unsigned)tid);
            if (st + 1 < nf + nb) { h0 = lds32(a_hdr + 8u * (st + 1)); h1 = lds32(a_hdr + 8u * (st + 1) + 4u); }   // next header
            const bool mine = tid < rows;
            const unsigned wi = a_w + 8u * (unsigned)(row0 + tid);
            double acc = 0.0, acc2 = 0.0;
            if (mine) acc = st >= nf ? lds64(wi) * lds64(a_dg + 8u * (unsigned)(row0 + tid)) : lds64(wi);
            int q = 0;
#pragma unroll 1
            for (; q + 1 < maxcnt; q += 2, ea += 128u) {   // 32 halfwords per entry slot
                const unsigned p0 = lds16u(ea), p1 = lds16u(ea + 64u);
                const double l0 = p0 != 0xffffu ? lds64(a_Lv + 8u * (p0 >> 7)) : 0.0, v0 = p0 != 0xffffu ? lds64(a_w + 8u * (p0 & 127u)) : 0.0;
                const double l1 = p1 != 0xffffu ? lds64(a_Lv + 8u * (p1 >> 7)) : 0.0, v1 = p1 != 0xffffu ? lds64(a_w + 8u * (p1 & 127u)) : 0.0;
                acc = fma(-l0, v0, acc);
                acc2 = fma(-l1, v1, acc2);
            }
            if (q < maxcnt) {
                const unsigned p0 = lds16u(ea);
                if (p0 != 0xffffu) acc = fma(-lds64(a_Lv + 8u * (p0 >> 7)), lds64(a_w + 8u * (p0 & 127u)), acc);
            }
            if (mine) sts64(wi, acc + acc2);
            __syncwarp();
        }
    }
    __syncthreads();
}

// the one-warp schedule when the structure has one (16-bit entries: < 511 fill entries, <= 128 rows, top block <= 32)
template <int TPE> __device__ __forceinline__ void hyb_solve_any(const PpnDevSparse& sp, unsigned tb, unsigned a_Lv, unsigned a_dg, const double* Z,
                                                                 int ldz, unsigned a_w, int tid) {
    if (sp.o_wsched >= 0) hyb_solve_warp<TPE>(sp, tb, a_Lv, a_dg, Z, ldz, a_w, tid);
    else hyb_solve<TPE>(sp, tb, a_Lv, a_dg, Z, ldz, a_w, tid);
}

// L D L^T x = w in place; dg holds the RECIPROCAL pivots.  The rows of a level are contiguous (the host sorts the
// elimination order by level), so ONE warp walks the levels with nothing but __syncwarp between them -- no CTA
// barrier per level -- while the other warps of the CTA wait at the closing barrier without using issue slots.
// Every address is a 32-bit shared-window address hoisted out of the loops; the tables hold byte offsets.
__device__ __forceinline__ void sps_solve_warp(unsigned a_lev, unsigned a_rowpk, unsigned a_colpk, unsigned a_rpack, unsigned a_rowoff,
                                               unsigned a_Lv, unsigned a_dg, unsigned a_w, int n_lev, int lane) {
    int s0 = lds32(a_lev + 4u);
    for (int lv = 1; lv < n_lev; lv++) {   // forward: rows of level 0 depend on nothing
        const int s1 = lds32(a_lev + 4u * (lv + 1));
#pragma unroll 1
        for (int i = s0 + lane; i < s1; i += 32) {
            const unsigned pk = (unsigned)lds32(a_rowpk + 4u * i);
            const unsigned wi = a_w + 8u * i;
            double acc = lds64(wi);
            unsigned ra = a_rpack + 4u * (pk >> 8);
            int cnt = (int)(pk & 255u);
#pragma unroll 1
            for (; cnt >= 2; cnt -= 2, ra += 8u) {   // two entries per trip, loads first
                const unsigned p0 = (unsigned)lds32(ra), p1 = (unsigned)lds32(ra + 4u);
                const double l0 = lds64(a_Lv + (p0 >> 16)), x0 = lds64(a_w + (p0 & 0xffffu));
                const double l1 = lds64(a_Lv + (p1 >> 16)), x1 = lds64(a_w + (p1 & 0xffffu));
                acc = fma(-l0, x0, acc);
                acc = fma(-l1, x1, acc);
            }
            if (cnt) {
                const unsigned p0 = (unsigned)lds32(ra);
                acc = fma(-lds64(a_Lv + (p0 >> 16)), lds64(a_w + (p0 & 0xffffu)), acc);
            }
            sts64(wi, acc);
        }
        __syncwarp();
        s0 = s1;
    }
    int e1 = lds32(a_lev + 4u * n_lev);
    for (int lv = n_lev - 1; lv >= 0; lv--) {   // diagonal and backward
        const int e0 = lds32(a_lev + 4u * lv);
#pragma unroll 1
        for (int i = e0 + lane; i < e1; i += 32) {
            const unsigned pk = (unsigned)lds32(a_colpk + 4u * i);
            const unsigned wi = a_w + 8u * i;
            double acc = lds64(wi) * lds64(a_dg + 8u * i);
            unsigned en = pk >> 8;
            int cnt = (int)(pk & 255u);
#pragma unroll 1
            for (; cnt >= 2; cnt -= 2, en += 2u) {
                const int r0 = lds16(a_rowoff + 2u * en), r1 = lds16(a_rowoff + 2u * en + 2u);
                const double l0 = lds64(a_Lv + 8u * en), l1 = lds64(a_Lv + 8u * en + 8u);
                const double x0 = lds64(a_w + r0), x1 = lds64(a_w + r1);
                acc = fma(-l0, x0, acc);
                acc = fma(-l1, x1, acc);
            }
            if (cnt) acc = fma(-lds64(a_Lv + 8u * en), lds64(a_w + lds16(a_rowoff + 2u * en)), acc);
            sts64(wi, acc);
        }
        __syncwarp();
        e1 = e0;
    }
}

template <int TPE> __device__ __forceinline__ void sps_solve(const PpnDevSparse& sp, unsigned tb, unsigned a_Lv, unsigned a_dg, unsigned w,
                                                             int tid) {
    if (tid < 32)
        sps_solve_warp(tb + 4u * sp.o_lev_rows_ptr, tb + 4u * sp.o_rowpk, tb + 4u * sp.o_colpk, tb + 4u * sp.o_rpack,
                       tb + 4u * sp.o_rowoff, a_Lv, a_dg, w, sp.n_lev, tid);
    __syncthreads();
}

// Explicit inverses of the active parts of the factored matrices, one column per thread (the columns of both
// matrices are spread over the env's threads in one go): forward substitution of a unit vector only walks the
// elimination-tree path of its row, the backward sweep is the same instruction stream for every thread (uniform
// index loads).  M: (n_act + 1) x ld, row n_act is the landing row of the identity rows.
template <int TPE> __device__ __forceinline__ void sp_invert(const SpView& sp, const SpFactor& f1, const SpFactor& f2,
                                                             const short* colbus1, const short* colbus2, double* M1, double* M2,
                                                             int n1, int n2, int ld1, int ld2, int tid, unsigned mask) {
    // reciprocal pivots and the offsets of every row / entry inside a column of the inverse
    for (int i = tid; i < sp.n; i += TPE) {
        f1.dg[i] = 1.0 / f1.dg[i]; f1.koff[i] = f1.cp[i] * ld1;
        if (n2 > 0) { f2.dg[i] = 1.0 / f2.dg[i]; f2.koff[i] = f2.cp[i] * ld2; }
    }
    for (int i = tid; i < sp.nnz; i += TPE) {
        const int r = sp.rowidx[i];
        f1.eoff[i] = f1.cp[r] * ld1;
        if (n2 > 0) f2.eoff[i] = f2.cp[r] * ld2;
    }
    env_sync<TPE>(mask);
    for (int cc = tid; cc < n1 + n2; cc += TPE) {
        const bool second = cc >= n1;
        const SpFactor& f = second ? f2 : f1;
        const int c = second ? cc - n1 : cc, n_act = second ? n2 : n1, ld = second ? ld2 : ld1;
        double* x = (second ? M2 : M1) + c;
        for (int r = 0; r <= n_act; r++) x[r * ld] = 0.0;
        const int jr = sp.bus_row[(second ? colbus2 : colbus1)[c]];
        x[f.koff[jr]] = 1.0;
        for (int k = jr; k >= 0; k = sp.parent[k]) {
            double* xk = x + f.koff[k];
            const double yk = *xk;
            const int e1 = sp.colptr[k + 1];
            for (int e = sp.colptr[k]; e < e1; e++) {
                double* t = x + f.eoff[e];
                *t = fma(-f.Lv[e], yk, *t);
            }
            *xk = yk * f.dg[k];
        }
        int e = sp.nnz;
        for (int k = sp.n - 1; k >= 0; k--) {
            double* t = x + f.koff[k];
            double sacc = *t;
            const int e0 = sp.colptr[k];
            for (int q = e0; q < e; q++) sacc = fma(-f.Lv[q], x[f.eoff[q]], sacc);
            e = e0;
            *t = sacc;
        }
    }
    env_sync<TPE>(mask);
}

// dot product of a matrix row with a vector, four independent accumulation chains; the vector (16-byte aligned, in
// shared memory: the mismatch vectors) is read two entries per load
__device__ __forceinline__ double row_dot(const double* m, const double* x, int n) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const double2* x2 = reinterpret_cast<const double2*>(x);
    int j = 0;
    for (; j + 3 < n; j += 4) {
        const double2 xa = x2[j >> 1], xb = x2[(j >> 1) + 1];
        const double m0 = m[j], m1 = m[j + 1], m2 = m[j + 2], m3 = m[j + 3];
        a0 = fma(m0, xa.x, a0); a1 = fma(m1, xa.y, a1);
        a2 = fma(m2, xb.x, a2); a3 = fma(m3, xb.y, a3);
    }
    if (j < n) a0 = fma(m[j], x[j], a0);
    if (j + 1 < n) a1 = fma(m[j + 1], x[j + 1], a1);
    if (j + 2 < n) a2 = fma(m[j + 2], x[j + 2], a2);
    return (a0 + a1) + (a2 + a3);
}

// ------------------------------------------------------------------------------------------------ chronic access
__device__ __forceinline__ const float* chronic_row(const PpnDevChronics& ch, int chronic, int row) {
    return ch.rows + (size_t)(ch.row_off[chronic] + row) * ch.row_words;
}

// ChronicLooper.get_next_chronic_folder (chronic.py:282-291): natural / random / fixed.
__device__ __forceinline__ int take_next_chronic(int* cursor, const PpnDevCfg& cfg, int n_chronics, int env) {
    const int cur = cursor[2];
    if (cfg.loop_mode == 0) {
        cursor[2] = (cur + 1) % n_chronics;
    } else if (cfg.loop_mode == 1) {  // counter-based hash (splitmix64) of (seed, env, draw index)
        unsigned long long z = cfg.seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(env + 1) +
                               0xBF58476D1CE4E5B9ull * (unsigned long long)(unsigned)cursor[3];
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        cursor[2] = (int)(z % (unsigned long long)n_chronics);
        cursor[3] += 1;
    }
    return cur;
}

// ------------------------------------------------------------------------------------------ topology-derived maps
// Current bus of every element (sub + S*node) and normalize_prods_voltages' positional baseKV (grid.py:266-271).
template <int TPE, class D>
__device__ __forceinline__ void refresh_element_buses(Env<TPE, D>& e, const PpnDevCase& c) {
    const int S = e.S;
    for (int l = e.tid; l < e.N; l += TPE) {
        e.fbus()[l] = (short)(c.lor_sub[l] + S * e.onode()[l]);
        e.tbus()[l] = (short)(c.lex_sub[l] + S * e.enode()[l]);
    }
    for (int g = e.tid; g < e.G; g += TPE) e.gbus()[g] = (short)(c.gen_sub[g] + S * e.gnode()[g]);
    for (int l = e.tid; l < e.L; l += TPE) e.lbus()[l] = (short)(c.load_sub[l] + S * e.lnode()[l]);
    env_sync<TPE>(e.mask);
    // baseKV of gen-hosting buses in bus-array order, applied positionally to the generators
    for (int g = e.tid; g < e.G; g += TPE) {
        const int b = e.gbus()[g];
        int rank = 0;
        for (int h = 0; h < e.G; h++) rank += (e.gbus()[h] < b);
        e.gkv()[rank] = c.bus_basekv[b];
    }
    env_sync<TPE>(e.mask);
}

// game.py:476-501 then :405-474 and grid.py:273-311.
template <int TPE, class D>
__device__ __forceinline__ void load_next_timestep(Env<TPE, D>& e, const PpnDevCase& c, const PpnDevChronics& ch, const PpnDevCfg& cfg,
                                   bool is_sim, int env) {
    int chronic = e.cursor()[0], row = e.cursor()[1];
    const int e_chronic = chronic, e_row = row;  // `current_timestep_entries` before the move (simulate reads its planned values)
    env_sync<TPE>(e.mask);
    int n_rows = ch.n_rows[chronic];
    bool at_last = (row >= 0 && row == n_rows - 1) || (row == -2 && ch.last_id_zero[chronic]);
    if (at_last && !is_sim) {
        if (e.tid == 0) {
            e.cursor()[0] = take_next_chronic(e.cursor(), cfg, ch.n_chronics, env);
            e.cursor()[1] = -2;
        }
        env_sync<TPE>(e.mask);
        chronic = e.cursor()[0];
        row = -2;
        n_rows = ch.n_rows[chronic];
        env_sync<TPE>(e.mask);
    }
    int nrow;
    if (row == -1) nrow = 0;
    else if (row == -2) nrow = max(ch.row_after_switch[chronic], 0);
    else nrow = min(row + 1, n_rows - 1);
    if (!is_sim) {
        for (int i = e.tid; i < e.N; i += TPE) {
            if (e.recon()[i] > 0) e.recon()[i] -= 1;
            if (e.lreact()[i] > 0) e.lreact()[i] -= 1;
        }
        for (int i = e.tid; i < e.S; i += TPE)
            if (e.nreact()[i] > 0) e.nreact()[i] -= 1;
    }
    const float* next = chronic_row(ch, chronic, nrow);
    const float* src;
    int opp, opv, olp, olq;
    if (!is_sim) {
        src = next; opp = ch.o_pp; opv = ch.o_pv; olp = ch.o_lp; olq = ch.o_lq;
    } else {
        src = chronic_row(ch, e_chronic, max(e_row, 0));
        opp = ch.o_ppp; opv = ch.o_pvp; olp = ch.o_lpp; olq = ch.o_lqp;
    }
    for (int g = e.tid; g < e.G; g += TPE) {
        const float pv = src[opv + g];
        e.gpg()[g] = (double)src[opp + g];
        e.gvg()[g] = (double)(pv <= 0.f ? 0.f : pv) / e.gkv()[g];
        e.gstat()[g] = pv > 0.f ? 1 : 0;
    }
    for (int l = e.tid; l < e.L; l += TPE) {
        e.lpd()[l] = (double)src[olp + l];
        e.lqd()[l] = (double)src[olq + l];
    }
    for (int i = e.tid; i < e.N; i += TPE) {   // same lane decremented recon[i] above
        const int m = (int)next[ch.o_mt + i];
        if (m > 0) { e.status()[i] = 0; e.recon()[i] = max(e.recon()[i], m); }
        if (!is_sim) {
            const int h = (int)next[ch.o_hz + i];
            if (h > 0) { e.status()[i] = 0; e.recon()[i] = max(e.recon()[i], h); }
        }
    }
    if (e.tid == 0) {
        e.cursor()[0] = chronic; e.cursor()[1] = nrow;
        // row whose planned values the observation reports (`current_timestep_entries`: not moved by simulate)
        e.misc()[6] = is_sim ? e_chronic : chronic;
        e.misc()[7] = is_sim ? max(e_row, 0) : nrow;
    }
    env_sync<TPE>(e.mask);
}

// ----------------------------------------------------------------------------------------------------- load-flow
// isolated buses (grid.py:176-210): mark[b] = 1 when no in-service line ends at b
template <int TPE, class D>
__device__ __forceinline__ void compute_isolated(Env<TPE, D>& e) {
    for (int b = e.tid; b < e.NB; b += TPE) e.mark()[b] = 1;
    env_sync<TPE>(e.mask);
    for (int l = e.tid; l < e.N; l += TPE)
        if (e.status()[l]) { e.mark()[e.fbus()[l]] = 0; e.mark()[e.tbus()[l]] = 0; }
    env_sync<TPE>(e.mask);
}

// Per-bus lists of the in-service line ends that sit on the bus, rebuilt for every load-flow.  The entries of the two
// buses of substation s share the substation's slice [adj_ptr[s], adj_ptr[s+1]) of the static adjacency: node 0 fills
// it from the front, node 1 from the back.  Entry k: eoth = bus at the other end, eline = line*2+end, (eyr, eyi) =
// the off-diagonal admittance seen from this end (yft from the origin, ytf from the extremity).
template <int TPE, class D>
__device__ __forceinline__ void build_entries(Env<TPE, D>& e, const PpnDevCase& c) {
    const int S = e.S;
    for (int b = e.tid; b < e.NB; b += TPE) {
        const int s = b >= S ? b - S : b, node = b >= S ? 1 : 0;
        const int k0 = c.adj_ptr[s], k1 = c.adj_ptr[s + 1];
        int n = 0;
        for (int k = k0; k < k1; k++) {
            const int a = c.adj[k], l = a >> 1, end = a & 1;
            if (!e.status()[l] || (end == 0 ? e.onode()[l] : e.enode()[l]) != node) continue;
            const int slot = node ? k1 - 1 - n : k0 + n;
            const double* y = c.line_y + 8 * l + (end ? 4 : 2);
            e.eoth()[slot] = end == 0 ? e.tbus()[l] : e.fbus()[l];
            e.eline()[slot] = (short)a;
            e.ey()[2 * slot] = y[0]; e.ey()[2 * slot + 1] = y[1];
            n++;
        }
        e.deg()[b] = (uint8_t)n;
    }
    env_sync<TPE>(e.mask);
}

// first entry slot and direction of bus b's list
#define PPN_ENTRIES(e, c, b, k0, step)                                  \
    const int _s = (b) >= (e).S ? (b) - (e).S : (b);                    \
    const int step = (b) >= (e).S ? -1 : 1;                             \
    const int k0 = (b) >= (e).S ? (c).adj_ptr[_s + 1] - 1 : (c).adj_ptr[_s];

// S_b = V_b conj((Ybus V)_b), gathered per bus: diagonal term (ydr, ydi: shunt + own-end admittances) + the
// off-diagonal entries of the in-service lines that end on this bus.
template <int TPE, class D>
__device__ __forceinline__ void bus_power(const Env<TPE, D>& e, const PpnDevCase& c, int b, double& sr, double& si) {
    PPN_ENTRIES(e, c, b, k0, step)
    const int n = e.deg()[b];
    const double vr = e.vri()[2 * b], vi = e.vri()[2 * b + 1];
    double ir = e.ydr()[b] * vr - e.ydi()[b] * vi, ii = e.ydr()[b] * vi + e.ydi()[b] * vr;
    double jr = 0.0, ji = 0.0;   // second accumulator pair: two independent chains
#pragma unroll 2
    for (int q = 0; q < n; q++) {
        const int k = k0 + step * q;
        const int o = e.eoth()[k];
        const double yr = e.ey()[2 * k], yi = e.ey()[2 * k + 1];
        const double wr = e.vri()[2 * o], wi = e.vri()[2 * o + 1];
        if (q & 1) { jr = fma(yr, wr, fma(-yi, wi, jr)); ji = fma(yr, wi, fma(yi, wr, ji)); }
        else { ir = fma(yr, wr, fma(-yi, wi, ir)); ii = fma(yr, wi, fma(yi, wr, ii)); }
    }
    ir += jr; ii += ji;
    sr = vr * ir + vi * ii;   // V conj(I)
    si = vi * ir - vr * ii;
}

// demand at bus b: the load of the substation when it sits on this node (its sister bus carries none)
template <int TPE, class D>
__device__ __forceinline__ void bus_demand(const Env<TPE, D>& e, const PpnDevCase& c, int b, double& pd, double& qd) {
    const int s = b >= e.S ? b - e.S : b, node = b >= e.S ? 1 : 0;
    const int l = c.load_of_sub[s];
    const bool here = l >= 0 && e.lnode()[l] == node;
    pd = here ? e.lpd()[l] : 0.0;
    qd = here ? e.lqd()[l] : 0.0;
}

// fdpf's mismatch: mis = (V conj(Ybus V) - Sbus)/Vm; P over pv+pq, Q over pq.  Returns whether both infinity norms
// are below tol (what fdpf tests after every half iteration); a NaN anywhere reads as "not converged".
template <int TPE, class D>
__device__ __forceinline__ bool mismatch(Env<TPE, D>& e, const PpnDevCase& c, double tol) {
    bool open = false;
    for (int b = e.tid; b < e.NB; b += TPE) {
        const int t = e.btype()[b];
        if (t == PPN_BT_ISOLATED || t == PPN_BT_REF) continue;
        double sr, si;
        bus_power(e, c, b, sr, si);
        const double rvm = 1.0 / e.vm()[b];
        const double p = (sr - e.pin()[b]) * rvm;
        e.P()[e.idxp()[b]] = p;
        open |= !(fabs(p) < tol);
        if (t == PPN_BT_PQ) {
            const double q = (si - e.qin()[b]) * rvm;
            e.Q()[e.idxq()[b]] = q;
            open |= !(fabs(q) < tol);
        }
    }
    const bool any_open = env_any<TPE>(open, e.mask);
    env_sync<TPE>(e.mask);
    return !any_open;
}

// ---- split buses under the hybrid factor: the bordered system -----------------------------------------------------
// The hybrid factor is built on the U structure (one row per substation: bus s).  When node-splitting actions put
// elements on sister buses, those few buses (k <= PPN_BORDER_MAX per matrix; one per split substation as a rule) are NOT
// given rows in the sparse structure -- that would be the two-rows-per-substation F structure, four times the fill and
// five times the levels -- but border it:   [A B; B^T C] [x; y] = [f; g]   with A the U-structure matrix (identity rows
// for buses that take no part), C the k x k block of the sister buses and B their couplings to A's rows.  Once per
// load-flow: W = A^-1 B (k solves with the hybrid factor) and the inverse of the Schur complement C - B^T W; per solve:
// t = A^-1 f, y = (C - B^T W)^-1 (g - B^T t), x = t - W y.  B is never stored: its entries are the line ends of the
// sister buses.  Storage (per env, in its slice of the HBM workspace): W1 | W2 (k x n each), S1 | S2 (k x k).
#define PPN_BORDER_MAX 8

struct Border {
    double* W[2];     // [PPN_BORDER_MAX][n]  A^-1 B of B' / B''
    double* Sc[2];    // [PPN_BORDER_MAX][PPN_BORDER_MAX]  C, then the inverse of the Schur complement
    int k[2];         // sister buses in the system of B' (pv + pq) / B'' (pq)
    int base[2];      // compact index of the first sister bus: n1 - k1 / n2 - k2
};

// Gauss-Jordan inverse of a k x k matrix (k <= PPN_BORDER_MAX, row stride PPN_BORDER_MAX) by one warp: lane i keeps row i
// in registers, the pivot row travels by shuffles.  No pivoting: Schur complements of positive definite matrices.
__device__ __forceinline__ void border_invert_warp(double* Sm, int k, int lane) {
    double row[PPN_BORDER_MAX];
#pragma unroll
    for (int j = 0; j < PPN_BORDER_MAX; j++) row[j] = (lane < k && j < k) ? Sm[lane * PPN_BORDER_MAX + j] : 0.0;
#pragma unroll
    for (int p = 0; p < PPN_BORDER_MAX; p++) {
        if (p < k) {   // uniform
            double pr[PPN_BORDER_MAX];
#pragma unroll
            for (int j = 0; j < PPN_BORDER_MAX; j++) pr[j] = __shfl_sync(PPN_FULL, row[j], p);
            const double piv = 1.0 / pr[p];
            if (lane == p) {
#pragma unroll
                for (int j = 0; j < PPN_BORDER_MAX; j++) row[j] = (j == p) ? piv : row[j] * piv;
            } else {
                const double ci = row[p] * piv;
#pragma unroll
                for (int j = 0; j < PPN_BORDER_MAX; j++) row[j] = (j == p) ? -ci : fma(-ci, pr[j], row[j]);
            }
        }
    }
    if (lane < k) {
#pragma unroll
        for (int j = 0; j < PPN_BORDER_MAX; j++)
            if (j < k) Sm[lane * PPN_BORDER_MAX + j] = row[j];
    }
}

// coupling of a sister bus's line-end entry k (other end o) in matrix m: B' -> -1/x, B'' -> -Im(off-diagonal admittance)
template <int TPE, class D>
__device__ __forceinline__ double border_coef(const Env<TPE, D>& e, const PpnDevCase& c, int k, int m) {
    return m == 0 ? -c.line_bp[e.eline()[k] >> 1] : -e.ey()[2 * k + 1];
}

// rundcpf on the prepared env: B theta = Pbus on pv+pq, Vm := 1.  M1 = n1 x n1 work matrix.
template <int TPE, int MAXR, class D>
__device__ __forceinline__ bool dc_solve(Env<TPE, D>& e, const PpnDevCase& c, double* M1, int n1, int ld1, int ref,
                                         const SpView* sp, double* spb, bool solve_mode) {
    const int NB = e.NB, S = e.S, tid = e.tid;
    const unsigned mask = e.mask;
    bool success;
    // ================= rundcpf: B theta = Pbus on pv+pq, Vm := 1
    SpFactor f1{}, f2{};
    if (sp) { f1 = sp_carve(spb, *sp, 0); sp_clear<TPE>(*sp, f1, n1, tid); }
    else for (int i = tid; i < n1 * ld1; i += TPE) M1[i] = 0.0;
    // the hybrid routines always work on two matrices side by side: the second one is an identity here
    if (sp && solve_mode && TPE > 32 && sp->hyb) { f2 = sp_carve(spb, *sp, 1); sp_clear<TPE>(*sp, f2, n1, tid); }
    if (solve_mode) for (int i = tid; i < NB; i += TPE) e.ydr()[i] = 0.0;   // right-hand side by row (ydr is free in DC mode)
    env_sync<TPE>(mask);
    const double va_ref = e.va()[ref] * (PPN_PI / 180.0);
    for (int i = tid; i < n1; i += TPE) {
        const int b = e.busp()[i];
        PPN_ENTRIES(e, c, b, k0, step)
        const int nent = e.deg()[b];
        const int row = sp ? sp->bus_row[b] : 0;
        double diag = 0.0, bref = 0.0;
        for (int q = 0; q < nent; q++) {
            const int k = k0 + step * q;
            const int o = e.eoth()[k];
            const int l = e.eline()[k] >> 1;
            const double w = c.line_bdc[l];
            diag += w;
            if (o == ref) bref -= w;
            else if (!sp) M1[i * ld1 + e.idxp()[o]] -= w;
            else if (row > sp->bus_row[o]) f1.Lv[sp->line_pos[sp->full ? 4 * l + 2 * e.onode()[l] + e.enode()[l] : l]] -= w;
        }
        if (sp) { f1.dg[row] = diag; f1.cp[row] = (short)i; }
        else M1[i * ld1 + i] += diag;
        // rhs = Pbus[pvpq] - B[pvpq, ref] Va0[ref], Pbus = Re(Sbus) - Gs/baseMVA
        const double rhs = (e.pin()[b] - c.bus_ysh_r[b]) - bref * va_ref;
        if (solve_mode) e.ydr()[row] = rhs; else e.P()[i] = rhs;
    }
    env_sync<TPE>(mask);
    if (solve_mode && TPE > 32 && sp->hyb) {
        const int ldz = sp->d->nt | 1;
        double* Z1 = spb + 2 * ppn_sp_factor_doubles(sp->n, sp->nnz);
        const SpsFactor s1{saddr(f1.T), saddr(f1.Lv), saddr(f1.dg)}, s2{saddr(f2.T), saddr(f2.Lv), saddr(f2.dg)};
        hyb_factor2<TPE>(*sp->d, sp->tb, s1, s2, Z1, Z1 + sp->d->nt * ldz, ldz, tid);
        hyb_invert2<TPE>(Z1, Z1 + sp->d->nt * ldz, sp->d->nt, ldz, e.vri(), tid);   // vri: 2 NB doubles of scratch in DC mode
        hyb_solve_any<TPE>(*sp->d, sp->tb, s1.Lv, s1.dg, Z1, ldz, saddr(e.ydr()), tid);
        for (int i = tid; i < n1; i += TPE) e.Q()[i] = e.ydr()[sp->bus_row[e.busp()[i]]];
    } else if (solve_mode) {
        sp_factor<TPE, 1>(*sp, f1, f1, tid, mask);
        sp_recip<TPE>(*sp, f1, f1, false, tid, mask);
        if (TPE > 32 && sp->tb) sps_solve<TPE>(*sp->d, sp->tb, saddr(f1.Lv), saddr(f1.dg), saddr(e.ydr()), tid);
        else sp_solve<TPE>(*sp, f1, e.ydr(), tid, mask);
        for (int i = tid; i < n1; i += TPE) e.Q()[i] = e.ydr()[sp->bus_row[e.busp()[i]]];
    } else {
        if (sp) {
            sp_factor<TPE, 1>(*sp, f1, f1, tid, mask);
            sp_invert<TPE>(*sp, f1, f1, e.busp(), e.busp(), M1, M1, n1, 0, ld1, ld1, tid, mask);
        }
        else gj_invert<TPE, MAXR>(M1, n1, ld1, tid, mask);
        for (int i = tid; i < n1; i += TPE) {
            e.Q()[i] = row_dot(M1 + i * ld1, e.P(), n1);  // theta (radians) of pvpq bus i
        }
    }
    env_sync<TPE>(mask);
    for (int b = tid; b < NB; b += TPE) {
        const int t = e.btype()[b];
        if (t == PPN_BT_ISOLATED) continue;
        e.theta()[b] = (t == PPN_BT_REF) ? va_ref : e.Q()[e.idxp()[b]];
    }
    env_sync<TPE>(mask);
    // branch flows, slack production
    for (int l = tid; l < e.N; l += TPE) {
        double p = 0.0;
        if (e.status()[l]) p = c.line_bdc[l] * (e.theta()[e.fbus()[l]] - e.theta()[e.tbus()[l]]) * c.base_mva;
        e.pf()[l] = p; e.pt()[l] = -p; e.qf()[l] = 0.0; e.qt()[l] = 0.0;
    }
    if (tid == 0) {
        // gen[refgen, PG] += (B[ref, :] Va - Pbus[ref]) baseMVA
        const int s = ref >= S ? ref - S : ref;
        PPN_ENTRIES(e, c, ref, k0, step)
        double acc = 0.0;
        for (int q = 0; q < e.deg()[ref]; q++) {
            const int k = k0 + step * q;
            acc += c.line_bdc[e.eline()[k] >> 1] * (e.theta()[ref] - e.theta()[e.eoth()[k]]);
        }
        const int g = c.gen_of_sub[s];
        e.gpg()[g] = e.gpg()[g] + (acc - (e.pin()[ref] - c.bus_ysh_r[ref])) * c.base_mva;
    }
    env_sync<TPE>(mask);
    for (int b = tid; b < NB; b += TPE) {
        if (e.btype()[b] == PPN_BT_ISOLATED) continue;
        e.vm()[b] = 1.0;
        e.va()[b] = e.theta()[b] * (180.0 / PPN_PI);
    }
    success = true;
    return success;
}

// runpf with PF_ALG=2 (fast-decoupled XB, fdpf + pfsoln of SURVEY.md Appendix A) on the prepared env.  M1 / M2 hold
// B' (n1 x n1) and B'' (n2 x n2); SMEM says whether they are in shared memory (the common case: the accesses then
// compile to LDS/STS) or in the env's slice of the HBM workspace.
template <int TPE, int MAXR, class D, bool SMEM>
__device__ __forceinline__ bool ac_solve(Env<TPE, D>& e, const PpnDevCase& c, const PpnDevCfg& cfg, double* M1, double* M2,
                                         int n1, int n2, int ld1, int ld2, int ref, int slot, int& n_iter,
                                         const SpView* sp, double* spb, bool solve_mode, Border* bd = nullptr) {
    const int NB = e.NB, S = e.S, tid = e.tid;
    const unsigned mask = e.mask;
    bool success;
    // ================= runpf, PF_ALG=2 (fast-decoupled XB)
    // Every thread OWNS the buses tid, tid+TPE (one per thread for IEEE-14 and IEEE-118): their magnitude, angle,
    // unit phasor, injections, Ybus diagonal and list position stay in registers for the whole iteration; only
    // the rectangular voltages (read by neighbours) and the mismatch vectors (read by the solves) go through
    // shared memory.
    constexpr int RB = D::NB_MAX > 0 ? (D::NB_MAX + TPE - 1) / TPE : (TPE == 256 ? 1 : 2);   // <= 256 buses (CTA), <= 64 (warp)
    // CTA-per-env kernels (two buses per thread, 168 registers, out-of-line solver calls) keep the loop-invariant part of
    // that state -- injections, Ybus diagonal -- in shared memory and recompute the bus powers for pfsoln: fewer live
    // registers across the solver calls, i.e. less local-memory traffic (ncu: 27 KB of spill write-backs per env-step)
    constexpr bool LEAN = TPE > 32;
    double r_vm[RB], r_rvm[RB], r_va[RB], r_cs[RB], r_sn[RB], r_pin[RB], r_qin[RB], r_ydr[RB], r_ydi[RB], r_sr[RB], r_si[RB];
    int r_t[RB], r_ip[RB], r_iq[RB], r_deg[RB], r_k0[RB], r_step[RB];
    // solve mode: the mismatch vectors are indexed by factor row (zero on the identity rows) and are solved in place
    if (solve_mode) {
        for (int i = tid; i < 2 * NB; i += TPE) e.P()[i] = 0.0;   // P | Q are adjacent
    }
    PPN_TICK(4);
    // V0 from the stored state; on-line generators impose their set-point magnitude
#pragma unroll
    for (int r = 0; r < RB; r++) {
        const int b = tid + r * TPE;
        const int t = b < NB ? e.btype()[b] : PPN_BT_ISOLATED;
        r_t[r] = t;
        r_vm[r] = 1.0; r_rvm[r] = 1.0; r_va[r] = 0.0; r_cs[r] = 1.0; r_sn[r] = 0.0; r_pin[r] = 0.0; r_qin[r] = 0.0;
        r_ydr[r] = 0.0; r_ydi[r] = 0.0; r_sr[r] = 0.0; r_si[r] = 0.0;
        r_ip[r] = 0; r_iq[r] = 0; r_deg[r] = 0; r_k0[r] = 0; r_step[r] = 1;
        if (t == PPN_BT_ISOLATED) continue;
        double sn, cs;
        sincos(e.va()[b] * (PPN_PI / 180.0), &sn, &cs);
        double vr = e.vm()[b] * cs, vi = e.vm()[b] * sn;
        const int s = b >= S ? b - S : b, node = b >= S ? 1 : 0;
        const int g = c.gen_of_sub[s];
        if (g >= 0 && e.gnode()[g] == node && e.gstat()[g] > 0) {
            const double sc = e.gvg()[g] / hypot(vr, vi);
            vr *= sc; vi *= sc;
        }
        reinterpret_cast<double2*>(e.vri())[b] = make_double2(vr, vi);
        const double vm = hypot(vr, vi);   // fdpf: Vm = abs(V0), Va = angle(V0)
        r_vm[r] = vm;
        r_rvm[r] = 1.0 / vm;
        r_va[r] = atan2(vi, vr);           // radians
        r_cs[r] = vr / vm; r_sn[r] = vi / vm;
        if (!LEAN) { r_pin[r] = e.pin()[b]; r_qin[r] = e.qin()[b]; }
        r_ip[r] = t != PPN_BT_REF ? (solve_mode ? (int)sp->bus_row[b] : (int)e.idxp()[b]) : 0;
        r_iq[r] = t == PPN_BT_PQ ? (solve_mode ? (int)sp->bus_row[b] : (int)e.idxq()[b]) : 0;
        r_deg[r] = e.deg()[b];
        r_step[r] = b >= S ? -1 : 1;
        r_k0[r] = b >= S ? c.adj_ptr[s + 1] - 1 : c.adj_ptr[s];
    }
    if (bd) {
        // sister buses of the bordered system: their entries of the solve vectors sit behind the factor's rows, in
        // bus order (= the order of their compact indices: sisters come last in the bus array)
        int c1 = 0, c2 = 0;
#pragma unroll
        for (int r = 0; r < RB; r++) {
            const int b = tid + r * TPE;
            if (b >= S && b < NB) { c1 += (r_t[r] == PPN_BT_PV || r_t[r] == PPN_BT_PQ); c2 += (r_t[r] == PPN_BT_PQ); }
        }
        bd->k[0] = env_sum_int<TPE>(c1, e.redi(), tid, mask);
        bd->k[1] = env_sum_int<TPE>(c2, e.redi() + Env<TPE, D>::NW, tid, mask);
        bd->base[0] = n1 - bd->k[0]; bd->base[1] = n2 - bd->k[1];
#pragma unroll
        for (int r = 0; r < RB; r++) {
            const int b = tid + r * TPE;
            if (b >= S && b < NB) {
                if (r_t[r] == PPN_BT_PV || r_t[r] == PPN_BT_PQ) r_ip[r] = sp->n + (int)e.idxp()[b] - bd->base[0];
                if (r_t[r] == PPN_BT_PQ) r_iq[r] = sp->n + (int)e.idxq()[b] - bd->base[1];
            }
        }
        for (int i = tid; i < bd->k[0] * sp->n; i += TPE) bd->W[0][i] = 0.0;
        for (int i = tid; i < bd->k[1] * sp->n; i += TPE) bd->W[1][i] = 0.0;
        for (int i = tid; i < PPN_BORDER_MAX * PPN_BORDER_MAX; i += TPE) { bd->Sc[0][i] = 0.0; bd->Sc[1][i] = 0.0; }
    }
    PPN_TICK(5);
    // B' (r = 0, no charging, no shunts, unit taps) over pv+pq; B'' = -Im(Ybus) over pq; Ybus diagonal
    // dense: assembled in place in M1 / M2; sparse: lower triangle into the static pattern, each entry written by the
    // thread that owns the bus with the larger row (parallel lines accumulate in that bus's list order)
    SpFactor f1{}, f2{};
    if (sp) {
        f1 = sp_carve(spb, *sp, 0); f2 = sp_carve(spb, *sp, 1);
        sp_clear<TPE>(*sp, f1, n1, tid); sp_clear<TPE>(*sp, f2, n2, tid);
    } else {
        for (int i = tid; i < n1 * ld1 + n2 * ld2; i += TPE) M1[i] = 0.0;
    }
    env_sync<TPE>(mask);
#pragma unroll
    for (int r = 0; r < RB; r++) {
        const int t = r_t[r];
        if (t == PPN_BT_ISOLATED) continue;
        const int b = tid + r * TPE;
        const bool ispq = t == PPN_BT_PQ, inp = t != PPN_BT_REF;
        const int i = r_ip[r], iq = r_iq[r];
        const int row = sp ? sp->bus_row[b] : 0;
        double d1 = 0.0, yr = c.bus_ysh_r[b], yi = c.bus_ysh_i[b];
        for (int q = 0; q < r_deg[r]; q++) {
            const int k = r_k0[r] + r_step[r] * q;
            const int a = e.eline()[k], l = a >> 1, end = a & 1;
            const int o = e.eoth()[k];
            const double w = c.line_bp[l];
            const double* y = c.line_y + 8 * l + (end ? 6 : 0);   // ytt or yff
            yr += y[0];
            yi += y[1];
            d1 += w;
            const int to = e.btype()[o];
            if (!sp) {
                if (inp && to != PPN_BT_REF) M1[i * ld1 + e.idxp()[o]] -= w;
                if (ispq && to == PPN_BT_PQ) M2[iq * ld2 + e.idxq()[o]] -= e.ey()[2 * k + 1];
            } else if (bd && (b >= S || o >= S)) {
                // bordered system: entries of B are written by the owner of the real bus (column = the sister), entries
                // of C by the owner of the sister bus that is the row
                if (b < S) {
                    if (inp && to != PPN_BT_REF) bd->W[0][((int)e.idxp()[o] - bd->base[0]) * sp->n + row] -= w;
                    if (ispq && to == PPN_BT_PQ) bd->W[1][((int)e.idxq()[o] - bd->base[1]) * sp->n + row] -= e.ey()[2 * k + 1];
                } else if (o >= S) {
                    if (inp && to != PPN_BT_REF) bd->Sc[0][(i - sp->n) * PPN_BORDER_MAX + ((int)e.idxp()[o] - bd->base[0])] -= w;
                    if (ispq && to == PPN_BT_PQ) bd->Sc[1][(iq - sp->n) * PPN_BORDER_MAX + ((int)e.idxq()[o] - bd->base[1])] -= e.ey()[2 * k + 1];
                }
            } else if (row > sp->bus_row[o]) {
                const int pos = sp->line_pos[sp->full ? 4 * l + 2 * e.onode()[l] + e.enode()[l] : l];
                if (inp && to != PPN_BT_REF) f1.Lv[pos] -= w;
                if (ispq && to == PPN_BT_PQ) f2.Lv[pos] -= e.ey()[2 * k + 1];
            }
        }
        if (LEAN) { e.ydr()[b] = yr; e.ydi()[b] = yi; } else { r_ydr[r] = yr; r_ydi[r] = yi; }
        if (!sp) {
            if (inp) M1[i * ld1 + i] += d1;
            if (ispq) M2[iq * ld2 + iq] += -yi;
        } else if (bd && b >= S) {
            if (inp) bd->Sc[0][(i - sp->n) * PPN_BORDER_MAX + (i - sp->n)] += d1;
            if (ispq) bd->Sc[1][(iq - sp->n) * PPN_BORDER_MAX + (iq - sp->n)] += -yi;
        } else {
            if (inp) { f1.dg[row] = d1; f1.cp[row] = (short)e.idxp()[b]; }
            if (ispq) { f2.dg[row] = -yi; f2.cp[row] = (short)e.idxq()[b]; }
        }
    }
    env_sync<TPE>(mask);
    // fdpf: evaluate, then alternate P (angle) and Q (magnitude) half-iterations, testing after each
    PPN_TICK(6);
    success = false;
    int half = 0;
    while (true) {
#ifdef PPN_TIMING
        long long t_a = clock64();
#endif
        if (half > 0) {
            if (solve_mode) {   // B'^-1 P or B''^-1 Q in place (every thread takes part)
#ifdef PPN_TIMING
                long long t_s = clock64();
#endif
                const SpFactor& fh = (half & 1) ? f1 : f2;
                double* wh = (half & 1) ? e.P() : e.Q();
                if (TPE > 32 && sp->hyb) {
                    const int ldz = sp->d->nt | 1;
                    const double* Zh = spb + 2 * ppn_sp_factor_doubles(sp->n, sp->nnz) + ((half & 1) ? 0 : sp->d->nt * ldz);
                    hyb_solve_any<TPE>(*sp->d, sp->tb, saddr(fh.Lv), saddr(fh.dg), Zh, ldz, saddr(wh), tid);
                    const int m = (half & 1) ? 0 : 1;
                    if (bd && bd->k[m] > 0) {   // y = S^-1 (g - B^T t), x = t - W y
                        const int nU = sp->n, kh = bd->k[m];
#pragma unroll
                        for (int r = 0; r < RB; r++) {
                            const int b = tid + r * TPE, t = r_t[r];
                            if (b < S || b >= NB || !(m == 0 ? (t == PPN_BT_PV || t == PPN_BT_PQ) : t == PPN_BT_PQ)) continue;
                            const int iw = m == 0 ? r_ip[r] : r_iq[r];
                            double acc = wh[iw];
                            for (int q = 0; q < r_deg[r]; q++) {
                                const int k = r_k0[r] + r_step[r] * q;
                                const int o = e.eoth()[k];
                                const int to = e.btype()[o];
                                if (o < S && (m == 0 ? to != PPN_BT_REF : to == PPN_BT_PQ))
                                    acc = fma(-border_coef(e, c, k, m), wh[sp->bus_row[o]], acc);
                            }
                            wh[iw] = acc;
                        }
                        __syncthreads();
                        double y = 0.0;
                        if (tid < kh)
                            for (int j = 0; j < kh; j++) y = fma(bd->Sc[m][tid * PPN_BORDER_MAX + j], wh[nU + j], y);
                        __syncthreads();
                        if (tid < kh) wh[nU + tid] = y;
                        __syncthreads();
                        for (int i = tid; i < nU; i += TPE) {
                            double acc = wh[i];
                            for (int j = 0; j < kh; j++) acc = fma(-bd->W[m][j * nU + i], wh[nU + j], acc);
                            wh[i] = acc;
                        }
                        __syncthreads();
                    }
                }
                else if (TPE > 32 && sp->tb) sps_solve<TPE>(*sp->d, sp->tb, saddr(fh.Lv), saddr(fh.dg), saddr(wh), tid);
                else sp_solve<TPE>(*sp, fh, wh, tid, mask);
                PPN_TICK_ACC(23, t_s);   // triangular solves
            }
#pragma unroll
            for (int r = 0; r < RB; r++) {
                const int t = r_t[r];
                const int b = tid + r * TPE;
                if (half & 1) {   // P iteration: Va[pvpq] -= B'^-1 P
                    if (t == PPN_BT_PV || t == PPN_BT_PQ) {
                        const double d = solve_mode ? e.P()[r_ip[r]] : row_dot(M1 + r_ip[r] * ld1, e.P(), n1);
                        r_va[r] -= d;
                        if (fabs(d) < 0.03125) {
                            // small step (every iteration but the first ones): rotate the unit phasor by -d with the Taylor
                            // polynomials of sin d and cos d (next terms d^9/9!, d^10/10! < 1e-19) instead of a full
                            // sincos of the new angle -- a quarter of its instructions on the dependent path
                            const double d2 = d * d;
                            const double sd = d * fma(d2, fma(d2, fma(d2, -1.0 / 5040.0, 1.0 / 120.0), -1.0 / 6.0), 1.0);
                            const double cd = fma(d2, fma(d2, fma(d2, fma(d2, 1.0 / 40320.0, -1.0 / 720.0), 1.0 / 24.0), -0.5), 1.0);
                            const double c0 = r_cs[r], s0 = r_sn[r];
                            r_cs[r] = fma(c0, cd, s0 * sd);
                            r_sn[r] = fma(s0, cd, -c0 * sd);
                        } else {
                            sincos(r_va[r], &r_sn[r], &r_cs[r]);
                        }
                        reinterpret_cast<double2*>(e.vri())[b] = make_double2(r_vm[r] * r_cs[r], r_vm[r] * r_sn[r]);
                    }
                } else if (t == PPN_BT_PQ) {   // Q iteration: Vm[pq] -= B''^-1 Q
                    r_vm[r] -= solve_mode ? e.Q()[r_iq[r]] : row_dot(M2 + r_iq[r] * ld2, e.Q(), n2);
                    r_rvm[r] = 1.0 / r_vm[r];
                    reinterpret_cast<double2*>(e.vri())[b] = make_double2(r_vm[r] * r_cs[r], r_vm[r] * r_sn[r]);
                }
            }
            env_sync<TPE>(mask);
        }
        PPN_TICK_ACC(20 + (half & 1), t_a);   // update half-steps: [20] Q, [21] P
#ifdef PPN_TIMING
        long long t_b = clock64();
#endif
        // mismatch: mis = (V conj(Ybus V) - Sbus)/Vm, P over pv+pq, Q over pq; both infinity norms < tol ?
        bool open = false;
#pragma unroll
        for (int r = 0; r < RB; r++) {
            const int t = r_t[r];
            if (t != PPN_BT_PV && t != PPN_BT_PQ) continue;
            const double vr = r_vm[r] * r_cs[r], vi = r_vm[r] * r_sn[r];
            const double ydr_ = LEAN ? e.ydr()[tid + r * TPE] : r_ydr[r], ydi_ = LEAN ? e.ydi()[tid + r * TPE] : r_ydi[r];
            double ir = ydr_ * vr - ydi_ * vi, ii = ydr_ * vi + ydi_ * vr;
            double jr = 0.0, ji = 0.0;
            {   // even entries accumulate in (ir, ii), odd ones in (jr, ji): two independent chains, one 16-byte load
                // per admittance and per neighbour voltage
                const double2* ey2 = reinterpret_cast<const double2*>(e.ey());
                const double2* v2 = reinterpret_cast<const double2*>(e.vri());
                const short* eo = e.eoth();
                const int deg = r_deg[r], st2 = r_step[r];
                int k = r_k0[r];
                int q = 0;
                for (; q + 1 < deg; q += 2, k += 2 * st2) {
                    const int o0 = eo[k], o1 = eo[k + st2];
                    const double2 y0 = ey2[k], y1 = ey2[k + st2];
                    const double2 w0 = v2[o0], w1 = v2[o1];
                    ir = fma(y0.x, w0.x, fma(-y0.y, w0.y, ir)); ii = fma(y0.x, w0.y, fma(y0.y, w0.x, ii));
                    jr = fma(y1.x, w1.x, fma(-y1.y, w1.y, jr)); ji = fma(y1.x, w1.y, fma(y1.y, w1.x, ji));
                }
                if (q < deg) {
                    const double2 y0 = ey2[k], w0 = v2[eo[k]];
                    ir = fma(y0.x, w0.x, fma(-y0.y, w0.y, ir)); ii = fma(y0.x, w0.y, fma(y0.y, w0.x, ii));
                }
            }
            ir += jr; ii += ji;
            const double sr = vr * ir + vi * ii, si = vi * ir - vr * ii;   // V conj(I)
            if (!LEAN) { r_sr[r] = sr; r_si[r] = si; }
            const double rvm = r_rvm[r];   // 1/Vm only changes in the Q half-iterations
            const double pm = (sr - (LEAN ? e.pin()[tid + r * TPE] : r_pin[r])) * rvm;
            e.P()[r_ip[r]] = pm;
            open |= !(fabs(pm) < cfg.tol);
            if (t == PPN_BT_PQ) {
                const double qm = (si - (LEAN ? e.qin()[tid + r * TPE] : r_qin[r])) * rvm;
                e.Q()[r_iq[r]] = qm;
                open |= !(fabs(qm) < cfg.tol);
            }
        }
        const bool any_open = env_any<TPE>(open, mask);
        env_sync<TPE>(mask);
        PPN_TICK_ACC(22, t_b);                  // mismatch evaluations
        if (!any_open) { success = true; break; }
        if (half == 2 * cfg.max_it) break;
        if (half == 0) {
            PPN_TICK(7);
            if (sp && TPE > 32 && sp->hyb) {
                const int ldz = sp->d->nt | 1;
                double* Z1 = spb + 2 * ppn_sp_factor_doubles(sp->n, sp->nnz);
                hyb_factor2<TPE>(*sp->d, sp->tb, SpsFactor{saddr(f1.T), saddr(f1.Lv), saddr(f1.dg)},
                                 SpsFactor{saddr(f2.T), saddr(f2.Lv), saddr(f2.dg)}, Z1, Z1 + sp->d->nt * ldz, ldz, tid);
                PPN_TICK(8);
                hyb_invert2<TPE>(Z1, Z1 + sp->d->nt * ldz, sp->d->nt, ldz, e.cs(), tid);   // cs | sn: 2 NB doubles of scratch (the Ybus diagonal sits in ydr | ydi)
                if (bd && (bd->k[0] | bd->k[1])) {
                    // W = A^-1 B column by column (the solve works on a vector in shared memory: cs | sn are free), then
                    // the Schur complements C - B^T W row by row (owner of the sister bus) and their inverses
                    const int nU = sp->n;
                    double* col = e.cs();
                    for (int m = 0; m < 2; m++) {
                        const SpFactor& fm = m == 0 ? f1 : f2;
                        const double* Zm = Z1 + (m == 0 ? 0 : sp->d->nt * ldz);
                        for (int j = 0; j < bd->k[m]; j++) {
                            for (int i = tid; i < nU; i += TPE) col[i] = bd->W[m][j * nU + i];
                            __syncthreads();
                            hyb_solve_any<TPE>(*sp->d, sp->tb, saddr(fm.Lv), saddr(fm.dg), Zm, ldz, saddr(col), tid);
                            for (int i = tid; i < nU; i += TPE) bd->W[m][j * nU + i] = col[i];
                            __syncthreads();
                        }
                    }
                    __threadfence_block();
#pragma unroll
                    for (int r = 0; r < RB; r++) {
                        const int b = tid + r * TPE, t = r_t[r];
                        if (b < S || b >= NB) continue;
                        for (int m = 0; m < 2; m++) {
                            if (!(m == 0 ? (t == PPN_BT_PV || t == PPN_BT_PQ) : t == PPN_BT_PQ)) continue;
                            const int ib = (m == 0 ? r_ip[r] : r_iq[r]) - nU;
                            for (int j = 0; j < bd->k[m]; j++) {
                                double acc = bd->Sc[m][ib * PPN_BORDER_MAX + j];
                                for (int q = 0; q < r_deg[r]; q++) {
                                    const int k = r_k0[r] + r_step[r] * q;
                                    const int o = e.eoth()[k];
                                    const int to = e.btype()[o];
                                    if (o < S && (m == 0 ? to != PPN_BT_REF : to == PPN_BT_PQ))
                                        acc = fma(-border_coef(e, c, k, m), bd->W[m][j * nU + sp->bus_row[o]], acc);
                                }
                                bd->Sc[m][ib * PPN_BORDER_MAX + j] = acc;
                            }
                        }
                    }
                    __syncthreads();
                    if (tid < 64) border_invert_warp(bd->Sc[tid >> 5], bd->k[tid >> 5], tid & 31);
                    __syncthreads();
                }
            } else if (sp) {
                if (TPE > 32 && sp->tb)
                    sps_factor2<TPE>(*sp->d, sp->tb, SpsFactor{saddr(f1.T), saddr(f1.Lv), saddr(f1.dg)},
                                     SpsFactor{saddr(f2.T), saddr(f2.Lv), saddr(f2.dg)}, tid);
                else sp_factor<TPE, 2>(*sp, f1, f2, tid, mask);
                PPN_TICK(8);
                if (solve_mode) sp_recip<TPE>(*sp, f1, f2, true, tid, mask);
                else sp_invert<TPE>(*sp, f1, f2, e.busp(), e.busq(), M1, M2, n1, n2, ld1, ld2, tid, mask);
            } else if (TPE == 32 && SMEM && n1 <= 16 && n2 <= 16 && NB >= 18) {   // the pivot buffers (ydr | ydi) hold 17 doubles each
                // both inverses at once: lanes 0-15 hold the rows of B', lanes 16-31 those of B''
                const int hf = tid >> 4;
                gj16_rows_in_registers(saddr(hf ? M2 : M1), hf ? n2 : n1, hf ? ld2 : ld1, tid & 15,
                                       n1 > n2 ? n1 : n2, saddr(hf ? e.ydi() : e.ydr()), mask);
            } else {
                gj_invert<TPE, MAXR>(M1, n1, ld1, tid, mask);
                gj_invert<TPE, MAXR>(M2, n2, ld2, tid, mask);
            }
            if (!sp) PPN_TICK(8);
            PPN_TICK(9);
        }
        half++;
    }
    n_iter = (half + 1) / 2;
    PPN_TICK(10);
#ifdef PPN_TIMING
    if (slot == 0 && tid == 0 && ppn_timing_buf[63] == 0) { ppn_timing_buf[30] = half; ppn_timing_buf[31] = n1; ppn_timing_buf[32] = n2; }
#endif
    // ---- pfsoln: generator Q (and the slack's P) by the thread that owns the generator's bus
    int n_on = 0;
    for (int g = tid; g < e.G; g += TPE) n_on += (e.gstat()[g] > 0 && e.btype()[e.gbus()[g]] != PPN_BT_ISOLATED);
    n_on = env_sum_int<TPE>(n_on, e.redi(), tid, mask);
#pragma unroll
    for (int r = 0; r < RB; r++) {
        const int t = r_t[r];
        if (t == PPN_BT_ISOLATED) continue;
        const int b = tid + r * TPE;
        const int s = b >= S ? b - S : b, node = b >= S ? 1 : 0;
        const int g = c.gen_of_sub[s];
        if (g >= 0 && e.gnode()[g] == node && e.gstat()[g] > 0) {
            double sr = LEAN ? 0.0 : r_sr[r], si = LEAN ? 0.0 : r_si[r];
            if (t == PPN_BT_REF || LEAN) {   // the reference bus is not part of the mismatch vectors
                const double vr = r_vm[r] * r_cs[r], vi = r_vm[r] * r_sn[r];
                const double ydr_ = LEAN ? e.ydr()[b] : r_ydr[r], ydi_ = LEAN ? e.ydi()[b] : r_ydi[r];
                double ir = ydr_ * vr - ydi_ * vi, ii = ydr_ * vi + ydi_ * vr;
                for (int q = 0; q < r_deg[r]; q++) {
                    const int k = r_k0[r] + r_step[r] * q;
                    const int o = e.eoth()[k];
                    const double yr = e.ey()[2 * k], yi = e.ey()[2 * k + 1];
                    const double wr = e.vri()[2 * o], wi = e.vri()[2 * o + 1];
                    ir = fma(yr, wr, fma(-yi, wi, ir));
                    ii = fma(yr, wi, fma(yi, wr, ii));
                }
                sr = vr * ir + vi * ii; si = vi * ir - vr * ii;
            }
            double pd, qd;
            bus_demand(e, c, b, pd, qd);
            double q = si * c.base_mva + qd;
            if (n_on > 1) {
                const double qmin = c.gen_qmin[g], qmax = c.gen_qmax[g];
                if (qmin != qmax) q = qmin + ((q - qmin) / (qmax - qmin + 2.220446049250313e-16)) * (qmax - qmin);
            }
            e.gqg()[g] = q;
            if (t == PPN_BT_REF) e.gpg()[g] = sr * c.base_mva + pd;
        }
        // adopted bus results: VM = |V|, VA = angle(V) in degrees
        const double vr = r_vm[r] * r_cs[r], vi = r_vm[r] * r_sn[r];
        e.vm()[b] = hypot(vr, vi);
        e.va()[b] = atan2(vi, vr) * (180.0 / PPN_PI);
    }
    env_sync<TPE>(mask);   // the branch results below reuse the storage of the mismatch vectors
    for (int l = tid; l < e.N; l += TPE) {
        double pf = 0.0, qf = 0.0, pt = 0.0, qt = 0.0;
        if (e.status()[l]) {
            const double* y = c.line_y + 8 * l;
            const int f = e.fbus()[l], t = e.tbus()[l];
            const double fr = e.vri()[2 * f], fi = e.vri()[2 * f + 1], tr = e.vri()[2 * t], ti = e.vri()[2 * t + 1];
            const double ifr = y[0] * fr - y[1] * fi + y[2] * tr - y[3] * ti;
            const double ifi = y[0] * fi + y[1] * fr + y[2] * ti + y[3] * tr;
            const double itr = y[4] * fr - y[5] * fi + y[6] * tr - y[7] * ti;
            const double iti = y[4] * fi + y[5] * fr + y[6] * ti + y[7] * tr;
            pf = (fr * ifr + fi * ifi) * c.base_mva; qf = (fi * ifr - fr * ifi) * c.base_mva;
            pt = (tr * itr + ti * iti) * c.base_mva; qt = (ti * itr - tr * iti) * c.base_mva;
        }
        e.pf()[l] = pf; e.qf()[l] = qf; e.pt()[l] = pt; e.qt()[l] = qt;
    }
    return success;
}

// ---- Newton-Raphson (PF_ALG = 1): the north star's named solver; the reference itself runs the fast-decoupled one -----
// Dense system A x = b, A n x n with the right-hand side in column n (row stride ld), Gaussian elimination with partial
// pivoting, one row per thread in the elimination.  The matrix lives in shared memory when the handle's plan has room
// for it (warp-per-env grids: IEEE-14 22 x 23, IEEE-30 53 x 54 doubles), else in the env's slice of the HBM workspace
// (an IEEE-118 Jacobian is 181 x 181 = 262 KB; this solver is an option, not the benchmarked path).
// Returns false on an exactly singular matrix (PYPOWER's spsolve then yields NaN: the iteration never converges).
template <int TPE>
__device__ __forceinline__ bool dense_solve_pivot(double* A, int n, int ld, int tid, unsigned mask, double* redd, int* redi) {
    for (int k = 0; k < n; k++) {
        double best = -1.0;
        int bi = 0x7fffffff;
        for (int i = k + tid; i < n; i += TPE) {
            const double v = fabs(A[(size_t)i * ld + k]);
            if (v > best) { best = v; bi = i; }
        }
        const double vmax = env_max_nan<TPE>(best, redd, tid, mask);
        if (!(vmax > 0.0)) return false;   // zero or NaN column
        const int p = env_min_int<TPE>(best == vmax ? bi : 0x7fffffff, redi, tid, mask);
        if (p != k)
            for (int j = k + tid; j <= n; j += TPE) {
                const double a = A[(size_t)k * ld + j], b = A[(size_t)p * ld + j];
                A[(size_t)k * ld + j] = b; A[(size_t)p * ld + j] = a;
            }
        env_sync<TPE>(mask);
        const double rp = 1.0 / A[(size_t)k * ld + k];
        for (int i = k + 1 + tid; i < n; i += TPE) {
            double* ai = A + (size_t)i * ld;
            const double* ak = A + (size_t)k * ld;
            const double f = ai[k] * rp;
            if (f != 0.0)
                for (int j = k + 1; j <= n; j++) ai[j] = fma(-f, ak[j], ai[j]);
        }
        env_sync<TPE>(mask);
    }
    for (int k = n - 1; k >= 0; k--) {
        const double xk = A[(size_t)k * ld + n] / A[(size_t)k * ld + k];
        env_sync<TPE>(mask);
        if (tid == 0) A[(size_t)k * ld + n] = xk;
        for (int i = tid; i < k; i += TPE) A[(size_t)i * ld + n] = fma(-A[(size_t)i * ld + k], xk, A[(size_t)i * ld + n]);
        env_sync<TPE>(mask);
    }
    return true;
}

// runpf with PF_ALG = 1 (PYPOWER newtonpf + pfsoln):
// F = [Re mis[pv+pq]; Im mis[pq]], mis = V conj(Ybus V) - Sbus; J = [[Re dS/dVa, Re dS/dVm], [Im dS/dVa, Im dS/dVm]]
// from dSbus_dV; dx = -J^-1 F; tolerance on |F|_inf; at most max_it_nr iterations.  Unknown / equation order: angles
// of pv+pq buses by compact index idxp, then magnitudes of pq buses by idxq (any consistent order gives the same dx).
// J: (n1+n2) x (n1+n2+1) doubles in the HBM workspace.
template <int TPE, class D>
__device__ __forceinline__ bool nr_solve(Env<TPE, D>& e, const PpnDevCase& c, const PpnDevCfg& cfg, double* J, int n1, int n2,
                                         int ref, int& n_iter) {
    const int NB = e.NB, S = e.S, tid = e.tid, nJ = n1 + n2, ld = nJ + 1;
    const unsigned mask = e.mask;
    // V0 from the stored state (on-line generators impose their set-point magnitude); va holds RADIANS until the end
    for (int b = tid; b < NB; b += TPE) {
        if (e.btype()[b] == PPN_BT_ISOLATED) continue;
        double sn, cs;
        sincos(e.va()[b] * (PPN_PI / 180.0), &sn, &cs);
        double vr = e.vm()[b] * cs, vi = e.vm()[b] * sn;
        const int s = b >= S ? b - S : b, node = b >= S ? 1 : 0;
        const int g = c.gen_of_sub[s];
        if (g >= 0 && e.gnode()[g] == node && e.gstat()[g] > 0) {
            const double sc = e.gvg()[g] / hypot(vr, vi);
            vr *= sc; vi *= sc;
        }
        e.vri()[2 * b] = vr; e.vri()[2 * b + 1] = vi;
        e.vm()[b] = hypot(vr, vi);
        e.va()[b] = atan2(vi, vr);
        // Ybus diagonal: shunt + own-end admittances of the in-service lines on this bus
        PPN_ENTRIES(e, c, b, k0, step)
        double yr = c.bus_ysh_r[b], yi = c.bus_ysh_i[b];
        for (int q = 0; q < e.deg()[b]; q++) {
            const int a = e.eline()[k0 + step * q];
            const double* y = c.line_y + 8 * (a >> 1) + ((a & 1) ? 6 : 0);
            yr += y[0]; yi += y[1];
        }
        e.ydr()[b] = yr; e.ydi()[b] = yi;
    }
    env_sync<TPE>(mask);
    bool success = false;
    int it = 0;
    while (true) {
        // bus currents (kept for the Jacobian), power mismatch
        bool open = false;
        for (int b = tid; b < NB; b += TPE) {
            const int t = e.btype()[b];
            if (t == PPN_BT_ISOLATED) continue;
            PPN_ENTRIES(e, c, b, k0, step)
            const double vr = e.vri()[2 * b], vi = e.vri()[2 * b + 1];
            double ir = e.ydr()[b] * vr - e.ydi()[b] * vi, ii = e.ydr()[b] * vi + e.ydi()[b] * vr;
            for (int q = 0; q < e.deg()[b]; q++) {
                const int k = k0 + step * q;
                const int o = e.eoth()[k];
                const double yr = e.ey()[2 * k], yi = e.ey()[2 * k + 1];
                const double wr = e.vri()[2 * o], wi = e.vri()[2 * o + 1];
                ir = fma(yr, wr, fma(-yi, wi, ir));
                ii = fma(yr, wi, fma(yi, wr, ii));
            }
            e.cs()[b] = ir; e.sn()[b] = ii;
            if (t == PPN_BT_REF) continue;
            const double fp = (vr * ir + vi * ii) - e.pin()[b];
            e.P()[e.idxp()[b]] = fp;
            open |= !(fabs(fp) < cfg.tol);
            if (t == PPN_BT_PQ) {
                const double fq = (vi * ir - vr * ii) - e.qin()[b];
                e.Q()[e.idxq()[b]] = fq;
                open |= !(fabs(fq) < cfg.tol);
            }
        }
        const bool any_open = env_any<TPE>(open, mask);
        env_sync<TPE>(mask);
        if (!any_open) { success = true; break; }
        if (it == cfg.max_it_nr) break;
        it++;
        // Jacobian, one bus (= up to two rows) per thread; right-hand side F
        for (int i = tid; i < nJ * ld; i += TPE) J[i] = 0.0;
        env_sync<TPE>(mask);
        for (int b = tid; b < NB; b += TPE) {
            const int t = e.btype()[b];
            if (t != PPN_BT_PV && t != PPN_BT_PQ) continue;
            const bool ispq = t == PPN_BT_PQ;
            double* rowp = J + (size_t)e.idxp()[b] * ld;
            double* rowq = ispq ? J + (size_t)(n1 + e.idxq()[b]) * ld : nullptr;
            PPN_ENTRIES(e, c, b, k0, step)
            const double vr = e.vri()[2 * b], vi = e.vri()[2 * b + 1], vm = e.vm()[b];
            double wr = 0.0, wi = 0.0;   // sum of the off-diagonal terms Y_bo V_o = I_b - Y_bb V_b
            for (int q = 0; q < e.deg()[b]; q++) {
                const int k = k0 + step * q;
                const int o = e.eoth()[k];
                const int to = e.btype()[o];
                const double yr = e.ey()[2 * k], yi = e.ey()[2 * k + 1];
                const double ur = yr * e.vri()[2 * o] - yi * e.vri()[2 * o + 1], ui = yr * e.vri()[2 * o + 1] + yi * e.vri()[2 * o];
                wr += ur; wi += ui;
                // T = V_b conj(Y_bo V_o):  dS_b/dVa_o = -j T,  dS_b/dVm_o = T / |V_o|
                const double tr = vr * ur + vi * ui, ti = vi * ur - vr * ui;
                if (to == PPN_BT_REF) continue;
                const int cp = e.idxp()[o];
                rowp[cp] += ti;
                if (rowq) rowq[cp] += -tr;
                if (to == PPN_BT_PQ) {
                    const int cq = n1 + e.idxq()[o];
                    const double rvo = 1.0 / e.vm()[o];
                    rowp[cq] += tr * rvo;
                    if (rowq) rowq[cq] += ti * rvo;
                }
            }
            // diagonal: dS_b/dVa_b = j V_b conj(I_b - Y_bb V_b),  dS_b/dVm_b = (|V_b|^2 conj(Y_bb) + conj(I_b) V_b) / |V_b|
            const double xr = vr * wr + vi * wi, xi = vi * wr - vr * wi;
            const int cp = e.idxp()[b];
            rowp[cp] += -xi;
            if (rowq) rowq[cp] += xr;
            if (ispq) {
                const double ir = e.cs()[b], ii = e.sn()[b];
                const double mr = (vm * vm * e.ydr()[b] + (ir * vr + ii * vi)) / vm;
                const double mi = (-vm * vm * e.ydi()[b] + (ir * vi - ii * vr)) / vm;
                const int cq = n1 + e.idxq()[b];
                rowp[cq] += mr;
                rowq[cq] += mi;
            }
            rowp[nJ] = e.P()[e.idxp()[b]];
            if (rowq) rowq[nJ] = e.Q()[e.idxq()[b]];
        }
        env_sync<TPE>(mask);
        const bool solved = dense_solve_pivot<TPE>(J, nJ, ld, tid, mask, e.redd(), e.redi());
        // dx = -J^-1 F; V = Vm e^{j Va}; Vm = |V|, Va = angle(V)
        for (int b = tid; b < NB; b += TPE) {
            const int t = e.btype()[b];
            if (t != PPN_BT_PV && t != PPN_BT_PQ) continue;
            const double nanv = nan("");
            double va = e.va()[b] - (solved ? J[(size_t)e.idxp()[b] * ld + nJ] : nanv);
            double vm = e.vm()[b];
            if (t == PPN_BT_PQ) vm -= solved ? J[(size_t)(n1 + e.idxq()[b]) * ld + nJ] : nanv;
            double sn, cs;
            sincos(va, &sn, &cs);
            const double vr = vm * cs, vi = vm * sn;
            e.vri()[2 * b] = vr; e.vri()[2 * b + 1] = vi;
            e.vm()[b] = hypot(vr, vi);
            e.va()[b] = atan2(vi, vr);
        }
        env_sync<TPE>(mask);
    }
    n_iter = it;
    // ---- pfsoln (as the fast-decoupled path): generator Q, the slack's P, bus results, branch flows
    int n_on = 0;
    for (int g = tid; g < e.G; g += TPE) n_on += (e.gstat()[g] > 0 && e.btype()[e.gbus()[g]] != PPN_BT_ISOLATED);
    n_on = env_sum_int<TPE>(n_on, e.redi(), tid, mask);
    for (int b = tid; b < NB; b += TPE) {
        const int t = e.btype()[b];
        if (t == PPN_BT_ISOLATED) continue;
        const int s = b >= S ? b - S : b, node = b >= S ? 1 : 0;
        const int g = c.gen_of_sub[s];
        const double vr = e.vri()[2 * b], vi = e.vri()[2 * b + 1];
        if (g >= 0 && e.gnode()[g] == node && e.gstat()[g] > 0) {
            const double ir = e.cs()[b], ii = e.sn()[b];
            const double sr = vr * ir + vi * ii, si = vi * ir - vr * ii;
            double pd, qd;
            bus_demand(e, c, b, pd, qd);
            double q = si * c.base_mva + qd;
            if (n_on > 1) {
                const double qmin = c.gen_qmin[g], qmax = c.gen_qmax[g];
                if (qmin != qmax) q = qmin + ((q - qmin) / (qmax - qmin + 2.220446049250313e-16)) * (qmax - qmin);
            }
            e.gqg()[g] = q;
            if (t == PPN_BT_REF) e.gpg()[g] = sr * c.base_mva + pd;
        }
        e.vm()[b] = hypot(vr, vi);
        e.va()[b] = atan2(vi, vr) * (180.0 / PPN_PI);
    }
    env_sync<TPE>(mask);   // the branch results below reuse the storage of the mismatch vectors
    for (int l = tid; l < e.N; l += TPE) {
        double pf = 0.0, qf = 0.0, pt = 0.0, qt = 0.0;
        if (e.status()[l]) {
            const double* y = c.line_y + 8 * l;
            const int f = e.fbus()[l], t = e.tbus()[l];
            const double fr = e.vri()[2 * f], fi = e.vri()[2 * f + 1], tr = e.vri()[2 * t], ti = e.vri()[2 * t + 1];
            const double ifr = y[0] * fr - y[1] * fi + y[2] * tr - y[3] * ti;
            const double ifi = y[0] * fi + y[1] * fr + y[2] * ti + y[3] * tr;
            const double itr = y[4] * fr - y[5] * fi + y[6] * tr - y[7] * ti;
            const double iti = y[4] * fi + y[5] * fr + y[6] * ti + y[7] * tr;
            pf = (fr * ifr + fi * ifi) * c.base_mva; qf = (fi * ifr - fr * ifi) * c.base_mva;
            pt = (tr * itr + ti * iti) * c.base_mva; qt = (ti * itr - tr * iti) * c.base_mva;
        }
        e.pf()[l] = pf; e.qf()[l] = qf; e.pt()[l] = pt; e.qt()[l] = qt;
    }
    return success;
}

// One load-flow on the current topology/injections (grid.py:244-264 around runpf / rundcpf).  Returns true when the
// reference raises DivergingLoadflowException.  On success the state (vm, va in degrees, gen pg/qg, flows) is the
// adopted output (`self.mpc = output`, grid.py:260).
template <int TPE, int MAXR, class D>
__device__ __forceinline__ bool loadflow(Env<TPE, D>& e, const PpnDevCase& c, const PpnDevCfg& cfg, const PpnStepArgs& args, int slot,
                         int& n_iter) {
    const int NB = e.NB, S = e.S, tid = e.tid;
    const unsigned mask = e.mask;
    n_iter = 0;
    PPN_TICK(0);
    compute_isolated(e);  // mark = isolated
    // ---- _synchronize_bus_types (grid.py:140-174) folded with bustypes: a bus whose generator is off is PQ
    int slack = c.slack_bus;
    if (e.mark()[slack]) {
        int found = -1;
        for (int g = 0; g < e.G; g++)
            if (e.gbus()[g] != c.slack_bus) { found = e.gbus()[g]; break; }
        slack = found;
    }
    for (int b = tid; b < NB; b += TPE) {
        const int s = b >= S ? b - S : b, node = b >= S ? 1 : 0;
        const int g = c.gen_of_sub[s];
        const bool has_gen = g >= 0 && e.gnode()[g] == node;
        int t = e.mark()[b] ? PPN_BT_ISOLATED : (has_gen ? PPN_BT_PV : PPN_BT_PQ);
        if (b == slack && !e.mark()[b] && has_gen) t = PPN_BT_REF;
        const bool on_gen = has_gen && e.gstat()[g] > 0;
        if (t != PPN_BT_ISOLATED && !on_gen) t = PPN_BT_PQ;
        e.btype()[b] = (uint8_t)t;
    }
    env_sync<TPE>(mask);
    // ---- bustypes: reference bus, compact indices (the group's first warp scans the buses in bus-array order)
    if (TPE <= 32 || tid < 32) {
        constexpr int CH = TPE < 32 ? TPE : 32;
        const unsigned gm = TPE <= 32 ? mask : PPN_FULL;
        const int sh = TPE <= 32 ? e.shift : 0;
        const int ln = TPE <= 32 ? tid : (tid & 31);
        const unsigned chm = CH == 32 ? 0xffffffffu : ((1u << CH) - 1u);
        int ref = -1, firstpv = -1, np_ = 0, nq_ = 0;
        for (int base = 0; base < NB; base += CH) {
            const int b = base + ln;
            const int t = b < NB ? e.btype()[b] : PPN_BT_ISOLATED;
            const unsigned mr = (__ballot_sync(gm, t == PPN_BT_REF) >> sh) & chm;
            const unsigned mv = (__ballot_sync(gm, t == PPN_BT_PV) >> sh) & chm;
            if (ref < 0 && mr) ref = base + __ffs(mr) - 1;
            if (firstpv < 0 && mv) firstpv = base + __ffs(mv) - 1;
        }
        if (ref < 0) ref = firstpv;  // ref = pv[0] (IndexError in PYPOWER when there is none)
        if (ref >= 0 && ln == 0) e.btype()[ref] = PPN_BT_REF;
        __syncwarp(gm);
        for (int base = 0; base < NB; base += CH) {
            const int b = base + ln;
            const int t = b < NB ? e.btype()[b] : PPN_BT_ISOLATED;
            const bool inp = (t == PPN_BT_PV || t == PPN_BT_PQ), inq = (t == PPN_BT_PQ);
            const unsigned mp = (__ballot_sync(gm, inp) >> sh) & chm, mq = (__ballot_sync(gm, inq) >> sh) & chm;
            const unsigned lt = (1u << ln) - 1u;
            if (inp) { const int i = np_ + __popc(mp & lt); e.idxp()[b] = (short)i; e.busp()[i] = (short)b; }
            if (inq) { const int i = nq_ + __popc(mq & lt); e.idxq()[b] = (short)i; e.busq()[i] = (short)b; }
            np_ += __popc(mp); nq_ += __popc(mq);
        }
        if (ln == 0) { e.misc()[0] = ref; e.misc()[1] = np_; e.misc()[2] = nq_; }
    }
    env_sync<TPE>(mask);
    const int ref = e.misc()[0], n1 = e.misc()[1], n2 = e.misc()[2];
    PPN_TICK(1);
    if (ref < 0) return true;
    // ---- connectivity: every non-isolated bus must be reachable from the reference bus over in-service lines
    for (int b = tid; b < NB; b += TPE) e.mark()[b] = (b == ref) ? 1 : 0;   // mark = reached
    env_sync<TPE>(mask);
    while (true) {
        bool changed = false;
        for (int l = tid; l < e.N; l += TPE) {
            if (!e.status()[l]) continue;
            const int f = e.fbus()[l], t = e.tbus()[l];
            const int rf = e.mark()[f], rt = e.mark()[t];
            if (rf != rt) { e.mark()[f] = 1; e.mark()[t] = 1; changed = true; }
        }
        if (!env_any<TPE>(changed, mask)) break;
        env_sync<TPE>(mask);
    }
    env_sync<TPE>(mask);
    bool floating = false;
    for (int b = tid; b < NB; b += TPE) floating |= (e.btype()[b] != PPN_BT_ISOLATED && !e.mark()[b]);
    // generators that are out of service (off, or on an isolated bus) end every adopted load-flow with Pg = Qg = 0
    if (env_any<TPE>(floating, mask)) {
        for (int g = tid; g < e.G; g += TPE)
            if (!(e.gstat()[g] > 0 && e.btype()[e.gbus()[g]] != PPN_BT_ISOLATED)) { e.gpg()[g] = 0.0; e.gqg()[g] = 0.0; }
        env_sync<TPE>(mask);
        return true;
    }
    if (n1 == 0 || (n2 == 0 && !cfg.dc)) return true;
    PPN_TICK(2);
    // ---- per-bus demand and makeSbus
    build_entries(e, c);
    PPN_TICK(3);
    for (int b = tid; b < NB; b += TPE) {
        if (e.btype()[b] == PPN_BT_ISOLATED) continue;
        const int s = b >= S ? b - S : b, node = b >= S ? 1 : 0;
        const int g = c.gen_of_sub[s];
        const bool on_gen = g >= 0 && e.gnode()[g] == node && e.gstat()[g] > 0;
        double pd, qd;
        bus_demand(e, c, b, pd, qd);
        double p = -pd, q = -qd;
        if (on_gen) { p += e.gpg()[g]; q += e.gqg()[g]; }
        e.pin()[b] = p / c.base_mva;
        e.qin()[b] = q / c.base_mva;
    }
    // ---- matrices: shared memory when they fit, else the env's slice of the global workspace
    const int ld1 = n1 | 1, ld2 = n2 | 1;
    bool success;
    if (!cfg.dc && cfg.alg == 1) {
        // Newton-Raphson (ppn_config.pf_alg = 1): dense Jacobian in the env's slice of the workspace
        env_sync<TPE>(mask);
        const int nJ = n1 + n2;
        double* Jbuf = nJ * (nJ + 1) <= args.mat_cap ? e.mat() : args.ws + (size_t)slot * args.ws_stride;   // shared memory when it fits
        success = nr_solve<TPE, D>(e, c, cfg, Jbuf, n1, n2, ref, n_iter);
    } else if (args.sparse) {
        // sparse LDL^T on the static pattern (U while no sister bus is in use, else F), then explicit inverses with
        // one landing row each; the factor storage sits behind the inverses when shared memory has room for it
        int n_sis = 0;
        for (int b = S + tid; b < NB; b += TPE) n_sis += e.btype()[b] != PPN_BT_ISOLATED;
        n_sis = env_sum_int<TPE>(n_sis, e.redi(), tid, mask);
        // hybrid factor, AC: a handful of sister buses border the one-row-per-substation factor (see Border) instead of
        // moving the env to the two-rows-per-substation structure
        const bool bordered = TPE > 32 && args.sparse == 3 && !cfg.dc && n_sis > 0 && n_sis <= PPN_BORDER_MAX &&
                              2 * ppn_sp_factor_doubles(c.sp[0].n, c.sp[0].nnz) + 2 * c.sp[0].nt * (c.sp[0].nt | 1) + c.sp[0].hyb_words / 2 <= args.mat_cap &&
                              2 * PPN_BORDER_MAX * c.sp[0].n + 2 * PPN_BORDER_MAX * PPN_BORDER_MAX <= args.ws_stride;   // the hybrid plan is in place
        const int which = (n_sis > 0 && !bordered) ? 1 : 0;
        const PpnDevSparse& spd = c.sp[which];
        const bool solve_mode = args.sparse >= 2;
        const int dense_need = solve_mode ? 0 : (n1 + 1) * ld1 + (cfg.dc ? 0 : (n2 + 1) * ld2);
        const int val_need = 2 * ppn_sp_factor_doubles(spd.n, spd.nnz) + (args.sparse == 3 ? 2 * spd.nt * (spd.nt | 1) : 0),
                  stage_words = args.sparse == 3 ? spd.hyb_words : spd.blob_words,   // the hybrid solver only reads a prefix
                  blob_dbl = stage_words / 2;
        double* wsrow = args.ws + (size_t)slot * args.ws_stride;
        // index tables: staged once per CTA at the end of the matrix area when everything fits (they stay there
        // across the load-flows of this step), else read from global memory
        const int* tables = spd.blob;
        if (dense_need + val_need + blob_dbl <= args.mat_cap) {
            int* dst = reinterpret_cast<int*>(e.mat() + (args.mat_cap - blob_dbl));
            if (e.misc()[3] != which) {
                env_sync<TPE>(mask);
                for (int i = tid; i < stage_words; i += TPE) dst[i] = spd.blob[i];
                if (tid == 0) e.misc()[3] = which;
                env_sync<TPE>(mask);
            }
            tables = dst;
        } else {
            env_sync<TPE>(mask);
            if (tid == 0) e.misc()[3] = -1;   // the dense part may overwrite a staged copy
        }
        SpView spv = sp_view(spd, tables, S);
        if (TPE > 32 && tables != spd.blob && dense_need + val_need <= args.mat_cap) {
            spv.tb = saddr(tables);
            spv.hyb = args.sparse == 3;
        }
        const SpView* sp = &spv;
        if (dense_need <= args.mat_cap) {
            double* M1 = e.mat();
            double* spb = dense_need + val_need <= args.mat_cap ? M1 + dense_need : wsrow + args.ws_dense;
            Border bd;
            if (bordered && spv.hyb) {   // border storage: the env's slice of the workspace (unused by the hybrid plan)
                bd.W[0] = wsrow; bd.W[1] = wsrow + PPN_BORDER_MAX * spd.n;
                bd.Sc[0] = wsrow + 2 * PPN_BORDER_MAX * spd.n; bd.Sc[1] = bd.Sc[0] + PPN_BORDER_MAX * PPN_BORDER_MAX;
                bd.k[0] = bd.k[1] = bd.base[0] = bd.base[1] = 0;
            }
            success = cfg.dc ? dc_solve<TPE, MAXR, D>(e, c, M1, n1, ld1, ref, sp, spb, solve_mode)
                             : ac_solve<TPE, MAXR, D, true>(e, c, cfg, M1, M1 + (n1 + 1) * ld1, n1, n2, ld1, ld2, ref, slot, n_iter, sp, spb, solve_mode,
                                                            (bordered && spv.hyb) ? &bd : nullptr);
        } else {
            double* M1 = wsrow;
            double* spb = wsrow + args.ws_dense;
            success = cfg.dc ? dc_solve<TPE, MAXR, D>(e, c, M1, n1, ld1, ref, sp, spb, false)
                             : ac_solve<TPE, MAXR, D, false>(e, c, cfg, M1, M1 + (n1 + 1) * ld1, n1, n2, ld1, ld2, ref, slot, n_iter, sp, spb, false);
        }
    } else if (n1 * ld1 + n2 * ld2 <= args.mat_cap) {
        double* M1 = e.mat();
        success = cfg.dc ? dc_solve<TPE, MAXR, D>(e, c, M1, n1, ld1, ref, nullptr, nullptr, false)
                         : ac_solve<TPE, MAXR, D, true>(e, c, cfg, M1, M1 + n1 * ld1, n1, n2, ld1, ld2, ref, slot, n_iter, nullptr, nullptr, false);
    } else {
        double* M1 = args.ws + (size_t)slot * args.ws_stride;
        success = cfg.dc ? dc_solve<TPE, MAXR, D>(e, c, M1, n1, ld1, ref, nullptr, nullptr, false)
                         : ac_solve<TPE, MAXR, D, false>(e, c, cfg, M1, M1 + n1 * ld1, n1, n2, ld1, ld2, ref, slot, n_iter, nullptr, nullptr, false);
    }
    PPN_TICK(11);
    // runpf tail: out-of-service generators report Pg = Qg = 0
    for (int g = tid; g < e.G; g += TPE)
        if (!(e.gstat()[g] > 0 && e.btype()[e.gbus()[g]] != PPN_BT_ISOLATED)) { e.gpg()[g] = 0.0; e.gqg()[g] = 0.0; }
    env_sync<TPE>(mask);
    // grid.py:103-110, 263: NaN or > 1e10 in bus Vm/Va, branch flows, bus Pd
    bool bad = false;
    for (int b = tid; b < NB; b += TPE) {
        const double a = e.vm()[b], d = e.va()[b];
        bad |= (a != a) || (d != d) || a > 1e10 || d > 1e10;
    }
    for (int l = tid; l < e.N; l += TPE) {
        const double a = e.pf()[l], b2 = e.qf()[l], c2 = e.pt()[l], d = e.qt()[l];
        bad |= (a != a) || (b2 != b2) || (c2 != c2) || (d != d) || a > 1e10 || b2 > 1e10 || c2 > 1e10 || d > 1e10;
    }
    for (int l = tid; l < e.L; l += TPE) { const double a = e.lpd()[l]; bad |= (a != a) || a > 1e10; }
    const bool any_bad = env_any<TPE>(bad, mask);
    PPN_TICK(12);
#ifdef PPN_TIMING
    if (slot == 0 && tid == 0) ppn_timing_buf[63] = 1;
#endif
    return !success || any_bad;
}

// grid.py:112-138, 29-36
template <int TPE, class D>
__device__ __forceinline__ void flows_ampere(Env<TPE, D>& e, const PpnDevCase& c) {
    for (int l = e.tid; l < e.N; l += TPE) {
        double a = 0.0;
        if (e.status()[l]) {
            const int f = e.fbus()[l];
            const double p = e.pf()[l], q = e.qf()[l];
            a = 1000. * sqrt(p * p + q * q) / (PPN_SQRT3 * (e.vm()[f] * c.bus_basekv[f]));
        }
        e.amp()[l] = a;
    }
    env_sync<TPE>(e.mask);
}

// game.py:503-589.  Returns true on DivergingLoadflowException.
template <int TPE, int MAXR, class D>
__device__ __forceinline__ bool cascade(Env<TPE, D>& e, const PpnDevCase& c, const PpnDevCfg& cfg, const PpnStepArgs& args, int slot,
                        int& n_lf, int& n_it, int& depth_out) {
    int depth = 0;
    for (int l = e.tid; l < e.N; l += TPE) e.over()[l] = 0;
    while (true) {
        int it;
        n_lf++;
        const bool div = loadflow<TPE, MAXR, D>(e, c, cfg, args, slot, it);
        n_it += it;
        if (div) { depth_out = depth; return true; }
        flows_ampere(e, c);
        bool any_over = false, any_trip = false;
        for (int l = e.tid; l < e.N; l += TPE) {
            const bool over = e.amp()[l] > c.thermal[l];
            e.over()[l] = over ? 1 : 0;
            any_over |= over;
        }
        if (!env_any<TPE>(any_over, e.mask)) break;
        for (int l = e.tid; l < e.N; l += TPE) {
            const double a = e.amp()[l], lim = c.thermal[l];
            if (a > cfg.hard_coef * lim) {
                e.status()[l] = 0; e.recon()[l] = cfg.n_hard_broken; any_trip = true; e.over()[l] = 0;
            } else if (e.over()[l] && (double)e.soft()[l] >= cfg.n_soft_consec) {
                e.status()[l] = 0; e.recon()[l] = cfg.n_soft_broken; any_trip = true; e.over()[l] = 0;
            }
        }
        depth++;
        if (!env_any<TPE>(any_trip, e.mask)) break;
        env_sync<TPE>(e.mask);
    }
    env_sync<TPE>(e.mask);
    for (int l = e.tid; l < e.N; l += TPE) e.soft()[l] = e.over()[l] ? e.soft()[l] + 1 : 0;
    env_sync<TPE>(e.mask);
    depth_out = depth;
    return false;
}

// reset_grid (game.py:782-797)
template <int TPE, class D>
__device__ __forceinline__ void reset_grid(Env<TPE, D>& e, const PpnDevCase& c) {
    for (int i = e.tid; i < e.N; i += TPE) {
        e.recon()[i] = 0; e.lreact()[i] = 0;
        e.onode()[i] = 0; e.enode()[i] = 0;
        e.status()[i] = c.line_status0[i];
    }
    for (int i = e.tid; i < e.S; i += TPE) e.nreact()[i] = 0;
    for (int i = e.tid; i < e.G; i += TPE) { e.gnode()[i] = 0; e.gstat()[i] = 1; }
    for (int i = e.tid; i < e.L; i += TPE) e.lnode()[i] = 0;
    for (int b = e.tid; b < e.NB; b += TPE) { e.vm()[b] = c.bus_vm0[b]; e.va()[b] = c.bus_va0[b]; }
    env_sync<TPE>(e.mask);
}

// Dynamic prefix of Observation.as_array (environment.py:451-466, 511-517), 7L+7G+13N+S+6 values.  The row is
// assembled in shared memory (`stage`: the matrix area, free once the cascade is over) and leaves in one linear,
// fully coalesced sweep -- `out` may be device memory or page-locked host memory written over PCIe (ppn_step_host).
// stage == nullptr: the fields are written to `out` directly.
// T = double (the reference's dtype) or float (ppn_step_host_f32: half the bytes over PCIe for host-side agents that feed
// a float32 network anyway; every value is the double one rounded once).
template <int TPE, class D, class T>
__device__ __forceinline__ void write_observation(Env<TPE, D>& e, const PpnDevCase& c, const PpnDevChronics& ch, T* out,
                                                  T* stage, bool bulk, bool wait_written = false) {
    const int G = e.G, L = e.L, N = e.N, S = e.S, tid = e.tid;
    compute_isolated(e);
    const float* row = chronic_row(ch, e.cursor()[0], max(e.cursor()[1], 0));   // maintenance horizon, date
    const float* prow = chronic_row(ch, e.misc()[6], e.misc()[7]);               // planned injections
    T* o = stage ? stage : out;
    for (int l = tid; l < L; l += TPE) {
        o[l] = e.lpd()[l];
        o[L + l] = e.mark()[e.lbus()[l]] ? 1.0 : 0.0;
        o[2 * L + l] = (double)prow[ch.o_lpp + l];
        o[3 * L + l] = (double)e.lnode()[l];
    }
    o += 4 * L;
    for (int g = tid; g < G; g += TPE) {
        o[g] = e.gpg()[g];
        o[G + g] = e.mark()[e.gbus()[g]] ? 1.0 : 0.0;
        o[2 * G + g] = (double)prow[ch.o_ppp + g];
        o[3 * G + g] = (double)e.gnode()[g];
    }
    o += 4 * G;
    const int* pm = reinterpret_cast<const int*>(row + ch.o_pm);
    for (int l = tid; l < N; l += TPE) {
        o[l] = (double)e.onode()[l];
        o[N + l] = (double)e.enode()[l];
        o[2 * N + l] = e.amp()[l];
        o[3 * N + l] = (double)e.status()[l];
        o[4 * N + l] = (double)e.recon()[l];
        o[5 * N + l] = (double)e.lreact()[l];
        o[6 * N + S + l] = (double)pm[l];
    }
    for (int s = tid; s < S; s += TPE) o[6 * N + s] = (double)e.nreact()[s];
    o += 7 * N + S;
    const int* dt = reinterpret_cast<const int*>(row + ch.o_dt);
    if (tid < 6) o[tid] = (double)dt[tid];
    o += 6;
    for (int l = tid; l < L; l += TPE) {
        o[l] = e.lqd()[l];
        o[L + l] = e.vm()[e.lbus()[l]];
        o[2 * L + 2 * G + 6 * N + l] = (double)prow[ch.o_lqp + l];
    }
    o += 2 * L;
    for (int g = tid; g < G; g += TPE) {
        o[g] = e.gqg()[g];
        o[G + g] = e.gvg()[g];
        const float pv = prow[ch.o_pvp + g];
        o[2 * G + 6 * N + L + g] = (double)(pv <= 0.f ? 0.f : pv) / e.gkv()[g];
    }
    o += 2 * G;
    for (int l = tid; l < N; l += TPE) {
        o[l] = e.pf()[l];
        o[N + l] = e.qf()[l];
        o[2 * N + l] = e.vm()[e.fbus()[l]];
        o[3 * N + l] = e.pt()[l];
        o[4 * N + l] = e.qt()[l];
        o[5 * N + l] = e.vm()[e.tbus()[l]];
    }
    if (stage) {
        const int n = 7 * L + 7 * G + 13 * N + S + 6;
        const int nb = n & ~(int)(16 / sizeof(T) - 1);   // bulk copies move multiples of 16 bytes
        if (bulk && (reinterpret_cast<size_t>(out) & 15) == 0) {
            // one TMA bulk store per row (cp.async.bulk shared -> global): full-width write transactions, which is
            // what makes page-locked host memory behind PCIe a usable destination, and no store loop for the warp
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staging writes -> visible to the async proxy
            env_sync<TPE>(e.mask);
            if (tid == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             :: "l"(out), "r"(saddr(stage)), "r"(nb * (int)sizeof(T)) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                for (int i = nb; i < n; i++) out[i] = stage[i];
                if (wait_written) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the row is in memory: a flag follows
                else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");          // the source may be reused / freed
            }
            env_sync<TPE>(e.mask);
        } else {
            env_sync<TPE>(e.mask);
            for (int i = tid; i < n; i += TPE) out[i] = stage[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------------ the kernel
template <int TPE, int MAXR, class D, int MINB>
__global__ void __launch_bounds__(TPE <= 32 ? 64 : TPE, MINB)
ppn_step_kernel(const __grid_constant__ PpnDevCase c, const __grid_constant__ PpnDevChronics ch, const __grid_constant__ PpnDevCfg cfg,
                const __grid_constant__ PpnDevState st, const __grid_constant__ PpnStepArgs args, int env_smem_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int envs_per_block = TPE <= 32 ? (blockDim.x / TPE) : 1;
    const int local = TPE <= 32 ? (threadIdx.x / TPE) : 0;
    const int slot = blockIdx.x * envs_per_block + local;          // output row
    const int n_rows_out = args.n_envs * args.n_cand;
    if (slot >= n_rows_out) return;
    const int env = args.env_off + slot / args.n_cand;   // state row; `slot` is the output row of this launch
    Env<TPE, D> e;
    e.init_dims(c);
    e.tid = TPE <= 32 ? (threadIdx.x & (TPE - 1)) : threadIdx.x;
    e.shift = TPE < 32 ? ((threadIdx.x & 31) & ~(TPE - 1)) : 0;
    e.mask = TPE < 32 ? (((1u << TPE) - 1u) << e.shift) : PPN_FULL;
    {   // the env's shared-memory image: its base address is made opaque to the optimiser, which otherwise re-derives it
        // from %tid / the CTA's shared window / env_smem_bytes at every use (7 instructions, dozens of times per iteration)
        unsigned bs = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)local * (unsigned)env_smem_bytes;
        asm volatile("" : "+r"(bs));
        e.base = reinterpret_cast<unsigned char*>(__cvta_shared_to_generic((size_t)bs));
    }
    e.fixed_bytes = env_smem_bytes - 8 * args.mat_cap;
    const int tid = e.tid, S = e.S, G = e.G, L = e.L, N = e.N, NB = e.NB;
    const unsigned mask = e.mask;
    const int mode = args.mode;
    const bool is_sim = mode == PPN_MODE_SIMULATE;
    if (mode == PPN_MODE_GAME_OVER && args.mask && !args.mask[env]) return;

    if (tid == 0) e.misc()[3] = -1;   // no sparse index tables staged yet
    const long long t_begin = args.trace ? clock64() : 0;
    // ---- state in
    if (mode == PPN_MODE_INIT) {
        for (int b = tid; b < NB; b += TPE) { e.vm()[b] = c.bus_vm0[b]; e.va()[b] = c.bus_va0[b]; }
        for (int l = tid; l < L; l += TPE) { e.lpd()[l] = c.load_pd0[l]; e.lqd()[l] = c.load_qd0[l]; e.lnode()[l] = 0; }
        for (int g = tid; g < G; g += TPE) {
            e.gpg()[g] = c.gen_pg0[g]; e.gqg()[g] = c.gen_qg0[g]; e.gvg()[g] = c.gen_vg0[g]; e.gnode()[g] = 0; e.gstat()[g] = 1;
        }
        for (int l = tid; l < N; l += TPE) {
            e.onode()[l] = 0; e.enode()[l] = 0; e.status()[l] = c.line_status0[l];
            e.recon()[l] = 0; e.lreact()[l] = 0; e.soft()[l] = 0;
        }
        for (int s = tid; s < S; s += TPE) e.nreact()[s] = 0;
        if (tid == 0) {
            const int start = args.init_chronic ? args.init_chronic[env] : 0;
            const int row0 = args.init_row0 ? args.init_row0[env] : 0;
            e.cursor()[2] = start; e.cursor()[3] = 0;
            e.cursor()[0] = take_next_chronic(e.cursor(), cfg, ch.n_chronics, env);
            e.cursor()[1] = row0 - 1;   // -1: no row played yet
        }
    } else {
        const double* sr = st.real + (size_t)env * st.rw;
        for (int b = tid; b < 2 * NB; b += TPE) e.vm()[b] = sr[b];            // vm | va are adjacent
        for (int l = tid; l < 2 * L; l += TPE) e.lpd()[l] = sr[2 * NB + l];   // lpd | lqd adjacent
        for (int g = tid; g < 3 * G; g += TPE) e.gpg()[g] = sr[2 * NB + 2 * L + g];  // gpg | gqg | gvg adjacent
        const uint8_t* tr = st.topo + (size_t)env * st.tw;
        for (int i = tid; i < 2 * G + L + 3 * N; i += TPE) e.gnode()[i] = tr[i];
        const int32_t* cr = st.cnt + (size_t)env * st.cw;
        for (int i = tid; i < 3 * N + S + 4; i += TPE) e.recon()[i] = cr[i];
    }
    env_sync<TPE>(mask);

    int flag = 0;
    bool done = false, illegal = false, too_much = false;
    int n_lf = 0, n_it = 0, depth = 0, n_resets = 0;
    int n_ill = 0, cost_nodes = 0, cost_lines = 0;

    if (mode == PPN_MODE_STEP || mode == PPN_MODE_SIMULATE) {
        // ---- action: legality (game.py:650-753), correction (game.py:809-854), application (game.py:591-648)
        const int NT = G + L + 2 * N;  // node-switch part
        bool any_set = false;
        if (args.act) {
            const uint8_t* a = args.act + (size_t)slot * e.A;
            for (int i = tid; i < e.A; i += TPE) { const uint8_t v = a[i]; e.act()[i] = v; any_set |= v != 0; }
        }
        any_set = env_any<TPE>(any_set, mask);
        if (any_set) {   // a do-nothing action changes nothing and is always legal
            for (int s = tid; s < S; s += TPE) e.subch()[s] = 0;
            env_sync<TPE>(mask);
            for (int i = tid; i < NT; i += TPE)
                if (e.act()[i]) e.subch()[c.elem_sub[i]] = 1;
            env_sync<TPE>(mask);
            int ns = 0, nl = 0;
            for (int s = tid; s < S; s += TPE) ns += e.subch()[s];
            for (int l = tid; l < N; l += TPE) nl += (e.act()[NT + l] == 1);
            ns = env_sum_int<TPE>(ns, e.redi(), tid, mask);
            nl = env_sum_int<TPE>(nl, e.redi() + (TPE + 31) / 32, tid, mask);
            too_much = ns > cfg.max_sub || nl > cfg.max_lines || ns + nl > cfg.max_total;
            for (int i = tid; i < 1 + 2 * N + S; i += TPE) e.ill()[i] = 0;
            env_sync<TPE>(mask);
            if (too_much) {
                illegal = true;
                if (tid == 0) e.ill()[0] = 1;
                for (int i = tid; i < e.A; i += TPE) e.act()[i] = 0;
                for (int s = tid; s < S; s += TPE) e.subch()[s] = 0;
            } else {
                int cnt = 0;
                for (int l = tid; l < N; l += TPE) {
                    const bool sw = e.act()[NT + l] == 1;
                    const bool r1 = sw && e.recon()[l] > 0, r2 = sw && e.lreact()[l] > 0;
                    e.ill()[1 + l] = r1; e.ill()[1 + N + l] = r2;
                    cnt += r1 + r2;
                    if (r1 || r2) e.act()[NT + l] = 0;
                }
                for (int s = tid; s < S; s += TPE) {
                    const bool r3 = e.subch()[s] && e.nreact()[s] > 0;
                    e.ill()[1 + 2 * N + s] = r3;
                    cnt += r3;
                }
                n_ill = env_sum_int<TPE>(cnt, e.redi(), tid, mask);
                illegal = n_ill > 0;
                env_sync<TPE>(mask);
                for (int i = tid; i < NT; i += TPE)
                    if (e.ill()[1 + 2 * N + c.elem_sub[i]]) e.act()[i] = 0;
                for (int s = tid; s < S; s += TPE)
                    if (e.ill()[1 + 2 * N + s]) e.subch()[s] = 0;
            }
            env_sync<TPE>(mask);
            // apply: new = where(bit, 1 - cur, cur); loads follow their node bit (the Pd/Qd swap of grid.py:405-421)
            int cn = 0, cl = 0;
            for (int i = tid; i < NT; i += TPE) {
                if (e.act()[i]) { e.gnode()[i] ^= 1; cn += e.act()[i]; }   // gnode|lnode|onode|enode are adjacent
            }
            for (int l = tid; l < N; l += TPE) {
                const int a = e.act()[NT + l];
                if (a) { e.status()[l] ^= 1; cl += a; }
                if (a == 1) e.lreact()[l] = cfg.n_line_react;
            }
            for (int s = tid; s < S; s += TPE)
                if (e.subch()[s]) e.nreact()[s] = cfg.n_node_react;
            cost_nodes = env_sum_int<TPE>(cn, e.redi(), tid, mask);
            if (cost_nodes > 0 && tid == 0 && args.split_flag && !is_sim) *args.split_flag = 1;
            cost_lines = env_sum_int<TPE>(cl, e.redi() + (TPE + 31) / 32, tid, mask);
            env_sync<TPE>(mask);
        } else if (args.illegal) {
            for (int i = tid; i < 1 + 2 * N + S; i += TPE) e.ill()[i] = 0;
            env_sync<TPE>(mask);
        }
    }

    // One loop, two kinds of pass: the step itself (not for PPN_MODE_GAME_OVER), then process_game_over passes
    // (game.py:762-780: reset, next row -- next chronic in hard mode --, cascade; again while it diverges).
    bool reset_pass = mode == PPN_MODE_GAME_OVER;
    int attempts = 0;
    while (true) {
        if (reset_pass) {
            reset_grid(e, c);
            if (cfg.hard_mode) {
                if (tid == 0) {
                    e.cursor()[0] = take_next_chronic(e.cursor(), cfg, ch.n_chronics, env);
                    e.cursor()[1] = -2;
                }
                env_sync<TPE>(mask);
            }
            n_resets++;
        }
        refresh_element_buses(e, c);
        load_next_timestep(e, c, ch, cfg, is_sim && !reset_pass, env);
        int d2 = 0;
        const bool div = cascade<TPE, MAXR, D>(e, c, cfg, args, args.env_off * args.n_cand + slot, n_lf, n_it, d2);   // global row: spill slice
        if (reset_pass) {
            if (!div || ++attempts >= cfg.max_reset_attempts) { done = false; break; }
            continue;
        }
        depth = d2;
        int nlc = 0, npc = 0;
        if (div) {
            flag = 2; done = true;
        } else if (mode != PPN_MODE_INIT) {   // Game.__init__ (game.py:339-340) only runs the cascade
            compute_isolated(e);
            for (int l = tid; l < L; l += TPE) nlc += e.mark()[e.lbus()[l]];
            for (int g = tid; g < G; g += TPE) npc += e.mark()[e.gbus()[g]];
            nlc = env_sum_int<TPE>(nlc, e.redi(), tid, mask);
            npc = env_sum_int<TPE>(npc, e.redi() + (TPE + 31) / 32, tid, mask);
            if (nlc > cfg.max_loads_go) { flag = 3; done = true; }
            else if (npc > cfg.max_prods_go) { flag = 4; done = true; }
        }
        if (flag == 0 && illegal) flag = 1;
        // ---- reward (parameters/default14/reward_signal.py:45-169)
        if (args.reward && mode != PPN_MODE_INIT) {
            const double k = cfg.reward_k;
            const double cost = -.1 * (double)cost_nodes + -.2 * (double)cost_lines;
            double r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
            if (flag == 2) { r2 = cost; r3 = -k; }
            else if (flag == 4) { r1 = -k; }
            else if (flag == 3) { r0 = -k; }
            else {
                int dist = 0;
                for (int i = tid; i < G + L + 2 * N; i += TPE) dist += e.gnode()[i];
                dist = env_sum_int<TPE>(dist, e.redi(), tid, mask);
                double u = 0.0;
                for (int l = tid; l < N; l += TPE) { const double x = e.amp()[l] / c.thermal[l]; u += x * x; }
                u = env_sum_double<TPE>(u, e.redd(), tid, mask);
                r0 = -k / 5. * (double)nlc;
                r1 = -k / 10. * (double)npc;
                r2 = cost;
                r3 = -.02 * (double)dist;
                r4 = -1. * u;
                if (flag == 1) {
                    if (too_much) r2 += -5 * k;
                    else r2 += (-k / 100.) * (double)n_ill;
                }
            }
            if (tid == 0) {
                double* r = args.reward + (size_t)slot * 5;
                r[0] = r0; r[1] = r1; r[2] = r2; r[3] = r3; r[4] = r4;
                if (args.pack) {
                    double* q = args.pack + (size_t)slot * 7;
                    q[0] = r0; q[1] = r1; q[2] = r2; q[3] = r3; q[4] = r4;
                }
            }
        }
        if (tid == 0) {
            if (args.done) args.done[slot] = done ? 1 : 0;
            if (args.flag) args.flag[slot] = flag;
            if (args.pack) { args.pack[(size_t)slot * 7 + 5] = done ? 1.0 : 0.0; args.pack[(size_t)slot * 7 + 6] = (double)flag; }
        }
        if (args.illegal && mode != PPN_MODE_INIT) {
            uint8_t* il = args.illegal + (size_t)slot * (1 + 2 * N + S);
            for (int i = tid; i < 1 + 2 * N + S; i += TPE) il[i] = e.ill()[i];
        }
        if (!(done && !is_sim && (args.auto_reset || mode == PPN_MODE_INIT))) break;
        reset_pass = true;
    }

    // ---- outputs
    if (args.obs && !done) {
        flows_ampere(e, c);
        if (args.obs_f32)
            write_observation<TPE, D, float>(e, c, ch, reinterpret_cast<float*>(args.obs) + (size_t)slot * args.obs_stride,
                                             args.mat_cap >= c.OBSD ? reinterpret_cast<float*>(e.mat()) : nullptr, args.obs_bulk != 0,
                                             args.row_flag != nullptr);
        else
            write_observation<TPE, D, double>(e, c, ch, args.obs + (size_t)slot * args.obs_stride,
                                              args.mat_cap >= c.OBSD ? e.mat() : nullptr, args.obs_bulk != 0, args.row_flag != nullptr);
    }
    if (args.row_flag) {
        // the row (written through the async proxy by the bulk store, or by plain stores) is complete: publish it
        env_sync<TPE>(mask);
        if (tid == 0) {
            asm volatile("fence.proxy.async;" ::: "memory");
            __threadfence();
            const unsigned v = args.epoch | ((args.obs && !done) ? 0u : 0x80000000u);
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(args.row_flag + slot), "r"(v) : "memory");
        }
    }
    if (!is_sim) {
        env_sync<TPE>(mask);
        double* sr = st.real + (size_t)env * st.rw;
        for (int b = tid; b < 2 * NB; b += TPE) sr[b] = e.vm()[b];
        for (int l = tid; l < 2 * L; l += TPE) sr[2 * NB + l] = e.lpd()[l];
        for (int g = tid; g < 3 * G; g += TPE) sr[2 * NB + 2 * L + g] = e.gpg()[g];
        uint8_t* tr = st.topo + (size_t)env * st.tw;
        for (int i = tid; i < 2 * G + L + 3 * N; i += TPE) tr[i] = e.gnode()[i];
        int32_t* cr = st.cnt + (size_t)env * st.cw;
        for (int i = tid; i < 3 * N + S + 4; i += TPE) cr[i] = e.recon()[i];
    }
    if (args.trace && tid == 0) {
        long long* tr4 = args.trace + 4 * (size_t)slot;
        tr4[0] = clock64() - t_begin; tr4[1] = n_lf; tr4[2] = n_it; tr4[3] = n_resets;
    }
    if (args.stats && tid == 0) {
        atomicAdd(args.stats + 0, (unsigned long long)n_lf);
        atomicAdd(args.stats + 1, (unsigned long long)n_it);
        atomicAdd(args.stats + 2, 1ull);
        atomicAdd(args.stats + 3, (unsigned long long)n_resets);
        atomicMax(args.stats + 4, (unsigned long long)depth);
        atomicMax(args.stats + 5, (unsigned long long)n_lf);
        atomicMax(args.stats + 6, (unsigned long long)n_it);
        if (mode == PPN_MODE_STEP) atomicAdd(args.stats + 8 + (depth < 7 ? depth : 7), 1ull);   // cascade-depth histogram
    }
}

template <int TPE, int MAXR, class D, int MINB>
int launch_group(const PpnDevCase* c, const PpnDevChronics* ch, const PpnDevCfg* cfg, const PpnDevState* st,
                 const PpnStepArgs* args, int envs_per_block, int env_smem_bytes, cudaStream_t stream) {
    const int rows = args->n_envs * args->n_cand;
    const int epb = TPE <= 32 ? envs_per_block : 1;
    const int grid = (rows + epb - 1) / epb;
    const size_t smem = (size_t)epb * env_smem_bytes;
    auto k = ppn_step_kernel<TPE, MAXR, D, MINB>;
    // opt in to large dynamic shared memory once per (kernel, device, size): the call costs more than a launch
    static int configured_smem[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || configured_smem[dev] < (int)smem) {
        cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
        if (dev >= 0 && dev < 64) configured_smem[dev] = (int)smem;
    }
    k<<<grid, TPE <= 32 ? epb * TPE : TPE, smem, stream>>>(*c, *ch, *cfg, *st, *args, env_smem_bytes);
    return (int)cudaGetLastError();
}

typedef StaticDims<14, 5, 11, 20> Dims14;      // IEEE-14  (parameters/default14)
typedef StaticDims<30, 6, 20, 41> Dims30;      // IEEE-30  (parameters/default30)
typedef StaticDims<118, 54, 99, 186> Dims118;  // IEEE-118 (parameters/default118)

template <class SD> bool dims_match(const PpnDevCase* c) {
    return c->S == SD::S && c->G == SD::G && c->L == SD::L && c->N == SD::N;
}

}  // namespace

#ifdef PPN_TIMING
extern "C" int ppn_debug_timing(long long* out_host, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out_host, ppn_timing_buf, sizeof(long long) * 64);
    if (reset) { long long z[64] = {0}; cudaMemcpyToSymbol(ppn_timing_buf, z, sizeof(z)); }
    return (int)e;
}
#endif

// -------------------------------------------------------------------------------------------------- launch wrapper
// tpe: 16 (<= 16 substations), 32 (<= 32 substations), 128 or 256 (one CTA per env, <= 128 substations).  The three
// IEEE families run kernels specialised on their sizes; any other grid runs the size-generic instantiation.
// The file can be compiled in parts so that the build runs in parallel (__graft_entry__.build): -DPPN_PART=0 holds the
// warp-per-env kernels and the dispatcher, 1 the 128-thread CTA kernels, 2 the 256-thread ones; without PPN_PART
// everything lands in one translation unit.
extern "C" int ppn_launch_step_cta128(const PpnDevCase* c, const PpnDevChronics* ch, const PpnDevCfg* cfg, const PpnDevState* st,
                                      const PpnStepArgs* args, int env_smem_bytes, cudaStream_t stream);
extern "C" int ppn_launch_step_cta256(const PpnDevCase* c, const PpnDevChronics* ch, const PpnDevCfg* cfg, const PpnDevState* st,
                                      const PpnStepArgs* args, int env_smem_bytes, cudaStream_t stream);

#if !defined(PPN_PART) || PPN_PART == 0
extern "C" int ppn_launch_step(const PpnDevCase* c, const PpnDevChronics* ch, const PpnDevCfg* cfg, const PpnDevState* st,
                               const PpnStepArgs* args, int tpe, int envs_per_block, int env_smem_bytes,
                               cudaStream_t stream) {
    if (args->n_envs * args->n_cand <= 0) return 0;
    switch (tpe) {
        case 16:   // two envs per warp: measured slower than a warp per env, kept size-generic only
            return launch_group<16, 2, DynDims, 1>(c, ch, cfg, st, args, envs_per_block, env_smem_bytes, stream);
        case 32:
            // IEEE-14: 96 registers, ten 64-thread CTAs per SM.  Measured alternatives (round 2: 128 registers, 8 CTAs:
            // +1.6 % at 4096 envs, -1 % at 65536 -- not kept); round 1: 80 registers (12 CTAs) 9.9 M
            // env-steps/s, 72 registers (14 CTAs, 4096 envs in one wave) 9.2 M, against 10.9 M -- the spills cost more
            // than the second wave
            if (dims_match<Dims14>(c)) return launch_group<32, 2, Dims14, 10>(c, ch, cfg, st, args, envs_per_block, env_smem_bytes, stream);
            if (dims_match<Dims30>(c)) return launch_group<32, 2, Dims30, 5>(c, ch, cfg, st, args, envs_per_block, env_smem_bytes, stream);
            return launch_group<32, 2, DynDims, 1>(c, ch, cfg, st, args, envs_per_block, env_smem_bytes, stream);
        case 128: return ppn_launch_step_cta128(c, ch, cfg, st, args, env_smem_bytes, stream);
        case 256: return ppn_launch_step_cta256(c, ch, cfg, st, args, env_smem_bytes, stream);
        default: return (int)cudaErrorInvalidValue;
    }
}
#endif

#if !defined(PPN_PART) || PPN_PART == 1
// half-size CTAs for the CTA-per-env grids: two buses per thread, three CTAs per SM (PPN_MINB=2: two, 255 registers)
extern "C" int ppn_launch_step_cta128(const PpnDevCase* c, const PpnDevChronics* ch, const PpnDevCfg* cfg, const PpnDevState* st,
                                      const PpnStepArgs* args, int env_smem_bytes, cudaStream_t stream) {
    if (dims_match<Dims118>(c)) {
        static const int minb128 = getenv("PPN_MINB") ? atoi(getenv("PPN_MINB")) : 3;
        if (minb128 == 2) return launch_group<128, 8, Dims118, 2>(c, ch, cfg, st, args, 1, env_smem_bytes, stream);
        return launch_group<128, 8, Dims118, 3>(c, ch, cfg, st, args, 1, env_smem_bytes, stream);
    }
    return launch_group<128, 8, DynDims, 1>(c, ch, cfg, st, args, 1, env_smem_bytes, stream);
}
#endif

#if !defined(PPN_PART) || PPN_PART == 2
// 256-thread CTAs: the explicit-inverse plan (whole SM per CTA) and, with PPN_TPE=256, the hybrid plan at two CTAs per SM
extern "C" int ppn_launch_step_cta256(const PpnDevCase* c, const PpnDevChronics* ch, const PpnDevCfg* cfg, const PpnDevState* st,
                                      const PpnStepArgs* args, int env_smem_bytes, cudaStream_t stream) {
    if (dims_match<Dims118>(c)) {
        if (args->sparse >= 2) return launch_group<256, 8, Dims118, 2>(c, ch, cfg, st, args, 1, env_smem_bytes, stream);
        return launch_group<256, 8, Dims118, 1>(c, ch, cfg, st, args, 1, env_smem_bytes, stream);
    }
    return launch_group<256, 8, DynDims, 1>(c, ch, cfg, st, args, 1, env_smem_bytes, stream);
}
#endif
