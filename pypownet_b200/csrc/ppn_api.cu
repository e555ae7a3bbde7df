// C ABI of the batched pypownet step path (include/pypownet_b200.h): handle management, upload of the grid family and
// of the chronic tables, and the entry points that enqueue the fused step kernel (ppn_kernels.cu).
// Everything numerical happens on the device; there is no CPU fallback behind these calls.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>
#include <complex>
#include <algorithm>

#include "../../include/pypownet_b200.h"
#include "ppn_device.cuh"

extern "C" int ppn_launch_step(const PpnDevCase* c, const PpnDevChronics* ch, const PpnDevCfg* cfg, const PpnDevState* st,
                               const PpnStepArgs* args, int tpe, int envs_per_block, int env_smem_bytes,
                               cudaStream_t stream);

struct ppn_env {
    int device = 0;
    int B = 0;
    int S = 0, G = 0, L = 0, N = 0, NB = 0, A = 0, OBS = 0, OBSD = 0;
    PpnDevCase dc{};
    PpnDevChronics dch{};
    PpnDevCfg dcfg{};
    PpnDevState st{};
    bool chronics_loaded = false, initialised = false;
    bool async_pending = false;   // work was enqueued on a caller stream since the last host-buffer call
    int tpe = 32, envs_per_block = 4, env_smem_bytes = 0, mat_cap = 0;
    double* ws = nullptr;
    long long ws_stride = 0, ws_rows = 0;
    int horizon = 20;
    int sp_need[2] = {0, 0};   // doubles of factor storage (both matrices) of the U / F sparse structures
    int sp_blob_dbl[2] = {0, 0};   // doubles of their index tables
    int sp_hyb_dbl[2] = {0, 0};    // ... of the prefix the hybrid solver stages
    int sparse = 0;            // solver mode of PpnStepArgs.sparse
    int alt_sparse = 0, alt_mat_cap = 0, alt_env_smem_bytes = 0, alt_tpe = 0;   // plan used once buses may be split
    double* pack_dev = nullptr;   // optional packed result rows written by ppn_step (ppn_set_result_pack)
    long long* trace_dev = nullptr;   // optional per-env trace rows written by ppn_step (ppn_set_env_trace)
    int* h_split = nullptr;    // page-locked, mapped: the kernel sets it when an env applies a node switch
    int* d_split = nullptr;    // its device alias
    long long ws_dense = 0;
    unsigned long long* stats = nullptr;
    long long launches = 0;
    std::vector<void*> allocs;
    std::vector<double> obs_static;
    std::vector<int> chronic_rows;
    // pinned staging for the host-buffer entry point
    uint8_t* h_act = nullptr; uint8_t* d_act = nullptr;
    double* h_obs = nullptr; double* d_obs = nullptr;
    double* h_reward = nullptr; double* d_reward = nullptr;
    uint8_t* h_done = nullptr; uint8_t* d_done = nullptr;
    int32_t* h_flag = nullptr; int32_t* d_flag = nullptr;
    uint8_t* h_ill = nullptr; uint8_t* d_ill = nullptr;
    int32_t* d_init = nullptr;   // [2B] chronic idx | row0
    cudaStream_t own_stream = nullptr;
    cudaStream_t drain_stream = nullptr;   // ppn_step_host: the kernel that moves finished rows to the host (see step_host_impl)
    unsigned* d_row_flag = nullptr; int* d_drain_err = nullptr; unsigned epoch = 0;
    // host-buffer entry point: the batch is cut into chunks, each with its own stream, so that the copies of one chunk
    // overlap the kernels of the others
    static const int MAX_CHUNKS = 16;
    int n_chunks = 0;
    cudaStream_t chunk_stream[MAX_CHUNKS] = {};
    std::string err;
};

static thread_local std::string g_err;

static int fail(ppn_env* e, int code, const std::string& msg) {
    if (e) e->err = msg;
    g_err = msg;
    return code;
}

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t _e = (call);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return fail(env, PPN_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));        \
    } while (0)

template <typename T> static T* upload(ppn_env* env, const std::vector<T>& v, cudaError_t* err) {
    T* d = nullptr;
    size_t bytes = (v.size() ? v.size() : 1) * sizeof(T);
    *err = cudaMalloc(&d, bytes);
    if (*err != cudaSuccess) return nullptr;
    env->allocs.push_back(d);
    if (v.size()) *err = cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}

#define UP(dst, vec)                                                                                 \
    do {                                                                                             \
        cudaError_t _e;                                                                              \
        dst = upload(env, vec, &_e);                                                                 \
        if (_e != cudaSuccess) return fail(env, PPN_E_CUDA, std::string("upload " #vec ": ") + cudaGetErrorString(_e)); \
    } while (0)

// ---------------------------------------------------------------------------------------- sparse symbolic analysis
// Minimum-degree order of the substation graph, symbolic LDL^T of the bus-level pattern (U: one row per substation,
// F: rows 2p, 2p+1 = bus s, s+S of the p-th substation), elimination-tree levels and the static update lists of the
// left-looking numeric factorisation run by the kernel (ppn_kernels.cu: sp_factor / sp_invert).
struct SparseHost {
    int n = 0, nnz = 0, n_lev = 0;
    std::vector<short> bus_row, rowidx, ecol, parent, row_lev, rowoff;
    std::vector<int> rpack, rowpk, colpk;
    std::vector<int> colptr, lev_ptr, lev_ent, trip_ptr, trip, line_pos, rowptr, rowent, lev_rows_ptr, lev_rows;
    std::vector<int> wsched;   // one-warp schedule of the hybrid solve (PpnDevSparse.o_wsched); empty: not available
};

// One-warp schedule of the hybrid solve for the cut at level `cut`: see PpnDevSparse.o_wsched.
static void build_warp_schedule(SparseHost& h, int cut) {
    h.wsched.clear();
    const int n = h.n, r0 = h.lev_rows_ptr[cut], nt = n - r0;
    if (h.nnz >= 511 || n > 128 || nt > 32 || cut < 1) return;   // 16-bit entries: 9 bits of entry id, 7 of row / column
    struct Step { int row0, rows, maxcnt; std::vector<std::vector<int>> ent; };   // ent[lane] = packed entries
    std::vector<Step> fwd, bwd;
    auto add_rows = [&](std::vector<Step>& dst, int a, int b, bool forward, int col_limit) {
        for (int base = a; base < b; base += 32) {
            Step st; st.row0 = base; st.rows = std::min(32, b - base); st.maxcnt = 0; st.ent.resize(32);
            for (int l = 0; l < st.rows; l++) {
                const int i = base + l;
                if (forward) {
                    for (int t = h.rowptr[i]; t < h.rowptr[i + 1]; t++) {
                        const int e = h.rowent[t], col = h.ecol[e];
                        if (col < col_limit) st.ent[l].push_back((e << 7) | col);
                    }
                } else {
                    for (int e = h.colptr[i]; e < h.colptr[i + 1]; e++) st.ent[l].push_back((e << 7) | h.rowidx[e]);
                }
                st.maxcnt = std::max(st.maxcnt, (int)st.ent[l].size());
            }
            dst.push_back(st);
        }
    };
    for (int lv = 1; lv < cut; lv++) add_rows(fwd, h.lev_rows_ptr[lv], h.lev_rows_ptr[lv + 1], true, n);
    add_rows(fwd, r0, n, true, r0);                                     // top block: y2 = w2 - L21 y1
    for (int lv = cut - 1; lv >= 0; lv--) add_rows(bwd, h.lev_rows_ptr[lv], h.lev_rows_ptr[lv + 1], false, 0);
    std::vector<Step> all(fwd);
    all.insert(all.end(), bwd.begin(), bwd.end());
    std::vector<unsigned short> ent;
    std::vector<int> hdr;
    for (const Step& st : all) {
        hdr.push_back(st.row0 | (st.rows << 16));
        hdr.push_back(st.maxcnt | ((int)ent.size() << 8));
        for (int q = 0; q < st.maxcnt; q++)
            for (int l = 0; l < 32; l++) ent.push_back(q < (int)st.ent[l].size() ? (unsigned short)st.ent[l][q] : (unsigned short)0xffff);
    }
    if (getenv("PPN_DEBUG_SCHED")) {
        fprintf(stderr, "[ppn] one-warp solve schedule: n %d, top block %d rows, %zu forward + %zu backward steps\n", n, nt, fwd.size(), bwd.size());
        for (const Step& st : all) {
            int tot = 0;
            for (int l = 0; l < 32; l++) tot += (int)st.ent[l].size();
            fprintf(stderr, "[ppn]   rows %3d..%3d  longest row %2d entries, %3d entries in all\n", st.row0, st.row0 + st.rows - 1, st.maxcnt, tot);
        }
    }
    h.wsched.push_back((int)fwd.size());
    h.wsched.push_back((int)bwd.size());
    h.wsched.insert(h.wsched.end(), hdr.begin(), hdr.end());
    const size_t w0 = h.wsched.size();
    h.wsched.resize(w0 + (ent.size() + 1) / 2, 0);
    memcpy(h.wsched.data() + w0, ent.data(), ent.size() * sizeof(unsigned short));
}

static std::vector<int> min_degree_order(int S, int N, const int* lor, const int* lex) {
    std::vector<char> a((size_t)S * S, 0), alive(S, 1);
    for (int l = 0; l < N; l++) if (lor[l] != lex[l]) { a[(size_t)lor[l] * S + lex[l]] = 1; a[(size_t)lex[l] * S + lor[l]] = 1; }
    std::vector<int> perm;
    for (int it = 0; it < S; it++) {
        int best = -1, bd = 1 << 30;
        for (int s = 0; s < S; s++) {
            if (!alive[s]) continue;
            int d = 0;
            for (int t = 0; t < S; t++) d += a[(size_t)s * S + t];
            if (d < bd) { bd = d; best = s; }
        }
        perm.push_back(best);
        alive[best] = 0;
        std::vector<int> nb;
        for (int t = 0; t < S; t++) if (a[(size_t)best * S + t] && alive[t]) nb.push_back(t);
        for (int x : nb) for (int y : nb) if (x != y) a[(size_t)x * S + y] = 1;
        for (int t = 0; t < S; t++) { a[(size_t)best * S + t] = 0; a[(size_t)t * S + best] = 0; }
    }
    return perm;
}

// `order`: bus of each row (elimination order).  Fills every table of SparseHost; `row_level` gets the level of each row.
static void build_sparse_ordered(int S, int N, const int* lor, const int* lex, const std::vector<int>& order, bool full, SparseHost& h,
                                 std::vector<int>& row_level) {
    const int NB = 2 * S, n = (int)order.size();
    h = SparseHost();
    h.n = n;
    h.bus_row.assign(NB, -1);
    for (int p = 0; p < n; p++) h.bus_row[order[p]] = (short)p;
    std::vector<char> a((size_t)n * n, 0);   // lower triangle: a[i*n+j], i > j
    auto couple = [&](int x, int y) {
        const int i = h.bus_row[x], j = h.bus_row[y];
        if (i < 0 || j < 0 || i == j) return;
        a[(size_t)(i > j ? i : j) * n + (i > j ? j : i)] = 1;
    };
    for (int l = 0; l < N; l++)
        for (int on = 0; on < (full ? 2 : 1); on++)
            for (int en = 0; en < (full ? 2 : 1); en++) couple(lor[l] + S * on, lex[l] + S * en);
    h.parent.assign(n, -1);
    for (int k = 0; k < n; k++) {
        std::vector<int> st;
        for (int i = k + 1; i < n; i++) if (a[(size_t)i * n + k]) st.push_back(i);
        if (!st.empty()) h.parent[k] = (short)st[0];
        for (size_t x = 0; x < st.size(); x++) for (size_t y = 0; y < x; y++) a[(size_t)st[x] * n + st[y]] = 1;
    }
    std::vector<int> pos((size_t)n * n, -1);
    h.colptr.assign(n + 1, 0);
    for (int k = 0; k < n; k++) {
        h.colptr[k] = (int)h.rowidx.size();
        for (int i = k + 1; i < n; i++)
            if (a[(size_t)i * n + k]) { pos[(size_t)i * n + k] = (int)h.rowidx.size(); h.rowidx.push_back((short)i); h.ecol.push_back((short)k); }
    }
    h.colptr[n] = (int)h.rowidx.size();
    h.nnz = (int)h.rowidx.size();
    std::vector<int> lev(n, 0);
    for (int k = 0; k < n; k++) if (h.parent[k] >= 0 && lev[h.parent[k]] < lev[k] + 1) lev[h.parent[k]] = lev[k] + 1;
    h.n_lev = 0;
    for (int k = 0; k < n; k++) if (lev[k] + 1 > h.n_lev) h.n_lev = lev[k] + 1;
    h.lev_ptr.assign(h.n_lev + 1, 0);
    for (int v = 0; v < h.n_lev; v++) {
        h.lev_ptr[v] = (int)h.lev_ent.size();
        for (int k = 0; k < n; k++) {
            if (lev[k] != v) continue;
            h.lev_ent.push_back(h.nnz + k);
            for (int e = h.colptr[k]; e < h.colptr[k + 1]; e++) h.lev_ent.push_back(e);
        }
    }
    h.lev_ptr[h.n_lev] = (int)h.lev_ent.size();
    // rows by level and the row-wise view of L (forward substitution gathers along rows)
    h.lev_rows_ptr.assign(h.n_lev + 1, 0);
    for (int v = 0; v < h.n_lev; v++) {
        h.lev_rows_ptr[v] = (int)h.lev_rows.size();
        for (int k = 0; k < n; k++) if (lev[k] == v) h.lev_rows.push_back(k);
    }
    h.lev_rows_ptr[h.n_lev] = (int)h.lev_rows.size();
    h.rowptr.assign(n + 1, 0);
    for (int i = 0; i < n; i++) {
        h.rowptr[i] = (int)h.rowent.size();
        for (int k = 0; k < i; k++) if (pos[(size_t)i * n + k] >= 0) h.rowent.push_back(pos[(size_t)i * n + k]);
    }
    h.rowptr[n] = (int)h.rowent.size();
    for (int en : h.rowent) h.rpack.push_back(((8 * en) << 16) | (8 * (int)h.ecol[en]));
    for (int i = 0; i < n; i++) {
        h.rowpk.push_back((h.rowptr[i] << 8) | (h.rowptr[i + 1] - h.rowptr[i]));
        h.colpk.push_back((h.colptr[i] << 8) | (h.colptr[i + 1] - h.colptr[i]));
    }
    for (int en = 0; en < h.nnz; en++) h.rowoff.push_back((short)(8 * h.rowidx[en]));
    for (int k = 0; k < n; k++) h.row_lev.push_back((short)lev[k]);
    row_level = lev;
    // update terms: target (i,j), i >= j, gets -T(i,k) L(j,k) for every k < j with both entries present
    h.trip_ptr.assign(h.nnz + n + 1, 0);
    auto terms = [&](int i, int j) {
        for (int k = 0; k < j; k++) {
            const int p1 = pos[(size_t)i * n + k], p2 = pos[(size_t)j * n + k];
            if (p1 >= 0 && p2 >= 0) h.trip.push_back((p1 << 16) | p2);
        }
    };
    for (int e = 0; e < h.nnz; e++) { h.trip_ptr[e] = (int)h.trip.size(); terms(h.rowidx[e], h.ecol[e]); }
    for (int k = 0; k < n; k++) { h.trip_ptr[h.nnz + k] = (int)h.trip.size(); terms(k, k); }
    h.trip_ptr[h.nnz + n] = (int)h.trip.size();
    for (int l = 0; l < N; l++)
        for (int on = 0; on < (full ? 2 : 1); on++)
            for (int en = 0; en < (full ? 2 : 1); en++) {
                const int i = h.bus_row[lor[l] + S * on], j = h.bus_row[lex[l] + S * en];
                h.line_pos.push_back(i == j ? -1 : pos[(size_t)(i > j ? i : j) * n + (i > j ? j : i)]);
            }
}

// Minimum-degree order first (F: the two buses of a substation next to each other), then the rows are re-sorted by
// their level in the elimination tree -- still a valid elimination order with the same fill -- so that the rows of
// one level are contiguous: the level-scheduled solves of the kernel walk plain ranges of rows.
static void build_sparse(int S, int N, const int* lor, const int* lex, const std::vector<int>& perm, bool full, SparseHost& h) {
    std::vector<int> order, lev;
    for (int p = 0; p < S; p++) { order.push_back(perm[p]); if (full) order.push_back(perm[p] + S); }
    build_sparse_ordered(S, N, lor, lex, order, full, h, lev);
    std::vector<int> idx(order.size());
    for (size_t i = 0; i < idx.size(); i++) idx[i] = (int)i;
    std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return lev[x] < lev[y]; });
    std::vector<int> order2(order.size());
    for (size_t i = 0; i < idx.size(); i++) order2[i] = order[idx[i]];
    build_sparse_ordered(S, N, lor, lex, order2, full, h, lev);
}

// hybrid cut: the lowest level from which at most cut_rows rows remain (at least one level stays sparse when there is one)
static int choose_cut(const SparseHost& h, int cut_rows = 24) {
    // cut_rows: measured on B200 for IEEE-118 -- 256-thread CTAs, two per SM: 12 / 20 / 28 / 40 rows give 1.60 / 1.62 /
    // 1.72 / 1.67 M env-steps/s; 128-thread CTAs, three per SM (the default): 12 / 20 rows gave 1.99 / 2.14 M in round 1.
    // With the warp-synchronous inverse of blocks of <= 24 rows (gj24_rows_in_registers) the optimum moved up: final
    // round-2 kernels, a top block of 17 / 22 / 28 rows: 3.39 / 3.19 / 4.36 ms per 8 192-env step (28 rows: tiled
    // inverse, and no longer three CTAs per SM).  At most 40: hyb_invert2 keeps a 5 x 5 tile per thread on an 8 x 8 grid.
    int cut = h.n_lev > 1 ? 1 : 0;
    if (const char* v = getenv("PPN_CUT_ROWS")) { cut_rows = atoi(v); if (cut_rows > 40) cut_rows = 40; if (cut_rows < 1) cut_rows = 1; }
    while (cut < h.n_lev - 1 && h.n - h.lev_rows_ptr[cut] > cut_rows) cut++;
    return cut;
}

// Diagnostic (host only, no GPU): builds the sparse tables of a grid exactly as ppn_create does, fills them with a random
// symmetric positive definite matrix on a random topology (lines off, buses inactive = identity rows, node bits in the
// F structure) and replays on the host, table by table, what the kernels do -- level factorisation + level-scheduled
// solve (sp_factor / sp_solve) and the hybrid sparse / dense-top-block factor and solve (hyb_*) -- against a dense
// Gaussian elimination.  Returns the largest deviation of the two solutions.  This is a check of the TABLES (ordering,
// fill pattern, update lists, levels, packed row/column views, cut); it is not a compute path of the library.
extern "C" int ppn_sparse_selfcheck(int n_sub, int n_line, const int32_t* line_or_sub, const int32_t* line_ex_sub, int full,
                                    uint32_t seed, double* max_err_out, int32_t* info_out /* n, nnz, n_lev, cut_lev, nt or NULL */) {
    if (n_sub <= 0 || n_line <= 0 || !line_or_sub || !line_ex_sub || !max_err_out) return fail(nullptr, PPN_E_INVALID, "ppn_sparse_selfcheck: bad arguments");
    const int S = n_sub, N = n_line, NB = 2 * S;
    SparseHost h;
    build_sparse(S, N, line_or_sub, line_ex_sub, min_degree_order(S, N, line_or_sub, line_ex_sub), full != 0, h);
    const int n = h.n, nnz = h.nnz, cut = choose_cut(h), r0 = h.lev_rows_ptr[cut], nt = n - r0, cut_ent = h.colptr[r0];
    if (info_out) { info_out[0] = n; info_out[1] = nnz; info_out[2] = h.n_lev; info_out[3] = cut; info_out[4] = nt; }
    uint64_t st = 0x9E3779B97F4A7C15ull ^ seed;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
    // topology: node bits (F only), line status, active buses
    std::vector<int> on(N), en(N), stat(N);
    for (int l = 0; l < N; l++) { on[l] = full ? rnd() < .2 : 0; en[l] = full ? rnd() < .2 : 0; stat[l] = rnd() < .9; }
    std::vector<char> active(NB, 0);
    for (int l = 0; l < N; l++) if (stat[l]) { active[line_or_sub[l] + S * on[l]] = 1; active[line_ex_sub[l] + S * en[l]] = 1; }
    for (int b = 0; b < NB; b++) if (active[b] && rnd() < .05) active[b] = 0;   // reference / PV rows of B''
    // assemble as the kernels do: off-diagonals at line_pos, diagonals per row; identity rows elsewhere
    std::vector<double> Lv(nnz, 0.0), dg(n, 1.0), T(nnz, 0.0), dense((size_t)n * n, 0.0), rhs(n, 0.0);
    for (int i = 0; i < n; i++) dense[(size_t)i * n + i] = 1.0;
    std::vector<double> diag(NB, 0.0);
    for (int l = 0; l < N; l++) {
        if (!stat[l]) continue;
        const int f = line_or_sub[l] + S * on[l], t = line_ex_sub[l] + S * en[l];
        const double w = 0.5 + 1.5 * rnd();
        diag[f] += w; diag[t] += w;
        if (active[f] && active[t]) {
            const int pos = h.line_pos[full ? 4 * l + 2 * on[l] + en[l] : l];
            if (pos < 0) return fail(nullptr, PPN_E_STATE, "ppn_sparse_selfcheck: line without a position in the pattern");
            Lv[pos] -= w;
            const int i = h.bus_row[f], j = h.bus_row[t];
            dense[(size_t)i * n + j] -= w; dense[(size_t)j * n + i] -= w;
        }
    }
    for (int b = 0; b < NB; b++) {
        const int i = h.bus_row[b];
        if (i < 0 || !active[b]) continue;
        dg[i] = diag[b] + 0.01;
        dense[(size_t)i * n + i] = dg[i];
        rhs[i] = rnd() - 0.5;
    }
    // dense reference solution (Gaussian elimination with partial pivoting)
    std::vector<double> a = dense, xref = rhs;
    for (int k = 0; k < n; k++) {
        int pv = k;
        for (int i = k + 1; i < n; i++) if (fabs(a[(size_t)i * n + k]) > fabs(a[(size_t)pv * n + k])) pv = i;
        if (pv != k) { for (int j = 0; j < n; j++) std::swap(a[(size_t)k * n + j], a[(size_t)pv * n + j]); std::swap(xref[k], xref[pv]); }
        for (int i = k + 1; i < n; i++) {
            const double m = a[(size_t)i * n + k] / a[(size_t)k * n + k];
            if (m == 0.0) continue;
            for (int j = k; j < n; j++) a[(size_t)i * n + j] -= m * a[(size_t)k * n + j];
            xref[i] -= m * xref[k];
        }
    }
    for (int k = n - 1; k >= 0; k--) {
        for (int j = k + 1; j < n; j++) xref[k] -= a[(size_t)k * n + j] * xref[j];
        xref[k] /= a[(size_t)k * n + k];
    }
    double err = 0.0;
    const std::vector<double> Lv0 = Lv, dg0 = dg;
    auto target = [&](int id, int limit_ent) {   // left-looking gather of one target; terms with entry (j,k) < limit_ent
        double acc = id < nnz ? Lv[id] : dg[id - nnz];
        for (int t = h.trip_ptr[id]; t < h.trip_ptr[id + 1]; t++) {
            const unsigned pk = (unsigned)h.trip[t];
            if ((int)(pk & 0xffffu) >= limit_ent) break;
            acc -= T[pk >> 16] * Lv[pk & 0xffffu];
        }
        return acc;
    };
    auto factor_levels = [&](int upto) {
        for (int lv = 0; lv < upto; lv++) {
            for (int q = h.lev_ptr[lv]; q < h.lev_ptr[lv + 1]; q++) {
                const int id = h.lev_ent[q];
                const double v = target(id, nnz);
                if (id < nnz) T[id] = v; else dg[id - nnz] = v;
            }
            for (int q = h.lev_ptr[lv]; q < h.lev_ptr[lv + 1]; q++) {
                const int id = h.lev_ent[q];
                if (id < nnz) Lv[id] = T[id] / dg[h.ecol[id]];
            }
        }
    };
    {   // (1) full level factorisation, level-scheduled solve through the packed row / column views
        factor_levels(h.n_lev);
        std::vector<double> w = rhs;
        for (int lv = 1; lv < h.n_lev; lv++)
            for (int i = h.lev_rows_ptr[lv]; i < h.lev_rows_ptr[lv + 1]; i++) {
                const unsigned pk = (unsigned)h.rowpk[i];
                double acc = w[i];
                for (unsigned t = pk >> 8, c = 0; c < (pk & 255u); c++, t++) {
                    const unsigned p0 = (unsigned)h.rpack[t];
                    acc -= Lv[(p0 >> 16) / 8] * w[(p0 & 0xffffu) / 8];
                }
                w[i] = acc;
            }
        for (int lv = h.n_lev - 1; lv >= 0; lv--)
            for (int i = h.lev_rows_ptr[lv]; i < h.lev_rows_ptr[lv + 1]; i++) {
                const unsigned pk = (unsigned)h.colpk[i];
                double acc = w[i] / dg[i];
                for (unsigned e2 = pk >> 8, c = 0; c < (pk & 255u); c++, e2++) acc -= Lv[e2] * w[h.rowoff[e2] / 8];
                w[i] = acc;
            }
        for (int i = 0; i < n; i++) err = fmax(err, fabs(w[i] - xref[i]));
    }
    {   // (2) hybrid: sparse levels below the cut, Schur complement of the top block, its dense inverse
        Lv = Lv0; dg = dg0; std::fill(T.begin(), T.end(), 0.0);
        factor_levels(cut);
        std::vector<double> Z((size_t)nt * nt, 0.0);
        for (int q = h.lev_ptr[cut]; q < h.lev_ptr[h.n_lev]; q++) {
            const int id = h.lev_ent[q];
            const int i = id < nnz ? h.rowidx[id] : id - nnz, j = id < nnz ? h.ecol[id] : id - nnz;
            const double v = target(id, cut_ent);
            Z[(size_t)(i - r0) * nt + (j - r0)] = v; Z[(size_t)(j - r0) * nt + (i - r0)] = v;
        }
        for (int k = 0; k < nt; k++) {   // Gauss-Jordan without pivoting, as hyb_invert2
            const double p = 1.0 / Z[(size_t)k * nt + k];
            for (int i = 0; i < nt; i++) {
                if (i == k) continue;
                const double ci = Z[(size_t)i * nt + k] * p;
                for (int j = 0; j < nt; j++) if (j != k) Z[(size_t)i * nt + j] -= ci * Z[(size_t)k * nt + j];
                Z[(size_t)i * nt + k] = -ci;
            }
            for (int j = 0; j < nt; j++) Z[(size_t)k * nt + j] = j == k ? p : Z[(size_t)k * nt + j] * p;
        }
        std::vector<double> w = rhs;
        for (int lv = 1; lv < cut; lv++)
            for (int i = h.lev_rows_ptr[lv]; i < h.lev_rows_ptr[lv + 1]; i++) {
                const unsigned pk = (unsigned)h.rowpk[i];
                double acc = w[i];
                for (unsigned t = pk >> 8, c = 0; c < (pk & 255u); c++, t++) {
                    const unsigned p0 = (unsigned)h.rpack[t];
                    acc -= Lv[(p0 >> 16) / 8] * w[(p0 & 0xffffu) / 8];
                }
                w[i] = acc;
            }
        std::vector<double> y2(nt);
        for (int t = 0; t < nt; t++) {
            const unsigned pk = (unsigned)h.rowpk[r0 + t];
            double acc = w[r0 + t];
            for (unsigned q = pk >> 8, c = 0; c < (pk & 255u); c++, q++) {
                const unsigned p0 = (unsigned)h.rpack[q];
                if ((p0 & 0xffffu) < 8u * (unsigned)r0) acc -= Lv[(p0 >> 16) / 8] * w[(p0 & 0xffffu) / 8];
            }
            y2[t] = acc;
        }
        for (int t = 0; t < nt; t++) {
            double x = 0.0;
            for (int j = 0; j < nt; j++) x += Z[(size_t)t * nt + j] * y2[j];
            w[r0 + t] = x;
        }
        for (int lv = cut - 1; lv >= 0; lv--)
            for (int i = h.lev_rows_ptr[lv]; i < h.lev_rows_ptr[lv + 1]; i++) {
                const unsigned pk = (unsigned)h.colpk[i];
                double acc = w[i] / dg[i];
                for (unsigned e2 = pk >> 8, c = 0; c < (pk & 255u); c++, e2++) acc -= Lv[e2] * w[h.rowoff[e2] / 8];
                w[i] = acc;
            }
        for (int i = 0; i < n; i++) err = fmax(err, fabs(w[i] - xref[i]));
        // (3) the same solve through the one-warp schedule (PpnDevSparse.o_wsched), step by step as the kernel walks it
        build_warp_schedule(h, cut);
        if (!h.wsched.empty()) {
            const int* ws = h.wsched.data();
            const int nf = ws[0], nb = ws[1];
            const int* S = ws + 2;
            const unsigned short* E = reinterpret_cast<const unsigned short*>(ws + 2 + 2 * (nf + nb));
            std::vector<double> v = rhs, tmp(32);
            auto run = [&](int s, bool backward) {
                const int row0 = S[2 * s] & 0xffff, rows = S[2 * s] >> 16, maxcnt = S[2 * s + 1] & 255, off = S[2 * s + 1] >> 8;
                for (int l = 0; l < rows; l++) {
                    double acc = backward ? v[row0 + l] / dg[row0 + l] : v[row0 + l];
                    for (int q = 0; q < maxcnt; q++) {
                        const unsigned p0 = E[off + q * 32 + l];
                        if (p0 != 0xffffu) acc -= Lv[p0 >> 7] * v[p0 & 127u];
                    }
                    tmp[l] = acc;
                }
                for (int l = 0; l < rows; l++) v[row0 + l] = tmp[l];
            };
            for (int s2 = 0; s2 < nf; s2++) run(s2, false);
            std::vector<double> x2(nt);
            for (int t = 0; t < nt; t++) {
                double x = 0.0;
                for (int j = 0; j < nt; j++) x += Z[(size_t)t * nt + j] * v[r0 + j];
                x2[t] = x;
            }
            for (int t = 0; t < nt; t++) v[r0 + t] = x2[t];
            for (int s2 = nf; s2 < nf + nb; s2++) run(s2, true);
            for (int i = 0; i < n; i++) err = fmax(err, fabs(v[i] - xref[i]));
        }
    }
    *max_err_out = err;
    return PPN_OK;
}

extern "C" const char* ppn_build_info(void) {
    return "pypownet_b200 step path; sm_100a; fused warp/CTA-per-env fast-decoupled XB + DC load-flow; built " __DATE__;
}

extern "C" const char* ppn_last_error(const ppn_env* env) { return env ? env->err.c_str() : g_err.c_str(); }

static int create_body(const ppn_case* g, const ppn_config* cfg, int n_envs, int device, ppn_env* env);

extern "C" int ppn_create(const ppn_case* g, const ppn_config* cfg, int n_envs, int device, ppn_env** out) {
    ppn_env* env = nullptr;
    if (!g || !cfg || !out || n_envs <= 0) return fail(nullptr, PPN_E_INVALID, "ppn_create: null argument or n_envs <= 0");
    const int S = g->n_sub, G = g->n_gen, L = g->n_load, N = g->n_line;
    if (S <= 0 || G <= 0 || L < 0 || N <= 0 || 2 * S > 32000) return fail(nullptr, PPN_E_INVALID, "ppn_create: bad grid sizes");
    if (!(g->base_mva > 0)) return fail(nullptr, PPN_E_INVALID, "ppn_create: base_mva must be positive");
    for (int i = 0; i < G; i++) {
        if (g->gen_sub[i] < 0 || g->gen_sub[i] >= S || (i && g->gen_sub[i] <= g->gen_sub[i - 1]))
            return fail(nullptr, PPN_E_INVALID, "ppn_create: gen_sub must be strictly ascending substation indices");
    }
    for (int i = 0; i < L; i++) {
        if (g->load_sub[i] < 0 || g->load_sub[i] >= S || (i && g->load_sub[i] <= g->load_sub[i - 1]))
            return fail(nullptr, PPN_E_INVALID, "ppn_create: load_sub must be strictly ascending substation indices");
    }
    for (int i = 0; i < N; i++) {
        if (g->line_or_sub[i] < 0 || g->line_or_sub[i] >= S || g->line_ex_sub[i] < 0 || g->line_ex_sub[i] >= S)
            return fail(nullptr, PPN_E_INVALID, "ppn_create: line end out of range");
        if (g->line_x[i] == 0.0) return fail(nullptr, PPN_E_INVALID, "ppn_create: line with zero reactance");
        if (g->line_or_sub[i] == g->line_ex_sub[i])   // no entry in the sparse patterns (and none in the shipped grids)
            return fail(nullptr, PPN_E_INVALID, "ppn_create: line with both ends in the same substation");
    }
    if (g->slack_sub < 0 || g->slack_sub >= S) return fail(nullptr, PPN_E_INVALID, "ppn_create: slack_sub out of range");
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, PPN_E_CUDA, "ppn_create: cudaSetDevice failed (no such CUDA device)");
    env = new ppn_env();
    const int rc = create_body(g, cfg, n_envs, device, env);
    if (rc != PPN_OK) {   // nothing of a half-built handle survives
        const std::string msg = env->err;
        ppn_destroy(env);
        return fail(nullptr, rc, msg);
    }
    *out = env;
    return PPN_OK;
}

static int create_body(const ppn_case* g, const ppn_config* cfg, int n_envs, int device, ppn_env* env) {
    const int S = g->n_sub, G = g->n_gen, L = g->n_load, N = g->n_line;
    env->device = device;
    env->B = n_envs;
    env->S = S; env->G = G; env->L = L; env->N = N; env->NB = 2 * S;
    env->A = G + L + 3 * N;
    env->OBSD = 7 * L + 7 * G + 13 * N + S + 6;
    env->OBS = 9 * L + 9 * G + 18 * N + 2 * S + 6;
    const int NB = 2 * S;

    // ---- derived static tables
    std::vector<int> gen_sub(g->gen_sub, g->gen_sub + G), load_sub(g->load_sub, g->load_sub + L),
        lor(g->line_or_sub, g->line_or_sub + N), lex(g->line_ex_sub, g->line_ex_sub + N);
    std::vector<int> gen_of_sub(S, -1), load_of_sub(S, -1), adj_ptr(S + 1, 0), adj(2 * N), elem_sub;
    for (int i = 0; i < G; i++) gen_of_sub[gen_sub[i]] = i;
    for (int i = 0; i < L; i++) load_of_sub[load_sub[i]] = i;
    for (int l = 0; l < N; l++) { adj_ptr[lor[l] + 1]++; adj_ptr[lex[l] + 1]++; }
    for (int s = 0; s < S; s++) adj_ptr[s + 1] += adj_ptr[s];
    {
        std::vector<int> fill(adj_ptr.begin(), adj_ptr.end() - 1);
        for (int l = 0; l < N; l++) { adj[fill[lor[l]]++] = 2 * l; adj[fill[lex[l]]++] = 2 * l + 1; }
    }
    elem_sub.insert(elem_sub.end(), gen_sub.begin(), gen_sub.end());
    elem_sub.insert(elem_sub.end(), load_sub.begin(), load_sub.end());
    elem_sub.insert(elem_sub.end(), lor.begin(), lor.end());
    elem_sub.insert(elem_sub.end(), lex.begin(), lex.end());
    // makeYbus per line: Ys = 1/(r+jx); Ytt = Ys + jb/2; Yff = Ytt/tap^2; Yft = Ytf = -Ys/tap
    std::vector<double> line_y(8 * N), bp(N), bdc(N);
    for (int l = 0; l < N; l++) {
        const double tap = (g->line_tap && g->line_tap[l] != 0.0) ? g->line_tap[l] : 1.0;
        const std::complex<double> ys = 1.0 / std::complex<double>(g->line_r[l], g->line_x[l]);
        const std::complex<double> ytt = ys + std::complex<double>(0.0, g->line_b[l] / 2);
        const std::complex<double> yff = ytt / (tap * tap);
        const std::complex<double> yft = -ys / tap, ytf = -ys / tap;
        line_y[8 * l + 0] = yff.real(); line_y[8 * l + 1] = yff.imag();
        line_y[8 * l + 2] = yft.real(); line_y[8 * l + 3] = yft.imag();
        line_y[8 * l + 4] = ytf.real(); line_y[8 * l + 5] = ytf.imag();
        line_y[8 * l + 6] = ytt.real(); line_y[8 * l + 7] = ytt.imag();
        bp[l] = 1.0 / g->line_x[l];
        bdc[l] = 1.0 / g->line_x[l] / tap;
    }
    std::vector<double> ysh_r(NB), ysh_i(NB), basekv(g->bus_basekv, g->bus_basekv + NB), vm0(g->bus_vm0, g->bus_vm0 + NB),
        va0(g->bus_va0, g->bus_va0 + NB);
    for (int b = 0; b < NB; b++) { ysh_r[b] = g->bus_gs[b] / g->base_mva; ysh_i[b] = g->bus_bs[b] / g->base_mva; }
    std::vector<double> qmin(g->gen_qmin, g->gen_qmin + G), qmax(g->gen_qmax, g->gen_qmax + G),
        pg0(g->gen_pg0, g->gen_pg0 + G), qg0(g->gen_qg0, g->gen_qg0 + G), vg0(g->gen_vg0, g->gen_vg0 + G),
        pd0(g->load_pd0, g->load_pd0 + L), qd0(g->load_qd0, g->load_qd0 + L), thermal(g->thermal_limits, g->thermal_limits + N);
    std::vector<uint8_t> status0(g->line_status0, g->line_status0 + N);

    PpnDevCase& c = env->dc;
    c.S = S; c.G = G; c.L = L; c.N = N; c.NB = NB; c.A = env->A; c.OBSD = env->OBSD;
    c.slack_bus = g->slack_sub;
    c.base_mva = g->base_mva;
    int* ip; double* dp; uint8_t* up8;
    UP(ip, gen_sub); c.gen_sub = ip; UP(ip, load_sub); c.load_sub = ip; UP(ip, lor); c.lor_sub = ip; UP(ip, lex); c.lex_sub = ip;
    UP(ip, gen_of_sub); c.gen_of_sub = ip; UP(ip, load_of_sub); c.load_of_sub = ip;
    UP(ip, adj_ptr); c.adj_ptr = ip; UP(ip, adj); c.adj = ip; UP(ip, elem_sub); c.elem_sub = ip;
    UP(dp, line_y); c.line_y = dp; UP(dp, bp); c.line_bp = dp; UP(dp, bdc); c.line_bdc = dp;
    UP(dp, ysh_r); c.bus_ysh_r = dp; UP(dp, ysh_i); c.bus_ysh_i = dp; UP(dp, basekv); c.bus_basekv = dp;
    UP(dp, vm0); c.bus_vm0 = dp; UP(dp, va0); c.bus_va0 = dp;
    UP(dp, qmin); c.gen_qmin = dp; UP(dp, qmax); c.gen_qmax = dp; UP(dp, pg0); c.gen_pg0 = dp; UP(dp, qg0); c.gen_qg0 = dp;
    UP(dp, vg0); c.gen_vg0 = dp; UP(dp, pd0); c.load_pd0 = dp; UP(dp, qd0); c.load_qd0 = dp; UP(dp, thermal); c.thermal = dp;
    UP(up8, status0); c.line_status0 = up8;

    // sparse LDL^T structures (U: un-split grid, F: any bus split)
    {
        const std::vector<int> perm = min_degree_order(S, N, lor.data(), lex.data());
        for (int f = 0; f < 2; f++) {
            SparseHost h;
            build_sparse(S, N, lor.data(), lex.data(), perm, f == 1, h);
            // byte offsets of entries / rows are packed into 16 bits
            if (h.nnz >= 8192 || 8 * h.n >= 32768) { return fail(env, PPN_E_UNSUPPORTED, "ppn_create: grid too large for the sparse factor tables"); }
            PpnDevSparse& d = c.sp[f];
            d.n = h.n; d.nnz = h.nnz; d.n_lev = h.n_lev;
            const int cut = choose_cut(h);
            d.cut_lev = cut; d.cut_row = h.lev_rows_ptr[cut]; d.cut_ent = h.colptr[d.cut_row]; d.nt = h.n - d.cut_row;
            std::vector<int> blob;
            auto put_i = [&](const std::vector<int>& v) { const int o = (int)blob.size(); blob.insert(blob.end(), v.begin(), v.end()); return o; };
            auto put_s = [&](const std::vector<short>& v) {
                const int o = (int)blob.size();
                blob.resize(o + (v.size() + 1) / 2, 0);
                if (!v.empty()) memcpy(blob.data() + o, v.data(), v.size() * sizeof(short));
                return o;
            };
            // the tables the hybrid solver reads come first: a CTA stages only that prefix (hyb_words)
            build_warp_schedule(h, cut);
            d.o_lev_ptr = put_i(h.lev_ptr); d.o_lev_ent = put_i(h.lev_ent);
            d.o_trip_ptr = put_i(h.trip_ptr); d.o_trip = put_i(h.trip); d.o_line_pos = put_i(h.line_pos);
            d.o_lev_rows_ptr = put_i(h.lev_rows_ptr);
            d.o_rowidx = put_s(h.rowidx); d.o_ecol = put_s(h.ecol); d.o_bus_row = put_s(h.bus_row);
            d.o_wsched = h.wsched.empty() ? -1 : put_i(h.wsched);
            if (blob.size() & 1) blob.push_back(0);
            d.hyb_words = (int)blob.size();
            d.o_colptr = put_i(h.colptr);
            d.o_rowptr = put_i(h.rowptr); d.o_rowent = put_i(h.rowent);
            d.o_lev_rows = put_i(h.lev_rows); d.o_rpack = put_i(h.rpack);
            d.o_rowpk = put_i(h.rowpk); d.o_colpk = put_i(h.colpk);
            d.o_parent = put_s(h.parent);
            d.o_row_lev = put_s(h.row_lev); d.o_rowoff = put_s(h.rowoff);
            if (blob.size() & 1) blob.push_back(0);
            d.blob_words = (int)blob.size();
            if (d.o_wsched < 0) d.hyb_words = d.blob_words;   // the level-by-level solve reads the packed row / column views
            int* bp32;
            UP(bp32, blob); d.blob = bp32;
            env->sp_need[f] = 2 * ppn_sp_factor_doubles(h.n, h.nnz) + 2 * d.nt * (d.nt | 1);   // + the dense top blocks
            env->sp_blob_dbl[f] = d.blob_words / 2;
            env->sp_hyb_dbl[f] = d.hyb_words / 2;
        }
    }

    // static tail of Observation.as_array (environment.py:583-595)
    {
        std::vector<double>& t = env->obs_static;
        for (int s = 0; s < S; s++) t.push_back((double)g->sub_ids[s]);
        for (int i = 0; i < L; i++) t.push_back((double)g->sub_ids[load_sub[i]]);
        for (int i = 0; i < G; i++) t.push_back((double)g->sub_ids[gen_sub[i]]);
        for (int i = 0; i < N; i++) t.push_back((double)g->sub_ids[lor[i]]);
        for (int i = 0; i < N; i++) t.push_back((double)g->sub_ids[lex[i]]);
        for (int i = 0; i < N; i++) t.push_back(thermal[i]);
        for (int i = 0; i < G + L + 2 * N; i++) t.push_back(0.0);
    }

    // ---- configuration
    PpnDevCfg& k = env->dcfg;
    k.dc = cfg->dc;
    k.hard_coef = cfg->hard_overflow_coefficient;
    k.n_hard_broken = cfg->n_timesteps_hard_overflow_is_broken;
    k.n_soft_consec = cfg->n_timesteps_consecutive_soft_overflow_breaks;
    k.n_soft_broken = cfg->n_timesteps_soft_overflow_is_broken;
    k.max_prods_go = cfg->max_number_prods_game_over;
    k.max_loads_go = cfg->max_number_loads_game_over;
    k.n_line_react = cfg->n_timesteps_actionned_line_reactionable;
    k.n_node_react = cfg->n_timesteps_actionned_node_reactionable;
    k.max_sub = cfg->max_number_actionned_substations;
    k.max_lines = cfg->max_number_actionned_lines;
    k.max_total = cfg->max_number_actionned_total;
    k.hard_mode = cfg->hard_game_over;
    k.loop_mode = cfg->loop_mode;
    k.tol = cfg->pf_tol > 0 ? cfg->pf_tol : 1e-6;
    k.max_it = cfg->pf_max_it > 0 ? cfg->pf_max_it : 25;
    k.alg = cfg->pf_alg == 1 ? 1 : 2;
    k.max_it_nr = 10;
    k.reward_k = cfg->reward_constant;
    k.seed = cfg->seed;
    k.max_reset_attempts = 64;

    // ---- threads per env and shared-memory plan
    int tpe = cfg->threads_per_env;
    // measured on B200: a full warp per env beats two envs per warp; CTA-per-env grids run 128-thread CTAs (two buses
    // per thread, three CTAs per SM: IEEE-118 2.14 M env-steps/s against 1.72 M with 256-thread CTAs, two per SM)
    if (tpe == 0) tpe = (NB <= 64) ? 32 : 128;
    if (const char* v = getenv("PPN_TPE")) { if (NB > 64 && (atoi(v) == 128 || atoi(v) == 256)) tpe = atoi(v); }
    if (tpe != 16 && tpe != 32 && tpe != 128 && tpe != 256) { return fail(env, PPN_E_INVALID, "threads_per_env must be 0, 16, 32, 128 or 256"); }
    if ((tpe == 16 && NB > 32) || (tpe == 32 && NB > 64) || NB > 256) { return fail(env, PPN_E_INVALID, "threads_per_env too small for this grid (16: <= 16 substations, 32: <= 32, 256: <= 128)"); }
    env->tpe = tpe;
    const int fixed = ppn_env_smem_fixed_bytes(S, G, L, N, tpe);
    int max_smem = 0;
    CK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    // Shared-memory room for B' and B'': the un-split grid (n1 = S-1 buses) with up to two generators off (n2 = PQ
    // buses); larger systems (node splitting, more generators off) use the env's slice of the HBM workspace.
    const int n1 = S - 1, n2 = S - 1 - (G > 3 ? G - 3 : 0);
    // Linear solver (PpnStepArgs.sparse), chosen per grid size from measurements on B200; PPN_SPARSE overrides:
    //   0 dense Gauss-Jordan inverses      IEEE-14 sized grids: both inverses by one warp with the rows in registers
    //   1 sparse LDL^T + explicit inverses warp per env above 16 substations (IEEE-30: +46 % over the dense inverse);
    //                                      also the CTA-per-env plan once a handle has received topology actions
    //   3 hybrid sparse / dense top block  CTA per env (IEEE-118): 4.7x the dense 117^3 inverse, two CTAs per SM
    //   (2 = sparse LDL^T with level-scheduled triangular solves: correct but latency-bound, kept for reference)
    env->sparse = tpe >= 128 ? 3 : (tpe == 32 && S > 16 ? 1 : 0);
    if (env->sparse == 3 && (env->dc.sp[0].nt > 40 || env->dc.sp[1].nt > 40)) env->sparse = 1;   // hyb_invert2 tiles hold <= 40 rows
    if (const char* v = getenv("PPN_SPARSE")) env->sparse = atoi(v);
    env->ws_dense = 2LL * NB * (NB | 1);
    int worst = (int)env->ws_dense + env->sp_need[1] + env->sp_blob_dbl[1];
    if (cfg->pf_alg == 1 && worst < 2 * NB * (2 * NB + 1)) worst = 2 * NB * (2 * NB + 1);   // the Newton-Raphson Jacobian + right-hand side
    env->envs_per_block = tpe == 16 ? 4 : (tpe == 32 ? 2 : 1);   // 64-thread CTAs for the sub-warp / warp kernels
    if (const char* v = getenv("PPN_EPB")) { if (tpe == 32 && atoi(v) >= 1 && atoi(v) <= 2) env->envs_per_block = atoi(v); }
    const int cap_bytes = max_smem / env->envs_per_block - fixed - 64;
    if (cap_bytes < 0) { return fail(env, PPN_E_UNSUPPORTED, "grid too large for the shared-memory plan"); }
    // doubles of shared memory per env for matrices / factors / tables under a given solver mode
    auto plan = [&](int mode) {
        int want = n1 * (n1 | 1) + n2 * (n2 | 1);
        if (mode == 1) want = (n1 + 1) * (n1 | 1) + (n2 + 1) * (n2 | 1) + env->sp_need[0] + env->sp_blob_dbl[0];
        if (mode == 1 && tpe <= 32 && !getenv("PPN_WARP_TABLES_SMEM")) {
            // warp per env: the inverses and the factor values live in shared memory, the read-only index tables are read
            // through L1 from their one copy in HBM (every env staging its own copy cost a third of the shared memory:
            // IEEE-30 went from 6 to 8 resident envs per SM), and the dense top blocks of the hybrid solver are not used
            const PpnDevSparse& u = env->dc.sp[0];
            want = (n1 + 1) * (n1 | 1) + (n2 + 1) * (n2 | 1) + 2 * ppn_sp_factor_doubles(u.n, u.nnz);
        }
        if (mode >= 2) want = env->sp_need[0] + (mode == 3 ? env->sp_hyb_dbl[0] : env->sp_blob_dbl[0]);
        if (cfg->pf_alg == 1 && tpe <= 32) {   // Newton-Raphson: the Jacobian of the un-split grid + right-hand side
            const int nj = n1 + n2;
            if (want < nj * (nj + 1)) want = nj * (nj + 1);
        }
        want = (want + 3) & ~1;   // the capacity below is rounded down to an even number of doubles
        if (want > worst) want = worst;
        int cap = cap_bytes / 8;
        if (tpe <= 32 || mode >= 2) {   // keep occupancy: the un-split grid fits, rare bigger systems spill to HBM
            if (want < env->OBSD) want = env->OBSD;   // the area also stages the observation row (write_observation)
            if (cap > want) cap = want;
        } else if (cap > worst) cap = worst;           // one env per CTA: take what the SM has
        return cap & ~1;
    };
    env->mat_cap = plan(env->sparse);
    env->env_smem_bytes = fixed + env->mat_cap * 8;
    // Second plan of CTA-per-env grids, used once the handle has received actions (buses may be split from then on):
    // the F structure does not fit the small hybrid plan, explicit inverses with the whole SM's shared memory do better
    // there (IEEE-118, random node-splitting agent: 0.56 M env-steps/s against 0.24 M).
    env->alt_sparse = env->sparse; env->alt_mat_cap = env->mat_cap; env->alt_env_smem_bytes = env->env_smem_bytes;
    env->alt_tpe = tpe;
    if (tpe >= 128 && env->sparse >= 2 && !getenv("PPN_NO_ALT_PLAN")) {
        // explicit inverses want one thread per column: 256-thread CTAs (the per-env state in HBM does not depend on
        // the thread count; only the reduction scratch of the shared-memory image does)
        const int fixed_b = ppn_env_smem_fixed_bytes(S, G, L, N, 256);
        env->alt_tpe = 256;
        env->alt_sparse = 1;
        int cap_b = (max_smem - fixed_b - 64) / 8;
        if (cap_b > worst) cap_b = worst;
        env->alt_mat_cap = cap_b & ~1;
        env->alt_env_smem_bytes = fixed_b + env->alt_mat_cap * 8;
    }
    env->ws_stride = worst;
    env->horizon = cfg->n_timesteps_horizon_maintenance > 0 ? cfg->n_timesteps_horizon_maintenance : 1;
    env->ws_rows = n_envs;
    CK(cudaMalloc(&env->ws, (size_t)env->ws_rows * worst * sizeof(double)));
    env->allocs.push_back(env->ws);

    // ---- per-env state
    env->st.rw = 2 * NB + 2 * L + 3 * G;
    env->st.tw = (2 * G + L + 3 * N + 15) & ~15;
    env->st.cw = (3 * N + S + 4 + 3) & ~3;
    CK(cudaMalloc(&env->st.real, (size_t)n_envs * env->st.rw * sizeof(double)));
    env->allocs.push_back(env->st.real);
    CK(cudaMalloc(&env->st.topo, (size_t)n_envs * env->st.tw));
    env->allocs.push_back(env->st.topo);
    CK(cudaMalloc(&env->st.cnt, (size_t)n_envs * env->st.cw * sizeof(int32_t)));
    env->allocs.push_back(env->st.cnt);
    CK(cudaMemset(env->st.real, 0, (size_t)n_envs * env->st.rw * sizeof(double)));
    CK(cudaMemset(env->st.topo, 0, (size_t)n_envs * env->st.tw));
    CK(cudaMemset(env->st.cnt, 0, (size_t)n_envs * env->st.cw * sizeof(int32_t)));
    CK(cudaMalloc(&env->stats, 16 * sizeof(unsigned long long)));
    env->allocs.push_back(env->stats);
    CK(cudaMemset(env->stats, 0, 16 * sizeof(unsigned long long)));
    CK(cudaMalloc(&env->d_init, (size_t)2 * n_envs * sizeof(int32_t)));
    env->allocs.push_back(env->d_init);
    CK(cudaStreamCreateWithFlags(&env->own_stream, cudaStreamNonBlocking));
    if (env->alt_sparse != env->sparse) {
        CK(cudaHostAlloc(&env->h_split, sizeof(int), cudaHostAllocMapped));
        *env->h_split = 0;
        CK(cudaHostGetDevicePointer(&env->d_split, env->h_split, 0));
    }
    return PPN_OK;
}

extern "C" void ppn_destroy(ppn_env* env) {
    if (!env) return;
    cudaSetDevice(env->device);
    for (void* p : env->allocs) cudaFree(p);
    void* pinned[] = {env->h_act, env->h_obs, env->h_reward, env->h_done, env->h_flag, env->h_ill};
    for (void* p : pinned) if (p) cudaFreeHost(p);
    void* dev[] = {env->d_act, env->d_obs, env->d_reward, env->d_done, env->d_flag, env->d_ill};
    for (void* p : dev) if (p) cudaFree(p);
    if (env->own_stream) cudaStreamDestroy(env->own_stream);
    if (env->drain_stream) cudaStreamDestroy(env->drain_stream);
    cudaFree(env->d_row_flag); cudaFreeHost(env->d_drain_err);
    if (env->h_split) cudaFreeHost(env->h_split);
    for (int i = 0; i < env->n_chunks; i++) if (env->chunk_stream[i]) cudaStreamDestroy(env->chunk_stream[i]);
    delete env;
}

extern "C" int ppn_load_chronics(ppn_env* env, int n_chronics, const ppn_chronic* t) {
    if (!env || !t || n_chronics <= 0) return fail(env, PPN_E_INVALID, "ppn_load_chronics: bad arguments");
    if (env->chronics_loaded) return fail(env, PPN_E_STATE, "ppn_load_chronics: chronics already loaded for this handle");
    CK(cudaSetDevice(env->device));
    const int G = env->G, L = env->L, N = env->N;
    PpnDevChronics& d = env->dch;
    int o = 0;
    d.o_pp = o; o += G; d.o_pv = o; o += G; d.o_lp = o; o += L; d.o_lq = o; o += L; d.o_mt = o; o += N; d.o_hz = o; o += N;
    d.o_ppp = o; o += G; d.o_pvp = o; o += G; d.o_lpp = o; o += L; d.o_lqp = o; o += L; d.o_pm = o; o += N; d.o_dt = o; o += 6;
    d.row_words = (o + 3) & ~3;
    d.n_chronics = n_chronics;
    std::vector<int> row_off(n_chronics), n_rows(n_chronics), ras(n_chronics), liz(n_chronics);
    long long total = 0;
    for (int c = 0; c < n_chronics; c++) {
        if (t[c].n_rows <= 0) return fail(env, PPN_E_INVALID, "ppn_load_chronics: chronic without rows");
        row_off[c] = (int)total; n_rows[c] = t[c].n_rows; total += t[c].n_rows;
    }
    if (total * d.row_words > 0x7fffffffLL * 2) return fail(env, PPN_E_UNSUPPORTED, "ppn_load_chronics: chronic table too large");
    std::vector<float> rows((size_t)total * d.row_words, 0.f);
    const int horizon = env->horizon;
    for (int c = 0; c < n_chronics; c++) {
        const ppn_chronic& s = t[c];
        const int T = s.n_rows;
        ras[c] = -1;
        for (int r = 0; r < T; r++)
            if (s.ids[r] == 0) { ras[c] = r + 1 < T ? r + 1 : T - 1; break; }
        liz[c] = s.ids[T - 1] == 0;
        for (int r = 0; r < T; r++) {
            float* w = rows.data() + (size_t)(row_off[c] + r) * d.row_words;
            memcpy(w + d.o_pp, s.prods_p + (size_t)r * G, G * sizeof(float));
            memcpy(w + d.o_pv, s.prods_v + (size_t)r * G, G * sizeof(float));
            memcpy(w + d.o_lp, s.loads_p + (size_t)r * L, L * sizeof(float));
            memcpy(w + d.o_lq, s.loads_q + (size_t)r * L, L * sizeof(float));
            memcpy(w + d.o_mt, s.maintenance + (size_t)r * N, N * sizeof(float));
            memcpy(w + d.o_hz, s.hazards + (size_t)r * N, N * sizeof(float));
            memcpy(w + d.o_ppp, s.prods_p_planned + (size_t)r * G, G * sizeof(float));
            memcpy(w + d.o_pvp, s.prods_v_planned + (size_t)r * G, G * sizeof(float));
            memcpy(w + d.o_lpp, s.loads_p_planned + (size_t)r * L, L * sizeof(float));
            memcpy(w + d.o_lqp, s.loads_q_planned + (size_t)r * L, L * sizeof(float));
            // timesteps before planned maintenance (chronic.py:239-246): first row index within the horizon
            int32_t* pm = reinterpret_cast<int32_t*>(w + d.o_pm);
            for (int l = 0; l < N; l++) {
                int first = 0;
                for (int h = 0; h < horizon && r + h < T; h++)
                    if (s.maintenance[(size_t)(r + h) * N + l] != 0.f) { first = h; break; }
                pm[l] = first;
            }
            int32_t* dt = reinterpret_cast<int32_t*>(w + d.o_dt);
            for (int q = 0; q < 6; q++) dt[q] = s.datetimes ? s.datetimes[(size_t)r * 6 + q] : 0;
        }
    }
    float* fp; int* ip;
    UP(fp, rows); d.rows = fp;
    UP(ip, row_off); d.row_off = ip; UP(ip, n_rows); d.n_rows = ip; UP(ip, ras); d.row_after_switch = ip; UP(ip, liz); d.last_id_zero = ip;
    env->chronic_rows = n_rows;
    env->chronics_loaded = true;
    return PPN_OK;
}

static int launch(ppn_env* env, PpnStepArgs& a, cudaStream_t s) {
    // Second plan (explicit inverses, whole SM per CTA) once buses have been split: only the DC path still needs it -- the AC
    // path borders the hybrid factor with the few sister buses in use and stays on the small plan, per env and per load-flow
    const bool alt = env->h_split && env->dcfg.dc && *(volatile int*)env->h_split != 0;   // may lag a launch or two
    a.split_flag = env->d_split;
    const int smem_bytes = alt ? env->alt_env_smem_bytes : env->env_smem_bytes;
    a.ws = env->ws; a.ws_stride = env->ws_stride; a.mat_cap = alt ? env->alt_mat_cap : env->mat_cap; a.stats = env->stats;
    a.sparse = alt ? env->alt_sparse : env->sparse; a.ws_dense = env->ws_dense;
    if (a.n_envs <= 0) { a.n_envs = env->B; a.env_off = 0; }   // whole batch unless the caller set a chunk
    {   // observation rows as TMA bulk stores when every row starts on a 16-byte boundary
        static const int no_bulk = getenv("PPN_NO_BULK") != nullptr;
        a.obs_bulk = !no_bulk && a.obs && (reinterpret_cast<size_t>(a.obs) & 15) == 0 && (a.obs_stride & (a.obs_f32 ? 3 : 1)) == 0;
    }
    if (a.n_cand <= 0) a.n_cand = 1;
    if (a.mode == PPN_MODE_SIMULATE && a.n_cand > 1) {
        // every candidate needs its own spill slice
        if ((long long)env->B * a.n_cand > env->ws_rows) {
            double* nws = nullptr;
            cudaError_t e2 = cudaMalloc(&nws, (size_t)env->B * a.n_cand * env->ws_stride * sizeof(double));
            if (e2 != cudaSuccess) return fail(env, PPN_E_CUDA, std::string("workspace for simulate: ") + cudaGetErrorString(e2));
            cudaFree(env->ws);   // synchronises the device: no launch still uses the previous workspace
            env->allocs.erase(std::remove(env->allocs.begin(), env->allocs.end(), (void*)env->ws), env->allocs.end());
            env->allocs.push_back(nws);
            env->ws = nws; env->ws_rows = (long long)env->B * a.n_cand;
            a.ws = nws;
        }
    }
    int rc = ppn_launch_step(&env->dc, &env->dch, &env->dcfg, &env->st, &a, alt ? env->alt_tpe : env->tpe, env->envs_per_block, smem_bytes, s);
    env->launches++;
    env->async_pending = true;
    if (rc != 0) return fail(env, PPN_E_CUDA, std::string("step kernel launch: ") + cudaGetErrorString((cudaError_t)rc));
    return PPN_OK;
}

extern "C" int ppn_reset(ppn_env* env, const int32_t* chronic_idx_host, const int32_t* row0_host, double* obs_dev,
                         int64_t obs_stride, int32_t* flag_dev, void* stream) {
    if (!env) return fail(nullptr, PPN_E_INVALID, "ppn_reset: null handle");
    if (!env->chronics_loaded) return fail(env, PPN_E_STATE, "ppn_reset: load chronics first");
    if (obs_dev && obs_stride < env->OBSD) return fail(env, PPN_E_INVALID, "ppn_reset: obs_stride smaller than the dynamic observation");
    CK(cudaSetDevice(env->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int B = env->B;
    std::vector<int32_t> init(2 * (size_t)B, 0);
    for (int e = 0; e < B; e++) {
        const int c = chronic_idx_host ? chronic_idx_host[e] : 0;
        if (c < 0 || c >= env->dch.n_chronics) return fail(env, PPN_E_INVALID, "ppn_reset: chronic index out of range");
        const int r = row0_host ? row0_host[e] : 0;
        if (r < 0 || r >= env->chronic_rows[c]) return fail(env, PPN_E_INVALID, "ppn_reset: first row out of range");
        init[e] = c; init[B + e] = r;
    }
    CK(cudaMemcpyAsync(env->d_init, init.data(), init.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));   // `init` is pageable and dies with this frame
    PpnStepArgs a{};
    a.mode = PPN_MODE_INIT;
    a.init_chronic = env->d_init; a.init_row0 = env->d_init + B;
    a.obs = obs_dev; a.obs_stride = obs_stride; a.flag = flag_dev;
    int rc = launch(env, a, s);
    if (rc == PPN_OK) env->initialised = true;
    return rc;
}

static int check_ready(ppn_env* env, const char* who) {
    if (!env) return fail(nullptr, PPN_E_INVALID, std::string(who) + ": null handle");
    if (!env->initialised) return fail(env, PPN_E_STATE, std::string(who) + ": call ppn_load_chronics and ppn_reset first");
    return PPN_OK;
}

extern "C" int ppn_step(ppn_env* env, const uint8_t* act_dev, double* obs_dev, int64_t obs_stride, double* reward_dev,
                        uint8_t* done_dev, int32_t* flag_dev, uint8_t* illegal_dev, int auto_reset, void* stream) {
    int rc = check_ready(env, "ppn_step");
    if (rc) return rc;
    if (obs_dev && obs_stride < env->OBSD) return fail(env, PPN_E_INVALID, "ppn_step: obs_stride smaller than the dynamic observation");
    CK(cudaSetDevice(env->device));
    PpnStepArgs a{};
    a.mode = PPN_MODE_STEP; a.auto_reset = auto_reset; a.act = act_dev;
    a.obs = obs_dev; a.obs_stride = obs_stride; a.reward = reward_dev; a.done = done_dev; a.flag = flag_dev; a.illegal = illegal_dev;
    a.pack = env->pack_dev;
    a.trace = env->trace_dev;
    return launch(env, a, (cudaStream_t)stream);
}

extern "C" int ppn_set_result_pack(ppn_env* env, double* pack_dev) {
    if (!env) return fail(nullptr, PPN_E_INVALID, "ppn_set_result_pack: null handle");
    env->pack_dev = pack_dev;
    return PPN_OK;
}

extern "C" int ppn_set_env_trace(ppn_env* env, int64_t* trace_dev) {
    if (!env) return fail(nullptr, PPN_E_INVALID, "ppn_set_env_trace: null handle");
    env->trace_dev = reinterpret_cast<long long*>(trace_dev);
    return PPN_OK;
}

extern "C" int ppn_simulate(ppn_env* env, int n_candidates, const uint8_t* act_dev, double* obs_dev, int64_t obs_stride,
                            double* reward_dev, uint8_t* done_dev, int32_t* flag_dev, uint8_t* illegal_dev, void* stream) {
    int rc = check_ready(env, "ppn_simulate");
    if (rc) return rc;
    if (n_candidates <= 0) return fail(env, PPN_E_INVALID, "ppn_simulate: n_candidates must be positive");
    if (obs_dev && obs_stride < env->OBSD) return fail(env, PPN_E_INVALID, "ppn_simulate: obs_stride smaller than the dynamic observation");
    CK(cudaSetDevice(env->device));
    PpnStepArgs a{};
    a.mode = PPN_MODE_SIMULATE; a.n_cand = n_candidates; a.act = act_dev;
    a.obs = obs_dev; a.obs_stride = obs_stride; a.reward = reward_dev; a.done = done_dev; a.flag = flag_dev; a.illegal = illegal_dev;
    return launch(env, a, (cudaStream_t)stream);
}

extern "C" int ppn_process_game_over(ppn_env* env, const uint8_t* mask_dev, double* obs_dev, int64_t obs_stride, void* stream) {
    int rc = check_ready(env, "ppn_process_game_over");
    if (rc) return rc;
    if (obs_dev && obs_stride < env->OBSD) return fail(env, PPN_E_INVALID, "ppn_process_game_over: obs_stride smaller than the dynamic observation");
    CK(cudaSetDevice(env->device));
    PpnStepArgs a{};
    a.mode = PPN_MODE_GAME_OVER; a.mask = mask_dev; a.obs = obs_dev; a.obs_stride = obs_stride;
    return launch(env, a, (cudaStream_t)stream);
}

// Game.is_action_valid (game.py:755-760): the legality tests of _verify_illegal_action, no state change.
__global__ void ppn_action_valid_kernel(PpnDevCase c, PpnDevCfg cfg, PpnDevState st, const uint8_t* act, uint8_t* valid, int B) {
    const int env = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (env >= B) return;
    const int lane = threadIdx.x & 31;
    const int NT = c.G + c.L + 2 * c.N;
    const uint8_t* a = act + (size_t)env * c.A;
    const int32_t* cnt = st.cnt + (size_t)env * st.cw;
    int ns = 0, nl = 0, bad = 0;
    for (int s = lane; s < c.S; s += 32) {
        bool ch = false;
        for (int i = 0; i < NT; i++) ch |= (a[i] != 0 && c.elem_sub[i] == s);
        ns += ch;
        bad += ch && cnt[3 * c.N + s] > 0;
    }
    for (int l = lane; l < c.N; l += 32) {
        const bool sw = a[NT + l] == 1;
        nl += sw;
        bad += sw && (cnt[l] > 0 || cnt[c.N + l] > 0);
    }
    ns = __reduce_add_sync(0xffffffffu, ns); nl = __reduce_add_sync(0xffffffffu, nl); bad = __reduce_add_sync(0xffffffffu, bad);
    if (lane == 0) valid[env] = !(ns > cfg.max_sub || nl > cfg.max_lines || ns + nl > cfg.max_total || bad > 0);
}

extern "C" int ppn_action_valid(ppn_env* env, const uint8_t* act_dev, uint8_t* valid_dev, void* stream) {
    int rc = check_ready(env, "ppn_action_valid");
    if (rc) return rc;
    if (!act_dev || !valid_dev) return fail(env, PPN_E_INVALID, "ppn_action_valid: null buffer");
    CK(cudaSetDevice(env->device));
    ppn_action_valid_kernel<<<(env->B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(env->dc, env->dcfg, env->st, act_dev, valid_dev, env->B);
    env->launches++;
    CK(cudaGetLastError());
    return PPN_OK;
}

static int ensure_staging(ppn_env* env) {
    if (env->d_act) return PPN_OK;
    const size_t B = env->B;
    const size_t iw = 1 + 2 * env->N + env->S;
    CK(cudaMallocHost(&env->h_act, B * env->A));
    CK(cudaMalloc(&env->d_act, B * env->A));
    CK(cudaMallocHost(&env->h_obs, B * env->OBSD * sizeof(double)));
    CK(cudaMalloc(&env->d_obs, B * env->OBSD * sizeof(double)));
    CK(cudaMallocHost(&env->h_reward, B * 5 * sizeof(double)));
    CK(cudaMalloc(&env->d_reward, B * 5 * sizeof(double)));
    CK(cudaMallocHost(&env->h_done, B));
    CK(cudaMalloc(&env->d_done, B));
    CK(cudaMallocHost(&env->h_flag, B * sizeof(int32_t)));
    CK(cudaMalloc(&env->d_flag, B * sizeof(int32_t)));
    CK(cudaMallocHost(&env->h_ill, B * iw));
    CK(cudaMalloc(&env->d_ill, B * iw));
    CK(cudaMalloc(&env->d_row_flag, B * sizeof(unsigned)));
    CK(cudaMemset(env->d_row_flag, 0, B * sizeof(unsigned)));
    CK(cudaHostAlloc(&env->d_drain_err, sizeof(int), cudaHostAllocMapped));   // read by the host after every step: no copy
    *env->d_drain_err = 0;
    {
        int lo = 0, hi = 0;   // "greatest" priority is the numerically lowest
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&env->drain_stream, cudaStreamNonBlocking, hi));
    }
    // two chunks measured best on B200 (every chunk ends with its own slowest env, so more chunks overlap no more copy
    // time and only add launches); PPN_HOST_CHUNKS overrides the count
    int n = B >= 1024 ? 2 : 1;
    if (const char* v = getenv("PPN_HOST_CHUNKS")) n = atoi(v);
    if (n < 1) n = 1;
    if (n > ppn_env::MAX_CHUNKS) n = ppn_env::MAX_CHUNKS;
    if (n > (int)B) n = (int)B;
    for (int i = 0; i < n; i++) CK(cudaStreamCreateWithFlags(&env->chunk_stream[i], cudaStreamNonBlocking));
    env->n_chunks = n;
    return PPN_OK;
}

// ---- drain kernel of the zero-copy host step.  A step CTA that stores its observation row straight into host memory holds
// its SM slot until PCIe has accepted the row, so the burst of the first wave (every env of a small grid finishes within
// ~100 us) delays the second wave by the time the link needs for it.  Instead the step kernel writes the row to device memory,
// release-stores the row's flag and retires; the few warps of this kernel -- launched right after the step kernel on a
// high-priority stream of their own -- poll the flags of their rows and move every finished row to the host while the other
// envs still iterate.  The step kernel never waits for this one, so there is no deadlock; a spin limit (about two seconds)
// turns a step kernel that hangs into an error instead of a second hang.
#define PPN_DRAIN_MAXJ 4   // rows a lane watches at most
__global__ void __launch_bounds__(1024) ppn_obs_drain_kernel(const unsigned long long* __restrict__ stage, long long stage_stride8,
                                                            unsigned long long* __restrict__ host, long long host_stride8, int n_rows,
                                                            int n8, const unsigned* flags, unsigned epoch, int n_warps, int* error) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= n_warps) return;
    // The warp owns rows gw, gw + n_warps, gw + 2 n_warps, ...  Interleaved, because the rows that finish last (the last
    // wave of step CTAs) are neighbours: contiguous ranges would leave that tail to a few warps.  It drains them in rounds
    // of 32 * PPN_DRAIN_MAXJ rows (one round for the default 128 rows per warp; more only for very large batches, whose
    // rows finish roughly in this order anyway): lane i watches the flags of rows gw + (i + 32 (jb + j)) * n_warps.
    const bool v16 = ((stage_stride8 | host_stride8) & 1) == 0 && ((reinterpret_cast<size_t>(stage) | reinterpret_cast<size_t>(host)) & 15) == 0;
    const int n16 = n8 >> 1;
    const long long t0 = clock64();
    for (int jb = 0; gw + (long long)(32 * jb) * n_warps < n_rows; jb += PPN_DRAIN_MAXJ) {
    unsigned pending[PPN_DRAIN_MAXJ];
    unsigned any = 0;
#pragma unroll
    for (int j = 0; j < PPN_DRAIN_MAXJ; j++) {
        pending[j] = __ballot_sync(0xffffffffu, gw + (long long)(lane + 32 * (jb + j)) * n_warps < n_rows);
        any |= pending[j];
    }
    while (any) {
        unsigned got = 0;
        any = 0;
#pragma unroll
        for (int j = 0; j < PPN_DRAIN_MAXJ; j++) {
            if (!pending[j]) continue;
            unsigned f = 0;
            if ((pending[j] >> lane) & 1u) f = *reinterpret_cast<const volatile unsigned*>(flags + gw + (long long)(lane + 32 * (jb + j)) * n_warps);
            const bool here = ((pending[j] >> lane) & 1u) && (f & 0x7fffffffu) == epoch;
            const unsigned ready = __ballot_sync(0xffffffffu, here);
            unsigned copy = __ballot_sync(0xffffffffu, here && !(f & 0x80000000u));   // rows that carry an observation
            pending[j] &= ~ready;
            any |= pending[j];
            got |= ready;
            if (ready) __threadfence();   // acquire side of the row's release store: the row is read after its flag
            while (copy) {
                const int i = __ffs(copy) - 1;
                copy &= copy - 1;
                const long long row = gw + (long long)(i + 32 * (jb + j)) * n_warps;
                const unsigned long long* src = stage + row * stage_stride8;
                unsigned long long* dst = host + row * host_stride8;
                if (v16) {
                    const ulonglong2* s2 = reinterpret_cast<const ulonglong2*>(src);
                    ulonglong2* d2 = reinterpret_cast<ulonglong2*>(dst);
                    int k = lane;
                    for (; k + 224 < n16; k += 256) {   // eight loads in flight per lane
                        ulonglong2 v[8];
#pragma unroll
                        for (int q = 0; q < 8; q++) v[q] = __ldcg(s2 + k + 32 * q);
#pragma unroll
                        for (int q = 0; q < 8; q++) d2[k + 32 * q] = v[q];
                    }
                    {   // the rest (< 256 elements) with predicated loads, still all in flight together
                        ulonglong2 v[8];
#pragma unroll
                        for (int q = 0; q < 8; q++) if (k + 32 * q < n16) v[q] = __ldcg(s2 + k + 32 * q);
#pragma unroll
                        for (int q = 0; q < 8; q++) if (k + 32 * q < n16) d2[k + 32 * q] = v[q];
                    }
                    if ((n8 & 1) && lane == 0) dst[n8 - 1] = __ldcg(src + n8 - 1);
                } else {
                    for (int k = lane; k < n8; k += 32) dst[k] = __ldcg(src + k);
                }
            }
        }
        if (any && !got) {
            if (clock64() - t0 > 4000000000ll) { if (lane == 0) atomicExch(error, 1); return; }
            __nanosleep(500);
        }
    }
    }
}

// true when p is page-locked host memory the copy engines can reach directly (cudaMallocHost / cudaHostRegister)
static bool is_pinned(const void* p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// RunEnv.step with HOST buffers.  The batch is cut into chunks; chunk i runs  actions H2D -> step kernel -> results D2H
// on its own stream, so the copies of the chunks that finished overlap the kernels of those still running.  Page-locked
// caller buffers are used in place; pageable ones go through the handle's pinned staging buffers.
static int step_host_impl(ppn_env* env, const uint8_t* act_host, double* obs_host, int64_t obs_stride, double* reward_host,
                          uint8_t* done_host, int32_t* flag_host, uint8_t* illegal_host, int auto_reset, bool f32);

extern "C" int ppn_step_host(ppn_env* env, const uint8_t* act_host, double* obs_host, int64_t obs_stride, double* reward_host,
                             uint8_t* done_host, int32_t* flag_host, uint8_t* illegal_host, int auto_reset) {
    return step_host_impl(env, act_host, obs_host, obs_stride, reward_host, done_host, flag_host, illegal_host, auto_reset, false);
}

// Same call with the observation rows delivered as float32 (obs_stride counts floats): half the bytes over PCIe.  Every
// value is the float64 one rounded once (the reference's Observation.as_array is float64; agents that feed a float32
// network convert anyway).  Page-locked result buffers only: the rows are written by the kernel itself.
extern "C" int ppn_step_host_f32(ppn_env* env, const uint8_t* act_host, float* obs_host, int64_t obs_stride, double* reward_host,
                                 uint8_t* done_host, int32_t* flag_host, uint8_t* illegal_host, int auto_reset) {
    return step_host_impl(env, act_host, reinterpret_cast<double*>(obs_host), obs_stride, reward_host, done_host, flag_host,
                          illegal_host, auto_reset, true);
}

static int step_host_impl(ppn_env* env, const uint8_t* act_host, double* obs_host, int64_t obs_stride, double* reward_host,
                          uint8_t* done_host, int32_t* flag_host, uint8_t* illegal_host, int auto_reset, bool f32) {
    int rc = check_ready(env, "ppn_step_host");
    if (rc) return rc;
    if (obs_host && obs_stride < env->OBSD) return fail(env, PPN_E_INVALID, "ppn_step_host: obs_stride smaller than the dynamic observation");
    CK(cudaSetDevice(env->device));
    rc = ensure_staging(env);
    if (rc) return rc;
    // the chunk streams do not order with the caller's streams: finish what device-pointer calls enqueued
    if (env->async_pending) CK(cudaDeviceSynchronize());
    const size_t B = env->B, iw = 1 + 2 * env->N + env->S, A = env->A, OD = env->OBSD;
    const bool act_direct = is_pinned(act_host);
    // ---- zero-copy: every output buffer is page-locked, so the step kernel writes its results straight into them over
    // PCIe as each env finishes (the transfer overlaps the envs that are still iterating); one launch, no D2H copies.
    {
        void* outs[5] = {obs_host, reward_host, done_host, flag_host, illegal_host};
        void* dev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        bool direct = f32 || getenv("PPN_HOST_STAGED") == nullptr;   // float32 rows only exist on the zero-copy path
        for (int i = 0; i < 5 && direct; i++) {
            if (!outs[i]) continue;
            if (!is_pinned(outs[i]) || cudaHostGetDevicePointer(&dev[i], outs[i], 0) != cudaSuccess) { cudaGetLastError(); direct = false; }
        }
        if (direct) {
            cudaStream_t s = env->own_stream;
            if (act_host) {
                if (!act_direct) memcpy(env->h_act, act_host, B * A);
                CK(cudaMemcpyAsync(env->d_act, act_direct ? act_host : env->h_act, B * A, cudaMemcpyHostToDevice, s));
            }
            PpnStepArgs a{};
            a.mode = PPN_MODE_STEP; a.auto_reset = auto_reset; a.act = act_host ? env->d_act : nullptr;
            a.obs = (double*)dev[0]; a.obs_stride = obs_stride; a.obs_f32 = f32 ? 1 : 0; a.reward = (double*)dev[1]; a.done = (uint8_t*)dev[2];
            a.flag = (int32_t*)dev[3]; a.illegal = (uint8_t*)dev[4];
            // rows through device memory + the drain kernel when the link could carry them within the kernel's run time only
            // if they did not hold SM slots (small rows: the IEEE-14 class); large rows keep the link busy either way
            const char* drain_env = getenv("PPN_HOST_DRAIN");   // 0 / 1 overrides the choice below
            const size_t row_bytes = OD * (f32 ? sizeof(float) : sizeof(double));
            const bool drain = dev[0] && (drain_env ? atoi(drain_env) != 0 : row_bytes <= 4096) && (row_bytes % 8) == 0 &&
                               (f32 ? (obs_stride % 2) == 0 : true);
            if (drain) {
                env->epoch = (env->epoch + 1) & 0x7fffffffu;
                if (env->epoch == 0) env->epoch = 1;
                a.obs = env->d_obs; a.obs_stride = (long long)OD; a.row_flag = env->d_row_flag; a.epoch = env->epoch;
            }
            rc = launch(env, a, s);
            if (rc) return rc;
            if (drain) {
                // AFTER the step kernel, on a stream of the highest priority: its CTAs take the first SM slots the step kernel
                // frees (about when the first rows are ready).  In this order a tool that serialises kernels (ncu, the sanitizer,
                // CUDA_LAUNCH_BLOCKING) only loses the overlap; the other order would spin until the limit there.
                const int n8 = (int)(row_bytes / 8);
                const long long ss8 = (long long)OD * (f32 ? 4 : 8) / 8, hs8 = obs_stride * (f32 ? 4 : 8) / 8;
                // 128 rows per warp, 4 warps per CTA measured best on B200 (more, or larger, drain CTAs slow the step kernel down)
                static const int rpw = []{ const char* v = getenv("PPN_DRAIN_ROWS"); int r = v ? atoi(v) : 128; return r < 1 ? 1 : r; }();
                static const int blk = []{ const char* v = getenv("PPN_DRAIN_BLOCK"); int r = v ? atoi(v) : 128; return r < 32 ? 32 : (r > 1024 ? 1024 : (r & ~31)); }();
                int warps = ((int)B + rpw - 1) / rpw;
                if (warps > 64) warps = 64;   // more drain CTAs take SM slots from the step kernel; each warp then drains in rounds
                const int wpb = blk / 32;
                ppn_obs_drain_kernel<<<(warps + wpb - 1) / wpb, blk, 0, env->drain_stream>>>(
                    reinterpret_cast<const unsigned long long*>(env->d_obs), ss8, reinterpret_cast<unsigned long long*>(dev[0]), hs8, (int)B,
                    n8, env->d_row_flag, env->epoch, warps, env->d_drain_err);
                CK(cudaGetLastError());
            }
            CK(cudaStreamSynchronize(s));
            if (drain) {
                CK(cudaStreamSynchronize(env->drain_stream));
                if (*(volatile int*)env->d_drain_err) { *env->d_drain_err = 0; return fail(env, PPN_E_CUDA, "ppn_step_host: the drain kernel timed out waiting for rows"); }
            }
            env->async_pending = false;
            return PPN_OK;
        }
    }
    if (f32) return fail(env, PPN_E_UNSUPPORTED, "ppn_step_host_f32: every result buffer must be page-locked (cudaHostRegister / pin_memory)");
    // ---- staged / chunked copies.  Rows of envs that ended without auto-reset hold no observation and must stay
    // untouched, which a plain copy cannot do: those go through the staging buffer.
    const bool obs_direct = obs_host && auto_reset && is_pinned(obs_host);
    const bool rew_direct = is_pinned(reward_host), done_direct = is_pinned(done_host), flag_direct = is_pinned(flag_host),
               ill_direct = is_pinned(illegal_host);
    if (act_host && !act_direct) memcpy(env->h_act, act_host, B * A);
    const int nc = env->n_chunks;
    for (int k = 0; k < nc; k++) {
        const size_t e0 = B * k / nc, e1 = B * (k + 1) / nc, n = e1 - e0;
        if (n == 0) continue;
        cudaStream_t s = env->chunk_stream[k];
        if (act_host)
            CK(cudaMemcpyAsync(env->d_act + e0 * A, (act_direct ? act_host : env->h_act) + e0 * A, n * A, cudaMemcpyHostToDevice, s));
        PpnStepArgs a{};
        a.mode = PPN_MODE_STEP; a.auto_reset = auto_reset; a.n_envs = (int)n; a.env_off = (int)e0;
        a.act = act_host ? env->d_act + e0 * A : nullptr;
        a.obs = obs_host ? env->d_obs + e0 * OD : nullptr; a.obs_stride = OD;
        a.reward = env->d_reward + e0 * 5; a.done = env->d_done + e0; a.flag = env->d_flag + e0;
        a.illegal = illegal_host ? env->d_ill + e0 * iw : nullptr;
        rc = launch(env, a, s);
        if (rc) return rc;
        if (obs_host) {
            if (obs_direct)
                CK(cudaMemcpy2DAsync(obs_host + e0 * obs_stride, obs_stride * sizeof(double), env->d_obs + e0 * OD, OD * sizeof(double),
                                     OD * sizeof(double), n, cudaMemcpyDeviceToHost, s));
            else
                CK(cudaMemcpyAsync(env->h_obs + e0 * OD, env->d_obs + e0 * OD, n * OD * sizeof(double), cudaMemcpyDeviceToHost, s));
        }
        CK(cudaMemcpyAsync((rew_direct ? reward_host : env->h_reward) + e0 * 5, env->d_reward + e0 * 5, n * 5 * sizeof(double), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync((done_direct ? done_host : env->h_done) + e0, env->d_done + e0, n, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync((flag_direct ? flag_host : env->h_flag) + e0, env->d_flag + e0, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        if (illegal_host)
            CK(cudaMemcpyAsync((ill_direct ? illegal_host : env->h_ill) + e0 * iw, env->d_ill + e0 * iw, n * iw, cudaMemcpyDeviceToHost, s));
    }
    for (int k = 0; k < nc; k++) CK(cudaStreamSynchronize(env->chunk_stream[k]));
    env->async_pending = false;
    if (done_host && !done_direct) memcpy(done_host, env->h_done, B);
    if (obs_host && !obs_direct) {
        const uint8_t* dn = done_direct ? done_host : env->h_done;
        for (size_t e = 0; e < B; e++)
            if (auto_reset || !dn[e])
                memcpy(obs_host + e * obs_stride, env->h_obs + e * OD, OD * sizeof(double));
    }
    if (reward_host && !rew_direct) memcpy(reward_host, env->h_reward, B * 5 * sizeof(double));
    if (flag_host && !flag_direct) memcpy(flag_host, env->h_flag, B * sizeof(int32_t));
    if (illegal_host && !ill_direct) memcpy(illegal_host, env->h_ill, B * iw);
    return PPN_OK;
}

// ------------------------------------------------------------------------------------------- peer memory (NVLink)
// Env-sharded runs: every rank's step kernel stores its packed result rows (reward[5] | done | flag,
// ppn_set_result_pack) STRAIGHT into a buffer that lives on the collecting rank's GPU, through a peer mapping of that
// buffer (CUDA IPC, NVLink / NVSwitch stores) -- compute and "gather" are one kernel, no collective sits between two
// steps.  A per-rank step counter written after each step (release, system scope) tells the collector which rows are
// complete; a "consumed" counter read back over NVLink gives the writers flow control over a ring of buffers.
#define CKD(call)                                                                                    \
    do {                                                                                             \
        cudaError_t _e = (call);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return fail(nullptr, PPN_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));    \
    } while (0)

extern "C" int ppn_peer_alloc(int device, uint64_t bytes, void** dev_out, uint8_t* handle_out) {
    if (!dev_out || !handle_out || bytes == 0) return fail(nullptr, PPN_E_INVALID, "ppn_peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    CKD(cudaSetDevice(device));
    void* p = nullptr;
    CKD(cudaMalloc(&p, bytes));
    CKD(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(nullptr, PPN_E_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
    memcpy(handle_out, &h, 64);
    *dev_out = p;
    return PPN_OK;
}

extern "C" int ppn_peer_open(int device, const uint8_t* handle, void** dev_out) {
    if (!handle || !dev_out) return fail(nullptr, PPN_E_INVALID, "ppn_peer_open: bad arguments");
    CKD(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CKD(cudaIpcOpenMemHandle(dev_out, h, cudaIpcMemLazyEnablePeerAccess));
    return PPN_OK;
}

extern "C" int ppn_peer_close(int device, void* dev_ptr) {
    CKD(cudaSetDevice(device));
    CKD(cudaIpcCloseMemHandle(dev_ptr));
    return PPN_OK;
}

extern "C" int ppn_peer_free(int device, void* dev_ptr) {
    CKD(cudaSetDevice(device));
    CKD(cudaFree(dev_ptr));
    return PPN_OK;
}

__global__ void ppn_peer_signal_kernel(unsigned long long* flag, unsigned long long value) {
    __threadfence_system();   // everything this stream wrote before (the step kernel's rows) is visible first
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(value) : "memory");
}

__global__ void ppn_peer_wait_kernel(const unsigned long long* flags, int n, unsigned long long value) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        // relaxed polls with a long sleep, one acquire fence at the end: the warp shares its SM with step CTAs
        unsigned long long v;
        while (true) {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + i) : "memory");
            if (v >= value) break;
            __nanosleep(4000);
        }
    }
    __threadfence_system();
}

extern "C" int ppn_peer_signal(int device, uint64_t* flag_dev, uint64_t value, void* stream) {
    if (!flag_dev) return fail(nullptr, PPN_E_INVALID, "ppn_peer_signal: null flag");
    CKD(cudaSetDevice(device));
    ppn_peer_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)flag_dev, (unsigned long long)value);
    CKD(cudaGetLastError());
    return PPN_OK;
}

extern "C" int ppn_peer_wait(int device, const uint64_t* flags_dev, int n, uint64_t value, void* stream) {
    if (!flags_dev || n <= 0) return fail(nullptr, PPN_E_INVALID, "ppn_peer_wait: bad arguments");
    CKD(cudaSetDevice(device));
    ppn_peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const unsigned long long*)flags_dev, n, (unsigned long long)value);
    CKD(cudaGetLastError());
    return PPN_OK;
}

extern "C" int ppn_peer_read(int device, void* host_dst, const void* dev_src, uint64_t bytes, void* stream) {
    if (!host_dst || !dev_src) return fail(nullptr, PPN_E_INVALID, "ppn_peer_read: null buffer");
    CKD(cudaSetDevice(device));
    CKD(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return PPN_OK;
}

extern "C" int ppn_state_width(const ppn_env* env, int field) {
    if (!env) return PPN_E_INVALID;
    switch (field) {
        case PPN_STATE_REAL: return env->st.rw;
        case PPN_STATE_TOPOLOGY: return 2 * env->G + env->L + 3 * env->N;
        case PPN_STATE_COUNTERS: return 3 * env->N + env->S + 4;
        default: return PPN_E_INVALID;
    }
}

static int copy_state(ppn_env* env, int field, void* user, bool to_user, cudaStream_t s) {
    if (!env || !user) return fail(env, PPN_E_INVALID, "ppn_get/set_state: null argument");
    CK(cudaSetDevice(env->device));
    const int w = ppn_state_width(env, field);
    if (w < 0) return fail(env, PPN_E_INVALID, "ppn_get/set_state: unknown field");
    size_t esz, stride; char* base;
    if (field == PPN_STATE_REAL) { esz = 8; stride = env->st.rw; base = (char*)env->st.real; }
    else if (field == PPN_STATE_TOPOLOGY) { esz = 1; stride = env->st.tw; base = (char*)env->st.topo; }
    else { esz = 4; stride = env->st.cw; base = (char*)env->st.cnt; }
    if (to_user) CK(cudaMemcpy2DAsync(user, w * esz, base, stride * esz, w * esz, env->B, cudaMemcpyDeviceToDevice, s));
    else CK(cudaMemcpy2DAsync(base, stride * esz, user, w * esz, w * esz, env->B, cudaMemcpyDeviceToDevice, s));
    return PPN_OK;
}

extern "C" int ppn_get_state(ppn_env* env, int field, void* out_dev, void* stream) {
    return copy_state(env, field, out_dev, true, (cudaStream_t)stream);
}
extern "C" int ppn_set_state(ppn_env* env, int field, const void* in_dev, void* stream) {
    return copy_state(env, field, const_cast<void*>(in_dev), false, (cudaStream_t)stream);
}

extern "C" int ppn_observation_static(ppn_env* env, double* out_host) {
    if (!env || !out_host) return fail(env, PPN_E_INVALID, "ppn_observation_static: null argument");
    memcpy(out_host, env->obs_static.data(), env->obs_static.size() * sizeof(double));
    return PPN_OK;
}

extern "C" int ppn_n_envs(const ppn_env* env) { return env ? env->B : PPN_E_INVALID; }
extern "C" int ppn_action_length(const ppn_env* env) { return env ? env->A : PPN_E_INVALID; }
extern "C" int ppn_obs_length(const ppn_env* env) { return env ? env->OBS : PPN_E_INVALID; }
extern "C" int ppn_obs_dynamic_length(const ppn_env* env) { return env ? env->OBSD : PPN_E_INVALID; }
extern "C" int ppn_device(const ppn_env* env) { return env ? env->device : PPN_E_INVALID; }

extern "C" int ppn_get_cascade_histogram(ppn_env* env, int64_t* out_host) {
    if (!env || !out_host) return fail(env, PPN_E_INVALID, "ppn_get_cascade_histogram: null argument");
    CK(cudaSetDevice(env->device));
    unsigned long long h[8];
    CK(cudaMemcpy(h, env->stats + 8, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 8; i++) out_host[i] = (int64_t)h[i];
    return PPN_OK;
}

extern "C" int ppn_get_counters(ppn_env* env, int64_t* out_host) {
    if (!env || !out_host) return fail(env, PPN_E_INVALID, "ppn_get_counters: null argument");
    CK(cudaSetDevice(env->device));
    unsigned long long h[8];
    CK(cudaMemcpy(h, env->stats, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 5; i++) out_host[i] = (int64_t)h[i];
    out_host[5] = env->launches;
    out_host[6] = env->env_smem_bytes;
    out_host[7] = env->tpe;
    out_host[8] = (int64_t)h[5];
    out_host[9] = (int64_t)h[6];
    // the two maxima restart after every read
    CK(cudaMemset(env->stats + 5, 0, 2 * sizeof(unsigned long long)));
    return PPN_OK;
}
