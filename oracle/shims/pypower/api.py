"""TEST INFRASTRUCTURE ONLY (oracle/) -- never imported by the product package `pypownet_b200`.

Re-statement, from the published algorithm, of the slice of PYPOWER 5.1.4 (the un-vendored third-party
dependency pinned in the reference's requirements.txt:9) that pypownet reaches:

    pypownet/grid.py:63-64   ppoption(PF_ALG=2, PF_MAX_IT_FD=25, PF_TOL=1e-6, VERBOSE=0, OUT_ALL=0)
    pypownet/grid.py:65      loadcase(path, expect_gencost=False)
    pypownet/grid.py:227-229 rundcpf / runpf (mpc, ppopt, pprint_fname, case_fname) -> (results, success)
    pypownet/grid.py:595     savecase(path, mpc)

With this module and the `gym` shim on sys.path the UNMODIFIED reference package (/root/reference/pypownet)
runs in this container; that stack is what generated tests/golden/* and what pins oracle/flat.py.
Operation order follows PYPOWER: ext2int -> bustypes -> V0 -> makeYbus/makeSbus -> makeB + fdpf (PF_ALG 2/3),
newtonpf (PF_ALG 1) or makeBdc + dcpf (PF_DC) -> pfsoln -> int2ext -> zero out-of-service result fields.
Linear algebra is scipy.sparse (splu / spsolve) exactly like PYPOWER, so failure modes are the same ones the
reference catches at grid.py:230 (RuntimeError "Factor is exactly singular", IndexError from an empty PV list,
ValueError from the infinity norm of an empty mismatch vector) or detects at grid.py:103-110 (NaN).

Pinning: the reference's own 26 tests (known-answer values at tests/test_core.py:351-372, 551-603, 917-976)
pass on top of this module (tests/golden/REFERENCE_TESTS.txt records the run).  "PYPOWER itself" is absent,
so parity below 1e-3 MW rests on this re-statement; DESIGN.md says so.
"""
from copy import deepcopy
import warnings

import numpy as np
from numpy import array, zeros, ones, exp, pi, conj, angle, r_, flatnonzero as find  # noqa: F401
from scipy.sparse import csr_matrix as sparse
from scipy.sparse.linalg import splu, spsolve

# --- column indices (MATPOWER case format v2) -------------------------------------------------------------------
BUS_I, BUS_TYPE, PD, QD, GS, BS, BUS_AREA, VM, VA, BASE_KV, ZONE, VMAX, VMIN = range(13)
PQ, PV, REF, NONE = 1, 2, 3, 4
GEN_BUS, PG, QG, QMAX, QMIN, VG, MBASE, GEN_STATUS, PMAX, PMIN = range(10)
F_BUS, T_BUS, BR_R, BR_X, BR_B, RATE_A, RATE_B, RATE_C, TAP, SHIFT, BR_STATUS, ANGMIN, ANGMAX, PF, QF, PT, QT = \
    range(17)
EPS = np.finfo(float).eps


def ppoption(ppopt=None, **kw):
    opt = {'PF_ALG': 1, 'PF_TOL': 1e-8, 'PF_MAX_IT': 10, 'PF_MAX_IT_FD': 30, 'PF_MAX_IT_GS': 1000,
           'ENFORCE_Q_LIMS': False, 'PF_DC': False, 'VERBOSE': 1, 'OUT_ALL': -1}
    if ppopt is not None:
        opt.update(ppopt)
    opt.update(kw)
    # test harness only: tools/make_golden.py records Newton-Raphson fixtures by forcing the algorithm the unmodified
    # reference asks for (it hard-codes PF_ALG=2, grid.py:63)
    import os
    if os.environ.get('PYPOWNET_SHIM_PF_ALG'):
        opt['PF_ALG'] = int(os.environ['PYPOWNET_SHIM_PF_ALG'])
    return opt


def loadcase(casefile, return_as_obj=True, expect_gencost=True, expect_areas=True):
    """A dict is deep-copied; a path to a .py case file is exec'd with numpy's `array` in scope and its single
    function called (the reference ships `reference_grid.py` defining `reference_grid()`)."""
    if isinstance(casefile, dict):
        ppc = deepcopy(casefile)
    else:
        path = casefile if str(casefile).endswith('.py') else str(casefile) + '.py'
        scope = {'array': np.array, 'np': np}
        with open(path) as f:
            exec(compile(f.read(), path, 'exec'), scope)
        funcs = [v for v in scope.values() if hasattr(v, '__code__') and v.__code__.co_filename == path]
        ppc = funcs[-1]()
    for k in ('bus', 'gen', 'branch'):
        ppc[k] = np.array(ppc[k], dtype=float)
    return ppc


def savecase(fname, ppc, *a, **kw):
    np.set_printoptions(threshold=10 ** 9)
    with open(fname if str(fname).endswith('.py') else str(fname) + '.py', 'w') as f:
        f.write('from numpy import array\n\ndef reference_grid():\n    ppc = {"version": "2"}\n')
        f.write('    ppc["baseMVA"] = %r\n' % float(ppc['baseMVA']))
        for k in ('bus', 'gen', 'branch'):
            f.write('    ppc[%r] = array(%s)\n' % (k, np.asarray(ppc[k]).tolist()))
        f.write('    return ppc\n')
    return fname


# --- ext2int / int2ext ------------------------------------------------------------------------------------------
def _ext2int(ppc):
    bus, gen, branch = ppc['bus'], ppc['gen'], ppc['branch']
    nb = bus.shape[0]
    o = {'ext': {'bus': bus.copy(), 'gen': gen.copy(), 'branch': branch.copy()}}
    maxb = int(bus[:, BUS_I].max())
    n2i = zeros(maxb + 1, dtype=int)          # PYPOWER builds a (maxb+1)-long lookup, 666119 entries for case118
    n2i[bus[:, BUS_I].astype(int)] = np.arange(nb)
    bs = bus[:, BUS_TYPE] != NONE
    gs = (gen[:, GEN_STATUS] > 0) & bs[n2i[gen[:, GEN_BUS].astype(int)]]
    brs = (branch[:, BR_STATUS].astype(int) & bs[n2i[branch[:, F_BUS].astype(int)]]
           & bs[n2i[branch[:, T_BUS].astype(int)]]).astype(bool)
    o['bus_on'], o['bus_off'] = find(bs), find(~bs)
    o['gen_on'], o['gen_off'] = find(gs), find(~gs)
    o['br_on'], o['br_off'] = find(brs), find(~brs)
    bus, gen, branch = bus[bs].copy(), gen[gs].copy(), branch[brs].copy()
    # consecutive internal numbering in bus-array order
    o['i2e'] = bus[:, BUS_I].copy()
    e2i = zeros(maxb + 1, dtype=int)
    e2i[o['i2e'].astype(int)] = np.arange(bus.shape[0])
    bus[:, BUS_I] = e2i[bus[:, BUS_I].astype(int)]
    gen[:, GEN_BUS] = e2i[gen[:, GEN_BUS].astype(int)]
    branch[:, F_BUS] = e2i[branch[:, F_BUS].astype(int)]
    branch[:, T_BUS] = e2i[branch[:, T_BUS].astype(int)]
    # gens sorted by internal bus (stable)
    o['gen_e2i'] = np.argsort(gen[:, GEN_BUS], kind='stable')
    o['gen_i2e'] = np.argsort(o['gen_e2i'], kind='stable')
    gen = gen[o['gen_e2i']]
    return bus, gen, branch, o


def _int2ext(bus, gen, branch, o):
    ebus, egen, ebr = o['ext']['bus'].copy(), o['ext']['gen'].copy(), o['ext']['branch'].copy()
    if ebr.shape[1] < branch.shape[1]:
        ebr = np.hstack([ebr, zeros((ebr.shape[0], branch.shape[1] - ebr.shape[1]))])
    ebus[o['bus_on'], :] = bus
    ebr[o['br_on'], :] = branch
    egen[o['gen_on'], :] = gen[o['gen_i2e'], :]
    ebus[o['bus_on'], BUS_I] = o['i2e'][ebus[o['bus_on'], BUS_I].astype(int)]
    ebr[o['br_on'], F_BUS] = o['i2e'][ebr[o['br_on'], F_BUS].astype(int)]
    ebr[o['br_on'], T_BUS] = o['i2e'][ebr[o['br_on'], T_BUS].astype(int)]
    egen[o['gen_on'], GEN_BUS] = o['i2e'][egen[o['gen_on'], GEN_BUS].astype(int)]
    return ebus, egen, ebr


def _bustypes(bus, gen):
    nb, ng = bus.shape[0], gen.shape[0]
    Cg = sparse((gen[:, GEN_STATUS] > 0, (gen[:, GEN_BUS].astype(int), range(ng))), (nb, ng))
    bus_gen_status = (Cg * ones(ng, int)).astype(bool)
    ref = find((bus[:, BUS_TYPE] == REF) & bus_gen_status)
    pv = find((bus[:, BUS_TYPE] == PV) & bus_gen_status)
    pq = find((bus[:, BUS_TYPE] == PQ) | ~bus_gen_status)
    if len(ref) == 0:
        ref = zeros(1, dtype=int)
        ref[0] = pv[0]                 # IndexError when there is no PV bus either (caught at grid.py:230)
        pv = pv[1:]
    return ref, pv, pq


def _makeYbus(baseMVA, bus, branch):
    nb, nl = bus.shape[0], branch.shape[0]
    stat = branch[:, BR_STATUS]
    Ys = stat / (branch[:, BR_R] + 1j * branch[:, BR_X])
    Bc = stat * branch[:, BR_B]
    tap = ones(nl)
    i = find(branch[:, TAP])
    tap[i] = branch[i, TAP]
    tap = tap * exp(1j * pi / 180 * branch[:, SHIFT])
    Ytt = Ys + 1j * Bc / 2
    Yff = Ytt / (tap * conj(tap))
    Yft = -Ys / conj(tap)
    Ytf = -Ys / tap
    Ysh = (bus[:, GS] + 1j * bus[:, BS]) / baseMVA
    f = branch[:, F_BUS].astype(int)
    t = branch[:, T_BUS].astype(int)
    Cf = sparse((ones(nl), (range(nl), f)), (nl, nb))
    Ct = sparse((ones(nl), (range(nl), t)), (nl, nb))
    i = r_[range(nl), range(nl)]
    Yf = sparse((r_[Yff, Yft], (i, r_[f, t])), (nl, nb))
    Yt = sparse((r_[Ytf, Ytt], (i, r_[f, t])), (nl, nb))
    Ybus = Cf.T * Yf + Ct.T * Yt + sparse((Ysh, (range(nb), range(nb))), (nb, nb))
    return Ybus, Yf, Yt


def _makeSbus(baseMVA, bus, gen):
    on = find(gen[:, GEN_STATUS] > 0)
    gbus = gen[on, GEN_BUS].astype(int)
    nb, ngon = bus.shape[0], on.shape[0]
    Cg = sparse((ones(ngon), (gbus, range(ngon))), (nb, ngon))
    return (Cg * (gen[on, PG] + 1j * gen[on, QG]) - (bus[:, PD] + 1j * bus[:, QD])) / baseMVA


def _makeB(baseMVA, bus, branch, alg):
    nb, nl = bus.shape[0], branch.shape[0]
    tbus, tbr = bus.copy(), branch.copy()
    tbus[:, BS] = zeros(nb)
    tbr[:, BR_B] = zeros(nl)
    tbr[:, TAP] = ones(nl)
    if alg == 2:
        tbr[:, BR_R] = zeros(nl)
    Bp = -1 * _makeYbus(baseMVA, tbus, tbr)[0].imag
    tbr = branch.copy()
    tbr[:, SHIFT] = zeros(nl)
    if alg == 3:
        tbr[:, BR_R] = zeros(nl)
    Bpp = -1 * _makeYbus(baseMVA, bus, tbr)[0].imag
    return Bp, Bpp


def _fdpf(Ybus, Sbus, V0, Bp, Bpp, ref, pv, pq, ppopt):
    tol, max_it = ppopt['PF_TOL'], ppopt['PF_MAX_IT_FD']
    converged, i = 0, 0
    V = V0
    Va, Vm = angle(V), abs(V)
    pvpq = r_[pv, pq]
    mis = (V * conj(Ybus * V) - Sbus) / Vm
    P, Q = mis[pvpq].real, mis[pq].imag
    normP, normQ = np.linalg.norm(P, np.inf), np.linalg.norm(Q, np.inf)   # ValueError on an empty vector
    if normP < tol and normQ < tol:
        converged = 1
    Bp = Bp[array([pvpq]).T, pvpq].tocsc()
    Bpp = Bpp[array([pq]).T, pq].tocsc()
    Bp_solver, Bpp_solver = splu(Bp), splu(Bpp)                         # RuntimeError when exactly singular
    while (not converged) and i < max_it:
        i += 1
        dVa = -Bp_solver.solve(P)
        Va[pvpq] = Va[pvpq] + dVa
        V = Vm * exp(1j * Va)
        mis = (V * conj(Ybus * V) - Sbus) / Vm
        P, Q = mis[pvpq].real, mis[pq].imag
        normP, normQ = np.linalg.norm(P, np.inf), np.linalg.norm(Q, np.inf)
        if normP < tol and normQ < tol:
            converged = 1
            break
        dVm = -Bpp_solver.solve(Q)
        Vm[pq] = Vm[pq] + dVm
        V = Vm * exp(1j * Va)
        mis = (V * conj(Ybus * V) - Sbus) / Vm
        P, Q = mis[pvpq].real, mis[pq].imag
        normP, normQ = np.linalg.norm(P, np.inf), np.linalg.norm(Q, np.inf)
        if normP < tol and normQ < tol:
            converged = 1
            break
    return V, converged, i


def _dSbus_dV(Ybus, V):
    ib = range(len(V))
    Ibus = Ybus * V
    diagV = sparse((V, (ib, ib)))
    diagIbus = sparse((Ibus, (ib, ib)))
    diagVnorm = sparse((V / abs(V), (ib, ib)))
    dS_dVm = diagV * conj(Ybus * diagVnorm) + conj(diagIbus) * diagVnorm
    dS_dVa = 1j * diagV * conj(diagIbus - Ybus * diagV)
    return dS_dVm, dS_dVa


def _newtonpf(Ybus, Sbus, V0, ref, pv, pq, ppopt):
    tol, max_it = ppopt['PF_TOL'], ppopt['PF_MAX_IT']
    converged, i = 0, 0
    V = V0
    Va, Vm = angle(V), abs(V)
    pvpq = r_[pv, pq]
    npv, npq = len(pv), len(pq)
    j1, j2, j3, j4, j5, j6 = 0, npv, npv, npv + npq, npv + npq, npv + 2 * npq
    mis = V * conj(Ybus * V) - Sbus
    F = r_[mis[pv].real, mis[pq].real, mis[pq].imag]
    if np.linalg.norm(F, np.inf) < tol:
        converged = 1
    while (not converged) and i < max_it:
        i += 1
        dS_dVm, dS_dVa = _dSbus_dV(Ybus, V)
        J11 = dS_dVa[array([pvpq]).T, pvpq].real
        J12 = dS_dVm[array([pvpq]).T, pq].real
        J21 = dS_dVa[array([pq]).T, pvpq].imag
        J22 = dS_dVm[array([pq]).T, pq].imag
        from scipy.sparse import hstack, vstack
        J = vstack([hstack([J11, J12]), hstack([J21, J22])], format='csr')
        dx = -1 * spsolve(J, F)
        if npv:
            Va[pv] = Va[pv] + dx[j1:j2]
        if npq:
            Va[pq] = Va[pq] + dx[j3:j4]
            Vm[pq] = Vm[pq] + dx[j5:j6]
        V = Vm * exp(1j * Va)
        Vm, Va = abs(V), angle(V)
        mis = V * conj(Ybus * V) - Sbus
        F = r_[mis[pv].real, mis[pq].real, mis[pq].imag]
        if np.linalg.norm(F, np.inf) < tol:
            converged = 1
    return V, converged, i


def _pfsoln(baseMVA, bus0, gen0, branch0, Ybus, Yf, Yt, V, ref, pv, pq):
    bus, gen, branch = bus0.copy(), gen0.copy(), branch0.copy()
    bus[:, VM] = abs(V)
    bus[:, VA] = angle(V) * 180 / pi
    on = find(gen[:, GEN_STATUS] > 0)
    gbus = gen[on, GEN_BUS].astype(int)
    Sbus = V[gbus] * conj(Ybus[gbus, :] * V)
    gen[:, QG] = zeros(gen.shape[0])
    gen[on, QG] = Sbus.imag * baseMVA + bus[gbus, QD]
    if len(on) > 1:
        nb, ngon = bus.shape[0], on.shape[0]
        Cg = sparse((ones(ngon), (range(ngon), gbus)), (ngon, nb))
        ngg = np.asarray(Cg * Cg.sum(0).T).flatten()
        gen[on, QG] = gen[on, QG] / ngg
        Cmin = sparse((gen[on, QMIN], (range(ngon), gbus)), (ngon, nb))
        Cmax = sparse((gen[on, QMAX], (range(ngon), gbus)), (ngon, nb))
        Qg_tot = Cg.T * gen[on, QG]
        Qg_min = np.asarray(Cmin.sum(0).T).flatten()
        Qg_max = np.asarray(Cmax.sum(0).T).flatten()
        ig = find(Cg * Qg_min == Cg * Qg_max)
        Qg_save = gen[on[ig], QG]
        gen[on, QG] = gen[on, QMIN] + (Cg * ((Qg_tot - Qg_min) / (Qg_max - Qg_min + EPS))) * \
            (gen[on, QMAX] - gen[on, QMIN])
        gen[on[ig], QG] = Qg_save
    for k in range(len(ref)):
        temp = find(gbus == ref[k])
        gen[on[temp[0]], PG] = Sbus[temp[0]].real * baseMVA + bus[ref[k], PD]
        if len(temp) > 1:
            gen[on[temp[0]], PG] = gen[on[temp[0]], PG] - np.sum(gen[on[temp[1:]], PG])
    out = find(branch[:, BR_STATUS] == 0)
    br = find(branch[:, BR_STATUS]).astype(int)
    Sf = V[branch[br, F_BUS].astype(int)] * conj(Yf[br, :] * V) * baseMVA
    St = V[branch[br, T_BUS].astype(int)] * conj(Yt[br, :] * V) * baseMVA
    branch[np.ix_(br, [PF, QF, PT, QT])] = np.c_[Sf.real, Sf.imag, St.real, St.imag]
    branch[np.ix_(out, [PF, QF, PT, QT])] = zeros((len(out), 4))
    return bus, gen, branch


def _makeBdc(baseMVA, bus, branch):
    nb, nl = bus.shape[0], branch.shape[0]
    stat = branch[:, BR_STATUS]
    b = stat / branch[:, BR_X]
    tap = ones(nl)
    i = find(branch[:, TAP])
    tap[i] = branch[i, TAP]
    b = b / tap
    f = branch[:, F_BUS].astype(int)
    t = branch[:, T_BUS].astype(int)
    i = r_[range(nl), range(nl)]
    Cft = sparse((r_[ones(nl), -ones(nl)], (i, r_[f, t])), (nl, nb))
    Bf = sparse((r_[b, -b], (i, r_[f, t])), (nl, nb))
    Bbus = Cft.T * Bf
    Pfinj = b * (-branch[:, SHIFT] * pi / 180)
    Pbusinj = Cft.T * Pfinj
    return Bbus, Bf, Pbusinj, Pfinj


def _dcpf(B, Pbus, Va0, ref, pv, pq):
    pvpq = r_[pv, pq]
    Va = np.copy(Va0)
    rhs = Pbus[pvpq] - B[pvpq][:, ref] * Va0[ref]
    Va[pvpq] = spsolve(B[pvpq][:, pvpq].tocsc(), rhs)        # singular -> MatrixRankWarning + NaN, as PYPOWER
    return Va


def runpf(casedata=None, ppopt=None, fname='', solvedcase=''):
    ppopt = ppoption(ppopt)
    dc = ppopt['PF_DC']
    ppc = loadcase(casedata)
    if ppc['branch'].shape[1] < QT + 1:
        ppc['branch'] = np.c_[ppc['branch'], zeros((ppc['branch'].shape[0], QT + 1 - ppc['branch'].shape[1]))]
    baseMVA = ppc['baseMVA']
    bus, gen, branch, order = _ext2int(ppc)
    ref, pv, pq = _bustypes(bus, gen)
    on = find(gen[:, GEN_STATUS] > 0)
    gbus = gen[on, GEN_BUS].astype(int)
    if dc:
        Va0 = bus[:, VA] * (pi / 180)
        B, Bf, Pbusinj, Pfinj = _makeBdc(baseMVA, bus, branch)
        Pbus = _makeSbus(baseMVA, bus, gen).real - Pbusinj - bus[:, GS] / baseMVA
        Va = _dcpf(B, Pbus, Va0, ref, pv, pq)
        branch[:, [QF, QT]] = zeros((branch.shape[0], 2))
        branch[:, PF] = (Bf * Va + Pfinj) * baseMVA
        branch[:, PT] = -branch[:, PF]
        bus[:, VM] = ones(bus.shape[0])
        bus[:, VA] = Va * (180 / pi)
        refgen = zeros(len(ref), dtype=int)
        for k in range(len(ref)):
            temp = find(gbus == ref[k])
            refgen[k] = on[temp[0]]
        gen[refgen, PG] = gen[refgen, PG] + (B[ref, :] * Va - Pbus[ref]) * baseMVA
        success = 1
    else:
        V0 = bus[:, VM] * exp(1j * pi / 180 * bus[:, VA])
        V0[gbus] = gen[on, VG] / abs(V0[gbus]) * V0[gbus]
        Ybus, Yf, Yt = _makeYbus(baseMVA, bus, branch)
        Sbus = _makeSbus(baseMVA, bus, gen)
        alg = ppopt['PF_ALG']
        if alg == 1:
            V, success, _ = _newtonpf(Ybus, Sbus, V0, ref, pv, pq, ppopt)
        elif alg in (2, 3):
            Bp, Bpp = _makeB(baseMVA, bus, branch, alg)
            V, success, _ = _fdpf(Ybus, Sbus, V0, Bp, Bpp, ref, pv, pq, ppopt)
        else:
            raise ValueError('Only Newton (1) and fast-decoupled (2, 3) power flow are restated')
        bus, gen, branch = _pfsoln(baseMVA, bus, gen, branch, Ybus, Yf, Yt, V, ref, pv, pq)
    ebus, egen, ebr = _int2ext(bus, gen, branch, order)
    if len(order['gen_off']) > 0:
        egen[np.ix_(order['gen_off'], [PG, QG])] = 0
    if len(order['br_off']) > 0:
        ebr[np.ix_(order['br_off'], [PF, QF, PT, QT])] = 0
    results = {k: v for k, v in ppc.items() if k not in ('bus', 'gen', 'branch')}
    results.update({'bus': ebus, 'gen': egen, 'branch': ebr, 'success': success, 'et': 0.0,
                    'order': {'state': 'e'}})
    return results, success


def rundcpf(casedata=None, ppopt=None, fname='', solvedcase=''):
    ppopt = ppoption(ppopt, PF_DC=True)
    return runpf(casedata, ppopt, fname, solvedcase)
