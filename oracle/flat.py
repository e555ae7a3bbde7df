"""TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of the reference's per-timestep hot path.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this file; the product package
`pypownet_b200` never does (its compute path is the CUDA library and fails loudly without it).

What is restated (reference file:line in each method):
  pypownet/game.py:799-885   Game.step            -> FlatEnv.step
  pypownet/game.py:650-753   _verify_illegal_action, :591-648 apply_action, :1088-1100 get_changed_substations
  pypownet/game.py:405-501   load_entries_from_next_timestep / _timestep_id
  pypownet/game.py:503-589   _compute_loadflow_cascading
  pypownet/game.py:762-797   process_game_over / reset_grid ; :887-943 simulate
  pypownet/grid.py:140-210   _synchronize_bus_types / _count_isolated_loads ; :244-264 compute_loadflow
  pypownet/grid.py:112-138   extract_flows_a ; :266-311 injections ; :360-423 apply_topology
  pypownet/grid.py:496-566 + game.py:945-978 + environment.py:451-595   observation export / as_array order
  parameters/default14/reward_signal.py:45-169   the shipped five-term reward
  PYPOWER 5.1.4 runpf/rundcpf (third-party, un-vendored; restated in oracle/shims/pypower/api.py and SURVEY.md
  Appendix A): ext2int, bustypes, makeYbus, makeSbus, makeB (XB), fdpf, pfsoln, int2ext, makeBdc, dcpf.

State is the struct-of-arrays model the CUDA kernels use (2S buses, node bits, line status, counters, cursor)
instead of the reference's `mpc` dict, so this file pins the DATA MODEL as well as the arithmetic.  Linear
algebra is dense LAPACK (the reference's is SuperLU through PYPOWER): results agree to rounding.

Pinning: tests/test_oracle_golden.py checks this file against tests/golden/*.npz, which were produced by the
UNMODIFIED reference package running on oracle/shims (tools/make_golden.py), including the known-answer values of
the reference's own tests (slack Pg 123.370285 MW at tests/test_core.py:351-372, the 15-step line-6 ampere
sequence at tests/test_core.py:917-934).  PYPOWER itself is not installable here, so agreement finer than the
reference's own test tolerances (1e-3 MW) is agreement with the restated PYPOWER, not with the original wheel.

One deliberate, documented normalisation: a load-flow on a grid with a multi-bus component that does not contain
the reference bus ("not connexe") is reported as diverging by an explicit connectivity test.  In exact arithmetic
B' of such a component is singular; the reference hands it to SuperLU (scipy splu through PYPOWER), which either
reports "Factor is exactly singular" (-> grid.py:230 -> DivergingLoadflowException 'The grid is not connexe') or
leaves a rounding-sized last pivot (~1e-16), in which case the component's common-mode update is (its net
mismatch)/1e-16 and the iteration fails to converge unless the pocket carries no injection at all.  Which of the
two happens depends on whether fl(b*fl(1/b)) == 1 for the pocket's line susceptances, i.e. on rounding inside
SuperLU, not on the physics; so step() returns DivergingLoadflowException for every such grid, which is what the
reference's own message says it means.  Measured deviation: see DESIGN.md ("floating pockets").
"""
import numpy as np
import scipy.linalg

FLAG_NONE, FLAG_ILLEGAL, FLAG_DIVERGING, FLAG_LOADS_CUT, FLAG_PRODS_CUT = 0, 1, 2, 3, 4
SQRT3 = 3. ** .5


class Config(object):
    """Scalars of configuration.yaml + RunEnv arguments (game.py:263-298)."""

    def __init__(self, cfg, game_over_mode='soft', without_overflow_cutoff=False, loop_mode='natural',
                 reward_constant=None, n_sub=None):
        self.dc = str(cfg['loadflow_mode']).lower() == 'dc'
        self.hard_coef = float(cfg['hard_overflow_coefficient'])
        self.n_hard_broken = int(cfg['n_timesteps_hard_overflow_is_broken'])
        self.n_soft_consec = float(cfg['n_timesteps_consecutive_soft_overflow_breaks'])
        self.n_soft_broken = int(cfg['n_timesteps_soft_overflow_is_broken'])
        if without_overflow_cutoff:                                   # game.py:268-275
            self.hard_coef = 1e9
            self.n_soft_consec = 1e12
        self.horizon = int(cfg['n_timesteps_horizon_maintenance'])
        self.max_prods_go = int(cfg['max_number_prods_game_over'])
        self.max_loads_go = int(cfg['max_number_loads_game_over'])
        self.n_line_react = int(cfg['n_timesteps_actionned_line_reactionable'])
        self.n_node_react = int(cfg['n_timesteps_actionned_node_reactionable'])
        self.max_sub = int(cfg['max_number_actionned_substations'])
        self.max_lines = int(cfg['max_number_actionned_lines'])
        self.max_total = int(cfg['max_number_actionned_total'])
        self.hard_mode = game_over_mode == 'hard'
        self.loop_mode = loop_mode
        self.tol, self.max_it = 1e-6, 25                               # grid.py:63
        # PF_ALG: 2 = fast-decoupled XB (what the reference runs, grid.py:63); 1 = Newton-Raphson (north star's named
        # variant; PYPOWER newtonpf, PF_MAX_IT = 10) -- selected by an extra `pf_alg` key, absent from the shipped configs
        self.pf_alg = int(cfg.get('pf_alg', 2))
        self.max_it_nr = 10
        self.reward_constant = float(reward_constant if reward_constant is not None else (n_sub or 0))


class FlatEnv(object):
    def __init__(self, case, config, chronics, start_id=0, thermal_limits=None, start_row=0):
        self.c, self.cfg, self.chronics = case, config, chronics
        c = case
        S = c.n_sub
        self.S, self.G, self.L, self.N = S, c.n_gen, c.n_load, c.n_line
        # static line admittances (makeYbus)
        Ys = 1.0 / (c.line_r + 1j * c.line_x)
        tap = c.line_tap.astype(complex)
        self.Ytt = Ys + 1j * c.line_b / 2
        self.Yff = self.Ytt / (tap * np.conj(tap))
        self.Yft = -Ys / np.conj(tap)
        self.Ytf = -Ys / tap
        self.Ysh = (c.bus_gs + 1j * c.bus_bs) / c.base_mva
        # B' (XB: r=0, b=0, tap=1, no shunts), B'' (everything, shift=0), Bdc
        self.bp = 1.0 / c.line_x
        self.bdc = 1.0 / c.line_x / c.line_tap
        # thermal limits: the first chronic's imaps, never refreshed (game.py:301-304)
        self.next_chronic = start_id
        self.chronic_id = self._take_next_chronic()
        self.thermal = np.asarray(chronics[self.chronic_id].imaps if thermal_limits is None else thermal_limits,
                                  dtype=np.float64).copy()
        self.planned_maint = {}
        # dynamic state
        self.gen_node = np.zeros(self.G, dtype=np.int64)
        self.load_node = np.zeros(self.L, dtype=np.int64)
        self.or_node = np.zeros(self.N, dtype=np.int64)
        self.ex_node = np.zeros(self.N, dtype=np.int64)
        self.status = c.line_status0.astype(np.int64).copy()
        self.vm, self.va = c.bus_vm0.copy(), c.bus_va0.copy()          # va in degrees (bus[:, VA])
        self.pd, self.qd = c.bus_pd0.copy(), c.bus_qd0.copy()
        self.gen_pg, self.gen_qg, self.gen_vg = c.gen_pg0.copy(), c.gen_qg0.copy(), c.gen_vg0.copy()
        self.gen_status = np.ones(self.G, dtype=np.int64)
        self.flows = np.zeros((self.N, 4))                              # Pf Qf Pt Qt
        self.t_reconnectable = np.zeros(self.N)
        self.t_line_react = np.zeros(self.N)
        self.t_node_react = np.zeros(S)
        self.soft_count = np.zeros(self.N)
        # index in the current chronic (None: none yet).  start_row > 0 is the batched harness' way of starting
        # envs at different offsets: as if row start_row-1 had just been played on the pristine grid
        self.row = None if start_row == 0 else start_row - 1
        self.entries_row = None                                         # row of `current_timestep_entries`
        self.entries_chronic = None
        self.last_depth = 0
        self.n_loadflows = 0
        # Game.__init__: first row + cascade (game.py:339-340); divergence here raises in the reference.  A batch
        # cannot raise per env: it runs process_game_over instead (include/pypownet_b200.h, ppn_reset)
        self._load_next_timestep(False)
        self.init_diverged = self._cascade()
        if self.init_diverged:
            self.process_game_over()

    # ----------------------------------------------------------------------------------------------- chronic cursor
    def _take_next_chronic(self):
        """ChronicLooper.get_next_chronic_folder (chronic.py:282-291); 'random' draws from np.random like the
        reference."""
        cur = self.next_chronic
        n = len(self.chronics)
        if self.cfg.loop_mode == 'natural':
            self.next_chronic = (cur + 1) % n
        elif self.cfg.loop_mode == 'random':
            self.next_chronic = int(np.random.choice(n))
        return cur

    def _switch_chronic(self):
        self.chronic_id = self._take_next_chronic()
        self.row = 'id0'                                                # current_timestep_id = 0 (game.py:399)

    def _next_row(self):
        ch = self.chronics[self.chronic_id]
        if self.row is None:
            return 0
        if self.row == 'id0':
            if ch.row_after_switch < 0:
                raise ValueError('chronic %s has no timestep id 0' % ch.name)
            return ch.row_after_switch
        return min(self.row + 1, ch.n_rows - 1)

    def _load_next_timestep(self, is_simulation):
        """game.py:476-501 then :405-474."""
        ch = self.chronics[self.chronic_id]
        at_last = self.row not in (None, 'id0') and self.row == ch.n_rows - 1
        if self.row == 'id0':
            at_last = ch.ids[-1] == 0
        if at_last and not is_simulation:
            self._switch_chronic()
            ch = self.chronics[self.chronic_id]
        row = self._next_row()
        if not is_simulation:
            for a in (self.t_reconnectable, self.t_line_react, self.t_node_react):
                a[a > 0] -= 1
        if not is_simulation:
            self.entries_row, self.entries_chronic = row, self.chronic_id
            self._load_injections(ch.prods_p[row], ch.prods_v[row], ch.loads_p[row], ch.loads_q[row])
        else:
            e = self.chronics[self.entries_chronic]
            r = self.entries_row
            self._load_injections(e.prods_p_planned[r], e.prods_v_planned[r], e.loads_p_planned[r],
                                  e.loads_q_planned[r])
        m = ch.maintenance[row].astype(np.float64)
        mask = m > 0
        self.status[mask] = 0
        self.t_reconnectable[mask] = np.maximum(self.t_reconnectable[mask], m[mask])
        if not is_simulation:
            h = ch.hazards[row].astype(np.float64)
            mask = h > 0
            self.status[mask] = 0
            self.t_reconnectable[mask] = np.maximum(self.t_reconnectable[mask], h[mask])
        self.row = row

    def _gen_basekv(self):
        """normalize_prods_voltages (grid.py:266-271): baseKV of the gen-hosting buses taken in BUS-ARRAY order and
        applied positionally to the generators."""
        gbus = self.c.gen_sub + self.S * self.gen_node
        hosts = np.zeros(2 * self.S, dtype=bool)
        hosts[gbus] = True
        return self.c.bus_basekv[np.flatnonzero(hosts)]

    def _load_injections(self, prods_p, prods_v, loads_p, loads_q):
        """grid.py:273-311."""
        pv = np.where(prods_v <= 0, np.float32(0), prods_v)
        self.gen_pg = prods_p.astype(np.float64)
        self.gen_vg = np.asarray(pv / self._gen_basekv())
        self.gen_status = (prods_v > 0).astype(np.int64)
        lbus = self.c.load_sub + self.S * self.load_node
        self.pd[lbus] = loads_p
        self.qd[lbus] = loads_q

    # ------------------------------------------------------------------------------------------------------ action
    def _split_action(self, action):
        a = np.asarray(action).astype(np.int64)
        if a.shape != (self.c.action_length,):
            raise ValueError('Expected action as a binary array of length %d, got %d' % (self.c.action_length,
                                                                                        a.size))
        G, L, N = self.G, self.L, self.N
        return a[:G + L + 2 * N].copy(), a[G + L + 2 * N:].copy()

    def _changed_substations(self, nodes_sw):
        ch = np.zeros(self.S, dtype=bool)
        ch[self.c.elem_sub[nodes_sw != 0]] = True
        return ch

    def _verify(self, nodes_sw, lines_sw):
        """game.py:650-753 -> (too_much, illegal_reconnect[N], illegal_line_cooldown[N], illegal_sub_cooldown[S])."""
        subs = self._changed_substations(nodes_sw)
        lines = lines_sw == 1
        n_s, n_l = int(subs.sum()), int(lines.sum())
        if n_s > self.cfg.max_sub or n_l > self.cfg.max_lines or n_s + n_l > self.cfg.max_total:
            return True, None, None, None
        return (False, lines & (self.t_reconnectable > 0), lines & (self.t_line_react > 0),
                subs & (self.t_node_react > 0))

    def is_action_valid(self, action):
        nodes_sw, lines_sw = self._split_action(action)
        too, a, b, c = self._verify(nodes_sw, lines_sw)
        return not (too or a.any() or b.any() or c.any())

    def _apply_action(self, nodes_sw, lines_sw):
        """game.py:628-648 + grid.py:360-423 (moving a load swaps Pd/Qd between the sister buses)."""
        G, L, N, S = self.G, self.L, self.N, self.S
        sw_g, sw_l = nodes_sw[:G] != 0, nodes_sw[G:G + L] != 0
        sw_o, sw_e = nodes_sw[G + L:G + L + N] != 0, nodes_sw[G + L + N:] != 0
        self.gen_node = np.where(sw_g, 1 - self.gen_node, self.gen_node)
        for lo in np.flatnonzero(sw_l):
            s = self.c.load_sub[lo]
            self.pd[s], self.pd[s + S] = self.pd[s + S], self.pd[s]
            self.qd[s], self.qd[s + S] = self.qd[s + S], self.qd[s]
        self.load_node = np.where(sw_l, 1 - self.load_node, self.load_node)
        self.or_node = np.where(sw_o, 1 - self.or_node, self.or_node)
        self.ex_node = np.where(sw_e, 1 - self.ex_node, self.ex_node)
        self.status = np.where(lines_sw != 0, 1 - self.status, self.status)
        self.t_line_react[lines_sw == 1] = self.cfg.n_line_react
        self.t_node_react[self._changed_substations(nodes_sw)] = self.cfg.n_node_react

    # ---------------------------------------------------------------------------------------------------- load-flow
    def _buses(self):
        S = self.S
        return (self.c.gen_sub + S * self.gen_node, self.c.load_sub + S * self.load_node,
                self.c.line_or_sub + S * self.or_node, self.c.line_ex_sub + S * self.ex_node)

    def _isolated(self):
        _, _, f, t = self._buses()
        on = self.status != 0
        iso = np.ones(2 * self.S, dtype=bool)
        iso[f[on]] = False
        iso[t[on]] = False
        return iso

    def _loadflow(self):
        """grid.py:244-264 around PYPOWER runpf/rundcpf.  Returns True when the reference would raise
        DivergingLoadflowException.  State is updated the way `self.mpc = output` does."""
        self.n_loadflows += 1
        c, S, NB = self.c, self.S, 2 * self.S
        gbus, lbus, f, t = self._buses()
        iso = self._isolated()
        # _synchronize_bus_types (grid.py:140-174)
        has_gen = np.zeros(NB, dtype=bool)
        has_gen[gbus] = True
        slack = c.slack_sub
        if iso[slack]:
            slack = gbus[gbus != slack][0]
        btype = np.where(iso, 4, np.where(has_gen, 2, 1))
        if not iso[slack] and has_gen[slack]:
            btype[slack] = 3
        # ext2int
        bs = btype != 4
        gs = (self.gen_status > 0) & bs[gbus]
        brs = (self.status != 0) & bs[f] & bs[t]
        e2i = np.cumsum(bs) - 1
        nb = int(bs.sum())
        ibus = np.flatnonzero(bs)
        gon = np.flatnonzero(gs)
        gon = gon[np.argsort(e2i[gbus[gon]], kind='stable')]
        gb = e2i[gbus[gon]]
        # bustypes
        on_gen = np.zeros(nb, dtype=bool)
        on_gen[gb] = True
        ty = btype[ibus]
        ref = np.flatnonzero((ty == 3) & on_gen)
        pv = np.flatnonzero((ty == 2) & on_gen)
        pq = np.flatnonzero((ty == 1) | ~on_gen)
        if len(ref) == 0:
            if len(pv) == 0:
                return True                                             # IndexError -> grid.py:230
            ref, pv = pv[:1], pv[1:]
        pvpq = np.r_[pv, pq]
        br = np.flatnonzero(brs)
        fi, ti = e2i[f[br]], e2i[t[br]]
        # explicit connectivity test (see module docstring)
        reach = np.zeros(nb, dtype=bool)
        reach[ref] = True
        while True:
            new = reach.copy()
            new[ti[reach[fi]]] = True
            new[fi[reach[ti]]] = True
            if (new == reach).all():
                break
            reach = new
        if not reach.all():
            # rundcpf on such a grid returns NaN/garbage that grid.py:260 adopts before raising; the only part of it
            # that survives the reset that follows is the zeroing of out-of-service generators (runpf tail)
            self.gen_pg[~gs] = 0
            self.gen_qg[~gs] = 0
            return True
        if len(pvpq) == 0 or (len(pq) == 0 and not self.cfg.dc):
            return True                                                 # ValueError (norm of empty) -> grid.py:230
        pd, qd = self.pd[ibus], self.qd[ibus]
        Sbus = -(pd + 1j * qd)
        np.add.at(Sbus, gb, self.gen_pg[gon] + 1j * self.gen_qg[gon])
        Sbus = Sbus / c.base_mva
        vm, va = self.vm[ibus].copy(), self.va[ibus].copy()
        gen_pg, gen_qg = self.gen_pg.copy(), self.gen_qg.copy()
        flows = np.zeros((self.N, 4))
        if self.cfg.dc:
            B = np.zeros((nb, nb))
            b = self.bdc[br]
            np.add.at(B, (fi, fi), b)
            np.add.at(B, (ti, ti), b)
            np.add.at(B, (fi, ti), -b)
            np.add.at(B, (ti, fi), -b)
            Pbus = Sbus.real - c.bus_gs[ibus] / c.base_mva
            Va0 = va * (np.pi / 180)
            Va = Va0.copy()
            rhs = Pbus[pvpq] - B[np.ix_(pvpq, ref)] @ Va0[ref]
            try:
                Va[pvpq] = np.linalg.solve(B[np.ix_(pvpq, pvpq)], rhs)
            except np.linalg.LinAlgError:
                Va[pvpq] = np.nan
            flows[br, 0] = b * (Va[fi] - Va[ti]) * c.base_mva
            flows[br, 2] = -flows[br, 0]
            vm[:] = 1.0
            va = Va * (180 / np.pi)
            refgen = gon[np.flatnonzero(gb == ref[0])[0]]
            gen_pg[refgen] = gen_pg[refgen] + (B[ref[0], :] @ Va - Pbus[ref[0]]) * c.base_mva
            success = True
        else:
            V0 = vm * np.exp(1j * np.pi / 180 * va)
            V0[gb] = self.gen_vg[gon] / abs(V0[gb]) * V0[gb]
            Ybus = np.zeros((nb, nb), dtype=complex)
            np.add.at(Ybus, (fi, fi), self.Yff[br])
            np.add.at(Ybus, (fi, ti), self.Yft[br])
            np.add.at(Ybus, (ti, fi), self.Ytf[br])
            np.add.at(Ybus, (ti, ti), self.Ytt[br])
            Ybus[np.arange(nb), np.arange(nb)] += self.Ysh[ibus]
            # makeB, XB
            Bp = np.zeros((nb, nb))
            w = self.bp[br]
            np.add.at(Bp, (fi, fi), w)
            np.add.at(Bp, (ti, ti), w)
            np.add.at(Bp, (fi, ti), -w)
            np.add.at(Bp, (ti, fi), -w)
            Bpp = -Ybus.imag
            V = V0
            Va, Vm = np.angle(V), abs(V)

            def mismatch(V, Vm):
                mis = (V * np.conj(Ybus @ V) - Sbus) / Vm
                P, Q = mis[pvpq].real, mis[pq].imag
                return P, Q, np.max(np.abs(P)), np.max(np.abs(Q))
            P, Q, nP, nQ = mismatch(V, Vm)
            tol = self.cfg.tol
            success = bool(nP < tol and nQ < tol)
            if self.cfg.pf_alg == 1:
                V, success, self.last_iterations = self._newton(Ybus, Sbus, V0, pv, pq)
            with np.errstate(all='ignore'):
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter('ignore')
                    lup = scipy.linalg.lu_factor(Bp[np.ix_(pvpq, pvpq)], check_finite=False)
                    lupp = scipy.linalg.lu_factor(Bpp[np.ix_(pq, pq)], check_finite=False)
                i = 0
                while self.cfg.pf_alg != 1 and not success and i < self.cfg.max_it:
                    i += 1
                    Va[pvpq] = Va[pvpq] - scipy.linalg.lu_solve(lup, P, check_finite=False)
                    V = Vm * np.exp(1j * Va)
                    P, Q, nP, nQ = mismatch(V, Vm)
                    if nP < tol and nQ < tol:
                        success = True
                        break
                    Vm[pq] = Vm[pq] - scipy.linalg.lu_solve(lupp, Q, check_finite=False)
                    V = Vm * np.exp(1j * Va)
                    P, Q, nP, nQ = mismatch(V, Vm)
                    if nP < tol and nQ < tol:
                        success = True
                        break
                if self.cfg.pf_alg != 1:
                    self.last_iterations = i
                # pfsoln
                vm = abs(V)
                va = np.angle(V) * 180 / np.pi
                Sg = V[gb] * np.conj(Ybus[gb, :] @ V)
                gen_qg[:] = 0
                q = Sg.imag * c.base_mva + qd[gb]
                if len(gon) > 1:                                        # proportional split, one gen per bus
                    qmin, qmax = c.gen_qmin[gon], c.gen_qmax[gon]
                    q2 = qmin + ((q - qmin) / (qmax - qmin + np.finfo(float).eps)) * (qmax - qmin)
                    q = np.where(qmin == qmax, q, q2)
                gen_qg[gon] = q
                k = np.flatnonzero(gb == ref[0])[0]
                gen_pg[gon[k]] = Sg[k].real * c.base_mva + pd[ref[0]]
                If = self.Yff[br] * V[fi] + self.Yft[br] * V[ti]
                It = self.Ytf[br] * V[fi] + self.Ytt[br] * V[ti]
                Sf = V[fi] * np.conj(If) * c.base_mva
                St = V[ti] * np.conj(It) * c.base_mva
                flows[br] = np.c_[Sf.real, Sf.imag, St.real, St.imag]
        # int2ext + zeroing of out-of-service results (runpf tail); self.mpc = output (grid.py:260)
        self.vm[ibus], self.va[ibus] = vm, va
        off = ~gs
        gen_pg[off] = 0
        gen_qg[off] = 0
        self.gen_pg, self.gen_qg = gen_pg, gen_qg
        self.flows = flows
        # grid.py:103-110, 263
        with np.errstate(invalid='ignore'):
            def bad(x):
                return bool(np.isnan(x).any() or np.any(x > 1e10))
            nan = bad(self.vm) or bad(self.va) or bad(self.flows) or bad(self.pd)
        return (not success) or nan

    def _newton(self, Ybus, Sbus, V0, pv, pq):
        """PYPOWER newtonpf (PF_ALG = 1; restated in oracle/shims/pypower/api.py:_newtonpf and SURVEY.md Appendix A):
        F = [Re mis[pv]; Re mis[pq]; Im mis[pq]], mis = V conj(Ybus V) - Sbus; J from dSbus_dV; tol on |F|_inf, 10 it."""
        tol, max_it = self.cfg.tol, self.cfg.max_it_nr
        V = V0.copy()
        Va, Vm = np.angle(V), abs(V)
        pvpq = np.r_[pv, pq]
        npv, npq = len(pv), len(pq)

        def F_of(V):
            mis = V * np.conj(Ybus @ V) - Sbus
            return np.r_[mis[pv].real, mis[pq].real, mis[pq].imag]
        F = F_of(V)
        converged = bool(np.max(np.abs(F)) < tol)
        i = 0
        with np.errstate(all='ignore'):
            while not converged and i < max_it:
                i += 1
                Ibus = Ybus @ V
                Vn = V / abs(V)
                dS_dVm = np.diag(V) @ np.conj(Ybus @ np.diag(Vn)) + np.conj(np.diag(Ibus)) @ np.diag(Vn)
                dS_dVa = 1j * np.diag(V) @ np.conj(np.diag(Ibus) - Ybus @ np.diag(V))
                J = np.block([[dS_dVa[np.ix_(pvpq, pvpq)].real, dS_dVm[np.ix_(pvpq, pq)].real],
                              [dS_dVa[np.ix_(pq, pvpq)].imag, dS_dVm[np.ix_(pq, pq)].imag]])
                try:
                    dx = -np.linalg.solve(J, F)
                except np.linalg.LinAlgError:
                    dx = np.full(len(F), np.nan)
                Va[pv] = Va[pv] + dx[:npv]
                Va[pq] = Va[pq] + dx[npv:npv + npq]
                Vm[pq] = Vm[pq] + dx[npv + npq:]
                V = Vm * np.exp(1j * Va)
                Vm, Va = abs(V), np.angle(V)
                F = F_of(V)
                converged = bool(np.max(np.abs(F)) < tol)
        return V, converged, i

    def _flows_a(self):
        """grid.py:112-138, 29-36."""
        _, _, f, _ = self._buses()
        v = self.vm[f] * self.c.bus_basekv[f]
        p, q = self.flows[:, 0], self.flows[:, 1]
        out = np.zeros(self.N)
        on = self.status != 0
        with np.errstate(all='ignore'):
            out[on] = (1000. * np.sqrt(p ** 2 + q ** 2) / (SQRT3 * v))[on]
        return out

    def _cascade(self):
        """game.py:503-589.  Returns True on DivergingLoadflowException."""
        depth = 0
        over = np.zeros(self.N, dtype=bool)
        while True:
            done = True
            if self._loadflow():
                self.last_depth = depth
                return True
            flows_a = self._flows_a()
            over = flows_a > self.thermal
            if over.sum() == 0:
                break
            hard = flows_a > self.cfg.hard_coef * self.thermal
            if hard.any():
                self.status[hard] = 0
                self.t_reconnectable[hard] = self.cfg.n_hard_broken
                done = False
            over[hard] = False
            if over.any():
                soft = over & (self.soft_count >= self.cfg.n_soft_consec)
                if soft.any():
                    self.status[soft] = 0
                    self.t_reconnectable[soft] = self.cfg.n_soft_broken
                    done = False
                    over[soft] = False
            depth += 1
            if done:
                break
        self.last_depth = depth
        self.soft_count[over] += 1
        self.soft_count[~over] = 0
        return False

    # ------------------------------------------------------------------------------------------------------- step
    def step(self, action, _is_simulation=False):
        """game.py:799-885 + environment.py:848-866.  Returns (obs_dyn | None, reward[5], done, flag, info) where
        info = dict(too_much, illegal_reconnect, illegal_line_cooldown, illegal_sub_cooldown, action_used)."""
        nodes_sw, lines_sw = self._split_action(action)
        too, ill_rec, ill_line, ill_sub = self._verify(nodes_sw, lines_sw)
        illegal = too or ill_rec.any() or ill_line.any() or ill_sub.any()
        if too:
            nodes_sw[:] = 0
            lines_sw[:] = 0
        elif illegal:
            lines_sw[ill_rec] = 0
            lines_sw[ill_line] = 0
            nodes_sw[ill_sub[self.c.elem_sub]] = 0
        info = {'too_much': bool(too), 'illegal_reconnect': ill_rec, 'illegal_line_cooldown': ill_line,
                'illegal_sub_cooldown': ill_sub, 'action_used': np.r_[nodes_sw, lines_sw]}
        self._apply_action(nodes_sw, lines_sw)
        self._load_next_timestep(_is_simulation)
        flag, done = FLAG_NONE, False
        if self._cascade():
            flag, done = FLAG_DIVERGING, True
        else:
            iso = self._isolated()
            gbus, lbus, _, _ = self._buses()
            if iso[lbus].sum() > self.cfg.max_loads_go:
                flag, done = FLAG_LOADS_CUT, True
            elif iso[gbus].sum() > self.cfg.max_prods_go:
                flag, done = FLAG_PRODS_CUT, True
        if flag == FLAG_NONE and illegal:
            flag = FLAG_ILLEGAL
        obs = None if done else self.observation_dynamic()
        reward = self.default_reward(obs, nodes_sw, lines_sw, flag, info)
        return obs, reward, done, flag, info

    def _snapshot(self):
        keys = ('gen_node', 'load_node', 'or_node', 'ex_node', 'status', 'vm', 'va', 'pd', 'qd', 'gen_pg', 'gen_qg',
                'gen_vg', 'gen_status', 'flows', 't_reconnectable', 't_line_react', 't_node_react', 'soft_count')
        snap = {k: getattr(self, k).copy() for k in keys}
        snap.update(row=self.row, chronic_id=self.chronic_id, next_chronic=self.next_chronic,
                    entries_row=self.entries_row, entries_chronic=self.entries_chronic)
        return snap

    def _restore(self, snap):
        for k, v in snap.items():
            setattr(self, k, v)

    def simulate(self, action):
        """game.py:887-943: planned injections, maintenance but no hazards, no counter decrement, no commit."""
        snap = self._snapshot()
        try:
            return self.step(action, _is_simulation=True)
        finally:
            self._restore(snap)

    def process_game_over(self):
        """game.py:762-797."""
        while True:
            self.t_reconnectable[:] = 0
            self.t_line_react[:] = 0
            self.t_node_react[:] = 0
            # apply_topology(initial) swaps the loads back (grid.py:405-421)
            self._apply_action_nodes_to_initial()
            self.gen_status[:] = 1
            self.status = self.c.line_status0.astype(np.int64).copy()
            self.va = self.c.bus_va0.copy()
            self.vm = self.c.bus_vm0.copy()
            if self.cfg.hard_mode:
                self._switch_chronic()
            self._load_next_timestep(False)
            if not self._cascade():
                return self.observation_dynamic()

    # ----------------------------------------------------------------------------- state rows (kernel layout)
    def export_rows(self):
        """The env state as the three rows of the CUDA library (include/pypownet_b200.h PPN_STATE_*): real
        (Vm | Va deg | load P | load Q | gen Pg | Qg | Vg), topology (node bits | line status | gen status), counters
        (reconnectable | line cooldown | soft-overflow count | node cooldown | cursor[4])."""
        _, lbus, _, _ = self._buses()
        real = np.concatenate((self.vm, self.va, self.pd[lbus], self.qd[lbus], self.gen_pg, self.gen_qg,
                               self.gen_vg)).astype(np.float64)
        topo = np.concatenate((self.gen_node, self.load_node, self.or_node, self.ex_node, self.status,
                               self.gen_status)).astype(np.uint8)
        row = -1 if self.row is None else (-2 if self.row == 'id0' else int(self.row))
        cnt = np.concatenate((self.t_reconnectable, self.t_line_react, self.soft_count, self.t_node_react,
                              [self.chronic_id, row, self.next_chronic, 0])).astype(np.int32)
        return real, topo, cnt

    def import_rows(self, real, topo, cnt):
        """Inverse of export_rows (test harness: re-synchronisation of a replay after a floating-pocket step)."""
        S, G, L, N, NB = self.S, self.G, self.L, self.N, 2 * self.S
        real, topo, cnt = np.asarray(real, dtype=np.float64), np.asarray(topo), np.asarray(cnt)
        self.vm, self.va = real[:NB].copy(), real[NB:2 * NB].copy()
        o = 2 * NB
        self.gen_node = topo[:G].astype(np.int64)
        self.load_node = topo[G:G + L].astype(np.int64)
        self.or_node = topo[G + L:G + L + N].astype(np.int64)
        self.ex_node = topo[G + L + N:G + L + 2 * N].astype(np.int64)
        self.status = topo[G + L + 2 * N:G + L + 3 * N].astype(np.int64)
        self.gen_status = topo[G + L + 3 * N:2 * G + L + 3 * N].astype(np.int64)
        lbus = self.c.load_sub + S * self.load_node
        self.pd, self.qd = np.zeros(NB), np.zeros(NB)
        self.pd[lbus], self.qd[lbus] = real[o:o + L], real[o + L:o + 2 * L]
        o += 2 * L
        self.gen_pg, self.gen_qg, self.gen_vg = real[o:o + G].copy(), real[o + G:o + 2 * G].copy(), \
            real[o + 2 * G:o + 3 * G].copy()
        self.t_reconnectable = cnt[:N].astype(np.float64)
        self.t_line_react = cnt[N:2 * N].astype(np.float64)
        self.soft_count = cnt[2 * N:3 * N].astype(np.float64)
        self.t_node_react = cnt[3 * N:3 * N + S].astype(np.float64)
        cur = cnt[3 * N + S:3 * N + S + 4]
        self.chronic_id, self.next_chronic = int(cur[0]), int(cur[2])
        self.row = None if cur[1] == -1 else ('id0' if cur[1] == -2 else int(cur[1]))
        self.entries_row, self.entries_chronic = (self.row if isinstance(self.row, int) else None), self.chronic_id
        self.flows = np.zeros((N, 4))

    def _apply_action_nodes_to_initial(self):
        S = self.S
        for lo in np.flatnonzero(self.load_node != 0):
            s = self.c.load_sub[lo]
            self.pd[s], self.pd[s + S] = self.pd[s + S], self.pd[s]
            self.qd[s], self.qd[s + S] = self.qd[s + S], self.qd[s]
        self.gen_node[:] = 0
        self.load_node[:] = 0
        self.or_node[:] = 0
        self.ex_node[:] = 0

    # ------------------------------------------------------------------------------------------ observation / reward
    def observation_dynamic(self):
        """The dynamic prefix (7L+7G+13N+S+6 values) of Observation.as_array (environment.py:451-466, 511-517)."""
        gbus, lbus, f, t = self._buses()
        iso = self._isolated()
        ch = self.chronics[self.entries_chronic]
        r = self.entries_row
        cur = self.chronics[self.chronic_id]
        pm = self.planned_maint.get(self.chronic_id)
        if pm is None:
            pm = self.planned_maint[self.chronic_id] = cur.planned_maintenance_table(self.cfg.horizon)
        ppv = ch.prods_v_planned[r]
        ppv = np.where(ppv <= 0, np.float32(0), ppv) / self._gen_basekv()
        return np.concatenate((
            self.pd[lbus], iso[lbus], ch.loads_p_planned[r], self.load_node,
            self.gen_pg, iso[gbus], ch.prods_p_planned[r], self.gen_node,
            self.or_node, self.ex_node,
            self._flows_a(), self.status, self.t_reconnectable, self.t_line_react, self.t_node_react, pm[self.row],
            cur.datetimes[self.row],
            self.qd[lbus], self.vm[lbus], self.gen_qg, self.gen_vg,
            self.flows[:, 0], self.flows[:, 1], self.vm[f], self.flows[:, 2], self.flows[:, 3], self.vm[t],
            ch.loads_q_planned[r], ppv)).astype(np.float64)

    def observation_static(self):
        """The static tail (S+2L+2G+5N values) of Observation.as_array (environment.py:583-595)."""
        c = self.c
        ids = c.sub_ids
        return np.concatenate((ids, ids[c.load_sub], ids[c.gen_sub], ids[c.line_or_sub], ids[c.line_ex_sub],
                               self.thermal, np.zeros(self.G), np.zeros(self.L), np.zeros(self.N),
                               np.zeros(self.N))).astype(np.float64)

    def observation(self):
        return np.concatenate((self.observation_dynamic(), self.observation_static()))

    def default_reward(self, obs_dyn, nodes_sw, lines_sw, flag, info):
        """parameters/default14/reward_signal.py:45-169 with constant = cfg.reward_constant."""
        k = self.cfg.reward_constant
        cost = -.1 * float(nodes_sw.sum()) + -.2 * float(lines_sw.sum())
        if flag == FLAG_DIVERGING:
            return np.array([0., 0., cost, -k, 0.])
        if flag == FLAG_PRODS_CUT:
            return np.array([0., -k, 0., 0., 0.])
        if flag == FLAG_LOADS_CUT:
            return np.array([-k, 0., 0., 0., 0.])
        gbus, lbus, _, _ = self._buses()
        iso = self._isolated()
        dist = int(self.gen_node.sum() + self.load_node.sum() + self.or_node.sum() + self.ex_node.sum())
        usage = self._flows_a() / self.thermal
        r = np.array([-k / 5. * iso[lbus].sum(), -k / 10. * iso[gbus].sum(), cost, -.02 * dist,
                      -1. * np.sum(np.square(usage))])
        if flag == FLAG_ILLEGAL:
            if info['too_much']:
                r[2] += -5 * k
            else:
                r[2] += (-k / 100.) * info['illegal_reconnect'].sum() + (-k / 100.) * \
                    info['illegal_line_cooldown'].sum() + (-k / 100.) * info['illegal_sub_cooldown'].sum()
        return r
