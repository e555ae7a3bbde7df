import sys, numpy as np, time
sys.path.insert(0,'/root/repo')
import bench
from oracle.flat import FlatEnv, Config
import oracle.flat as F
grid='case14'
case,cfg,chronics,imaps=bench.build_workload(grid)
B=int(sys.argv[1]) if len(sys.argv)>1 else 1024
steps=int(sys.argv[2]) if len(sys.argv)>2 else 6
sc,sr=bench.env_starts(B)
ocfg=Config(cfg,reward_constant=float(case.n_sub),n_sub=case.n_sub)
# instrument _loadflow
log=[]
orig=FlatEnv._loadflow
def lf(self):
    self.last_iterations=-1
    r=orig(self)
    log.append(self.last_iterations)   # -1: no iteration (pocket/no ref)
    return r
FlatEnv._loadflow=lf
envs=[FlatEnv(case,ocfg,chronics,start_id=int(sc[e]),thermal_limits=imaps,start_row=int(sr[e])) for e in range(B)]
a=np.zeros(case.action_length,dtype=np.uint8)
def cost(its):   # cycles
    return 12e3 if its<0 else 38e3+2*2.7e3*max(its,0.5)
# warm 3 steps
for w in range(3):
    for e in envs:
        if e.step(a)[2]: e.process_game_over()
res=[]
for t in range(steps):
    step_c=np.zeros(B); att=[[] for _ in range(B)]
    for i,e in enumerate(envs):
        log.clear()
        d=e.step(a)[2]
        step_c[i]=sum(cost(x) for x in log)
        if d:
            # attempts one by one: replicate process_game_over loop manually to tag attempts
            log.clear()
            e.process_game_over()
            # attempts are separated by divergence: each attempt = one cascade; a cascade that diverges ends with a diverging LF.
            # approximate: split log at LFs with its==-1 or its==25 (diverged)... use sequential: an attempt ends when diverged or success at the end
            cur=[]
            for x in log:
                cur.append(x)
                if x<0 or x>=25:   # diverged LF ends the attempt
                    att[i].append(sum(cost(y) for y in cur)); cur=[]
            if cur: att[i].append(sum(cost(y) for y in cur))
    tot=np.array([step_c[i]+sum(att[i]) for i in range(B)])
    T_in=tot.max()
    def split(K):
        l2=max([max(a_[:K]) for a_ in att if a_]+[0])
        l3=max([sum(a_[K:]) for a_ in att if len(a_)>K]+[0])
        return step_c.max()+l2+l3
    def dyn(K):
        return max(step_c[i]+(max(att[i][:K]) if att[i] else 0)+sum(att[i][K:]) for i in range(B))
    res.append((T_in,step_c.max(),split(1),split(4),dyn(4),dyn(8), np.mean(tot), max(len(a_) for a_ in att)))
    print('step %d: T_inplace %.0fk | L1max %.0fk split1 %.0fk split4 %.0fk | dyn4 %.0fk dyn8 %.0fk | mean %.0fk | longest chain %d'%((t,)+tuple(x/1e3 for x in res[-1][:7])+(res[-1][7],)))
