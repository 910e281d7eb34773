import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from conftest import Golden
from emul import Emul
from cnt_film_monte_carlo_b200.engine import Engine
g=Golden('small_forster')
e=Engine(g.mc); e.set_mesh(g.pos_nm,g.orient); e.kubo_init()
m=Emul(g.mc); m.kubo_init(g.pos_nm,g.orient)
P=64
e.kubo_create_particles(P,seed=3); m.create_philox(P,3)
for n in (1,1,2,8):
    e.kubo_step(g.dt,n); m.kubo_step(g.dt,n)
    pe,pm=e.particles(),m.particles()
    bad=np.nonzero((pe['site']!=pm['site'])|(pe['ndraw']!=pm['ndraw']))[0]
    print('after',n,'steps: mismatching excitons',len(bad),'hops engine',e.hops(),'emul',m.hops())
    for i in bad[:5]:
        print('  exciton',i,'site',pe['site'][i],pm['site'][i],'ndraw',pe['ndraw'][i],pm['ndraw'][i],'ff',pe['ff'][i],pm['ff'][i],'pos',pe['pos'][:,i],pm['pos'][:,i])
