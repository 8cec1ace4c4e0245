import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'tests/golden'); sys.path.insert(0,'.')
import make_golden as MG, oracle_lib as O
from tfg_pathtracer_b200 import renderer as R, scenes as S
for light in (True, False):
  for spp in (8,):
    sc = S.cornell_box(96, env_size=(66,33), light=light, tilt=(3.0,7.0,2.0))
    orc = O.Oracle(sc); orc.render(spp)
    r = R.Renderer(**R.PARITY).render_setup(sc); r.render_cuda(spp)
    a = r.film()[...,:3]; b = orc.film(0)[...,:3]
    bad = ~(np.abs(a-b) <= 1e-3+1e-3*np.abs(b)).all(-1)
    _, pc = r.get_buffers((0,)); s_, opc = orc.counts()
    print("light",light,"spp",spp,"bad frac",bad.mean(),"pc equal",(pc.astype(np.uint32)==opc).mean(), "means", a.mean(), b.mean())
    ys,xs = np.nonzero(bad)
    for k in range(0,len(ys),max(1,len(ys)//12)):
        y,x=ys[k],xs[k]; i=y*96+x
        print("  px",x,y,"ours",a[y,x],"orc",b[y,x],"pc",pc[i],opc[i], "cnt", s_[i])
    # row histogram of bad
    print("  bad rows:", np.nonzero(bad.any(1))[0][:40])
