import sys, time, json
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from monte_carlompi_b200 import engine
from oracle import oracle
import golden_util as gu
print(engine.query(0))
ctx = engine.Context([0])
# 1. replay
import os
for name in (gu.CASES if not os.environ.get('SKIP_REPLAY') else ()):
    c = gu.load_case(name); cfg = c['cfg']
    P = engine.make_params(np.pi*cfg['theta_0']/180., cfg['tau_tot'], cfg['rho_snw'], cfg['Lambertian_reflectance'], cfg['wvl0'], cfg['half_width']/2.355, 0, lambert_bottom=cfg['Lambertian_bottom'])
    t=time.time()
    out = ctx.replay(P, c['wvl'], c['ssa_ice'], c['ssa_imp'], c['g'], c['ext_cff_mss'], c['p_ext_imp'], c['init_draws'], c['offsets'], c['stream'])
    try:
        print('replay', name, gu.compare_replay(out, c), '%.2fs'%(time.time()-t))
    except AssertionError as e:
        print('replay FAIL', name, e)
# 2. production vs oracle, same Philox stream
from monte_carlompi_b200 import ssp_fixtures
def table_for(kind, r, k_lo, k_hi):
    wvl, ssa, ext, g = ssp_fixtures.ice_table(kind, r)
    rows = np.zeros(k_hi-k_lo+1, engine.ROW_DTYPE)
    for j,k in enumerate(range(k_lo,k_hi+1)):
        w = k/100.0
        i = int(np.argmin(np.abs(wvl*1e6 - w)))
        rows[j] = (w, ssa[i], 0.3, g[i], ext[i], 0.0)
    return rows
rows = table_for('spectral', 100, 104, 156)
for (tau, lb, R, th, n) in [(1e6, True, .5, 15., 200000), (3.0, True, .5, 15., 200000), (0.5, True, 0.5, 0., 200000), (2.0, False, 1., 0., 200000)]:
    P = engine.make_params(np.pi*th/180., tau, 300., R, 1.3, 0.085/2.355, 104, lambert_bottom=lb, n_theta_bins=137)
    rec, tally, st = ctx.run(P, rows, 12345, 1000, n)
    Po = oracle.make_params(np.pi*th/180., tau, 300., R, 1.3, 0.085/2.355, 104, lambert_bottom=lb, n_theta_bins=137)
    o = oracle.philox(Po, rows, 12345, 1000, n, n_threads=8)
    same = (rec['condition']==o['condition']) & (rec['n_scat']==o['n_scat']) & (rec['wvl_row']==o['wvl_row'])
    print('prod tau',tau,'lb',lb,'same frac',same.mean(),'row same',(rec['wvl_row']==o['wvl_row']).mean(), 'events', st['n_events'], o['n_events'], 'kernel_ms',st['kernel_ms'])
    print('   cond gpu',np.bincount(rec['condition'],minlength=6)[1:], 'oracle', np.bincount(o['condition'],minlength=6)[1:])
    for col in ('theta_n','phi_n','path_length'):
        a=rec[col][same].astype(np.float64); b=o[col][same]
        d=np.abs(a-b)/np.maximum(np.abs(b),1e-30); print('   ',col,'median rel',np.median(d),'p99',np.percentile(d,99),'max',d.max())
    # tally vs numpy histogram of the records
    h = np.zeros_like(tally)
    for r_ in range(len(rows)):
        m = rec['wvl_row']==r_
        h[r_,0]=m.sum()
        for cnd in range(1,6): h[r_,cnd]=(m&(rec['condition']==cnd)).sum()
        h[r_,8:] = np.histogram(rec['theta_n'][m&(rec['condition']==1)].astype(np.float64), bins=137, range=(0,np.pi/2))[0]
    print('   tally exact:', np.array_equal(h, tally), 'events ok', st['n_events']==int(rec['n_scat'].astype(np.int64).sum()+n))
# 3. timing sweep
P = engine.make_params(np.pi*15/180., 1e6, 300., .5, 1.3, 0.085/2.355, 104, lambert_bottom=True, n_theta_bins=137)
for n in (1000000, 10000000):
  for (bps, bt) in [(4,256),(5,256),(8,128)]:
    for thr in (2,4,6,8,12):
        ctx.set_launch(bps, bt, thr)
        best=None
        for rep in range(3):
            rec, tally, st = ctx.run(P, rows, 777, 0, n, records=False)
            best = st if best is None or st['kernel_ms']<best['kernel_ms'] else best
        print('n',n,'bps',bps,'bt',bt,'thr',thr,'kernel_ms %.3f'%best['kernel_ms'],'events/s %.3e'%(best['n_events']/best['kernel_ms']*1e3),'photons/s %.3e'%(n/best['kernel_ms']*1e3), 'grid', best['grid_blocks'])
