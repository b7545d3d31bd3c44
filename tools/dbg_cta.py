"""GPU: per-CTA timeline of the forward render kernel (cfg3): duration vs work, SM occupancy over time."""
import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("lidar-gs_b200","tests"): sys.path.insert(0, os.path.join(ROOT,p))
import numpy as np, torch, util
from lgs_b200 import synth
from lgs_b200.inspect import frame_views, _al
sc=synth.make_config(3)
sys.path.insert(0, ROOT)
import bench
POSE=int(sys.argv[1]) if len(sys.argv)>1 else 0
sc["viewmatrix"]=bench.rank_pose(sc, POSE)
from lgs_b200 import capi
L=capi.load()
if len(sys.argv)>2: L.lgs_set_forward_split(int(sys.argv[2]))
for it in range(4):
    res, fr = util.run_abi(sc, backward=False)
print('pose',POSE,'forward mode',L.lgs_last_forward_mode(),'longest walk',L.lgs_last_longest_walk())
H,W,P=sc["H"],sc["W"],sc["P"]
v=frame_views(fr,P,H,W)
nb=v["nbins"]
img=fr.image
o=_al(4*H*W); o=_al(o+4*H*W); o=_al(o+4*nb); o=_al(o+16*H*W)
prof=img[o:o+16*nb].view(torch.int32).view(nb,4).cpu().numpy().astype(np.int64) & 0xffffffff
se=v["sorted_end"].cpu().numpy()
t0=prof[:,0]-prof[:,0].min(); dur=prof[:,1]; sm=prof[:,2]; nbat=prof[:,3]  # chunks of 32 (entry, row) pairs evaluated by the workers of the CTA
print("CTAs",nb,"start us: min/median/max",t0.min(),np.median(t0),t0.max())
print("duration kcycles: mean %.1f median %.1f p90 %.1f max %.1f"%(dur.mean()/1e3,np.median(dur)/1e3,np.percentile(dur,90)/1e3,dur.max()/1e3))
idx=np.argsort(-dur)[:12]
print("heaviest CTAs: (dur kcyc, sorted entries, chunks, start us, sm)")
for i in idx: print("   %.1f %d %d %d %d"%(dur[i]/1e3, se[i], nbat[i], t0[i], sm[i]))
# cycles per batch for big vs small
big=nbat>40; small=(nbat>0)&(nbat<=12)
print("kcycles per chunk: big bins %.2f, small bins %.2f"%((dur[big]/nbat[big]).mean()/1e3,(dur[small]/nbat[small]).mean()/1e3))
# per-SM busy sum
busy=np.zeros(sm.max()+1); 
for s_,d_ in zip(sm,dur): busy[s_]+=d_
print("per-SM sum of CTA durations kcyc: min %.0f mean %.0f max %.0f (CTAs overlap, 4/SM)"%(busy.min()/1e3,busy.mean()/1e3,busy.max()/1e3))
end=t0+dur/1965.0
print("kernel span us (from CTA timers): %.1f"%(end.max()))
hist=np.histogram(se,bins=[0,100,200,300,400,600,800,1200,1600,2400])
print("sorted-entries histogram:",list(zip(hist[1][1:],hist[0])))
# how many CTAs running over time
ts=np.linspace(0,end.max(),12)
print("running CTAs over time:",[(round(t),int(((t0<=t)&(end>t)).sum())) for t in ts])
