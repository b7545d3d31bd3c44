"""GPU: how much of each bin's walk is spent with how many live pixels (cfg3)."""
import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("lidar-gs_b200","tests"): sys.path.insert(0, os.path.join(ROOT,p))
import numpy as np, torch, util
from lgs_b200 import synth
from lgs_b200.inspect import frame_views, _al
sc=synth.make_config(3)
res, fr = util.run_abi(sc, backward=False)
H,W,P=sc["H"],sc["W"],sc["P"]
v=frame_views(fr,P,H,W)
nb=v["nbins"]; RB=v["rows_per_bin"]; gx=(W+15)//16
img=fr.image
o=_al(4*H*W); o=_al(o+4*H*W); o=_al(o+4*nb)
fin=img[o:o+16*H*W].view(torch.float32).view(H,W,4)
stop=fin[...,3].contiguous().view(torch.int32).long()
nc=v["n_contrib"]; se=v["sorted_end"]
binid=(torch.arange(H,device=img.device)[:,None]//RB)*gx + (torch.arange(W,device=img.device)[None,:]//16)
end=se[binid]
walk=torch.where(stop>0, stop, end)   # entries each pixel stays live for
print("pixels", H*W, "terminated frac", float((stop>0).float().mean()))
print("sum n_contrib", int(nc.sum()), "sum walk", int(walk.sum()), "sum bin_end (all pixels live to the end)", int(end.sum()))
print("mean n_contrib", float(nc.float().mean()), "mean stop (terminated)", float(stop[stop>0].float().mean()), "mean bin end", float(end.float().mean()))
nt=(stop==0)
print("non-terminated: count", int(nt.sum()), "mean final_T", float(v["final_T"][nt].mean()) if nt.any() else None, "mean bin_end there", float(end[nt].float().mean()) if nt.any() else None)
# per bin: ratio of mean walk to end
per=torch.zeros(nb,device=img.device).index_add_(0,binid.flatten(),walk.flatten().float())/ (RB*16)
print("mean over bins of (mean walk / end):", float((per/ end.new_tensor(1.0) / se.clamp(min=1).float()).mean()))
tot=torch.zeros(nb,device=img.device).index_add_(0,binid.flatten(),torch.ones(H*W,device=img.device))
print("bins with any non-terminated pixel:", int((torch.zeros(nb,device=img.device).index_add_(0,binid.flatten(),nt.flatten().float())>0).sum()), "of", nb)
bins_total=(v["binbase"][1:]-v["binbase"][:-1])
print("bins walked to the very end:", int((se>=bins_total).sum()), "mean total per bin", float(bins_total.float().mean()))
