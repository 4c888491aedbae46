import json,sys
d=json.load(open("gpurun_out/frame_timeline.json"))
k=d["kernels"]; r=d["ranges"]
name=sys.argv[1]; which=int(sys.argv[2])
rg=[x for x in r if x[0]==name][which]
t0,t1=rg[1],rg[1]+rg[2]
prev=None; tot=0
for n,t,du in k:
    if t>=t0 and t<=t1:
        gap = 0 if prev is None else t-prev
        print(f"{t-t0:8.1f} +{du:7.1f} gap {gap:6.1f}  {n[:60]}")
        prev=t+du; tot+=du
print("busy",tot,"range",rg[2])
