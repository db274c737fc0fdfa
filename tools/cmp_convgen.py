import json,sys
tag=sys.argv[1]
rows={}
for P in ("halo","tap"):
    for l in open(f"gpurun_out/{tag}_layers_{P}.jsonl"):
        r=json.loads(l)
        if "layer" in r: rows.setdefault(r["layer"],{})[P]=r
        else: print(P, r)
for k,v in rows.items():
    h,t=v.get("halo"),v.get("tap")
    print(k, h["ci"],h["co"],h["stride"], "| fwd halo/tap/cudnn", h["fwd"]["ours_ms"], t["fwd"]["ours_ms"], h["fwd"]["cudnn_ms"], "| dgrad", h["dgrad"]["ours_ms"], t["dgrad"]["ours_ms"], h["dgrad"]["cudnn_ms"], "| wgrad", h["wgrad"]["ours_ms"], h["wgrad"]["cudnn_ms"])
