import json,sys
d=json.load(open(sys.argv[1]))
sw=d["single_world"]
for k in ("pile100k","addpair20k","mixed10k"):
    for m in ("mode1","mode2"):
        if m in sw[k]:
            st=sw[k][m]["stage_ms_next_2_steps"]
            print(k,m,"%.1f ms"%sw[k][m]["ms_per_step"],"cpu %.1f"%sw[k]["cpu_ms_per_step"],"x%.2f"%sw[k][m]["speedup_vs_cpu_thread"], {a:round(b,2) for a,b in st.items() if b>0.3})
