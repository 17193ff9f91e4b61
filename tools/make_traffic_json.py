"""profiles/traffic.json from ncu --set full captures: DRAM read + write bytes per launch of every kernel.
python tools/make_traffic_json.py l216=gpurun_out/prof_r2u_l216.ncu-rep default=... velprofile=... > profiles/traffic.json"""
import csv, json, subprocess, sys
out = {"_source": {}}
ALIAS = (("k_vel_res", "k_vel"), ("k_velprofile", "k_velprofile"), ("k_plan", "k_plan"), ("k_path", "k_path"),
         ("k_export", "k_export"), ("k_state", "k_state"), ("k_ref", "k_ref"), ("k_prefix", "k_prefix"), ("k_backup", "k_backup"))
for arg in sys.argv[1:]:
    tag, rep = arg.split("=", 1)
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    d = {}
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        key = next((v for k, v in ALIAS if k in name), name)
        mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for col in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):   # every column carries its own unit
            tot += float(r[hdr.index(col)]) * mul[rows[1][hdr.index(col)]]
        d[key] = tot
    out[tag] = d
    out["_source"][tag] = rep.split("/")[-1]
print(json.dumps(out, indent=1, sort_keys=True))
